#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02za_pytest_dp.log; cat gpurun_out/r02za_pytest_dp.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02za_bench_n2.json 2> gpurun_out/r02za_bench_n2.err
grep '^{' gpurun_out/r02za_bench_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), d['e2e'] and round(d['e2e']['value'],1), d.get('dp_mode'), d['dp_check'])" || grep -v "^\[W\|^W1\|^\*\*\*\|^frame" gpurun_out/r02za_bench_n2.err | grep -E "Error|what\(\)|rank0\]:" | head -20
