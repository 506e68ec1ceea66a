#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02zb_pytest_dp.log; cat gpurun_out/r02zb_pytest_dp.log
for ov in 1 0; do
  VITAE_DP_OVERLAP_GATHER=$ov timeout 400 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02zb_bench_n2_overlap$ov.json 2> gpurun_out/r02zb_bench_n2.err
  grep '^{' gpurun_out/r02zb_bench_n2_overlap$ov.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('overlap_gather=$ov', round(d['value'],1), round(d['ms_per_step'],3), d['dp_check']['ok'], d['dp_check']['sharded_vs_allreduce'])" || grep -v "^\[W\|^W1\|^\*\*\*\|^frame" gpurun_out/r02zb_bench_n2.err | grep -E "Error|what\(\)|rank0\]:" | head -20
done
VITAE_DP_GATHER_BLOCKS=148 timeout 400 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gather_blocks=148', round(d['value'],1), round(d['ms_per_step'],3))"
