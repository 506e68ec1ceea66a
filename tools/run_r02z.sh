#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for rep in 1 2; do
  timeout 400 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02z_bench2_rep$rep.json 2> gpurun_out/r02z_bench2.err
  grep '^{' gpurun_out/r02z_bench2_rep$rep.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rep=$rep', round(d['value'],1), round(d['ms_per_step'],3), d.get('dp_mode'), d['dp_check']['ok'], d['dp_check']['sharded_vs_allreduce'])" || grep -v "^\[W\|^W1\|^\*\*\*\|^frame" gpurun_out/r02z_bench2.err | grep -E "Error|what\(\)|rank0\]:" | head -20
done
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 $TR --master-port 29581 tools/dp_timeline.py 2> gpurun_out/r02z_tl.err | grep '^{' > gpurun_out/r02z_dp_timeline_2gpu.txt; cat gpurun_out/r02z_dp_timeline_2gpu.txt
