#!/bin/bash
# final 2-GPU sanity of the committed code: full bench line (e2e included) and the contrastive workload through the sharded step
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02zd_bench_n2.json 2> gpurun_out/r02zd_bench_n2.err
grep '^{' gpurun_out/r02zd_bench_n2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), d['e2e'] and (round(d['e2e']['value'],1), d['e2e']['ms_per_step_repeats']), d.get('dp_mode'), d['dp_check']['ok'])" || grep -v "^\[W\|^W1\|^\*\*\*\|^frame" gpurun_out/r02zd_bench_n2.err | grep -E "Error|what\(\)|rank0\]:" | head -20
timeout 400 $TR --master-port 29583 bench.py --gpus 2 --steps 20 --warmup 5 --workload contr_vit_base_128 --no-e2e > gpurun_out/r02zd_bench_n2_contr.json 2> gpurun_out/r02zd_bench_n2_contr.err
grep '^{' gpurun_out/r02zd_bench_n2_contr.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('contr', round(d['value'],1), round(d['ms_per_step'],3), d.get('dp_mode'), d.get('final_loss'))" || grep -v "^\[W\|^W1\|^\*\*\*\|^frame" gpurun_out/r02zd_bench_n2_contr.err | grep -E "Error|what\(\)|rank0\]:" | head -20
