#!/bin/bash
# sharded data-parallel step on 2 GPUs: replica check, timeline against the all-reduce path, bench lines
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29580 tests/dp_check.py > gpurun_out/r02w_dp_check.txt 2> gpurun_out/r02w_dp_check.err; echo "dp_check rc=$?" >> gpurun_out/r02w_dp_check.txt
tail -3 gpurun_out/r02w_dp_check.txt
grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02w_dp_check.err | grep -B2 -A12 "Traceback" | head -40
out=gpurun_out/r02w_dp_timeline.txt; : > $out
tl() { env "$@" timeout 300 $TR --master-port 29581 tools/dp_timeline.py 2> gpurun_out/r02w_last.err | grep '^{' >> $out; [ ${PIPESTATUS[0]} -ne 0 ] && grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02w_last.err | tail -8 >> $out; }
tl A=1
tl VITAE_DP_SHARDED=0
tl VITAE_DP_REDUCE_BLOCKS=16
tl VITAE_DP_REDUCE_BLOCKS=96
cat $out
for mode in 1 0; do
  VITAE_DP_SHARDED=$mode timeout 400 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02w_bench2_sharded$mode.json 2> gpurun_out/r02w_bench2_sharded$mode.err
  grep '^{' gpurun_out/r02w_bench2_sharded$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('sharded=$mode', round(d['value'],1), round(d['ms_per_step'],3), d.get('dp_check'))" || grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02w_bench2_sharded$mode.err | tail -12
done
