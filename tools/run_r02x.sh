#!/bin/bash
# sharded step on 2 GPUs: grid of the overlapped reduce
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
out=gpurun_out/r02x_dp_timeline.txt; : > $out
tl() { env "$@" timeout 300 $TR --master-port 29581 tools/dp_timeline.py 2> gpurun_out/r02x_last.err | grep '^{' >> $out; [ ${PIPESTATUS[0]} -ne 0 ] && grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02x_last.err | tail -8 >> $out; }
for b in 48 64 96 128 192 296; do tl VITAE_DP_REDUCE_BLOCKS=$b; done
cat $out
for b in 64 96 128; do
  VITAE_DP_REDUCE_BLOCKS=$b timeout 400 $TR --master-port 29582 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02x_bench2_blocks$b.json 2> gpurun_out/r02x_bench2.err
  grep '^{' gpurun_out/r02x_bench2_blocks$b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('blocks=$b', round(d['value'],1), round(d['ms_per_step'],3), d.get('dp_mode'), d['dp_check']['ok'])" || grep -v "^\[W\|^W1\|^\*\*\*" gpurun_out/r02x_bench2.err | tail -12
done
