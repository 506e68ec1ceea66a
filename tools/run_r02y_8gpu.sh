#!/bin/bash
# one 8-GPU box: the sharded data-parallel step at 8 and 4 ranks against the all-reduce path, with a timeline at 8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # n, name, flags, env...
  n=$1; name=$2; flags=$3; shift; shift; shift
  env "$@" timeout 500 $TR --nproc-per-node $n --master-port 29540 bench.py --gpus $n --steps 20 --warmup 5 $flags > gpurun_out/r02y_$name.json 2> gpurun_out/r02y_$name.err
  grep '^{' gpurun_out/r02y_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d.get('e2e') and round(d['e2e']['value'],1), d.get('dp_mode'), d['dp_check'])" || grep -v "^\[W\|^W1\|^\*\*\*\|^frame" gpurun_out/r02y_$name.err | tail -12
}
run 8 n8_sharded '' A=1
run 8 n8_sharded_blocks48 --no-e2e VITAE_DP_REDUCE_BLOCKS=48
run 4 n4_sharded --no-e2e CUDA_VISIBLE_DEVICES=0,1,2,3
timeout 300 $TR --nproc-per-node 8 --master-port 29541 tools/dp_timeline.py 2> gpurun_out/r02y_tl.err | grep '^{' > gpurun_out/r02y_dp_timeline_8gpu.txt
cat gpurun_out/r02y_dp_timeline_8gpu.txt
