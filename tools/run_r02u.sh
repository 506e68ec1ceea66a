#!/bin/bash
# 2-GPU timeline of the data-parallel step under a few NCCL / staging settings
mkdir -p gpurun_out
out=gpurun_out/r02u_dp_timeline.txt; : > $out
tl() { # env...
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 tools/dp_timeline.py 2> gpurun_out/r02u_last.err | grep '^{' >> $out
  [ ${PIPESTATUS[0]} -ne 0 ] && tail -5 gpurun_out/r02u_last.err >> $out
}
timeout 300 python tools/dp_timeline.py 2> gpurun_out/r02u_n1.err | grep '^{' >> $out
tl A=1
tl NCCL_MAX_CTAS=8
tl NCCL_MAX_CTAS=4
tl NCCL_MAX_CTAS=16
tl VITAE_DP_GROUP=1
tl VITAE_DP_GROUP=2 NCCL_MAX_CTAS=8
tl VITAE_GRAD_EXCHANGE=bf16
cat $out
