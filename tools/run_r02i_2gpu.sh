#!/bin/bash
# two-GPU validation: DP self-check test, bench with bf16 and fp32 gradient exchange, NCCL algorithm / protocol log
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02i_gpus.txt
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q > gpurun_out/r02i_pytest_dp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02i_pytest_dp.log
run() { # name, extra env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02i_bench2_$name.json 2> gpurun_out/r02i_bench2_$name.err
}
run bf16 VITAE_GRAD_EXCHANGE=bf16
run fp32 VITAE_GRAD_EXCHANGE=fp32
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
   bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e 2>&1 | grep -E "NCCL INFO (AllReduce|Broadcast|Connected|Channel|comm|Using|NVLS|nranks|Algo|algo)" | head -150 > gpurun_out/r02i_nccl_info.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_bench1.json 2> gpurun_out/r02i_bench1.err
