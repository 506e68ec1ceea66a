#!/bin/bash
# one 8-GPU box: scaling 1/2/4/8 with the defaults, then exchange variants at 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02r_gpus.txt
run() { # n, name, extra bench flags, extra env...
  n=$1; name=$2; flags=$3; shift; shift; shift
  if [ $n -eq 1 ]; then
    env "$@" timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02r_$name.json 2> gpurun_out/r02r_$name.err
  else
    env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29540 \
        bench.py --gpus $n --steps 20 --warmup 5 $flags > gpurun_out/r02r_$name.json 2> gpurun_out/r02r_$name.err
  fi
}
run 8 n8_fp32 '' VITAE_GRAD_EXCHANGE=fp32
run 8 n8_bf16 --no-e2e VITAE_GRAD_EXCHANGE=bf16
run 4 n4 --no-e2e CUDA_VISIBLE_DEVICES=0,1,2,3
run 2 n2 --no-e2e CUDA_VISIBLE_DEVICES=0,1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
   bench.py --gpus 8 --steps 2 --warmup 3 --no-e2e 2>&1 | grep -E "NCCL INFO" | grep -E "AllReduce|Algo|algo|NVLS|nranks|Connected|comm 0x" | head -80 > gpurun_out/r02r_nccl_info_8gpu.txt
for f in gpurun_out/r02r_n*.json; do echo $f; grep '^{' $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), d.get('e2e') and round(d['e2e']['value'],1), (d.get('dp_check') or {}).get('ok'))"; done
