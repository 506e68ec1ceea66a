#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02o_ab.txt; : > $out
for rep in 1 2; do
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  line=$(VITAE_FUSED_LOSS=$1 VITAE_NORM_PARTS=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['gemm_ms_per_step'],3))")
  echo "fused_loss=$1 norm_parts=$2 rep=$rep ms_per_step,gemm_ms: $line" >> $out
done
done
for rep in 1 2; do
  line=$(VITAE_ATTN_LEGACY=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4))")
  echo "mma_sync_attention rep=$rep: $line" >> $out
done
cat $out
