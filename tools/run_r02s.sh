#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02s_attention_ab.txt; : > $out
for rep in 1 2 3; do
for mode in 0 bwd 1; do
  line=$(VITAE_ATTN_LEGACY=$mode timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1))")
  echo "VITAE_ATTN_LEGACY=$mode (0: tcgen05 fwd+bwd, bwd: tcgen05 fwd + mma.sync bwd, 1: mma.sync fwd+bwd) rep=$rep ms_per_step,vol/s: $line" >> $out
done
done
for mode in 0 1; do
  line=$(VITAE_ATTN_LEGACY=$mode timeout 300 python bench.py --steps 10 --warmup 5 --batch 16 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1))")
  echo "batch 16: VITAE_ATTN_LEGACY=$mode ms_per_step,vol/s: $line" >> $out
done
cat $out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02s_pytest.log
