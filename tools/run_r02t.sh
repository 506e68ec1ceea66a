#!/bin/bash
# multi-issuer backward kernels: correctness (attn_check, attention tests), isolated timings, in-step A/B against mma.sync
mkdir -p gpurun_out
timeout 300 python tools/attn_check.py > gpurun_out/r02t_attn_check.txt 2>&1; echo "attn_check rc=$?" >> gpurun_out/r02t_attn_check.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention > gpurun_out/r02t_pytest_attn.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_pytest_attn.log
out=gpurun_out/r02t_attention_ab.txt; : > $out
for rep in 1 2 3; do
for mode in 0 bwd; do
  line=$(VITAE_ATTN_LEGACY=$mode timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1))")
  echo "VITAE_ATTN_LEGACY=$mode (0: tcgen05 fwd+bwd [multi-issuer bwd], bwd: tcgen05 fwd + mma.sync bwd) rep=$rep ms_per_step,vol/s: $line" >> $out
done
done
for mode in 0 bwd; do
  line=$(VITAE_ATTN_LEGACY=$mode timeout 300 python bench.py --steps 10 --warmup 5 --batch 16 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1))")
  echo "batch 16: VITAE_ATTN_LEGACY=$mode ms_per_step,vol/s: $line" >> $out
done
cat $out
tail -20 gpurun_out/r02t_attn_check.txt
tail -3 gpurun_out/r02t_pytest_attn.log
