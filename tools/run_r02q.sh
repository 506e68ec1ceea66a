#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/r02q_ab.txt; : > $out
for rep in 1 2 3; do
for cfg in 1 0; do
  line=$(VITAE_FUSED_LOSS=$cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['roofline']['gemm_ms_per_step'],3))")
  echo "fused_loss=$cfg rep=$rep ms_per_step,gemm_ms: $line" >> $out
done
done
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off -k regex:gemm_bf16 --csv --log-file gpurun_out/r02q_gemm_traffic_cold.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-range > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none --cache-control none --profile-from-start off -k regex:gemm_bf16 --csv --log-file gpurun_out/r02q_gemm_traffic_insitu.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-range > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r02q_all_traffic_insitu.csv \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-range > /dev/null 2>&1
ls -la gpurun_out | grep r02q
