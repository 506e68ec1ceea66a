#!/bin/bash
# evidence pass: full tests, bench lines, launch list, ncu of attention + the decoder-phase kernels, compute-sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
VITAE_ATTN_LEGACY=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_mma_sync_attention.json 2>/dev/null
timeout 300 python bench.py --steps 10 --warmup 5 --batch 16 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_b16.json 2>/dev/null
VITAE_ATTN_LEGACY=1 timeout 300 python bench.py --steps 10 --warmup 5 --batch 16 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_b16_mma_sync_attention.json 2>/dev/null
# launch list of two steps (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 1200 --csv --log-file gpurun_out/r02k_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02k_ncu_launches.log 2>&1
# attention kernels, --set full
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_' -s 14 -c 14 -o /tmp/r02k_attn python tools/attn_ncu_driver.py > gpurun_out/r02k_ncu_attn.log 2>&1
ncu -i /tmp/r02k_attn.ncu-rep --page raw --csv > gpurun_out/r02k_attn_raw.csv 2>> gpurun_out/r02k_ncu_attn.log
# decoder-phase small kernels that the round's first pass missed
timeout 900 ncu --set full --clock-control none -k regex:'masked_mse|layernorm_fwd_kernel<4>|layernorm_bwd_kernel<4>|ingest_|bn_relu|cosine' -s 0 -c 24 \
   -o /tmp/r02k_small python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_ncu_small.log 2>&1
ncu -i /tmp/r02k_small.ncu-rep --page raw --csv > gpurun_out/r02k_small_raw.csv 2>> gpurun_out/r02k_ncu_small.log
# compute-sanitizer on a tiny step (eager launches, lanes on)
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py > gpurun_out/r02k_sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step.py > gpurun_out/r02k_sanitizer_racecheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_step.py > gpurun_out/r02k_sanitizer_synccheck.log 2>&1
du -sh gpurun_out
