#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/host_profile.py > gpurun_out/r02p_host_profile.txt 2>&1
timeout 300 python tools/step_breakdown.py > gpurun_out/r02p_step_breakdown.txt 2>&1
VITAE_ATTN_LEGACY=1 timeout 300 python tools/step_breakdown.py > gpurun_out/r02p_step_breakdown_mma_sync.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02p_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profile-range > gpurun_out/r02p_ncu_launches.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --workload contr_vit_base_128 --cpu-sample-steps 2 > gpurun_out/r02p_bench_contr.json 2> gpurun_out/r02p_bench_contr.err
