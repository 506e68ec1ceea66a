#!/bin/bash
# round-2 GPU pass b: (1) everything except attention on the round-1 attention kernels, (2) the tcgen05 attention kernels
mkdir -p gpurun_out
python -m vit_ae_plus_plus_b200.build > gpurun_out/r02b_build.log 2>&1
VITAE_ATTN_LEGACY=1 timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_legacy.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b_pytest_legacy.log
VITAE_ATTN_LEGACY=1 timeout 120 python tools/attn_check.py > gpurun_out/r02b_attn_legacy.txt 2>&1
timeout 120 python tools/attn_check.py > gpurun_out/r02b_attn_tc.txt 2>&1
echo "attn_check rc=$?" >> gpurun_out/r02b_attn_tc.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention > gpurun_out/r02b_pytest_attn.log 2>&1
VITAE_ATTN_LEGACY=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_legacy.json 2> gpurun_out/r02b_bench_legacy.err
for mr in 0.25 0.5; do
  VITAE_ATTN_LEGACY=1 timeout 400 python bench.py --steps 20 --warmup 5 --mask-ratio $mr --cpu-sample-steps 2 > gpurun_out/r02b_bench_mask$mr.json 2> gpurun_out/r02b_bench_mask$mr.err
done
VITAE_ATTN_LEGACY=1 timeout 400 python bench.py --steps 20 --warmup 5 --workload vit_large_96 --batch 2 --cpu-sample-steps 2 > gpurun_out/r02b_bench_vitl_b2.json 2> gpurun_out/r02b_bench_vitl_b2.err
VITAE_ATTN_LEGACY=1 timeout 400 python bench.py --steps 20 --warmup 5 --workload vit_large_96 --batch 16 --no-cpu-baseline > gpurun_out/r02b_bench_vitl_b16.json 2> gpurun_out/r02b_bench_vitl_b16.err
# non-GEMM kernels: one --set full pass (report stays on the box; the raw page comes back as csv)
VITAE_ATTN_LEGACY=1 timeout 900 ncu --set full --clock-control none \
   -k regex:'attn_|layernorm_|masked_mse|block_colreduce|adamw_flat|grad_sqnorm|im2col' -s 300 -c 60 \
   -o /tmp/r02b_small python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02b_ncu.log 2>&1
ncu -i /tmp/r02b_small.ncu-rep --page raw --csv > gpurun_out/r02b_small_raw.csv 2>> gpurun_out/r02b_ncu.log
du -sh gpurun_out
