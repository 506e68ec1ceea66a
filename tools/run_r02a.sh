#!/bin/bash
# round-2 first GPU pass: tests, the BASELINE configs' bench lines, ncu of the non-GEMM kernels
mkdir -p gpurun_out
python -m vit_ae_plus_plus_b200.build > gpurun_out/r02a_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
for mr in 0.25 0.5; do
  timeout 400 python bench.py --steps 20 --warmup 5 --mask-ratio $mr --cpu-sample-steps 2 > gpurun_out/r02a_bench_mask$mr.json 2> gpurun_out/r02a_bench_mask$mr.err
done
timeout 400 python bench.py --steps 20 --warmup 5 --workload vit_large_96 --batch 2 --cpu-sample-steps 2 > gpurun_out/r02a_bench_vitl_b2.json 2> gpurun_out/r02a_bench_vitl_b2.err
timeout 400 python bench.py --steps 20 --warmup 5 --workload vit_large_96 --batch 16 --no-cpu-baseline > gpurun_out/r02a_bench_vitl_b16.json 2> gpurun_out/r02a_bench_vitl_b16.err
# non-GEMM kernels: one --set full pass, a handful of launches of each
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'attn_|layernorm_|masked_mse|block_colreduce|adamw_flat|grad_sqnorm|im2col' -s 300 -c 60 \
   -o gpurun_out/r02a_small python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
ls -la gpurun_out | tail -20
