#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r02l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
VITAE_GEMM_ATOM_TMA=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02l_bench_noatoms.json 2>/dev/null
timeout 300 python tools/gemm_bench.py > gpurun_out/r02l_gemm_shapes.txt 2>&1
