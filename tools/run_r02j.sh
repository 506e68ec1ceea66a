#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_check.py > gpurun_out/r02j_attn.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
