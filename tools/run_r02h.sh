#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_check.py > gpurun_out/r02h_attn.txt 2>&1
VITAE_LIB=$PWD/vit_ae_plus_plus_b200/libvitae_b200_attntrace.so timeout 200 python tools/attn_trace.py > gpurun_out/r02h_attn_trace.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x > gpurun_out/r02h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
