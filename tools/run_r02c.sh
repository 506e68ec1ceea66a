#!/bin/bash
mkdir -p gpurun_out
VITAE_LIB=$PWD/vit_ae_plus_plus_b200/libvitae_b200_attntrace.so timeout 200 python tools/attn_trace.py > gpurun_out/r02c_attn_trace.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "ingest or casts or attention" > gpurun_out/r02c_pytest.log 2>&1
