#!/bin/bash
mkdir -p gpurun_out
export VITAE_LIB=$PWD/vit_ae_plus_plus_b200/libvitae_b200_attntrace.so
VITAE_ATTN_FWD=v1 VITAE_ATTN_FWD_CFG=0 timeout 200 python tools/attn_trace.py events > gpurun_out/r02e_events_cfg0.txt 2>&1
VITAE_ATTN_FWD=v1 VITAE_ATTN_FWD_CFG=1 timeout 200 python tools/attn_trace.py events > gpurun_out/r02e_events_cfg1.txt 2>&1
unset VITAE_LIB
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_kfold_replay_gpu.py -m gpu -q -k "bn_relu or cosine or contrastive or latent or kfold" > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
