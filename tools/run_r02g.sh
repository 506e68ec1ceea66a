#!/bin/bash
mkdir -p gpurun_out
VITAE_LIB=$PWD/vit_ae_plus_plus_b200/libvitae_b200_attntrace.so timeout 200 python tools/attn_trace.py > gpurun_out/r02g_attn_trace_res.txt 2>&1
