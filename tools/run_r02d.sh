#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/attn_check.py > gpurun_out/r02d_attn_tc8.txt 2>&1
VITAE_ATTN_FWD=v1 timeout 200 python tools/attn_check.py > gpurun_out/r02d_attn_tc_v1fwd.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
