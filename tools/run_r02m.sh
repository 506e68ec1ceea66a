#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02m_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
VITAE_ATTN_LEGACY=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02m_bench_mma_sync_attention.json 2>/dev/null
