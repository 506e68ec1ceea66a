#!/bin/bash
# final single-GPU pass of the round: whole GPU suite, the bench lines of record, ncu launch list, step breakdown
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02zc_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zc_pytest_gpu.log
tail -4 gpurun_out/r02zc_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02zc_bench_vit_base.json 2> gpurun_out/r02zc_bench_vit_base.err
b() { name=$1; shift; timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02zc_bench_$name.json 2> gpurun_out/r02zc_bench_$name.err; }
b mask050 --mask-ratio 0.5 --no-e2e
b mask025 --mask-ratio 0.25 --no-e2e
b vit_large_b2 --workload vit_large_96 --batch 2 --no-e2e
b vit_large_b16 --workload vit_large_96 --batch 16 --no-e2e
b vit_base_b16 --batch 16 --no-e2e
b contr --workload contr_vit_base_128
for f in gpurun_out/r02zc_bench_*.json; do echo $f; grep '^{' $f | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), d.get('e2e') and round(d['e2e']['value'],1), d.get('roofline') and round(d['roofline']['frac'],3), d.get('cpu_baseline') and d['cpu_baseline'].get('value'), d.get('gpu_launches'))"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02zc_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --profile-range > gpurun_out/r02zc_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r02zc_launches.csv > gpurun_out/r02zc_launches_summary.txt 2>&1; head -12 gpurun_out/r02zc_launches_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" > gpurun_out/r02zc_smoke.log 2>&1; tail -2 gpurun_out/r02zc_smoke.log
timeout 300 python tools/step_breakdown.py > gpurun_out/r02zc_step_breakdown.txt 2>&1; tail -12 gpurun_out/r02zc_step_breakdown.txt
