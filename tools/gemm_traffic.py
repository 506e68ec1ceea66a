#!/usr/bin/env python
"""DRAM traffic of the step's GEMM launches from an ncu metrics pass over ONE step (bench.py --profile-range):
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none [--cache-control none]
      --profile-from-start off -k regex:gemm_bf16 --csv --log-file x.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-range
usage: python tools/gemm_traffic.py x.csv "<how it was taken>" > profiles/rNN_gemm_dram_traffic.json"""
import csv
import json
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
    rd = wr = dur = 0.0
    ids = set()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        name = row["Metric Name"]
        if name.startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            if "read" in name:
                rd += v
            else:
                wr += v
        elif name == "gpu__time_duration.sum":
            dur += v * {"ns": 1e-3, "us": 1, "ms": 1e3}[u]
        ids.add(row["ID"])
    print(json.dumps({"source": sys.argv[2] if len(sys.argv) > 2 else "", "gemm_launches": len(ids), "gemm_dram_read_bytes": rd,
                      "gemm_dram_write_bytes": wr, "gemm_dram_bytes": rd + wr, "gemm_time_us_under_ncu": dur}, indent=1))


if __name__ == "__main__":
    main()
