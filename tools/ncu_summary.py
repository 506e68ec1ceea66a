#!/usr/bin/env python
"""Summarises an `ncu --set full` report (read here, no GPU needed): one row per profiled launch with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [shape names...] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "dur_us", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1.0),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1.0),
    ("dram__bytes_read.sum", "dram_rd_MB", 1.0),
    ("dram__bytes_write.sum", "dram_wr_MB", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__shared_mem_per_block_dynamic", "smem_dyn", 1.0),
]


def main():
    rep = sys.argv[1]
    labels = sys.argv[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised, ~40 replays per launch)")
    print("# " + "  ".join(f"{n}[{units[col[m]]}]" for m, n, _ in METRICS if m in col))
    for k, r in enumerate(data):
        name = r[col["Kernel Name"]]
        name = name[name.find("gemm_bf16"):][:60] if "gemm_bf16" in name else name[:60]
        vals = []
        for m, n, _ in METRICS:
            if m in col:
                try:
                    vals.append(f"{n}={float(r[col[m]].replace(',', '')):.2f}")
                except ValueError:
                    vals.append(f"{n}={r[col[m]]}")
        lab = labels[k // max(1, len(data) // max(1, len(labels)))] if labels else ""
        print(f"{lab:14s} grid={r[col['Grid Size']]:>14s} {name:62s} " + " ".join(vals))


if __name__ == "__main__":
    main()
