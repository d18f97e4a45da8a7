#!/usr/bin/env python
"""Summarise an .ncu-rep into text: headline metrics (raw page) plus the instructions with the most warp-stall
samples (source page, SASS).  The second part is what exposed the single-thread TMA producer as the bottleneck of the
3x3 wgrad (profiles/README.md).      python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/ncu_prof_x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]} {units[hdr.index(k)]}")
    src = page(rep, "source", ("--print-source", "sass"))
    if len(src) > 2:
        hdr = src[1]
        rows = [dict(zip(hdr, r)) for r in src[2:] if len(r) == len(hdr)]
        total = sum(int(r["# Samples"]) for r in rows) or 1
        print(f"\nwarp-stall samples: {total} over {len(rows)} SASS instructions; top {top}:")
        for i, r in sorted(enumerate(rows), key=lambda t: -int(t[1]["# Samples"]))[:top]:
            print(f"  #{i:5d} {int(r['# Samples']):6d} ({100 * int(r['# Samples']) / total:4.1f}%) exec {r['Instructions Executed']:>9s}  {r['Source'].strip()[:90]}")


if __name__ == "__main__":
    main()
