#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a small markdown table (profiles/):
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/rN_full_summary.md"""
import csv
import re
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__shared_mem_per_block_dynamic", "dyn smem")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
cols = [(hdr.index(k), lab) for k, lab in WANT if k in hdr]
ki = hdr.index("Kernel Name")
print("| kernel | " + " | ".join(lab for _, lab in cols) + " |")
print("|---|" + "---:|" * len(cols))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    print("| %s | " % name + " | ".join("%s %s" % (r[i], units[i]) for i, _ in cols) + " |")
