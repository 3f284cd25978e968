#!/usr/bin/env python
"""Summarise an `ncu --set full` report into a small markdown table (profiles/):
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv > profiles/rN_full_summary.md
    python tools/ncu_summary.py raw.csv --traffic ENGINE BATCH N_GEMM_PER_STEP "<command>" > profiles/rN_gemm_traffic.json
--traffic sums dram__bytes_read + dram__bytes_write over the LAST N_GEMM_PER_STEP GEMM launches of the capture (one train step) —
the file bench.py's roofline.traffic is read from (never a constant in the code)."""
import csv
import json
import re
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("smsp__inst_executed.sum", "warp inst"),
        ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__shared_mem_per_block_dynamic", "dyn smem")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")


def to_bytes(v, unit):
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


if "--traffic" in sys.argv:
    a = sys.argv.index("--traffic")
    engine, batch, n = sys.argv[a + 1], int(sys.argv[a + 2]), int(sys.argv[a + 3])
    cmd = sys.argv[a + 4] if len(sys.argv) > a + 4 else ""
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    gemm = [r for r in rows[2:] if "tc_gemm" in r[ki]]
    step = gemm[-n:]
    total = sum(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in step)
    print(json.dumps({"engine": engine, "batch": batch, "dram_bytes_per_step": total, "gemm_launches": len(step), "command": cmd,
                      "source": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the GEMM launches of one train step (ncu --set full, cold cache per launch)"}, indent=1))
    sys.exit(0)

cols = [(hdr.index(k), lab) for k, lab in WANT if k in hdr]
print("| kernel | " + " | ".join(lab for _, lab in cols) + " |")
print("|---|" + "---:|" * len(cols))
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("wd::", "").replace("<unnamed>::", "")
    m = re.search(r"kernel<([^>]*)>", r[ki])
    if m:
        name = re.sub(r"<.*", "", name) + "<" + m.group(1).replace("(int)", "") + ">"
    print("| %s | " % name + " | ".join("%s %s" % (r[i], units[i]) for i, _ in cols) + " |")
