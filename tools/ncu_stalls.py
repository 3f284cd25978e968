#!/usr/bin/env python
"""Top stalled SASS lines of one kernel of an ncu report:  ncu -i X.ncu-rep --page source --csv --kernel-id :::N > src.csv ; python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
I = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[I["# Samples"]].isdigit()]
tot = sum(int(r[I["# Samples"]]) for r in data)
inst = sum(int(r[I["Instructions Executed"]]) for r in data)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[I[s]]) for r in data) for s in stalls}
print("samples %d, warp instructions %d, SASS lines %d" % (tot, inst, len(data)))
print("stall reasons:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -int(r[I["# Samples"]]))[:top]:
    st = sorted([(int(r[I[s]]), s[6:]) for s in stalls], reverse=True)[:2]
    print("%6s %8s  %-64s %s" % (r[I["# Samples"]], r[I["Instructions Executed"]], r[I["Source"]].strip()[:64], st))
