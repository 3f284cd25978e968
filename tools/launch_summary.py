#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last complete train step.
    python tools/launch_summary.py gpurun_out/launches.csv [--md]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
names = [re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("wd::", "").replace("(anonymous namespace)::", "") for r in rows]
t = [float(r[-1]) for r in rows]
streams = [r[6] for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith("ids_count")]
a, b = idx[-2], idx[-1]
agg = collections.OrderedDict()
for i in range(a, b):
    k = (names[i], streams[i])
    agg.setdefault(k, [0.0, 0])
    agg[k][0] += t[i]
    agg[k][1] += 1
tot = sum(v[0] for v in agg.values())
md = "--md" in sys.argv
if md:
    print("| kernel | stream | launches | us / step | share |\n|---|---|---:|---:|---:|")
for (k, s), v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if md:
        print("| `%s` | %s | %d | %.1f | %.1f%% |" % (k[:70], s, v[1], v[0] / 1e3, 100 * v[0] / tot))
    else:
        print("%-64s s%-3s %3d %8.1f us %5.1f%%" % (k[:64], s, v[1], v[0] / 1e3, 100 * v[0] / tot))
by_stream = collections.defaultdict(float)
for (k, s), v in agg.items():
    by_stream[s] += v[0]
print(("\n" if md else "") + "total %.1f us over %d launches; per stream: %s" % (tot / 1e3, b - a, ", ".join("%s: %.1f us" % (s, v / 1e3) for s, v in sorted(by_stream.items()))))
