#!/usr/bin/env python
"""Host-core probe for the CPU arm (by hand on the GPU box): what the box offers (logical CPUs, affinity, cgroup quota) and how
the optimised CPU step (oracle/fast.py) scales with torch threads.  `python tools/cpu_probe.py`"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cgroup_quota():
    for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            t = open(p).read().split()
            if p.endswith("cpu.max"):
                return None if t[0] == "max" else float(t[0]) / float(t[1])
            q = float(t[0])
            return None if q <= 0 else q / float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        except Exception:
            continue
    return None


def main():
    print("os.cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "cgroup quota (cpus)", cgroup_quota(),
          "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), flush=True)
    try:
        print(open("/proc/loadavg").read().strip(), flush=True)
    except Exception:
        pass
    import torch
    from oracle import fast as OF, model as OM
    from wide_deep_b200 import synthetic
    print("torch threads default", torch.get_num_threads(), "interop", torch.get_num_interop_threads(), flush=True)
    fc, cross, model, emb = synthetic.criteo_conf(scale=1e-2)
    B = 8192
    cats = [f for f, c in fc.items() if c["type"] == "category"]
    dn = [f for f, c in fc.items() if c["type"] == "continuous"]
    om = OM.OracleModel(fc, cross, model, "wide_deep", embedding_dim_override=emb, acc=np.float32).init(1)
    keys, dense, label = synthetic.criteo_batch_arrays(fc, B, step=0)
    raw = {f: (np.arange(B + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
    for j, f in enumerate(dn):
        raw[f] = np.ascontiguousarray(dense[:, j])
    n = os.cpu_count() or 1
    for t in sorted({n, max(1, n // 2), max(1, n // 4), max(1, n // 8), 16, 8}, reverse=True):
        if t > n:
            continue
        fm = OF.FastCpuModel(om, threads=t)
        fm.train_step(raw, label)
        t0 = time.perf_counter()
        for _ in range(2):
            fm.train_step(raw, label)
        dt = (time.perf_counter() - t0) / 2
        a = torch.randn(8192, 1024)
        w = torch.randn(1024, 1024)
        t1 = time.perf_counter()
        for _ in range(5):
            a @ w
        mm = (time.perf_counter() - t1) / 5
        print("threads %3d: step %.3f s = %.0f examples/s; sgemm 8192x1024x1024 %.1f GFLOP/s" % (t, dt, B / dt, 2 * 8192 * 1024 * 1024 / mm / 1e9), flush=True)


if __name__ == "__main__":
    main()
