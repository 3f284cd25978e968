"""Diagnostic (by hand on a GPU box): how far a TRAINING RUN of each GEMM engine drifts from the oracle on the benchmark shape
(Criteo towers 845-1024-512-256, tables scaled 1e-3, 2048 examples per step).  Two protocols:
  init    50 steps from the TF initialisers: with the reference's SUM-reduced loss and Adagrad(0.05) the first steps are a violent
          transient (oracle losses 1.5e3 -> 4.2e5 -> 1.6e4 -> 4.9e2 -> 4.7e3 ...), which amplifies product rounding
  warm    the oracle alone trains 10 steps first; its state is copied to the GPU model; then 50 steps in the settled regime
Prints the worst per-step relative loss error and the fresh-batch logit error after the run.
`python tests/engine_drift_report.py [engines...]`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as OM  # noqa: E402
from tests.helpers import copy_params_to_product  # noqa: E402
from wide_deep_b200 import synthetic  # noqa: E402
from wide_deep_b200.model import Batch, WideDeepModel  # noqa: E402
from wide_deep_b200.plan import Plan  # noqa: E402

B = 2048
fc, cross, model, emb = synthetic.criteo_conf(scale=1e-3)
n_cat = sum(1 for c in fc.values() if c["type"] == "category")
cats = [f for f, c in fc.items() if c["type"] == "category"]
dn = [f for f, c in fc.items() if c["type"] == "continuous"]


def batch(step):
    keys, dense, label = synthetic.criteo_batch_arrays(fc, B, step=step)
    raw = {f: (np.arange(B + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
    for j, f in enumerate(dn):
        raw[f] = np.ascontiguousarray(dense[:, j])
    return raw, label, Batch(B, keys.reshape(-1), None, dense, label)


def run(engine, warm):
    om = OM.OracleModel(fc, cross, model, "wide_deep", embedding_dim_override=emb).init(11)
    for s in range(warm):
        raw, label, _ = batch(1000 + s)
        om.train_step(raw, label)
    plan = Plan(fc, cross, model, "wide_deep", max_batch=B, embedding_dim_override=emb, gemm_engine=engine,
                max_nnz=B * (len(fc) + len(cross)), max_keys=B * n_cat)
    pm = WideDeepModel(plan)
    copy_params_to_product(om, pm)
    errs = []
    for s in range(50):
        raw, label, b = batch(s)
        loss = pm.train_step(b)
        ref, _ = om.train_step(raw, label)
        errs.append(abs(loss - ref) / max(abs(ref), 1.0))
    raw, label, b = batch(999)
    logits, _ = pm.forward(b)
    _, cache = om.forward(raw)
    ref = cache["logits"]
    e = np.abs(logits - ref) / np.maximum(np.abs(ref), 1.0)
    pm.close()
    print("engine %-7s %-5s worst loss err %.3g (step %d), steps>1e-4: %d, first 8: %s | final logits max %.3g rms %.3g" % (
        engine, "warm" if warm else "init", max(errs), int(np.argmax(errs)), sum(1 for x in errs if x > 1e-4),
        " ".join("%.1e" % x for x in errs[:8]), float(e.max()), float(np.sqrt((e ** 2).mean()))), flush=True)


if __name__ == "__main__":
    engines = sys.argv[1:] or ["ffma", "tc3x", "bf16x3"]
    for warm in (0, 10):
        for eng in engines:
            run(eng, warm)
