"""Diagnostic (run by hand on a GPU box, not collected by pytest): logit / loss error of every GEMM engine against the oracle on the
benchmark model with tables scaled 1e-3, over four train steps of 2048 examples.  `python tests/engine_error_report.py`"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # repo root
from oracle import model as OM
from wide_deep_b200 import synthetic
from wide_deep_b200.model import Batch, WideDeepModel
from wide_deep_b200.plan import Plan
from tests.helpers import copy_params_to_product

fc, cross, model, emb = synthetic.criteo_conf(scale=1e-3, hidden=(1024, 512, 256))
B = 2048
n_cat = sum(1 for c in fc.values() if c["type"] == "category")
cats = [f for f, c in fc.items() if c["type"] == "category"]
dn = [f for f, c in fc.items() if c["type"] == "continuous"]
for eng in ["ffma", "tc3x", "bf16x3", "tc1x"]:
    om = OM.OracleModel(fc, cross, model, "wide_deep", embedding_dim_override=emb).init(61)
    plan = Plan(fc, cross, model, "wide_deep", max_batch=B, embedding_dim_override=emb, max_nnz=B * (len(fc) + len(cross)), max_keys=B * n_cat, gemm_engine=eng)
    pm = WideDeepModel(plan)
    copy_params_to_product(om, pm)
    out = []
    for step in range(4):
        keys, dense, label = synthetic.criteo_batch_arrays(fc, B, step=step)
        raw = {f: (np.arange(B + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
        for j, f in enumerate(dn):
            raw[f] = np.ascontiguousarray(dense[:, j])
        b = Batch(B, keys.reshape(-1), None, dense, label)
        logits, _ = pm.forward(b)
        _, cache = om.forward(raw)
        ref = cache["logits"]
        err = np.abs(logits - ref) / np.maximum(np.abs(ref), 1.0)
        loss = pm.train_step(b)
        rl, _ = om.train_step(raw, label)
        out.append((float(err.max()), float(np.sqrt((err ** 2).mean())), abs(loss - rl) / max(abs(rl), 1)))
    perr = max(float(np.max(np.abs(pm.get_tensor(n) - om.params[n]))) for n in pm.tensor_names())
    print(eng, " ".join("step%d max %.2e rms %.2e loss %.1e |" % (i, a, b_, c) for i, (a, b_, c) in enumerate(out)), "param max abs %.2e" % perr, flush=True)
    del pm
