"""Diagnostic (by hand on a GPU box): sharded run (G virtual ranks on cuda:0) vs the oracle, tensor by tensor after every step.
`python tests/shard_debug.py [G] [model_type]`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as OM  # noqa: E402
from tests.helpers import random_raw_batch, to_product_batch  # noqa: E402
from tests.test_gpu_parity import small_conf  # noqa: E402
from tests.test_gpu_sharded import make_group  # noqa: E402
from tests.test_parallel_gloo import slice_raw  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model_type = sys.argv[2] if len(sys.argv) > 2 else "wide_deep"
fc, cross, model = small_conf(hidden=(64, 32))
per = 40
B = per * G
om = OM.OracleModel(fc, cross, model, model_type).init(3 + G)
rng = np.random.default_rng(100 + G)
if om.use_wide:
    for c in om.wide_cols:
        om.params[om.wname(c)][:] = rng.standard_normal(c.num_buckets).astype(np.float32) * 0.1
grp = make_group(fc, cross, model, model_type, G, per, om, dense_rows=400)
plan0 = grp.models[0].plan
for step in range(4):
    raw = random_raw_batch(fc, B, rng)
    label = (rng.random(B) < 0.3).astype(np.float32)
    weight = (rng.random(B).astype(np.float32) + 0.5) if step == 1 else None
    shards = [to_product_batch(plan0, slice_raw(raw, r * per, (r + 1) * per), label[r * per:(r + 1) * per],
                               None if weight is None else weight[r * per:(r + 1) * per]) for r in range(G)]
    loss = grp.train_step(shards)
    ref, _ = om.train_step(raw, label, weight)
    print("step %d loss %.6f oracle %.6f rel %.2e" % (step, loss, ref, abs(loss - ref) / max(abs(ref), 1)), flush=True)
    rows = []
    for name in grp.models[0].tensor_names():
        got, exp = grp.get_tensor(name), om.params[name]
        scale = max(float(np.abs(exp).max()), 1e-3)
        d = float(np.max(np.abs(got - exp))) / scale
        sd = []
        for si, key in enumerate([k for k in ("acc", "n", "z") if k in om.slots[name]]):
            g2, e2 = grp.get_tensor(name, slot=si + 1), om.slots[name][key]
            sd.append(float(np.max(np.abs(g2 - e2))) / max(float(np.abs(e2).max()), 1e-3))
        rows.append((max([d] + sd), d, sd, name, grp.models[0].plan.is_sharded_tensor(name), int(np.argmax(np.abs(got - exp).reshape(-1)))))
    rows.sort(reverse=True)
    for worst, d, sd, name, sh, arg in rows[:8]:
        print("   %-90s %s w %.2e slots %s argmax %d" % (name[-90:], "SHARDED" if sh else "replic.", d, ["%.1e" % x for x in sd], arg), flush=True)
