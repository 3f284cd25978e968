"""Diagnostic (by hand on a GPU box): max relative logit error of the tensor-core engines per dnn_connected_mode, on the small
unit-test towers (128-96-64, the worst-conditioned case seen) and on bench-sized towers, identical parameters, several seeds.
`python tests/engine_mode_report.py`"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as OM  # noqa: E402
from tests.helpers import copy_params_to_product, random_raw_batch, to_product_batch  # noqa: E402
from tests.test_gpu_parity import small_conf  # noqa: E402
from wide_deep_b200.model import WideDeepModel  # noqa: E402
from wide_deep_b200.plan import Plan  # noqa: E402

B = 512
for hidden in [(128, 96, 64), (512, 256, 128)]:
    for mode in ["simple", "first_dense", "last_dense", "dense", "resnet"]:
        for eng in ["bf16x3", "tc3x"]:
            worst = 0.0
            for seed in range(4):
                fc, cross, model = small_conf(hidden=hidden, mode=mode)
                om = OM.OracleModel(fc, cross, model, "wide_deep").init(seed)
                plan = Plan(fc, cross, model, "wide_deep", max_batch=B, max_nnz=B * 64, max_keys=B * 64, gemm_engine=eng)
                pm = WideDeepModel(plan)
                copy_params_to_product(om, pm)
                rng = np.random.default_rng(100 + seed)
                raw = random_raw_batch(fc, B, rng)
                logits, _ = pm.forward(to_product_batch(plan, raw, (rng.random(B) < 0.3).astype(np.float32)))
                _, cache = om.forward(raw)
                ref = cache["logits"]
                worst = max(worst, float((np.abs(logits - ref) / np.maximum(np.abs(ref), 1.0)).max()))
                pm.close()
            print("hidden %-16s mode %-12s engine %-7s max rel logit err %.3g" % (hidden, mode, eng, worst), flush=True)
