"""The optimised CPU restatement bench.py times (oracle/fast.py: torch-CPU kernels) against the checker (oracle/model.py):
same losses and the same parameters after several steps, for every connection mode and optimizer family."""
from collections import OrderedDict

import numpy as np
import pytest

from oracle import fast as OF, model as OM
from tests.helpers import random_raw_batch


def _conf(mode="simple", act="relu", bn=1, dnn_opt="Adagrad",
          lin_opt="tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)", hidden=(48, 32, 16)):
    fc = OrderedDict()
    fc["h1"] = dict(type="category", transform="hash_bucket", parameter=1000)
    fc["h2"] = dict(type="category", transform="hash_bucket", parameter=37)
    fc["v1"] = dict(type="category", transform="vocab", parameter=[0, 1, 2, 3, 4])
    fc["id1"] = dict(type="category", transform="identity", parameter=15)
    fc["x1"] = dict(type="continuous", transform="min_max", parameter=dict(normalization=[10, 90], boundaries=[15, 25, 35, 45]))
    fc["x2"] = dict(type="continuous", transform="standard", parameter=dict(normalization=[40.0, 30.0], boundaries=None))
    cross = [(["h1", "h2"], 1000, 1), (["h1", "x1"], 500, 1), (["id1", "v1"], 100, 0)]
    model = dict(linear_optimizer=lin_opt, linear_initial_learning_rate=0.05, dnn_hidden_units=list(hidden),
                 dnn_connected_mode=mode, dnn_optimizer=dnn_opt, dnn_initial_learning_rate=0.05,
                 dnn_activation_function=act, dnn_dropout=None, dnn_batch_normalization=bn)
    return fc, cross, model


@pytest.mark.parametrize("mode,model_type,dnn_opt,lin_opt", [
    ("simple", "wide_deep", "Adagrad", "Ftrl"),
    ("dense", "wide_deep", "Adagrad", "tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)"),
    ("resnet", "deep", "tf.train.FtrlOptimizer(learning_rate=0.05,l1_regularization_strength=0.001,l2_regularization_strength=0.01)", "SGD"),
    ("first_dense", "wide_deep", "tf.train.GradientDescentOptimizer(learning_rate=0.0002)", "Adagrad"),
    ("last_dense", "wide_deep", "Adagrad", "Ftrl"),
    ("simple", "wide", "Adagrad", "Ftrl"),
])
def test_fast_cpu_step_equals_checker(mode, model_type, dnn_opt, lin_opt):
    fc, cross, model = _conf(mode=mode, dnn_opt=dnn_opt, lin_opt=lin_opt)
    B = 96
    a = OM.OracleModel(fc, cross, model, model_type).init(3)
    b = OM.OracleModel(fc, cross, model, model_type).init(3)
    fb = OF.FastCpuModel(b)
    rng = np.random.default_rng(5)
    for step in range(3):
        raw = random_raw_batch(fc, B, rng)
        label = (rng.random(B) < 0.3).astype(np.float32)
        weight = (rng.random(B).astype(np.float32) + 0.5) if step == 1 else None
        la, _ = a.train_step(raw, label, weight)
        lb, _ = fb.train_step(raw, label, weight)
        assert abs(la - lb) <= 2e-5 * max(abs(la), 1.0), (step, la, lb)
    for name in a.params:
        sc = max(float(np.abs(a.params[name]).max()), 1e-3)
        assert np.max(np.abs(a.params[name] - b.params[name])) <= 2e-4 * sc, name
        for k in a.slots[name]:
            s2 = max(float(np.abs(a.slots[name][k]).max()), 1e-3)
            assert np.max(np.abs(a.slots[name][k] - b.slots[name][k])) <= 5e-4 * s2, (name, k)
    raw = random_raw_batch(fc, B, rng)
    la, _ = a.forward(raw)
    np.testing.assert_allclose(fb.forward(raw), la, rtol=0, atol=2e-4)
