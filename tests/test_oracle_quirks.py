"""The reference's documented quirks (SURVEY.md section "Quirks", Q1-Q12), pinned on the CPU oracle — the GPU path is checked against
the oracle, so each quirk the oracle honours is a quirk the product honours.  Reference lines are cited per test."""
from collections import OrderedDict

import numpy as np

from oracle import columns as C
from oracle import model as OM
from tests.helpers import random_raw_batch


def _conf(bn=1, lr=0.05, hidden=(8, 4), l1=None, l2=None):
    fc = OrderedDict()
    fc["age"] = dict(type="continuous", transform="min_max",
                     parameter=dict(normalization=[10, 90], boundaries=[15, 20, 25, 30, 35, 40, 45, 50, 55]))
    fc["h"] = dict(type="category", transform="hash_bucket", parameter=50)
    model = dict(linear_optimizer="tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)",
                 linear_initial_learning_rate=0.05, dnn_hidden_units=list(hidden), dnn_connected_mode="simple",
                 dnn_optimizer="Adagrad", dnn_initial_learning_rate=lr, dnn_activation_function="relu", dnn_dropout=None,
                 dnn_batch_normalization=bn, dnn_l1=l1, dnn_l2=l2)
    return fc, [(["h", "age"], 100, 1)], model


def test_q3_standalone_bucketized_column_sees_the_normalised_value():
    """build_estimator.py:127-134: the wide bucketized column wraps the NORMALISED numeric column, so with min_max every age
    lands in bucket 0 (28 -> 0.225 < 15), while a cross key buckets the raw value (build_estimator.py:145: 28 -> bucket 3)."""
    fc, cross, _ = _conf()
    cols = C.build_columns(fc, cross)
    wide = {c.name: c for c in cols[0]}
    batch = {"age": np.array([28.0, 52.0, 10.0], dtype=np.float32), "h": (np.arange(4, dtype=np.int64), np.array([1, 2, 3], dtype=np.uint64))}
    offs, ids = wide["age_bucketized"].ids(batch)
    assert ids.tolist() == [0, 0, 0]
    raw_bucket = np.searchsorted(np.asarray(fc["age"]["parameter"]["boundaries"], dtype=np.float32), batch["age"], side="right")
    assert raw_bucket.tolist() == [3, 8, 0]                     # what the cross key uses (checked end to end by the id KATs)


def test_q4_batch_norm_is_an_inference_mode_affine_after_the_activation():
    """dnn.py:96-114: tf.layers.batch_normalization(training=False-equivalent, no update ops) => h = relu(z) * gamma / sqrt(1 + 1e-3) + beta."""
    fc, cross, model = _conf(bn=1)
    om = OM.OracleModel(fc, cross, model, "deep").init(3)
    rng = np.random.default_rng(1)
    raw = random_raw_batch(fc, 16, rng)
    g = "dnn/dnn_1/hiddenlayer_0/batch_normalization/gamma"
    om.params[g][:] = 1.7
    om.params["dnn/dnn_1/hiddenlayer_0/batch_normalization/beta"][:] = -0.3
    _, cache = om.forward(raw)
    tc = cache["towers"][0]
    np.testing.assert_allclose(tc["H"][0], np.maximum(tc["Z"][0], 0) * (1.7 / np.sqrt(1.001)) - 0.3, rtol=1e-6, atol=1e-6)


def test_q11_loss_is_a_batch_sum_so_gradients_scale_with_the_batch():
    """joint.py:404-406 (_binary_logistic_head..., default SUM reduction): duplicating the batch doubles loss and gradients."""
    fc, cross, model = _conf()
    om = OM.OracleModel(fc, cross, model, "wide_deep").init(4)
    rng = np.random.default_rng(2)
    raw = random_raw_batch(fc, 8, rng)
    label = (rng.random(8) < 0.5).astype(np.float32)
    twice = {k: (np.concatenate([v[0], v[0][1:] + v[0][-1]]), np.concatenate([v[1], v[1]])) if isinstance(v, tuple) else np.concatenate([v, v])
             for k, v in raw.items()}
    _, c1 = om.forward(raw)
    _, c2 = om.forward(twice)
    l1, l2 = om.loss(c1["logits"], label), om.loss(c2["logits"], np.concatenate([label, label]))
    assert abs(l2 - 2 * l1) <= 1e-9 * abs(l1)
    g1, g2 = om.backward(c1, label), om.backward(c2, np.concatenate([label, label]))
    k = "dnn/dnn_1/hiddenlayer_0/kernel"
    np.testing.assert_allclose(np.asarray(g2[k]), 2 * np.asarray(g1[k]), rtol=1e-5, atol=1e-7)


def test_q1_learning_rate_is_constant():
    """joint.py:144-154 vs 227: the decay schedule reads a step counter that is never incremented, so the rate never changes: the
    oracle (and plan.py) carry only the initial rate; global_step advancing must not change the update for the same gradient."""
    fc, cross, model = _conf()
    om = OM.OracleModel(fc, cross, model, "deep").init(5)
    assert om.opt_dnn["lr"] == 0.05
    om.global_step = 10 ** 6
    k = "dnn/dnn_1/logits/bias"
    p0, acc0 = om.params[k].copy(), om.slots[k]["acc"].copy()
    om.apply({k: np.ones_like(p0)})
    np.testing.assert_allclose(om.params[k], p0 - 0.05 / np.sqrt(acc0 + 1.0), rtol=1e-6)


def test_q5_l1_l2_regularisers_are_inert():
    """dnn.py:30-40: kernel_regularizer objects are created but create_estimator_spec gets no regularization_losses
    (joint.py:264-269) => loss and gradients do not depend on dnn_l1 / dnn_l2."""
    fc, cross, m0 = _conf()
    _, _, m1 = _conf(l1=0.5, l2=0.5)
    a = OM.OracleModel(fc, cross, m0, "deep").init(6)
    b = OM.OracleModel(fc, cross, m1, "deep").init(6)
    rng = np.random.default_rng(3)
    raw = random_raw_batch(fc, 12, rng)
    label = (rng.random(12) < 0.5).astype(np.float32)
    la, _ = a.train_step(raw, label)
    lb, _ = b.train_step(raw, label)
    assert la == lb
    for k in a.params:
        np.testing.assert_array_equal(a.params[k], b.params[k])


def test_q12_embedding_width_heuristic_uses_the_natural_log():
    """build_estimator.py:57-59: int(2 ** ceil(ln(n ** 0.25)))."""
    for n, d in ((100, 4), (1000, 4), (10000, 8), (10 ** 6, 16), (10 ** 7, 32)):
        assert C.embedding_dim(n) == int(2 ** np.ceil(np.log(n ** 0.25))) == d
