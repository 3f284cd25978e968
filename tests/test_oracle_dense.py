"""Independent cross-check of the oracle's floating-point math against torch-CPU autograd.

torch supplies: EmbeddingBag(mode='mean'), linear layers, binary_cross_entropy_with_logits(reduction='sum'),
autograd gradients, torch.optim.Adagrad(initial_accumulator_value=0.1, eps=0).  The FTRL update is checked
against a scalar transcription of TensorFlow's training_ops.cc formula.
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import columns as C
from oracle import model as OM
from tests.helpers import random_raw_batch


def conf(mode="simple", act="relu", bn=1, hidden=(16, 8)):
    fc = OrderedDict()
    fc["h1"] = dict(type="category", transform="hash_bucket", parameter=50)
    fc["h2"] = dict(type="category", transform="hash_bucket", parameter=400)
    fc["v1"] = dict(type="category", transform="vocab", parameter=[0, 1, 2])
    fc["id1"] = dict(type="category", transform="identity", parameter=7)
    fc["x1"] = dict(type="continuous", transform="min_max", parameter=dict(normalization=[10, 90], boundaries=[20, 40, 60]))
    cross = [(["h1", "h2"], 300, 1), (["id1", "x1"], 40, 0)]
    model = dict(linear_optimizer="Ftrl", linear_initial_learning_rate=0.05, dnn_hidden_units=list(hidden),
                 dnn_connected_mode=mode, dnn_optimizer="Adagrad", dnn_initial_learning_rate=0.05,
                 dnn_activation_function=act, dnn_dropout=None, dnn_batch_normalization=bn)
    return fc, cross, model


def torch_forward(om, raw, P, drop_step=None):
    """Re-implementation with torch ops, parameters P (dict name -> torch tensor requiring grad)."""
    ids = om.transform(raw)
    B = len(next(iter(ids.values()))[0]) - 1
    logit = torch.zeros(B, dtype=torch.float64)
    if om.use_wide:
        logit = logit + P["linear/linear_model/bias_weights"][0]
        for c in om.wide_cols:
            offs, cid = ids[c.name]
            if len(cid):
                bag = torch.nn.functional.embedding_bag(torch.from_numpy(cid), P[om.wname(c)].unsqueeze(1),
                                                        torch.from_numpy(offs[:-1]), mode="sum")
                logit = logit + bag[:, 0]
    if om.use_deep:
        cols = []
        for c in om.deep_cols:
            if isinstance(c, C.Numeric):
                cols.append(torch.from_numpy(c.values(raw).astype(np.float64)).unsqueeze(1))
            elif isinstance(c, C.Indicator):
                offs, cid = ids[c.cat.name]
                oh = torch.zeros(B, c.width, dtype=torch.float64)
                for b in range(B):
                    for i in cid[offs[b]:offs[b + 1]]:
                        oh[b, i] += 1
                cols.append(oh)
            else:
                offs, cid = ids[c.cat.name]
                cols.append(torch.nn.functional.embedding_bag(torch.from_numpy(cid), P[om.ename(c)],
                                                              torch.from_numpy(offs[:-1]), mode="mean"))
        x = torch.cat(cols, 1)
        for t in range(len(om.towers)):
            hu = om.towers[t]
            srcs = OM.layer_sources(om.modes[t], len(hu))
            H = []
            pick = lambda s: x if s == "x" else H[s]
            act = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "elu": torch.nn.functional.elu,
                   "softplus": torch.nn.functional.softplus,
                   "crelu": lambda z: torch.cat([torch.relu(z), torch.relu(-z)], 1)}[om.act]        # tf.nn.crelu
            for l in range(len(hu)):
                sc = "dnn/dnn_%d/hiddenlayer_%d" % (t + 1, l)
                inp = torch.cat([pick(s) for s in srcs[l]], 1)
                a = act(inp @ P[sc + "/kernel"] + P[sc + "/bias"])
                if drop_step is not None and om.dropout > 0:      # tf.layers.dropout, the oracle's counter-based keep mask
                    keep = OM.drop_keep(om.dropout_seed, drop_step, t * 64 + l, a.shape[0], a.shape[1], om.dropout)
                    a = a * torch.from_numpy(keep.astype(np.float64) / (1.0 - float(np.float32(om.dropout))))
                if om.bn:
                    a = a * (P[sc + "/batch_normalization/gamma"] / np.sqrt(1 + 1e-3)) + P[sc + "/batch_normalization/beta"]
                H.append(a)
            sc = "dnn/dnn_%d/logits" % (t + 1)
            inp = torch.cat([pick(s) for s in srcs[-1]], 1)
            logit = logit + (inp @ P[sc + "/kernel"] + P[sc + "/bias"])[:, 0]
    return logit


@pytest.mark.parametrize("mode,act", [("simple", "relu"), ("dense", "tanh"), ("resnet", "sigmoid"), ("first_dense", "elu"),
                                      ("last_dense", "softplus"), ("simple", "crelu"), ("dense", "crelu")])
def test_oracle_logits_and_grads_match_torch(mode, act):
    fc, cross, model = conf(mode, act)
    rng = np.random.default_rng(4)
    om = OM.OracleModel(fc, cross, model, "wide_deep").init(2)
    for c in om.wide_cols:
        om.params[om.wname(c)][:] = rng.standard_normal(c.num_buckets).astype(np.float32) * 0.1
    B = 40
    raw = random_raw_batch(fc, B, rng)
    label = (rng.random(B) < 0.4).astype(np.float32)
    weight = rng.random(B).astype(np.float32) + 0.5
    P = {k: torch.tensor(v.astype(np.float64), requires_grad=True) for k, v in om.params.items()}
    logit = torch_forward(om, raw, P)
    loss = (torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.from_numpy(label.astype(np.float64)),
                                                                 reduction="none") * torch.from_numpy(weight.astype(np.float64))).sum()
    loss.backward()
    logits, cache = om.forward(raw)
    np.testing.assert_allclose(cache["logits"], logit.detach().numpy(), rtol=1e-10, atol=1e-10)
    assert abs(om.loss(cache["logits"], label, weight) - float(loss)) < 1e-8 * max(1.0, abs(float(loss)))
    grads = om.backward(cache, label, weight)
    for name, g in grads.items():
        ref = P[name].grad.numpy()
        if isinstance(g, tuple):
            rows, gr = g
            dense = np.zeros_like(ref).reshape(ref.shape[0], -1)
            dense[rows] = np.asarray(gr).reshape(len(rows), -1)
            g = dense.reshape(ref.shape)
        np.testing.assert_allclose(np.asarray(g).reshape(ref.shape), ref, rtol=1e-8, atol=1e-10, err_msg=name)


def test_adagrad_matches_torch_optim():
    fc, cross, model = conf()
    rng = np.random.default_rng(5)
    om = OM.OracleModel(fc, cross, model, "deep").init(3)
    P = {k: torch.tensor(v.copy(), requires_grad=True) for k, v in om.params.items()}
    opt = torch.optim.Adagrad(P.values(), lr=0.05, initial_accumulator_value=0.1, eps=0)
    for _ in range(2):
        raw = random_raw_batch(fc, 30, rng)
        label = (rng.random(30) < 0.4).astype(np.float32)
        P64 = {k: v.double() for k, v in P.items()}
        opt.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(torch_forward(om, raw, P64), torch.from_numpy(label).double(), reduction="sum")
        loss.backward()
        # torch's dense Adagrad touches every row (g = 0 leaves w unchanged), same fixed point as the sparse apply
        opt.step()
        om.train_step(raw, label)
    for k in P:
        np.testing.assert_allclose(om.params[k], P[k].detach().numpy(), rtol=2e-5, atol=2e-6, err_msg=k)


def test_ftrl_formula():
    """Scalar transcription of tensorflow/core/kernels/training_ops.cc FtrlCompute (lr_power = -0.5)."""
    o = dict(kind="ftrl", lr=0.1, l1=0.5, l2=1.0, lr_power=-0.5, init_acc=0.1)
    rng = np.random.default_rng(6)
    w, n, z = 0.0, 0.1, 0.0
    om = OM.OracleModel(*conf(), "wide")
    om.params = {"linear/t": np.zeros(1, dtype=np.float32)}
    om.opt_lin = o
    om.reset_slots()
    for _ in range(20):
        g = float(rng.standard_normal() * 3)
        n1 = n + g * g
        z = z + g - (np.sqrt(n1) - np.sqrt(n)) / o["lr"] * w
        w = 0.0 if abs(z) <= o["l1"] else (np.sign(z) * o["l1"] - z) / (np.sqrt(n1) / o["lr"] + 2 * o["l2"])
        n = n1
        om.apply({"linear/t": np.array([g])})
        assert abs(om.params["linear/t"][0] - w) < 1e-5 * max(1, abs(w))
        assert abs(om.slots["linear/t"]["n"][0] - n) < 1e-4 * n


def test_eval_metrics_against_sklearn():
    from sklearn.metrics import roc_auc_score
    from oracle.metrics import EvalAccumulator
    rng = np.random.default_rng(8)
    logits = rng.standard_normal(4000).astype(np.float32) * 2
    labels = (rng.random(4000) < 1 / (1 + np.exp(-logits))).astype(np.float32)
    acc = EvalAccumulator()
    for i in range(0, 4000, 500):
        acc.update(logits[i:i + 500], labels[i:i + 500])
    r = acc.result()
    assert abs(r["auc"] - roc_auc_score(labels, logits)) < 2e-3       # 200-threshold trapezoid vs exact
    p = 1 / (1 + np.exp(-logits.astype(np.float64)))
    assert abs(r["accuracy"] - np.mean((logits > 0) == (labels > 0.5))) < 1e-12
    assert abs(r["prediction/mean"] - p.mean()) < 1e-6
    assert abs(r["label/mean"] - labels.mean()) < 1e-12
    lm = float(labels.astype(np.float64).mean())
    assert abs(r["accuracy_baseline"] - max(lm, 1 - lm)) < 1e-12
    assert set(r) == {"accuracy", "accuracy_baseline", "auc", "auc_precision_recall", "average_loss", "label/mean", "loss",
                      "precision", "prediction/mean", "recall"}


def test_dropout_forward_and_grads_match_torch_autograd():
    """dnn_dropout: the keep mask comes from oracle.model.drop_keep (counter-based, what the CUDA path reproduces); scaling by
    1 / (1 - rate), applied after the activation and before batch norm, TRAIN only (reference dnn.py:111-112)."""
    fc, cross, model = conf("dense", "relu")
    model["dnn_dropout"] = 0.3
    rng = np.random.default_rng(8)
    om = OM.OracleModel(fc, cross, model, "wide_deep").init(4)
    om.global_step = 5
    B = 40
    raw = random_raw_batch(fc, B, rng)
    label = (rng.random(B) < 0.4).astype(np.float32)
    P = {k: torch.tensor(v.astype(np.float64), requires_grad=True) for k, v in om.params.items()}
    logit = torch_forward(om, raw, P, drop_step=5)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logit, torch.from_numpy(label.astype(np.float64)), reduction="sum")
    loss.backward()
    logits, cache = om.forward(raw, train=True)
    np.testing.assert_allclose(cache["logits"], logit.detach().numpy(), rtol=1e-6, atol=1e-6)
    ev, _ = om.forward(raw)                                         # eval / predict: no dropout
    assert np.abs(ev - logits).max() > 1e-3
    grads = om.backward(cache, label)
    for name, g in grads.items():
        ref = P[name].grad.numpy()
        if isinstance(g, tuple):
            rows, gr = g
            dense = np.zeros_like(ref).reshape(ref.shape[0], -1)
            dense[rows] = np.asarray(gr).reshape(len(rows), -1)
            g = dense.reshape(ref.shape)
        np.testing.assert_allclose(np.asarray(g).reshape(ref.shape), ref, rtol=1e-5, atol=1e-7, err_msg=name)
