"""Row-sharded tables (wide_deep_b200/csrc/shard.cu): G ranks, each holding 1/G of the rows of every large table and a
1/G row shard of the batch, must compute what the oracle (and a single GPU) computes on the whole batch.

The G ranks are G model handles in ONE process on cuda:0 (`LocalShardGroup`: the phases of a step are ordered with events),
so the whole exchange — routing by owner, owner-side pooling into the requesters' buffers, the combine, the owners' pulled
gradient sums + optimizers, the two-shot all-reduce of the dense gradients — runs on the single-GPU test box.  The same
kernels run under flag barriers between processes (tests/_shard_worker.py, below: two processes sharing cuda:0 through CUDA
IPC, and on >= 2 GPUs one process per GPU).

Miniatures of BASELINE.json configs[2] (replicated small tables + sharded large ones), configs[3] (one multihot slot, ResDnn,
row-sharded) and configs[4] (wide-only hashed crosses + FTRL, row-sharded)."""
import os
import subprocess
import sys
from collections import OrderedDict

import numpy as np
import pytest

from oracle import model as OM
from tests.helpers import random_raw_batch, to_product_batch
from tests.test_gpu_parity import small_conf
from tests.test_parallel_gloo import slice_raw
from wide_deep_b200.model import Batch, WideDeepModel
from wide_deep_b200.plan import Plan
from wide_deep_b200.sharded import LocalShardGroup

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-4


def make_group(fc, cross, model, model_type, G, per_rank, om, dense_rows, emb_dim=None, engine="ffma", max_ids=64):
    models = []
    for r in range(G):
        plan = Plan(fc, cross, model, model_type, max_batch=per_rank, embedding_dim_override=emb_dim, gemm_engine=engine,
                    max_nnz=per_rank * max_ids, max_keys=per_rank * max_ids, dense_exchange_max_rows=dense_rows,
                    shard_world=G, shard_rank=r, shard_slack=float(G))          # slack G: a rank could own every id of a tiny test batch
        models.append(WideDeepModel(plan))
    grp = LocalShardGroup(models)
    for name in models[0].tensor_names():
        grp.set_tensor(name, om.params[name])
        slots = om.slots[name]
        if "acc" in slots:
            grp.set_tensor(name, slots["acc"], slot=1)
        if "n" in slots:
            grp.set_tensor(name, slots["n"], slot=1)
            grp.set_tensor(name, slots["z"], slot=2)
    return grp


def compare_params(grp, om, tol=2e-4, slot_tol=5e-4):
    for name in grp.models[0].tensor_names():
        got, exp = grp.get_tensor(name), om.params[name]
        scale = max(float(np.abs(exp).max()), 1e-3)
        assert np.max(np.abs(got - exp)) <= tol * scale, "%s: max abs diff %g (scale %g)" % (name, np.max(np.abs(got - exp)), scale)
        for si, key in enumerate([k for k in ("acc", "n", "z") if k in om.slots[name]]):
            g2, e2 = grp.get_tensor(name, slot=si + 1), om.slots[name][key]
            sc = max(float(np.abs(e2).max()), 1e-3)
            assert np.max(np.abs(g2 - e2)) <= slot_tol * sc, "%s slot %s" % (name, key)


@pytest.mark.parametrize("G", [2, 3, 4])
@pytest.mark.parametrize("model_type", ["wide_deep", "deep", "wide"])
def test_sharded_ranks_equal_oracle_on_whole_batch(G, model_type):
    """Sharded: h1 (1000 rows), h3 (200 000), the crosses of 1000 / 500 / 2000 buckets, their embeddings; replicated + dense
    block: everything <= 400 rows.  Multihot bags, empty bags, dropped ids; 3 train steps, then a fresh forward."""
    fc, cross, model = small_conf(hidden=(64, 32))
    per = 40
    B = per * G
    om = OM.OracleModel(fc, cross, model, model_type).init(3 + G)
    rng = np.random.default_rng(100 + G)
    if om.use_wide:                                   # zero-initialised wide weights carry no signal: give them some
        for c in om.wide_cols:
            om.params[om.wname(c)][:] = rng.standard_normal(c.num_buckets).astype(np.float32) * 0.1
    grp = make_group(fc, cross, model, model_type, G, per, om, dense_rows=400)
    assert any(grp.models[0].plan.is_sharded_tensor(n) for n in grp.models[0].tensor_names())
    plan0 = grp.models[0].plan
    for step in range(3):
        raw = random_raw_batch(fc, B, rng)
        label = (rng.random(B) < 0.3).astype(np.float32)
        weight = (rng.random(B).astype(np.float32) + 0.5) if step == 1 else None
        shards = [to_product_batch(plan0, slice_raw(raw, r * per, (r + 1) * per), label[r * per:(r + 1) * per],
                                   None if weight is None else weight[r * per:(r + 1) * per]) for r in range(G)]
        if step == 0:                                 # forward parity first (identical parameters)
            logits = np.concatenate(grp.forward(shards))
            _, cache = om.forward(raw)
            np.testing.assert_array_less(np.abs(logits - cache["logits"]), RTOL * np.maximum(np.abs(cache["logits"]), 1.0))
        loss = grp.train_step(shards)
        ref, _ = om.train_step(raw, label, weight)
        assert abs(loss - ref) <= RTOL * max(abs(ref), 1.0), "step %d: loss %g vs oracle %g" % (step, loss, ref)
    compare_params(grp, om)
    raw = random_raw_batch(fc, B, rng)
    label = (rng.random(B) < 0.3).astype(np.float32)
    shards = [to_product_batch(plan0, slice_raw(raw, r * per, (r + 1) * per), label[r * per:(r + 1) * per]) for r in range(G)]
    logits = np.concatenate(grp.forward(shards))
    _, cache = om.forward(raw)
    np.testing.assert_array_less(np.abs(logits - cache["logits"]), 5 * RTOL * np.maximum(np.abs(cache["logits"]), 1.0))


def test_sharded_multihot_slot_resdnn():
    """BASELINE.json configs[3] in miniature, row-sharded over 4 ranks: one hashed multihot slot (Poisson(30) ids per example,
    skewed, duplicates inside bags), 64-wide embedding, ResDnn 4 x 64.  Owner-side partial pooling of multi-id bags, bags
    spread over all owners, hot rows (chunked sums) on the owners."""
    from oracle import hashing as OH
    G, per = 4, 64
    B = G * per
    fc = OrderedDict()
    fc["tags"] = dict(type="category", transform="hash_bucket", parameter=5000)
    fc["x"] = dict(type="continuous", transform="standard", parameter=dict(normalization=[0.0, 1.0], boundaries=[-1, 0, 1]))
    model = dict(linear_optimizer="Ftrl", linear_initial_learning_rate=0.05, dnn_hidden_units=[64, 64, 64, 64],
                 dnn_connected_mode="resnet", dnn_optimizer="Adagrad", dnn_initial_learning_rate=0.05,
                 dnn_activation_function="relu", dnn_dropout=None, dnn_batch_normalization=1)
    om = OM.OracleModel(fc, [], model, "wide_deep", embedding_dim_override=64).init(71)
    grp = make_group(fc, [], model, "wide_deep", G, per, om, dense_rows=100, emb_dim=64, max_ids=128)
    rng = np.random.default_rng(73)
    vocab = OH.fingerprint64_tokens(["t%d" % i for i in range(2000)])
    for step in range(3):
        lens = np.clip(rng.poisson(30, size=B), 1, 96)
        offs = np.zeros(B + 1, dtype=np.int64)
        offs[1:] = np.cumsum(lens)
        fps = vocab[(rng.zipf(1.3, size=int(offs[-1])) - 1) % len(vocab)]
        x = rng.standard_normal(B).astype(np.float32)
        label = (rng.random(B) < 0.3).astype(np.float32)
        raw = {"tags": (offs, fps), "x": x}
        shards = []
        for r in range(G):
            lo, hi = r * per, (r + 1) * per
            shards.append(Batch(per, fps[offs[lo]:offs[hi]], (offs[lo:hi + 1] - offs[lo]).astype(np.int32), x[lo:hi].reshape(per, 1), label[lo:hi]))
        loss = grp.train_step(shards)
        ref, _ = om.train_step(raw, label)
        assert abs(loss - ref) <= RTOL * max(abs(ref), 1.0), "step %d loss %g vs %g" % (step, loss, ref)
    name = "dnn/input_from_feature_columns/input_layer/tags_embedding/embedding_weights"
    assert grp.models[0].plan.is_sharded_tensor(name)
    got, exp = grp.get_tensor(name), om.params[name]
    assert np.max(np.abs(got - exp)) <= 0.03 * 0.05


def test_sharded_wide_only_ftrl():
    """BASELINE.json configs[4] in miniature, row-sharded over 4 ranks: 'wide' model, 15 hashed crosses into large bucket
    spaces, FTRL: ids bit-exact per rank, FTRL state (w, n, z) of every shard after three steps."""
    G, per = 4, 256
    B = G * per
    fc = OrderedDict()
    for i in range(6):
        fc["k%d" % i] = dict(type="category", transform="hash_bucket", parameter=1000 + 17 * i)
    cross = [(["k%d" % a, "k%d" % b], 200000 + 1000 * (a + b), 0) for a in range(6) for b in range(a + 1, 6)]
    model = dict(linear_optimizer="tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)",
                 linear_initial_learning_rate=0.05, dnn_hidden_units=[8], dnn_connected_mode="simple", dnn_optimizer="Adagrad",
                 dnn_initial_learning_rate=0.05, dnn_activation_function="relu", dnn_dropout=None, dnn_batch_normalization=0)
    rng = np.random.default_rng(81)
    om = OM.OracleModel(fc, cross, model, "wide").init(83)
    grp = make_group(fc, cross, model, "wide", G, per, om, dense_rows=1005, max_ids=32)     # k0 (1000 rows) replicated, the rest sharded
    plan0 = grp.models[0].plan
    for step in range(3):
        raw = random_raw_batch(fc, B, rng, multihot_max=1, na_rate=0.05)
        label = (rng.random(B) < 0.3).astype(np.float32)
        shards = [to_product_batch(plan0, slice_raw(raw, r * per, (r + 1) * per), label[r * per:(r + 1) * per]) for r in range(G)]
        loss = grp.train_step(shards)
        ref, _ = om.train_step(raw, label)
        assert abs(loss - ref) <= RTOL * max(abs(ref), 1.0), "step %d loss %g vs %g" % (step, loss, ref)
    for name in grp.models[0].tensor_names():
        for slot, key in ((0, None), (1, "n"), (2, "z")):
            got = grp.get_tensor(name, slot=slot)
            exp = om.params[name] if key is None else om.slots[name][key]
            sc = max(float(np.abs(exp).max()), 1e-3)
            assert np.max(np.abs(got - exp)) <= 5e-4 * sc, "%s slot %d: %g (scale %g)" % (name, slot, np.max(np.abs(got - exp)), sc)


def test_sharded_bench_engine_criteo_shape():
    """The benchmark's multi-GPU configuration in miniature (Criteo shape, tables scaled 1e-3, the 8 large tables and 16 large
    wide columns row-sharded over 2 ranks, bf16x3 towers): losses of 4 steps against the oracle at the 1e-4 bar."""
    from wide_deep_b200 import synthetic
    G, per = 2, 1024
    B = G * per
    fc, cross, model, emb = synthetic.criteo_conf(scale=1e-3, hidden=(256, 128, 64))
    om = OM.OracleModel(fc, cross, model, "wide_deep", embedding_dim_override=emb).init(61)
    cats = [f for f, c in fc.items() if c["type"] == "category"]
    dn = [f for f, c in fc.items() if c["type"] == "continuous"]
    # (the first steps from the TF initialisers are a violent transient that amplifies ANY rounding difference ~1000x — see
    # tests/test_gpu_bench_engine.py; the oracle walks through it alone and the comparison starts in the settled regime)
    for s in range(10):
        keys, dense, label = synthetic.criteo_batch_arrays(fc, B, step=1000 + s)
        raw = {f: (np.arange(B + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
        for j, f in enumerate(dn):
            raw[f] = np.ascontiguousarray(dense[:, j])
        om.train_step(raw, label)
    grp = make_group(fc, cross, model, "wide_deep", G, per, om, dense_rows=100, emb_dim=emb, engine="bf16x3", max_ids=len(fc) + len(cross))
    for step in range(4):
        keys, dense, label = synthetic.criteo_batch_arrays(fc, B, step=step, zipf=1.2 if step == 1 else None)
        raw = {f: (np.arange(B + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
        for j, f in enumerate(dn):
            raw[f] = np.ascontiguousarray(dense[:, j])
        shards = [Batch(per, keys[r * per:(r + 1) * per].reshape(-1), None, dense[r * per:(r + 1) * per], label[r * per:(r + 1) * per]) for r in range(G)]
        loss = grp.train_step(shards)
        ref, _ = om.train_step(raw, label)
        assert abs(loss - ref) <= RTOL * max(abs(ref), 1.0), "step %d loss %g vs %g" % (step, loss, ref)


@pytest.mark.parametrize("same_gpu", [True, False])
def test_sharded_ranks_in_separate_processes(same_gpu):
    """The multi-process path: CUDA IPC mapping of the peers' exchange segments, flag barriers in peer memory, the step replayed
    from a CUDA graph.  same_gpu: two processes share cuda:0 (runs on the single-GPU box; the contexts time-slice, so barriers
    are slow but the protocol is the real one).  Otherwise one process per GPU (needs >= 2 GPUs)."""
    import torch
    n = torch.cuda.device_count()
    if not same_gpu and n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2 if same_gpu else min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29653", os.path.join(ROOT, "tests", "_shard_worker.py")]
    env = dict(os.environ)
    if same_gpu:
        env["WD_SHARD_SAME_GPU"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0 and "SHARD_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
