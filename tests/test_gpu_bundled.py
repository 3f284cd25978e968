"""BASELINE.json configs[0]: the bundled conf/*.yaml + data/ fixtures through the product (C++ loader + CUDA step)
against the oracle (its own TSV parser + numpy step), and the drop-in entry points end to end."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bundled_conf_train_and_eval_parity():
    from oracle import model as OM, tsv as otsv
    from oracle.metrics import EvalAccumulator
    from tests.helpers import copy_params_to_product
    from wide_deep_b200.config import Config
    from wide_deep_b200.dataset import TsvReader
    from wide_deep_b200.model import WideDeepModel
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    B = 64
    fc, cc = cfg.read_feature_conf(), cfg.read_cross_feature_conf()
    plan = compile_plan(cfg, "wide_deep", B, tf_compat_pad=True, max_nnz=B * 2048, max_keys=B * 512)
    om = OM.OracleModel(fc, cc, cfg.model, "wide_deep", tf_compat_pad=True).init(123)
    pm = WideDeepModel(plan)
    copy_params_to_product(om, pm)
    reader = TsvReader(cfg, plan)
    lines = open(os.path.join(ROOT, "data", "train", "train1")).read().split("\n")
    lines = [l for l in lines if l]
    for step in range(4):
        chunk = lines[step * B:(step + 1) * B]
        batch = reader.parse(chunk)
        raw, lab = otsv.parse_lines(chunk, cfg.read_schema(), fc)
        if step == 0:                       # ids of every wide column, bit-exact, incl. multihot crosses with '' padding (Q2)
            pm.forward(batch)
            offs, ids = pm.column_ids()
            ref = om.transform(raw)
            C = len(plan.columns)
            for ci, col in enumerate(plan.columns):
                if col.name in ref:
                    ro, ri = ref[col.name]
                    for b in range(B):
                        assert np.array_equal(ids[offs[b * C + ci]:offs[b * C + ci + 1]], ri[ro[b]:ro[b + 1]]), (col.name, b)
        loss = pm.train_step(batch)
        ref_loss, _ = om.train_step(raw, lab)
        assert abs(loss - ref_loss) <= 1e-4 * max(abs(ref_loss), 1.0), (step, loss, ref_loss)
    # evaluation metrics on the eval fixture
    elines = [l for l in open(os.path.join(ROOT, "data", "eval", "eval1")).read().split("\n") if l][:128]
    acc = EvalAccumulator()
    pm.eval_reset()
    for i in range(0, len(elines), B):
        chunk = elines[i:i + B]
        pm.eval_accumulate(reader.parse(chunk))
        raw, lab = otsv.parse_lines(chunk, cfg.read_schema(), fc)
        _, cache = om.forward(raw)
        acc.update(cache["logits"].astype(np.float32), lab)
    got, exp = pm.eval_finish(), acc.result()
    for k in exp:
        assert abs(got[k] - exp[k]) <= 2e-4 * max(abs(exp[k]), 1.0), (k, got[k], exp[k])


def test_entry_points_train_then_eval(tmp_path):
    """python/train.py (dynamic_train over data/train/{train1,train2}) then python/eval.py: same flags, sorted
    metric printout with the head's ten keys (reference train.py:147-148, eval.py:82-83)."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    mdir = str(tmp_path / "model")
    r = subprocess.run([sys.executable, "train.py", "--model_dir", mdir, "--train_epochs", "1", "--batch_size", "64"],
                       cwd=os.path.join(ROOT, "python"), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for key in ["accuracy:", "accuracy_baseline:", "auc:", "auc_precision_recall:", "average_loss:", "label/mean:", "loss:",
                "precision:", "prediction/mean:", "recall:"]:
        assert key in r.stdout, key
    assert "Using dynamic train mode." in r.stdout
    assert any(f.startswith("model.ckpt-") for f in os.listdir(os.path.join(mdir, "wide_deep")))
    r2 = subprocess.run([sys.executable, "eval.py", "--model_dir", mdir, "--batch_size", "64"],
                        cwd=os.path.join(ROOT, "python"), env=env, capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert "average_loss:" in r2.stdout and "global_step:" in r2.stdout
    r3 = subprocess.run([sys.executable, "pred.py", "--model_dir", mdir, "--data_dir", "../data/pred", "--batch_size", "64"],
                        cwd=os.path.join(ROOT, "python"), env=env, capture_output=True, text=True, timeout=600)
    assert r3.returncode == 0 and r3.stdout.count("Prediction is") == 5000, r3.stdout[-1000:] + r3.stderr[-1000:]


def test_checkpoint_roundtrip(tmp_path):
    from wide_deep_b200.config import Config
    from wide_deep_b200.dataset import input_fn
    from wide_deep_b200.estimator import build_custom_estimator
    cfg = Config()
    est = build_custom_estimator(str(tmp_path / "m"), "wide_deep", config=cfg, max_batch=64)
    data = os.path.join(ROOT, "data", "test", "test2")
    est.train(input_fn=lambda: input_fn(data, None, "train", 64, config=cfg, plan=est.plan))
    m1 = est.evaluate(input_fn=lambda: input_fn(data, None, "eval", 64, config=cfg, plan=est.plan))
    est2 = build_custom_estimator(str(tmp_path / "m"), "wide_deep", config=cfg, max_batch=64)       # restores latest
    m2 = est2.evaluate(input_fn=lambda: input_fn(data, None, "eval", 64, config=cfg, plan=est2.plan))
    assert m1 == m2 and m2["global_step"] == 1


def test_estimator_train_equals_direct_steps(tmp_path):
    """estimator.train (step i enqueued, batch i+1 parsed and prefetched into the other slot meanwhile, then the loss of step i
    read) must leave exactly the parameters that plain train_step calls over the same batches leave (same seed, same order)."""
    from wide_deep_b200.config import Config
    from wide_deep_b200.dataset import input_fn
    from wide_deep_b200.estimator import build_custom_estimator
    cfg = Config()
    data = os.path.join(ROOT, "data", "eval", "eval1")                      # 5000 rows -> 79 batches of 64
    est_a = build_custom_estimator(str(tmp_path / "a"), "wide_deep", config=cfg, max_batch=64)
    est_a.train(input_fn=lambda: input_fn(data, None, "train", 64, config=cfg, plan=est_a.plan))
    est_b = build_custom_estimator(str(tmp_path / "b"), "wide_deep", config=cfg, max_batch=64)
    mb = est_b._ensure_model()                                             # fresh model, same seed as est_a's
    losses = [mb.train_step(b) for b in input_fn(data, None, "train", 64, config=cfg, plan=est_b.plan)]
    ma = est_a._ensure_model()
    assert ma.global_step == mb.global_step == 79 and np.isfinite(losses).all()
    for name in ma.tensor_names()[:12]:
        np.testing.assert_array_equal(ma.get_tensor(name), mb.get_tensor(name))


def test_device_fingerprint64_bit_exact(native_lib):
    import random
    from oracle import hashing as H
    rnd = random.Random(11)
    strs = [bytes(rnd.randrange(256) for _ in range(n)) for n in list(range(0, 140)) + [300, 1000, 5000]]
    buf = np.frombuffer(b"".join(strs) + b"\0", dtype=np.uint8).copy()
    offs = np.zeros(len(strs) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(s) for s in strs])
    out = np.zeros(len(strs), dtype=np.uint64)
    rc = native_lib.wd_fingerprint64_device(buf.ctypes.data, offs.ctypes.data, len(strs), out.ctypes.data)
    assert rc == 0
    assert [int(v) for v in out] == [H.fingerprint64(s) for s in strs]
    # ... and against the Abseil-derived known answers (0..32 bytes) directly, not only through the oracle
    import json
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "cityhash_le32_kat.json")))["vectors"]
    strs = [bytes.fromhex(v["hex"]) for v in kat]
    buf = np.frombuffer(b"".join(strs) + b"\0", dtype=np.uint8).copy()
    offs = np.zeros(len(strs) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([len(s) for s in strs])
    out = np.zeros(len(strs), dtype=np.uint64)
    assert native_lib.wd_fingerprint64_device(buf.ctypes.data, offs.ctypes.data, len(strs), out.ctypes.data) == 0
    assert [int(v) for v in out] == [int(v["hash"]) for v in kat]


def test_checkpoint_cadence_by_steps(tmp_path):
    """RunConfig's save_checkpoints_steps (reference conf/train.yaml:80-98): a checkpoint every N global steps during train() plus
    the final one; keep_checkpoint_max bounds what stays on disk; giving both cadences is the same error TensorFlow raises."""
    from wide_deep_b200.config import Config
    from wide_deep_b200.dataset import input_fn
    from wide_deep_b200.estimator import build_custom_estimator
    cfg = Config()
    run = cfg.runconfig                                  # (cached dict of this Config instance)
    run["save_checkpoints_steps"], run["save_checkpoints_secs"], run["keep_checkpoint_max"] = 20, None, 3
    data = os.path.join(ROOT, "data", "eval", "eval1")   # 5000 rows -> 79 batches of 64
    est = build_custom_estimator(str(tmp_path / "m"), "wide_deep", config=cfg, max_batch=64)
    est.train(input_fn=lambda: input_fn(data, None, "train", 64, config=cfg, plan=est.plan))
    have = sorted(int(f[len("model.ckpt-"):-4]) for f in os.listdir(str(tmp_path / "m")) if f.endswith(".npz"))
    assert have == [40, 60, 79], have                    # 20 was rotated out by keep_checkpoint_max = 3
    est2 = build_custom_estimator(str(tmp_path / "m"), "wide_deep", config=cfg, max_batch=64)
    assert est2._ensure_model().global_step == 79
    run["save_checkpoints_secs"] = 5
    est3 = build_custom_estimator(str(tmp_path / "m3"), "wide_deep", config=cfg, max_batch=64)
    with pytest.raises(ValueError):
        est3.train(input_fn=lambda: input_fn(data, None, "train", 64, config=cfg, plan=est3.plan))
