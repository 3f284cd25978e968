"""Shared test helpers: build the SAME inputs for the oracle (raw batch dict) and the product (Batch)."""
import numpy as np

from oracle import hashing as OH
from wide_deep_b200.model import Batch


def random_raw_batch(feature_conf, B, rng, multihot_max=3, na_rate=0.1, vocab_oov_rate=0.2):
    """Random raw batch in the oracle's format (see oracle/columns.py): string features as CSR of
    fingerprints of random tokens, identity ints (sometimes out of range / -1), floats."""
    raw = {}
    for f, c in feature_conf.items():
        if c["type"] == "category" and c["transform"] != "identity":
            rows = []
            for _ in range(B):
                if rng.random() < na_rate:
                    rows.append([])
                    continue
                n = int(rng.integers(1, multihot_max + 1))
                if c["transform"] == "vocab":
                    toks = [str(c["parameter"][int(rng.integers(len(c["parameter"])))]) if rng.random() > vocab_oov_rate
                            else "oov%d" % rng.integers(100) for _ in range(n)]
                else:
                    toks = ["tok%d" % rng.integers(0, 50) for _ in range(n)]
                rows.append(toks)
            offs = np.zeros(B + 1, dtype=np.int64)
            flat = []
            for i, r in enumerate(rows):
                flat.extend(r)
                offs[i + 1] = len(flat)
            raw[f] = (offs, OH.fingerprint64_tokens(flat))
        elif c["type"] == "category":
            v = rng.integers(-2, c["parameter"] + 3, size=B)
            raw[f] = v.astype(np.int64)
        else:
            raw[f] = (rng.standard_normal(B) * 30 + 40).astype(np.float32)
    return raw


def to_product_batch(plan, raw, label=None, weight=None, tf_compat_pad=False):
    """Oracle raw batch -> product Batch (CSR over (row, cat field) in plan.cat_fields order)."""
    B = len(label) if label is not None else len(next(v for v in raw.values() if not isinstance(v, tuple)))
    F = len(plan.cat_fields)
    per_field = []
    for f, is_str in zip(plan.cat_fields, plan.cat_is_string):
        if is_str:
            offs, fp = raw[f]
            if tf_compat_pad:
                lens = np.diff(offs)
                L = int(lens.max()) if len(lens) else 0
                pad = np.full((len(lens), L), np.uint64(OH.FP_EMPTY), dtype=np.uint64)
                for b in range(len(lens)):
                    pad[b, :lens[b]] = fp[offs[b]:offs[b + 1]]
                offs, fp = np.arange(len(lens) + 1, dtype=np.int64) * L, pad.reshape(-1)
            per_field.append((offs, fp))
        else:
            v = np.asarray(raw[f], dtype=np.int64)
            per_field.append((np.arange(len(v) + 1, dtype=np.int64), v.astype(np.uint64)))
    B = len(per_field[0][0]) - 1 if per_field else B
    offsets = np.zeros(B * F + 1, dtype=np.int32)
    chunks = []
    n = 0
    for b in range(B):
        for j, (offs, vals) in enumerate(per_field):
            offsets[b * F + j] = n
            seg = vals[offs[b]:offs[b + 1]]
            chunks.append(seg)
            n += len(seg)
    offsets[B * F] = n
    keys = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint64)
    dense = None
    if plan.dense_fields:
        dense = np.stack([np.asarray(raw[f], dtype=np.float32) for f in plan.dense_fields], axis=1)
    return Batch(B, keys, offsets, dense, label, weight)


def copy_params_to_product(oracle_model, model):
    """Upload the oracle's parameters and optimizer slots under the shared TensorFlow variable names."""
    for name in model.tensor_names():
        model.set_tensor(name, oracle_model.params[name])
        slots = oracle_model.slots[name]
        if "acc" in slots:
            model.set_tensor(name, slots["acc"], slot=1)
        if "n" in slots:
            model.set_tensor(name, slots["n"], slot=1)
            model.set_tensor(name, slots["z"], slot=2)
        if "m" in slots:
            model.set_tensor(name, slots["m"], slot=1)
            model.set_tensor(name, slots["v"], slot=2)
        if "ms" in slots:
            model.set_tensor(name, slots["ms"], slot=1)
            model.set_tensor(name, slots["mom"], slot=2)


def rel_err(a, b, floor=1.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
