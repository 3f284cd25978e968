"""Synthetic Criteo-shaped configuration and batches (BASELINE.json configs[1]/[2]).

13 continuous + 26 single-valued categorical features (hash_bucket, Criteo-Kaggle cardinalities), embedding
width forced to 32, wide part = the 26 hash columns + 13 bucketized dense columns (10 boundaries) + 8 pairwise
crosses at 1M buckets, MLP 1024-512-256 simple/relu/BN, Adagrad(0.05) deep + FTRL(0.1, l1 0.5, l2 1) wide.
Expressed as ordinary feature/cross/model conf dicts, i.e. exactly what conf/*.yaml would hold.
Batches are generated with numpy (host) from a counter-based seed so every rank/step is reproducible.
"""
from collections import OrderedDict

import numpy as np

CRITEO_CARDINALITIES = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
                        5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]
CROSS_PAIRS = [(1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12), (13, 14), (15, 16)]


def criteo_conf(cardinalities=None, n_dense=13, emb_dim=32, hidden=(1024, 512, 256), mode="simple", n_cross=8,
                cross_buckets_k=1000, scale=1.0):
    """-> (feature_conf, cross_conf list, model_conf, embedding_dim_override).  ``scale`` shrinks every
    cardinality (tests use small tables)."""
    card = list(cardinalities or CRITEO_CARDINALITIES)
    fc = OrderedDict()
    for i in range(n_dense):
        fc["i%d" % (i + 1)] = dict(type="continuous", transform="standard",
                                   parameter=dict(normalization=[0.0, 1.0],
                                                  boundaries=[-1.5, -1.0, -0.6, -0.3, 0.0, 0.3, 0.6, 1.0, 1.5, 2.0]))
    for i, n in enumerate(card):
        fc["c%d" % (i + 1)] = dict(type="category", transform="hash_bucket", parameter=max(2, int(n * scale)))
    cross = []
    for a, b in CROSS_PAIRS[:n_cross]:
        if a <= len(card) and b <= len(card):
            cross.append((["c%d" % a, "c%d" % b], max(100, int(cross_buckets_k * 1000 * scale)), 0))
    model = dict(linear_optimizer="tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)",
                 linear_initial_learning_rate=0.05, linear_decay_rate=0.8,
                 dnn_hidden_units=list(hidden), dnn_connected_mode=mode, dnn_optimizer="Adagrad",
                 dnn_initial_learning_rate=0.05, dnn_decay_rate=0.8, dnn_activation_function="relu",
                 dnn_l1=0.1, dnn_l2=0.1, dnn_dropout=None, dnn_batch_normalization=1, cnn_use_flag=0)
    return fc, cross, model, emb_dim


def criteo_batch_arrays(feature_conf, batch_size, seed=0x5EED0001, step=0, zipf=None, pos_rate=0.03):
    """Host arrays of one batch: keys uint64 [B, n_cat] (one key per field, row-major), dense float32 [B, n_dense],
    label float32 [B].  Keys are uniform 64-bit fingerprints (=> uniform ids after the modulo: the HBM
    worst case the roofline is quoted on) or, with ``zipf`` = alpha, ids drawn Zipf(alpha) and lifted to a
    fingerprint with the same residue."""
    cats = [(f, c["parameter"]) for f, c in feature_conf.items() if c["type"] == "category"]
    n_dense = sum(1 for c in feature_conf.values() if c["type"] == "continuous")
    rng = np.random.Generator(np.random.Philox(key=[seed, step]))
    B = batch_size
    if zipf is None:
        keys = rng.integers(0, np.iinfo(np.uint64).max, size=(B, len(cats)), dtype=np.uint64, endpoint=True)
    else:
        keys = np.empty((B, len(cats)), dtype=np.uint64)
        for j, (_, n) in enumerate(cats):
            ids = (rng.zipf(zipf, size=B) - 1) % n
            mult = rng.integers(0, (2 ** 63) // n, size=B, dtype=np.uint64)
            keys[:, j] = ids.astype(np.uint64) + mult * np.uint64(n)
    dense = rng.standard_normal((B, n_dense), dtype=np.float32)
    label = (rng.random(B) < pos_rate).astype(np.float32)
    return keys, dense, label


# ---------------------------------------------------------------------------------- BASELINE.json configs[3]: multihot slot
def multihot_conf(rows=12_500_000, emb_dim=64, hidden=(512, 512, 512, 512)):
    """One hashed multihot slot (avg 30 ids per example) with a 64-wide embedding and a ResDnn 4 x 512 ('resnet' connections)
    on top — the pure embedding-bag workload.  model_type 'deep'.  100 M rows over 8 GPUs = 12.5 M rows
    per GPU; single-GPU runs use the 12.5 M-row slice."""
    fc = OrderedDict()
    fc["tags"] = dict(type="category", transform="hash_bucket", parameter=int(rows))
    model = dict(linear_optimizer="Ftrl", linear_initial_learning_rate=0.05, dnn_hidden_units=list(hidden),
                 dnn_connected_mode="resnet", dnn_optimizer="Adagrad", dnn_initial_learning_rate=0.05,
                 dnn_activation_function="relu", dnn_dropout=None, dnn_batch_normalization=1)
    return fc, [], model, emb_dim


def multihot_batch_arrays(batch_size, seed=0x5EED0002, step=0, mean_ids=30, max_ids=128, pos_rate=0.03):
    """-> (keys uint64[nnz], offsets int32[B + 1], label float32[B]): ids per example ~ Poisson(mean) clipped to [1, max],
    keys uniform 64-bit fingerprints (uniform rows: the HBM worst case)."""
    rng = np.random.Generator(np.random.Philox(key=[seed, step]))
    lens = np.clip(rng.poisson(mean_ids, size=batch_size), 1, max_ids).astype(np.int64)
    offs = np.zeros(batch_size + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    keys = rng.integers(0, np.iinfo(np.uint64).max, size=int(offs[-1]), dtype=np.uint64, endpoint=True)
    label = (rng.random(batch_size) < pos_rate).astype(np.float32)
    return keys, offs.astype(np.int32), label


# ------------------------------------------------------------------------------ BASELINE.json configs[4]: wide-only crosses
def wide_conf(total_cross_rows=125_000_000, n_fields=9, field_rows=100_000, n_cross=32):
    """Wide-only model (model_type 'wide'): n_fields hashed key fields and n_cross pairwise hashed crosses into one large weight
    space (n_cross x total/n_cross buckets), FTRL(0.1, l1 0.5, l2 1) — the pure sparse-linear path.
    1 B cross buckets over 8 GPUs = 125 M per GPU; single-GPU runs use the 125 M slice."""
    fc = OrderedDict()
    for i in range(n_fields):
        fc["k%d" % i] = dict(type="category", transform="hash_bucket", parameter=int(field_rows + 1009 * i))
    pairs = [(a, b) for a in range(n_fields) for b in range(a + 1, n_fields)][:n_cross]
    per = max(100, int(total_cross_rows // max(len(pairs), 1)))
    cross = [(["k%d" % a, "k%d" % b], per, 0) for a, b in pairs]
    model = dict(linear_optimizer="tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)",
                 linear_initial_learning_rate=0.05, dnn_hidden_units=[8], dnn_connected_mode="simple", dnn_optimizer="Adagrad",
                 dnn_initial_learning_rate=0.05, dnn_activation_function="relu", dnn_dropout=None, dnn_batch_normalization=0)
    return fc, cross, model, None


def wide_batch_arrays(feature_conf, batch_size, seed=0x5EED0003, step=0, pos_rate=0.03):
    """-> (keys uint64[B, n_fields], label float32[B]); one uniform key per field."""
    n = sum(1 for c in feature_conf.values() if c["type"] == "category")
    rng = np.random.Generator(np.random.Philox(key=[seed, step]))
    keys = rng.integers(0, np.iinfo(np.uint64).max, size=(batch_size, n), dtype=np.uint64, endpoint=True)
    label = (rng.random(batch_size) < pos_rate).astype(np.float32)
    return keys, label
