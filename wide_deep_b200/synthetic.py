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
