"""ORACLE (test infrastructure): feature columns.

Restates what ``_build_model_columns`` wires up (reference python/lib/build_estimator.py:49-169) and what
the TensorFlow feature-column calls it makes compute.  Input is the plain
feature-conf dict and cross list; nothing here depends on ``wide_deep_b200``.

A *raw batch* is a dict ``feature -> value``:
  string features (hash_bucket / vocab)  -> CSR ``(offsets int64[B+1], fingerprints uint64[nnz])``
                                            (tokens already through Fingerprint64; see ``encode_tokens``)
  identity features                      -> int64[B]
  continuous features                    -> float32[B]
``tf_compat_pad`` reproduces quirk Q2: string keys of a cross arrive as DENSE padded tensors,
so each row contributes ``max_len`` entries, the missing ones being Fingerprint64("").
"""
import math

import numpy as np

from . import hashing as H


def embedding_dim(n):
    # reference build_estimator.py:57-59 — natural log (quirk Q12)
    return int(np.power(2, np.ceil(np.log(n ** 0.25))))


def encode_tokens(rows):
    """list (per example) of list of str tokens -> CSR of fingerprints (empty tokens already removed)."""
    offs = np.zeros(len(rows) + 1, dtype=np.int64)
    flat = []
    for i, r in enumerate(rows):
        flat.extend(r)
        offs[i + 1] = len(flat)
    return offs, H.fingerprint64_tokens(flat)


def _csr_from_single(vals, keep):
    """dense per-row values + keep mask -> CSR"""
    offs = np.zeros(len(vals) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(keep)
    return offs, np.asarray(vals)[keep]


class Column(object):
    name = None


class HashBucket(Column):
    def __init__(self, feature, size):
        self.feature, self.size, self.name = feature, int(size), feature
        self.num_buckets = int(size)

    def ids(self, batch, **_):
        offs, fp = batch[self.feature]
        keep = fp != np.uint64(H.FP_EMPTY)          # A.1: '' entries are "missing" (2^-64 caveat)
        ids = (fp % np.uint64(self.size)).astype(np.int64)   # A.2: unsigned modulo
        return _filter_csr(offs, ids, keep)


class Vocab(Column):
    def __init__(self, feature, vocab):
        self.feature, self.name = feature, feature
        self.vocab = [str(v) for v in vocab]     # reference build_estimator.py:103 map(str, ...)
        self.num_buckets = len(self.vocab)
        self._fp = {H.fingerprint64(t): i for i, t in enumerate(self.vocab)}

    def ids(self, batch, **_):
        offs, fp = batch[self.feature]
        ids = np.fromiter((self._fp.get(int(v), -1) for v in fp), dtype=np.int64, count=len(fp))
        return _filter_csr(offs, ids, ids >= 0)   # OOV -> -1 -> pruned by every consumer (A.3)


class Identity(Column):
    def __init__(self, feature, n):
        self.feature, self.name, self.num_buckets = feature, feature, int(n)

    def ids(self, batch, **_):
        v = np.asarray(batch[self.feature], dtype=np.int64)
        keep = v != -1                                            # A.1 dense->sparse drops -1
        ids = np.where((v < 0) | (v >= self.num_buckets), 0, v)   # A.4 default_value=0
        return _csr_from_single(ids, keep)


class Numeric(Column):
    def __init__(self, feature, transform=None, normalization=None):
        self.feature, self.name = feature, feature
        self.transform, self.norm = transform, normalization

    def values(self, batch):
        x = np.asarray(batch[self.feature], dtype=np.float32)
        if self.transform is None:
            return x
        if self.transform == "min_max":     # build_estimator.py:63-64, fp32 tensor arithmetic
            a, b = self.norm
            return ((x - np.float32(a)) / np.float32(b - a)).astype(np.float32)
        if self.transform == "standard":
            m, s = self.norm
            return ((x - np.float32(m)) / np.float32(s)).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.log(x).astype(np.float32)


class Bucketized(Column):
    def __init__(self, source, boundaries):
        self.source = source
        self.boundaries = np.asarray(boundaries, dtype=np.float32)
        self.name = source.name + "_bucketized"
        self.num_buckets = len(boundaries) + 1

    def ids(self, batch, **_):
        ids = H.bucketize(self.source.values(batch), self.boundaries)   # A.6
        return _csr_from_single(ids, np.ones(len(ids), dtype=bool))


class Crossed(Column):
    def __init__(self, keys, size):
        # keys: list of str (raw string feature) | Identity | Bucketized, in conf order
        self.keys, self.num_buckets = keys, int(size)
        leaf = [k if isinstance(k, str) else k.name for k in keys]
        self.name = "_X_".join(sorted(leaf))

    def op_order(self):
        """A.5: SparseTensor inputs (categorical-column keys) first, then dense (raw string) inputs."""
        return [k for k in self.keys if not isinstance(k, str)] + [k for k in self.keys if isinstance(k, str)]

    def ids(self, batch, tf_compat_pad=False, **_):
        cols = []
        for k in self.op_order():
            if isinstance(k, str):
                offs, fp = batch[k]
                if tf_compat_pad:
                    offs, fp = _pad_csr(offs, fp)
                cols.append((offs, fp))
            else:
                offs, ids = k.ids(batch)
                cols.append((offs, ids.astype(np.uint64)))    # ints enter the chain raw
        return H.cross_rows(cols, self.num_buckets)


class Embedding(Column):
    def __init__(self, cat, dim):
        self.cat, self.dim, self.name = cat, int(dim), cat.name + "_embedding"
        self.width = self.dim


class Indicator(Column):
    def __init__(self, cat):
        self.cat, self.name, self.width = cat, cat.name + "_indicator", cat.num_buckets


def _filter_csr(offs, vals, keep):
    csum = np.concatenate([[0], np.cumsum(keep)]).astype(np.int64)
    return csum[offs], vals[keep]


def _pad_csr(offs, fp):
    B = len(offs) - 1
    lens = np.diff(offs)
    L = int(lens.max()) if B else 0
    out = np.full((B, L), np.uint64(H.FP_EMPTY), dtype=np.uint64)
    for b in range(B):
        out[b, :lens[b]] = fp[offs[b]:offs[b + 1]]
    return np.arange(B + 1, dtype=np.int64) * L, out.reshape(-1)


def build_columns(feature_conf, cross_conf, embedding_dim_override=None):
    """-> (wide_columns, deep_columns), restating reference build_estimator.py:70-158.
    cross_conf: list of (feature names, bucket count, is_deep) as produced by read_cross_feature_conf."""
    edim = (lambda n: embedding_dim_override) if embedding_dim_override else embedding_dim
    wide, deep = [], []
    for f, conf in feature_conf.items():
        t, tr, p = conf["type"], conf["transform"], conf["parameter"]
        if t == "category":
            if tr == "hash_bucket":
                c = HashBucket(f, p)
                wide.append(c)
                deep.append(Embedding(c, edim(p)))
            elif tr == "vocab":
                c = Vocab(f, p)
                wide.append(c)
                deep.append(Indicator(c))
            elif tr == "identity":
                c = Identity(f, p)
                wide.append(c)
                deep.append(Indicator(c))
        else:
            c = Numeric(f, tr, tuple(p["normalization"]) if tr else None)
            if p["boundaries"]:
                wide.append(Bucketized(c, p["boundaries"]))      # buckets the NORMALISED value (quirk Q3)
            deep.append(c)
    for names, size, is_deep in cross_conf:
        keys = []
        for f in names:
            conf = feature_conf[f]
            if conf["type"] == "continuous":
                keys.append(Bucketized(Numeric(f), conf["parameter"]["boundaries"]))   # raw value (145)
            elif conf["transform"] == "identity":
                keys.append(Identity(f, conf["parameter"]))
            else:
                keys.append(f)
        c = Crossed(keys, size)
        wide.append(c)
        if is_deep:
            deep.append(Embedding(c, edim(size)))
    return wide, deep
