"""Input pipeline: TSV files -> CSR batches in pinned host memory.

Host-side mirror of ``input_fn`` / ``_CsvDataset`` (reference python/lib/dataset.py:62-195, 293-310): same
arguments (``csv_data_file`` may be a file or a directory, ``mode`` in {'train','eval','pred'}, ``batch_size``),
same parsing rules (TAB-separated, no quoting, NA token ``-`` -> per-type default, multi-valued string fields
split on ','), but the parser is the multi-threaded C++ loader in libwd_b200 (``wd_tsv_parse``): strings leave
it as Fingerprint64 values, so what crosses PCIe is the compact CSR the CUDA stage consumes.

Differences kept explicit:
  * shuffle: the reference shuffles with tf.data (buffer = num_examples, seed 123, dataset.py:180); that RNG
    stream cannot be reproduced outside TensorFlow, so 'train' mode shuffles whole files' lines with
    numpy's Philox(seed 123) instead — same intent (one seeded pass), different permutation;
  * distributed sharding: every ``world``-th line starting at ``rank`` (dataset.shard, dataset.py:173-174).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _native
from ._native import TsvSpecC
from .model import Batch


def list_files(path):
    """File or directory -> sorted list of data files, hidden files skipped (reference lib/utils/util.py:36-45,
    dataset.py:31-33)."""
    if not os.path.exists(path):
        raise AssertionError("data file: {} not found. Please check input data path".format(path))
    if os.path.isdir(path):
        return sorted(os.path.join(path, f) for f in os.listdir(path) if not f.startswith("."))
    return [path]


class TsvReader(object):
    """Parses TSV text into ``Batch`` objects for a given Plan."""

    def __init__(self, config, plan, is_pred=False, n_threads=None):
        self.plan = plan
        schema = config.read_schema()
        fconf = config.read_feature_conf()
        train = config.train
        names = [schema[k] for k in sorted(schema)]
        if is_pred:
            names = [n for n in names if n != "clk"]
        cat_idx = {f: i for i, f in enumerate(plan.cat_fields)}
        dense_idx = {f: i for i, f in enumerate(plan.dense_fields)}
        role, target = [], []
        for n in names:
            if n == "clk":
                role.append(0); target.append(0)
            elif n in cat_idx:
                role.append(1 if plan.cat_is_string[cat_idx[n]] else 2); target.append(cat_idx[n])
            elif n in dense_idx:
                role.append(3); target.append(dense_idx[n])
            else:
                role.append(-1); target.append(0)          # unused schema column: parsed and discarded
        self._role = np.asarray(role, dtype=np.int32)
        self._target = np.asarray(target, dtype=np.int32)
        pos, neg = train.get("pos_sample_loss_weight"), train.get("neg_sample_loss_weight")
        self.use_weight = pos is not None and neg is not None      # dataset.py:70-72 (both must be set)
        spec = TsvSpecC()
        spec.n_columns = len(names)
        spec.col_role, spec.col_target = self._role.ctypes.data, self._target.ctypes.data
        spec.n_cat_fields, spec.n_dense_fields = len(plan.cat_fields), len(plan.dense_fields)
        spec.multivalue = 1 if train.get("multivalue") else 0
        spec.tf_compat_pad = 1 if plan.tf_compat_pad else 0
        spec.pos_weight, spec.neg_weight = float(pos or 1), float(neg or 1)
        spec.use_weight = 1 if self.use_weight else 0
        spec.has_label = 0 if is_pred else 1
        self._spec = spec
        self.is_pred = is_pred
        self.n_threads = n_threads or min(os.cpu_count() or 1, 16)
        self._lib = _native.lib()

    def parse(self, lines):
        """list of text lines (str or bytes, no trailing newline needed) -> Batch"""
        if lines and isinstance(lines[0], bytes):
            text = b"\n".join(lines)                                  # input_fn's path: no decode / encode round trip
        else:
            text = ("\n".join(l.rstrip("\n") for l in lines)).encode("utf-8")
        n = len(lines)
        F, Nd = len(self.plan.cat_fields), len(self.plan.dense_fields)
        offsets = np.zeros(n * F + 1, dtype=np.int32)
        dense = np.zeros((n, max(Nd, 1)), dtype=np.float32)
        label = np.zeros(n, dtype=np.float32)
        weight = np.ones(n, dtype=np.float32)
        if self.plan.tf_compat_pad:
            # padded string fields (quirk Q2) can exceed the token count: ask for the size first (keys_cap = 0), then fill
            nnz = self._lib.wd_tsv_parse(ctypes.byref(self._spec), text, len(text), n, offsets.ctypes.data, None, 0,
                                         dense.ctypes.data, label.ctypes.data, weight.ctypes.data, self.n_threads)
            if nnz < 0:
                raise ValueError(self._lib.wd_last_error().decode())
            cap = max(nnz, 1)
        else:
            cap = n * max(F, 1) + text.count(b",") + 1               # every field holds at most (commas + 1) tokens
        keys = np.empty(cap, dtype=np.uint64)
        nnz = self._lib.wd_tsv_parse(ctypes.byref(self._spec), text, len(text), n, offsets.ctypes.data, keys.ctypes.data,
                                     keys.size, dense.ctypes.data, label.ctypes.data, weight.ctypes.data, self.n_threads)
        if nnz < 0:
            raise ValueError(self._lib.wd_last_error().decode())
        return Batch(n, keys[:nnz], offsets, dense[:, :Nd] if Nd else None, None if self.is_pred else label,
                     weight if (self.use_weight and not self.is_pred) else None)


def input_fn(csv_data_file, img_data_file, mode, batch_size, config=None, plan=None, rank=0, world=1, seed=123):
    """Generator of ``Batch`` for one pass over the data (one epoch), mirroring the reference's
    ``input_fn(csv_data_file, img_data_file, mode, batch_size)`` (dataset.py:293-310).  ``img_data_file`` is
    accepted for signature compatibility and must be None (the CNN branch is out of scope)."""
    assert mode in ("train", "eval", "pred"), "mode must in `train`, `eval`, or `pred`, found {}".format(mode)
    if img_data_file:
        raise ValueError("image inputs are not supported by the B200 path (cnn_use_flag: 0)")
    reader = TsvReader(config, plan, is_pred=(mode == "pred"))
    lines = []
    for f in list_files(csv_data_file):
        with open(f, "rb") as fh:
            lines.extend(l for l in fh.read().split(b"\n") if l != b"")
    if world > 1:
        lines = lines[rank::world]
    if mode == "train":
        perm = np.random.Generator(np.random.Philox(seed)).permutation(len(lines))
        lines = [lines[i] for i in perm]
    for i in range(0, len(lines), batch_size):
        yield reader.parse(lines[i:i + batch_size])
