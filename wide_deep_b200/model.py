"""Python handle on a ``WdModel`` (libwd_b200) plus the host-side batch container.

``WideDeepModel`` is the thin object the estimator shim drives: it owns the native handle created from a
compiled ``Plan`` and exposes one call per C-ABI step (train_step / forward / eval), tensor IO by the
TensorFlow variable names of the reference's checkpoints, and the split backward/apply used for
data-parallel training.  All arithmetic happens in the CUDA library; numpy is only the container for host
buffers.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native
from ._native import BatchC, check
from .plan import Plan

METRIC_KEYS = ["accuracy", "accuracy_baseline", "auc", "auc_precision_recall", "average_loss", "label/mean", "loss",
               "precision", "prediction/mean", "recall"]


class Batch(object):
    """One batch in host memory: CSR of uint64 keys over (row, cat field) + dense matrix + labels.
    ``offsets`` may be None when every (row, field) holds exactly one key (Criteo-style)."""

    def __init__(self, batch_size, keys, offsets=None, dense=None, label=None, weight=None):
        self.batch_size = int(batch_size)
        self.keys = np.ascontiguousarray(keys, dtype=np.uint64)
        self.offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32)
        self.dense = None if dense is None else np.ascontiguousarray(dense, dtype=np.float32)
        self.label = None if label is None else np.ascontiguousarray(label, dtype=np.float32)
        self.weight = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)

    def h2d_bytes(self):
        n = self.keys.nbytes
        for a in (self.offsets, self.dense, self.label, self.weight):
            if a is not None:
                n += a.nbytes
        return n

    def to_c(self):
        c = BatchC()
        c.batch_size = self.batch_size
        c.cat_offsets = self.offsets.ctypes.data if self.offsets is not None else None
        c.cat_keys = self.keys.ctypes.data
        c.nnz = int(self.keys.shape[0])
        c.dense = self.dense.ctypes.data if self.dense is not None else None
        c.label = self.label.ctypes.data if self.label is not None else None
        c.weight = self.weight.ctypes.data if self.weight is not None else None
        return c

    def rows(self, lo, hi, n_fields):
        """Row slice [lo, hi) as a new Batch (used to shard a global batch over ranks)."""
        if self.offsets is None:
            keys, offs = self.keys[lo * n_fields:hi * n_fields], None
        else:
            s, e = int(self.offsets[lo * n_fields]), int(self.offsets[hi * n_fields])
            keys = self.keys[s:e]
            offs = self.offsets[lo * n_fields:hi * n_fields + 1] - s
        nd = 0 if self.dense is None else self.dense.size // max(self.batch_size, 1)
        return Batch(hi - lo, keys, offs,
                     None if self.dense is None else self.dense.reshape(self.batch_size, nd)[lo:hi],
                     None if self.label is None else self.label[lo:hi],
                     None if self.weight is None else self.weight[lo:hi])


class WideDeepModel(object):
    def __init__(self, plan: Plan, device=0):
        self.plan = plan
        self._lib = _native.lib()
        desc, self._keep = plan.to_c()
        h = ctypes.c_void_p()
        check(self._lib.wd_model_create(ctypes.byref(desc), int(device), ctypes.byref(h)))
        self._h = h
        self.device = int(device)
        self.global_step = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.wd_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ parameters
    def init(self, seed=0):
        check(self._lib.wd_model_init(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF))
        return self

    def set_opt_step(self, steps):
        """Restore the optimizers' step count (Adam's beta powers) after loading a checkpoint."""
        check(self._lib.wd_set_opt_step(self._h, int(steps)))

    def tensor_names(self):
        return list(self.plan.tensor_names.keys())

    def n_slots(self, name):
        o = self.plan.lin_opt if name.startswith("linear/") else self.plan.dnn_opt
        return {"sgd": 0, "adagrad": 1, "ftrl": 2, "adam": 2, "rmsprop": 2}[o["kind"]]

    def get_tensor(self, name, slot=0):
        """Parameter (slot 0) or optimizer slot by TensorFlow variable name.  Row-sharded tensors: this rank's rows
        (global rows rank, rank + G, ...; Plan.local_shape)."""
        kind, index, sub, _ = self.plan.tensor_names[name]
        shape = self.plan.local_shape(name)
        out = np.empty(shape, dtype=np.float32)
        check(self._lib.wd_tensor_io(self._h, kind, index, sub, slot, out.ctypes.data, out.size, 0))
        return out

    def set_tensor(self, name, value, slot=0):
        """``value`` has the GLOBAL shape; of a row-sharded tensor only this rank's rows are uploaded."""
        kind, index, sub, shape = self.plan.tensor_names[name]
        v = np.ascontiguousarray(value, dtype=np.float32).reshape(shape)
        if self.plan.is_sharded_tensor(name):
            v = np.ascontiguousarray(v[self.plan.shard_rank::self.plan.shard_world])
        check(self._lib.wd_tensor_io(self._h, kind, index, sub, slot, v.ctypes.data, v.size, 1))

    # ------------------------------------------------------------------ steps
    def train_step(self, batch: Batch):
        """One optimizer step on a host batch; returns the sum-reduced loss (reference joint.py:404-406)."""
        c = batch.to_c()
        self._rows_hint = batch.batch_size
        loss = ctypes.c_float()
        check(self._lib.wd_train_step(self._h, ctypes.byref(c), ctypes.byref(loss)))
        self.global_step += 1
        return loss.value

    def upload(self, batch: Batch):
        c = batch.to_c()
        check(self._lib.wd_batch_upload(self._h, ctypes.byref(c)))
        self._rows_hint = batch.batch_size

    def train_step_resident(self, want_loss=True):
        loss = ctypes.c_float()
        check(self._lib.wd_train_step_resident(self._h, ctypes.byref(loss) if want_loss else None))
        self.global_step += 1
        return loss.value

    def forward(self, batch: Batch):
        """-> (logits float32[B], loss or None)"""
        c = batch.to_c()
        self._rows_hint = batch.batch_size
        logits = np.empty(batch.batch_size, dtype=np.float32)
        loss = ctypes.c_float()
        check(self._lib.wd_forward(self._h, ctypes.byref(c), logits.ctypes.data, ctypes.byref(loss)))
        return logits, (loss.value if batch.label is not None else None)

    def step_backward(self, batch: Batch | None, want_loss=True):
        loss = ctypes.c_float()
        c = batch.to_c() if batch is not None else None
        if batch is not None:
            self._rows_hint = batch.batch_size
        check(self._lib.wd_step_backward(self._h, ctypes.byref(c) if c is not None else None,
                                         ctypes.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def step_backward_slot(self, slot, want_loss=True):
        loss = ctypes.c_float()
        check(self._lib.wd_step_backward_slot(self._h, int(slot), ctypes.byref(loss) if want_loss else None))
        return loss.value if want_loss else None

    def step_apply(self):
        check(self._lib.wd_step_apply(self._h))
        self.global_step += 1

    def dense_grad(self):
        """(device pointer, float count) of the dense gradient arena after step_backward."""
        return int(self._lib.wd_dense_grad_ptr(self._h) or 0), int(self._lib.wd_dense_grad_count(self._h))

    def sparse_grads(self, which, want_count=True):
        """(rows ptr, grads ptr, n or None, width, capacity).  want_count=False does not synchronise."""
        rows, grads = ctypes.c_void_p(), ctypes.c_void_p()
        n, cap, width = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        check(self._lib.wd_sparse_grads(self._h, which, ctypes.byref(rows), ctypes.byref(grads),
                                        ctypes.byref(n) if want_count else None, ctypes.byref(width), ctypes.byref(cap)))
        return rows.value, grads.value, (n.value if want_count else None), width.value, cap.value

    def sparse_set(self, which, rows_ptr, grads_ptr, n):
        check(self._lib.wd_sparse_set(self._h, which, ctypes.c_void_p(rows_ptr), ctypes.c_void_p(grads_ptr), int(n)))

    def sparse_set_sorted(self, which, rows_ptr, grads_ptr, n_lists, list_len):
        """Merge n_lists sorted, duplicate-free, INVALID_ROW-padded lists of list_len rows (all-gathered wd_sparse_grads buffers)."""
        check(self._lib.wd_sparse_set_sorted(self._h, which, ctypes.c_void_p(rows_ptr), ctypes.c_void_p(grads_ptr), int(n_lists), int(list_len)))

    # ------------------------------------------------------------------ eval
    def eval_reset(self):
        check(self._lib.wd_eval_reset(self._h))

    def eval_accumulate(self, batch: Batch):
        c = batch.to_c()
        check(self._lib.wd_eval_accumulate(self._h, ctypes.byref(c)))

    def eval_finish(self):
        out = np.zeros(10, dtype=np.float64)
        check(self._lib.wd_eval_finish(self._h, out.ctypes.data))
        return dict(zip(METRIC_KEYS, (float(v) for v in out)))

    # ------------------------------------------------------------------ introspection
    def column_ids(self):
        """CSR (offsets int32[B*C+1], ids int64[nnz]) of the last batch, for parity tests."""
        nnz = ctypes.c_int64()
        check(self._lib.wd_debug_column_ids(self._h, None, 0, None, 0, ctypes.byref(nnz)))
        B = self._rows_hint
        offs = np.empty(B * len(self.plan.columns) + 1, dtype=np.int32)
        ids = np.empty(max(nnz.value, 1), dtype=np.int64)
        check(self._lib.wd_debug_column_ids(self._h, offs.ctypes.data, offs.size, ids.ctypes.data, ids.size, ctypes.byref(nnz)))
        return offs, ids[:nnz.value]

    def deep_input(self, batch_size):
        out = np.empty((batch_size, self.plan.d0_phys), dtype=np.float32)
        check(self._lib.wd_debug_deep_input(self._h, out.ctypes.data, out.size))
        return out

    def hidden_output(self, tower, layer, batch_size):
        out = np.empty(batch_size * 4096, dtype=np.float32)
        n = self._lib.wd_debug_hidden(self._h, tower, layer, out.ctypes.data, out.size)
        if n < 0:
            check(n)
        return out[:batch_size * n].reshape(batch_size, n)

    def launch_count(self):
        return int(self._lib.wd_launch_count(self._h))

    def gemm_fallback_count(self):
        """Tensor-core-engine GEMMs that ran on the FFMA kernel instead (must stay 0)."""
        return int(self._lib.wd_gemm_fallback_count(self._h))

    def set_profile(self, on=True):
        check(self._lib.wd_set_profile(self._h, 1 if on else 0))

    def last_timings(self):
        """OrderedDict phase -> ms of the last synchronised step ('total' first); needs set_profile(True)."""
        from collections import OrderedDict
        out = np.zeros(64, dtype=np.float32)
        n = self._lib.wd_last_timings(self._h, out.ctypes.data, 64)
        res = OrderedDict()
        for i in range(max(n, 0)):
            name = self._lib.wd_timing_name(self._h, i).decode()
            res[name] = res.get(name, 0.0) + float(out[i])
        return res

    def upload_slot(self, slot, batch: Batch):
        c = batch.to_c()
        check(self._lib.wd_batch_upload_slot(self._h, int(slot), ctypes.byref(c)))
        self._rows_hint = batch.batch_size

    def prefetch_slot(self, slot, batch: Batch):
        """Asynchronous refill of a batch slot on the upload stream (overlaps the step running on another slot).  The batch's
        host arrays (pinned for a truly asynchronous copy) are kept alive here until the slot is refilled again."""
        c = batch.to_c()
        check(self._lib.wd_batch_prefetch_slot(self._h, int(slot), ctypes.byref(c)))
        if not hasattr(self, "_prefetched"):
            self._prefetched = {}
        self._prefetched[int(slot)] = (batch, c)
        self._rows_hint = batch.batch_size

    def last_loss(self):
        loss = ctypes.c_float()
        check(self._lib.wd_last_loss(self._h, ctypes.byref(loss)))
        return loss.value

    def train_step_slot(self, slot, want_loss=True):
        loss = ctypes.c_float()
        check(self._lib.wd_train_step_slot(self._h, int(slot), ctypes.byref(loss) if want_loss else None))
        self.global_step += 1
        return loss.value if want_loss else None

    def stream(self):
        return int(self._lib.wd_stream(self._h) or 0)

    def stream_sparse(self, which):
        return int(self._lib.wd_stream_sparse(self._h, int(which)) or 0)

    def sync(self):
        check(self._lib.wd_sync(self._h))
