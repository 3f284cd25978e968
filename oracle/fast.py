"""ORACLE (test infrastructure): the OPTIMISED CPU restatement of the train step — what bench.py times as the CPU arm.

``oracle/model.py`` is the checker: numpy + scipy, float64 accumulation, written for clarity.  Timing it says little
about what host cores can do, so the reference arm / ``cpu_baseline`` leg of bench.py times THIS module instead: the
same step (same ids, same wiring, same optimizers, "sum duplicates, apply once") on the host's best kernels —

  ids              the plain-C hashing restatement (oracle/wd_oracle_hash.c) through ``OracleModel.transform``
  gather + pool    ``torch.nn.functional.embedding_bag`` (mode mean / sum, multi-threaded, sparse gradients)
  towers           torch-CPU matmuls (oneDNN / MKL sgemm on every host thread), autograd for the backward
  sparse update    coalesced sparse gradients (unique rows, summed) -> Adagrad / FTRL / SGD on the touched rows only
  dense update     in-place torch elementwise kernels

It follows the reference files the checker follows (reference python/lib/linear.py:29-36, dnn.py:83-233, joint.py:216-262,
lib/utils/model_util.py:62-105) and is itself checked against ``oracle/model.py`` on CPU (tests/test_oracle_fast.py): same
losses, same parameters after several steps.  It shares the parameter / slot arrays of the ``OracleModel`` it wraps
(zero-copy ``torch.from_numpy`` views), so either implementation can continue from the other's state.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import columns as C
from .model import BN_EPS, layer_sources

_ACT = {
    "relu": torch.relu,
    "relu6": F.relu6,
    "sigmoid": torch.sigmoid,
    "tanh": torch.tanh,
    "leaky_relu": lambda z: F.leaky_relu(z, 0.2),
    "elu": F.elu,
    "selu": F.selu,
    "softplus": F.softplus,
    "softsign": F.softsign,
    "crelu": lambda z: torch.cat([torch.relu(z), torch.relu(-z)], 1),     # tf.nn.crelu
}


class FastCpuModel(object):
    """Optimised CPU step over the parameters of an ``OracleModel`` (fp32 arithmetic, all host threads)."""

    def __init__(self, oracle_model, threads=None):
        self.om = oracle_model
        for o in (oracle_model.opt_lin, oracle_model.opt_dnn):
            if o["kind"] not in ("adagrad", "ftrl", "sgd"):
                raise NotImplementedError("the timed CPU arm covers Adagrad / Ftrl / SGD (the benchmark's optimizers), not %s" % o["kind"])
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        om = oracle_model
        self.P = {k: torch.from_numpy(v) for k, v in om.params.items()}                       # views: updates land in om.params
        self.S = {k: {s: torch.from_numpy(a) for s, a in d.items()} for k, d in om.slots.items()}
        self.inv = float(1.0 / np.sqrt(1.0 + BN_EPS))

    # ------------------------------------------------------------------ forward (+ autograd graph)
    def _forward(self, batch, ids, train):
        om, P = self.om, self.P
        B = len(next(iter(ids.values()))[0]) - 1
        leaves = {}

        def leaf(name, as_column=False):
            if name not in leaves:
                t = P[name].view(-1, 1) if as_column else P[name]                           # wide weights: [n, 1] "embedding"
                leaves[name] = t.detach().requires_grad_(True) if train else t
            return leaves[name]

        logits = torch.zeros(B, dtype=torch.float32)
        if om.use_wide:
            wl = leaf("linear/linear_model/bias_weights").expand(B).clone()
            for c in om.wide_cols:
                offs, cid = ids[c.name]
                if len(cid) == 0:
                    continue
                w = leaf(om.wname(c), as_column=True)
                wl = wl + F.embedding_bag(torch.from_numpy(cid), w, torch.from_numpy(offs[:-1]), mode="sum", sparse=train).squeeze(1)
            logits = logits + wl
        if om.use_deep:
            parts = []
            for c in om.deep_cols:
                if isinstance(c, C.Numeric):
                    parts.append(torch.from_numpy(np.ascontiguousarray(c.values(batch), dtype=np.float32)).unsqueeze(1))
                elif isinstance(c, C.Indicator):
                    offs, cid = ids[c.cat.name]
                    x = torch.zeros(B, c.width, dtype=torch.float32)
                    if len(cid):
                        rows = torch.from_numpy(np.repeat(np.arange(B, dtype=np.int64), np.diff(offs)))
                        x.index_put_((rows, torch.from_numpy(cid)), torch.ones(len(cid)), accumulate=True)
                    parts.append(x)
                else:
                    offs, cid = ids[c.cat.name]
                    if len(cid) == 0:
                        parts.append(torch.zeros(B, c.dim, dtype=torch.float32))
                        continue
                    parts.append(F.embedding_bag(torch.from_numpy(cid), leaf(om.ename(c)), torch.from_numpy(offs[:-1]),
                                                 mode="mean", sparse=train))                  # empty bag -> zeros (A.7)
            X = torch.cat(parts, 1)
            act = _ACT[om.act]
            for t, hu in enumerate(om.towers):
                srcs = layer_sources(om.modes[t], len(hu))
                H = []
                pick = lambda s: X if s == "x" else H[s]
                for l in range(len(hu)):
                    scope = "dnn/dnn_%d/hiddenlayer_%d" % (t + 1, l)
                    inp = torch.cat([pick(s) for s in srcs[l]], 1) if len(srcs[l]) > 1 else pick(srcs[l][0])
                    a = act(torch.addmm(leaf(scope + "/bias"), inp, leaf(scope + "/kernel")))
                    if om.bn:                                                                # inference-mode affine (quirk Q4)
                        a = a * (leaf(scope + "/batch_normalization/gamma") * self.inv) + leaf(scope + "/batch_normalization/beta")
                    H.append(a)
                scope = "dnn/dnn_%d/logits" % (t + 1)
                inp = torch.cat([pick(s) for s in srcs[-1]], 1) if len(srcs[-1]) > 1 else pick(srcs[-1][0])
                logits = logits + torch.addmm(leaf(scope + "/bias"), inp, leaf(scope + "/kernel")).squeeze(1)
        return logits, leaves

    def forward(self, batch):
        with torch.no_grad():
            logits, _ = self._forward(batch, self.om.transform(batch), False)
        return logits.numpy()

    # ------------------------------------------------------------------ optimizers
    def _apply(self, name, g):
        om = self.om
        o = om.opt_lin if name.startswith("linear/") else om.opt_dnn
        p, s = self.P[name], self.S[name]
        lr = o["lr"]
        if g.is_sparse:                                        # touched rows only; duplicates already summed by coalesce()
            g = g.coalesce()
            rows, gr = g.indices()[0], g.values()
            if rows.numel() == 0:
                return
            gr = gr.reshape((rows.numel(),) + tuple(p.shape[1:]))
            get = lambda a: a.index_select(0, rows)
            put = lambda a, v: a.index_copy_(0, rows, v)
        else:
            gr = g.reshape(p.shape)
            get = lambda a: a
            put = lambda a, v: a.copy_(v)
        if o["kind"] == "adagrad":
            acc = get(s["acc"]) + gr * gr
            put(s["acc"], acc)
            put(p, get(p) - lr * gr / acc.sqrt())
        elif o["kind"] == "ftrl":
            n0, z0, w0 = get(s["n"]), get(s["z"]), get(p)
            n1 = n0 + gr * gr
            z1 = z0 + gr - (n1.sqrt() - n0.sqrt()) / lr * w0
            w1 = torch.where(z1.abs() > o["l1"], (torch.sign(z1) * o["l1"] - z1) / (n1.sqrt() / lr + 2.0 * o["l2"]), torch.zeros_like(z1))
            put(s["n"], n1)
            put(s["z"], z1)
            put(p, w1)
        else:
            put(p, get(p) - lr * gr)

    def train_step(self, batch, labels, weights=None):
        """-> (loss, logits): the same step as OracleModel.train_step, in fp32 on the host's fast kernels."""
        om = self.om
        ids = om.transform(batch)
        logits, leaves = self._forward(batch, ids, True)
        y = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.float32))
        w = None if weights is None else torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float32))
        loss = F.binary_cross_entropy_with_logits(logits, y, weight=w, reduction="sum")       # SUM reduction (Q11)
        names = list(leaves)
        grads = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
        with torch.no_grad():
            for n, g in zip(names, grads):
                if g is not None:
                    self._apply(n, g)
        om.global_step += 1
        return float(loss), logits.detach().numpy()
