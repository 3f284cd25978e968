"""ORACLE (test infrastructure): the Wide&Deep train / eval step on CPU (numpy + scipy.sparse).

Restates, for ONE process, what one ``sess.run(train_op)`` of the reference computes:
  wide logit      reference python/lib/linear.py:29-36       
  deep input      reference python/lib/dnn.py:83-91          (A.7, a11)
  MLP             reference python/lib/dnn.py:92-233         (A.8, a12; 5 connection modes, BN = inference affine)
  multi tower     reference python/lib/dnn.py:237-275        (a14)
  logits + head   reference python/lib/joint.py:216-222, 264-269, 402-406   (A.10, a16; loss = SUM)
  optimizers      reference python/lib/joint.py:224-262, lib/utils/model_util.py:62-105  (A.9; constant LR, quirk Q1)
Parameter names follow the TensorFlow variable names the reference's checkpoint would hold.

Arithmetic: parameters are float32; reductions run in ``acc`` (float64 by default so the checker is
tighter than either fp32 implementation; float32 for the timed cpu_baseline).
"""
import ast
import re

import numpy as np
import scipy.sparse as sp

from . import columns as C

BN_EPS = 1e-3


# ------------------------------------------------------------------------------------------- helpers
def parse_optimizer(spec, default_lr):
    """'Adagrad' | 'Ftrl' | 'SGD' (use default_lr) or 'tf.train.FtrlOptimizer(learning_rate=..., ...)'
    (whose own learning_rate wins, model_util.py:95-101).  -> dict(kind, lr, l1, l2, lr_power, init_acc)"""
    names = {"Adagrad": "adagrad", "Ftrl": "ftrl", "SGD": "sgd", "Adam": "adam", "RMSProp": "rmsprop"}
    base = dict(l1=0.0, l2=0.0, lr_power=-0.5, init_acc=0.1, beta1=0.9, beta2=0.999, epsilon=1e-8, rho=0.9, momentum=0.0)
    if spec in names:
        o = dict(base, kind=names[spec], lr=float(default_lr))
        if o["kind"] == "rmsprop":
            o["epsilon"] = 1e-10                      # tf.train.RMSPropOptimizer default
        return o
    m = re.match(r"^\s*tf\.train\.(\w+)Optimizer\((.*)\)\s*$", spec)
    if not m:
        raise ValueError("Unsupported optimizer option: `{}`".format(spec))
    cls = {"Adagrad": "adagrad", "Ftrl": "ftrl", "GradientDescent": "sgd", "Adam": "adam", "RMSProp": "rmsprop"}.get(m.group(1))
    if cls is None:
        raise ValueError("Unsupported optimizer option: `{}`".format(spec))
    call = ast.parse("f(" + m.group(2) + ")", mode="eval").body
    kw = {k.arg: ast.literal_eval(k.value) for k in call.keywords}
    if call.args:
        kw.setdefault("learning_rate", ast.literal_eval(call.args[0]))
    if cls == "rmsprop" and kw.get("centered"):
        raise ValueError("centered RMSProp is not supported")
    return dict(kind=cls, lr=float(kw.get("learning_rate", 0.001 if cls == "adam" else default_lr)),
                l1=float(kw.get("l1_regularization_strength", 0.0)),
                l2=float(kw.get("l2_regularization_strength", 0.0)),
                lr_power=float(kw.get("learning_rate_power", -0.5)),
                init_acc=float(kw.get("initial_accumulator_value", 0.1)),
                beta1=float(kw.get("beta1", 0.9)), beta2=float(kw.get("beta2", 0.999)),
                epsilon=float(kw.get("epsilon", 1e-8 if cls == "adam" else 1e-10)),
                rho=float(kw.get("decay", 0.9)), momentum=float(kw.get("momentum", 0.0)))


def act_fwd(name, z):
    if name == "relu":
        return np.maximum(z, 0)
    if name == "relu6":
        return np.minimum(np.maximum(z, 0), 6)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    if name == "tanh":
        return np.tanh(z)
    if name == "leaky_relu":
        return np.where(z > 0, z, 0.2 * z)
    if name == "elu":
        return np.where(z > 0, z, np.expm1(np.minimum(z, 0)))
    if name == "selu":
        a, s = 1.6732632423543772, 1.0507009873554805
        return s * np.where(z > 0, z, a * np.expm1(np.minimum(z, 0)))
    if name == "softplus":
        return np.logaddexp(z, 0)
    if name == "softsign":
        return z / (1 + np.abs(z))
    if name == "crelu":          # tf.nn.crelu (reference model_util.py:45-50): concat([relu(z), relu(-z)], axis=-1)
        return np.concatenate([np.maximum(z, 0), np.maximum(-z, 0)], axis=-1)
    raise ValueError("Unsupported activation name: {}".format(name))


def act_bwd(name, z, a):
    """d act / d z, given pre-activation z and post-activation a."""
    if name == "relu":
        return (z > 0).astype(z.dtype)
    if name == "relu6":
        return ((z > 0) & (z < 6)).astype(z.dtype)
    if name == "sigmoid":
        return a * (1 - a)
    if name == "tanh":
        return 1 - a * a
    if name == "leaky_relu":
        return np.where(z > 0, 1.0, 0.2).astype(z.dtype)
    if name == "elu":
        return np.where(z > 0, 1.0, a + 1.0)
    if name == "selu":
        al, s = 1.6732632423543772, 1.0507009873554805
        return np.where(z > 0, s, a + s * al)
    if name == "softplus":
        return 1.0 / (1.0 + np.exp(-z))
    if name == "softsign":
        return 1.0 / (1 + np.abs(z)) ** 2
    raise ValueError(name)


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def drop_keep(seed, step, layer_id, rows, cols, rate):
    """Keep mask [rows, cols] (bool) of the dropout after hidden layer `layer_id` (= tower * 64 + layer) in train step `step`:
    element (m, n) is kept iff u(m, n) >= rate, u = top 24 bits of splitmix64(key ^ (m * 65536 + n)) / 2^24,
    key = splitmix64(seed ^ step * 0x9E3779B97F4A7C15 ^ layer_id << 48).  (tf.layers.dropout keeps an element iff its uniform
    draw >= rate and scales the kept ones by 1 / (1 - rate); the draw itself is TensorFlow's stream, replaced by this one.)"""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.uint64(seed) ^ ((np.uint64(step) * np.uint64(0x9E3779B97F4A7C15)) & _M64) ^ (np.uint64(layer_id) << np.uint64(48)))
        m = np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(65536)
        n = np.arange(cols, dtype=np.uint64)[None, :]
        r = _splitmix64(key ^ (m + n))
    u = (r >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u >= np.float32(rate)


def layer_sources(mode, L):
    """Concat order of every hidden layer's input and of the logits layer's input.
    -> list of L+1 lists whose items are 'x' or int j (output of hidden layer j)."""
    hid = []
    for l in range(L):
        if l == 0:
            hid.append(["x"])
        elif mode in ("simple", "last_dense"):
            hid.append([l - 1])
        elif mode == "first_dense":
            hid.append([l - 1, "x"])
        elif mode == "dense":
            hid.append(["x"] + list(range(l)))
        elif mode == "resnet":
            hid.append(list(range(l - 1, -1, -1)) + ["x"])
        else:
            raise AssertionError("Invalid connected_mode: {}".format(mode))
    if L == 0:
        last = ["x"]
    elif mode == "simple":
        last = [L - 1]
    elif mode == "first_dense":
        last = [L - 1, "x"]
    elif mode in ("last_dense", "dense"):
        last = ["x"] + list(range(L))
    elif mode == "resnet":
        last = list(range(L - 1, -1, -1)) + ["x"]
    else:
        raise AssertionError("Invalid connected_mode: {}".format(mode))
    return hid + [last]


def _csr(offs, ids, ncols, data=None, dtype=np.float64):
    nnz = len(ids)
    d = np.ones(nnz, dtype=dtype) if data is None else data.astype(dtype)
    return sp.csr_matrix((d, ids.astype(np.int64), offs.astype(np.int64)), shape=(len(offs) - 1, ncols))


# --------------------------------------------------------------------------------------------- model
class OracleModel(object):
    def __init__(self, feature_conf, cross_conf, model_conf, model_type="wide_deep",
                 embedding_dim_override=None, tf_compat_pad=False, acc=np.float64):
        assert model_type in ("wide", "deep", "wide_deep"), "Invalid model type: {}".format(model_type)
        self.model_type, self.acc, self.tf_compat_pad = model_type, acc, tf_compat_pad
        self.wide_cols, deep_cols = C.build_columns(feature_conf, cross_conf, embedding_dim_override)
        self.deep_cols = sorted(deep_cols, key=lambda c: c.name)          # A.7 sorted column-name order
        self.use_wide = model_type != "deep"
        self.use_deep = model_type != "wide"
        self.act = model_conf.get("dnn_activation_function") or "relu"
        self.bn = bool(model_conf.get("dnn_batch_normalization"))
        # tf.layers.dropout(net, rate, training=True) after every hidden layer's activation, TRAIN mode only (reference dnn.py:
        # 111-112 and the sibling blocks).  TensorFlow's random stream cannot be reproduced, so the keep mask is DEFINED here by a
        # counter-based generator (drop_keep) that the CUDA path implements bit for bit: same mask, same scaling by 1 / (1 - rate).
        self.dropout = float(model_conf.get("dnn_dropout") or 0.0)
        self.dropout_seed = 0x5EED0006
        hu = model_conf["dnn_hidden_units"]
        self.towers = [list(h) for h in hu] if hu and isinstance(hu[0], (list, tuple)) else [list(hu)]
        cm = model_conf.get("dnn_connected_mode") or "simple"
        self.modes = [cm] * len(self.towers) if isinstance(cm, str) else list(cm)
        self.opt_lin = parse_optimizer(model_conf["linear_optimizer"], model_conf.get("linear_initial_learning_rate") or 0.005)
        self.opt_dnn = parse_optimizer(model_conf["dnn_optimizer"], model_conf.get("dnn_initial_learning_rate") or 0.001)
        self.deep_offsets, off = {}, 0
        for c in self.deep_cols:
            self.deep_offsets[c.name] = off
            off += c.width if not isinstance(c, C.Numeric) else 1
        self.D0 = off
        self.params, self.slots = {}, {}
        self.global_step = 0

    # ---- names
    @staticmethod
    def wname(col):
        return "linear/linear_model/%s/weights" % col.name

    @staticmethod
    def ename(col):
        return "dnn/input_from_feature_columns/input_layer/%s/embedding_weights" % col.name

    def layer_dims(self, t):
        """[(in_dim, out_dim)] for hidden layers then logits of tower t."""
        hu, srcs = self.towers[t], layer_sources(self.modes[t], len(self.towers[t]))
        w = lambda s: self.D0 if s == "x" else self.out_width(hu[s])
        return [(sum(w(s) for s in srcs[l]), hu[l] if l < len(hu) else 1) for l in range(len(hu) + 1)]

    def out_width(self, units):
        """Features a hidden layer of `units` units hands on: tf.nn.crelu doubles them."""
        return 2 * units if self.act == "crelu" else units

    # ---- init (same distributions as TF: A.7 truncated normal, A.8 glorot uniform, zeros)
    def init(self, seed=0):
        rng = np.random.RandomState(seed)
        P = self.params = {}
        if self.use_wide:
            for c in self.wide_cols:
                P[self.wname(c)] = np.zeros(c.num_buckets, dtype=np.float32)
            P["linear/linear_model/bias_weights"] = np.zeros(1, dtype=np.float32)
        if self.use_deep:
            for c in self.deep_cols:
                if isinstance(c, C.Embedding):
                    n, d = c.cat.num_buckets, c.dim
                    g = np.random.default_rng([seed, len(P)])
                    w = g.standard_normal((n, d), dtype=np.float32)
                    bad = np.flatnonzero(np.abs(w.reshape(-1)) > 2)
                    while bad.size:
                        w.reshape(-1)[bad] = g.standard_normal(bad.size, dtype=np.float32)
                        bad = bad[np.abs(w.reshape(-1)[bad]) > 2]
                    w *= np.float32(1.0 / np.sqrt(d))
                    P[self.ename(c)] = w
            for t in range(len(self.towers)):
                dims = self.layer_dims(t)
                for l, (i, o) in enumerate(dims):
                    scope = "dnn/dnn_%d/" % (t + 1) + ("hiddenlayer_%d" % l if l < len(dims) - 1 else "logits")
                    lim = np.sqrt(6.0 / (i + o))
                    P[scope + "/kernel"] = rng.uniform(-lim, lim, size=(i, o)).astype(np.float32)
                    P[scope + "/bias"] = np.zeros(o, dtype=np.float32)
                    if self.bn and l < len(dims) - 1:
                        P[scope + "/batch_normalization/gamma"] = np.ones(self.out_width(o), dtype=np.float32)
                        P[scope + "/batch_normalization/beta"] = np.zeros(self.out_width(o), dtype=np.float32)
        self.reset_slots()
        return self

    def reset_slots(self):
        self.slots = {}
        for k, v in self.params.items():
            o = self.opt_lin if k.startswith("linear/") else self.opt_dnn
            if o["kind"] == "adagrad":
                self.slots[k] = {"acc": np.full_like(v, o["init_acc"])}
            elif o["kind"] == "ftrl":
                self.slots[k] = {"n": np.full_like(v, o["init_acc"]), "z": np.zeros_like(v)}
            elif o["kind"] == "adam":                 # tf.train.AdamOptimizer slots m, v = 0
                self.slots[k] = {"m": np.zeros_like(v), "v": np.zeros_like(v)}
            elif o["kind"] == "rmsprop":              # tf.train.RMSPropOptimizer slots rms = 1, momentum = 0
                self.slots[k] = {"ms": np.ones_like(v), "mom": np.zeros_like(v)}
            else:
                self.slots[k] = {}
        # Adam's non-slot variables, one pair per optimizer (linear / dnn): beta^t, multiplied in fp32 after every step
        self.beta_pow = {"lin": [np.float32(self.opt_lin.get("beta1", 0.9)), np.float32(self.opt_lin.get("beta2", 0.999))],
                         "dnn": [np.float32(self.opt_dnn.get("beta1", 0.9)), np.float32(self.opt_dnn.get("beta2", 0.999))]}

    # ---- forward
    def transform(self, batch):
        """column name -> CSR (offsets, ids): every categorical column the model reads."""
        out = {}
        cats = []
        if self.use_wide:
            cats += self.wide_cols
        if self.use_deep:
            cats += [c.cat for c in self.deep_cols if not isinstance(c, C.Numeric)]
        for c in cats:
            if c.name not in out:
                out[c.name] = c.ids(batch, tf_compat_pad=self.tf_compat_pad)
        return out

    def forward(self, batch, ids=None, train=False):
        A = self.acc
        ids = ids if ids is not None else self.transform(batch)
        B = len(next(iter(ids.values()))[0]) - 1 if ids else len(next(iter(batch.values())))
        cache = {"ids": ids, "B": B}
        logits = np.zeros(B, dtype=A)
        if self.use_wide:
            wl = np.full(B, self.params["linear/linear_model/bias_weights"][0], dtype=A)
            for c in self.wide_cols:
                offs, cid = ids[c.name]
                wl += _csr(offs, cid, c.num_buckets, dtype=A) @ self.params[self.wname(c)].astype(A, copy=False)
            cache["wide_logit"] = wl
            logits += wl
        if self.use_deep:
            X = np.zeros((B, self.D0), dtype=A)
            for c in self.deep_cols:
                o = self.deep_offsets[c.name]
                if isinstance(c, C.Numeric):
                    X[:, o] = c.values(batch)
                elif isinstance(c, C.Indicator):
                    offs, cid = ids[c.cat.name]
                    X[:, o:o + c.width] = _csr(offs, cid, c.width, dtype=A).toarray()
                else:
                    offs, cid = ids[c.cat.name]
                    cnt = np.diff(offs)
                    w = np.repeat(1.0 / np.maximum(cnt, 1), cnt)
                    X[:, o:o + c.dim] = _csr(offs, cid, c.cat.num_buckets, w, dtype=A) @ self.params[self.ename(c)].astype(A, copy=False)
            cache["X"] = X
            dl = np.zeros(B, dtype=A)
            cache["towers"] = []
            for t in range(len(self.towers)):
                tc = self._tower_fwd(t, X, train)
                cache["towers"].append(tc)
                dl += tc["logit"]
            cache["deep_logit"] = dl
            logits += dl
        cache["logits"] = logits
        return logits.astype(np.float32), cache

    def _tower_fwd(self, t, X, train=False):
        A, hu = self.acc, self.towers[t]
        srcs = layer_sources(self.modes[t], len(hu))
        H, Z, Aact, INP, D = [], [], [], [], []
        pick = lambda s: X if s == "x" else H[s]
        inv = 1.0 / np.sqrt(1.0 + BN_EPS)
        for l in range(len(hu)):
            scope = "dnn/dnn_%d/hiddenlayer_%d" % (t + 1, l)
            inp = np.concatenate([pick(s) for s in srcs[l]], axis=1)
            z = inp @ self.params[scope + "/kernel"].astype(A, copy=False) + self.params[scope + "/bias"].astype(A, copy=False)
            a = act_fwd(self.act, z)
            D.append(None)
            if train and self.dropout > 0:
                D[-1] = drop_keep(self.dropout_seed, self.global_step, t * 64 + l, a.shape[0], a.shape[1], self.dropout).astype(A) \
                    * A(1.0 / (1.0 - np.float32(self.dropout)))
                a_in = a * D[-1]
            else:
                a_in = a
            if self.bn:   # A.8: always inference mode, moving mean 0 / var 1 (quirk Q4)
                h = a_in * (self.params[scope + "/batch_normalization/gamma"].astype(A, copy=False) * inv) \
                    + self.params[scope + "/batch_normalization/beta"].astype(A, copy=False)
            else:
                h = a_in
            INP.append(inp); Z.append(z); Aact.append(a); H.append(h)
        scope = "dnn/dnn_%d/logits" % (t + 1)
        inp = np.concatenate([pick(s) for s in srcs[-1]], axis=1)
        logit = (inp @ self.params[scope + "/kernel"].astype(A, copy=False) + self.params[scope + "/bias"].astype(A, copy=False))[:, 0]
        INP.append(inp)
        return dict(H=H, Z=Z, A=Aact, INP=INP, D=D, logit=logit, srcs=srcs)

    # ---- loss
    @staticmethod
    def loss_terms(logits, labels):
        x, z = logits.astype(np.float64), labels.astype(np.float64)
        return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))

    def loss(self, logits, labels, weights=None):
        w = np.ones_like(labels, dtype=np.float64) if weights is None else weights.astype(np.float64)
        return float((w * self.loss_terms(logits, labels)).sum())        # SUM reduction (Q11)

    # ---- backward: returns grads dict name -> dense ndarray | (unique_rows, grad_rows)
    def backward(self, cache, labels, weights=None):
        A = self.acc
        x = cache["logits"].astype(np.float64)
        w = np.ones_like(x) if weights is None else weights.astype(np.float64)
        with np.errstate(over="ignore"):                 # exp(-x) -> inf for very negative logits: 1 / inf = 0 is the right sigmoid
            dlogit = ((1.0 / (1.0 + np.exp(-x)) - labels.astype(np.float64)) * w).astype(A)
        grads, ids, B = {}, cache["ids"], cache["B"]
        if self.use_wide:
            grads["linear/linear_model/bias_weights"] = np.array([dlogit.sum()])
            for c in self.wide_cols:
                offs, cid = ids[c.name]
                u, inv = np.unique(cid, return_inverse=True)
                g = _csr(offs, inv, len(u), dtype=A).T @ dlogit
                grads[self.wname(c)] = (u, np.asarray(g))
        if self.use_deep:
            dX = np.zeros_like(cache["X"])
            for t, tc in enumerate(cache["towers"]):
                self._tower_bwd(t, tc, dlogit, dX, grads)
            for c in self.deep_cols:
                if isinstance(c, C.Embedding):
                    o = self.deep_offsets[c.name]
                    offs, cid = ids[c.cat.name]
                    cnt = np.diff(offs)
                    wgt = np.repeat(1.0 / np.maximum(cnt, 1), cnt)
                    u, inv = np.unique(cid, return_inverse=True)
                    g = _csr(offs, inv, len(u), wgt, dtype=A).T @ dX[:, o:o + c.dim]
                    grads[self.ename(c)] = (u, np.asarray(g))
            cache["dX"] = dX
        cache["dlogit"] = dlogit
        return grads

    def _tower_bwd(self, t, tc, dlogit, dX, grads):
        A, hu, srcs = self.acc, self.towers[t], tc["srcs"]
        L = len(hu)
        dH = [np.zeros_like(h) for h in tc["H"]]
        inv = 1.0 / np.sqrt(1.0 + BN_EPS)

        def scatter(dinp, sources):
            o = 0
            for s in sources:
                wd = dX.shape[1] if s == "x" else self.out_width(hu[s])
                if s == "x":
                    dX[:, :] += dinp[:, o:o + wd]
                else:
                    dH[s] += dinp[:, o:o + wd]
                o += wd

        scope = "dnn/dnn_%d/logits" % (t + 1)
        K = self.params[scope + "/kernel"].astype(A, copy=False)
        grads[scope + "/kernel"] = tc["INP"][L].T @ dlogit[:, None]
        grads[scope + "/bias"] = np.array([dlogit.sum()])
        scatter(dlogit[:, None] @ K.T, srcs[L])
        for l in range(L - 1, -1, -1):
            scope = "dnn/dnn_%d/hiddenlayer_%d" % (t + 1, l)
            dh = dH[l]
            if self.bn:
                gam = self.params[scope + "/batch_normalization/gamma"].astype(A, copy=False)
                a_drop = tc["A"][l] if tc["D"][l] is None else tc["A"][l] * tc["D"][l]
                grads[scope + "/batch_normalization/gamma"] = (dh * a_drop).sum(0) * inv
                grads[scope + "/batch_normalization/beta"] = dh.sum(0)
                da = dh * (gam * inv)
            else:
                da = dh
            if tc["D"][l] is not None:
                da = da * tc["D"][l]                                  # d(dropout) = mask / keep_prob
            if self.act == "crelu":                                   # da is [B, 2u]: d relu(z) on the first half, d relu(-z) = -1[z < 0] on the second
                z, u = tc["Z"][l], tc["Z"][l].shape[1]
                dz = da[:, :u] * (z > 0) - da[:, u:] * (z < 0)
            else:
                dz = da * act_bwd(self.act, tc["Z"][l], tc["A"][l])
            grads[scope + "/kernel"] = tc["INP"][l].T @ dz
            grads[scope + "/bias"] = dz.sum(0)
            scatter(dz @ self.params[scope + "/kernel"].astype(A, copy=False).T, srcs[l])

    # ---- optimizers (A.9)
    def apply(self, grads):
        for name, g in grads.items():
            o = self.opt_lin if name.startswith("linear/") else self.opt_dnn
            p, s = self.params[name], self.slots[name]
            if isinstance(g, tuple):
                rows, gr = g
                if len(rows) == 0:
                    continue
                gr = gr.reshape((len(rows),) + p.shape[1:]).astype(np.float32)
                view = lambda arr: arr[rows]
            else:
                rows, gr = slice(None), np.asarray(g).reshape(p.shape).astype(np.float32)
                view = lambda arr: arr
            lr = np.float32(o["lr"])
            if o["kind"] == "adagrad":
                acc = view(s["acc"]) + gr * gr
                s["acc"][rows] = acc
                p[rows] = view(p) - lr * gr / np.sqrt(acc)
            elif o["kind"] == "ftrl":
                assert o["lr_power"] == -0.5
                n0, z0, w0 = view(s["n"]), view(s["z"]), view(p)
                n1 = n0 + gr * gr
                z1 = z0 + gr - (np.sqrt(n1) - np.sqrt(n0)) / lr * w0
                l1, l2 = np.float32(o["l1"]), np.float32(o["l2"])
                w1 = np.where(np.abs(z1) > l1, (np.sign(z1) * l1 - z1) / (np.sqrt(n1) / lr + np.float32(2.0) * l2), np.float32(0.0))
                s["n"][rows], s["z"][rows], p[rows] = n1, z1, w1.astype(np.float32)
            elif o["kind"] == "rmsprop":
                # ApplyRMSProp / SparseApplyRMSProp (touched rows only): ms += (g^2 - ms)(1 - rho); mom = mom*momentum + lr*g*rsqrt(ms + eps)
                rho, mo, eps = np.float32(o["rho"]), np.float32(o["momentum"]), np.float32(o["epsilon"])
                ms = view(s["ms"]) + (gr * gr - view(s["ms"])) * (np.float32(1) - rho)
                mom = view(s["mom"]) * mo + (gr * lr) / np.sqrt(ms + eps)
                s["ms"][rows], s["mom"][rows] = ms, mom
                p[rows] = view(p) - mom
            elif o["kind"] == "adam":
                # ApplyAdam; sparse gradients (AdamOptimizer._apply_sparse_shared): m and v decay over the WHOLE variable, the summed
                # gradients are scatter-added, then every row moves
                b1, b2, eps = np.float32(o["beta1"]), np.float32(o["beta2"]), np.float32(o["epsilon"])
                b1p, b2p = self.beta_pow["lin" if name.startswith("linear/") else "dnn"]
                lr_t = lr * np.sqrt(np.float32(1) - b2p) / (np.float32(1) - b1p)
                if isinstance(g, tuple):
                    s["m"] *= b1
                    s["v"] *= b2
                    s["m"][rows] += gr * (np.float32(1) - b1)
                    s["v"][rows] += gr * gr * (np.float32(1) - b2)
                else:
                    s["m"] += (gr - s["m"]) * (np.float32(1) - b1)
                    s["v"] += (gr * gr - s["v"]) * (np.float32(1) - b2)
                p -= lr_t * s["m"] / (np.sqrt(s["v"]) + eps)
            else:
                p[rows] = view(p) - lr * gr
        for key, o in (("lin", self.opt_lin), ("dnn", self.opt_dnn)):      # AdamOptimizer._finish
            if o["kind"] == "adam":
                self.beta_pow[key][0] = np.float32(self.beta_pow[key][0] * np.float32(o["beta1"]))
                self.beta_pow[key][1] = np.float32(self.beta_pow[key][1] * np.float32(o["beta2"]))
        self.global_step += 1

    def train_step(self, batch, labels, weights=None):
        logits, cache = self.forward(batch, train=True)
        loss = self.loss(cache["logits"], labels, weights)
        self.apply(self.backward(cache, labels, weights))
        return loss, logits

    def predict(self, batch):
        logits, _ = self.forward(batch)
        p = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
        return dict(logits=logits, logistic=p.astype(np.float32),
                    probabilities=np.stack([1 - p, p], 1).astype(np.float32), class_ids=(logits > 0).astype(np.int64))
