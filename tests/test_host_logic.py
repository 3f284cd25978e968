"""CPU tests of the host side: YAML config surface, plan compiler, TSV loader vs the oracle's parser."""
import os

import numpy as np
import pytest

from wide_deep_b200.config import Config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_surface_matches_reference_shapes():
    cfg = Config()
    schema = cfg.read_schema()
    assert len(schema) == 61 and schema[1] == "clk"
    fc = cfg.read_feature_conf()
    assert len(fc) == 39 and fc["adplan_id"] == {"type": "category", "transform": "hash_bucket", "parameter": 10000}
    cc = cfg.read_cross_feature_conf()
    assert len(cc) == 31
    d = {"&".join(n): (s, deep) for n, s, deep in cc}
    assert d["adplan_id&category"] == (100000, 1)
    assert d["age&ugender"] == (100, 1)            # hash_bucket_size 0.1 -> 100 (quirk Q8, int here)
    assert cfg.train["batch_size"] == 64 and cfg.train["dynamic_train"] is True
    assert cfg.model["dnn_hidden_units"] == [1024, 512, 256]
    assert len(cfg.get_feature_name("all")) == 60 and len(cfg.get_feature_name("used")) == 39
    assert len(cfg.get_feature_name("category")) == 36 and len(cfg.get_feature_name("continuous")) == 3
    with pytest.raises(ValueError):
        cfg.get_feature_name("bogus")


def test_config_validation_errors(tmp_path):
    import shutil
    d = tmp_path / "conf"
    shutil.copytree(os.path.join(ROOT, "conf"), d)
    (d / "feature.yaml").write_text("adplan_id: {type: category, transform: hash_bucket, parameter: abc}\n")
    with pytest.raises(TypeError):
        Config(conf_dir=str(d)).read_feature_conf()
    (d / "feature.yaml").write_text("nope: {type: category, transform: hash_bucket, parameter: 10}\n")
    with pytest.raises(ValueError):
        Config(conf_dir=str(d)).read_feature_conf()
    (d / "feature.yaml").write_text("age: {type: continuous, transform: min_max, parameter: {normalization: [90, 10], boundaries: [1]}}\n")
    with pytest.raises(AssertionError):
        Config(conf_dir=str(d)).read_feature_conf()


def test_plan_matches_reference_dimensions(native_lib):
    from wide_deep_b200.plan import compile_plan, embedding_dim
    p = compile_plan(Config())
    s = p.summary()
    # 70 wide columns / 12,714,809 wide rows; 47 tables / 12,714,400 rows / 353.7M params; deep input 734
    assert s["wide_columns"] == 70 and s["wide_rows"] == 12714809
    assert s["tables"] == 47 and s["table_rows"] == 12714400 and s["table_params"] == 353669600
    assert s["deep_dim"] == 734 and s["deep_dim_phys"] % 32 == 0
    # reference heuristic int(2**ceil(ln(n**0.25))) (natural log, quirk Q12)
    assert [embedding_dim(n) for n in (100, 10000, 20000, 500000, 10000000)] == [4, 8, 8, 16, 32]
    names = list(p.deep_layout)
    assert names == sorted(names)                      # input_layer concatenates in sorted column-name order
    assert p.layer_dims(0) == [(734, 1024), (1024, 512), (512, 256), (256, 1)]
    # cross key order = SparseCross op order: categorical-column keys first, then raw string keys
    c = next(c for c in p.columns if c.name == "age_bucketized_X_scheduling_id")
    assert [k[0] for k in c.keys] == [1, 0]
    assert {k: p.lin_opt[k] for k in ("kind", "lr", "l1", "l2", "lr_power", "init_acc")} == dict(kind="ftrl", lr=0.1, l1=0.5, l2=1.0, lr_power=-0.5, init_acc=0.1)
    assert p.dnn_opt["kind"] == "adagrad" and p.dnn_opt["lr"] == 0.05


def test_optimizer_parsing():
    from wide_deep_b200.plan import parse_optimizer
    assert parse_optimizer("Adagrad", 0.05)["lr"] == 0.05
    o = parse_optimizer("tf.train.FtrlOptimizer(learning_rate=0.1,l1_regularization_strength=0.5,l2_regularization_strength=1)", 9.9)
    assert (o["kind"], o["lr"], o["l1"], o["l2"]) == ("ftrl", 0.1, 0.5, 1.0)
    with pytest.raises(ValueError):
        parse_optimizer("__import__('os').system('true')", 0.1)     # never eval()'ed
    a = parse_optimizer("Adam", 0.1)                                 # the reference's five names are all offered (model_util.py:84-90)
    assert (a["kind"], a["lr"], a["beta1"], a["beta2"], a["epsilon"]) == ("adam", 0.1, 0.9, 0.999, 1e-8)
    r = parse_optimizer("tf.train.RMSPropOptimizer(learning_rate=0.01, decay=0.8, momentum=0.5)", 9.9)
    assert (r["kind"], r["lr"], r["rho"], r["momentum"], r["epsilon"]) == ("rmsprop", 0.01, 0.8, 0.5, 1e-10)
    with pytest.raises(ValueError):
        parse_optimizer("tf.train.MomentumOptimizer(0.1, 0.9)", 0.1)


def test_crelu_shapes_match_oracle_and_all_ten_activation_names_parse():
    """The reference's ten activation names (model_util.py:28-59) are all accepted; crelu doubles what a layer hands on while its
    kernel / bias keep the conf's units (tf.layers.dense(units=u, activation=tf.nn.crelu) + batch_normalization over 2u)."""
    from oracle.model import OracleModel
    from tests.test_gpu_parity import small_conf
    from wide_deep_b200.plan import ACTS, Plan
    assert sorted(ACTS) == sorted(["sigmoid", "tanh", "relu", "relu6", "leaky_relu", "crelu", "elu", "selu", "softplus", "softsign"])
    for mode in ["simple", "first_dense", "last_dense", "dense", "resnet"]:
        hidden = (32, 32, 32) if mode == "resnet" else (64, 48, 16)
        fc, cross, model = small_conf(hidden=hidden, mode=mode, act="crelu", bn=1)
        p = Plan(fc, cross, model, "wide_deep", max_batch=8)
        om = OracleModel(fc, cross, model, "wide_deep").init(0)
        shapes = {k: tuple(v[3]) for k, v in p.tensor_names.items() if k.startswith("dnn/dnn_1/")}
        assert shapes == {k: v.shape for k, v in om.params.items() if k.startswith("dnn/dnn_1/")}, mode
        assert shapes["dnn/dnn_1/hiddenlayer_0/kernel"] == (p.d0, hidden[0])
        assert shapes["dnn/dnn_1/hiddenlayer_0/batch_normalization/gamma"] == (2 * hidden[0],)
        desc, keep = p.to_c()                                        # the library receives the conf's units and WD_ACT_CRELU (= 9)
        assert desc.activation == 9 and ACTS[desc.activation] == "crelu"
        import ctypes
        hu = ctypes.cast(desc.hidden_units, ctypes.POINTER(ctypes.c_int32))
        assert [hu[i] for i in range(len(hidden))] == list(hidden)
    with pytest.raises(ValueError):
        small = small_conf(act="swish")
        Plan(small[0], small[1], small[2], "wide_deep", max_batch=8)


def test_layer_sources_match_oracle():
    from oracle.model import layer_sources as o_src
    from wide_deep_b200.plan import Plan
    for mode in ["simple", "first_dense", "last_dense", "dense", "resnet"]:
        for L in range(0, 5):
            assert Plan.layer_sources(mode, L) == o_src(mode, L), (mode, L)


@pytest.mark.parametrize("pad", [False, True])
def test_tsv_loader_matches_oracle_parser(native_lib, pad):
    """C++ loader (product) vs pure-Python restatement of dataset.py (oracle) on the bundled fixtures."""
    from oracle import tsv as otsv
    from tests.helpers import to_product_batch
    from wide_deep_b200.dataset import TsvReader
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    plan = compile_plan(cfg, tf_compat_pad=pad)
    lines = open(os.path.join(ROOT, "data", "train", "train1")).read().split("\n")[:64]
    lines[3] = "\t".join("-" if i in (3, 21, 56, 50) else f for i, f in enumerate(lines[3].split("\t")))   # NA tokens
    got = TsvReader(cfg, plan).parse(lines)
    raw, lab = otsv.parse_lines(lines, cfg.read_schema(), cfg.read_feature_conf())
    exp = to_product_batch(plan, raw, lab, tf_compat_pad=pad)
    assert np.array_equal(got.offsets, exp.offsets)
    assert np.array_equal(got.keys, exp.keys)
    assert np.array_equal(got.dense, exp.dense)
    assert np.array_equal(got.label, exp.label)


def test_tsv_loader_errors(native_lib):
    from wide_deep_b200.dataset import TsvReader
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    r = TsvReader(cfg, compile_plan(cfg))
    with pytest.raises(ValueError):
        r.parse(["0\tonly\tthree"])


def test_pred_files_have_no_label(native_lib):
    from wide_deep_b200.dataset import input_fn
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    plan = compile_plan(cfg)
    b = next(input_fn(os.path.join(ROOT, "data", "pred"), None, "pred", 16, config=cfg, plan=plan))
    assert b.batch_size == 16 and b.label is None


def test_shard_rows_cover_batch():
    from wide_deep_b200.parallel import shard_rows
    for n in (1, 7, 64, 65536):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def test_dense_exchange_layout_and_list_bounds():
    """dense_exchange_max_rows: the plan orders the wide columns large-first (small ones at the end of the row space, from
    wide_small_base on) without changing any column's size, and exchange_rows() counts only the large tables / columns."""
    from wide_deep_b200 import synthetic
    from wide_deep_b200.plan import Plan
    fc, cross, model, emb = synthetic.criteo_conf()
    p0 = Plan(fc, cross, model, "wide_deep", max_batch=512, embedding_dim_override=emb)
    p1 = Plan(fc, cross, model, "wide_deep", max_batch=512, embedding_dim_override=emb, dense_exchange_max_rows=16384)
    assert p0.wide_rows == p1.wide_rows and p0.wide_small_base == p0.wide_rows      # nothing is small without the option
    assert [c.name for c in p0.wide_columns] == [c.name for c in p1.wide_columns]
    spans = sorted((c.wide_base, c.wide_base + c.buckets, c.buckets) for c in p1.wide_columns)
    assert spans[0][0] == 0 and spans[-1][1] == p1.wide_rows
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))                     # the columns tile the row space
    for lo, hi, n in spans:
        assert (n <= 16384) == (lo >= p1.wide_small_base)                          # small columns exactly behind the base
    n_big_emb = sum(1 for t in p1.tables if t["rows"] > 16384)
    n_big_wide = sum(1 for c in p1.wide_columns if c.buckets > 16384)
    assert (n_big_emb, n_big_wide) == (8, 16)
    assert p1.exchange_rows(512) == (512 * 8, 512 * 16)
    assert p0.exchange_rows(512) == (512 * len(p0.tables), 512 * len(p0.wide_columns))
    d, _keep = p1.to_c()
    assert d.dense_exchange_max_rows == 16384 and d.wide_small_base == p1.wide_small_base


@pytest.mark.parametrize("tf_compat_pad", [False, True])
def test_pinned_ring_and_prefetch_thread_yield_the_same_batches(native_lib, tf_compat_pad):
    """input_fn(pinned=True) parses into a ring of (page-locked, when a GPU exists) buffer sets and the estimator drains it from a
    prefetch thread: same batches as the plain path, batch by batch; sharded input (rank, world) has equal batch counts."""
    from wide_deep_b200.dataset import Prefetcher, input_fn
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    plan = compile_plan(cfg, "wide_deep", 64, tf_compat_pad=tf_compat_pad)
    data = os.path.join(ROOT, "data", "test", "test1")
    plain = list(input_fn(data, None, "eval", 64, config=cfg, plan=plan))[:12]
    got = 0
    for i, b in enumerate(Prefetcher(input_fn(data, None, "eval", 64, config=cfg, plan=plan, pinned=True), depth=2)):
        if i >= len(plain):
            break
        a = plain[i]
        assert b.batch_size == a.batch_size
        np.testing.assert_array_equal(b.keys, a.keys)
        np.testing.assert_array_equal(b.offsets, a.offsets)
        np.testing.assert_array_equal(b.dense, a.dense)
        np.testing.assert_array_equal(b.label, a.label)
        got += 1
    assert got == len(plain)
    counts = [sum(1 for _ in input_fn(data, None, "train", 64, config=cfg, plan=plan, rank=r, world=3)) for r in range(3)]
    assert counts[0] == counts[1] == counts[2] == -(-(5000 // 3) // 64)


@pytest.mark.parametrize("tf_compat_pad", [False, True])
@pytest.mark.parametrize("mode,rank,world", [("eval", 0, 1), ("train", 0, 1), ("train", 1, 3)])
def test_input_fn_over_the_line_index_equals_the_line_list_path(native_lib, tmp_path, tf_compat_pad, mode, rank, world):
    """input_fn indexes the file image once (wd_tsv_index_lines) and parses each batch's lines in place (wd_tsv_parse_lines); what
    it yields must be, batch by batch, what the straightforward procedure gives: split the files into lines, drop empty ones,
    take every world-th line from `rank`, shuffle with Philox(seed) in train mode, join and parse batch_size lines at a time
    (reference dataset.py:167-184: TextLineDataset -> shard -> shuffle -> batch).  Two files, CRLF line ends, blank lines and a
    missing final newline included."""
    from wide_deep_b200.dataset import TsvReader, input_fn
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    plan = compile_plan(cfg, "wide_deep", 100, tf_compat_pad=tf_compat_pad)
    src = open(os.path.join(ROOT, "data", "test", "test1"), "rb").read().split(b"\n")
    src = [l for l in src if l][:700]
    d = tmp_path / "data"
    d.mkdir()
    (d / "part-0").write_bytes(b"\n".join(src[:300]) + b"\n\n\n")                 # blank lines at the end
    (d / "part-1").write_bytes(b"\r\n".join(src[300:]))                            # CRLF, no final newline
    (d / ".hidden").write_bytes(b"not data\n")
    lines = src[:300] + [l + b"\r" for l in src[300:-1]] + [src[-1]]
    if world > 1:
        lines = lines[rank::world][:len(lines) // world]
    if mode == "train":
        perm = np.random.Generator(np.random.Philox(123)).permutation(len(lines))
        lines = [lines[i] for i in perm]
    reader = TsvReader(cfg, plan)
    want = [reader.parse(lines[i:i + 100]) for i in range(0, len(lines), 100)]
    for pinned in (False, True):
        got = 0
        for a, b in zip(want, input_fn(str(d), None, mode, 100, config=cfg, plan=plan, rank=rank, world=world, pinned=pinned)):
            assert b.batch_size == a.batch_size
            np.testing.assert_array_equal(b.keys, a.keys)
            np.testing.assert_array_equal(b.offsets, a.offsets)
            np.testing.assert_array_equal(b.dense, a.dense)
            np.testing.assert_array_equal(b.label, a.label)
            got += 1
        assert got == len(want) == -(-len(lines) // 100)


def test_tsv_two_call_protocol_parses_once_and_never_mixes_batches(native_lib):
    """wd_tsv_parse with no key buffer (or one that is too small) only counts and keeps its parse; the follow-up call with the same
    arguments copies the keys out.  A follow-up call for DIFFERENT text (same sizes, same output arrays) must parse afresh."""
    from wide_deep_b200.dataset import TsvReader
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    plan = compile_plan(cfg, "wide_deep", 64, tf_compat_pad=True)
    reader = TsvReader(cfg, plan)
    src = [l for l in open(os.path.join(ROOT, "data", "test", "test1"), "rb").read().split(b"\n") if l]
    a_lines, b_lines = src[:64], src[64:128]
    want_a, want_b = reader.parse(a_lines), reader.parse(b_lines)
    lib, spec = reader._lib, reader._spec
    F, Nd = len(plan.cat_fields), len(plan.dense_fields)
    offs = np.zeros(64 * F + 1, dtype=np.int32); dense = np.zeros((64, Nd), dtype=np.float32)
    label = np.zeros(64, dtype=np.float32); weight = np.ones(64, dtype=np.float32)
    ta, tb = b"\n".join(a_lines), b"\n".join(b_lines)

    def call(text, keys):
        import ctypes
        return lib.wd_tsv_parse(ctypes.byref(spec), text, len(text), 64, offs.ctypes.data, keys.ctypes.data if keys is not None else None,
                                keys.size if keys is not None else 0, dense.ctypes.data, label.ctypes.data, weight.ctypes.data, 2)
    small = np.empty(8, dtype=np.uint64)
    nnz = call(ta, small)                                             # too small: counts only
    assert nnz == len(want_a.keys) > 8
    keys = np.empty(nnz, dtype=np.uint64)
    assert call(ta, keys) == nnz                                      # follow-up: the kept parse
    np.testing.assert_array_equal(keys, want_a.keys)
    np.testing.assert_array_equal(offs, want_a.offsets)
    assert call(ta, None) == nnz                                      # counting call for A ...
    nb = call(tb, np.empty(len(want_b.keys), dtype=np.uint64))        # ... followed by a call for B: parsed afresh
    kb = np.empty(nb, dtype=np.uint64)
    assert call(tb, kb) == nb == len(want_b.keys)
    np.testing.assert_array_equal(kb, want_b.keys)
    np.testing.assert_array_equal(offs, want_b.offsets)
    np.testing.assert_array_equal(dense, want_b.dense)


def test_tsv_fast_number_fields_equal_strtof_and_strtoll(native_lib):
    """The loader decodes plain [-]digits[.digits] fields itself and leaves every other shape to strtof / strtoll; its floats must
    be the correctly rounded float of the decimal — what libc's strtof returns — including decimals that sit on or next to the
    midpoint between two floats (where rounding through double would go wrong), and its ints what strtoll returns."""
    import ctypes
    import ctypes.util
    from wide_deep_b200._native import TsvSpecC
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.strtof.restype, libc.strtof.argtypes = ctypes.c_float, [ctypes.c_char_p, ctypes.c_void_p]
    libc.strtoll.restype, libc.strtoll.argtypes = ctypes.c_longlong, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
    rng = np.random.default_rng(17)
    floats = ["0", "-0", "-0.0", "1.", ".5", "-.5", "00012.50", "31.0", "1e5", "-2.5E-3", "16777217", "16777219", "33554434", "8388608.5",
              "8388609.5", "0.1", "0.30000000000000004", "123456789012345", "1234567890123456", "3.4028235e38", "1e-45", "7.006492321624085e-46",
              "4294967296.000001", "0.000000059604644775390625", "1.00000005960464477539", "9007199254740993"]
    for _ in range(20000):
        kind = int(rng.integers(0, 5))
        if kind == 0:                                                 # plain decimals, 1-15 significant digits
            digs = "".join(str(int(x)) for x in rng.integers(0, 10, size=int(rng.integers(1, 16))))
            k = int(rng.integers(0, len(digs) + 1))
            t = digs[:k] + "." + digs[k:] if rng.random() < 0.8 else digs
        elif kind == 1:                                               # integers around 2^24 .. 2^40: exact float midpoints are common
            t = str(int(rng.integers(1 << 24, 1 << 40)))
        elif kind == 2:                                               # midpoints between neighbouring floats, printed exactly when short
            f = np.float32(rng.uniform(0.001, 4096.0))
            mid = (float(f) + float(np.nextafter(f, np.float32(np.inf)))) / 2
            t = repr(mid) if rng.random() < 0.5 else "%.15g" % mid
        elif kind == 3:                                               # what %g / repr print for random floats
            t = repr(float(np.float32(rng.standard_normal() * 10 ** int(rng.integers(-6, 7)))))
        else:
            t = "%.*f" % (int(rng.integers(0, 12)), rng.uniform(-1e6, 1e6))
        floats.append(("-" if rng.random() < 0.2 and not t.startswith("-") else "") + t)
    ints = ["0", "-0", "7", "-7", "+7", "007", "999999999999999999", "-999999999999999999", "9223372036854775807", "-9223372036854775808",
            "1234567890123456789"] + [str(int(x)) for x in rng.integers(-2 ** 62, 2 ** 62, size=2000)]
    floats = [t for t in floats if t not in ("", "-")]               # (empty and '-' are the NA tokens)
    n = max(len(floats), len(ints))
    floats += ["1"] * (n - len(floats))
    ints += ["1"] * (n - len(ints))
    text = "\n".join("%s\t%s" % (a, b) for a, b in zip(floats, ints)).encode()
    role, target = np.asarray([3, 2], dtype=np.int32), np.asarray([0, 0], dtype=np.int32)
    spec = TsvSpecC()
    spec.n_columns, spec.col_role, spec.col_target = 2, role.ctypes.data, target.ctypes.data
    spec.n_cat_fields, spec.n_dense_fields, spec.multivalue, spec.tf_compat_pad = 1, 1, 0, 0
    spec.pos_weight, spec.neg_weight, spec.use_weight, spec.has_label = 1.0, 1.0, 0, 0
    offs, keys = np.zeros(n + 1, dtype=np.int32), np.zeros(n, dtype=np.uint64)
    dense, weight = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
    got = native_lib.wd_tsv_parse(ctypes.byref(spec), text, len(text), n, offs.ctypes.data, keys.ctypes.data, n, dense.ctypes.data, None,
                                  weight.ctypes.data, 3)
    assert got == n, native_lib.wd_last_error()
    want_f = np.asarray([libc.strtof(t.encode(), None) for t in floats], dtype=np.float32)
    want_i = np.asarray([libc.strtoll(t.encode(), None, 10) for t in ints], dtype=np.int64)
    bad = np.flatnonzero(dense.view(np.uint32) != want_f.view(np.uint32))
    assert bad.size == 0, [(floats[i], float(dense[i]), float(want_f[i])) for i in bad[:5]]
    np.testing.assert_array_equal(keys.view(np.int64), want_i)


def test_tsv_worker_pool_concurrent_callers_and_fork(native_lib):
    """The loader's persistent worker threads: several Python threads parsing at once with changing thread counts get the same
    batches as a single-threaded parse, and a forked child (whose copy of the pool has no threads) parses too."""
    import threading
    from wide_deep_b200.dataset import TsvReader
    from wide_deep_b200.plan import compile_plan
    cfg = Config()
    plan = compile_plan(cfg, "wide_deep", 4096, tf_compat_pad=True)
    src = [l for l in open(os.path.join(ROOT, "data", "test", "test1"), "rb").read().split(b"\n") if l]
    want = {}
    for o, n in [(0, 1), (3, 255), (100, 900), (500, 2500), (7, 4000)]:
        b = TsvReader(cfg, plan, n_threads=1).parse(src[o:o + n])
        want[(o, n)] = (b.keys.copy(), b.offsets.copy(), b.dense.copy())
    errors = []

    def job(tid):
        rng = np.random.default_rng(tid)
        try:
            for _ in range(40):
                (o, n), nt = list(want)[int(rng.integers(0, len(want)))], int(rng.integers(1, 17))
                b = TsvReader(cfg, plan, n_threads=nt).parse(src[o:o + n])
                k, f, d = want[(o, n)]
                assert np.array_equal(b.keys, k) and np.array_equal(b.offsets, f) and np.array_equal(b.dense, d), (o, n, nt)
        except Exception as e:                                          # surfaced on the main thread
            errors.append(e)
    ths = [threading.Thread(target=job, args=(i,)) for i in range(3)]
    [t.start() for t in ths]
    [t.join(120) for t in ths]
    assert not errors and not any(t.is_alive() for t in ths), errors
    pid = os.fork()
    if pid == 0:
        ok = 1
        try:
            b = TsvReader(cfg, plan, n_threads=8).parse(src[500:3000])
            ok = 0 if np.array_equal(b.keys, want[(500, 2500)][0]) else 2
        finally:
            os._exit(ok)
    for _ in range(600):                                                # a hung child must fail the test, not hang it
        done, status = os.waitpid(pid, os.WNOHANG)
        if done:
            break
        import time
        time.sleep(0.05)
    else:
        os.kill(pid, 9)
        os.waitpid(pid, 0)
        raise AssertionError("forked child hung in the TSV loader")
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0
    assert np.array_equal(TsvReader(cfg, plan, n_threads=8).parse(src[500:3000]).keys, want[(500, 2500)][0])


def test_sharded_plan_marks_large_tables_only(native_lib):
    from wide_deep_b200 import synthetic
    from wide_deep_b200.plan import Plan
    fc, cross, model, emb = synthetic.criteo_conf()
    p = Plan(fc, cross, model, "wide_deep", max_batch=512, embedding_dim_override=emb, shard_world=8, shard_rank=3)
    assert p.dense_exchange_max_rows == 16384
    assert sum(1 for t in p.tables if t["sharded"]) == 8 and sum(p.wide_sharded) == 16
    name = next(n for n in p.tensor_names if p.is_sharded_tensor(n) and "embedding_weights" in n)
    rows = p.tensor_names[name][3][0]
    assert p.local_shape(name) == ((rows - 3 + 7) // 8, 32)
    assert p.exchange_rows(512) == (0, 0)                 # nothing travels as (row, gradient) lists in a sharded run
    p1 = Plan(fc, cross, model, "wide_deep", max_batch=512, embedding_dim_override=emb)
    assert not any(t["sharded"] for t in p1.tables) and not any(p1.wide_sharded)
