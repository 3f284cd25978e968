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
