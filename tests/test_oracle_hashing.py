"""Oracle pinned on TensorFlow-upstream known-answer vectors (tests/golden/hash_kat.json) — CPU only."""
import json
import os
import random

import numpy as np

from oracle import hashing as H

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hash_kat.json")))


def test_fingerprint64_kat():
    for s, v in KAT["fingerprint64"].items():
        assert H.fingerprint64(s) == int(v), s
        assert H.py_fingerprint64(s) == int(v), s
    for s, v in KAT["fingerprint64_signed"].items():
        assert int(np.uint64(H.fingerprint64(s)).astype(np.int64)) == int(v), s


def test_fingerprint_cat64_kat():
    for a, b, v in KAT["fingerprint_cat64"]:
        assert H.fingerprint_cat64(H.fingerprint64(a), H.fingerprint64(b)) == int(v)
        assert H.py_fingerprint_cat64(H.fingerprint64(a), H.fingerprint64(b)) == int(v)


def test_hash_bucket_kat():
    for s, v in KAT["hash_bucket_10"].items():
        assert H.fingerprint64(s) % 10 == v, s


def test_cross_chain_kat():
    k = KAT["cross_chain"]
    h = H.HASH_KEY
    for s in k["keys"]:
        h = H.fingerprint_cat64(h, H.fingerprint64(s))
    assert h == int(k["raw"])
    assert h % (2 ** 63 - 1) == int(k["mod_int64_max"])
    assert h % 100 == k["mod_100"]
    cols = [(np.array([0, 1]), np.array([H.fingerprint64(s)], dtype=np.uint64)) for s in k["keys"]]
    offs, ids = H.cross_rows(cols, 100)
    assert list(ids) == [83] and list(offs) == [0, 1]


def test_crossed_bucketized_string_kat():
    """Pins: integer keys enter the chain raw, chain order = key order, product order = last key innermost."""
    k = KAT["crossed_bucketized_string"]
    ints = (np.array([0, 2, 4]), np.array(k["row0"]["ints"] + k["row1"]["ints"], dtype=np.uint64))
    strs = (np.array([0, 1, 3]), H.fingerprint64_tokens(k["row0"]["strings"] + k["row1"]["strings"]))
    offs, ids = H.cross_rows([ints, strs], k["buckets"], hash_key=k["hash_key"])
    assert list(ids[offs[0]:offs[1]]) == k["row0"]["ids"]
    assert list(ids[offs[1]:offs[2]]) == k["row1"]["ids"]


def test_bucketize_left_closed():
    assert list(H.bucketize(np.array([-1, 0.5, 1.0, 0.0], dtype=np.float32), [0.0, 1.0])) == [0, 1, 2, 1]


def test_c_and_python_restatements_agree_on_all_length_classes():
    rnd = random.Random(7)
    for n in list(range(0, 200)) + [255, 256, 257, 1000, 4097]:
        s = bytes(rnd.randrange(256) for _ in range(n))
        assert H.fingerprint64(s) == H.py_fingerprint64(s), n


def test_fixture_row_ids():
    from oracle import model as OM, tsv
    from wide_deep_b200.config import Config
    cfg = Config()
    fc, cc = cfg.read_feature_conf(), cfg.read_cross_feature_conf()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = open(os.path.join(root, "data", "test", "test2")).readlines()[:1]
    raw, lab = tsv.parse_lines(lines, cfg.read_schema(), fc)
    om = OM.OracleModel(fc, cc, cfg.model, "wide")
    ids = om.transform(raw)
    for name, exp in KAT["fixture_test2_row0"].items():
        if name.startswith("_"):
            continue
        assert list(ids[name][1]) == exp, name
    assert lab[0] == 0.0
