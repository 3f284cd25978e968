"""The C-ABI library loads without a GPU and exports every symbol include/wd_b200.h declares."""
import os
import re

import pytest


def header_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "include", "wd_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wd_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(native_lib):
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(native_lib, s), "libwd_b200.so does not export %s" % s


def test_python_binding_covers_header():
    from wide_deep_b200 import _native
    assert sorted(_native.SYMBOLS) == header_symbols()


def test_host_hash_matches_oracle(native_lib):
    from oracle import hashing as H
    from wide_deep_b200 import _hashing_host as HH
    import random
    rnd = random.Random(3)
    for n in list(range(0, 150)) + [300, 1000]:
        s = bytes(rnd.randrange(256) for _ in range(n))
        assert HH.fingerprint64(s) == H.fingerprint64(s), n
    assert HH.fingerprint_cat64(1, 2) == H.fingerprint_cat64(1, 2)


def test_no_cpu_fallback(native_lib):
    """Without a CUDA device the product refuses to create a model (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wide_deep_b200 import _native
    from wide_deep_b200.config import Config
    from wide_deep_b200.model import WideDeepModel
    from wide_deep_b200.plan import compile_plan
    with pytest.raises(_native.NativeError) as e:
        WideDeepModel(compile_plan(Config()))
    assert e.value.code == _native.ENODEVICE


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "wide_deep_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
