"""ORACLE (test infrastructure): hashing primitives.

C restatement (wd_oracle_hash.c) bound with ctypes, plus an independent pure-Python restatement
(`py_fingerprint64`, `py_fingerprint_cat64`) following SURVEY.md Appendix B.1; tests require both to
agree with each other and with the TensorFlow-upstream known-answer vectors.
Reference call sites: python/lib/build_estimator.py:86-88 (hash_bucket), :153 (crossed_column).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libwd_oracle.so")
HASH_KEY = 0xDECAFCAFFE  # tf.feature_column.crossed_column default hash_key
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "wd_oracle_hash.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u64, i64, vp = ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p
        L.wdo_fingerprint64.restype = u64
        L.wdo_fingerprint64.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.wdo_fingerprint_cat64.restype = u64
        L.wdo_fingerprint_cat64.argtypes = [u64, u64]
        L.wdo_fingerprint64_batch.restype = None
        L.wdo_fingerprint64_batch.argtypes = [vp, vp, i64, vp]
        L.wdo_cross_rows.restype = i64
        L.wdo_cross_rows.argtypes = [ctypes.c_int, vp, vp, i64, u64, u64, vp, vp]
        L.wdo_bucketize.restype = None
        L.wdo_bucketize.argtypes = [vp, i64, vp, ctypes.c_int, vp]
        _lib = L
    return _lib


def fingerprint64(s):
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    return int(lib().wdo_fingerprint64(b, len(b)))


def fingerprint_cat64(a, b):
    return int(lib().wdo_fingerprint_cat64(a, b))


FP_EMPTY = 0x9AE16A3B2F90404F  # Fingerprint64("")


def fingerprint64_tokens(tokens):
    """list[str] -> uint64 array."""
    enc = [t.encode("utf-8") for t in tokens]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        offs[1:] = np.cumsum([len(e) for e in enc])
    buf = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
    out = np.empty(len(enc), dtype=np.uint64)
    lib().wdo_fingerprint64_batch(buf.ctypes.data, offs.ctypes.data, len(enc), out.ctypes.data)
    return out


def cross_rows(cols, num_buckets, hash_key=HASH_KEY):
    """cols: list of (offsets int64[B+1], values uint64[nnz]) in OP order -> (offsets, ids int64)."""
    B = len(cols[0][0]) - 1
    n = len(cols)
    offs = [np.ascontiguousarray(c[0], dtype=np.int64) for c in cols]
    vals = [np.ascontiguousarray(c[1], dtype=np.uint64) for c in cols]
    po = (ctypes.c_void_p * n)(*[o.ctypes.data for o in offs])
    pv = (ctypes.c_void_p * n)(*[v.ctypes.data for v in vals])
    total = lib().wdo_cross_rows(n, po, pv, B, int(num_buckets), hash_key, None, None)
    out_offs = np.empty(B + 1, dtype=np.int64)
    out_ids = np.empty(total, dtype=np.int64)
    lib().wdo_cross_rows(n, po, pv, B, int(num_buckets), hash_key, out_offs.ctypes.data, out_ids.ctypes.data)
    return out_offs, out_ids


def bucketize(x, boundaries):
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = np.ascontiguousarray(boundaries, dtype=np.float32)
    out = np.empty(x.shape[0], dtype=np.int64)
    lib().wdo_bucketize(x.ctypes.data, x.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data)
    return out


# --------------------------------------------------------------------------- pure-Python restatement
_M = (1 << 64) - 1
_K0, _K1, _K2 = 0xC3A5C85C97CB3127, 0xB492B66FBE98F273, 0x9AE16A3B2F90404F


def _rot(v, s):
    return v if s == 0 else ((v >> s) | (v << (64 - s))) & _M


def _sm(v):
    return v ^ (v >> 47)


def _f64(b, i):
    return int.from_bytes(b[i:i + 8], "little")


def _f32(b, i):
    return int.from_bytes(b[i:i + 4], "little")


def _h16(u, v, m):
    a = ((u ^ v) * m) & _M
    a ^= a >> 47
    b = ((v ^ a) * m) & _M
    b ^= b >> 47
    return (b * m) & _M


def _weak(b, p, a, c):
    w, x, y, z = _f64(b, p), _f64(b, p + 8), _f64(b, p + 16), _f64(b, p + 24)
    a = (a + w) & _M
    c = _rot((c + a + z) & _M, 21)
    t = a
    a = (a + x + y) & _M
    c = (c + _rot(a, 44)) & _M
    return (a + z) & _M, (c + t) & _M


def py_fingerprint64(s):
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    n = len(b)
    if n == 0:
        return _K2
    if n <= 3:
        y = (b[0] + (b[n >> 1] << 8)) & 0xFFFFFFFF
        z = (n + (b[n - 1] << 2)) & 0xFFFFFFFF
        return (_sm(((y * _K2) & _M) ^ ((z * _K0) & _M)) * _K2) & _M
    if n <= 7:
        m = (_K2 + 2 * n) & _M
        return _h16((n + (_f32(b, 0) << 3)) & _M, _f32(b, n - 4), m)
    if n <= 16:
        m = (_K2 + 2 * n) & _M
        a = (_f64(b, 0) + _K2) & _M
        c0 = _f64(b, n - 8)
        c = (_rot(c0, 37) * m + a) & _M
        d = ((_rot(a, 25) + c0) * m) & _M
        return _h16(c, d, m)
    if n <= 32:
        m = (_K2 + 2 * n) & _M
        a = (_f64(b, 0) * _K1) & _M
        bb = _f64(b, 8)
        c = (_f64(b, n - 8) * m) & _M
        d = (_f64(b, n - 16) * _K2) & _M
        return _h16((_rot((a + bb) & _M, 43) + _rot(c, 30) + d) & _M, (a + _rot((bb + _K2) & _M, 18) + c) & _M, m)
    if n <= 64:
        m = (_K2 + 2 * n) & _M
        a = (_f64(b, 0) * _K2) & _M
        bb = _f64(b, 8)
        c = (_f64(b, n - 8) * m) & _M
        d = (_f64(b, n - 16) * _K2) & _M
        y = (_rot((a + bb) & _M, 43) + _rot(c, 30) + d) & _M
        z = _h16(y, (a + _rot((bb + _K2) & _M, 18) + c) & _M, m)
        e = (_f64(b, 16) * m) & _M
        f = _f64(b, 24)
        g = ((y + _f64(b, n - 32)) * m) & _M
        h = ((z + _f64(b, n - 24)) * m) & _M
        return _h16((_rot((e + f) & _M, 43) + _rot(g, 30) + h) & _M, (e + _rot((f + a) & _M, 18) + g) & _M, m)
    x = 81
    y = (81 * _K1 + 113) & _M
    z = (_sm((y * _K2 + 113) & _M) * _K2) & _M
    v0 = v1 = w0 = w1 = 0
    x = (x * _K2 + _f64(b, 0)) & _M
    end = ((n - 1) // 64) * 64
    last64 = end + ((n - 1) & 63) - 63
    p = 0
    while True:
        x = (_rot((x + y + v0 + _f64(b, p + 8)) & _M, 37) * _K1) & _M
        y = (_rot((y + v1 + _f64(b, p + 48)) & _M, 42) * _K1) & _M
        x ^= w1
        y = (y + v0 + _f64(b, p + 40)) & _M
        z = (_rot((z + w0) & _M, 33) * _K1) & _M
        v0, v1 = _weak(b, p, (v1 * _K1) & _M, (x + w0) & _M)
        w0, w1 = _weak(b, p + 32, (z + w1) & _M, (y + _f64(b, p + 16)) & _M)
        z, x = x, z
        p += 64
        if p == end:
            break
    m = (_K1 + ((z & 0xFF) << 1)) & _M
    p = last64
    w0 = (w0 + ((n - 1) & 63)) & _M
    v0 = (v0 + w0) & _M
    w0 = (w0 + v0) & _M
    x = (_rot((x + y + v0 + _f64(b, p + 8)) & _M, 37) * m) & _M
    y = (_rot((y + v1 + _f64(b, p + 48)) & _M, 42) * m) & _M
    x ^= (w1 * 9) & _M
    y = (y + v0 * 9 + _f64(b, p + 40)) & _M
    z = (_rot((z + w0) & _M, 33) * m) & _M
    v0, v1 = _weak(b, p, (v1 * m) & _M, (x + w0) & _M)
    w0, w1 = _weak(b, p + 32, (z + w1) & _M, (y + _f64(b, p + 16)) & _M)
    z, x = x, z
    return _h16((_h16(v0, w0, m) + ((_sm(y) * _K0) & _M) + z) & _M, (_h16(v1, w1, m) + x) & _M, m)


def py_fingerprint_cat64(a, b):
    K = 0xC6A4A7935BD1E995
    r = a ^ K
    r ^= (_sm((b * K) & _M) * K) & _M
    r = (r * K) & _M
    r = (_sm(r) * K) & _M
    return _sm(r)
