"""Host-side Fingerprint64 / FingerprintCat64 from libwd_b200 (same source as the device kernels)."""
from . import _native


def fingerprint64(s):
    b = s.encode("utf-8") if isinstance(s, str) else bytes(s)
    return int(_native.lib().wd_fingerprint64(b, len(b)))


def fingerprint_cat64(a, b):
    return int(_native.lib().wd_fingerprint_cat64(a, b))


FP_EMPTY = 0x9AE16A3B2F90404F
