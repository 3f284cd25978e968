#!/usr/bin/env python
"""Generates tests/golden/cityhash_le32_kat.json: known-answer vectors for strings of 0..32 bytes from an INDEPENDENT, third-party
compiled implementation — Abseil's `absl::hash_internal::CityHash64` (CityHash v1.1) as exported by pyarrow's libarrow_compute.so
in this image.  farmhashna::Hash64 (= FarmHash Fingerprint64, what TensorFlow's string_to_hash_bucket_fast / crossed_column use)
shares CityHash v1.1's code for the length classes 0-16 and 17-32 (HashLen0to16, HashLen17to32), so these vectors pin those two
classes of the oracle and of the CUDA kernel; FarmHash has its own HashLen33to64 and >64-byte loop, which Abseil cannot pin
(verified: the two functions disagree from 33 bytes on).  Run here (needs pyarrow); the JSON is the committed fixture."""
import ctypes
import glob
import json
import os
import random

import pyarrow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def abseil_cityhash64():
    for so in sorted(glob.glob(os.path.join(os.path.dirname(pyarrow.__file__), "libarrow_compute.so*"))):
        lib = ctypes.CDLL(so)
        for sym in ("_ZN4absl12lts_2025081413hash_internal10CityHash64EPKcm",):
            try:
                f = getattr(lib, sym)
            except AttributeError:
                continue
            f.restype, f.argtypes = ctypes.c_uint64, [ctypes.c_char_p, ctypes.c_size_t]
            return f, os.path.basename(so)
    raise SystemExit("no Abseil CityHash64 symbol found in pyarrow")


def main():
    f, so = abseil_cityhash64()
    rnd = random.Random(20260924)
    vecs = []
    for n in range(0, 33):
        for rep in range(3):
            s = bytes(rnd.randrange(256) for _ in range(n)) if rep else bytes((65 + (i * 7 + n) % 58) for i in range(n))
            vecs.append({"hex": s.hex(), "len": n, "hash": str(f(s, n))})
    out = {"_comment": "Abseil CityHash64 (v1.1) of byte strings with 0..32 bytes, computed by pyarrow's %s; identical to FarmHash "
                       "Fingerprint64 for these length classes (tools/make_cityhash_kat.py)" % so,
           "vectors": vecs}
    path = os.path.join(ROOT, "tests", "golden", "cityhash_le32_kat.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote %d vectors to %s" % (len(vecs), path))


if __name__ == "__main__":
    main()
