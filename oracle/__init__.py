"""ORACLE — test infrastructure only.

CPU restatement of the reference's Wide&Deep train/eval step (Lapis-Hong/wide_deep), used ONLY as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs.  Nothing under ``wide_deep_b200/`` may import this package: the product path is CUDA-only and fails
loudly when its extension is missing.

The reference's arithmetic lives in the un-vendored third-party dependency ``tensorflow`` (reference
requirements.txt:1, pin ">=1.4"; asserted at reference python/train.py:176).  TensorFlow is not installable
here (no wheel, no network) and the reference is Python-2-only, so the reference itself cannot be run to
generate vectors => **parity unpinned against a live reference run**.  What pins this oracle instead:
TensorFlow-upstream known-answer vectors for every integer function on the path
(``tests/golden/hash_kat.json``: Fingerprint64, FingerprintCat64, hash-bucket ids, SparseCross ids incl.
the bucketized x string cross of TF's feature_column_test) and an independent torch-CPU cross-check of
the floating-point math (``tests/test_oracle_dense.py``).

Modules
  hashing  - ctypes wrapper of wd_oracle_hash.c (FarmHash Fingerprint64, FingerprintCat64, SparseCross,
             Bucketize) plus a second, pure-Python restatement used to cross-check the C one.
  columns  - feature/cross conf -> feature-column objects (reference python/lib/build_estimator.py:49-169).
  tsv      - TSV line -> feature dict (reference python/lib/dataset.py:86-195).
  model    - logits / loss / gradients / Adagrad / FTRL (reference python/lib/{linear,dnn,joint}.py).
  metrics  - eval metric dict of the binary head (reference python/lib/joint.py:402-406).
"""
