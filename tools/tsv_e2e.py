#!/usr/bin/env python
"""End to end FROM TSV BYTES (the product input path, not a bench.py line): lines/s of `estimator.train` over the bundled
training files repeated R times — file bytes -> wd_tsv_parse (C++ worker threads, pinned ring) -> prefetch thread ->
wd_batch_prefetch_slot -> wd_train_step_slot, loss read every step — next to the parser alone and the train step alone.

    python tools/tsv_e2e.py [--repeat 40] [--batch 2048] [--model_type wide_deep]

The bundled configuration (conf/*.yaml: 43 raw fields, 20 string crosses, towers 1024-512-256) is the reference's own; the
synthetic Criteo workload of bench.py has no text form, so this is the only number that includes parsing + hashing."""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wide_deep_b200.config import Config  # noqa: E402
from wide_deep_b200.dataset import input_fn, list_files  # noqa: E402
from wide_deep_b200.estimator import build_custom_estimator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--repeat", type=int, default=40)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--model_type", default="wide_deep")
    args = ap.parse_args()
    cfg = Config()
    run = cfg.runconfig
    run["save_checkpoints_steps"], run["save_checkpoints_secs"] = None, 10 ** 9          # no checkpoint inside the timed pass
    src = b"".join(open(f, "rb").read() for f in list_files(os.path.join(ROOT, "data", "train")))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "train.tsv")
        with open(path, "wb") as fh:
            for _ in range(args.repeat):
                fh.write(src)
        n_lines = src.count(b"\n") * args.repeat
        nbytes = len(src) * args.repeat
        est = build_custom_estimator(os.path.join(tmp, "model"), args.model_type, config=cfg, max_batch=args.batch)
        # host pipeline alone (same input_fn, same batch size, pageable buffers): read + index + shuffle + parse / hash
        t0 = time.time()
        for _ in input_fn(path, None, "train", args.batch, config=cfg, plan=est.plan):      # file image -> line index -> shuffled batches
            pass
        t_parse = time.time() - t0
        # warm-up pass (graph captures), then the timed pass from file bytes
        est.train(input_fn=lambda: input_fn(path, None, "train", args.batch, config=cfg, plan=est.plan, pinned=True), steps=12)
        save, est.save = est.save, (lambda: None)          # the end-of-pass checkpoint (all tables -> npz) is not input-path work
        t0 = time.time()
        est.train(input_fn=lambda: input_fn(path, None, "train", args.batch, config=cfg, plan=est.plan, pinned=True))
        t_e2e = time.time() - t0
        est.save = save
        # the train step alone on one resident batch
        m = est._ensure_model()
        b = next(iter(input_fn(path, None, "train", args.batch, config=cfg, plan=est.plan)))
        m.upload_slot(0, b)
        for _ in range(5):
            m.train_step_slot(0, want_loss=False)
        m.sync()
        t0 = time.time()
        for _ in range(50):
            m.train_step_slot(0, want_loss=False)
        m.sync()
        t_step = (time.time() - t0) / 50
    print({"lines": n_lines, "mbytes": round(nbytes / 1e6, 1), "batch": args.batch,
           "host_pipeline_lines_per_s": round(n_lines / t_parse), "host_pipeline_MB_per_s": round(nbytes / 1e6 / t_parse, 1),
           "e2e_tsv_lines_per_s": round(n_lines / t_e2e), "step_only_lines_per_s": round(args.batch / t_step),
           "note": "e2e = read file + line index + shuffle + parse/hash (C++ worker pool, pinned ring, prefetch thread) + H2D + train step + "
                   "loss readback every step; no checkpoint inside the timed pass"})


if __name__ == "__main__":
    main()
