#!/usr/bin/env python
"""bench.py — CTR examples/sec of the Wide&Deep train step on N B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference ...                      (CPU restatement of the reference step)

Workload (config.workload): BASELINE.json configs[1] / [2] — synthetic Criteo shape, 13 dense + 26 categorical
(Criteo-Kaggle cardinalities, 33.76 M embedding rows x 32), wide = 26 hash columns + 13 bucketized + 8 crosses
@ 1 M buckets, MLP 1024-512-256 (relu, BN affine), Adagrad deep / FTRL wide, 8192 examples per GPU per step
(weak scaling).  One "step" = ids + forward + sum-reduced sigmoid-CE + backward + both optimizers.

JSON keys beyond the base contract:
  value     examples/s with the step's inputs already resident in HBM (a ring of distinct batches, so the
            rows each step touches are not the ones left in L2 by the previous step)
  e2e       the same metric fed from pinned host memory the way estimator.train feeds it: wd_batch_prefetch_slot refills two
            alternating slots on the upload stream (the copy of step i+1 overlaps step i, like dataset.prefetch in the reference),
            wd_train_step_slot runs the step, and every step ends with a device -> host read of its loss — all timed
  roofline  dominant kernel group (the nine MLP GEMM launches) as achieved fp32-equivalent TFLOP/s vs the measured dense-bf16
            tensor peak, timed live with CUDA events on the model stream; `kernels` carries the same for the embedding gather (HBM)
  gemm_engine   bf16x3 by default here (2^-16 products); `parity` re-checks it against the oracle in this run (bar 1e-4) and
            `strict_engine` reports the same step on the fp32-faithful tc3x engine (the library default)
  cpu_baseline  the oracle (CPU restatement of the reference; TensorFlow itself cannot run here) on the host cores
Multi-GPU (config.exchange): one all-reduce for dense gradients + small-table gradient blocks, all-gather + on-device merge of the
large tables' (row, gradient) lists; WD_DP_PROFILE=1 prints the phase timeline of a data-parallel step to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 8192
RING = 8
# arithmetic the MLP GEMMs run in (tables, optimizers, pooling and every reduction are fp32 in all engines)
DTYPE_OF_ENGINE = {"bf16x3": "bf16x3 (fp32 operands split into bf16 hi+lo, 3 tensor-core products, fp32 accumulate; 2^-16)",
                   "tc3x": "tf32x3 (fp32 operands split into tf32 hi+lo, 3 tensor-core products, fp32 accumulate; 2^-21)",
                   "tc1x": "tf32", "ffma": "f32", "auto": "tf32x3 (library default)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


def gemm_traffic_from_profile(engine, batch):
    """DRAM bytes of the GEMM launches of one step, from the newest committed ncu capture that matches (engine, batch):
    profiles/*_gemm_traffic.json = {"engine", "batch", "dram_bytes_per_step", "command", "source"} written by
    tools/ncu_summary.py --traffic from the `ncu --set full` raw CSV.  (None, reason) when no capture matches — never a constant."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json"))):
        try:
            d = json.load(open(path))
        except Exception:
            continue
        if d.get("engine") == engine and int(d.get("batch", -1)) == int(batch):
            best = (float(d["dram_bytes_per_step"]), os.path.relpath(path, ROOT))
    return best if best else (None, "no committed ncu capture for engine=%s batch=%d" % (engine, batch))


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).

    A step is ~1 ms, so `nvidia-smi -lms` (>= 100 ms per sample) would miss short runs: NVML is polled in-process every 2 ms from a
    thread (same counters nvidia-smi reads); nvidia-smi is the fallback when the NVML binding is missing."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop_flag = index, [], None, None, False
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip()]
        if ids and self.index < len(ids) and ids[self.index].strip().isdigit():
            return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = [float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))]
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == str(self._physical_index()):
                self.rows.append(f)

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_min_mhz": min(self.sm) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons), "samples": len(self.sm),
                    "power_w": float(np.median(self.power)) if self.power else None, "source": "nvml, 2 ms poll during both timed regions"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": "nvidia-smi -lms 100"}


def workload(n_gpus, per_gpu_batch):
    from wide_deep_b200 import synthetic
    fc, cross, model, emb = synthetic.criteo_conf()
    n_cat = sum(1 for c in fc.values() if c["type"] == "category")
    n_dense = len(fc) - n_cat
    P = (n_cat * emb + n_dense) * 1024 + 1024 * 512 + 512 * 256 + 256
    return fc, cross, model, emb, n_cat, n_dense, P


def config_dict(n_gpus, per_gpu_batch):
    return {"workload": "synthetic Criteo shape: 13 dense + 26 categorical (Criteo-Kaggle cardinalities, 33.76M rows), emb 32, "
                        "wide 26 hash + 13 bucketized + 8 crosses@1M, MLP 1024-512-256 relu+BN, Adagrad/FTRL; train step",
            "global_batch": per_gpu_batch * n_gpus, "per_gpu_batch": per_gpu_batch,
            "parallelism": "dp%d" % n_gpus if n_gpus > 1 else "single",
            "tables": "replicated", "ids": "uniform",
            "exchange": ("dense all-reduce of MLP/wide-bias gradients + gradient blocks of tables <= %s rows; all-gather + on-device "
                         "re-reduction of (row, gradient) lists for the larger tables" % os.environ.get("WD_DENSE_EXCHANGE_ROWS", "16384"))
            if n_gpus > 1 else "none",
            "l2": "ring of %d distinct resident batches; touched rows per step ~60 MB, tables 8.6 GB >> 126 MB L2" % RING}


# ------------------------------------------------------------------------------------------- reference arm
def oracle_examples_per_sec(batch_rows, steps, warmup, threads, acc=np.float32, budget_s=None):
    """Time the OPTIMISED CPU restatement (oracle/fast.py: torch-CPU matmuls + embedding_bag + sparse row updates, C hashing;
    checked against oracle/model.py by tests/test_oracle_fast.py) on the same workload with every host thread; returns
    (examples/s, seconds per step, steps timed).  torch.distributed.run exports OMP_NUM_THREADS=1: overridden explicitly."""
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    import torch
    torch.set_num_threads(threads)
    from oracle import fast as OF, model as OM
    from wide_deep_b200 import synthetic
    fc, cross, model, emb, n_cat, n_dense, _ = workload(1, batch_rows)
    om = OF.FastCpuModel(OM.OracleModel(fc, cross, model, "wide_deep", embedding_dim_override=emb, acc=acc).init(1), threads=threads)
    cats = [f for f, c in fc.items() if c["type"] == "category"]
    dense_names = [f for f, c in fc.items() if c["type"] == "continuous"]
    times = []
    for s in range(warmup + steps):
        keys, dense, label = synthetic.criteo_batch_arrays(fc, batch_rows, step=s)
        raw = {f: (np.arange(batch_rows + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
        for j, f in enumerate(dense_names):
            raw[f] = np.ascontiguousarray(dense[:, j])
        t0 = time.perf_counter()
        om.train_step(raw, label)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
            if budget_s is not None and sum(times) > budget_s:
                break
    sec = float(np.mean(times))
    return batch_rows / sec, sec, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the same step as the GPU arm: same tables, same GLOBAL batch (N x 8192: the CPU arm is the whole box's host cores, whatever N
    # is), a bounded number of steps
    rows = args.batch * max(1, args.gpus)
    warm = max(1, min(args.warmup, 2))
    v, sec, steps = oracle_examples_per_sec(rows, max(1, args.steps), warm, threads, budget_s=90.0)   # K steps or 90 s of CPU work
    sample = "%d steps of %d examples (same tables/config as the GPU arm at N=%d), %d threads" % (steps, rows, args.gpus, threads)
    out = {"impl": "reference", "metric": "CTR examples/sec (train step)", "value": v, "unit": "examples/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_dict(max(1, args.gpus), args.batch),
           "cpu_baseline": {"value": v, "unit": "examples/s", "cores": threads, "kind": "port", "sample": sample,
                            "note": "optimised CPU restatement of the reference step (oracle/fast.py: torch-CPU sgemm + embedding_bag + "
                                    "sparse row updates, C hashing); TensorFlow 1.x is not installable here"},
           "e2e": {"value": v, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def parity_check(engine, rows=2048):
    """Checker leg (oracle = test infrastructure): max relative logit error of `engine` against the CPU oracle on the benchmark
    model with every table scaled by 1e-3 (same columns, same 845-1024-512-256 towers, same kernels), parameters copied from the
    oracle.  The bar is BASELINE.json's: |gpu - oracle| <= 1e-4 * max(|oracle|, 1)."""
    from oracle import model as OM
    from tests.helpers import copy_params_to_product
    from wide_deep_b200 import synthetic
    from wide_deep_b200.model import Batch, WideDeepModel
    from wide_deep_b200.plan import Plan
    fc, cross, model, emb = synthetic.criteo_conf(scale=1e-3)
    n_cat = sum(1 for c in fc.values() if c["type"] == "category")
    om = OM.OracleModel(fc, cross, model, "wide_deep", embedding_dim_override=emb).init(7)
    plan = Plan(fc, cross, model, "wide_deep", max_batch=rows, embedding_dim_override=emb, gemm_engine=engine,
                max_nnz=rows * (len(fc) + len(cross)), max_keys=rows * n_cat)
    pm = WideDeepModel(plan)
    copy_params_to_product(om, pm)
    keys, dense, label = synthetic.criteo_batch_arrays(fc, rows, step=123)
    cats = [f for f, c in fc.items() if c["type"] == "category"]
    dn = [f for f, c in fc.items() if c["type"] == "continuous"]
    raw = {f: (np.arange(rows + 1, dtype=np.int64), np.ascontiguousarray(keys[:, j])) for j, f in enumerate(cats)}
    for j, f in enumerate(dn):
        raw[f] = np.ascontiguousarray(dense[:, j])
    b = Batch(rows, keys.reshape(-1), None, dense, label)
    logits, _ = pm.forward(b)
    _, cache = om.forward(raw)
    ref = cache["logits"]
    err = np.abs(logits - ref) / np.maximum(np.abs(ref), 1.0)
    loss = pm.train_step(b)
    ref_loss, _ = om.train_step(raw, label)
    del pm
    return {"engine": engine, "max_rel_logit_err": float(err.max()), "rms_rel_logit_err": float(np.sqrt((err ** 2).mean())),
            "rel_loss_err": float(abs(loss - ref_loss) / max(abs(ref_loss), 1.0)), "bar": 1e-4, "pass": bool(err.max() <= 1e-4),
            "sample": "%d examples, benchmark model with tables scaled 1e-3, parameters copied from the oracle" % rows}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="examples per GPU per step")
    ap.add_argument("--engine", default=os.environ.get("WD_GEMM_ENGINE", "bf16x3"),
                    help="MLP GEMM engine: bf16x3 (tcgen05 kind::f16 on bf16 hi/lo copies, 2^-16 products; re-checked against the "
                         "oracle in this run) | tc3x (tcgen05 kind::tf32 3-pass, 2^-21, the library default) | ffma (fp32 CUDA cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from wide_deep_b200 import synthetic
    from wide_deep_b200.model import Batch, WideDeepModel
    from wide_deep_b200.plan import Plan
    B = args.batch
    fc, cross, model_conf, emb, n_cat, n_dense, P = workload(world, B)
    n_cols = n_cat + n_dense + len(cross)
    # data-parallel runs: tables / wide columns of <= 16384 rows (18 of the 26 embedding tables, 31 of the 47 wide columns) are
    # exchanged as a dense gradient block inside the dense all-reduce; only the large tables' touched rows travel as lists
    dense_rows = int(os.environ.get("WD_DENSE_EXCHANGE_ROWS", "16384")) if world > 1 else 0
    plan = Plan(fc, cross, model_conf, "wide_deep", max_batch=B, embedding_dim_override=emb, gemm_engine=args.engine,
                max_nnz=B * n_cols * (world if world > 1 else 1), max_keys=B * n_cat, dense_exchange_max_rows=dense_rows)
    model = WideDeepModel(plan, device=local)
    model.init(seed=0x5EED0005)          # identical replicas on every rank
    trainer = None
    if world > 1:
        from wide_deep_b200.parallel import DataParallelTrainer
        trainer = DataParallelTrainer(model, fixed_rows=plan.exchange_rows(B))

    # distinct batches per (rank, ring slot) in pinned host memory
    host = []
    for s in range(RING):
        keys, dense, label = synthetic.criteo_batch_arrays(fc, B, step=rank * 1000 + s)
        tk = torch.from_numpy(keys.view(np.int64).reshape(-1).copy()).pin_memory()
        td = torch.from_numpy(dense.copy()).pin_memory()
        tl = torch.from_numpy(label.copy()).pin_memory()
        host.append((Batch(B, tk.numpy().view(np.uint64), None, td.numpy(), tl.numpy()), (tk, td, tl)))
    for s in range(RING):
        model.upload_slot(s, host[s][0])
    model.sync()
    stream = torch.cuda.ExternalStream(model.stream(), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        e1.record(stream)
        model.sync()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step_resident(i):
        if trainer:
            trainer.step_slot(i % RING)
        else:
            model.train_step_slot(i % RING, want_loss=False)

    # end to end: every step's inputs come from pinned host memory and its loss goes back to the host.  Two extra slots are
    # refilled alternately with wd_batch_prefetch_slot — the copy of step i+1 runs on the upload stream while step i computes, as
    # tf.data's prefetch does in the reference's input_fn — and each step ends with a device->host read of its loss.
    E2E0 = RING

    def step_e2e(i):
        model.prefetch_slot(E2E0 + (i + 1) % 2, host[(i + 1) % RING][0])
        if trainer:
            trainer.step_slot(E2E0 + i % 2, want_loss=False)
            return model.last_loss()
        return model.train_step_slot(E2E0 + i % 2, want_loss=True)

    launches0 = model.launch_count()
    for i in range(args.warmup):
        step_resident(i)
    model.sync()
    per_step_launches = (model.launch_count() - launches0) // max(args.warmup, 1)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step_resident, args.steps)
    model.prefetch_slot(E2E0, host[0][0])
    for i in range(8):                                   # both e2e slots past their eager steps (graphs captured)
        step_e2e(i)
    ms_e2e = timed(lambda i: step_e2e(i + 8), args.steps)
    clk = clocks.stop() if rank == 0 else None

    if trainer and os.environ.get("WD_DP_PROFILE") and rank == 0:
        for i in range(3):
            prof = trainer.profile_step(i % RING)
        sys.stderr.write("dp phases (ms from step start): %s\n" % json.dumps({k: round(v, 3) for k, v in prof.items()}))
    elif trainer and os.environ.get("WD_DP_PROFILE"):
        for i in range(3):
            trainer.profile_step(i % RING)

    # per-kernel timings (CUDA events between stages on the model stream), a few profiled steps
    phases = {}
    if not trainer:
        model.set_profile(True)
        nprof = 5
        for i in range(nprof):
            model.train_step_slot(i % RING, want_loss=True)
            for k, v in model.last_timings().items():
                phases[k] = phases.get(k, 0.0) + v / nprof
        model.set_profile(False)

    if rank == 0:
        peaks = load_peaks()
        gb = B * world
        value = gb * args.steps / (ms / 1e3)
        e2e = gb * args.steps / (ms_e2e / 1e3)
        out = {"metric": "CTR examples/sec (train step)", "value": value, "unit": "examples/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_OF_ENGINE.get(args.engine, "f32"), "data": "synthetic",
               "config": config_dict(world, B),
               "e2e": {"value": e2e, "unit": "examples/s", "ms_per_step": ms_e2e / args.steps,
                       "h2d_bytes_per_step": host[0][0].h2d_bytes(), "d2h_bytes_per_step": 8,
                       "input": "pinned host batches, wd_batch_prefetch_slot into two alternating slots (copy of step i+1 overlaps step i), loss read every step"},
               "gpu_launches": int(per_step_launches * args.steps), "launches_per_step": int(per_step_launches),
               "clocks": clk, "gemm_engine": args.engine}
        if phases:
            gemm_ms = sum(v for k, v in phases.items() if k.startswith("gemm_"))
            flops = 6.0 * B * P                                   # 2BP forward + 4BP backward
            ach = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
            traffic, traffic_src = gemm_traffic_from_profile(args.engine, B)
            out["roofline"] = {"kernel": "mlp gemm (fwd+dgrad+wgrad)", "bound": "tensor", "achieved": ach, "peak": peaks["bf16_sustained"],
                               "unit": "TFLOP/s", "frac": ach / peaks["bf16_sustained"],
                               # the GEMMs run inside a long step, back to back with the rest of it: the SUSTAINED peak applies; the
                               # fraction against the burst figure (a kernel timed alone) is given beside it
                               "peak_burst": peaks["bf16_burst"], "frac_burst": ach / peaks["bf16_burst"], "peak_applies": "sustained",
                               # dram__bytes_read.sum + dram__bytes_write.sum summed over the GEMM launches of one step, read from the
                               # committed ncu capture of this engine / batch size (profiles/*_gemm_traffic.json); null when there is none
                               "traffic": traffic, "traffic_unit": "bytes per step (all GEMM launches of one step)", "traffic_source": traffic_src,
                               "peak_source": peaks["source"] + " dense bf16.  achieved = algorithmic fp32 FLOPs (6*B*P) / GEMM time; "
                                              "both split engines issue 3 tensor-core products per algorithmic one, so the fp32-"
                                              "equivalent ceiling is 1/3 of the bf16 peak for bf16x3 and 1/6 for tc3x",
                               "tensor_pipe_frac": 3.0 * ach / peaks["bf16_sustained"] * (2.0 if args.engine == "tc3x" else 1.0),
                               "share_of_step": gemm_ms / phases.get("total", 1.0)}
            gather_bytes = B * (n_cat * (4 * emb + 4) + 4 * n_cat + 4 * n_cat * emb)       # SURVEY 8(d) K3 formula
            g_ms = phases.get("emb_fwd", 0.0)
            g_ach = gather_bytes / (g_ms * 1e-3) / 1e9 if g_ms > 0 else 0.0
            out["kernels"] = {"emb_gather_pool_fwd": {"bound": "hbm", "achieved": g_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                      "frac": g_ach / peaks["hbm_gbs"], "algorithmic_bytes": gather_bytes, "ms": g_ms},
                              "phases_ms": {k: round(v, 4) for k, v in phases.items()}}
        if not args.no_cpu_baseline and world == 1 and args.engine != "tc3x":
            # the same step on the fp32-faithful engine (3xTF32, the library default), for reference next to the headline
            plan2 = Plan(fc, cross, model_conf, "wide_deep", max_batch=B, embedding_dim_override=emb, gemm_engine="tc3x",
                         max_nnz=B * n_cols, max_keys=B * n_cat)
            m2 = WideDeepModel(plan2, device=local)
            m2.init(seed=0x5EED0005)
            for s_ in range(RING):
                m2.upload_slot(s_, host[s_][0])
            st2 = torch.cuda.ExternalStream(m2.stream(), device=torch.device("cuda", local))
            for i in range(6):
                m2.train_step_slot(i % RING, want_loss=False)
            m2.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st2)
            for i in range(40):
                m2.train_step_slot(i % RING, want_loss=False)
            e1.record(st2)
            m2.sync()
            out["strict_engine"] = {"gemm_engine": "tc3x", "value": B * 40 / (e0.elapsed_time(e1) / 1e3), "unit": "examples/s",
                                    "ms_per_step": e0.elapsed_time(e1) / 40, "steps": 40}
            del m2
        if not args.no_cpu_baseline and world == 1:
            out["parity"] = parity_check(args.engine)
            threads = os.cpu_count() or 1
            v, sec, nst = oracle_examples_per_sec(B, 10, 2, threads, budget_s=20.0)
            out["cpu_baseline"] = {"value": v, "unit": "examples/s", "cores": threads, "kind": "port",
                                   "sample": "%d steps of %d examples, same tables/config (oracle/fast.py: torch-CPU + C hashing, fp32)" % (nst, B)}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
