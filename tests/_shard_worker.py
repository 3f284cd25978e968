"""Worker for tests/test_gpu_sharded.py::test_sharded_ranks_in_separate_processes (launched by torch.distributed.run): every
rank trains its row shard through wd_shard_train_step_slot (CUDA IPC + flag barriers, CUDA-graph replay after two eager
steps) and the global result is compared with the oracle on the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import model as OM
    from tests.helpers import random_raw_batch, to_product_batch
    from tests.test_gpu_parity import small_conf
    from tests.test_parallel_gloo import slice_raw
    from wide_deep_b200.model import WideDeepModel
    from wide_deep_b200.plan import Plan
    from wide_deep_b200.sharded import ShardedTrainer
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    same = bool(os.environ.get("WD_SHARD_SAME_GPU"))
    dev = 0 if same else local
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")                    # plumbing only: the 64-byte IPC handles and the final verdict
    fc, cross, model = small_conf(hidden=(64, 32))
    per = 48
    B = per * world
    om = OM.OracleModel(fc, cross, model, "wide_deep").init(5)
    rng = np.random.default_rng(77)
    for c in om.wide_cols:
        om.params[om.wname(c)][:] = rng.standard_normal(c.num_buckets).astype(np.float32) * 0.1
    plan = Plan(fc, cross, model, "wide_deep", max_batch=per, max_nnz=per * 64, max_keys=per * 64, dense_exchange_max_rows=400,
                shard_world=world, shard_rank=rank, shard_slack=float(world), gemm_engine="ffma")
    pm = WideDeepModel(plan, device=dev)
    for name in pm.tensor_names():
        pm.set_tensor(name, om.params[name])
        slots = om.slots[name]
        if "acc" in slots:
            pm.set_tensor(name, slots["acc"], slot=1)
        if "n" in slots:
            pm.set_tensor(name, slots["n"], slot=1)
            pm.set_tensor(name, slots["z"], slot=2)
    trainer = ShardedTrainer(pm)
    ok = True
    for step in range(6):                              # steps 0-1 eager, 2 captured, 3-5 replayed (one slot)
        raw = random_raw_batch(fc, B, rng)
        label = (rng.random(B) < 0.3).astype(np.float32)
        lo, hi = rank * per, (rank + 1) * per
        loss = trainer.step(to_product_batch(plan, slice_raw(raw, lo, hi), label[lo:hi]))
        t = torch.tensor([loss], dtype=torch.float64)
        dist.all_reduce(t)
        ref, _ = om.train_step(raw, label)
        if abs(t.item() - ref) > 1e-4 * max(abs(ref), 1.0):
            print("LOSS MISMATCH step", step, t.item(), ref, flush=True)
            ok = False
    for name in pm.tensor_names():
        got, exp = trainer.get_tensor(name), om.params[name]
        scale = max(float(np.abs(exp).max()), 1e-3)
        if np.max(np.abs(got - exp)) > 2e-4 * scale:
            print("MISMATCH", name, np.max(np.abs(got - exp)), scale, flush=True)
            ok = False
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if rank == 0:
        print("SHARD_OK" if flag.item() == 0 else "SHARD_FAIL", flush=True)
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
