"""Worker for tests/test_gpu_multi.py (launched by torch.distributed.run): every rank trains its row shard with the
data-parallel trainer; rank 0 also trains a single-GPU model on the concatenated batch and compares parameters."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import model as OM
    from tests.helpers import copy_params_to_product, random_raw_batch, to_product_batch
    from tests.test_gpu_parity import small_conf
    from tests.test_parallel_gloo import slice_raw
    from wide_deep_b200.model import WideDeepModel
    from wide_deep_b200.parallel import DataParallelTrainer, shard_rows
    from wide_deep_b200.plan import Plan
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fc, cross, model = small_conf(hidden=(64, 32))
    B = 96 * world
    om = OM.OracleModel(fc, cross, model, "wide_deep").init(5)
    # WD_DP_DENSE: tables / wide columns of <= 1000 rows travel as a dense block inside the dense all-reduce, the rest as lists
    dense_rows = 1000 if os.environ.get("WD_DP_DENSE") else 0
    plan = Plan(fc, cross, model, "wide_deep", max_batch=B, max_nnz=B * 64 * world, max_keys=B * 64, dense_exchange_max_rows=dense_rows)
    plan_list = Plan(fc, cross, model, "wide_deep", max_batch=B, max_nnz=B * 64 * world, max_keys=B * 64)
    pm = WideDeepModel(plan, device=local)
    copy_params_to_product(om, pm)
    fixed = (96 * 64, 96 * 64) if os.environ.get("WD_DP_FIXED") else None
    trainer = DataParallelTrainer(pm, fixed_rows=fixed)
    single = None
    if rank == 0:
        single = WideDeepModel(plan_list, device=local)     # reference: one GPU, whole batch, plain list path
        copy_params_to_product(om, single)
    use_slot = bool(os.environ.get("WD_DP_SLOT"))           # resident-slot steps: forward+backward replayed from a CUDA graph
    rng = np.random.default_rng(77)
    for step in range(6 if use_slot else 3):
        raw = random_raw_batch(fc, B, rng)
        label = (rng.random(B) < 0.3).astype(np.float32)
        lo, hi = shard_rows(B, rank, world)
        shard = to_product_batch(plan, slice_raw(raw, lo, hi), label[lo:hi])
        if use_slot:
            pm.upload_slot(0, shard)
            trainer.step_slot(0)
        else:
            trainer.step(shard)
        if single is not None:
            single.train_step(to_product_batch(plan_list, raw, label))
    pm.sync()
    ok = True
    if rank == 0:
        for name in pm.tensor_names():
            a, b = pm.get_tensor(name), single.get_tensor(name)
            scale = max(float(np.abs(b).max()), 1e-3)
            if np.max(np.abs(a - b)) > 2e-5 * scale:
                print("MISMATCH", name, np.max(np.abs(a - b)), scale)
                ok = False
    # replicas must stay identical across ranks
    for name in pm.tensor_names()[:6]:
        t = torch.from_numpy(pm.get_tensor(name)).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(t, ref):
            print("REPLICA DIVERGED", name, rank)
            ok = False
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    if rank == 0:
        print("DP_OK" if flag.item() == 0 else "DP_FAIL")
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
