"""N>1 path on CPU: world_size-2 gloo run of the data-parallel exchange logic (wide_deep_b200/parallel.py) with the
oracle standing in for the device step.  Checks the claim in DESIGN.md §6: G ranks on G row shards == 1 rank on the
concatenated batch (dense gradients SUM-allreduced, sparse (row, grad) lists all-gathered and re-reduced by row)."""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as OM
from tests.helpers import random_raw_batch
from tests.test_oracle_dense import conf


def slice_raw(raw, lo, hi):
    out = {}
    for f, v in raw.items():
        if isinstance(v, tuple):
            offs, fp = v
            out[f] = (offs[lo:hi + 1] - offs[lo], fp[offs[lo]:offs[hi]])
        else:
            out[f] = v[lo:hi]
    return out


def _worker(rank, world, port, path):
    from wide_deep_b200.parallel import exchange_sparse, merge_sparse_host, shard_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fc, cross, model = conf()
    om = OM.OracleModel(fc, cross, model, "wide_deep").init(9)
    rng = np.random.default_rng(21)
    B = 50
    for step in range(2):
        raw = random_raw_batch(fc, B, rng)
        label = (rng.random(B) < 0.4).astype(np.float32)
        lo, hi = shard_rows(B, rank, world)
        _, cache = om.forward(slice_raw(raw, lo, hi))
        grads = om.backward(cache, label[lo:hi])
        merged = {}
        for name in sorted(grads):
            g = grads[name]
            if isinstance(g, tuple):
                rows, gr = g
                gr = np.asarray(gr).reshape(len(rows), -1)
                r_t = torch.from_numpy(rows.astype(np.int32))
                g_t = torch.from_numpy(gr.astype(np.float32))
                all_r, all_g = exchange_sparse(r_t, g_t, len(rows))
                u, s = merge_sparse_host(all_r, all_g)
                merged[name] = (u.numpy(), s.numpy().astype(np.float64))
            else:
                t = torch.from_numpy(np.asarray(g, dtype=np.float64).copy())
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                merged[name] = t.numpy()
        om.apply(merged)
    if rank == 0:
        np.savez(path, **om.params)
    dist.destroy_process_group()


def test_two_rank_exchange_equals_single_rank():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    path = os.path.join(tempfile.mkdtemp(), "params.npz")
    mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
    fc, cross, model = conf()
    ref = OM.OracleModel(fc, cross, model, "wide_deep").init(9)
    rng = np.random.default_rng(21)
    for step in range(2):
        raw = random_raw_batch(fc, 50, rng)
        label = (rng.random(50) < 0.4).astype(np.float32)
        ref.train_step(raw, label)
    got = np.load(path)
    for k, v in ref.params.items():
        np.testing.assert_allclose(got[k], v, rtol=2e-5, atol=2e-6, err_msg=k)


def test_merge_skips_invalid_rows():
    from wide_deep_b200.parallel import INVALID_ROW, merge_sparse_host
    rows = torch.tensor([5, INVALID_ROW, 2, 5, INVALID_ROW], dtype=torch.int32)
    grads = torch.tensor([[1.0], [9.0], [2.0], [3.0], [9.0]])
    u, s = merge_sparse_host(rows, grads)
    assert u.tolist() == [2, 5] and s[:, 0].tolist() == [2.0, 4.0]


def test_sorted_list_merge_positions_equal_stable_sort():
    """Host restatement of merge_rank_kernel (csrc/sparse.cu): for G sorted, duplicate-free, INVALID-padded lists the merged
    position of an element = its index in its own list + (elements <= it in lower-numbered lists) + (elements < it in
    higher-numbered lists).  Must reproduce a stable sort of the concatenation, and the row-wise sums of merge_sparse_host."""
    from wide_deep_b200.parallel import INVALID_ROW, merge_sparse_host
    INV = 0xFFFFFFFF                                                      # the marker as the device sees it (uint32 maximum)
    rng = np.random.default_rng(11)
    G, K = 4, 64
    rows = np.full((G, K), INV, dtype=np.int64)
    for g in range(G):
        n = int(rng.integers(0, K + 1))                                   # ragged, possibly empty, possibly full lists
        rows[g, :n] = np.sort(rng.choice(200, size=n, replace=False))
    grads = rng.standard_normal((G * K, 3)).astype(np.float32)
    flat = rows.reshape(-1)
    pos = np.full(G * K, -1, dtype=np.int64)
    for e in range(G * K):
        x = flat[e]
        if x == INV:
            continue
        r = e // K
        p = e - r * K
        for q in range(G):
            if q != r:
                p += np.searchsorted(rows[q], x, side="right" if q < r else "left")
        pos[e] = p
    valid = flat != INV
    nv = int(valid.sum())
    assert sorted(pos[valid].tolist()) == list(range(nv))                 # a permutation of [0, n_valid)
    order = np.empty(nv, dtype=np.int64)
    order[pos[valid]] = np.nonzero(valid)[0]
    ref = np.nonzero(valid)[0][np.argsort(flat[valid], kind="stable")]    # stable sort of the concatenation
    np.testing.assert_array_equal(order, ref)
    # unique rows + sums in merged order == the reference merge
    keys = flat[order]
    uniq, start = np.unique(keys, return_index=True)
    sums = np.add.reduceat(grads[order].astype(np.float64), start, axis=0)
    as_i32 = np.where(valid, flat, INVALID_ROW).astype(np.int32)           # what torch sees: int32 with -1 for the marker
    u2, s2 = merge_sparse_host(torch.from_numpy(as_i32), torch.from_numpy(grads))
    np.testing.assert_array_equal(uniq, u2.numpy())
    np.testing.assert_allclose(sums, s2.numpy().astype(np.float64), rtol=0, atol=1e-5)
