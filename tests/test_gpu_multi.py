"""Multi-GPU data-parallel step on real GPUs (needs >= 2): G ranks on G row shards == 1 GPU on the whole batch."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fixed,dense,slot", [(False, False, False), (True, False, False), (True, True, False), (True, True, True)])
def test_two_gpu_data_parallel_equals_single_gpu(fixed, dense, slot):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tests", "_dp_worker.py")]
    env = dict(os.environ)
    if fixed:
        env["WD_DP_FIXED"] = "1"          # asynchronous fixed-size exchange (no host-side counts)
    if dense:
        env["WD_DP_DENSE"] = "1"          # small tables exchanged as a dense block inside the dense all-reduce
    if slot:
        env["WD_DP_SLOT"] = "1"           # resident batch slot: forward + backward replayed from a CUDA graph
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
