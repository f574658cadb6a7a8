"""The N > 1 paths on real GPUs: one process per GPU under torch.distributed.run, NCCL.
Needs at least two GPUs on the box (`gpurun --gpus 2`); on a one-GPU box the test is skipped - the host-side
logic of the same paths runs on CPU with gloo in tests/test_dist.py, and the band / shard arithmetic on one GPU
in tests/test_gpu_grid.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks_match_one(gpu):
    import ctypes
    n = ctypes.c_int(0)
    gpu.pdsb_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip("needs >= 2 GPUs (found %d)" % n.value)
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multirank_worker.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith(("PASS", "FAIL"))]
    print("\n".join(lines))
    assert p.returncode == 0, p.stdout[-4000:]
    assert lines and not any(l.startswith("FAIL") for l in lines)
