"""GPU, >= 2 devices: the element-partitioned NCCL path (interface exchange, all-reduced Krylov dots) against the
single-GPU path and the oracle.  Skipped on a 1-GPU box (the host-side logic is covered on CPU by
tests/test_partition_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_partitioned_path(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    port = 29400 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0 and "MGPU_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
