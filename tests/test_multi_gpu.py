"""GPU, >= 2 devices: the element-partitioned NCCL path (interface exchange, all-reduced Krylov dots, distributed fast
diagonalisation, partitioned NS preconditioner) against the single-GPU path and the oracle.  Skipped when the box has fewer
GPUs than ranks (the host-side logic is covered on CPU by tests/test_partition_gloo.py).  world = 3 exercises ranks with two
neighbours and slabs of unequal width; the two environment variants cover the NCCL send/recv fallback of the interface
exchange, the eager (no CUDA graph) path and the three-launch + exchange-kernel path that the one-launch apply (exchange inside
the operator kernel) replaced as the default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,env", [(2, {}), (2, {"SEM_B200_NO_P2P": "1"}), (2, {"SEM_B200_NO_GRAPH": "1"}), (2, {"SEM_B200_FUSED_XCH": "0"}),
                                       (3, {}), (4, {}),
                                       (8, {})])
def test_partitioned_path(world, env):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    port = 29400 + (os.getpid() + 7 * world + len(env)) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "mgpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=root, env={**os.environ, **env})
    assert res.returncode == 0 and "MGPU_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
