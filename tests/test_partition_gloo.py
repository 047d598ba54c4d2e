"""CPU, world_size 2 and 3 over gloo: the host-side logic of the multi-GPU path -- element-column partition, slab
slicing, interface exchange-and-add, ownership weights of the global dot products -- with the oracle as the local
operator.  (The NCCL path itself runs in tests/test_multi_gpu.py on the GPU box.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sem_b200.partition import Partition, exchange_add_lines


def _worker(rank, world, port, P, nx, ny, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sem_oracle as so
        Lx, Ly = 1.3, 0.8
        dx, dy = Lx / nx, Ly / ny
        part = Partition(nx, ny, P, rank, world)
        rng = np.random.default_rng(5)
        xg = rng.standard_normal((nx * P + 1) * (ny * P + 1))
        yg = rng.standard_normal(xg.size)
        _, Kg, Gxg, Gyg = so.global_operators(P, nx, ny, dx, dy)
        # local slab operators = element sums of this rank's elements only
        Ml, Kl, Gxl, Gyl = so.global_operators(P, part.m_end - part.m_begin, ny, dx, dy)
        xl = part.local_slice(xg)
        errs = []
        for A_l, A_g in ((Kl, Kg), (Gxl, Gxg), (Gyl, Gyg), (Ml, None)):
            if A_g is None:   # mass: diagonal vector
                loc = torch.from_numpy((A_l * xl).reshape(part.NX_local, part.NY).copy())
                ref = part.local_slice(so.global_operators(P, nx, ny, dx, dy)[0] * xg)
            else:
                loc = torch.from_numpy((A_l @ xl).reshape(part.NX_local, part.NY).copy())
                ref = part.local_slice(A_g @ xg)
            exchange_add_lines(part, loc, dist)
            errs.append(float(np.linalg.norm(loc.numpy().reshape(-1) - ref) / np.linalg.norm(ref)))
        # global dot product: owned lines only, then all-reduce
        m = part.owned_mask()
        a = part.local_slice(xg).reshape(part.NX_local, part.NY)[m]
        b = part.local_slice(yg).reshape(part.NX_local, part.NY)[m]
        t = torch.tensor([float((a * b).sum())], dtype=torch.float64)
        dist.all_reduce(t)
        errs.append(abs(float(t) - float(xg @ yg)) / abs(float(xg @ yg)))
        # slabs reassemble to the global vector
        slabs = [None] * world
        dist.all_gather_object(slabs, np.array(xl))
        errs.append(float(np.abs(Partition.gather(slabs, nx, ny, P) - xg).max()))
        out[rank] = max(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,P,nx,ny", [(2, 4, 6, 3), (3, 3, 7, 2), (2, 8, 2, 2)])
def test_partitioned_apply_matches_global(world, P, nx, ny):
    port = 29500 + (os.getpid() % 500) + world
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, P, nx, ny, out), nprocs=world, join=True)
    assert len(out) == world
    assert max(out.values()) < 1e-13, dict(out)


def test_partition_ranges():
    p = [Partition(10, 4, 3, r, 4) for r in range(4)]
    assert [(q.m_begin, q.m_end) for q in p] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert p[0].line_end == p[1].line_begin == 9
    assert sum(q.owned_mask().sum() for q in p) == 10 * 3 + 1
    with pytest.raises(ValueError):
        Partition(2, 2, 2, 0, 3)
