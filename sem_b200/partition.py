"""Element-column partition of the mesh across ranks (one process per GPU).

Rank r owns the contiguous element columns [m_begin, m_end); its slab of a global field is the contiguous line range
[m_begin*P, m_end*P] (inclusive) of the x-slow / y-fast global vector (SEM.py:110).  The interface line between two
neighbouring ranks is duplicated on both; after a local operator apply each copy holds the element sums of its own
rank and the exchange adds the neighbour's (NCCL on the GPU path, see csrc/sem_comm.cu; ``exchange_add_lines`` below is
the same step written with torch.distributed for host tensors, used by the CPU/gloo tests).
"""
import numpy as np


class Partition:
    def __init__(self, N_ex, N_ey, P, rank=0, world=1):
        if world < 1 or not (0 <= rank < world):
            raise ValueError("bad rank / world size")
        if world > N_ex:
            raise ValueError("more ranks than element columns")
        self.N_ex, self.N_ey, self.P, self.rank, self.world = N_ex, N_ey, P, rank, world
        base, rem = divmod(N_ex, world)
        counts = [base + (1 if r < rem else 0) for r in range(world)]
        starts = np.concatenate(([0], np.cumsum(counts)))
        self.m_begin, self.m_end = int(starts[rank]), int(starts[rank + 1])
        self.all_ranges = [(int(starts[r]), int(starts[r + 1])) for r in range(world)]
        self.NY = N_ey * P + 1
        self.NX_global = N_ex * P + 1
        self.line_begin = self.m_begin * P                 # first global node line of the slab
        self.line_end = self.m_end * P                     # last global node line (inclusive)
        self.NX_local = self.line_end - self.line_begin + 1
        self.has_left = rank > 0
        self.has_right = rank < world - 1

    @property
    def N_local(self):
        return self.NX_local * self.NY

    def local_slice(self, global_vec):
        """This rank's slab (a view) of a global vector of length NX_global*NY."""
        g = np.asarray(global_vec)
        if g.size != self.NX_global * self.NY:
            raise ValueError("not a global vector of this mesh")
        return g.reshape(self.NX_global, self.NY)[self.line_begin:self.line_end + 1].reshape(-1)

    def owned_mask(self):
        """Lines of the local slab that this rank counts in global sums (the interface line goes to the left rank)."""
        m = np.ones(self.NX_local, dtype=bool)
        if self.has_left:
            m[0] = False
        return m

    @staticmethod
    def gather(slabs, N_ex, N_ey, P):
        """Reassemble a global vector from the per-rank slabs (interface lines must agree)."""
        world = len(slabs)
        NY = N_ey * P + 1
        out = np.empty(((N_ex * P + 1), NY))
        for r in range(world):
            p = Partition(N_ex, N_ey, P, r, world)
            out[p.line_begin:p.line_end + 1] = np.asarray(slabs[r]).reshape(p.NX_local, NY)
        return out.reshape(-1)


def exchange_add_lines(part, field2d, dist, group=None):
    """Host/any-backend version of the interface exchange: field2d is a torch tensor [NX_local, >=NY] whose first/last
    line hold this rank's partial sums; adds the neighbour's partial (lower rank's term first)."""
    import torch
    ops, bufs = [], {}
    NY = part.NY
    if part.has_left:
        bufs["L"] = torch.empty(NY, dtype=field2d.dtype)
        ops += [dist.P2POp(dist.isend, field2d[0, :NY].contiguous(), part.rank - 1, group),
                dist.P2POp(dist.irecv, bufs["L"], part.rank - 1, group)]
    if part.has_right:
        bufs["R"] = torch.empty(NY, dtype=field2d.dtype)
        ops += [dist.P2POp(dist.isend, field2d[-1, :NY].contiguous(), part.rank + 1, group),
                dist.P2POp(dist.irecv, bufs["R"], part.rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if part.has_left:
        field2d[0, :NY] = bufs["L"] + field2d[0, :NY]
    if part.has_right:
        field2d[-1, :NY] = field2d[-1, :NY] + bufs["R"]
    return field2d
