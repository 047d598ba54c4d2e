"""Device-side mesh context: owns a ``sem_ctx`` (C ABI), allocates padded field vectors as torch tensors and moves
the reference's dense numpy vectors across the host/device boundary.

Layout: the reference stores a field as a 1-D array of length N with node (ix, iy) at ``iy + NY*ix`` (SEM.py:110).
On the device the same field is a ``[NX, LD]`` fp64 tensor with ``LD = round_up(NY, 16)`` (zero pads) so that every
node line starts 128-byte aligned.
"""
import ctypes as C
import os
import weakref

import numpy as np
import torch

from . import GLL
from . import _lib as L


def default_device():
    """CUDA device ordinal of this process: SEM_B200_DEVICE, else LOCAL_RANK (torchrun), else 0."""
    for key in ("SEM_B200_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


# Page-locked result arrays.  Every call hands out a FRESH numpy array (the reference returns fresh arrays too), backed by
# a pinned block so that the D2H copy runs at the full link rate (54 GB/s measured against 21 GB/s into pageable memory);
# a weakref finaliser puts the block back into a small pool once the array and all its views are gone, because
# cudaHostAlloc of a large block costs ~130 ms.
_PINNED_POOL = {}
_PINNED_KEEP = 4


def _pinned_release(n, block):
    pool = _PINNED_POOL.setdefault(n, [])
    if len(pool) < _PINNED_KEEP:
        pool.append(block)


def _pinned_result(n):
    pool = _PINNED_POOL.get(n)
    block = pool.pop() if pool else _alloc_pinned(n)
    arr = block.numpy()                  # views of `arr` have `.base is arr`: it dies only after all of them
    weakref.finalize(arr, _pinned_release, n, block)
    return arr


def _alloc_pinned(n):
    return torch.empty(n, dtype=torch.float64).pin_memory()


class SemDevice:
    def __init__(self, P, N_ex, N_ey, dx, dy, device=None, m_begin=0, m_end=None, partition=None):
        if not torch.cuda.is_available():
            raise L.SemError("sem_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = L.load()
        self.P, self.N_ex, self.N_ey = int(P), int(N_ex), int(N_ey)
        self.dx, self.dy = float(dx), float(dy)
        self.device = default_device() if device is None else int(device)
        self.part = None
        if partition is not None:                       # (rank, world): strip of element columns of this rank
            from .partition import Partition
            self.part = Partition(self.N_ex, self.N_ey, self.P, int(partition[0]), int(partition[1]))
            m_begin, m_end = self.part.m_begin, self.part.m_end
        self.m_begin = int(m_begin)
        self.m_end = self.N_ex if m_end is None else int(m_end)
        self.tdev = torch.device("cuda", self.device)
        torch.cuda.set_device(self.tdev)
        self._D = np.ascontiguousarray(GLL.standard_differentiation_matrix(self.P))
        self._Ks = np.ascontiguousarray(GLL.standard_stiffness_matrix(self.P))
        self._w = np.ascontiguousarray(GLL.standard_nodes(self.P)[1])
        desc = L.sem_mesh_desc(self.P, self.N_ex, self.N_ey, self.dx, self.dy,
                               self._D.ctypes.data, self._Ks.ctypes.data, self._w.ctypes.data,
                               self.device, self.m_begin, self.m_end)
        ctx = C.c_void_p()
        L.check(self.lib.sem_ctx_create(C.byref(ctx), C.byref(desc)), "sem_ctx_create")
        self.ctx = ctx
        self.LD = self.lib.sem_ctx_ld(ctx)
        self.NX = self.lib.sem_ctx_nx(ctx)
        self.NY = self.lib.sem_ctx_ny(ctx)
        self.vec_len = self.lib.sem_ctx_vec_len(ctx)
        self.N_local = self.NX * self.NY
        self._pinned = {}
        self.has_fdm = False
        self.has_pbb = False
        if self.part is not None and self.part.world > 1:
            self._attach_comm()

    def _attach_comm(self):
        """Create the NCCL communicator of the partition: rank 0 draws the unique id, torch.distributed spreads it."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise L.SemError("partition=(rank, world) needs torch.distributed to be initialised (torchrun)")
        idbuf = (C.c_ubyte * 128)()
        if self.part.rank == 0:
            L.check(self.lib.sem_nccl_unique_id(idbuf), "sem_nccl_unique_id")
        t = torch.tensor(list(idbuf), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.to(self.tdev)
        dist.broadcast(t, src=0)
        idbuf = (C.c_ubyte * 128)(*t.cpu().tolist())
        L.check(self.lib.sem_ctx_attach_comm(self.ctx, idbuf, self.part.rank, self.part.world), "sem_ctx_attach_comm")

    @property
    def comm_mode(self):
        """'none' | 'nccl' (send/recv fallback) | 'p2p' (peer-memory mailboxes over NVLink) -- how the interface lines travel."""
        return ("none", "nccl", "p2p")[self.lib.sem_ctx_comm_mode(self.ctx)]

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.sem_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- storage -------------------------------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def zeros(self, k=None):
        shape = (self.NX, self.LD) if k is None else (k, self.NX, self.LD)
        return torch.zeros(shape, dtype=torch.float64, device=self.tdev)

    def set_tiling(self, Ty=0, Mx=0):
        L.check(self.lib.sem_ctx_set_tiling(self.ctx, int(Ty), int(Mx)), "sem_ctx_set_tiling")

    def _host(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.size != self.N_local:
            raise ValueError(f"expected a vector of length {self.N_local}, got {a.size}")
        return a

    def to_device(self, arr, out=None):
        """numpy (N,) -> padded device vector (pads untouched, i.e. zero)."""
        a = self._host(arr)
        if out is None:
            out = self.zeros()
        L.check(self.lib.sem_h2d(self.ctx, a.ctypes.data, out.data_ptr(), self.stream), "sem_h2d")
        # the async copy reads pageable host memory: make sure it is done before `a` can be released
        torch.cuda.current_stream(self.tdev).synchronize()
        return out

    def host_result(self):
        """A fresh page-locked numpy vector of the local length (see ``_pinned_result``)."""
        return _pinned_result(self.N_local)

    def to_host(self, vec, out=None):
        """padded device vector -> fresh numpy (N,) (synchronises)."""
        if out is None:
            out = _pinned_result(self.N_local)
        L.check(self.lib.sem_d2h(self.ctx, vec.data_ptr(), out.ctypes.data, self.stream), "sem_d2h")
        return out

    # ---- tensor-product interpolation (SEM.eval_interpolation, SEM.py:248-273; mesh-to-mesh transfer of the couplers) ------
    def interpolate(self, vec, xs, ys):
        """Values of the SEM interpolant of the device vector ``vec`` on the ij-meshgrid xs x ys: I_x F I_y^T with the 1-D
        interpolation matrices of ``SEM.interp_matrix_1d``, two small dense products on the device.  Whole mesh only."""
        from . import SEM
        if self.part is not None and self.part.world > 1:
            raise L.SemError("interpolation needs the whole mesh on one GPU")
        Ix = torch.from_numpy(SEM.interp_matrix_1d(self.P, self.N_ex, self.dx, xs)).to(self.tdev)
        Iy = torch.from_numpy(SEM.interp_matrix_1d(self.P, self.N_ey, self.dy, ys)).to(self.tdev)
        return ((Ix @ vec[:, :self.NY]) @ Iy.T).cpu().numpy()

    # ---- fast-diagonalisation preconditioner ---------------------------------------------------------------------------
    def _fdm_1d(self, nel, h, dir_lo, dir_hi):
        """Generalised eigenpairs of the assembled 1-D pencil (K1, M1) of ``nel`` elements of size ``h`` (SEM.py:186-203
        restricted to one direction) with the Dirichlet end nodes eliminated.  Returns (Q [n, n] row-major with zero
        rows at eliminated nodes and zero unused columns, lam [n]) on the device: Q^T M1 Q = I, Q^T K1 Q = diag(lam)."""
        P, n = self.P, nel * self.P + 1
        Ks = torch.from_numpy(self._Ks).to(self.tdev) * (2.0 / h)
        w = torch.from_numpy(self._w).to(self.tdev) * (0.5 * h)
        rows = (torch.arange(nel, device=self.tdev)[:, None] * P + torch.arange(P + 1, device=self.tdev)[None, :])
        K = torch.zeros((n, n), dtype=torch.float64, device=self.tdev)
        flat = (rows[:, :, None] * n + rows[:, None, :]).reshape(-1)
        K.view(-1).index_add_(0, flat, Ks.expand(nel, P + 1, P + 1).reshape(-1))
        M = torch.zeros(n, dtype=torch.float64, device=self.tdev)
        M.index_add_(0, rows.reshape(-1), w.expand(nel, P + 1).reshape(-1))
        lo, hi = (1 if dir_lo else 0), (n - 1 if dir_hi else n)
        s = 1.0 / torch.sqrt(M[lo:hi])
        A = K[lo:hi, lo:hi] * s[:, None] * s[None, :]
        lam, V = torch.linalg.eigh(0.5 * (A + A.T))
        Q = torch.zeros((n, n), dtype=torch.float64, device=self.tdev)
        Q[lo:hi, :hi - lo] = V * s[:, None]
        lam_full = torch.ones(n, dtype=torch.float64, device=self.tdev)
        lam_full[:hi - lo] = lam.clamp_min(0.0)
        return Q.contiguous(), lam_full

    def setup_fdm(self, dirichlet_wesn):
        """Build and hand over the fast-diagonalisation preconditioner for the given Dirichlet sides (W, E, S, N).  On a
        partitioned mesh the x transform is distributed (GEMM + reduce-scatter, all-gather + GEMM): still the exact inverse."""
        dW, dE, dS, dN = (bool(v) for v in dirichlet_wesn)
        Qx, lx = self._fdm_1d(self.N_ex, self.dx, dW, dE)        # the GLOBAL x pencil (every rank solves the same small problem)
        if self.part is not None and self.part.world > 1:
            # exact distributed FDM: this rank keeps the rows of Qx that belong to its slab (all modes)
            Qx = Qx[self.part.line_begin:self.part.line_end + 1].contiguous()
        Qy, ly = self._fdm_1d(self.N_ey, self.dy, dS, dN)
        flags = (C.c_int * 4)(int(dW), int(dE), int(dS), int(dN))
        torch.cuda.current_stream(self.tdev).synchronize()
        L.check(self.lib.sem_ctx_set_fdm(self.ctx, Qx.data_ptr(), lx.data_ptr(), Qy.data_ptr(), ly.data_ptr(), flags),
                "sem_ctx_set_fdm")
        self.has_fdm = True

    def setup_pressure_boundary_block(self, pin):
        """EXPERIMENTAL (not validated on a GPU in round 1; used only by ``NavierStokesSolver(precond='fdm+bb')``): dense
        inverse of the stiffness matrix restricted to the boundary pressure nodes, for the block elimination of the
        pressure-Neumann rows in the NS preconditioner.  Whole mesh on one GPU, at most 8192 boundary
        nodes (a banded ring solve is the production answer)."""
        from . import SEM
        if self.part is not None and self.part.world > 1:
            raise L.SemError("the pressure boundary block needs the whole mesh on one GPU")
        ix, iy, KBB = SEM.pressure_boundary_block(self.P, self.N_ex, self.N_ey, self.dx, self.dy, pin)
        if ix.size > 8192:
            raise L.SemError(f"pressure boundary block: {ix.size} boundary nodes exceed the dense limit of 8192")
        inv = torch.linalg.inv(torch.from_numpy(KBB).to(self.tdev)).contiguous()
        idx = np.ascontiguousarray(ix.astype(np.int64) * self.LD + iy.astype(np.int64))
        torch.cuda.current_stream(self.tdev).synchronize()
        L.check(self.lib.sem_ctx_set_pbb(self.ctx, idx.ctypes.data, int(idx.size), inv.data_ptr()), "sem_ctx_set_pbb")
        self.has_pbb = True

    # ---- single operators ------------------------------------------------------------------------------------------
    def apply_stiffness(self, x, y):
        L.check(self.lib.sem_apply_stiffness(self.ctx, x.data_ptr(), y.data_ptr(), self.stream), "sem_apply_stiffness")
        return y

    def apply_gradient(self, x, gx, gy, scale=1.0):
        L.check(self.lib.sem_apply_gradient(self.ctx, x.data_ptr(), float(scale),
                                            gx.data_ptr() if gx is not None else None,
                                            gy.data_ptr() if gy is not None else None, self.stream),
                "sem_apply_gradient")
        return gx, gy

    def apply_mass(self, x, y):
        L.check(self.lib.sem_apply_mass(self.ctx, x.data_ptr(), y.data_ptr(), self.stream), "sem_apply_mass")
        return y

    def mass_diag(self, m=None):
        m = self.zeros() if m is None else m
        L.check(self.lib.sem_mass_diag(self.ctx, m.data_ptr(), self.stream), "sem_mass_diag")
        return m

    def gather_scatter(self, elem, y=None):
        """Colour-ordered assembly of a device element array [m, n, i, j] into a padded vector."""
        y = self.zeros() if y is None else y
        L.check(self.lib.sem_gather_scatter(self.ctx, elem.data_ptr(), y.data_ptr(), self.stream), "sem_gather_scatter")
        return y

    def scatter(self, x, elem=None):
        if elem is None:
            elem = torch.empty((self.m_end - self.m_begin, self.N_ey, self.P + 1, self.P + 1), dtype=torch.float64,
                               device=self.tdev)
        L.check(self.lib.sem_scatter(self.ctx, x.data_ptr(), elem.data_ptr(), self.stream), "sem_scatter")
        return elem

    def dot(self, x, y):
        out = C.c_double()
        L.check(self.lib.sem_dot(self.ctx, x.data_ptr(), y.data_ptr(), x.numel(), C.byref(out), self.stream), "sem_dot")
        return out.value

    def axpby(self, a, x, b, y):
        """y = a*x + b*y on the device (own kernel)."""
        L.check(self.lib.sem_axpby(self.ctx, float(a), x.data_ptr(), float(b), y.data_ptr(), x.numel(), self.stream),
                "sem_axpby")
        return y
