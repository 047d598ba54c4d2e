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
        self.has_ns_schur = False
        self.ns_singular = False
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
        """Values of the SEM interpolant of the device vector ``vec`` on the ij-meshgrid xs x ys (SEM.eval_interpolation,
        SEM.py:248-273) by the tensor-product kernel ``sem_interpolate``: the element of every plot column / row comes from
        ``x2xi`` (SEM.py:23-36) and the Lagrange basis values from ``GLL.standard_evaluation_matrix`` (GLL.py:105-116), like
        the reference.  Works on a partitioned mesh too (every rank gets the whole array)."""
        xs, ys = np.asarray(xs, dtype=np.float64).ravel(), np.asarray(ys, dtype=np.float64).ravel()
        mx, Sx, ny, Sy = self.transfer_tables(xs, ys)
        out = torch.empty((xs.size, ys.size), dtype=torch.float64, device=self.tdev)
        L.check(self.lib.sem_interpolate(self.ctx, vec.data_ptr(), int(xs.size), mx.data_ptr(), Sx.data_ptr(), int(ys.size),
                                         ny.data_ptr(), Sy.data_ptr(), out.data_ptr(), self.stream), "sem_interpolate")
        return out.cpu().numpy()

    def transfer_tables(self, xs, ys):
        """Device tables of the interpolation from THIS mesh onto the tensor grid xs x ys: element column / row of every
        target line (``x2xi``, SEM.py:23-36) and the Lagrange basis values there (GLL.py:105-116): (mx, Sx, ny, Sy)."""
        from . import SEM
        tabs = []
        for pts, h, nel in ((xs, self.dx, self.N_ex), (ys, self.dy, self.N_ey)):
            e, xi = SEM.x2xi(np.asarray(pts, dtype=np.float64).ravel(), h)
            e = np.clip(e, 0, nel - 1)
            S = GLL.standard_evaluation_matrix(self.P, xi)
            tabs += [torch.from_numpy(e.astype(np.int32)).to(self.tdev), torch.from_numpy(np.ascontiguousarray(S)).to(self.tdev)]
        return tuple(tabs)

    # ---- fast-diagonalisation plans (sem_ctx_set_fdm) ----------------------------------------------------------------------
    _EIG_HOST_MAX = 2100     # pencils up to this size are diagonalised on the host (LAPACK), larger ones with torch on the GPU

    def _pencil_1d(self, nel, h):
        """Dense assembled 1-D stiffness matrix, mass vector and weak-gradient matrix of ``nel`` elements of size ``h``
        (the x or y factor of the Kronecker-structured global matrices, SEM.py:170-223), as float64 torch tensors on the
        host (small) or on the device (large)."""
        P, n = self.P, nel * self.P + 1
        dev = torch.device("cpu") if n <= self._EIG_HOST_MAX else self.tdev
        Ks = torch.from_numpy(self._Ks).to(dev) * (2.0 / h)
        Gs = torch.from_numpy(self._w[:, None] * self._D).to(dev)                  # G_s = diag(w) D, GLL.py:62-70
        w = torch.from_numpy(self._w).to(dev) * (0.5 * h)
        rows = (torch.arange(nel, device=dev)[:, None] * P + torch.arange(P + 1, device=dev)[None, :])
        flat = (rows[:, :, None] * n + rows[:, None, :]).reshape(-1)
        K = torch.zeros((n, n), dtype=torch.float64, device=dev)
        K.view(-1).index_add_(0, flat, Ks.expand(nel, P + 1, P + 1).reshape(-1))
        G = torch.zeros((n, n), dtype=torch.float64, device=dev)
        G.view(-1).index_add_(0, flat, Gs.expand(nel, P + 1, P + 1).reshape(-1))
        M = torch.zeros(n, dtype=torch.float64, device=dev)
        M.index_add_(0, rows.reshape(-1), w.expand(nel, P + 1).reshape(-1))
        return K, M, G

    @staticmethod
    def _eigh_scaled(A, m):
        """Generalised eigenpairs of (A, diag(m)): returns (lam, C) with C^T diag(m) C = I; tiny |lam| clamped to 0."""
        s = 1.0 / torch.sqrt(m)
        B = A * s[:, None] * s[None, :]
        lam, V = torch.linalg.eigh(0.5 * (B + B.T))
        lam = torch.where(lam.abs() < 1e-11 * lam.abs().max(), torch.zeros_like(lam), lam)
        return lam, V * s[:, None]

    def _fdm_dir(self, A, M, lo, hi, fold, keep):
        """One direction of a plan: eigenpairs of the pencil (A[lo:hi, lo:hi], diag(M[lo:hi])).  fold: parity-split (the pencil
        must be symmetric about the centre).  Returns a ``sem_fdm_dir`` whose device arrays are appended to ``keep``."""
        cnt = hi - lo
        Aa, Ma = A[lo:hi, lo:hi], M[lo:hi]
        if fold:
            h, ne = cnt // 2, (cnt + 1) // 2
            Fe = torch.zeros((cnt, ne), dtype=torch.float64, device=A.device)
            idx = torch.arange(h, device=A.device)
            Fe[idx, idx] = 1.0
            Fe[cnt - 1 - idx, idx] = 1.0
            if cnt % 2:
                Fe[h, h] = 1.0
            Fo = torch.zeros((cnt, h), dtype=torch.float64, device=A.device)
            Fo[idx, idx] = 1.0
            Fo[cnt - 1 - idx, idx] = -1.0
            le, Ce = self._eigh_scaled(Fe.T @ Aa @ Fe, Fe.T @ Ma)
            if h:
                lo_, Co = self._eigh_scaled(Fo.T @ Aa @ Fo, Fo.abs().T @ Ma)
            else:
                lo_, Co = le[:0], Ce[:0, :0]
            lam = torch.cat([le, lo_])
        else:
            lam, Ce = self._eigh_scaled(Aa, Ma)
            Co = Ce[:0, :0]
        Qe, Qo, lam = (t.to(self.tdev).contiguous() for t in (Ce, Co, lam))
        keep += [Qe, Qo, lam]
        return L.sem_fdm_dir(int(lo), int(cnt), int(bool(fold)), Qe.data_ptr(), Qo.data_ptr() if Qo.numel() else None,
                             lam.data_ptr(), int(cnt))

    def _fdm_dir_global(self, A, M, lo, hi, keep):
        """Unfolded x direction of a PARTITIONED plan: the slab's rows of the global eigenvector matrix."""
        part = self.part
        lam, C = self._eigh_scaled(A[lo:hi, lo:hi], M[lo:hi])
        n = A.shape[0]
        Q = torch.zeros((n, hi - lo), dtype=torch.float64, device=A.device)
        Q[lo:hi] = C
        Qs = Q[part.line_begin:part.line_end + 1].to(self.tdev).contiguous()
        lam = lam.to(self.tdev).contiguous()
        keep += [Qs, lam]
        l0 = max(lo, part.line_begin) - part.line_begin
        l1 = min(hi, part.line_end + 1) - part.line_begin
        return L.sem_fdm_dir(int(l0), int(l1 - l0), 0, Qs.data_ptr(), None, lam.data_ptr(), int(hi - lo))

    def _set_plan(self, slot, Ax, Mx, xr, Ay, My, yr, outside):
        keep = []
        partitioned = self.part is not None and self.part.world > 1
        if partitioned:
            dx_ = self._fdm_dir_global(Ax, Mx, xr[0], xr[1], keep)
        else:
            dx_ = self._fdm_dir(Ax, Mx, xr[0], xr[1], xr[0] + xr[1] == Ax.shape[0], keep)
        dy_ = self._fdm_dir(Ay, My, yr[0], yr[1], yr[0] + yr[1] == Ay.shape[0], keep)
        torch.cuda.current_stream(self.tdev).synchronize()
        L.check(self.lib.sem_ctx_set_fdm(self.ctx, int(slot), C.byref(dx_), C.byref(dy_), int(outside), 0.0), "sem_ctx_set_fdm")

    def setup_fdm(self, dirichlet_wesn):
        """Fast-diagonalisation plan of the Laplacian with the given Dirichlet sides (W, E, S, N) -- slot 0: the preconditioner
        of the CD solve and the velocity block of the NS preconditioner.  Identity rows (z = r) on the Dirichlet nodes."""
        dW, dE, dS, dN = (bool(v) for v in dirichlet_wesn)
        Kx, Mx, _ = self._pencil_1d(self.N_ex, self.dx)
        Ky, My, _ = self._pencil_1d(self.N_ey, self.dy)
        nx, ny = Kx.shape[0], Ky.shape[0]
        self._set_plan(0, Kx, Mx, (int(dW), nx - int(dE)), Ky, My, (int(dS), ny - int(dN)), 1)
        self.has_fdm = True

    def setup_ns_schur(self, pin):
        """Everything the Schur-complement stages of the NS preconditioner need (sem_krylov.precond 3 / 4, DESIGN.md section 4;
        CPU mirror: oracle/ns_precond.py): the all-Neumann pressure Laplacian (slot 1), the structured coarse operator
        Shat = R(sigma) (x) M + M (x) R(sigma) on the interior pressure nodes (slot 2), the tables of the separable coarse-space
        projector, the 1-D factors of the Jacobian's left null vector and the Chebyshev parameters of the ring block."""
        from numpy.polynomial import legendre as npl
        P = self.P
        xi = GLL.standard_nodes(P)[0]
        LP = npl.legval(xi, [0] * P + [1])
        xn = 0.5 * (xi + 1.0)
        wl = np.ascontiguousarray((1.0 - xn) * LP)
        wr = np.ascontiguousarray(xn * LP)
        two_level = P >= 2
        dirs = []
        for nel, h in ((self.N_ex, self.dx), (self.N_ey, self.dy)):
            K, M, G = self._pencil_1d(nel, h)
            n = K.shape[0]
            sign = np.array([(-1.0) ** m if P % 2 else 1.0 for m in range(nel)])
            s = np.zeros(n)
            lf = np.zeros(n)
            for m in range(nel):
                s[m * P:m * P + P + 1] = sign[m] * LP
                lf[m * P:m * P + P + 1] = 1.0 + sign[m] * LP if P % 2 else 1.0 - LP
            Mh = M.cpu().numpy()
            # tridiagonal Gram matrix T = W_I^T M_II W_I of the coarse-space functions (hat_k * s), interior nodes only
            W = np.zeros((n, nel + 1))
            for m in range(nel):
                W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1.0 - xn)
                W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xn)
            W *= s[:, None]
            W[0] = 0.0
            W[-1] = 0.0
            T = W.T @ (Mh[:, None] * W)
            Tinv = np.linalg.inv(T) if two_level else np.zeros_like(T)     # P = 1: no two-level stage (T is singular)
            if two_level:
                # coarse pencil (E^T R(sigma) E, M_II), R(sigma) = G E (K_II + sigma M_II)^-1 E^T G^T, sigma = lambda_s / 4
                st = torch.from_numpy(s).to(K.device)
                lam_s = float((st @ (K @ st)) / (st @ (M * st)))
                X = torch.linalg.solve(K[1:-1, 1:-1] + 0.25 * lam_s * torch.diag(M[1:-1]), G[:, 1:-1].T)
                R = G[:, 1:-1] @ X
                R = 0.5 * (R + R.T)
            else:
                R = K
            glf = (G.T @ torch.from_numpy(lf).to(G.device)).cpu().numpy()       # G1^T lf: enters the velocity part of the null vector
            dirs.append(dict(K=K, M=M, R=R, n=n, lf=lf, Tinv=Tinv, Mh=Mh, glf=glf))
        X_, Y_ = dirs
        self._set_plan(1, X_["K"], X_["M"], (0, X_["n"]), Y_["K"], Y_["M"], (0, Y_["n"]), 0)
        if two_level:
            self._set_plan(2, X_["R"], X_["M"], (1, X_["n"] - 1), Y_["R"], Y_["M"], (1, Y_["n"] - 1), 0)
        NYn = Y_["n"]
        pin_ix, pin_iy = divmod(int(pin), NYn)
        lcx, lcy = X_["lf"], Y_["lf"]
        on_boundary = max(abs(lcx[0]), abs(lcx[-1]), abs(lcy[0]), abs(lcy[-1]))
        singular = bool(on_boundary < 1e-12 and abs(lcx[pin_ix] * lcy[pin_iy]) < 1e-12)
        den = float((X_["Mh"] * lcx * lcx).sum() * (Y_["Mh"] * lcy * lcy).sum())
        arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in
                (wl, wr, X_["Tinv"], Y_["Tinv"], lcx, lcy)]
        from . import SEM
        cheb_lo, cheb_hi, cheb_steps = SEM.ring_chebyshev_parameters(P, self.N_ex, self.N_ey, self.dx, self.dy, int(pin))
        desc = L.sem_ns_schur_desc(*[v.ctypes.data for v in arrs], int(singular), int(two_level),
                                   1.0 / den if den > 0 else 0.0, cheb_lo, cheb_hi, cheb_steps)
        L.check(self.lib.sem_ctx_set_ns_schur(self.ctx, C.byref(desc)), "sem_ctx_set_ns_schur")
        self.has_ns_schur = True
        self.ns_singular = singular
        self._ns_null_1d = (lcx, lcy, X_["glf"], Y_["glf"], X_["Mh"], Y_["Mh"])

    def ns_left_null_vector(self):
        """The left null vector l = (l_u, l_v, l_c) of the singular NS Jacobian as three device vectors, or None when the
        Jacobian is regular.  It is separable and independent of the linearisation point: l_c = lfx (x) lfy,
        l_u = -(G1x^T lfx) (x) (M1y lfy), l_v = -(M1x lfx) (x) (G1y^T lfy) (non-zero on the W/E resp. S/N walls only)
        -- pinned against the sparse-LU null vector of the CPU restatement in tests/test_oracle.py."""
        if not self.has_ns_schur:
            raise L.SemError("ns_left_null_vector needs setup_ns_schur")
        if not self.ns_singular:
            return None
        lfx, lfy, gx, gy, Mx, My = self._ns_null_1d
        NXg, NY = lfx.size, lfy.size
        bnd = np.zeros((NXg, NY), dtype=bool)
        bnd[0], bnd[-1], bnd[:, 0], bnd[:, -1] = True, True, True, True
        fields = (np.where(bnd, -np.outer(gx, My * lfy), 0.0), np.where(bnd, -np.outer(Mx * lfx, gy), 0.0), np.outer(lfx, lfy))
        out = self.zeros(3)
        for k, f in enumerate(fields):
            f = f.ravel() if self.part is None else self.part.local_slice(f.ravel())
            self.to_device(f, out[k])
        return out

    def fdm_apply(self, slot, r, z, nf=1):
        L.check(self.lib.sem_fdm_apply(self.ctx, int(slot), r.data_ptr(), z.data_ptr(), int(nf), self.stream), "sem_fdm_apply")
        return z

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
