"""GPU convection-diffusion solver -- same class, constructor and method signatures as the reference's
``Solvers/ConvectionDiffusion_Solver.py`` (cited below as CD:line), so the OpenMDAO component and the couplers that
wrap it run unchanged.  numpy arrays in, numpy arrays out; in between the fields live on the B200 as padded fp64
tensors and every operation is a call into ``libsem_b200.so`` (no CPU fallback, no host matrices).
"""
import ctypes as C
import typing

import numpy as np
import torch

from . import SEM
from . import _lib as L
from .device import SemDevice


class ConvectionDiffusionSolver:
    def __init__(self, L_x: float, L_y: float, Pe: float, P: int, N_ex: int, N_ey: int,
                 T_W: float = None, T_E: float = None, T_S: float = None, T_N: float = None,
                 mtol=1e-7, iprint: list = [], device: int = None, restart: int = None, precond: str = 'auto',
                 partition=None):
        """
        Steady convection-diffusion ``Pe [u, v].grad T = lap T`` on [0,L_x]x[0,L_y] with Dirichlet (value) or
        homogeneous Neumann (None) sides -- arguments as CD:10-35.  Extra, optional: ``device`` (CUDA ordinal),
        ``restart`` (Krylov basis size), ``precond`` ('auto' | 'fdm' | 'jacobi' | 'none'; auto = fast diagonalisation of
        the Laplacian -- of the whole mesh on one GPU, of each rank's slab on a partitioned mesh), ``partition`` = (rank, world): this process
        owns one strip of element columns and takes/returns the matching slab of every global vector.
        """
        self._iprint = iprint
        self._Pe = Pe
        self._mtol = mtol
        self._L_x, self._L_y = L_x, L_y
        self._P, self._N_ex, self._N_ey = P, N_ex, N_ey
        self._dx, self._dy = L_x / N_ex, L_y / N_ey
        self.N = (N_ex * P + 1) * (N_ey * P + 1)
        self._points = None
        self._points_e = None

        self._dev = SemDevice(P, N_ex, N_ey, self._dx, self._dy, device=device, partition=partition)
        self._lib = self._dev.lib
        self._bc = L.sem_cd_bc()
        for k, val in enumerate((T_W, T_E, T_S, T_N)):          # W, E, S, N: later sides win (CD:62-71)
            self._bc.active[k] = 0 if val is None else 1
            self._bc.value[k] = 0.0 if val is None else float(val)
        # linearisation state == self._Sys / self._Jac_T_u / self._Jac_T_v of the reference (CD:56-58)
        d = self._dev
        self._u, self._v = d.zeros(), d.zeros()
        self._gxT, self._gyT = d.zeros(), d.zeros()
        self._have_sys = False
        self._have_jac = False
        self._buf = [d.zeros() for _ in range(4)]
        self._restart = restart
        if precond == 'auto':
            precond = 'fdm'
        self._precond = {'none': 0, 'jacobi': 1, 'fdm': 2}[precond]
        self._work = None
        self.last_iters = 0
        self.last_resnorm = float('nan')

    # ---- host-side mesh data, built on demand (2N and 2 N_e (P+1)^2 doubles: large on the big meshes) -----------------
    @property
    def points(self):
        if self._points is None:
            pts = SEM.global_nodes(self._P, self._N_ex, self._N_ey, self._dx, self._dy)
            part = self._dev.part
            self._points = pts if part is None else np.stack([part.local_slice(pts[0]), part.local_slice(pts[1])])
        return self._points

    @property
    def points_e(self):
        if self._points_e is None:
            self._points_e = SEM.element_nodes(self._P, self._N_ex, self._N_ey, self._dx, self._dy)
        return self._points_e

    # ---- internals -------------------------------------------------------------------------------------------------------
    def _state(self, with_jac=True):
        st = L.sem_cd_state()
        st.bc = self._bc
        st.Pe = float(self._Pe)
        st.u, st.v = self._u.data_ptr(), self._v.data_ptr()
        st.gxT = self._gxT.data_ptr() if (with_jac and self._have_jac) else None
        st.gyT = self._gyT.data_ptr() if (with_jac and self._have_jac) else None
        return st

    def _krylov(self):
        if self._restart is None:
            free = torch.cuda.mem_get_info(self._dev.tdev)[0]
            cap = max(20, int(0.25 * free / (8 * self._dev.vec_len)) - 3)
            self._restart = int(min(1000, max(30, int(0.3 * self.N)), cap))
        need = self._lib.sem_cd_work_len(self._dev.ctx, self._restart)
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.float64, device=self._dev.tdev)
        if self._precond == 2 and not self._dev.has_fdm:
            self._dev.setup_fdm([self._bc.active[k] for k in range(4)])
        kr = L.sem_krylov()
        kr.atol = float(self._mtol * np.sqrt(self.N))            # CD:147
        kr.restart = self._restart
        kr.max_iters = max(1000, 20 * self._restart)
        kr.precond = self._precond
        kr.verbose = 2 if 'LGMRES_iter' in self._iprint else 0
        return kr

    def _solve_dev(self, rhs, x):
        """Solve J x = rhs on the device (x holds the guess); raises like CD:149-150 when not converged."""
        if not self._have_sys:
            raise RuntimeError('ConvectionDiffusion: _get_residuals must be called before a linear solve')
        kr = self._krylov()
        st = self._state(with_jac=False)
        code = L.check(self._lib.sem_cd_solve(self._dev.ctx, C.byref(st), rhs.data_ptr(), x.data_ptr(), C.byref(kr),
                                              self._work.data_ptr(), self._work.numel(), self._dev.stream),
                       "sem_cd_solve")
        self.last_iters, self.last_resnorm = kr.iters, kr.resnorm
        if code != 0:
            raise RuntimeError(f'ConvectionDiffusion LGMRES: Failed to converge in {kr.iters} iterations')   # text of CD:150
        if 'LGMRES_suc' in self._iprint:                                                                     # text of CD:152-154
            print(f'ConvectionDiffusion LGMRES: Converged in {kr.iters} evaluations with max-norm {self._residual_maxnorm(rhs, x)}')
        return x

    def _residual_maxnorm(self, rhs, x):
        """max-norm of J x - rhs, the number the reference prints on success (CD:153); only evaluated for 'LGMRES_suc'."""
        d = self._dev
        st = self._state(with_jac=False)
        r = d.zeros()
        L.check(self._lib.sem_cd_jvp(d.ctx, C.byref(st), x.data_ptr(), None, None, r.data_ptr(), d.stream), "sem_cd_jvp")
        return float((r - rhs).abs().max())

    # ---- reference API -----------------------------------------------------------------------------------------------------
    def _get_residuals(self, T: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
        """res = (K + Pe(diag(u)G_x + diag(v)G_y)) T with Dirichlet rows T - T_dir  (CD:73-92); caches (u, v)."""
        d = self._dev
        Td = d.to_device(T, self._buf[0])
        d.to_device(u, self._u)
        d.to_device(v, self._v)
        self._have_sys = True
        st = self._state(with_jac=False)
        L.check(self._lib.sem_cd_residual(d.ctx, C.byref(st), Td.data_ptr(), self._buf[1].data_ptr(), d.stream),
                "sem_cd_residual")
        return d.to_host(self._buf[1])

    def _calc_jacobians(self, T: np.ndarray):
        """Jac_T_u = Pe diag(G_x T), Jac_T_v = Pe diag(G_y T)  (CD:94-102), kept as device vectors."""
        d = self._dev
        Td = d.to_device(T, self._buf[0])
        L.check(self._lib.sem_cd_jacobians(d.ctx, float(self._Pe), Td.data_ptr(), self._gxT.data_ptr(),
                                           self._gyT.data_ptr(), d.stream), "sem_cd_jacobians")
        self._have_jac = True

    def _get_dresiduals(self, dT: np.ndarray, du: np.ndarray = None, dv: np.ndarray = None) -> np.ndarray:
        """dres = Sys dT (+ Jac_T_u du + Jac_T_v dv), Dirichlet rows = dT  (CD:104-121)."""
        d = self._dev
        if not self._have_sys:
            raise RuntimeError('ConvectionDiffusion: _get_residuals must be called before _get_dresiduals')
        if (du is not None or dv is not None) and not self._have_jac:
            raise RuntimeError('ConvectionDiffusion: _calc_jacobians must be called before passing du/dv')
        if du is None and dv is None:
            # host vector in, host vector out: upload, apply and download pipelined over segments of element columns
            a, out = d._host(dT), d.host_result()
            st = self._state(with_jac=False)
            L.check(self._lib.sem_cd_jvp_host(d.ctx, C.byref(st), a.ctypes.data, out.ctypes.data,
                                              self._buf[0].data_ptr(), self._buf[1].data_ptr(), d.stream),
                    "sem_cd_jvp_host")
            return out
        dTd = d.to_device(dT, self._buf[0])
        dud = d.to_device(du, self._buf[2]) if du is not None else None
        dvd = d.to_device(dv, self._buf[3]) if dv is not None else None
        st = self._state()
        L.check(self._lib.sem_cd_jvp(d.ctx, C.byref(st), dTd.data_ptr(),
                                     dud.data_ptr() if dud is not None else None,
                                     dvd.data_ptr() if dvd is not None else None,
                                     self._buf[1].data_ptr(), d.stream), "sem_cd_jvp")
        return d.to_host(self._buf[1])

    def _get_update(self, dres: np.ndarray, dT0: np.ndarray = None) -> np.ndarray:
        """Solve J dT = dres by preconditioned GMRES on the device to ||r|| <= mtol sqrt(N)  (CD:123-156)."""
        d = self._dev
        rhs = d.to_device(dres, self._buf[0])
        x = self._buf[1]
        if dT0 is not None:
            d.to_device(dT0, x)
        else:
            x.zero_()
        self._solve_dev(rhs, x)
        return d.to_host(x)

    def _get_solution(self, u: np.ndarray, v: np.ndarray, T0: np.ndarray = None) -> np.ndarray:
        """T = T0 + dT with J dT = -res(T0): one Newton step, the problem is linear  (CD:158-170).  Device resident."""
        d = self._dev
        T = self._buf[2]
        if T0 is not None:
            d.to_device(T0, T)
        else:
            T.zero_()
        d.to_device(u, self._u)
        d.to_device(v, self._v)
        self._have_sys = True
        st = self._state(with_jac=False)
        res, dT = self._buf[0], self._buf[1]
        L.check(self._lib.sem_cd_residual(d.ctx, C.byref(st), T.data_ptr(), res.data_ptr(), d.stream),
                "sem_cd_residual")
        d.axpby(-1.0, res, 0.0, self._buf[3])          # rhs = -res
        dT.zero_()
        self._solve_dev(self._buf[3], dT)
        d.axpby(1.0, dT, 1.0, T)                       # T + dT
        return d.to_host(T)

    def _get_vector(self, f_func: typing.Callable[[np.ndarray, np.ndarray], np.ndarray]) -> np.ndarray:
        """f evaluated at the global nodes  (CD:172-178)."""
        return f_func(self.points[0], self.points[1])

    def _get_interpol(self, f: np.ndarray, points_plot: typing.Tuple[np.ndarray, np.ndarray]) -> np.ndarray:
        """Interpolation of the global vector f at plotting points  (CD:180-188)."""
        d = self._dev
        xs, ys = np.asarray(points_plot[0])[:, 0], np.asarray(points_plot[1])[0, :]     # ij-meshgrid, as SEM.py:262-263
        return d.interpolate(d.to_device(f, self._buf[3]), xs, ys)

    def run(self, u_func, v_func, points_plot):
        """Solution at plotting points  (CD:190-203)."""
        u = self._get_vector(u_func)
        v = self._get_vector(v_func)
        T = self._get_solution(u, v)
        return self._get_interpol(T, points_plot)
