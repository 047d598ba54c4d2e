"""GPU Navier-Stokes solver -- same class, constructor and method signatures as the reference's
``Solvers/NavierStokes_Solver.py`` (cited below as NS:line).  numpy in / numpy out; fields, the linearisation state
and the Krylov basis live on the B200; every operation is a call into ``libsem_b200.so`` (no CPU fallback).

Linear solve.  The reference factorises the velocity block with SuperLU and runs LGMRES on the pressure Schur
complement with a diagonal-mass right preconditioner (NS:162-236).  Here the whole 3-field system is solved by
right-preconditioned GMRES with the block lower-triangular preconditioner [[P_a, 0], [C, M_p]] (P_a = the exact inverse
of the Laplacian by fast diagonalisation, or Jacobi, on the velocity block, C = continuity rows, M_p = the reference's mass preconditioner).  The linearised system is singular
but consistent on most meshes (equal-order spaces); this structure makes GMRES pick the same member of
the solution family as the reference's Schur iteration, so pressures agree too.
"""
import ctypes as C
import typing

import numpy as np
import torch

from . import SEM
from . import _lib as L
from .device import SemDevice


class NavierStokesSolver:
    def __init__(self, L_x: float, L_y: float, Re: float, Gr: float, P: int, N_ex: int, N_ey: int,
                 v_W: float = 0, v_E: float = 0, u_S: float = 0, u_N: float = 0,
                 mtol=1e-7, mtol_newton=1e-5, iprint: list = ['NEWTON_suc', 'NEWTON_iter'],
                 device: int = None, restart: int = None, max_newton: int = 50, partition=None, precond: str = 'auto'):
        """Arguments as NS:11-41.  Extra, optional: ``device``, ``restart`` (Krylov basis size), ``max_newton``,
        ``precond`` ('auto' | 'full' | 'fdm+bb' | 'fdm' | 'jacobi'): 'jacobi' / 'fdm' = Jacobi / fast-diagonalisation velocity
        block with the reference's diagonal-mass Schur preconditioner (NS:208-212); 'fdm+bb' adds the block elimination of the
        pressure-Neumann boundary rows; 'full' (= 'auto') the two-level Schur preconditioner with a pressure convection-
        diffusion stage on top."""
        self._iprint = iprint
        self._Re = Re
        self._Gr = Gr
        if self._Re == 0 and self._Gr != 0:
            raise ValueError('Cannot have Re == 0 and Gr != 0')                     # NS:46-47
        self._Gr_over_Re = self._Gr / self._Re if self._Re != 0 else 0.
        self._mtol = mtol
        self._mtol_newton = mtol_newton
        self._L_x, self._L_y = L_x, L_y
        self._P, self._N_ex, self._N_ey = P, N_ex, N_ey
        self._dx, self._dy = L_x / N_ex, L_y / N_ey
        self.N = (N_ex * P + 1) * (N_ey * P + 1)
        self._points = None
        self._points_e = None
        self._k = 0
        self._max_newton = max_newton

        self._dev = SemDevice(P, N_ex, N_ey, self._dx, self._dy, device=device, partition=partition)
        self._lib = self._dev.lib
        self._bc = L.sem_ns_bc(float(v_W), float(v_E), float(u_S), float(u_N))
        d = self._dev
        self._uv = d.zeros(2)        # advecting velocity of the last _get_residuals  (self._Sys of NS:106)
        self._jac = d.zeros(4)       # Re G_x u, Re G_y u, Re G_x v, Re G_y v          (self._Jac_* of NS:131-136)
        self._have_sys = False
        self._have_jac = False
        self._in = d.zeros(4)        # staging of (u, v, p, T) style inputs
        self._out = d.zeros(3)
        self._x = d.zeros(3)
        self._restart = restart
        if precond == 'auto':
            # the reference's own meshes (a few thousand nodes) are launch-latency bound: the ~70 extra launches of the two-level
            # Schur stages cost more than the iterations they save there (C2: 0.79 s / 3969 its with 'fdm', 1.21 s / 1634 its with
            # 'full'); from ~10^5 nodes on the mesh-independent iteration count wins (profiles/README.md)
            precond = 'full' if self.N >= 30000 and P >= 2 else 'fdm'
        self._precond = {'jacobi': 1, 'fdm': 2, 'fdm+bb': 3, 'full': 4}[precond]
        self._work = None
        self._null = None            # left null vector of the Jacobian (coupled drivers: _get_update_inexact)
        self._null_ready = False
        self.last_iters = 0
        self.last_resnorm = float('nan')
        self.krylov_iters = []       # iterations of every linear solve since construction

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
    def _state(self):
        st = L.sem_ns_state()
        st.bc = self._bc
        st.Re, st.Gr_over_Re = float(self._Re), float(self._Gr_over_Re)
        st.u, st.v = self._uv[0].data_ptr(), self._uv[1].data_ptr()
        if self._have_jac:
            st.gxu, st.gyu = self._jac[0].data_ptr(), self._jac[1].data_ptr()
            st.gxv, st.gyv = self._jac[2].data_ptr(), self._jac[3].data_ptr()
        return st

    def _residual_dev(self, u, v, p, T, out3):
        """Device residual; also records (u, v) as the linearisation point like NS:103-106."""
        d = self._dev
        self._uv[0].copy_(u)
        self._uv[1].copy_(v)
        self._have_sys = True
        st = self._state()
        L.check(self._lib.sem_ns_residual(d.ctx, C.byref(st), u.data_ptr(), v.data_ptr(), p.data_ptr(), T.data_ptr(),
                                          out3[0].data_ptr(), out3[1].data_ptr(), out3[2].data_ptr(), d.stream),
                "sem_ns_residual")

    def _jacobians_dev(self, u, v):
        d = self._dev
        L.check(self._lib.sem_ns_jacobians(d.ctx, float(self._Re), u.data_ptr(), v.data_ptr(),
                                           self._jac[0].data_ptr(), self._jac[1].data_ptr(),
                                           self._jac[2].data_ptr(), self._jac[3].data_ptr(), d.stream),
                "sem_ns_jacobians")
        self._have_jac = True

    def _krylov(self):
        if self._restart is None:
            free = torch.cuda.mem_get_info(self._dev.tdev)[0]
            cap = max(20, int(0.4 * free / (8 * 3 * self._dev.vec_len)) - 3)
            self._restart = int(min(4000, max(50, self.N), cap))
        need = self._lib.sem_ns_work_len(self._dev.ctx, self._restart)
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.float64, device=self._dev.tdev)
        kr = L.sem_krylov()
        kr.atol = float(self._mtol * np.sqrt(self.N))            # NS:223
        kr.restart = self._restart
        kr.max_iters = max(2000, 10 * self._restart)
        if self._precond >= 2 and not self._dev.has_fdm:
            self._dev.setup_fdm([1, 1, 1, 1])                    # velocity Dirichlet rows on all four sides (NS:78-88)
        if self._precond >= 3 and not self._dev.has_ns_schur:
            self._dev.setup_ns_schur(int(self.N / 2))            # pin node int(N/2), NS:89
        kr.precond = self._precond
        kr.verbose = 2 if 'LGMRES_iter' in self._iprint else 0
        return kr

    def _solve_dev(self, rhs3, x3):
        if not (self._have_sys and self._have_jac):
            raise RuntimeError('NavierStokes: _get_residuals and _calc_jacobians must precede a linear solve')
        kr = self._krylov()
        st = self._state()
        code = L.check(self._lib.sem_ns_solve(self._dev.ctx, C.byref(st), rhs3.data_ptr(), x3.data_ptr(),
                                              C.byref(kr), self._work.data_ptr(), self._work.numel(),
                                              self._dev.stream), "sem_ns_solve")
        self.last_iters, self.last_resnorm = kr.iters, kr.resnorm
        self.krylov_iters.append(kr.iters)
        if code != 0:
            raise RuntimeError(f'NavierStokes LGMRES: Failed to converge in {kr.iters} iterations')   # text of NS:226
        if 'LGMRES_suc' in self._iprint:                                                              # text of NS:228-230
            print(f'NavierStokes LGMRES: Converged in {kr.iters} evaluations with max-norm {self._residual_maxnorm(rhs3, x3)}')
        return x3

    def _residual_maxnorm(self, rhs3, x3):
        """max-norm of J x - rhs over the three fields (the reference prints the max-norm of its Schur residual, NS:229);
        only evaluated for 'LGMRES_suc'."""
        d = self._dev
        st = self._state()
        r = d.zeros(3)
        L.check(self._lib.sem_ns_jvp(d.ctx, C.byref(st), x3[0].data_ptr(), x3[1].data_ptr(), x3[2].data_ptr(), None,
                                     r[0].data_ptr(), r[1].data_ptr(), r[2].data_ptr(), d.stream), "sem_ns_jvp")
        return float((r - rhs3).abs().max())

    def _precond_debug(self, what, level, a, b=None):
        """One application of the preconditioner (what = 0: a = (r_u, r_v, r_c) -> (z_u, z_v, z_p) at ``level``) or of one of
        its Schur stages (1 coarse(a), 2 ring elimination of b for right-hand side a, 3 a - S_0 b, 4 pcd(a)); numpy in / out.
        For the stage-by-stage parity tests against oracle/ns_precond.py.  Not part of the reference API."""
        d = self._dev
        if not (self._have_sys and self._have_jac):
            raise RuntimeError('NavierStokes: _get_residuals and _calc_jacobians must precede the preconditioner')
        self._precond, keep = max(self._precond, level), self._precond
        self._krylov()
        self._precond = keep
        st = self._state()
        if what == 0:
            for k in range(3):
                d.to_device(a[k], self._out[k])
            L.check(self._lib.sem_ns_precond_debug(d.ctx, C.byref(st), 0, int(level), self._out.data_ptr(), None,
                                                   self._x.data_ptr(), d.stream), "sem_ns_precond_debug")
            return tuple(d.to_host(self._x[k]) for k in range(3))
        d.to_device(a, self._in[0])
        if b is not None:
            d.to_device(b, self._in[1])
        L.check(self._lib.sem_ns_precond_debug(d.ctx, C.byref(st), int(what), int(level), self._in[0].data_ptr(),
                                               self._in[1].data_ptr() if b is not None else None,
                                               self._in[2].data_ptr(), d.stream), "sem_ns_precond_debug")
        return d.to_host(self._in[2])

    def _spectral_norm(self, r3):
        """2-norm of the 3 x N residual array as a MATRIX (largest singular value) -- what
        ``np.linalg.norm((res_u, res_v, res_cont), ord=2)`` computes at NS:255 -- from its 3x3 Gram matrix."""
        d = self._dev
        G = np.empty((3, 3))
        for i in range(3):
            for j in range(i, 3):
                G[i, j] = G[j, i] = d.dot(r3[i], r3[j])
        return float(np.sqrt(max(np.linalg.eigvalsh(G)[-1], 0.0)))

    # ---- reference API -----------------------------------------------------------------------------------------------------
    def _get_residuals(self, u, v, p, T):
        """Momentum and continuity residuals with their boundary rows  (NS:93-121)."""
        d = self._dev
        for k, a in enumerate((u, v, p, T)):
            d.to_device(a, self._in[k])
        self._residual_dev(self._in[0], self._in[1], self._in[2], self._in[3], self._out)
        return tuple(d.to_host(self._out[k]) for k in range(3))

    def _calc_jacobians(self, u, v):
        """Jacobian diagonals Re G_x u, Re G_y u, Re G_x v, Re G_y v about (u, v)  (NS:123-136)."""
        d = self._dev
        d.to_device(u, self._in[0])
        d.to_device(v, self._in[1])
        self._jacobians_dev(self._in[0], self._in[1])

    def _get_dresiduals(self, du, dv, dp, dT=None):
        """3-field Jacobian-vector product with boundary rows  (NS:138-160)."""
        d = self._dev
        if not (self._have_sys and self._have_jac):
            raise RuntimeError('NavierStokes: _get_residuals and _calc_jacobians must precede _get_dresiduals')
        for k, a in enumerate((du, dv, dp)):
            d.to_device(a, self._in[k])
        dTd = d.to_device(dT, self._in[3]) if dT is not None else None
        st = self._state()
        L.check(self._lib.sem_ns_jvp(d.ctx, C.byref(st), self._in[0].data_ptr(), self._in[1].data_ptr(),
                                     self._in[2].data_ptr(), dTd.data_ptr() if dTd is not None else None,
                                     self._out[0].data_ptr(), self._out[1].data_ptr(), self._out[2].data_ptr(),
                                     d.stream), "sem_ns_jvp")
        return tuple(d.to_host(self._out[k]) for k in range(3))

    def _get_update(self, dres_u, dres_v, dres_cont, du0=None, dv0=None, dp0=None):
        """Velocity and pressure differentials for given residual differentials  (NS:162-236)."""
        d = self._dev
        for k, a in enumerate((dres_u, dres_v, dres_cont)):
            d.to_device(a, self._out[k])
        for k, a in enumerate((du0, dv0, dp0)):
            if a is not None:
                d.to_device(a, self._x[k])
            else:
                self._x[k].zero_()
        self._solve_dev(self._out, self._x)
        return tuple(d.to_host(self._x[k]) for k in range(3))

    def _get_update_inexact(self, dres_u, dres_v, dres_cont, rtol=1e-10):
        """``_get_update`` for right-hand sides that need not lie in the range of the (singular) Jacobian --
        the block-Jacobi preconditioner of a coupled Newton-Krylov driver hands over arbitrary Krylov vectors.  The right-hand
        side is first projected onto the range, b <- b - l (l.b) / (l.l) with the analytic left null vector
        (``SemDevice.ns_left_null_vector``), which makes the system consistent; the Krylov solver then converges to the same
        member of the solution set as for a Newton step (tolerance ``max(mtol sqrt(N), rtol |rhs|)``).  Not part of the
        reference API."""
        if not (self._have_sys and self._have_jac):
            raise RuntimeError('NavierStokes: _get_residuals and _calc_jacobians must precede a linear solve')
        d = self._dev
        for k, a in enumerate((dres_u, dres_v, dres_cont)):
            d.to_device(a, self._out[k])
        self._x.zero_()
        kr = self._krylov()
        if not d.has_ns_schur:
            d.setup_ns_schur(int(self.N / 2))
        if not self._null_ready:
            self._null = d.ns_left_null_vector()
            self._null_nrm2 = d.dot(self._null, self._null) if self._null is not None else 0.0
            self._null_ready = True
        if self._null is not None:
            d.axpby(-d.dot(self._null, self._out) / self._null_nrm2, self._null, 1.0, self._out)
        rhs_norm = float(np.sqrt(d.dot(self._out, self._out)))
        kr.atol = max(kr.atol, rtol * rhs_norm)
        st = self._state()
        code = L.check(self._lib.sem_ns_solve(d.ctx, C.byref(st), self._out.data_ptr(), self._x.data_ptr(), C.byref(kr),
                                              self._work.data_ptr(), self._work.numel(), d.stream), "sem_ns_solve")
        self.last_iters, self.last_resnorm = kr.iters, kr.resnorm
        self.krylov_iters.append(kr.iters)
        if code != 0:
            raise RuntimeError(f'NavierStokes LGMRES: Failed to converge in {kr.iters} iterations')
        return tuple(d.to_host(self._x[k]) for k in range(3))

    def _get_solution(self, T, u0=None, v0=None, p0=None):
        """Newton iteration, device resident; updates u0, v0, p0 in place like NS:248-267 and returns them."""
        d = self._dev
        u = u0 if u0 is not None else np.zeros(self._dev.N_local)
        v = v0 if v0 is not None else np.zeros(self._dev.N_local)
        p = p0 if p0 is not None else np.zeros(self._dev.N_local)
        state = d.zeros(3)
        for k, a in enumerate((u, v, p)):
            d.to_device(a, state[k])
        Td = d.to_device(T, self._in[3])
        res, rhs = self._out, d.zeros(3)
        self._k = 0
        while True:
            self._residual_dev(state[0], state[1], state[2], Td, res)
            norm = self._spectral_norm(res)
            if 'NEWTON_iter' in self._iprint:
                print(f'NavierStokes NEWTON: {self._k}\t{norm}')
            if norm <= self._mtol_newton * np.sqrt(self.N * 3):                      # NS:258
                if 'NEWTON_suc' in self._iprint:
                    r = np.array([d.to_host(res[k]) for k in range(3)])
                    print(f'NavierStokes NEWTON: Converged in {self._k} iterations'
                          f' with max-norm {np.linalg.norm(r, ord=np.inf)}')
                break
            if self._k >= self._max_newton:
                raise RuntimeError(f'NavierStokes NEWTON: Failed to converge in {self._k} iterations')
            self._jacobians_dev(state[0], state[1])
            d.axpby(-1.0, res, 0.0, rhs)
            self._x.zero_()
            self._solve_dev(rhs, self._x)
            d.axpby(1.0, self._x, 1.0, state)                                        # u += du; v += dv; p += dp
            self._k += 1
        for k, a in enumerate((u, v, p)):
            np.copyto(a, d.to_host(state[k]))
        return u, v, p

    def _get_vector(self, f_func):
        """f evaluated at the global nodes  (NS:272-278)."""
        return f_func(self.points[0], self.points[1])

    def _get_interpol(self, f, points_plot):
        """Interpolation of the global vector f at plotting points  (NS:280-288)."""
        d = self._dev
        xs, ys = np.asarray(points_plot[0])[:, 0], np.asarray(points_plot[1])[0, :]     # ij-meshgrid, as SEM.py:262-263
        return d.interpolate(d.to_device(f, self._in[3]), xs, ys)

    def run(self, T_func, points_plot):
        """Solution at plotting points  (NS:290-303)."""
        T = self._get_vector(T_func)
        u, v, p = self._get_solution(T)
        return self._get_interpol(u, points_plot), self._get_interpol(v, points_plot), \
            self._get_interpol(p, points_plot)
