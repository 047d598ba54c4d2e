"""Coupled Boussinesq driver -- same ``run`` signature and return values as the reference's
``OpenMDAO/Boussinesq_SequentialCoupler.py`` (cited as BSC:line), without OpenMDAO.

The reference wires its two solvers into an OpenMDAO group and lets OpenMDAO's NewtonSolver / ScipyKrylov / LinearBlockJac /
NonlinearBlockGS / ArmijoGoldsteinLS drive them (BSC:66-97; the adapters are ``OpenMDAO/*_Component.py``).  OpenMDAO cannot be
installed here and the author ran a locally patched copy (BSC:75,79), so iteration-by-iteration parity is unpinned (SURVEY
section 8c).  What is well defined is the coupled fixed point, and this module reaches it with the same three strategies, the
same tolerances and the same component semantics, calling the GPU solvers directly:

* ``mode='GS'``  nonlinear block Gauss-Seidel: ``solve_nonlinear`` of the CD component, then of the NS component
  (``CD_Component.py:59-61``, ``NS_Component.py:62-65``), until the coupled residual 2-norm <= mtol_nonlin * sqrt(DOF).
* ``mode='JNK'`` Newton on the coupled residual; every Newton step solves the coupled Jacobian system with restarted GMRES
  (atol = mtol_gmres * sqrt(DOF)), preconditioned by block Jacobi = one ``solve_linear`` per component.
* ``mode='NJ'``  Newton with the block-Jacobi solve as the (inexact) linear solver and an Armijo-Goldstein line search
  (AGi, AGr, AGc as BSC:16).

The mesh-to-mesh transfer between the CD and NS spaces (``change_inputs``, ``CD_Component.py:23-36``) is the tensor-product
interpolation ``I_x F I_y^T`` on the device (``SemDevice.interpolate``); it is skipped when both meshes coincide.

With the GPU solvers the linear solve of ``mode='JNK'`` is device resident (``sem_coupled_solve``): coupled vector, Krylov basis,
operator kernels, mesh transfers and both inner block solves stay in HBM.  The host-vector GMRES below remains for solver
objects without a device (the CPU check of this module's logic in tests/test_oracle.py).
"""
import typing

import numpy as np

from .ConvectionDiffusion_Solver import ConvectionDiffusionSolver
from .NavierStokes_Solver import NavierStokesSolver


class _Coupled:
    """Residual, Jacobian-vector product and block-Jacobi solve of the coupled system in the unknowns x = [T | u | v | p]."""

    def __init__(self, cd, ns):
        self.cd, self.ns = cd, ns
        self.same = (cd._P, cd._N_ex, cd._N_ey) == (ns._P, ns._N_ex, ns._N_ey)
        self.nT, self.nN = cd.N, ns.N
        self.iter_cd = self.iter_ns = 0

    # mesh-to-mesh transfer: evaluate a field of one solver at the nodes of the other (CD_Component.py:23-36)
    def _to(self, src, dst, f):
        if self.same:
            return f
        shape = (2, dst._P * dst._N_ex + 1, dst._P * dst._N_ey + 1)
        grid = np.reshape(dst.points, shape)
        return src._get_interpol(f, grid).ravel()

    def split(self, x):
        a, b = self.nT, self.nN
        return x[:a], x[a:a + b], x[a + b:a + 2 * b], x[a + 2 * b:]

    def residual(self, x):
        """apply_nonlinear of both components; also the linearisation point of the next linearize()."""
        T, u, v, p = self.split(x)
        rT = self.cd._get_residuals(T, self._to(self.ns, self.cd, u), self._to(self.ns, self.cd, v))
        ru, rv, rc = self.ns._get_residuals(u, v, p, self._to(self.cd, self.ns, T))
        return np.concatenate((rT, ru, rv, rc))

    def linearize(self, x):
        T, u, v, p = self.split(x)
        self.cd._calc_jacobians(T)
        self.ns._calc_jacobians(u, v)

    # ---- the linear solve of the Newton-Krylov coupling on the device (sem_coupled_solve) ------------------------------------
    def device_ready(self):
        """Both solvers are the GPU classes on one (unpartitioned) device."""
        return all(getattr(s, '_dev', None) is not None and s._dev.part is None for s in (self.cd, self.ns))

    def device_solve(self, rhs, atol, restart, maxiter):
        """GMRES(restart) on the coupled Jacobian with one block-Jacobi sweep as preconditioner (BSC:91-94), entirely on the
        device: the coupled vector [dT | du | dv | dp], the Krylov basis, both operator kernels, the mesh-to-mesh transfers and
        the two inner block solves stay in HBM; per iteration only the Hessenberg column travels to the host.  Must follow
        ``residual`` + ``linearize`` (they set the linearisation state of both solvers).  Returns (dx, iterations)."""
        import ctypes as C
        import torch
        from . import SEM
        from . import _lib as L
        from .device import SemDevice
        cd, ns = self.cd, self.ns
        dc, dn = cd._dev, ns._dev
        if not hasattr(self, '_dev_state'):
            tabs = []
            for src, dst in ((ns, cd), (cd, ns)):
                xs = SEM.global_nodes_1d(dst._P, dst._N_ex, dst._dx)
                ys = SEM.global_nodes_1d(dst._P, dst._N_ey, dst._dy)
                tabs.append(src._dev.transfer_tables(xs, ys))
            outer = SemDevice(ns._P, ns._N_ex, ns._N_ey, ns._dx, ns._dy, device=dn.device)
            ns._krylov()
            if not dn.has_ns_schur:
                dn.setup_ns_schur(int(ns.N / 2))
            null = dn.ns_left_null_vector()
            self._dev_state = dict(tabs=tabs, outer=outer, null=null, nrm2=dn.dot(null, null) if null is not None else 0.0, work=None)
        S = self._dev_state
        kr_ns, kr_cd = ns._krylov(), cd._krylov()
        st_ns, st_cd = ns._state(), cd._state()
        q = L.sem_coupled()
        q.ns, q.cd = dn.ctx, dc.ctx
        q.ns_state, q.cd_state = C.pointer(st_ns), C.pointer(st_cd)
        for name, (mx, Sx, ny, Sy), dst in (("ns_to_cd", S["tabs"][0], dc), ("cd_to_ns", S["tabs"][1], dn)):
            setattr(q, name, L.sem_transfer(dst.NX, dst.NY, mx.data_ptr(), ny.data_ptr(), Sx.data_ptr(), Sy.data_ptr()))
        q.kr_ns, q.kr_cd = C.pointer(kr_ns), C.pointer(kr_cd)
        q.ns_work, q.ns_work_len = ns._work.data_ptr(), ns._work.numel()
        q.cd_work, q.cd_work_len = cd._work.data_ptr(), cd._work.numel()
        q.ns_null = S["null"].data_ptr() if S["null"] is not None else None
        q.ns_null_nrm2 = float(S["nrm2"])
        lib = dn.lib
        n = lib.sem_coupled_vec_len(C.byref(q))
        need = lib.sem_coupled_work_len(C.byref(q), int(restart))
        if S["work"] is None or S["work"].numel() < need + 2 * n:
            S["work"] = torch.zeros(need + 2 * n, dtype=torch.float64, device=dn.tdev)
        b, x, work = S["work"][:n], S["work"][n:2 * n], S["work"][2 * n:]
        x.zero_()
        vc, vn = dc.vec_len, dn.vec_len
        parts = self.split(rhs)
        dc.to_device(parts[0], b[:vc].view(dc.NX, dc.LD))
        for k in range(3):
            dn.to_device(parts[1 + k], b[vc + k * vn:vc + (k + 1) * vn].view(dn.NX, dn.LD))
        kr = L.sem_krylov()
        kr.atol, kr.restart, kr.max_iters, kr.precond, kr.verbose = float(atol), int(restart), int(maxiter), 0, 0
        code = L.check(lib.sem_coupled_solve(S["outer"].ctx, C.byref(q), b.data_ptr(), x.data_ptr(), C.byref(kr), work.data_ptr(),
                                             work.numel(), dn.stream), "sem_coupled_solve")
        self.iter_cd += q.solves
        self.iter_ns += q.solves
        ns.krylov_iters.append(q.iters_ns)
        if code != 0:
            raise RuntimeError(f'Boussinesq GMRES: Failed to converge in {kr.iters} iterations')
        dx = np.concatenate([dc.to_host(x[:vc].view(dc.NX, dc.LD))] +
                            [dn.to_host(x[vc + k * vn:vc + (k + 1) * vn].view(dn.NX, dn.LD)) for k in range(3)])
        return dx, kr.iters

    def jvp(self, dx):
        """apply_linear of both components (forward mode)."""
        dT, du, dv, dp = self.split(dx)
        rT = self.cd._get_dresiduals(dT, self._to(self.ns, self.cd, du), self._to(self.ns, self.cd, dv))
        ru, rv, rc = self.ns._get_dresiduals(du, dv, dp, self._to(self.cd, self.ns, dT))
        return np.concatenate((rT, ru, rv, rc))

    def block_jacobi(self, r):
        """LinearBlockJac with one sweep: solve_linear of each component with a zero guess."""
        rT, ru, rv, rc = self.split(r)
        dT = self.cd._get_update(rT)
        # the NS Jacobian is singular (one spurious pressure mode) and a Krylov vector need not lie in its range
        du, dv, dp = getattr(self.ns, '_get_update_inexact', self.ns._get_update)(ru, rv, rc)
        self.iter_cd += 1
        self.iter_ns += 1
        return np.concatenate((dT, du, dv, dp))

    def gauss_seidel_sweep(self, x):
        """solve_nonlinear of the CD component, then of the NS component with the new temperature."""
        T, u, v, p = (a.copy() for a in self.split(x))
        T = self.cd._get_solution(self._to(self.ns, self.cd, u), self._to(self.ns, self.cd, v), T0=T)
        u, v, p = self.ns._get_solution(self._to(self.cd, self.ns, T), u0=u, v0=v, p0=p)
        self.iter_cd += 1
        self.iter_ns += self.ns._k
        return np.concatenate((T, u, v, p))


def _gmres(A, M, b, atol, restart, maxiter):
    """Right-preconditioned restarted GMRES on host vectors (the operator and the preconditioner run on the GPU)."""
    x = np.zeros_like(b)
    r = b.copy()
    beta = np.linalg.norm(r)
    its = 0
    while beta > atol and its < maxiter:
        m = restart
        V = np.zeros((m + 1, b.size))
        Z = np.zeros((m, b.size))
        H = np.zeros((m + 1, m))
        V[0] = r / beta
        g = np.zeros(m + 1)
        g[0] = beta
        cs, sn = np.zeros(m), np.zeros(m)
        k = 0
        for j in range(m):
            Z[j] = M(V[j])
            w = A(Z[j])
            its += 1
            for i in range(j + 1):                       # modified Gram-Schmidt
                H[i, j] = V[i] @ w
                w = w - H[i, j] * V[i]
            H[j + 1, j] = np.linalg.norm(w)
            if H[j + 1, j] > 0:
                V[j + 1] = w / H[j + 1, j]
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            d = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = (H[j, j] / d, H[j + 1, j] / d) if d > 0 else (1.0, 0.0)
            H[j, j], H[j + 1, j] = d, 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            k = j + 1
            if abs(g[j + 1]) <= atol or its >= maxiter:
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), g[:k])
        x = x + y @ Z[:k]
        r = b - A(x)
        beta = np.linalg.norm(r)
    if beta > atol:
        raise RuntimeError(f'Boussinesq GMRES: Failed to converge in {its} iterations')
    return x, its


def solve(cd: ConvectionDiffusionSolver, ns: NavierStokesSolver, mode='JNK', mtol_nonlin=1e-9, AGi=8, AGr=0.8, AGc=0.2,
          mtol_gmres=1e-10, restart=20, maxiter=None, iprint=False):
    """Drive the two solvers to the coupled fixed point; returns (T, u, v, p, info)."""
    sys_ = _Coupled(cd, ns)
    DOF = 3 * ns.N + cd.N                                          # BSC:61
    atol_gmres = mtol_gmres * np.sqrt(DOF)
    atol_nonlin = mtol_nonlin * np.sqrt(DOF)
    x = np.zeros(DOF)
    info = {'mode': mode, 'nonlinear_its': 0, 'gmres_its': []}
    if mode == 'GS':
        for it in range(maxiter or 1000):                          # BSC:77
            x = sys_.gauss_seidel_sweep(x)
            rn = np.linalg.norm(sys_.residual(x))
            info['nonlinear_its'] = it + 1
            if iprint:
                print(f'NL: NLBGS {it + 1} ; {rn}')
            if rn <= atol_nonlin:
                break
        else:
            raise RuntimeError('Boussinesq block Gauss-Seidel: Failed to converge')
    elif mode in ('JNK', 'NJ'):
        r = sys_.residual(x)
        rn = np.linalg.norm(r)
        for it in range(maxiter or (100 if mode == 'JNK' else 1000)):   # BSC:86,90
            if iprint:
                print(f'NL: Newton {it} ; {rn}')
            if rn <= atol_nonlin:
                break
            sys_.linearize(x)
            if mode == 'JNK' and sys_.device_ready():
                dx, its = sys_.device_solve(-r, atol_gmres, restart, 5000)                      # BSC:91-94, device resident
                info['gmres_its'].append(its)
                x = x + dx
                r = sys_.residual(x)
                rn = np.linalg.norm(r)
            elif mode == 'JNK':
                dx, its = _gmres(sys_.jvp, sys_.block_jacobi, -r, atol_gmres, restart, 5000)    # host vectors (CPU checks)
                info['gmres_its'].append(its)
                x = x + dx
                r = sys_.residual(x)
                rn = np.linalg.norm(r)
            else:
                dx = sys_.block_jacobi(-r)                         # LinearBlockJac, maxiter=1 (BSC:89)
                # Armijo-Goldstein backtracking on the residual norm (om.ArmijoGoldsteinLS(rho=AGr, c=AGc, maxiter=AGi))
                alpha, slope = 1.0, -rn
                best = None
                for _ in range(AGi + 1):
                    xt = x + alpha * dx
                    rt = sys_.residual(xt)
                    rtn = np.linalg.norm(rt)
                    if best is None or rtn < best[2]:
                        best = (xt, rt, rtn)
                    if rtn <= rn + AGc * alpha * slope:
                        break
                    alpha *= AGr
                else:
                    xt, rt, rtn = best                             # no trial met the Armijo condition: keep the best one
                x, r, rn = xt, rt, rtn
            info['nonlinear_its'] = it + 1
        else:
            if rn > atol_nonlin:                                   # the last allowed update may have converged
                raise RuntimeError('Boussinesq Newton: Failed to converge')
    else:
        raise ValueError('Unknown method')                         # BSC:96
    info['iter_cd'], info['iter_ns'] = sys_.iter_cd, sys_.iter_ns
    T, u, v, p = sys_.split(x)
    return T, u, v, p, info


def run(points_plot: typing.Tuple[np.ndarray, np.ndarray], L_x: float, L_y: float,
        Re=1.e3, Ra=1.e3, Pr=0.71,
        P_cd=4, N_ex_cd=8, N_ey_cd=8,
        P_ns=4, N_ex_ns=8, N_ey_ns=8,
        mode='JNK',
        mtol_nonlin=1e-9, AGi=8, AGr=0.8, AGc=0.2,
        mtol_gmres=1e-10, restart=20,
        mtol_internal=1e-13) -> typing.Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Steady Boussinesq flow in the differentially heated cavity; arguments and returns as BSC:10-53,99-108."""
    cd = ConvectionDiffusionSolver(L_x=L_x, L_y=L_y, Pe=Re * Pr, P=P_cd, N_ex=N_ex_cd, N_ey=N_ey_cd,
                                   T_W=0.5, T_E=-0.5, mtol=mtol_internal)                       # BSC:56-59
    ns = NavierStokesSolver(L_x=L_x, L_y=L_y, Re=Re, Gr=Ra / Pr, P=P_ns, N_ex=N_ex_ns, N_ey=N_ey_ns,
                            mtol=mtol_internal, mtol_newton=mtol_internal, iprint=[])           # BSC:60-62
    T, u, v, p, _ = solve(cd, ns, mode=mode, mtol_nonlin=mtol_nonlin, AGi=AGi, AGr=AGr, AGc=AGc,
                          mtol_gmres=mtol_gmres, restart=restart)
    return cd._get_interpol(T, points_plot), ns._get_interpol(u, points_plot), ns._get_interpol(v, points_plot)


def study_title(mode, Re, Ra, Pr, P, N_e, mtol_nonlin, AGi, AGr, AGc, mtol_gmres, restart, mtol_internal):
    """File stem of a study run, as study/Boussinesq_run.py:35-43."""
    title = f"Boussinesq{mode}_{Re:.1e}~{Ra:.1e}~{Pr}_{P}~{N_e}_"
    if mode == 'GS':
        return title + f"{mtol_nonlin:.0e}_{mtol_internal:.0e}"
    if mode == 'NJ':
        return title + f"{mtol_nonlin:.0e}~{AGi}~{AGr}~{AGc}_{mtol_internal:.0e}"
    if mode == 'JNK':
        return title + f"{mtol_nonlin:.0e}_{mtol_gmres:.0e}~{restart}_{mtol_internal:.0e}"
    raise RuntimeError('Unknown method')


def run_study(save=True, out_dir="Boussinesq_study", L_x=1., L_y=1., Re=1.e3, Ra=1.e3, Pr=0.71, P=4, N_e=8, mode='JNK',
              mtol_nonlin=1e-10, AGi=8, AGr=0.8, AGc=0.2, mtol_gmres=1e-13, restart=20, mtol_internal=1e-13):
    """One run of the parameter study (study/Boussinesq_run.py:26-135): CD on N_e/2 x N_e/2 elements, NS on N_e x N_e, and
    the same output file -- ``np.savez`` with positional arrays ``arr_0..arr_3 = T_e, u_e, v_e, [iters_cd, iters_ns,
    iter_nonlin]`` in the element layout ``[m, n, i, j]`` of ``SEM.scatter`` -- so the study's post-processing reads it
    unchanged.  Returns (title, T_e, u_e, v_e, iters)."""
    import os
    from . import SEM
    title = study_title(mode, Re, Ra, Pr, P, N_e, mtol_nonlin, AGi, AGr, AGc, mtol_gmres, restart, mtol_internal)
    cd = ConvectionDiffusionSolver(L_x=L_x, L_y=L_y, Pe=Re * Pr, P=P, N_ex=int(N_e / 2), N_ey=int(N_e / 2),
                                   T_W=0.5, T_E=-0.5, mtol=mtol_internal)
    ns = NavierStokesSolver(L_x=L_x, L_y=L_y, Re=Re, Gr=Ra / Pr, P=P, N_ex=N_e, N_ey=N_e,
                            mtol=mtol_internal, mtol_newton=mtol_internal, iprint=[])
    T, u, v, p, info = solve(cd, ns, mode=mode, mtol_nonlin=mtol_nonlin, AGi=AGi, AGr=AGr, AGc=AGc,
                             mtol_gmres=mtol_gmres, restart=restart)
    T_e = SEM.scatter(T, cd._P, cd._N_ex, cd._N_ey)
    u_e = SEM.scatter(u, ns._P, ns._N_ex, ns._N_ey)
    v_e = SEM.scatter(v, ns._P, ns._N_ex, ns._N_ey)
    iters = [info['iter_cd'], info['iter_ns'], info['nonlinear_its']]
    if save:
        os.makedirs(out_dir, exist_ok=True)
        np.savez(os.path.join(out_dir, title), T_e, u_e, v_e, iters)
    return title, T_e, u_e, v_e, iters
