"""Generate the committed golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz   (about 2-3 minutes)

The reference at /root/reference is imported through ``oracle/ref_shim.py`` (stand-in ``sparse`` module and the
``lgmres(tol=)`` adapter; the reference sources are not touched or copied).  Every array stored here is an output of
the reference's own functions on seeded inputs; the inputs are stored beside them so the tests need neither the
reference nor the RNG.  Solves use ``mtol = mtol_newton = 1e-13`` (the tolerances the reference's couplers use,
``OpenMDAO/Boussinesq_SequentialCoupler.py:17,56,59``) because the default tolerances leave the reference itself
3e-4 .. 3e-3 away from its own converged answer.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_shim  # noqa: E402
from tests.golden.make_golden_cases import CD_CASES, MESHES, NS_CASES  # noqa: E402

GLL, SEM, CDS, NSS = ref_shim.install()


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def gll_tables():
    out = {}
    for P in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16):
        x, w, V = GLL.standard_nodes(P)
        out[f"x{P}"], out[f"w{P}"] = x, w
        out[f"D{P}"] = GLL.standard_differentiation_matrix(P)
        out[f"K{P}"] = GLL.standard_stiffness_matrix(P)
        out[f"G{P}"] = GLL.standard_gradient_matrix(P)
        xi = np.linspace(-1, 1, 7)
        out[f"S{P}"] = GLL.standard_evaluation_matrix(P, xi)
    save("gll", **out)




def operator_applies():
    """K, G_x, G_y, M applied to seeded vectors; assemble / scatter / global_index; interpolation."""
    out = {}
    for tag, P, nx, ny, Lx, Ly in MESHES:
        dx, dy = Lx / nx, Ly / ny
        N = (nx * P + 1) * (ny * P + 1)
        rng = np.random.default_rng(hash(tag) % 2**31 if False else sum(map(ord, tag)))
        x = rng.standard_normal(N)
        M = SEM.global_mass_matrix(P, nx, ny, dx, dy)
        K = SEM.global_stiffness_matrix(P, nx, ny, dx, dy)
        Gx, Gy = SEM.global_gradient_matrices(P, nx, ny, dx, dy)
        out[f"{tag}/x"] = x
        out[f"{tag}/Mdiag"] = M.diagonal()
        out[f"{tag}/Mx"] = M @ x
        out[f"{tag}/Kx"] = K @ x
        out[f"{tag}/Gxx"] = Gx @ x
        out[f"{tag}/Gyx"] = Gy @ x
        A_e = rng.standard_normal((nx, ny, P + 1, P + 1))
        out[f"{tag}/A_e"] = A_e
        out[f"{tag}/assembled"] = SEM.assemble(A_e)
        out[f"{tag}/scattered"] = SEM.scatter(x, P, nx, ny)
        out[f"{tag}/points"] = SEM.global_nodes(P, nx, ny, dx, dy)
        xp, yp = np.meshgrid(np.linspace(0, Lx, 13), np.linspace(0, Ly, 9), indexing='ij')
        out[f"{tag}/xp"], out[f"{tag}/yp"] = xp, yp
        out[f"{tag}/interp"] = SEM.eval_interpolation(SEM.scatter(x, P, nx, ny),
                                                      SEM.element_nodes(P, nx, ny, dx, dy), (xp, yp))
    save("operators", **out)




def cd_cases():
    out = {}
    for tag, kw in CD_CASES:
        cd = CDS(mtol=1e-13, **kw)
        rng = np.random.default_rng(100 + sum(map(ord, tag)))
        Lx, Ly = kw["L_x"], kw["L_y"]
        u = cd._get_vector(lambda x, y: y - Ly / 2)
        v = cd._get_vector(lambda x, y: Lx / 2 - x)
        ur, vr = rng.standard_normal(cd.N), rng.standard_normal(cd.N)
        T, dT, du, dv = (rng.standard_normal(cd.N) for _ in range(4))
        out[f"{tag}/u"], out[f"{tag}/v"], out[f"{tag}/ur"], out[f"{tag}/vr"] = u, v, ur, vr
        out[f"{tag}/T"], out[f"{tag}/dT"], out[f"{tag}/du"], out[f"{tag}/dv"] = T, dT, du, dv
        out[f"{tag}/res_r"] = cd._get_residuals(T, ur, vr)
        cd._calc_jacobians(T)
        out[f"{tag}/dres_r"] = cd._get_dresiduals(dT)
        out[f"{tag}/dres_r_uv"] = cd._get_dresiduals(dT, du, dv)
        out[f"{tag}/dres_r_u"] = cd._get_dresiduals(dT, du=du)
        out[f"{tag}/res"] = cd._get_residuals(T, u, v)
        out[f"{tag}/dres"] = cd._get_dresiduals(dT)
        t = time.time()
        out[f"{tag}/T_sol"] = cd._get_solution(u, v)
        print(f"  CD {tag}: N={cd.N} solve {time.time() - t:.1f}s")
        # update from a random right-hand side about the same linearisation point
        rhs = rng.standard_normal(cd.N)
        out[f"{tag}/rhs"] = rhs
        out[f"{tag}/dT_sol"] = cd._get_update(rhs)
        out[f"{tag}/mask_dir"] = cd._mask_dir
        out[f"{tag}/dirichlet"] = cd._dirichlet
    save("cd", **out)




def ns_cases():
    out = {}
    for tag, kw, solve in NS_CASES:
        ns = NSS(mtol=1e-13, mtol_newton=1e-13, iprint=[], **kw)
        rng = np.random.default_rng(200 + sum(map(ord, tag)))
        u, v, p, T, du, dv, dp, dT = (rng.standard_normal(ns.N) for _ in range(8))
        for k, a in zip("u v p T du dv dp dT".split(), (u, v, p, T, du, dv, dp, dT)):
            out[f"{tag}/{k}"] = a
        ru, rv, rc = ns._get_residuals(u, v, p, T)
        out[f"{tag}/res_u"], out[f"{tag}/res_v"], out[f"{tag}/res_c"] = ru, rv, rc
        ns._calc_jacobians(u, v)
        a, b, c = ns._get_dresiduals(du, dv, dp)
        out[f"{tag}/dres_u"], out[f"{tag}/dres_v"], out[f"{tag}/dres_c"] = a, b, c
        a, b, c = ns._get_dresiduals(du, dv, dp, dT)
        out[f"{tag}/dresT_u"], out[f"{tag}/dresT_v"], out[f"{tag}/dresT_c"] = a, b, c
        out[f"{tag}/mask_bound"] = ns._mask_bound
        out[f"{tag}/dirichlet_u"], out[f"{tag}/dirichlet_v"] = ns._dirichlet_u, ns._dirichlet_v
        if solve:
            Lx = kw["L_x"]
            Tf = ns._get_vector(lambda x, y: 0.5 - x / Lx)
            out[f"{tag}/T_in"] = Tf
            t = time.time()
            us, vs, ps = ns._get_solution(Tf)
            print(f"  NS {tag}: N={ns.N} Newton its {ns._k}  {time.time() - t:.1f}s")
            out[f"{tag}/u_sol"], out[f"{tag}/v_sol"], out[f"{tag}/p_sol"] = us, vs, ps
            out[f"{tag}/newton_its"] = np.array(ns._k)
            # one linear update about the converged state with a random right-hand side
            ns._get_residuals(us, vs, ps, Tf)
            ns._calc_jacobians(us, vs)
            # a consistent right-hand side (the Jacobian is singular: a random one is not in its range)
            ra, rb, rc2 = ns._get_dresiduals(*(rng.standard_normal(ns.N) for _ in range(3)))
            out[f"{tag}/rhs_u"], out[f"{tag}/rhs_v"], out[f"{tag}/rhs_c"] = ra, rb, rc2
            ns._mtol = 1e-11
            a, b, c = ns._get_update(ra, rb, rc2)
            out[f"{tag}/upd_u"], out[f"{tag}/upd_v"], out[f"{tag}/upd_p"] = a, b, c
    save("ns", **out)


def boussinesq_fixed_point():
    """C3 physics (Examples/Boussinesq_Sequential_Example.py:22-37): coupled fixed point by block Gauss-Seidel over
    the reference's own solve calls, mirroring ``solve_nonlinear`` of the two OpenMDAO components
    (ConvectionDiffusion_Component.py:59-61, NavierStokes_Component.py:62-65) on identical meshes (change_inputs is
    then the identity).  OpenMDAO itself is not installable here, so the coupler's iteration path is unpinned; the
    fixed point it converges to does not depend on that path."""
    Re, Ra, Pr, P, ne = 1e3, 1e3, 0.71, 4, 8
    cd = CDS(L_x=1., L_y=1., Pe=Re * Pr, P=P, N_ex=ne, N_ey=ne, T_W=0.5, T_E=-0.5, mtol=1e-13)
    ns = NSS(L_x=1., L_y=1., Re=Re, Gr=Ra / Pr, P=P, N_ex=ne, N_ey=ne, mtol=1e-13, mtol_newton=1e-13, iprint=[])
    N = cd.N
    T, u, v, p = (np.zeros(N) for _ in range(4))
    for sweep in range(60):
        T = cd._get_solution(u, v, T0=T)
        u, v, p = ns._get_solution(T, u0=u, v0=v, p0=p)
        r = np.hstack((cd._get_residuals(T, u, v),) + ns._get_residuals(u, v, p, T))
        nr = np.linalg.norm(r)
        print(f"  GS sweep {sweep}: |res| = {nr:.3e}")
        if nr <= 1e-11 * np.sqrt(4 * N):
            break
    xp, yp = np.meshgrid(np.linspace(0, 1, 101), np.linspace(0, 1, 101), indexing='ij')
    up, vp = ns._get_interpol(u, (xp, yp)), ns._get_interpol(v, (xp, yp))
    save("boussinesq_c3", T=T, u=u, v=v, p=p, sweeps=np.array(sweep + 1),
         umax_RePr=np.array(up.max() * Re * Pr), vmax_RePr=np.array(vp.max() * Re * Pr))


if __name__ == "__main__":
    which = sys.argv[1:] or ["gll", "operators", "cd", "ns", "boussinesq"]
    if "gll" in which:
        gll_tables()
    if "operators" in which:
        operator_applies()
    if "cd" in which:
        cd_cases()
    if "ns" in which:
        ns_cases()
    if "boussinesq" in which:
        boussinesq_fixed_point()
