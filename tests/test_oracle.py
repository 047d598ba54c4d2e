"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference, plus the analytic
known answers of SURVEY.md section 4.  This is what pins the oracle (it is the checker for every -m gpu test)."""
import numpy as np
import pytest

from tests.conftest import relerr
from oracle import sem_oracle as so
from tests.golden.make_golden_cases import CD_CASES, MESHES, NS_CASES


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16])
def test_gll_tables_match_reference(golden, P):
    g = golden("gll")
    x, w, _ = so.gll(P)
    assert np.array_equal(x, g[f"x{P}"]) and np.array_equal(w, g[f"w{P}"])
    assert np.array_equal(so.diff_matrix(P), g[f"D{P}"])
    assert np.array_equal(so.stiff_1d(P), g[f"K{P}"])
    assert np.array_equal(so.grad_1d(P), g[f"G{P}"])
    assert np.allclose(so.eval_matrix(P, np.linspace(-1, 1, 7)), g[f"S{P}"], rtol=0, atol=1e-13)


def test_gll_known_answers():
    x, w, _ = so.gll(4)
    assert np.allclose(x, [-1, -np.sqrt(3 / 7), 0, np.sqrt(3 / 7), 1], atol=1e-15)
    assert np.allclose(w, [1 / 10, 49 / 90, 32 / 45, 49 / 90, 1 / 10], atol=1e-15)
    for P in range(1, 13):
        x, w, _ = so.gll(P)
        D = so.diff_matrix(P)
        assert abs(w.sum() - 2) < 1e-14
        assert np.max(np.abs(D.sum(axis=1))) < 1e-12
        assert np.max(np.abs(D @ x**P - P * x**(P - 1))) < 1e-12
        K = so.stiff_1d(P)
        assert np.max(np.abs(K - K.T)) < 1e-13 and np.max(np.abs(K @ np.ones(P + 1))) < 1e-12


@pytest.mark.parametrize("tag,P,nx,ny,Lx,Ly", MESHES)
def test_operators_match_reference(golden, tag, P, nx, ny, Lx, Ly):
    g = golden("operators")
    x = g[f"{tag}/x"]
    M, K, Gx, Gy = so.global_operators(P, nx, ny, Lx / nx, Ly / ny)
    assert relerr(M, g[f"{tag}/Mdiag"]) < 1e-15
    assert relerr(K @ x, g[f"{tag}/Kx"]) < 1e-13
    assert relerr(Gx @ x, g[f"{tag}/Gxx"]) < 1e-13
    assert relerr(Gy @ x, g[f"{tag}/Gyx"]) < 1e-13
    assert abs(M.sum() - Lx * Ly) < 1e-13 * Lx * Ly
    assert np.array_equal(so.global_nodes(P, nx, ny, Lx / nx, Ly / ny), g[f"{tag}/points"])
    assert relerr(so.assemble_vector(g[f"{tag}/A_e"]), g[f"{tag}/assembled"]) < 1e-15
    assert np.array_equal(so.scatter(x, P, nx, ny), g[f"{tag}/scattered"])
    val = so.interpolate(x, P, nx, ny, Lx / nx, Ly / ny, (g[f"{tag}/xp"], g[f"{tag}/yp"]))
    assert relerr(val, g[f"{tag}/interp"]) < 1e-13


def test_global_index_errors():
    with pytest.raises(ValueError):
        so.global_index(4, 2, 2, 2, 0, 0, 0)
    with pytest.raises(ValueError):
        so.scatter(np.zeros(7), 2, 2, 2)
    assert so.global_index(4, 16, 16, 3, 5, 2, 1) == 5 * 4 + 1 + 65 * (3 * 4 + 2)


@pytest.mark.parametrize("tag,kw", CD_CASES)
def test_cd_matches_reference(golden, tag, kw):
    g = golden("cd")
    cd = so.CDOracle(mtol=1e-13, **kw)
    k = lambda s: g[f"{tag}/{s}"]
    assert np.array_equal(cd._mask_dir, k("mask_dir"))
    assert np.array_equal(np.nan_to_num(cd._dirichlet, nan=9e99), np.nan_to_num(k("dirichlet"), nan=9e99))
    assert relerr(cd._get_residuals(k("T"), k("ur"), k("vr")), k("res_r")) < 1e-13
    cd._calc_jacobians(k("T"))
    assert relerr(cd._get_dresiduals(k("dT")), k("dres_r")) < 1e-13
    assert relerr(cd._get_dresiduals(k("dT"), k("du"), k("dv")), k("dres_r_uv")) < 1e-13
    assert relerr(cd._get_dresiduals(k("dT"), du=k("du")), k("dres_r_u")) < 1e-13
    assert relerr(cd._get_residuals(k("T"), k("u"), k("v")), k("res")) < 1e-13
    assert relerr(cd._get_dresiduals(k("dT")), k("dres")) < 1e-13
    assert relerr(cd._get_solution(k("u"), k("v")), k("T_sol")) < 1e-9
    assert relerr(cd._get_update(k("rhs")), k("dT_sol")) < 1e-9


@pytest.mark.parametrize("tag,kw,solve", NS_CASES)
def test_ns_matches_reference(golden, tag, kw, solve):
    g = golden("ns")
    ns = so.NSOracle(mtol=1e-13, mtol_newton=1e-13, **kw)
    k = lambda s: g[f"{tag}/{s}"]
    assert np.array_equal(ns._mask_bound, k("mask_bound"))
    ru, rv, rc = ns._get_residuals(k("u"), k("v"), k("p"), k("T"))
    assert relerr(ru, k("res_u")) < 1e-13 and relerr(rv, k("res_v")) < 1e-13 and relerr(rc, k("res_c")) < 1e-13
    ns._calc_jacobians(k("u"), k("v"))
    a, b, c = ns._get_dresiduals(k("du"), k("dv"), k("dp"))
    assert relerr(a, k("dres_u")) < 1e-13 and relerr(b, k("dres_v")) < 1e-13 and relerr(c, k("dres_c")) < 1e-13
    a, b, c = ns._get_dresiduals(k("du"), k("dv"), k("dp"), k("dT"))
    assert relerr(a, k("dresT_u")) < 1e-13 and relerr(b, k("dresT_v")) < 1e-13 and relerr(c, k("dresT_c")) < 1e-13
    # the assembled Jacobian matrix reproduces the JVP (it is what the oracle's direct solve factorises)
    J = ns.jacobian_matrix()
    jv = J @ np.hstack((k("du"), k("dv"), k("dp")))
    assert relerr(jv, np.hstack((k("dres_u"), k("dres_v"), k("dres_c")))) < 1e-13
    if not solve:
        return
    u, v, p = ns._get_solution(k("T_in"))
    # The stored reference fields stop at the reference's own tolerance (|res| <= 1e-13 sqrt(3N)); on the 8x8 mesh
    # its smallest singular values (~1e-5) turn that into ~6e-8 in p -- a floor of the reference, not of the oracle.
    ptol = 2e-7 if tag == "c3" else 1e-8
    assert relerr(u, k("u_sol")) < 1e-8 and relerr(v, k("v_sol")) < 1e-8 and relerr(p, k("p_sol")) < ptol
    ns._get_residuals(k("u_sol"), k("v_sol"), k("p_sol"), k("T_in"))
    ns._calc_jacobians(k("u_sol"), k("v_sol"))
    a, b, c = ns._get_update(k("rhs_u"), k("rhs_v"), k("rhs_c"))
    # the stored update was solved by the reference to mtol = 1e-11 only (see make_golden.py)
    assert relerr(a, k("upd_u")) < 1e-7 and relerr(b, k("upd_v")) < 1e-7 and relerr(c, k("upd_p")) < 1e-5


@pytest.mark.parametrize("P,nx,ny,Lx,Ly", [(4, 5, 3, 1.3, 0.7), (3, 4, 4, 1.0, 1.0), (8, 3, 2, 2.0, 1.0)])
def test_kronecker_operators_match_the_assembled_ones(P, nx, ny, Lx, Ly):
    """oracle.KronOps (the full-size checker of tests/test_gpu_parity.py::test_full_size_against_kronecker_oracle) reproduces the
    assembled operators / Jacobian-vector products of the pinned oracle classes."""
    k = so.KronOps(P, nx, ny, Lx / nx, Ly / ny)
    sh = k.shape
    cd = so.CDOracle(Lx, Ly, 7.0, P, nx, ny, T_W=0.5, T_E=-0.5, T_S=0.2)
    rng = np.random.default_rng(0)
    T, u, v, p, a, b, c = (rng.standard_normal(cd.N) for _ in range(7))
    M, K, Gx, Gy = so.global_operators(P, nx, ny, Lx / nx, Ly / ny)
    assert relerr(k.K(a.reshape(sh)).ravel(), K @ a) < 1e-14 and relerr(k.Gx(a.reshape(sh)).ravel(), Gx @ a) < 1e-14
    assert relerr(k.Gy(a.reshape(sh)).ravel(), Gy @ a) < 1e-14 and relerr(k.M(a.reshape(sh)).ravel(), M * a) < 1e-14
    cd._get_residuals(T, u, v)
    assert relerr(k.cd_jvp(7.0, a.reshape(sh), u.reshape(sh), v.reshape(sh), (1, 1, 1, 0)).ravel(), cd._get_dresiduals(a)) < 1e-14
    ns = so.NSOracle(Lx, Ly, 30.0, 0.0, P, nx, ny, u_N=1.0)
    ns._get_residuals(u, v, p, T)
    ns._calc_jacobians(u, v)
    got = k.ns_jvp(30.0, u.reshape(sh), v.reshape(sh), a.reshape(sh), b.reshape(sh), c.reshape(sh))
    for x, y in zip(got, ns._get_dresiduals(a, b, c)):
        assert relerr(x.ravel(), y) < 1e-14


def test_reference_algorithm_ports_reach_the_direct_solution():
    """The line-by-line ports of the reference's own linear solvers (SciPy LGMRES on the CD operator, CD:123-156; SuperLU +
    LGMRES on the pressure Schur complement, NS:162-236) -- the host timing baselines of bench.py -- land on the oracle's
    direct solution, and the NS port needs the number of Schur evaluations the survey probed on the reference."""
    kw = dict(L_x=1.0, L_y=1.0, Pe=40.0, P=4, N_ex=8, N_ey=8, T_W=0.5, T_E=-0.5)
    cd = so.CDOracle(mtol=1e-12, **kw)
    u = cd._get_vector(lambda x, y: y - 0.5)
    v = cd._get_vector(lambda x, y: 0.5 - x)
    T, evals = cd._get_solution_lgmres(u, v)
    assert relerr(T, cd._get_solution(u, v)) < 1e-9 and evals > 50
    ns = so.NSOracle(1.0, 1.0, 100.0, 0.0, 4, 6, 6, u_N=1.0, mtol=1e-12, mtol_newton=1e-11)
    ua, va, pa = ns._get_solution(np.zeros(ns.N), algorithm='reference')
    ns2 = so.NSOracle(1.0, 1.0, 100.0, 0.0, 4, 6, 6, u_N=1.0, mtol=1e-12, mtol_newton=1e-11)
    ub, vb, pb = ns2._get_solution(np.zeros(ns.N))
    assert ns._k == ns2._k and relerr(ua, ub) < 1e-8 and relerr(va, vb) < 1e-8 and relerr(pa, pb) < 1e-7


def test_readme_helmholtz_known_answer():
    """Solvers/README.md:49-96: (lambda M + K) u = M f with f = cos(pi x/Lx) cos(pi y/Ly), homogeneous Neumann;
    exact u = f / (lambda + pi^2 (1/Lx^2 + 1/Ly^2)).  Max error 9.5e-7 at P=4, 2x3 elements."""
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    Lx, Ly, lam, P, nx, ny = 2.0, 1.0, 1.0, 4, 2, 3
    M, K, _, _ = so.global_operators(P, nx, ny, Lx / nx, Ly / ny)
    pts = so.global_nodes(P, nx, ny, Lx / nx, Ly / ny)
    f = np.cos(np.pi * pts[0] / Lx) * np.cos(np.pi * pts[1] / Ly)
    u = spla.spsolve((lam * sps.diags(M) + K).tocsc(), M * f)
    exact = f / (lam + np.pi**2 * (1 / Lx**2 + 1 / Ly**2))
    assert np.max(np.abs(u - exact)) < 5e-5


def test_boussinesq_fixed_point_residual(golden):
    """The stored C3 coupled state is a root of the oracle's four residuals and reproduces de Vahl Davis (1983)."""
    g = golden("boussinesq_c3")
    Re, Ra, Pr = 1e3, 1e3, 0.71
    cd = so.CDOracle(1., 1., Re * Pr, 4, 8, 8, T_W=0.5, T_E=-0.5)
    ns = so.NSOracle(1., 1., Re, Ra / Pr, 4, 8, 8)
    r = np.hstack((cd._get_residuals(g["T"], g["u"], g["v"]),) + ns._get_residuals(g["u"], g["v"], g["p"], g["T"]))
    assert np.linalg.norm(r) < 1e-9 * np.sqrt(r.size)
    assert abs(float(g["umax_RePr"]) - 3.649) < 5e-3 and abs(float(g["vmax_RePr"]) - 3.697) < 1e-2


def test_boussinesq_coupler_logic_on_the_oracle_solvers():
    """The native coupled driver (sem_b200/Boussinesq_SequentialCoupler.py) only talks to its two solvers through the
    reference's method names, so its host logic can be exercised on the CPU with the oracle classes: block Gauss-Seidel
    and Jacobian-free Newton-Krylov reach the same coupled state (the C3 physics on P = 4, 4 x 4 elements)."""
    import importlib.util
    import os
    import sys
    import types
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sem_b200", "Boussinesq_SequentialCoupler.py")
    src = open(path).read()
    # the module imports the GPU classes only for type annotations and for run(): strip them for the CPU check
    src = src.replace("from .ConvectionDiffusion_Solver import ConvectionDiffusionSolver\n", "ConvectionDiffusionSolver = object\n")
    src = src.replace("from .NavierStokes_Solver import NavierStokesSolver\n", "NavierStokesSolver = object\n")
    mod = types.ModuleType("bsc_host")
    exec(compile(src, path, "exec"), mod.__dict__)
    Re, Ra, Pr = 1e3, 1e3, 0.71
    out = {}
    for mode in ("GS", "JNK"):
        cd = so.CDOracle(1., 1., Re * Pr, 4, 4, 4, T_W=0.5, T_E=-0.5, mtol=1e-13)
        ns = so.NSOracle(1., 1., Re, Ra / Pr, 4, 4, 4, mtol=1e-13, mtol_newton=1e-13)
        for s in (cd, ns):
            s._P, s._N_ex, s._N_ey = 4, 4, 4
        out[mode] = mod.solve(cd, ns, mode=mode, mtol_nonlin=1e-11, mtol_gmres=1e-13, maxiter=40)
    for a, b in zip(out["GS"][:3], out["JNK"][:3]):
        assert np.linalg.norm(a - b) <= 1e-8 * max(np.linalg.norm(b), 1e-30)
    assert out["JNK"][4]["nonlinear_its"] < out["GS"][4]["nonlinear_its"]
    assert mod.study_title('JNK', 1e3, 1e3, 0.71, 4, 8, 1e-10, 8, 0.8, 0.2, 1e-13, 20, 1e-13) == \
        "BoussinesqJNK_1.0e+03~1.0e+03~0.71_4~8_1e-10_1e-13~20_1e-13"


def _gmres_right(A, b, Pinv, tol, maxit):
    """Plain right-preconditioned GMRES (CGS2), numpy: x = sum_k y_k Pinv(v_k)."""
    n = b.size
    beta = np.linalg.norm(b)
    V = np.zeros((maxit + 1, n)); Z = np.zeros((maxit, n)); H = np.zeros((maxit + 1, maxit))
    V[0] = b / beta
    for k in range(maxit):
        Z[k] = Pinv(V[k]); w = A(Z[k])
        for _ in range(2):
            h = V[:k + 1] @ w; w -= h @ V[:k + 1]; H[:k + 1, k] += h
        H[k + 1, k] = np.linalg.norm(w); V[k + 1] = w / H[k + 1, k]
        e1 = np.zeros(k + 2); e1[0] = beta
        y, res = np.linalg.lstsq(H[:k + 2, :k + 1], e1, rcond=None)[:2]
        if np.linalg.norm(H[:k + 2, :k + 1] @ y - e1) <= tol:
            break
    return y @ Z[:k + 1], k + 1


def test_null_vector_structure_and_member_selection():
    """Facts about the singular NS Jacobian that the device solver's preconditioner relies on:
    (1) the pressure part of the left null vector is the tensor product (1 - L_P(xi)) x (1 - L_P(eta)), element by element,
    independent of the linearisation point; (2) right-preconditioned GMRES with the block lower-triangular preconditioner
    lands on the reference's member of the solution set for ANY Schur-block preconditioner once that is given the rank-one
    correction z <- z - l_c (m_c.z - l_c.r) / (m_c.l_c), m_c = M_p l_c -- and on a different member without it."""
    import scipy.sparse.linalg as spla
    from numpy.polynomial import legendre as npl
    P, ne = 4, 4
    ns = so.NSOracle(1.0, 1.0, 50.0, 0.0, P, ne, ne, u_N=1.0, mtol=1e-13, mtol_newton=1e-13)
    N, n1 = ns.N, ne * P + 1
    T = np.zeros(N)
    u, v, p = ns._get_solution(T, max_newton=2)
    ru, rv, rc = ns._get_residuals(u, v, p, T)
    ns._calc_jacobians(u, v)
    J = ns.jacobian_matrix().tocsr()
    l = ns._left_null(J.tocsc())
    assert l is not None
    lc = l[2 * N:]
    one_minus_LP = np.tile((1.0 - npl.legval(so.gll(P)[0], [0] * P + [1]))[:-1], ne)
    one_minus_LP = np.append(one_minus_LP, 0.0)
    model = np.outer(one_minus_LP, one_minus_LP).ravel()
    c = (model @ lc) / (model @ model)
    assert np.linalg.norm(lc - c * model) < 1e-10 * np.linalg.norm(lc)
    # (1b) the whole null vector is separable: l_u = -(G1^T f) (x) (M1 f), l_v = -(M1 f) (x) (G1^T f) on the walls, f = 1 - L_P
    # (what SemDevice.ns_left_null_vector builds for the range projection of the coupled drivers)
    from oracle.ns_precond import Dir1D
    d1 = Dir1D(P, ne, 1.0 / ne)
    mb = ns._mask_bound.reshape(n1, n1)
    lu = np.where(mb, -np.outer(d1.G.T @ d1.lfac, d1.M * d1.lfac), 0.0)
    lv = np.where(mb, -np.outer(d1.M * d1.lfac, d1.G.T @ d1.lfac), 0.0)
    lm = np.hstack((lu.ravel(), lv.ravel(), np.outer(d1.lfac, d1.lfac).ravel()))
    assert np.linalg.norm(J.T @ lm) < 1e-12 * np.linalg.norm(lm)
    assert np.linalg.norm(l - lm * (lm @ l) / (lm @ lm)) < 1e-10 * np.linalg.norm(l)
    b = -np.hstack((ru, rv, rc))
    x_ref = np.hstack(ns._get_update(-ru, -rv, -rc))
    Mp = ns._M.copy(); Mp[ns._pin] = 1.0
    mc = Mp * lc
    lu = spla.splu(J[:2 * N, :2 * N].tocsc()); C = J[2 * N:, :2 * N].tocsr()
    Dr = np.exp(np.random.default_rng(1).uniform(-1, 1, N))
    other = lambda y: Dr * y / Mp
    def corrected(y):
        z = other(y)
        return z - lc * ((mc @ z - lc @ y) / (mc @ lc))
    def tri(Sinv):
        def f(r):
            za = lu.solve(r[:2 * N])
            return np.hstack((za, Sinv(r[2 * N:] - C @ za)))
        return f
    tol = 1e-11 * np.linalg.norm(b)      # stop AT convergence: on a singular system further steps drift along the null vector
    errs = {}
    for name, Si in (("mass", lambda y: y / Mp), ("other", other), ("corrected", corrected)):
        x, its = _gmres_right(lambda z: J @ z, b, tri(Si), tol, 400)
        assert its < 400 and np.linalg.norm(J @ x - b) < 1e-9 * np.linalg.norm(b), name
        errs[name] = relerr(x[2 * N:], x_ref[2 * N:])
    assert errs["mass"] < 1e-8 and errs["corrected"] < 1e-8
    assert errs["other"] > 1e-4
