"""Probe for the two remaining relaxed gates: (1) linear update about the converged C2 / C3 state and (2) the Boussinesq block
Gauss-Seidel fixed point -- GPU against the ORACLE driven to tight tolerances (direct solves), not against the stored
reference outputs, which stop at the reference's own tolerances."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import sem_b200
from oracle import sem_oracle as so
from tests.golden.make_golden_cases import NS_CASES
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
g = np.load('tests/golden/ns.npz')
for tag in ('c2', 'c3'):
    kw = [c for c in NS_CASES if c[0] == tag][0][1]
    k = lambda s: g[f"{tag}/{s}"]
    o = so.NSOracle(mtol=1e-15, mtol_newton=1e-15, **kw)
    o._get_residuals(k("u_sol"), k("v_sol"), k("p_sol"), k("T_in")); o._calc_jacobians(k("u_sol"), k("v_sol"))
    ao, bo, co = o._get_update(k("rhs_u"), k("rhs_v"), k("rhs_c"))
    print(tag, 'oracle vs stored upd', rel(ao, k("upd_u")), rel(bo, k("upd_v")), rel(co, k("upd_p")), flush=True)
    for precond in ('fdm', 'full'):
        for mt in (1e-11, 1e-13, 1e-14):
            try:
                ns = sem_b200.NavierStokesSolver(mtol=mt, mtol_newton=mt, iprint=[], precond=precond, **kw)
                ns._get_residuals(k("u_sol"), k("v_sol"), k("p_sol"), k("T_in")); ns._calc_jacobians(k("u_sol"), k("v_sol"))
                a, b, c = ns._get_update(k("rhs_u"), k("rhs_v"), k("rhs_c"))
                print(tag, precond, mt, 'gpu vs oracle upd', rel(a, ao), rel(b, bo), rel(c, co), flush=True)
            except Exception as e:
                print(tag, precond, mt, 'FAILED', str(e)[:120], flush=True)

Re, Ra, Pr = 1e3, 1e3, 0.71
def gs(cd, ns, tol, maxs=80):
    N = cd.N
    T, u, v, p = (np.zeros(N) for _ in range(4))
    for sweep in range(maxs):
        T = cd._get_solution(u, v, T0=T)
        u, v, p = ns._get_solution(T, u0=u, v0=v, p0=p)
        r = np.hstack((cd._get_residuals(T, u, v),) + tuple(ns._get_residuals(u, v, p, T)))
        if np.linalg.norm(r) <= tol * np.sqrt(4 * N):
            break
    return T, u, v, p, sweep + 1, float(np.linalg.norm(r))
t0 = time.time()
cdo = so.CDOracle(1., 1., Re * Pr, 4, 8, 8, T_W=0.5, T_E=-0.5, mtol=1e-15)
nso = so.NSOracle(1., 1., Re, Ra / Pr, 4, 8, 8, mtol=1e-15, mtol_newton=1e-15)
To, uo, vo, po, so_, ro = gs(cdo, nso, 1e-14)
print('oracle GS', so_, ro, time.time() - t0, flush=True)
gb = np.load('tests/golden/boussinesq_c3.npz')
print('oracle-tight vs stored', rel(To, gb['T']), rel(uo, gb['u']), rel(vo, gb['v']), rel(po, gb['p']), flush=True)
for mt, tol in ((1e-13, 1e-11), (1e-14, 1e-13), (1e-14, 1e-14)):
    cd = sem_b200.ConvectionDiffusionSolver(1., 1., Re * Pr, 4, 8, 8, T_W=0.5, T_E=-0.5, mtol=mt)
    ns = sem_b200.NavierStokesSolver(1., 1., Re, Ra / Pr, 4, 8, 8, mtol=mt, mtol_newton=mt, iprint=[])
    try:
        T, u, v, p, sw, r = gs(cd, ns, tol)
        print('gpu GS', mt, tol, sw, r, 'vs oracle-tight', rel(T, To), rel(u, uo), rel(v, vo), rel(p, po), flush=True)
    except Exception as e:
        print('gpu GS', mt, tol, 'FAILED', str(e)[:150], flush=True)
