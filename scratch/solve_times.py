"""Solve timings: NS C2 / C3 (reference defaults and tight), CD C1, larger CD / NS meshes; fdm vs jacobi preconditioner."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import sem_b200
from tests.conftest import load_golden, relerr
g = load_golden('ns')
pre = sys.argv[1].split(',') if len(sys.argv) > 1 else ['fdm', 'jacobi']
big = len(sys.argv) > 2
for precond in pre:
    for tag, kw in (('c3', dict(L_x=1.0, L_y=1.0, Re=1e3, Gr=1e3 / 0.71, P=4, N_ex=8, N_ey=8)), ('c2', dict(L_x=1, L_y=1, Re=400, Gr=0, P=4, N_ex=16, N_ey=16, u_N=1))):
        for tol in (None, 1e-13):
            kws = dict(kw)
            if tol: kws.update(mtol=tol, mtol_newton=tol)
            ns = sem_b200.NavierStokesSolver(iprint=[], precond=precond, **kws)
            T = g[f'{tag}/T_in']
            ns._get_solution(T)  # warm
            torch.cuda.synchronize(); t = time.perf_counter()
            u, v, p = ns._get_solution(T)
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            print(f"[{precond}] NS {tag} tol={tol}: {dt:.3f} s  newton {ns._k}  krylov its {ns.krylov_iters[-ns._k:]}  err u {relerr(u, g[f'{tag}/u_sol']):.2e} p {relerr(p, g[f'{tag}/p_sol']):.2e}", flush=True)
    gc = load_golden('cd')
    cd = sem_b200.ConvectionDiffusionSolver(1, 1, 40, 4, 16, 16, T_E=-0.5, T_W=0.5, precond=precond)
    for tol in (1e-7, 1e-13):
        cd._mtol = tol
        cd._get_solution(gc['c1/u'], gc['c1/v'])
        torch.cuda.synchronize(); t = time.perf_counter()
        T = cd._get_solution(gc['c1/u'], gc['c1/v'])
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"[{precond}] CD c1 tol={tol}: {dt:.4f} s its {cd.last_iters} err {relerr(T, gc['c1/T_sol']):.2e}", flush=True)
    # larger meshes
    for P, ne in (((4, 128), (8, 128)) if precond == 'jacobi' else ((4, 128), (8, 128), (8, 512))):
        cd = sem_b200.ConvectionDiffusionSolver(1, 1, 40, P, ne, ne, T_E=-0.5, T_W=0.5, mtol=1e-10, precond=precond, restart=60)
        u = cd._get_vector(lambda x, y: y - 0.5); v = cd._get_vector(lambda x, y: 0.5 - x)
        torch.cuda.synchronize(); t = time.perf_counter()
        try:
            T = cd._get_solution(u, v)
            t1 = time.perf_counter() - t
            t = time.perf_counter(); T = cd._get_solution(u, v); t2 = time.perf_counter() - t
            print(f"[{precond}] CD P={P} ne={ne} N={cd.N}: first {t1:.3f} s, again {t2:.3f} s its {cd.last_iters} res {cd.last_resnorm:.2e} restart {cd._restart}", flush=True)
        except RuntimeError as e:
            print(f"[{precond}] CD P={P} ne={ne}: {e} after {time.perf_counter()-t:.2f}s res {cd.last_resnorm:.2e} restart {cd._restart}", flush=True)
    if precond == 'fdm':
        for P, ne, Re in ((4, 128, 400.0), (8, 128, 400.0)) + (((8, 288, 400.0),) if big else ()):
            ns = sem_b200.NavierStokesSolver(1, 1, Re, 0, P, ne, ne, u_N=1, iprint=[], precond=precond, restart=200)
            T0 = np.zeros(ns.N)
            torch.cuda.synchronize(); t = time.perf_counter()
            try:
                ns._get_solution(T0)
                print(f"[{precond}] NS lid Re={Re:g} P={P} ne={ne} DOF={3*ns.N}: {time.perf_counter()-t:.2f} s newton {ns._k} krylov {ns.krylov_iters[-ns._k:]}", flush=True)
            except RuntimeError as e:
                print(f"[{precond}] NS P={P} ne={ne}: {e} after {time.perf_counter()-t:.2f}s its {ns.krylov_iters[-3:]}", flush=True)
