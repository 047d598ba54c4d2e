"""Steady NS solve of BASELINE config 2 (16x16 elements, P=4, Re=400, lid-driven): wall time, Krylov iterations, parity of
the converged fields with the committed golden vectors of the unmodified reference."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch, sem_b200
from tests.conftest import load_golden
g = load_golden('ns')
for rep in range(3):
    ns = sem_b200.NavierStokesSolver(1, 1, 400, 0, 4, 16, 16, u_N=1, iprint=[])
    T0 = np.zeros(ns.N)
    torch.cuda.synchronize()
    t = time.perf_counter(); u, v, p = ns._get_solution(T0); torch.cuda.synchronize()
    print("solve", round(time.perf_counter() - t, 4), "s", ns.krylov_iters, flush=True)
ns = sem_b200.NavierStokesSolver(1, 1, 400, 0, 4, 16, 16, u_N=1, mtol=1e-13, mtol_newton=1e-13, iprint=[])
t = time.perf_counter(); u, v, p = ns._get_solution(np.zeros(ns.N)); torch.cuda.synchronize()
print("tight solve", round(time.perf_counter() - t, 4), "s", ns.krylov_iters)
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
print("rel err vs reference golden: u %.2e v %.2e p %.2e" % (rel(u, g['c2/u_sol']), rel(v, g['c2/v_sol']), rel(p, g['c2/p_sol'])))
