"""C3 (8x8, P=4, Re=1000) pressure floor: GPU solution at several tolerances against the ORACLE converged to 5e-15 (direct
solves), not against the reference's stored fields (which stop at the reference's own floor, 6e-8 in p)."""
import sys
sys.path.insert(0, '.')
import numpy as np
import sem_b200
from oracle import sem_oracle as so
from tests.golden.make_golden_cases import NS_CASES
g = np.load('tests/golden/ns.npz')
kw = [c for c in NS_CASES if c[0] == 'c3'][0][1]
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
o = so.NSOracle(mtol=1e-15, mtol_newton=1e-15, **kw)
uo, vo, po = o._get_solution(g['c3/T_in'])
for precond in ('fdm', 'full'):
    for mt in (1e-13, 1e-14, 1e-15):
        try:
            ns = sem_b200.NavierStokesSolver(mtol=mt, mtol_newton=mt, iprint=[], precond=precond, **kw)
            u, v, p = ns._get_solution(g['c3/T_in'])
            print(precond, mt, 'newton', ns._k, 'krylov', sum(ns.krylov_iters), 'vs oracle', rel(u, uo), rel(v, vo), rel(p, po),
                  'vs ref', rel(p, g['c3/p_sol']), flush=True)
        except Exception as e:
            print(precond, mt, 'FAILED', str(e)[:150], flush=True)
