import sys
sys.path.insert(0, '.')
import numpy as np
import sem_b200
from sem_b200 import Boussinesq_SequentialCoupler as bsc
from tests.conftest import load_golden, relerr
g = load_golden("boussinesq_c3")
Re, Ra, Pr = 1e3, 1e3, 0.71
for precond in sys.argv[1:] or ["fdm", "fdm+bb", "full"]:
    cd = sem_b200.ConvectionDiffusionSolver(1., 1., Re * Pr, 4, 8, 8, T_W=0.5, T_E=-0.5, mtol=1e-13)
    ns = sem_b200.NavierStokesSolver(1., 1., Re, Ra / Pr, 4, 8, 8, mtol=1e-13, mtol_newton=1e-13, iprint=[], precond=precond)
    T, u, v, p, info = bsc.solve(cd, ns, mode='JNK', mtol_nonlin=1e-11, mtol_gmres=1e-13, iprint=True)
    print(precond, 'T', relerr(T, g["T"]), 'u', relerr(u, g["u"]), 'v', relerr(v, g["v"]), info['gmres_its'], 'ns krylov', sum(ns.krylov_iters), len(ns.krylov_iters), flush=True)
