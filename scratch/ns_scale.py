"""Steady lid-driven cavity (NS example physics: Re = 400, u_N = 1, T = 0) on larger meshes.
usage: python scratch/ns_scale.py P ne precond [Re] [mtol] [mtol_newton] [restart]"""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import torch
import sem_b200
P, ne, precond = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
Re = float(sys.argv[4]) if len(sys.argv) > 4 else 400.0
mtol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-7
mtol_newton = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-5
restart = int(sys.argv[7]) if len(sys.argv) > 7 else None
t0 = time.time()
ns = sem_b200.NavierStokesSolver(1., 1., Re, 0., P, ne, ne, u_N=1., mtol=mtol, mtol_newton=mtol_newton, precond=precond,
                                 restart=restart, iprint=['NEWTON_iter', 'NEWTON_suc'])
T = np.zeros(ns.N)
t1 = time.time()
u, v, p = ns._get_solution(T)
torch.cuda.synchronize()
t2 = time.time()
ru, rv, rc = ns._get_residuals(u, v, p, T)
rn = np.sqrt(np.linalg.norm(ru) ** 2 + np.linalg.norm(rv) ** 2 + np.linalg.norm(rc) ** 2)
print(json.dumps(dict(P=P, ne=ne, N=ns.N, Re=Re, precond=precond, mtol=mtol, mtol_newton=mtol_newton, newton=ns._k,
                      krylov=ns.krylov_iters, total_krylov=sum(ns.krylov_iters), setup_s=t1 - t0, solve_s=t2 - t1,
                      res_frobenius=rn, res_limit=mtol_newton * np.sqrt(3 * ns.N), restart=ns._restart,
                      mem_gb=torch.cuda.max_memory_allocated() / 2**30)), flush=True)
