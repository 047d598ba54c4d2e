import sys, time
sys.path.insert(0, '.')
import numpy as np, torch, sem_b200
ns = sem_b200.NavierStokesSolver(1, 1, 400, 0, 4, 16, 16, u_N=1, iprint=[])
T0 = np.zeros(ns.N)
t = time.perf_counter(); ns._get_solution(T0); torch.cuda.synchronize(); print("solve", time.perf_counter() - t, ns.krylov_iters)
