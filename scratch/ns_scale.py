"""Steady lid-driven cavity (NS example physics: Re = 400, u_N = 1, T = 0) on larger meshes.
usage: python scratch/ns_scale.py P ne precond [Re] [mtol] [mtol_newton] [restart]"""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import torch
import os
import sem_b200
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
P, ne, precond = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
Re = float(sys.argv[4]) if len(sys.argv) > 4 else 400.0
mtol = float(sys.argv[5]) if len(sys.argv) > 5 else 1e-7
mtol_newton = float(sys.argv[6]) if len(sys.argv) > 6 else 1e-5
restart = int(sys.argv[7]) if len(sys.argv) > 7 else None
t0 = time.time()
ns = sem_b200.NavierStokesSolver(1., 1., Re, 0., P, ne, ne, u_N=1., mtol=mtol, mtol_newton=mtol_newton, precond=precond,
                                 restart=restart, iprint=['NEWTON_iter'] if rank == 0 else [], device=local,
                                 **({"partition": (rank, world)} if world > 1 else {}))
T = np.zeros(ns._dev.N_local)
t1 = time.time()
u, v, p = ns._get_solution(T)
torch.cuda.synchronize()
t2 = time.time()
d = ns._dev
st3, res = d.zeros(3), d.zeros(3)
for k, a in enumerate((u, v, p)):
    d.to_device(a, st3[k])
ns._residual_dev(st3[0], st3[1], st3[2], d.to_device(T, ns._in[3]), res)
rn = float(np.sqrt(d.dot(res, res)))
if rank == 0:
  print(json.dumps(dict(P=P, ne=ne, N=ns.N, Re=Re, precond=precond, mtol=mtol, mtol_newton=mtol_newton, newton=ns._k,
                      krylov=ns.krylov_iters, total_krylov=sum(ns.krylov_iters), setup_s=t1 - t0, solve_s=t2 - t1,
                      res_frobenius=rn, res_limit=mtol_newton * np.sqrt(3 * ns.N), restart=ns._restart,
                      mem_gb=torch.cuda.max_memory_allocated() / 2**30, n_gpus=world)), flush=True)
if world > 1:
    dist.destroy_process_group()
