"""Worker of tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU): the partitioned NCCL
path against the single-GPU path and the oracle, on identical inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import sem_b200  # noqa: E402
from sem_b200.partition import Partition  # noqa: E402
from oracle import sem_oracle as so  # noqa: E402  (checker)


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fails = []

    def check(name, val, tol):
        if not (val < tol):
            fails.append(f"rank {rank}: {name} = {val:.3e} (tol {tol:.0e})")

    # ---- CD: applies, Jacobian, solve --------------------------------------------------------------------------------
    P, nx, ny = 4, max(8, world + 1), 6          # world = 3: slabs of unequal width, a middle rank with two neighbours
    kw = dict(L_x=1.5, L_y=1.0, Pe=25.0, P=P, N_ex=nx, N_ey=ny, T_W=0.5, T_E=-0.5, T_S=0.1)
    part = Partition(nx, ny, P, rank, world)
    cd = sem_b200.ConvectionDiffusionSolver(mtol=1e-12, device=local, partition=(rank, world), **kw)
    cd_o = so.CDOracle(mtol=1e-12, **kw)
    rng = np.random.default_rng(3)
    T, u, v, dT, du, dv = (rng.standard_normal(cd_o.N) for _ in range(6))
    sl = part.local_slice
    check("cd residual", relerr(cd._get_residuals(sl(T), sl(u), sl(v)), sl(cd_o._get_residuals(T, u, v))), 1e-12)
    # host vector in / out without du, dv: the pipelined upload / apply / download path with the interface exchange at its end
    check("cd jvp (host pipeline)", relerr(cd._get_dresiduals(sl(dT)), sl(cd_o._get_dresiduals(dT))), 1e-12)
    cd._calc_jacobians(sl(T))
    cd_o._calc_jacobians(T)
    check("cd jvp", relerr(cd._get_dresiduals(sl(dT), sl(du), sl(dv)), sl(cd_o._get_dresiduals(dT, du, dv))), 1e-12)
    uu = cd_o._get_vector(lambda x, y: y - 0.5)
    vv = cd_o._get_vector(lambda x, y: 0.75 - x)
    check("cd solve", relerr(cd._get_solution(sl(uu), sl(vv)), sl(cd_o._get_solution(uu, vv))), 1e-8)
    check("cd points", float(np.abs(cd.points - np.stack([sl(cd_o.points[0]), sl(cd_o.points[1])])).max()), 1e-15)

    # ---- NS: applies and a Newton solve --------------------------------------------------------------------------------
    nex_ns = max(4, world + world % 2)
    nkw = dict(L_x=1.0, L_y=1.0, Re=50.0, Gr=200.0, P=3, N_ex=nex_ns, N_ey=4, u_N=1.0)
    part = Partition(nex_ns, 4, 3, rank, world)
    sl = part.local_slice
    ns = sem_b200.NavierStokesSolver(mtol=1e-13, mtol_newton=1e-13, iprint=[], device=local, partition=(rank, world), **nkw)
    ns_o = so.NSOracle(mtol=1e-13, mtol_newton=1e-13, **nkw)
    a, b, c, d, e = (rng.standard_normal(ns_o.N) for _ in range(5))
    for x, y in zip(ns._get_residuals(sl(a), sl(b), sl(c), sl(d)), ns_o._get_residuals(a, b, c, d)):
        check("ns residual", relerr(x, sl(y)), 1e-12)
    ns._calc_jacobians(sl(a), sl(b))
    ns_o._calc_jacobians(a, b)
    for x, y in zip(ns._get_dresiduals(sl(c), sl(d), sl(e), sl(a)), ns_o._get_dresiduals(c, d, e, a)):
        check("ns jvp", relerr(x, sl(y)), 1e-12)
    Tin = ns_o._get_vector(lambda x, y: 0.5 - x)
    us, vs, ps = ns._get_solution(sl(Tin))
    uo, vo, po = ns_o._get_solution(Tin)
    check("ns solve u", relerr(us, sl(uo)), 1e-8)
    check("ns solve v", relerr(vs, sl(vo)), 1e-8)
    check("ns solve p", relerr(ps, sl(po)), 1e-8)
    if ns._k != ns_o._k:
        fails.append(f"rank {rank}: Newton its {ns._k} vs {ns_o._k}")
    # run() / _get_interpol on a partitioned solver: every rank gets the whole plot array (sem_interpolate + all-reduce)
    xp, yp = np.meshgrid(np.linspace(0, 1, 23), np.linspace(0, 1, 17), indexing='ij')
    check("partitioned interpolation", relerr(ns._get_interpol(us, (xp, yp)), ns_o._get_interpol(uo, (xp, yp))), 1e-8)

    # ---- the partitioned NS preconditioner (distributed fast-diagonalisation plans, ring Chebyshev with interface exchange,
    #      all-reduced projector coefficients and member dots): every stage against the CPU mirror, then a Newton solve with it
    from oracle.ns_precond import NSPrecondMirror
    # 6 element columns per rank: wide enough for the one-launch (in-kernel exchange) applies inside the stages; the second
    # mesh has slabs of 1-2 columns (exchange kernels)
    for (Pq, nxq, nyq, Re) in ((4, 6 * world, 6, 80.0), (3, nex_ns, 4, 40.0)):
        qkw = dict(L_x=1.2, L_y=0.9, Re=Re, Gr=0.0, P=Pq, N_ex=nxq, N_ey=nyq, u_N=1.0, v_W=0.2)
        part = Partition(nxq, nyq, Pq, rank, world)
        sl = part.local_slice
        nsq = sem_b200.NavierStokesSolver(mtol=1e-13, mtol_newton=1e-13, iprint=[], device=local, partition=(rank, world),
                                          precond='full', **qkw)
        nsq_o = so.NSOracle(mtol=1e-13, mtol_newton=1e-13, **qkw)
        N = nsq_o.N
        uq, vq, pq, Tq = (0.3 * rng.standard_normal(N) for _ in range(4))
        nsq._get_residuals(sl(uq), sl(vq), sl(pq), sl(Tq))
        nsq._calc_jacobians(sl(uq), sl(vq))
        nsq_o._get_residuals(uq, vq, pq, Tq)
        nsq_o._calc_jacobians(uq, vq)
        m = NSPrecondMirror(nsq_o)
        aa, bb = rng.standard_normal(N), rng.standard_normal(N)
        tag = f"P{Pq} {nxq}x{nyq}"
        ref1 = np.where(m.pin == np.arange(N), aa, m.coarse(aa))
        check(f"{tag} coarse", relerr(nsq._precond_debug(1, 4, sl(aa)), sl(ref1)), 1e-9)
        zin = np.where(m.inner, bb, 0.0)
        zin[m.pin] = aa[m.pin]
        check(f"{tag} ring", relerr(nsq._precond_debug(2, 4, sl(aa), sl(zin)), sl(m.add_ring(zin, aa))), 1e-9)
        check(f"{tag} stokes residual", relerr(nsq._precond_debug(3, 4, sl(aa), sl(zin)), sl(aa - m.schur_stokes(zin))), 1e-9)
        ref4 = m.pcd(aa)
        ref4[m.pin] = aa[m.pin]
        check(f"{tag} pcd", relerr(nsq._precond_debug(4, 4, sl(aa)), sl(ref4)), 1e-9)
        r3 = tuple(rng.standard_normal(N) for _ in range(3))
        for level in (2, 3, 4):
            z3 = nsq._precond_debug(0, level, tuple(sl(x) for x in r3))
            ref = np.split(m.apply(np.hstack(r3), level), 3)
            check(f"{tag} precond level {level}", max(relerr(x, sl(y)) for x, y in zip(z3, ref)), 1e-9)
        Tin = nsq_o._get_vector(lambda x, y: 0.0 * x)
        us, vs, ps = nsq._get_solution(sl(Tin))
        uo, vo, po = nsq_o._get_solution(Tin)
        check(f"{tag} ns solve (full preconditioner)", max(relerr(us, sl(uo)), relerr(vs, sl(vo)), relerr(ps, sl(po))), 1e-8)

    # ---- larger stiffness apply: partitioned result == single-GPU result, bitwise away from / to rounding at the interface
    P, ne = 8, 16 * world
    dall = sem_b200.SemDevice(P, ne, ne, 1.0 / ne, 1.0 / ne, device=local)
    dpar = sem_b200.SemDevice(P, ne, ne, 1.0 / ne, 1.0 / ne, device=local, partition=(rank, world))
    xg = np.random.default_rng(11).standard_normal(dall.N_local)
    yg = dall.to_host(dall.apply_stiffness(dall.to_device(xg), dall.zeros()))
    yl = dpar.to_host(dpar.apply_stiffness(dpar.to_device(dpar.part.local_slice(xg)), dpar.zeros()))
    check("partitioned K apply", relerr(yl, dpar.part.local_slice(yg)), 1e-13)
    xl = dpar.to_device(dpar.part.local_slice(xg))
    check("global dot", abs(dpar.dot(xl, xl) - float(xg @ xg)) / float(xg @ xg), 1e-13)

    # ---- a mid-size CD solve with the distributed fast-diagonalisation preconditioner (GEMM + reduce-scatter, all-gather +
    #      GEMM): the exact inverse of the Laplacian on any partition -- same answer, same iteration count as on one GPU
    kw = dict(L_x=1.0, L_y=1.0, Pe=40.0, P=8, N_ex=32 * world, N_ey=48, T_W=0.5, T_E=-0.5, mtol=1e-11)
    part = Partition(kw["N_ex"], kw["N_ey"], 8, rank, world)
    sl = part.local_slice
    cd1 = sem_b200.ConvectionDiffusionSolver(device=local, restart=200, **kw)
    uu = cd1._get_vector(lambda x, y: y - 0.5)
    vv = cd1._get_vector(lambda x, y: 0.5 - x)
    T1 = cd1._get_solution(uu, vv)
    cdp = sem_b200.ConvectionDiffusionSolver(device=local, partition=(rank, world), restart=200, **kw)
    Tp = cdp._get_solution(sl(uu), sl(vv))
    check("partitioned fdm solve", relerr(Tp, sl(T1)), 1e-8)
    if not (cdp.last_iters <= cd1.last_iters + 2):
        fails.append(f"rank {rank}: distributed FDM took {cdp.last_iters} iterations (whole mesh: {cd1.last_iters})")
    xh = np.random.default_rng(5).standard_normal(cd1.N)
    cd1._get_residuals(xh, uu, vv)
    cdp._get_residuals(sl(xh), sl(uu), sl(vv))
    check("partitioned host-pipeline jvp (several segments)", relerr(cdp._get_dresiduals(sl(xh)), sl(cd1._get_dresiduals(xh))), 1e-13)
    if rank == 0:
        print(f"cd solve {cd1.N} nodes: one GPU {cd1.last_iters} its, {world} GPUs {cdp.last_iters} its")
        print(f"interface exchange: {dpar.comm_mode}")

    # ---- many back-to-back exchanges (graph replays included): the mailbox epochs / parities must never slip.  The K apply
    #      is linear, so y_k = K^k x (rescaled) on the partition must track the single-GPU sequence.
    xs, xp = dall.to_device(xg), dpar.to_device(dpar.part.local_slice(xg))
    ys, yp = dall.zeros(), dpar.zeros()
    for k in range(40):
        dall.apply_stiffness(xs, ys)
        dpar.apply_stiffness(xp, yp)
        s = 1.0 / float(torch.linalg.vector_norm(ys))
        xs, ys = ys.mul_(s), xs
        xp, yp = yp.mul_(s), xp
    check("40 chained partitioned K applies", relerr(dpar.to_host(xp), dpar.part.local_slice(dall.to_host(xs))), 1e-11)

    # which partitioned-apply path ran: with peer-memory mailboxes the K / G / DIV / CD applies are one launch (in-kernel exchange)
    nfused, nsplit = (int(dpar.lib.sem_ctx_partitioned_applies(dpar.ctx, k)) for k in (1, 0))
    if rank == 0:
        print(f"partitioned K applies: {nfused} one-launch, {nsplit} three-launch")
    want_fused = dpar.comm_mode == "p2p" and os.environ.get("SEM_B200_FUSED_XCH", "1") != "0"
    if want_fused != (nfused > 0) or (want_fused and nsplit > 0):
        fails.append(f"rank {rank}: expected the {'one' if want_fused else 'three'}-launch path, counters {nfused}/{nsplit}")

    allf = [None] * world
    dist.all_gather_object(allf, fails)
    dist.destroy_process_group()
    flat = [f for fl in allf for f in fl]
    if rank == 0:
        print("MGPU_OK" if not flat else "MGPU_FAIL\n" + "\n".join(flat))
    sys.exit(1 if flat else 0)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:                      # the launcher's own error report buries a worker's stderr: say it on stdout
        import traceback
        print(f"MGPU_EXC rank {os.environ.get('RANK')}:\n{traceback.format_exc()[-1500:]}", flush=True)
        raise
