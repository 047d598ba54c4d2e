"""bench.py -- headline benchmark of the sem_b200 hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): "GDOF/s element apply".  One step = one fused matrix-free application of the convection-
diffusion Jacobian (`ConvectionDiffusionSolver._get_dresiduals`, reference CD:104-121: K + Pe(diag(u)G_x + diag(v)G_y)
with Dirichlet rows) on BASELINE config 5: synthetic mesh of 1024 x 1024 elements, P = 8, 67.1M nodes per field.
`value` = nodes processed by all ranks / device time with inputs resident in HBM; `e2e` = the same apply through the
public Python class with pinned HOST buffers (H2D of the input field and D2H of the result inside the timed region).
For N > 1 the element columns are partitioned across the ranks (strong scaling of the fixed config-5 mesh) with an
NCCL exchange of the interface node lines.

Secondary numbers of the same run go to `extra` (BASELINE.json metric part ii and configs 1-4): a steady Navier-Stokes solve
of a 19.7M-DOF lid-driven cavity partitioned over the N ranks (`extra.ns_solve`, every N), the NS / CD / Boussinesq examples
at their own sizes with the reference's algorithm timed on the host IN THE SAME RUN (N = 1), a weak-scaling apply, and -- for
N > 1 -- a parity field: the partitioned applies against the same applies on one GPU.

`--impl reference` times the reference's own CPU implementation of this apply -- scipy CSR mat-vec on the host, on
matrices value-identical to the reference's (oracle port, the reference itself cannot build this mesh: SURVEY 8c) --
on a bounded sample mesh.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "GDOF/s element apply (CD Jacobian-vector product, reference CD:104-121)"
UNIT = "GDOF/s"
P_ORDER = 8
NE = 1024                 # config 5: 1024 x 1024 elements
PE = 40.0
CPU_SAMPLE_NE = 128       # bounded CPU sample: 128 x 128 elements, P = 8 (1.05M nodes)
ALG_BYTES_PER_NODE = 32   # fp64: read dT, u, v, write dres


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason samples while the timed region runs.  NVML in-process (the same counters nvidia-smi
    prints): forking nvidia-smi from a thread of a process that holds NCCL communicators can dead-lock in fork()."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml is None:
            return
        n = self.nvml
        while not self._halt.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(
                    n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((int(sm), int(mx), int(rs)))
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(r[0] for r in self.rows)
        mx = [r[1] for r in self.rows]
        reasons = sorted({name for r in self.rows for name, bit in self.REASONS if r[2] & bit})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": "nvml"}


def cpu_apply_sample(steps, warmup):
    """The reference's CPU path for this apply: CSR mat-vec + Dirichlet rows (CD:112-119) on the sample mesh."""
    import numpy as np
    from oracle import sem_oracle as so
    ne = CPU_SAMPLE_NE
    cd = so.CDOracle(1.0, 1.0, PE, P_ORDER, ne, ne, T_W=0.5, T_E=-0.5)
    u = cd._get_vector(lambda x, y: y - 0.5)
    v = cd._get_vector(lambda x, y: 0.5 - x)
    rng = np.random.default_rng(0)
    dT = rng.standard_normal(cd.N)
    cd._get_residuals(dT, u, v)          # builds Sys like the reference (outside the timed region)
    for _ in range(warmup):
        cd._get_dresiduals(dT)
    t0 = time.perf_counter()
    for _ in range(steps):
        cd._get_dresiduals(dT)
    dt = (time.perf_counter() - t0) / steps
    return cd.N / dt / 1e9, dt, cd.N


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)             # exactly K timed steps, like the GPU arm
    val, dt, n = cpu_apply_sample(steps, max(args.warmup, 3))
    cores = len(os.sched_getaffinity(0))
    sample = f"{CPU_SAMPLE_NE}x{CPU_SAMPLE_NE} elements, P={P_ORDER} ({n} nodes), scipy CSR mat-vec, {steps} applies"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(), "note": "CPU arm runs the bounded sample mesh: " + sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "host_cores_available": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_name():
    return (f"BASELINE config 5: synthetic {NE}x{NE} elements, P={P_ORDER} ({(NE * P_ORDER + 1) ** 2} nodes/field), "
            f"fused CD Jacobian-vector apply, Pe={PE:g}, Dirichlet W/E; inputs (1.6 GB) larger than L2")


def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    import sem_b200
    from sem_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    steps, warmup = args.steps, max(args.warmup, 3)
    cd = sem_b200.ConvectionDiffusionSolver(1.0, 1.0, PE, P_ORDER, NE, NE, T_W=0.5, T_E=-0.5, device=local,
                                            **({"partition": (rank, world)} if world > 1 else {}))
    d = cd._dev
    n_local = d.NX * d.NY
    n_global = (NE * P_ORDER + 1) ** 2
    gen = torch.Generator(device=d.tdev).manual_seed(rank)

    def rnd():
        x = d.zeros()
        x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64)
        return x

    dT, out = rnd(), d.zeros()
    cd._u.copy_(rnd())
    cd._v.copy_(rnd())
    cd._have_sys = True
    st = cd._state(with_jac=False)
    lib = d.lib

    def step():
        # for world > 1 the call ends with the exchange of the interface node lines (peer-memory mailboxes)
        L.check(lib.sem_cd_jvp(d.ctx, C.byref(st), dT.data_ptr(), None, None, out.data_ptr(), d.stream), "sem_cd_jvp")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    # the clock sampler (NVML, rank 0 only) starts BEFORE the barrier: started after it, rank 0 entered the timed region
    # late and the other ranks' first exchange waited for it inside their timed region
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fused0 = lib.sem_ctx_partitioned_applies(d.ctx, 1)
    if world > 1:
        # the host barrier above releases the ranks up to ~100 us apart, which a 20-step region of ~0.1 ms steps does not
        # amortise (every step waits for the neighbours): line the DEVICES up with an in-stream all-reduce right before
        # the start event, so the timed region is K steps from a common start on every rank
        dist.all_reduce(torch.zeros(1, device=d.tdev))
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync_all()
    fused_steps = lib.sem_ctx_partitioned_applies(d.ctx, 1) - fused0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=d.tdev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # a short run leaves the clock sampler without a sample under load: keep the same load up for ~1.5 s more.
    # The count is derived from the all-reduced time so that EVERY rank issues the same number of (collective) steps.
    extra_steps = int(min(5000, max(0, 1500.0 / max(ms / steps, 1e-3))))
    for _ in range(extra_steps):
        step()
    sync_all()
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / steps
    value = n_global / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public class with pinned host buffers ------------------------------------------------------
    e2e_steps = max(2, min(steps, 5))
    host_in = torch.empty(n_local, dtype=torch.float64).pin_memory().numpy()
    host_in[:] = np.random.default_rng(rank).standard_normal(n_local)
    for _ in range(3):                       # warm-up: the pinned result blocks are allocated once and recycled
        res_host = cd._get_dresiduals(host_in)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res_host = cd._get_dresiduals(host_in)
    sync_all()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=d.tdev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.isfinite(res_host).all()
    # the same call with an ordinary (pageable) numpy input, what OpenMDAO hands the solvers (views into its root vectors)
    pageable_in = np.random.default_rng(rank + 100).standard_normal(n_local)
    cd._get_dresiduals(pageable_in)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res_host = cd._get_dresiduals(pageable_in)
    sync_all()
    e2e_pageable_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_pageable_s], device=d.tdev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pageable_s = float(t.item())
    # The host link under the same load, without any of our code: every rank copies its pinned slab up and another pinned
    # slab down concurrently (full duplex), all ranks at once.  e2e can at best move 2 x n_local x 8 bytes in that time.
    link = None
    try:
        up_dev = torch.empty(n_local, dtype=torch.float64, device=d.tdev)
        dn_dev = torch.randn(n_local, dtype=torch.float64, device=d.tdev)
        dn_host = torch.empty(n_local, dtype=torch.float64).pin_memory()
        up_host = torch.from_numpy(host_in)
        s_up, s_dn = torch.cuda.Stream(d.tdev), torch.cuda.Stream(d.tdev)

        def duplex():
            with torch.cuda.stream(s_up):
                up_dev.copy_(up_host, non_blocking=True)
            with torch.cuda.stream(s_dn):
                dn_host.copy_(dn_dev, non_blocking=True)

        duplex()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(3):
            duplex()
        sync_all()
        link_s = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([link_s], device=d.tdev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            link_s = float(t.item())
        link = {"ms_per_duplex_copy": link_s * 1e3, "gb_s_each_way_per_rank": n_local * 8 / link_s / 1e9,
                "gb_s_each_way_aggregate": world * n_local * 8 / link_s / 1e9,
                "e2e_ceiling_gdof_s": n_global / link_s / 1e9,
                "note": "torch pinned copies of one slab up and one down on two streams, all ranks at once; no sem_b200 code"}
        del up_dev, dn_dev, dn_host, up_host
    except Exception as exc:                          # a probe must not lose the headline line
        link = {"error": str(exc)[:200]}
    del res_host, pageable_in, host_in

    parity = partition_parity(sem_b200, rank, world, local) if world > 1 else None

    # ---- N > 1: a steady CD solve of the whole config-5 mesh (67.1 M nodes), partitioned, distributed FDM preconditioner ------
    part_solve = None
    if world > 1 and not args.no_extra:
        try:
            del dT, out
            torch.cuda.empty_cache()
            cd._mtol, cd._restart = 1e-8, 100
            xs, ys = cd.points[0], cd.points[1]
            ub, vb = ys - 0.5, 0.5 - xs
            cd._get_solution(ub, vb)                     # includes the one-off eigen-decompositions of the pencils
            sync_all()
            t0 = time.perf_counter()
            cd._get_solution(ub, vb)
            sync_all()
            part_solve = {"wall_s": time.perf_counter() - t0, "krylov_its": cd.last_iters, "nodes": n_global,
                          "resnorm": cd.last_resnorm, "atol": 1e-8 * float(np.sqrt(n_global)), "restart": 100,
                          "note": "config-5 mesh partitioned over the ranks; distributed fast diagonalisation (GEMM + "
                                  "reduce-scatter, all-gather + GEMM); host slabs in / out included"}
        except Exception as exc:                         # a solver failure must not lose the headline line
            part_solve = {"error": str(exc)[:200]}

    shared_extra = {}
    if not args.no_extra:
        del cd
        torch.cuda.empty_cache()
        shared_extra["ns_solve"] = ns_solve_large(sem_b200, rank, world, local, sync_all)
        shared_extra["weak_scaling_apply"] = weak_scaling_apply(sem_b200, rank, world, local, sync_all, steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    achieved = ALG_BYTES_PER_NODE * n_local / (ms_per_step * 1e-3) / 1e9
    traffic = None                                   # per launch of THIS line: known from the ncu capture of the whole mesh only
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("cd_jvp_dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(), "partition": f"{world} strips of element columns" if world > 1 else "none",
                   "interface_exchange": {"p2p": ("inside the operator kernel: edge CTAs store their segments of the interface "
                                                  "lines into the neighbour's mailbox over NVLink (per-strip epoch flags) "
                                                  "and add what arrived; one launch per apply, no NCCL on the apply path"
                                                  if fused_steps == steps else
                                                  "peer-memory mailboxes over NVLink (push kernel + epoch flag), no NCCL "
                                                  "on the apply path"),
                                          "nccl": "ncclSend/ncclRecv", "none": "none"}[d.comm_mode],
                   "l2": "inputs larger than L2 (3 x 537 MB read per step)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "sem_march3_kernel<8, MODE_CD, false>", "peak_source": peak_src,
                     "algorithmic_bytes_per_node": ALG_BYTES_PER_NODE, "nodes_per_launch": n_local},
        "e2e": {"value": n_global / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n_local * 8),
                "d2h_bytes_per_step": int(n_local * 8), "ms_per_step": e2e_s * 1e3,
                "api": "ConvectionDiffusionSolver._get_dresiduals(numpy pinned) -> numpy",
                "host_link": link,
                "pageable_input": {"value": n_global / e2e_pageable_s / 1e9, "ms_per_step": e2e_pageable_s * 1e3,
                                   "note": "same call with an ordinary numpy array (what OpenMDAO passes); result still lands "
                                           "in a recycled page-locked block"}},
        # one operator kernel per step; a partitioned apply that does not take the one-launch path is left edge + right edge +
        # interior + exchange kernel (replayed as one CUDA graph)
        "gpu_launches": steps if (world == 1 or fused_steps == steps) else 4 * steps,
        "clocks": clocks,
    }
    if parity is not None:
        line["parity"] = parity
    line["extra"] = dict(shared_extra)
    if part_solve is not None:
        line["extra"]["cd_solve_config5_partitioned"] = part_solve
    if world == 1:
        cval, cdt, cn = cpu_apply_sample(20, 3)
        line["cpu_baseline"] = {"value": cval, "unit": UNIT, "cores": 1,
                                "host_cores_available": len(os.sched_getaffinity(0)), "kind": "port",
                                "sample": f"{CPU_SAMPLE_NE}x{CPU_SAMPLE_NE} elements, P={P_ORDER} ({cn} nodes), scipy CSR "
                                          f"mat-vec of the reference-identical Sys matrix, 20 applies, {cdt * 1e3:.1f} ms each"}
        if not args.no_extra:
            line["extra"].update(extra_numbers(sem_b200, local))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


NS_BIG_NE = 320           # extra.ns_solve: 320 x 320 elements, P = 8 -> 2561^2 nodes/field, 19.7M DOF (north star: >= 16M DOF)


def ns_solve_large(sem_b200, rank, world, local, sync_all):
    """Steady incompressible NS (NS example physics: lid-driven cavity, Re = 400, T = 0; reference defaults mtol = 1e-7,
    mtol_newton = 1e-5) on NS_BIG_NE^2 elements of order 8, element columns partitioned over the ranks: distributed
    fast-diagonalisation plans, two-level Schur preconditioner, NVLink interface exchange, all-reduced Krylov dots."""
    import numpy as np
    import torch
    try:
        ne = NS_BIG_NE
        t0 = time.perf_counter()
        ns = sem_b200.NavierStokesSolver(1.0, 1.0, 400.0, 0.0, P_ORDER, ne, ne, u_N=1.0, iprint=[], device=local, restart=600,
                                         **({"partition": (rank, world)} if world > 1 else {}))
        T = np.zeros(ns._dev.N_local)
        ns._krylov()                                      # plans, Schur tables (one-off set-up, reported separately)
        sync_all()
        t1 = time.perf_counter()
        u, v, p = ns._get_solution(T)
        sync_all()
        t2 = time.perf_counter()
        d = ns._dev
        res = d.zeros(3)
        st3 = d.zeros(3)
        for k, a in enumerate((u, v, p)):
            d.to_device(a, st3[k])
        ns._residual_dev(st3[0], st3[1], st3[2], d.to_device(T, ns._in[3]), res)
        rn = ns._spectral_norm(res)
        out = {"wall_s": t2 - t1, "setup_s": t1 - t0, "newton_its": ns._k, "krylov_its": list(ns.krylov_iters),
               "nodes_per_field": ns.N, "dof": 3 * ns.N, "mesh": f"{ne}x{ne} elements, P={P_ORDER}", "Re": 400.0,
               "newton_residual": rn, "newton_limit": ns._mtol_newton * float(np.sqrt(3 * ns.N)),
               "converged": bool(rn <= ns._mtol_newton * np.sqrt(3 * ns.N)), "precond": "full (two-level Schur + PCD)",
               "peak_mem_gb_rank0": torch.cuda.max_memory_allocated() / 2 ** 30, "n_gpus": world,
               "tolerances": "reference defaults (mtol 1e-7, mtol_newton 1e-5)"}
        del ns, res, st3
        torch.cuda.empty_cache()
        return out
    except Exception as exc:                              # a solver failure must not lose the headline line
        return {"error": str(exc)[:300]}


def weak_scaling_apply(sem_b200, rank, world, local, sync_all, steps):
    """Weak scaling of the headline apply: config 5 PER GPU (global mesh 1024 N x 1024 elements, one 1024-column slab per rank)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from sem_b200 import _lib as L
    try:
        cd = sem_b200.ConvectionDiffusionSolver(float(world), 1.0, PE, P_ORDER, NE * world, NE, T_W=0.5, T_E=-0.5, device=local,
                                                **({"partition": (rank, world)} if world > 1 else {}))
        d = cd._dev
        gen = torch.Generator(device=d.tdev).manual_seed(rank)
        x, y = d.zeros(), d.zeros()
        for t in (x, cd._u, cd._v):
            t[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64)
        cd._have_sys = True
        st = cd._state(with_jac=False)
        step = lambda: L.check(d.lib.sem_cd_jvp(d.ctx, C.byref(st), x.data_ptr(), None, None, y.data_ptr(), d.stream), "sem_cd_jvp")
        for _ in range(5):
            step()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=d.tdev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        n_global = (NE * world * P_ORDER + 1) * (NE * P_ORDER + 1)
        out = {"gdof_s": n_global / ms / 1e6, "ms_per_step": ms, "nodes": n_global,
               "mesh": f"{NE * world}x{NE} elements, P={P_ORDER}: config 5 per GPU"}
        del cd, x, y
        torch.cuda.empty_cache()
        return out
    except Exception as exc:
        return {"error": str(exc)[:300]}


def partition_parity(sem_b200, rank, world, local):
    """Correctness carried on every N > 1 line: the partitioned fused applies (CD Jacobian, 3-field NS Jacobian) on a mid-size mesh
    against the SAME applies on one GPU (each rank also holds the whole mesh; the one-GPU path is gated against the reference's
    golden vectors and the CPU restatement by tests/), relative L2 error of the slab, maximum over ranks.  Bar: 1e-12."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from sem_b200.partition import Partition
    try:
        P, nx, ny = P_ORDER, 32 * world, 96
        kw = dict(L_x=1.0, L_y=1.0, P=P, N_ex=nx, N_ey=ny)
        part = Partition(nx, ny, P, rank, world)
        sl = part.local_slice
        rng = np.random.default_rng(7)
        N = (nx * P + 1) * (ny * P + 1)
        T, u, v, p, a, b, c = (rng.standard_normal(N) for _ in range(7))
        errs = {}
        rel = lambda x, y: float(np.linalg.norm(x - y) / np.linalg.norm(y))
        cd1 = sem_b200.ConvectionDiffusionSolver(Pe=PE, T_W=0.5, T_E=-0.5, device=local, **kw)
        cdp = sem_b200.ConvectionDiffusionSolver(Pe=PE, T_W=0.5, T_E=-0.5, device=local, partition=(rank, world), **kw)
        cd1._get_residuals(T, u, v)
        cdp._get_residuals(sl(T), sl(u), sl(v))
        errs["cd_jvp"] = rel(cdp._get_dresiduals(sl(a)), sl(cd1._get_dresiduals(a)))
        ns1 = sem_b200.NavierStokesSolver(Re=400.0, Gr=10.0, u_N=1.0, iprint=[], device=local, **kw)
        nsp = sem_b200.NavierStokesSolver(Re=400.0, Gr=10.0, u_N=1.0, iprint=[], device=local, partition=(rank, world), **kw)
        r1 = ns1._get_residuals(u, v, p, T)
        rp = nsp._get_residuals(sl(u), sl(v), sl(p), sl(T))
        errs["ns_residual"] = max(rel(x, sl(y)) for x, y in zip(rp, r1))
        ns1._calc_jacobians(u, v)
        nsp._calc_jacobians(sl(u), sl(v))
        j1 = ns1._get_dresiduals(a, b, c)
        jp = nsp._get_dresiduals(sl(a), sl(b), sl(c))
        errs["ns_jvp"] = max(rel(x, sl(y)) for x, y in zip(jp, j1))
        t = torch.tensor([errs[k] for k in sorted(errs)], device=torch.device("cuda", local), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        errs = {k: float(t[i]) for i, k in enumerate(sorted(errs))}
        del cd1, cdp, ns1, nsp
        torch.cuda.empty_cache()
        return {"mesh": f"{nx}x{ny} elements, P={P} ({N} nodes)", "against": "the same applies on one GPU", "max_rel_err": errs,
                "bar": 1e-12, "ok": bool(max(errs.values()) <= 1e-12)}
    except Exception as exc:
        return {"error": str(exc)[:300]}


def extra_numbers(sem_b200, local):
    """Secondary numbers of the N = 1 run (not the headline): the other fused applies, the reference's examples (BASELINE
    configs 1-4) with the reference's own algorithm timed on the host in this run, and a large CD solve."""
    import ctypes as C
    import numpy as np
    import torch
    from sem_b200 import _lib as L
    from sem_b200 import Boussinesq_SequentialCoupler as bsc
    out = {}
    peak, _ = measured_peaks()

    def timeit(fn, k=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / k

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as exc:
            out[name] = {"error": str(exc)[:300]}
        torch.cuda.empty_cache()

    def applies():
        d = sem_b200.SemDevice(P_ORDER, NE, NE, 1.0 / NE, 1.0 / NE, device=local)
        n = d.NX * d.NY
        gen = torch.Generator(device=d.tdev).manual_seed(1)
        x, y = d.zeros(), d.zeros()
        x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64)
        t = timeit(lambda: d.apply_stiffness(x, y))
        out["stiffness_apply"] = {"gdof_s": n / t / 1e6, "hbm_frac": 16 * n / t / 1e6 / peak, "ms": t, "bytes_per_node": 16}
        del d, x, y
        torch.cuda.empty_cache()
        ns = sem_b200.NavierStokesSolver(1.0, 1.0, 400.0, 0.0, P_ORDER, NE, NE, u_N=1.0, iprint=[], device=local)
        nd = ns._dev
        for k in range(3):
            ns._in[k][:, :nd.NY] = torch.randn((nd.NX, nd.NY), generator=gen, device=nd.tdev, dtype=torch.float64)
        ns._uv.copy_(ns._in[:2])
        ns._have_sys = True
        ns._jacobians_dev(ns._in[0], ns._in[1])
        nst = ns._state()
        t = timeit(lambda: L.check(nd.lib.sem_ns_jvp(nd.ctx, C.byref(nst), ns._in[0].data_ptr(), ns._in[1].data_ptr(),
                                                     ns._in[2].data_ptr(), None, ns._out[0].data_ptr(), ns._out[1].data_ptr(),
                                                     ns._out[2].data_ptr(), nd.stream), "sem_ns_jvp"), k=5)
        return {"gdof_s": 3 * n / t / 1e6, "hbm_frac": 96 * n / t / 1e6 / peak, "ms": t, "bytes_per_node": 96}

    guarded("ns_jvp_apply", applies)

    def cd_config1():
        """BASELINE config 1: Examples/ConvectionDiffusion_Example.py (P=4, 16x16, Pe=40, u = y - 1/2, v = 1/2 - x)."""
        kw = dict(L_x=1.0, L_y=1.0, Pe=40.0, P=4, N_ex=16, N_ey=16, T_W=0.5, T_E=-0.5)
        cd = sem_b200.ConvectionDiffusionSolver(device=local, **kw)
        u = cd._get_vector(lambda x, y: y - 0.5)
        v = cd._get_vector(lambda x, y: 0.5 - x)
        cd._get_solution(u, v)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        Tg = cd._get_solution(u, v)
        torch.cuda.synchronize()
        tg = time.perf_counter() - t0
        from oracle import sem_oracle as so               # host leg: the reference's algorithm, timed in this run
        cdo = so.CDOracle(**kw)
        t0 = time.perf_counter()
        Th, evals = cdo._get_solution_lgmres(u, v)
        th = time.perf_counter() - t0
        return {"wall_s": tg, "krylov_its": cd.last_iters, "host_wall_s": th, "host_operator_evals": evals,
                "host_kind": "port of CD:123-156 (SciPy LGMRES, no preconditioner, inner_m = 0.3 N) on this box, 1 core",
                "rel_diff_vs_host": float(np.linalg.norm(Tg - Th) / np.linalg.norm(Th))}

    guarded("cd_solve_config1", cd_config1)

    def ns_config2():
        """BASELINE config 2: Examples/NavierStokes_Example.py (P=4, 16x16, Re=400, lid u_N=1), reference default tolerances."""
        res = {"tolerances": "reference defaults (mtol 1e-7, mtol_newton 1e-5)"}
        T0 = np.zeros(4225)
        for precond in ("auto", "full"):
            ns2 = sem_b200.NavierStokesSolver(1, 1, 400, 0, 4, 16, 16, u_N=1, iprint=[], device=local, precond=precond)
            ns2._get_solution(T0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ug, vg, pg = ns2._get_solution(T0)
            torch.cuda.synchronize()
            res[precond] = {"wall_s": time.perf_counter() - t0, "newton_its": ns2._k, "krylov_its": ns2.krylov_iters[-ns2._k:],
                            "total_krylov": int(sum(ns2.krylov_iters[-ns2._k:]))}
            del ns2
        res["wall_s"] = res["auto"]["wall_s"]
        from oracle import sem_oracle as so               # host leg: the reference's algorithm, timed in this run
        nso = so.NSOracle(1, 1, 400, 0, 4, 16, 16, u_N=1)
        t0 = time.perf_counter()
        uh, vh, ph = nso._get_solution(T0, algorithm='reference')
        res["host_wall_s"] = time.perf_counter() - t0
        res["host_newton_its"], res["host_schur_matvecs"] = nso._k, nso.schur_matvecs
        res["host_kind"] = "port of NS:162-270 (SuperLU velocity block + SciPy LGMRES on the Schur complement) on this box, 1 core"
        res["rel_diff_vs_host"] = {"u": float(np.linalg.norm(ug - uh) / np.linalg.norm(uh)),
                                   "p": float(np.linalg.norm(pg - ph) / np.linalg.norm(ph)),
                                   "note": "both stop at the reference's loose default tolerances"}
        return res

    guarded("ns_solve_config2", ns_config2)

    def boussinesq(ne, mode, tol):
        def f():
            t0 = time.perf_counter()
            title, T_e, u_e, v_e, iters = bsc.run_study(save=False, P=4, N_e=ne, mode=mode, mtol_nonlin=tol, mtol_gmres=1e-13 if
                                                        mode == 'JNK' else 1e-10, mtol_internal=1e-13)
            torch.cuda.synchronize()
            return {"wall_s": time.perf_counter() - t0, "mode": mode, "mesh": f"NS {ne}x{ne}, CD {ne // 2}x{ne // 2}, P=4",
                    "dof": 3 * (4 * ne + 1) ** 2 + (2 * ne + 1) ** 2, "iters_cd_ns_nonlin": [int(i) for i in iters],
                    "umax_RePr": float(u_e.max() * 1e3 * 0.71), "mtol_nonlin": tol,
                    "note": "study/Boussinesq_run.py physics (Re=1e3, Ra=1e3, Pr=0.71); set-up included"}
        return f

    guarded("boussinesq_config3", boussinesq(8, 'JNK', 1e-10))       # Examples/Boussinesq_Sequential_Example.py size
    guarded("boussinesq_config4_1M_dof", boussinesq(128, 'GS', 1e-8))  # study run scaled to ~1 M DOF (BASELINE config 4)

    def cd_16m():
        cdb = sem_b200.ConvectionDiffusionSolver(1, 1, PE, P_ORDER, 512, 512, T_W=0.5, T_E=-0.5, mtol=1e-10, restart=60, device=local)
        ub = cdb._get_vector(lambda x, y: y - 0.5)
        vb = cdb._get_vector(lambda x, y: 0.5 - x)
        cdb._get_solution(ub, vb)                  # includes the one-off eigen-decompositions
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cdb._get_solution(ub, vb)
        torch.cuda.synchronize()
        return {"wall_s": time.perf_counter() - t0, "krylov_its": cdb.last_iters, "nodes": cdb.N, "resnorm": cdb.last_resnorm,
                "atol": 1e-10 * float(np.sqrt(cdb.N)),
                "note": "host vectors in / out included; preconditioner: fast diagonalisation on the hand-written DMMA GEMM"}

    guarded("cd_solve_16M_nodes", cd_16m)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="headline only (used for the ncu captures)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
