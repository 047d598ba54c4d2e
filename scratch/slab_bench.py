"""Device time of the fused K / CD / NS applies on ONE GPU for slab-shaped meshes (nex x ney elements): the
no-communication floor of the partitioned apply at 2/4/8 GPUs, the latency of an edge-sized launch, chunk-length sweeps.
usage: slab_bench.py [P] [ney] [nex list] [Mx list] [modes]"""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
import sem_b200

def timeit(fn, n=100, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ney = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
nexs = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [4, 8, 120, 128, 256, 512]
mxs = [int(v) for v in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0, 16, 8, 4]
modes = sys.argv[5].split(',') if len(sys.argv) > 5 else ['CD']
for nex in nexs:
    cd = sem_b200.ConvectionDiffusionSolver(nex / ney, 1.0, 40.0, P, nex, ney, T_W=0.5, T_E=-0.5)
    d = cd._dev
    N = d.NX * d.NY
    gen = torch.Generator(device=d.tdev).manual_seed(0)
    def rnd():
        x = d.zeros(); x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64); return x
    # rotate over enough vector sets to exceed L2 (126 MB) so that small slabs are not timed out of cache
    nset = max(1, int(400e6 // (4 * 8 * d.vec_len)) + 1)
    sets = [(rnd(), rnd(), rnd(), d.zeros()) for _ in range(nset)]
    cd._have_sys = True
    calls = []
    for x, u, v, y in sets:
        cd._u = u; cd._v = v
        calls.append((C.byref(cd._state(with_jac=False)), x.data_ptr(), y.data_ptr()))
    k = [0]
    f, fk, ctx, stream = d.lib.sem_cd_jvp, d.lib.sem_apply_stiffness, d.ctx, d.stream
    def run_cd():
        s, xp, yp = calls[k[0] % nset]; k[0] += 1
        f(ctx, s, xp, None, None, yp, stream)
    def run_k():
        x, u, v, y = sets[k[0] % nset]; k[0] += 1
        d.apply_stiffness(x, y)
    if 'NS' in modes:
        ns = sem_b200.NavierStokesSolver(nex / ney, 1.0, 400.0, 0.0, P, nex, ney, u_N=1.0, iprint=[], device_obj=d) \
            if False else sem_b200.NavierStokesSolver(nex / ney, 1.0, 400.0, 0.0, P, nex, ney, u_N=1.0, iprint=[])
        x, u, v, y = sets[0]
        ns._uv[0].copy_(u); ns._uv[1].copy_(v); ns._have_sys = True
        ns._jacobians_dev(u, v)
        nst = ns._state()
        x3, y3 = ns._in, ns._out
        x3[0].copy_(x); x3[1].copy_(u); x3[2].copy_(v)
        fn, nctx, nstream = d.lib.sem_ns_jvp, ns._dev.ctx, ns._dev.stream
        def run_ns():
            fn(nctx, C.byref(nst), x3[0].data_ptr(), x3[1].data_ptr(), x3[2].data_ptr(), None,
               y3[0].data_ptr(), y3[1].data_ptr(), y3[2].data_ptr(), nstream)
    for mode in modes:
        line = f"{mode:2s} nex={nex:4d} N={N:9d}"
        fn_ = run_ns if mode == 'NS' else (run_cd if mode == 'CD' else run_k)
        dev_ = ns._dev if mode == 'NS' else d
        ts = {Mx: [] for Mx in mxs}
        for rep in range(7):                     # interleaved repetitions: clock drift hits every Mx alike
            for Mx in mxs:
                dev_.set_tiling(0, Mx)
                ts[Mx].append(timeit(fn_, n=20 if mode != 'NS' else 8, warm=2))
        for Mx in mxs:
            v = sorted(ts[Mx])
            line += f" | {Mx:2d}: {v[len(v)//2]*1e3:6.1f}/{v[0]*1e3:6.1f}"
        print(line + "  (us median/min)", flush=True)
    if 'NS' in modes: del ns
    del sets, cd, d, calls
    torch.cuda.empty_cache()
