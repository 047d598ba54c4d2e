"""Device time of the CD Jacobian apply on an inner slab of config 5 on ONE GPU: plain (no exchange: the floor) against the
loopback communicator (the slab is its own neighbour: the full partitioned-apply path incl. the peer-memory exchange).
SEM_B200_FUSED_XCH=0 in the environment selects the three-launch + exchange-kernel path for the loopback context.
usage: loopback_bench.py [nex list] [reps]"""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
import sem_b200
from sem_b200 import _lib as L

P, ney = 8, 1024
nexs = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [128, 256, 512]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
mxs = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0]      # chunk lengths (0: the library's choice)

def timeit(fn, n, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for nex, Mx in [(a, b) for a in nexs for b in mxs]:
    N_ex = 1024
    res = {}
    for kind in ('plain', 'loopback'):
        d = sem_b200.SemDevice(P, N_ex, ney, 1.0 / N_ex, 1.0 / ney, m_begin=nex, m_end=2 * nex)
        if kind == 'loopback':
            L.check(d.lib.sem_ctx_attach_loopback(d.ctx), 'loopback')
        if Mx:
            d.set_tiling(0, Mx)
        gen = torch.Generator(device=d.tdev).manual_seed(0)
        def rnd():
            x = d.zeros(); x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64); return x
        nset = max(1, int(400e6 // (4 * 8 * d.vec_len)) + 1)     # rotate over more data than L2 holds
        sets = [(rnd(), rnd(), rnd(), d.zeros()) for _ in range(nset)]
        bc = L.sem_cd_bc()
        for k, (a, val) in enumerate([(1, 0.5), (1, -0.5), (0, 0.0), (0, 0.0)]):
            bc.active[k], bc.value[k] = a, val
        calls = []
        for x, u, v, y in sets:
            st = L.sem_cd_state(); st.bc, st.Pe = bc, 40.0
            st.u, st.v, st.gxT, st.gyT = u.data_ptr(), v.data_ptr(), None, None
            calls.append((st, x.data_ptr(), y.data_ptr()))
        k = [0]
        f, ctx, stream = d.lib.sem_cd_jvp, d.ctx, d.stream
        def run():
            st, xp, yp = calls[k[0] % nset]; k[0] += 1
            f(ctx, C.byref(st), xp, None, None, yp, stream)
        res[kind] = timeit(run, reps)
        if kind == 'loopback':
            res['fused'] = d.lib.sem_ctx_partitioned_applies(d.ctx, 1) > 0
        del sets, calls, d
        torch.cuda.empty_cache()
    nodes = (nex * P + 1) * (ney * P + 1)
    ideal = 32 * nodes / 6531.9e9 * 1e6
    print(f"CD slab nex={nex:4d} Mx={Mx:2d}: plain {res['plain']:7.1f} us, loopback ({'one launch' if res['fused'] else 'three launches + exchange kernel'}) "
          f"{res['loopback']:7.1f} us, HBM-roofline time {ideal:6.1f} us", flush=True)
