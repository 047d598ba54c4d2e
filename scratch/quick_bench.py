"""Quick device timing of the fused applies at BASELINE config 5 (P=8, 1024x1024 elements) for several tilings.
usage: quick_bench.py [P] [ne] [modes e.g. K,CD,NS] [tilings e.g. 0:0,16:16,16:32]"""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
import sem_b200

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

P = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
modes = sys.argv[3].split(',') if len(sys.argv) > 3 else ['K', 'CD', 'NS']
tilings = [tuple(int(v) for v in s.split(':')) for s in sys.argv[4].split(',')] if len(sys.argv) > 4 else [(0, 0)]
d = sem_b200.SemDevice(P, ne, ne, 1.0 / ne, 1.0 / ne)
N = d.NX * d.NY
gen = torch.Generator(device=d.tdev).manual_seed(0)
def rnd():
    x = d.zeros(); x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64); return x
x, u, v, y = rnd(), rnd(), rnd(), d.zeros()
t = timeit(lambda: y.copy_(x)); print(f"torch copy: {t:.3f} ms  {16*d.vec_len/t/1e6:.0f} GB/s", flush=True)
lib = d.lib
cd = sem_b200.ConvectionDiffusionSolver(1.0, 1.0, 40.0, P, ne, ne, T_W=0.5, T_E=-0.5)
cd._u.copy_(u); cd._v.copy_(v); cd._have_sys = True
st = cd._state(with_jac=False)
if 'NS' in modes:
    ns = sem_b200.NavierStokesSolver(1.0, 1.0, 400.0, 0.0, P, ne, ne, u_N=1.0, iprint=[])
    ns._uv[0].copy_(u); ns._uv[1].copy_(v); ns._have_sys = True
    ns._jacobians_dev(u, v)
    nst = ns._state()
    x3, y3 = ns._in, ns._out
    x3[0].copy_(x); x3[1].copy_(u); x3[2].copy_(v)
for Ty, Mx in tilings:
    line = f"Ty={Ty:3d} Mx={Mx:3d} "
    if 'K' in modes:
        d.set_tiling(Ty, Mx)
        tK = timeit(lambda: d.apply_stiffness(x, y))
        line += f" K {tK:7.3f} ms {N/tK/1e6:7.1f} GDOF/s {16*N/tK/1e6:6.0f} GB/s |"
    if 'CD' in modes:
        cd._dev.set_tiling(Ty, Mx)
        tC = timeit(lambda: lib.sem_cd_jvp(cd._dev.ctx, C.byref(st), x.data_ptr(), None, None, y.data_ptr(), cd._dev.stream))
        line += f" CD {tC:7.3f} ms {N/tC/1e6:7.1f} GDOF/s {32*N/tC/1e6:6.0f} GB/s |"
    if 'NS' in modes:
        ns._dev.set_tiling(Ty, Mx)
        tN = timeit(lambda: lib.sem_ns_jvp(ns._dev.ctx, C.byref(nst), x3[0].data_ptr(), x3[1].data_ptr(), x3[2].data_ptr(), None,
                                           y3[0].data_ptr(), y3[1].data_ptr(), y3[2].data_ptr(), ns._dev.stream), n=5)
        line += f" NS {tN:7.3f} ms {3*N/tN/1e6:7.1f} GDOF/s {96*N/tN/1e6:6.0f} GB/s"
    print(line, flush=True)
