"""Launch each fused apply a few times at config 5 (for ncu captures)."""
import ctypes as C, sys
sys.path.insert(0, '.')
import torch, sem_b200
P, ne = 8, 1024
modes = sys.argv[1].split(',') if len(sys.argv) > 1 else ['K', 'CD', 'NS']
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = sem_b200.SemDevice(P, ne, ne, 1.0 / ne, 1.0 / ne)
gen = torch.Generator(device=d.tdev).manual_seed(0)
def rnd():
    x = d.zeros(); x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64); return x
x, u, v, y = rnd(), rnd(), rnd(), d.zeros()
lib = d.lib
if 'K' in modes:
    for _ in range(reps): d.apply_stiffness(x, y)
if 'CD' in modes:
    cd = sem_b200.ConvectionDiffusionSolver(1.0, 1.0, 40.0, P, ne, ne, T_W=0.5, T_E=-0.5)
    cd._u.copy_(u); cd._v.copy_(v); cd._have_sys = True
    st = cd._state(with_jac=False)
    for _ in range(reps): lib.sem_cd_jvp(cd._dev.ctx, C.byref(st), x.data_ptr(), None, None, y.data_ptr(), cd._dev.stream)
if 'NS' in modes:
    ns = sem_b200.NavierStokesSolver(1.0, 1.0, 400.0, 0.0, P, ne, ne, u_N=1.0, iprint=[])
    ns._uv[0].copy_(u); ns._uv[1].copy_(v); ns._have_sys = True
    ns._jacobians_dev(u, v)
    nst = ns._state()
    x3, y3 = ns._in, ns._out
    x3[0].copy_(x); x3[1].copy_(u); x3[2].copy_(v)
    for _ in range(reps):
        lib.sem_ns_jvp(ns._dev.ctx, C.byref(nst), x3[0].data_ptr(), x3[1].data_ptr(), x3[2].data_ptr(), None,
                       y3[0].data_ptr(), y3[1].data_ptr(), y3[2].data_ptr(), ns._dev.stream)
torch.cuda.synchronize()
print("done")
