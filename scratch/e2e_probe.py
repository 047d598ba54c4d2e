"""Break the end-to-end CD apply (host numpy in -> host numpy out) into its pieces at config 5."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch, sem_b200
P, ne = 8, 1024
cd = sem_b200.ConvectionDiffusionSolver(1.0, 1.0, 40.0, P, ne, ne, T_W=0.5, T_E=-0.5)
d = cd._dev
cd._have_sys = True
n = d.N_local
host_in = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
host_in[:] = 1.0
pageable = np.ones(n)
def t(fn, k=5):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3
buf = d.zeros()
print("to_device pinned  ", t(lambda: d.to_device(host_in, buf)))
print("to_device pageable", t(lambda: d.to_device(pageable, buf)))
print("to_host (pinned result alloc)", t(lambda: d.to_host(buf)))
keep = []
def th():
    keep.append(d.to_host(buf))
    if len(keep) > 2: keep.pop(0)
print("to_host keep 2 alive", t(th))
out = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
print("to_host into given pinned", t(lambda: d.to_host(buf, out)))
outp = np.empty(n)
print("to_host into pageable", t(lambda: d.to_host(buf, outp)))
print("full _get_dresiduals pinned in", t(lambda: cd._get_dresiduals(host_in)))
print("torch raw H2D pinned", t(lambda: buf.view(-1)[:n].copy_(torch.from_numpy(host_in), non_blocking=True)))
