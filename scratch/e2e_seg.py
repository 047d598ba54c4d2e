"""e2e CD apply (pinned numpy in -> numpy out) at config 5; the segment count comes from SEM_B200_HOST_SEGMENTS."""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch, sem_b200
P, ne = 8, 1024
cd = sem_b200.ConvectionDiffusionSolver(1.0, 1.0, 40.0, P, ne, ne, T_W=0.5, T_E=-0.5)
d = cd._dev
cd._have_sys = True
n = d.N_local
host_in = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
host_in[:] = np.random.default_rng(0).standard_normal(n)
for _ in range(3): r = cd._get_dresiduals(host_in)
torch.cuda.synchronize()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); r = cd._get_dresiduals(host_in); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print('segments', os.environ.get('SEM_B200_HOST_SEGMENTS', '8'), 'ms median', sorted(ts)[len(ts) // 2], 'min', min(ts), flush=True)
