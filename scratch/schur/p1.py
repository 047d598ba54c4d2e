"""Throw-away: structure of the NS Schur complement (null vectors, spectrum) on the oracle's matrices."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
from oracle import sem_oracle as so

P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]) if len(sys.argv) > 3 else 400.0
ns = so.NSOracle(1.0, 1.0, Re, 0.0, P, ne, ne, u_N=1.0, mtol=1e-13, mtol_newton=1e-10)
N = ns.N; NX = ne * P + 1
T = np.zeros(N)
t = time.time()
u, v, p = ns._get_solution(T, max_newton=3)   # a few Newton steps: representative state
print('newton state', time.time() - t, 'k', ns._k)
ns._get_residuals(u, v, p, T); ns._calc_jacobians(u, v)
J = ns.jacobian_matrix().tocsr()
l = ns._left_null(J.tocsc())
lc = l[2 * N:].reshape(NX, NX)
sv = np.linalg.svd(lc, compute_uv=False)
print('l_c singular values', sv[:5] / sv[0])
print('l_a norm interior', np.linalg.norm(l[:2 * N][~np.hstack((ns._mask_bound,) * 2)]), 'bnd', np.linalg.norm(l[:2 * N]))
# right null vector
r_star = 2 * N + ns._interior_probe()
keep = np.ones(3 * N); keep[r_star] = 0; unit = np.zeros(3 * N); unit[r_star] = 1
lu = spla.splu((sps.diags(keep) @ J + sps.diags(unit)).tocsc())
q = lu.solve(unit)
print('J q', np.linalg.norm(J @ q), 'q vel', np.linalg.norm(q[:2 * N]), np.linalg.norm(q[2 * N:]))
ps = q[2 * N:].reshape(NX, NX)
sv = np.linalg.svd(ps, compute_uv=False)
print('p_s singular values', sv[:5] / sv[0])
U, s, Vt = np.linalg.svd(ps)
print('p_s x-factor (first 2 elements)', U[:2 * P + 1, 0] / np.abs(U[:, 0]).max())
print('p_s y-factor', Vt[0, :2 * P + 1] / np.abs(Vt[0]).max())
U, s, Vt = np.linalg.svd(lc)
print('l_c x-factor', U[:2 * P + 1, 0] / np.abs(U[:, 0]).max())
print('l_c y-factor', Vt[0, :2 * P + 1] / np.abs(Vt[0]).max())
np.savez(f'/tmp/state_{P}_{ne}.npz', u=u, v=v, p=p)
