"""Velocity block of the Newton-state Jacobian, A = K + Re (conv + diag terms) with Dirichlet rows: how good are cheap
approximate inverses?  GMRES iterations on A alone (to 1e-8) with: Stokes (K^-1, today's fast diagonalisation), shifted
(K + a M)^-1 (also diagonalisable), k inner steps."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
from oracle import sem_oracle as so
from proto_ns_krylov import gmres_right
P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3])
ns = so.NSOracle(1.0, 1.0, Re, 0.0, P, ne, ne, u_N=1.0, mtol=1e-13, mtol_newton=1e-13)
N = ns.N; T = np.zeros(N)
u, v, p = ns._get_solution(T, max_newton=3)
ns._get_residuals(u, v, p, T); ns._calc_jacobians(u, v)
J = ns.jacobian_matrix().tocsr()
A = J[:2 * N, :2 * N].tocsc()
inner = (~ns._mask_bound).astype(float); bnd = ns._mask_bound.astype(float)
I2 = sps.identity(2 * N, format='csc')
Din = sps.diags(np.hstack((inner, inner))); Dbn = sps.diags(np.hstack((bnd, bnd)))
K2 = sps.block_diag((ns._K, ns._K), format='csc'); M2 = sps.diags(np.hstack((ns._M, ns._M)))
def shifted(a):
    return spla.splu((Din @ (K2 + a * M2) + Dbn).tocsc())
rng = np.random.default_rng(0)
b = A @ rng.standard_normal(2 * N); tol = 1e-8 * np.linalg.norm(b)
print(f'P={P} ne={ne} Re={Re}: max |Re grad u| = {max(np.abs(ns._gxu).max(), np.abs(ns._gyv).max()) / ns._M.max():.1f} (per unit mass), |u|max {np.abs(u).max():.2f}')
for name, lu in [('Stokes K^-1', shifted(0.0))] + [(f'(K + {a} M)^-1', shifted(a)) for a in (Re * 0.5, Re * 2.0, Re * 8.0)]:
    x, its, hist = gmres_right(lambda z: A @ z, b, lu.solve, tol, 600)
    print(f'{name:22s} its {its:4d} res {np.linalg.norm(A @ x - b) / np.linalg.norm(b):.1e}')
