"""Check of the member-selection claim: the 3N x 3N Jacobian is singular-consistent; right-preconditioned GMRES with the
block lower-triangular preconditioner [[A,0],[C,S~]] converges to the reference's member (l_c . M_p (dp - dp0) = 0) for ANY
Schur preconditioner S~0^-1 once it is corrected to  z <- z - l_c (m_c.z - l_c.r)/(m_c.l_c),  m_c = M_p l_c."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
from oracle import sem_oracle as so
from proto_ns_krylov import gmres_right

P, ne, Re = 4, 8, 100.0
ns = so.NSOracle(1.0, 1.0, Re, 0.0, P, ne, ne, u_N=1.0, mtol=1e-13, mtol_newton=1e-13)
N = ns.N
T = np.zeros(N)
u, v, p = ns._get_solution(T, max_newton=2)
ru, rv, rc = ns._get_residuals(u, v, p, T); ns._calc_jacobians(u, v)
b = -np.hstack((ru, rv, rc))
J = ns.jacobian_matrix().tocsr()
x_ref = np.hstack(ns._get_update(-ru, -rv, -rc))          # the reference's member (oracle: direct solve + constraint)
print('|J x_ref - b| =', np.linalg.norm(J @ x_ref - b))
l = ns._left_null(J.tocsc()); lc = l[2 * N:]
Mp = ns._M.copy(); Mp[ns._pin] = 1.0
mc = Mp * lc
Aa = J[:2 * N, :2 * N].tocsc(); C = J[2 * N:, :2 * N].tocsr()
lu = spla.splu(Aa)
rng = np.random.default_rng(1)
Dr = np.exp(rng.uniform(-1, 1, N))                          # an arbitrary SPD diagonal "other" Schur preconditioner
def tri(Sinv):
    def f(r):
        za = lu.solve(r[:2 * N]); zp = Sinv(r[2 * N:] - C @ za)
        return np.hstack((za, zp))
    return f
mass = lambda y: y / Mp
other = lambda y: Dr * y / Mp
def corrected(y):
    z = other(y)
    return z - lc * ((mc @ z - lc @ y) / (mc @ lc))
tol = 1e-12 * np.linalg.norm(b)
for name, Si in (('mass (reference)', mass), ('other, uncorrected', other), ('other + rank-one correction', corrected)):
    x, its, hist = gmres_right(lambda z: J @ z, b, tri(Si), tol, 1500)
    dp = x[2 * N:] - x_ref[2 * N:]
    print(f'{name:30s} its {its:4d} |Jx-b| {np.linalg.norm(J @ x - b):.1e}  velocity diff {np.linalg.norm(x[:2*N] - x_ref[:2*N]) / np.linalg.norm(x_ref[:2*N]):.1e}'
          f'  pressure diff {np.linalg.norm(dp) / np.linalg.norm(x_ref[2*N:]):.1e}  constraint l_c.M_p dp = {mc @ x[2*N:]:.2e}')
