"""Two-level test: mass preconditioner + exact coarse correction on the span of {s(x) e_k(x) (x) delta_y} and
{delta_x (x) s(y) e_k(y)}, s = L_P(xi) element-wise, e_k = hat functions on the element vertices."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps
from numpy.polynomial import legendre as npl
from p2 import build, dense_schur
from p4 import gmres
from oracle import sem_oracle as so

P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); stokes = len(sys.argv) > 4 and sys.argv[4] == 's'
ns, J = build(P, ne, Re, stokes)
N = ns.N; n1 = ne * P + 1
S, lu = dense_schur(ns, J)
Mp = ns._M.copy(); Mp[ns._pin] = 1
xi = so.gll(P)[0]
LP = npl.legval(xi, [0] * P + [1])           # L_P at the GLL nodes
s1 = np.zeros(n1)
for m in range(ne):
    sign = 1.0 if P % 2 == 0 else (-1.0) ** m
    s1[m * P:m * P + P + 1] = sign * LP
def hats(ne, P):
    W = np.zeros((n1, ne + 1))
    t = np.linspace(0, 1, P + 1)   # hat in reference coordinate of node index (linear in node index ~ fine for a test)
    xnodes = (xi + 1) / 2
    for m in range(ne):
        W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1 - xnodes)
        W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xnodes)
    return W
Wx = hats(ne, P) * s1[:, None]               # n1 x (ne+1): s(x) e_k(x)
I1 = np.eye(n1)
Zx = np.kron(Wx, I1)                         # x slow, y fast
Zy = np.kron(I1, Wx)
Zxy = np.kron(Wx, Wx)
rng = np.random.default_rng(0)
xt = rng.standard_normal(N); b = S @ xt
tol = 1e-10 * np.linalg.norm(b)
A = lambda x: S @ x
def mass(r): return r / Mp
def two_level(Z):
    Sc = Z.T @ S @ Z
    Sci = np.linalg.pinv(Sc, rcond=1e-10)
    def f(r):
        return r / Mp + Z @ (Sci @ (Z.T @ r))
    return f
def two_level_mult(Z):
    Sc = Z.T @ S @ Z
    Sci = np.linalg.pinv(Sc, rcond=1e-10)
    def f(r):
        z = Z @ (Sci @ (Z.T @ r))
        return z + (r - S @ z) / Mp
    return f
print('N', N, 'coarse dims', Zx.shape[1], Zy.shape[1])
for name, Pi in (('mass', mass), ('add x+y', two_level(np.hstack((Zx, Zy)))), ('mult x+y', two_level_mult(np.hstack((Zx, Zy)))),
                 ('mult x+y+xy', two_level_mult(np.hstack((Zx, Zy, Zxy))))):
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, min(N, 1200))
    print(f'{name:12s} its {its:5d} relres {np.linalg.norm(b - S @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}',
          [f'{h:.0e}' for h in hist[::max(1, len(hist) // 8)]])
