"""Two-level Schur preconditioner with STRUCTURED coarse solves (Kronecker products of 1-D operators)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps, scipy.linalg as sla
from numpy.polynomial import legendre as npl
from p2 import build, dense_schur
from p4 import gmres
from oracle import sem_oracle as so

P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); stokes = len(sys.argv) > 4 and sys.argv[4] == 's'
ns, J = build(P, ne, Re, stokes)
N = ns.N; n1 = ne * P + 1; h = 1.0 / ne
S, lu = dense_schur(ns, J)
Mp = ns._M.copy(); Mp[ns._pin] = 1
xi = so.gll(P)[0]
LP = npl.legval(xi, [0] * P + [1])
s1 = np.zeros(n1)
for m in range(ne):
    sign = 1.0 if P % 2 == 0 else (-1.0) ** m
    s1[m * P:m * P + P + 1] = sign * LP
W = np.zeros((n1, ne + 1)); xn = (xi + 1) / 2
for m in range(ne):
    W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1 - xn)
    W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xn)
W = W * s1[:, None]
# 1-D operators
M1 = so._assembled_1d(h / 2 * so.mass_1d(P), ne).toarray()
K1 = so._assembled_1d(2 / h * so.stiff_1d(P), ne).toarray()
G1 = so._assembled_1d(so.grad_1d(P), ne).toarray()
m1 = np.diag(M1)
I = np.arange(1, n1 - 1)
E = np.zeros((n1, n1 - 2)); E[I, np.arange(n1 - 2)] = 1
KII = E.T @ K1 @ E; MII = E.T @ M1 @ E
lam, Q = sla.eigh(KII, MII)                 # Q^T MII Q = I
def R1(mu):                                  # G1 E (K_II + mu M_II)^-1 E^T G1^T   (n1 x n1, PSD)
    X = np.linalg.solve(KII + mu * MII, E.T @ G1.T)
    return G1 @ E @ X
Mt = M1 @ E @ np.linalg.solve(MII, E.T @ M1)   # M E M_II^-1 E^T M = interior mass
T1 = M1 @ E @ np.linalg.solve(KII, E.T @ M1)   # M E K_II^-1 E^T M
lam_s = (s1 @ K1 @ s1) / (s1 @ M1 @ s1)
print('lam_s', lam_s, 'lam range', lam[0], lam[-1])
I1 = np.eye(n1)
Zx = np.kron(W, I1); Zy = np.kron(I1, W)
rng = np.random.default_rng(0)
xt = rng.standard_normal(N); b = S @ xt
tol = 1e-10 * np.linalg.norm(b)
A = lambda x: S @ x

def mult(coarse):
    def f(r):
        z = coarse(r)
        return z + (r - S @ z) / Mp
    return f
# (1) exact coarse, x and y blocks coupled
Z = np.hstack((Zx, Zy)); Sci = np.linalg.pinv(Z.T @ S @ Z, rcond=1e-10)
c_exact = lambda r: Z @ (Sci @ (Z.T @ r))
# (2) exact, block diagonal (no x-y coupling)
Sxi = np.linalg.pinv(Zx.T @ S @ Zx, rcond=1e-10); Syi = np.linalg.pinv(Zy.T @ S @ Zy, rcond=1e-10)
c_bd = lambda r: Zx @ (Sxi @ (Zx.T @ r)) + Zy @ (Syi @ (Zy.T @ r))
# (3) Kronecker model:  Zx^T S Zx ~ (W^T Mt W) (x) R1(lam_s)
def kron_coarse(Ax, By, tag):
    Axi = np.linalg.pinv(Ax, rcond=1e-10); Byi = np.linalg.pinv(By, rcond=1e-10)
    Px = W @ Axi @ W.T                     # n1 x n1
    def f(r):
        Rm = r.reshape(n1, n1)
        return (Px @ Rm @ Byi + Byi @ Rm @ Px).ravel()
    return f
Sxx = Zx.T @ S @ Zx
for tag, Ax, By in (('Mt(x)R1(lam_s)', W.T @ Mt @ W, R1(lam_s)), ('T1(x)A', W.T @ T1 @ W, G1.T @ E @ np.linalg.solve(MII, E.T @ G1)),
                    ('Mt(x)R1(lam_max)', W.T @ Mt @ W, R1(lam[-1])), ('Mt(x)R1(lam_s/2)', W.T @ Mt @ W, R1(lam_s / 2))):
    model = np.kron(Ax, By)
    print(f'model {tag:18s} rel err of Zx^T S Zx: {np.linalg.norm(model - Sxx) / np.linalg.norm(Sxx):.3f}')
cands = [('mass', lambda r: r / Mp), ('exact coupled', mult(c_exact)), ('exact block-diag', mult(c_bd)),
         ('kron Mt,R1(lam_s)', mult(kron_coarse(W.T @ Mt @ W, R1(lam_s), ''))),
         ('kron T1,A', mult(kron_coarse(W.T @ T1 @ W, G1.T @ E @ np.linalg.solve(MII, E.T @ G1), '')))]
for name, Pi in cands:
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, min(N, 1200))
    print(f'{name:20s} its {its:5d} relres {np.linalg.norm(b - S @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}',
          [f'{h_:.0e}' for h_ in hist[::max(1, len(hist) // 8)]])
print('--- diagnostics')
By = R1(lam_s); Ax = W.T @ Mt @ W
model = np.kron(Ax, By)
d1 = np.diag(Sxx); d2 = np.diag(model)
print('diag ratio Sxx/model: min', (d1 / d2).min(), 'median', np.median(d1 / d2), 'max', (d1 / d2).max())
# restrict to interior y nodes and interior hats
ky = np.arange(1, n1 - 1); kx = np.arange(1, ne)
idx = (kx[:, None] * n1 + ky[None, :]).ravel()
A_ = Sxx[np.ix_(idx, idx)]; B_ = model[np.ix_(idx, idx)]
print('interior block: ||S||', np.linalg.norm(A_), '||model||', np.linalg.norm(B_), 'rel err', np.linalg.norm(A_ - B_) / np.linalg.norm(A_), 'rel err (sign flipped)', np.linalg.norm(A_ + B_) / np.linalg.norm(A_))
ev = np.linalg.eigvals(np.linalg.solve(B_, A_)); print('gen eig (interior) range', np.sort(ev.real)[[0, 1, -2, -1]])
