"""Convective Schur complement (Newton state, Re > 0): mass vs pressure-convection-diffusion (PCD) preconditioning,
on the reduced (interior-pressure) dense Schur complement with exact velocity solves."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps
from p2 import build, dense_schur
from p4 import gmres
P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3])
ns, J = build(P, ne, Re, False)
N = ns.N
S, lu = dense_schur(ns, J)
Mp = ns._M.copy(); Mp[ns._pin] = 1
bset = ns._mask_bound.copy(); bset[ns._pin] = True
Ii = np.where(~bset)[0]; Bi = np.where(bset)[0]
Sred = S[np.ix_(Ii, Ii)] - S[np.ix_(Ii, Bi)] @ np.linalg.solve(S[np.ix_(Bi, Bi)], S[np.ix_(Bi, Ii)])
MI = Mp[Ii]
K = ns._K.toarray()
Fp = K + Re * (np.diag(ns._u) @ ns._G_x.toarray() + np.diag(ns._v) @ ns._G_y.toarray())
# pressure Laplacian with the boundary rows eliminated the same way (Neumann rows K_B p = 0): K_red = K_II - K_IB K_BB^-1 K_BI
KBBi = np.linalg.inv(S[np.ix_(Bi, Bi)])        # S_BB = K_BB (+ pin)
Kred = K[np.ix_(Ii, Ii)] - K[np.ix_(Ii, Bi)] @ KBBi @ S[np.ix_(Bi, Ii)]
Fred = Fp[np.ix_(Ii, Ii)] - Fp[np.ix_(Ii, Bi)] @ KBBi @ S[np.ix_(Bi, Ii)]
Kri = np.linalg.pinv(Kred, rcond=1e-11)
rng = np.random.default_rng(0)
b = Sred @ rng.standard_normal(Ii.size); tol = 1e-10 * np.linalg.norm(b)
A = lambda x: Sred @ x
mass = lambda r: r / MI
pcd = lambda r: (Fred @ (Kri @ r)) / MI
pcd2 = lambda r: Kri @ (Fred @ (r / MI))
for name, Pi in (('mass', mass), ('PCD  M^-1 F K^+', pcd), ('PCD  K^+ F M^-1', pcd2)):
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, 1200)
    print(f'{name:20s} its {its:5d} relres {np.linalg.norm(b - Sred @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}', flush=True)
# ---- two-level (structured, p10) with PCD instead of the mass sweep
import scipy.linalg as sla
from numpy.polynomial import legendre as npl
from oracle import sem_oracle as so
n1 = ne * P + 1; h = 1.0 / ne; nI = n1 - 2
xi = so.gll(P)[0]; LP = npl.legval(xi, [0] * P + [1])
s1 = np.zeros(n1)
for m in range(ne): s1[m * P:m * P + P + 1] = (1.0 if P % 2 == 0 else (-1.0) ** m) * LP
W = np.zeros((n1, ne + 1)); xn = (xi + 1) / 2
for m in range(ne):
    W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1 - xn)
    W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xn)
W = W * s1[:, None]
M1 = so._assembled_1d(h / 2 * so.mass_1d(P), ne).toarray(); K1 = so._assembled_1d(2 / h * so.stiff_1d(P), ne).toarray()
G1 = so._assembled_1d(so.grad_1d(P), ne).toarray()
E = np.zeros((n1, nI)); E[np.arange(1, n1 - 1), np.arange(nI)] = 1
KII = E.T @ K1 @ E; MII = E.T @ M1 @ E
lam_s = (s1 @ K1 @ s1) / (s1 @ M1 @ s1)
Rs = G1 @ E @ np.linalg.solve(KII + 0.25 * lam_s * MII, E.T @ G1.T)
rho, Vv = sla.eigh(E.T @ Rs @ E, MII)
WI = E.T @ W; PW = WI @ np.linalg.pinv(WI.T @ MII @ WI) @ WI.T @ MII
den = rho[:, None] + rho[None, :]; inv = np.where(den > 1e-9 * den.max(), 1.0 / np.where(den > 0, den, 1), 0.0)
def proj(Xm, Pm):
    PX = Pm @ Xm
    return PX + Xm @ Pm.T - PX @ Pm.T
full_int = np.where(~ns._mask_bound)[0]; int_pos = np.searchsorted(full_int, Ii)
def coarse(yI):
    Y = np.zeros(nI * nI); Y[int_pos] = yI
    X = Vv @ ((Vv.T @ proj(Y.reshape(nI, nI), PW.T) @ Vv) * inv) @ Vv.T
    return proj(X, PW).ravel()[int_pos]
def two(stage2):
    def f(r):
        z = coarse(r)
        return z + stage2(r - Sred @ z)
    return f
for name, Pi in (('two-level -> mass', two(mass)), ('two-level -> PCD', two(pcd)), ('PCD -> two-level -> PCD', lambda r: (lambda z0: (lambda z1: z1 + pcd(r - Sred @ z1))(z0 + coarse(r - Sred @ z0)))(pcd(r)))):
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, 1200)
    print(f'{name:26s} its {its:5d} relres {np.linalg.norm(b - Sred @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}', flush=True)
