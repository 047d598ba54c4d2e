"""Structured two-level Schur preconditioner on the reduced (interior-pressure) system: iteration counts vs mesh and Re.
 stage 1: z1 = Pi Shat^+ Pi^T r   (Pi = P_W(x)I + I(x)P_W - P_W(x)P_W, Shat = R(lam_s)(x)M + M(x)R(lam_s) by fast diagonalisation)
 stage 2: z  = z1 + M^-1 (r - S z1)"""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.linalg as sla
from numpy.polynomial import legendre as npl
from p2 import build, dense_schur
from p4 import gmres
from oracle import sem_oracle as so

P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); stokes = len(sys.argv) > 4 and sys.argv[4] == 's'
t0 = time.time()
ns, J = build(P, ne, Re, stokes)
N = ns.N; n1 = ne * P + 1; h = 1.0 / ne; nI = n1 - 2
S, lu = dense_schur(ns, J)
print('dense S', time.time() - t0, flush=True)
Mp = ns._M.copy(); Mp[ns._pin] = 1
bset = ns._mask_bound.copy(); bset[ns._pin] = True
Ii = np.where(~bset)[0]; Bi = np.where(bset)[0]
Sred = S[np.ix_(Ii, Ii)] - S[np.ix_(Ii, Bi)] @ np.linalg.solve(S[np.ix_(Bi, Bi)], S[np.ix_(Bi, Ii)])
del S
MI = Mp[Ii]
xi = so.gll(P)[0]
LP = npl.legval(xi, [0] * P + [1])
s1 = np.zeros(n1)
for m in range(ne):
    s1[m * P:m * P + P + 1] = (1.0 if P % 2 == 0 else (-1.0) ** m) * LP
W = np.zeros((n1, ne + 1)); xn = (xi + 1) / 2
for m in range(ne):
    W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1 - xn)
    W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xn)
W = W * s1[:, None]
M1 = so._assembled_1d(h / 2 * so.mass_1d(P), ne).toarray()
K1 = so._assembled_1d(2 / h * so.stiff_1d(P), ne).toarray()
G1 = so._assembled_1d(so.grad_1d(P), ne).toarray()
E = np.zeros((n1, nI)); E[np.arange(1, n1 - 1), np.arange(nI)] = 1
KII = E.T @ K1 @ E; MII = E.T @ M1 @ E
lam_s = (s1 @ K1 @ s1) / (s1 @ M1 @ s1)
def stage(shift_scale):
    Rs = G1 @ E @ np.linalg.solve(KII + shift_scale * lam_s * MII, E.T @ G1.T)
    R_II = E.T @ Rs @ E
    rho, Vv = sla.eigh(R_II, MII)
    WI = E.T @ W
    PW = WI @ np.linalg.pinv(WI.T @ MII @ WI) @ WI.T @ MII
    PWt = PW.T
    den = rho[:, None] + rho[None, :]
    inv = np.where(den > 1e-9 * den.max(), 1.0 / np.where(den > 0, den, 1), 0.0)
    def proj(Xm, Pm):
        PX = Pm @ Xm
        return PX + Xm @ Pm.T - PX @ Pm.T
    def f(r):
        Rm = proj(r.reshape(nI, nI), PWt)
        X = Vv @ ((Vv.T @ Rm @ Vv) * inv) @ Vv.T
        return proj(X, PW).ravel()
    return f
full_int = np.where(~ns._mask_bound)[0]
pin_k = int(np.searchsorted(full_int, ns._pin))
def wrap(st):
    return lambda r: np.delete(st(np.insert(r, pin_k, 0.0)), pin_k)
rng = np.random.default_rng(0)
xt2 = rng.standard_normal(Ii.size); b2 = Sred @ xt2; tol2 = 1e-10 * np.linalg.norm(b2)
A2 = lambda x: Sred @ x
def mult2(coarse):
    def f(r):
        z = coarse(r)
        return z + (r - Sred @ z) / MI
    return f
def sym2(coarse):       # mass - coarse - mass
    def f(r):
        z = r / MI
        z = z + coarse(r - Sred @ z)
        return z + (r - Sred @ z) / MI
    return f
cands = [('mass', lambda r: r / MI)]
for sc in (1.0, 0.5, 0.25):
    cands.append((f'structured shift {sc} lam_s -> mass', mult2(wrap(stage(sc)))))
cands.append(('mass -> structured(1.0) -> mass', sym2(wrap(stage(1.0)))))
for name, Pi in cands:
    t = time.time()
    x, its, hist = gmres(A2, b2, Pi, tol2, min(Ii.size, 1500))
    print(f'{name:36s} its {its:5d} relres {np.linalg.norm(b2 - Sred @ x) / np.linalg.norm(b2):.2e} t {time.time() - t:.1f}', flush=True)
