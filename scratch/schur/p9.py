"""Coupled coarse solve with a MODEL operator: Z^T S_hat Z, S_hat = R(lam_s) (x) Mt + Mt (x) R(lam_s) (a global Kronecker sum),
variants with R(0) in the diagonal blocks; multiplicative with the mass preconditioner."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.linalg as sla
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
    s1[m * P:m * P + P + 1] = (1.0 if P % 2 == 0 else (-1.0) ** m) * LP
W = np.zeros((n1, ne + 1)); xn = (xi + 1) / 2
for m in range(ne):
    W[m * P:m * P + P + 1, m] = np.maximum(W[m * P:m * P + P + 1, m], 1 - xn)
    W[m * P:m * P + P + 1, m + 1] = np.maximum(W[m * P:m * P + P + 1, m + 1], xn)
W = W * s1[:, None]
M1 = so._assembled_1d(h / 2 * so.mass_1d(P), ne).toarray()
K1 = so._assembled_1d(2 / h * so.stiff_1d(P), ne).toarray()
G1 = so._assembled_1d(so.grad_1d(P), ne).toarray()
E = np.zeros((n1, n1 - 2)); E[np.arange(1, n1 - 1), np.arange(n1 - 2)] = 1
KII = E.T @ K1 @ E; MII = E.T @ M1 @ E
R1 = lambda mu: G1 @ E @ np.linalg.solve(KII + mu * MII, E.T @ G1.T)
Mt = M1 @ E @ np.linalg.solve(MII, E.T @ M1)
lam_s = (s1 @ K1 @ s1) / (s1 @ M1 @ s1)
I1 = np.eye(n1)
Zx = np.kron(W, I1); Zy = np.kron(I1, W); Z = np.hstack((Zx, Zy))
rng = np.random.default_rng(0)
xt = rng.standard_normal(N); b = S @ xt
tol = 1e-10 * np.linalg.norm(b)
A = lambda x: S @ x
def mult(coarse):
    def f(r):
        z = coarse(r)
        return z + (r - S @ z) / Mp
    return f
def galerkin(Sh):
    Sci = np.linalg.pinv(Z.T @ Sh @ Z, rcond=1e-9)
    return lambda r: Z @ (Sci @ (Z.T @ r))
Rs, R0 = R1(lam_s), R1(0.0)
Sh1 = np.kron(Rs, Mt) + np.kron(Mt, Rs)
# block model: diagonal blocks with R(0) for the W-direction operator, cross blocks with R(lam_s)
nx_ = (ne + 1) * n1
Cx = W.T @ R0 @ W; Dx = W.T @ Mt @ W
Sxx = np.kron(Cx, Mt) + np.kron(Dx, Rs)
Syy = np.kron(Mt, Cx) + np.kron(Rs, Dx)
Sxy = np.kron(W.T @ Rs, Mt @ W) + np.kron(W.T @ Mt, Rs @ W)
Sc2 = np.block([[Sxx, Sxy], [Sxy.T, Syy]])
Sc2i = np.linalg.pinv(Sc2, rcond=1e-9)
c2 = lambda r: Z @ (Sc2i @ (Z.T @ r))
Sc_exact = Z.T @ S @ Z
for nm, Mdl in (('Z^T Shat Z', Z.T @ Sh1 @ Z), ('block model', Sc2)):
    ev = np.sort(np.linalg.eigvals(np.linalg.pinv(Mdl, rcond=1e-9) @ Sc_exact).real)
    ev = ev[np.abs(ev) > 1e-8]
    print(f'{nm}: gen eig of exact coarse vs model: min {ev[0]:.3f} 5% {ev[len(ev) // 20]:.3f} median {ev[len(ev) // 2]:.3f} 95% {ev[-len(ev) // 20]:.3f} max {ev[-1]:.3f}')
for name, Pi in (('mass', lambda r: r / Mp), ('exact Galerkin', mult(galerkin(S))), ('Galerkin of Shat(lam_s)', mult(galerkin(Sh1))),
                 ('block model R0/Rs', mult(c2))):
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, min(N, 1200))
    print(f'{name:26s} its {its:5d} relres {np.linalg.norm(b - S @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}',
          [f'{h_:.0e}' for h_ in hist[::max(1, len(hist) // 8)]])
print('--- coarse space restricted to interior pressure nodes')
inner = (~ns._mask_bound).astype(float); inner[ns._pin] = 0
Zi = Z * inner[:, None]
def galerkin_Z(Sh, Zm):
    Sci = np.linalg.pinv(Zm.T @ Sh @ Zm, rcond=1e-9)
    return lambda r: Zm @ (Sci @ (Zm.T @ r))
ev = np.sort(np.linalg.eigvals(np.linalg.pinv(Zi.T @ Sh1 @ Zi, rcond=1e-9) @ (Zi.T @ S @ Zi)).real); ev = ev[np.abs(ev) > 1e-8]
print(f'gen eig exact vs Shat on interior coarse: min {ev[0]:.3f} 5% {ev[len(ev) // 20]:.3f} median {ev[len(ev) // 2]:.3f} 95% {ev[-len(ev) // 20]:.3f} max {ev[-1]:.3f}')
for name, Pi in (('exact Galerkin (interior Z)', mult(galerkin_Z(S, Zi))), ('Galerkin of Shat (interior Z)', mult(galerkin_Z(Sh1, Zi)))):
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, min(N, 1200))
    print(f'{name:30s} its {its:5d} relres {np.linalg.norm(b - S @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}',
          [f'{h_:.0e}' for h_ in hist[::max(1, len(hist) // 8)]])
print('--- reduced system (boundary / pin pressure rows eliminated exactly)')
bset = ns._mask_bound.copy(); bset[ns._pin] = True
Ii = np.where(~bset)[0]; Bi = np.where(bset)[0]
Sred = S[np.ix_(Ii, Ii)] - S[np.ix_(Ii, Bi)] @ np.linalg.solve(S[np.ix_(Bi, Bi)], S[np.ix_(Bi, Ii)])
MI = Mp[Ii]
ZI = Z[Ii]
xt2 = rng.standard_normal(Ii.size); b2 = Sred @ xt2; tol2 = 1e-10 * np.linalg.norm(b2)
A2 = lambda x: Sred @ x
def mult2(coarse):
    def f(r):
        z = coarse(r)
        return z + (r - Sred @ z) / MI
    return f
def gal2(Sh):
    Sci = np.linalg.pinv(ZI.T @ Sh @ ZI, rcond=1e-9)
    return lambda r: ZI @ (Sci @ (ZI.T @ r))
Sh_II = Sh1[np.ix_(Ii, Ii)]
ev = np.sort(np.linalg.eigvals(np.linalg.pinv(ZI.T @ Sh_II @ ZI, rcond=1e-9) @ (ZI.T @ Sred @ ZI)).real); ev = ev[np.abs(ev) > 1e-8]
print(f'gen eig Sred vs Shat_II on coarse: min {ev[0]:.3f} 5% {ev[len(ev) // 20]:.3f} median {ev[len(ev) // 2]:.3f} 95% {ev[-len(ev) // 20]:.3f} max {ev[-1]:.3f}')
for name, Pi in (('mass (reduced)', lambda r: r / MI), ('exact Galerkin (reduced)', mult2(gal2(Sred))), ('Galerkin of Shat_II (reduced)', mult2(gal2(Sh_II)))):
    t = time.time()
    x, its, hist = gmres(A2, b2, Pi, tol2, min(Ii.size, 1200))
    print(f'{name:30s} its {its:5d} relres {np.linalg.norm(b2 - Sred @ x) / np.linalg.norm(b2):.2e} t {time.time() - t:.1f}',
          [f'{h_:.0e}' for h_ in hist[::max(1, len(hist) // 8)]])
print('--- fully structured coarse stage on the reduced system: z1 = Pi Shat_II^+ Pi^T r, Pi = P_W(x)I + I(x)P_W - P_W(x)P_W')
nI = n1 - 2
WI = E.T @ W                                  # interior rows of W
R_II = E.T @ Rs @ E; M_II1 = MII
rho, Vv = sla.eigh(R_II, M_II1)               # R v = rho M v, V^T M V = I
print('rho range', rho[:4], rho[-2:])
def make_stage(metric):
    if metric == 'euclid':
        PW = WI @ np.linalg.pinv(WI.T @ WI) @ WI.T
        PWt = PW
    else:                                     # M-orthogonal projector (acts on pressure vectors), transpose acts on residuals
        PW = WI @ np.linalg.pinv(WI.T @ M_II1 @ WI) @ WI.T @ M_II1
        PWt = PW.T
    den = rho[:, None] + rho[None, :]
    inv = np.where(den > 1e-9 * den.max(), 1.0 / np.where(den > 0, den, 1), 0.0)
    def proj(Xm, Pm):                         # (P(x)I + I(x)P - P(x)P) X
        PX = Pm @ Xm
        return PX + Xm @ Pm.T - PX @ Pm.T
    def f(r):
        Rm = r.reshape(nI, nI)
        Rm = proj(Rm, PWt)
        Y = Vv.T @ Rm @ Vv
        Y = Y * inv
        X = Vv @ Y @ Vv.T
        return proj(X, PW).ravel()
    return f
# interior nodes in Ii are ordered (ix, iy) over the interior minus the pin: handle the pin by padding
full_int = np.where(~ns._mask_bound)[0]
pos = {g_: k for k, g_ in enumerate(full_int)}
pin_k = pos[ns._pin]
def wrap(stage):
    def f(r):
        rr = np.insert(r, pin_k, 0.0)
        z = stage(rr)
        return np.delete(z, pin_k)
    return f
for metric in ('euclid', 'mass'):
    st = wrap(make_stage(metric))
    for name, Pi in ((f'structured ({metric}) -> mass', mult2(st)),):
        t = time.time()
        x, its, hist = gmres(A2, b2, Pi, tol2, min(Ii.size, 1200))
        print(f'{name:30s} its {its:5d} relres {np.linalg.norm(b2 - Sred @ x) / np.linalg.norm(b2):.2e} t {time.time() - t:.1f}',
              [f'{h_:.0e}' for h_ in hist[::max(1, len(hist) // 8)]])
