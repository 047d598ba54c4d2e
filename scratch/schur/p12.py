"""Round-2 blueprint, end to end on the CPU oracle: monolithic right-preconditioned GMRES on the 3N x 3N NS Jacobian with
the block lower-triangular preconditioner  [[P_a, 0], [C, S~]]:
  P_a  = Stokes velocity block (what the device's fast diagonalisation inverts exactly), identity boundary rows
  S~^-1: (a) the reference's mass preconditioner (today's device solver)
         (b) structured two-level: boundary/pin pressure rows by block elimination with K_BB, interior by
             z1 = Pi Shat^+ Pi^T y,  z = z1 + M^-1 (y - S_stokes z1)   [S_stokes applied with P_a: no exact Jacobian solves]
         + the rank-one member-selection correction.
Reports iteration counts; checks that (b) lands on the reference's member."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla, scipy.linalg as sla
from numpy.polynomial import legendre as npl
from oracle import sem_oracle as so
from proto_ns_krylov import gmres_right

P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); newton = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ns = so.NSOracle(1.0, 1.0, Re, 0.0, P, ne, ne, u_N=1.0, mtol=1e-13, mtol_newton=1e-13)
N = ns.N; n1 = ne * P + 1; h = 1.0 / ne; nI = n1 - 2
T = np.zeros(N)
u, v, p = ns._get_solution(T, max_newton=newton) if newton > 0 else (np.zeros(N),) * 3
ru, rv, rc = ns._get_residuals(u, v, p, T); ns._calc_jacobians(u, v)
b = -np.hstack((ru, rv, rc))
J = ns.jacobian_matrix().tocsr()
x_ref = np.hstack(ns._get_update(-ru, -rv, -rc))
l = ns._left_null(J.tocsc()); lc = l[2 * N:]
Mp = ns._M.copy(); Mp[ns._pin] = 1.0; mc = Mp * lc
# Stokes Jacobian blocks (state 0): what the device preconditioner can invert exactly
ns0 = so.NSOracle(1.0, 1.0, Re, 0.0, P, ne, ne, u_N=1.0)
z0 = np.zeros(N); ns0._get_residuals(z0, z0, z0, z0); ns0._calc_jacobians(z0, z0)
J0 = ns0.jacobian_matrix().tocsr()
A0 = J0[:2 * N, :2 * N].tocsc(); G0 = J0[:2 * N, 2 * N:].tocsr(); C0 = J0[2 * N:, :2 * N].tocsr(); D0 = J0[2 * N:, 2 * N:].tocsr()
luA0 = spla.splu(A0)
C = J[2 * N:, :2 * N].tocsr()
S0 = lambda q: D0 @ q - C0 @ luA0.solve(G0 @ q)          # Stokes Schur complement, applied matrix-free
# pressure index sets
bset = ns._mask_bound.copy(); bset[ns._pin] = True
Ii = np.where(~bset)[0]; Bi = np.where(bset)[0]
DBB = D0[Bi][:, Bi].tocsc(); DBI = D0[Bi][:, Ii].tocsr()
luBB = spla.splu(DBB)
# 1-D pieces of the structured stage
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
Rs = G1 @ E @ np.linalg.solve(KII + 0.25 * lam_s * MII, E.T @ G1.T)
rho, Vv = sla.eigh(E.T @ Rs @ E, MII)
WI = E.T @ W
PW = WI @ np.linalg.pinv(WI.T @ MII @ WI) @ WI.T @ MII
den = rho[:, None] + rho[None, :]
inv = np.where(den > 1e-9 * den.max(), 1.0 / np.where(den > 0, den, 1), 0.0)
def proj(Xm, Pm):
    PX = Pm @ Xm
    return PX + Xm @ Pm.T - PX @ Pm.T
full_int = np.where(~ns._mask_bound)[0]
int_pos = np.searchsorted(full_int, Ii)                 # positions of Ii inside the (n1-2)^2 interior grid
def coarse_interior(yI):
    Y = np.zeros(nI * nI); Y[int_pos] = yI
    Rm = proj(Y.reshape(nI, nI), PW.T)
    X = Vv @ ((Vv.T @ Rm @ Vv) * inv) @ Vv.T
    return proj(X, PW).ravel()[int_pos]
def schur_two_level(y):
    """block elimination of the boundary / pin rows (K_BB), structured two-level on the interior"""
    z = np.zeros(N)
    # interior: right-hand side with the boundary part eliminated approximately (S_IB neglected)
    z1 = np.zeros(N); z1[Ii] = coarse_interior(y[Ii])
    z1[Bi] = luBB.solve(y[Bi] - DBI @ z1[Ii])
    r1 = y - S0(z1)
    z2 = np.zeros(N); z2[Ii] = r1[Ii] / Mp[Ii]
    z2[Bi] = luBB.solve(r1[Bi] - DBI @ z2[Ii])
    return z1 + z2
def corrected(Sinv):
    def f(y):
        z = Sinv(y)
        return z - lc * ((mc @ z - lc @ y) / (mc @ lc))
    return f
def tri(Sinv):
    def f(r):
        za = luA0.solve(r[:2 * N]); zp = Sinv(r[2 * N:] - C @ za)
        return np.hstack((za, zp))
    return f
mass = lambda y: y / Mp
def mass_bb(y):     # mass on the interior, K_BB block elimination on the boundary rows
    z = np.zeros(N); z[Ii] = y[Ii] / Mp[Ii]; z[Bi] = luBB.solve(y[Bi] - DBI @ z[Ii]); return z
tol = 1e-10 * np.linalg.norm(b)
print(f'P={P} ne={ne} Re={Re} newton state {newton}: N={N}, |b|={np.linalg.norm(b):.2e}')
cands = [('mass (today)', mass), ('mass + K_BB elimination', corrected(mass_bb)), ('structured two-level', corrected(schur_two_level))]
if len(sys.argv) > 5: cands = cands[int(sys.argv[5]):]
for name, Si in cands:
    t = time.time()
    x, its, hist = gmres_right(lambda z: J @ z, b, tri(Si), tol, 700 if N > 10000 else 1500)
    print(f'{name:28s} its {its:5d} |Jx-b|/|b| {np.linalg.norm(J @ x - b) / np.linalg.norm(b):.1e}  p diff vs reference member '
          f'{np.linalg.norm(x[2*N:] - x_ref[2*N:]) / np.linalg.norm(x_ref[2*N:]):.1e}  vel diff {np.linalg.norm(x[:2*N] - x_ref[:2*N]) / max(np.linalg.norm(x_ref[:2*N]), 1e-300):.1e}  t {time.time() - t:.1f}', flush=True)
# ---- with a pressure-convection-diffusion fine stage instead of the mass sweep (interior), Re > 0
if newton > 0:
    K = ns._K.tocsr()
    Fp = (K + Re * (sps.diags(ns._u) @ ns._G_x + sps.diags(ns._v) @ ns._G_y)).tocsr()
    KIB_BBinv_DBI = None
    # reduced operators on the interior (boundary values slaved by the Neumann rows): X_red = X_II - X_IB K_BB^-1 K_BI
    def red(X):
        XII = X[Ii][:, Ii].toarray(); XIB = X[Ii][:, Bi].toarray()
        return XII - XIB @ luBB.solve(DBI.toarray())
    Kred = red(K); Fred = red(Fp)
    Kri = np.linalg.pinv(Kred, rcond=1e-11)
    def schur_two_level_pcd(y):
        z1 = np.zeros(N); z1[Ii] = coarse_interior(y[Ii])
        z1[Bi] = luBB.solve(y[Bi] - DBI @ z1[Ii])
        r1 = y - S0(z1)
        z2 = np.zeros(N); z2[Ii] = (Fred @ (Kri @ r1[Ii])) / Mp[Ii]
        z2[Bi] = luBB.solve(r1[Bi] - DBI @ z2[Ii])
        return z1 + z2
    def pcd_bb(y):
        z = np.zeros(N); z[Ii] = (Fred @ (Kri @ y[Ii])) / Mp[Ii]; z[Bi] = luBB.solve(y[Bi] - DBI @ z[Ii]); return z
    luA = spla.splu(J[:2 * N, :2 * N].tocsc())
    def tri_exact(Sinv):
        def f(r):
            za = luA.solve(r[:2 * N]); zp = Sinv(r[2 * N:] - C @ za)
            return np.hstack((za, zp))
        return f
    for name, pre in (('PCD + K_BB (Stokes velocity block)', tri(corrected(pcd_bb))),
                      ('two-level + PCD (Stokes velocity block)', tri(corrected(schur_two_level_pcd))),
                      ('mass, EXACT velocity block', tri_exact(mass)),
                      ('two-level + PCD, EXACT velocity block', tri_exact(corrected(schur_two_level_pcd)))):
        t = time.time()
        x, its, hist = gmres_right(lambda z: J @ z, b, pre, tol, 1500)
        print(f'{name:42s} its {its:5d} |Jx-b|/|b| {np.linalg.norm(J @ x - b) / np.linalg.norm(b):.1e}  p diff '
              f'{np.linalg.norm(x[2*N:] - x_ref[2*N:]) / np.linalg.norm(x_ref[2*N:]):.1e}  t {time.time() - t:.1f}', flush=True)
# ---- the boundary block without a factorisation: K_BB is uniformly well conditioned after diagonal scaling (spectrum in
#      [0.58, 1.46] for every mesh), so a fixed Chebyshev polynomial replaces the exact solve
if len(sys.argv) > 6:
    dB = DBB.diagonal()
    lo, hi = 0.55, 1.50
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    def cheb_solve(rhs, steps):
        # Chebyshev iteration for (D^-1/2 K D^-1/2) y = D^-1/2 rhs, x = D^-1/2 y  (fixed step count: a linear operator)
        bb = rhs / np.sqrt(dB)
        Aop = lambda y: (DBB @ (y / np.sqrt(dB))) / np.sqrt(dB)
        y = np.zeros_like(bb); r = bb.copy()
        sigma = theta / delta; rho_ = 1.0 / sigma
        d = r / theta
        for k in range(steps):
            y = y + d
            r = r - Aop(d)
            rho_new = 1.0 / (2.0 * sigma - rho_)
            d = rho_new * rho_ * d + (2.0 * rho_new / delta) * r
            rho_ = rho_new
        return y / np.sqrt(dB)
    for steps in (3, 5, 8):
        def mass_bb_cheb(y, steps=steps):
            z = np.zeros(N); z[Ii] = y[Ii] / Mp[Ii]; z[Bi] = cheb_solve(y[Bi] - DBI @ z[Ii], steps); return z
        t = time.time()
        x, its, hist = gmres_right(lambda z: J @ z, b, tri(mass_bb_cheb), tol, 1500)
        print(f'mass + K_BB by {steps}-step Chebyshev      its {its:5d} |Jx-b|/|b| {np.linalg.norm(J @ x - b) / np.linalg.norm(b):.1e}  p diff '
              f'{np.linalg.norm(x[2*N:] - x_ref[2*N:]) / np.linalg.norm(x_ref[2*N:]):.1e}  t {time.time() - t:.1f}', flush=True)
