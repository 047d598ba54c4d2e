"""Multiplicative three-stage Schur preconditioner: coarse_x -> coarse_y -> mass; exact and Kronecker-sum coarse blocks."""
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
M1 = so._assembled_1d(h / 2 * so.mass_1d(P), ne).toarray()
K1 = so._assembled_1d(2 / h * so.stiff_1d(P), ne).toarray()
G1 = so._assembled_1d(so.grad_1d(P), ne).toarray()
I = np.arange(1, n1 - 1)
E = np.zeros((n1, n1 - 2)); E[I, np.arange(n1 - 2)] = 1
KII = E.T @ K1 @ E; MII = E.T @ M1 @ E
def R1(mu):
    return G1 @ E @ np.linalg.solve(KII + mu * MII, E.T @ G1.T)
Mt = M1 @ E @ np.linalg.solve(MII, E.T @ M1)
lam_s = (s1 @ K1 @ s1) / (s1 @ M1 @ s1)
I1 = np.eye(n1)
Zx = np.kron(W, I1); Zy = np.kron(I1, W)
rng = np.random.default_rng(0)
xt = rng.standard_normal(N); b = S @ xt
tol = 1e-10 * np.linalg.norm(b)
A = lambda x: S @ x
def chain(*stages):
    def f(r):
        z = np.zeros_like(r); rr = r.copy()
        for k, st in enumerate(stages):
            dz = st(rr); z += dz
            if k + 1 < len(stages): rr = r - S @ z
        return z
    return f
mass = lambda r: r / Mp
Sxi = np.linalg.pinv(Zx.T @ S @ Zx, rcond=1e-10); Syi = np.linalg.pinv(Zy.T @ S @ Zy, rcond=1e-10)
cx = lambda r: Zx @ (Sxi @ (Zx.T @ r)); cy = lambda r: Zy @ (Syi @ (Zy.T @ r))
# Kronecker-sum models of the coarse blocks:  Zx^T S Zx ~ Cx (x) Mt + Dx (x) Ry
Cx = W.T @ R1(0.0) @ W; Dx = W.T @ Mt @ W; Ry = R1(lam_s)
def ksum_inv(A1, B1, A2, B2):
    """inverse (pseudo) of A1 (x) B2 + B1 (x) A2 ... here: Cx (x) Mt + Dx (x) Ry, via two generalised eigenproblems"""
    # x: Cx v = a Dx v   (Dx SPD?)   y: Ry u = b Mt u  (Mt singular at boundary nodes -> regularise)
    return None
Sx_model = np.kron(Cx, Mt) + np.kron(Dx, Ry)
Sxx = Zx.T @ S @ Zx
# compare on interior hats / interior y nodes, away from the pin
ky = np.arange(1, n1 - 1); kx = np.arange(1, ne)
pin_x, pin_y = divmod(ns._pin, n1)
idx = np.array([a * n1 + b_ for a in kx for b_ in ky if not (abs(a * P - pin_x) <= P and b_ == pin_y)])
A_ = Sxx[np.ix_(idx, idx)]; B_ = Sx_model[np.ix_(idx, idx)]
print('interior coarse block: ||S||', np.linalg.norm(A_), '||model||', np.linalg.norm(B_), 'rel err', np.linalg.norm(A_ - B_) / np.linalg.norm(A_))
ev = np.sort(np.linalg.eigvals(np.linalg.solve(B_, A_)).real); print('  gen eig range', ev[[0, 1, len(ev) // 2, -2, -1]])
Smi = np.linalg.pinv(Sx_model, rcond=1e-10)
mx = lambda r: Zx @ (Smi @ (Zx.T @ r)); my = lambda r: (Zy @ (Smi_y @ (Zy.T @ r)))
# y block: same with roles swapped: Zy = I (x) W -> Mt (x) Cx + Ry (x) Dx
Smi_y = np.linalg.pinv(np.kron(Mt, Cx) + np.kron(Ry, Dx), rcond=1e-10)
for name, Pi in (('mass', mass), ('exact x->y->mass', chain(cx, cy, mass)), ('exact x->y->x->mass', chain(cx, cy, cx, mass)),
                 ('model x->y->mass', chain(mx, my, mass)), ('model x->y->x->y->mass', chain(mx, my, mx, my, mass))):
    t = time.time()
    x, its, hist = gmres(A, b, Pi, tol, min(N, 1200))
    print(f'{name:24s} its {its:5d} relres {np.linalg.norm(b - S @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}',
          [f'{h_:.0e}' for h_ in hist[::max(1, len(hist) // 8)]])
