"""Throw-away: GMRES iteration counts on the dense Schur complement with candidate right preconditioners."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla, scipy.linalg as sla
from p2 import build, dense_schur

def gmres(A, b, Pinv, tol, maxit):
    n = b.size
    r = b.copy(); beta = np.linalg.norm(r)
    V = np.zeros((maxit + 1, n)); Z = np.zeros((maxit, n)); H = np.zeros((maxit + 1, maxit))
    V[0] = r / beta; g = np.zeros(maxit + 1); g[0] = beta
    cs = np.zeros(maxit); sn = np.zeros(maxit); hist = [beta]
    for k in range(maxit):
        Z[k] = Pinv(V[k]); w = A(Z[k])
        for _ in range(2):
            h = V[:k + 1] @ w; w -= h @ V[:k + 1]; H[:k + 1, k] += h
        H[k + 1, k] = np.linalg.norm(w); V[k + 1] = w / H[k + 1, k]
        for i in range(k):
            t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]; H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]; H[i, k] = t
        d = np.hypot(H[k, k], H[k + 1, k]); cs[k] = H[k, k] / d; sn[k] = H[k + 1, k] / d
        H[k, k] = d; H[k + 1, k] = 0; g[k + 1] = -sn[k] * g[k]; g[k] = cs[k] * g[k]
        hist.append(abs(g[k + 1]))
        if abs(g[k + 1]) <= tol: break
    y = np.linalg.solve(np.triu(H[:k + 1, :k + 1]), g[:k + 1])
    return y @ Z[:k + 1], k + 1, hist

if __name__ == '__main__':
    P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); stokes = len(sys.argv) > 4 and sys.argv[4] == 's'
    ns, J = build(P, ne, Re, stokes)
    N = ns.N
    S, lu = dense_schur(ns, J)
    Mp = ns._M.copy(); Mp[ns._pin] = 1
    mb = ns._mask_bound; inner = (~mb).astype(float)
    Gx, Gy = ns._G_x, ns._G_y
    Mu_inv = inner / ns._M
    # L = D Mu^-1 G on the full pressure space (velocity interior)
    L = (Gx @ sps.diags(Mu_inv) @ Gx + Gy @ sps.diags(Mu_inv) @ Gy).toarray()
    print('L symmetry', np.abs(L - L.T).max(), 'L eig range', np.linalg.eigvalsh((L + L.T) / 2)[[0, 1, 2, -1]])
    Lp = np.linalg.pinv(L, rcond=1e-11, hermitian=True)
    Aa = J[:2 * N, :2 * N]
    G2 = sps.vstack((sps.diags(inner) @ Gx, sps.diags(inner) @ Gy)).tocsr()
    D2 = sps.hstack((Gx @ sps.diags(inner), Gy @ sps.diags(inner))).tocsr()
    Mu2 = np.hstack((Mu_inv, Mu_inv))
    def F(z):
        return D2 @ (Mu2 * (Aa @ (Mu2 * (G2 @ z))))
    def lsc(r):
        return -(Lp @ F(Lp @ r))
    rng = np.random.default_rng(0)
    xt = rng.standard_normal(N); b = S @ xt     # consistent rhs
    tol = 1e-10 * np.linalg.norm(b)
    A = lambda x: S @ x
    def mass(r): return r / Mp
    def lsc_b(r):      # LSC on interior rows, mass on boundary rows
        z = lsc(r * inner); z += (r * mb) / Mp; return z
    def lsc_plus_mass(r): return lsc(r) + r / Mp
    for name, Pi in (('mass', mass), ('lsc', lsc), ('lsc_b', lsc_b), ('lsc+mass', lsc_plus_mass)):
        t = time.time()
        x, its, hist = gmres(A, b, Pi, tol, min(N, 1500))
        print(f'{name:10s} its {its:5d} relres {np.linalg.norm(b - S @ x) / np.linalg.norm(b):.2e} t {time.time() - t:.1f}',
              [f'{h:.0e}' for h in hist[::max(1, len(hist) // 8)]])
