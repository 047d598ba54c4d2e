"""Throw-away prototype: monolithic right-preconditioned GMRES on the NS Jacobian with block lower-triangular
preconditioner [[P_a,0],[C,M_p]] -- iteration counts for different P_a choices (CPU, oracle matrices)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
from oracle import sem_oracle as so
from tests.conftest import load_golden

def gmres_right(A, b, Pinv, tol, maxit, x0=None):
    n = b.size
    x = np.zeros(n) if x0 is None else x0.copy()
    r = b - A(x); beta = np.linalg.norm(r)
    V = np.zeros((maxit + 1, n)); Z = np.zeros((maxit, n)); H = np.zeros((maxit + 1, maxit))
    V[0] = r / beta
    g = np.zeros(maxit + 1); g[0] = beta
    cs = np.zeros(maxit); sn = np.zeros(maxit)
    hist = [beta]
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
    x += y @ Z[:k + 1]
    return x, k + 1, hist

def setup(tag, kw):
    g = load_golden('ns')
    ns = so.NSOracle(mtol=1e-13, mtol_newton=1e-13, **kw)
    return ns, g

if __name__ == '__main__':
    case = sys.argv[1] if len(sys.argv) > 1 else 'c3'
    from tests.golden.make_golden_cases import NS_CASES
    kw = [c for c in NS_CASES if c[0] == case][0][1]
    ns, g = setup(case, kw)
    N = ns.N
    Tin = g[f'{case}/T_in']; us, vs, ps = (g[f'{case}/{k}_sol'] for k in 'uvp')
    # linearise about a half-way state (harder than the start, representative of Newton steps)
    u, v, p = 0.7 * us, 0.7 * vs, 0.7 * ps
    ru, rv, rc = ns._get_residuals(u, v, p, Tin); ns._calc_jacobians(u, v)
    b = -np.hstack((ru, rv, rc))
    J = ns.jacobian_matrix().tocsr()
    A = lambda x: J @ x
    Jd = J.diagonal()
    Aa = J[:2 * N, :2 * N].tocsc(); C = J[2 * N:, :2 * N].tocsr()
    Mp = ns._M.copy(); Mp[ns._pin] = 1.0
    lu = spla.splu(Aa)
    def tri(Pa_inv):
        def f(r):
            za = Pa_inv(r[:2 * N]); zp = (r[2 * N:] - C @ za) / Mp
            return np.hstack((za, zp))
        return f
    tol = 1e-13 * np.sqrt(3 * N)
    print('N', N, '|b|', np.linalg.norm(b), 'tol', tol)
    for name, Pa in (('exactLU', lu.solve), ('jacobi', lambda r: r / Jd[:2 * N])):
        t = time.time()
        x, its, hist = gmres_right(A, b, tri(Pa), tol, 6000 if name == 'jacobi' else 1500)
        print(name, 'its', its, 'res', np.linalg.norm(b - J @ x), 'time', time.time() - t,
              'constraint', np.sum((ns._M * x[2 * N:])[~ns._mask_bound & (np.arange(N) != ns._pin)]))
        print('   hist', [f'{h:.1e}' for h in hist[::max(1, len(hist) // 12)]])
