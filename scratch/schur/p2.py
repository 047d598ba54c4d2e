"""Throw-away: spectrum of M_p^-1 S."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
from oracle import sem_oracle as so

def build(P, ne, Re, stokes=False):
    ns = so.NSOracle(1.0, 1.0, Re, 0.0, P, ne, ne, u_N=1.0, mtol=1e-13, mtol_newton=1e-10)
    N = ns.N
    T = np.zeros(N)
    if stokes:
        u = v = p = np.zeros(N)
    else:
        u, v, p = ns._get_solution(T, max_newton=3)
    ns._get_residuals(u, v, p, T); ns._calc_jacobians(u, v)
    J = ns.jacobian_matrix().tocsr()
    return ns, J

def dense_schur(ns, J):
    N = ns.N
    Aa = J[:2 * N, :2 * N].tocsc(); B = J[:2 * N, 2 * N:]; C = J[2 * N:, :2 * N].tocsr(); D = J[2 * N:, 2 * N:]
    lu = spla.splu(Aa)
    X = lu.solve(B.toarray())
    return D.toarray() - C @ X, lu

if __name__ == '__main__':
    P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); stokes = len(sys.argv) > 4
    ns, J = build(P, ne, Re, stokes)
    S, lu = dense_schur(ns, J)
    N = ns.N
    Mp = ns._M.copy(); Mp[ns._pin] = 1
    ev = np.linalg.eigvals(S / Mp[None, :])   # S M^-1 (right precond)
    a = np.sort(np.abs(ev))
    print('N', N, 'elements', ne * ne)
    print('smallest |ev|', a[:12])
    print('largest', a[-5:])
    for th in (1e-3, 1e-2, 3e-2, 0.1, 0.3, 0.5):
        print(f'  #|ev| < {th}: {np.sum(a < th)}')
    np.save(f'/tmp/S_{P}_{ne}_{int(Re)}{"s" if stokes else ""}.npy', S)
