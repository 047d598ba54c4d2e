import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import sem_oracle as so
P, ne = 4, 8
S = np.load('/tmp/S_4_8_400s.npy')
ns = so.NSOracle(1.0, 1.0, 400.0, 0.0, P, ne, ne, u_N=1.0)
N = ns.N; NX = ne * P + 1
Mp = ns._M.copy(); Mp[ns._pin] = 1
w, V = np.linalg.eig(S / Mp[None, :])
idx = np.argsort(np.abs(w))
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for k in idx[:6]:
    y = np.real(V[:, k]); x = y / Mp       # pressure-space vector
    X = x.reshape(NX, NX)
    # element blocks
    E = np.zeros((ne * ne, (P + 1) ** 2))
    for m in range(ne):
        for n in range(ne):
            E[m * ne + n] = X[m * P:m * P + P + 1, n * P:n * P + P + 1].ravel()
    sv = np.linalg.svd(E, compute_uv=False)
    print('ev', w[k], 'element-block singular values', sv[:5] / sv[0])
    U, s, Vt = np.linalg.svd(E)
    print(' pattern\n', (Vt[0] / np.abs(Vt[0]).max()).reshape(P + 1, P + 1))
    print(' envelope\n', (U[:, 0] / np.abs(U[:, 0]).max()).reshape(ne, ne))
