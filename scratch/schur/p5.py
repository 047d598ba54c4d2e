import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/scratch/schur')
import numpy as np, scipy.sparse as sps
from p2 import build, dense_schur
from p4 import gmres
P = int(sys.argv[1]); ne = int(sys.argv[2]); Re = float(sys.argv[3]); stokes = len(sys.argv) > 4 and sys.argv[4] == 's'
ns, J = build(P, ne, Re, stokes)
N = ns.N
S, lu = dense_schur(ns, J)
Mp = ns._M.copy(); Mp[ns._pin] = 1
mb = ns._mask_bound.copy()
bset = mb.copy(); bset[ns._pin] = True      # rows that are not continuity rows
I = np.where(~bset)[0]; B = np.where(bset)[0]
Sred = S[np.ix_(I, I)] - S[np.ix_(I, B)] @ np.linalg.solve(S[np.ix_(B, B)], S[np.ix_(B, I)])
ev = np.sort(np.abs(np.linalg.eigvals(Sred / Mp[I][None, :])))
print('reduced: smallest', ev[:8], 'largest', ev[-4:])
for th in (1e-3, 1e-2, 3e-2, 0.1, 0.3): print(f'  #<{th}: {np.sum(ev < th)}')
# how do the interior rows alone look (S_II)?
ev = np.sort(np.abs(np.linalg.eigvals(S[np.ix_(I, I)] / Mp[I][None, :])))
print('S_II: smallest', ev[:8], 'largest', ev[-4:])
for th in (1e-3, 1e-2, 3e-2, 0.1, 0.3): print(f'  #<{th}: {np.sum(ev < th)}')
