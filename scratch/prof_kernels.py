"""Small drivers for the ncu captures of round 2 (profiles/README.md).
   python scratch/prof_kernels.py fdm      -- fast-diagonalisation apply at 4097^2 nodes (k_dgemm, big tile), 2 fields, 3 applies
   python scratch/prof_kernels.py krylov   -- Jacobi-GMRES(64) on the CD operator at 4097^2 = 16.8 M nodes, 130 iterations
                                              (k_multi_dot / k_multi_axpy with up to 64 basis vectors of 135 MB)"""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
import sem_b200
what = sys.argv[1]
P, ne = 8, 512
if what == 'fdm':
    d = sem_b200.SemDevice(P, ne, ne, 1.0 / ne, 1.0 / ne)
    d.setup_fdm([1, 1, 1, 1])
    r = d.zeros(2)
    r[:, :, :d.NY] = torch.randn((2, d.NX, d.NY), device=d.tdev, dtype=torch.float64)
    z = d.zeros(2)
    for _ in range(3):
        d.fdm_apply(0, r, z, 2)
    torch.cuda.synchronize()
    print('fdm ok', float(z.abs().max()))
else:
    cd = sem_b200.ConvectionDiffusionSolver(1, 1, 40.0, P, ne, ne, T_W=0.5, T_E=-0.5, mtol=1e-12, restart=64, precond='jacobi')
    cd._krylov()
    u = cd._get_vector(lambda x, y: y - 0.5)
    v = cd._get_vector(lambda x, y: 0.5 - x)
    import ctypes as C
    from sem_b200 import _lib as L
    d = cd._dev
    d.to_device(u, cd._u); d.to_device(v, cd._v); cd._have_sys = True
    rhs = d.zeros(); rhs[:, :d.NY] = torch.randn((d.NX, d.NY), device=d.tdev, dtype=torch.float64)
    x = d.zeros()
    kr = cd._krylov(); kr.max_iters = 130
    st = cd._state(with_jac=False)
    code = cd._lib.sem_cd_solve(d.ctx, C.byref(st), rhs.data_ptr(), x.data_ptr(), C.byref(kr), cd._work.data_ptr(), cd._work.numel(), d.stream)
    torch.cuda.synchronize()
    print('krylov ok: code', code, 'iterations', kr.iters, 'residual', kr.resnorm)
