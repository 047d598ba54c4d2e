"""Body of test_partitioned_apply_loopback_exchange (also runnable as a child process with another exchange path selected
by environment variables, which the library reads once per process)."""
import ctypes as C
import os
import sys

import numpy as np
import torch


def check(sem, P, nex, ney, expect_fused):
    from sem_b200 import _lib as L
    N_ex = 3 * nex
    dx, dy = 1.0 / N_ex, 1.0 / ney
    plain = sem.SemDevice(P, N_ex, ney, dx, dy, m_begin=nex, m_end=2 * nex)
    loop = sem.SemDevice(P, N_ex, ney, dx, dy, m_begin=nex, m_end=2 * nex)
    L.check(loop.lib.sem_ctx_attach_loopback(loop.ctx), "sem_ctx_attach_loopback")
    assert loop.comm_mode == "p2p"
    gen = torch.Generator(device=plain.tdev).manual_seed(1234 + P)

    def rnd(d):
        x = d.zeros()
        x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64)
        return x

    x, u, v, du, dv, gx, gy = (rnd(plain) for _ in range(7))

    def expected(y):
        e = y.clone()
        s = y[0] + y[-1]
        e[0] = s
        e[-1] = s
        return e

    def both(fn, nout):
        outs = []
        for d in (plain, loop):
            ys = [torch.full_like(x, float("nan")) for _ in range(nout)]
            for y in ys:
                y[:, d.NY:] = 0.0
            for _rep in range(5 if d is loop else 1):   # repeated exchanges: epochs advance, both parities are used
                fn(d, ys)
            torch.cuda.synchronize()
            outs.append(ys)
        for yp, yl in zip(*outs):
            assert torch.isfinite(yl).all()
            assert torch.equal(expected(yp), yl), float((expected(yp) - yl).abs().max())

    both(lambda d, ys: d.apply_stiffness(x, ys[0]), 1)
    both(lambda d, ys: d.apply_gradient(x, ys[0], ys[1], 1.7), 2)

    def cd(d, ys, residual, pointwise):
        st = L.sem_cd_state()
        bc = L.sem_cd_bc()
        for k, (a, val) in enumerate([(1, 0.5), (1, -0.5), (1, 0.25), (0, 0.0)]):
            bc.active[k], bc.value[k] = a, val
        st.bc, st.Pe = bc, 40.0
        st.u, st.v = u.data_ptr(), v.data_ptr()
        st.gxT = gx.data_ptr() if pointwise else None
        st.gyT = gy.data_ptr() if pointwise else None
        if residual:
            L.check(d.lib.sem_cd_residual(d.ctx, C.byref(st), x.data_ptr(), ys[0].data_ptr(), d.stream), "sem_cd_residual")
        else:
            L.check(d.lib.sem_cd_jvp(d.ctx, C.byref(st), x.data_ptr(), du.data_ptr() if pointwise else None,
                                     dv.data_ptr() if pointwise else None, ys[0].data_ptr(), d.stream), "sem_cd_jvp")

    both(lambda d, ys: cd(d, ys, True, False), 1)
    both(lambda d, ys: cd(d, ys, False, False), 1)
    both(lambda d, ys: cd(d, ys, False, True), 1)
    nf, ns = (loop.lib.sem_ctx_partitioned_applies(loop.ctx, k) for k in (1, 0))
    if expect_fused is not None:
        assert (nf > 0 and ns == 0) if expect_fused else (nf == 0), (nf, ns)
    return nf, ns


if __name__ == "__main__":
    sys.path.insert(0, os.getcwd())
    import sem_b200
    P, nex, ney = (int(t) for t in os.environ["SEM_LOOPBACK_CHILD"].split(","))
    check(sem_b200, P, nex, ney, expect_fused=False)
    print("LOOPBACK_OK")
