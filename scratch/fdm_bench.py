"""Time the fast-diagonalisation apply (hand-written DMMA GEMM, folded transforms) at large sizes.
usage: python scratch/fdm_bench.py [ne ...]   (P = 8; nodes per direction = 8 ne + 1)"""
import json, sys, time
sys.path.insert(0, '.')
import torch
import sem_b200
out = {}
for ne in [int(a) for a in sys.argv[1:]] or [64, 256, 512]:
    P = 8
    d = sem_b200.SemDevice(P, ne, ne, 1.0 / ne, 1.0 / ne)
    t0 = time.time()
    d.setup_fdm([1, 1, 1, 1])
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    r = d.zeros(2)
    r[:, :, :d.NY] = torch.randn((2, d.NX, d.NY), device=d.tdev, dtype=torch.float64)
    z = d.zeros(2)
    res = {}
    for nf in (1, 2):
        for _ in range(2):
            d.fdm_apply(0, r, z, nf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            d.fdm_apply(0, r, z, nf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        n = d.NX - 2
        flops_full = 4 * 2.0 * n ** 3 * nf           # the four unfolded products
        res[f"nf{nf}_ms"] = ms
        res[f"nf{nf}_equiv_tflops_unfolded"] = flops_full / ms * 1e-9
        res[f"nf{nf}_actual_tflops"] = flops_full / 2 / ms * 1e-9
    # check: K z = r on interior
    Kz = d.zeros()
    d.apply_stiffness(z[0], Kz)
    err = float((Kz[1:-1, 1:d.NY - 1] - r[0, 1:-1, 1:d.NY - 1]).norm() / r[0, 1:-1, 1:d.NY - 1].norm())
    res["residual"] = err
    res["setup_s"] = t_setup
    out[f"{d.NX}x{d.NY}"] = res
    print(d.NX, res, flush=True)
    del d, r, z, Kz
    torch.cuda.empty_cache()
print(json.dumps(out))
