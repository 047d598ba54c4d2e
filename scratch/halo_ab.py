"""Device time of the partitioned CD Jacobian apply on 2 ranks for slabs of `nex_total / 2` element columns each (1024 rows, P = 8):
slab-for-slab the work of config 5 on 2 * (1024 / (nex_total / 2)) GPUs.  Run once as is (fused halo exchange) and once with
SEM_B200_NO_FUSED_HALO=1 (round-1 three-launch sequence).  usage: torchrun --nproc-per-node 2 scratch/halo_ab.py [nex_total ...]"""
import ctypes as C, os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import sem_b200
from sem_b200 import _lib as L
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tag = "unfused" if os.environ.get("SEM_B200_NO_FUSED_HALO") else "fused"
for nex in [int(a) for a in sys.argv[1:]] or [256, 512, 1024]:
    ney = 1024
    cd = sem_b200.ConvectionDiffusionSolver(nex / ney, 1.0, 40.0, 8, nex, ney, T_W=0.5, T_E=-0.5, device=local, partition=(rank, world))
    d = cd._dev
    gen = torch.Generator(device=d.tdev).manual_seed(rank)
    nset = max(1, int(400e6 // (4 * 8 * d.vec_len)) + 1)       # rotate over more than L2
    sets = []
    for _ in range(nset):
        t = [d.zeros() for _ in range(4)]
        for x in t[:3]:
            x[:, :d.NY] = torch.randn((d.NX, d.NY), generator=gen, device=d.tdev, dtype=torch.float64)
        sets.append(t)
    cd._have_sys = True
    calls = []
    for x, u, v, y in sets:
        cd._u, cd._v = u, v
        calls.append((cd._state(with_jac=False), x, y))
    k = [0]
    def step():
        st, x, y = calls[k[0] % nset]; k[0] += 1
        L.check(d.lib.sem_cd_jvp(d.ctx, C.byref(st), x.data_ptr(), None, None, y.data_ptr(), d.stream), "jvp")
    res = []
    for rep in range(5):
        for _ in range(5): step()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): step()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 50], device=d.tdev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.append(float(t))
    if rank == 0:
        res.sort()
        ideal = 0.3672 * (nex / 2) / 1024 * 1e3
        print(f"{tag:8s} nex/rank {nex // 2:4d}: median {res[2] * 1e3:7.1f} us  min {res[0] * 1e3:7.1f} us   (bandwidth time {ideal:6.1f} us)", flush=True)
    del cd, d, sets, calls
    torch.cuda.empty_cache()
dist.destroy_process_group()
