"""torchrun script (N >= 2 GPUs): K4b with the median selection sharded over the ranks (distributed.svgd_direction_sharded)
against the replicated evaluation of the same ensemble on every rank: identical bandwidth bit for bit, the rank's rows of the
update within the kernel tolerance; timing of both at the C4 shape (n = 4096 particles per GPU, d = 128).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tests/multi_gpu/check_sharded_svgd.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from brancher_b200 import distributed, _cuda as cu

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
report = []
for n_local, d in ((512, 32), (1024, 100), (4096, 128)):
    g = torch.Generator(device="cuda").manual_seed(7 + n_local)          # same ensemble on every rank
    theta = torch.randn(world * n_local, d, device=dev, generator=g)
    grad = torch.randn(world * n_local, d, device=dev, generator=g)
    mine = slice(rank * n_local, (rank + 1) * n_local)
    full, bw_full = cu.svgd_direction(theta, grad, row0=rank * n_local, rows=n_local)      # replicated selection
    full, bw_full = full.clone(), bw_full.clone()
    out, bw = distributed.svgd_direction_sharded(theta[mine].contiguous(), grad[mine].contiguous())
    same_bw = bw.item() == bw_full.item()
    err = (out - full).abs().max().item() / full.abs().max().item()
    bws = [torch.empty_like(bw) for _ in range(world)]
    dist.all_gather(bws, bw)
    same_all = all(torch.equal(bws[0], b) for b in bws)
    ok = ok and same_bw and same_all and err <= 1e-5

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    t_rep = timed(lambda: cu.svgd_direction(theta, grad, row0=rank * n_local, rows=n_local))
    t_sh = timed(lambda: distributed.svgd_direction_sharded(theta[mine].contiguous(), grad[mine].contiguous()))
    report.append({"n_local": n_local, "d": d, "world": world, "same_bandwidth": same_bw and same_all, "max_err_over_scale": err,
                   "ms_replicated": t_rep, "ms_sharded_incl_allgather": t_sh})
if rank == 0:
    import json
    print("RESULT " + json.dumps({"ok": bool(ok), "cases": report}))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
