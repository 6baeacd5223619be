"""torchrun script (N >= 2 GPUs): brn_allreduce_oneshot against torch.distributed.all_reduce on the same buffers, eagerly and
replayed from a CUDA graph; bit-identical results on every rank; timing of both.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu/check_oneshot_allreduce.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from brancher_b200 import distributed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for n in (4, 260, 159024, 668952):
    g = torch.Generator(device="cuda").manual_seed(1000 * rank + n)
    for it in range(5):
        x = torch.randn(n, device=dev, generator=g)
        ref = x.double()
        dist.all_reduce(ref)
        y = distributed.all_reduce_flat(x.clone())
        assert distributed._oneshot[(local, n)], "one-shot path not active"
        err = (y.double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        gathered = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(gathered, y)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if err > 1e-6 or not same:
            ok = False
            print("rank %d n=%d it=%d: rel err %.2e identical=%s" % (rank, n, it, err, same), flush=True)
# graph replay + timing at the C3 gradient size
n = 159024
x = torch.randn(n, device=dev)
buf = x.clone()
distributed.all_reduce_flat(buf)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    buf.copy_(x)
    distributed.all_reduce_flat(buf)
ref = x.double()
dist.all_reduce(ref)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
ok = ok and (buf.double() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


t_graph = timeit(g.replay)
t_eager = timeit(lambda: distributed.all_reduce_flat(buf))
t_nccl = timeit(lambda: dist.all_reduce(buf))
timeouts = distributed._oneshot[(local, n)].timeouts()
if rank == 0:
    print("RESULT ok=%s world=%d  636 KB all-reduce: one-shot in graph %.1f us, one-shot eager %.1f us, NCCL %.1f us, time-outs %d" % (
        ok and timeouts == 0, world, t_graph, t_eager, t_nccl, timeouts), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
