"""One process per GPU: sample / row sharding and the small gradient all-reduce.

torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo in CPU tests).  Every rank evaluates
its shard of the MC samples with GLOBAL sample indices (so the noise, hence the result, does not depend
on the number of ranks), produces partial loss / gradients already scaled by 1/S_total, and one
all-reduce(sum) of a single flat fp32 buffer [grads..., loss_hi, loss_lo] finishes the evaluation.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(total, world=None, r=None):
    """(first, count) of a contiguous balanced split of `total` items; the first `total % world` ranks
    get one extra item."""
    world = world_size() if world is None else world
    r = rank() if r is None else r
    base, rem = divmod(int(total), world)
    count = base + (1 if r < rem else 0)
    first = r * base + min(r, rem)
    return first, count


def all_reduce_partials(loss, grads):
    """Sum partial (loss fp64 [1], grads list of fp32 tensors) over ranks, in place.  The fp64 loss
    travels as a hi/lo fp32 pair inside the same flat buffer as the gradients: one collective."""
    if world_size() == 1:
        return loss, grads
    hi = loss.to(torch.float32)
    lo = (loss - hi.to(torch.float64)).to(torch.float32)
    flat = torch.cat([g.reshape(-1) for g in grads] + [hi.reshape(1), lo.reshape(1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].reshape(g.shape))
        off += n
    loss = flat[off].to(torch.float64).reshape(1) + flat[off + 1].to(torch.float64).reshape(1)
    return loss, grads
