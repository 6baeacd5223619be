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


# ---------------------------------------------------------------------------------------------------
# one-shot all-reduce over NVLink peer memory (brn_allreduce_oneshot) for flat CUDA gradient buffers
# ---------------------------------------------------------------------------------------------------
_oneshot = {}
oneshot_enabled = True       # False: always use the NCCL collective (A/B measurements)


class OneShotAllReduce:
    """Symmetric buffers (torch.distributed._symmetric_memory provides the allocation and the peer mappings only) + the
    hand-written kernels of csrc/allreduce.cu: every rank reads all peers' partial gradients straight over NVLink and sums
    them in rank order.  ~10 us for the 636 KB of the C3 gradient, where the NCCL ring costs ~66 us at 8 ranks."""

    def __init__(self, numel, device):
        import torch.distributed._symmetric_memory as symm
        from brancher_b200 import _cuda as cu
        cu.lib()
        self.world, self.rank, self.numel = world_size(), rank(), int(numel)
        pad = (self.numel + 7) // 8 * 8
        total = 2 * pad + 2 * self.world + 8
        self.buf = symm.empty(total, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        bufs = [ptrs[r] + par * pad * 4 for par in (0, 1) for r in range(self.world)]
        flags = [ptrs[r] + 2 * pad * 4 for r in range(self.world)]
        self.bufs_dev = torch.tensor(bufs, dtype=torch.int64, device=device)
        self.flags_dev = torch.tensor(flags, dtype=torch.int64, device=device)
        self.state = torch.zeros(3, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier()                      # every rank's flags are zero before anybody raises one

    def __call__(self, flat, loss=None):
        from brancher_b200 import _cuda as cu
        cu._check(cu.lib().brn_allreduce_oneshot(flat.data_ptr(), flat.data_ptr(), self.numel, self.bufs_dev.data_ptr(),
                                                 self.flags_dev.data_ptr(), self.rank, self.world, self.state.data_ptr(),
                                                 None if loss is None else loss.data_ptr(), cu._stream(flat.device)),
                  "brn_allreduce_oneshot")
        return flat

    def timeouts(self):
        return int(self.state.view(torch.int32)[3].item())


def all_reduce_flat(flat, loss=None):
    """Sum a flat fp32 buffer over ranks in place: the one-shot peer-memory kernel on CUDA (NCCL backend), else the
    backend's own all_reduce (gloo in the CPU tests).  loss (fp64 [1], optional) is summed too, in place: it travels in
    the buffer's last quad (which the caller keeps spare, see _cuda.flat_grad_views) as a (hi, lo) fp32 pair."""
    if world_size() == 1:
        return flat
    if loss is not None and (flat.numel() % 4 or loss.dtype != torch.float64):
        raise ValueError("all_reduce_flat: a loss needs an fp64 scalar and a buffer whose length is a multiple of 4")
    if oneshot_enabled and flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous() and flat.data_ptr() % 16 == 0 \
            and dist.get_backend() == "nccl":
        key = (flat.device.index, flat.numel())
        ar = _oneshot.get(key)
        if ar is None:
            try:
                ar = OneShotAllReduce(flat.numel(), flat.device)
            except Exception as exc:          # symmetric memory unavailable on this system: remember and use NCCL
                import warnings
                warnings.warn("one-shot all-reduce unavailable (%s); using the NCCL collective" % exc)
                ar = False
            _oneshot[key] = ar
        if ar:
            return ar(flat, loss)
    if loss is not None:
        hi = loss.to(torch.float32)
        flat[-4:-2] = torch.cat([hi.reshape(1), (loss - hi.to(torch.float64)).to(torch.float32).reshape(1)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if loss is not None:
        loss.copy_((flat[-4].to(torch.float64) + flat[-3].to(torch.float64)).reshape(loss.shape))
    return flat


def svgd_direction_sharded(theta_local, grad_local, group=None):
    """K4b for particles sharded over the ranks (equal shards): all-gather of (theta, grad), then the exact-median selection
    sharded by rows -- each rank histograms its own rows of the pairwise distances and the histograms are added
    (_cuda.svgd_direction_sharded).  Returns (update of the LOCAL particles [n_local, d], bandwidth [1], identical on all ranks).
    Replaces SteinVariationalGradientDescent.correct_gradient (brancher/inference.py:301-324) for an ensemble spread over GPUs."""
    from brancher_b200 import _cuda as cu
    w, r = world_size(), rank()
    n_local, d = theta_local.shape
    if w == 1:
        return cu.svgd_direction(theta_local, grad_local)
    theta_all = torch.empty((w * n_local, d), dtype=theta_local.dtype, device=theta_local.device)
    grad_all = torch.empty_like(theta_all)
    dist.all_gather_into_tensor(theta_all, theta_local.contiguous(), group=group)
    dist.all_gather_into_tensor(grad_all, grad_local.contiguous(), group=group)
    return cu.svgd_direction_sharded(theta_all, grad_all, r * n_local, n_local, group=group)
