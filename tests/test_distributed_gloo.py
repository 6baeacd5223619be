"""World-size-2 CPU (gloo) coverage of the N>1 host logic: sample sharding + the single flat all-reduce of
[grads | loss_hi | loss_lo] (brancher_b200/distributed.py).  Each rank plays the role of one GPU: it evaluates
the oracle on ITS shard of the global MC samples with the global 1/S_total scaling (the contract of
include/brancher_cuda.h), and the all-reduced result must equal the single-process evaluation."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.RandomState(5)
    N, F, S = 40, 6, 7            # S odd: the shards are ragged (4 + 3)
    X = rng.randn(N, F).astype("f4")
    y = (rng.rand(N) < 0.5).astype("f4")
    params = {"weights": ((0.3 * rng.randn(1, F)).astype("f4"), (rng.randn(1, F) - 1).astype("f4"))}
    eps = {"weights": rng.randn(S, 1, F).astype("f4")}
    return X, y, params, eps, S


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brancher_b200 import distributed as D
    from oracle import elbo_oracle as O
    X, y, params, eps, S = _problem()
    assert D.world_size() == world and D.rank() == rank
    s0, n = D.shard(S)
    # partial of this rank: mean over its samples, rescaled to the global 1/S; prior/entropy is sample-independent
    # in expectation only through eps, so evaluate it per shard as well and weight by n/S.
    loss_r, g_r = O.logreg_elbo(X, y, params, {"weights": eps["weights"][s0:s0 + n]}, prior={"weights": (0.0, 0.5)},
                                dtype=torch.float64)
    w = n / S
    loss = torch.tensor([loss_r * w], dtype=torch.float64)
    grads = [torch.tensor(g_r["weights_loc"] * w, dtype=torch.float32), torch.tensor(g_r["weights_scale"] * w, dtype=torch.float32)]
    loss, grads = D.all_reduce_partials(loss, grads)
    if rank == 0:
        torch.save({"loss": loss, "grads": grads, "shards": [D.shard(S, world, r) for r in range(world)]}, out)
    dist.destroy_process_group()


def test_two_rank_sample_sharding_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import elbo_oracle as O
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    X, y, params, eps, S = _problem()
    assert got["shards"] == [(0, 4), (4, 3)]
    want_loss, want_g = O.logreg_elbo(X, y, params, eps, prior={"weights": (0.0, 0.5)}, dtype=torch.float64)
    assert abs(got["loss"].item() - want_loss) <= 1e-6 * abs(want_loss)      # hi/lo pair keeps ~fp64 through fp32 wire
    for g, k in zip(got["grads"], ["weights_loc", "weights_scale"]):
        np.testing.assert_allclose(g.numpy().reshape(-1), want_g[k].reshape(-1), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# C4: particles sharded over ranks -- each rank evaluates the loss gradients of ITS particles, all ranks all-gather theta
# and G, and each computes its rows of the pairwise SVGD update from the gathered copies (bench.py's svgd
# workload does exactly this with NCCL and the K4a / K4b kernels).
# ---------------------------------------------------------------------------------------------------
def _svgd_problem():
    rng = np.random.RandomState(9)
    N, F, C, n = 30, 5, 1, 6            # 3 particles per rank
    X = rng.randn(N, F).astype("f4")
    y = (rng.rand(N) < 0.5).astype("f4")
    theta = (0.7 * rng.randn(n, C, F)).astype("f4")
    prior = (np.zeros((C, F), "f4"), np.ones((C, F), "f4"))
    return X, y, theta, prior


def _svgd_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from brancher_b200 import distributed as D
    from oracle import elbo_oracle as O
    X, y, theta, prior = _svgd_problem()
    n, d = theta.shape[0], theta.shape[1] * theta.shape[2]
    p0, cnt = D.shard(n)
    _, G_local = O.particles_loss_grad(X, y, theta[p0:p0 + cnt], prior, dtype=torch.float64, likelihood="binomial")
    th_local = torch.tensor(theta[p0:p0 + cnt].reshape(cnt, d), dtype=torch.float64)
    g_local = torch.tensor(G_local.reshape(cnt, d), dtype=torch.float64)
    th_all, g_all = torch.empty((n, d), dtype=torch.float64), torch.empty((n, d), dtype=torch.float64)
    dist.all_gather_into_tensor(th_all, th_local)
    dist.all_gather_into_tensor(g_all, g_local)
    full, bw = O.svgd_direction(th_all.numpy(), g_all.numpy())       # every rank: bandwidth from the gathered particles
    mine = torch.tensor(full[p0:p0 + cnt])                            # ... and its own rows of the update
    rows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine)
    # the row-sharded median selection (brn_svgd_sharded_phase): this rank histograms only its rows, the histograms are added
    # by all-reduce -- every rank must end with np.median of ALL pairwise distances
    th32 = th_all.numpy().astype("f4")
    d2 = ((th32[:, None] - th32[None]) ** 2).sum(-1).astype("f4")

    def red(op):
        def f(a):
            t = torch.tensor(a)
            dist.all_reduce(t, op=op)
            return t.numpy()
        return f
    med = O.svgd_median_radix_sharded(d2, [(p0, cnt)], red(dist.ReduceOp.SUM), red(dist.ReduceOp.MIN))
    meds = [None] * world
    dist.all_gather_object(meds, float(med))
    if rank == 0:
        torch.save({"update": torch.cat(rows), "bw": bw, "medians": meds,
                    "median_want": float(np.median(np.sqrt(d2[np.triu_indices(n, 1)])))}, out)
    dist.destroy_process_group()


def test_two_rank_particle_sharding_matches_single_process(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import elbo_oracle as O
    out = str(tmp_path / "svgd.pt")
    mp.spawn(_svgd_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    X, y, theta, prior = _svgd_problem()
    _, G = O.particles_loss_grad(X, y, theta, prior, dtype=torch.float64, likelihood="binomial")
    n = theta.shape[0]
    want, bw = O.svgd_direction(theta.reshape(n, -1), G.reshape(n, -1))
    assert abs(got["bw"] - bw) <= 1e-12 * abs(bw)
    np.testing.assert_allclose(got["update"].numpy(), want, rtol=1e-10, atol=1e-12)
    assert got["medians"] == [got["median_want"]] * 2        # sharded radix selection == np.median, identical on both ranks
