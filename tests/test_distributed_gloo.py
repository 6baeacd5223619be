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
