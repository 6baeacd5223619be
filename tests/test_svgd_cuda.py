"""K4 parity (GPU): SVGD particle loss/gradients and the pairwise direction through the C ABI vs the golden vectors of
the live reference and vs the fp64 oracle."""
import numpy as np
import pytest
import torch

from helpers import load_golden, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from brancher_b200 import _cuda
    _cuda.lib()
    assert torch.cuda.is_available()
    return _cuda


def dev(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device="cuda")


def test_svgd_direction_golden(cu):
    z = load_golden("svgd_small")["raw"]
    out, bw = cu.svgd_direction(dev(z["theta"]), dev(z["grad"]))
    assert abs(bw.item() - float(z["bandwidth"])) <= 1e-6 * float(z["bandwidth"])
    assert_close(out.cpu().numpy(), z["out"], "svgd out vs reference", rtol=1e-5, atol=1e-6, scale=np.abs(z["out"]).max())


def test_svgd_full_iteration_golden(cu):
    """compute_loss + backward + correct_gradient of the reference == K4a + K4b."""
    z = load_golden("svgd_softmax")["raw"]
    n, C, F = z["theta"].shape
    loss, G = cu.linear_particles_loss_grad(dev(z["X"]), dev(z["y"], torch.int32), cu.CATEGORICAL, dev(z["theta"]).reshape(n, -1),
                                            C, dev(z["prior_loc"]).reshape(-1), dev(z["prior_scale"]).reshape(-1))
    assert_close(loss.item(), z["loss"], "svgd loss", rtol=1e-5, atol=1e-6)
    assert_close(G.cpu().numpy().reshape(n, C, F), z["raw_grad"], "raw grad", rtol=1e-5, atol=1e-6, scale=np.abs(z["raw_grad"]).max())
    out, bw = cu.svgd_direction(dev(z["theta"]).reshape(n, -1), G)
    assert abs(bw.item() - float(z["bandwidth"])) <= 2e-6 * float(z["bandwidth"])
    assert_close(out.cpu().numpy().reshape(n, C, F), z["out"], "svgd direction", rtol=1e-5, atol=1e-6, scale=np.abs(z["out"]).max())


@pytest.mark.parametrize("n,d", [(2, 1), (3, 4), (64, 7), (100, 128), (257, 130), (1000, 33), (2048, 128)])
def test_svgd_direction_vs_oracle(cu, n, d):
    from oracle import elbo_oracle as O
    rng = np.random.RandomState(n * 31 + d)
    theta = rng.randn(n, d).astype("f4")
    grad = rng.randn(n, d).astype("f4")
    want, bw64 = O.svgd_direction(theta, grad)
    out, bw = cu.svgd_direction(dev(theta), dev(grad))
    # exact order statistics: the bandwidth only differs by fp32 rounding of the distances themselves
    assert abs(bw.item() - bw64) <= 2e-6 * bw64, (bw.item(), bw64)
    assert_close(out.cpu().numpy(), want, "svgd n=%d d=%d" % (n, d), scale=np.abs(want).max())


def test_svgd_duplicate_particles_and_ties(cu):
    """Collisions: identical particles (zero distances) and tied distances must follow np.median."""
    from oracle import elbo_oracle as O
    rng = np.random.RandomState(3)
    base = rng.randn(5, 3).astype("f4")
    theta = np.concatenate([base, base, base[:2]], 0)          # many exact duplicates
    grad = rng.randn(theta.shape[0], 3).astype("f4")
    want, bw64 = O.svgd_direction(theta, grad)
    out, bw = cu.svgd_direction(dev(theta), dev(grad))
    assert abs(bw.item() - bw64) <= 2e-6 * bw64
    assert_close(out.cpu().numpy(), want, "svgd duplicates", scale=np.abs(want).max())


def test_svgd_row_sharding_is_invariant(cu):
    """Particles sharded over ranks: every rank computes its rows from the all-gathered theta / grad."""
    rng = np.random.RandomState(11)
    n, d = 301, 20
    theta, grad = dev(rng.randn(n, d)), dev(rng.randn(n, d))
    full, bw = cu.svgd_direction(theta, grad)
    parts = []
    for r0, rows in [(0, 100), (100, 0), (100, 150), (250, 51)]:
        o, bw_r = cu.svgd_direction(theta, grad, row0=r0, rows=rows)
        assert bw_r.item() == bw.item()
        parts.append(o)
    got = torch.cat(parts)
    assert_close(got.cpu().numpy(), full.cpu().numpy(), "sharded rows", rtol=1e-6, atol=1e-6, scale=float(full.abs().max()))


@pytest.mark.parametrize("N,F,C,n,lik", [(50, 9, 1, 7, "binomial"), (300, 128, 1, 130, "binomial"), (77, 5, 4, 33, "categorical"),
                                         (0, 3, 2, 4, "categorical")])
def test_particles_loss_grad_vs_oracle(cu, N, F, C, n, lik):
    from oracle import elbo_oracle as O
    rng = np.random.RandomState(N + n)
    X = rng.randn(N, F).astype("f4")
    theta = (0.5 * rng.randn(n, C, F)).astype("f4")
    pl, ps = (0.1 * rng.randn(C, F)).astype("f4"), (0.5 + rng.rand(C, F)).astype("f4")
    if lik == "binomial":
        y = (rng.rand(N) < 0.5).astype("f4")
        yd, code = dev(y), cu.BERNOULLI
    else:
        y = rng.randint(0, C, size=N)
        yd, code = dev(y, torch.int32), cu.CATEGORICAL
    l64, g64 = O.particles_loss_grad(X, y, theta, (pl, ps), dtype=torch.float64, likelihood=lik)
    loss, G = cu.linear_particles_loss_grad(dev(X).reshape(N, F), yd, code, dev(theta).reshape(n, -1), C, dev(pl).reshape(-1),
                                            dev(ps).reshape(-1))
    assert_close(loss.item(), l64, "particles loss")
    assert_close(G.cpu().numpy().reshape(g64.shape), g64, "particles grad", scale=np.abs(g64).max())
    # no prior
    l64, g64 = O.particles_loss_grad(X, y, theta, None, dtype=torch.float64, likelihood=lik)
    loss, G = cu.linear_particles_loss_grad(dev(X).reshape(N, F), yd, code, dev(theta).reshape(n, -1), C)
    assert_close(loss.item(), l64, "particles loss (no prior)")
    assert_close(G.cpu().numpy().reshape(g64.shape), g64, "particles grad (no prior)", scale=max(np.abs(g64).max(), 1e-8))


@pytest.mark.parametrize("N,F,n,force", [(300, 128, 130, True), (1000, 20, 70, True), (5000, 128, 200, False), (4200, 64, 257, False)])
def test_particles_loss_grad_tcgen05_variant(cu, monkeypatch, N, F, n, force):
    """K4a through the K2 tcgen05 GEMM pair (particles as the weight vectors): ragged row/particle counts, both the
    automatic selection (large N, n) and the forced one at small shapes; same tolerance as the SIMT path."""
    from oracle import elbo_oracle as O
    if force:
        monkeypatch.setenv("BRN_LINEAR_VARIANT", "tcgen05")
    rng = np.random.RandomState(N + n)
    X = rng.randn(N, F).astype("f4")
    theta = (0.5 * rng.randn(n, 1, F)).astype("f4")
    pl, ps = (0.1 * rng.randn(1, F)).astype("f4"), (0.5 + rng.rand(1, F)).astype("f4")
    y = (rng.rand(N) < 0.5).astype("f4")
    l64, g64 = O.particles_loss_grad(X, y, theta, (pl, ps), dtype=torch.float64, likelihood="binomial")
    loss, G = cu.linear_particles_loss_grad(dev(X), dev(y), cu.BERNOULLI, dev(theta).reshape(n, -1), 1, dev(pl).reshape(-1),
                                            dev(ps).reshape(-1))
    assert cu.last_variant().startswith("tcgen05")
    assert_close(loss.item(), l64, "particles loss (tcgen05)")
    assert_close(G.cpu().numpy().reshape(g64.shape), g64, "particles grad (tcgen05)", scale=np.abs(g64).max())


@pytest.mark.parametrize("n,d,rows", [(512, 128, None), (1000, 33, None), (4096, 128, None), (2048, 100, (512, 700))])
def test_svgd_direction_tcgen05_variant(cu, monkeypatch, n, d, rows):
    """K4b on the tensor cores: D2 through the theta.theta^T GEMM (norms in the epilogue), the update as a GEMM whose A
    operand is exponentiated by the converter warps; exact median unchanged.  Same tolerance as the SIMT variant; a row
    shard (particles sharded over ranks) equals the corresponding rows of the full update."""
    from oracle import elbo_oracle as O
    monkeypatch.setenv("BRN_SVGD_VARIANT", "tcgen05")
    rng = np.random.RandomState(n + d)
    theta = rng.randn(n, d).astype("f4")
    grad = rng.randn(n, d).astype("f4")
    want, bw64 = O.svgd_direction(theta, grad)
    if rows is None:
        out, bw = cu.svgd_direction(dev(theta), dev(grad))
        sel = slice(None)
    else:
        out, bw = cu.svgd_direction(dev(theta), dev(grad), row0=rows[0], rows=rows[1])
        sel = slice(rows[0], rows[0] + rows[1])
    assert cu.last_variant() == "tcgen05"
    assert abs(bw.item() - bw64) <= 2e-6 * bw64, (bw.item(), bw64)
    assert_close(out.cpu().numpy(), want[sel], "svgd tcgen05 n=%d d=%d" % (n, d), scale=np.abs(want).max())


@pytest.mark.parametrize("n,d,shards", [(1024, 128, [(0, 1024)]), (2048, 64, [(0, 1024), (1024, 1024)]),
                                        (1536, 100, [(0, 256), (256, 768), (1024, 512)]), (1024, 32, [(0, 1024), (1024, 0)])])
def test_svgd_sharded_median_equals_replicated(cu, n, d, shards):
    """K4b sharded over ranks (brn_svgd_sharded_phase): every rank histograms only ITS rows of the distance matrix and the
    histograms are added between the select passes.  The ranks are played by several ShardedSvgd instances on one device
    (the additions torch.distributed performs in the product are done here by hand): the bandwidth must equal the replicated
    evaluation's BIT FOR BIT on every rank, and the concatenated row blocks must equal its update (and the oracle's)."""
    from oracle import elbo_oracle as O
    rng = np.random.RandomState(n + d)
    theta = rng.randn(n, d).astype("f4")
    theta[5] = theta[17]                                   # duplicate particles: zero distances, ties in the selection
    grad = rng.randn(n, d).astype("f4")
    th, gg = dev(theta), dev(grad)
    full, bw_full = cu.svgd_direction(th, gg)
    full, bw_full = full.clone(), bw_full.clone()
    ranks = []
    for row0, rows in shards:
        nbytes = cu.lib().brn_svgd_sharded_workspace_bytes(n, d, rows)
        ranks.append(cu.ShardedSvgd(th, gg, row0, rows, workspace=torch.empty(nbytes, dtype=torch.uint8, device=th.device)))
    for k in range(3):
        for r in ranks:
            r.phase(k)
        tot = sum(r.hist.clone() for r in ranks)
        for r in ranks:
            r.hist.copy_(tot)
    for r in ranks:
        r.phase(3)
    cnt = sum(r.cnt_le.clone() for r in ranks)
    nxt = torch.stack([r.next.clone() for r in ranks]).min(dim=0).values
    for r in ranks:
        r.cnt_le.copy_(cnt)
        r.next.copy_(nxt)
        r.phase(4)
    for r in ranks:
        assert r.bw.item() == bw_full.item(), (r.bw.item(), bw_full.item())
    out = torch.cat([r.out for r in ranks if r.rows > 0]).cpu().numpy()
    want, bw64 = O.svgd_direction(theta, grad)
    assert abs(bw_full.item() - bw64) <= 2e-6 * bw64
    assert_close(out, full.cpu().numpy(), "sharded vs replicated update", scale=np.abs(want).max())
    assert_close(out, want, "sharded update vs oracle", scale=np.abs(want).max())
    # the single-process driver (no process group: no collectives) is the one-rank case
    o1, b1 = cu.svgd_direction_sharded(th, gg, 0, n)
    assert b1.item() == bw_full.item()
    assert_close(o1.cpu().numpy(), want, "svgd_direction_sharded (one rank)", scale=np.abs(want).max())
