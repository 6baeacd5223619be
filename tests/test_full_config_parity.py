"""GPU parity at every BASELINE configuration's OWN size: the CUDA path through the C ABI against the
fp64 evaluation of the oracle on the same inputs and the same noise.

Each test asserts the repo's parity rule (tests/helpers.check_against_oracle: rtol 1e-5, atol 1e-6 x the tensor's own
scale) and additionally REPORTS how many elements fall outside the literal `rtol 1e-5 / atol 1e-6`
box -- for the CUDA result and for the oracle's own fp32 evaluation (what the reference's arithmetic itself delivers),
both measured against fp64.  The report is appended to gpurun_out/parity_full_config.jsonl (profiles/ keeps a copy).
"""
import json
import os

import numpy as np
import pytest
import torch

from helpers import check_against_oracle, assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cu():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from brancher_b200 import _cuda
    _cuda.lib()
    return _cuda


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).to(DEV).contiguous()


def literal_report(config, got, g64, g32=None):
    """fraction of elements outside |x - x64| <= 1e-6 + 1e-5 |x64| (no tensor scale), per tensor"""
    rec = {"config": config, "tensors": {}}
    for k in g64:
        w = np.asarray(g64[k], dtype=np.float64)
        a = np.asarray(got[k], dtype=np.float64).reshape(w.shape)
        box = 1e-6 + 1e-5 * np.abs(w)
        sc = max(np.abs(w).max(), 1e-300)
        t = {"n": int(w.size), "scale": float(sc), "cuda_outside_literal": int((np.abs(a - w) > box).sum()),
             "cuda_max_err_over_scale": float(np.abs(a - w).max() / sc)}
        if g32 is not None:
            b = np.asarray(g32[k], dtype=np.float64).reshape(w.shape)
            t["fp32_oracle_outside_literal"] = int((np.abs(b - w) > box).sum())
            t["fp32_oracle_max_err_over_scale"] = float(np.abs(b - w).max() / sc)
        rec["tensors"][k] = t
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_full_config.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    print(json.dumps(rec))
    return rec


# ------------------------------------------------------------------------------------------------ C3
@pytest.mark.parametrize("noise", ["philox", "injected"])
def test_c3_bnn_full_config(cu, noise):
    """C3: 784-100-10 tanh BNN, B = 1024, S = 256, q init (mu = 0 -> small random, sigma = 0.01), tied prior as in the playground.
    Philox noise is re-materialised with brn_philox_normal_fill for the oracle; the injected run feeds the same noise back."""
    from oracle import elbo_oracle as O
    import test_cuda_kernels as T
    B, P, H, C, S = 1024, 784, 100, 10, 256
    X, y, params, _, shapes = T.random_bnn(2026, B, P, H, C, 1, sigma=0.01, mu_scale=0.3)
    r = cu.sample_range(S, seed=7, offset=11)
    eps = {n: cu.philox_normal(int(np.prod(shapes[n])), i, r, DEV).cpu().numpy().reshape((S,) + shapes[n])
           for i, n in enumerate(T.BNN_NAMES)}
    o64 = O.bnn_elbo(X, y, params, eps, None, dtype=torch.float64, sample_chunk=8)
    o32 = O.bnn_elbo(X, y, params, eps, None, sample_chunk=8)
    if noise == "philox":
        loss, grads, _ = T.run_bnn(cu, X, y, params, None, r=r)
    else:
        loss, grads, _ = T.run_bnn(cu, X, y, params, eps)
    assert cu.last_variant() == "tcgen05"
    literal_report("C3 bnn B=1024 S=256 (%s)" % noise, dict(grads, loss=np.array(loss)), dict(o64[1], loss=np.array(o64[0])),
                   dict(o32[1], loss=np.array(o32[0])))
    check_against_oracle(loss, grads, o32, o64, "C3 full config (%s)" % noise)


# ------------------------------------------------------------------------------------------------ C2
def test_c2_logreg_full_config(cu):
    """C2: N = 10^6 rows x 128 features, S = 1024, prior N(0, 0.5), q init (0, 1): the chunked tcgen05 pair (54 row chunks)."""
    from oracle import elbo_oracle as O
    N, F, S = 1_000_000, 128, 1024
    g = torch.Generator().manual_seed(0)
    X = torch.randn(N, F, generator=g)
    wstar = torch.randn(F, generator=g) / F ** 0.5
    y = (torch.rand(N, generator=g) < torch.sigmoid(X @ wstar)).float()
    params = {"weights": (np.zeros((1, F), "f4"), O.softplus_inverse(np.ones((1, F))).astype("f4"))}
    prior = {"weights": (0.0, 0.5)}
    r = cu.sample_range(S, seed=3, offset=5)
    eps = {"weights": cu.philox_normal(F, 0, r, DEV).cpu().numpy().reshape(S, 1, F)}
    o64 = O.logreg_elbo_streamed(X.numpy(), y.numpy(), params, eps, prior)
    o32 = O.logreg_elbo_streamed(X.numpy(), y.numpy(), params, eps, prior, dtype=torch.float32)
    w = cu.MeanFieldVar(dev(params["weights"][0]), dev(params["weights"][1]), var_id=0,
                        prior_loc=dev(np.zeros((1, F), "f4")), prior_scale=dev(np.full((1, F), 0.5, "f4")))
    loss = cu.linear_elbo_fwd_bwd(X.to(DEV), y.to(DEV), cu.BERNOULLI, w, 1, r).item()
    assert cu.last_variant() == "tcgen05"
    grads = {"weights_loc": w.dmu.cpu().numpy().reshape(1, F), "weights_scale": w.drho.cpu().numpy().reshape(1, F)}
    literal_report("C2 logreg N=1e6 S=1024", dict(grads, loss=np.array(loss)), dict(o64[1], loss=np.array(o64[0])),
                   dict(o32[1], loss=np.array(o32[0])))
    check_against_oracle(loss, grads, o32, o64, "C2 full config")


# ------------------------------------------------------------------------------------------------ C4
def test_c4_svgd_full_config(cu):
    """C4: n = 4096 particles, d = 128, 65536 rows: particle gradients (K4a), exact median bandwidth and direction (K4b)."""
    from oracle import elbo_oracle as O
    n, F, B = 4096, 128, 65536
    g = torch.Generator().manual_seed(1)
    X = torch.randn(B, F, generator=g)
    wstar = torch.randn(F, generator=g) / F ** 0.5
    y = (torch.rand(B, generator=g) < torch.sigmoid(X @ wstar)).float()
    theta = torch.randn(n, 1, F, generator=g)
    pl, ps = np.zeros((1, F), "f4"), np.full((1, F), 0.5, "f4")
    l64, G64 = O.particles_loss_grad_streamed(X.numpy(), y.numpy(), theta.numpy(), (pl, ps))
    loss, G = cu.linear_particles_loss_grad(X.to(DEV), y.to(DEV), cu.BERNOULLI, theta.reshape(n, F).to(DEV).contiguous(), 1,
                                            dev(pl).reshape(-1), dev(ps).reshape(-1))
    assert_close(loss.item(), l64, "C4 particle loss")
    Gn = G.cpu().numpy().reshape(G64.shape)
    assert_close(Gn, G64, "C4 particle gradients", scale=np.abs(G64).max())
    # the pairwise stage on the fp32 gradients the device produced (so its own error is measured, not K4a's)
    want, bw64 = O.svgd_direction(theta.reshape(n, F).numpy(), Gn.reshape(n, F))
    out, bw = cu.svgd_direction(theta.reshape(n, F).to(DEV).contiguous(), G.reshape(n, F))
    assert abs(bw.item() - bw64) <= 2e-6 * bw64, (bw.item(), bw64)
    literal_report("C4 svgd n=4096 d=128 B=65536", {"particle_grad": Gn, "direction": out.cpu().numpy()},
                   {"particle_grad": G64, "direction": want})
    assert_close(out.cpu().numpy(), want, "C4 svgd direction", scale=np.abs(want).max())


# ------------------------------------------------------------------------------------------------ C5
def test_c5_vae_full_config(cu):
    """C5: VAE_playground widths (784-256-512-(2+2) / 2-512-256-784), B = 4096, S = 16, Philox noise re-materialised."""
    from oracle import elbo_oracle as O
    import test_vae_cuda as V
    B, D, L, S = 4096, 784, 2, 16
    X, enc, dec, _ = V.random_vae(31, B, D, L, (256, 512), (512, 256), 1)
    r = cu.sample_range(S, seed=13, offset=2)
    net = V.make_net(cu, enc, dec)
    loss = cu.vae_elbo_fwd_bwd(dev(X), net, r, var_id=5).item()
    eps = cu.philox_normal(B * L, 5, r, DEV).cpu().numpy().reshape(S, B, L)
    o64 = O.vae_elbo(X, enc, dec, eps, dtype=torch.float64, row_chunk=512)
    o32 = O.vae_elbo(X, enc, dec, eps, row_chunk=512)
    grads = V.grads_of(net)
    literal_report("C5 vae B=4096 S=16", dict(grads, loss=np.array(loss)), dict(o64[1], loss=np.array(o64[0])),
                   dict(o32[1], loss=np.array(o32[0])))
    check_against_oracle(loss, grads, o32, o64, "C5 full config")
