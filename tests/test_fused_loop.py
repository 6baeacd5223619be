"""GPU: the CUDA-graph-captured training iteration (fused evaluation + brn_opt_step, brancher_b200/inference._fused_loop)
against the step-by-step loop (per-iteration host check + torch.optim): same Philox noise sequence, so the loss curves and
the trained parameters must agree to fp32 rounding (replaces brancher/inference.py:95-108, optimizers.py:69-73)."""
import numpy as np
import pytest
import torch

import model_zoo as zoo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ns():
    assert torch.cuda.is_available()
    from brancher_b200 import config
    config.set_device("cuda:0")
    return zoo.namespace("brancher_b200")


def learnable_values(model):
    return {v.name: v.link.parameter.detach().cpu().numpy().copy() for v in model.posterior_model.flatten()
            if getattr(v, "learnable", False) and hasattr(v.link, "parameter")}


BUILDERS = {
    "ar1": lambda ns: zoo.ar1(ns, 0, 20)[0],
    "logreg": lambda ns: zoo.logreg(ns, 3, B=300, F=16, tied=False)[0],
    "softmax": lambda ns: zoo.softmax_reg(ns, 5, B=64, F=6, C=3)[0],
    "bnn": lambda ns: zoo.bnn(ns, 1, B=130, P=40, H=100, C=10)[0],
}


@pytest.mark.parametrize("model_name,S,opt,kw,iters", [
    ("ar1", 300, "SGD", dict(lr=1e-3), 60), ("ar1", 50, "Adam", dict(lr=0.01), 40),
    ("logreg", 32, "Adam", dict(lr=0.05), 30), ("logreg", 32, "SGD", dict(lr=1e-4, momentum=0.9), 30),
    ("softmax", 16, "Adam", dict(lr=0.02, betas=(0.8, 0.99), eps=1e-6), 25), ("bnn", 6, "Adam", dict(lr=1e-3), 12)])
def test_graph_loop_equals_stepwise_loop(ns, model_name, S, opt, kw, iters):
    from brancher_b200 import config, inference
    out = {}
    for mode in ("graph", "eager"):
        config.set_seed(123)
        model = BUILDERS[model_name](ns)
        inference.fused_loop_enabled = mode == "graph"
        try:
            inference.perform_inference(model, number_iterations=iters, number_samples=S, optimizer=opt,
                                        inference_method=inference.ReverseKL(), **kw)
        finally:
            inference.fused_loop_enabled = True
        assert inference.last_loop == mode
        out[mode] = (np.asarray(model.diagnostics["loss curve"], dtype=np.float64).reshape(-1), learnable_values(model))
    cg, ce = out["graph"][0], out["eager"][0]
    assert cg.shape == ce.shape == (iters,)
    assert np.isfinite(cg).all()
    np.testing.assert_allclose(cg, ce, rtol=2e-5, atol=1e-5 * np.abs(ce).max())
    for k, v in out["eager"][1].items():
        sc = max(np.abs(v).max(), 1e-6)
        np.testing.assert_allclose(out["graph"][1][k], v, rtol=1e-4, atol=1e-5 * sc, err_msg=k)


def test_graph_loop_skips_non_finite_iterations(ns):
    """a non-finite loss leaves the parameters untouched and is counted (inference.py:98,106-107)"""
    from brancher_b200 import _cuda as cu
    dev = torch.device("cuda:0")
    p = torch.ones(5, device=dev)
    g = torch.full((5,), 2.0, device=dev)
    fo = cu.FusedOptimizer([p], [g], cu.SGD, lr=0.5, curve_len=3)
    off = torch.zeros(1, dtype=torch.int64, device=dev)
    for val in (1.0, float("nan"), float("inf")):
        fo.step(torch.tensor([val], dtype=torch.float64, device=dev), off)
    assert fo.counters.tolist() == [1, 3, 2] and off.item() == 3
    assert torch.allclose(p, torch.zeros(5, device=dev))
    c = fo.curve.cpu().numpy()
    assert c[0] == 1.0 and np.isnan(c[1]) and np.isinf(c[2])


def test_fused_adam_matches_torch_adam(ns):
    from brancher_b200 import _cuda as cu
    dev = torch.device("cuda:0")
    g0 = torch.Generator(device="cuda").manual_seed(0)
    ps = [torch.randn(n, device=dev, generator=g0) for n in (1, 7, 1000, 33)]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    grads = [torch.zeros_like(p) for p in ps]
    fo = cu.FusedOptimizer(ps, grads, cu.ADAM, lr=0.01, weight_decay=0.1)
    opt = torch.optim.Adam(ref, lr=0.01, weight_decay=0.1)
    one = torch.ones(1, dtype=torch.float64, device=dev)
    for it in range(20):
        for g, r in zip(grads, ref):
            g.copy_(torch.randn(g.shape, device=dev, generator=g0))
            r.grad = g.clone()
        fo.step(one)
        opt.step()
    for p, r in zip(ps, ref):
        assert torch.allclose(p, r.detach(), rtol=2e-6, atol=2e-7), (p - r).abs().max()
