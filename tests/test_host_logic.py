"""CPU: host-side mirror -- model construction, shapes, graph lowering, error behaviour."""
import numpy as np
import pytest
import torch

import model_zoo as zoo


@pytest.fixture(scope="module")
def ns():
    from brancher_b200 import config
    config.set_device("cpu")
    yield zoo.namespace("brancher_b200")
    config.set_device("cuda:0" if torch.cuda.is_available() else "cpu")


def test_shape_convention(ns):
    from brancher_b200.utilities import coerce_to_dtype
    assert coerce_to_dtype(1.5).shape == (1, 1, 1, 1)
    assert coerce_to_dtype(np.zeros((3, 4))).shape == (1, 1, 3, 4)
    assert coerce_to_dtype(np.zeros((7,)), is_observed=True).shape == (1, 7, 1, 1)
    assert coerce_to_dtype(np.zeros((7, 5)), is_observed=True).shape == (1, 7, 5, 1)
    assert coerce_to_dtype(np.zeros((7, 5, 1)), is_observed=True).shape == (1, 7, 5, 1)
    assert coerce_to_dtype([1, 2]) == [1, 2]


def test_auto_named_roots_and_softplus_storage(ns):
    v = ns.NormalVariable(0.5, 2.0, "w", learnable=True)
    names = sorted(p.name for p in v.parents)
    assert names == ["w_loc", "w_scale"]
    rho = v.roots["scale"].value
    assert isinstance(rho, torch.nn.Parameter) and rho.shape == (1, 1, 1, 1)
    np.testing.assert_allclose(torch.nn.functional.softplus(rho).item(), 2.0, rtol=1e-6)
    np.testing.assert_allclose(rho.item(), np.log(np.exp(2.0) - 1), rtol=1e-6)


def test_bnn_lowers_to_k3_tied(ns):
    from brancher_b200 import lowering
    model, Q, d = zoo.bnn(ns, 1, B=12, P=20, H=7, C=4)
    plan = lowering.get_plan(model, model.posterior_model)
    assert plan.family.startswith("bnn")
    assert [s.name for s in plan.latents] == ["weights1", "b1", "weights2", "b2"]
    assert [s.shape for s in plan.latents] == [(7, 20), (7, 1), (4, 7), (4, 1)]
    assert all(s.tied for s in plan.latents)       # numeric hyper-parameters on both sides => name collision
    assert lowering.get_plan(model, model.posterior_model) is plan      # cached


@pytest.mark.parametrize("tied", [True, False])
def test_logreg_lowers_to_k2(ns, tied):
    from brancher_b200 import lowering
    model, Q, d = zoo.logreg(ns, 3, B=40, F=8, tied=tied)
    plan = lowering.get_plan(model, model.posterior_model)
    assert plan.family.startswith("linear") and plan.C == 1
    spec = plan.latents[0]
    assert spec.tied == tied
    if not tied:
        np.testing.assert_allclose(spec.prior_scale.cpu().numpy().reshape(-1), 0.5)
        np.testing.assert_allclose(spec.prior_loc.cpu().numpy().reshape(-1), 0.0)


def test_unsupported_graph_raises(ns):
    from brancher_b200 import lowering
    x = ns.RootVariable(np.random.rand(5, 3, 1), "x", is_observed=True)
    w = ns.NormalVariable(np.zeros((1, 3)), np.ones((1, 3)), "weights")
    k = ns.BinomialVariable(1, logits=ns.BF.sin(ns.BF.matmul(w, x)), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(np.ones((5, 1)))
    model.set_posterior_model(ns.ProbabilisticModel([ns.NormalVariable(np.zeros((1, 3)), np.ones((1, 3)), "weights",
                                                                       learnable=True)]))
    with pytest.raises(lowering.UnsupportedModelError):
        lowering.get_plan(model, model.posterior_model)


def test_no_cpu_fallback_for_elbo(ns):
    model, Q, d = zoo.logreg(ns, 3, B=40, F=8, tied=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ns.inference.ReverseKL().compute_loss(model, model.posterior_model, None, 4)


def test_eager_sampling_api(ns):
    model, Q, d = zoo.bnn(ns, 1, B=12, P=20, H=7, C=4)
    s = model._get_posterior_sample(3)
    by_name = {v.name: t for v, t in s.items() if torch.is_tensor(t)}
    assert by_name["weights1"].shape == (3, 1, 7, 20)
    assert by_name["k"].shape == (3, 12, 4, 1)
    lp = model.calculate_log_probability(model._get_sample(2))
    assert lp.shape == (2, 1)          # observed node summed over the data axis
    frame = model.get_sample(2)
    assert frame.shape[0] == 2 and "k" in frame.columns


def test_optimizer_collects_parameters(ns):
    from brancher_b200.optimizers import ProbabilisticOptimizer
    model, Q, d = zoo.bnn(ns, 1, B=12, P=20, H=7, C=4)
    opt = ProbabilisticOptimizer(model.posterior_model, "SGD", lr=0.1)
    assert len(list(opt.module.parameters())) == 8
    assert sum(p.numel() for p in opt.module.parameters()) == 2 * (7 * 20 + 7 + 4 * 7 + 4)


def test_shard_is_balanced_partition():
    from brancher_b200 import distributed as D
    for total, world in [(256, 8), (10, 3), (5, 8), (0, 2)]:
        parts = [D.shard(total, world, r) for r in range(world)]
        assert sum(c for _, c in parts) == total
        pos = 0
        for first, count in parts:
            assert first == pos
            pos += count
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_svgd_particles_lower_to_k4(ns):
    """The SVGD playground's particle ensemble is recognised by the graph matcher (no GPU needed)."""
    from brancher_b200 import lowering
    model, particles, d = zoo.svgd_softmax(ns, 8, B=30, F=5, C=3, n=6)
    plan = lowering.get_particle_plan(model, particles)
    assert plan.family == "particles (K4)" and plan.shape == (3, 5) and plan.C == 3
    assert plan.stacked().shape == (6, 15)
    np.testing.assert_allclose(plan.prior_scale.numpy(), 10.0, rtol=1e-6)
    bad = [ns.ProbabilisticModel([ns.RootVariable(np.zeros((2, 2)), name="weights", learnable=True)])]
    with pytest.raises(lowering.UnsupportedModelError):
        lowering.lower_particles(model, bad)


def test_wvgd_lowering_and_no_cpu_fallback():
    """WVGD ensembles lower to the K6 plan on the host (tied prior detected by root-name collision) and refuse to run on CPU."""
    import model_zoo as zoo
    from brancher_b200 import config, lowering
    config.set_device("cpu")
    ns = zoo.namespace("brancher_b200")
    model, particles, samplers, d = zoo.wvgd_softmax(ns, 21, 30, 4, 3, 3)
    m = ns.inference.WassersteinVariationalGradientDescent(variational_samplers=samplers, particles=particles)
    m.check_model_compatibility(model, particles, m.sampler_model)
    plan = lowering.get_wvgd_plan(model, particles, samplers)
    assert plan.family == "wvgd (K6)" and plan.tied
    assert [tuple(p.shape) for p in plan.parameters()] == [(1, 1, 3, 4)] * 6 + [(1, 1, 1, 1)] * 3
    with pytest.raises(RuntimeError):
        m.compute_loss(model, particles, m.sampler_model, 5)
    with pytest.raises(lowering.UnsupportedModelError):
        lowering.get_wvgd_plan(model, particles, samplers[:2])
    with pytest.raises(NotImplementedError):
        ns.inference.WassersteinVariationalGradientDescent(samplers, particles, cost_function=lambda a, b: 0)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints exactly ONE JSON line on stdout with the
    keys the measurement contract names -- runs on CPU, bounded sample."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "ar1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def _logreg_with(ns, prior_loc, prior_scale, q_scale):
    rng = np.random.RandomState(0)
    F, B = 5, 12
    x = ns.RootVariable(rng.randn(B, F, 1).astype("float32"), "x", is_observed=True)
    weights = ns.NormalVariable(prior_loc, prior_scale, "weights")
    k = ns.BinomialVariable(1, logits=ns.BF.matmul(weights, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe((rng.rand(B, 1) < 0.5).astype("float32"))
    model.set_posterior_model(ns.ProbabilisticModel([ns.NormalVariable(np.zeros((1, F)), q_scale, "weights", learnable=True)]))
    return model


def test_learnable_untied_prior_is_rejected(ns):
    """ReverseKL optimises the joint model's learnable parameters too (inference.py:132); the fused kernels return no
    gradient for a prior's own hyper-parameters, so such a model must not lower silently (it would train with them frozen)."""
    from brancher_b200 import lowering
    F = 5
    model = _logreg_with(ns, ns.RootVariable(np.zeros((1, F)), "prior_loc", learnable=True),
                         ns.RootVariable(0.5 * np.ones((1, F)), "prior_scale"), np.ones((1, F)))
    with pytest.raises(lowering.UnsupportedModelError, match="learnable hyper-parameters"):
        lowering.get_plan(model, model.posterior_model)


def test_scalar_q_scale_is_rejected_with_a_clear_message(ns):
    """a scalar scale is valid in the reference; the mean-field kernels need one scale per element (no out-of-bounds read)"""
    from brancher_b200 import lowering
    F = 5
    model = _logreg_with(ns, np.zeros((1, F)), 0.5 * np.ones((1, F)), 1.0)
    with pytest.raises(lowering.UnsupportedModelError, match="one\\s+scale per element|scale has 1 elements"):
        lowering._lower_dense(model, model.posterior_model)


def test_declared_prior_is_read_at_evaluation_time(ns):
    from brancher_b200 import lowering
    F = 5
    loc = ns.RootVariable(np.zeros((1, F)), "prior_loc")
    model = _logreg_with(ns, loc, ns.RootVariable(0.5 * np.ones((1, F)), "prior_scale"), np.ones((1, F)))
    spec = lowering.get_plan(model, model.posterior_model).latents[0]
    np.testing.assert_allclose(spec.prior_loc.numpy().reshape(-1), 0.0)
    loc._value = loc._value + 2.0                                    # re-assigned constant: the plan must see it
    np.testing.assert_allclose(spec.prior_loc.numpy().reshape(-1), 2.0)


def test_plan_cache_follows_observation_changes(ns):
    from brancher_b200 import lowering
    model, Q, d = zoo.logreg(ns, 3, B=40, F=8, tied=True)
    p1 = lowering.get_plan(model, model.posterior_model)
    assert lowering.get_plan(model, model.posterior_model) is p1
    sig = lowering._observation_signature(model)
    assert ("k", True) in sig


def test_vae_lowers_to_k5(ns):
    """examples/VAE_playground.py:65-80 written against the package API: BrancherFunction(nn.Module) encoder / decoder with
    dict outputs and Index links lower to the K5 family; the module structure is recovered by a verified numerical probe."""
    from brancher_b200 import lowering
    model, Q, d = zoo.vae(ns, 3, B=10, D=12, L=2, h_enc=(8, 6), h_dec=(6, 8))
    plan = lowering.get_plan(model, model.posterior_model)
    assert plan.family.startswith("vae")
    assert abs(plan.sd_offset - 0.1) < 1e-9
    assert [tuple(l.weight.shape) for l in plan.layers()] == [(8, 12), (6, 8), (2, 6), (2, 6), (6, 2), (8, 6), (12, 8)]
    assert plan.enc_mean is d["enc"].l_mean and plan.enc_sd is d["enc"].l_sd and plan.dec_out is d["dec"].l_out
    assert len(plan.parameters()) == 14


def test_vae_lowering_rejects_a_module_that_is_not_a_relu_mlp(ns):
    import torch.nn as nn
    from brancher_b200 import lowering

    class Odd(nn.Module):
        def __init__(self):
            super().__init__()
            self.a, self.m, self.s = nn.Linear(12, 8), nn.Linear(8, 2), nn.Linear(8, 2)

        def __call__(self, x):
            h = torch.tanh(self.a(x.squeeze(-1)))          # tanh, not ReLU
            return {"mean": self.m(h), "sd": nn.functional.softplus(self.s(h)) + 0.1}

    with pytest.raises(lowering.UnsupportedModelError, match="probe mismatch"):
        lowering._identify_relu_mlp(Odd(), 12, {"mean": "id", "sd": "softplus+c"}, "encoder")
