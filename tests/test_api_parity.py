"""GPU: the SAME model-construction code (tests/model_zoo.py) that produced the golden vectors with the
reference package is run against brancher_b200; loss and every `.grad` must match the reference's
outputs (stored in tests/golden/*.npz) and the oracle."""
import numpy as np
import pytest
import torch

import model_zoo as zoo
from helpers import load_golden, mf_params, check_against_oracle, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ns():
    assert torch.cuda.is_available()
    from brancher_b200 import config
    config.set_device("cuda:0")
    return zoo.namespace("brancher_b200")


def loss_and_grads(ns, model, S, eps):
    """Drive our package exactly as the oracle harness drives the reference."""
    from brancher_b200 import lowering
    with lowering.inject_noise(eps):
        loss = ns.inference.ReverseKL().compute_loss(model, model.posterior_model, None, S)
    loss.backward()
    grads = {v.name: v.link.parameter.grad.detach().cpu().numpy()[0, 0]
             for v in model.posterior_model.flatten() if getattr(v, "learnable", False) and hasattr(v.link, "parameter")}
    return float(loss.detach()), grads


def params_match(model, g):
    """Same construction code + same seeds => same initial parameters as the reference run."""
    for v in model.posterior_model.flatten():
        if getattr(v, "learnable", False) and hasattr(v.link, "parameter"):
            np.testing.assert_allclose(v.link.parameter.detach().cpu().numpy()[0, 0].reshape(g["param"][v.name].shape),
                                       g["param"][v.name], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag,kw", [("bnn_small", dict(seed=1, B=12, P=20, H=7, C=4)),
                                    ("bnn_small_wide", dict(seed=2, B=9, P=16, H=5, C=3, q_sigma=0.3, q_mu_scale=0.5)),
                                    ("bnn_relu", dict(seed=31, B=11, P=14, H=6, C=3, q_sigma=0.3, q_mu_scale=0.5, activation="relu")),
                                    ("bnn_sigmoid", dict(seed=32, B=10, P=12, H=5, C=4, q_sigma=0.3, q_mu_scale=0.5,
                                                         activation="sigmoid"))])
def test_bnn_same_script_same_numbers(ns, tag, kw):
    from oracle import elbo_oracle as O
    g = load_golden(tag)
    model, Q, d = zoo.bnn(ns, **kw)
    params_match(model, g)
    S = g["eps"]["b1"].shape[0]
    loss, grads = loss_and_grads(ns, model, S, g["eps"])
    names = ["weights1", "b1", "weights2", "b2"]
    o64 = O.bnn_elbo(g["raw"]["X"], g["raw"]["y"], mf_params(g, names), g["eps"], dtype=torch.float64,
                     activation=kw.get("activation", "tanh"))
    check_against_oracle(loss, grads, (float(g["raw"]["loss"]), g["grad"]), o64, tag + " via API vs reference")


@pytest.mark.parametrize("tag,tied,seed", [("logreg_tied", True, 3), ("logreg_declared_prior", False, 4)])
def test_logreg_same_script_same_numbers(ns, tag, tied, seed):
    from oracle import elbo_oracle as O
    g = load_golden(tag)
    model, Q, d = zoo.logreg(ns, seed, B=40, F=8, tied=tied)
    params_match(model, g)
    prior = None if tied else {"weights": (g["raw"]["prior_loc"], g["raw"]["prior_scale"])}
    loss, grads = loss_and_grads(ns, model, 16, g["eps"])
    o64 = O.logreg_elbo(g["raw"]["X"], g["raw"]["y"], mf_params(g, ["weights"]), g["eps"], prior, dtype=torch.float64)
    check_against_oracle(loss, grads, (float(g["raw"]["loss"]), g["grad"]), o64, tag + " via API vs reference")


def test_softmax_regression_same_script_same_numbers(ns):
    from oracle import elbo_oracle as O
    g = load_golden("softmax_reg")
    model, Q, d = zoo.softmax_reg(ns, 5, B=24, F=6, C=3)
    params_match(model, g)
    loss, grads = loss_and_grads(ns, model, 8, g["eps"])
    o64 = O.logreg_elbo(g["raw"]["X"], g["raw"]["y"], mf_params(g, ["weights"]), g["eps"], likelihood="categorical",
                        dtype=torch.float64)
    check_against_oracle(loss, grads, (float(g["raw"]["loss"]), g["grad"]), o64, "softmax_reg via API vs reference")


def test_perform_inference_logreg_posterior_mean(ns):
    """End to end under independent (Philox) noise: perform_inference drives the loss down and the
    posterior mean lands on the fp64 oracle-trained optimum within MC error."""
    from brancher_b200 import config
    config.set_seed(123)
    model, Q, d = zoo.logreg(ns, 11, B=400, F=4, tied=False)
    ns.inference.perform_inference(model, number_iterations=300, number_samples=64, optimizer="Adam", lr=0.05)
    curve = model.diagnostics["loss curve"]
    assert curve.shape == (300,) and np.isfinite(curve).all()
    assert curve[-20:].mean() < curve[:20].mean()
    mu = Q[0].roots["loc"].value.detach().cpu().numpy().reshape(-1)
    # reference optimum: maximise the same ELBO with the fp64 oracle by plain Adam on CPU
    from oracle import elbo_oracle as O
    p = {"weights": (np.zeros((1, 4)), O.softplus_inverse(np.ones((1, 4))))}
    t_mu = torch.tensor(p["weights"][0], dtype=torch.float64, requires_grad=True)
    t_rho = torch.tensor(p["weights"][1], dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([t_mu, t_rho], lr=0.05)
    rng = np.random.RandomState(0)
    for _ in range(300):
        eps = {"weights": rng.randn(64, 1, 4)}
        _, gr = O.logreg_elbo(d["X"], d["y"], {"weights": (t_mu.detach().numpy(), t_rho.detach().numpy())}, eps,
                              prior={"weights": (0.0, 0.5)}, dtype=torch.float64)
        opt.zero_grad()
        t_mu.grad = torch.tensor(gr["weights_loc"]); t_rho.grad = torch.tensor(gr["weights_scale"])
        opt.step()
    sd = torch.nn.functional.softplus(t_rho).detach().numpy().reshape(-1)
    assert np.all(np.abs(mu - t_mu.detach().numpy().reshape(-1)) < 4 * sd / np.sqrt(64) + 0.05)


def test_minibatch_model_runs(ns):
    """examples/minibatch_logistic_regression.py structure (RandomIndices + EmpiricalVariable)."""
    rng = np.random.RandomState(0)
    N, F = 50, 2
    xin = np.concatenate([rng.normal(1.5, 1.5, (N // 2, F, 1)), rng.normal(-1.5, 1.5, (N // 2, F, 1))])
    lab = np.concatenate([np.zeros((N // 2, 1)), np.ones((N // 2, 1))])
    idx = ns.RandomIndices(dataset_size=N, batch_size=30, name="indices", is_observed=True)
    x = ns.EmpiricalVariable(xin, indices=idx, name="x", is_observed=True)
    labels = ns.EmpiricalVariable(lab, indices=idx, name="labels", is_observed=True)
    w = ns.NormalVariable(np.zeros((1, F)), 0.5 * np.ones((1, F)), "weights")
    k = ns.BinomialVariable(1, logits=ns.BF.matmul(w, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(labels)
    Qw = ns.NormalVariable(np.zeros((1, F)), np.ones((1, F)), "weights", learnable=True)
    model.set_posterior_model(ns.ProbabilisticModel([Qw]))
    ns.inference.perform_inference(model, number_iterations=100, number_samples=50, optimizer="Adam", lr=0.05)
    curve = model.diagnostics["loss curve"]
    assert np.isfinite(curve).all() and curve[-10:].mean() < curve[:10].mean()
    post = model._get_posterior_sample(50)
    assert post[w].shape[0] == 50
    mu = Qw.roots["loc"].value.detach().cpu().numpy().reshape(-1)
    assert mu[0] < 0 and mu[1] < 0          # class 1 sits at negative coordinates


def test_svgd_same_script_same_numbers(ns):
    """SteinVariationalGradientDescent through the mirrored API: compute_loss -> backward -> correct_gradient on the
    playground's particle ensemble == the live reference's numbers (tests/golden/svgd_softmax.npz)."""
    from helpers import assert_close
    z = load_golden("svgd_softmax")["raw"]
    model, particles, d = zoo.svgd_softmax(ns, 8, B=30, F=5, C=3, n=6)
    m = ns.inference.SteinVariationalGradientDescent()
    m.check_model_compatibility(model, particles, None)
    loss = m.compute_loss(model, particles, None, 1)
    loss.backward()
    raw = np.stack([list(p.flatten())[0].value.grad.detach().cpu().numpy().reshape(3, 5) for p in particles])
    assert_close(float(loss.detach()), z["loss"], "svgd loss via API", rtol=1e-5, atol=1e-6)
    assert_close(raw, z["raw_grad"], "svgd raw grads via API", rtol=1e-5, atol=1e-6, scale=np.abs(z["raw_grad"]).max())
    m.correct_gradient(model, particles, None, 1)
    out = np.stack([list(p.flatten())[0].value.grad.detach().cpu().numpy().reshape(3, 5) for p in particles])
    assert abs(float(m.bandwidth) - float(z["bandwidth"])) <= 2e-6 * float(z["bandwidth"])
    assert_close(out, z["out"], "svgd direction via API", rtol=1e-5, atol=1e-6, scale=np.abs(z["out"]).max())


def test_svgd_perform_inference_runs(ns):
    model, particles, d = zoo.svgd_softmax(ns, 9, B=40, F=4, C=3, n=8)
    before = np.stack([list(p.flatten())[0].value.detach().cpu().numpy().copy() for p in particles])
    ns.inference.perform_inference(model, inference_method=ns.inference.SteinVariationalGradientDescent(),
                                   number_iterations=20, number_samples=1, optimizer="SGD", lr=0.0025,
                                   posterior_model=particles)
    curve = model.diagnostics["loss curve"]
    after = np.stack([list(p.flatten())[0].value.detach().cpu().numpy() for p in particles])
    assert curve.shape == (20,) and np.isfinite(curve).all() and curve[-1] < curve[0]
    assert np.abs(after - before).max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n,B,F,C,S,seed", [("wvgd_softmax", 3, 30, 4, 3, 20, 21), ("wvgd_softmax4", 4, 16, 5, 2, 24, 22)])
def test_wvgd_api_matches_reference(tag, n, B, F, C, S, seed):
    """The reference's WVGD script shape through the mirrored API: WassersteinVariationalGradientDescent.compute_loss +
    backward land the reference's gradients in every sampler / particle parameter."""
    import os
    from helpers import GOLDEN
    from brancher_b200 import config, lowering
    config.set_device("cuda:0")
    ns = zoo.namespace("brancher_b200")
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    model, particles, samplers, d = zoo.wvgd_softmax(ns, seed, B, F, C, n)
    m = ns.inference.WassersteinVariationalGradientDescent(variational_samplers=samplers, particles=particles, biased=False)
    m.check_model_compatibility(model, particles, m.sampler_model)
    with lowering.inject_noise({"elbo": g["eps_elbo"], "particle": g["eps_particle"]}):
        loss = m.compute_loss(model, particles, m.sampler_model, S)
    loss.backward()
    assert_close(float(loss.detach()), g["loss"], tag + " loss", rtol=1e-5, atol=1e-6)
    by = lambda mdl, name: [v for v in mdl.flatten() if v.name == name][0]
    got_loc = np.stack([by(s_, "weights_loc").value.grad.cpu().numpy().reshape(C, F) for s_ in samplers])
    got_rho = np.stack([by(s_, "weights_scale").value.grad.cpu().numpy().reshape(()) for s_ in samplers])
    got_th = np.stack([by(p_, "weights").value.grad.cpu().numpy().reshape(C, F) for p_ in particles])
    for got, key in ((got_loc, "grad_loc"), (got_rho, "grad_rho"), (got_th, "grad_theta")):
        assert_close(got, g[key], tag + " " + key, rtol=1e-5, atol=1e-6, scale=np.abs(g[key]).max())


@pytest.mark.gpu
def test_wvgd_perform_inference_runs():
    from brancher_b200 import config
    config.set_device("cuda:0")
    ns = zoo.namespace("brancher_b200")
    model, particles, samplers, d = zoo.wvgd_softmax(ns, 23, 40, 4, 3, 4)
    m = ns.inference.WassersteinVariationalGradientDescent(variational_samplers=samplers, particles=particles)
    ns.inference.perform_inference(model, inference_method=m, number_iterations=30, number_samples=50, optimizer="Adam", lr=0.01,
                                   posterior_model=particles)
    curve = model.diagnostics["loss curve"]
    assert curve.shape == (30,) and np.isfinite(curve).all()
    assert curve[-5:].mean() < curve[:5].mean()


@pytest.mark.parametrize("tag,kw", [("vae_small", dict(seed=12, B=6, D=12, L=2, h_enc=(5, 7), h_dec=(7, 5))),
                                    ("vae_deep", dict(seed=13, B=10, D=20, L=3, h_enc=(9,), h_dec=(6, 8, 5)))])
def test_vae_same_script_same_numbers(ns, tag, kw):
    """C5 through the Brancher API (examples/VAE_playground.py:65-80): the model-construction script that produced the golden
    vectors with the reference -- BrancherFunction(nn.Module) encoder / decoder, Index links -- runs against brancher_b200,
    lowers to K5 (lowering.VaePlan) and reproduces the reference's loss (incl. the -ln S constant) and every module
    parameter's gradient."""
    from brancher_b200 import lowering
    g = load_golden(tag)
    model, Q, d = zoo.vae(ns, **kw)
    d["enc"].to("cuda:0"); d["dec"].to("cuda:0")
    S = g["eps"]["z"].shape[0]
    with lowering.inject_noise({"z": g["eps"]["z"]}):
        loss = ns.inference.ReverseKL().compute_loss(model, model.posterior_model, None, S)
    assert lowering.get_plan(model, model.posterior_model).family.startswith("vae")
    loss.backward()
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden_names", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    grads = {}
    for side, net, heads in (("enc", d["enc"], (("W_mean", "l_mean"), ("W_sd", "l_sd"))), ("dec", d["dec"], (("W_out", "l_out"),))):
        for i, l in enumerate(net.hidden):
            grads["%s.W.%d" % (side, i)] = l.weight.grad.cpu().numpy()
            grads["%s.b.%d" % (side, i)] = l.bias.grad.cpu().numpy()
        for key, attr in heads:
            grads["%s.%s" % (side, key)] = getattr(net, attr).weight.grad.cpu().numpy()
            grads["%s.b%s" % (side, key[1:])] = getattr(net, attr).bias.grad.cpu().numpy()
    from oracle import elbo_oracle as O
    from helpers import vae_nets
    enc, dec = vae_nets(g)
    o64 = O.vae_elbo(g["raw"]["X"], enc, dec, g["eps"]["z"], dtype=torch.float64)
    check_against_oracle(float(loss.detach()), grads, (float(g["raw"]["loss"]), g["grad"]), o64, tag + " via API vs reference")


def test_vae_perform_inference_runs_and_improves(ns):
    """the VAE trains through perform_inference (graph-captured loop: K5 evaluation + fused Adam over the modules' parameters)"""
    from brancher_b200 import config, inference
    config.set_seed(5)
    model, Q, d = zoo.vae(ns, 21, B=64, D=30, L=2, h_enc=(16,), h_dec=(16,))
    d["enc"].to("cuda:0"); d["dec"].to("cuda:0")
    inference.perform_inference(model, number_iterations=150, number_samples=8, optimizer="Adam", lr=0.01,
                                inference_method=inference.ReverseKL())
    curve = np.asarray(model.diagnostics["loss curve"]).reshape(-1)
    assert inference.last_loop == "graph"
    assert np.isfinite(curve).all() and curve[-20:].mean() < curve[:20].mean()


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n,B,F,C,seed", [("wvgd_post", 3, 30, 4, 3, 21), ("wvgd_post4", 4, 16, 5, 2, 22)])
def test_wvgd_post_process_matches_reference(tag, n, B, F, C, seed):
    """WassersteinVariationalGradientDescent.post_process (inference.py:234-247): ensemble weights and the per-sampler log
    normalisers of the importance weights, on the device, against the LIVE reference's own output on the same injected draws
    (tests/golden/make_golden.py: wvgd_post) and against the fp64 oracle; then with a declared (untied) prior and with the
    kernel's own Philox draws (consistency: weights are a distribution, every sampler accepted something)."""
    import os
    from helpers import GOLDEN
    from brancher_b200 import config, lowering
    from oracle import elbo_oracle as O
    config.set_device("cuda:0")
    ns = zoo.namespace("brancher_b200")
    g = np.load(os.path.join(GOLDEN, tag + ".npz"))
    S = g["eps_post"].shape[1]
    model, particles, samplers, d = zoo.wvgd_softmax(ns, seed, B, F, C, n)
    m = ns.inference.WassersteinVariationalGradientDescent(variational_samplers=samplers, particles=particles, biased=False,
                                                           number_post_samples=S)
    m.check_model_compatibility(model, particles, m.sampler_model)
    with lowering.inject_noise({"post": g["eps_post"]}):
        m.post_process(model)
    w64, lz64, c64 = O.wvgd_ensemble_weights(g["X"], g["y"], g["theta"], g["loc"], g["rho"], g["eps_post"])
    assert (m.accepted_post == c64).all()
    assert_close(m.log_normalizers, g["logZ"], tag + " logZ vs reference", rtol=1e-5, atol=1e-5)
    assert_close(m.log_normalizers, lz64, tag + " logZ vs fp64 oracle", rtol=1e-5, atol=1e-5)
    assert_close(m.weights, g["weights"], tag + " weights vs reference", rtol=1e-4, atol=1e-7)
    assert_close(m.weights, w64, tag + " weights vs fp64 oracle", rtol=1e-4, atol=1e-7)
    # Philox draws, many samples, several chunks
    m.number_post_samples = 20000
    m.post_process(model)
    assert m.weights.shape == (n,) and abs(m.weights.sum() - 1) < 1e-6 and (m.accepted_post > 0).all()
    assert m.accepted_post.sum() <= 20000 * n


@pytest.mark.gpu
def test_map_logistic_regression_matches_oracle():
    """examples/MAP_logistic_regression.py shape: MAP().compute_loss = -log p(data, theta) and its gradient through K4a vs the
    fp64 oracle; perform_inference drives it down."""
    from brancher_b200 import config
    from oracle import elbo_oracle as O
    config.set_device("cuda:0")
    ns = zoo.namespace("brancher_b200")
    rng = np.random.RandomState(3)
    B, F, C = 60, 5, 3
    X = rng.randn(B, F, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = ns.RootVariable(X, "x", is_observed=True)
    weights = ns.NormalVariable(np.zeros((C, F)), 10 * np.ones((C, F)), "weights")
    k = ns.CategoricalVariable(logits=ns.BF.matmul(weights, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(y)
    w0 = rng.randn(C, F)
    point = ns.ProbabilisticModel([ns.RootVariable(w0, name="weights", learnable=True)])
    model.set_posterior_model(point)
    m = ns.inference.MAP()
    m.check_model_compatibility(model, model.posterior_model, None)
    loss = m.compute_loss(model, model.posterior_model, None, 1)
    loss.backward()
    root = [v for v in point.flatten() if v.name == "weights"][0]
    l64, g64 = O.particles_loss_grad(X[:, :, 0], y, w0[None].astype("f4"), (np.zeros((C, F), "f4"), np.full((C, F), 10.0, "f4")),
                                     dtype=torch.float64, likelihood="categorical")
    assert_close(float(loss.detach()), l64, "MAP loss")
    assert_close(root.value.grad.cpu().numpy().reshape(C, F), g64[0], "MAP gradient", scale=np.abs(g64).max())
    ns.inference.perform_inference(model, inference_method=ns.inference.MAP(), number_iterations=40, number_samples=1,
                                   optimizer="SGD", lr=0.01)
    curve = model.diagnostics["loss curve"]
    assert curve.shape == (40,) and curve[-1] < curve[0]


@pytest.mark.gpu
def test_posterior_predictive_api():
    """ProbabilisticModel.get_posterior_predictive: one batched launch sequence for all test rows and posterior samples"""
    from brancher_b200 import config
    config.set_device("cuda:0")
    ns = zoo.namespace("brancher_b200")
    model, Q, d = zoo.bnn(ns, 1, B=40, P=30, H=100, C=5, q_sigma=0.05, q_mu_scale=0.3)
    x = [v for v in model._flatten() if v.name == "x"][0]
    out = model.get_posterior_predictive(12, {x: d["X"][:25]})
    assert tuple(out["logits"].shape) == (12, 25, 5) and tuple(out["samples"].shape) == (12, 25)
    p = out["probs"].cpu().numpy()
    assert p.shape == (25, 5) and np.allclose(p.sum(1), 1.0, atol=1e-5)
    want = torch.softmax(out["logits"], -1).mean(0).cpu().numpy()
    np.testing.assert_allclose(p, want, rtol=1e-5, atol=1e-6)
