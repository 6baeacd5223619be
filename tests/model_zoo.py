"""Model-construction code shared VERBATIM between the reference and brancher_b200.

Every builder takes `ns`, a namespace exposing the public names of either package
(`brancher.*` in tests/golden/make_golden.py, `brancher_b200.*` in the parity tests), and builds the
model through that package's public API only -- the drop-in claim, executed.
Each returns (model, q_variables, data dict).
"""
import types

import numpy as np


def namespace(pkg):
    """pkg = 'brancher' (reference) or 'brancher_b200'."""
    import importlib
    V = importlib.import_module(pkg + ".variables")
    SV = importlib.import_module(pkg + ".standard_variables")
    ns = types.SimpleNamespace(
        pkg=pkg, RootVariable=V.RootVariable, ProbabilisticModel=V.ProbabilisticModel,
        NormalVariable=SV.NormalVariable, CategoricalVariable=SV.CategoricalVariable,
        BinomialVariable=SV.BinomialVariable, DeterministicVariable=SV.DeterministicVariable,
        LogNormalVariable=SV.LogNormalVariable, EmpiricalVariable=SV.EmpiricalVariable, RandomIndices=SV.RandomIndices,
        CauchyVariable=SV.CauchyVariable, LaplaceVariable=SV.LaplaceVariable,
        LogitNormalVariable=getattr(SV, "LogitNormalVariable", None),
        BF=importlib.import_module(pkg + ".functions"), inference=importlib.import_module(pkg + ".inference"))
    return ns


def bnn(ns, seed, B, P, H, C, q_sigma=0.01, q_mu_scale=0.0, activation="tanh"):
    """development_playgrounds/MNIST_bayesian_neural_network.py:26-57 on synthetic data."""
    rng = np.random.RandomState(seed)
    X = rng.rand(B, P, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = ns.RootVariable(X, "x", is_observed=True)
    shapes = {"b1": (H, 1), "b2": (C, 1), "weights1": (H, P), "weights2": (C, H)}
    pv = {n: ns.NormalVariable(np.zeros(s), 10 * np.ones(s), n) for n, s in shapes.items()}
    h = getattr(ns.BF, activation)(ns.BF.matmul(pv["weights1"], x) + pv["b1"])
    a = ns.BF.matmul(pv["weights2"], h) + pv["b2"]
    k = ns.CategoricalVariable(logits=a, name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(y)
    mu0 = {n: (q_mu_scale * rng.randn(*s)).astype("float32") for n, s in shapes.items()}
    sg0 = {n: (q_sigma * (1 + rng.rand(*s))).astype("float32") for n, s in shapes.items()}
    Q = [ns.NormalVariable(mu0[n].astype("float64"), sg0[n].astype("float64"), n, learnable=True) for n in shapes]
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    return model, Q, {"X": X[:, :, 0], "y": y, "shapes": shapes, "rng": rng}


def logreg(ns, seed, B, F, tied):
    """examples/minibatch_logistic_regression.py:27-43 shape, batch passed as an observed root
    (development_playgrounds/bayesian_logistic_regression_playground.py:22-30)."""
    rng = np.random.RandomState(seed)
    X = rng.randn(B, F, 1).astype("float32")
    wtrue = rng.randn(F) / np.sqrt(F)
    y = (rng.rand(B) < 1 / (1 + np.exp(-X[:, :, 0] @ wtrue))).astype("float32").reshape(B, 1)
    x = ns.RootVariable(X, "x", is_observed=True)
    if tied:   # numeric hyper-parameters on both sides: roots collide by name (every example does this)
        weights = ns.NormalVariable(np.zeros((1, F)), 0.5 * np.ones((1, F)), "weights")
    else:      # p's roots named distinctly -> the declared prior N(0, 0.5) is what is evaluated
        weights = ns.NormalVariable(ns.RootVariable(np.zeros((1, F)), "prior_loc"),
                                    ns.RootVariable(0.5 * np.ones((1, F)), "prior_scale"), "weights")
    k = ns.BinomialVariable(1, logits=ns.BF.matmul(weights, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(y)
    mu0 = (0.3 * rng.randn(1, F)).astype("float32")
    sg0 = (0.5 + rng.rand(1, F)).astype("float32")
    Q = [ns.NormalVariable(mu0.astype("float64"), sg0.astype("float64"), "weights", learnable=True)]
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    return model, Q, {"X": X[:, :, 0], "y": y[:, 0], "rng": rng}


def softmax_reg(ns, seed, B, F, C):
    """examples/MNIST_logistic_regression.py shape: Categorical(logits = W x), W [C,F]."""
    rng = np.random.RandomState(seed)
    X = rng.randn(B, F, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = ns.RootVariable(X, "x", is_observed=True)
    weights = ns.NormalVariable(np.zeros((C, F)), 10 * np.ones((C, F)), "weights")
    k = ns.CategoricalVariable(logits=ns.BF.matmul(weights, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(y)
    mu0 = (0.3 * rng.randn(C, F)).astype("float32")
    sg0 = (0.1 + 0.2 * rng.rand(C, F)).astype("float32")
    Q = [ns.NormalVariable(mu0.astype("float64"), sg0.astype("float64"), "weights", learnable=True)]
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    return model, Q, {"X": X[:, :, 0], "y": y, "rng": rng}


def svgd_softmax(ns, seed, B, F, C, n):
    """development_playgrounds/SVGD_logistic_regression.py:34-48: softmax regression, one single-root
    ProbabilisticModel per particle (logits= instead of the playground's broken softmax_p=)."""
    rng = np.random.RandomState(seed)
    X = rng.randn(B, F, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = ns.RootVariable(X, "x", is_observed=True)
    weights = ns.NormalVariable(np.zeros((C, F)), 10 * np.ones((C, F)), "weights")
    k = ns.CategoricalVariable(logits=ns.BF.matmul(weights, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(y)
    theta0 = rng.randn(n, C, F).astype("float32")
    particles = [ns.ProbabilisticModel([ns.RootVariable(theta0[i].astype("float64"), name="weights", learnable=True)])
                 for i in range(n)]
    return model, particles, {"X": X[:, :, 0], "y": y, "theta": theta0, "rng": rng}


def lognormal_normal(ns, seed, N):
    """examples/logNormal_normal.py:11-30 with synthetic observations (N rows)."""
    rng = np.random.RandomState(seed)
    data = (-2.0 + 1.0 * rng.randn(N, 1, 1)).astype("float32")
    nu = ns.LogNormalVariable(0., 1., "nu")
    mu = ns.NormalVariable(0., 10., "mu")
    x = ns.NormalVariable(mu, nu, "x")
    model = ns.ProbabilisticModel([x])
    x.observe(data)
    Qnu = ns.LogNormalVariable(0., 1., "nu", learnable=True)
    Qmu = ns.NormalVariable(0., 1., "mu", learnable=True)
    model.set_posterior_model(ns.ProbabilisticModel([Qmu, Qnu]))
    return model, [Qmu, Qnu], {"x": data[:, 0, 0], "rng": rng}


def multivariate_regression(ns, seed, n):
    """examples/multivariate_regression.py:11-44: observed deterministic regressors, 4 Normal weights, LogNormal noise."""
    rng = np.random.RandomState(seed)
    x_range = np.linspace(-1., 1., n)
    x1v, x2v = np.sin(2 * np.pi * 2 * x_range), x_range
    x1 = ns.DeterministicVariable(x1v, name="x1", is_observed=True)
    x2 = ns.DeterministicVariable(x2v, name="x2", is_observed=True)
    b = ns.NormalVariable(0., 1., name="b")
    w1 = ns.NormalVariable(0., 1., name="w1")
    w2 = ns.NormalVariable(0., 1., name="w2")
    w12 = ns.NormalVariable(0., 1., name="w12")
    nu = ns.LogNormalVariable(0.2, 0.5, name="nu")
    mean = b + w1 * x1 + w2 * x2 + w12 * x1 * x2
    y = ns.NormalVariable(mean, nu, name="y")
    model = ns.ProbabilisticModel([y])
    Q = [ns.NormalVariable(0.1 * k, 1., name=nm, learnable=True) for k, nm in enumerate(["b", "w1", "w2", "w12"])]
    Q.append(ns.LogNormalVariable(0.2, 0.5, name="nu", learnable=True))
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    ydata = (0.3 + 0.8 * x1v - 0.5 * x2v + 0.2 * x1v * x2v + 0.4 * rng.randn(n)).astype("float32")
    y.observe(ydata.reshape(n, 1, 1))
    return model, Q, {"y": ydata, "x1": x1v.astype("float32"), "x2": x2v.astype("float32"), "rng": rng}


def robust_regression(ns, seed, n):
    """Heavy-tailed variant of examples/multivariate_regression.py: Laplace priors on the weights
    (standard_variables.py:171-183), Cauchy likelihood (standard_variables.py:156-168), LogNormal noise scale; mean-field
    Normal / LogNormal posterior."""
    rng = np.random.RandomState(seed)
    xv = np.linspace(-1., 1., n)
    x = ns.DeterministicVariable(xv, name="x", is_observed=True)
    b = ns.LaplaceVariable(0., 1., name="b")
    w = ns.LaplaceVariable(0., 2., name="w")
    nu = ns.LogNormalVariable(-1., 0.5, name="nu")
    y = ns.CauchyVariable(b + w * x, nu, name="y")
    model = ns.ProbabilisticModel([y])
    Q = [ns.NormalVariable(0.2, 0.7, name="b", learnable=True), ns.NormalVariable(-0.1, 0.9, name="w", learnable=True),
         ns.LogNormalVariable(-1., 0.5, name="nu", learnable=True)]
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    ydata = (0.3 + 0.8 * xv + 0.3 * rng.standard_cauchy(n)).astype("float32")
    y.observe(ydata.reshape(n, 1, 1))
    return model, Q, {"y": ydata, "x": xv.astype("float32"), "rng": rng}


def scalar_logistic(ns, seed, n):
    """One-feature Bayesian logistic regression written with scalar links (no matmul): Binomial(1, logits = w x + b) observed,
    Normal priors and posterior -- the Bernoulli/Binomial node of distributions.py:561-592 in the scalar-DAG family."""
    rng = np.random.RandomState(seed)
    xv = np.linspace(-2., 2., n)
    x = ns.DeterministicVariable(xv, name="x", is_observed=True)
    w = ns.NormalVariable(0., 1., name="w")
    b = ns.NormalVariable(0., 1., name="b")
    k = ns.BinomialVariable(1, logits=w * x + b, name="k")
    model = ns.ProbabilisticModel([k])
    Q = [ns.NormalVariable(0.3, 0.6, name="w", learnable=True), ns.NormalVariable(-0.2, 0.8, name="b", learnable=True)]
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    labels = (rng.rand(n) < 1.0 / (1.0 + np.exp(-(1.5 * xv - 0.3)))).astype("float32")
    k.observe(labels.reshape(n, 1, 1))
    return model, Q, {"k": labels, "x": xv.astype("float32"), "rng": rng}


def op_zoo(ns, seed, n):
    """Regression whose mean exercises every unary / binary op of the scalar-DAG family that the example models do not
    (sin, cos, tanh, relu, abs, sqrt, exp, log, log1p, sigmoid, softplus, pow, neg, div): functions.py:28-62 lifts each torch
    function to a link, variables.py:246-295 the operators."""
    rng = np.random.RandomState(seed)
    xv = np.linspace(0.2, 2.0, n)
    x = ns.DeterministicVariable(xv, name="x", is_observed=True)
    a = ns.NormalVariable(0., 1., name="a")
    b = ns.NormalVariable(0., 1., name="b")
    c = ns.LogNormalVariable(0., 0.3, name="c")
    BF = ns.BF
    mean = (BF.sin(a * x) + BF.tanh(b) * BF.cos(x) + BF.relu(a - 0.1) - BF.abs(b) * 0.3 + BF.sqrt(c * x) + BF.exp(-c)
            + a / (c + 1.0) + b ** 2 + BF.log1p(BF.exp(a)) * 0.1 + BF.sigmoid(b) + BF.softplus(a) * 0.2 + BF.log(c + 1.0))
    y = ns.NormalVariable(mean, 0.5, name="y")
    model = ns.ProbabilisticModel([y])
    Q = [ns.NormalVariable(0.4, 0.5, name="a", learnable=True), ns.NormalVariable(-0.3, 0.6, name="b", learnable=True),
         ns.LogNormalVariable(0.1, 0.3, name="c", learnable=True)]
    model.set_posterior_model(ns.ProbabilisticModel(Q))
    ydata = (1.0 + np.sin(0.5 * xv) + 0.5 * rng.randn(n)).astype("float32")
    y.observe(ydata.reshape(n, 1, 1))
    return model, Q, {"y": ydata, "x": xv.astype("float32"), "rng": rng}


def ar1(ns, seed, T):
    """README.md:22-75 model, y0 named 'y0' (the README reuses 'x0')."""
    rng = np.random.RandomState(seed)
    driving, measure, btrue = 1.0, 0.3, 0.7
    xs = [rng.randn() * driving]
    for t in range(1, T):
        xs.append(btrue * xs[-1] + driving * rng.randn())
    ydata = np.array(xs) + measure * rng.randn(T)
    x0 = ns.NormalVariable(0., driving, "x0")
    y0 = ns.NormalVariable(x0, measure, "y0")
    b = ns.LogitNormalVariable(0.5, 1., "b")
    x, y = [x0], [y0]
    for t in range(1, T):
        x.append(ns.NormalVariable(b * x[t - 1], driving, "x%d" % t))
        y.append(ns.NormalVariable(x[t], measure, "y%d" % t))
    model = ns.ProbabilisticModel(x + y)
    for t, yt in enumerate(y):
        yt.observe(float(ydata[t]))
    Qb = ns.LogitNormalVariable(0.5, 0.5, "b", learnable=True)
    logit_b_post = ns.DeterministicVariable(0., "logit_b_post", learnable=True)
    Qx = [ns.NormalVariable(0., 1., "x0", learnable=True)]
    Qx_mean = [ns.DeterministicVariable(0., "x0_mean", learnable=True)]
    for t in range(1, T):
        Qx_mean.append(ns.DeterministicVariable(0.1 * rng.randn(), "x%d_mean" % t, learnable=True))
        Qx.append(ns.NormalVariable(ns.BF.sigmoid(logit_b_post) * Qx[t - 1] + Qx_mean[t], 1., "x%d" % t, learnable=True))
    model.set_posterior_model(ns.ProbabilisticModel([Qb] + Qx))
    return model, [Qb] + Qx, {"y": ydata.astype("float32"), "measure_noise": measure, "rng": rng}


def vae_modules(D, L, h_enc, h_dec, seed):
    """Encoder / decoder of examples/VAE_playground.py:30-63 (ReLU MLPs; encoder heads mean and softplus(.)+0.1) with
    configurable widths; plain torch.nn so both packages wrap the same objects with BF.BrancherFunction."""
    import torch
    import torch.nn as nn

    class EncoderArchitecture(nn.Module):
        def __init__(self):
            super().__init__()
            dims = [D] + list(h_enc)
            self.hidden = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
            self.f = nn.ReLU()
            self.l_mean = nn.Linear(dims[-1], L)
            self.l_sd = nn.Linear(dims[-1], L)
            self.softplus = nn.Softplus()

        def __call__(self, x):
            h = x.squeeze(-1)
            for l in self.hidden:
                h = self.f(l(h))
            return {"mean": self.l_mean(h), "sd": self.softplus(self.l_sd(h)) + 0.1}

    class DecoderArchitecture(nn.Module):
        def __init__(self):
            super().__init__()
            dims = [L] + list(h_dec)
            self.hidden = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])
            self.f = nn.ReLU()
            self.l_out = nn.Linear(dims[-1], D)

        def __call__(self, z):
            h = z
            for l in self.hidden:
                h = self.f(l(h))
            return {"mean": self.l_out(h)}

    torch.manual_seed(seed)
    return EncoderArchitecture(), DecoderArchitecture()


def vae(ns, seed, B, D, L, h_enc, h_dec):
    """examples/VAE_playground.py:65-80 on synthetic binarised data; the minibatch is the fixed index list 0..B-1 shared
    by all MC samples (EmpiricalVariable(indices=...), distributions.py:443-455)."""
    rng = np.random.RandomState(seed)
    dataset = (rng.rand(B, D, 1) < 0.5).astype("int32")
    enc, dec = vae_modules(D, L, h_enc, h_dec, seed)
    encoder = ns.BF.BrancherFunction(enc)
    decoder = ns.BF.BrancherFunction(dec)
    z = ns.NormalVariable(np.zeros((L,)), np.ones((L,)), name="z")
    decoder_output = ns.DeterministicVariable(decoder(z), name="decoder_output")
    x = ns.BinomialVariable(total_count=1, logits=decoder_output["mean"], name="x")
    model = ns.ProbabilisticModel([x, z])
    Qx = ns.EmpiricalVariable(dataset, indices=list(range(B)), name="x", is_observed=True)
    encoder_output = ns.DeterministicVariable(encoder(Qx), name="encoder_output")
    Qz = ns.NormalVariable(encoder_output["mean"], encoder_output["sd"], name="z")
    model.set_posterior_model(ns.ProbabilisticModel([Qx, Qz]))
    return model, [Qz], {"X": dataset[:, :, 0].astype("float32"), "enc": enc, "dec": dec, "rng": rng}


def wvgd_softmax(ns, seed, B, F, C, n, spread=1.0, q_sigma=0.3):
    """development_playgrounds/WVGD_logistic_regression.py:33-58: softmax regression; one single-root particle model and
    one single-Normal sampler model (learnable loc/scale, same variable name "weights") per particle."""
    rng = np.random.RandomState(seed)
    X = rng.randn(B, F, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = ns.RootVariable(X, "x", is_observed=True)
    weights = ns.NormalVariable(np.zeros((C, F)), 10 * np.ones((C, F)), "weights")
    k = ns.CategoricalVariable(logits=ns.BF.matmul(weights, x), name="k")
    model = ns.ProbabilisticModel([k])
    k.observe(y)
    theta0 = (spread * rng.randn(n, C, F)).astype("float32")
    loc0 = (theta0 + 0.05 * rng.randn(n, C, F)).astype("float32")
    particles = [ns.ProbabilisticModel([ns.RootVariable(theta0[i].astype("float64"), name="weights", learnable=True)])
                 for i in range(n)]
    samplers = [ns.ProbabilisticModel([ns.NormalVariable(loc=loc0[i].astype("float64"), scale=q_sigma, name="weights",
                                                         learnable=True)]) for i in range(n)]
    return model, particles, samplers, {"X": X[:, :, 0], "y": y, "theta": theta0, "loc": loc0, "q_sigma": q_sigma, "rng": rng}
