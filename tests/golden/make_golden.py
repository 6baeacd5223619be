"""Generate golden input/output vectors from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference (read-only), builds each config-shaped model
at a small size through the reference's own public API, injects fixed standard-normal noise per q
variable (by overriding that variable's `distribution._get_sample`, the hook SURVEY §8c describes),
evaluates `ReverseKL().compute_loss(...)` + `.backward()` (inference.py:140-144,100) and stores
inputs + loss + every parameter gradient (keyed by the reference's variable names) in
tests/golden/*.npz.  The GPU box has no /root/reference: tests read only the .npz files.
"""
import os
import sys
import warnings

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import numpy as np
import torch

from brancher.variables import RootVariable, ProbabilisticModel
from brancher.standard_variables import (NormalVariable, CategoricalVariable, BinomialVariable,
                                         DeterministicVariable, VariableConstructor)
from brancher import inference, distributions, geometric_ranges
import brancher.functions as BF

HERE = os.path.dirname(os.path.abspath(__file__))


def inject(qvars, eps, transform=None):
    """Replace rsample of each q variable by loc + eps*scale with OUR eps (shape (S,1,*event))."""
    for q in qvars:
        def f(differentiable, _n=q.name, **p):
            z = p["loc"] + eps[_n] * p["scale"]
            return transform[_n](z) if transform and _n in transform else z
        q.distribution._get_sample = f


def loss_and_grads(model, S):
    model.update_observed_submodel()
    loss = inference.ReverseKL().compute_loss(model, model.posterior_model, None, S)
    loss.backward()
    grads = {v.name: v.link.parameter.grad.detach().numpy().copy()
             for v in model.posterior_model.flatten() if getattr(v, "learnable", False)}
    values = {v.name: v.link.parameter.detach().numpy().copy()
              for v in model.posterior_model.flatten() if getattr(v, "learnable", False)}
    return float(loss.detach()), grads, values


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: np.asarray(v).shape for k, v in arrays.items() if not k.startswith("grad_")})


def flat(prefix, d):
    return {prefix + k: np.asarray(v) for k, v in d.items()}


# ---------------------------------------------------------------------------------------------
def bnn(seed, B, P, H, C, S, q_sigma=0.01, q_mu_scale=0.0, tag="bnn_small"):
    rng = np.random.RandomState(seed)
    X = rng.rand(B, P, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = RootVariable(X, "x", is_observed=True)
    shapes = {"b1": (H, 1), "b2": (C, 1), "weights1": (H, P), "weights2": (C, H)}
    pv = {n: NormalVariable(np.zeros(s), 10 * np.ones(s), n) for n, s in shapes.items()}
    h = BF.tanh(BF.matmul(pv["weights1"], x) + pv["b1"])
    a = BF.matmul(pv["weights2"], h) + pv["b2"]
    k = CategoricalVariable(logits=a, name="k")
    model = ProbabilisticModel([k])
    k.observe(y)
    mu0 = {n: (q_mu_scale * rng.randn(*s)).astype("float32") for n, s in shapes.items()}
    sg0 = {n: (q_sigma * (1 + rng.rand(*s))).astype("float32") for n, s in shapes.items()}
    Q = [NormalVariable(mu0[n].astype("float64"), sg0[n].astype("float64"), n, learnable=True) for n in shapes]
    model.set_posterior_model(ProbabilisticModel(Q))
    eps = {n: torch.tensor(rng.randn(S, 1, *s).astype("float32")) for n, s in shapes.items()}
    inject(Q, eps)
    loss, grads, values = loss_and_grads(model, S)
    save(tag, X=X[:, :, 0], y=y, loss=loss,
         **flat("eps_", {n: e.numpy()[:, 0] for n, e in eps.items()}),
         **flat("param_", {n: v[0, 0] for n, v in values.items()}),
         **flat("grad_", {n: g[0, 0] for n, g in grads.items()}))


def logreg(seed, B, F, S, tied, tag):
    rng = np.random.RandomState(seed)
    X = rng.randn(B, F, 1).astype("float32")
    wtrue = rng.randn(F) / np.sqrt(F)
    y = (rng.rand(B) < 1 / (1 + np.exp(-X[:, :, 0] @ wtrue))).astype("float32").reshape(B, 1)
    x = RootVariable(X, "x", is_observed=True)
    if tied:   # numeric hyper-parameters on both sides: roots collide by name (every example does this)
        weights = NormalVariable(np.zeros((1, F)), 0.5 * np.ones((1, F)), "weights")
    else:      # p's roots named distinctly -> the declared prior N(0, 0.5) is what is evaluated
        weights = NormalVariable(RootVariable(np.zeros((1, F)), "prior_loc"),
                                 RootVariable(0.5 * np.ones((1, F)), "prior_scale"), "weights")
    k = BinomialVariable(1, logits=BF.matmul(weights, x), name="k")
    model = ProbabilisticModel([k])
    k.observe(y)
    mu0 = (0.3 * rng.randn(1, F)).astype("float32")
    sg0 = (0.5 + rng.rand(1, F)).astype("float32")
    Q = [NormalVariable(mu0.astype("float64"), sg0.astype("float64"), "weights", learnable=True)]
    model.set_posterior_model(ProbabilisticModel(Q))
    eps = {"weights": torch.tensor(rng.randn(S, 1, 1, F).astype("float32"))}
    inject(Q, eps)
    loss, grads, values = loss_and_grads(model, S)
    save(tag, X=X[:, :, 0], y=y[:, 0], loss=loss, tied=int(tied),
         prior_loc=np.zeros((1, F), "float32"), prior_scale=0.5 * np.ones((1, F), "float32"),
         eps_weights=eps["weights"].numpy()[:, 0],
         **flat("param_", {n: v[0, 0] for n, v in values.items()}),
         **flat("grad_", {n: g[0, 0] for n, g in grads.items()}))


def softmax_reg(seed, B, F, C, S, tag):
    """MNIST_logistic_regression-shaped: Categorical(logits = W x), W [C,F]."""
    rng = np.random.RandomState(seed)
    X = rng.randn(B, F, 1).astype("float32")
    y = rng.randint(0, C, size=(B,))
    x = RootVariable(X, "x", is_observed=True)
    weights = NormalVariable(np.zeros((C, F)), 10 * np.ones((C, F)), "weights")
    k = CategoricalVariable(logits=BF.matmul(weights, x), name="k")
    model = ProbabilisticModel([k])
    k.observe(y)
    mu0 = (0.3 * rng.randn(C, F)).astype("float32")
    sg0 = (0.1 + 0.2 * rng.rand(C, F)).astype("float32")
    Q = [NormalVariable(mu0.astype("float64"), sg0.astype("float64"), "weights", learnable=True)]
    model.set_posterior_model(ProbabilisticModel(Q))
    eps = {"weights": torch.tensor(rng.randn(S, 1, C, F).astype("float32"))}
    inject(Q, eps)
    loss, grads, values = loss_and_grads(model, S)
    save(tag, X=X[:, :, 0], y=y, loss=loss, eps_weights=eps["weights"].numpy()[:, 0],
         **flat("param_", {n: v[0, 0] for n, v in values.items()}),
         **flat("grad_", {n: g[0, 0] for n, g in grads.items()}))


# ---------------------------------------------------------------------------------------------
class _LogitNormalDistribution(distributions.ContinuousDistribution, distributions.UnivariateDistribution):
    """Shim for the LogitNormal the README uses but the reference commented out
    (standard_variables.py:201-213): same pattern as LogNormalDistribution (distributions.py:493-507)
    with torch's SigmoidTransform; no analytic entropy."""
    def __init__(self):
        super().__init__()
        self.torchdist = lambda loc, scale: torch.distributions.TransformedDistribution(
            torch.distributions.Normal(loc, scale), [torch.distributions.transforms.SigmoidTransform()])
        self.required_parameters = {"loc", "scale"}
        self.has_differentiable_samples = True
        self.is_finite = False
        self.is_discrete = False
        self.has_analytic_entropy = False
        self.has_analytic_mean = False
        self.has_analytic_var = False


class _LogitNormalVariable(VariableConstructor):
    def __init__(self, loc, scale, name, learnable=False, is_observed=False):
        self._type = "Logit Normal"
        ranges = {"loc": geometric_ranges.UnboundedRange(), "scale": geometric_ranges.RightHalfLine(0.)}
        super().__init__(name, loc=loc, scale=scale, learnable=learnable, ranges=ranges, is_observed=is_observed)
        self.distribution = _LogitNormalDistribution()


def ar1(seed, T, S, tag="ar1_readme"):
    """README.md:22-75 model, y0 named 'y0'."""
    rng = np.random.RandomState(seed)
    driving, measure, btrue = 1.0, 0.3, 0.7
    xs = [rng.randn() * driving]
    for t in range(1, T):
        xs.append(btrue * xs[-1] + driving * rng.randn())
    ydata = np.array(xs) + measure * rng.randn(T)
    x0 = NormalVariable(0., driving, "x0")
    y0 = NormalVariable(x0, measure, "y0")
    b = _LogitNormalVariable(0.5, 1., "b")
    x, y = [x0], [y0]
    for t in range(1, T):
        x.append(NormalVariable(b * x[t - 1], driving, "x%d" % t))
        y.append(NormalVariable(x[t], measure, "y%d" % t))
    model = ProbabilisticModel(x + y)
    for t, yt in enumerate(y):
        yt.observe(float(ydata[t]))
    Qb = _LogitNormalVariable(0.5, 0.5, "b", learnable=True)
    logit_b_post = DeterministicVariable(0., "logit_b_post", learnable=True)
    Qx = [NormalVariable(0., 1., "x0", learnable=True)]
    Qx_mean = [DeterministicVariable(0., "x0_mean", learnable=True)]
    for t in range(1, T):
        Qx_mean.append(DeterministicVariable(0.1 * rng.randn(), "x%d_mean" % t, learnable=True))
        Qx.append(NormalVariable(BF.sigmoid(logit_b_post) * Qx[t - 1] + Qx_mean[t], 1., "x%d" % t, learnable=True))
    model.set_posterior_model(ProbabilisticModel([Qb] + Qx))
    eps = {"b": torch.tensor(rng.randn(S, 1, 1, 1).astype("float32"))}
    for t in range(T):
        eps["x%d" % t] = torch.tensor(rng.randn(S, 1, 1, 1).astype("float32"))
    inject([Qb] + Qx, eps, transform={"b": torch.sigmoid})
    loss, grads, values = loss_and_grads(model, S)
    save(tag, y=ydata.astype("float32"), loss=loss, measure_noise=measure,
         **flat("eps_", {n: e.numpy().reshape(S) for n, e in eps.items()}),
         **flat("param_", {n: v.reshape(()) for n, v in values.items()}),
         **flat("grad_", {n: g.reshape(()) for n, g in grads.items()}))


def svgd(seed, n, d, tag="svgd_small"):
    """SteinVariationalGradientDescent.correct_gradient (inference.py:301-324) on n particles of dim d
    with arbitrary incoming gradients."""
    rng = np.random.RandomState(seed)
    theta = rng.randn(n, d).astype("float32")
    grad = rng.randn(n, d).astype("float32")
    particles = [ProbabilisticModel([RootVariable(theta[i].astype("float64"), name="weights", learnable=True)])
                 for i in range(n)]
    for i, p in enumerate(particles):
        for v in p.flatten():
            v.value.grad = torch.tensor(grad[i]).reshape(v.value.shape).clone()
    m = inference.SteinVariationalGradientDescent()
    m.correct_gradient(None, particles, None, 1)
    out = np.stack([list(p.flatten())[0].value.grad.detach().numpy().reshape(d) for p in particles])
    save(tag, theta=theta, grad=grad, out=out, bandwidth=float(m.bandwidth))


if __name__ == "__main__":
    torch.manual_seed(0)
    bnn(1, B=12, P=20, H=7, C=4, S=6, tag="bnn_small")
    bnn(2, B=9, P=16, H=5, C=3, S=5, q_sigma=0.3, q_mu_scale=0.5, tag="bnn_small_wide")
    logreg(3, B=40, F=8, S=16, tied=True, tag="logreg_tied")
    logreg(4, B=40, F=8, S=16, tied=False, tag="logreg_declared_prior")
    softmax_reg(5, B=24, F=6, C=3, S=8, tag="softmax_reg")
    ar1(6, T=20, S=32, tag="ar1_readme")
    svgd(7, n=7, d=5, tag="svgd_small")
