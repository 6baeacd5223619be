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
sys.path.insert(0, os.path.dirname(HERE))
import model_zoo as zoo


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


NS = zoo.namespace("brancher")          # the model builders are shared with the brancher_b200 parity tests
NS.LogitNormalVariable = _LogitNormalVariable


def bnn(seed, B, P, H, C, S, q_sigma=0.01, q_mu_scale=0.0, tag="bnn_small", activation="tanh"):
    model, Q, d = zoo.bnn(NS, seed, B, P, H, C, q_sigma, q_mu_scale, activation)
    eps = {n: torch.tensor(d["rng"].randn(S, 1, *s).astype("float32")) for n, s in d["shapes"].items()}
    inject(Q, eps)
    loss, grads, values = loss_and_grads(model, S)
    save(tag, X=d["X"], y=d["y"], loss=loss,
         **flat("eps_", {n: e.numpy()[:, 0] for n, e in eps.items()}),
         **flat("param_", {n: v[0, 0] for n, v in values.items()}),
         **flat("grad_", {n: g[0, 0] for n, g in grads.items()}))


def logreg(seed, B, F, S, tied, tag):
    model, Q, d = zoo.logreg(NS, seed, B, F, tied)
    eps = {"weights": torch.tensor(d["rng"].randn(S, 1, 1, F).astype("float32"))}
    inject(Q, eps)
    loss, grads, values = loss_and_grads(model, S)
    save(tag, X=d["X"], y=d["y"], loss=loss, tied=int(tied),
         prior_loc=np.zeros((1, F), "float32"), prior_scale=0.5 * np.ones((1, F), "float32"),
         eps_weights=eps["weights"].numpy()[:, 0],
         **flat("param_", {n: v[0, 0] for n, v in values.items()}),
         **flat("grad_", {n: g[0, 0] for n, g in grads.items()}))


def softmax_reg(seed, B, F, C, S, tag):
    model, Q, d = zoo.softmax_reg(NS, seed, B, F, C)
    eps = {"weights": torch.tensor(d["rng"].randn(S, 1, C, F).astype("float32"))}
    inject(Q, eps)
    loss, grads, values = loss_and_grads(model, S)
    save(tag, X=d["X"], y=d["y"], loss=loss, eps_weights=eps["weights"].numpy()[:, 0],
         **flat("param_", {n: v[0, 0] for n, v in values.items()}),
         **flat("grad_", {n: g[0, 0] for n, g in grads.items()}))


def ar1(seed, T, S, tag="ar1_readme"):
    model, Q, d = zoo.ar1(NS, seed, T)
    eps = {"b": torch.tensor(d["rng"].randn(S, 1, 1, 1).astype("float32"))}
    for t in range(T):
        eps["x%d" % t] = torch.tensor(d["rng"].randn(S, 1, 1, 1).astype("float32"))
    inject(Q, eps, transform={"b": torch.sigmoid})
    loss, grads, values = loss_and_grads(model, S)
    save(tag, y=d["y"], loss=loss, measure_noise=d["measure_noise"],
         **flat("eps_", {n: e.numpy().reshape(S) for n, e in eps.items()}),
         **flat("param_", {n: v.reshape(()) for n, v in values.items()}),
         **flat("grad_", {n: g.reshape(()) for n, g in grads.items()}))


def scalar_model(builder, tag, S, transforms, **kw):
    """Scalar-DAG family (K1): every q variable is a scalar; noise injected per variable by name."""
    model, Q, d = builder(NS, **kw)
    eps = {q.name: torch.tensor(d["rng"].randn(S, 1, 1, 1).astype("float32")) for q in Q}
    inject(Q, eps, transform={n: t for n, t in transforms.items()})
    loss, grads, values = loss_and_grads(model, S)
    extra = {k: v for k, v in d.items() if k != "rng"}
    save(tag, loss=loss, **extra,
         **flat("eps_", {n: e.numpy().reshape(S) for n, e in eps.items()}),
         **flat("param_", {n: v.reshape(()) for n, v in values.items()}),
         **flat("grad_", {n: g.reshape(()) for n, g in grads.items()}))


def vae_nets(d):
    """numpy view of the torch modules of zoo.vae_modules in the oracle's / C-ABI's layout."""
    enc, dec = d["enc"], d["dec"]
    n = lambda t: t.detach().cpu().numpy().copy()
    e = {"W": [n(l.weight) for l in enc.hidden], "b": [n(l.bias) for l in enc.hidden], "W_mean": n(enc.l_mean.weight),
         "b_mean": n(enc.l_mean.bias), "W_sd": n(enc.l_sd.weight), "b_sd": n(enc.l_sd.bias)}
    dd = {"W": [n(l.weight) for l in dec.hidden], "b": [n(l.bias) for l in dec.hidden], "W_out": n(dec.l_out.weight),
          "b_out": n(dec.l_out.bias)}
    return e, dd


def vae_grads(d):
    enc, dec = d["enc"], d["dec"]
    g = lambda t: t.grad.detach().cpu().numpy().copy()
    out = {}
    for side, net, heads in (("enc", enc, (("W_mean", "l_mean"), ("W_sd", "l_sd"))), ("dec", dec, (("W_out", "l_out"),))):
        for i, l in enumerate(net.hidden):
            out["%s.W.%d" % (side, i)] = g(l.weight)
            out["%s.b.%d" % (side, i)] = g(l.bias)
        for key, attr in heads:
            out["%s.%s" % (side, key)] = g(getattr(net, attr).weight)
            out["%s.b%s" % (side, key[1:])] = g(getattr(net, attr).bias)
    return out


def vae(seed, B, D, L, h_enc, h_dec, S, tag):
    model, Q, d = zoo.vae(NS, seed, B, D, L, h_enc, h_dec)
    eps = torch.tensor(d["rng"].randn(S, B, L).astype("float32"))
    inject(Q, {"z": eps})
    loss = inference.ReverseKL().compute_loss(model, model.posterior_model, None, S)
    loss.backward()
    e, dd = vae_nets(d)
    flatnet = {}
    for side, net in (("enc", e), ("dec", dd)):
        for k, v in net.items():
            if isinstance(v, list):
                for i, a in enumerate(v):
                    flatnet["param_%s.%s.%d" % (side, k, i)] = a
            else:
                flatnet["param_%s.%s" % (side, k)] = v
    save(tag, X=d["X"], loss=float(loss.detach()), eps_z=eps.numpy(), **flatnet, **flat("grad_", vae_grads(d)))


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


def svgd_model(seed, B, F, C, n, tag="svgd_softmax"):
    """One full SVGD gradient evaluation through the reference: compute_loss (inference.py:292-299), backward,
    correct_gradient (:301-315) on the softmax-regression particles of the playground."""
    model, particles, d = zoo.svgd_softmax(NS, seed, B, F, C, n)
    m = inference.SteinVariationalGradientDescent()
    model.update_observed_submodel()
    loss = m.compute_loss(model, particles, None, 1)
    loss.backward()
    raw = np.stack([list(p.flatten())[0].value.grad.detach().numpy().reshape(C, F).copy() for p in particles])
    m.correct_gradient(model, particles, None, 1)
    out = np.stack([list(p.flatten())[0].value.grad.detach().numpy().reshape(C, F).copy() for p in particles])
    save(tag, X=d["X"], y=d["y"], theta=d["theta"], loss=float(loss.detach().sum()), raw_grad=raw, out=out,
         bandwidth=float(m.bandwidth), prior_loc=np.zeros((C, F), "float32"), prior_scale=10 * np.ones((C, F), "float32"))


def wvgd(seed, B, F, C, n, S, tag="wvgd_softmax", tied=True):
    """One WassersteinVariationalGradientDescent.compute_loss + backward (inference.py:203-229) through the reference:
    per-sampler truncated ELBOs + importance-weighted particle loss.  Two noise draws per sampler (ELBO, particle loss),
    injected in call order; the generator asserts no sampler needed a rejection re-draw."""
    model, particles, samplers, d = zoo.wvgd_softmax(NS, seed, B, F, C, n)
    rng = d["rng"]
    eps = rng.randn(n, 2, S, C, F).astype("float32")
    used = [[] for _ in range(n)]
    for k, smp in enumerate(samplers):
        v = [v for v in smp.flatten() if v.name == "weights"][0]

        def f(differentiable, _k=k, **p):
            e = torch.tensor(eps[_k, len(used[_k])]).reshape(S, 1, C, F)
            used[_k].append(1)
            return p["loc"] + e * p["scale"]
        v.distribution._get_sample = f
    m = inference.WassersteinVariationalGradientDescent(variational_samplers=samplers, particles=particles, biased=False)
    model.update_observed_submodel()
    loss = m.compute_loss(model, particles, m.sampler_model, S)
    loss.backward()
    assert all(len(u) == 2 for u in used), "a sampler re-drew (no accepted sample): pick another seed %s" % [len(u) for u in used]
    by = lambda smp, name: [v for v in smp.flatten() if v.name == name][0]
    g_loc = np.stack([by(s_, "weights_loc").link.parameter.grad.numpy().reshape(C, F) for s_ in samplers])
    g_rho = np.stack([by(s_, "weights_scale").link.parameter.grad.numpy().reshape(()) for s_ in samplers])
    rho = np.stack([by(s_, "weights_scale").link.parameter.detach().numpy().reshape(()) for s_ in samplers])
    g_theta = np.stack([by(p_, "weights").value.grad.numpy().reshape(C, F) for p_ in particles])
    save(tag, X=d["X"], y=d["y"], theta=d["theta"], loc=d["loc"], rho=rho, eps_elbo=eps[:, 0], eps_particle=eps[:, 1],
         loss=float(loss.detach()), grad_loc=g_loc, grad_rho=g_rho, grad_theta=g_theta)


def wvgd_post(seed, B, F, C, n, S, tag):
    """WassersteinVariationalGradientDescent.post_process (inference.py:234-247) through the reference: ensemble weights =
    softmax over samplers of log sum_s exp(log p(z_ks, data) - log q_k(z_ks)) over the accepted draws of each truncated
    sampler; the per-sampler log normalisers are recomputed with the same injected noise (variables.py:821-841)."""
    model, particles, samplers, d = zoo.wvgd_softmax(NS, seed, B, F, C, n)
    rng = d["rng"]
    eps = rng.randn(n, S, C, F).astype("float32")
    used = [0] * n
    for k, smp in enumerate(samplers):
        v = [v for v in smp.flatten() if v.name == "weights"][0]

        def f(differentiable, _k=k, **p):
            used[_k] += 1
            return p["loc"] + torch.tensor(eps[_k]).reshape(S, 1, C, F) * p["scale"]
        v.distribution._get_sample = f
    m = inference.WassersteinVariationalGradientDescent(variational_samplers=samplers, particles=particles, biased=False,
                                                        number_post_samples=S)
    model.update_observed_submodel()
    m.post_process(model)
    assert used == [1] * n, "a sampler re-drew (no accepted sample): pick another seed %s" % used
    logZ = []
    for smp in m.sampler_model:
        _, lz = model.get_importance_weights(q_samples=smp._get_sample(S, max_itr=1), q_model=smp, for_gradient=False,
                                             give_normalization=True)
        logZ.append(lz)
    by = lambda smp, name: [v for v in smp.flatten() if v.name == name][0]
    rho = np.stack([by(s_, "weights_scale").link.parameter.detach().numpy().reshape(()) for s_ in samplers])
    save(tag, X=d["X"], y=d["y"], theta=d["theta"], loc=d["loc"], rho=rho, eps_post=eps, weights=np.asarray(m.weights),
         logZ=np.asarray(logZ, dtype=np.float64))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "wvgd":
        wvgd(21, B=30, F=4, C=3, n=3, S=20, tag="wvgd_softmax")
        wvgd(22, B=16, F=5, C=2, n=4, S=24, tag="wvgd_softmax4")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bnn_act":
        bnn(31, B=11, P=14, H=6, C=3, S=5, q_sigma=0.3, q_mu_scale=0.5, tag="bnn_relu", activation="relu")
        bnn(32, B=10, P=12, H=5, C=4, S=6, q_sigma=0.3, q_mu_scale=0.5, tag="bnn_sigmoid", activation="sigmoid")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "wvgd_post":
        wvgd_post(21, B=30, F=4, C=3, n=3, S=64, tag="wvgd_post")
        wvgd_post(22, B=16, F=5, C=2, n=4, S=48, tag="wvgd_post4")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "robust":
        scalar_model(zoo.robust_regression, "robust_regression", S=16, transforms={"nu": torch.exp}, seed=14, n=40)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "op_zoo":
        scalar_model(zoo.op_zoo, "op_zoo", S=12, transforms={"c": torch.exp}, seed=16, n=24)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "scalar_logistic":
        scalar_model(zoo.scalar_logistic, "scalar_logistic", S=20, transforms={}, seed=15, n=30)
        sys.exit(0)
    torch.manual_seed(0)
    bnn(1, B=12, P=20, H=7, C=4, S=6, tag="bnn_small")
    bnn(2, B=9, P=16, H=5, C=3, S=5, q_sigma=0.3, q_mu_scale=0.5, tag="bnn_small_wide")
    logreg(3, B=40, F=8, S=16, tied=True, tag="logreg_tied")
    logreg(4, B=40, F=8, S=16, tied=False, tag="logreg_declared_prior")
    softmax_reg(5, B=24, F=6, C=3, S=8, tag="softmax_reg")
    ar1(6, T=20, S=32, tag="ar1_readme")
    svgd(7, n=7, d=5, tag="svgd_small")
    scalar_model(zoo.lognormal_normal, "lognormal_normal", S=24, transforms={"nu": torch.exp}, seed=10, N=20)
    scalar_model(zoo.multivariate_regression, "multivariate_regression", S=16, transforms={"nu": torch.exp}, seed=11, n=50)
    scalar_model(zoo.robust_regression, "robust_regression", S=16, transforms={"nu": torch.exp}, seed=14, n=40)
    scalar_model(zoo.scalar_logistic, "scalar_logistic", S=20, transforms={}, seed=15, n=30)
    scalar_model(zoo.op_zoo, "op_zoo", S=12, transforms={"c": torch.exp}, seed=16, n=24)
    svgd_model(8, B=30, F=5, C=3, n=6, tag="svgd_softmax")
    vae(12, B=6, D=12, L=2, h_enc=(5, 7), h_dec=(7, 5), S=3, tag="vae_small")
    vae(13, B=10, D=20, L=3, h_enc=(9,), h_dec=(6, 8, 5), S=4, tag="vae_deep")
    wvgd(21, B=30, F=4, C=3, n=3, S=20, tag="wvgd_softmax")
    wvgd(22, B=16, F=5, C=2, n=4, S=24, tag="wvgd_softmax4")
    wvgd_post(21, B=30, F=4, C=3, n=3, S=64, tag="wvgd_post")
    wvgd_post(22, B=16, F=5, C=2, n=4, S=48, tag="wvgd_post4")
    bnn(31, B=11, P=14, H=6, C=3, S=5, q_sigma=0.3, q_mu_scale=0.5, tag="bnn_relu", activation="relu")
    bnn(32, B=10, P=12, H=5, C=4, S=6, q_sigma=0.3, q_mu_scale=0.5, tag="bnn_sigmoid", activation="sigmoid")
