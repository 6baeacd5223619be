"""Graph matcher: lowers a (joint model, posterior model) pair to a fused CUDA kernel family.

The reference evaluates the ELBO by walking the model graph in Python, one torch op at a time
(variables.py:843-870, 486-570).  Here the pair is inspected ONCE: its expression trees are
pattern-matched against the kernel families of include/brancher_cuda.h and a `Plan` is cached on the
joint model.  `Plan.elbo(S, empirical_samples)` then runs one fused forward+backward evaluation and
returns a 0-d tensor wired into autograd (so `loss.backward()` fills `.grad` of every learnable
`ParameterModule.parameter`, the contract of optimizers.py:34-51).

Families
  LinearPlan  K2  observed Binomial(1)/Bernoulli/Categorical with logits = BF.matmul(W, x), W mean-field Normal
  BNNPlan     K3  observed Categorical with logits = BF.matmul(W2, BF.tanh(BF.matmul(W1, x) + b1)) + b2
Unrecognised graphs raise `UnsupportedModelError` -- there is no eager / CPU fallback for the ELBO.
"""
import contextlib

import numpy as np
import torch

from brancher_b200 import config, distributed
from brancher_b200.variables import (RootVariable, RandomVariable, VarRef, Const, Call, evaluate)


class UnsupportedModelError(NotImplementedError):
    pass


# ---------------------------------------------------------------------------------------------------
# noise injection (parity testing): {q variable name: tensor (S, *event)} of standard normals
# ---------------------------------------------------------------------------------------------------
_INJECTED = None


@contextlib.contextmanager
def inject_noise(eps_by_name):
    """Within the context every fused evaluation uses the given standard-normal draws instead of Philox
    (the counterpart of overriding `Qvar.distribution._get_sample` in the reference)."""
    global _INJECTED
    prev, _INJECTED = _INJECTED, dict(eps_by_name)
    try:
        yield
    finally:
        _INJECTED = prev


# ---------------------------------------------------------------------------------------------------
# expression pattern helpers
# ---------------------------------------------------------------------------------------------------
def _is_call(e, name, nargs=None):
    return isinstance(e, Call) and e.name == name and (nargs is None or len(e.args) == nargs) and not e.kwargs


def _const_value(e):
    if isinstance(e, Const) and isinstance(e.value, (int, float)):
        return float(e.value)
    return None


def _strip_zero_add(e):
    """x + 0.0 / 0.0 + x -> x   (RightHalfLine(0.).forward_transform writes `0. + softplus(x)`)."""
    if _is_call(e, "add", 2):
        a, b = e.args
        if _const_value(b) == 0.0:
            return a
        if _const_value(a) == 0.0:
            return b
    return e


def _as_root(e):
    return e.var if isinstance(e, VarRef) and isinstance(e.var, RootVariable) else None


def _softplus_root(e):
    e = _strip_zero_add(e)
    if _is_call(e, "softplus", 1):
        return _as_root(e.args[0])
    return None


def _commutative_add(e):
    if _is_call(e, "add", 2):
        a, b = e.args
        return [(a, b), (b, a)]
    return []


def _var(e):
    return e.var if isinstance(e, VarRef) else None


# ---------------------------------------------------------------------------------------------------
class MeanFieldSpec:
    """One latent: q = Normal(root_loc, softplus(root_scale)) (learnable roots) and p = Normal prior of the
    same name, tied by root-name collision or declared."""

    def __init__(self, name, q_var, p_var, q_names):
        self.name = name
        lq = q_var.partial_links
        if q_var.distribution.kind != "normal" or p_var.distribution.kind != "normal":
            raise UnsupportedModelError("latent %r: only Normal q / Normal prior are lowered to the mean-field kernels" % name)
        self.loc_root = _as_root(lq["loc"].expr)
        self.scale_root = _softplus_root(lq["scale"].expr)
        if self.loc_root is None or self.scale_root is None:
            raise UnsupportedModelError("latent %r: q's loc/scale must be (learnable) roots" % name)
        self.shape = tuple(self.loc_root._value.shape[2:])
        lp = p_var.partial_links
        p_loc_root, p_scale_sp = _as_root(lp["loc"].expr), _softplus_root(lp["scale"].expr)
        p_scale_raw = _as_root(lp["scale"].expr)
        p_scale_root = p_scale_sp or p_scale_raw
        if p_loc_root is None or p_scale_root is None:
            raise UnsupportedModelError("latent %r: the prior's loc/scale must be constants" % name)
        tied_loc = p_loc_root.name in q_names
        tied_scale = p_scale_root.name in q_names
        if tied_loc != tied_scale:
            raise UnsupportedModelError("latent %r: prior loc/scale roots are tied to q inconsistently" % name)
        if tied_loc and not (p_loc_root.name == self.loc_root.name and p_scale_root.name == self.scale_root.name
                             and p_scale_sp is not None):
            raise UnsupportedModelError("latent %r: unexpected root-name collision pattern" % name)
        self.tied = tied_loc
        self.prior_loc = self.prior_scale = None
        if not self.tied:
            self.prior_loc = p_loc_root.value.detach()
            sc = p_scale_root.value.detach()
            self.prior_scale = torch.nn.functional.softplus(sc) if p_scale_sp is not None else sc

    def parameters(self):
        return [self.loc_root.value, self.scale_root.value]

    def make(self, cu, var_id, eps, sinks=(None, None)):
        mu, rho = self.loc_root.value, self.scale_root.value
        if not self.tied:
            pl = self.prior_loc.expand(mu.shape).reshape(-1).contiguous()
            ps = self.prior_scale.expand(mu.shape).reshape(-1).contiguous()
        else:
            pl = ps = None
        return cu.MeanFieldVar(mu, rho, var_id=var_id, prior_loc=pl, prior_scale=ps, eps=eps, dmu=sinks[0], drho=sinks[1])


class _FusedELBO(torch.autograd.Function):
    """autograd node of one fused evaluation: forward runs the kernels (loss and all gradients at once),
    backward hands the stored gradients out."""

    @staticmethod
    def forward(ctx, runner, *params):
        loss, grads = runner()
        ctx.grads = grads
        return (-loss).to(torch.float32).reshape(())          # the reference's estimator returns +ELBO

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple((-g) * gr if gr is not None else None for gr in ctx.grads)


class Plan:
    family = "?"

    def __init__(self, joint, posterior):
        self.joint, self.posterior = joint, posterior
        self.latents = []        # MeanFieldSpec in kernel order

    def _eps(self, S_local, s0):
        out = []
        for spec in self.latents:
            if _INJECTED is None:
                out.append(None)
                continue
            if spec.name not in _INJECTED:
                raise KeyError("inject_noise: no noise given for q variable %r" % spec.name)
            e = torch.as_tensor(_INJECTED[spec.name], dtype=torch.float32, device=config.device)
            e = e.reshape(e.shape[0], -1)
            out.append(e[s0:s0 + S_local].contiguous())
        return out

    def elbo(self, number_samples, empirical_samples):
        if config.device.type != "cuda":
            raise RuntimeError("brancher_b200 evaluates the ELBO only on CUDA devices (no CPU fallback); "
                               "config.device is %s" % config.device)
        from brancher_b200 import _cuda as cu
        cu.lib()
        s0, S_local = distributed.shard(number_samples)
        r = cu.sample_range(number_samples, s0=s0, s_local=S_local, seed=config.seed, offset=config.next_offset())
        params = [p for spec in self.latents for p in spec.parameters()]

        def runner():
            _, sinks = cu.flat_grad_views([int(np.prod(spec.shape)) for spec in self.latents], config.device)
            mvars = [spec.make(cu, i, eps, sk) for i, (spec, eps, sk) in enumerate(zip(self.latents, self._eps(S_local, s0), sinks))]
            loss = self._launch(cu, mvars, r, empirical_samples)
            grads = []
            for spec, v in zip(self.latents, mvars):
                grads += [v.dmu.reshape(spec.loc_root.value.shape), v.drho.reshape(spec.scale_root.value.shape)]
            loss, grads = distributed.all_reduce_partials(loss, grads)
            return loss, [g if p.requires_grad else None for g, p in zip(grads, params)]

        return _FusedELBO.apply(runner, *params)

    def _launch(self, cu, mvars, r, empirical_samples):
        raise NotImplementedError


def _data_matrix(t, what):
    """(1, B, P, 1) / (1, B, P) observed tensor -> [B, P] fp32 contiguous."""
    if not torch.is_tensor(t) or t.shape[0] != 1:
        raise UnsupportedModelError("%s: expected an observed tensor with a singleton sample axis" % what)
    return t.reshape(t.shape[1], -1).to(torch.float32).contiguous()


class LinearPlan(Plan):
    family = "linear (K2)"

    def __init__(self, joint, posterior, k, w_spec, x_var, likelihood, C):
        super().__init__(joint, posterior)
        self.k, self.x_var, self.likelihood, self.C = k, x_var, likelihood, C
        self.latents = [w_spec]

    def _launch(self, cu, mvars, r, empirical):
        X = _data_matrix(empirical[self.x_var], "x")
        yv = empirical[self.k].reshape(-1)
        if self.likelihood == cu.BERNOULLI:
            y = yv.to(torch.float32).contiguous()
        else:
            y = yv.to(torch.int32).contiguous()
        return cu.linear_elbo_fwd_bwd(X, y, self.likelihood, mvars[0], self.C, r)


class BNNPlan(Plan):
    family = "bnn (K3)"

    def __init__(self, joint, posterior, k, specs, x_var):
        super().__init__(joint, posterior)
        self.k, self.x_var = k, x_var
        self.latents = specs          # weights1, b1, weights2, b2

    def _launch(self, cu, mvars, r, empirical):
        X = _data_matrix(empirical[self.x_var], "x")
        y = empirical[self.k].reshape(-1).to(torch.int32).contiguous()
        return cu.bnn_elbo_fwd_bwd(X, y, mvars, r)


# ---------------------------------------------------------------------------------------------------
def _latent_random_variables(model):
    return [v for v in model._flatten() if isinstance(v, RandomVariable) and not v.is_observed
            and v.distribution.kind not in ("deterministic", "empirical")]


def _observed_likelihood_nodes(model):
    return [v for v in model._flatten() if isinstance(v, RandomVariable) and v.is_observed
            and v.distribution.kind not in ("deterministic", "empirical")]


def _is_data(var):
    return var is not None and var.is_observed and (isinstance(var, RootVariable) or
                                                    var.distribution.kind in ("deterministic", "empirical"))


def lower(joint, posterior):
    from brancher_b200 import _cuda as cu
    q_names = {v.name for v in posterior._flatten()}
    q_by_name = {v.name: v for v in posterior._flatten()}
    latents = _latent_random_variables(joint)
    liks = _observed_likelihood_nodes(joint)
    for v in latents:
        if v.name not in q_by_name:
            raise UnsupportedModelError("latent variable %r has no counterpart in the posterior model" % v.name)
    if len(liks) != 1:
        raise UnsupportedModelError("fused families need exactly one observed likelihood node, found %d" % len(liks))
    k = liks[0]
    kind = k.distribution.kind
    if "logits" not in k.partial_links:
        raise UnsupportedModelError("likelihood %r: only logits= parametrisations are lowered" % k.name)
    if kind == "binomial":
        tc = _as_root(k.partial_links["total_count"].expr)
        if tc is None or float(tc.value.reshape(-1)[0]) != 1.0 or tc.value.numel() != 1:
            raise UnsupportedModelError("Binomial likelihood is lowered only for total_count=1")
    if kind not in ("binomial", "bernoulli", "categorical"):
        raise UnsupportedModelError("likelihood kind %r is not lowered" % kind)
    logits = k.partial_links["logits"].expr
    spec = lambda var: MeanFieldSpec(var.name, q_by_name[var.name], var, q_names)

    # K2: logits = matmul(W, x)
    if _is_call(logits, "matmul", 2):
        W, x = _var(logits.args[0]), _var(logits.args[1])
        if W in latents and _is_data(x) and len(latents) == 1:
            ws = spec(W)
            if len(ws.shape) != 2:
                raise UnsupportedModelError("linear family: weights must be a [C, F] matrix")
            C = ws.shape[0]
            if kind != "categorical" and C != 1:
                raise UnsupportedModelError("Binomial/Bernoulli logistic regression needs weights of shape [1, F]")
            return LinearPlan(joint, posterior, k, ws, x, cu.CATEGORICAL if kind == "categorical" else cu.BERNOULLI, C)

    # K3: logits = matmul(W2, tanh(matmul(W1, x) + b1)) + b2
    if kind == "categorical":
        for outer, b2e in _commutative_add(logits):
            b2 = _var(b2e)
            if not (_is_call(outer, "matmul", 2) and b2 in latents):
                continue
            W2, hid = _var(outer.args[0]), outer.args[1]
            if not (W2 in latents and _is_call(hid, "tanh", 1)):
                continue
            for inner, b1e in _commutative_add(hid.args[0]):
                b1 = _var(b1e)
                if not (_is_call(inner, "matmul", 2) and b1 in latents):
                    continue
                W1, x = _var(inner.args[0]), _var(inner.args[1])
                if W1 in latents and _is_data(x) and len({W1, b1, W2, b2}) == 4 and len(latents) == 4:
                    return BNNPlan(joint, posterior, k, [spec(W1), spec(b1), spec(W2), spec(b2)], x)
    raise UnsupportedModelError("model graph is not recognised by any fused kernel family "
                                "(linear K2, bnn K3); likelihood link: %s" % k.partial_links["logits"].string)


# ---------------------------------------------------------------------------------------------------
# particle ensembles (SVGD, K4)
# ---------------------------------------------------------------------------------------------------
class _ParticleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, *params):
        loss, G = runner()
        ctx.G, ctx.shapes = G, [p.shape for p in params]
        return loss.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, g):
        return (None,) + tuple(g * ctx.G[i].reshape(sh) for i, sh in enumerate(ctx.shapes))


class ParticlePlan:
    """SVGD over (multi-class) logistic-regression particles: the joint model is the linear family (K2's graph), every
    particle is a ProbabilisticModel holding ONE learnable RootVariable named like the weights latent
    (development_playgrounds/SVGD_logistic_regression.py:34-48)."""
    family = "particles (K4)"

    def __init__(self, joint, particles, k, W, x_var, likelihood, C):
        self.joint, self.particles, self.k, self.x_var, self.likelihood, self.C = joint, particles, k, x_var, likelihood, C
        lp = W.partial_links
        loc, sc_sp = _as_root(lp["loc"].expr), _softplus_root(lp["scale"].expr)
        sc = sc_sp or _as_root(lp["scale"].expr)
        if loc is None or sc is None:
            raise UnsupportedModelError("particle family: the prior's loc/scale must be constants")
        self.shape = tuple(loc._value.shape[2:])
        full = lambda t: t.detach().to(torch.float32).expand((1, 1) + self.shape).reshape(-1).contiguous()
        self.prior_loc = full(loc.value)
        self.prior_scale = full(torch.nn.functional.softplus(sc.value) if sc_sp is not None else sc.value)
        self.roots = []
        for p in particles:
            roots = [v for v in p._flatten() if isinstance(v, RootVariable) and v.name == W.name]
            if len(roots) != 1 or tuple(roots[0]._value.shape[2:]) != self.shape:
                raise UnsupportedModelError("every particle must hold one root named %r of shape %s" % (W.name, self.shape))
            self.roots.append(roots[0])

    def parameters(self):
        return [r.value for r in self.roots]

    def stacked(self):
        return torch.stack([p.detach().reshape(-1) for p in self.parameters()]).contiguous()

    def loss(self, empirical):
        if config.device.type != "cuda":
            raise RuntimeError("brancher_b200 evaluates particle losses only on CUDA devices (no CPU fallback)")
        from brancher_b200 import _cuda as cu
        cu.lib()

        def runner():
            X = _data_matrix(empirical[self.x_var], "x")
            yv = empirical[self.k].reshape(-1)
            y = yv.to(torch.float32).contiguous() if self.likelihood == cu.BERNOULLI else yv.to(torch.int32).contiguous()
            return cu.linear_particles_loss_grad(X, y, self.likelihood, self.stacked(), self.C, self.prior_loc, self.prior_scale)

        return _ParticleLoss.apply(runner, *self.parameters())


def lower_particles(joint, particles):
    from brancher_b200 import _cuda as cu
    latents = _latent_random_variables(joint)
    liks = _observed_likelihood_nodes(joint)
    if len(liks) != 1 or len(latents) != 1:
        raise UnsupportedModelError("particle family needs one observed likelihood node and one latent weight matrix")
    k, W = liks[0], latents[0]
    kind = k.distribution.kind
    if kind not in ("binomial", "bernoulli", "categorical") or "logits" not in k.partial_links:
        raise UnsupportedModelError("particle family: likelihood %r is not lowered" % kind)
    if kind == "binomial":
        tc = _as_root(k.partial_links["total_count"].expr)
        if tc is None or tc.value.numel() != 1 or float(tc.value.reshape(-1)[0]) != 1.0:
            raise UnsupportedModelError("Binomial likelihood is lowered only for total_count=1")
    logits = k.partial_links["logits"].expr
    if not (_is_call(logits, "matmul", 2) and _var(logits.args[0]) is W and _is_data(_var(logits.args[1]))):
        raise UnsupportedModelError("particle family: logits must be BF.matmul(weights, x)")
    if W.distribution.kind != "normal":
        raise UnsupportedModelError("particle family: only a Normal prior on the weights is lowered")
    shape = tuple(W.partial_links["loc"].expr.var._value.shape[2:]) if _as_root(W.partial_links["loc"].expr) else ()
    if len(shape) != 2:
        raise UnsupportedModelError("particle family: weights must be a [C, F] matrix")
    C = shape[0]
    if kind != "categorical" and C != 1:
        raise UnsupportedModelError("Binomial/Bernoulli logistic regression needs weights of shape [1, F]")
    return ParticlePlan(joint, particles, k, W, _var(logits.args[1]), cu.CATEGORICAL if kind == "categorical" else cu.BERNOULLI, C)


def get_particle_plan(joint, particles):
    key = ("particles",) + tuple(id(p) for p in particles)
    plan = joint._plans.get(key)
    if plan is None:
        plan = lower_particles(joint, list(particles))
        joint._plans[key] = plan
    return plan


def get_plan(joint, posterior):
    key = id(posterior)
    plan = joint._plans.get(key)
    if plan is None or plan.posterior is not posterior:
        plan = lower(joint, posterior)
        joint._plans[key] = plan
    return plan
