"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Vectorised torch-CPU restatement of the reference's Monte-Carlo ELBO
(`ProbabilisticModel.estimate_log_model_evidence`, /root/reference/brancher/variables.py:843-870,
driven by `PathwiseDerivativeEstimator.__call__`, gradient_estimators.py:39-44) for the
model families of BASELINE.json's configs.  Gradients come from torch autograd exactly as
in the reference (`loss.backward()`, inference.py:100).  All element arithmetic is
delegated to the same third-party code the reference calls -- `torch.distributions`
(reference call sites distributions.py:108,122,166,180,310; pin `torch>=1.0.0`,
requirements.txt:1; version in this image 2.11.0).

What is *restated* (instead of executed from the reference) is the graph walk:
  * ancestral sampling of q with reparameterisation `loc + eps*scale`
    (variables.py:527-570, distributions.py:111-124 -> torch Normal.rsample),
  * `scale = softplus(rho)` for learnable scales (geometric_ranges.py:48-57),
  * q->p reassignment BY NAME, so that p's auto-created `<name>_loc/<name>_scale` roots
    take q's values ("collision"/tied mode; utilities.py:282-309, variables.py:367-371),
  * observed nodes summed over the data axis (variables.py:513-514), latent nodes kept
    per-sample, analytic entropies (variables.py:156-162,744-749),
  * the final `.mean()` over the sample axis (gradient_estimators.py:44).
The reference materialises every operand at (S*B, ...) (variables.py:436-449); here the
same contraction is a broadcasted matmul/einsum, which is what lets the oracle run at the
BASELINE shapes.

Every public function returns `(loss, grads)` with `loss = -ELBO` (inference.py:140-144) as a
python float and `grads` a dict keyed by the REFERENCE's parameter names
(`<var>_loc`, `<var>_scale` -- the latter is d loss / d rho, rho = softplus^-1(sigma)).
"""
import math

import numpy as np
import torch
import torch.distributions as D
import torch.nn.functional as F


def _t(x, dtype, requires_grad=False):
    if isinstance(x, torch.Tensor):
        t = x.detach().to("cpu", dtype).clone()
    else:
        t = torch.tensor(np.asarray(x), dtype=dtype)
    if requires_grad:
        t.requires_grad_(True)
    return t


def F_softplus(x):
    return torch.nn.functional.softplus(x)


def softplus_inverse(sigma):
    """rho such that softplus(rho) = sigma, as the reference stores learnable scales
    (geometric_ranges.py:56-57, numpy fp64 then cast to fp32 by utilities.py:241-247)."""
    return np.log(np.exp(np.asarray(sigma, dtype=np.float64)) - 1.0)


class MeanFieldNormal:
    """A set of mean-field Normal q variables (NormalVariable(loc, scale, name, learnable=True),
    standard_variables.py:133-145) with their p-side Normal priors of the same names.

    `params`: {name: (mu, rho)} arrays of the variable's event shape.
    `eps`:    {name: array (S, *event)} standard-normal noise (injected).
    `prior`:  None -> tied/collision mode (p's roots take q's values);
              or {name: (loc, scale)} numeric -> declared prior (roots named distinctly).
    """

    def __init__(self, params, eps, prior=None, dtype=torch.float32):
        self.dtype = dtype
        self.names = sorted(params)
        self.mu = {n: _t(params[n][0], dtype, True) for n in self.names}
        self.rho = {n: _t(params[n][1], dtype, True) for n in self.names}
        self.eps = {n: _t(eps[n], dtype) for n in self.names}
        self.prior = prior
        self.S = next(iter(self.eps.values())).shape[0]

    def sample(self):
        """w_s = mu + softplus(rho) * eps_s (torch Normal.rsample = loc + eps*scale)."""
        self.sigma = {n: F.softplus(self.rho[n]) for n in self.names}
        self.w = {n: self.mu[n].unsqueeze(0) + self.eps[n] * self.sigma[n].unsqueeze(0) for n in self.names}
        return self.w

    def log_prior(self):
        """sum over variables and event dims of Normal(pi_loc, pi_scale).log_prob(w_s) -> [S]."""
        total = 0.0
        for n in self.names:
            if self.prior is None:
                loc, scale = self.mu[n], self.sigma[n]
            else:
                loc = _t(np.broadcast_to(self.prior[n][0], self.mu[n].shape).copy(), self.dtype)
                scale = _t(np.broadcast_to(self.prior[n][1], self.mu[n].shape).copy(), self.dtype)
            lp = D.Normal(loc.unsqueeze(0), scale.unsqueeze(0)).log_prob(self.w[n])
            total = total + lp.reshape(self.S, -1).sum(dim=1)
        return total

    def entropy(self):
        """sum over variables and event dims of Normal(mu, sigma).entropy() -> scalar
        (variables.py:156-162: analytic entropy where torch has it)."""
        total = 0.0
        for n in self.names:
            total = total + D.Normal(self.mu[n], self.sigma[n]).entropy().sum()
        return total

    def grads(self):
        g = {}
        for n in self.names:
            g[n + "_loc"] = self.mu[n].grad.detach().numpy().copy()
            g[n + "_scale"] = self.rho[n].grad.detach().numpy().copy()
        return g


def mean_field_prior_entropy(params, eps, prior=None, dtype=torch.float32):
    """-(mean_s log p(w_s) + H[q]) alone: the K1 elementwise stage of the LogReg/BNN configs."""
    q = MeanFieldNormal(params, eps, prior, dtype)
    q.sample()
    loss = -(q.log_prior().mean() + q.entropy())
    loss.backward()
    return float(loss.detach()), q.grads()


def logreg_elbo(X, y, params, eps, prior=None, dtype=torch.float32, likelihood="binomial",
                row_chunk=None, with_prior=True):
    """Bayesian (multi-class) logistic regression, `minibatch_logistic_regression.py:27-43` shape.

    X [B,F]; weights variable named "weights" with event shape [C,F] (C=1 for the Binomial
    case, minibatch_logistic_regression.py:27-29); logits_sbc = sum_f W_scf X_bf  (BF.matmul(weights, x)).
    likelihood: "binomial"  -> Binomial(total_count=1, logits).log_prob(y)  (C must be 1)
                "bernoulli" -> Bernoulli(logits).log_prob(y)
                "categorical" -> Categorical(logits).log_prob(label)     (distributions.py:294-311)
    Observed node => summed over the data axis b (variables.py:513-514).
    """
    q = MeanFieldNormal(params, eps, prior, dtype)
    w = q.sample()["weights"]                       # [S,C,F]
    Xt = _t(X, dtype)
    B = Xt.shape[0]
    S = q.S
    yt_all = _t(y, dtype if likelihood != "categorical" else torch.int64)
    ll = torch.zeros(S, dtype=dtype)
    step = row_chunk or B
    for r0 in range(0, B, step):
        xb = Xt[r0:r0 + step]
        yb = yt_all[r0:r0 + step]
        logits = torch.einsum("scf,bf->sbc", w, xb)
        if likelihood == "binomial":
            lp = D.Binomial(total_count=1, logits=logits[..., 0]).log_prob(yb.unsqueeze(0))
        elif likelihood == "bernoulli":
            lp = D.Bernoulli(logits=logits[..., 0]).log_prob(yb.unsqueeze(0))
        elif likelihood == "categorical":
            lp = D.Categorical(logits=logits).log_prob(yb.unsqueeze(0))
        else:
            raise ValueError(likelihood)
        ll = ll + lp.sum(dim=1)
    per_sample = ll
    if with_prior:
        per_sample = per_sample + q.log_prior()
    elbo = per_sample.mean()
    if with_prior:
        elbo = elbo + q.entropy()
    loss = -elbo
    loss.backward()
    return float(loss.detach()), q.grads()


def bnn_elbo(X, y, params, eps, prior=None, dtype=torch.float32, sample_chunk=None, with_prior=True, activation="tanh"):
    """One-hidden-layer Bayesian neural network, `development_playgrounds/MNIST_bayesian_neural_network.py:26-57`.

    variables "weights1" [H,P], "b1" [H,1], "weights2" [C,H], "b2" [C,1];
    h = tanh(W1 x + b1) (:38), a = W2 h + b2 (:39), k ~ Categorical(logits=a) (:40) observed
    with integer labels y [B]; X [B,P].
    """
    q = MeanFieldNormal(params, eps, prior, dtype)
    w = q.sample()
    Xt = _t(X, dtype)
    yt = _t(y, torch.int64)
    S = q.S
    step = sample_chunk or S
    lls = []
    for s0 in range(0, S, step):
        sl = slice(s0, s0 + step)
        pre = torch.einsum("shp,bp->sbh", w["weights1"][sl], Xt) + w["b1"][sl, :, 0].unsqueeze(1)
        h = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[activation](pre)     # BF.tanh / BF.relu / BF.sigmoid
        a = torch.einsum("sch,sbh->sbc", w["weights2"][sl], h) + w["b2"][sl, :, 0].unsqueeze(1)
        lp = D.Categorical(logits=a).log_prob(yt.unsqueeze(0))      # [s,B]
        lls.append(lp.sum(dim=1))
    per_sample = torch.cat(lls)
    if with_prior:
        per_sample = per_sample + q.log_prior()
    elbo = per_sample.mean()
    if with_prior:
        elbo = elbo + q.entropy()
    loss = -elbo
    loss.backward()
    return float(loss.detach()), q.grads()


# ----------------------------------------------------------------------------------------------
# Amortised VAE (config C5): /root/reference/examples/VAE_playground.py:30-80
# ----------------------------------------------------------------------------------------------
def vae_elbo(X, enc, dec, eps, sd_offset=0.1, dtype=torch.float32, row_chunk=None):
    """Amortised VAE with ReLU-MLP encoder / decoder, z ~ N(0, I), x ~ Binomial(1, logits) (VAE_playground.py:65-76).

    X [B, D] in {0,1}; eps [S, B, L] injected noise of Qz; enc / dec: dicts of numpy arrays
      enc: "W" list of hidden weights [n_out, n_in] (torch.nn.Linear layout), "b" list of biases, "W_mean","b_mean",
           "W_sd","b_sd"  (mean = l3(h), sd = softplus(l4(h)) + 0.1, VAE_playground.py:44-46)
      dec: "W","b" hidden lists, "W_out","b_out"  (logits = l3(h), :58-62)
    The minibatch is the SAME B rows for every MC sample.  What the reference computes (SURVEY §8 a'):
      x is not observed in p  => no data-axis sum (variables.py:513-514) => mean over S AND B (gradient_estimators.py:44);
      log p(z) = sum_lat N(z;0,1); H[Qz] analytic = sum_lat (1/2 + c + log sd) (variables.py:156-162);
      H[Qx] = log(S): EmpiricalDistribution._get_entropy sees the S-times-tiled dataset's leading axis
      (distributions.py:464-473).
    Returns (loss, grads) with grads keyed "enc.W.0", "enc.b.0", ..., "enc.W_mean", ..., "dec.W_out", "dec.b_out".
    """
    Xt = _t(X, dtype)
    et = _t(eps, dtype)
    S, B, L = et.shape
    P = {}
    for side, net in (("enc", enc), ("dec", dec)):
        for k, v in net.items():
            if isinstance(v, (list, tuple)):
                for i, a in enumerate(v):
                    P["%s.%s.%d" % (side, k, i)] = _t(a, dtype, True)
            else:
                P["%s.%s" % (side, k)] = _t(v, dtype, True)
    n_enc, n_dec = len(enc["W"]), len(dec["W"])
    c = 0.5 * math.log(2 * math.pi)
    step = row_chunk or B
    total = 0.0
    for b0 in range(0, B, step):
        xb, eb = Xt[b0:b0 + step], et[:, b0:b0 + step]
        h = xb
        for i in range(n_enc):
            h = torch.relu(F.linear(h, P["enc.W.%d" % i], P["enc.b.%d" % i]))
        mean = F.linear(h, P["enc.W_mean"], P["enc.b_mean"])
        sd = F.softplus(F.linear(h, P["enc.W_sd"], P["enc.b_sd"])) + sd_offset
        z = mean.unsqueeze(0) + eb * sd.unsqueeze(0)                                   # Normal.rsample: loc + eps*scale
        g = z
        for i in range(n_dec):
            g = torch.relu(F.linear(g, P["dec.W.%d" % i], P["dec.b.%d" % i]))
        logits = F.linear(g, P["dec.W_out"], P["dec.b_out"])                          # [S, b, D]
        ll = D.Binomial(total_count=1, logits=logits).log_prob(xb.unsqueeze(0)).sum(-1)     # [S, b]
        lpz = D.Normal(torch.zeros((), dtype=dtype), torch.ones((), dtype=dtype)).log_prob(z).sum(-1)
        ent = D.Normal(mean, sd).entropy().sum(-1).unsqueeze(0)                        # [1, b]
        part = -(ll + lpz + ent).sum() / (S * B)
        part.backward()
        total += float(part.detach())
    loss = total - math.log(S)
    return loss, {k: v.grad.numpy().copy() for k, v in P.items()}


def vae_relu_margins(X, enc, dec, eps, sd_offset=0.1, row_chunk=512):
    """fp64 forward pass of `vae_elbo`: for every DATA row b the smallest relative distance of a ReLU pre-activation from
    its kink,  min over layers / units / samples of |pre| / (sum_k |w_k a_k| + |bias|).
    A ReLU network's gradient is discontinuous where a pre-activation crosses zero: two correct fp32 evaluations whose
    forward passes differ by rounding pick different branches there (the reference's own fp32 run included), so parity of
    gradients is only defined on rows whose margin exceeds the forward error of the implementations being compared."""
    dt = torch.float64
    Xt, et = _t(X, dt), _t(eps, dt)
    S, B, L = et.shape
    W = lambda a: _t(a, dt)
    out = np.full(B, np.inf)
    with torch.no_grad():
        for b0 in range(0, B, row_chunk):
            xb, eb = Xt[b0:b0 + row_chunk], et[:, b0:b0 + row_chunk]
            m = torch.full((xb.shape[0],), float("inf"), dtype=dt)
            h = xb
            for Wi, bi in zip(enc["W"], enc["b"]):
                Wi, bi = W(Wi), W(bi)
                pre = F.linear(h, Wi, bi)
                sc = F.linear(h.abs(), Wi.abs(), bi.abs())
                m = torch.minimum(m, (pre.abs() / sc).min(dim=1).values)
                h = torch.relu(pre)
            mean = F.linear(h, W(enc["W_mean"]), W(enc["b_mean"]))
            sd = F.softplus(F.linear(h, W(enc["W_sd"]), W(enc["b_sd"]))) + sd_offset
            g = mean.unsqueeze(0) + eb * sd.unsqueeze(0)
            for Wi, bi in zip(dec["W"], dec["b"]):
                Wi, bi = W(Wi), W(bi)
                pre = F.linear(g, Wi, bi)
                sc = F.linear(g.abs(), Wi.abs(), bi.abs())
                m = torch.minimum(m, (pre.abs() / sc).min(dim=2).values.min(dim=0).values)
                g = torch.relu(pre)
            out[b0:b0 + row_chunk] = m.numpy()
    return out


# ----------------------------------------------------------------------------------------------
# README AR(1) (config C1): /root/reference/README.md:22-75 with y0 named 'y0' (README reuses 'x0')
# and LogitNormalVariable defined as torch TransformedDistribution(Normal, SigmoidTransform)
# following the LogNormal pattern (distributions.py:493-507); see SURVEY §8 a13.
# ----------------------------------------------------------------------------------------------
def ar1_elbo(y, params, eps, measure_noise=0.3, dtype=torch.float32):
    """params keys (reference names): b_loc, b_scale, logit_b_post_value, x0_loc, x0_scale,
    x{t}_mean_value (t>=1), x{t}_scale (t>=1); eps keys: "b" [S], "x0".."x{T-1}" [S].

    Collision mode (every hyper-parameter is numeric on both sides): p's `x{t}_scale` root takes q's
    learnable sigma_t, p's `b_loc/b_scale` take q's, so log p(b) - log q(b) cancels identically
    (LogitNormal has no analytic entropy => entropy term is -log q(b), variables.py:156-162).
    """
    T = len(y)
    P = {k: _t(v, dtype, True) for k, v in params.items()}
    E = {k: _t(v, dtype) for k, v in eps.items()}
    yt = _t(y, dtype)
    sig_b = F.softplus(P["b_scale"])
    base_b = D.Normal(P["b_loc"], sig_b)
    qb = D.TransformedDistribution(base_b, [D.transforms.SigmoidTransform()])
    u = P["b_loc"] + sig_b * E["b"]
    b = torch.sigmoid(u)                                  # [S]
    coef = torch.sigmoid(P["logit_b_post_value"])
    sig = [F.softplus(P["x0_scale"])] + [F.softplus(P["x%d_scale" % t]) for t in range(1, T)]
    x = [P["x0_loc"] + sig[0] * E["x0"]]
    qloc = [P["x0_loc"].expand_as(x[0])]
    for t in range(1, T):
        m = coef * x[t - 1] + P["x%d_mean_value" % t]
        qloc.append(m)
        x.append(m + sig[t] * E["x%d" % t])
    logp = qb.log_prob(b)                                  # p(b) with q's (collided) parameters
    logp = logp + D.Normal(P["x0_loc"], sig[0]).log_prob(x[0])         # x0_loc/x0_scale collide as well
    for t in range(1, T):
        logp = logp + D.Normal(b * x[t - 1], sig[t]).log_prob(x[t])
    for t in range(T):
        logp = logp + D.Normal(x[t], torch.tensor(measure_noise, dtype=dtype)).log_prob(yt[t])
    ent = -qb.log_prob(b)                                   # no analytic entropy -> -log q  [S]
    for t in range(T):
        ent = ent + D.Normal(qloc[t], sig[t]).entropy()
    loss = -(logp + ent).mean()
    loss.backward()
    return float(loss.detach()), {k: v.grad.detach().numpy().copy() for k, v in P.items()}


# ----------------------------------------------------------------------------------------------
# SVGD particle loss (config C4): SteinVariationalGradientDescent.compute_loss, inference.py:292-299:
#   loss = sum_particles -joint.calculate_log_probability(sample_k)   (prior + summed observed log-lik)
# ----------------------------------------------------------------------------------------------
def particles_loss_grad(X, y, theta, prior=None, dtype=torch.float32, likelihood="categorical"):
    """theta [n, C, F] particles of (multi-class) logistic regression; prior = (loc, scale) arrays [C,F] or None.
    Returns (loss, G [n, C, F] = d loss / d theta)."""
    Xt = _t(X, dtype)
    th = _t(theta, dtype, requires_grad=True)
    logits = torch.einsum("ncf,bf->nbc", th, Xt)
    if Xt.shape[0] == 0:                         # empty minibatch: torch.distributions cannot validate 0-row logits
        ll = (th * 0).sum()
    elif likelihood == "categorical":
        yt = torch.as_tensor(np.asarray(y), dtype=torch.long)
        ll = D.Categorical(logits=logits).log_prob(yt[None, :]).sum()
    else:
        yt = _t(y, dtype)
        ll = D.Binomial(total_count=1, logits=logits[..., 0]).log_prob(yt[None, :]).sum()
    loss = -ll
    if prior is not None:
        loss = loss - D.Normal(_t(prior[0], dtype), _t(prior[1], dtype)).log_prob(th).sum()
    loss.backward()
    return float(loss.detach()), th.grad.detach().numpy().copy()


# ----------------------------------------------------------------------------------------------
# SVGD direction (config C4): inference.py:301-324, vectorised.
# ----------------------------------------------------------------------------------------------
def svgd_direction(theta, grad, dtype=np.float64):
    """theta, grad: [n,d] (grad = d loss/d theta = -d log p).  Returns (new_grad [n,d], bandwidth).

    update_bandwidth (inference.py:317-324): bw = 2*median_{i!=j}(||theta_i-theta_j||)^2 / ln n
    kernel (:284,304-307):   K_ij = exp(-||theta_i-theta_j||^2 / (2 bw))
    interaction (:308-311):  I[j][i] = -(theta_i - theta_j) K[j][i] / bw  (indexed [index2][index1])
    new grad (:312-315):     g_i = sum_j K[i][j] grad_j + I[i][j]
                                 = (K @ grad)_i + sum_j -(theta_j - theta_i) K_ij / bw
                                 = (K @ grad)_i + (rowsum(K)_i theta_i - (K @ theta)_i) / bw
    (sign of the second term is the reference's, opposite to canonical SVGD; no 1/n.)
    """
    th = np.asarray(theta, dtype=dtype)
    g = np.asarray(grad, dtype=dtype)
    n = th.shape[0]
    sq = (th * th).sum(1)
    d2 = np.maximum(sq[:, None] + sq[None, :] - 2.0 * th @ th.T, 0.0)
    np.fill_diagonal(d2, 0.0)
    off = ~np.eye(n, dtype=bool)
    dist = np.sqrt(d2[off])
    bw = 2.0 * np.median(dist) ** 2 / np.log(n)
    K = np.exp(-d2 / (2.0 * bw))
    out = K @ g + (K.sum(1, keepdims=True) * th - K @ th) / bw
    return out.astype(dtype), float(bw)


def svgd_median_radix_sharded(d2, shards, sum_over_ranks=None, min_over_ranks=None):
    """The exact-median selection of update_bandwidth (inference.py:317-324: np.median over all pairwise distances) restated
    the way the row-sharded K4b evaluates it (csrc/svgd.cu, brn_svgd_sharded_phase): the order statistic(s) of the fp32 bit
    patterns of the n(n-1)/2 upper-triangle squared distances by three radix passes (bits 31..20, 19..8, 7..0) plus a
    successor pass, where every rank only looks at ITS rows [row0, row0 + rows) and the 4096-bin histograms (then the
    (count <= selected, smallest value above) pair) are combined over the ranks.

    d2 [n, n] float32 squared distances; shards = this process's list of (row0, rows) blocks (all blocks of all ranks partition
    the rows); sum_over_ranks / min_over_ranks combine an int64 array over the ranks (identity when None: single process
    holding every shard).  Returns the median DISTANCE as np.float32 arithmetic does: (sqrt(v_k1) + sqrt(v_k2)) / 2."""
    d2 = np.ascontiguousarray(d2, dtype=np.float32)
    n = d2.shape[0]
    bits = d2.view(np.uint32)
    mine = np.concatenate([bits[i, i + 1:] for row0, rows in shards for i in range(row0, row0 + rows)] + [np.zeros(0, np.uint32)])
    m = n * (n - 1) // 2
    k1, k2 = (m - 1) // 2, m // 2
    ident = lambda a: a
    sum_over_ranks = sum_over_ranks or ident
    min_over_ranks = min_over_ranks or ident
    prefix, rank = np.uint32(0), k1
    for shift, nbits in ((20, 12), (8, 12), (0, 8)):
        hshift = shift + nbits
        sel = mine if hshift >= 32 else mine[(mine >> np.uint32(hshift)) == (prefix >> np.uint32(hshift))]
        hist = np.bincount(((sel >> np.uint32(shift)) & np.uint32((1 << nbits) - 1)).astype(np.int64), minlength=1 << nbits)
        hist = sum_over_ranks(hist.astype(np.int64))
        cum = np.cumsum(hist)
        b = int(np.searchsorted(cum, rank, side="right"))
        rank -= int(cum[b - 1]) if b > 0 else 0
        prefix = np.uint32(prefix | np.uint32(b << shift))
    cnt_le = sum_over_ranks(np.array([np.count_nonzero(mine <= prefix)], np.int64))[0]
    above = mine[mine > prefix]
    nxt = min_over_ranks(np.array([above.min() if above.size else 0x7f800000], np.int64))[0]
    v1 = np.array([prefix], np.uint32).view(np.float32)[0]
    v2 = v1 if cnt_le >= k2 + 1 else np.array([nxt], np.uint32).view(np.float32)[0]
    return np.float32(0.5) * (np.sqrt(v1) + np.sqrt(v2))


def svgd_direction_loops(theta, grad):
    """Literal O(n^2) loop transcription of the index pattern in inference.py:301-324 for tiny n,
    used only to pin `svgd_direction`'s vectorised algebra (fp64 numpy)."""
    th = np.asarray(theta, dtype=np.float64)
    g = np.asarray(grad, dtype=np.float64)
    n = th.shape[0]
    dev = lambda a, b: float(((a - b) ** 2).sum())
    dists = [math.sqrt(dev(th[i], th[j])) for i in range(n) for j in range(n) if i != j]
    bw = 2 * np.median(dists) ** 2 / np.log(n)
    kern = [[math.exp(-dev(th[p1], th[p2]) / (2 * bw)) for p1 in range(n)] for p2 in range(n)]
    inter = [[-(th[p1] - th[p2]) * kern[p1][p2] / bw for p1 in range(n)] for p2 in range(n)]
    out = np.zeros_like(th)
    for i in range(n):
        out[i] = sum(kern[i][j] * g[j] + inter[i][j] for j in range(n))
    return out, float(bw)


# ----------------------------------------------------------------------------------------------
# WVGD: WassersteinVariationalGradientDescent.compute_loss, inference.py:154-229, for the model class of
# development_playgrounds/WVGD_logistic_regression.py (one Normal sampler + one root particle per ensemble member).
# ----------------------------------------------------------------------------------------------
def voronoi_owner(z, theta, first_column_only=True):
    """z [m, C, F] samples, theta [n, C, F] particle locations -> owner [m] = argmin_j cost(z, theta_j).

    The reference evaluates its cost on NUMPY arrays (inference.py:177-186): `sum_from_dim` then takes
    `np.sum(x, axis=(1, ..., ndim-2))[:, 0]` (utilities.py:125-126), i.e. for values of shape (S,1,C,F) it sums over the
    C axis and keeps ONLY index 0 of the last axis: cost = sum_c (z[c,0] - theta_j[c,0])^2.  `first_column_only=True`
    reproduces that; False is the evident intent (full squared distance).  fp32 like the reference's arrays;
    np.argmin = first minimal index (inference.py:188-189)."""
    z = np.asarray(z, np.float32)
    th = np.asarray(theta, np.float32)
    if first_column_only:
        z, th = z[:, :, :1], th[:, :, :1]
    d2 = ((z[:, None] - th[None]) ** 2).sum(axis=(2, 3), dtype=np.float32)
    return np.argmin(d2, axis=1)


def wvgd_loss(X, y, theta, loc, rho, eps_elbo, eps_particle, prior=None, dtype=torch.float32, likelihood="categorical",
              biased=False, first_column_only=True):
    """theta, loc [n,C,F]; rho [n] (scalar scale per sampler) or [n,C,F]; eps_* [n,S,C,F] (two independent draws per
    sampler: one for the sampler ELBO, one for the particle loss -- inference.py:203-212).

    Per sampler k, with z = loc_k + softplus(rho_k) eps and A_k = {s : voronoi_owner(z_s) == k} (rejection with max_itr=1,
    transformations.py:28-43; the caller must supply noise with at least one accepted sample per draw):
      -ELBO_k = -mean_{A_k}[ log p(z, data) - log q_k(z) + mean_{A_k} log q_k(z.detach()) ]
                (entropy of a transformed model = -truncated log-prob, variables.py:744-749, whose for_gradient branch adds the
                 detached-mean normaliser, transformations.py:12-22; PathwiseDerivativeEstimator .mean(), gradient_estimators.py:44)
      particle_k = sum_{A'_k} w_s ||theta_k - z'_s.detach()||^2, w = softmax_{A'_k}(log p(z', data) - log q_k(z'))  (1/S if biased)
                (inference.py:211-229, variables.py:821-841)
    prior=None: tied mode (p's auto-named weights_loc/weights_scale roots take q_k's values, utilities.py:282-309), else the
    declared (loc, scale).  Returns (loss, {"loc": [n,C,F], "rho": like rho, "theta": [n,C,F]}, accepted counts [n,2])."""
    Xt = _t(X, dtype)
    n = len(theta)
    th = _t(theta, dtype, True)
    lc = _t(loc, dtype, True)
    rh = _t(rho, dtype, True)
    e1, e2 = _t(eps_elbo, dtype), _t(eps_particle, dtype)
    if likelihood == "categorical":
        yt = torch.as_tensor(np.asarray(y), dtype=torch.long)
        loglik = lambda z: D.Categorical(logits=torch.einsum("scf,bf->sbc", z, Xt)).log_prob(yt[None, :]).sum(1)
    else:
        yt = _t(y, dtype)
        loglik = lambda z: D.Binomial(total_count=1, logits=torch.einsum("scf,bf->sbc", z, Xt)[..., 0]).log_prob(yt[None, :]).sum(1)
    total = 0.0
    counts = np.zeros((n, 2), np.int64)
    for k in range(n):
        sg = F.softplus(rh[k])
        q = D.Normal(lc[k], sg.expand_as(lc[k]))
        pr = q if prior is None else D.Normal(_t(prior[0], dtype), _t(prior[1], dtype))
        # sampler ELBO
        z = lc[k] + sg * e1[k]
        acc = torch.as_tensor(voronoi_owner(z.detach().float().numpy(), th.detach().float().numpy(), first_column_only) == k)
        counts[k, 0] = int(acc.sum())
        za = z[acc]
        logq = q.log_prob(za).sum((1, 2))
        norm = -q.log_prob(za.detach()).sum((1, 2)).mean().detach()     # `nondiff_values` detaches the root samples too
        elbo = (loglik(za) + pr.log_prob(za).sum((1, 2)) - (logq + norm)).mean()
        # particle loss (fresh draw, everything about the samples detached)
        z2 = (lc[k] + sg * e2[k]).detach()
        acc2 = torch.as_tensor(voronoi_owner(z2.float().numpy(), th.detach().float().numpy(), first_column_only) == k)
        counts[k, 1] = int(acc2.sum())
        z2a = z2[acc2]
        if biased:
            w = torch.full((z2a.shape[0],), 1.0 / eps_particle.shape[1], dtype=dtype)
        else:
            logw = (loglik(z2a) + pr.log_prob(z2a).sum((1, 2)) - q.log_prob(z2a).sum((1, 2))).detach()
            w = torch.softmax(logw, 0)
        ploss = (w * ((th[k][None] - z2a) ** 2).sum((1, 2))).sum()
        total = total - elbo + ploss
    total.backward()
    return float(total.detach()), {"loc": lc.grad.numpy().copy(), "rho": rh.grad.numpy().copy(), "theta": th.grad.numpy().copy()}, counts


def wvgd_ensemble_weights(X, y, theta, loc, rho, eps, prior=None, dtype=torch.float64, likelihood="categorical",
                          first_column_only=True):
    """WassersteinVariationalGradientDescent.post_process (inference.py:234-247) + get_importance_weights
    (variables.py:821-841): per sampler k, z = loc_k + softplus(rho_k) eps_k [S,C,F], A_k = {s : voronoi_owner(z_s) == k}
    (rejection with max_itr = 1), logZ_k = log sum_{A_k} exp(log p(z, data) - log q_k(z)) -- the UNNORMALISED q log-prob
    (normalized=False) and no division by the count; weights = softmax_k(logZ_k).  prior=None: tied mode (log p(z) = log q_k(z)).
    Returns (weights [n], logZ [n], accepted counts [n])."""
    Xt = _t(X, dtype)
    n = len(theta)
    if likelihood == "categorical":
        yt = torch.as_tensor(np.asarray(y), dtype=torch.long)
        loglik = lambda z: D.Categorical(logits=torch.einsum("scf,bf->sbc", z, Xt)).log_prob(yt[None, :]).sum(1)
    else:
        yt = _t(y, dtype)
        loglik = lambda z: D.Binomial(total_count=1, logits=torch.einsum("scf,bf->sbc", z, Xt)[..., 0]).log_prob(yt[None, :]).sum(1)
    logZ, counts = np.zeros(n), np.zeros(n, np.int64)
    with torch.no_grad():
        for k in range(n):
            lc = _t(loc[k], dtype)
            sg = F.softplus(_t(rho[k], dtype))
            q = D.Normal(lc, sg.expand_as(lc))
            pr = q if prior is None else D.Normal(_t(prior[0], dtype), _t(prior[1], dtype))
            z = lc + sg * _t(eps[k], dtype)
            acc = torch.as_tensor(voronoi_owner(z.float().numpy(), np.asarray(theta, np.float32), first_column_only) == k)
            counts[k] = int(acc.sum())
            za = z[acc]
            logw = loglik(za) + pr.log_prob(za).sum((1, 2)) - q.log_prob(za).sum((1, 2))
            logZ[k] = float(torch.logsumexp(logw, 0)) if counts[k] else -np.inf
    w = np.exp(logZ - logZ.max())
    return w / w.sum(), logZ, counts


# ----------------------------------------------------------------------------------------------
# Full-size evaluations (BASELINE configs C2 / C4): the same objectives with the gradient written out by hand and the
# rows streamed in chunks, so that 10^6 rows x 1024 samples fit in host memory (autograd would keep every chunk's
# [S, rows] logits alive).  Each is pinned against the autograd restatement above at small sizes in
# tests/test_oracle_golden.py before it is trusted at the configured size.
# ----------------------------------------------------------------------------------------------
def logreg_elbo_streamed(X, y, params, eps, prior=None, dtype=torch.float64, row_chunk=65536):
    """Binomial(1, logits) logistic regression, C = 1.  Same value and gradients as `logreg_elbo` (reference lines cited
    there): loss = -(mean_s[sum_b (y l - softplus(l)) + log p(w_s)] + H[q]); d/dmu, d/drho through w = mu + softplus(rho) eps,
    the prior (tied: q's own loc / scale) and the analytic entropy."""
    mu, rho = (_t(a, dtype) for a in params["weights"])
    e = _t(eps["weights"], dtype)                                   # [S, 1, F]
    S, F = e.shape[0], e.shape[-1]
    sg = F_softplus(rho)
    w = (mu + sg * e).reshape(S, F)
    Xt = torch.as_tensor(np.asarray(X))
    yt = torch.as_tensor(np.asarray(y))
    ll = torch.zeros(S, dtype=dtype)
    gw = torch.zeros(S, F, dtype=dtype)
    with torch.no_grad():
        for r0 in range(0, Xt.shape[0], row_chunk):
            xb = Xt[r0:r0 + row_chunk].to(dtype)
            yb = yt[r0:r0 + row_chunk].to(dtype)
            L = w @ xb.T                                            # [S, rows]
            ll += (yb[None, :] * L - torch.nn.functional.softplus(L)).sum(1)
            gw += (yb[None, :] - torch.sigmoid(L)) @ xb             # d ll_s / d w_s
    c = 0.5 * math.log(2 * math.pi)
    e2 = e.reshape(S, -1)
    muf, sgf, rhof = mu.reshape(-1), sg.reshape(-1), rho.reshape(-1)
    if prior is None:                                               # tied: log N(w; mu, sg) = -eps^2/2 - log sg - c
        lp = (-0.5 * e2 * e2 - torch.log(sgf) - c).sum(1)
        dlp_dmu = torch.zeros_like(muf)
        dlp_dsg = -1.0 / sgf                                        # per sample; the eps-dependent parts cancel
        dlp_dsg = dlp_dsg.expand(S, -1)
        dlp_dmu = dlp_dmu.expand(S, -1)
    else:
        a = torch.as_tensor(np.broadcast_to(np.asarray(prior["weights"][0], dtype=np.float64), tuple(mu.shape)).copy()).to(dtype).reshape(-1)
        b = torch.as_tensor(np.broadcast_to(np.asarray(prior["weights"][1], dtype=np.float64), tuple(mu.shape)).copy()).to(dtype).reshape(-1)
        d = w - a
        lp = (-0.5 * d * d / (b * b) - torch.log(b) - c).sum(1)
        g = -d / (b * b)
        dlp_dmu, dlp_dsg = g, g * e2
    ent = (0.5 + c + torch.log(sgf)).sum()
    loss = -((ll + lp).mean() + ent)
    dE_dmu = (gw + dlp_dmu).mean(0)
    dE_dsg = (gw * e2 + dlp_dsg).mean(0) + 1.0 / sgf
    grads = {"weights_loc": (-dE_dmu).reshape(tuple(mu.shape)).numpy(),
             "weights_scale": (-dE_dsg * torch.sigmoid(rhof)).reshape(tuple(mu.shape)).numpy()}
    return float(loss), grads


def particles_loss_grad_streamed(X, y, theta, prior=None, dtype=torch.float64, row_chunk=16384):
    """Binomial(1, logits) particles (C = 1), hand-written gradient, rows streamed: same as `particles_loss_grad`."""
    th = _t(theta, dtype).reshape(theta.shape[0], -1)               # [n, F]
    Xt = torch.as_tensor(np.asarray(X))
    yt = torch.as_tensor(np.asarray(y))
    ll = torch.zeros((), dtype=dtype)
    G = torch.zeros_like(th)
    with torch.no_grad():
        for r0 in range(0, Xt.shape[0], row_chunk):
            xb = Xt[r0:r0 + row_chunk].to(dtype)
            yb = yt[r0:r0 + row_chunk].to(dtype)
            L = th @ xb.T
            ll += (yb[None, :] * L - torch.nn.functional.softplus(L)).sum()
            G -= (yb[None, :] - torch.sigmoid(L)) @ xb
        loss = -ll
        if prior is not None:
            a, b = _t(prior[0], dtype).reshape(-1), _t(prior[1], dtype).reshape(-1)
            d = th - a
            loss = loss - (-0.5 * d * d / (b * b) - torch.log(b) - 0.5 * math.log(2 * math.pi)).sum()
            G += d / (b * b)
    return float(loss), G.numpy().reshape(np.asarray(theta).shape)
