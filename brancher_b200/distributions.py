"""Distribution adaptors for the EAGER sampling API (mirror of brancher/distributions.py).

Only the sampling / log-prob-of-given-values API (`get_sample`, `_get_posterior_sample`,
`calculate_log_probability`) goes through these torch.distributions adaptors; the ELBO hot path uses
the fused kernels and only reads the class-level metadata here (`kind`, flags) when lowering a graph.
"""
import numpy as np
import torch
from torch import distributions as D

from brancher_b200 import config
from brancher_b200.utilities import broadcast_all, sum_from_dim, is_tensor, is_discrete, batch_sizes, flatten_batch


class Distribution:
    kind = "generic"
    has_differentiable_samples = False
    is_finite = False
    is_discrete = False
    has_analytic_entropy = False
    has_analytic_mean = False
    has_analytic_var = False
    required_parameters = ()

    def check_parameters(self, **parameters):
        for req in self.required_parameters:
            names = req if isinstance(req, tuple) else (req,)
            assert any(n in parameters for n in names), "missing parameter %s" % (names,)

    def torchdist(self, **parameters):
        raise NotImplementedError

    # univariate default: broadcast everything, sum event dims
    def calculate_log_probability(self, x, **parameters):
        self.check_parameters(**parameters)
        keys = list(parameters)
        vals = broadcast_all(x, *[parameters[k] for k in keys])
        lp = self.torchdist(**dict(zip(keys, vals[1:]))).log_prob(vals[0])
        return sum_from_dim(lp, 2)

    def _broadcast(self, parameters):
        keys = list(parameters)
        vals = broadcast_all(*[parameters[k] for k in keys])
        return dict(zip(keys, vals))

    def get_sample(self, differentiable=True, **parameters):
        self.check_parameters(**parameters)
        dist = self.torchdist(**self._broadcast(parameters))
        return dist.rsample() if (self.has_differentiable_samples and differentiable) else dist.sample()

    def get_entropy(self, **parameters):
        if not self.has_analytic_entropy:
            raise ValueError("The entropy of the distribution cannot be computed analytically")
        return self.torchdist(**self._broadcast(parameters)).entropy()

    def get_mean(self, **parameters):
        if not self.has_analytic_mean:
            raise ValueError("The mean of the distribution cannot be computed analytically")
        return self.torchdist(**self._broadcast(parameters)).mean

    def get_variance(self, **parameters):
        if not self.has_analytic_var:
            raise ValueError("The variance of the distribution cannot be computed analytically")
        return self.torchdist(**self._broadcast(parameters)).variance


class DeterministicDistribution(Distribution):
    """brancher/distributions.py:334-390"""
    kind = "deterministic"
    required_parameters = ("value",)
    has_differentiable_samples = True
    is_finite = True
    is_discrete = True
    has_analytic_entropy = True
    has_analytic_mean = True
    has_analytic_var = True

    def calculate_log_probability(self, x, **parameters):
        return torch.zeros((1, 1), device=config.device)

    def get_sample(self, differentiable=True, **parameters):
        return parameters["value"]

    def get_mean(self, **parameters):
        return parameters["value"]

    def get_entropy(self, **parameters):
        return torch.zeros((1, 1, 1), device=config.device)

    def get_variance(self, **parameters):
        return torch.zeros((1, 1, 1), device=config.device)


class EmpiricalDistribution(Distribution):
    """Minibatch sampler without replacement (brancher/distributions.py:393-473)."""
    kind = "empirical"
    required_parameters = ("dataset",)
    is_finite = True
    is_discrete = True
    has_analytic_entropy = True

    def __init__(self, batch_size, is_observed):
        self.batch_size = batch_size
        self.is_observed = is_observed

    def calculate_log_probability(self, x, **parameters):
        return torch.zeros((1, 1), device=config.device)

    def get_sample(self, differentiable=True, **parameters):
        dataset = parameters["dataset"]
        if "indices" in parameters:
            indices = parameters["indices"]
        else:
            p = None
            if "weights" in parameters:
                p = np.asarray(parameters["weights"], dtype="float64")
                p = p / p.sum()
            if is_tensor(dataset):
                size = dataset.shape[1] if self.is_observed else dataset.shape[2]
            else:
                size = len(dataset)
            if size < self.batch_size:
                raise ValueError("It is impossible to have more samples than the size of the dataset without replacement")
            if (p is None and is_tensor(dataset) and dataset.is_cuda and 2 * self.batch_size <= size and self.batch_size <= 8192):
                # device-side sampling (brn_minibatch_indices): the reference permutes all `size` row ids on the HOST for every
                # iteration (np.random.choice(range(N), B, replace=False): ~140 ms at N = 10^6); here the
                # distinct row ids are drawn by one small kernel and never leave the GPU.  Independent Philox stream per
                # draw (config.seed, config.next_minibatch_offset()); same distribution, different RNG than numpy's.
                from brancher_b200 import _cuda as cu
                draw_dev = lambda: cu.minibatch_indices(size, self.batch_size, dataset.device, seed=config.seed,
                                                        offset=config.next_minibatch_offset())
                # one independent index set per leading (MC-sample) row of the dataset, as the reference draws them
                rows = [dataset[n].index_select(0 if self.is_observed else 1, draw_dev()) for n in range(dataset.shape[0])]
                return torch.stack(rows, dim=0)
            draw = lambda: np.random.choice(size, size=self.batch_size, replace=False, p=p)
            indices = draw() if is_discrete(dataset) else [draw() for _ in range(dataset.shape[0])]
        if not is_tensor(dataset):
            return list(np.array(dataset)[indices])
        if isinstance(indices, list) and len(indices) and isinstance(indices[0], np.ndarray):
            rows = [dataset[n, torch.as_tensor(k, device=dataset.device)] if self.is_observed
                    else dataset[n, :, torch.as_tensor(k, device=dataset.device)] for n, k in enumerate(indices)]
            return torch.stack(rows, dim=0)
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=dataset.device)
        return dataset.index_select(1 if self.is_observed else 2, idx)

    def get_entropy(self, **parameters):
        if "weights" in parameters:
            probs = torch.as_tensor(parameters["weights"], dtype=torch.float32, device=config.device)
        else:
            ds = parameters["dataset"]
            n = int(ds.shape[0]) if is_tensor(ds) else len(ds)
            probs = torch.ones(n, device=config.device)
        return D.Categorical(probs=probs).entropy()


class NormalDistribution(Distribution):
    """brancher/distributions.py:476-490"""
    kind = "normal"
    required_parameters = ("loc", "scale")
    has_differentiable_samples = True
    has_analytic_entropy = True
    has_analytic_mean = True
    has_analytic_var = True

    def torchdist(self, loc, scale):
        return D.Normal(loc, scale)


class LogNormalDistribution(NormalDistribution):
    """brancher/distributions.py:493-507"""
    kind = "lognormal"

    def torchdist(self, loc, scale):
        return D.LogNormal(loc, scale)


class LogitNormalDistribution(NormalDistribution):
    """The LogitNormal the reference's README uses but its code has commented out
    (standard_variables.py:201-213): Normal pushed through a sigmoid, following the LogNormal pattern;
    no analytic entropy (=> -log q, variables.py:156-162)."""
    kind = "logitnormal"
    has_analytic_entropy = False
    has_analytic_mean = False
    has_analytic_var = False

    def torchdist(self, loc, scale):
        return D.TransformedDistribution(D.Normal(loc, scale), [D.transforms.SigmoidTransform()])


class CauchyDistribution(NormalDistribution):
    kind = "cauchy"
    has_analytic_var = False

    def torchdist(self, loc, scale):
        return D.Cauchy(loc, scale)


class LaplaceDistribution(NormalDistribution):
    kind = "laplace"

    def torchdist(self, loc, scale):
        return D.Laplace(loc, scale)


class BetaDistribution(Distribution):
    kind = "beta"
    required_parameters = ("concentration1", "concentration0")
    has_differentiable_samples = True
    has_analytic_entropy = True
    has_analytic_mean = True
    has_analytic_var = True

    def torchdist(self, concentration1, concentration0):
        return D.Beta(concentration1, concentration0)


class BinomialDistribution(Distribution):
    """brancher/distributions.py:561-575"""
    kind = "binomial"
    required_parameters = ("total_count", ("probs", "logits"))
    is_finite = True
    is_discrete = True
    has_analytic_mean = True
    has_analytic_var = True

    def torchdist(self, **p):
        return D.Binomial(**p)


class BernulliDistribution(Distribution):
    """brancher/distributions.py:578-592 (the reference's spelling is kept)."""
    kind = "bernoulli"
    required_parameters = (("probs", "logits"),)
    is_finite = True
    is_discrete = True
    has_analytic_entropy = True
    has_analytic_mean = True
    has_analytic_var = True

    def torchdist(self, **p):
        return D.Bernoulli(**p)


class CategoricalDistribution(Distribution):
    """Integer-label Categorical over the event's first axis (brancher/distributions.py:275-311; only the
    label branch of the reference works)."""
    kind = "categorical"
    required_parameters = (("probs", "logits"),)
    is_finite = True
    is_discrete = True

    def _flat(self, t, S, B):
        t = flatten_batch(t, S, B)
        return t.reshape(t.shape[0], -1)

    def calculate_log_probability(self, x, **parameters):
        self.check_parameters(**parameters)
        S, B = batch_sizes({**parameters, "x": x})
        flat = {k: self._flat(v, S, B) for k, v in parameters.items()}
        labels = self._flat(x, S, B)[:, 0]
        return D.Categorical(**flat).log_prob(labels).reshape(S, B)

    def get_sample(self, differentiable=True, **parameters):
        S, B = batch_sizes(parameters)
        flat = {k: self._flat(v, S, B) for k, v in parameters.items()}
        shape = next(iter(parameters.values())).shape[2:]
        return D.OneHotCategorical(**flat).sample().reshape((S, B) + tuple(shape))
