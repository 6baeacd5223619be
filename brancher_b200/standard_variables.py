"""User-facing variable constructors (mirror of brancher/standard_variables.py): same names, argument
order and meaning.  Numeric hyper-parameters become auto-named root parents `<name>_<param>` stored in
the unconstrained space of the parameter's geometric range (standard_variables.py:57-68) -- the naming
is what produces the reference's p/q root collision, and it is preserved on purpose.
"""
import numbers

import numpy as np
import torch.nn as nn

import brancher_b200.distributions as distributions
import brancher_b200.geometric_ranges as geometric_ranges
from brancher_b200.variables import var2link, Variable, RootVariable, RandomVariable, PartialLink


class LinkConstructor(nn.ModuleList):
    """{parameter name: link} evaluated together; an nn.ModuleList of the nn.Modules the links use so that
    `ProbabilisticOptimizer` finds their parameters (standard_variables.py:14-29)."""

    def __init__(self, **kwargs):
        links = {k: var2link(v) for k, v in kwargs.items()}
        modules = []
        for l in links.values():
            if isinstance(l, PartialLink):
                for m in l.links:
                    if isinstance(m, nn.Module) and all(m is not x for x in modules):
                        modules.append(m)
        super().__init__(modules)
        self.kwargs = kwargs
        self._links = links

    def forward(self, values):
        return {k: (l.fn(values) if isinstance(l, PartialLink) else l) for k, l in self._links.items()}


class VariableConstructor(RandomVariable):
    """Builds a RandomVariable from keyword parameters that may be numbers, arrays, variables or links
    (standard_variables.py:32-68)."""

    def __init__(self, name, learnable, ranges, is_observed=False, **kwargs):
        self.name = name
        self._init_state(is_observed)
        self.roots = {}
        for pname, value in list(kwargs.items()):
            if isinstance(value, (Variable, PartialLink)):
                continue
            if isinstance(value, np.ndarray):
                dim = value.shape[0] if value.ndim else 1
            elif isinstance(value, numbers.Number):
                dim = 1
            else:
                dim = []
            rng = ranges[pname]
            root = RootVariable(rng.inverse_transform(value, dim), name + "_" + pname, learnable,
                                is_observed=self._observed)
            self.roots[pname] = root
            kwargs[pname] = rng.forward_transform(root, dim)
        self.partial_links = {k: var2link(v) for k, v in kwargs.items()}
        self.parents = set().union(*[l.vars for l in self.partial_links.values() if isinstance(l, PartialLink)])
        self.ancestors = set(self.parents).union(*[p.ancestors for p in self.parents]) if self.parents else set()
        self.link = LinkConstructor(**kwargs)
        self.ranges = ranges
        self.learnable = learnable
        self.is_normalized = True


class EmpiricalVariable(VariableConstructor):
    """Minibatch view of a dataset (standard_variables.py:71-96)."""

    def __init__(self, dataset, name, learnable=False, is_observed=False, batch_size=None, indices=None, weights=None):
        self._type = "Empirical"
        given = {k: v for k, v in (("dataset", dataset), ("batch_size", batch_size), ("indices", indices),
                                   ("weights", weights)) if v is not None}
        ranges = {k: geometric_ranges.UnboundedRange() for k in given}
        super().__init__(name, **given, learnable=learnable, ranges=ranges, is_observed=is_observed)
        if not batch_size:
            if indices is not None:
                batch_size = len(indices)
            else:
                raise ValueError("Either the indices or the batch size has to be given as input")
        self.batch_size = batch_size
        self.distribution = distributions.EmpiricalDistribution(batch_size=batch_size, is_observed=is_observed)


class RandomIndices(EmpiricalVariable):
    """standard_variables.py:99-112"""

    def __init__(self, dataset_size, batch_size, name, is_observed=False):
        super().__init__(dataset=list(range(dataset_size)), batch_size=batch_size, is_observed=is_observed, name=name)
        self._type = "Random Index"
        self.dataset_size = dataset_size

    def __len__(self):
        return self.batch_size


class DeterministicVariable(VariableConstructor):
    """standard_variables.py:115-130"""

    def __init__(self, value, name, learnable=False, is_observed=False, variable_range=geometric_ranges.UnboundedRange()):
        self._type = "Deterministic node"
        super().__init__(name, value=value, learnable=learnable, ranges={"value": variable_range}, is_observed=is_observed)
        self.distribution = distributions.DeterministicDistribution()

    @property
    def value(self):
        return self._get_sample(1)[self]


def _loc_scale(cls_name, type_name, dist_cls, doc):
    def __init__(self, loc, scale, name, learnable=False, is_observed=False):
        self._type = type_name
        ranges = {"loc": geometric_ranges.UnboundedRange(), "scale": geometric_ranges.RightHalfLine(0.)}
        VariableConstructor.__init__(self, name, loc=loc, scale=scale, learnable=learnable, ranges=ranges,
                                     is_observed=is_observed)
        self.distribution = dist_cls()
    return type(cls_name, (VariableConstructor,), {"__init__": __init__, "__doc__": doc})


NormalVariable = _loc_scale("NormalVariable", "Normal", distributions.NormalDistribution,
                            "standard_variables.py:133-145")
LogNormalVariable = _loc_scale("LogNormalVariable", "Log Normal", distributions.LogNormalDistribution,
                               "standard_variables.py:186-198")
LogitNormalVariable = _loc_scale("LogitNormalVariable", "Logit Normal", distributions.LogitNormalDistribution,
                                 "README.md:30,56 (commented out in the reference, standard_variables.py:201-213)")
CauchyVariable = _loc_scale("CauchyVariable", "Cauchy", distributions.CauchyDistribution, "standard_variables.py:156-168")
LaplaceVariable = _loc_scale("LaplaceVariable", "Laplace", distributions.LaplaceDistribution,
                             "standard_variables.py:171-183")


class BetaVariable(VariableConstructor):
    """standard_variables.py:216-231"""

    def __init__(self, alpha, beta, name, learnable=False, is_observed=False):
        self._type = "Beta"
        ranges = {"concentration1": geometric_ranges.RightHalfLine(0.), "concentration0": geometric_ranges.RightHalfLine(0.)}
        super().__init__(name, concentration1=alpha, concentration0=beta, learnable=learnable, ranges=ranges,
                         is_observed=is_observed)
        self.distribution = distributions.BetaDistribution()


def _probs_or_logits(probs, logits):
    if (probs is None) == (logits is None):
        raise ValueError("Either probs or logits needs to be provided as input")
    if probs is not None:
        return "probs", probs, geometric_ranges.Interval(0., 1.)
    return "logits", logits, geometric_ranges.UnboundedRange()


class BinomialVariable(VariableConstructor):
    """standard_variables.py:234-255"""

    def __init__(self, total_count, probs=None, logits=None, name="Binomial", learnable=False, is_observed=False):
        self._type = "Binomial"
        key, val, rng = _probs_or_logits(probs, logits)
        ranges = {"total_count": geometric_ranges.UnboundedRange(), key: rng}
        super().__init__(name, total_count=total_count, **{key: val}, learnable=learnable, ranges=ranges,
                         is_observed=is_observed)
        self.distribution = distributions.BinomialDistribution()


class BernulliVariable(VariableConstructor):
    """standard_variables.py:258-277"""

    def __init__(self, probs=None, logits=None, name="Bernulli", learnable=False, is_observed=False):
        self._type = "Bernulli"
        key, val, rng = _probs_or_logits(probs, logits)
        super().__init__(name, **{key: val}, learnable=learnable, ranges={key: rng}, is_observed=is_observed)
        self.distribution = distributions.BernulliDistribution()


class CategoricalVariable(VariableConstructor):
    """standard_variables.py:280-299 (logits= with integer labels is the branch that works in the reference)."""

    def __init__(self, probs=None, logits=None, name="Categorical", learnable=False, is_observed=False):
        self._type = "Categorical"
        if (probs is None) == (logits is None):
            raise ValueError("Either probs or logits needs to be provided as input")
        key, val = ("probs", probs) if probs is not None else ("logits", logits)
        super().__init__(name, **{key: val}, learnable=learnable, ranges={key: geometric_ranges.UnboundedRange()},
                         is_observed=is_observed)
        self.distribution = distributions.CategoricalDistribution()
