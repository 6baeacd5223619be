"""Model graph: variables, symbolic links and probabilistic models.

Host-side mirror of brancher/variables.py (same public names and argument meaning), re-designed so
that the graph is *inspectable*: a `PartialLink` carries an explicit expression tree (`Expr`) instead
of an opaque Python closure (reference: variables.py:977-1072, functions.py:28-41).  The tree is what
`brancher_b200.lowering` pattern-matches to pick a fused CUDA kernel family; the same tree can be
evaluated eagerly with torch for the sampling API (`get_sample`, `_get_posterior_sample`), which is not
part of the ELBO hot path.

ELBO evaluation (`estimate_log_model_evidence`, reference variables.py:843-870) never walks the graph
in Python: it is lowered once and executed by the kernels behind `brancher_b200._cuda`.
"""
import numbers
import operator
import warnings
from abc import ABC, abstractmethod
from collections.abc import Iterable

import numpy as np
import torch

from brancher_b200 import config
from brancher_b200 import distributions
from brancher_b200 import gradient_estimators
from brancher_b200.modules import ParameterModule
from brancher_b200.utilities import (coerce_to_dtype, tile_parameter, is_discrete, is_tensor, contains_tensors,
                                     batch_sizes, flatten_batch, unflatten_batch, map_structure, sum_from_dim,
                                     partial_broadcast, get_model_mapping, reassign_samples, to_numpy)


# ---------------------------------------------------------------------------------------------------
# expression tree
# ---------------------------------------------------------------------------------------------------
class Expr:
    """Node of a symbolic link."""
    __slots__ = ()


class VarRef(Expr):
    __slots__ = ("var",)

    def __init__(self, var):
        self.var = var


class Const(Expr):
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


class Call(Expr):
    """`fn(*args, **kwargs)`; `name` is 'add'/'sub'/'mul'/'truediv'/'pow' for operators, the torch
    function name for `BF.<name>`."""
    __slots__ = ("name", "fn", "args", "kwargs")

    def __init__(self, name, fn, args, kwargs=None):
        self.name, self.fn, self.args, self.kwargs = name, fn, list(args), dict(kwargs or {})


class ModuleCall(Call):
    """Call of a torch.nn.Module wrapped by BF.BrancherFunction (functions.py:15-22)."""
    __slots__ = ()


class Index(Expr):
    __slots__ = ("base", "key")

    def __init__(self, base, key):
        self.base, self.key = base, key


class TupleOf(Expr):
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = list(items)


class ShapeOf(Expr):
    __slots__ = ("base",)

    def __init__(self, base):
        self.base = base


class Opaque(Expr):
    """A user-supplied closure: evaluable eagerly, not lowerable."""
    __slots__ = ("fn",)

    def __init__(self, fn):
        self.fn = fn


def evaluate(expr, values):
    """Eager torch evaluation of an expression tree on `values` {Variable: tensor}."""
    if isinstance(expr, VarRef):
        return values[expr.var]
    if isinstance(expr, Const):
        v = expr.value
        if isinstance(v, np.ndarray):
            return torch.as_tensor(v, dtype=torch.float32, device=config.device)
        return v
    if isinstance(expr, Call):
        args = [evaluate(a, values) if isinstance(a, Expr) else a for a in expr.args]
        kwargs = {k: evaluate(a, values) if isinstance(a, Expr) else a for k, a in expr.kwargs.items()}
        return expr.fn(*args, **kwargs)
    if isinstance(expr, Index):
        return evaluate(expr.base, values)[expr.key]
    if isinstance(expr, TupleOf):
        return tuple(evaluate(e, values) for e in expr.items)
    if isinstance(expr, ShapeOf):
        return evaluate(expr.base, values).shape
    if isinstance(expr, Opaque):
        return expr.fn(values)
    raise TypeError("not an expression: %r" % (expr,))


def expr_variables(expr, out=None):
    out = set() if out is None else out
    if isinstance(expr, VarRef):
        out.add(expr.var)
    elif isinstance(expr, Call):
        for a in list(expr.args) + list(expr.kwargs.values()):
            if isinstance(a, Expr):
                expr_variables(a, out)
    elif isinstance(expr, (Index, ShapeOf)):
        expr_variables(expr.base, out)
    elif isinstance(expr, TupleOf):
        for e in expr.items:
            expr_variables(e, out)
    return out


_OP_SYMBOL = {"add": "+", "sub": "-", "mul": "*", "truediv": "/", "pow": "**"}
_OP_FN = {"add": operator.add, "sub": operator.sub, "mul": operator.mul, "truediv": operator.truediv,
          "pow": operator.pow}


# ---------------------------------------------------------------------------------------------------
class BrancherClass(ABC):
    """Abstract superclass of variables, links and models (variables.py:47-100)."""

    @abstractmethod
    def _flatten(self):
        pass

    def flatten(self):
        return set(self._flatten())

    def get_variable(self, var_name):
        for var in self._flatten():
            if var.name == var_name:
                return var
        raise KeyError("The variable {} is not present in the model".format(var_name))


class _Operators:
    """Arithmetic on variables / links builds links (variables.py:230-277, 995-1035)."""

    def _apply_operator(self, other, op):
        return var2link(self)._apply_operator(other, op)

    def __neg__(self):
        return -1 * self

    def __add__(self, other):
        return self._apply_operator(other, "add")

    def __radd__(self, other):
        return self.__add__(other)

    def __sub__(self, other):
        return self._apply_operator(other, "sub")

    def __rsub__(self, other):
        return -1 * self.__sub__(other)

    def __mul__(self, other):
        return self._apply_operator(other, "mul")

    def __rmul__(self, other):
        return self.__mul__(other)

    def __truediv__(self, other):
        return self._apply_operator(other, "truediv")

    def __rtruediv__(self, other):
        return self.__truediv__(other) ** (-1)

    def __pow__(self, other):
        return self._apply_operator(other, "pow")

    def __rpow__(self, other):
        raise NotImplementedError


class Variable(_Operators, BrancherClass):
    """Abstract superclass of deterministic and random variables (variables.py:103-295)."""
    __hash__ = object.__hash__

    @property
    @abstractmethod
    def is_observed(self):
        pass

    def __str__(self):
        return self.name

    def __getitem__(self, key):
        if isinstance(key, str):
            idx = key
        elif isinstance(key, Iterable):
            idx = (slice(None), *key)
        else:
            idx = (slice(None), key)
        return PartialLink(vars={self}, links=set(), expr=Index(VarRef(self), idx), string="%s[%s]" % (self.name, key))

    def shape(self):
        return PartialLink(vars={self}, links=set(), expr=ShapeOf(VarRef(self)))

    def _get_entropy(self, input_values={}):
        """Analytic entropy where the distribution has one, else -log q (variables.py:156-162)."""
        if self.distribution.has_analytic_entropy:
            params = self._get_parameters_from_input_values(input_values)
            return sum_from_dim(self.distribution.get_entropy(**params), 2)
        return -self.calculate_log_probability(input_values, include_parents=False)

    def get_sample(self, number_samples, input_values={}):
        from brancher_b200.pandas_interface import reformat_sample_to_pandas
        formatted = _reformat_sampler_input(input_values, number_samples)
        raw = {self: self._get_sample(number_samples, resample=False, observed=self.is_observed,
                                      differentiable=False, input_values=formatted)[self]}
        self.reset()
        return reformat_sample_to_pandas(raw)


def _reformat_sampler_input(input_values, number_samples):
    return {var: tile_parameter(coerce_to_dtype(value, is_observed=var.is_observed), number_samples)
            for var, value in input_values.items()}


class RootVariable(Variable):
    """Constant / learnable parameter node (variables.py:298-381).  Learnable roots own an
    `nn.Parameter` of shape (1, 1, *event) inside a `ParameterModule`: the gradient sink."""

    def __init__(self, data, name, learnable=False, is_observed=False):
        self.name = name
        self.distribution = distributions.DeterministicDistribution()
        self._observed = is_observed
        self.parents = set()
        self.ancestors = set()
        self._type = "Deterministic"
        self.learnable = learnable
        self.link = None
        self._value = coerce_to_dtype(data, is_observed)
        if learnable:
            if is_discrete(data):
                self.learnable = False
                warnings.warn("Currently discrete parameters are not learnable. Learnable set to False")
            else:
                self._value = torch.nn.Parameter(self._value, requires_grad=True)
                self.link = ParameterModule(self._value)

    @property
    def value(self):
        return self.link() if self.learnable else self._value

    @property
    def is_observed(self):
        return self._observed

    def calculate_log_probability(self, values, reevaluate=True, for_gradient=False, normalized=True,
                                  include_parents=False):
        return torch.zeros((1, 1), device=config.device)

    def _get_parameters_from_input_values(self, input_values):
        return {"value": self.value}

    def _get_sample(self, number_samples, resample=False, observed=False, input_values={}, differentiable=True,
                    _memo=None):
        value = input_values[self] if self in input_values else self.value
        if is_discrete(value):
            return {self: value}
        return {self: tile_parameter(value, number_samples)}

    def reset(self, recursive=False):
        pass

    def _flatten(self):
        return []


class RandomVariable(Variable):
    """Random node: distribution + parents + link mapping parent values to distribution parameters
    (variables.py:384-622)."""

    def __init__(self, distribution, name, parents, link):
        self.name = name
        self.distribution = distribution
        self.link = link
        self.parents = parents
        self.ancestors = None
        self._type = "Random"
        self._init_state()

    def _init_state(self, is_observed=False):
        self._observed = is_observed
        self._observed_value = None
        self.dataset = None
        self.has_random_dataset = False
        self.has_observed_value = False

    @property
    def value(self):
        if self._observed:
            return self._observed_value
        raise AttributeError("RandomVariable has to be observed to receive value.")

    @property
    def is_observed(self):
        return self._observed

    # -- link application --------------------------------------------------------------------------
    def _apply_link(self, parents_values):
        """Parents (s|1, b|1, *event) are broadcast to (S, B), flattened to (S*B, *event), pushed through
        the link and un-flattened (variables.py:436-449)."""
        S, B = batch_sizes(parents_values)
        feed = {}
        for var, val in parents_values.items():
            if is_discrete(val) and not contains_tensors(val):
                feed[var] = val
            else:
                feed[var] = map_structure(lambda t: flatten_batch(t, S, B), val)
        out = self.link(feed)
        return {k: map_structure(lambda t: unflatten_batch(t, S, B), v) if (is_tensor(v) or contains_tensors(v)) else v
                for k, v in out.items()}

    def _get_parameters_from_input_values(self, input_values):
        S, _ = batch_sizes(input_values) if input_values else (1, 1)
        S = S or 1
        vals = {}
        for parent in self.parents:
            if parent in input_values:
                vals[parent] = input_values[parent]
            elif isinstance(parent, RootVariable) or parent._type == "Deterministic node":
                vals[parent] = parent._get_sample(S, input_values=input_values)[parent]
        return self._apply_link(vals)

    def _own_value(self, input_values):
        if self in input_values:
            return input_values[self]
        if self._type == "Deterministic node":
            return self._get_sample(1, input_values=input_values)[self]
        return self.value

    def calculate_log_probability(self, input_values, reevaluate=True, for_gradient=False, include_parents=True,
                                  normalized=True, _done=None):
        """Eager log-probability (variables.py:486-520): observed nodes are summed over the data axis."""
        done = set() if _done is None else _done
        if self in done and not reevaluate:
            return 0.
        done.add(self)
        value = self._own_value(input_values)
        params = self._get_parameters_from_input_values(input_values)
        log_prob = self.distribution.calculate_log_probability(value, **params)
        if self.is_observed:
            log_prob = log_prob.sum(dim=1, keepdim=True)
        if not include_parents:
            return log_prob
        parents_lp = 0.
        for parent in self.parents:
            if isinstance(parent, RandomVariable):
                parents_lp = parents_lp + parent.calculate_log_probability(input_values, reevaluate, for_gradient,
                                                                           normalized=normalized, _done=done)
        if is_tensor(parents_lp):
            log_prob, parents_lp = partial_broadcast(log_prob, parents_lp)
        return log_prob + parents_lp

    def _get_sample(self, number_samples=1, resample=True, observed=False, input_values={}, differentiable=True,
                    _memo=None):
        """Ancestral sampling with per-call memoisation (variables.py:527-570)."""
        memo = {} if _memo is None else _memo
        if self in memo and not resample:
            return {self: memo[self]}
        target = self
        if not observed:
            if self in input_values:
                return {self: input_values[self]}
        elif self.has_observed_value:
            return {self: self._observed_value}
        elif self.has_random_dataset:
            target = self.dataset
        out = {}
        for parent in target.parents:
            out.update(parent._get_sample(number_samples, resample, observed, input_values,
                                          differentiable=differentiable, _memo=memo))
        params = target._apply_link({p: out[p] for p in target.parents})
        sample = target.distribution.get_sample(differentiable=differentiable, **params)
        memo[self] = sample
        out[self] = sample
        return out

    def observe(self, data):
        from brancher_b200.pandas_interface import pandas_frame2value
        data = pandas_frame2value(data, self.name)
        if isinstance(data, RandomVariable):
            self.dataset = data
            self.has_random_dataset = True
        else:
            self._observed_value = coerce_to_dtype(data, is_observed=True)
            self.has_observed_value = True
        self._observed = True

    def unobserve(self):
        self._init_state(False)

    def reset(self, recursive=True):
        pass   # sampling / log-prob memoisation is per call here, not object state

    def _flatten(self):
        return sorted(list(self.ancestors) + [self], key=lambda v: v.name)


# ---------------------------------------------------------------------------------------------------
class ProbabilisticModel(BrancherClass):
    """Collection of variables (variables.py:625-881)."""

    def __init__(self, variables):
        for var in variables:
            if not isinstance(var, (RootVariable, RandomVariable, ProbabilisticModel)):
                raise ValueError("Invalid input type: {}".format(type(var)))
        self._input_variables = list(variables)
        self.variables = self.flatten()
        self.posterior_model = None
        self.posterior_sampler = None
        self.observed_submodel = None
        self.is_transformed = False
        self.diagnostics = {}
        self._plans = {}
        if all(var.is_observed for var in self._input_variables):
            self.observed_submodel = self
        else:
            self.update_observed_submodel()

    def __str__(self):
        return str(self.model_summary)

    @property
    def model_summary(self):
        from brancher_b200.pandas_interface import reformat_model_summary
        flat = self._flatten()
        return reformat_model_summary([[v._type, v.parents, v.is_observed] for v in flat], [v.name for v in flat],
                                      ["Distribution", "Parents", "Observed"])

    @property
    def is_observed(self):
        return all(var.is_observed for var in self._flatten())

    def _flatten(self):
        seen = {}
        for var in self._input_variables:
            for v in list(var.ancestors) + [var]:
                seen[id(v)] = v
        return sorted(seen.values(), key=lambda v: v.name)

    def observe(self, data):
        if hasattr(data, "columns"):
            from brancher_b200.pandas_interface import pandas_frame2value
            data = {name: pandas_frame2value(data, index=name) for name in data}
        if not isinstance(data, dict):
            raise ValueError("The input data should be either a dictionary of values or a pandas dataframe")
        for key, value in data.items():
            var = self.get_variable(key) if isinstance(key, str) else key
            if isinstance(var, RandomVariable):
                var.observe(value)

    def update_observed_submodel(self):
        self.observed_submodel = ProbabilisticModel([v for v in self._flatten() if v.is_observed])

    def set_posterior_model(self, model, sampler=None):
        self.posterior_model = PosteriorModel(posterior_model=model, joint_model=self)
        self._plans.clear()
        if sampler:
            if isinstance(sampler, ProbabilisticModel):
                self.posterior_sampler = PosteriorModel(sampler, joint_model=self)
            elif isinstance(sampler, Variable):
                self.posterior_sampler = PosteriorModel(ProbabilisticModel([sampler]), joint_model=self)
            elif isinstance(sampler, Iterable) and all(isinstance(s, (ProbabilisticModel, Variable)) for s in sampler):
                self.posterior_sampler = [PosteriorModel(ProbabilisticModel([s]) if isinstance(s, Variable) else s,
                                                         joint_model=self) for s in sampler]
            else:
                raise ValueError("The sampler should be ither a probabilistic model, a brancher variable or an "
                                 "iterable of variables and/or models")

    # -- eager API (sampling / log-prob of given values; NOT the ELBO hot path) ---------------------
    def calculate_log_probability(self, rv_values, for_gradient=False, normalized=True):
        done = set()
        total = 0.
        for var in self._input_variables:
            total = total + var.calculate_log_probability(rv_values, reevaluate=False, for_gradient=for_gradient,
                                                          normalized=normalized, _done=done) \
                if isinstance(var, RandomVariable) else total
        return total

    def _get_sample(self, number_samples, observed=False, input_values={}, differentiable=True):
        memo, out = {}, {}
        for var in self._input_variables:
            out.update(var._get_sample(number_samples=number_samples, resample=False, observed=observed,
                                       input_values=input_values, differentiable=differentiable, _memo=memo))
        out.update(input_values)
        return out

    def _get_entropy(self, input_values={}, for_gradient=True):
        if self.is_transformed:
            return -self.calculate_log_probability(input_values, for_gradient=for_gradient)
        return sum(sum_from_dim(var._get_entropy(input_values), 2) for var in self.variables)

    def get_sample(self, number_samples, input_values={}):
        from brancher_b200.pandas_interface import reformat_sample_to_pandas
        formatted = _reformat_sampler_input(input_values, number_samples)
        return reformat_sample_to_pandas(self._get_sample(number_samples, observed=False, input_values=formatted,
                                                          differentiable=False))

    def check_posterior_model(self):
        if not self.posterior_model:
            raise AttributeError("The posterior model has not been initialized.")

    def _get_posterior_sample(self, number_samples, input_values={}, differentiable=True):
        self.check_posterior_model()
        post = self.posterior_model._get_posterior_sample(number_samples=number_samples, input_values=input_values,
                                                          differentiable=differentiable)
        return self._get_sample(number_samples, input_values=post, differentiable=differentiable)

    def get_posterior_predictive(self, number_samples, input_values):
        """Batched posterior-predictive pass on the GPU for model pairs that lower to the BNN family: `input_values` maps the
        observed input variable to a batch of rows; returns dict(logits [S, B, C], samples [S, B], probs [B, C]).
        One launch sequence for the whole batch instead of one graph walk per image and sample
        (reference: the `_get_posterior_sample` loop of tests/test_MNIST_bayesian_neural_network.py:75-81)."""
        from brancher_b200 import lowering
        self.check_posterior_model()
        plan = lowering.get_plan(self, self.posterior_model)
        if not hasattr(plan, "predict"):
            raise lowering.UnsupportedModelError("posterior predictive is lowered for the BNN family only (plan: %s)" % plan.family)
        if plan.x_var not in input_values:
            raise KeyError("input_values must hold the model's input variable %r" % plan.x_var.name)
        return plan.predict(input_values[plan.x_var], number_samples)

    def get_posterior_sample(self, number_samples, input_values={}):
        from brancher_b200.pandas_interface import reformat_sample_to_pandas
        formatted = _reformat_sampler_input(input_values, number_samples)
        return reformat_sample_to_pandas(self._get_posterior_sample(number_samples, input_values=formatted))

    def get_p_log_probabilities_from_q_samples(self, q_samples, q_model, empirical_samples={}, for_gradient=False,
                                               normalized=True):
        p_samples = reassign_samples(q_samples, source_model=q_model, target_model=self)
        p_samples.update(empirical_samples)
        return self.calculate_log_probability(p_samples, for_gradient=for_gradient, normalized=normalized)

    # -- the hot path ------------------------------------------------------------------------------
    def estimate_log_model_evidence(self, number_samples, method="ELBO", input_values={}, for_gradient=False,
                                    posterior_model=(), gradient_estimator=None):
        """Monte-Carlo ELBO (variables.py:843-870).  The observed sub-model is sampled once (the
        minibatch), then the ELBO *and* its pathwise gradient are produced by ONE fused CUDA evaluation
        selected by `lowering.lower(joint, posterior)`; `for_gradient=False` returns the same value
        detached."""
        if not posterior_model:
            self.check_posterior_model()
            posterior_model = self.posterior_model
        if method != "ELBO":
            raise NotImplementedError("The requested estimation method is currently not implemented.")
        observed_now = [v for v in self._flatten() if v.is_observed]
        if self.observed_submodel is None or [id(v) for v in observed_now] != \
                [id(v) for v in self.observed_submodel._input_variables]:
            self.observed_submodel = ProbabilisticModel(observed_now)     # observe() was called after construction
        empirical_samples = self.observed_submodel._get_sample(1, observed=True, differentiable=False)
        function = gradient_estimators.ELBOFunction(self, posterior_model, empirical_samples)
        if for_gradient:
            estimator = (gradient_estimator or gradient_estimators.PathwiseDerivativeEstimator)(
                function, posterior_model, empirical_samples)
            return estimator(number_samples)
        with torch.no_grad():
            return gradient_estimators.PathwiseDerivativeEstimator(function, posterior_model,
                                                                   empirical_samples)(number_samples).detach()

    def reset(self):
        pass


class PosteriorModel(ProbabilisticModel):
    """q, with its by-name mapping onto the joint model (variables.py:884-907)."""

    def __init__(self, posterior_model, joint_model):
        super().__init__(sorted(posterior_model.variables, key=lambda v: v.name))
        self._input_variables = list(posterior_model._input_variables)
        self.variables = self.flatten()
        self.posterior_model = None
        self.model_mapping = get_model_mapping(self, joint_model)
        self._is_trained = False

    def posterior_sample2joint_sample(self, posterior_sample):
        return reassign_samples(posterior_sample, self.model_mapping)

    def _get_posterior_sample(self, number_samples, observed=False, input_values={}, differentiable=True):
        sample = self.posterior_sample2joint_sample(self._get_sample(number_samples, observed, input_values,
                                                                     differentiable=differentiable))
        sample.update(input_values)
        return sample


# ---------------------------------------------------------------------------------------------------
def var2link(var):
    """Lift a variable / number / array / tuple of variables to a PartialLink (variables.py:910-922)."""
    if isinstance(var, PartialLink):
        return var
    if isinstance(var, Variable):
        return PartialLink(vars={var}, links=set(), expr=VarRef(var), string=str(var))
    if isinstance(var, (numbers.Number, np.ndarray, torch.Tensor)):
        return PartialLink(vars=set(), links=set(), expr=Const(var), string=str(var))
    if isinstance(var, (tuple, list)) and all(isinstance(v, (Variable, PartialLink)) for v in var):
        parts = [var2link(v) for v in var]
        return PartialLink(vars=set().union(*[p.vars for p in parts]), links=set().union(*[p.links for p in parts]),
                           expr=TupleOf([p.expr for p in parts]), string=str(var))
    return var


class PartialLink(_Operators, BrancherClass):
    """Symbolic function of variables: `vars` it depends on, `links` (nn.Modules with parameters) it
    uses, and the expression tree `expr`; `fn(values)` evaluates it (variables.py:977-1072)."""
    __hash__ = object.__hash__

    def __init__(self, vars, fn=None, links=None, string="", expr=None):
        self.vars = set(vars)
        self.links = set(links or ())
        self.string = string
        self.expr = expr if expr is not None else Opaque(fn)
        self.fn = fn if fn is not None else (lambda values, _e=self.expr: evaluate(_e, values))

    def __str__(self):
        return self.string

    def _apply_operator(self, other, op):
        other = var2link(other)
        name = op if isinstance(op, str) else {v: k for k, v in _OP_FN.items()}.get(op, "?")
        return PartialLink(vars=self.vars | other.vars, links=self.links | other.links,
                           expr=Call(name, _OP_FN[name], [self.expr, other.expr]),
                           string="(" + str(self) + _OP_SYMBOL.get(name, "?") + str(other) + ")")

    def __getitem__(self, key):
        if isinstance(key, Iterable) and not isinstance(key, str) and all(isinstance(k, int) for k in key):
            idx = (slice(None), *key)
        elif isinstance(key, int):
            idx = (slice(None), key)
        else:
            idx = key
        return PartialLink(vars=self.vars, links=self.links, expr=Index(self.expr, idx),
                           string="%s[%s]" % (self.string, key))

    def shape(self):
        return PartialLink(vars=self.vars, links=self.links, expr=ShapeOf(self.expr))

    def _flatten(self):
        out = []
        for var in self.vars:
            out += var._flatten()
        return out + [self]
