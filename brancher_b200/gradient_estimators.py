"""Gradient estimators (mirror of brancher/gradient_estimators.py:17-56).

`PathwiseDerivativeEstimator(function, sampler, empirical_samples)(n_samples)` keeps the reference's
signature and returns a 0-d tensor supporting `.backward()`, but instead of sampling q and walking the
graph (gradient_estimators.py:39-44) it runs the fused CUDA evaluation the model pair lowers to.
BlackBoxEstimator / Taylor1Estimator are outside the hot path and raise.
"""
from abc import ABC, abstractmethod


class ELBOFunction:
    """What the reference passes as `function` (variables.py:851-855): log p(x, z) + H[q], here kept as
    a description (joint, posterior, minibatch) for the estimator to lower -- not a Python closure."""

    def __init__(self, joint_model, posterior_model, empirical_samples):
        self.joint_model = joint_model
        self.posterior_model = posterior_model
        self.empirical_samples = empirical_samples

    def __call__(self, samples):
        raise NotImplementedError("eager evaluation of the ELBO integrand is not part of brancher_b200: "
                                  "use PathwiseDerivativeEstimator (fused CUDA path)")


class GradientEstimator(ABC):
    def __init__(self, function, sampler, empirical_samples={}):
        self.function = function
        self.sampler = sampler
        self.empirical_samples = empirical_samples

    @abstractmethod
    def __call__(self, n_samples):
        pass


class PathwiseDerivativeEstimator(GradientEstimator):
    def __call__(self, n_samples):
        from brancher_b200 import lowering
        if not isinstance(self.function, ELBOFunction):
            raise NotImplementedError("PathwiseDerivativeEstimator needs the ELBO description built by "
                                      "ProbabilisticModel.estimate_log_model_evidence")
        plan = lowering.get_plan(self.function.joint_model, self.sampler)
        return plan.elbo(n_samples, self.empirical_samples)


class BlackBoxEstimator(GradientEstimator):
    def __call__(self, n_samples):
        raise NotImplementedError("BlackBoxEstimator is outside the fused hot path (SURVEY.md §8f)")


class Taylor1Estimator(GradientEstimator):
    def __call__(self, n_samples):
        raise NotImplementedError("Taylor1Estimator is outside the fused hot path (SURVEY.md §8f)")
