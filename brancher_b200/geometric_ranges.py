"""Constraint transforms of learnable parameters (mirror of brancher/geometric_ranges.py).

A numeric `scale=sigma` is stored as the unconstrained root rho = log(exp(sigma) - 1) and used as
softplus(rho) (RightHalfLine, geometric_ranges.py:48-57); probabilities are stored as logits
(Interval, :34-45).  The fused kernels apply the same transforms on chip.
"""
from abc import ABC, abstractmethod

import numpy as np

import brancher_b200.functions as BF


class GeometricRange(ABC):
    @abstractmethod
    def forward_transform(self, x, dim):
        pass

    @abstractmethod
    def inverse_transform(self, x, dim):
        pass


class UnboundedRange(GeometricRange):
    def forward_transform(self, x, dim):
        return x

    def inverse_transform(self, y, dim):
        return y


class Interval(GeometricRange):
    def __init__(self, lower_bound, upper_bound):
        self.lower_bound, self.upper_bound = lower_bound, upper_bound

    def forward_transform(self, x, dim):
        return self.lower_bound + (self.upper_bound - self.lower_bound) * BF.sigmoid(x)

    def inverse_transform(self, y, dim):
        z = (y - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return np.log(z / (1 - z))


class RightHalfLine(GeometricRange):
    def __init__(self, lower_bound):
        self.lower_bound = lower_bound

    def forward_transform(self, x, dim):
        return self.lower_bound + BF.softplus(x)

    def inverse_transform(self, y, dim):
        return np.log(np.exp(y - self.lower_bound) - 1)


class LeftHalfLine(GeometricRange):
    def __init__(self, upper_bound):
        self.upper_bound = upper_bound

    def forward_transform(self, x, dim):
        return self.upper_bound - BF.softplus(x)

    def inverse_transform(self, y, dim):
        return np.log(np.exp(-y + self.upper_bound) - 1)
