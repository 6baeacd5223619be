"""Gradient-sink modules (mirror of brancher/modules.py:9-26)."""
from torch import nn


class ParameterModule(nn.Module):
    """Holds one learnable `nn.Parameter`; calling it returns the parameter (modules.py:9-18).
    `ProbabilisticOptimizer` collects these, and the fused kernels' gradients land in `.parameter.grad`."""

    def __init__(self, parameter):
        super().__init__()
        self.parameter = parameter

    def forward(self, *args, **kwargs):
        return self.parameter


class EmptyModule(nn.ModuleList):
    def __init__(self):
        super().__init__([])
