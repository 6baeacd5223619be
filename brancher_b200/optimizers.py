"""ProbabilisticOptimizer: collects the learnable parameters of a model and drives torch.optim
(mirror of brancher/optimizers.py:19-73).  Gradients arrive in `ParameterModule.parameter.grad` from the
fused kernels' autograd node."""
from collections.abc import Iterable

import torch

from brancher_b200 import config
from brancher_b200.modules import ParameterModule, EmptyModule
from brancher_b200.standard_variables import LinkConstructor
from brancher_b200.variables import BrancherClass, Variable, ProbabilisticModel


class ProbabilisticOptimizer:
    def __init__(self, model, optimizer="SGD", **kwargs):
        assert isinstance(optimizer, str), "Optimizer should be a name of available pytoch optimizers"
        self.link_set = set()
        self.module = None
        self.optimizer = None
        self.setup(model, optimizer, **kwargs)

    def _update_link_set(self, model):
        assert isinstance(model, BrancherClass)
        variables = model.flatten() if isinstance(model, ProbabilisticModel) else (model.ancestors | {model})
        for var in variables:
            link = getattr(var, "link", None)
            if isinstance(link, (ParameterModule, LinkConstructor)):
                self.link_set.add(link)

    def add_variable2module(self, random_variable):
        self._update_link_set(random_variable)
        present = {id(m) for m in self.module}
        for link in sorted(self.link_set, key=id):
            mods = [link] if isinstance(link, ParameterModule) else list(link)
            for m in mods:
                if id(m) not in present:
                    self.module.append(m)
                    present.add(id(m))

    def setup(self, model, optimizer, **kwargs):
        self.module = EmptyModule()
        optimizer_class = getattr(torch.optim, optimizer)
        if isinstance(model, (Variable, ProbabilisticModel)):
            self.add_variable2module(model)
        elif isinstance(model, Iterable) and all(isinstance(m, (Variable, ProbabilisticModel)) for m in model):
            for m in model:
                self.add_variable2module(m)
        else:
            raise ValueError("Only brancher variables and iterable of variables can be added to a probabilistic optimizer")
        params = list(self.module.parameters())
        self.optimizer = optimizer_class(params, **kwargs) if params else None
        self.module.to(config.device)

    def update(self):
        self.optimizer.step()

    def zero_grad(self):
        self.optimizer.zero_grad()
