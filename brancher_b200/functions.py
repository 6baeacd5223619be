"""`BF.<torch function>` lifts any torch function or nn.Module to a symbolic link (mirror of
brancher/functions.py:9-66).  Names are resolved lazily (module `__getattr__`) from
torch._C._VariableFunctions, then torch.nn.functional -- the reference's precedence
(functions.py:50-62) -- and each call records a `Call` node in the link's expression tree, which is
what the lowering pattern-matches (`BF.matmul`, `BF.tanh`, `BF.sigmoid`, `BF.softplus`, ...).
"""
import types

import torch

from brancher_b200.variables import var2link, Variable, PartialLink, Call, ModuleCall, Const, Expr

_MODULE_TYPES = (torch.nn.Module, torch.nn.Parameter, torch.nn.ParameterDict, torch.nn.ParameterList)


class BrancherFunction(object):
    """Wrapper on backend functions (torch) for the user interface."""

    def __init__(self, fn, name="f_?"):
        self.fn = fn
        self.name = name
        self.links = {fn} if isinstance(fn, _MODULE_TYPES) else set()

    def _get_string(self, *args, **kwargs):
        parts = []
        for a in list(args) + list(kwargs.values()):
            l = var2link(a)
            parts.append(l.string if isinstance(l, PartialLink) else str(l))
        return self.name + "(" + ", ".join(parts) + ")"

    def __call__(self, *args, **kwargs):
        def lift(a):
            l = var2link(a)
            return l if isinstance(l, PartialLink) else a
        largs = [lift(a) for a in args]
        lkwargs = {k: lift(a) for k, a in kwargs.items()}
        parts = [l for l in largs + list(lkwargs.values()) if isinstance(l, PartialLink)]
        vars_ = set().union(*[l.vars for l in parts]) if parts else set()
        links = set(self.links).union(*[l.links for l in parts]) if parts else set(self.links)
        node_cls = ModuleCall if self.links else Call
        expr = node_cls(self.name, self.fn, [l.expr if isinstance(l, PartialLink) else l for l in largs],
                        {k: (l.expr if isinstance(l, PartialLink) else l) for k, l in lkwargs.items()})
        return PartialLink(vars=vars_, links=links, expr=expr, string=self._get_string(*args, **kwargs))


def _resolve(name):
    for ns in (torch._C._VariableFunctions, torch.nn.functional):
        fn = getattr(ns, name, None)
        if fn is not None and isinstance(fn, (types.FunctionType, types.BuiltinFunctionType)):
            return fn
    return None


def __getattr__(name):
    if name.startswith("_"):
        raise AttributeError(name)
    fn = _resolve(name)
    if fn is None:
        raise AttributeError("brancher_b200.functions has no backend function %r" % name)
    bf = BrancherFunction(fn, name)
    globals()[name] = bf
    return bf


def _batch_meshgrid(tensor1, tensor2):
    """utilities.py:341-350"""
    assert tensor1.dim() == 2 and tensor2.dim() == 2, \
        "You can use batch_meshgrid only on 2D tensor (The first dimension is the batch dimension)"
    shape = [tensor1.shape[0], tensor1.shape[1], tensor2.shape[1]]
    return tensor1.unsqueeze(2).expand(*shape), tensor2.unsqueeze(1).expand(*shape)


batch_meshgrid = BrancherFunction(_batch_meshgrid, "batch_meshgrid")
delta = BrancherFunction(lambda x, y: (x == y).float(), "delta")
