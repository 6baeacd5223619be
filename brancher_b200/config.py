"""Global device selection (mirror of brancher/config.py:10-30).

Unlike the reference, other modules read `config.device` at call time, so `set_device` may be called
after imports (the reference binds `device` at import, config.py / variables.py:44).
The default is cuda:0 when a GPU is visible (the ELBO hot path is CUDA-only), else cpu (model
construction, graph lowering and eager sampling still work there).
"""
import torch

seed = 0           # Philox key of the fused kernels
_iteration = 0     # Philox offset: bumped once per ELBO evaluation


def set_device(device_):
    global device
    if isinstance(device_, int):
        device = torch.device("cuda", device_)
        return
    if isinstance(device_, torch.device):
        device = device_
        return
    name = str(device_).lower()
    if "cuda" in name or name == "gpu":
        assert torch.cuda.is_available(), "Cuda requested but not available"
        device = torch.device("cuda:0" if name == "gpu" else name)
    elif name == "cpu":
        device = torch.device("cpu")
    else:
        raise ValueError("Device is not recongnized")


def set_seed(value):
    global seed, _iteration
    seed = int(value)
    _iteration = 0
    global _minibatch_draws
    _minibatch_draws = 0


_minibatch_draws = 0     # Philox offset of the device-side minibatch sampler: bumped once per index set


def next_minibatch_offset():
    global _minibatch_draws
    _minibatch_draws += 1
    return _minibatch_draws


def next_offset():
    global _iteration
    _iteration += 1
    return _iteration


device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
