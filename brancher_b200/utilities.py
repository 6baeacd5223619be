"""Shape conventions of the model graph (the semantics of brancher/utilities.py, restated).

Every tensor is laid out (MC sample, datapoint, *event):  axis 0 = sample, axis 1 = datapoint
(brancher/variables.py:140-143).  Non-observed values are stored (1, 1, *shape); observed data
(N, ...) becomes (1, N, ...) padded to 4-D (utilities.py:223-254).
"""
from collections.abc import Iterable

import numpy as np
import torch

from brancher_b200 import config

DISCRETE_TYPES = (list, set, tuple, dict, str)


def is_tensor(x):
    return torch.is_tensor(x)


def is_discrete(x):
    return type(x) in DISCRETE_TYPES


def contains_tensors(x):
    if isinstance(x, dict):
        return all(is_tensor(v) for v in x.values())
    if isinstance(x, Iterable) and not is_tensor(x):
        return all(is_tensor(v) for v in x)
    return False


def coerce_to_dtype(data, is_observed=False):
    """number / ndarray / tensor -> fp32 tensor with the (sample, datapoint, *event) layout
    (utilities.py:223-254); python containers are "discrete" values and pass through."""
    if is_discrete(data):
        return data
    if torch.is_tensor(data):
        t = data.float()
    elif isinstance(data, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(data)).float()
    elif isinstance(data, (int, float, np.floating, np.integer)):
        t = torch.full((1, 1), float(data))
    elif hasattr(data, "values") and hasattr(data, "columns"):      # pandas DataFrame
        t = torch.as_tensor(np.asarray(data.values)).float()
    else:
        raise TypeError("Invalid input dtype {} - expected float, integer, np.ndarray, or torch var.".format(type(data)))
    if is_observed:
        t = t.unsqueeze(0)
        if t.dim() == 2:
            t = t.reshape(t.shape + (1, 1))
        elif t.dim() == 3:
            t = t.reshape(t.shape + (1,))
    else:
        t = t.unsqueeze(0).unsqueeze(0)
    return t.to(config.device)


def tile_parameter(t, number_samples):
    """(1, ...) -> (S, ...) along the sample axis; a view instead of the reference's physical
    .repeat (utilities.py:257-266) -- values are identical."""
    if t.shape[0] == number_samples:
        return t
    if t.shape[0] == 1:
        return t.expand((number_samples,) + tuple(t.shape[1:]))
    raise ValueError("The parameter cannot be broadcasted to the required number of samples")


def sum_from_dim(t, dim_index):
    if t.dim() <= dim_index:
        return t
    return t.sum(dim=tuple(range(dim_index, t.dim())))


def batch_sizes(values):
    """(number_samples, number_datapoints) = max over the leading two axes of all tensors."""
    S = B = None
    def visit(v):
        nonlocal S, B
        if is_tensor(v):
            S = v.shape[0] if S is None else max(S, v.shape[0])
            B = v.shape[1] if B is None else max(B, v.shape[1])
        elif isinstance(v, dict):
            [visit(x) for x in v.values()]
        elif isinstance(v, (list, tuple)):
            [visit(x) for x in v]
    for v in values.values() if isinstance(values, dict) else values:
        visit(v)
    return S, B


def flatten_batch(t, S, B):
    """(s|1, b|1, *event) -> (S*B, *event): how parents are fed to a link (variables.py:436-449)."""
    return t.expand((S, B) + tuple(t.shape[2:])).reshape((S * B,) + tuple(t.shape[2:]))


def unflatten_batch(t, S, B):
    return t.reshape((S, B) + tuple(t.shape[1:]))


def map_structure(fn, x):
    if is_tensor(x):
        return fn(x)
    if isinstance(x, dict):
        return {k: map_structure(fn, v) for k, v in x.items()}
    if isinstance(x, tuple):
        return tuple(map_structure(fn, v) for v in x)
    if isinstance(x, list):
        return [map_structure(fn, v) for v in x]
    return x


def broadcast_all(*tensors):
    """Align ranks by appending trailing singleton axes, collapse all-scalar events to (.,.,1,1), then
    broadcast (utilities.py:143-159,274-279)."""
    if all(int(np.prod(t.shape[2:])) == 1 for t in tensors):
        tensors = [t.reshape(tuple(t.shape[:2]) + (1, 1)) for t in tensors]
    rank = max(t.dim() for t in tensors)
    out = []
    for t in tensors:
        while t.dim() < rank:
            t = t.unsqueeze(-1)
        out.append(t)
    return torch.broadcast_tensors(*out)


def partial_broadcast(*tensors):
    s0 = max(t.shape[0] for t in tensors)
    s1 = max(t.shape[1] for t in tensors)
    return [t.expand((s0, s1) + tuple(t.shape[2:])) for t in tensors]


def get_model_mapping(source_model, target_model):
    """{source variable -> target variable} matched BY NAME (utilities.py:282-293) -- including the
    auto-created hyper-parameter roots, which is what ties p's prior parameters to q's."""
    targets = list(target_model.keys()) if isinstance(target_model, dict) else target_model._flatten()
    by_name = {v.name: v for v in source_model._flatten()}
    return {by_name[t.name]: t for t in targets if t.name in by_name}


def reassign_samples(samples, model_mapping=None, source_model=None, target_model=None):
    if not model_mapping:
        if source_model is None or target_model is None:
            raise ValueError("Either a model mapping or both source and target models have to be provided as input")
        model_mapping = get_model_mapping(source_model, target_model)
    return {model_mapping[k]: v for k, v in samples.items() if k in model_mapping}


def to_numpy(t):
    return t.detach().cpu().numpy()
