"""CPU: the C-ABI library builds, loads, and exports every symbol include/brancher_cuda.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    from brancher_b200._cuda import build
    path = build.build()
    assert os.path.exists(path)
    return ctypes.CDLL(path)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "brancher_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(brn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(library):
    names = _declared_symbols()
    assert "brn_bnn_elbo_fwd_bwd" in names and "brn_linear_elbo_fwd_bwd" in names
    for n in names:
        assert hasattr(library, n), "missing export " + n


def test_binding_covers_header(library):
    from brancher_b200 import _cuda
    assert sorted(_cuda.SYMBOLS) == _declared_symbols()
    assert _cuda.lib().brn_abi_version() == _cuda.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    from brancher_b200 import _cuda
    with pytest.raises(_cuda.BrancherCudaError):
        _cuda._ptr(torch.zeros(3), what="x")


def test_argument_errors_are_reported(library):
    from brancher_b200 import _cuda
    r = _cuda.sample_range(4)
    # NULL pointers -> status < 0 and a message, without touching a GPU
    st = _cuda.lib().brn_bnn_elbo_fwd_bwd(None, None, 1, 1, 1, 1, None, ctypes.byref(r), None, 0, 1, None, None)
    assert st < 0 and b"NULL" in _cuda.lib().brn_last_error()
    assert _cuda.lib().brn_bnn_workspace_bytes(1024, 784, 100, 10, 256) > 0
