"""Shared helpers for parity tests: golden-fixture loading and the tolerance rule."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: ELBO and every parameter gradient within rtol 1e-5 / atol 1e-6 in fp32.
RTOL = 1e-5
ATOL = 1e-6


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {"raw": z, "eps": {}, "param": {}, "grad": {}}
    for k in z.files:
        for pre in ("eps", "param", "grad"):
            if k.startswith(pre + "_"):
                out[pre][k[len(pre) + 1:]] = z[k]
    return out


def mf_params(g, names):
    """{name: (mu, rho)} from golden 'param_<name>_loc/_scale' entries."""
    return {n: (g["param"][n + "_loc"], g["param"][n + "_scale"]) for n in names}


def assert_close(got, want, what, rtol=RTOL, atol=ATOL, scale=None):
    """|got - want| <= atol*scale + rtol*|want| elementwise.

    `scale` (default 1) lets gradient tensors use an absolute floor relative to the tensor's own
    magnitude: the reference's fp32 autograd result itself carries cancellation noise of that
    order (SURVEY §7 hard parts), measured in tests/test_oracle_golden.py against the fp64 oracle.
    """
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    s = 1.0 if scale is None else float(scale)
    err = np.abs(got - want)
    bound = atol * s + rtol * np.abs(want)
    bad = err > bound
    assert not bad.any(), "%s: max err %.3e (bound %.3e) at %s; %d/%d outside" % (
        what, err.max(), bound.flat[err.argmax()], np.unravel_index(err.argmax(), err.shape), bad.sum(), bad.size)
