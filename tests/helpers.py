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


def check_against_oracle(got_loss, got_grads, o32, o64, what=""):
    """The parity rule of this repo (stated in DESIGN.md §Parity):

    gate A  CUDA vs the fp64 evaluation of the reference's formulas:
            |cuda - o64| <= ATOL*scale + RTOL*|o64|,  scale = max|o64| over the tensor (1 for the loss)
    gate B  CUDA vs the fp32 evaluation (what the reference itself returns), allowing for the fp32
            evaluation's own measured distance from fp64:
            |cuda - o32| <= ATOL*scale + RTOL*|o32| + 2*max|o32 - o64|
    """
    l32, g32 = o32
    l64, g64 = o64
    assert_close(got_loss, l64, what + " loss vs fp64 oracle")
    assert_close(got_loss, l32, what + " loss vs fp32 oracle", atol=ATOL + 2 * abs(l32 - l64))
    for k in g64:
        a = np.asarray(got_grads[k], dtype=np.float64).reshape(g64[k].shape)
        # tensor scale; when the exact gradient is identically zero (tied prior + entropy) fp64 returns
        # pure rounding noise, so the fp32 evaluation's magnitude and an absolute floor bound the scale
        sc = max(np.abs(g64[k]).max(), np.abs(np.asarray(g32[k], dtype=np.float64)).max(), 1e-8)
        assert_close(a, g64[k], "%s grad %s vs fp64 oracle" % (what, k), scale=sc)
        noise = np.abs(np.asarray(g32[k], dtype=np.float64) - g64[k]).max()
        assert_close(a, g32[k], "%s grad %s vs fp32 oracle" % (what, k), atol=ATOL + 2 * noise / sc, scale=sc)


def vae_nets(g):
    """(enc, dec) dicts in the oracle's / C-ABI's layout from golden 'param_enc.*' / 'param_dec.*' entries."""
    nets = {"enc": {"W": [], "b": []}, "dec": {"W": [], "b": []}}
    for k in sorted(g["param"]):
        side, rest = k.split(".", 1)
        if rest.startswith(("W.", "b.")):
            nets[side][rest[0]].append((int(rest[2:]), g["param"][k]))
        else:
            nets[side][rest] = g["param"][k]
    for side in nets:
        for wb in ("W", "b"):
            nets[side][wb] = [a for _, a in sorted(nets[side][wb], key=lambda t: t[0])]
    return nets["enc"], nets["dec"]
