"""GPU: device-side minibatch index sampling (brn_minibatch_indices) -- the counterpart of the reference's host-side
np.random.choice(range(N), B, replace=False) (brancher/distributions.py:410-462).  The RNG differs from numpy's, so parity is
distributional: distinct ids in range, determinism in (seed, offset), uniform marginals, uniform first-slot distribution."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def cu():
    from brancher_b200 import _cuda
    _cuda.lib()
    return _cuda


@pytest.mark.parametrize("N,B", [(10, 1), (10, 5), (1000, 500), (10 ** 6, 1024), (10 ** 6, 8192), (2 ** 33, 4096), (7, 0)])
def test_distinct_in_range_and_deterministic(cu, N, B):
    a, rounds = cu.minibatch_indices(N, B, DEV, seed=3, offset=11, return_rounds=True)
    b = cu.minibatch_indices(N, B, DEV, seed=3, offset=11)
    c = cu.minibatch_indices(N, B, DEV, seed=3, offset=12)
    an = a.cpu().numpy()
    assert an.shape == (B,) and an.dtype == np.int64
    if B:
        assert an.min() >= 0 and an.max() < N and len(set(an.tolist())) == B
        assert torch.equal(a, b)                                   # pure function of (N, B, seed, offset)
        if B > 1 or N > 1000:
            assert not torch.equal(a, c)
        assert 1 <= int(rounds.item()) <= 64


def test_uniform_without_replacement(cu):
    """N = 40, B = 20 over 4000 independent draws: every row is included with probability B / N = 1/2 (binomial tolerance
    of 5 sigma), and the FIRST slot is uniform over the rows (chi-square against the uniform law)."""
    N, B, T = 40, 20, 4000
    draws = np.stack([cu.minibatch_indices(N, B, DEV, seed=1, offset=t).cpu().numpy() for t in range(T)])
    incl = np.zeros(N)
    for row in draws:
        incl[row] += 1
    sigma = np.sqrt(T * 0.25)
    assert np.abs(incl - T / 2).max() < 5 * sigma, incl
    first = np.bincount(draws[:, 0], minlength=N)
    chi2 = ((first - T / N) ** 2 / (T / N)).sum()
    assert chi2 < 90.0, chi2                     # 39 degrees of freedom: P(chi2 > 90) ~ 1e-5
    # pairs of slots: slot 1 given slot 0 is uniform over the remaining rows => P(slot1 < slot0) = 1/2
    lt = (draws[:, 1] < draws[:, 0]).mean()
    assert abs(lt - 0.5) < 5 * 0.5 / np.sqrt(T), lt


def test_argument_validation(cu):
    for N, B in [(10, 6), (0, 0), (10 ** 6, 9000), (5, -1)]:
        with pytest.raises(cu.BrancherCudaError):
            cu.minibatch_indices(N, B, DEV)


def test_empirical_variable_draws_on_device(cu):
    """EmpiricalVariable(batch_size=...) on a CUDA dataset: rows come from the device-side sampler (no host permutation),
    are distinct rows of the dataset, and change from draw to draw."""
    from brancher_b200 import config
    import model_zoo as zoo
    config.set_device("cuda:0")
    config.set_seed(5)
    ns = zoo.namespace("brancher_b200")
    N, F, B = 5000, 3, 64
    data = torch.arange(N * F, dtype=torch.float32).reshape(N, F)
    x = ns.EmpiricalVariable(data, name="x", batch_size=B, is_observed=True)
    s1 = x._get_sample(1)[x]
    s2 = x._get_sample(1)[x]
    assert s1.is_cuda and s1.shape[1] == B
    rows1 = (s1.reshape(B, F)[:, 0] / F).long().cpu().numpy()
    assert len(set(rows1.tolist())) == B and rows1.min() >= 0 and rows1.max() < N
    assert torch.equal(s1.reshape(B, F), data.to(DEV)[torch.as_tensor(rows1, device=DEV)])
    assert not torch.equal(s1, s2)
