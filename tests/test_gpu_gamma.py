"""Direct distribution tests of the device Gamma / Dirichlet draws (``brta_gamma_fill``).

The posterior update of the sampler (basicrta/gibbs.py:210-211: ``rng.dirichlet(whypers + Ns)``,
``rng.gamma(1 + Ns, 1 / (3 + Ts))``) is drawn on the device by Marsaglia-Tsang trials on the Philox
stream, returned as log2 of the variate, with the shape < 1 boost applied in log2 space
(csrc/brta_math.cuh: trial_randoms / trial_finish / log2_gamma).  Teacher-forced parity tests bypass
these functions, so they are tested here on their own: one-sample Kolmogorov-Smirnov against
``scipy.stats.gamma`` / ``scipy.stats.beta`` for the shapes the sampler meets -- 1/30 and 1/15 (Dirichlet
shape of an empty component at K = 30 / 15), 1 (rate prior of an empty component), 0.5, 2.5, 101 and
1e5 + 1 (populated components) -- plus first and second moments.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

from basicrta_b200 import _cabi

pytestmark = pytest.mark.gpu
N_DRAWS = 200_000
P_MIN = 1e-3


def gamma_fill(shapes, n, chain=7, purpose=1, seed=12345):
    import torch
    lib = _cabi.load()
    sh = torch.tensor(np.asarray(shapes, dtype=np.float32), device='cuda:0')
    out = torch.empty(n, dtype=torch.float32, device='cuda:0')
    with torch.cuda.device(0):
        rc = lib.brta_gamma_fill(C.c_void_p(out.data_ptr()), n, C.c_void_p(sh.data_ptr()), len(shapes), chain, purpose,
                                 seed, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _cabi.check(rc, 'brta_gamma_fill')
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)


@pytest.mark.parametrize('shape', [1.0 / 30, 1.0 / 15, 0.5, 1.0, 2.5, 101.0, 100001.0])
def test_log2_gamma_matches_scipy(shape):
    l2g = gamma_fill([shape], N_DRAWS, chain=int(shape * 1000) % 97 + 1)
    assert np.all(np.isfinite(l2g))
    # KS in log2 space: P(log2 G <= x) = F_gamma(2^x); tiny shapes put most of the mass below float32's range
    # in linear space, which is exactly why the sampler works with log2 G
    with np.errstate(over='ignore', under='ignore'):
        res = stats.kstest(l2g, lambda x: stats.gamma.cdf(np.exp2(x), shape))
    assert res.pvalue > P_MIN, (shape, res)
    if shape >= 0.5:                                           # moments where linear space is representable
        g = np.exp2(l2g)
        se = np.sqrt(shape / N_DRAWS)
        assert abs(g.mean() - shape) < 5 * se, (shape, g.mean())
        assert abs(g.var() / shape - 1) < 0.03, (shape, g.var())
    else:                                                      # E[ln G] = digamma(a), Var[ln G] = trigamma(a)
        from scipy.special import digamma, polygamma
        lng = l2g * np.log(2.0)
        se = np.sqrt(polygamma(1, shape) / N_DRAWS)
        assert abs(lng.mean() - digamma(shape)) < 5 * se, (shape, lng.mean(), digamma(shape))
        assert abs(lng.var() / polygamma(1, shape) - 1) < 0.03


def test_streams_are_independent_of_each_other():
    """Different (chain, purpose) words give unrelated variates; the same words repeat bit for bit."""
    a = gamma_fill([3.0], 50_000, chain=1, purpose=1)
    b = gamma_fill([3.0], 50_000, chain=1, purpose=2)
    c = gamma_fill([3.0], 50_000, chain=2, purpose=1)
    again = gamma_fill([3.0], 50_000, chain=1, purpose=1)
    assert np.array_equal(a, again)
    for x in (b, c):
        assert abs(np.corrcoef(a, x)[0, 1]) < 0.02
    lag = np.corrcoef(a[:-1], a[1:])[0, 1]                     # consecutive counters
    assert abs(lag) < 0.02


@pytest.mark.parametrize('K,counts', [(15, [4000, 700, 290, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]),
                                      (30, [12000, 5000, 2000, 800, 200] + [0] * 25)])
def test_dirichlet_marginals(K, counts):
    """Weights = normalised gammas with shapes 1/K + n_k (gibbs.py:173, 210): component k is
    Beta(a_k, sum a - a_k).  Populated components are tested in linear space, empty ones through
    log2 w_k (their weights are ~1e-10 and below)."""
    shapes = 1.0 / K + np.asarray(counts, dtype=np.float64)
    n_sets = 40_000
    l2g = gamma_fill(shapes, n_sets * K, chain=K).reshape(n_sets, K)
    mx = l2g.max(axis=1, keepdims=True)
    l2w = l2g - mx - np.log2(np.exp2(l2g - mx).sum(axis=1, keepdims=True))
    a0 = shapes.sum()
    for k in (0, 1, 2, 3, K - 1):
        a = shapes[k]
        with np.errstate(over='ignore', under='ignore'):
            res = stats.kstest(l2w[:, k], lambda x: stats.beta.cdf(np.exp2(x), a, a0 - a))
        assert res.pvalue > P_MIN, (K, k, res)
    w = np.exp2(l2w)
    assert np.allclose(w.sum(axis=1), 1.0, atol=1e-5)
    assert np.abs(w.mean(axis=0) - shapes / a0).max() < 5e-4
