"""CPU oracle for the Gaussian-mixture clustering of the posterior samples (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the
product (``basicrta_b200.gmm``) never does and has no CPU path.

What it restates.  ``Gibbs.cluster`` (``/root/reference/basicrta/gibbs.py:221-257``) fits
``sklearn.mixture.GaussianMixture(n_init=117, n_components=lmode)`` (``gibbs.py:296``) to the retained
``(log weight, log rate)`` samples and predicts a label for every retained sample.  The arithmetic lives in a
third-party dependency that is not under ``/root/reference``: **scikit-learn**, unpinned in the reference's
``pyproject.toml`` (the authors' environment ``basicrta.yml`` lists 1.0.2); the build container has **1.9.0**,
whose ``sklearn/mixture/_base.py`` (``BaseMixture.fit_predict``, ``_initialize_parameters``, ``_e_step``) and
``sklearn/mixture/_gaussian_mixture.py`` (``_estimate_gaussian_parameters``,
``_estimate_gaussian_covariances_full``, ``_compute_precision_cholesky``, ``_estimate_log_gaussian_prob``)
define the algorithm restated here for ``covariance_type='full'`` in two dimensions:

* ``m_step``       weights / means / covariances from responsibilities (``nk = sum resp + 10 eps``, centred
                   covariances ``+ reg_covar`` on the diagonal), precision Cholesky factors;
* ``em_fit``       the EM loop of ``fit_predict``: E step, M step, ``lower_bound = mean(logsumexp)`` of the
                   parameters *before* the M step, stop when ``|change| < tol``; returned parameters are the
                   ones *after* the last M step;
* ``predict``      ``argmax_k`` of the weighted log probabilities;
* ``best_of``      the restart rule: the first restart with the strictly largest lower bound.

Pinned against scikit-learn itself in ``tests/test_gmm_oracle.py`` (same injected initial parameters ->
same weights, means, covariances, lower bound and iteration count).

The *initialisation* cannot be shared with scikit-learn: it draws k-means++ seeds from a NumPy
``RandomState`` that the reference never seeds (``random_state=None``).  The device draws its seeds from the
sampler's Philox stream instead; ``kmeans_init`` restates that procedure -- scikit-learn's greedy k-means++
(``sklearn/cluster/_kmeans.py::_kmeans_plusplus``, ``2 + int(log k)`` local trials) followed by Lloyd iterations
with scikit-learn's stopping rule (squared centre shift ``<= tol * mean(var(X))``) -- **in the device's own
order of floating-point operations** (thread t of 128 owns the points t, t + 128, ...; per-thread partial sums
are sequential; a block sum is a butterfly over the 32 lanes of each warp followed by the four warps in order),
so that every discrete decision (sampled candidate, chosen candidate, nearest centre) is reproduced exactly.
"""
import numpy as np

from . import philox

THREADS = 128
PURPOSE = 0x474D4D00          # Philox counter word 3 of the clustering stream ("GMM")
EPS10 = 10 * np.finfo(np.float64).eps
LOG_2PI = float(np.log(2 * np.pi))


# --------------------------------------------------------------------------------------
# scikit-learn's EM, restated

def precision_cholesky(cov):
    """Upper-triangular P with P P^T = cov^-1 for 2x2 covariances [K, 2, 2]; raises ValueError if a
    covariance is not positive definite (sklearn: "ill-defined empirical covariance")."""
    cxx, cxy, cyy = cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]
    if not np.all(cxx > 0):
        raise ValueError('ill-defined empirical covariance')
    l00 = np.sqrt(cxx)
    l10 = cxy / l00
    d = cyy - l10 * l10
    if not np.all(d > 0):
        raise ValueError('ill-defined empirical covariance')
    l11 = np.sqrt(d)
    p = np.zeros_like(cov)
    p[:, 0, 0] = 1.0 / l00
    p[:, 1, 1] = 1.0 / l11
    p[:, 0, 1] = -l10 / (l00 * l11)
    return p


def weighted_log_prob(X, weights, means, pchol):
    """[M, K]: log w_k + log N(x | mu_k, Sigma_k) (``_estimate_weighted_log_prob``)."""
    out = np.empty((X.shape[0], len(weights)))
    for k in range(len(weights)):
        d = X - means[k]
        y0 = d[:, 0] * pchol[k, 0, 0]
        y1 = d[:, 0] * pchol[k, 0, 1] + d[:, 1] * pchol[k, 1, 1]
        out[:, k] = -0.5 * (2 * LOG_2PI + (y0 * y0 + y1 * y1)) + (np.log(pchol[k, 0, 0]) + np.log(pchol[k, 1, 1])) \
            + np.log(weights[k])
    return out


def logsumexp(a):
    m = a.max(axis=1)
    return m + np.log(np.exp(a - m[:, None]).sum(axis=1))


def m_step(X, resp, reg_covar):
    """``_estimate_gaussian_parameters`` (full covariances); weights are the raw ``nk``."""
    nk = resp.sum(axis=0) + EPS10
    means = (resp.T @ X) / nk[:, None]
    cov = np.empty((len(nk), 2, 2))
    for k in range(len(nk)):
        d = X - means[k]
        cov[k] = (resp[:, k] * d.T) @ d / nk[k]
        cov[k, 0, 0] += reg_covar
        cov[k, 1, 1] += reg_covar
    return nk, means, cov


def init_from_labels(X, labels, n_components, reg_covar=1e-6):
    """``BaseMixture._initialize_parameters`` + ``GaussianMixture._initialize`` for ``init_params='kmeans'``:
    an M step on the one-hot responsibilities of the k-means labels, weights = nk / n_samples."""
    resp = np.zeros((X.shape[0], n_components))
    resp[np.arange(X.shape[0]), labels] = 1.0
    nk, means, cov = m_step(X, resp, reg_covar)
    return nk / X.shape[0], means, cov


def em_fit(X, weights, means, cov, tol=1e-3, reg_covar=1e-6, max_iter=100):
    """The EM loop of ``BaseMixture.fit_predict`` for one initialisation."""
    X = np.asarray(X, dtype=np.float64)
    weights, means, cov = np.array(weights, float), np.array(means, float), np.array(cov, float)
    pchol = precision_cholesky(cov)
    lower_bound, converged, n_iter = -np.inf, False, 0
    for n_iter in range(1, max_iter + 1):
        prev = lower_bound
        wlp = weighted_log_prob(X, weights, means, pchol)
        lse = logsumexp(wlp)
        resp = np.exp(wlp - lse[:, None])
        nk, means, cov = m_step(X, resp, reg_covar)
        weights = nk / nk.sum()
        pchol = precision_cholesky(cov)
        lower_bound = lse.mean()
        if abs(lower_bound - prev) < tol:
            converged = True
            break
    return dict(weights=weights, means=means, covariances=cov, precisions_cholesky=pchol,
                lower_bound=lower_bound, n_iter=n_iter, converged=converged)


def predict(X, weights, means, cov):
    return weighted_log_prob(np.asarray(X, float), weights, means, precision_cholesky(cov)).argmax(axis=1)


def best_of(lower_bounds):
    """``lower_bound > max_lower_bound or max_lower_bound == -inf`` over the restarts in order."""
    best, best_lb = 0, -np.inf
    for r, lb in enumerate(lower_bounds):
        if lb > best_lb or best_lb == -np.inf:
            best, best_lb = r, lb
    return best


# --------------------------------------------------------------------------------------
# the device's initialisation, in the device's order of operations

def uniform53(seed, problem, restart, draw):
    """Uniform double in [0, 1) number ``draw`` of (problem, restart): 53 bits from two Philox words."""
    k0, k1 = philox.seed_key(seed)
    w = philox.philox4x32_10(draw, restart, problem, PURPOSE, k0, k1)
    return float((int(w[0]) >> 5) * 67108864 + (int(w[1]) >> 6)) / 9007199254740992.0


def _layout(X):
    """[THREADS, J] view of the points: thread t owns i = t + THREADS j; padding is masked."""
    M = X.shape[0]
    J = -(-M // THREADS)
    idx = np.arange(THREADS)[:, None] + THREADS * np.arange(J)[None, :]
    valid = idx < M
    idx = np.where(valid, idx, 0)
    return X[idx, 0], X[idx, 1], valid, idx


def thread_sum(a, valid):
    """Sequential per-thread sums over j of a[t, j, ...] (masked)."""
    acc = np.zeros(a.shape[:1] + a.shape[2:])
    for j in range(a.shape[1]):
        v = valid[:, j].reshape((-1,) + (1,) * (a.ndim - 2))
        acc = np.where(v, acc + a[:, j], acc)
    return acc


def block_sum(per_thread):
    """Butterfly over the lanes of each warp (xor 16, 8, 4, 2, 1), then the warps in order."""
    v = per_thread.reshape((THREADS // 32, 32) + per_thread.shape[1:]).copy()
    lanes = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        v = v + v[:, lanes ^ off]
    total = v[0, 0]
    for w in range(1, THREADS // 32):
        total = total + v[w, 0]
    return total


def _dist2(x0, x1, c):
    d0, d1 = x0 - c[0], x1 - c[1]
    return d0 * d0 + d1 * d1


def _sample_position(closest, valid, idx, target):
    """First position, in thread-major order, whose running sum reaches ``target`` (the device's scan:
    sequential inside a thread, thread totals accumulated in order)."""
    within = np.zeros_like(closest)
    acc = np.zeros(closest.shape[0])
    for j in range(closest.shape[1]):
        acc = np.where(valid[:, j], acc + closest[:, j], acc)
        within[:, j] = acc
    totals = within[:, -1]
    prefix = np.zeros(THREADS)
    run = 0.0
    for t in range(THREADS):
        prefix[t] = run
        run = run + totals[t]
    cum = prefix[:, None] + within
    hit = (cum >= target) & valid
    if not hit.any():
        return int(idx[valid].max())
    t = int(np.argmax(hit.any(axis=1)))
    j = int(np.argmax(hit[t]))
    return int(idx[t, j])


def _scan_total(closest, valid):
    tot = thread_sum(closest, valid)
    run = 0.0
    for t in range(THREADS):
        run = run + tot[t]
    return run


def kmeans_init(X, n_components, seed, problem, restart, kmeans_tol=1e-4, kmeans_max_iter=300):
    """Centres and labels of the device's k-means initialisation of (problem, restart)."""
    X = np.asarray(X, dtype=np.float64)
    M, K = X.shape[0], int(n_components)
    x0, x1, valid, idx = _layout(X)
    # scikit-learn's tolerance: tol * mean of the per-feature variances
    mean = np.array([block_sum(thread_sum(x0, valid)), block_sum(thread_sum(x1, valid))]) / M
    var0 = block_sum(thread_sum((x0 - mean[0]) * (x0 - mean[0]), valid)) / M
    var1 = block_sum(thread_sum((x1 - mean[1]) * (x1 - mean[1]), valid)) / M
    tol = 0.5 * (var0 + var1) * kmeans_tol
    # greedy k-means++
    n_trials = 2 + int(np.log(K))
    centers = np.empty((K, 2))
    first = min(int(uniform53(seed, problem, restart, 0) * M), M - 1)
    centers[0] = X[first]
    closest = _dist2(x0, x1, centers[0])
    for c in range(1, K):
        pot = _scan_total(closest, valid)
        best_pot, best_d = np.inf, None
        for trial in range(n_trials):
            target = uniform53(seed, problem, restart, 8 * c + trial) * pot
            cand = _sample_position(closest, valid, idx, target)
            d = np.minimum(closest, _dist2(x0, x1, X[cand]))
            new_pot = block_sum(thread_sum(d, valid))
            if new_pot < best_pot:
                best_pot, best_d, best_cand = new_pot, d, cand
        centers[c] = X[best_cand]
        closest = best_d
    # Lloyd iterations
    for _ in range(kmeans_max_iter):
        labels = _nearest(x0, x1, centers)
        new = centers.copy()
        for k in range(K):
            m = (labels == k) & valid
            cnt = block_sum(thread_sum(m.astype(np.float64), valid))
            if cnt > 0:
                new[k, 0] = block_sum(thread_sum(np.where(m, x0, 0.0), valid)) / cnt
                new[k, 1] = block_sum(thread_sum(np.where(m, x1, 0.0), valid)) / cnt
        d = new - centers
        shift = 0.0
        for k in range(K):
            shift = shift + (d[k, 0] * d[k, 0] + d[k, 1] * d[k, 1])
        centers = new
        if shift <= tol:
            break
    labels = _nearest(x0, x1, centers)
    flat = np.empty(M, dtype=np.int64)
    flat[idx[valid]] = labels[valid]
    return centers, flat


def _nearest(x0, x1, centers):
    best = _dist2(x0, x1, centers[0])
    lab = np.zeros(x0.shape, dtype=np.int64)
    for k in range(1, len(centers)):
        d = _dist2(x0, x1, centers[k])
        closer = d < best
        best = np.where(closer, d, best)
        lab = np.where(closer, k, lab)
    return lab


def fit_restart(X, n_components, seed, problem, restart, tol=1e-3, reg_covar=1e-6, max_iter=100):
    """One (problem, restart) of the device batch: k-means initialisation + EM."""
    _, labels = kmeans_init(X, n_components, seed, problem, restart)
    w, mu, cov = init_from_labels(np.asarray(X, float), labels, n_components, reg_covar)
    return em_fit(X, w, mu, cov, tol=tol, reg_covar=reg_covar, max_iter=max_iter)
