"""Gaussian-mixture clustering of the posterior samples on the GPU.

``Gibbs.cluster`` fits ``sklearn.mixture.GaussianMixture(n_init=117, n_components=lmode)`` to the retained
``(log weight, log rate)`` samples of a residue and predicts a label for each of them
(basicrta/gibbs.py:255-257, 296); ``ProcessProtein.reprocess`` repeats that for every residue over a process
pool (basicrta/cluster.py:54-76).  This module is the device form: :func:`fit_batch` runs every
(residue, restart) pair as one CTA of ``brta_gmm_fit_batch`` and picks the restart by scikit-learn's rule;
:class:`GaussianMixture` wraps a batch of one behind the estimator's names (``fit``, ``predict``,
``weights_``, ``means_``, ``covariances_``, ``precisions_cholesky_``, ``converged_``, ``n_iter_``,
``lower_bound_``), so ``Gibbs.cluster(device=...)`` reads like the reference.

Only what the reference uses is implemented: ``covariance_type='full'``, ``init_params='kmeans'``, two
features, float64.  The k-means seeds come from the sampler's Philox stream -- the reference leaves
``random_state=None``, so there is no NumPy stream to reproduce; an integer ``random_state`` makes a fit
repeatable.  No CPU fallback: without the CUDA library the calls raise.
"""
import ctypes as C

import numpy as np

from . import _cabi

KMAX = _cabi.GMM_MAX_COMPONENTS


def _torch():
    import torch
    return torch


def _fresh_seed():
    return int(np.random.SeedSequence().entropy) & 0xFFFFFFFFFFFFFFFF


class GmmFit(object):
    """The selected restart of one problem."""
    __slots__ = ('weights', 'means', 'covariances', 'lower_bound', 'n_iter', 'converged', 'restart',
                 'lower_bounds', 'error')

    def __init__(self):
        self.error = None

    @property
    def precisions_cholesky(self):
        return _precisions_cholesky(self.covariances)


def _precisions_cholesky(cov):
    """Upper-triangular factors P with P P^T = cov^-1 (sklearn's ``precisions_cholesky_``), 2 x 2."""
    l00 = np.sqrt(cov[:, 0, 0])
    l10 = cov[:, 0, 1] / l00
    l11 = np.sqrt(cov[:, 1, 1] - l10 * l10)
    p = np.zeros_like(cov)
    p[:, 0, 0], p[:, 1, 1], p[:, 0, 1] = 1.0 / l00, 1.0 / l11, -l10 / (l00 * l11)
    return p


def _pack(problems):
    xs = [np.ascontiguousarray(x, dtype=np.float64) for x in problems]
    for x in xs:
        if x.ndim != 2 or x.shape[1] != 2:
            raise ValueError('every problem must be an [M, 2] array (log weight, log rate)')
    sizes = np.array([len(x) for x in xs], dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    flat = np.concatenate(xs) if xs else np.zeros((0, 2))
    return flat, offsets, sizes


def _unpack_params(block, k):
    """[16, 6] parameter block -> weights [k], means [k, 2], covariances [k, 2, 2]."""
    b = block[:k]
    cov = np.empty((k, 2, 2))
    cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 0], cov[:, 1, 1] = b[:, 3], b[:, 4], b[:, 4], b[:, 5]
    return b[:, 0].copy(), b[:, 1:3].copy(), cov


def pack_params(weights, means, covariances):
    """The inverse of :func:`_unpack_params`: one [16, 6] block."""
    k = len(weights)
    block = np.zeros((KMAX, 6))
    block[:k, 0] = weights
    block[:k, 1:3] = means
    block[:k, 3], block[:k, 4], block[:k, 5] = covariances[:, 0, 0], covariances[:, 0, 1], covariances[:, 1, 1]
    return block


def fit_batch(problems, n_components, n_init=1, tol=1e-3, reg_covar=1e-6, max_iter=100, seed=None, device=0,
              problem_ids=None, init_params=None, kmeans_tol=1e-4, kmeans_max_iter=300, return_all=False):
    """Fit one Gaussian mixture per problem, ``n_init`` restarts each, in one launch.

    ``problems``: list of [M_p, 2] float arrays; ``n_components``: int or one int per problem (1..16);
    ``problem_ids``: the Philox counter word of each problem (default: its position; the dispatcher passes a
    hash of the residue name so that a residue's fit does not depend on its place in the batch);
    ``init_params``: optional [P, n_init, 16, 6] initial parameters instead of the k-means initialisation.

    Returns one :class:`GmmFit` per problem.  A problem one of whose restarts lost positive definiteness has
    ``error`` set to the ``ValueError`` scikit-learn would have raised (basicrta/cluster.py:46 catches it per
    residue); with ``return_all`` the raw per-restart arrays are returned as well."""
    torch = _torch()
    lib = _cabi.load()
    dev = torch.device('cuda', int(device))
    flat, offsets, sizes = _pack(problems)
    P = len(sizes)
    ks = np.broadcast_to(np.asarray(n_components, dtype=np.int32), (P,)).copy()
    if P == 0:
        return ([], {}) if return_all else []
    if ks.min() < 1 or ks.max() > KMAX:
        raise ValueError(f'n_components must be in 1..{KMAX}')
    if np.any(sizes < np.maximum(ks, 2)):
        raise ValueError('Expected n_samples >= n_components (and >= 2) in every problem')      # sklearn's check
    if sizes.max() > _cabi.GMM_MAX_POINTS:
        raise ValueError(f'at most {_cabi.GMM_MAX_POINTS} points per problem')
    n_init = int(n_init)
    seed = _fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
    ids = np.arange(P, dtype=np.uint32) if problem_ids is None else np.asarray(problem_ids, dtype=np.uint32)
    with torch.cuda.device(dev):
        x_d = torch.from_numpy(flat).to(dev)
        off_d = torch.from_numpy(offsets).to(dev)
        k_d = torch.from_numpy(ks).to(dev)
        id_d = torch.from_numpy(ids).to(dev)
        lb_d = torch.empty((P, n_init), dtype=torch.float64, device=dev)
        it_d = torch.empty((P, n_init), dtype=torch.int32, device=dev)
        st_d = torch.empty((P, n_init), dtype=torch.int32, device=dev)
        par_d = torch.zeros((P, n_init, KMAX, 6), dtype=torch.float64, device=dev)
        init_out_d = torch.zeros((P, n_init, KMAX, 6), dtype=torch.float64, device=dev) if return_all else None
        init_d = None
        if init_params is not None:
            ip = np.ascontiguousarray(init_params, dtype=np.float64)
            if ip.shape != (P, n_init, KMAX, 6):
                raise ValueError(f'init_params must have shape {(P, n_init, KMAX, 6)}')
            init_d = torch.from_numpy(ip).to(dev)
        b = _cabi.GmmBatch()
        b.n_problems, b.n_init, b.max_iter, b.kmeans_max_iter = P, n_init, int(max_iter), int(kmeans_max_iter)
        b.max_points = int(sizes.max())
        b.class_mask = int((1 if ks.min() <= 4 else 0) | (2 if np.any((ks > 4) & (ks <= 8)) else 0) | (4 if ks.max() > 8 else 0))
        b.tol, b.reg_covar, b.kmeans_tol, b.seed = float(tol), float(reg_covar), float(kmeans_tol), seed
        b.x, b.offsets, b.n_components, b.problem_id = x_d.data_ptr(), off_d.data_ptr(), k_d.data_ptr(), id_d.data_ptr()
        b.init_params = init_d.data_ptr() if init_d is not None else None
        b.lower_bound, b.n_iter, b.status, b.params = lb_d.data_ptr(), it_d.data_ptr(), st_d.data_ptr(), par_d.data_ptr()
        b.init_out = init_out_d.data_ptr() if init_out_d is not None else None
        _cabi.check(lib.brta_gmm_fit_batch(C.byref(b), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                    'brta_gmm_fit_batch')
        lb, its, st = lb_d.cpu().numpy(), it_d.cpu().numpy(), st_d.cpu().numpy()
        # restart rule of BaseMixture.fit_predict: `lower_bound > max_lower_bound or max_lower_bound == -inf`
        # over the restarts in order = the first strictly largest lower bound
        best = np.argmax(np.where(np.isnan(lb), -np.inf, lb), axis=1)
        best_d = torch.from_numpy(best).to(dev)
        chosen = par_d[torch.arange(P, device=dev), best_d].cpu().numpy()
    fits = []
    for p in range(P):
        f = GmmFit()
        f.restart, f.lower_bounds = int(best[p]), lb[p].copy()
        f.weights, f.means, f.covariances = _unpack_params(chosen[p], int(ks[p]))
        f.lower_bound, f.n_iter = float(lb[p, best[p]]), int(its[p, best[p]])
        f.converged = bool(st[p, best[p]] == _cabi.GMM_CONVERGED)
        if np.any(st[p] == _cabi.GMM_ILL_DEFINED):
            f.error = ValueError('Fitting the mixture model failed because some components have ill-defined '
                                 'empirical covariance (for instance caused by singleton or collapsed samples). '
                                 'Try to decrease the number of components, increase reg_covar, or scale the input data.')
        fits.append(f)
    if return_all:
        raw = dict(lower_bound=lb, n_iter=its, status=st, params=par_d.cpu().numpy(), init=init_out_d.cpu().numpy())
        return fits, raw
    return fits


def predict_batch(problems, fits, device=0):
    """``GaussianMixture.predict`` for every problem with its own fitted parameters: one int64 label array per
    problem (gibbs.py:257)."""
    torch = _torch()
    lib = _cabi.load()
    dev = torch.device('cuda', int(device))
    flat, offsets, sizes = _pack(problems)
    P = len(sizes)
    if P == 0:
        return []
    out = []
    for lo in range(0, P, 65535):                            # gridDim.y
        hi = min(P, lo + 65535)
        ks = np.array([len(f.weights) for f in fits[lo:hi]], dtype=np.int32)
        blocks = np.stack([pack_params(f.weights, f.means, f.covariances) for f in fits[lo:hi]])
        sub_off = offsets[lo:hi + 1] - offsets[lo]
        sub = flat[offsets[lo]:offsets[hi]]
        if len(sub) == 0:
            out += [np.zeros(0, dtype=np.int64) for _ in range(lo, hi)]
            continue
        with torch.cuda.device(dev):
            x_d = torch.from_numpy(sub).to(dev)
            off_d = torch.from_numpy(np.ascontiguousarray(sub_off)).to(dev)
            k_d = torch.from_numpy(ks).to(dev)
            par_d = torch.from_numpy(blocks).to(dev)
            lab_d = torch.empty(len(sub), dtype=torch.uint8, device=dev)
            _cabi.check(lib.brta_gmm_predict(C.c_void_p(x_d.data_ptr()), C.c_void_p(off_d.data_ptr()), hi - lo,
                                             int(sizes[lo:hi].max()), C.c_void_p(k_d.data_ptr()),
                                             C.c_void_p(par_d.data_ptr()), C.c_void_p(lab_d.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'brta_gmm_predict')
            lab = lab_d.cpu().numpy().astype(np.int64)
        out += [lab[sub_off[i]:sub_off[i + 1]] for i in range(hi - lo)]
    return out


class GaussianMixture(object):
    """The subset of ``sklearn.mixture.GaussianMixture`` that ``Gibbs.cluster`` uses, on the GPU."""

    def __init__(self, n_components=1, *, covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100, n_init=1,
                 init_params='kmeans', random_state=None, device=0):
        if covariance_type != 'full' or init_params != 'kmeans':
            raise ValueError("basicrta_b200.gmm.GaussianMixture supports covariance_type='full', init_params='kmeans' "
                             '(what basicrta/gibbs.py:255-256 uses)')
        self.n_components, self.covariance_type, self.tol, self.reg_covar = n_components, covariance_type, tol, reg_covar
        self.max_iter, self.n_init, self.init_params, self.random_state, self.device = max_iter, n_init, init_params, random_state, device

    def fit(self, X, y=None):
        if self.random_state is not None and not isinstance(self.random_state, (int, np.integer)):
            raise ValueError('random_state must be None or an integer')
        f = fit_batch([X], self.n_components, n_init=self.n_init, tol=self.tol, reg_covar=self.reg_covar,
                      max_iter=self.max_iter, seed=self.random_state, device=self.device)[0]
        if f.error is not None:
            raise f.error
        self._fit = f
        self.weights_, self.means_, self.covariances_ = f.weights, f.means, f.covariances
        self.precisions_cholesky_ = f.precisions_cholesky
        self.precisions_ = np.einsum('kij,klj->kil', self.precisions_cholesky_, self.precisions_cholesky_)
        self.converged_, self.n_iter_, self.lower_bound_ = f.converged, f.n_iter, f.lower_bound
        if not f.converged and self.max_iter > 0:
            import warnings
            warnings.warn('Best performing initialization did not converge. Try different init parameters, or '
                          'increase max_iter, tol, or check for degenerate data.', RuntimeWarning)
        return self

    def predict(self, X):
        return predict_batch([X], [self._fit], device=self.device)[0]

    def fit_predict(self, X, y=None):
        return self.fit(X).predict(X)
