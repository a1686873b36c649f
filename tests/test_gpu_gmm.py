"""brta_gmm_fit_batch / brta_gmm_predict against the oracle and against scikit-learn --
the code Gibbs.cluster calls (basicrta/gibbs.py:255-257; n_init=117 at gibbs.py:296)."""
import os

import numpy as np
import pytest

from oracle import gmm_oracle as G

pytestmark = pytest.mark.gpu


def cloud(seed, n_rows=600, comps=((0.7, 4.0), (0.25, 0.2), (0.05, 0.004))):
    """(log weight, log rate) samples shaped like the retained posterior rows of a residue."""
    rng = np.random.default_rng(seed)
    pts = []
    for w, r in comps:
        lw = np.log(w) + 0.05 / np.sqrt(w) * rng.standard_normal(n_rows)
        lr = np.log(r) + 0.03 / np.sqrt(w) * rng.standard_normal(n_rows) + 0.3 * (lw - np.log(w))
        pts.append(np.stack((lw, lr), axis=1))
    x = np.concatenate(pts)
    return x[rng.permutation(len(x))]


FIVE = ((0.6, 10.0), (0.25, 1.0), (0.1, 0.1), (0.04, 0.01), (0.01, 0.001))


def _aligned(means_a, means_b):
    """permutation that sorts both sets of means by their rate coordinate"""
    return np.argsort(means_a[:, 1]), np.argsort(means_b[:, 1])


def test_em_from_injected_parameters_equals_oracle_and_sklearn():
    """Parity of the EM proper: same initial parameters -> same iteration count, lower bound and parameters as
    the oracle (tolerance 1e-9: float64 with a different summation order and CUDA's exp/log) and as
    scikit-learn's own fit."""
    from sklearn.cluster import KMeans
    from sklearn.mixture import GaussianMixture

    from basicrta_b200 import gmm
    cases = [(cloud(1), 3), (cloud(2, 300), 2), (cloud(3), 4), (cloud(4, 250, FIVE), 5), (cloud(5, 400, FIVE), 7),
             (cloud(6, 500, FIVE), 11), (cloud(7, 40), 1)]
    init = np.zeros((len(cases), 1, gmm.KMAX, 6))
    inits = []
    for p, (x, k) in enumerate(cases):
        labels = KMeans(n_clusters=k, n_init=1, random_state=p).fit(x).labels_
        w0, mu0, cov0 = G.init_from_labels(x, labels, k)
        inits.append((w0, mu0, cov0))
        init[p, 0] = gmm.pack_params(w0, mu0, cov0)
    fits = gmm.fit_batch([c[0] for c in cases], [c[1] for c in cases], n_init=1, init_params=init)
    for (x, k), (w0, mu0, cov0), f in zip(cases, inits, fits):
        ref = G.em_fit(x, w0, mu0, cov0)
        assert f.error is None and f.n_iter == ref['n_iter'] and f.converged == ref['converged']
        assert abs(f.lower_bound - ref['lower_bound']) < 1e-9
        np.testing.assert_allclose(f.weights, ref['weights'], rtol=1e-8)
        np.testing.assert_allclose(f.means, ref['means'], rtol=1e-8)
        np.testing.assert_allclose(f.covariances, ref['covariances'], rtol=1e-7, atol=1e-13)
        sk = GaussianMixture(n_components=k, weights_init=w0 / w0.sum(), means_init=mu0,
                             precisions_init=np.linalg.inv(cov0)).fit(x)
        if k > 1:                                           # weights_init is renormalised by 1 + 1e-15 for sklearn
            assert f.n_iter == sk.n_iter_ and abs(f.lower_bound - sk.lower_bound_) < 1e-9
            np.testing.assert_allclose(f.means, sk.means_, rtol=1e-7)
            np.testing.assert_allclose(f.precisions_cholesky, sk.precisions_cholesky_, rtol=1e-6, atol=1e-10)


def blobs(seed, n=500, sep=1.2, sig=0.5, k=3):
    """overlapping clusters: EM needs many iterations"""
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.normal((sep * i, -0.7 * sep * i), sig, size=(n, 2)) for i in range(k)])


@pytest.mark.parametrize('seed,k,tol,max_iter', [(1, 3, 1e-8, 100), (2, 4, 1e-8, 100), (4, 5, 1e-9, 40)])
def test_long_em_runs_and_the_iteration_cap(seed, k, tol, max_iter):
    """~100 EM iterations on overlapping clusters, including runs that end at max_iter (scikit-learn's
    ConvergenceWarning): same iteration count, flag and parameters as the oracle."""
    from sklearn.cluster import KMeans

    from basicrta_b200 import gmm
    x = blobs(seed)
    labels = KMeans(n_clusters=k, n_init=1, random_state=0).fit(x).labels_
    w0, mu0, cov0 = G.init_from_labels(x, labels, k)
    ref = G.em_fit(x, w0, mu0, cov0, tol=tol, max_iter=max_iter)
    f = gmm.fit_batch([x], [k], n_init=1, tol=tol, max_iter=max_iter,
                      init_params=gmm.pack_params(w0, mu0, cov0)[None, None])[0]
    assert f.n_iter == ref['n_iter'] and f.converged == ref['converged']
    assert abs(f.lower_bound - ref['lower_bound']) < 1e-10
    np.testing.assert_allclose(f.means, ref['means'], atol=1e-9)
    np.testing.assert_allclose(f.covariances, ref['covariances'], atol=1e-9)
    np.testing.assert_allclose(f.weights, ref['weights'], atol=1e-10)


def test_device_initialisation_equals_its_restatement():
    """k-means++ seeding and Lloyd iterations on the Philox stream: the initial parameters of every restart
    equal the oracle's restatement (identical discrete decisions; sums in the same order), and so does the
    whole fit."""
    from basicrta_b200 import gmm
    # incl. the edges: one component, fewer points than threads, as many components as points
    problems = [cloud(11), cloud(12, 150, FIVE), cloud(13, 50), cloud(14, 43), cloud(15, 9)[:5], cloud(16, 9)[:2]]
    ks, ids, seed, n_init = [3, 5, 2, 1, 3, 2], [7, 123456, 99, 0, 4294967295, 17], 2024, 6
    fits, raw = gmm.fit_batch(problems, ks, n_init=n_init, seed=seed, problem_ids=ids, return_all=True)
    for p, (x, k) in enumerate(zip(problems, ks)):
        for r in range(n_init):
            _, labels = G.kmeans_init(x, k, seed, ids[p], r)
            w0, mu0, cov0 = G.init_from_labels(x, labels, k)
            got = raw['init'][p, r, :k]
            np.testing.assert_allclose(got[:, 0], w0, rtol=1e-12)
            np.testing.assert_allclose(got[:, 1:3], mu0, rtol=1e-11)
            np.testing.assert_allclose(got[:, 3], cov0[:, 0, 0], rtol=1e-9)
            np.testing.assert_allclose(got[:, 4], cov0[:, 0, 1], rtol=1e-8, atol=1e-14)
            np.testing.assert_allclose(got[:, 5], cov0[:, 1, 1], rtol=1e-9)
            ref = G.em_fit(x, w0, mu0, cov0)
            assert raw['n_iter'][p, r] == ref['n_iter']
            assert abs(raw['lower_bound'][p, r] - ref['lower_bound']) < 1e-9
        assert fits[p].restart == G.best_of(list(raw['lower_bound'][p]))


def test_fit_does_not_depend_on_the_batch():
    from basicrta_b200 import gmm
    problems = [cloud(21), cloud(22, 200, FIVE), cloud(23, 90)]
    ks, ids = [3, 5, 3], [5, 6, 7]
    together = gmm.fit_batch(problems, ks, n_init=9, seed=1, problem_ids=ids)
    alone = gmm.fit_batch([problems[1]], [5], n_init=9, seed=1, problem_ids=[6])[0]
    assert together[1].restart == alone.restart and together[1].lower_bound == alone.lower_bound
    assert np.array_equal(together[1].means, alone.means)
    again = gmm.fit_batch(problems, ks, n_init=9, seed=1, problem_ids=ids)
    assert all(np.array_equal(a.covariances, b.covariances) for a, b in zip(together, again))


@pytest.mark.parametrize('seed,comps,k', [(31, ((0.7, 4.0), (0.25, 0.2), (0.05, 0.004)), 3), (32, FIVE, 5),
                                          (33, ((0.8, 2.0), (0.2, 0.02)), 2)])
def test_best_of_117_restarts_reaches_sklearns_optimum(seed, comps, k):
    """The reference's call: n_init = 117.  Different random initialisations, same optimum: lower bounds within
    the EM stopping tolerance (1e-3 per iteration), the same clusters and the same labels."""
    from sklearn.mixture import GaussianMixture

    from basicrta_b200 import gmm
    x = cloud(seed, 1000, comps)
    sk = GaussianMixture(n_components=k, n_init=117, random_state=seed).fit(x)
    model = gmm.GaussianMixture(n_components=k, n_init=117, random_state=seed).fit(x)
    assert model.converged_ and abs(model.lower_bound_ - sk.lower_bound_) < 2e-3
    a, b = _aligned(model.means_, sk.means_)
    np.testing.assert_allclose(model.means_[a], sk.means_[b], atol=5e-3)
    np.testing.assert_allclose(model.weights_[a], sk.weights_[b], atol=2e-3)
    to_rank, sk_rank = np.empty(k, int), np.empty(k, int)
    to_rank[a], sk_rank[b] = np.arange(k), np.arange(k)
    assert np.mean(to_rank[model.predict(x)] == sk_rank[sk.predict(x)]) > 0.998


def test_predict_and_estimator_attributes():
    from basicrta_b200 import gmm
    x = cloud(41)
    model = gmm.GaussianMixture(n_components=3, n_init=4, random_state=5).fit(x)
    lab = model.predict(x)
    assert lab.dtype == np.int64 and np.array_equal(lab, G.predict(x, model.weights_, model.means_, model.covariances_))
    assert np.array_equal(model.fit_predict(x), lab)
    np.testing.assert_allclose(model.precisions_cholesky_, G.precision_cholesky(model.covariances_), rtol=1e-13)
    np.testing.assert_allclose(model.precisions_ @ model.covariances_, np.broadcast_to(np.eye(2), (3, 2, 2)), atol=1e-9)
    assert abs(model.weights_.sum() - 1) < 1e-12 and model.n_iter_ >= 1
    with pytest.raises(ValueError):
        gmm.GaussianMixture(n_components=2, covariance_type='diag')
    with pytest.raises(ValueError):
        gmm.fit_batch([x[:2]], [3])                         # fewer samples than components: sklearn's ValueError
    with pytest.raises(ValueError):
        gmm.fit_batch([x], [17])


def test_collapsed_component_is_reported_like_sklearns_value_error():
    from sklearn.mixture import GaussianMixture

    from basicrta_b200 import gmm
    x = np.concatenate([np.zeros((40, 2)), np.ones((40, 2))])              # two point masses, no regularisation
    with pytest.raises(ValueError):
        GaussianMixture(n_components=2, reg_covar=0.0, random_state=0).fit(x)
    f = gmm.fit_batch([x, cloud(51)], [2, 3], n_init=3, reg_covar=0.0, seed=0)
    assert isinstance(f[0].error, ValueError) and f[1].error is None       # failures stay per problem
    with pytest.raises(ValueError):
        gmm.GaussianMixture(n_components=2, reg_covar=0.0, random_state=0).fit(x)


def _short_chain(tmp_path, name, seed, n=4000):
    from basicrta_b200.gibbs import Gibbs
    from oracle import gibbs_oracle as O
    times = O.synth_times(n, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=seed)
    g = Gibbs(times, name, 0, ncomp=15, niter=20000, cutoff=7.0)
    g.burnin, g.seed = 5000, seed
    g.run()
    return g


def test_process_gibbs_with_the_device_mixture(tmp_path, monkeypatch):
    """Gibbs.process_gibbs end to end with the mixture fitted on the GPU against the same call with
    scikit-learn (gibbs.py:275-308): same number of clusters, same labels, same tau."""
    import copy

    from basicrta_b200 import postprocess
    monkeypatch.chdir(tmp_path)
    g = _short_chain(tmp_path, 'X5', 3)
    h = copy.deepcopy(g)
    g.process_gibbs(save=False, device=0, gmm_device=0)
    h.process_gibbs(save=False)
    assert g.processed_results.ncomp == h.processed_results.ncomp
    assert np.mean(g.processed_results.labels == h.processed_results.labels) > 0.995
    np.testing.assert_allclose(g.processed_results.indicator, h.processed_results.indicator, atol=0.02)
    tg, th = postprocess.estimate_tau(g), postprocess.estimate_tau(h)
    np.testing.assert_allclose([tg[0], tg[2]], [th[0], th[2]], rtol=0.03)      # 95 % interval of tau
    assert abs(tg[1] - th[1]) <= (th[2] - th[0]) / 5                            # mode: within a histogram bin


def test_reprocess_fans_every_residue_into_one_batch(tmp_path, monkeypatch):
    """ProcessProtein.reprocess (cluster.py:54-76): every residue directory is re-clustered and saved; a
    residue's result does not depend on which other residues were in the batch."""
    from basicrta_b200 import postprocess
    from basicrta_b200.cluster import ProcessProtein
    from basicrta_b200.gibbs import Gibbs
    monkeypatch.chdir(tmp_path)
    for i, name in enumerate(('A10', 'B2', 'C33')):
        _short_chain(tmp_path, name, 10 + i, n=2500)
    os.makedirs('basicrta-7.0/D4')                                          # a residue without results
    pp = ProcessProtein(20000, 'prot', 7.0)
    done = pp.reprocess(device=0, seed=77)
    assert done == ['B2', 'A10', 'C33']                                      # sorted by residue number
    assert pp.collect_results()['D4'] is None and pp.residues['B2'].endswith('B2/gibbs_20000.pkl')
    loaded = Gibbs.load('basicrta-7.0/A10/gibbs_20000.pkl')
    pr = loaded.processed_results
    assert pr.ncomp >= 2 and pr.indicator.shape == (2500, pr.ncomp) and len(pr.labels) == len(pr.rates)
    assert pr.parameters.shape == (pr.ncomp, 2) and pr.intervals.shape == (2, pr.ncomp, 2)
    taus, bars = pp.get_taus()                                               # cluster.py:97-127
    missing = list(pp.residues).index('D4')
    assert taus.shape == (4,) and bars.shape == (2, 4) and taus[missing] == 0 and np.all(np.delete(taus, missing) > 0)
    alone = Gibbs.load('basicrta-7.0/A10/gibbs_20000.pkl')
    postprocess.reprocess_batch([alone], device=0, seed=77, save=False)
    assert np.array_equal(alone.processed_results.labels, pr.labels)
    assert np.array_equal(alone.processed_results.parameters, pr.parameters)
