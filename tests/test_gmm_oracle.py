"""The Gaussian-mixture oracle (oracle/gmm_oracle.py) against scikit-learn itself -- the third-party code
that ``Gibbs.cluster`` calls (basicrta/gibbs.py:255-257, ``n_init=117`` at gibbs.py:296).  CPU only."""
import numpy as np
import pytest

from oracle import gmm_oracle as G

sklearn = pytest.importorskip('sklearn')
from sklearn.cluster import KMeans                      # noqa: E402
from sklearn.mixture import GaussianMixture             # noqa: E402


def posterior_cloud(seed, n_rows=600, comps=((0.7, 4.0), (0.25, 0.2), (0.05, 0.004))):
    """(log weight, log rate) samples shaped like the retained posterior rows of a 3-component residue."""
    rng = np.random.default_rng(seed)
    pts = []
    for w, r in comps:
        lw = np.log(w) + 0.05 / np.sqrt(w) * rng.standard_normal(n_rows)
        lr = np.log(r) + 0.03 / np.sqrt(w) * rng.standard_normal(n_rows) + 0.3 * (lw - np.log(w))
        pts.append(np.stack((lw, lr), axis=1))
    x = np.concatenate(pts)
    return x[rng.permutation(len(x))]


@pytest.mark.parametrize('seed,k', [(1, 3), (2, 2), (3, 4), (4, 6)])
def test_em_equals_sklearn_from_injected_parameters(seed, k):
    x = posterior_cloud(seed)
    labels = KMeans(n_clusters=k, n_init=1, random_state=seed).fit(x).labels_
    w0, mu0, cov0 = G.init_from_labels(x, labels, k)
    ref = GaussianMixture(n_components=k, n_init=1, weights_init=w0 / w0.sum(), means_init=mu0,
                          precisions_init=np.linalg.inv(cov0)).fit(x)
    got = G.em_fit(x, w0 / w0.sum(), mu0, cov0)
    assert got['n_iter'] == ref.n_iter_ and got['converged'] == ref.converged_
    assert abs(got['lower_bound'] - ref.lower_bound_) < 1e-10
    np.testing.assert_allclose(got['weights'], ref.weights_, rtol=1e-9)
    np.testing.assert_allclose(got['means'], ref.means_, rtol=1e-9)
    np.testing.assert_allclose(got['covariances'], ref.covariances_, rtol=1e-8, atol=1e-14)
    np.testing.assert_allclose(got['precisions_cholesky'], ref.precisions_cholesky_, rtol=1e-8, atol=1e-12)
    assert np.array_equal(G.predict(x, got['weights'], got['means'], got['covariances']), ref.predict(x))


def test_restart_loop_and_selection_equal_sklearn():
    """fit_predict's restart loop: one RandomState feeds the k-means of every restart; the first restart with
    the strictly largest lower bound wins."""
    x, k, n_init = posterior_cloud(9), 3, 6
    rs = np.random.RandomState(7)
    fits = []
    for _ in range(n_init):
        labels = KMeans(n_clusters=k, n_init=1, random_state=rs).fit(x).labels_
        fits.append(G.em_fit(x, *G.init_from_labels(x, labels, k)))
    best = fits[G.best_of([f['lower_bound'] for f in fits])]
    ref = GaussianMixture(n_components=k, n_init=n_init, random_state=np.random.RandomState(7)).fit(x)
    assert abs(best['lower_bound'] - ref.lower_bound_) < 1e-10 and best['n_iter'] == ref.n_iter_
    np.testing.assert_allclose(best['means'], ref.means_, rtol=1e-9)
    np.testing.assert_allclose(best['covariances'], ref.covariances_, rtol=1e-8, atol=1e-14)


def test_best_of_keeps_the_first_maximum():
    assert G.best_of([-3.0, -1.0, -1.0, -2.0]) == 1
    assert G.best_of([-np.inf, -np.inf]) == 1          # sklearn: `or max_lower_bound == -inf` takes each in turn


def test_ill_defined_covariance_raises_like_sklearn():
    with pytest.raises(ValueError):
        G.precision_cholesky(np.array([[[1.0, 2.0], [2.0, 1.0]]]))


def test_device_style_initialisation_is_a_sound_kmeans():
    """The Philox-seeded k-means++ / Lloyd restatement lands on the same partition as scikit-learn's k-means
    on well-separated clouds, and its restarts reach scikit-learn's best-of-117 lower bound."""
    x, k = posterior_cloud(5), 3
    centres, labels = G.kmeans_init(x, k, seed=3, problem=0, restart=0)
    ref = KMeans(n_clusters=k, n_init=5, random_state=0).fit(x)
    order = np.argsort(centres[:, 1]); ref_order = np.argsort(ref.cluster_centers_[:, 1])
    np.testing.assert_allclose(centres[order], ref.cluster_centers_[ref_order], rtol=1e-3, atol=1e-3)
    remap = np.empty(k, int); remap[order] = np.arange(k)
    ref_remap = np.empty(k, int); ref_remap[ref_order] = np.arange(k)
    assert np.mean(remap[labels] == ref_remap[ref.labels_]) > 0.999
    lbs = [G.fit_restart(x, k, seed=3, problem=0, restart=r)['lower_bound'] for r in range(8)]
    full = GaussianMixture(n_components=k, n_init=20, random_state=0).fit(x)
    assert abs(max(lbs) - full.lower_bound_) < 1e-3


def test_uniform53_is_a_unit_interval_stream():
    u = np.array([G.uniform53(11, p, r, d) for p in range(4) for r in range(8) for d in range(16)])
    assert u.min() >= 0.0 and u.max() < 1.0 and len(np.unique(u)) == len(u)
    assert abs(u.mean() - 0.5) < 0.06
