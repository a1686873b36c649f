"""Host-side posterior processing of a finished chain, without the plotting.

The north star keeps this on the host ("process_gibbs and its GaussianMixture clustering,
tau estimates ... stay on the host").  It is the *consumer* of the sampler's arrays and is
what the statistical parity tests compare on: clustered weights / rates and the slowest
tau.  Restates, plot-free:

* ``Gibbs.cluster``          basicrta/gibbs.py:221-273
* ``Gibbs.process_gibbs``    basicrta/gibbs.py:275-308 (its ``mixture_and_plot`` call only
                             contributes the label re-sorting of basicrta/util.py:738-756)
* ``Gibbs._estimate_params`` basicrta/gibbs.py:667-689
* ``Gibbs.estimate_tau``     basicrta/gibbs.py:691-715
"""
import numpy as np

from .util import confidence_interval


def _filtered(gibbs):
    burn = gibbs.burnin // gibbs.g
    wcutoff = 10 / len(gibbs.times)
    weights, rates = gibbs.mcweights[burn:], gibbs.mcrates[burn:]
    keep = weights > wcutoff
    return burn, wcutoff, weights, rates, keep


def _mode(values):
    from scipy import stats
    return stats.mode(values).mode


def cluster_inputs(gibbs):
    """What ``Gibbs.cluster`` feeds the mixture model (gibbs.py:230-252): the training points (rows with the
    modal number of components above the weight cutoff), all retained points, their (row, component)
    indices and the modal count.  Both point sets are (weight, rate) pairs, *not* yet in log space."""
    burn, wcutoff, weights, rates, keep = _filtered(gibbs)
    lens = keep.sum(axis=1)
    lmode = int(_mode(lens))
    train_rows = lens == lmode
    train = np.stack((weights[train_rows][keep[train_rows]], rates[train_rows][keep[train_rows]]), axis=1)
    rows, comps = np.where(keep)
    data = np.stack((weights[rows, comps], rates[rows, comps]), axis=1)
    return train, data, rows, comps, lmode


def cluster(gibbs, method='GaussianMixture', device=None, gmm_device=None, **kwargs):
    """Cluster the retained (weight, rate) samples in log space and accumulate, per datum,
    how often its label fell in each cluster (gibbs.py:221-273).  ``device``: GPU index to run the
    accumulation on (``engine.pindicator_counts``, same integers); None keeps it in NumPy.
    ``gmm_device``: GPU index to fit the mixture on (:mod:`basicrta_b200.gmm`) instead of
    scikit-learn; only for ``method='GaussianMixture'``."""
    train, data, rows, comps, lmode = cluster_inputs(gibbs)
    if gmm_device is not None and method == 'GaussianMixture':
        from .gmm import GaussianMixture
        model = GaussianMixture(device=gmm_device, **kwargs)
    else:
        from sklearn import mixture
        model = getattr(mixture, method)(**kwargs)
    model.fit(np.log(train))
    labels = model.predict(np.log(data))
    return _accumulate(gibbs, labels, rows, comps, lmode, device)


def _accumulate(gibbs, labels, rows, comps, lmode, device):
    """gibbs.py:259-273: per-datum cluster membership frequencies from the stored label rows."""
    burn = gibbs.burnin // gibbs.g
    resident = getattr(gibbs, '_device_indicator', None)     # Gibbs.run(keep_indicator_on_device=True)
    if gibbs.indicator is not None:
        indicator = gibbs.indicator[burn:]
    elif resident is not None:                               # reduce on the GPU that holds the labels
        indicator, device = resident[burn:], resident.device.index
    else:
        indicator = gibbs._sample_indicator()
    pind = pindicator_counts_host(indicator, rows, comps, labels, lmode, gibbs.ncomp, device).astype(np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        pind = (pind.T / pind.sum(axis=1)).T
    gibbs.processed_results.indicator = pind
    gibbs.processed_results.labels = labels
    return labels


def pindicator_counts_host(indicator, rows, comps, labels, n_clusters, ncomp, device=None):
    """counts[i, c] = number of retained (row, component) pairs with mixture label c whose component
    labels datum i in that row (gibbs.py:264-268).  NumPy, or the GPU kernel if ``device`` is given."""
    if isinstance(indicator, np.ndarray):
        ncomp = max(int(ncomp), int(indicator.max(initial=0)) + 1)
    lut = np.full((indicator.shape[0], ncomp), -1, dtype=np.int8)
    lut[rows, comps] = labels
    if device is not None:
        from .engine import pindicator_counts
        return pindicator_counts(indicator, lut, n_clusters, device=device)
    counts = np.zeros((indicator.shape[1], n_clusters), dtype=np.int64)
    for lo in range(0, indicator.shape[0], 64):                                   # bounded scratch
        mapped = np.take_along_axis(lut[lo:lo + 64], indicator[lo:lo + 64].astype(np.int64), axis=1)
        for c in range(n_clusters):
            counts[:, c] += (mapped == c).sum(axis=0)
    return counts


def sort_labels(gibbs):
    """Order clusters by mean rate, fastest first, noise clusters (no datum assigned with
    probability >= ``_noise_cutoff``) last: basicrta/util.py:738-756."""
    _, _, weights, rates, keep = _filtered(gibbs)
    arates = rates[keep]
    labels = gibbs.processed_results.labels
    uniq = np.unique(labels)
    imaxs = gibbs.processed_results.indicator.max(axis=0)
    noise = np.where(imaxs < gibbs._noise_cutoff)[0]
    means = np.array([arates[labels == i].mean() for i in uniq])
    valid = np.delete(uniq, noise)
    vsorts = means[valid].argsort()[::-1]
    nsorts = means[noise].argsort()[::-1]
    presorts = np.concatenate([valid[vsorts], noise[nsorts]]).astype(int)
    sorts = np.array([np.where(presorts == i)[0][0] for i in uniq])
    return sorts[labels], presorts


def process_gibbs(gibbs, save=True, device=None, gmm_device=None, _labels=None):
    """gibbs.py:275-308 without figures.  ``device`` / ``gmm_device``: see :func:`cluster`; ``_labels``: mixture
    labels of the retained samples already fitted elsewhere (:func:`reprocess_batch`)."""
    burn, wcutoff, weights, rates, keep = _filtered(gibbs)
    rows, comps = np.where(keep)
    iteration = np.arange(gibbs.burnin, gibbs.niter + 1, gibbs.g)[rows] // gibbs.g
    lmode = int(_mode(keep.sum(axis=1)))

    if _labels is None:
        cluster(gibbs, n_init=117, n_components=lmode, device=device, gmm_device=gmm_device)
    else:
        _accumulate(gibbs, _labels, rows, comps, lmode, device)
    labels, presorts = sort_labels(gibbs)
    pr = gibbs.processed_results
    pr.labels = labels
    pr.indicator = pr.indicator[:, presorts]
    pr.weights, pr.rates = weights[rows, comps], rates[rows, comps]
    pr.ncomp, pr.residue, pr.iteration, pr.niter = lmode, gibbs.residue, iteration, gibbs.niter
    estimate_params(gibbs)
    if save:
        gibbs.save()


def reprocess_batch(gibbs_list, device=0, n_init=117, seed=None, save=True, pindicator_on_device=True):
    """``ProcessProtein.reprocess`` (basicrta/cluster.py:54-76: ``process_gibbs`` of every residue over a process
    pool) with the mixture fits of ALL residues in one launch: 117 restarts x every residue = one CTA each
    (``gmm.fit_batch``), one ``predict`` launch, then the per-residue tail of ``process_gibbs``.  A residue whose
    fit fails the way scikit-learn's would (``ValueError``) is skipped like cluster.py:46 does; returns the list
    of residues processed."""
    import zlib

    from . import gmm
    usable, inputs = [], []
    for g in gibbs_list:                                   # a residue that cannot be clustered fails alone
        try:
            inp = cluster_inputs(g)
            if len(inp[0]) < max(inp[4], 2):               # scikit-learn's "n_samples >= n_components" ValueError
                raise ValueError('too few retained samples')
        except (ValueError, IndexError) as err:
            print(f'{g.residue}: not clustered ({err})')
            continue
        usable.append(g)
        inputs.append(inp)
    gibbs_list = usable
    ids = [zlib.crc32(f'{g.residue}|{g.cutoff}'.encode()) & 0xFFFFFFFF for g in gibbs_list]
    fits = gmm.fit_batch([np.log(i[0]) for i in inputs], [i[4] for i in inputs], n_init=n_init, seed=seed,
                         device=device, problem_ids=ids)
    labels = gmm.predict_batch([np.log(i[1]) for i in inputs], fits, device=device)
    done = []
    for g, f, lab in zip(gibbs_list, fits, labels):
        if f.error is not None:
            continue
        process_gibbs(g, save=save, device=device if pindicator_on_device else None, _labels=lab)
        done.append(g)
    return done


def estimate_params(gibbs):
    """Mode-of-log-histogram estimates and 95 % intervals per cluster (gibbs.py:667-689)."""
    pr = gibbs.processed_results
    params, wb, rb = [], [], []
    for i in range(pr.ncomp):
        w, r = pr.weights[pr.labels == i], pr.rates[pr.labels == i]
        row = []
        for x in (w, r):
            bins = np.exp(np.linspace(np.log(x.min()), np.log(x.max()), 20))
            h = np.histogram(x, bins=bins)
            row.append(h[1][np.argmax(h[0])])
        params.append(row)
        wb.append(confidence_interval(w))
        rb.append(confidence_interval(r))
    pr.parameters = np.array(params)
    pr.intervals = np.array([np.array(wb), np.array(rb)])


def estimate_tau(gibbs):
    """[CI low, posterior mode, CI high] of tau = 1/rate of the slowest non-noise cluster
    (gibbs.py:691-715)."""
    pr = gibbs.processed_results
    imaxs = pr.indicator.max(axis=0)
    noise = np.where(imaxs < gibbs._noise_cutoff)[0]
    inds = np.delete(np.unique(pr.labels), noise)
    index = pr.parameters[inds, 1].argmin()
    taus = 1 / pr.rates[pr.labels == index]
    ci = confidence_interval(taus)
    h = np.histogram(taus, bins=15)
    imax = h[0].argmax()
    return [ci[0], 0.5 * (h[1][:-1][imax] + h[1][1:][imax]), ci[1]]
