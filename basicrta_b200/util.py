"""Host-side helpers that sit next to the sampler in the reference's ``basicrta/util.py``.

Only what the hot path and its pickle need is here; plotting and trajectory tools of the
reference are out of scope.
"""
import numpy as np


def get_bins(x, ts):
    """Histogram edges ts, 2 ts, ... covering ``x`` (basicrta/util.py:653-660)."""
    if isinstance(x, list):
        x = np.asarray(x)
    elif not isinstance(x, np.ndarray):
        raise TypeError('Input should be a list or array')
    return np.arange(1, int(x.max() // ts) + 3) * ts


def make_surv(ahist):
    """Empirical survival function from a histogram (basicrta/util.py:611-620): keep the
    non-empty bins, prepend t = 0, s = 1 - cumulative fraction."""
    counts, edges = ahist
    keep = counts != 0
    t = np.insert(edges[:-1][keep], 0, 0)
    y = np.insert(np.cumsum(counts[keep]), 0, 0)
    y = y / y[-1]
    return t, 1 - y


DENSE_BINS_MAX = 1 << 22


def get_s(x, ts):
    """``t, s`` stored in the Gibbs pickle (basicrta/util.py:116-120); not used by the sampler.

    The reference histograms over ``int(x.max() // ts) + 2`` edges ``k * ts`` and keeps the non-empty
    bins (util.py:611-620, 653-660) -- a billion bins (43 s, 8.6 GB) for continuous times whose
    first gap is tiny.  Beyond DENSE_BINS_MAX edges the same ``t, s`` are computed
    from the occupied bins only: bit-identical edges (``float64(k) * ts`` as ``np.arange(...) * ts``
    gives), same half-open bins with a closed last one."""
    x = np.asarray(x) if isinstance(x, list) else x
    if not isinstance(x, np.ndarray):
        raise TypeError('Input should be a list or array')
    n_edges = int(x.max() // ts) + 2
    if n_edges <= DENSE_BINS_MAX:
        return make_surv(np.histogram(x, bins=get_bins(x, ts)))
    ts = np.float64(ts)
    xs = np.sort(np.asarray(x, dtype=np.float64))
    k = np.floor(xs / ts).astype(np.int64)                    # candidate: edge k*ts <= x, fixed up below
    for _ in range(3):
        k = np.where(k.astype(np.float64) * ts > xs, k - 1, k)
        k = np.where((k + 1).astype(np.float64) * ts <= xs, k + 1, k)
    last = n_edges                                             # edges are k = 1 .. n_edges; the last bin is closed
    k = np.where(k >= last, last - 1, k)
    k = k[k >= 1]                                              # below the first edge: not counted
    occupied, counts = np.unique(k, return_counts=True)
    t = np.insert(occupied.astype(np.float64) * ts, 0, 0)
    y = np.insert(np.cumsum(counts), 0, 0)
    y = y / y[-1]
    return t, 1 - y


def confidence_interval(data, percentage=95):
    """Empirical central interval (basicrta/util.py:78-91)."""
    ds = np.sort(data)
    perc = np.arange(1, len(ds) + 1) / len(ds)
    lower = (100 - percentage) / 200
    upper = (percentage + (100 - percentage) / 2) / 100
    try:
        lo = ds[np.where(perc <= lower)[0][-1]]
        hi = ds[np.where(perc >= upper)[0][0]]
    except IndexError:
        lo, hi = ds[0], ds[-1]
    return [lo, hi]


def simulate_hn(n, weights, rates, seed=None, ts=None):
    """Hyper-exponential sample (basicrta/util.py:596-608), seeded; with ``ts`` the values
    are ceil-quantised to the trajectory grid as real contact durations are."""
    rng = np.random.default_rng(seed)
    n = int(n)
    weights = np.asarray(weights, dtype=float)
    comp = rng.choice(len(weights), size=n, p=weights / weights.sum())
    x = rng.exponential(1.0 / np.asarray(rates, dtype=float)[comp])
    if ts is not None:
        x = np.maximum(np.ceil(x / ts), 1.0) * ts
    x.sort()
    return x


def run_residue(residue, time, proc, ncomp, niter, cutoff):
    """Worker shim of the reference's pool (basicrta/util.py:475-485): build a ``Gibbs`` for
    one residue and run it.  ``proc`` selects the GPU (``proc % device_count``)."""
    from .gibbs import Gibbs
    x = np.array(time)
    gib = Gibbs(x, residue, proc, ncomp=ncomp, niter=niter, cutoff=cutoff)
    gib.run()


def get_bars(tau):
    """Error bars (distance of the 95 % bounds from the estimate) of ``[low, tau, high]`` rows: util.py:94-99."""
    tau = np.asarray(tau, dtype=np.float64)
    return np.array([tau[:, 1] - tau[:, 0], tau[:, 2] - tau[:, 1]])


def label_bits(ncomp):
    """Bits per label of the compact side-car: 1, 2, 4 or 8."""
    return 1 if ncomp <= 2 else 2 if ncomp <= 4 else 4 if ncomp <= 16 else 8


def pack_labels(indicator, ncomp):
    """uint8 label rows [S, N] -> [S, ceil(N / per)] bytes holding ``per = 8 // label_bits(ncomp)`` labels each,
    the first label of a group in the low bits (the opt-in compact form of ``indicator``)."""
    ind = np.ascontiguousarray(indicator, dtype=np.uint8)
    bits = label_bits(ncomp)
    if bits == 8:
        return ind.copy()
    if ind.size and int(ind.max()) >= (1 << bits):
        raise ValueError('a label does not fit the bit width of ncomp')
    per = 8 // bits
    rows, n = ind.shape
    padded = np.zeros((rows, -(-n // per) * per), dtype=np.uint8)
    padded[:, :n] = ind
    groups = padded.reshape(rows, -1, per)
    out = np.zeros(groups.shape[:2], dtype=np.uint8)
    for k in range(per):
        out |= groups[:, :, k] << np.uint8(bits * k)
    return out


def unpack_labels(packed, ncomp, n_data):
    """Inverse of :func:`pack_labels`."""
    packed = np.asarray(packed, dtype=np.uint8)
    bits = label_bits(ncomp)
    if bits == 8:
        return packed[:, :n_data].copy()
    per = 8 // bits
    out = np.empty((packed.shape[0], packed.shape[1], per), dtype=np.uint8)
    for k in range(per):
        out[:, :, k] = (packed >> np.uint8(bits * k)) & np.uint8((1 << bits) - 1)
    return out.reshape(packed.shape[0], -1)[:, :n_data].copy()
