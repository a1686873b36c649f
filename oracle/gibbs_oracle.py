"""CPU oracle for the exponential-mixture Gibbs sweep (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  It is the *checker*, never the
product: ``basicrta_b200`` has no CPU path and fails loudly without its CUDA library.

What it restates (all citations relative to ``/root/reference/``):

* ``basicrta/gibbs.py:186-188``  deterministic initial weights / rates       -> :func:`init_state`
* ``basicrta/gibbs.py:196-197``  responsibilities  w_k r_k exp(-r_k t_i)     -> :func:`responsibilities_f64`,
                                                                               :func:`draw_indicators_f32`
* ``basicrta/gibbs.py:200``      one categorical draw per datum               -> :func:`draw_indicators_f32`
* ``basicrta/gibbs.py:203-207``  N_k and T_k = sum of times per component     -> :func:`sufficient_stats`
* ``basicrta/gibbs.py:210-211``  Dirichlet weights / Gamma rates              -> :func:`posterior_numpy`
* ``basicrta/gibbs.py:214-217``  thinning: row j//g-1 gets post-update (w, r)
                                 and the indicators drawn in iteration j      -> both ``run_*`` functions
* ``basicrta/gibbs.py:167-174``  output shapes / dtypes / hyper-parameters    -> :func:`allocate`

Two oracles live here.

``run_reference_order``
    fp64, linear space, NumPy ``Generator`` calls in the reference's order.  With the
    same seeded generator it reproduces the real reference bit for bit; that is how the
    restatement is pinned (``tests/golden/make_golden.py`` ran the unmodified reference
    in the build container, ``tests/test_oracle_golden.py`` replays it).  It is also the
    CPU baseline that ``bench.py`` times, because it executes the reference's arithmetic.

``run_teacher_forced``
    the bit-exact target for the CUDA kernel's EXACT mode: fp32 log2-space logits,
    max-subtracted, a fixed IEEE-only exp2 (``soft_exp2``), a sequential cumulative sum,
    one Philox uniform per datum, inverse-CDF in component order, integer tick sums.
    The posterior draw comes from a NumPy generator and is handed to the kernel as
    per-iteration coefficient rows (``coef_c``, ``coef_a``), so one flipped indicator can
    never cascade ("teacher forcing").  Replacing NumPy's conditional-binomial
    ``multinomial`` by a single-uniform inverse CDF is a deliberate, distribution-
    preserving deviation; the two oracles are tied together statistically in
    ``tests/test_oracle_statistics.py``.

Parity status: the reference's own tests pin nothing on this path (its only sampler
test is commented out, ``basicrta/tests/test_functions.py:6-40``); parity is pinned by
outputs of the reference itself generated in the build container (``tests/golden``).
"""
import numpy as np

from . import philox

LOG2E = 1.4426950408889634
_F = np.float32

# Taylor coefficients of 2**f = exp(f ln 2), rounded to float32, degree 7.
# |f| <= 0.5  ->  truncation error < 4e-9 relative, below float32 rounding.
_LN2 = 0.6931471805599453
EXP2_COEF = tuple(_F(_LN2 ** n / float(np.prod(np.arange(1, n + 1)) if n else 1.0))
                  for n in range(8))
_MAGIC = _F(12582912.0)          # 1.5 * 2**23: float32 add/sub rounds to nearest integer
_EXP2_FLOOR = _F(-125.0)         # below this the term is flushed to exactly 0


# --------------------------------------------------------------------------------------
# deterministic pieces
# --------------------------------------------------------------------------------------
def init_state(ncomp):
    """Initial (weights, rates): basicrta/gibbs.py:186-188."""
    rates = (0.5 * 10 ** np.arange(-ncomp + 2, 2, dtype=float))[::-1]   # 5, 0.5, ..., 5e-(K-1)
    unnorm = 9 * 10 ** (-np.arange(1, ncomp + 1, dtype=float))          # 0.9, 0.09, ...
    return unnorm / unnorm.sum(), rates


def allocate(niter, g, n, ncomp):
    """Output arrays and hyper-parameters: basicrta/gibbs.py:167-174."""
    rows = (niter + 1) // g
    return dict(indicator=np.zeros((rows, n), dtype=np.uint8),
                mcweights=np.zeros((rows, ncomp)),
                mcrates=np.zeros((rows, ncomp)),
                whypers=np.ones(ncomp) / ncomp,
                rhypers=np.ones((ncomp, 2)) * [1.0, 3.0])


def time_step(times):
    """``ts`` = first non-zero gap of the sorted times: basicrta/gibbs.py:147-151."""
    srt = np.sort(times)
    diff = srt[1:] - srt[:-1]
    nz = diff[diff != 0]
    return nz[0] if len(nz) else times.min()


def to_ticks(times, ts):
    """Integer multiples of ``ts`` (contact durations are frame counts, contacts.py:222-229)."""
    ticks = np.rint(np.asarray(times, dtype=np.float64) / float(ts)).astype(np.int64)
    return ticks


def responsibilities_f64(times, weights, rates):
    """z[i, k] of basicrta/gibbs.py:196-197 (fp64, linear space, no underflow guard)."""
    tmp = weights * rates * np.exp(np.outer(-rates, times)).T
    return (tmp.T / tmp.sum(axis=1)).T


def sufficient_stats(s, ticks, ncomp):
    """(N_k, sum of ticks with label k): basicrta/gibbs.py:203-207 with integer tick sums."""
    s = np.asarray(s, dtype=np.int64)
    nk = np.bincount(s, minlength=ncomp).astype(np.int64)
    tk = np.zeros(ncomp, dtype=np.int64)
    np.add.at(tk, s, np.asarray(ticks, dtype=np.int64))
    return nk, tk


# --------------------------------------------------------------------------------------
# oracle O1': the reference's update order on a NumPy generator (fp64)
# --------------------------------------------------------------------------------------
def run_reference_order(times, ncomp, niter, rng, g=100, whypers=None, rhypers=None):
    """fp64 sweep consuming ``rng`` exactly as basicrta/gibbs.py:191-217 does.

    Stream order per iteration: N categorical draws (``multinomial``), K Dirichlet
    gammas (``dirichlet``), K rate gammas (``gamma``).
    """
    times = np.asarray(times, dtype=np.float64)
    out = allocate(niter, g, len(times), ncomp)
    wh = out['whypers'] if whypers is None else np.asarray(whypers, float)
    rh = out['rhypers'] if rhypers is None else np.asarray(rhypers, float)
    weights, rates = init_state(ncomp)
    comps = np.arange(ncomp)
    for j in range(1, niter + 1):
        z = responsibilities_f64(times, weights, rates)
        s = rng.multinomial(1, z).argmax(axis=1)
        onehot = s[None, :] == comps[:, None]
        nk = onehot.sum(axis=1)
        tk = np.array([times[row].sum() for row in onehot])
        weights = rng.dirichlet(wh + nk)
        rates = rng.gamma(rh[:, 0] + nk, 1.0 / (rh[:, 1] + tk))
        if j % g == 0:
            row = j // g - 1
            out['mcweights'][row], out['mcrates'][row] = weights, rates
            out['indicator'][row] = s
    return out


# --------------------------------------------------------------------------------------
# oracle O2: the device arithmetic, IEEE-only (fp32), bit-exact target of EXACT mode
# --------------------------------------------------------------------------------------
def coefficients(weights, rates, ts):
    """(coef_c, coef_a) float32 rows: logit_k(tick) = coef_c[k] - coef_a[k]*tick  (log2 units).

    coef_c = log2(w_k r_k), coef_a = r_k * ts * log2(e); computed in fp64, rounded once.
    A component with zero weight or rate gets coef_c = -inf (never drawn).
    """
    w = np.asarray(weights, dtype=np.float64)
    r = np.asarray(rates, dtype=np.float64)
    with np.errstate(divide='ignore'):
        c = np.log2(w) + np.log2(r)
    a = r * (float(ts) * LOG2E)
    return c.astype(np.float32), a.astype(np.float32)


def soft_exp2(x):
    """2**x for x <= 0 using only float32 add/mul and integer exponent arithmetic.

    Mirrors ``brta_soft_exp2`` in basicrta_b200/csrc/brta_math.cuh operation by operation:
    x < -125 (or -inf/NaN) -> 0; n = rint(x) by the magic-number add; f = x - n in
    [-0.5, 0.5]; degree-7 Horner with separate multiply and add; exponent insert by
    integer add of n << 23.
    """
    x = np.asarray(x, dtype=np.float32)
    alive = x >= _EXP2_FLOOR                      # False for NaN and -inf too
    xc = np.where(alive, x, _F(0.0)).astype(np.float32)
    z = (xc + _MAGIC).astype(np.float32)
    nf = (z - _MAGIC).astype(np.float32)
    f = (xc - nf).astype(np.float32)
    p = np.full(x.shape, EXP2_COEF[7], dtype=np.float32)
    for c in EXP2_COEF[6::-1]:
        p = (p * f).astype(np.float32)
        p = (p + c).astype(np.float32)
    bits = p.view(np.int32) + (nf.astype(np.int32) << 23)
    return np.where(alive, bits.view(np.float32), _F(0.0)).astype(np.float32)


def draw_indicators_f32(ticks, coef_c, coef_a, u, chunk=1 << 16):
    """One inverse-CDF categorical draw per datum, float32 IEEE-only (EXACT-mode mirror).

    logit = c - (a * tick)  [separate multiply and subtract], m = max_k, p_k =
    soft_exp2(logit_k - m), cum_k sequential over k ascending, thr = u * cum_{K-1},
    s = #{k: cum_k <= thr} clamped to K-1.
    """
    ticks = np.asarray(ticks)
    c = np.asarray(coef_c, dtype=np.float32)
    a = np.asarray(coef_a, dtype=np.float32)
    u = np.asarray(u, dtype=np.float32)
    K = len(c)
    out = np.empty(len(ticks), dtype=np.uint8)
    for lo in range(0, len(ticks), chunk):
        tf = ticks[lo:lo + chunk].astype(np.float32)
        prod = (a[None, :] * tf[:, None]).astype(np.float32)
        logit = (c[None, :] - prod).astype(np.float32)
        m = logit.max(axis=1)
        p = soft_exp2((logit - m[:, None]).astype(np.float32))
        cum = np.empty_like(p)
        acc = np.zeros(len(tf), dtype=np.float32)
        for k in range(K):
            acc = (acc + p[:, k]).astype(np.float32)
            cum[:, k] = acc
        thr = (u[lo:lo + chunk] * acc).astype(np.float32)
        s = (cum <= thr[:, None]).sum(axis=1)
        out[lo:lo + chunk] = np.minimum(s, K - 1).astype(np.uint8)
    return out


def posterior_numpy(nk, tk, ts, whypers, rhypers, rng):
    """Dirichlet / Gamma update of basicrta/gibbs.py:210-211 with T_k = tk * ts."""
    weights = rng.dirichlet(whypers + nk)
    rates = rng.gamma(rhypers[:, 0] + nk, 1.0 / (rhypers[:, 1] + tk * float(ts)))
    return weights, rates


def canonical_order(ticks):
    """The sampler walks a chain's data in ascending tick order (stable): position p of that
    order consumes Philox word p.  Data are exchangeable in the model (basicrta/gibbs.py:196-207
    never looks at their order), so this only fixes which uniform a datum gets."""
    return np.argsort(np.asarray(ticks), kind='stable')


def run_teacher_forced(ticks, ts, ncomp, niter, seed, chain_id, rng, g=100,
                       whypers=None, rhypers=None, uniforms=None):
    """Whole chain in device arithmetic with a host-side posterior draw.  Data are processed in
    :func:`canonical_order`; ``indicator`` (and ``uniforms``) are in the caller's original order."""
    order = canonical_order(ticks)
    out = _run_teacher_forced_canonical(np.asarray(ticks)[order], ts, ncomp, niter, seed, chain_id, rng, g=g,
                                        whypers=whypers, rhypers=rhypers,
                                        uniforms=None if uniforms is None else np.asarray(uniforms)[:, order])
    ind = np.empty_like(out['indicator'])
    ind[:, order] = out['indicator']
    out['indicator'] = ind
    return out


def _run_teacher_forced_canonical(ticks, ts, ncomp, niter, seed, chain_id, rng, g=100,
                                  whypers=None, rhypers=None, uniforms=None):
    """Whole chain in device arithmetic with a host-side posterior draw.

    Returns the outputs of :func:`allocate` plus the per-iteration traces the CUDA
    kernel is compared against: ``coef_c``/``coef_a`` [niter, K] float32 (row j-1 holds
    the coefficients *used* in iteration j), ``nk``/``tk`` [niter, K] int64.

    ``uniforms`` (optional, [niter, N] float32) replaces the Philox stream.
    """
    ticks = np.asarray(ticks, dtype=np.int64)
    n = len(ticks)
    out = allocate(niter, g, n, ncomp)
    wh = out['whypers'] if whypers is None else np.asarray(whypers, float)
    rh = out['rhypers'] if rhypers is None else np.asarray(rhypers, float)
    weights, rates = init_state(ncomp)
    out.update(coef_c=np.zeros((niter, ncomp), np.float32),
               coef_a=np.zeros((niter, ncomp), np.float32),
               nk=np.zeros((niter, ncomp), np.int64),
               tk=np.zeros((niter, ncomp), np.int64))
    for j in range(1, niter + 1):
        c, a = coefficients(weights, rates, ts)
        out['coef_c'][j - 1], out['coef_a'][j - 1] = c, a
        if uniforms is None:
            u = philox.indicator_uniforms(seed, chain_id, j, n)
        else:
            u = uniforms[j - 1]
        s = draw_indicators_f32(ticks, c, a, u)
        nk, tk = sufficient_stats(s, ticks, ncomp)
        out['nk'][j - 1], out['tk'][j - 1] = nk, tk
        weights, rates = posterior_numpy(nk, tk, ts, wh, rh, rng)
        if j % g == 0:
            row = j // g - 1
            out['mcweights'][row], out['mcrates'][row] = weights, rates
            out['indicator'][row] = s
    return out


# --------------------------------------------------------------------------------------
# synthetic data: seeded, quantised up to the ts grid
# --------------------------------------------------------------------------------------
def synth_times(n, weights, rates, seed, ts=0.1):
    """Hyper-exponential sample, ceil-quantised to ``ts`` (min = ts).

    Re-specification of ``basicrta/util.py:596-608`` (``simulate_hn``) with a seeded
    generator and the grid real contact durations live on (contacts.py:222-229).
    """
    rng = np.random.default_rng(seed)
    comp = rng.choice(len(weights), size=int(n), p=np.asarray(weights) / np.sum(weights))
    x = rng.exponential(1.0 / np.asarray(rates, dtype=float)[comp])
    ticks = np.maximum(np.ceil(x / ts), 1.0)
    return ticks * ts


# ---- label-invariant posterior functionals (statistical parity tests) -------------------------
FUNCTIONAL_QUANTILES = (0.5, 0.75, 0.9, 0.97, 0.99, 0.997, 0.999)
FUNCTIONAL_NAMES = tuple(f'S(t_q{q:g})' for q in FUNCTIONAL_QUANTILES) + (
    'mean rate', 'mean time', 'slowest significant rate', 'heaviest weight', 'components above cutoff')


def well_determined(n, min_tail=30):
    """Mask over FUNCTIONAL_NAMES: a survival functional S(t_q) counts only if at least ``min_tail`` data
    lie beyond t_q -- further out the posterior of S(t) is as wide as S(t) itself, and a comparison of two
    Monte-Carlo means to 2 % says nothing about the sampler."""
    keep = [(1.0 - q) * n >= min_tail for q in FUNCTIONAL_QUANTILES]
    return np.array(keep + [True] * (len(FUNCTIONAL_NAMES) - len(FUNCTIONAL_QUANTILES)))


def functional_times(times):
    """Time grid of the survival functionals: upper quantiles of the data, so that the mixture survival
    S(t) = sum_k w_k exp(-r_k t) is evaluated where the data constrain it (S from 0.5 down to 1e-3)."""
    return np.quantile(np.asarray(times, dtype=np.float64), FUNCTIONAL_QUANTILES)


def posterior_functionals(weights, rates, times):
    """Per stored sample, functionals of (weights, rates) that do not depend on the labelling of the
    components (the reference's chains switch labels, and dead components draw from the prior,
    basicrta/gibbs.py:210-211): the mixture survival on ``functional_times``, the mean rate
    sum w r, the mean residence time sum w / r, the slowest rate among components above the
    10/N weight cut-off of ``process_gibbs`` (gibbs.py:284-296) -- what ``estimate_tau`` is built on --,
    the heaviest weight and the number of components above the cut-off.
    Returns float64 [samples, len(FUNCTIONAL_NAMES)]."""
    w = np.asarray(weights, dtype=np.float64)
    r = np.asarray(rates, dtype=np.float64)
    n = len(times)
    cols = [(w * np.exp(-r * t)).sum(axis=1) for t in functional_times(times)]
    cols.append((w * r).sum(axis=1))
    cols.append((w / r).sum(axis=1))
    sig = w > 10.0 / n
    cols.append(np.where(sig, r, np.inf).min(axis=1))
    cols.append(w.max(axis=1))
    cols.append(sig.sum(axis=1).astype(np.float64))
    return np.stack(cols, axis=1)
