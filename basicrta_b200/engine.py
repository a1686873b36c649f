"""Host side of the batch sampler: packs chains into the device layout, builds the
schedule, calls the C ABI (``brta_gibbs_run_batch``) and brings the results back.

PyTorch is used for device memory, pinned host memory and streams only; the compute is
the hand-written kernel in ``csrc/brta_gibbs.cu``.  There is no CPU path: every entry
point raises if CUDA or the built library is unavailable.

Layout in HBM (one batch = the residues one GPU runs in one launch):

=================  ==========================  ==========================================
buffer             shape / dtype               notes
=================  ==========================  ==========================================
ticks              uint16 or uint32 [sum N8]   times as multiples of ts, chain-concatenated,
                                               each chain padded to 8 elements (16 B)
per-chain scalars  n_data, tick_offset, ...    [R]
whyper, rhyper     float32 [R,K], [R,K,2]      basicrta/gibbs.py:173-174
init_c, init_a     float32 [R,K]               log2(w r), r ts log2(e) of gibbs.py:186-188
mcweights/mcrates  float64 [R,S,K]             basicrta/gibbs.py:169-170
indicator          uint8, chain r = [S,N_r]    dense, the reference's layout (gibbs.py:167)
exchange           BRTA_EXCH_BYTES(team)/chain tagged mailboxes of integer team partials
=================  ==========================  ==========================================
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _cabi
from .memo import COST_DIRECT, COST_SERVED, ChainCost
from .plan import build_plan, rebalance_team

LOG2E = 1.4426950408889634


@dataclass
class ChainInput:
    """One residue's chain: the reference's ``Gibbs`` inputs in device units."""
    ticks: np.ndarray                 # integer multiples of ts (any integer dtype)
    ts: float
    chain_id: int = 0
    whypers: np.ndarray = None        # [K]   (default 1/K, gibbs.py:173)
    rhypers: np.ndarray = None        # [K,2] (default (1,3), gibbs.py:174)
    init_weights: np.ndarray = None   # [K]   (default gibbs.py:186-188)
    init_rates: np.ndarray = None     # [K]


@dataclass
class ChainResult:
    mcweights: np.ndarray = None      # [S,K] float64
    mcrates: np.ndarray = None        # [S,K] float64
    indicator: np.ndarray = None      # [S,N] uint8
    status: int = 0
    trace_nk: np.ndarray = None       # [niter,K] int64 (FLAG_TRACE)
    trace_tk: np.ndarray = None


def initial_state(ncomp):
    """Deterministic start of every chain (basicrta/gibbs.py:186-188): rates 5, 0.5, ...
    descending by decades, weights proportional to 0.9, 0.09, ..."""
    k = np.arange(ncomp, dtype=np.float64)
    rates = 0.5 * 10.0 ** (1 - k)
    weights = 9.0 * 10.0 ** (-(k + 1))
    return weights / weights.sum(), rates


def coefficients(weights, rates, ts):
    """Logit coefficients in log2 units: logit_k(tick) = c_k - a_k * tick."""
    w = np.asarray(weights, dtype=np.float64)
    r = np.asarray(rates, dtype=np.float64)
    with np.errstate(divide='ignore'):
        c = np.log2(w) + np.log2(r)
    a = r * (float(ts) * LOG2E)
    return c.astype(np.float32), a.astype(np.float32)


def times_to_ticks(times, ts):
    """Residence times -> integer tick counts on the grid ``ts`` (strict: every time must lie within
    1e-3 of a tick).  :func:`tick_grid` finds the grid of arbitrary data."""
    t = np.asarray(times, dtype=np.float64) / float(ts)
    ticks = np.rint(t)
    if len(ticks) and np.max(np.abs(t - ticks)) > 1e-3:
        raise ValueError('times are not integer multiples of ts; quantise them to the trajectory '
                         'time step first (basicrta/contacts.py:222-229)')
    if len(ticks) and (ticks.min() < 0 or ticks.max() >= _cabi.TICK_LIMIT):
        raise ValueError(f'tick range exceeds [0, 2^23): max {ticks.max():.0f}')
    return ticks.astype(np.int64)


GRID_TOL = 1e-3                 # a time counts as "on the grid" within this many ticks
GRID_MAX_DIVISOR = 4096         # finest sub-multiple of the hint tried before the data count as continuous
SLICE_SUM_DATA = 1 << 17        # the tick sum of this many of the largest data must stay below 2^32


def _on_grid(t, grid, tol):
    r = t / grid
    return float(np.max(np.abs(r - np.rint(r)))) <= tol


def tick_grid(times, ts_hint=None):
    """The integer grid the device works on: ``(ticks int64, grid)`` with ``times ~= ticks * grid``.

    The reference's ``ts`` (first non-zero gap of the sorted times, gibbs.py:147-151) is only a hint:
    contact durations are multiples of the trajectory step dt (contacts.py:222-229), but the first
    gap of a sparse residue is usually a MULTIPLE of dt ({0.3, 0.7, 0.8} gives ts = 0.4), so the grid is
    the coarsest ``ts_hint / m`` (m = 1, 2, ...) every time is a multiple of.  Data on no such grid
    (continuous times, e.g. the reference's ``util.simulate_hn``, util.py:596-608) are put on a
    fixed-point grid of up to 2^23 - 1 levels of the largest time -- the resolution float32 offers --
    coarsened if needed so that 32-bit per-slice tick sums cannot overflow; the returned ``ticks`` are
    then a rounding of the data, exact on the grid returned.  ``self.ts`` of the pickle is untouched."""
    t = np.asarray(times, dtype=np.float64)
    if t.ndim != 1 or len(t) == 0:
        raise ValueError('times must be a non-empty 1-D array')
    if not np.all(np.isfinite(t)) or t.min() < 0:
        raise ValueError('times must be finite and non-negative')
    tmax = float(t.max())
    if tmax <= 0:
        raise ValueError('all times are zero')
    tol = GRID_TOL if np.asarray(times).dtype != np.float32 else 2e-2
    if ts_hint is None or not np.isfinite(ts_hint) or ts_hint <= 0:
        pos = t[t > 0]
        ts_hint = float(pos.min())
    ts_hint = float(ts_hint)
    probe = t[:: max(1, len(t) // 256)]                              # cheap rejection before the full check
    for m in range(1, GRID_MAX_DIVISOR + 1):
        grid = ts_hint / m
        if tmax / grid >= _cabi.TICK_LIMIT:
            break
        if _on_grid(probe, grid, tol) and _on_grid(t, grid, tol):
            return np.rint(t / grid).astype(np.int64), grid
    # continuous data: fixed point
    srt = np.sort(t)[-SLICE_SUM_DATA:]
    levels = float(_cabi.TICK_LIMIT - 1)
    while levels > 2 and float(np.sum(np.rint(srt / (tmax / levels)))) >= float(1 << 32):
        levels = np.floor(levels / 2)
    grid = tmax / levels
    return np.rint(t / grid).astype(np.int64), grid


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _cabi.BrtaError('basicrta_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch


@dataclass
class DeviceBatch:
    """Everything one launch needs, resident on the device."""
    batch: _cabi.Batch
    plan: object
    tensors: dict = field(default_factory=dict)      # keeps device memory alive
    host: dict = field(default_factory=dict)         # pinned host copies of the inputs
    order: list = None                                # per chain: canonical position -> original index
    n_data: np.ndarray = None
    ind_offset: np.ndarray = None
    rows: int = 0
    ncomp: int = 0
    niter: int = 0
    flags: int = 0
    h2d_bytes: int = 0
    units: float = 0.0                                # sum_r N_r * K * niter
    executed_ex2_share: float = 1.0                   # ex2 actually executed / units (memoised rows make it < 1)
    segments: list = None                              # iteration ends of the launches of one run (None: one launch)
    cal: tuple = None                                  # (costs, ticks, tick_offset) for re-slicing between segments
    kernel_choice: dict = None                         # measured choice between the 4- and 3-CTAs-per-SM builds
    progress: object = None                            # pinned int32 [R]: rows complete per chain, written by the kernel


CALIBRATE_MIN_NITER = 2000     # shorter runs are not worth the calibration launches
CALIBRATE_ITERS = 96           # iterations per calibration launch (the kernel times the second half)
CALIBRATE_ROUNDS = 2
CHOICE_ITERS = 384             # iterations of the launches that decide between the two kernel builds
SEGMENT_MIN_NITER = 20000      # longer calibrated runs are cut into launches and re-sliced in between
SEGMENT_FRACTIONS = (0.02, 0.06, 0.15, 0.3, 0.5, 0.75)
SEGMENT_DAMPING = 0.5          # move the boundaries half of the way to the measured optimum (noise, one-segment lag)
SEGMENT_MIN_SPREAD = 0.15      # teams whose members are within 15 % of each other are left alone


def _canonical_order(ticks):
    """Stable ascending-tick order (oracle.gibbs_oracle.canonical_order).  16-bit keys take NumPy's radix
    sort, 4x faster than the comparison sort of int64 -- the sort is the largest host cost of a batch."""
    t = np.asarray(ticks)
    if t.size and 0 <= int(t.min()) and int(t.max()) < 65536:
        t = t.astype(np.uint16)
    return np.argsort(t, kind='stable').astype(np.int32)


class GibbsEngine:
    """One engine per GPU.  ``prepare`` (pack + H2D), ``launch`` (async kernel),
    ``fetch`` (D2H) are separate so callers can time and overlap them."""

    def __init__(self, device=0, ctas_per_sm=None, overhead_quads=None):
        self.torch = _torch()
        self.lib = _cabi.load()
        self.device = int(device)
        self.caps = _cabi.query(self.device)
        if (self.caps.cc_major, self.caps.cc_minor) != (10, 0):
            raise _cabi.BrtaError(f'built for sm_100a (B200); device {device} is '
                                  f'sm_{self.caps.cc_major}{self.caps.cc_minor}')
        self.ctas_per_sm = ctas_per_sm
        self.overhead_quads = overhead_quads

    # ---- schedule ----------------------------------------------------------------------
    def _plan(self, n_data, ncomp, flags, n_waves=None, costs=None, narrow=None):
        """``costs``: per chain a :class:`memo.ChainCost` (ticks in canonical order) -- slices are then
        cut to equal time under the memoised-row cost model instead of equal length.  ``narrow[r]``:
        chain r's ticks fit 16 bits, so its slices take 8 B per quad of shared memory instead of 16."""
        n_quads = (np.asarray(n_data, dtype=np.int64) + 3) // 4
        # occupancy is set by registers for small slices; ask with a small slice first,
        # then shrink the capacity to what that occupancy leaves per CTA.
        info = _cabi.launch_info(self.device, ncomp, flags, 64)
        per_sm = info.ctas_per_sm if self.ctas_per_sm is None else min(self.ctas_per_sm, info.ctas_per_sm)
        if per_sm < 1:
            raise _cabi.BrtaError('sampler kernel does not fit on an SM')
        smem_sm = 227 * 1024
        cap16 = (smem_sm // per_sm - info.static_smem - 1024) // 16          # 16-byte units per CTA
        narrow = np.zeros(len(n_quads), bool) if narrow is None else np.asarray(narrow, bool)
        cap = np.where(narrow, 2 * cap16, cap16).astype(np.int64)             # quads per CTA, per chain
        grid = self.caps.sm_count * per_sm
        kw = {} if self.overhead_quads is None else {'overhead_quads': self.overhead_quads}
        if costs is None:
            plan = build_plan(n_quads, grid, cap, n_waves=n_waves, **kw)
        else:
            # pass 1 (equal quads) fixes provisional team sizes; pass 2 balances time with the cost model
            first = build_plan(n_quads, grid, cap, n_waves=n_waves, **kw)
            cost = [costs[r].total_cost(int(first.team_size[r]), int(cap[r])) for r in range(len(n_quads))]
            plan = build_plan(n_quads, grid, cap, n_waves=n_waves, cost=cost,
                              splitter=lambda r, c: costs[r].split(c, int(cap[r])), **kw)
        # dynamic shared memory in 16-byte units: the largest task slice
        units = np.where(narrow[plan.tasks['chain']], (plan.tasks['quad_count'] + 1) // 2, plan.tasks['quad_count'])
        plan.smem_units = int(units.max())
        plan.cap_quads, plan.cap_units, plan.tick_total = cap, int(cap16), {}
        check = _cabi.launch_info(self.device, ncomp, flags, plan.smem_units)
        if check.ctas_per_sm < per_sm:
            raise _cabi.BrtaError('schedule assumes more co-resident CTAs than the device grants')
        return plan

    @staticmethod
    def _wide_plan(n_data):
        """ncomp > 32: the general kernel runs one CTA per chain and needs no schedule; this is the trivial plan
        the rest of the host code reads (one task per chain, teams of one, one wave)."""
        from .plan import Plan, TASK_DTYPE
        n_quads = (np.asarray(n_data, dtype=np.int64) + 3) // 4
        R = len(n_quads)
        tasks = np.zeros(R, dtype=TASK_DTYPE)
        tasks['chain'], tasks['team_size'], tasks['quad_count'], tasks['order'] = np.arange(R), 1, n_quads, np.arange(R)
        plan = Plan(tasks=tasks, cta_task_begin=np.arange(R + 1, dtype=np.int32), grid=R,
                    slice_cap_quads=int(n_quads.max()), n_waves=1, team_size=np.ones(R, dtype=np.int32),
                    wave_of_chain=np.zeros(R, dtype=np.int32), est_efficiency=1.0, smem_units=1)
        plan.cap_quads, plan.cap_units, plan.tick_total = n_quads.copy(), int(n_quads.max()), {}
        return plan

    @staticmethod
    def _check_slice_sums(plan, ticks, tick_offset, d0, n_local):
        """Per-iteration slice sums are accumulated in 32 bits inside a CTA."""
        for task in plan.tasks:
            r = task['chain']
            lo = tick_offset[r] + 4 * int(task['quad_begin'])
            hi = min(lo + 4 * int(task['quad_count']), tick_offset[r] + d0 + n_local[r])
            if int(ticks[lo:hi].sum(dtype=np.int64)) >= (1 << 32):
                raise ValueError('a slice holds more than 2^32 ticks; use a coarser ts')

    @staticmethod
    def _watchdog_ns(plan, n_data, niter):
        """Rendezvous watchdog of the launch: a CTA may legitimately wait for a team mate that is still
        finishing its previous wave, i.e. up to about the whole run.  Ten times a generous estimate of the
        run (20 cycles per quad and iteration at 1.5 GHz along the schedule's critical path), at least 60 s."""
        quads = float(np.sum((np.asarray(n_data, dtype=np.int64) + 3) // 4))
        makespan = quads / max(plan.grid * max(plan.est_efficiency, 1e-3), 1.0) + 400.0 * plan.n_waves
        est_s = makespan * float(niter) * 20.0 / 1.5e9
        return int(max(60.0, 10.0 * est_s) * 1e9)

    # ---- pack + upload -------------------------------------------------------------------
    def prepare(self, chains, ncomp, niter, thin=100, seed=0, flags=0, inject=None, n_waves=None, shard=None,
                calibrate=None, segments=None, choose_kernel=None, progress_rows=0):
        """``chains``: list of :class:`ChainInput`.  ``inject``: dict with optional
        ``coef_c``/``coef_a`` (list of [niter,K] float32) and ``u`` (list of [niter,N]).
        ``shard`` (internal, see :func:`run_sharded`): this GPU's part of ONE chain split over
        several GPUs: dict(rank, n_shards, quad_begin, quad_count, mailbox_table)."""
        torch = self.torch
        dev = torch.device('cuda', self.device)
        R = len(chains)
        K = int(ncomp)
        if not 1 <= K <= _cabi.MAX_NCOMP:
            raise ValueError(f'ncomp must be in 1..{_cabi.MAX_NCOMP}')
        if niter < 1 or thin < 1:
            raise ValueError('niter and thin must be >= 1')
        rows = (niter + 1) // thin
        flags = int(flags)
        costs = None
        n_data = np.array([len(c.ticks) for c in chains], dtype=np.int64)
        if R == 0 or n_data.min() < 1:
            raise ValueError('every chain needs at least one datum')
        if n_data.max() >= (1 << 27):
            raise ValueError('chain too long (N must be < 2^27)')
        max_tick = max(int(np.max(c.ticks)) for c in chains)
        min_tick = min(int(np.min(c.ticks)) for c in chains)
        if min_tick < 0 or max_tick >= _cabi.TICK_LIMIT:
            raise ValueError('ticks must lie in [0, 2^23)')
        tick_dtype = np.uint16 if max_tick < 65536 else np.uint32

        # canonical order: ascending ticks (stable).  Philox word p belongs to canonical position p.
        order = [_canonical_order(ch.ticks) for ch in chains]
        wide = K > _cabi.LANE_MAX_NCOMP                    # the general kernel: one CTA per chain, no schedule
        if wide and (shard is not None or flags & _cabi.FLAG_CTAS3):
            raise ValueError(f'ncomp > {_cabi.LANE_MAX_NCOMP} supports neither sharded chains nor FLAG_CTAS3')
        if shard is not None:
            if R != 1:
                raise ValueError('a sharded launch holds exactly one chain')
            sq0, snq = int(shard['quad_begin']), int(shard['quad_count'])
            d0 = 4 * sq0                                              # first datum of this GPU's shard
            n_local = np.array([min(int(n_data[0]), 4 * (sq0 + snq)) - d0], dtype=np.int64)
            plan = self._plan(np.array([4 * snq]), K, flags, n_waves=n_waves, narrow=[max_tick < 65536])
            plan.tasks['quad_begin'] += sq0                          # tasks carry GLOBAL quad indices
        else:
            d0 = 0
            n_local = n_data
            costs = None
            if wide:
                plan = self._wide_plan(n_data)
            else:
                if not flags & _cabi.FLAG_NO_TABLE:
                    costs = [ChainCost(np.asarray(ch.ticks)[o], K) for ch, o in zip(chains, order)]
                plan = self._plan(n_data, K, flags, n_waves=n_waves, costs=costs,
                                  narrow=[int(np.max(ch.ticks)) < 65536 for ch in chains])
        if int(plan.team_size.max()) * int(niter) >= (1 << 31):
            raise ValueError('team_size * niter overflows the arrive counter')

        pad8 = (n_local + 7) // 8 * 8
        tick_offset = np.concatenate(([0], np.cumsum(pad8)[:-1])).astype(np.int64) - d0
        ticks = np.zeros(int(pad8.sum()), dtype=tick_dtype)
        ts = np.zeros(R, dtype=np.float32)
        chain_id = np.zeros(R, dtype=np.uint32)
        whyper = np.zeros((R, K), dtype=np.float32)
        rhyper = np.zeros((R, K, 2), dtype=np.float32)
        init_c = np.zeros((R, K), dtype=np.float32)
        init_a = np.zeros((R, K), dtype=np.float32)
        w0, r0 = initial_state(K)
        for r, ch in enumerate(chains):
            t = np.asarray(ch.ticks)[order[r]][d0:d0 + int(n_local[r])]
            ticks[tick_offset[r] + d0:tick_offset[r] + d0 + len(t)] = t
            ts[r] = ch.ts
            chain_id[r] = np.uint32(ch.chain_id & 0xFFFFFFFF)
            whyper[r] = np.ones(K) / K if ch.whypers is None else ch.whypers
            rhyper[r] = np.ones((K, 2)) * [1.0, 3.0] if ch.rhypers is None else ch.rhypers
            iw = w0 if ch.init_weights is None else ch.init_weights
            ir = r0 if ch.init_rates is None else ch.init_rates
            init_c[r], init_a[r] = coefficients(iw, ir, ch.ts)
        ind_stride = n_local.astype(np.int32)
        ind_bytes = rows * n_local
        ind_offset = np.concatenate(([0], np.cumsum(ind_bytes)[:-1])).astype(np.int64) - d0

        T = {}
        h2d = 0

        H = {}

        def up(name, arr):
            nonlocal h2d
            src = torch.from_numpy(np.ascontiguousarray(arr))
            pinned = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            pinned.copy_(src)
            H[name] = pinned
            t = pinned.to(dev, non_blocking=True)
            T[name] = t
            h2d += t.numel() * t.element_size()
            return t.data_ptr()

        b = _cabi.Batch()
        b.n_chains, b.ncomp, b.niter, b.thin = R, K, int(niter), int(thin)
        b.tick_bytes = ticks.dtype.itemsize
        b.flags = int(flags)
        b.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        b.device = self.device
        b.watchdog_ns = self._watchdog_ns(plan, n_local, niter)
        b.ticks = up('ticks', ticks.view(np.int16 if tick_dtype == np.uint16 else np.int32))
        b.tick_offset = up('tick_offset', tick_offset)
        if shard is None:                                  # sharded runs keep canonical order, see run_sharded
            perm_offset = np.concatenate(([0], np.cumsum(n_data)[:-1])).astype(np.int64)
            b.perm = up('perm', np.concatenate(order))
            b.perm_offset = up('perm_offset', perm_offset)
        b.n_data = up('n_data', n_data.astype(np.int32))
        b.max_tick = up('max_tick', np.array([int(np.max(ch.ticks)) for ch in chains], dtype=np.int64).astype(np.int32))
        b.chain_id = up('chain_id', chain_id.view(np.int32))
        b.ts = up('ts', ts)
        b.whyper = up('whyper', whyper)
        b.rhyper = up('rhyper', rhyper)
        b.init_c = up('init_c', init_c)
        b.init_a = up('init_a', init_a)
        b.ind_offset = up('ind_offset', ind_offset)
        b.ind_stride = up('ind_stride', ind_stride)
        b.tasks = up('tasks', plan.tasks.view(np.int32).reshape(-1, 6))
        b.cta_task_begin = up('cta_task_begin', plan.cta_task_begin)
        b.grid_ctas = plan.grid
        b.slice_cap_quads = plan.smem_units
        if not wide:                                       # the general kernel sums ticks in 64 bits
            self._check_slice_sums(plan, ticks, tick_offset, d0, n_local)

        def dev_zeros(name, shape, dtype):
            T[name] = torch.zeros(shape, dtype=dtype, device=dev)
            return T[name].data_ptr()

        b.mcweights = dev_zeros('mcweights', (R, max(rows, 1), K), torch.float64)
        b.mcrates = dev_zeros('mcrates', (R, max(rows, 1), K), torch.float64)
        b.indicator = dev_zeros('indicator', (max(int(ind_bytes.sum()), 1),), torch.uint8)
        b.status = dev_zeros('status', (R,), torch.int32)
        exch_size = np.array([(max(_cabi.exch_bytes(int(c)), 1280) + 127) // 128 * 128 for c in plan.team_size],
                             dtype=np.int64)
        exch_offset = np.concatenate(([0], np.cumsum(exch_size)[:-1])).astype(np.int64)
        b.exchange = dev_zeros('exchange', (int(exch_size.sum()),), torch.uint8)
        b.exch_offset = up('exch_offset', exch_offset)

        if shard is not None:
            b.n_shards, b.shard_rank = int(shard['n_shards']), int(shard['rank'])
            b.shard_mailbox = shard['mailbox_table'].data_ptr()
            T['shard_mailbox_table'] = shard['mailbox_table']

        inject = inject or {}
        if flags & _cabi.FLAG_INJECT_COEF:
            b.inj_c = up('inj_c', np.stack([np.asarray(x, np.float32).reshape(niter, K) for x in inject['coef_c']]))
            b.inj_a = up('inj_a', np.stack([np.asarray(x, np.float32).reshape(niter, K) for x in inject['coef_a']]))
        if flags & _cabi.FLAG_INJECT_U:
            pitch = (n_data + 3) // 4 * 4
            u_off = np.concatenate(([0], np.cumsum(pitch * niter)[:-1])).astype(np.int64)
            u_all = np.zeros(int((pitch * niter).sum()), dtype=np.float32)
            for r, u in enumerate(inject['u']):                   # given per original datum -> canonical order
                blk = u_all[u_off[r]:u_off[r] + pitch[r] * niter].reshape(niter, pitch[r])
                blk[:, :n_data[r]] = np.asarray(u, np.float32)[:, order[r]]
            b.inj_u = up('inj_u', u_all)
            b.inj_u_offset = up('inj_u_offset', u_off)
        if flags & _cabi.FLAG_TRACE:
            b.trace_nk = dev_zeros('trace_nk', (R, niter, K), torch.int64)
            b.trace_tk = dev_zeros('trace_tk', (R, niter, K), torch.int64)

        progress = None
        if progress_rows and shard is None and K <= _cabi.LANE_MAX_NCOMP and \
                not flags & (_cabi.FLAG_INJECT_COEF | _cabi.FLAG_INJECT_U | _cabi.FLAG_TRACE):
            # mapped pinned host memory: with unified addressing the device writes through the same pointer
            progress = torch.zeros(R, dtype=torch.int32).pin_memory()
            b.progress, b.progress_rows = progress.data_ptr(), int(progress_rows)
        torch.cuda.current_stream(dev).synchronize()
        db = DeviceBatch(batch=b, plan=plan, tensors=T, host=H, order=order, n_data=n_data, ind_offset=ind_offset,
                         rows=rows, ncomp=K, niter=int(niter), flags=int(flags), h2d_bytes=h2d,
                         units=float(n_data.sum()) * K * int(niter))
        db.progress = progress
        if wide:
            if calibrate:
                raise ValueError(f'ncomp > {_cabi.LANE_MAX_NCOMP} has no schedule to calibrate')
            calibrate, db.executed_ex2_share = False, 3.0   # three passes over the components per datum
        if calibrate is None:
            calibrate = (shard is None and niter >= CALIBRATE_MIN_NITER and int(plan.team_size.max()) > 1 and
                         not flags & (_cabi.FLAG_INJECT_COEF | _cabi.FLAG_INJECT_U | _cabi.FLAG_TRACE))
        if calibrate:
            if shard is not None or flags & (_cabi.FLAG_INJECT_COEF | _cabi.FLAG_INJECT_U | _cabi.FLAG_TRACE):
                raise ValueError('calibration needs a plain (not sharded, not injected, not traced) batch')
            if choose_kernel is None:
                choose_kernel = (K <= 16 and n_waves is None and not wide and
                                 not flags & (_cabi.FLAG_EXACT | _cabi.FLAG_CTAS3))
            if choose_kernel:                              # calibrates both candidates itself
                self._choose_kernel(db, n_data, costs, [int(np.max(ch.ticks)) < 65536 for ch in chains], ticks, tick_offset)
                plan = db.plan
            # every launch ends with a tail in which CTAs run dry; with several waves that tail is long (the
            # waves are only balanced over the whole run), so only single-wave schedules are segmented by default
            if segments is None and niter >= SEGMENT_MIN_NITER and plan.n_waves == 1:
                segments = SEGMENT_FRACTIONS
            if segments:
                ends = sorted({int(round(f * niter / thin)) * int(thin) for f in segments} | {int(niter)})
                ends = [e for e in ends if 0 < e <= niter]
                if len(ends) > 2:
                    for name in ('c', 'a'):
                        T['final_' + name] = torch.zeros((R, K), dtype=torch.float32, device=dev)
                        T['init_' + name + '0'] = T['init_' + name].clone()
                    b.final_c, b.final_a = T['final_c'].data_ptr(), T['final_a'].data_ptr()
                    db.segments, db.cal = ends, (costs, ticks, tick_offset)
            if not choose_kernel:
                for _ in range(CALIBRATE_ROUNDS):
                    self._calibrate(db, costs, ticks, tick_offset)
        if shard is None and costs is not None:
            db.executed_ex2_share = self._executed_ex2_share(plan, costs, float(n_data.sum()))
        if shard is not None:
            db.n_data, db.ind_offset = n_local, ind_offset + d0      # host-side views are shard-local
        return db

    @staticmethod
    def _executed_ex2_share(plan, costs, n_total):
        """ex2 actually issued per algorithmic unit: recomputed data + memoised rows built."""
        done = 0.0
        for task in plan.tasks:
            cc = costs[int(task['chain'])]
            s0, e0 = int(task['quad_begin']), int(task['quad_begin'] + task['quad_count'])
            k = cc.served(s0, e0)
            rows_built = 0
            if k > 0:
                hi_tick = int(cc.qmax[min(e0, cc.n_full) - 1]) if min(e0, cc.n_full) > s0 else int(cc.qmin[s0])
                rows_built = min(cc.rows, hi_tick - int(cc.qmin[s0]) + 1)
            done += 4.0 * (e0 - s0 - k) + rows_built
        return done / n_total

    def _time_short(self, db, iters=CALIBRATE_ITERS, reps=2):
        """Device time (ms) of a short launch of ``db`` (no row saved, workspace reset afterwards)."""
        torch = self.torch
        cal = _cabi.Batch.from_buffer_copy(db.batch)
        cal.niter, cal.thin = int(iters), int(iters) + 2
        cal.progress, cal.progress_rows = None, 0
        best = float('inf')
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream()
            for rep in range(reps + 1):                             # the first launch warms up (module load, L2)
                self.reset(db)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = self.lib.brta_gibbs_run_batch(C.byref(cal), C.c_void_p(s.cuda_stream))
                e1.record()
                _cabi.check(rc, 'brta_gibbs_run_batch (kernel choice)')
                s.synchronize()
                if rep:
                    best = min(best, e0.elapsed_time(e1))
        if int(db.tensors['status'].max().item()) != 0:
            raise _cabi.BrtaError('short launch failed (status != 0)')
        self.reset(db)
        return best

    def _choose_kernel(self, db, n_data, costs, narrow, ticks, tick_offset):
        """K <= 16 comes in two builds: 4 CTAs per SM at 128 registers, and 3 CTAs per SM at 168 registers
        with a third more shared memory per slice (BRTA_FLAG_CTAS3).  Fewer, larger slices mean less
        per-iteration fixed work per chain; how that trades against occupancy, and how many waves the 3-CTA
        build wants, is not something the cost model predicts (C2: 4 CTAs x 2 waves 4.21e12 units/s, 3 CTAs x
        2 waves 4.06e12, 3 CTAs x 3 waves 4.60e12; 200 residues: 3 CTAs x 2 waves wins; 50-100 residues: 3 CTAs x
        1 wave).  Rather than guess, time a short launch of each candidate -- the planner's schedule for the
        4-CTA build, and the 3-CTA build with the same number of waves and, for multi-wave batches, one more --
        after the measured slicing of each, and keep the fastest.  Results do not depend on the choice."""
        torch = self.torch
        T, H, b = db.tensors, db.host, db.batch
        dev = T['tasks'].device
        keys_t = ('tasks', 'cta_task_begin', 'exchange', 'exch_offset')
        keys_h = ('tasks', 'cta_task_begin', 'exch_offset')

        def snapshot():
            return dict(plan=db.plan, flags=int(b.flags), tensors={k: T[k] for k in keys_t},
                        host={k: H[k] for k in keys_h}, watchdog=b.watchdog_ns)

        def restore(snap):
            T.update(snap['tensors'])
            H.update(snap['host'])
            b.flags, b.watchdog_ns = snap['flags'], snap['watchdog']
            b.tasks, b.cta_task_begin = T['tasks'].data_ptr(), T['cta_task_begin'].data_ptr()
            b.exchange, b.exch_offset = T['exchange'].data_ptr(), T['exch_offset'].data_ptr()
            db.plan, db.flags = snap['plan'], int(snap['flags'])
            b.grid_ctas, b.slice_cap_quads = db.plan.grid, db.plan.smem_units

        def up(name, arr):
            pinned = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
            H[name] = pinned
            T[name] = pinned.to(dev, non_blocking=True)
            return T[name].data_ptr()

        def install(plan, flags):
            exch_size = np.array([(max(_cabi.exch_bytes(int(c)), 1280) + 127) // 128 * 128 for c in plan.team_size],
                                 dtype=np.int64)
            b.flags = flags
            b.tasks = up('tasks', plan.tasks.view(np.int32).reshape(-1, 6))
            b.cta_task_begin = up('cta_task_begin', plan.cta_task_begin)
            b.exch_offset = up('exch_offset', np.concatenate(([0], np.cumsum(exch_size)[:-1])).astype(np.int64))
            T['exchange'] = torch.zeros(int(exch_size.sum()), dtype=torch.uint8, device=dev)
            b.exchange = T['exchange'].data_ptr()
            b.grid_ctas, b.slice_cap_quads = plan.grid, plan.smem_units
            b.watchdog_ns = self._watchdog_ns(plan, n_data, db.niter)
            db.plan, db.flags = plan, flags
            torch.cuda.current_stream(dev).synchronize()

        def measure():
            for _ in range(CALIBRATE_ROUNDS):
                self._calibrate(db, costs, ticks, tick_offset)
            return self._time_short(db, CHOICE_ITERS)

        base_waves = int(db.plan.n_waves)
        tried = [dict(ctas_per_sm=int(db.plan.grid // self.caps.sm_count), n_waves=base_waves, ms=measure())]
        snaps = [snapshot()]
        alt_flags = int(b.flags) | _cabi.FLAG_CTAS3
        for waves in ([base_waves] if base_waves == 1 else [base_waves, base_waves + 1]):
            try:
                plan = self._plan(n_data, db.ncomp, alt_flags, n_waves=waves, costs=costs, narrow=narrow)
                if any(t['ctas_per_sm'] == 3 and t['n_waves'] == plan.n_waves for t in tried):
                    continue                                 # the wave count was raised to the feasible minimum
                self._check_slice_sums(plan, ticks, tick_offset, 0, np.asarray(n_data))
            except (ValueError, _cabi.BrtaError):
                continue
            install(plan, alt_flags)
            tried.append(dict(ctas_per_sm=int(plan.grid // self.caps.sm_count), n_waves=int(plan.n_waves), ms=measure()))
            snaps.append(snapshot())
        best = int(np.argmin([t['ms'] for t in tried]))
        if best != len(tried) - 1:
            restore(snaps[best])
        alt = [t['ms'] for t in tried[1:]]
        db.kernel_choice = {'ctas_per_sm': tried[best]['ctas_per_sm'], 'n_waves': tried[best]['n_waves'],
                            'ms_4': tried[0]['ms'], 'ms_3': min(alt) if alt else None, 'candidates': tried}

    def _measure(self, db, iters):
        """Cycles per task from the start of an iteration to the post of its partials (a short launch
        with ``task_cycles`` set; outputs untouched, workspace reset afterwards)."""
        torch = self.torch
        T = db.tensors
        cyc = torch.zeros(len(db.plan.tasks), dtype=torch.int64, device=T['tasks'].device)
        cal = _cabi.Batch.from_buffer_copy(db.batch)
        cal.niter, cal.thin = int(iters), int(iters) + 2            # no row is saved
        cal.progress, cal.progress_rows = None, 0
        cal.task_cycles = cyc.data_ptr()
        self.reset(db)
        with torch.cuda.device(self.device):
            rc = self.lib.brta_gibbs_run_batch(C.byref(cal), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _cabi.check(rc, 'brta_gibbs_run_batch (calibration)')
        times = cyc.cpu().numpy().astype(np.float64)
        if int(T['status'].max().item()) != 0:
            raise _cabi.BrtaError('calibration launch failed (status != 0)')
        self.reset(db)
        return times

    @staticmethod
    def _rebalanced(plan, tasks, times, costs, ticks, tick_offset, damping=1.0, min_spread=0.0):
        """Host side of the measured slicing: a copy of ``tasks`` whose slice boundaries inside every team
        cut the measured time density (``times[i]`` = cycles of task i) into equal parts (or move
        ``damping`` of the way there), and the dynamic shared memory (16-byte units) the new slices need."""
        tasks = tasks.copy()
        narrow = plan.cap_quads > plan.cap_units                    # chains at 8 B per quad
        by_chain = [[] for _ in range(len(plan.team_size))]
        for i, c in enumerate(tasks['chain']):
            by_chain[int(c)].append(i)
        for r, idx in enumerate(by_chain):
            if len(idx) < 2:
                continue
            idx.sort(key=lambda i: int(tasks[i]['team_rank']))
            t_team = times[idx]
            if min_spread > 0.0 and np.ptp(t_team) < min_spread * t_team.mean():
                continue                                           # balanced well enough: leave it alone
            last = tasks[idx[-1]]
            bounds = np.array([int(tasks[i]['quad_begin']) for i in idx] + [int(last['quad_begin'] + last['quad_count'])])
            weight = None
            if costs is not None:
                weight = np.full(int(bounds[-1]), COST_DIRECT)
                for k in range(len(idx)):
                    s0, e0 = int(bounds[k]), int(bounds[k + 1])
                    weight[s0:s0 + costs[r].served(s0, e0)] = COST_SERVED
            new = rebalance_team(bounds, np.maximum(times[idx], 1.0), weight, cap=int(plan.cap_quads[r]))
            if damping < 1.0:                                      # move only part of the way (noisy measurements)
                new = np.rint(bounds + damping * (new - bounds)).astype(np.int64)
                if np.any(np.diff(new) < 1):
                    continue
            # per-iteration slice sums are accumulated in 32 bits inside a CTA: nothing to check if the
            # whole chain sums to less (cached), else every new slice
            total = plan.tick_total.get(r)
            if total is None:
                total = plan.tick_total[r] = int(np.sum(ticks[tick_offset[r]:tick_offset[r] + 4 * int(bounds[-1])], dtype=np.int64))
            if total >= (1 << 32):
                cs = np.concatenate(([0], np.cumsum(ticks[tick_offset[r]:tick_offset[r] + 4 * int(bounds[-1])], dtype=np.int64)))
                hi = np.minimum(4 * new[1:], len(cs) - 1)
                if np.any(cs[hi] - cs[4 * new[:-1]] >= (1 << 32)):
                    continue
            for k, i in enumerate(idx):
                tasks[i]['quad_begin'], tasks[i]['quad_count'] = int(new[k]), int(new[k + 1] - new[k])
        units = np.where(narrow[tasks['chain']], (tasks['quad_count'] + 1) // 2, tasks['quad_count'])
        assert int(units.max()) <= plan.cap_units
        return tasks, int(units.max())

    def _calibrate(self, db, costs, ticks, tick_offset, iters=None):
        """Measured slicing.  The members of a team wait for the slowest one every iteration, and the
        time a slice takes depends on more than the cost model knows (memoised rows built, label mix
        behind the shared-memory reductions, co-resident CTAs).  So: run a few iterations with
        ``task_cycles`` set, then move the slice boundaries inside every team so that the *measured*
        time density is cut into equal parts.  Team sizes and CTA assignment stay (re-planning them
        from measured chain costs was tried and lost: the measurement is only valid for the team
        sizes it was taken with).  Results do not depend on the schedule."""
        torch = self.torch
        plan, T, H = db.plan, db.tensors, db.host
        times = self._measure(db, CALIBRATE_ITERS if iters is None else int(iters))
        plan.tasks, plan.smem_units = self._rebalanced(plan, plan.tasks, times, costs, ticks, tick_offset)
        plan.slice_cap_quads = int(plan.tasks['quad_count'].max())
        db.batch.slice_cap_quads = plan.smem_units
        H['tasks'].copy_(torch.from_numpy(np.ascontiguousarray(plan.tasks.view(np.int32).reshape(-1, 6))))
        T['tasks'].copy_(H['tasks'], non_blocking=True)
        torch.cuda.current_stream(T['tasks'].device).synchronize()
        return times

    # ---- run -----------------------------------------------------------------------------
    def reset(self, db):
        """Zero the exchange workspace and status so the same DeviceBatch can be launched again."""
        db.tensors['exchange'].zero_()
        db.tensors['status'].zero_()
        if db.progress is not None:
            self.torch.cuda.current_stream(db.tensors['status'].device).synchronize()
            db.progress.zero_()

    def launch(self, db, stream=None):
        """Enqueue the whole run.  Long calibrated runs go as several back-to-back launches
        (``db.segments``) so that the slicing can follow the slowly drifting cost of the slices; the
        call then returns once the last segment is enqueued (it waits for the earlier ones)."""
        torch = self.torch
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream() if stream is None else stream
            if db.segments and len(db.segments) > 2:
                with torch.cuda.stream(s):
                    self._launch_segmented(db, s)
                return
            rc = self.lib.brta_gibbs_run_batch(C.byref(db.batch), C.c_void_p(s.cuda_stream))
        _cabi.check(rc, 'brta_gibbs_run_batch')

    def _side_stream(self):
        if getattr(self, '_side', None) is None:
            with self.torch.cuda.device(self.device):
                self._side = self.torch.cuda.Stream()
        return self._side

    def _launch_segmented(self, db, s):
        """Launch k runs iterations ends[k-1]+1 .. ends[k] from the state launch k-1 left in (final_c, final_a).
        Launches 0 and 1 use the current slicing; launch k >= 2 uses a slicing re-cut from the task cycles
        of launch k-2, computed on the host while launch k-1 runs (two task buffers, two cycle buffers), so
        the GPU never waits for the host.  Philox counters, row indices and integer statistics are those of
        the single launch: same bits."""
        torch = self.torch
        T, plan = db.tensors, db.plan
        costs, ticks, tick_offset = db.cal
        ends = db.segments
        dev = T['tasks'].device
        n_tasks = len(plan.tasks)
        if 'tasks_b' not in T or T['tasks_b'].shape != T['tasks'].shape:
            T['tasks_b'] = torch.empty_like(T['tasks'])
            T['cycles'] = torch.zeros((2, n_tasks), dtype=torch.int64, device=dev)
        T['init_c'].copy_(T['init_c0'])
        T['init_a'].copy_(T['init_a0'])
        bufs = (T['tasks'], T['tasks_b'])
        used = [None] * len(ends)                      # (host tasks, smem units) each launch ran with
        events = []
        cur = (plan.tasks, plan.smem_units)
        begin = 0
        for k, end in enumerate(ends):
            which = 0 if k < 2 else (k & 1) ^ 1        # launches 0, 1: buffer 0; then the buffer launch k-2 used
            if k >= 2:
                events[k - 2].synchronize()
                with torch.cuda.stream(self._side_stream()):        # launch k-1 is running on s: stay off it
                    times = T['cycles'][k & 1].cpu().numpy().astype(np.float64)
                    cur = self._rebalanced(plan, used[k - 2][0], times, costs, ticks, tick_offset, SEGMENT_DAMPING, SEGMENT_MIN_SPREAD)
                    bufs[which].copy_(torch.from_numpy(np.ascontiguousarray(cur[0].view(np.int32).reshape(-1, 6))))
                    ready = torch.cuda.Event()
                    ready.record()
                s.wait_event(ready)
            elif k == 0:
                bufs[0].copy_(torch.from_numpy(np.ascontiguousarray(cur[0].view(np.int32).reshape(-1, 6))))
            used[k] = cur
            if k > 0:
                T['init_c'].copy_(T['final_c'])
                T['init_a'].copy_(T['final_a'])
                T['exchange'].zero_()
            T['cycles'][k & 1].zero_()
            seg = _cabi.Batch.from_buffer_copy(db.batch)
            seg.iter_begin, seg.iter_end = int(begin), int(end)
            seg.tasks = bufs[which].data_ptr()
            seg.slice_cap_quads = int(cur[1])
            seg.task_cycles = T['cycles'][k & 1].data_ptr()
            rc = self.lib.brta_gibbs_run_batch(C.byref(seg), C.c_void_p(s.cuda_stream))
            _cabi.check(rc, 'brta_gibbs_run_batch (segment)')
            ev = torch.cuda.Event()
            ev.record(s)
            events.append(ev)
            begin = end
        # keep the latest slicing for the next run of this batch
        plan.tasks, plan.smem_units = cur
        plan.slice_cap_quads = int(plan.tasks['quad_count'].max())
        db.batch.slice_cap_quads = plan.smem_units
        db.host['tasks'].copy_(torch.from_numpy(np.ascontiguousarray(plan.tasks.view(np.int32).reshape(-1, 6))))
        if bufs[which] is not T['tasks']:
            T['tasks'], T['tasks_b'] = T['tasks_b'], T['tasks']
            db.batch.tasks = T['tasks'].data_ptr()

    # ---- host <-> device legs, separately callable so they can be timed / overlapped -------
    def upload(self, db):
        """H2D of every input from its pinned host copy (async on the current stream)."""
        for name, pinned in db.host.items():
            db.tensors[name].copy_(pinned, non_blocking=True)
        return db.h2d_bytes

    def alloc_host_outputs(self, db):
        """Pinned host buffers for the results of ``db`` (reused across runs)."""
        torch = self.torch
        return {name: torch.empty(db.tensors[name].shape, dtype=db.tensors[name].dtype, pin_memory=True)
                for name in ('mcweights', 'mcrates', 'indicator', 'status')}

    def download(self, db, host_out):
        """D2H of the results into pinned buffers (async on the current stream); returns bytes."""
        n = 0
        for name, dst in host_out.items():
            dst.copy_(db.tensors[name], non_blocking=True)
            n += dst.numel() * dst.element_size()
        return n

    def download_through_ring(self, db):
        """D2H of every result through the ring of pinned staging buffers (async on the current stream; the
        bytes are not kept): what a batch too large for one pinned host buffer costs to bring home.
        Returns bytes."""
        torch = self.torch
        ring = self._ring()
        flat = db.tensors['indicator']
        n = 0
        for k, a in enumerate(range(0, int(flat.numel()), self.STREAM_CHUNK_BYTES)):
            b = min(a + self.STREAM_CHUNK_BYTES, int(flat.numel()))
            ring[k % len(ring)][:b - a].copy_(flat[a:b], non_blocking=True)
            n += b - a
        for name in ('mcweights', 'mcrates', 'status'):
            raw = db.tensors[name].reshape(-1).view(torch.uint8)
            for k, a in enumerate(range(0, int(raw.numel()), self.STREAM_CHUNK_BYTES)):
                b = min(a + self.STREAM_CHUNK_BYTES, int(raw.numel()))
                ring[k % len(ring)][:b - a].copy_(raw[a:b], non_blocking=True)
            n += int(raw.numel())
        return n

    def results_from_host(self, db, host_out):
        """Views (no copy) of pinned result buffers as per-chain arrays."""
        mcw, mcr = host_out['mcweights'].numpy(), host_out['mcrates'].numpy()
        ind, status = host_out['indicator'].numpy(), host_out['status'].numpy()
        res = []
        for r, n in enumerate(db.n_data):
            o = int(db.ind_offset[r])
            res.append(ChainResult(mcweights=mcw[r, :db.rows], mcrates=mcr[r, :db.rows],
                                   indicator=ind[o:o + db.rows * int(n)].reshape(db.rows, int(n)),
                                   status=int(status[r])))
        return res

    def fetch(self, db, out=None):
        """D2H of the results.  Returns a list of :class:`ChainResult`."""
        torch = self.torch
        T = db.tensors
        torch.cuda.synchronize(self.device)
        status = T['status'].cpu().numpy()
        mcw = T['mcweights'].cpu().numpy()
        mcr = T['mcrates'].cpu().numpy()
        ind = T['indicator'].cpu().numpy()
        tr_nk = T['trace_nk'].cpu().numpy() if 'trace_nk' in T else None
        tr_tk = T['trace_tk'].cpu().numpy() if 'trace_tk' in T else None
        res = []
        for r, n in enumerate(db.n_data):
            o = int(db.ind_offset[r])
            res.append(ChainResult(
                mcweights=mcw[r, :db.rows].copy(), mcrates=mcr[r, :db.rows].copy(),
                indicator=ind[o:o + db.rows * int(n)].reshape(db.rows, int(n)),
                status=int(status[r]),
                trace_nk=None if tr_nk is None else tr_nk[r],
                trace_tk=None if tr_tk is None else tr_tk[r]))
        return res

    # ---- overlapped output path ----------------------------------------
    STREAM_CHUNK_BYTES = 128 << 20
    STREAM_RING = 6

    def _ring(self):
        """Pinned staging buffers, allocated once per engine (page-locking memory is slow: ~0.3 s/GB)."""
        if getattr(self, '_ring_lock', None) is None:
            import threading
            self._ring_lock = threading.Lock()
        if getattr(self, '_ring_bufs', None) is None:
            torch = self.torch
            self._ring_bufs = [torch.empty(self.STREAM_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True)
                               for _ in range(self.STREAM_RING)]
            with torch.cuda.device(self.device):
                self._copy_stream = torch.cuda.Stream()
        return self._ring_bufs

    def stream_results(self, db, on_chain, pool, dest=None, keep_on_device=False):
        """D2H of a launched batch, overlapped with whatever ``on_chain`` does with a finished chain
        (normally: assign the arrays to the ``Gibbs`` object and write its pickle).

        The dense label buffer (chain r = ``rows * N_r`` contiguous bytes) is copied in chunks through a
        ring of pinned staging buffers on a copy stream; worker threads of ``pool`` (a
        ``ThreadPoolExecutor``) scatter each chunk into the per-chain destination arrays ``dest[r]``
        (``uint8 [rows, N_r]``, allocated here if None) and call ``on_chain(r, ChainResult)`` once the
        last byte of chain r has landed.  ``keep_on_device``: no label bytes cross PCIe at all;
        ``ChainResult.indicator`` is then a CUDA tensor view ``[rows, N_r]`` of the resident buffer
        (SURVEY.md 8 f-1).  Returns the list of futures of the ``on_chain`` calls."""
        import queue
        import threading
        torch = self.torch
        T = db.tensors
        R = len(db.n_data)
        rows = db.rows
        sizes = np.array([rows * int(n) for n in db.n_data], dtype=np.int64)
        offs = np.asarray(db.ind_offset, dtype=np.int64)
        if dest is None and not keep_on_device:
            dest = [np.empty((rows, int(n)), dtype=np.uint8) for n in db.n_data]
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()                       # the sampler is done
            status = T['status'].cpu().numpy()
            mcw = T['mcweights'].cpu().numpy()
            mcr = T['mcrates'].cpu().numpy()

        def result(r):
            if keep_on_device:
                o = int(offs[r])
                ind = T['indicator'][o:o + int(sizes[r])].view(rows, int(db.n_data[r]))
            else:
                ind = dest[r]
            return ChainResult(mcweights=mcw[r, :rows].copy(), mcrates=mcr[r, :rows].copy(), indicator=ind,
                               status=int(status[r]))

        if keep_on_device or int(sizes.sum()) == 0:
            return [pool.submit(on_chain, r, result(r)) for r in range(R)]

        ring = self._ring()
        self._ring_lock.acquire()                                           # released when the last chunk has been scattered
        free = queue.Queue()                                                # indices of staging buffers not in use
        for k in range(len(ring)):
            free.put(k)
        remaining = sizes.copy()
        lock = threading.Lock()
        futures = []
        total = int(sizes.sum())
        flat = T['indicator']
        ends = offs + sizes

        def scatter(slot, a, b, ev):
            try:
                ev.synchronize()
                src = ring[slot].numpy()
                first = int(np.searchsorted(ends, a, side='right'))
                done = []
                for r in range(first, R):
                    lo, hi = max(a, int(offs[r])), min(b, int(ends[r]))
                    if lo >= hi:
                        if int(offs[r]) >= b:
                            break
                        continue
                    np.copyto(dest[r].reshape(-1)[lo - int(offs[r]):hi - int(offs[r])], src[lo - a:hi - a])
                    with lock:
                        remaining[r] -= hi - lo
                        if remaining[r] == 0:
                            done.append(r)
            finally:
                free.put(slot)
            for r in done:
                on_chain(r, result(r))

        try:
            with torch.cuda.device(self.device):
                for a in range(0, total, self.STREAM_CHUNK_BYTES):
                    b = min(a + self.STREAM_CHUNK_BYTES, total)
                    slot = free.get()
                    with torch.cuda.stream(self._copy_stream):
                        ring[slot][:b - a].copy_(flat[a:b], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record()
                    futures.append(pool.submit(scatter, slot, a, b, ev))
        except BaseException:
            self._ring_lock.release()
            raise
        for r in np.nonzero(sizes == 0)[0]:
            futures.append(pool.submit(on_chain, int(r), result(int(r))))
        chunk_futures = list(futures)

        def release():
            for f in chunk_futures:
                f.exception()                                               # wait; errors surface through the caller's result()
            self._ring_lock.release()
        threading.Thread(target=release, daemon=True).start()
        return futures

    def start_live_stream(self, db, on_chain, pool, dest=None, progress=None, sinks=None):
        """Start bringing results home WHILE the sweep runs (call right before ``launch``; ``db`` must have been
        prepared with ``progress_rows``).  Returns a :class:`LiveStream`; its ``finish()`` waits for the
        launch, copies what is left and returns the futures of the ``on_chain`` calls.  ``sinks[r]`` = ``(fd, offset)``
        sends chain r's labels straight into a file (``os.pwrite`` from the staging buffer: 2-3 times the
        throughput of stores into a memory map of the same pages) instead of into ``dest[r]``."""
        return LiveStream(self, db, on_chain, pool, dest=dest, progress=progress, sinks=sinks)

    def run(self, chains, ncomp, niter, thin=100, seed=0, flags=0, inject=None, n_waves=None):
        db = self.prepare(chains, ncomp, niter, thin=thin, seed=seed, flags=flags, inject=inject,
                          n_waves=n_waves)
        self.launch(db)
        return self.fetch(db)


class LiveStream:
    """Output path overlapped with the sweep.

    The kernel publishes, per chain, how many saved rows are complete (``brta_batch.progress`` in mapped pinned
    memory).  A poller thread turns newly finished row blocks -- contiguous byte ranges of the dense label
    buffer -- into device-to-host copies on a copy stream, packed into the engine's ring of pinned staging
    buffers; worker threads of ``pool`` scatter each staged buffer into the per-chain arrays and call
    ``on_chain(r, ChainResult)`` (normally: write the pickle) when the last byte of chain r has landed.  When the
    launch ends only the rows since the last publication are left to copy."""

    POLL_SECONDS = 0.003
    MIN_FLUSH_BYTES = 8 << 20

    def __init__(self, engine, db, on_chain, pool, dest=None, progress=None, sinks=None):
        import queue
        import threading
        if db.progress is None:
            raise ValueError('the batch was prepared without progress_rows')
        self.eng, self.db, self.on_chain, self.pool, self.report = engine, db, on_chain, pool, progress
        torch = engine.torch
        self.rows = db.rows
        self.n = np.asarray(db.n_data, dtype=np.int64)
        self.offs = np.asarray(db.ind_offset, dtype=np.int64)
        self.sinks = sinks if sinks is not None else [None] * len(self.n)
        self.dest = list(dest) if dest is not None else [None] * len(self.n)
        for r, n in enumerate(self.n):
            if self.dest[r] is None and self.sinks[r] is None:
                self.dest[r] = np.empty((self.rows, int(n)), dtype=np.uint8)
        self.done_rows = np.zeros(len(self.n), dtype=np.int64)
        self.remaining = self.rows * self.n
        self.lock = threading.Lock()
        self.futures = []
        self.ring = engine._ring()
        engine._ring_lock.acquire()                         # one batch at a time uses the engine's staging ring
        self.free = queue.Queue()
        for k in range(len(self.ring)):
            self.free.put(k)
        self.finished = []                                  # chains whose labels are all home (results need mcw / mcr)
        self.stop = threading.Event()
        self.error = None
        with torch.cuda.device(engine.device):
            self.launched = torch.cuda.Event()
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    # -- staging --------------------------------------------------------------------------
    def _flush(self, slot, pieces, used):
        torch = self.eng.torch
        with torch.cuda.stream(self.eng._copy_stream):
            ev = torch.cuda.Event()
            ev.record()
        self.futures.append(self.pool.submit(self._scatter, slot, pieces, ev))

    def _scatter(self, slot, pieces, ev):
        try:
            ev.synchronize()
            src = self.ring[slot].numpy()
            done = []
            for r, lo, pos, nbytes in pieces:
                if self.sinks[r] is not None:
                    fd, base = self.sinks[r]
                    view, at = memoryview(src[pos:pos + nbytes]), base + lo
                    while len(view):                                 # pwrite may write less than asked
                        k = os.pwrite(fd, view, at)
                        view, at = view[k:], at + k
                else:
                    np.copyto(self.dest[r].reshape(-1)[lo:lo + nbytes], src[pos:pos + nbytes])
                with self.lock:
                    self.remaining[r] -= nbytes
                    if self.remaining[r] == 0:
                        done.append(r)
        finally:
            self.free.put(slot)
        with self.lock:
            self.finished.extend(done)

    def _copy_rows(self, upto):
        """Enqueue D2H of rows [done_rows[r], upto[r]) of every chain, packed into staging buffers."""
        torch = self.eng.torch
        flat = self.db.tensors['indicator']
        chunk = self.eng.STREAM_CHUNK_BYTES
        slot, pieces, used = None, [], 0
        with torch.cuda.device(self.eng.device), torch.cuda.stream(self.eng._copy_stream):
            for r in np.nonzero(upto > self.done_rows)[0]:
                lo = int(self.done_rows[r] * self.n[r])
                hi = int(upto[r] * self.n[r])
                while lo < hi:
                    if slot is None:
                        slot, pieces, used = self.free.get(), [], 0
                    take = min(hi - lo, chunk - used)
                    a = int(self.offs[r]) + lo
                    self.ring[slot][used:used + take].copy_(flat[a:a + take], non_blocking=True)
                    pieces.append((int(r), lo, used, take))
                    used += take
                    lo += take
                    if used == chunk:
                        self._flush(slot, pieces, used)
                        slot = None
                self.done_rows[r] = upto[r]
            if slot is not None:
                self._flush(slot, pieces, used)

    def _poll(self):
        import time
        try:
            prog = self.db.progress.numpy()
            while not self.stop.is_set():
                now = np.minimum(prog.astype(np.int64), self.rows)          # snapshot of the device's words
                ready = int(((now - self.done_rows) * self.n).sum())
                if ready >= self.MIN_FLUSH_BYTES:
                    self._copy_rows(now)
                    if self.report is not None:
                        self.report(int(now.min()) * (self.db.niter // max(self.rows, 1)), self.db.niter)
                time.sleep(self.POLL_SECONDS)
        except BaseException as e:                                          # surfaced by finish()
            self.error = e

    # -- end of the launch -------------------------------------------------------------------
    def finish(self):
        """Wait for the sweep, bring the rest home, hand every chain to ``on_chain``.  Returns the futures."""
        torch = self.eng.torch
        T = self.db.tensors
        with torch.cuda.device(self.eng.device):
            torch.cuda.current_stream().synchronize()                       # the sampler is done
        self.stop.set()
        self.thread.join()
        if self.error is not None:
            self.eng._ring_lock.release()
            raise self.error
        with torch.cuda.device(self.eng.device):
            status = T['status'].cpu().numpy()
            mcw = T['mcweights'].cpu().numpy()
            mcr = T['mcrates'].cpu().numpy()
        try:
            self._copy_rows(np.full(len(self.n), self.rows, dtype=np.int64))
            for fut in list(self.futures):
                fut.result()                                                # all labels are home
        finally:
            self.eng._ring_lock.release()
        rows = self.rows

        def deliver(r):
            self.on_chain(r, ChainResult(mcweights=mcw[r, :rows].copy(), mcrates=mcr[r, :rows].copy(),
                                         indicator=self.dest[r], status=int(status[r])))
        return [self.pool.submit(deliver, r) for r in range(len(self.n))]


def shard_bounds(n_data, n_shards):
    """Quad boundaries of a chain of ``n_data`` data cut into ``n_shards`` contiguous shards of (nearly) equal
    length; even quad counts keep every shard 16-byte aligned.  Returns n_shards + 1 quad indices."""
    nq = (int(n_data) + 3) // 4
    per = -(-nq // int(n_shards))
    per += per & 1
    bounds = [min(g * per, nq) for g in range(int(n_shards) + 1)]
    if bounds[-2] >= nq:
        raise ValueError('chain too short to shard over that many devices')
    return bounds


def run_sharded(chain, ncomp, niter, devices, thin=100, seed=0, flags=0, inject=None):
    """ONE chain with its residence times sharded over several GPUs of a box (config C4 of
    BASELINE.json), all GPUs driven by THIS process: GPU g sweeps a contiguous range of the data; once
    per iteration the GPUs exchange their integer (n_k, sum tick_k) through tagged words written into
    each other's memory over NVLink, inside the persistent kernels -- there is no host round trip and
    no NCCL call per iteration.  Every GPU then draws the identical posterior update from the same
    Philox key.  Results are bit-identical to the single-GPU run (Philox is keyed by the global datum
    index; the statistics are integers).  :func:`run_sharded_dist` is the one-process-per-GPU form.

    Returns one :class:`ChainResult` for the whole chain."""
    import torch
    devices = [int(d) for d in devices]
    G = len(devices)
    if not 1 <= G <= _cabi.MAX_SHARDS:
        raise ValueError(f'1..{_cabi.MAX_SHARDS} devices')
    engines = [get_engine(d) for d in devices]
    lib = _cabi.load()
    for a in devices:
        for bdev in devices:
            if a != bdev:
                _cabi.check(lib.brta_enable_peer_access(a, bdev), 'brta_enable_peer_access')
    bounds = shard_bounds(len(chain.ticks), G)
    mailboxes = [torch.zeros(_cabi.shard_mailbox_bytes(G), dtype=torch.uint8, device=f'cuda:{d}') for d in devices]
    ptrs = np.array([m.data_ptr() for m in mailboxes], dtype=np.int64)
    batches = []
    for g, (d, eng) in enumerate(zip(devices, engines)):
        with torch.cuda.device(d):
            table = torch.from_numpy(ptrs).to(f'cuda:{d}')
            shard = dict(rank=g, n_shards=G, quad_begin=bounds[g], quad_count=bounds[g + 1] - bounds[g],
                         mailbox_table=table)
            batches.append(eng.prepare([chain], ncomp, niter, thin=thin, seed=seed, flags=flags, inject=inject,
                                       shard=shard))
    for d in devices:
        torch.cuda.synchronize(d)
    for eng, db in zip(engines, batches):                         # asynchronous launches: all kernels run concurrently
        eng.launch(db)
    parts = [eng.fetch(db)[0] for eng, db in zip(engines, batches)]
    del mailboxes
    first = parts[0]
    status = 0
    for p in parts:
        status |= p.status
    canonical = np.concatenate([p.indicator for p in parts], axis=1)     # shards hold canonical (sorted) order
    indicator = np.empty_like(canonical)
    indicator[:, batches[0].order[0]] = canonical
    return ChainResult(mcweights=first.mcweights, mcrates=first.mcrates, indicator=indicator, status=status,
                       trace_nk=first.trace_nk, trace_tk=first.trace_tk)


class ShardedChain:
    """This rank's part of ONE chain sharded over the ranks of a ``torch.distributed`` group, one
    process and one GPU per rank (``torchrun``): the one-process-per-GPU form of :func:`run_sharded`.

    Every rank allocates its mailbox with ``brta_shard_mailbox_create`` and maps its peers' through
    CUDA IPC handles exchanged with ONE ``all_gather`` at set-up; after that the ranks only meet
    inside the kernels (tagged words over NVLink, SURVEY.md 8e: an allreduce of 2K integers per
    iteration).  Collective: every rank constructs it with the same chain and parameters.

    ``launch()`` is asynchronous; ``fetch()`` returns this rank's :class:`ChainResult` whose
    ``indicator`` holds the rank's shard of the data in CANONICAL (ascending-tick) order --
    columns ``data_begin : data_end`` of the canonical order (``order`` maps back)."""

    def __init__(self, chain, ncomp, niter, thin=100, seed=0, flags=0, group=None, device=None, inject=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not 1 <= self.world <= _cabi.MAX_SHARDS:
            raise ValueError(f'1..{_cabi.MAX_SHARDS} ranks')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.engine = get_engine(self.device)
        self.lib = _cabi.load()
        G = self.world
        own = C.c_void_p()
        handle = C.create_string_buffer(_cabi.IPC_HANDLE_BYTES)
        _cabi.check(self.lib.brta_shard_mailbox_create(self.device, G, C.byref(own), handle), 'brta_shard_mailbox_create')
        self._own = own.value
        # the only host-side collective: everybody learns everybody's mailbox handle
        mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8)
        on_gpu = dist.get_backend(group) == 'nccl'
        if on_gpu:
            mine = mine.to(f'cuda:{self.device}')
        gathered = [torch.empty_like(mine) for _ in range(G)]
        dist.all_gather(gathered, mine, group=group)
        self._opened = []
        ptrs = []
        for g in range(G):
            if g == self.rank:
                ptrs.append(self._own)
                continue
            peer = C.c_void_p()
            raw = bytes(gathered[g].cpu().numpy().tobytes())
            _cabi.check(self.lib.brta_shard_mailbox_open(self.device, raw, C.byref(peer)), 'brta_shard_mailbox_open')
            self._opened.append(peer.value)
            ptrs.append(peer.value)
        bounds = shard_bounds(len(chain.ticks), G)
        with torch.cuda.device(self.device):
            table = torch.from_numpy(np.array(ptrs, dtype=np.int64)).to(f'cuda:{self.device}')
            shard = dict(rank=self.rank, n_shards=G, quad_begin=bounds[self.rank],
                         quad_count=bounds[self.rank + 1] - bounds[self.rank], mailbox_table=table)
            self.db = self.engine.prepare([chain], ncomp, niter, thin=thin, seed=seed, flags=flags, inject=inject,
                                          shard=shard)
        self.order = self.db.order[0]
        self.data_begin = 4 * bounds[self.rank]
        self.data_end = min(4 * bounds[self.rank + 1], len(chain.ticks))
        self.niter = int(niter)
        self._launches = 0

    def launch(self, before=None, after=None):
        """Enqueue this rank's kernel.  The G kernels wait for each other on the device, so every rank must
        call it (a host barrier first keeps a late rank from eating into the watchdog).  The mailbox tags
        are iteration numbers, so a re-run first zeroes the mailboxes (between two barriers: no rank may
        still be reading, none may already be writing).  ``before`` / ``after`` are called right around
        the launch itself (after the barriers), e.g. to record timing events on the stream."""
        torch = self.torch
        torch.cuda.synchronize(self.device)
        if self._launches:
            self.dist.barrier(group=self.group)
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.brta_shard_mailbox_clear(self.device, C.c_void_p(self._own), self.world,
                                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                            'brta_shard_mailbox_clear')
                self.engine.reset(self.db)
            torch.cuda.synchronize(self.device)
        self._launches += 1
        self.dist.barrier(group=self.group)
        torch.cuda.synchronize(self.device)                          # the barrier's own kernel is done
        with torch.cuda.device(self.device):
            if before is not None:
                before()
            self.engine.launch(self.db)
            if after is not None:
                after()

    def fetch(self):
        return self.engine.fetch(self.db)[0]

    def close(self):
        self.torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)                           # nobody unmaps memory a peer still writes to
        for p in self._opened:
            _cabi.check(self.lib.brta_shard_mailbox_close(self.device, C.c_void_p(p)), 'brta_shard_mailbox_close')
        self._opened = []
        self.dist.barrier(group=self.group)
        if self._own is not None:
            _cabi.check(self.lib.brta_shard_mailbox_destroy(self.device, C.c_void_p(self._own)), 'brta_shard_mailbox_destroy')
            self._own = None


def run_sharded_dist(chain, ncomp, niter, thin=100, seed=0, flags=0, group=None, device=None):
    """Collective convenience wrapper of :class:`ShardedChain`: run, fetch, close.  Returns
    (this rank's ChainResult, (data_begin, data_end) of its shard in canonical order, order)."""
    sc = ShardedChain(chain, ncomp, niter, thin=thin, seed=seed, flags=flags, group=group, device=device)
    try:
        sc.launch()
        res = sc.fetch()
    finally:
        sc.close()
    return res, (sc.data_begin, sc.data_end), sc.order


_engines = {}


def pindicator_counts(indicator, cluster_of, n_clusters, device=0):
    """Cluster-membership counts of every datum (the accumulation loop of
    ``Gibbs.cluster``, basicrta/gibbs.py:264-268) on the GPU.

    ``indicator``: uint8 [S, N] label rows -- a NumPy array (copied to the device) or a CUDA
    tensor (used in place, e.g. a view of the sampler's ``indicator`` output that never left the GPU);
    ``cluster_of``: int8 [S, K], the mixture label of (row, component) or -1 if that pair is below the
    weight cutoff.  Returns int32 counts [N, n_clusters] as a NumPy array."""
    torch = _torch()
    lib = _cabi.load()
    dev = torch.device('cuda', int(device))
    if isinstance(indicator, np.ndarray):
        if indicator.dtype != np.uint8 or indicator.ndim != 2:
            raise ValueError('indicator must be uint8 [S, N]')
        ind = torch.from_numpy(np.ascontiguousarray(indicator)).to(dev, non_blocking=True)
    else:
        ind = indicator
        if ind.dtype != torch.uint8 or ind.dim() != 2 or ind.stride(1) != 1 or ind.device != dev:
            raise ValueError(f'indicator must be a uint8 [S, N] tensor with unit column stride on {dev}')
    n_rows, n_data = int(ind.shape[0]), int(ind.shape[1])
    cmap = np.ascontiguousarray(cluster_of, dtype=np.int8)
    if cmap.ndim != 2 or cmap.shape[0] != n_rows:
        raise ValueError('cluster_of must be int8 [S, K] with one row per indicator row')
    if not 1 <= int(n_clusters) <= 32:
        raise ValueError('n_clusters must be in 1..32')
    counts = torch.zeros((n_data, int(n_clusters)), dtype=torch.int32, device=dev)
    if n_rows == 0 or n_data == 0:
        return counts.cpu().numpy()
    cmap_d = torch.from_numpy(cmap).to(dev)
    with torch.cuda.device(dev):
        rc = lib.brta_pindicator_counts(C.c_void_p(ind.data_ptr()), int(ind.stride(0)) if n_rows > 1 else max(n_data, 1),
                                        n_rows, n_data, C.c_void_p(cmap_d.data_ptr()), int(cmap.shape[1]),
                                        int(n_clusters), C.c_void_p(counts.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _cabi.check(rc, 'brta_pindicator_counts')
    return counts.cpu().numpy()


def get_engine(device=0):
    if device not in _engines:
        _engines[device] = GibbsEngine(device)
    return _engines[device]
