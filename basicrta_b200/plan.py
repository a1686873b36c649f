"""Static schedule of chains (residues) onto the persistent grid of one GPU.

The reference fans residues out to a ``multiprocessing.Pool`` with ``chunksize=1``
(basicrta/gibbs.py:73-86): one chain per worker process, greedy dynamic balance.  On a
B200 one chain cannot fill 148 SMs and 400 chains do not divide evenly over them, so the
schedule gives every chain a *team* of CTAs sized in proportion to its cost
(N_r * K * niter; K and niter are common to a batch, so cost ~ N_r):

* chains are grouped into *waves*; inside a wave every CTA of the grid holds at most one
  task, and team sizes are chosen greedily so that the largest per-CTA slice is as small
  as possible -- all teams of a wave then finish together, whatever the mix of sizes;
* the number of waves trades per-iteration rendezvous overhead (small slices) against
  integer rounding of team sizes (few CTAs per chain); it is picked by a cost model;
* all members of a team hold their chain in the same wave and every CTA walks its waves
  in order, which makes the in-kernel rendezvous deadlock-free under a cooperative launch.

Pure NumPy, no GPU needed: the host logic is covered by the CPU test-suite.
"""
import heapq
import os
from dataclasses import dataclass

import numpy as np

TASK_DTYPE = np.dtype([('chain', np.int32), ('team_size', np.int32), ('team_rank', np.int32),
                       ('quad_begin', np.int32), ('quad_count', np.int32), ('order', np.int32)])

MIN_SLICE_QUADS = int(os.environ.get('BRTA_MIN_SLICE_QUADS', 32))   # a CTA has 128 threads; below this most of them idle
DEFAULT_OVERHEAD_QUADS = 192.0   # rendezvous + posterior draw per iteration, in quad-times


@dataclass
class Plan:
    tasks: np.ndarray            # TASK_DTYPE, grouped by CTA, ascending order within a CTA
    cta_task_begin: np.ndarray   # int32 [grid + 1]
    grid: int
    slice_cap_quads: int         # largest quad_count
    n_waves: int
    team_size: np.ndarray        # int32 [R]
    wave_of_chain: np.ndarray    # int32 [R]
    est_efficiency: float        # sum(cost) / (grid * simulated makespan)
    smem_units: int = 0          # dynamic shared memory per CTA in 16-byte units; set by the engine

    def tasks_of_cta(self, b):
        return self.tasks[self.cta_task_begin[b]:self.cta_task_begin[b + 1]]


def _allocate_wave(quads, cmin, grid):
    """Team sizes for one wave: start at the minimum, then hand spare CTAs to the team with
    the largest slice.  Returns (team sizes, largest slice)."""
    team = cmin.copy()
    spare = grid - int(team.sum())
    assert spare >= 0
    heap = [(-int(-(-int(q) // int(c))), i) for i, (q, c) in enumerate(zip(quads, team))]
    heapq.heapify(heap)
    while spare > 0 and heap:
        negd, i = heapq.heappop(heap)
        q, c = int(quads[i]), int(team[i]) + 1
        d = -(-q // c)
        if d < MIN_SLICE_QUADS:
            # the largest slice of the wave is already tiny: more CTAs would only add
            # rendezvous traffic; leave the spare CTAs without a task in this wave
            heapq.heappush(heap, (negd, i))
            break
        team[i] = c
        spare -= 1
        heapq.heappush(heap, (-d, i))
    dmax = max(-(-int(q) // int(c)) for q, c in zip(quads, team))
    return team, dmax


def _assign_ctas(n_quads, cost_q, order, groups, alloc, grid, overhead_quads, splitter):
    """Gang list-scheduling of the teams on the CTAs.  CTAs never synchronise between waves,
    so a CTA that finishes its wave-w task early can start its wave-(w+1) task early: within a
    wave the teams with the longest slices take the CTAs that become free first, which absorbs
    most of the integer-rounding slack of the previous wave.  Time unit: one quad of one
    iteration (+ a fixed per-iteration overhead); a team starts when its last member is free.
    Returns (team sizes, wave of chain, per-CTA task lists, simulated makespan)."""
    R = len(n_quads)
    team_size = np.zeros(R, dtype=np.int32)
    wave_of = np.zeros(R, dtype=np.int32)
    per_cta = [[] for _ in range(grid)]
    free_at = np.zeros(grid, dtype=np.float64)
    for w, ((a, b), team) in enumerate(zip(groups, alloc)):
        members = [(int(order[a + i]), min(int(c), int(n_quads[int(order[a + i])]))) for i, c in enumerate(team)]
        slice_len = {r: -(-int(cost_q[r]) // c) for r, c in members}
        members.sort(key=lambda rc: -slice_len[rc[0]])
        avail = list(np.argsort(free_at, kind='stable'))         # earliest-free CTAs first
        pos = 0
        for r, c in members:
            ctas = avail[pos:pos + c]
            pos += c
            q = int(n_quads[r])
            team_size[r], wave_of[r] = c, w
            if splitter is None:
                base, rem = divmod(q, c)
                bounds = [rank * base + min(rank, rem) for rank in range(c + 1)]
            else:
                bounds = [int(x) for x in splitter(r, c)]
                assert len(bounds) == c + 1 and bounds[0] == 0 and bounds[-1] == q
            start = max(free_at[x] for x in ctas)
            for rank, x in enumerate(ctas):
                cnt = bounds[rank + 1] - bounds[rank]
                assert cnt >= 1
                per_cta[int(x)].append((r, c, rank, bounds[rank], cnt, w))
                free_at[x] = start + slice_len[r] + overhead_quads
        assert pos <= grid
    return team_size, wave_of, per_cta, float(free_at.max())


def _group_contiguous(quads_sorted, cmin_sorted, n_waves, grid):
    """Split the (descending) chain list into n_waves contiguous groups of similar total
    work, each feasible for one wave (count <= grid, sum of minimum teams <= grid)."""
    total = float(quads_sorted.sum())
    groups, start, acc, acc_c = [], 0, 0.0, 0
    target = total / n_waves
    for i, (q, c) in enumerate(zip(quads_sorted, cmin_sorted)):
        full = (acc_c + c > grid) or (i - start >= grid)
        if i > start and (full or (acc + 0.5 * q > target * (len(groups) + 1) and len(groups) < n_waves - 1)):
            groups.append((start, i))
            start, acc_c = i, 0
        acc += float(q)
        acc_c += int(c)
    groups.append((start, len(quads_sorted)))
    return groups


def build_plan(n_quads, grid, cap_quads, overhead_quads=DEFAULT_OVERHEAD_QUADS, n_waves=None,
               cost=None, splitter=None):
    """Schedule chains with ``n_quads[r]`` quads (4 data each) on ``grid`` CTAs whose shared
    memory holds at most ``cap_quads`` quads.

    ``cost[r]`` (optional, in quad-times) is what the time balance uses instead of the quad
    count -- quads served from the memoised rows are cheaper than recomputed ones -- and
    ``splitter(r, c)`` (optional) returns the c+1 quad boundaries of chain r's slices; the
    default cuts equal quad counts."""
    n_quads = np.asarray(n_quads, dtype=np.int64)
    cost_q = n_quads if cost is None else np.maximum(1, np.rint(np.asarray(cost, dtype=np.float64))).astype(np.int64)
    R = len(n_quads)
    if R == 0:
        raise ValueError('empty batch')
    if np.any(n_quads < 1):
        raise ValueError('every chain needs at least one datum')
    grid = int(grid)
    cap_quads = np.broadcast_to(np.asarray(cap_quads, dtype=np.int64), n_quads.shape)   # per chain
    cmin = np.maximum(1, -(-n_quads // cap_quads)).astype(np.int64)
    if cmin.max() > grid:
        raise ValueError(f'a chain needs {int(cmin.max())} CTAs; the grid has {grid}')

    order = np.argsort(-cost_q, kind='stable')
    qs, cs = cost_q[order], cmin[order]
    w_min = max(1, -(-R // grid), -(-int(cs.sum()) // grid))
    if n_waves is not None:
        candidates = [max(int(n_waves), w_min)]
    else:
        w_hi = min(R, max(w_min + 12, 2 * w_min))
        candidates = sorted(set(range(w_min, w_hi + 1)))

    best = None
    for W in candidates:
        groups = _group_contiguous(qs, cs, W, grid)
        alloc = [_allocate_wave(qs[a:b], cs[a:b], grid)[0] for (a, b) in groups]
        span = _assign_ctas(n_quads, cost_q, order, groups, alloc, grid, overhead_quads, None)[3]
        if best is None or span < best[0] - 1e-9:
            best = (span, groups, alloc)
    span, groups, alloc = best

    team_size, wave_of, per_cta, makespan = _assign_ctas(n_quads, cost_q, order, groups, alloc, grid,
                                                         overhead_quads, splitter)

    flat, begin = [], np.zeros(grid + 1, dtype=np.int32)
    for bidx, lst in enumerate(per_cta):
        flat.extend(lst)
        begin[bidx + 1] = len(flat)
    tasks = np.array(flat, dtype=TASK_DTYPE)
    return Plan(tasks=tasks, cta_task_begin=begin, grid=grid,
                slice_cap_quads=int(tasks['quad_count'].max()), n_waves=len(groups),
                team_size=team_size, wave_of_chain=wave_of,
                est_efficiency=float(cost_q.sum()) / (grid * max(makespan, 1.0)))


def density_from_times(bounds, times, quad_weight=None):
    """Per-quad time density of one chain from measured slice times: inside slice i (quads
    bounds[i] .. bounds[i+1]) the time ``times[i]`` is spread over the quads in proportion to
    ``quad_weight`` (model cost per quad; uniform if None)."""
    bounds = np.asarray(bounds, dtype=np.int64)
    nq = int(bounds[-1])
    w = np.ones(nq) if quad_weight is None else np.asarray(quad_weight, dtype=np.float64)
    dens = np.empty(nq)
    for i in range(len(bounds) - 1):
        s, e = int(bounds[i]), int(bounds[i + 1])
        dens[s:e] = w[s:e] * (float(times[i]) / w[s:e].sum())
    return dens


def cut_density(dens, c, cap=None):
    """c+1 quad boundaries that cut a per-quad time density into c parts of equal time, every part
    holding at least one quad.  Returns None if a part would exceed ``cap`` quads."""
    nq = len(dens)
    c = int(min(c, nq))
    cum = np.cumsum(dens)
    targets = cum[-1] * np.arange(1, c) / c
    inner = np.searchsorted(cum, targets, side='left') + 1
    new = np.concatenate(([0], inner, [nq])).astype(np.int64)
    for i in range(1, c):                                    # strictly increasing, room for the rest
        new[i] = min(max(new[i], new[i - 1] + 1), nq - (c - i))
    if cap is not None and int(np.diff(new).max()) > int(cap):
        return None
    return new


def rebalance_team(bounds, times, quad_weight=None, cap=None):
    """New slice boundaries of one team from the measured times of its current slices (same team
    size): the measured time density is cut into equal parts.  Falls back to the old boundaries
    if a slice would exceed ``cap`` quads or the measurement is unusable."""
    bounds = np.asarray(bounds, dtype=np.int64)
    times = np.asarray(times, dtype=np.float64)
    c = len(bounds) - 1
    if c < 2 or int(bounds[-1]) < 2 * c or not np.all(times > 0):
        return bounds.copy()
    new = cut_density(density_from_times(bounds, times, quad_weight), c, cap)
    return bounds.copy() if new is None else new


def shard_chains(costs, n_shards):
    """Longest-processing-time-first assignment of whole chains to GPUs (no collective:
    chains are independent, basicrta/gibbs.py:73-86).  Returns a list of index arrays."""
    costs = np.asarray(costs, dtype=np.float64)
    loads = [(0.0, s) for s in range(n_shards)]
    heapq.heapify(loads)
    shards = [[] for _ in range(n_shards)]
    for r in np.argsort(-costs, kind='stable'):
        load, s = heapq.heappop(loads)
        shards[s].append(int(r))
        heapq.heappush(loads, (load + float(costs[r]), s))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]
