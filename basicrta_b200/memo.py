"""Host-side cost model of the memoised-row sweep, used to cut a chain into slices of equal
*time* rather than equal length.

In the kernel a CTA computes, once per iteration, the cumulative rows of the tick values
``lo .. lo + rows - 1`` of its slice (``lo`` = smallest tick of the slice's full quads, ``rows`` =
``TABLE_FLOATS / row_floats(K)``); a quad whose four ticks fall in that window reads its rows, any
other quad recomputes them.  With the chain in ascending-tick order the served quads are a prefix
of the slice, so the cost of a slice is piecewise linear in its length.  Team members wait for each
other every iteration, hence slices must take equal time, not hold equal counts.
"""
import numpy as np

TABLE_FLOATS = 4096                      # must match TABLE_FLOATS in csrc/brta_gibbs.cu
MAX_ROWS = 256                           # table_rows_max of csrc/brta_sweep.cuh: tick - lo fits 8 bits (packed statistics)
COST_SERVED = 0.5                        # quad-times of a quad read from the table (measured optimum on B200)
COST_DIRECT = 1.0                        # quad-times of a recomputed quad


def table_rows(ncomp):
    """Rows of the kernel's table: TABLE_FLOATS / table_row_stride(K) (odd stride, see brta_gibbs.cu)."""
    return min(TABLE_FLOATS // ((int(ncomp) + 3) // 4 * 4 + 1), MAX_ROWS)


class ChainCost:
    """Cost model of one chain given its ticks in canonical (ascending) order."""

    def __init__(self, sorted_ticks, ncomp, enabled=True):
        t = np.asarray(sorted_ticks, dtype=np.int64)
        n = len(t)
        self.nq = (n + 3) // 4
        self.rows = table_rows(ncomp) if enabled else 0
        first = t[0::4]
        last = t[np.minimum(np.arange(self.nq) * 4 + 3, n - 1)]
        self.qmin = first
        # the chain's partial last quad is never served from the table
        self.qmax = last.copy()
        if n % 4:
            self.qmax[-1] = np.iinfo(np.int64).max
        self.n_full = n // 4
        # first quad at or after s that is NOT served by the rows starting at quad s's first tick (the chain is in
        # ascending-tick order, so qmax is sorted): one vectorised search instead of one per query
        self._reach = np.searchsorted(self.qmax, self.qmin + self.rows, side='left') if self.rows else None
        self._splits = {}

    def served(self, s, e):
        """Number of quads of slice [s, e) read from the table (a prefix of the slice)."""
        if self.rows == 0 or s >= self.n_full:
            return 0
        return max(0, min(int(self._reach[s]), e) - s)

    def slice_cost(self, s, e):
        k = self.served(s, e)
        return COST_SERVED * k + COST_DIRECT * (e - s - k)

    def _advance(self, s, tau, cap):
        """Largest e with slice_cost(s, e) <= tau (at least s + 1, at most s + cap)."""
        hi = min(self.nq, s + cap)
        k = self.served(s, hi)
        if tau <= COST_SERVED * k:
            e = s + int(tau / COST_SERVED)
        else:
            e = s + k + int((tau - COST_SERVED * k) / COST_DIRECT)
        return max(s + 1, min(e, hi))

    def _count(self, tau, cap):
        s, m = 0, 0
        while s < self.nq:
            s = self._advance(s, tau, cap)
            m += 1
        return m

    def split(self, c, cap=1 << 30):
        """c + 1 quad boundaries of c slices of (nearly) equal cost, each <= cap quads.  Memoised: the planner asks
        for the same (c, cap) when it prices a team size and again when it cuts the slices, and again for every
        candidate schedule the engine measures."""
        key = (int(min(c, self.nq)), int(cap))
        if key not in self._splits:
            self._splits[key] = self._split(*key)
        return list(self._splits[key])

    def _split(self, c, cap):
        lo, hi = 0.0, COST_DIRECT * self.nq + 1.0
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            if self._count(mid, cap) <= c:
                hi = mid
            else:
                lo = mid
        bounds, s = [0], 0
        while s < self.nq:
            s = self._advance(s, hi, cap)
            bounds.append(s)
        while len(bounds) - 1 < c:                           # fewer slices than members: halve the longest
            i = int(np.argmax(np.diff(bounds)))
            if bounds[i + 1] - bounds[i] < 2:
                break
            bounds.insert(i + 1, (bounds[i] + bounds[i + 1]) // 2)
        return bounds

    def total_cost(self, c, cap=1 << 30):
        b = self.split(c, cap)
        return sum(self.slice_cost(b[i], b[i + 1]) for i in range(len(b) - 1))
