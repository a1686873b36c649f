"""Philox4x32-10 counter-based generator, NumPy restatement (TEST INFRASTRUCTURE).

This file is part of the CPU oracle.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product path (``basicrta_b200``) never does.

The reference (``/root/reference/basicrta/gibbs.py:17``) draws from an *unseeded*
``numpy.random.default_rng()`` (PCG64), so it defines no stream a device could
share.  The B200 sampler keys every random number by
``(seed, chain, iteration, datum)`` with Philox4x32-10 (Salmon et al., SC'11;
the same round/Weyl constants as ``curand_philox4x32_x.h``).  This module is the
host mirror of that stream; it is pinned by the Random123 known-answer vectors
(``tests/test_philox.py``).

Stream layout (shared with ``basicrta_b200/csrc/brta_rng.cuh``)::

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (x, iteration, chain_id, purpose)

    purpose 0            indicator uniforms: x = datum_index >> 2, datum i uses
                         output word (i & 3)
    purpose 1 + 4*k      Dirichlet gamma of component k, x = rejection trial
    purpose 2 + 4*k      rate gamma of component k,      x = rejection trial
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All arguments broadcast; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *[np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0                      # 32x32 -> 64 bit, exact in uint64
        p1 = M1 * c2
        n0 = (p1 >> _S32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> _S32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def seed_key(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def indicator_words(seed, chain_id, iteration, n):
    """The uint32 word each of the ``n`` data of a chain consumes in ``iteration``."""
    k0, k1 = seed_key(seed)
    nq = (n + 3) // 4
    q = np.arange(nq, dtype=np.uint64)
    w = philox4x32_10(q, iteration, chain_id, 0, k0, k1)
    return np.stack(w, axis=1).reshape(-1)[:n]


def word_to_uniform(x):
    """uint32 -> float32 in [0, 1): top 23 bits, exact (mantissa of a float in [1,2) minus 1)."""
    x = np.asarray(x, dtype=np.uint32)
    return ((x >> np.uint32(9)).astype(np.float32) * np.float32(2.0 ** -23)).astype(np.float32)


def indicator_uniforms(seed, chain_id, iteration, n):
    return word_to_uniform(indicator_words(seed, chain_id, iteration, n))
