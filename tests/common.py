"""Seeded synthetic inputs shared by the tests (SplitMix64, seed 0xB2000254 + config)."""
from __future__ import annotations

import numpy as np

from oracle import bn254_ref as o
from oracle import port

SEED = 0xB2000254


def scalars(n, seed=SEED, edges=True):
    rng = o.SplitMix64(seed)
    ks = [rng.scalar() for _ in range(n)]
    if edges:
        for i, e in enumerate(o.EDGE_SCALARS):
            if i < n:
                ks[i] = e
    return ks


def scalar_bytes(ks):
    return np.frombuffer(b"".join(o.scalar_to_bytes(k) for k in ks), dtype=np.uint8).copy()


def points(n, seed=SEED, threads=4):
    """P[i] = [a_i]G1, Q[i] = [b_i]G2 in gnark layout, plus the scalars."""
    a = scalars(n, seed, edges=False)
    b = scalars(n, seed + 1, edges=False)
    g1, g2 = port.generators()
    P = port.g1_mul_base_batch(g1, scalar_bytes(a), n, threads)
    Q = port.g2_mul_base_batch(g2, scalar_bytes(b), n, threads)
    return P, Q, a, b


def with_infinities(P, Q):
    """Copy with P[0] = inf, Q[1] = inf, both inf at index 2."""
    P, Q = P.copy(), Q.copy()
    P[0:64] = 0
    Q[128:256] = 0
    P[128:192] = 0
    Q[256:384] = 0
    return P, Q
