"""Batch partitioning across GPUs: contiguous chunks, no data-path collective.

The only exchange step north_star names -- combining per-GPU partial Miller products of ONE very large
multi-pairing -- is `combine_partials`: a gather of <= 8 x 384 B followed by GT products and one final
exponentiation on a single GPU."""
from __future__ import annotations


def shard_range(n, rank, world):
    """Contiguous [lo, hi) owned by `rank`: ceil(n/world) per rank, the tail ranks may be short or empty."""
    per = -(-n // world)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def shard_pairs_of_product(k, rank, world):
    """Which of the k pairs of one big multi-pairing a rank accumulates."""
    return shard_range(k, rank, world)


def combine_partials(engine, partials):
    """partials: list of (384,) uint8 Miller-loop partial products -> final GT (384,) uint8."""
    acc = partials[0]
    for p in partials[1:]:
        acc = engine.gt_mul_batch(acc, p)[0]
    return engine.final_exp_batch(acc)[0]
