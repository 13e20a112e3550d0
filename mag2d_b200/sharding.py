"""Particle sharding across GPUs (new: the reference is single-process).

Particles of a species are independent within a step; only the charge grid couples them.  Every rank
owns a contiguous block of each species' particle array, deposits into its own fixed-point grid and
the int64 grids are summed with one all-reduce per step.  Integer sums are associative, so the
reduced grid is bit-identical for any rank count and any split.
"""


def shard_range(n, rank, world):
    """[lo, hi) of the block of n items owned by rank; sizes differ by at most one"""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seed(seed, rank):
    """per-rank Philox seed for the device-side loaders (distinct streams, reproducible)"""
    return (int(seed) * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03 * (rank + 1)) & 0xFFFFFFFFFFFFFFFF
