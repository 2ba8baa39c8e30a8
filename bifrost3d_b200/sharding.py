"""Sample-index sharding across GPUs.

The value of a sample is a pure function of (pixel, accumulation index, scene): the RNG is counter based
(Types.h:452-459, RNG.h:280-292). Rank r of G therefore renders its own contiguous block of accumulation indices into a
local fp64 sum buffer; one SUM reduce combines the blocks. Index 0 (pixel-centre sample, SimpleRGPs.cu:68) belongs to rank 0.
"""


def sample_range(rank, samples_per_rank, warmup=0):
    """(first accumulation index, count) of `rank`, leaving `warmup` untimed indices in front of each block."""
    block = samples_per_rank + warmup
    return rank * block + warmup, samples_per_rank
