"""Multi-GPU decomposition of the render (DESIGN.md section e).

The path shards naturally: camera samples are independent (integrator.rs:235-245 already exploits this with
tiles) and with the box filter every film pixel only receives its own samples (film.rs:569-573).  Each rank
holds a full scene replica, renders a contiguous range of SAMPLE INDICES of every pixel (distinct
(pixel, sample) RNG streams -> no correlation between ranks) and the f64 films are summed onto rank 0 with
ONE collective per render (NCCL reduce over NVLink on GPUs; gloo in the CPU tests).  No other data-path
communication exists.
"""
from typing import Tuple


def sample_range_for_rank(spp: int, rank: int, world: int, mode: str = "strong") -> Tuple[int, int]:
    """strong: the `spp` samples of the image are split across ranks (remainder to the low ranks);
    weak: every rank renders `spp` NEW sample indices (image ends up with world*spp samples)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if mode == "weak":
        return rank * spp, (rank + 1) * spp
    if mode != "strong":
        raise ValueError(mode)
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_film(film_tensor, dst: int = 0):
    """Sum the (n_pixels, 4) float64 film over all ranks onto `dst` (one collective per render)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_tensor, dst=dst, op=dist.ReduceOp.SUM)
    return film_tensor
