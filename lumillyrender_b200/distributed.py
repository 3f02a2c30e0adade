"""spp-range sharding across the GPUs of one box.

Samples are independent and counter-indexed, so rank r of G renders sample indices
[r*spp/G, (r+1)*spp/G) of EVERY pixel with the full scene replicated, accumulates per-pixel sums in a
W*H*3 fp32 buffer, and the buffers are summed with ONE reduce to rank 0 (NCCL over NVLink on GPUs,
gloo in the CPU tests); rank 0 divides by the total spp.  One process per GPU (torchrun).
The per-rank renderer is injected so that the host logic is testable without a GPU.
"""
import os


def shard_range(total, rank, world):
    """[begin, end) of `total` sample indices owned by `rank`; ranges tile [0,total) exactly, sizes differ by <= 1."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request total=%r rank=%r world=%r" % (total, rank, world))
    per, rem = divmod(total, world)
    begin = rank * per + min(rank, rem)
    return begin, begin + per + (1 if rank < rem else 0)


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def render_sharded(accumulate, accum, spp_total, rank, world, dist=None, dst=0):
    """Renders this rank's share of `spp_total` into the zeroed tensor `accum` through
    `accumulate(spp_begin, spp_count)` (which ADDS per-pixel sums into accum), reduces to `dst` and
    normalises there.  Returns the mean image on rank dst, None elsewhere."""
    begin, end = shard_range(spp_total, rank, world)
    if end > begin:
        accumulate(begin, end - begin)
    if world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    if rank != dst:
        return None
    # `estimated_sum / spp as f32` (main.rs:104) after the cross-GPU sum.  A tensor divisor keeps it a true
    # division (torch turns division by a Python scalar into a multiplication by the reciprocal on CUDA).
    import torch
    accum.div_(torch.full((), float(spp_total), dtype=accum.dtype, device=accum.device))
    return accum
