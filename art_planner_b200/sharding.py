"""Multi-GPU plumbing of the pose-validity path: poses are independent given the read-only map,
so every rank checks its own contiguous shard of the sample stream and the ranks exchange only the ordered indices of
the valid samples -- one count all-gather plus one padded index all-gather (NCCL over NVLink on GPUs; the same code
runs on gloo/CPU tensors, which is what tests/test_sharding_cpu.py exercises)."""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of n_total samples for `rank` of `world`."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_valid_indices(idx, cnt, world: int, group=None, out_idx=None, out_cnt=None):
    """All-gather the per-rank ordered valid-index lists.

    idx: int64 tensor [cap] (first cnt entries meaningful, cap identical on all ranks); cnt: int32 tensor [1].
    Returns (all_idx [world*cap], counts list of [1] tensors). Collective: every rank must call it."""
    import torch
    import torch.distributed as dist
    if out_cnt is None:
        out_cnt = [torch.zeros_like(cnt) for _ in range(world)]
    if out_idx is None:
        out_idx = torch.empty(world * idx.numel(), dtype=idx.dtype, device=idx.device)
    dist.all_gather(out_cnt, cnt, group=group)
    dist.all_gather_into_tensor(out_idx, idx, group=group)
    return out_idx, out_cnt


def merge_gathered(all_idx, counts, cap: int):
    """Concatenate the meaningful prefix of every rank's slice (ranks own increasing index ranges, so the result is
    globally sorted)."""
    import torch
    parts = [all_idx[r * cap: r * cap + int(c.item())] for r, c in enumerate(counts)]
    return torch.cat(parts) if parts else all_idx[:0]
