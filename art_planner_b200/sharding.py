"""Multi-GPU plumbing of the pose-validity path: poses are independent given the read-only map.
Two layouts: (1) pose shards against a replicated map (configs[1]); (2) SPATIAL shards (configs[4]): the map is cut into
row slabs, rank r holds slab r plus a halo (artp_set_map_window) and checks the samples whose x falls into its slab
(slab_window / rank_of_x below). Either way the ranks exchange only verdicts: one all-gather of the bit-packed masks (NCCL
over NVLink on GPUs; the same code runs on gloo/CPU tensors, which is what tests/test_sharding_cpu.py exercises)."""
from __future__ import annotations


def slab_window(rows: int, rank: int, world: int, halo: int, align: int = 4):
    """Row slab [s0, s1) of rank `rank` and the window [lo, hi) it uploads (slab + halo, clipped to the map; lo aligned down
    to `align` rows because artp_set_map_window wants row0 % 4 == 0). halo >= largest box half-diagonal + box offsets, in rows."""
    base, rem = divmod(rows, world)
    s0 = rank * base + min(rank, rem)
    s1 = s0 + base + (1 if rank < rem else 0)
    lo = max(0, s0 - halo)
    lo -= lo % align
    return s0, s1, lo, min(rows, s1 + halo)


def row_of_x(x, cx: float, length_x: float, res: float, rows: int):
    """grid_map row index of a position x (rows run towards -x), clamped."""
    import numpy as np
    return np.clip(np.floor((cx + 0.5 * length_x - np.asarray(x)) / res).astype(np.int64), 0, rows - 1)


def rank_of_x(x, cx: float, length_x: float, res: float, rows: int, world: int):
    """The rank whose slab holds position x (the routing rule of the spatially sharded path)."""
    import numpy as np
    r = row_of_x(x, cx, length_x, res, rows)
    base, rem = divmod(rows, world)
    edge = (base + 1) * rem           # first `rem` slabs have base + 1 rows
    return np.where(r < edge, r // (base + 1), rem + (r - edge) // max(base, 1)).astype(np.int64)


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


def gather_valid_bits(bits, world: int, group=None, out=None):
    """All-gather the per-rank bit-packed validity masks (one collective, no padding, no count exchange).

    bits: int32 tensor [words] (same length on every rank: shards are padded to a multiple of 32 samples by the
    caller, see words_per_shard). Returns int32 [world * words]: rank r's samples are bits [r*words*32, ...)."""
    import torch
    import torch.distributed as dist
    if out is None:
        out = torch.empty(world * bits.numel(), dtype=bits.dtype, device=bits.device)
    dist.all_gather_into_tensor(out, bits, group=group)
    return out


def words_per_shard(n_per_rank: int) -> int:
    return (n_per_rank + 31) // 32


def indices_from_bits(all_bits, n_total: int, world: int):
    """Reference (torch) decoding of a gathered bit mask into ordered global sample indices: rank r owns
    shard_range(n_total, r, world), its words start at r * words_per_shard(cap), cap = ceil(n_total / world).
    The CUDA path does the same with artp_compact_bits_device per rank slice; the gloo tests check this form."""
    import torch
    cap = (n_total + world - 1) // world
    w = words_per_shard(cap)
    out = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        words = all_bits[r * w:(r + 1) * w].to(torch.int64) & 0xFFFFFFFF
        bit = (words[:, None] >> torch.arange(32, device=all_bits.device)[None, :]) & 1
        local = torch.nonzero(bit.reshape(-1)[: hi - lo]).reshape(-1)
        out.append(local + lo)
    return torch.cat(out) if out else torch.zeros(0, dtype=torch.int64)


def pack_bits_reference(valid, cap=None):
    """torch restatement of artp_pack_valid_bits_device (CPU tensors; used by the gloo tests); cap pads the shard."""
    import torch
    n = valid.numel()
    w = words_per_shard(n if cap is None else cap)
    pad = torch.zeros(w * 32, dtype=torch.int64)
    pad[:n] = (valid != 0).to(torch.int64)
    words = (pad.reshape(w, 32) << torch.arange(32)[None, :]).sum(dim=1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32)
