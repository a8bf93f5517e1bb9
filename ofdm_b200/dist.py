"""Multi-GPU plumbing: one process per GPU, streams sharded statically, no collective on the data path.

A frame's decode depends only on its own samples (src/receiver.rs:9-96 is a pure function of one Vec), so streams
are block-partitioned over the ranks. The only collective is the sum-reduction of the four BER
counters {bit_errs, byte_errs, bits_compared, frames_failed} (the fields of utils::Analysis, src/utils.rs:39-43,
plus a failure count) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple


def stream_shard(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r owns streams [r*n/W, (r+1)*n/W)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


def capture_shards(n_samples: int, world: int, frame_len: int, sym_len: int = 80) -> List[Tuple[int, int]]:
    """Split one long capture into `world` contiguous ranges overlapping by 2*L + frame_len samples, so that any
    frame lies wholly inside one rank's range (duplicates in the overlap are de-duplicated by offset)."""
    overlap = 2 * sym_len + frame_len
    out = []
    for r in range(world):
        a = (r * n_samples) // world
        b = min(n_samples, ((r + 1) * n_samples) // world + overlap)
        out.append((a, b))
    return out


def allreduce_counters(counters, group=None):
    """In-place SUM all-reduce of the 4 x int64 BER counters (torch tensor on the backend's device)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def err_rate(counters) -> float:
    """Analysis.err_rate over the whole job (src/utils.rs:61)."""
    c = [int(x) for x in counters]
    return c[0] / c[2] if c[2] else 0.0
