"""Multi-GPU plumbing: one process per GPU, streams sharded statically, no collective on the data path.

A frame's decode depends only on its own samples (src/receiver.rs:9-96 is a pure function of one Vec), so streams
are block-partitioned over the ranks. The only collective is the sum-reduction of the four BER
counters {bit_errs, byte_errs, bits_compared, frames_failed} (the fields of utils::Analysis, src/utils.rs:39-43,
plus a failure count) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, NamedTuple, Tuple


def stream_shard(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r owns streams [r*n/W, (r+1)*n/W)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


class CaptureShard(NamedTuple):
    """One rank's part of a long capture: it READS samples [read_lo, read_hi) and OWNS the frames whose offset lies in
    [own_lo, own_hi). The owned ranges partition the capture, so every frame is reported by exactly one rank."""
    read_lo: int
    read_hi: int
    own_lo: int
    own_hi: int


# A search that starts in the middle of a frame sees a Schmidl-Cox plateau at its very first lags and refines it to some
# offset inside the first 2 * 80 + 176 samples it read (ramp window of docs/SPEC.md 4); a rank therefore starts reading
# `guard` samples before the first offset it owns, and whatever it finds in front of own_lo belongs to its left neighbour.
CAPTURE_GUARD = 512


def capture_shards(n_samples: int, world: int, frame_len: int, sym_len: int = 80, guard: int = CAPTURE_GUARD) -> List[CaptureShard]:
    """Split one long capture over `world` ranks (SURVEY.md 8e partition 2; the reference searches one radio buffer at a
    time, src/receiver.rs:20-21, examples/jetson_rx.rs:46-57). Rank r owns offsets [r n / W, (r + 1) n / W); it reads
    `guard` samples more on the left and 2 L + frame_len more on the right, so a frame that starts at its last owned sample
    still lies wholly inside what it reads. No exchange between ranks: only the final gather of the peak lists."""
    overlap = 2 * sym_len + frame_len
    out = []
    for r in range(world):
        own_lo = (r * n_samples) // world
        own_hi = ((r + 1) * n_samples) // world
        out.append(CaptureShard(max(0, own_lo - guard), min(n_samples, own_hi + overlap), own_lo, own_hi))
    return out


def owned_peaks(local_offsets, shard: CaptureShard):
    """Global offsets of this rank's detections (local offsets are relative to read_lo) and the mask of those it owns --
    the de-duplication of the overlap regions."""
    import numpy as np

    g = np.asarray(local_offsets, dtype=np.int64) + shard.read_lo
    return g, (g >= shard.own_lo) & (g < shard.own_hi)


def gather_peaks(records, group=None):
    """All ranks' owned peak records (any picklable / numpy array), concatenated in rank order = ascending offset order."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [records]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, records, group=group)
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
