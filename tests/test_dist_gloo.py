"""world_size-2 gloo test of the N>1 host logic: static stream sharding + the BER counter all-reduce.
The per-rank decode runs on the CPU oracle here (no GPU in this container); on GPUs the same logic runs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_batch(oo, n_streams):
    rng = np.random.default_rng(77)
    cfg = oo.make_cfg(True, oo.QAM64, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, 1024)
    pays = [rng.integers(0, 256, 120, dtype=np.uint8) for _ in range(n_streams)]
    caps = [oo.channel(oo.tx(p, cfg), 27.0, 0.02, 1, 50 + i) for i, p in enumerate(pays)]      # low SNR: some bit errors / lost headers
    stride = max(c.size for c in caps)
    iq = np.zeros((n_streams, stride), np.complex64)
    ns = np.zeros(n_streams, np.uint32)
    for i, c in enumerate(caps):
        iq[i, : c.size] = c
        ns[i] = c.size
    ns[3] = 500                                                                                 # one failed stream (TOO_SHORT / NO_SYNC)
    return cfg, pays, iq, ns


def _counters(oo, cfg, pays, iq, ns, lo, hi):
    out, out_len, status, _ = oo.decode_batch_fc32(iq[lo:hi].view(np.float32).reshape(hi - lo, iq.shape[1], 2), ns[lo:hi], cfg, 128, 1)
    c = np.zeros(4, np.int64)
    for i in range(hi - lo):
        p = pays[lo + i]
        if status[i] != 0 or out_len[i] != p.size:
            c += [8 * p.size, p.size, 8 * p.size, 1]
        else:
            e, be, _ = oo.analysis(p, out[i, : p.size])
            c += [e, be, 8 * p.size, 0]
    return c


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as oo
    from ofdm_b200 import dist as od
    cfg, pays, iq, ns = _make_batch(oo, 10)
    lo, hi = od.stream_shard(10, rank, world)
    c = torch.from_numpy(_counters(oo, cfg, pays, iq, ns, lo, hi))
    od.allreduce_counters(c)
    q.put((rank, c.tolist(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_ber_reduction_world2(oo):
    world, port = 2, 29500 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, pays, iq, ns = _make_batch(oo, 10)
    want = _counters(oo, cfg, pays, iq, ns, 0, 10).tolist()
    spans = sorted(r[2] for r in res)
    assert spans == [(0, 5), (5, 10)]
    for _, got, _ in res:
        assert got == want                      # every rank holds the job-wide counters
    assert 1 <= want[3] < 10 and want[2] == 10 * 120 * 8 and want[0] > 0


def _make_capture(oo):
    """One capture with frames all over it -- also right on the shard boundary and inside the overlap region."""
    rng = np.random.default_rng(91)
    cfg = oo.make_cfg(True, oo.QPSK, True)
    n = 400_000
    cap = (0.003 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    # 2 720-sample frames; 200 001 has offset 200 000 = the first sample rank 1 owns (both ranks detect it), 197 000 ends inside
    # rank 1's guard region, 203 100 starts just past what rank 0 reads
    starts = [900, 60_000, 133_000, 197_000, 200_001, 203_100, 260_000, 330_000, 395_000]
    frame_len = None
    for i, p in enumerate(starts):
        tx = oo.tx(rng.integers(0, 256, 150, dtype=np.uint8), cfg)
        frame_len = tx.size
        if p + tx.size > n:
            continue
        cap[p: p + tx.size] += (tx * np.exp(1j * 0.004 * i * np.arange(tx.size))).astype(np.complex64)
    return cap, starts, frame_len


def _capture_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as oo
    from ofdm_b200 import dist as od
    cap, starts, frame_len = _make_capture(oo)
    shard = od.capture_shards(cap.size, world, frame_len)[rank]
    local = oo.sync_search(cap[shard.read_lo: shard.read_hi])              # on GPUs: ofdm_sync_search on the rank's samples
    g, mine = od.owned_peaks(local["offset"], shard)
    parts = od.gather_peaks(g[mine].tolist())
    q.put((rank, parts, int(len(local)), int(mine.sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_one_capture_split_over_two_ranks_finds_every_frame_once(oo):
    """SURVEY.md 8(e) partition 2: one long capture, contiguous sample ranges with a 2 L + frame_len overlap, duplicates
    de-duplicated by ownership of the offset, peak lists gathered. The union equals the search over the whole capture."""
    world, port = 2, 30500 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_capture_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cap, starts, frame_len = _make_capture(oo)
    whole = [int(x) for x in oo.sync_search(cap)["offset"]]
    assert [s - 1 for s in starts if s + frame_len <= cap.size] == whole    # lag - 1 rule, every frame found by the plain search
    for _, parts, n_local, n_mine in res:
        merged = [x for part in parts for x in part]
        assert merged == whole                                              # each frame exactly once, ascending, on every rank
    assert sum(r[2] for r in res) > len(whole)                             # the overlap really produced duplicates ...
    assert sum(r[3] for r in res) == len(whole)                            # ... and ownership removed them
