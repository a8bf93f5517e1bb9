#!/usr/bin/env python
"""Concurrent host->device copy ceiling per GPU for different pinned-memory flavours (run under torchrun, one rank per GPU):
default pinned, write-combined, and two copy streams per GPU. Prints one line per rank-0 summary (slowest rank, GB/s)."""
import ctypes as C, os, time
import torch, torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rt = C.CDLL("libcudart.so.12")
N = 2 << 30
dev = torch.empty(N, dtype=torch.uint8, device="cuda")


def bar():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(flags, streams):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(flags)) == 0
    C.memset(p, 1, N)
    ss = [torch.cuda.Stream() for _ in range(streams)]
    part = N // streams

    def go():
        for i, s in enumerate(ss):
            assert rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + i * part), C.c_void_p(p.value + i * part), C.c_size_t(part), 1, C.c_void_p(s.cuda_stream)) == 0
    go(); bar()
    t0 = time.perf_counter()
    for _ in range(4):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 4
    rt.cudaFreeHost(p)
    t = torch.tensor([N / dt / 1e9], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


for name, flags, streams in (("pinned", 0, 1), ("pinned, 2 streams", 0, 2), ("write-combined", 4, 1), ("write-combined, 2 streams", 4, 2), ("portable", 1, 1)):
    v = run(flags, streams)
    if rank == 0:
        print(f"n_gpus {world}: {name:28s} {v:6.1f} GB/s per GPU (slowest rank), {v * world:7.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
