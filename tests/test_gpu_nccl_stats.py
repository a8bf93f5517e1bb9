"""ofdm_stats_allreduce (the path's only collective) over a real NCCL communicator: needs two GPUs, so it is
skipped on the single-GPU test box and run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl_stats.py -m gpu`."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)            # only to hand the NCCL unique id around
    import ofdm_b200 as ob
    nccl = C.CDLL("libnccl.so.2")                                           # the NCCL torch has loaded, else the system one
    uid = NcclUniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    blob = [bytes(uid.internal)]
    dist.broadcast_object_list(blob, src=0)
    C.memmove(C.byref(uid), blob[0], 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclUniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    eng = ob.Engine(ob.Config(modulation=2, guard_bands=True, fec=True), rank)
    mine = np.array([10 + rank, 3 * rank + 1, 1000 * (rank + 1), rank], np.uint64)
    host = eng.stats_allreduce(mine.copy(), comm.value)                      # OFDM_MEM_HOST
    dev = torch.from_numpy(mine.astype(np.int64)).cuda()
    eng._check(eng.lib.ofdm_stats_allreduce(eng._h, dev.data_ptr(), comm, ob.engine.MEM_DEVICE, None), "ofdm_stats_allreduce")
    torch.cuda.synchronize()
    q.put((rank, host.tolist(), dev.cpu().tolist()))
    dist.barrier()
    nccl.ncclCommDestroy(comm)
    eng.close()
    dist.destroy_process_group()


def test_stats_allreduce_over_nccl():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, 31500 + os.getpid() % 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [10 + 11, 1 + 4, 1000 + 2000, 0 + 1]
    for _, host, dev in res:
        assert host == want and dev == want
