#!/usr/bin/env python
"""Can the host -> device feed skip the cyclic prefixes? Rate of cudaMemcpy2DAsync (512 useful bytes of every 640-byte symbol)
against the contiguous copy of the same pinned buffer: one big 2-D copy, one 2-D copy per stream (2038 rows), on 1 / 2 CUDA
streams; and of a zero-copy kernel read (torch index over mapped pinned memory) for comparison. Useful GB/s = bytes that arrive."""
import ctypes as C, time
import torch

rt = C.CDLL("libcudart.so.12")
ROWS, W, P = 2038 + 10, 512, 640
NSTREAMS = 1024
N = NSTREAMS * ROWS * P
p = C.c_void_p()
assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(0)) == 0
C.memset(p, 1, N)
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
ss = [torch.cuda.Stream() for _ in range(4)]


def timed(fn, useful):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    return useful / dt / 1e9, dt * 1e3


def contiguous():
    assert rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr()), p, C.c_size_t(N), 1, C.c_void_p(ss[0].cuda_stream)) == 0


def big2d():
    assert rt.cudaMemcpy2DAsync(C.c_void_p(dev.data_ptr()), C.c_size_t(W), p, C.c_size_t(P), C.c_size_t(W), C.c_size_t(NSTREAMS * ROWS), 1,
                                C.c_void_p(ss[0].cuda_stream)) == 0


def per_stream(k, w=W, pitch_dst=W):
    def go():
        for i in range(NSTREAMS):
            s = ss[i % k]
            assert rt.cudaMemcpy2DAsync(C.c_void_p(dev.data_ptr() + i * ROWS * pitch_dst), C.c_size_t(pitch_dst), C.c_void_p(p.value + i * ROWS * P + 8 * (i % 37)),
                                        C.c_size_t(P), C.c_size_t(w), C.c_size_t(ROWS - 1), 1, C.c_void_p(s.cuda_stream)) == 0
    return go


def per_stream_contig(k):
    def go():
        for i in range(NSTREAMS):
            s = ss[i % k]
            assert rt.cudaMemcpyAsync(C.c_void_p(dev.data_ptr() + i * ROWS * P), C.c_void_p(p.value + i * ROWS * P), C.c_size_t(ROWS * P), 1, C.c_void_p(s.cuda_stream)) == 0
    return go


print("contiguous, one copy            : %.1f GB/s useful (%.1f ms)" % timed(contiguous, N))
print("contiguous, one copy per stream : %.1f GB/s useful (%.1f ms)" % timed(per_stream_contig(1), N))
print("2-D 512/640, one copy           : %.1f GB/s useful (%.1f ms)  [x1.25 = contiguous-equivalent]" % timed(big2d, NSTREAMS * ROWS * W))
for k in (1, 2, 4):
    print("2-D 512/640, per stream, %d strm : %.1f GB/s useful (%.1f ms)" % ((k,) + timed(per_stream(k), NSTREAMS * (ROWS - 1) * W)))
