"""fc32 file / ring-buffer ingest in front of the streaming receiver.

The reference reads captures with `bytes_to_sig` (src/utils.rs:238-254: interleaved native-endian f32 pairs -- the format
`rx_samples_to_file --type float` writes, data/receive.sh:1), slices them with `[start..stop]` (examples/lab3c.rs:57-74)
and, on the radio, decodes one frame per 2 M-sample buffer and drops failures (examples/jetson_rx.rs:46-57,83-108).
Here a capture of any length -- a file slice or samples pushed by a radio thread -- is cut into chunks that overlap by one
maximum frame length, each chunk crosses PCIe once from pinned memory and goes through ofdm_sync_search +
ofdm_rx_decode_capture on the device; every frame is reported once with its absolute sample offset.

The chunk arithmetic (`plan_chunks`, `accept_frame`) is plain host logic and is unit-tested without a GPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple

import numpy as np

from . import engine as _e

HEAD_SYMS = 10          # lock | preamble x4 | training x5 (src/transmitter.rs:21-34)


@dataclass
class Frame:
    offset: int          # absolute sample index of the frame start (lag - 1 rule, src/receiver.rs:21)
    f_delta: float       # CFO estimate, rad/sample
    metric: float        # Schmidl-Cox metric at the detection
    status: int          # engine status (0 = OK)
    data: bytes          # decoded payload (empty unless OK)


def fc32_file_samples(path: str) -> int:
    """Number of complex samples in an fc32 file (8 bytes per sample, src/utils.rs:228-254)."""
    return os.path.getsize(path) // 8


def read_fc32(path: str, start: int = 0, stop: Optional[int] = None) -> np.ndarray:
    """`bytes_to_sig(read(path))[start..stop]` (examples/lab3c.rs:57-74) as a complex64 memory map -- nothing is copied."""
    n = fc32_file_samples(path)
    stop = n if stop is None else min(int(stop), n)
    start = min(int(start), stop)
    if stop == start:
        return np.zeros(0, np.complex64)
    return np.memmap(path, dtype=np.complex64, mode="r", offset=8 * start, shape=(stop - start,))


def plan_chunks(n_samples: int, chunk: int, overlap: int) -> List[Tuple[int, int]]:
    """[begin, end) of every chunk of a capture of n_samples: chunks advance by chunk - overlap so that a frame of up to
    `overlap` samples starting anywhere lies wholly inside the chunk that owns its start."""
    if chunk <= overlap:
        raise ValueError("chunk must be longer than the overlap (one maximum frame)")
    out, a = [], 0
    while True:
        b = min(n_samples, a + chunk)
        out.append((a, b))
        if b >= n_samples:
            return out
        a = b - overlap


def accept_frame(local_offset: int, chunk_len: int, overlap: int, final: bool) -> bool:
    """A chunk owns the frames that start before its last `overlap` samples; the final chunk owns everything it sees."""
    return final or local_offset < chunk_len - overlap


class StreamReceiver:
    """Continuous capture -> frames. `push()` samples as they arrive (any block size), `flush()` at the end.

    chunk_samples: device buffer size in samples (default 32 Mi samples = 256 MiB). max_frame_samples: longest frame the
    link carries (its whole length, head included) = the chunk overlap and the per-frame sample cap. out_stride: room per
    decoded payload."""

    def __init__(self, cfg: _e.Config, device: int = 0, chunk_samples: int = 1 << 25, max_frame_samples: int = 1 << 18,
                 out_stride: int = 1 << 16, max_peaks: int = 4096, hold_off: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise _e.EngineError("StreamReceiver needs a CUDA device (there is no CPU fallback)")
        if chunk_samples <= max_frame_samples:
            raise ValueError("chunk_samples must exceed max_frame_samples")
        self.torch = torch
        self.cfg, self.dev = cfg, torch.device("cuda", device)
        self.eng = _e.Engine(cfg, device)
        if hold_off is None:
            hold_off = 10 * cfg.sym_len                                   # lock + preamble + training: one detection per frame
        self.chunk, self.overlap, self.out_stride, self.max_peaks, self.hold_off = chunk_samples, max_frame_samples, out_stride, max_peaks, hold_off
        with torch.cuda.device(self.dev):
            self.h_bufs = [torch.empty((chunk_samples, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
            self.h_iq = self.h_bufs[0]                                  # push() fills this one
            self.d_iq = torch.empty((chunk_samples, 2), dtype=torch.float32, device=self.dev)
            self.d_peaks = torch.zeros((max_peaks, 2), dtype=torch.int64, device=self.dev)          # 16-byte ofdm_peak records
            self.d_npk = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.d_out = torch.zeros((max_peaks, out_stride), dtype=torch.uint8, device=self.dev)
            self.d_len = torch.zeros(max_peaks, dtype=torch.int32, device=self.dev)
            self.d_st = torch.zeros(max_peaks, dtype=torch.int32, device=self.dev)
        self.h_np = self.h_iq.numpy().view(np.complex64).reshape(-1)       # the pinned chunk as complex64
        self.fill = 0              # samples currently in the pinned chunk
        self.base = 0              # absolute index of its first sample
        self.last = None           # absolute offset of the last reported frame
        self.samples_in = 0
        self.bytes_h2d = 0

    def close(self):
        self.eng.close()

    # ---- feeding ------------------------------------------------------------------------------------------------
    def push(self, samples: np.ndarray) -> List[Frame]:
        """Append complex samples (complex64, or anything castable the way sig_to_bytes casts, src/utils.rs:228-236)."""
        out: List[Frame] = []
        x = np.asarray(samples).reshape(-1)
        pos = 0
        while pos < x.size:
            k = min(x.size - pos, self.chunk - self.fill)
            self.h_np[self.fill: self.fill + k] = x[pos: pos + k]      # casts to fc32
            self.fill += k
            pos += k
            self.samples_in += k
            if self.fill == self.chunk:
                out += self._process(final=False)
        return out

    def push_file(self, f, n_samples: int) -> List[Frame]:
        """Read n_samples fc32 samples from a binary file object straight into the pinned chunk (no intermediate copy)."""
        out: List[Frame] = []
        left = n_samples
        raw = self.h_iq.numpy().view(np.uint8).reshape(-1)
        while left > 0:
            k = min(left, self.chunk - self.fill)
            got = f.readinto(memoryview(raw[8 * self.fill: 8 * (self.fill + k)]))
            if not got:
                break
            got //= 8
            self.fill += got
            left -= got
            self.samples_in += got
            if self.fill == self.chunk:
                out += self._process(final=False)
        return out

    def flush(self) -> List[Frame]:
        """Decode what is buffered as the end of the capture."""
        out = self._process(final=True) if self.fill else []
        self.base += self.fill
        self.fill = 0
        return out

    # ---- one chunk ----------------------------------------------------------------------------------------------
    def _process(self, final: bool) -> List[Frame]:
        n = self.fill
        frames = self.process_buffer(self.h_iq, n, self.base, final)
        if not final:
            keep_from = n - self.overlap                               # carry the overlap into the next chunk
            self.h_np[: self.overlap] = self.h_np[keep_from: n].copy()
            self.base += keep_from
            self.fill = self.overlap
        return frames

    def process_buffer(self, h_buf, n: int, base: int, final: bool) -> List[Frame]:
        """One chunk: n samples in the pinned tensor h_buf, whose first sample has absolute index `base`."""
        torch = self.torch
        frames: List[Frame] = []
        if n < HEAD_SYMS * self.cfg.sym_len:
            return frames
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream()
            self.d_iq[:n].copy_(h_buf[:n], non_blocking=True)
            self.bytes_h2d += 8 * n
            self.eng.sync_search_device(self.d_iq.data_ptr(), n, self.d_peaks.data_ptr(), self.max_peaks, self.d_npk.data_ptr(), st.cuda_stream)
            k = min(int(self.d_npk.item()), self.max_peaks)
            if not k:
                return frames
            self.eng.decode_capture_device(self.d_iq.data_ptr(), n, self.d_peaks.data_ptr(), k, self.overlap, self.d_out.data_ptr(),
                                           self.out_stride, self.d_len.data_ptr(), self.d_st.data_ptr(), st.cuda_stream)
            peaks = self.d_peaks[:k].cpu().numpy().view(_e.PEAK_DTYPE).reshape(-1)
            lens = self.d_len[:k].cpu().numpy()
            stat = self.d_st[:k].cpu().numpy()
            ok = stat == _e.OK
            width = int(lens[ok].max()) if ok.any() else 0
            out = self.d_out[:k, :width].cpu().numpy() if width else None      # one device -> host copy for the chunk
        for i in range(k):
            if peaks["metric"][i] < 0 or not accept_frame(int(peaks["offset"][i]), n, self.overlap, final):
                continue
            absolute = base + int(peaks["offset"][i])
            if self.last is not None and absolute < self.last + self.hold_off:
                continue                                               # already reported from the previous chunk's body
            self.last = absolute
            data = out[i, : int(lens[i])].tobytes() if ok[i] else b""
            frames.append(Frame(absolute, float(peaks["f_delta"][i]), float(peaks["metric"][i]), int(stat[i]), data))
        return frames


def _read_into(fd: int, file_offset: int, raw: np.ndarray, pool, pieces: int) -> None:
    """pread `raw.size` bytes at file_offset into the (pinned) byte array, split over the pool's threads."""
    total = raw.size
    step = -(-total // pieces)

    def one(a):
        b = min(total, a + step)
        mv = memoryview(raw[a:b])
        done = 0
        while done < b - a:
            got = os.preadv(fd, [mv[done:]], file_offset + a + done)
            if got <= 0:
                raise IOError("short read")
            done += got

    list(pool.map(one, range(0, total, step)))


def decode_file(path: str, cfg: _e.Config, start: int = 0, stop: Optional[int] = None, device: int = 0,
                receiver: Optional[StreamReceiver] = None, read_threads: int = 8, **kw) -> List[Frame]:
    """Every frame in `bytes_to_sig(read(path))[start..stop]` (examples/lab3c.rs:57-74); offsets are relative to `start`.
    The file is read straight into two pinned chunks by `read_threads` preads while the previous chunk is on the GPU."""
    from concurrent.futures import ThreadPoolExecutor
    n = fc32_file_samples(path)
    stop = n if stop is None else min(int(stop), n)
    start = min(int(start), stop)
    rx = receiver or StreamReceiver(cfg, device, **kw)
    rx.last = None
    frames: List[Frame] = []
    fd = os.open(path, os.O_RDONLY)
    try:
        chunks = plan_chunks(stop - start, rx.chunk, rx.overlap) if stop > start else []
        raws = [b.numpy().view(np.uint8).reshape(-1) for b in rx.h_bufs]
        with ThreadPoolExecutor(read_threads) as readers, ThreadPoolExecutor(1) as ahead:
            def load(i):
                a, b = chunks[i]
                _read_into(fd, 8 * (start + a), raws[i & 1][: 8 * (b - a)], readers, read_threads)

            fut = ahead.submit(load, 0) if chunks else None
            for i, (a, b) in enumerate(chunks):
                fut.result()
                if i + 1 < len(chunks):
                    fut = ahead.submit(load, i + 1)                     # overlaps this chunk's PCIe copy and kernels
                frames += rx.process_buffer(rx.h_bufs[i & 1], b - a, a, i == len(chunks) - 1)
                rx.samples_in += b - a
        return frames
    finally:
        os.close(fd)
        if receiver is None:
            rx.close()
