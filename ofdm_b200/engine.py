"""ctypes binding of the engine's C ABI (include/ofdm_engine.h -> ofdm_b200/libofdm_b200.so).

This is the same binding a Rust `ofdm-sys` shim performs. There is no CPU fallback: if the
shared library is missing, or there is no CUDA device, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _build

MOD_BPSK, MOD_QPSK, MOD_QAM64 = 0, 1, 2
SYNC_REFERENCE, SYNC_SCHMIDL_COX = 0, 1
CFO_REFERENCE, CFO_ANGLE_OF_SUM = 0, 1
PHASE_REFERENCE, PHASE_ANGLE_OF_SUM = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
OK, TOO_SHORT, NO_SYNC, BAD_HEADER, NEG_OFFSET = 0, 1, 2, 3, 4
STATUS_NAMES = {OK: "OK", TOO_SHORT: "TOO_SHORT", NO_SYNC: "NO_SYNC", BAD_HEADER: "BAD_HEADER", NEG_OFFSET: "NEG_OFFSET"}

EXPORTS = [
    "ofdm_abi_version", "ofdm_cfg_default", "ofdm_status_name", "ofdm_engine_create", "ofdm_engine_destroy",
    "ofdm_last_error", "ofdm_get_tables", "ofdm_host_alloc", "ofdm_host_free", "ofdm_coded_len",
    "ofdm_frame_data_syms", "ofdm_frame_len", "ofdm_max_payload", "ofdm_tx_encode_batch", "ofdm_rx_decode_batch",
    "ofdm_channel_apply_batch", "ofdm_ber_accumulate", "ofdm_kernel_launches", "ofdm_profile_begin", "ofdm_profile_read",
    "ofdm_sync_search", "ofdm_sync_counts", "ofdm_engine_reserve", "ofdm_rx_decode_capture", "ofdm_rx_decode_file",
    "ofdm_stats_allreduce", "ofdm_last_h2d_bytes",
    "ofdm_rs_encoded_len", "ofdm_rs_decoded_len", "ofdm_rs_encode_batch", "ofdm_rs_decode_batch",
]


class EngineError(RuntimeError):
    pass


class CCfg(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("nfft", C.c_uint32), ("cp", C.c_uint32), ("modulation", C.c_uint32),
        ("guard_bands", C.c_uint32), ("fec", C.c_uint32), ("sync_mode", C.c_uint32), ("cfo_mode", C.c_uint32),
        ("phase_mode", C.c_uint32), ("sync_window", C.c_uint32),
        ("locking", C.c_void_p), ("preamble", C.c_void_p), ("training", C.c_void_p),
    ]


class CRxDiag(C.Structure):
    _fields_ = [
        ("offset", C.c_void_p), ("f_delta", C.c_void_p), ("h_k", C.c_void_p), ("n_data_syms", C.c_void_p),
        ("points", C.c_void_p), ("points_stride", C.c_uint32),
    ]


class CChannelParams(C.Structure):
    _fields_ = [
        ("snr_db", C.c_float), ("cfo_max", C.c_float), ("lead_min", C.c_uint32), ("lead_max", C.c_uint32),
        ("multipath", C.c_uint32), ("noise_mode", C.c_uint32), ("seed", C.c_uint64),
    ]


PEAK_DTYPE = np.dtype([("offset", np.uint64), ("f_delta", np.float32), ("metric", np.float32)])      # = ofdm_peak
FRAME_DTYPE = np.dtype([("offset", np.uint64), ("f_delta", np.float32), ("metric", np.float32), ("status", np.int32), ("out_len", np.uint32)])   # = ofdm_frame_info

_lib = None


def load_library(build: bool = False) -> C.CDLL:
    """dlopen the engine. `build=True` (re)builds it with nvcc first when stale."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        _build.build_engine()
    if not os.path.exists(_build.LIB_PATH):
        raise EngineError(
            f"{_build.LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback.")
    L = C.CDLL(_build.LIB_PATH)
    vp, u32, i32, u64, sz = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64, C.c_size_t
    L.ofdm_abi_version.restype = u32
    L.ofdm_cfg_default.argtypes = [C.POINTER(CCfg)]
    L.ofdm_status_name.argtypes = [C.c_int32]
    L.ofdm_status_name.restype = C.c_char_p
    L.ofdm_engine_create.argtypes = [C.POINTER(CCfg), i32, C.POINTER(vp)]
    L.ofdm_engine_create.restype = i32
    L.ofdm_engine_destroy.argtypes = [vp]
    L.ofdm_last_error.argtypes = [vp]
    L.ofdm_last_error.restype = C.c_char_p
    L.ofdm_get_tables.argtypes = [vp, vp, vp, vp]
    L.ofdm_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.ofdm_host_alloc.restype = i32
    L.ofdm_host_free.argtypes = [vp]
    for f in (L.ofdm_coded_len, L.ofdm_frame_data_syms, L.ofdm_frame_len, L.ofdm_max_payload):
        f.argtypes = [C.POINTER(CCfg), u32]
        f.restype = u32
    L.ofdm_tx_encode_batch.argtypes = [vp, vp, vp, u32, u32, vp, u32, vp, i32, vp]
    L.ofdm_tx_encode_batch.restype = i32
    L.ofdm_rx_decode_batch.argtypes = [vp, vp, vp, u32, u32, u32, vp, u32, vp, vp, C.POINTER(CRxDiag), i32, vp]
    L.ofdm_rx_decode_batch.restype = i32
    L.ofdm_channel_apply_batch.argtypes = [vp, vp, vp, u32, u32, C.POINTER(CChannelParams), vp, u32, vp, vp, vp, i32, vp]
    L.ofdm_channel_apply_batch.restype = i32
    L.ofdm_ber_accumulate.argtypes = [vp, vp, vp, u32, vp, vp, u32, vp, u32, vp, i32, vp]
    L.ofdm_ber_accumulate.restype = i32
    L.ofdm_profile_begin.argtypes = [vp, u32]
    L.ofdm_profile_begin.restype = i32
    L.ofdm_profile_read.argtypes = [vp, vp, vp, vp]
    L.ofdm_profile_read.restype = i32
    L.ofdm_sync_search.argtypes = [vp, vp, u64, vp, u32, vp, i32, vp]
    L.ofdm_sync_search.restype = i32
    L.ofdm_rx_decode_file.argtypes = [vp, C.c_char_p, u64, u64, u32, u32, vp, u32, vp, u32, vp]
    L.ofdm_rx_decode_file.restype = i32
    L.ofdm_stats_allreduce.argtypes = [vp, vp, vp, i32, vp]
    L.ofdm_stats_allreduce.restype = i32
    L.ofdm_sync_counts.argtypes = [vp, vp, vp]
    L.ofdm_sync_counts.restype = i32
    L.ofdm_engine_reserve.argtypes = [vp, u32, u64]
    L.ofdm_engine_reserve.restype = i32
    L.ofdm_rx_decode_capture.argtypes = [vp, vp, u64, vp, u32, u32, vp, u32, vp, vp, i32, vp]
    L.ofdm_rx_decode_capture.restype = i32
    for f in (L.ofdm_rs_encoded_len, L.ofdm_rs_decoded_len):
        f.argtypes = [sz]
        f.restype = sz
    L.ofdm_rs_encode_batch.argtypes = [vp, vp, vp, u32, u32, vp, u32, vp, i32, vp]
    L.ofdm_rs_encode_batch.restype = i32
    L.ofdm_rs_decode_batch.argtypes = [vp, vp, vp, u32, u32, vp, u32, vp, vp, vp, i32, vp]
    L.ofdm_rs_decode_batch.restype = i32
    L.ofdm_kernel_launches.argtypes = [vp]
    L.ofdm_kernel_launches.restype = u64
    L.ofdm_last_h2d_bytes.argtypes = [vp]
    L.ofdm_last_h2d_bytes.restype = u64
    _lib = L
    return L


@dataclass(frozen=True)
class Config:
    """Engine configuration = the optional arguments of the reference's encode!/decode! macros plus the batch knobs."""
    modulation: int = MOD_BPSK          # src/transmitter.rs:17
    guard_bands: bool = False           # src/transmitter.rs:16
    fec: bool = False
    sync_mode: int = SYNC_REFERENCE
    cfo_mode: int = CFO_REFERENCE
    phase_mode: int = PHASE_REFERENCE
    sync_window: int = 0
    nfft: int = 64
    cp: int = 16

    def to_c(self) -> CCfg:
        c = CCfg()
        load_library().ofdm_cfg_default(C.byref(c))
        c.nfft, c.cp = self.nfft, self.cp
        c.modulation, c.guard_bands, c.fec = int(self.modulation), int(self.guard_bands), int(self.fec)
        c.sync_mode, c.cfo_mode, c.phase_mode = int(self.sync_mode), int(self.cfo_mode), int(self.phase_mode)
        c.sync_window = int(self.sync_window)
        return c

    @property
    def bits_per_carrier(self) -> int:
        return (1, 2, 6)[self.modulation]

    @property
    def data_carriers(self) -> int:
        if self.nfft == 1024:
            return 768 if self.guard_bands else 1024
        return 48 if self.guard_bands else 64

    @property
    def sym_len(self) -> int:
        return self.nfft + self.cp

    def coded_len(self, n: int) -> int:
        return int(load_library().ofdm_coded_len(C.byref(self.to_c()), n))

    def frame_data_syms(self, n: int) -> int:
        return int(load_library().ofdm_frame_data_syms(C.byref(self.to_c()), n))

    def frame_len(self, n: int) -> int:
        return int(load_library().ofdm_frame_len(C.byref(self.to_c()), n))

    def max_payload(self, n_data_syms: int) -> int:
        return int(load_library().ofdm_max_payload(C.byref(self.to_c()), n_data_syms))


@dataclass
class ChannelParams:
    snr_db: float = 30.0                # src/channel.rs:40
    cfo_max: float = -1.0               # < 0: no CFO
    lead_min: int = 0
    lead_max: int = 0
    multipath: bool = True
    noise_mode: int = 0
    seed: int = 1

    def to_c(self) -> CChannelParams:
        return CChannelParams(self.snr_db, self.cfo_max, self.lead_min, self.lead_max, int(self.multipath),
                              self.noise_mode, self.seed)


@dataclass
class RxResult:
    data: list                 # per stream: bytes (empty unless status OK)
    status: np.ndarray         # int32 [n]
    out_len: np.ndarray        # uint32 [n]
    out: np.ndarray            # uint8 [n, out_stride]
    offset: Optional[np.ndarray] = None
    f_delta: Optional[np.ndarray] = None
    h_k: Optional[np.ndarray] = None         # complex64 [n, 64]
    n_data_syms: Optional[np.ndarray] = None
    points: Optional[np.ndarray] = None      # complex64 [n, points_stride]


def _ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(None)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class Engine:
    """One handle per GPU (not thread-safe), mirroring ofdm_engine_create/destroy."""

    def __init__(self, cfg: Config = Config(), device: int = 0):
        self.lib = load_library()
        self.cfg = cfg
        self._ccfg = cfg.to_c()
        h = C.c_void_p()
        rc = self.lib.ofdm_engine_create(C.byref(self._ccfg), device, C.byref(h))
        if rc != 0:
            raise EngineError(f"ofdm_engine_create failed ({rc}): {self.lib.ofdm_last_error(None).decode()}")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ofdm_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed ({rc}): {self.lib.ofdm_last_error(self._h).decode()}")

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.ofdm_kernel_launches(self._h))

    @property
    def last_h2d_bytes(self) -> int:
        """Host -> device bytes of the last host-mode rx_decode call (the cyclic prefixes may have stayed on the host)."""
        return int(self.lib.ofdm_last_h2d_bytes(self._h))

    def profile_begin(self, max_calls: int):
        self._check(self.lib.ofdm_profile_begin(self._h, max_calls), "ofdm_profile_begin")

    def profile_read(self, max_calls: int):
        """Returns (acquire_ms[], decode_ms[]) of the rx_decode_device calls since profile_begin."""
        a = np.zeros(max_calls, np.float32)
        d = np.zeros(max_calls, np.float32)
        n = C.c_uint32(0)
        self._check(self.lib.ofdm_profile_read(self._h, _ptr(a), _ptr(d), C.byref(n)), "ofdm_profile_read")
        return a[: n.value], d[: n.value]

    def tables(self):
        lock = np.zeros(self.cfg.sym_len, np.complex64)
        pre = np.zeros(self.cfg.sym_len, np.complex64)
        tr = np.zeros(self.cfg.nfft, np.complex64)
        self._check(self.lib.ofdm_get_tables(self._h, _ptr(lock), _ptr(pre), _ptr(tr)), "ofdm_get_tables")
        return lock, pre, tr

    # ---- host-buffer API (numpy) --------------------------------------------------------------------------
    def tx_encode(self, payloads: Sequence[bytes], iq_stride: Optional[int] = None):
        """Batch `encode` (src/transmitter.rs:11-58). Returns (iq complex64 [n, iq_stride], frame_len uint32 [n])."""
        n = len(payloads)
        lens = np.array([len(p) for p in payloads], np.uint32)
        pstride = max(1, int(lens.max()))
        pay = np.zeros((n, pstride), np.uint8)
        for i, p in enumerate(payloads):
            pay[i, : len(p)] = np.frombuffer(bytes(p), np.uint8)
        need = max(self.cfg.frame_len(int(x)) for x in lens)
        if iq_stride is None:
            iq_stride = need
        iq = np.zeros((n, iq_stride), np.complex64)
        flen = np.zeros(n, np.uint32)
        self._check(self.lib.ofdm_tx_encode_batch(self._h, _ptr(pay), _ptr(lens), pstride, n, _ptr(iq), iq_stride,
                                                  _ptr(flen), MEM_HOST, None), "ofdm_tx_encode_batch")
        return iq, flen

    def rx_decode(self, iq: np.ndarray, n_samples=None, out_stride: Optional[int] = None, diag: bool = False,
                  points: bool = False) -> RxResult:
        """Batch `decode` (src/receiver.rs:9-96). iq: complex64 [n, iq_stride] (fc32)."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        if iq.ndim == 1:
            iq = iq[None, :]
        n, stride = iq.shape
        if n_samples is None:
            n_samples = np.full(n, stride, np.uint32)
        n_samples = np.ascontiguousarray(n_samples, np.uint32)
        if out_stride is None:
            out_stride = max(16, (stride // self.cfg.sym_len) * self.cfg.data_carriers * self.cfg.bits_per_carrier // 8)
        out = np.zeros((n, out_stride), np.uint8)
        out_len = np.zeros(n, np.uint32)
        status = np.zeros(n, np.int32)
        res = RxResult([], status, out_len, out)
        d = None
        if diag or points:
            res.offset = np.zeros(n, np.int32)
            res.f_delta = np.zeros(n, np.float32)
            res.h_k = np.zeros((n, self.cfg.nfft), np.complex64)
            res.n_data_syms = np.zeros(n, np.uint32)
            ps = 0
            if points:
                ps = (stride // self.cfg.sym_len + 1) * self.cfg.data_carriers
                res.points = np.zeros((n, ps), np.complex64)
            d = CRxDiag(_ptr(res.offset), _ptr(res.f_delta), _ptr(res.h_k), _ptr(res.n_data_syms),
                        _ptr(res.points) if points else None, ps)
        self._check(self.lib.ofdm_rx_decode_batch(self._h, _ptr(iq), _ptr(n_samples), n, stride, int(n_samples.max()),
                                                  _ptr(out), out_stride, _ptr(out_len), _ptr(status),
                                                  C.byref(d) if d is not None else None, MEM_HOST, None),
                    "ofdm_rx_decode_batch")
        res.data = [bytes(out[i, : out_len[i]]) if status[i] == OK else b"" for i in range(n)]
        return res

    def channel(self, tx: np.ndarray, tx_len=None, params: ChannelParams = ChannelParams(), rx_stride: Optional[int] = None):
        """Batched, seeded `channel` (src/channel.rs:33-74). Returns (rx complex64 [n, rx_stride], rx_len, lead, cfo)."""
        tx = np.ascontiguousarray(tx, dtype=np.complex64)
        if tx.ndim == 1:
            tx = tx[None, :]
        n, stride = tx.shape
        if tx_len is None:
            tx_len = np.full(n, stride, np.uint32)
        tx_len = np.ascontiguousarray(tx_len, np.uint32)
        if rx_stride is None:
            rx_stride = int(tx_len.max()) + 63 + params.lead_max
        rx = np.zeros((n, rx_stride), np.complex64)
        rx_len = np.zeros(n, np.uint32)
        lead = np.zeros(n, np.uint32)
        cfo = np.zeros(n, np.float32)
        cp = params.to_c()
        self._check(self.lib.ofdm_channel_apply_batch(self._h, _ptr(tx), _ptr(tx_len), stride, n, C.byref(cp), _ptr(rx),
                                                      rx_stride, _ptr(rx_len), _ptr(lead), _ptr(cfo), MEM_HOST, None),
                    "ofdm_channel_apply_batch")
        return rx, rx_len, lead, cfo

    def sync_search(self, iq: np.ndarray, max_peaks: int = 4096) -> np.ndarray:
        """Preamble search over one long capture (complex64). Returns a structured array (offset, f_delta, metric)."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        peaks = np.zeros(max_peaks, PEAK_DTYPE)
        n = C.c_uint32(0)
        self._check(self.lib.ofdm_sync_search(self._h, _ptr(iq), iq.size, _ptr(peaks), max_peaks, C.byref(n), MEM_HOST, None),
                    "ofdm_sync_search")
        return peaks[: n.value].copy()

    def decode_capture(self, iq: np.ndarray, max_frame_samples: int = 0, out_stride: int = 4096, max_peaks: int = 4096):
        """Streaming receiver (examples/jetson_rx.rs loop): find every frame of one long capture and decode them all.
        Returns (peaks, list of payload bytes, status array)."""
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        peaks = self.sync_search(iq, max_peaks)
        n = len(peaks)
        out = np.zeros((max(n, 1), out_stride), np.uint8)
        out_len = np.zeros(max(n, 1), np.uint32)
        status = np.zeros(max(n, 1), np.int32)
        if n:
            self._check(self.lib.ofdm_rx_decode_capture(self._h, _ptr(iq), iq.size, _ptr(peaks), n, max_frame_samples, _ptr(out), out_stride,
                                                        _ptr(out_len), _ptr(status), MEM_HOST, None), "ofdm_rx_decode_capture")
        data = [bytes(out[i, : out_len[i]]) if status[i] == OK else b"" for i in range(n)]
        return peaks, data, status[:n]

    def sync_search_device(self, iq_ptr, n_samples, peaks_ptr, max_peaks, n_peaks_ptr, stream=0):
        self._check(self.lib.ofdm_sync_search(self._h, iq_ptr, n_samples, peaks_ptr, max_peaks, n_peaks_ptr, MEM_DEVICE, stream or None),
                    "ofdm_sync_search")

    def decode_file(self, path: str, start: int = 0, stop: int = 0, chunk_samples: int = 0, max_frame_samples: int = 1 << 18,
                    out_stride: int = 1 << 16, max_frames: int = 4096):
        """ofdm_rx_decode_file: every frame of an fc32 capture file (examples/lab3c.rs:57-74 for any length / frame count).
        Returns (frame records [offset, f_delta, metric, status, out_len], list of payload bytes)."""
        frames = np.zeros(max_frames, FRAME_DTYPE)
        out = np.zeros((max_frames, out_stride), np.uint8)
        n = C.c_uint32(0)
        self._check(self.lib.ofdm_rx_decode_file(self._h, os.fsencode(path), start, stop, chunk_samples, max_frame_samples, _ptr(out),
                                                 out_stride, _ptr(frames), max_frames, C.byref(n)), "ofdm_rx_decode_file")
        k = n.value
        return frames[:k].copy(), [bytes(out[i, : frames["out_len"][i]]) if frames["status"][i] == OK else b"" for i in range(k)]

    def stats_allreduce(self, counters: np.ndarray, nccl_comm: int) -> np.ndarray:
        """ofdm_stats_allreduce on host counters (4 x uint64); nccl_comm is the raw ncclComm_t."""
        c = np.ascontiguousarray(counters, np.uint64)
        self._check(self.lib.ofdm_stats_allreduce(self._h, _ptr(c), C.c_void_p(nccl_comm), MEM_HOST, None), "ofdm_stats_allreduce")
        return c

    def sync_counts(self, stream=0):
        """(threshold crossings, frames detected, entries written to peaks[], overflowed tiles) of the last sync search;
        waits for `stream`."""
        c = (C.c_uint32 * 4)()
        self._check(self.lib.ofdm_sync_counts(self._h, c, stream or None), "ofdm_sync_counts")
        return int(c[0]), int(c[1]), int(c[2]), int(c[3])

    def reserve(self, max_streams: int, max_capture_samples: int = 0):
        self._check(self.lib.ofdm_engine_reserve(self._h, max_streams, max_capture_samples), "ofdm_engine_reserve")

    def decode_capture_device(self, iq_ptr, n_samples, peaks_ptr, n_frames, max_frame_samples, out_ptr, out_stride, out_len_ptr,
                              status_ptr, stream=0):
        self._check(self.lib.ofdm_rx_decode_capture(self._h, iq_ptr, n_samples, peaks_ptr, n_frames, max_frame_samples, out_ptr,
                                                    out_stride, out_len_ptr, status_ptr, MEM_DEVICE, stream or None),
                    "ofdm_rx_decode_capture")

    def ber(self, ref: np.ndarray, ref_len, got: np.ndarray, got_len, status) -> np.ndarray:
        """Batch utils::Analysis (src/utils.rs:45-68) -> [bit_errs, byte_errs, bits_compared, frames_failed]."""
        ref = np.ascontiguousarray(ref, np.uint8)
        got = np.ascontiguousarray(got, np.uint8)
        ref_len = np.ascontiguousarray(ref_len, np.uint32)
        got_len = np.ascontiguousarray(got_len, np.uint32)
        status = np.ascontiguousarray(status, np.int32)
        counters = np.zeros(4, np.uint64)
        self._check(self.lib.ofdm_ber_accumulate(self._h, _ptr(ref), _ptr(ref_len), ref.shape[1], _ptr(got), _ptr(got_len),
                                                 got.shape[1], _ptr(status), ref.shape[0], _ptr(counters), MEM_HOST, None),
                    "ofdm_ber_accumulate")
        return counters

    # ---- Reed-Solomon outer code (src/utils.rs:97-137, 152-180) ---------------------------------------
    def rs_encode(self, payloads: Sequence[bytes]):
        """create_transmission_bytes for a batch -> (coded [n, stride] u8, coded_len [n])."""
        n = len(payloads)
        lens = np.array([len(p) for p in payloads], np.uint32)
        in_stride = max(1, int(lens.max()))
        out_stride = int(self.lib.ofdm_rs_encoded_len(int(lens.max())))
        data = np.zeros((n, in_stride), np.uint8)
        for i, p in enumerate(payloads):
            data[i, :len(p)] = np.frombuffer(bytes(p), np.uint8)
        coded = np.zeros((n, out_stride), np.uint8)
        coded_len = np.zeros(n, np.uint32)
        self._check(self.lib.ofdm_rs_encode_batch(self._h, _ptr(data), _ptr(lens), n, in_stride, _ptr(coded), out_stride,
                                                  _ptr(coded_len), MEM_HOST, None), "ofdm_rs_encode_batch")
        return coded, coded_len

    def rs_decode(self, coded: np.ndarray, coded_len=None):
        """decipher_transmission_bytes for a batch -> (data [n, stride] u8, data_len, n_corrected, n_failed)."""
        coded = np.ascontiguousarray(coded, np.uint8)
        if coded.ndim == 1:
            coded = coded[None, :]
        n, in_stride = coded.shape
        lens = np.full(n, in_stride, np.uint32) if coded_len is None else np.ascontiguousarray(coded_len, np.uint32)
        if in_stride == 0:
            coded = np.zeros((n, 1), np.uint8)
            in_stride = 1
        out_stride = int(self.lib.ofdm_rs_decoded_len(int(lens.max())))
        data = np.zeros((n, out_stride), np.uint8)
        data_len = np.zeros(n, np.uint32)
        n_corr = np.zeros(n, np.uint32)
        n_fail = np.zeros(n, np.uint32)
        self._check(self.lib.ofdm_rs_decode_batch(self._h, _ptr(coded), _ptr(lens), n, in_stride, _ptr(data), out_stride,
                                                  _ptr(data_len), _ptr(n_corr), _ptr(n_fail), MEM_HOST, None),
                    "ofdm_rs_decode_batch")
        return data, data_len, n_corr, n_fail

    def rs_encode_device(self, data_ptr, data_len_ptr, n_streams, data_stride, coded_ptr, coded_stride, coded_len_ptr, stream=0):
        self._check(self.lib.ofdm_rs_encode_batch(self._h, data_ptr, data_len_ptr, n_streams, data_stride, coded_ptr, coded_stride,
                                                  coded_len_ptr, MEM_DEVICE, stream or None), "ofdm_rs_encode_batch")

    def rs_decode_device(self, coded_ptr, coded_len_ptr, n_streams, coded_stride, data_ptr, data_stride, data_len_ptr,
                         n_corrected_ptr, n_failed_ptr, stream=0):
        self._check(self.lib.ofdm_rs_decode_batch(self._h, coded_ptr, coded_len_ptr, n_streams, coded_stride, data_ptr, data_stride,
                                                  data_len_ptr, n_corrected_ptr, n_failed_ptr, MEM_DEVICE, stream or None),
                    "ofdm_rs_decode_batch")

    # ---- device-pointer API (raw addresses, e.g. torch.Tensor.data_ptr(); `stream` = cudaStream_t handle) ---
    def tx_encode_device(self, payload_ptr, payload_len_ptr, payload_stride, n_streams, iq_ptr, iq_stride,
                         frame_len_ptr=0, stream=0):
        self._check(self.lib.ofdm_tx_encode_batch(self._h, payload_ptr, payload_len_ptr, payload_stride, n_streams, iq_ptr,
                                                  iq_stride, frame_len_ptr or None, MEM_DEVICE, stream or None),
                    "ofdm_tx_encode_batch")

    def rx_decode_device(self, iq_ptr, n_samples_ptr, n_streams, iq_stride, max_n_samples, out_ptr, out_stride,
                         out_len_ptr, status_ptr, stream=0, diag: Optional[CRxDiag] = None):
        self._check(self.lib.ofdm_rx_decode_batch(self._h, iq_ptr, n_samples_ptr, n_streams, iq_stride, max_n_samples,
                                                  out_ptr, out_stride, out_len_ptr, status_ptr,
                                                  C.byref(diag) if diag is not None else None, MEM_DEVICE, stream or None),
                    "ofdm_rx_decode_batch")

    def channel_device(self, tx_ptr, tx_len_ptr, tx_stride, n_streams, params: ChannelParams, rx_ptr, rx_stride,
                       rx_len_ptr, lead_ptr=0, cfo_ptr=0, stream=0):
        cp = params.to_c()
        self._check(self.lib.ofdm_channel_apply_batch(self._h, tx_ptr, tx_len_ptr, tx_stride, n_streams, C.byref(cp), rx_ptr,
                                                      rx_stride, rx_len_ptr, lead_ptr or None, cfo_ptr or None, MEM_DEVICE,
                                                      stream or None), "ofdm_channel_apply_batch")

    def ber_device(self, ref_ptr, ref_len_ptr, ref_stride, got_ptr, got_len_ptr, got_stride, status_ptr, n_streams,
                   counters_ptr, stream=0):
        self._check(self.lib.ofdm_ber_accumulate(self._h, ref_ptr, ref_len_ptr, ref_stride, got_ptr, got_len_ptr, got_stride,
                                                 status_ptr, n_streams, counters_ptr, MEM_DEVICE, stream or None),
                    "ofdm_ber_accumulate")
