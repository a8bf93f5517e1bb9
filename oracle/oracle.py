"""ctypes front end of the CPU oracle (oracle/libofdm_oracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module. The product (ofdm_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libofdm_oracle.so")

BPSK, QPSK, QAM64 = 0, 1, 2
SYNC_REFERENCE, SYNC_SCHMIDL_COX = 0, 1
CFO_REFERENCE, CFO_ANGLE_OF_SUM = 0, 1
PHASE_REFERENCE, PHASE_ANGLE_OF_SUM = 0, 1
OK, TOO_SHORT, NO_SYNC, BAD_HEADER, NEG_OFFSET = 0, 1, 2, 3, 4


class OoCfg(C.Structure):
    _fields_ = [
        ("guard_bands", C.c_int32),
        ("modulation", C.c_int32),
        ("fec", C.c_int32),
        ("sync_mode", C.c_int32),
        ("cfo_mode", C.c_int32),
        ("phase_mode", C.c_int32),
        ("sync_window", C.c_int32),
        ("xcorr_fft", C.c_int32),
        ("nfft", C.c_int32),
    ]


class OoDiag(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("offset", C.c_int32),
        ("f_delta", C.c_double),
        ("h_k", C.c_double * 128),
        ("n_data_syms", C.c_int64),
        ("n_points", C.c_int64),
        ("packet_length", C.c_uint64),
        ("h_full", C.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with its committed Makefile (gcc)."""
    src = os.path.join(_HERE, "ofdm_oracle.c")
    hdr = os.path.join(_HERE, "ofdm_oracle.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libofdm_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp, sz, i32, u64, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_double
    L.oo_locking_signal.argtypes = [vp, i32]
    L.oo_preamble.argtypes = [vp, i32]
    L.oo_training_signals.argtypes = [vp, i32]
    L.oo_stdrng_uniform_pm1.argtypes = [u64, vp, i32]
    L.oo_fft.argtypes = [vp, sz, i32]
    L.oo_fft_shift.argtypes = [vp, sz]
    L.oo_ifft_shift.argtypes = [vp, sz]
    L.oo_xcorr_fft.argtypes = [vp, sz, vp, sz, vp]
    L.oo_xcorr_fft.restype = sz
    L.oo_convolve.argtypes = [vp, sz, vp, sz, vp]
    L.oo_analysis.argtypes = [vp, vp, sz, vp, vp, vp]
    L.oo_sig_to_fc32.argtypes = [vp, sz, vp]
    L.oo_fc32_to_sig.argtypes = [vp, sz, vp]
    for f in (L.oo_hamming74_encoded_len, L.oo_hamming74_decoded_len):
        f.argtypes = [sz]
        f.restype = sz
    L.oo_hamming74_encode.argtypes = [vp, sz, vp]
    L.oo_hamming74_decode.argtypes = [vp, sz, vp]
    for f in (L.oo_rs_encoded_len, L.oo_rs_decoded_len):
        f.argtypes = [sz]
        f.restype = sz
    L.oo_rs_encode_block.argtypes = [vp, i32, i32, vp]
    L.oo_rs_encode_block.restype = None
    L.oo_rs_correct_block.argtypes = [vp, i32, i32]
    L.oo_rs_correct_block.restype = i32
    L.oo_rs_encode.argtypes = [vp, sz, vp]
    L.oo_rs_encode.restype = None
    L.oo_rs_decode.argtypes = [vp, sz, vp, vp, vp]
    L.oo_rs_decode.restype = i32
    L.oo_modulate.argtypes = [vp, sz, i32, vp]
    L.oo_modulate.restype = sz
    L.oo_demodulate.argtypes = [vp, sz, i32, vp]
    L.oo_demodulate.restype = sz
    L.oo_frame_data_syms.argtypes = [sz, i32, i32]
    L.oo_frame_data_syms.restype = sz
    L.oo_frame_len.argtypes = [sz, i32, i32]
    L.oo_frame_len.restype = sz
    L.oo_encode.argtypes = [vp, sz, i32, i32, vp]
    L.oo_encode.restype = sz
    L.oo_encode_n.argtypes = [vp, sz, i32, i32, i32, vp]
    L.oo_encode_n.restype = sz
    L.oo_frame_len_n.argtypes = [sz, i32, i32, i32]
    L.oo_frame_len_n.restype = sz
    L.oo_frame_data_syms_n.argtypes = [sz, i32, i32, i32]
    L.oo_frame_data_syms_n.restype = sz
    L.oo_tx.argtypes = [vp, sz, C.POINTER(OoCfg), vp]
    L.oo_tx.restype = sz
    L.oo_tx_len.argtypes = [sz, C.POINTER(OoCfg)]
    L.oo_tx_len.restype = sz
    L.oo_channel.argtypes = [vp, sz, dbl, dbl, i32, u64, vp]
    L.oo_decode.argtypes = [vp, sz, C.POINTER(OoCfg), vp, sz, C.POINTER(sz), vp, sz, C.POINTER(OoDiag)]
    L.oo_decode.restype = i32
    L.oo_decode_batch_fc32.argtypes = [vp, vp, C.c_uint32, sz, C.POINTER(OoCfg), vp, sz, vp, vp, vp, i32]
    L.oo_decode_batch_fc32.restype = i32
    L.oo_max_threads.restype = i32
    L.oo_sync_search_fc32.argtypes = [vp, sz, vp, sz]
    L.oo_sync_search_fc32.restype = sz
    L.oo_sync_search_fc32_n.argtypes = [vp, sz, vp, sz, C.c_int]
    L.oo_sync_search_fc32_n.restype = sz
    L.oo_angle.argtypes = [C.c_double * 2]
    _lib = L
    return L


def _p(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def make_cfg(guard_bands=True, modulation=BPSK, fec=False, sync_mode=SYNC_REFERENCE,
             cfo_mode=CFO_REFERENCE, phase_mode=PHASE_REFERENCE, sync_window=0, xcorr_fft=False, nfft=64) -> OoCfg:
    return OoCfg(int(guard_bands), int(modulation), int(fec), int(sync_mode), int(cfo_mode),
                 int(phase_mode), int(sync_window), int(xcorr_fft), int(nfft))


def _c128(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.complex128))


def _bytes(x) -> np.ndarray:
    if isinstance(x, (bytes, bytearray)):
        return np.frombuffer(bytes(x), dtype=np.uint8).copy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint8))


# ---- tables --------------------------------------------------------------------------------
def locking_signal(n=80):
    o = np.zeros(n, np.complex128)
    lib().oo_locking_signal(_p(o), n)
    return o


def preamble(n=80):
    o = np.zeros(n, np.complex128)
    lib().oo_preamble(_p(o), n)
    return o


def training_signals(n=64):
    o = np.zeros(n, np.complex128)
    lib().oo_training_signals(_p(o), n)
    return o


def stdrng_uniform_pm1(seed, n):
    o = np.zeros(n, np.float64)
    lib().oo_stdrng_uniform_pm1(seed, _p(o), n)
    return o


# ---- primitives ----------------------------------------------------------------------------
def fft(x, inverse=False):
    o = _c128(x).copy()
    lib().oo_fft(_p(o), o.size, int(inverse))
    return o


def fft_shift(x):
    o = _c128(x).copy()
    lib().oo_fft_shift(_p(o), o.size)
    return o


def ifft_shift(x):
    o = _c128(x).copy()
    lib().oo_ifft_shift(_p(o), o.size)
    return o


def xcorr_fft(a, b):
    a, b = _c128(a), _c128(b)
    o = np.zeros(2 * a.size - 1, np.complex128)
    idx = lib().oo_xcorr_fft(_p(a), a.size, _p(b), b.size, _p(o))
    return int(idx), o


def convolve(a, b):
    a, b = _c128(a), _c128(b)
    o = np.zeros(a.size + b.size - 1, np.complex128)
    lib().oo_convolve(_p(a), a.size, _p(b), b.size, _p(o))
    return o


def angle(z: complex) -> float:
    return float(np.arctan2(z.imag, z.real))  # same libm atan2 as oo_angle; kept for KAT symmetry


def analysis(left, right):
    l, r = _bytes(left), _bytes(right)
    assert l.size == r.size
    e, be, rate = C.c_uint32(), C.c_uint32(), C.c_double()
    lib().oo_analysis(_p(l), _p(r), l.size, C.byref(e), C.byref(be), C.byref(rate))
    return e.value, be.value, rate.value


def sig_to_fc32(x):
    x = _c128(x)
    o = np.zeros(2 * x.size, np.float32)
    lib().oo_sig_to_fc32(_p(x), x.size, _p(o))
    return o


def fc32_to_sig(f):
    f = np.ascontiguousarray(np.asarray(f, np.float32))
    o = np.zeros(f.size // 2, np.complex128)
    lib().oo_fc32_to_sig(_p(f), f.size // 2, _p(o))
    return o


# ---- FEC -----------------------------------------------------------------------------------
def hamming74_encode(data):
    d = _bytes(data)
    o = np.zeros(lib().oo_hamming74_encoded_len(d.size), np.uint8)
    lib().oo_hamming74_encode(_p(d), d.size, _p(o))
    return o


def hamming74_decode(coded):
    d = _bytes(coded)
    o = np.zeros(lib().oo_hamming74_decoded_len(d.size), np.uint8)
    lib().oo_hamming74_decode(_p(d), d.size, _p(o))
    return o


def rs_encode_block(msg, nsym=32):
    """parity bytes of one block (`Encoder::new(nsym).encode`, src/utils.rs:108,118)"""
    d = _bytes(msg)
    o = np.zeros(nsym, np.uint8)
    lib().oo_rs_encode_block(_p(d), d.size, nsym, _p(o))
    return o


def rs_correct_block(word, nsym=32):
    """(corrected word, n_corrected) or (word, -1): `Decoder::new(nsym).correct`, src/utils.rs:154,165"""
    w = _bytes(word).copy()
    r = lib().oo_rs_correct_block(_p(w), w.size, nsym)
    return w, int(r)


def rs_encode(data):
    """create_transmission_bytes, src/utils.rs:97-137"""
    d = _bytes(data)
    o = np.zeros(lib().oo_rs_encoded_len(d.size), np.uint8)
    lib().oo_rs_encode(_p(d), d.size, _p(o))
    return o


def rs_decode(coded):
    """decipher_transmission_bytes, src/utils.rs:152-180 -> (bytes, n_corrected, n_failed); the reference returns
    None when n_failed > 0"""
    d = _bytes(coded)
    o = np.zeros(lib().oo_rs_decoded_len(d.size), np.uint8)
    nc = np.zeros(1, np.uint32)
    nf = np.zeros(1, np.uint32)
    lib().oo_rs_decode(_p(d), d.size, _p(o), _p(nc), _p(nf))
    return o, int(nc[0]), int(nf[0])


# ---- TX / channel / RX ---------------------------------------------------------------------
def modulate(data, scheme):
    d = _bytes(data)
    o = np.zeros(8 * d.size + 8, np.complex128)
    n = lib().oo_modulate(_p(d), d.size, scheme, _p(o))
    return o[:n].copy()


def demodulate(syms, scheme):
    s = _c128(syms)
    o = np.zeros(s.size + 8, np.uint8)
    n = lib().oo_demodulate(_p(s), s.size, scheme, _p(o))
    if n == C.c_size_t(-1).value:
        raise ValueError("symbol count is not a multiple of 8")
    return o[:n].copy()


def encode(data, guard_bands=False, modulation=BPSK, nfft=64):
    d = _bytes(data)
    n = lib().oo_frame_len_n(d.size, int(guard_bands), modulation, nfft)
    o = np.zeros(n, np.complex128)
    m = lib().oo_encode_n(_p(d), d.size, int(guard_bands), modulation, nfft, _p(o))
    assert m == n
    return o


def tx(payload, cfg: OoCfg):
    d = _bytes(payload)
    n = lib().oo_tx_len(d.size, C.byref(cfg))
    o = np.zeros(n, np.complex128)
    m = lib().oo_tx(_p(d), d.size, C.byref(cfg), _p(o))
    assert m == n
    return o


def channel(txsig, snr_db=30.0, f_delta=-1.0, noise_mode=0, seed=1):
    t = _c128(txsig)
    o = np.zeros(t.size + 63, np.complex128)
    lib().oo_channel(_p(t), t.size, float(snr_db), float(f_delta), int(noise_mode), int(seed), _p(o))
    return o


@dataclass
class DecodeResult:
    status: int
    data: np.ndarray
    offset: int
    f_delta: float
    h_k: np.ndarray
    points: np.ndarray
    n_data_syms: int
    packet_length: int


def decode(samples, cfg: OoCfg, want_points=True, out_cap=None) -> DecodeResult:
    s = _c128(samples)
    nf = 1024 if cfg.nfft == 1024 else 64
    cap = out_cap if out_cap is not None else max(16, s.size)
    out = np.zeros(cap, np.uint8)
    pts_cap = (s.size // (nf + nf // 4) + 2) * nf if want_points else 0
    pts = np.zeros(max(pts_cap, 1), np.complex128)
    ol = C.c_size_t(0)
    diag = OoDiag()
    hfull = np.zeros(nf, np.complex128)
    diag.h_full = hfull.ctypes.data
    st = lib().oo_decode(_p(s), s.size, C.byref(cfg), _p(out), cap, C.byref(ol),
                         _p(pts) if want_points else None, pts_cap, C.byref(diag))
    npts = min(int(diag.n_points), pts_cap)
    return DecodeResult(st, out[: ol.value].copy(), diag.offset, diag.f_delta, hfull, pts[:npts].copy(),
                        int(diag.n_data_syms), int(diag.packet_length))


def decode_batch_fc32(iq: np.ndarray, n_samples: np.ndarray, cfg: OoCfg, out_stride: int, threads: int = 0):
    """iq: float32 [n_streams, iq_stride, 2] (fc32). Returns (out, out_len, status, offsets)."""
    iq = np.ascontiguousarray(iq, dtype=np.float32)
    n_streams, iq_stride = iq.shape[0], iq.shape[1]
    n_samples = np.ascontiguousarray(n_samples, dtype=np.uint32)
    out = np.zeros((n_streams, out_stride), np.uint8)
    out_len = np.zeros(n_streams, np.uint32)
    status = np.zeros(n_streams, np.int32)
    offsets = np.zeros(n_streams, np.int32)
    lib().oo_decode_batch_fc32(_p(iq), _p(n_samples), n_streams, iq_stride, C.byref(cfg), _p(out), out_stride,
                               _p(out_len), _p(status), _p(offsets), int(threads))
    return out, out_len, status, offsets


def max_threads() -> int:
    return int(lib().oo_max_threads())


PEAK_DTYPE = np.dtype([("offset", np.uint64), ("f_delta", np.float64), ("metric", np.float64)])


def sync_search(iq_c64: np.ndarray, max_peaks: int = 4096, nfft: int = 64) -> np.ndarray:
    """Capture search (docs/SPEC.md 4; nfft = 1024: section 9, every length scaled by 16). iq_c64: complex64 capture.
    Returns a structured array (offset, f_delta, metric)."""
    x = np.ascontiguousarray(iq_c64, dtype=np.complex64)
    peaks = np.zeros(max_peaks, PEAK_DTYPE)
    n = lib().oo_sync_search_fc32_n(_p(x), x.size, _p(peaks), max_peaks, int(nfft))
    return peaks[:n].copy()
