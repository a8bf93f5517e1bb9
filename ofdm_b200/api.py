"""Host-side mirror of the reference crate's public modem API, backed by the CUDA engine.

Same names, argument meaning and error behaviour as the reference so tests read like its own:

  encode(data, guard_bands=None, modulation=None) -> complex128 array    src/transmitter.rs:10-58
  decode(samples, guard_bands=None, modulation=None) -> bytes            src/receiver.rs:8-96
  ModulationScheme {Bpsk, Qpsk, Qam}                                      src/transmitter.rs:98-104
  Header(packet_length)  (bincode u128 LE)                                src/packets/mod.rs:20-32
  Analysis.new(left, right) -> num_errs, num_block_errs, err_rate        src/utils.rs:38-69
  sig_to_bytes / bytes_to_sig (fc32 wire format)                          src/utils.rs:228-254

plus the batched forms (`encode_batch`, `decode_batch`) a GPU needs. Nothing here falls back to the CPU: every
call goes through libofdm_b200.so and raises if it (or a CUDA device) is missing.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import engine as _e


class ModulationScheme(enum.IntEnum):
    Bpsk = _e.MOD_BPSK
    Qpsk = _e.MOD_QPSK
    Qam = _e.MOD_QAM64      # empty arm in the reference; 64QAM here (docs/SPEC.md)


class DecodeError(RuntimeError):
    """The reference's anyhow::Error / panics of `decode`, one status per stream."""

    def __init__(self, status: int):
        self.status = int(status)
        msg = {
            _e.TOO_SHORT: "Input not long enough, bailing early",      # src/receiver.rs:28
            _e.NO_SYNC: "no frame found",
            _e.BAD_HEADER: "header length does not fit the received frame",
            _e.NEG_OFFSET: "negative frame offset",                    # src/receiver.rs:25 panics
        }.get(self.status, "decode failed")
        super().__init__(msg)


@dataclass(frozen=True)
class Header:
    packet_length: int

    def serialize(self) -> bytes:            # bincode fixint LE u128
        return int(self.packet_length).to_bytes(16, "little")

    @staticmethod
    def deserialize(b: bytes) -> "Header":
        return Header(int.from_bytes(bytes(b[:16]), "little"))


@dataclass(frozen=True)
class Analysis:
    num_errs: int
    num_block_errs: int
    err_rate: float

    @staticmethod
    def new(left: bytes, right: bytes, device: int = 0) -> "Analysis":
        left, right = bytes(left), bytes(right)
        assert len(left) == len(right)          # src/utils.rs:46
        l = np.frombuffer(left, np.uint8)[None, :]
        r = np.frombuffer(right, np.uint8)[None, :]
        n = np.array([len(left)], np.uint32)
        c = _engine(_e.Config(), device).ber(l, n, r, n, np.zeros(1, np.int32))
        return Analysis(int(c[0]), int(c[1]), float(c[0]) / (len(left) * 8.0))


_engines: dict = {}


def _engine(cfg: _e.Config, device: int = 0) -> _e.Engine:
    key = (cfg, device)
    if key not in _engines:
        _engines[key] = _e.Engine(cfg, device)
    return _engines[key]


def _cfg(guard_bands, modulation, **kw) -> _e.Config:
    return _e.Config(modulation=int(modulation if modulation is not None else ModulationScheme.Bpsk),
                     guard_bands=bool(guard_bands) if guard_bands is not None else False, **kw)


def encode_batch(payloads: Sequence[bytes], guard_bands: Optional[bool] = None,
                 modulation: Optional[ModulationScheme] = None, device: int = 0, **kw):
    """Returns a list of complex128 frames (one per payload)."""
    eng = _engine(_cfg(guard_bands, modulation, **kw), device)
    iq, flen = eng.tx_encode(payloads)
    return [iq[i, : flen[i]].astype(np.complex128) for i in range(len(payloads))]


def encode(data: bytes, guard_bands: Optional[bool] = None, modulation: Optional[ModulationScheme] = None,
           device: int = 0, **kw) -> np.ndarray:
    return encode_batch([bytes(data)], guard_bands, modulation, device, **kw)[0]


def decode_batch(captures: Sequence[np.ndarray], guard_bands: Optional[bool] = None,
                 modulation: Optional[ModulationScheme] = None, device: int = 0, **kw):
    """Returns (list of bytes, status array); a failed stream yields b'' and its status."""
    eng = _engine(_cfg(guard_bands, modulation, **kw), device)
    n = len(captures)
    lens = np.array([len(c) for c in captures], np.uint32)
    stride = int(lens.max())
    iq = np.zeros((n, stride), np.complex64)
    for i, c in enumerate(captures):
        iq[i, : len(c)] = np.asarray(c).astype(np.complex64)     # Complex64 -> fc32 like sig_to_bytes
    res = eng.rx_decode(iq, lens)
    return res.data, res.status


def decode(samples, guard_bands: Optional[bool] = None, modulation: Optional[ModulationScheme] = None,
           device: int = 0, **kw) -> bytes:
    data, status = decode_batch([np.asarray(samples)], guard_bands, modulation, device, **kw)
    if status[0] != _e.OK:
        raise DecodeError(int(status[0]))
    return data[0]


def create_transmission_bytes(data: bytes, device: int = 0) -> bytes:
    """RS(255,223) outer code with the reference's framing (src/utils.rs:97-137)."""
    coded, n = _engine(_e.Config(), device).rs_encode([bytes(data)])
    return coded[0, :int(n[0])].tobytes()


def decipher_transmission_bytes(data: bytes, device: int = 0) -> Optional[bytes]:
    """Inverse of create_transmission_bytes (src/utils.rs:152-180): None when any block is beyond repair."""
    out, n, _, n_failed = _engine(_e.Config(), device).rs_decode(np.frombuffer(bytes(data), np.uint8))
    return None if int(n_failed[0]) else out[0, :int(n[0])].tobytes()


def sig_to_bytes(sig) -> bytes:
    """src/utils.rs:228-236: interleaved native-endian f32 re, im."""
    return np.asarray(sig).astype(np.complex64).tobytes()


def bytes_to_sig(b: bytes) -> np.ndarray:
    """src/utils.rs:239-254 (drops a trailing partial sample like chunks_exact)."""
    n = len(b) // 8
    return np.frombuffer(bytes(b[: 8 * n]), np.complex64).astype(np.complex128)
