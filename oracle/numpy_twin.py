"""numpy twin of the C oracle (TEST INFRASTRUCTURE): an independent restatement of the reference's encode/decode
(src/transmitter.rs:11-58, src/receiver.rs:9-96) written against numpy.fft, used only to cross-check
oracle/ofdm_oracle.c. Reference modes only (ramp-correlation sync, mean-of-angles CFO and pilot phase), BPSK/QPSK
as in the reference plus the SPEC's 64QAM. The three tables are inputs (taken from the C oracle) because the
StdRng restatement lives there.
"""
from __future__ import annotations

import numpy as np

NULLS = [i for i in range(64) if i >= 59 or i <= 5 or i == 32]       # src/transmitter.rs:150
PILOTS = [6, 25, 39, 58]                                             # src/transmitter.rs:153
DATA_G = [i for i in range(64) if i not in NULLS and i not in PILOTS]
QAM_LEVEL = {0: -7, 1: -5, 3: -3, 2: -1, 6: 1, 7: 3, 5: 5, 4: 7}     # docs/SPEC.md 2


def bits_lsb_first(data: bytes) -> np.ndarray:
    return np.unpackbits(np.frombuffer(bytes(data), np.uint8), bitorder="little")      # src/utils.rs:21-27


def modulate(data: bytes, scheme: int) -> np.ndarray:
    b = bits_lsb_first(data).astype(int)
    if scheme == 0:
        return (2.0 * b - 1.0).astype(np.complex128)                                  # src/transmitter.rs:112-118
    if scheme == 1:
        b = b.reshape(-1, 2)
        return (2.0 * b[:, 0] - 1.0) + 1j * (2.0 * b[:, 1] - 1.0)                     # src/transmitter.rs:122-132
    pad = (-b.size) % 6
    b = np.concatenate([b, np.zeros(pad, int)]).reshape(-1, 6)
    ci = b[:, 0] + 2 * b[:, 1] + 4 * b[:, 2]
    cq = b[:, 3] + 2 * b[:, 4] + 4 * b[:, 5]
    lv = np.vectorize(QAM_LEVEL.get)
    return lv(ci) / 7.0 + 1j * lv(cq) / 7.0


def encode(data: bytes, guard_bands: bool, scheme: int, lock, pre, train) -> np.ndarray:
    out = [np.asarray(lock)] + [np.asarray(pre)] * 4                                  # src/transmitter.rs:22-29
    t = np.fft.ifft(np.asarray(train))
    out += [np.concatenate([t[48:], t])] * 5                                          # :32-34, :168-181
    syms = modulate(len(data).to_bytes(16, "little") + bytes(data), scheme)          # :37-47
    D = 48 if guard_bands else 64
    carriers = DATA_G if guard_bands else list(range(64))
    for k in range(0, syms.size, D):                                                  # :49-54
        blk = np.zeros(64, np.complex128)
        chunk = syms[k:k + D]
        blk[carriers[:chunk.size]] = chunk
        if guard_bands:
            blk[PILOTS] = 1.0
        t = np.fft.ifft(blk)
        out.append(np.concatenate([t[48:], t]))
    x = np.concatenate(out)
    m = max(0.0, x.real.max(), x.imag.max())                                          # :183-194
    return x / m


def demodulate(points: np.ndarray, scheme: int) -> bytes:
    re, im = points.real, points.imag
    if scheme == 0:
        bits = (re > 0.0).astype(np.uint8)                                            # src/receiver.rs:162
    elif scheme == 1:
        l = re >= 0.0
        r = np.where(l, im >= 0.0, (re < 0.0) & (im > 0.0))                           # src/receiver.rs:169-175
        bits = np.stack([l, r], 1).astype(np.uint8).reshape(-1)
    else:
        def axis(v):
            i = np.clip(np.floor(3.5 * v + 4.0), 0, 7).astype(int)
            return i ^ (i >> 1)
        ci, cq = axis(re), axis(im)
        bits = np.stack([(ci >> 0) & 1, (ci >> 1) & 1, (ci >> 2) & 1, (cq >> 0) & 1, (cq >> 1) & 1, (cq >> 2) & 1], 1)
        bits = bits.astype(np.uint8).reshape(-1)
    return np.packbits(bits[: bits.size // 8 * 8], bitorder="little").tobytes()


def decode(samples: np.ndarray, guard_bands: bool, scheme: int, lock, train):
    """Returns dict(offset, f_delta, h_k, points, data) following src/receiver.rs:9-96 step by step."""
    a = np.asarray(samples, np.complex128)
    M = a.size
    P = 2 * M - 1
    # xcorr_fft, src/signals/mod.rs:186-217: literal length-P transforms
    fa = np.fft.fft(np.concatenate([a, np.zeros(P - M)]))
    fb = np.fft.fft(np.concatenate([np.asarray(lock, np.complex128), np.zeros(P - len(lock))]))
    c = np.fft.ifft(fa * np.conj(fb))
    mid = (P + 1) // 2
    c = np.concatenate([c[mid:], c[:mid]])                                            # fft_shift :61-77
    idx = int(np.argmax(np.abs(c) ** 2))              # first strict maximum, src/signals/mod.rs:205-214
    offset = idx - ((P - 1) // 2 + 1)                                                 # src/receiver.rs:21
    x = a[offset:]
    rows = -(-x.size // 80)
    x = np.concatenate([x, np.zeros(rows * 80 - x.size)]).reshape(rows, 80)           # :192-210
    f_delta = abs(np.mean(np.angle(x[4] / x[3])) / 80.0)                              # :231-240
    n = np.arange(rows * 80).reshape(rows, 80)
    x = x * np.exp(-1j * f_delta * n)                                                 # :44-50
    H = np.mean(np.fft.fft(x[5:10, 16:], axis=1) / np.asarray(train)[None, :], axis=0)   # :212-229
    Y = np.fft.fft(x[10:, 16:], axis=1) / H[None, :]                                  # :64-70
    if guard_bands:
        phi = np.mean(np.angle(Y[:, PILOTS]), axis=1)                                 # :125-137
        pts = Y[:, DATA_G] * np.exp(-1j * phi)[:, None]                               # :140-144
    else:
        pts = Y
    pts = pts.reshape(-1)
    raw = demodulate(pts, scheme)
    plen = int.from_bytes(raw[:16], "little")                                         # :86-93
    return {"offset": offset, "f_delta": f_delta, "h_k": H, "points": pts, "data": raw[16:16 + plen]}
