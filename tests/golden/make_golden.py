"""Generates tests/golden/golden_v1.npz from the CPU oracle (run in the build container; committed with its output).

The reference crate cannot be built or run here (no rustc/cargo), and it ships no golden IQ, so these vectors pin the
ORACLE's behaviour (regression pin for oracle + engine), not the reference's. The config-1 payload
(support/dancing.bytes, 576 palette ids) is read from /root/reference when present and embedded in the fixture.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as oo  # noqa: E402

CASES = [
    # name, modulation, guard, fec, sync, cfo, phase, window, snr_db, cfo_rad, noise_mode, n_payload
    ("bpsk_guard_ref", 0, 1, 0, 0, 0, 0, 0, 30.0, 0.02, 0, 100),
    ("qpsk_noguard_ref", 1, 0, 0, 0, 0, 0, 0, 30.0, 0.01, 0, 64),
    ("qpsk_guard_ref", 1, 1, 0, 0, 0, 0, 0, 30.0, 0.035, 0, 8),       # "alskdjas"-sized
    ("qam64_guard_fec_sc", 2, 1, 1, 1, 1, 1, 1024, 35.0, 0.02, 1, 576),  # config 1 payload size
    ("qam64_noguard_ref", 2, 0, 0, 0, 0, 0, 0, 40.0, 0.0, 1, 99),
    ("bpsk_guard_fec_sc", 0, 1, 1, 1, 1, 0, 512, 30.0, 0.03, 0, 31),
]


def main():
    rng = np.random.default_rng(0x0FD0)
    out = {}
    dancing = None
    p = "/root/reference/support/dancing.bytes"
    if os.path.exists(p):
        dancing = np.fromfile(p, np.uint8)
    for i, (name, mod, guard, fec, sync, cfo, phase, win, snr, f, nm, n) in enumerate(CASES):
        cfg = oo.make_cfg(guard, mod, fec, sync, cfo, phase, win)
        if n == 576 and dancing is not None:
            pay = dancing.copy()
        elif n == 8:
            pay = np.frombuffer(b"alskdjas", np.uint8).copy()
        else:
            pay = rng.integers(0, 256, n, dtype=np.uint8)
        tx = oo.tx(pay, cfg)
        cap = oo.channel(tx, snr, f, nm, 0xD0FD0001 + i).astype(np.complex64)      # fc32 wire format
        res = oo.decode(cap.astype(np.complex128), cfg)
        assert res.status == 0 and res.data.tobytes() == pay.tobytes(), name
        out[f"{name}.cfg"] = np.array([mod, guard, fec, sync, cfo, phase, win], np.int32)
        out[f"{name}.payload"] = pay
        out[f"{name}.tx"] = tx
        out[f"{name}.capture"] = cap
        out[f"{name}.offset"] = np.int32(res.offset)
        out[f"{name}.f_delta"] = np.float64(res.f_delta)
        out[f"{name}.h_k"] = res.h_k
        out[f"{name}.points"] = res.points.astype(np.complex64)
        out[f"{name}.data"] = res.data
    out["tables.lock"] = oo.locking_signal()
    out["tables.preamble"] = oo.preamble()
    out["tables.training"] = oo.training_signals()
    out["names"] = np.array([c[0] for c in CASES])
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote golden_v1.npz", {k: v.shape for k, v in out.items() if k.endswith("capture")})


if __name__ == "__main__":
    main()
