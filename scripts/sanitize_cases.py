#!/usr/bin/env python
"""Small invocations of every kernel family, meant to run under compute-sanitizer (scripts/sanitize.sh): N=64 TX / channel /
RX (several modes, one frame spanning four 224-symbol tiles so the TMA prefetch crosses tile boundaries), the wideband
variant, capture search + streaming decode, RS(255,223) with errors, BER. Every result is checked against what was sent."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ofdm_b200 as ob

rng = np.random.default_rng(2026)


def loop(cfg, lens, snr=45.0, cfo=0.02, lead=(8, 300), out_stride=None):
    eng = ob.Engine(cfg, 0)
    pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    iq, flen = eng.tx_encode(pays)
    rx, rl, _, _ = eng.channel(iq, flen, ob.ChannelParams(snr_db=snr, cfo_max=cfo, lead_min=lead[0], lead_max=lead[1], noise_mode=1, seed=7))
    res = eng.rx_decode(rx, rl, out_stride=out_stride or (max(lens) + 16), points=True)
    ok = all(res.status[i] == 0 and res.data[i] == p for i, p in enumerate(pays))
    ref = np.zeros((len(pays), max(lens) + 16), np.uint8)
    got = np.zeros_like(ref)
    for i, p in enumerate(pays):
        ref[i, : len(p)] = np.frombuffer(p, np.uint8)
        got[i, : len(res.data[i])] = np.frombuffer(res.data[i], np.uint8)
    c = eng.ber(ref, np.array(lens, np.uint32), got, np.array([len(d) for d in res.data], np.uint32), np.asarray(res.status, np.int32))
    eng.close()
    return ok and c[0] == 0 and c[3] == 0


results = {}
sc = dict(sync_mode=ob.SYNC_SCHMIDL_COX, cfo_mode=ob.CFO_ANGLE_OF_SUM, phase_mode=ob.PHASE_ANGLE_OF_SUM, sync_window=1024)
c64 = ob.Config(modulation=ob.MOD_QAM64, guard_bands=True, fec=True, **sc)
results["n64_qam64_fec_4tiles"] = loop(c64, [c64.max_payload(700), 576, 0, 1, 3000])
results["n64_qam64_nofec_noguard"] = loop(ob.Config(modulation=ob.MOD_QAM64, guard_bands=False, fec=False, **sc), [4000, 5], snr=60.0, cfo=0.002)
results["n64_qpsk_fec"] = loop(ob.Config(modulation=ob.MOD_QPSK, guard_bands=True, fec=True, **sc), [2000, 77])
results["n64_bpsk_reference_modes"] = loop(ob.Config(modulation=ob.MOD_BPSK, guard_bands=True, fec=False), [576, 33], cfo=0.01, lead=(0, 40))
cw = ob.Config(modulation=ob.MOD_QAM64, guard_bands=True, fec=True, sync_mode=1, cfo_mode=1, phase_mode=1, sync_window=4096, nfft=1024, cp=256)
results["wide_qam64_fec"] = loop(cw, [cw.max_payload(30), 577], snr=60.0, cfo=0.001, lead=(0, 500))

# capture search + streaming receiver
eng = ob.Engine(c64, 0)
pays = [rng.integers(0, 256, int(rng.integers(1, 5000)), dtype=np.uint8).tobytes() for _ in range(9)]
iq, flen = eng.tx_encode(pays)
cap = (0.0003 * (rng.standard_normal(500_000) + 1j * rng.standard_normal(500_000))).astype(np.complex64)
pos, p = [], 1234
for i in range(len(pays)):
    cap[p: p + flen[i]] += iq[i, : flen[i]]
    pos.append(p)
    p += int(flen[i]) + int(rng.integers(900, 20_000))
peaks, data, status = eng.decode_capture(cap, out_stride=5008)
results["capture_search_and_decode"] = (len(peaks) == len(pays) and [int(x) for x in peaks["offset"]] == [q - 1 for q in pos]
                                         and all(d == q for d, q in zip(data, pays)) and bool((status == 0).all()))
# Reed-Solomon with errors (3 or 16 symbol errors per block: the Berlekamp-Massey / Chien / Forney path)
msg = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (576, 0, 223, 5000)]
coded, clen = eng.rs_encode(msg)
bad = coded.copy()
for i in range(len(msg)):
    for b in range(int(clen[i]) // 255):
        idx = rng.choice(255, 16 if b % 2 else 3, replace=False)
        bad[i, 255 * b + idx] ^= rng.integers(1, 256, idx.size, dtype=np.uint8)
dec, dlen, ncorr, nfail = eng.rs_decode(bad, clen)
results["rs_255_223_errors"] = (all(bytes(dec[i, : len(m)]) == m for i, m in enumerate(msg)) and int(np.sum(nfail)) == 0
                                and int(np.sum(ncorr)) > 0)
eng.close()
print(results)
sys.exit(0 if all(results.values()) else 1)
