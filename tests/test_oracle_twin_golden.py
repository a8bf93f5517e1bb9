"""Oracle vs its independent numpy twin, oracle vs the committed golden vectors, and oracle-level properties."""
import numpy as np
import pytest


@pytest.mark.parametrize("mod", [0, 1, 2])
@pytest.mark.parametrize("guard", [False, True])
def test_twin_agrees_with_c_oracle(oo, mod, guard):
    from oracle import numpy_twin as tw
    lock, pre, tr = oo.locking_signal(), oo.preamble(), oo.training_signals()
    rng = np.random.default_rng(10 * mod + guard)
    pay = rng.integers(0, 256, 211, dtype=np.uint8).tobytes()
    a, b = oo.encode(pay, guard, mod), tw.encode(pay, guard, mod, lock, pre, tr)
    np.testing.assert_allclose(a, b, atol=1e-13)
    cap = oo.channel(a, 30.0, 0.02, 0, 9)
    r = oo.decode(cap, oo.make_cfg(guard, mod, xcorr_fft=True))
    t = tw.decode(cap, guard, mod, lock, tr)
    assert r.status == 0 and r.offset == t["offset"] == 8          # main tap at delay 9 -> lag 9 -> offset 8
    assert r.f_delta == pytest.approx(t["f_delta"], abs=1e-15)
    np.testing.assert_allclose(r.h_k, t["h_k"], atol=1e-12)
    n = t["points"].size
    np.testing.assert_allclose(r.points[:n], t["points"], atol=1e-11)
    assert r.data.tobytes() == t["data"] == pay


def test_direct_and_fft_sync_agree(oo):
    rng = np.random.default_rng(5)
    pay = rng.integers(0, 256, 50, dtype=np.uint8)
    cap = oo.channel(oo.encode(pay, True, oo.QPSK), 25.0, 0.03, 0, 3)
    a = oo.decode(cap, oo.make_cfg(True, oo.QPSK, xcorr_fft=True))
    b = oo.decode(cap, oo.make_cfg(True, oo.QPSK, xcorr_fft=False))
    assert a.offset == b.offset and a.data.tobytes() == b.data.tobytes() == pay.tobytes()


def test_golden_vectors_reproduce(oo, golden):
    for name in golden["names"]:
        mod, guard, fec, sync, cfo, phase, win = [int(v) for v in golden[f"{name}.cfg"]]
        cfg = oo.make_cfg(guard, mod, fec, sync, cfo, phase, win)
        tx = oo.tx(golden[f"{name}.payload"], cfg)
        np.testing.assert_allclose(tx, golden[f"{name}.tx"], atol=1e-15)
        r = oo.decode(golden[f"{name}.capture"].astype(np.complex128), cfg)
        assert r.status == 0 and r.offset == int(golden[f"{name}.offset"])
        assert r.f_delta == pytest.approx(float(golden[f"{name}.f_delta"]), abs=1e-15)
        np.testing.assert_allclose(r.h_k, golden[f"{name}.h_k"], atol=1e-13)
        np.testing.assert_allclose(r.points, golden[f"{name}.points"], atol=1e-6)
        assert r.data.tobytes() == golden[f"{name}.data"].tobytes() == golden[f"{name}.payload"].tobytes()
    np.testing.assert_array_equal(golden["tables.preamble"], oo.preamble())
    np.testing.assert_array_equal(golden["tables.training"], oo.training_signals())


def test_config1_frame_sizes(oo):
    # BPSK + 765 B -> 11 280 samples; QPSK raw 576 B -> 4 800; 64QAM+Hamming -> 3 120; 64QAM raw -> 2 160
    L = oo.lib()
    assert L.oo_frame_len(765, 1, oo.BPSK) == 11280
    assert L.oo_frame_len(576, 1, oo.QPSK) == 4800
    assert L.oo_frame_len(576, 1, oo.QAM64) == 2160
    import ctypes as C
    cfg = oo.make_cfg(True, oo.QAM64, True)
    assert L.oo_tx_len(576, C.byref(cfg)) == 3120


def test_status_paths(oo):
    rng = np.random.default_rng(8)
    pay = rng.integers(0, 256, 40, dtype=np.uint8)
    tx = oo.encode(pay, True, oo.BPSK)
    cfg = oo.make_cfg(True, oo.BPSK)
    # no channel delay: lag 0 -> offset -1 -> the reference panics (src/receiver.rs:25)
    assert oo.decode(tx, cfg).status == oo.NEG_OFFSET
    # one leading sample: offset 0, clean decode
    r = oo.decode(np.concatenate([[0], tx]), cfg)
    assert r.status == oo.OK and r.offset == 0 and r.data.tobytes() == pay.tobytes()
    # short input (src/receiver.rs:27-29)
    assert oo.decode(np.concatenate([np.zeros(5), tx[:700]]), cfg).status == oo.TOO_SHORT
    # capture cut inside the payload: header length does not fit
    assert oo.decode(np.concatenate([[0], tx[:1200]]), cfg).status == oo.BAD_HEADER
    # capture cut mid-symbol after the payload's last needed symbol: zero padded tail row still decodes
    full = np.concatenate([[0], tx, 0.001 * np.ones(37)])
    assert oo.decode(full, cfg).data.tobytes() == pay.tobytes()
    # Schmidl-Cox: noise only -> NO_SYNC
    noise = 0.01 * (rng.standard_normal(4000) + 1j * rng.standard_normal(4000))
    assert oo.decode(noise, oo.make_cfg(True, oo.BPSK, sync_mode=oo.SYNC_SCHMIDL_COX)).status == oo.NO_SYNC


@pytest.mark.parametrize("snr,cfo", [(15, 0.0), (20, 0.01), (25, 0.02), (30, 0.035), (40, 0.02)])
def test_loopback_sweep_bpsk_qpsk(oo, snr, cfo):
    # config 1 operating points: BPSK / QPSK decode error free over the lab channel
    rng = np.random.default_rng(int(snr * 100 + cfo * 1000))
    pay = rng.integers(0, 256, 576, dtype=np.uint8)
    for mod in (oo.BPSK, oo.QPSK):
        cap = oo.channel(oo.encode(pay, True, mod), snr, cfo, 0, 0xD0FD0001)
        r = oo.decode(cap, oo.make_cfg(True, mod))
        assert r.status == 0 and r.offset == 8
        errs = oo.analysis(pay, r.data)[0]
        assert errs == 0 or snr < 25


def test_schmidl_cox_matches_reference_sync(oo):
    rng = np.random.default_rng(12)
    pay = rng.integers(0, 256, 300, dtype=np.uint8)
    for lead in (0, 1, 57, 400, 1031):
        tx = oo.encode(pay, True, oo.QAM64)
        cap = oo.channel(tx, 35.0, 0.02, 1, lead + 1)
        noise = 0.005 * (rng.standard_normal(lead) + 1j * rng.standard_normal(lead))
        cap = np.concatenate([noise, cap])
        a = oo.decode(cap, oo.make_cfg(True, oo.QAM64, sync_mode=oo.SYNC_REFERENCE))
        b = oo.decode(cap, oo.make_cfg(True, oo.QAM64, sync_mode=oo.SYNC_SCHMIDL_COX, sync_window=2048))
        assert a.status == b.status == 0
        assert a.offset == b.offset == lead + 8


@pytest.mark.parametrize("mod", [0, 1, 2])
def test_wideband_1024_oracle_loopback(oo, mod):
    """docs/SPEC.md 9: nfft = 1024, CP = 256. TX -> lab channel -> RX recovers the payload; frame sizes scale by 16."""
    import ctypes as C
    rng = np.random.default_rng(mod)
    pay = rng.integers(0, 256, 2000, dtype=np.uint8)
    for guard in (False, True):
        cfg = oo.make_cfg(guard, mod, True, oo.SYNC_SCHMIDL_COX, 1, 1, 4096, nfft=1024)
        tx = oo.tx(pay, cfg)
        D, bpc = (768 if guard else 1024), (1, 2, 6)[mod]
        coded = (14 * 2000 + 7) // 8
        S = -(-(-(-(128 + 8 * coded) // bpc)) // D)
        assert tx.size == (10 + S) * 1280 == oo.lib().oo_tx_len(2000, C.byref(cfg))
        assert max(tx.real.max(), tx.imag.max()) == pytest.approx(1.0)
        cap = np.concatenate([np.zeros(333), oo.channel(tx, 60.0, 0.0017, 1, 4)])
        r = oo.decode(cap, cfg)
        assert r.status == 0 and r.offset == 333 + 8 and abs(r.f_delta - 0.0017) < 1e-6
        assert r.data.tobytes() == pay.tobytes()
        assert r.h_k.size == 1024
    # carrier map: 192 nulls, 64 pilots, 768 data
    x = oo.encode(bytes(5000), True, oo.BPSK, nfft=1024)
    sym = np.fft.fft(x[12800 + 256: 12800 + 1280])
    mag = np.abs(sym)
    nulls = [k for k in range(1024) if k <= 95 or k == 512 or k >= 929]
    assert len(nulls) == 192 and mag[nulls].max() < 1e-9 * mag.max()
    used = [k for k in range(1024) if k not in set(nulls)]
    pilots = used[::13]
    assert len(used) == 832 and len(pilots) == 64 and pilots[0] == 96
    np.testing.assert_allclose(sym[pilots] / sym[pilots][0], 1.0, atol=1e-9)          # pilots are 1+0j, data is -1 (all-zero bits)
    data = [k for k in used if k not in set(pilots)]
    assert len(data) == 768
    np.testing.assert_allclose(sym[data][16 * 8:] / sym[pilots][0], -1.0, atol=1e-9)


def test_oracle_capture_search_scales_with_the_symbol_length(oo):
    """docs/SPEC.md 4 / 9: the capture search of the nfft = 1024 layout is the nfft = 64 one with every length scaled by 16.
    Planted frames are found at `start - 1` (lag - 1 rule) with their CFO, in both layouts; the hold-off keeps one detection
    per frame; a frame head cut by the end of the capture is not reported."""
    rng = np.random.default_rng(12)
    for nfft, gap in ((64, 900), (1024, 14_000)):
        cfg = oo.make_cfg(True, oo.QAM64, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, 0, nfft=nfft)
        L = nfft + nfft // 4
        parts, want, pos = [], [], 0
        for k in range(6):
            g = int(rng.integers(gap, 3 * gap))
            parts.append(1e-3 * (rng.standard_normal(g) + 1j * rng.standard_normal(g)))
            pos += g
            tx = oo.tx(rng.integers(0, 256, int(rng.integers(1, 800)), dtype=np.uint8).tobytes(), cfg)
            f = float(rng.uniform(-0.5, 0.5) * np.pi / L)
            parts.append(tx * np.exp(1j * f * np.arange(tx.size)))
            want.append((pos - 1, f))
            pos += tx.size
        cap = np.concatenate(parts).astype(np.complex64)
        pk = oo.sync_search(cap, nfft=nfft)
        assert [int(x) for x in pk["offset"]] == [p for p, _ in want]
        np.testing.assert_allclose(pk["f_delta"], [f for _, f in want], atol=2e-4 / (L / 80))
        assert (pk["metric"] > 0.5).all()
        cut = want[-1][0] + 10 * L                                         # the last frame's head just fits: offset + 10 L <= M
        assert len(oo.sync_search(cap[:cut - 1], nfft=nfft)) == len(want) - 1
        assert len(oo.sync_search(cap[:cut], nfft=nfft)) == len(want)
