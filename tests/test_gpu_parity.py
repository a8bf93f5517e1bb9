"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle on identical inputs.

Bars (north star): Hamming / QAM bit mapping / frame + packet indexing bit-exact given the same hard decisions;
equalised constellation points within 1e-4 relative (fp32 engine vs f64 oracle); offset identical; f_delta within 1e-6.
"""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL_POINTS = 1e-4        # north star tolerance for equalised constellation points
ABS_TOL_FDELTA = 1e-6        # rad/sample
MODES = [(0, 0, 0), (1, 1, 1), (1, 0, 0), (0, 1, 1)]      # (sync, cfo, phase)


@pytest.fixture(scope="module")
def ob():
    import ofdm_b200
    return ofdm_b200


def _mk(ob, oo, mod, guard, fec, sync, cfo, phase, window=None):
    window = (0 if sync == 0 else 1024) if window is None else window
    cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec, sync_mode=sync, cfo_mode=cfo, phase_mode=phase, sync_window=window)
    return ob.Engine(cfg, 0), cfg, oo.make_cfg(guard, mod, fec, sync, cfo, phase, window)


def _batch(caps):
    n = np.array([c.size for c in caps], np.uint32)
    iq = np.zeros((len(caps), int(n.max())), np.complex64)
    for i, c in enumerate(caps):
        iq[i, : c.size] = c
    return iq, n


def _near_boundary(points, mod, tol):
    """True where a hard decision of the point could flip under a perturbation of `tol` (relative to scale 1)."""
    re, im = points.real, points.imag
    if mod == 0:
        return np.abs(re) < tol
    if mod == 1:
        return (np.abs(re) < tol) | (np.abs(im) < tol)
    fr = lambda v: np.abs((3.5 * v + 4.0) - np.round(3.5 * v + 4.0)) < 3.5 * tol
    return fr(re) | fr(im)


@pytest.mark.parametrize("mod,guard,fec", list(itertools.product((0, 1, 2), (False, True), (False, True))))
@pytest.mark.parametrize("modes", MODES)
def test_tx_rx_parity_matrix(ob, oo, mod, guard, fec, modes):
    sync, cfo, phase = modes
    eng, cfg, ocfg = _mk(ob, oo, mod, guard, fec, sync, cfo, phase)
    rng = np.random.default_rng(1000 * mod + 100 * guard + 10 * fec + sync)
    pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (576, 0, 1, 100, 333)]   # ragged + empty
    iq, flen = eng.tx_encode(pays)
    caps = []
    for i, p in enumerate(pays):
        ref = oo.tx(p, ocfg)
        assert ref.size == flen[i]
        np.testing.assert_allclose(iq[i, : flen[i]], ref, atol=2e-6)           # TX parity (fp32 IFFT vs f64)
        assert not iq[i, flen[i]:].any()                                        # zero fill past the frame
        caps.append(oo.channel(ref, 40.0, 0.02 + 0.003 * i, 1, 100 + i))
    batch, n = _batch(caps)
    res = eng.rx_decode(batch, n, points=True)
    for i, p in enumerate(pays):
        ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg)
        assert ref.status == res.status[i] == 0
        assert ref.offset == res.offset[i] == 8
        assert abs(ref.f_delta - res.f_delta[i]) < ABS_TOL_FDELTA
        np.testing.assert_allclose(res.h_k[i], ref.h_k, atol=REL_TOL_POINTS * np.abs(ref.h_k).max())
        npts = cfg.frame_data_syms(len(p)) * cfg.data_carriers
        scale = max(1.0, np.abs(ref.points[:npts]).max())
        assert np.abs(res.points[i, :npts] - ref.points[:npts]).max() <= REL_TOL_POINTS * scale
        assert res.data[i] == ref.data.tobytes() == p                           # bit exact, ragged lengths
        assert res.out_len[i] == len(p)
    eng.close()


def test_low_snr_decisions_identical_away_from_boundaries(ob, oo):
    """At 27 dB (many symbol errors) the engine's hard decisions equal the oracle's wherever the oracle's point is
    not within 1e-4 of a decision boundary; with identical decisions the bytes are identical."""
    eng, cfg, ocfg = _mk(ob, oo, 2, True, False, 0, 0, 0)
    rng = np.random.default_rng(4)
    pays = [rng.integers(0, 256, 2000, dtype=np.uint8).tobytes() for _ in range(24)]
    caps = [oo.channel(oo.tx(p, ocfg), 27.0, 0.02, 1, 7 + i) for i, p in enumerate(pays)]
    batch, n = _batch(caps)
    res = eng.rx_decode(batch, n, points=True)
    n_err = 0
    for i, p in enumerate(pays):
        ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg)
        if ref.status != 0:
            assert res.status[i] == ref.status
            continue
        assert res.status[i] == 0 and res.out_len[i] == ref.data.size
        npts = cfg.frame_data_syms(len(p)) * 48
        gb = np.unpackbits(np.frombuffer(res.data[i], np.uint8), bitorder="little")
        rb = np.unpackbits(ref.data, bitorder="little")
        if ref.data.size == len(p):                     # (a corrupted header can still give a valid, shorter length)
            n_err += int((rb != np.unpackbits(np.frombuffer(p, np.uint8), bitorder="little")).sum())
        diff = np.flatnonzero(gb != rb)
        risky = _near_boundary(ref.points[:npts], 2, 1e-4)
        for b in diff:                                  # payload bit b -> stream bit 128 + b -> carrier
            assert risky[(128 + b) // 6], f"stream {i}: bit {b} differs away from any decision boundary"
    assert n_err > 0                                    # the channel really produced bit errors


def test_long_frame_hot_kernel_tiling(ob, oo):
    """Frames spanning many 224-symbol tiles, Hamming tile alignment, every modulation."""
    for mod, guard, fec in ((2, True, True), (2, True, False), (1, True, True), (0, False, True), (2, False, True)):
        eng, cfg, ocfg = _mk(ob, oo, mod, guard, fec, 1, 1, 1, 2048)
        rng = np.random.default_rng(mod)
        S = 700 if mod == 2 else 500
        lens = [cfg.max_payload(S), cfg.max_payload(S) - 1, cfg.max_payload(225), cfg.max_payload(224), cfg.max_payload(217) + 1, 5]
        pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
        caps = []
        for i, p in enumerate(pays):
            lead = int(rng.integers(0, 900))
            c = oo.channel(oo.tx(p, ocfg), 60.0, 0.025, 1, i)       # ~noise free: without pilots the phase walks with the CFO estimate error
            caps.append(np.concatenate([0.002 * (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)), c]))
        batch, n = _batch(caps)
        res = eng.rx_decode(batch, n, diag=True)
        for i, p in enumerate(pays):
            ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg, want_points=False)
            assert ref.status == res.status[i] == 0 and ref.offset == res.offset[i]
            assert res.data[i] == ref.data.tobytes() == p, (mod, guard, fec, i)
        eng.close()


def test_status_paths_and_mixed_batch(ob, oo):
    """TOO_SHORT / NEG_OFFSET / BAD_HEADER / NO_SYNC never abort the batch; good streams still decode."""
    rng = np.random.default_rng(8)
    pay = rng.integers(0, 256, 40, dtype=np.uint8).tobytes()
    for sync in (0, 1):
        eng, cfg, ocfg = _mk(ob, oo, 0, True, False, sync, 0, 0)
        tx = oo.tx(pay, ocfg)
        noise = 0.01 * (rng.standard_normal(4000) + 1j * rng.standard_normal(4000))
        caps = [
            tx,                                                   # no delay: lag 0 -> offset -1 (reference panics)
            np.concatenate([[0], tx]),                            # offset 0
            np.concatenate([np.zeros(5), tx[:700]]),              # too short
            np.concatenate([[0], tx[:1200]]),                     # header length does not fit
            np.concatenate([[0], tx, 0.001 * np.ones(37)]),       # partial (zero padded) tail row
            noise,                                                # nothing there
            oo.channel(tx, 30.0, 0.01, 0, 3),
        ]
        batch, n = _batch(caps)
        res = eng.rx_decode(batch, n, diag=True)
        for i in range(len(caps)):
            ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg, want_points=False)
            assert res.status[i] == ref.status, (sync, i, res.status[i], ref.status)
            if ref.status == 0:
                assert res.offset[i] == ref.offset and res.data[i] == ref.data.tobytes() == pay
            else:
                assert res.out_len[i] == 0 and res.data[i] == b""
        if sync == 0:
            assert list(res.status[:6]) == [ob.NEG_OFFSET, ob.OK, ob.TOO_SHORT, ob.BAD_HEADER, ob.OK, res.status[5]]
        else:
            assert res.status[5] == ob.NO_SYNC
        eng.close()


def test_out_stride_too_small_is_bad_header(ob, oo):
    eng, cfg, ocfg = _mk(ob, oo, 1, True, False, 0, 0, 0)
    cap = oo.channel(oo.tx(bytes(100), ocfg), 30.0, 0.0, 0, 1)
    batch, n = _batch([cap])
    assert eng.rx_decode(batch, n, out_stride=64).status[0] == ob.BAD_HEADER
    assert eng.rx_decode(batch, n, out_stride=100).status[0] == ob.OK
    eng.close()


def test_golden_vectors(ob, oo, golden):
    for name in golden["names"]:
        mod, guard, fec, sync, cfo, phase, win = [int(v) for v in golden[f"{name}.cfg"]]
        eng, cfg, _ = _mk(ob, oo, mod, bool(guard), bool(fec), sync, cfo, phase, win)
        pay = golden[f"{name}.payload"].tobytes()
        iq, flen = eng.tx_encode([pay])
        np.testing.assert_allclose(iq[0, : flen[0]], golden[f"{name}.tx"], atol=2e-6)
        cap = golden[f"{name}.capture"]
        res = eng.rx_decode(cap[None, :], points=True)
        assert res.status[0] == 0 and res.offset[0] == int(golden[f"{name}.offset"])
        assert abs(res.f_delta[0] - float(golden[f"{name}.f_delta"])) < ABS_TOL_FDELTA
        pts = golden[f"{name}.points"]
        npts = cfg.frame_data_syms(len(pay)) * cfg.data_carriers
        assert np.abs(res.points[0, :npts] - pts[:npts]).max() <= REL_TOL_POINTS * max(1.0, np.abs(pts[:npts]).max())
        assert res.data[0] == golden[f"{name}.data"].tobytes() == pay
        lock, pre, tr = eng.tables()
        np.testing.assert_allclose(lock, golden["tables.lock"], atol=1e-7)
        np.testing.assert_allclose(pre, golden["tables.preamble"], atol=1e-7)
        np.testing.assert_allclose(tr, golden["tables.training"], atol=1e-7)
        eng.close()


def test_ber_counters_match_analysis(ob, oo):
    eng = ob.Engine(ob.Config(), 0)
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 256, (9, 500), dtype=np.uint8)
    got = ref.copy()
    flips = rng.random(got.shape) < 0.01
    got[flips] ^= rng.integers(1, 256, int(flips.sum()), dtype=np.uint8)
    lens = np.array([500, 0, 1, 499, 500, 123, 500, 500, 77], np.uint32)
    glen = lens.copy()
    status = np.zeros(9, np.int32)
    status[6] = 2                                              # failed stream
    glen[7] = 499                                              # length mismatch = failed
    c = eng.ber(ref, lens, got, glen, status)
    want = np.zeros(4, np.int64)
    for i in range(9):
        n = int(lens[i])
        if status[i] != 0 or glen[i] != n:
            want += [8 * n, n, 8 * n, 1]
        elif n:
            e, be, _ = oo.analysis(ref[i, :n], got[i, :n])
            want += [e, be, 8 * n, 0]
    assert c.tolist() == want.tolist()
    eng.close()


def test_channel_harness_statistics(ob, oo):
    """The device channel is a seeded restatement of src/channel.rs: taps, lead-in, CFO and SNR are checked statistically;
    the oracle then decodes what it produced."""
    eng, cfg, ocfg = _mk(ob, oo, 2, True, True, 1, 1, 1, 2048)
    rng = np.random.default_rng(5)
    pays = [rng.integers(0, 256, 576, dtype=np.uint8).tobytes() for _ in range(32)]
    iq, flen = eng.tx_encode(pays)
    # noiseless, no CFO: exactly the 12-tap convolution
    rx0, rl0, lead0, _ = eng.channel(iq, flen, ob.ChannelParams(snr_db=200.0, cfo_max=-1.0, lead_min=3, lead_max=3, seed=1))
    ref0 = oo.convolve(iq[0, : flen[0]].astype(np.complex128), np.r_[np.zeros(7), [-0.0, -0.1912, 0.9316, 0.2821, -0.1990, 0.1630, -0.1017, 0.0544, -0.0261, 0.0090, 0.0, -0.0034]])
    assert rl0[0] == flen[0] + 63 + 3 and lead0[0] == 3
    np.testing.assert_allclose(rx0[0, 3: 3 + flen[0] + 18], ref0[: flen[0] + 18], atol=1e-5)
    # full channel
    prm = ob.ChannelParams(snr_db=35.0, cfo_max=0.9 * np.pi / 80, lead_min=8, lead_max=1031, noise_mode=1, seed=99)
    rx, rl, lead, cfo = eng.channel(iq, flen, prm)
    assert (rl == flen + 63 + lead).all() and lead.min() >= 8 and lead.max() <= 1031 and len(set(lead.tolist())) > 16
    assert (cfo >= 0).all() and (cfo < 0.9 * np.pi / 80).all()
    res = eng.rx_decode(rx, rl, out_stride=640, diag=True)
    for i, p in enumerate(pays):
        ref = oo.decode(rx[i, : rl[i]].astype(np.complex128), ocfg, want_points=False, out_cap=640)
        assert ref.status == res.status[i] == 0
        assert ref.offset == res.offset[i] == lead[i] + 8
        assert abs(res.f_delta[i] - cfo[i]) < 2e-4                                  # estimator accuracy at 35 dB
        assert res.data[i] == ref.data.tobytes() == p
    # measured SNR of the noise-only lead-in vs the frame
    i = int(np.argmax(lead))
    pn = np.mean(np.abs(rx[i, : lead[i]]) ** 2)
    ps = np.mean(np.abs(rx[i, lead[i]: rl[i]]) ** 2)
    assert 10 * np.log10(ps / pn) == pytest.approx(35.0, abs=1.5)
    eng.close()


def test_channel_cfo_sample_exact(ob, oo):
    """C2 (src/channel.rs:54-62): y[i] *= exp(+j f (i + 1)) with the 1-based sample index, after the 12-tap convolution --
    the device channel against the oracle sample for sample (a 0-based index would be off by f ~ 0.02 rad, 1e4 x the bar)."""
    eng, cfg, ocfg = _mk(ob, oo, 2, True, True, 1, 1, 1, 2048)
    rng = np.random.default_rng(15)
    pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (576, 3000, 1, 576, 2000, 100, 576, 40)]
    iq, flen = eng.tx_encode(pays)
    for noise_mode in (0, 1):
        prm = ob.ChannelParams(snr_db=200.0, cfo_max=0.9 * np.pi / 80, lead_min=0, lead_max=37, noise_mode=noise_mode, seed=21 + noise_mode)
        rx, rl, lead, cfo = eng.channel(iq, flen, prm)
        assert len(set(cfo.tolist())) == len(pays) and (cfo > 0).all()
        for i in range(len(pays)):
            ref = oo.channel(iq[i, : flen[i]].astype(np.complex128), 200.0, float(cfo[i]), noise_mode, 5)
            assert rl[i] == lead[i] + ref.size
            got = rx[i, lead[i]: rl[i]].astype(np.complex128)
            assert np.abs(got - ref).max() < 2e-6
            # the check has teeth: the same signal with a 0-based phase index is far outside the bar
            assert np.abs(got - ref * np.exp(-1j * float(cfo[i]))).max() > 1e-3
            assert np.abs(rx[i, : lead[i]]).max(initial=0.0) < 1e-6          # lead-in: noise only
    eng.close()


def test_channel_reference_noise_mode(ob, oo):
    """C3 (src/channel.rs:66-71, src/signals/mod.rs:239-249): noise = sqrt(0.5 var / snr) (U(-1,1) + j U(-1,1)) with the
    reference's complex-valued, un-conjugated "variance" and the complex square root. The device draws come from Philox, not
    from the oracle's generator, so the amplitude is checked exactly and the draws through their support and moments:
    (noisy - noiseless) / amp_oracle must fill the square [-1, 1]^2 uniformly. A wrong |amp| or arg(amp) breaks the support."""
    import ctypes as C
    eng, cfg, ocfg = _mk(ob, oo, 2, True, True, 1, 1, 1, 2048)
    rng = np.random.default_rng(16)
    pays = [rng.integers(0, 256, 20000, dtype=np.uint8).tobytes() for _ in range(6)]
    iq, flen = eng.tx_encode(pays)
    kw = dict(cfo_max=0.03, lead_min=11, lead_max=11, noise_mode=0, seed=33)
    clean, rl, lead, cfo = eng.channel(iq, flen, ob.ChannelParams(snr_db=200.0, **kw))
    snr_db = 3.0                          # |var| of a circular signal is ~ power / sqrt(n): a low nominal SNR keeps the noise well above fp32 rounding
    noisy, rl2, lead2, cfo2 = eng.channel(iq, flen, ob.ChannelParams(snr_db=snr_db, **kw))
    assert (rl == rl2).all() and (lead == lead2).all() and (cfo == cfo2).all()

    class Z(C.Structure):
        _fields_ = [("re", C.c_double), ("im", C.c_double)]
    oo.lib().oo_variance.restype = Z
    oo.lib().oo_variance.argtypes = [C.c_void_p, C.c_size_t]
    for i in range(len(pays)):
        y = np.ascontiguousarray(clean[i, lead[i]: rl[i]].astype(np.complex128))
        v = oo.lib().oo_variance(y.ctypes.data, y.size)                     # sum (mean - y)^2 / len, no conjugate
        var = complex(v.re, v.im)
        assert abs(var - ((y.mean() - y) ** 2).sum() / y.size) < 1e-12 * abs(var)
        amp = np.sqrt(0.5 * var / 10 ** (snr_db / 10))                      # principal complex square root
        assert abs(amp) > 1e-4
        u = (noisy[i, lead[i]: rl[i]].astype(np.complex128) - y) / amp
        tol = 5e-7 / abs(amp)                                               # fp32 rounding of y + n relative to the amplitude
        for comp in (u.real, u.imag):
            assert comp.max() <= 1 + tol and comp.min() >= -1 - tol         # support [-1, 1]: exact amplitude, modulus and angle
            assert comp.max() > 0.9995 and comp.min() < -0.9995             # ... and it is filled to the edges (8e4 draws)
            assert abs(comp.mean()) < 0.01 and abs(comp.var() - 1 / 3) < 0.01
        assert abs(np.mean(u.real * u.imag)) < 0.01                          # independent components
        # the lead-in carries the same noise process
        ul = noisy[i, : lead[i]].astype(np.complex128) / amp
        assert np.abs(ul.real).max() <= 1 + tol and np.abs(ul.imag).max() <= 1 + tol
    eng.close()


def _config1_cfos():
    rng = np.random.default_rng(0xD0FD0001)
    return [0.0, 0.01, 0.02, 0.035] + (np.pi * rng.random(100) / 80).tolist()      # src/channel.rs:54: f = pi U(0,1) / 80


def _assert_bytes_equal_away_from_boundaries(got: bytes, ref: np.ndarray, points: np.ndarray, mod: int, fec: bool, tol: float, what):
    """Decoded bytes may differ from the oracle's only where a hard decision could flip under a `tol` perturbation: byte j
    is made of stream bits [128 + NB j, 128 + NB (j + 1)) (NB = 14 with Hamming, 8 without), i.e. of known carriers."""
    g = np.frombuffer(got, np.uint8)
    assert g.size == ref.size, what
    bad = np.flatnonzero(g != ref)
    if bad.size == 0:
        return 0
    bpc, nb = (1, 2, 6)[mod], (14 if fec else 8)
    risky = _near_boundary(points, mod, tol)
    for j in bad:
        c0, c1 = (128 + nb * j) // bpc, (128 + nb * j + nb - 1) // bpc
        assert risky[c0: c1 + 1].any(), f"{what}: byte {j} differs away from any decision boundary"
    return int(bad.size)


@pytest.mark.parametrize("mod,fec,modes", [(2, True, (1, 1, 1)), (2, False, (1, 1, 1)), (0, False, (0, 0, 0))])
def test_config1_sweep(ob, oo, golden, mod, fec, modes):
    """BASELINE.json configs[0] exactly as SURVEY.md 8(d) states it: support/dancing.bytes -> (FEC) -> encode(guard_bands) ->
    channel(snr, timing_error) -> decode, SNR {15, 20, 25, 30, 40} dB x CFO {0, .01, .02, .035} + 100 seeded draws of
    pi U(0,1)/80 (src/channel.rs:54), reference-faithful noise (src/channel.rs:66-71), seed 0xD0FD_0001. 520 captures per
    variant: 64QAM + Hamming(7,4) (north star), raw 64QAM, and what examples/lab3c_image.rs:15-42 really runs --
    BPSK + RS(255,223), reference sync / CFO / phase modes. At every point status, offset, f_delta and the equalised points
    agree with the oracle; bytes are identical except where the oracle's own point sits on a decision boundary."""
    sync, cfo_mode, phase = modes
    eng, cfg, ocfg = _mk(ob, oo, mod, True, fec, sync, cfo_mode, phase, 1024 if sync else 0)
    pay = golden["qam64_guard_fec_sc.payload"].tobytes()
    assert len(pay) == 576
    sent = pay if mod == 2 else oo.rs_encode(np.frombuffer(pay, np.uint8)).tobytes()        # examples/lab3c_image.rs:19-21
    tx = oo.tx(sent, ocfg)
    assert tx.size == {(2, True): 3120, (2, False): 2160, (0, False): 11280}[(mod, fec)]      # SURVEY.md 8(d) frame sizes
    iq_tx, flen = eng.tx_encode([sent])
    np.testing.assert_allclose(iq_tx[0, : flen[0]], tx, atol=2e-6)
    caps, meta = [], []
    for snr in (15, 20, 25, 30, 40):
        for k, f in enumerate(_config1_cfos()):
            caps.append(oo.channel(tx, float(snr), float(f), 0, 0xD0FD0001 + 1000 * snr + k))
            meta.append((snr, f))
    batch, n = _batch(caps)
    res = eng.rx_decode(batch, n, out_stride=1024, points=True)
    n_ok = n_clean = n_boundary = 0
    for i, (snr, f) in enumerate(meta):
        ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg, out_cap=1024)
        what = f"snr {snr} dB cfo {f:.5f}"
        assert res.status[i] == ref.status, what
        assert res.offset[i] == ref.offset, what
        if ref.status != 0:
            continue
        n_ok += 1
        assert abs(ref.f_delta - res.f_delta[i]) < ABS_TOL_FDELTA, what
        # the engine demodulates the symbols the header asks for, the oracle every symbol of the capture (channel tail included)
        bpc, dcar = (1, 2, 6)[mod], 48
        nsy = -(-(-(-(128 + 8 * ref.packet_length) // bpc)) // dcar)
        npts = min(nsy * dcar, ref.points.size, res.points.shape[1])
        scale = max(1.0, np.abs(ref.points[:npts]).max())
        assert np.abs(res.points[i, :npts] - ref.points[:npts]).max() <= REL_TOL_POINTS * scale, what
        n_boundary += _assert_bytes_equal_away_from_boundaries(res.data[i], ref.data, ref.points[:npts], mod, fec, REL_TOL_POINTS * scale, what)
        if ref.data.tobytes() == sent:
            assert res.data[i] == sent, what                                         # the pass criterion of SURVEY.md 8(d)
            n_clean += 1
    assert n_ok >= 400 and n_clean >= (100 if mod == 2 else 300), (n_ok, n_clean, n_boundary)
    eng.close()


def test_repeated_runs_are_bit_identical(ob, oo):
    """Race evidence without a sanitizer (compute-sanitizer is closed on this GPU pool, profiles/r2_sanitizer.txt): the RX
    kernels order their shared-memory exchanges with __syncwarp, named barriers and mbarriers (TMA staging); a missing
    ordering shows up as run-to-run differences. The same batch -- frames spanning several tiles, every output the kernels
    write (payload, points, h_k, offsets) -- is decoded 12 times, for nfft 64 and 1024, and must be bit-identical each time."""
    for nfft in (64, 1024):
        kw = dict(nfft=1024, cp=256, sync_window=4096) if nfft == 1024 else dict(sync_window=2048)
        cfg = ob.Config(modulation=2, guard_bands=True, fec=True, sync_mode=1, cfo_mode=1, phase_mode=1, **kw)
        eng = ob.Engine(cfg, 0)
        rng = np.random.default_rng(1234 + nfft)
        lens = [cfg.max_payload(700 if nfft == 64 else 60)] * 6 + [577, 1, 0, cfg.max_payload(225 if nfft == 64 else 29)]
        pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
        iq, flen = eng.tx_encode(pays)
        rx, rl, _, _ = eng.channel(iq, flen, ob.ChannelParams(snr_db=50.0 if nfft == 1024 else 38.0, cfo_max=0.9 * np.pi / (nfft * 5 // 4),
                                                              lead_min=8, lead_max=900, noise_mode=1, seed=5))
        first = None
        for rep in range(12):
            res = eng.rx_decode(rx, rl, out_stride=max(lens) + 16, points=True)
            blob = (tuple(res.data), res.points.tobytes(), res.h_k.tobytes(), tuple(res.offset), tuple(res.status), tuple(res.f_delta))
            if first is None:
                first = blob
                assert all(d == p for d, p in zip(res.data, pays))
            else:
                assert blob == first, f"nfft {nfft}: run {rep} differs from run 0"
        eng.close()


def test_host_and_device_paths_agree(ob, oo):
    import torch
    eng, cfg, ocfg = _mk(ob, oo, 2, True, True, 1, 1, 1, 2048)
    rng = np.random.default_rng(6)
    pays = [rng.integers(0, 256, 3000, dtype=np.uint8).tobytes() for _ in range(40)]
    iq, flen = eng.tx_encode(pays)
    rx, rl, _, _ = eng.channel(iq, flen, ob.ChannelParams(snr_db=40.0, cfo_max=0.03, lead_min=8, lead_max=500, noise_mode=1, seed=4))
    host = eng.rx_decode(rx, rl, out_stride=3008)
    d_iq = torch.from_numpy(rx.view(np.float32).reshape(40, rx.shape[1], 2)).cuda()
    d_n = torch.from_numpy(rl.astype(np.int32)).cuda()
    d_out = torch.zeros((40, 3008), dtype=torch.uint8, device="cuda")
    d_ol = torch.zeros(40, dtype=torch.int32, device="cuda")
    d_st = torch.zeros(40, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    eng.rx_decode_device(d_iq.data_ptr(), d_n.data_ptr(), 40, rx.shape[1], int(rl.max()), d_out.data_ptr(), 3008, d_ol.data_ptr(), d_st.data_ptr(), stream)
    torch.cuda.synchronize()
    assert (d_st.cpu().numpy() == host.status).all() and (host.status == 0).all()
    d_o, d_l = d_out.cpu().numpy(), d_ol.cpu().numpy()
    assert (d_l == host.out_len).all()
    assert all((d_o[i, : d_l[i]] == host.out[i, : d_l[i]]).all() for i in range(40))     # bytes past out_len are unspecified
    assert all(host.data[i] == pays[i] for i in range(40))
    eng.close()


def _capture_with_frames(oo, rng, n, stride, first, payload_len=300, sigma=0.01, mod=2):
    cfg = oo.make_cfg(True, mod, True)
    cap = (sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    truth = []
    p = first
    while p + oo.lib().oo_tx_len(payload_len, __import__("ctypes").byref(cfg)) + 200 < n:
        pay = rng.integers(0, 256, payload_len, dtype=np.uint8)
        tx = oo.tx(pay, cfg)
        f = float(rng.uniform(-0.035, 0.035))
        cap[p: p + tx.size] += (tx * np.exp(1j * f * np.arange(tx.size))).astype(np.complex64)
        truth.append((p, f))
        p += stride
    return cap, truth


def test_capture_sync_search_matches_oracle(ob, oo):
    """BASELINE.json configs[2] in small: frames every 100 003 samples (prime stride: all alignments) in a noise floor,
    each with its own CFO. Every inserted offset is found exactly (lag - 1 rule), CFO within 1e-4 rad/sample, and the
    engine's peak list equals the oracle's."""
    eng = ob.Engine(ob.Config(modulation=2, guard_bands=True, fec=True), 0)
    rng = np.random.default_rng(33)
    cap, truth = _capture_with_frames(oo, rng, 3_000_000, 100_003, 777)
    got = eng.sync_search(cap)
    ref = oo.sync_search(cap)
    assert len(got) == len(ref) == len(truth) >= 29
    for (p, f), g, r in zip(truth, got, ref):
        assert int(g["offset"]) == int(r["offset"]) == p - 1            # no channel: lock peak at lag p, offset = lag - 1
        assert abs(float(g["f_delta"]) - float(r["f_delta"])) < 1e-6
        assert abs(float(g["f_delta"]) - f) < 1e-4
        assert float(g["metric"]) == pytest.approx(float(r["metric"]), rel=1e-4) and float(g["metric"]) > 0.5
    # ascending order, and the found offsets decode
    assert (np.diff(got["offset"].astype(np.int64)) > 0).all()
    eng.close()


def test_capture_sync_search_edges(ob, oo):
    eng = ob.Engine(ob.Config(modulation=0, guard_bands=True), 0)
    rng = np.random.default_rng(34)
    noise = (0.01 * (rng.standard_normal(200_000) + 1j * rng.standard_normal(200_000))).astype(np.complex64)
    assert len(eng.sync_search(noise)) == len(oo.sync_search(noise)) == 0          # nothing there
    assert len(eng.sync_search(noise[:100])) == 0                                  # shorter than two preamble periods
    assert len(eng.sync_search(np.zeros(5000, np.complex64))) == 0                 # all zero: 0 > 0 is false
    # frames at the very start (offset would be -1 -> dropped), through the lab channel, and cut by the capture end
    cfg = oo.make_cfg(True, 0, False)
    tx = oo.tx(bytes(50), cfg)
    cap = noise[:60_000].copy()
    cap[: tx.size] += tx.astype(np.complex64)
    ch = oo.channel(tx, 30.0, 0.02, 1, 5).astype(np.complex64)
    cap[20_000: 20_000 + ch.size] += ch
    cap[-700:] += tx[:700].astype(np.complex64)
    got, ref = eng.sync_search(cap), oo.sync_search(cap)
    assert [int(x) for x in got["offset"]] == [int(x) for x in ref["offset"]] == [20_008]
    np.testing.assert_allclose(got["f_delta"], ref["f_delta"], atol=1e-6)
    # unaligned capture pointer (odd sample offset into a larger buffer) takes the 8-byte load path
    cap2, truth = _capture_with_frames(oo, rng, 400_001, 50_021, 1001, mod=1)
    got, ref = eng.sync_search(cap2[1:]), oo.sync_search(cap2[1:])
    assert [int(x) for x in got["offset"]] == [int(x) for x in ref["offset"]] == [p - 2 for p, _ in truth]
    eng.close()


def test_device_mode_sync_search_truncates_to_max_peaks(ob, oo):
    """with OFDM_MEM_DEVICE *n_peaks is clamped to max_peaks on the device (it is handed straight to
    ofdm_rx_decode_capture), and ofdm_sync_counts reports the truncation."""
    import torch
    eng = ob.Engine(ob.Config(modulation=1, guard_bands=True, fec=True, cfo_mode=1, phase_mode=1), 0)
    rng = np.random.default_rng(35)
    cap, truth = _capture_with_frames(oo, rng, 800_000, 50_021, 600, mod=1)
    assert len(truth) >= 12
    d_cap = torch.from_numpy(cap.view(np.float32).reshape(-1, 2)).cuda()
    for max_peaks in (5, 64):
        peaks = torch.zeros((max_peaks, 2), dtype=torch.int64, device="cuda")
        n_peaks = torch.full((1,), 12345, dtype=torch.int32, device="cuda")
        eng.sync_search_device(d_cap.data_ptr(), cap.size, peaks.data_ptr(), max_peaks, n_peaks.data_ptr())
        crossings, detections, written, overflowed = eng.sync_counts()
        assert overflowed == 0
        k = int(n_peaks.item())
        assert k == written == min(len(truth), max_peaks) and detections == len(truth) and crossings >= detections
        rec = peaks[:k].cpu().numpy().view(ob.engine.PEAK_DTYPE).reshape(-1)
        assert [int(x) for x in rec["offset"]] == [p - 1 for p, _ in truth[:k]]
        # ... and the clamped count decodes without touching anything past the max_peaks-sized buffers
        out = torch.zeros((max_peaks + 1, 1024), dtype=torch.uint8, device="cuda")
        out_len = torch.zeros(max_peaks + 1, dtype=torch.int32, device="cuda")
        status = torch.full((max_peaks + 1,), -7, dtype=torch.int32, device="cuda")
        eng.decode_capture_device(d_cap.data_ptr(), cap.size, peaks.data_ptr(), k, 0, out.data_ptr(), 1024, out_len.data_ptr(), status.data_ptr())
        torch.cuda.synchronize()
        assert (status[:k] == 0).all().item() and int(status[k].item()) == -7
    eng.close()


def test_more_than_65535_streams_per_call(ob, oo):
    """The stream index rides on gridDim.y (<= 65535): larger batches are launched in chunks (TX, channel, RX)."""
    import torch
    n = 66_000
    cfg = ob.Config(modulation=1, guard_bands=False, fec=False, sync_mode=0, cfo_mode=1, phase_mode=1, sync_window=16)
    eng = ob.Engine(cfg, 0)
    ocfg = oo.make_cfg(False, 1, False, 0, 1, 1, 16)
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    plen_b = 12
    flen_b = cfg.frame_len(plen_b)
    payload = torch.randint(0, 256, (n, 16), dtype=torch.uint8, device="cuda", generator=g)
    plen = torch.full((n,), plen_b, dtype=torch.int32, device="cuda")
    tx = torch.zeros((n, flen_b, 2), dtype=torch.float32, device="cuda")
    flen = torch.zeros(n, dtype=torch.int32, device="cuda")
    eng.tx_encode_device(payload.data_ptr(), plen.data_ptr(), 16, n, tx.data_ptr(), flen_b, flen.data_ptr())
    stride = flen_b + 63 + 4
    rx = torch.zeros((n, stride, 2), dtype=torch.float32, device="cuda")
    rl = torch.zeros(n, dtype=torch.int32, device="cuda")
    eng.channel_device(tx.data_ptr(), flen.data_ptr(), flen_b, n, ob.ChannelParams(snr_db=45.0, cfo_max=0.02, lead_min=0, lead_max=4, noise_mode=1, seed=3),
                       rx.data_ptr(), stride, rl.data_ptr())
    out = torch.zeros((n, 16), dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(n, dtype=torch.int32, device="cuda")
    status = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    eng.rx_decode_device(rx.data_ptr(), rl.data_ptr(), n, stride, stride, out.data_ptr(), 16, out_len.data_ptr(), status.data_ptr())
    torch.cuda.synchronize()
    assert (flen == flen_b).all().item() and (status == 0).all().item() and (out_len == plen_b).all().item()
    assert (out[:, :plen_b] == payload[:, :plen_b]).all().item()
    for i in (0, 65_534, 65_535, 65_999):                        # either side of the chunk boundary, against the oracle
        cap = rx[i, : int(rl[i])].cpu().numpy().view(np.complex64).reshape(-1).astype(np.complex128)
        ref = oo.decode(cap, ocfg, want_points=False)
        assert ref.status == 0 and ref.data.tobytes() == bytes(out[i, :plen_b].cpu().numpy())
    eng.close()


def test_capture_search_dense_short_frames(ob, oo):
    """More than 50 000 back-to-back short frames in one capture (the old design capped a search at 8 192 candidates): every
    frame start is found, in order, and the list equals the oracle's."""
    eng = ob.Engine(ob.Config(modulation=0, guard_bands=True), 0)
    cfg = oo.make_cfg(True, 0, False)
    tx = oo.tx(b"", cfg).astype(np.complex64)                                       # 13 symbols = 1040 samples
    assert tx.size == 1040
    rng = np.random.default_rng(51)
    n_frames = 50_500
    gaps = rng.integers(0, 200, n_frames)
    starts = 700 + np.cumsum(tx.size + gaps) - (tx.size + gaps[0])
    n = int(starts[-1]) + tx.size + 500
    cap = (0.004 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    cfos = rng.uniform(-0.03, 0.03, n_frames)
    t = np.arange(tx.size)
    for p, f in zip(starts, cfos):
        cap[p: p + tx.size] += tx * np.exp(1j * f * t).astype(np.complex64)
    got = eng.sync_search(cap, max_peaks=60_000)
    crossings, detections, written, overflowed = eng.sync_counts()
    assert overflowed == 0 and detections == written == len(got) and crossings >= detections
    ref = oo.sync_search(cap, max_peaks=60_000)
    assert len(got) == len(ref) >= n_frames
    assert (got["offset"] == ref["offset"]).all()
    np.testing.assert_allclose(got["f_delta"], ref["f_delta"], atol=1e-6)
    found = set(int(x) for x in got["offset"])
    # lag - 1 rule, no channel; a frame that follows its predecessor within a few samples can be detected a little off
    assert sum((int(p) - 1) in found for p in starts) >= 0.99 * n_frames
    assert (np.diff(got["offset"].astype(np.int64)) > 0).all()
    eng.close()


def test_capture_search_beyond_2_pow_32_samples(ob, oo):
    """64-bit offsets: frames on both sides of sample 2^32 of a 4.4e9-sample device capture (35 GB of HBM)."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 45 << 30:
        pytest.skip("needs 45 GB of free device memory")
    eng = ob.Engine(ob.Config(modulation=1, guard_bands=True, fec=True, cfo_mode=1, phase_mode=1), 0)
    cfg = oo.make_cfg(True, 1, True)
    pay = bytes(range(200))
    tx = torch.from_numpy(oo.tx(pay, cfg).astype(np.complex64).view(np.float32).reshape(-1, 2)).cuda()
    n = 4_400_000_123
    cap = torch.zeros((n, 2), dtype=torch.float32, device="cuda")                   # an all-zero floor never crosses the threshold (0 > 0)
    starts = [5_000, (1 << 32) - 1_000, (1 << 32) + 77_777, n - tx.shape[0] - 3]
    for p in starts:
        cap[p: p + tx.shape[0]] += tx
    peaks = torch.zeros((16, 2), dtype=torch.int64, device="cuda")
    n_peaks = torch.zeros(1, dtype=torch.int32, device="cuda")
    eng.sync_search_device(cap.data_ptr(), n, peaks.data_ptr(), 16, n_peaks.data_ptr())
    k = int(n_peaks.item())
    rec = peaks[:k].cpu().numpy().view(ob.engine.PEAK_DTYPE).reshape(-1)
    assert [int(x) for x in rec["offset"]] == [p - 1 for p in starts]
    out = torch.zeros((k, 256), dtype=torch.uint8, device="cuda")
    out_len = torch.zeros(k, dtype=torch.int32, device="cuda")
    status = torch.full((k,), -1, dtype=torch.int32, device="cuda")
    eng.decode_capture_device(cap.data_ptr(), n, peaks.data_ptr(), k, 1 << 16, out.data_ptr(), 256, out_len.data_ptr(), status.data_ptr())
    torch.cuda.synchronize()
    # the lag - 1 rule puts the frame start one sample early; the decoder then finds lag 1 itself
    assert (status == 0).all().item() and all(bytes(out[i, :200].cpu().numpy()) == pay for i in range(k))
    del cap
    eng.close()


def test_streaming_receiver_decodes_every_frame_of_a_capture(ob, oo):
    """SURVEY 8f rank 3 / examples/jetson_rx.rs: one long capture -> every frame found and decoded in two launches groups.
    Each decoded payload equals what the oracle decodes from the same frame slice."""
    cfgk = dict(modulation=2, guard_bands=True, fec=True, cfo_mode=1, phase_mode=1)
    eng = ob.Engine(ob.Config(**cfgk), 0)
    ocfg = oo.make_cfg(True, 2, True, oo.SYNC_REFERENCE, 1, 1)
    rng = np.random.default_rng(77)
    n = 1_200_000
    cap = (0.0003 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)     # floor ~49 dB below the frames
    sent, p = [], 4321
    while p + 40_000 < n:
        pay = rng.integers(0, 256, int(rng.integers(1, 9000)), dtype=np.uint8).tobytes()
        ch = oo.channel(oo.tx(pay, ocfg), 38.0, float(rng.uniform(0, 0.03)), 1, p).astype(np.complex64)
        cap[p: p + ch.size] += ch
        sent.append((p, pay))
        p += ch.size + int(rng.integers(900, 30_000))
    peaks, data, status = eng.decode_capture(cap, out_stride=9008)
    assert len(peaks) == len(sent) and (status == 0).all()
    for (p, pay), pk, got in zip(sent, peaks, data):
        assert int(pk["offset"]) == p + 8                       # main channel tap at delay 9, lag - 1 rule
        assert got == pay
    # against the oracle on one frame slice (known start -> reference sync finds lag 1 -> offset 0)
    i = 3
    o = int(peaks[i]["offset"])
    ref = oo.decode(np.concatenate([[0], cap[o: int(peaks[i + 1]["offset"])]]).astype(np.complex128), ocfg, want_points=False, out_cap=9008)
    assert ref.status == 0 and ref.data.tobytes() == data[i]
    eng.close()


@pytest.mark.parametrize("mod,guard,fec,modes", [(2, True, True, (1, 1, 1)), (2, True, False, (0, 0, 0)), (1, False, True, (1, 1, 0)),
                                                 (0, True, True, (0, 1, 1)), (2, False, False, (1, 1, 1)), (1, True, False, (0, 0, 0))])
def test_wideband_1024_parity(ob, oo, mod, guard, fec, modes):
    """BASELINE.json configs[3]: 1024-subcarrier variant (docs/SPEC.md 9), TX + RX round trip against the oracle."""
    sync, cfo, phase = modes
    win = 0 if sync == 0 else 4096
    cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec, sync_mode=sync, cfo_mode=cfo, phase_mode=phase, sync_window=win, nfft=1024, cp=256)
    eng = ob.Engine(cfg, 0)
    ocfg = oo.make_cfg(guard, mod, fec, sync, cfo, phase, win, nfft=1024)
    rng = np.random.default_rng(7 * mod + guard)
    # lengths around the 28-symbol tile and 7-symbol Hamming boundaries, empty and ragged
    lens = [cfg.max_payload(60), cfg.max_payload(29) + 1, cfg.max_payload(28), 0, 1, 577]
    pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    iq, flen = eng.tx_encode(pays)
    caps = []
    for i, p in enumerate(pays):
        ref = oo.tx(p, ocfg)
        assert ref.size == flen[i]
        np.testing.assert_allclose(iq[i, : flen[i]], ref, atol=2e-6)
        lead = int(rng.integers(0, 900))
        c = oo.channel(ref, 60.0, 0.0012 + 0.0002 * i, 1, 50 + i)
        caps.append(np.concatenate([1e-4 * (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)), c]))
    batch, n = _batch(caps)
    res = eng.rx_decode(batch, n, points=True)
    for i, p in enumerate(pays):
        ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg)
        assert ref.status == res.status[i] == 0 and ref.offset == res.offset[i]
        assert abs(ref.f_delta - res.f_delta[i]) < ABS_TOL_FDELTA
        np.testing.assert_allclose(res.h_k[i], ref.h_k, atol=REL_TOL_POINTS * np.abs(ref.h_k).max())
        npts = cfg.frame_data_syms(len(p)) * cfg.data_carriers
        assert np.abs(res.points[i, :npts] - ref.points[:npts]).max() <= REL_TOL_POINTS * max(1.0, np.abs(ref.points[:npts]).max())
        assert res.data[i] == ref.data.tobytes() == p
    # status paths: too short, nothing there
    short = eng.rx_decode(batch[:1, :9000], np.array([9000], np.uint32))
    assert short.status[0] in (ob.TOO_SHORT, ob.NO_SYNC, ob.NEG_OFFSET) and short.out_len[0] == 0
    assert oo.decode(batch[0, :9000].astype(np.complex128), ocfg).status == short.status[0]
    eng.close()


@pytest.mark.parametrize("mod,guard,fec", [(2, True, True), (2, False, False), (1, True, True), (0, False, True), (0, True, False)])
def test_tx_resident_kernel_parity(ob, oo, monkeypatch, mod, guard, fec):
    """The one-pass TX kernel (frames resident in tensor memory, tx_resident.cuh) is what large batches run; here it is forced
    for a ragged batch: every frame against the oracle's encode (src/transmitter.rs:11-58), zero fill past the frame, and
    value-for-value against the two-pass kernel. Lengths include the empty payload, one byte, and frames of 1, 2 and several CTAs'
    worth of symbols; iq_stride admits frames that need a group of 3 CTAs."""
    rng = np.random.default_rng(77 + 10 * mod + 2 * guard + fec)
    cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec)
    longest = cfg.max_payload(1300)
    lens = [0, 1, 2, 15, 16, 17, 100, 333, 1000, longest // 3, longest // 2, longest - 1, longest] + [int(v) for v in rng.integers(0, longest + 1, 40)]
    pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    # payloads of equal bytes: identical carriers, a transform that piles up in one sample -- the data maximum beats the head's,
    # which is the case the speculative kernel has to notice and hand to the redo pass
    # (all-ones bits put +1 on every BPSK / QPSK carrier, also through the Hamming code; "100100..." is 64QAM's (+1, +1) point)
    pays += [bytes(1000), b"\xff" * 3000, bytes(longest // 2), b"\x55" * (longest - 3), b"\x49\x92\x24" * 700, b"\xff" * (longest - 1)]
    out = {}
    for path in ("resident", "warp", "spec", "twopass"):
        monkeypatch.setenv("OFDM_TX_PATH", path)
        eng = ob.Engine(cfg, 0)
        l0 = eng.kernel_launches
        out[path] = eng.tx_encode(pays)
        assert eng.kernel_launches - l0 == (2 if path in ("twopass", "spec") else 1)
        eng.close()
    ocfg = oo.make_cfg(guard, mod, fec, 0, 0, 0, 0)
    beaten = 0
    for i, p in enumerate(pays):
        ref = oo.tx(p, ocfg)
        head, data = ref[:800], ref[800:]
        beaten += data.size > 0 and max(data.real.max(), data.imag.max()) > max(head.real.max(), head.imag.max())
    assert beaten >= 1 or (mod == 2 and fec)                # (64QAM through the Hamming code: no crafted payload here; the other four exercise the redo pass)
    for path in ("resident", "warp", "spec"):               # "warp": barrier-free frame loop; "spec": bet on the head maximum + redo pass
        iq, flen = out[path]
        for i, p in enumerate(pays):
            ref = oo.tx(p, ocfg)
            assert ref.size == flen[i]
            np.testing.assert_allclose(iq[i, : flen[i]], ref, atol=2e-6)
            assert not iq[i, flen[i]:].any()
        assert np.array_equal(flen, out["twopass"][1])
        assert np.array_equal(iq, out["twopass"][0])        # the same values (an exact zero may carry the other sign: conj vs swap transform)


@pytest.mark.parametrize("db", [0, 1])
@pytest.mark.parametrize("mod,guard,fec", [(2, True, True), (2, False, False), (1, True, True), (0, False, True), (2, True, False)])
def test_wide_tx_resident_kernel_parity(ob, oo, monkeypatch, mod, guard, fec, db):
    """nfft = 1024: the one-pass TX kernel (frames resident in tensor memory, wide_tx_resident.cuh) forced for a ragged batch:
    every frame against the oracle's encode (src/transmitter.rs:11-58 scaled by 16, docs/SPEC.md 9), zero fill past the frame,
    frame lengths and (to rounding: the two kernels use different FFT factorisations) the two-pass kernel. Lengths include the
    empty payload, one byte, the header-only symbol, and frames of 1, 2 and 3 CTAs' worth of symbols (32 symbols per CTA)."""
    rng = np.random.default_rng(177 + 10 * mod + 2 * guard + fec)
    cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec, nfft=1024, cp=256)
    longest = cfg.max_payload(70)
    lens = [0, 1, 2, 15, 16, 17, 100, 333, 577, cfg.max_payload(1), cfg.max_payload(1) + 1, cfg.max_payload(32), cfg.max_payload(33),
            longest // 2, longest - 1, longest] + [int(v) for v in rng.integers(0, longest + 1, 12)]
    pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
    # payloads whose carriers are all equal pile the inverse transform up in one sample: the data maximum beats the head's, which
    # the speculative kernel (wide_tx_spec_kernel, the large-batch default) has to notice and hand to its redo pass
    pays += [b"\xff" * 3000, b"\x49\x92\x24" * 1500, b"\xff" * (longest - 1)]
    out = {}
    monkeypatch.setenv("OFDM_WTX_DB", str(db))          # 0: two symbols of a frame per warp; 1: one symbol of two frames in flight
    for path in ("resident", "spec", "twopass"):
        monkeypatch.setenv("OFDM_TX_PATH", path)
        eng = ob.Engine(cfg, 0)
        l0 = eng.kernel_launches
        out[path] = eng.tx_encode(pays)
        assert eng.kernel_launches - l0 == (1 if path == "resident" else 2)
        eng.close()
    ocfg = oo.make_cfg(guard, mod, fec, 0, 0, 0, 0, nfft=1024)
    refs = [oo.tx(p, ocfg) for p in pays]
    beaten = sum(r.size > 12800 and max(r[12800:].real.max(), r[12800:].imag.max()) > max(r[:12800].real.max(), r[:12800].imag.max()) for r in refs)
    assert beaten >= 1 or (mod == 2 and fec)                # (the redo pass of the speculative path is exercised)
    for path in ("resident", "spec"):
        iq, flen = out[path]
        for i, p in enumerate(pays):
            ref = refs[i]
            assert ref.size == flen[i]
            np.testing.assert_allclose(iq[i, : flen[i]], ref, atol=2e-6)
            assert not iq[i, flen[i]:].any()
            assert abs(max(iq[i].real.max(), iq[i].imag.max()) - 1.0) < 1e-6      # normalize, src/transmitter.rs:183-194
        assert np.array_equal(flen, out["twopass"][1])
        np.testing.assert_allclose(iq, out["twopass"][0], atol=2e-6)


def test_wide_tx_resident_kernel_is_the_large_batch_path(ob, oo):
    """nfft = 1024: a batch of a few hundred frames takes the one-pass kernel by itself (one launch); frames that do not fit
    iq_stride come back zeroed with their required length, as from the two-pass kernel; RX decodes what TX produced."""
    cfg = ob.Config(modulation=2, guard_bands=True, fec=True, sync_mode=ob.SYNC_SCHMIDL_COX, cfo_mode=ob.CFO_ANGLE_OF_SUM,
                    phase_mode=ob.PHASE_ANGLE_OF_SUM, sync_window=4096, nfft=1024, cp=256)
    rng = np.random.default_rng(6)
    n = 330
    plen = cfg.max_payload(40)
    pays = [rng.integers(0, 256, plen - (i % 9) * 311, dtype=np.uint8).tobytes() for i in range(n)]
    eng = ob.Engine(cfg, 0)
    l0 = eng.kernel_launches
    iq, flen = eng.tx_encode(pays)
    assert eng.kernel_launches - l0 == 2                   # the speculative one-pass kernel + the redo pass (which exits at once here)
    ocfg = oo.make_cfg(True, 2, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, 4096, nfft=1024)
    for i in (0, 1, 36, 37, 147, 148, 329):
        np.testing.assert_allclose(iq[i, : flen[i]], oo.tx(pays[i], ocfg), atol=2e-6)
    mx = np.maximum(iq.real.max(axis=1), iq.imag.max(axis=1))
    np.testing.assert_allclose(mx, 1.0, atol=1e-6)
    pick = list(range(0, n, 10))
    lead = 1e-4 * (rng.standard_normal(60) + 1j * rng.standard_normal(60))
    batch, ns = _batch([np.concatenate([lead, iq[i, : flen[i]]]) for i in pick])
    res = eng.rx_decode(batch, ns)
    assert all(res.status[j] == 0 and res.data[j] == pays[i] for j, i in enumerate(pick))
    eng.close()


def test_tx_resident_kernel_is_the_large_batch_path(ob, oo):
    """A batch of a few hundred frames takes the one-pass kernel by itself (one launch instead of two); spot-check frames
    against the oracle and the normalisation of every frame (max positive component 1, src/transmitter.rs:183-194)."""
    cfg = ob.Config(modulation=2, guard_bands=True, fec=True)
    rng = np.random.default_rng(5)
    n = 600
    plen = cfg.max_payload(64)
    pays = [rng.integers(0, 256, plen - (i % 7), dtype=np.uint8).tobytes() for i in range(n)]
    eng = ob.Engine(cfg, 0)
    l0 = eng.kernel_launches
    iq, flen = eng.tx_encode(pays)
    assert eng.kernel_launches - l0 == 2                   # the speculative one-pass kernel + the redo pass (which exits at once here)
    eng.close()
    ocfg = oo.make_cfg(True, 2, True, 0, 0, 0, 0)
    for i in (0, 1, 147, 148, 299, 599):
        np.testing.assert_allclose(iq[i, : flen[i]], oo.tx(pays[i], ocfg), atol=2e-6)
    mx = np.maximum(iq.real.max(axis=1), iq.imag.max(axis=1))
    np.testing.assert_allclose(mx, 1.0, atol=1e-6)


@pytest.mark.parametrize("nfft", [64, 1024])
def test_host_feed_without_cyclic_prefixes_matches_full_copy(ob, oo, monkeypatch, nfft):
    """Host-mode rx_decode fetches the head region of every stream, locates the frame, and then copies only the useful nfft
    samples of the data symbols the header asks for (rx_host_skip_cp: `unprefix_block`, src/receiver.rs:104-118, discards the
    prefixes anyway). Same status / lengths / bytes / offset / f_delta / h_k / points as the whole-capture copy, and as the
    oracle, on a ragged batch: frames of different lengths and lead-ins, a capture cut in the middle of a data symbol (zero-padded
    tail row, src/receiver.rs:206-210), one cut inside the head, pure noise, noise after the frame; several chunks of streams."""
    L = nfft + nfft // 4
    win = 2048 if nfft == 64 else 4096
    cfg = ob.Config(modulation=2, guard_bands=True, fec=True, sync_mode=ob.SYNC_SCHMIDL_COX, cfo_mode=ob.CFO_ANGLE_OF_SUM,
                    phase_mode=ob.PHASE_ANGLE_OF_SUM, sync_window=win, nfft=nfft, cp=nfft // 4)
    ocfg = oo.make_cfg(True, 2, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, win, nfft=nfft)
    rng = np.random.default_rng(900 + nfft)
    syms = [400, 399, 250, 1, 2, 120] if nfft == 64 else [62, 61, 12, 1, 2, 25]
    lens = [cfg.max_payload(s) - int(rng.integers(0, 40)) for s in syms] + [0]
    pays = [rng.integers(0, 256, max(n, 0), dtype=np.uint8).tobytes() for n in lens]
    caps = []
    for i, p in enumerate(pays):
        tx = oo.tx(p, ocfg)
        lead = int(rng.integers(0, win - 200))
        c = oo.channel(tx, 45.0 if nfft == 64 else 60.0, 0.0015 / (nfft // 64), 1, 300 + i)
        noise = lambda k: 1e-4 * (rng.standard_normal(k) + 1j * rng.standard_normal(k))
        caps.append(np.concatenate([noise(lead), c, noise(int(rng.integers(0, 5 * L)))]))
    caps.append(caps[0][: caps[0].size - 3 * L - L // 3])      # cut in the middle of a data symbol: the header asks for more than is there
    caps.append(caps[1][: 7 * L])                              # cut inside the head
    caps.append(1e-3 * (rng.standard_normal(20 * L) + 1j * rng.standard_normal(20 * L)))       # nothing there
    reps = 70 if nfft == 64 else 24                            # > one 128 MB chunk of streams
    import torch
    pageable, ns = _batch(caps * reps)
    pinned = torch.empty(pageable.shape, dtype=torch.complex64, pin_memory=True)   # the gather kernel reads the capture in place
    batch = pinned.numpy()
    batch[:] = pageable
    assert 4 * (win + 13 * L) <= batch.shape[1] and batch.nbytes > (160 << 20)
    res = {}
    for feed in ("full", "skipcp", "gather", "", "pageable"):
        if feed in ("", "pageable"):
            monkeypatch.delenv("OFDM_RX_FEED", raising=False)
        else:
            monkeypatch.setenv("OFDM_RX_FEED", feed)
        eng = ob.Engine(cfg, 0)
        res[feed] = eng.rx_decode(pageable if feed == "pageable" else batch, ns, points=True)
        res[feed].moved = eng.last_h2d_bytes
        eng.close()
    full, skip = res["full"], res["skipcp"]
    assert full.moved == batch.nbytes and skip.moved < 0.9 * batch.nbytes
    assert res["gather"].moved == skip.moved
    # pinned capture: the automatic choice is the gather kernel for symbols of >= 4 KB (nfft 1024); with nfft 64 the 128-byte
    # fetch granularity of host memory eats the saving (measured), so the capture is copied whole
    assert res[""].moved == (skip.moved if nfft == 1024 else batch.nbytes)
    assert res["pageable"].moved == batch.nbytes                                      # pageable capture: copied whole
    for r in (skip, res["gather"], res[""], res["pageable"]):
        assert np.array_equal(r.status, full.status) and np.array_equal(r.out_len, full.out_len)
        assert np.array_equal(r.offset, full.offset) and np.array_equal(r.f_delta, full.f_delta)
        assert np.array_equal(r.h_k, full.h_k) and np.array_equal(r.n_data_syms, full.n_data_syms)
        assert r.data == full.data
        assert np.array_equal(r.points, full.points)
    for i, c in enumerate(caps):
        ref = oo.decode(batch[i, : ns[i]].astype(np.complex128), ocfg, want_points=False)
        assert ref.status == skip.status[i], (i, ref.status, skip.status[i])
        if ref.status == 0:
            assert ref.offset == skip.offset[i] and skip.data[i] == ref.data.tobytes()
            if i < len(pays):
                assert skip.data[i] == pays[i]
    assert (full.status[: len(pays)] == 0).sum() >= len(pays) - 1 and full.status[len(pays) + 2] == ob.NO_SYNC
