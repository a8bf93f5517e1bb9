"""Capture search, streaming receiver and file ingest for the 1024-subcarrier layout (docs/SPEC.md 4 and 9): the nfft = 64
machinery with every length scaled by 16, checked against the CPU oracle's search and the transmitted payloads."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfgs(ob, oo, mod=2, fec=True):
    cfg = ob.Config(nfft=1024, cp=256, modulation=mod, guard_bands=True, fec=fec, sync_mode=1, cfo_mode=1, phase_mode=1, sync_window=4096)
    ocfg = oo.make_cfg(True, mod, fec, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, 4096, nfft=1024)
    return cfg, ocfg


def _wide_capture(oo, rng, n, ocfg, max_pay, gap=(3000, 60_000), sigma=3e-4, first=5001, back_to_back=False):
    cap = (sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    sent, p = [], first
    while True:
        pay = rng.integers(0, 256, int(rng.integers(1, max_pay)), dtype=np.uint8).tobytes()
        tx = oo.tx(pay, ocfg)
        f = float(rng.uniform(-0.002, 0.002))
        if p + tx.size + 100 > n:
            break
        cap[p: p + tx.size] += (tx * np.exp(1j * f * np.arange(tx.size))).astype(np.complex64)
        sent.append((p - 1, pay, f))                                    # no channel: lock peak at lag p, offset = lag - 1
        p += tx.size + (0 if back_to_back and len(sent) % 3 else int(rng.integers(*gap)))
    return cap, sent


def test_wide_sync_search_matches_oracle(oo):
    import ofdm_b200 as ob
    cfg, ocfg = _cfgs(ob, oo)
    eng = ob.Engine(cfg, 0)
    rng = np.random.default_rng(101)
    cap, sent = _wide_capture(oo, rng, 1_500_000, ocfg, 9000, back_to_back=True)
    assert len(sent) >= 20
    got, ref = eng.sync_search(cap), oo.sync_search(cap, nfft=1024)
    assert [int(x) for x in got["offset"]] == [int(x) for x in ref["offset"]] == [p for p, _, _ in sent]
    np.testing.assert_allclose(got["f_delta"], ref["f_delta"], atol=1e-6)
    np.testing.assert_allclose(got["f_delta"], [f for _, _, f in sent], atol=2e-5)
    np.testing.assert_allclose(got["metric"], ref["metric"], rtol=1e-4)
    # unaligned capture pointer, a frame cut by the capture end, nothing / too little to search
    got, ref = eng.sync_search(cap[3:700_001]), oo.sync_search(cap[3:700_001], nfft=1024)
    assert [int(x) for x in got["offset"]] == [int(x) for x in ref["offset"]] and len(got) > 5
    # aligned capture whose length is not a multiple of 8 (the tail samples are patched into the TMA-staged tile), ending inside a frame head
    for cut in (sent[9][0] + 12_801, sent[9][0] + 12_799, sent[15][0] + 2563):
        got, ref = eng.sync_search(cap[:cut]), oo.sync_search(cap[:cut], nfft=1024)
        assert [int(x) for x in got["offset"]] == [int(x) for x in ref["offset"]], cut
    noise = cap[:4000].copy()
    assert len(eng.sync_search(noise)) == len(oo.sync_search(noise, nfft=1024)) == 0
    assert len(eng.sync_search(noise[:2000])) == 0                      # shorter than two symbol lengths
    assert len(eng.sync_search(np.zeros(50_000, np.complex64))) == 0
    eng.close()


def test_wide_decode_capture_round_trip(oo):
    import ofdm_b200 as ob
    for mod, fec in ((2, True), (1, False)):
        cfg, ocfg = _cfgs(ob, oo, mod, fec)
        eng = ob.Engine(cfg, 0)
        rng = np.random.default_rng(7 + mod)
        cap, sent = _wide_capture(oo, rng, 1_200_000, ocfg, 12_000)
        peaks, data, status = eng.decode_capture(cap, max_frame_samples=120_000, out_stride=12_032)
        assert [int(x) for x in peaks["offset"]] == [p for p, _, _ in sent]
        assert (status == 0).all() and data == [pay for _, pay, _ in sent]
        # each frame decoded on its own by the batch path gives the same bytes (known start vs searched start)
        i = len(sent) // 2
        p = sent[i][0]
        one = eng.rx_decode(cap[p - 2000: p + 110_000][None, :], out_stride=12_032, diag=True)
        assert one.status[0] == 0 and one.data[0] == sent[i][1] and one.offset[0] == 2000
        eng.close()


def test_wide_host_search_is_pipelined_in_chunks(oo):
    """More than one 64 MB copy chunk: the tiles a chunk completes are scanned while the next chunk is on the link."""
    import ofdm_b200 as ob
    cfg, ocfg = _cfgs(ob, oo)
    eng = ob.Engine(cfg, 0)
    rng = np.random.default_rng(11)
    n = (8 << 20) + 300_000
    cap = np.zeros(n, np.complex64)
    cap[:] = 3e-4 * (rng.standard_normal(n).astype(np.float32) + 1j * rng.standard_normal(n).astype(np.float32))
    sent = []
    for p in (4097, (8 << 20) - 30_000, (8 << 20) - 700, (8 << 20) + 150_000):        # around the chunk boundary
        pay = rng.integers(0, 256, 2000, dtype=np.uint8).tobytes()
        tx = oo.tx(pay, ocfg).astype(np.complex64)
        if sent and p < sent[-1][0] + sent[-1][2] + 13_000:
            continue
        cap[p: p + tx.size] += tx
        sent.append((p - 1, pay, tx.size))
    peaks, data, status = eng.decode_capture(cap, max_frame_samples=40_000, out_stride=2048)
    assert [int(x) for x in peaks["offset"]] == [p for p, _, _ in sent]
    assert (status == 0).all() and data == [pay for _, pay, _ in sent]
    eng.close()


def test_wide_file_ingest(tmp_path, oo):
    import ofdm_b200 as ob
    from ofdm_b200 import ingest
    cfg, ocfg = _cfgs(ob, oo)
    rng = np.random.default_rng(21)
    cap, sent = _wide_capture(oo, rng, 2_000_000, ocfg, 6000)
    assert len(sent) > 15
    path = tmp_path / "wide.dat"
    path.write_bytes(ob.sig_to_bytes(cap))
    eng = ob.Engine(cfg, 0)
    for chunk in (300_000, 1 << 22):
        rec, data = eng.decode_file(str(path), chunk_samples=chunk, max_frame_samples=70_000, out_stride=6016, max_frames=256)
        assert [int(x) for x in rec["offset"]] == [p for p, _, _ in sent] and (rec["status"] == 0).all()
        assert data == [pay for _, pay, _ in sent]
    eng.close()
    frames = ingest.decode_file(str(path), cfg, chunk_samples=300_000, max_frame_samples=70_000, out_stride=6016)
    assert [(f.offset, f.status, f.data) for f in frames] == [(p, 0, pay) for p, pay, _ in sent]
    rx = ingest.StreamReceiver(cfg, chunk_samples=262_144, max_frame_samples=70_000, out_stride=6016)
    got, pos = [], 0
    while pos < cap.size:
        k = int(rng.integers(1, 90_000))
        got += rx.push(cap[pos: pos + k])
        pos += k
    got += rx.flush()
    rx.close()
    assert [(f.offset, f.status, f.data) for f in got] == [(p, 0, pay) for p, pay, _ in sent]
