"""fc32 file / ring-buffer ingest: chunk arithmetic on CPU, file and push paths on the GPU."""
import numpy as np
import pytest


# ---- CPU: host logic --------------------------------------------------------------------------------------------------
def test_plan_chunks_every_frame_start_is_owned_exactly_once():
    from ofdm_b200.ingest import plan_chunks, accept_frame
    rng = np.random.default_rng(1)
    for _ in range(200):
        overlap = int(rng.integers(1, 500))
        chunk = overlap + int(rng.integers(1, 2000))
        n = int(rng.integers(1, 20000))
        chunks = plan_chunks(n, chunk, overlap)
        assert chunks[0][0] == 0 and chunks[-1][1] == n
        for (a0, b0), (a1, b1) in zip(chunks, chunks[1:]):
            assert a1 == b0 - overlap and b0 - a0 == chunk               # full chunks, advancing by chunk - overlap
        for pos in rng.integers(0, n, 50):
            owners = [(a, b) for i, (a, b) in enumerate(chunks) if a <= pos < b and accept_frame(int(pos) - a, b - a, overlap, i == len(chunks) - 1)]
            assert len(owners) == 1, (n, chunk, overlap, int(pos), chunks)
            a, b = owners[0]
            assert min(int(pos) + overlap, n) <= b                      # a frame of up to `overlap` samples lies inside its owner


def test_plan_chunks_rejects_chunk_not_longer_than_a_frame():
    from ofdm_b200.ingest import plan_chunks
    with pytest.raises(ValueError):
        plan_chunks(1000, 100, 100)


def test_read_fc32_is_bytes_to_sig_sliced(tmp_path, oo):
    # bytes_to_sig (src/utils.rs:238-254) + [start..stop] (examples/lab3c.rs:57-74)
    from ofdm_b200.ingest import read_fc32, fc32_file_samples
    rng = np.random.default_rng(2)
    sig = (rng.standard_normal(1000) + 1j * rng.standard_normal(1000))
    p = tmp_path / "cap.dat"
    p.write_bytes(oo.sig_to_fc32(sig).tobytes())
    assert fc32_file_samples(str(p)) == 1000
    want = oo.fc32_to_sig(np.frombuffer(p.read_bytes(), np.float32))
    np.testing.assert_array_equal(np.asarray(read_fc32(str(p))).astype(np.complex128), want)
    np.testing.assert_array_equal(np.asarray(read_fc32(str(p), 10, 500)).astype(np.complex128), want[10:500])
    assert read_fc32(str(p), 400, 5000).size == 600 and read_fc32(str(p), 2000).size == 0


# ---- GPU --------------------------------------------------------------------------------------------------------------
def _capture(oo, rng, n, ocfg, max_pay):
    cap = (0.0003 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    sent, p = [], 1234
    while True:
        pay = rng.integers(0, 256, int(rng.integers(1, max_pay)), dtype=np.uint8).tobytes()
        ch = oo.channel(oo.tx(pay, ocfg), 38.0, float(rng.uniform(0, 0.03)), 1, p).astype(np.complex64)
        if p + ch.size + 100 > n:
            break
        cap[p: p + ch.size] += ch
        sent.append((p + 8, pay))                                       # main channel tap at delay 9, lag - 1 rule
        p += ch.size + int(rng.integers(900, 20_000))
    return cap, sent


@pytest.mark.gpu
def test_decode_file_chunked_equals_whole_capture(tmp_path, oo):
    import ofdm_b200 as ob
    from ofdm_b200 import ingest
    cfg = ob.Config(modulation=2, guard_bands=True, fec=True, cfo_mode=1, phase_mode=1)
    ocfg = oo.make_cfg(True, 2, True, oo.SYNC_REFERENCE, 1, 1)
    rng = np.random.default_rng(5)
    cap, sent = _capture(oo, rng, 900_000, ocfg, 6000)
    assert len(sent) > 25
    path = tmp_path / "rx.dat"
    path.write_bytes(ob.sig_to_bytes(cap))
    max_frame = 30_000                                                  # > the longest frame (6000 B -> ~23.6 k samples + channel tail)
    # many small chunks: frames straddle chunk boundaries and are found again in the overlap
    frames = ingest.decode_file(str(path), cfg, chunk_samples=100_000, max_frame_samples=max_frame, out_stride=6016)
    assert [f.offset for f in frames] == [p for p, _ in sent]
    assert all(f.status == 0 for f in frames) and [f.data for f in frames] == [pay for _, pay in sent]
    # one big chunk gives the same list
    whole = ingest.decode_file(str(path), cfg, chunk_samples=1 << 20, max_frame_samples=max_frame, out_stride=6016)
    assert [(f.offset, f.data) for f in whole] == [(f.offset, f.data) for f in frames]
    np.testing.assert_allclose([f.f_delta for f in whole], [f.f_delta for f in frames], atol=1e-6)
    # [start..stop] slicing like lab3c --start/--stop: offsets are relative to start; frames cut by the slice are dropped
    start, stop = sent[3][0] - 500, sent[9][0] + 300
    part = ingest.decode_file(str(path), cfg, start=start, stop=stop, chunk_samples=100_000, max_frame_samples=max_frame, out_stride=6016)
    good = [f for f in part if f.status == 0]
    assert [(f.offset + start, f.data) for f in good] == [(p, pay) for p, pay in sent[3:9]]


@pytest.mark.gpu
def test_stream_receiver_push_in_odd_blocks(oo):
    # the radio loop of examples/jetson_rx.rs:46-57: samples arrive in blocks unrelated to frames or chunks
    import ofdm_b200 as ob
    from ofdm_b200 import ingest
    cfg = ob.Config(modulation=1, guard_bands=True, fec=False, cfo_mode=1, phase_mode=1)
    ocfg = oo.make_cfg(True, 1, False, oo.SYNC_REFERENCE, 1, 1)
    rng = np.random.default_rng(6)
    cap, sent = _capture(oo, rng, 500_000, ocfg, 1500)
    rx = ingest.StreamReceiver(cfg, chunk_samples=65_536, max_frame_samples=24_000, out_stride=2048)
    frames, pos = [], 0
    while pos < cap.size:
        k = int(rng.integers(1, 40_000))
        frames += rx.push(cap[pos: pos + k].astype(np.complex128))      # Complex64 in, cast like sig_to_bytes
        pos += k
    frames += rx.flush()
    rx.close()
    assert rx.samples_in == cap.size
    assert [(f.offset, f.status, f.data) for f in frames] == [(p, 0, pay) for p, pay in sent]


@pytest.mark.gpu
def test_c_abi_decode_file_equals_python_ingest(tmp_path, oo):
    """ofdm_rx_decode_file (the C entry a Rust caller binds) against the Python ingest and the transmitted payloads: chunked,
    whole, and [start..stop]-sliced (examples/lab3c.rs:57-74)."""
    import ofdm_b200 as ob
    cfg = ob.Config(modulation=2, guard_bands=True, fec=True, cfo_mode=1, phase_mode=1)
    ocfg = oo.make_cfg(True, 2, True, oo.SYNC_REFERENCE, 1, 1)
    rng = np.random.default_rng(15)
    cap, sent = _capture(oo, rng, 700_000, ocfg, 5000)
    assert len(sent) > 20
    path = tmp_path / "rx.dat"
    path.write_bytes(ob.sig_to_bytes(cap))
    eng = ob.Engine(cfg, 0)
    for chunk in (90_000, 1 << 20):
        rec, data = eng.decode_file(str(path), chunk_samples=chunk, max_frame_samples=30_000, out_stride=5008, max_frames=256)
        assert [int(x) for x in rec["offset"]] == [p for p, _ in sent] and (rec["status"] == 0).all()
        assert data == [pay for _, pay in sent] and [int(x) for x in rec["out_len"]] == [len(pay) for _, pay in sent]
        assert (rec["metric"] > 0.5).all()
    start, stop = sent[2][0] - 700, sent[7][0] + 200
    rec, data = eng.decode_file(str(path), start=start, stop=stop, chunk_samples=90_000, max_frame_samples=30_000, out_stride=5008, max_frames=64)
    good = rec["status"] == 0
    assert [(int(o) + start, d) for o, d, g in zip(rec["offset"], data, good) if g] == [(p, pay) for p, pay in sent[2:7]]
    with pytest.raises(ob.EngineError, match="max_frames"):
        eng.decode_file(str(path), chunk_samples=90_000, max_frame_samples=30_000, out_stride=5008, max_frames=5)
    with pytest.raises(ob.EngineError, match="cannot open"):
        eng.decode_file(str(tmp_path / "missing.dat"))
    eng.close()
