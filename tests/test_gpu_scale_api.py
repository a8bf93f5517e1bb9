"""Full-size properties (BASELINE.json configs[1] shapes) and the host API mirror, on the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_full_size_round_trip_4096_streams():
    """4096 streams x S=2038 64QAM+Hamming frames (5.4 GB of IQ): TX -> channel (40 dB) -> RX recovers every payload.
    Size-independent property: encode -> channel -> decode round trip, checked with the BER counters on the device."""
    import torch
    import ofdm_b200 as ob
    cfg = ob.Config(modulation=ob.MOD_QAM64, guard_bands=True, fec=True, sync_mode=ob.SYNC_SCHMIDL_COX,
                    cfo_mode=ob.CFO_ANGLE_OF_SUM, phase_mode=ob.PHASE_ANGLE_OF_SUM, sync_window=2048)
    eng = ob.Engine(cfg, 0)
    n, S = 4096, 2038
    plen = cfg.max_payload(S)
    flen_want = cfg.frame_len(plen)
    assert flen_want == 163840
    stride = (flen_want + 1031 + 63 + 31) // 32 * 32
    ostride = (plen + 15) // 16 * 16
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    st = torch.cuda.current_stream().cuda_stream
    pay = torch.randint(0, 256, (n, ostride), dtype=torch.uint8, device=dev, generator=g)
    pl = torch.full((n,), plen, dtype=torch.int32, device=dev)
    tx = torch.empty((n, flen_want, 2), dtype=torch.float32, device=dev)
    fl = torch.zeros(n, dtype=torch.int32, device=dev)
    eng.tx_encode_device(pay.data_ptr(), pl.data_ptr(), ostride, n, tx.data_ptr(), flen_want, fl.data_ptr(), st)
    rx = torch.empty((n, stride, 2), dtype=torch.float32, device=dev)
    rl = torch.zeros(n, dtype=torch.int32, device=dev)
    lead = torch.zeros(n, dtype=torch.int32, device=dev)
    prm = ob.ChannelParams(snr_db=40.0, cfo_max=0.9 * np.pi / 80, lead_min=8, lead_max=1031, noise_mode=1, seed=2)
    eng.channel_device(tx.data_ptr(), fl.data_ptr(), flen_want, n, prm, rx.data_ptr(), stride, rl.data_ptr(), lead.data_ptr(), 0, st)
    torch.cuda.synchronize()
    assert (fl == flen_want).all()
    del tx
    out = torch.zeros((n, ostride), dtype=torch.uint8, device=dev)
    ol = torch.zeros(n, dtype=torch.int32, device=dev)
    stt = torch.zeros(n, dtype=torch.int32, device=dev)
    off = torch.zeros(n, dtype=torch.int32, device=dev)
    from ofdm_b200.engine import CRxDiag
    diag = CRxDiag(off.data_ptr(), None, None, None, None, 0)
    eng.rx_decode_device(rx.data_ptr(), rl.data_ptr(), n, stride, int(rl.max()), out.data_ptr(), ostride, ol.data_ptr(), stt.data_ptr(), st, diag)
    counters = torch.zeros(4, dtype=torch.int64, device=dev)
    eng.ber_device(pay.data_ptr(), pl.data_ptr(), ostride, out.data_ptr(), ol.data_ptr(), ostride, stt.data_ptr(), n, counters.data_ptr(), st)
    torch.cuda.synchronize()
    assert (stt == 0).all() and (ol == plen).all()
    assert (off == lead + 8).all()                       # offset = lag - 1 with the main tap at delay 9
    c = counters.tolist()
    assert c == [0, 0, 8 * plen * n, 0], c
    assert (out[:, :plen] == pay[:, :plen]).all()
    # idempotence: decoding the same capture again gives the same bytes
    out2 = torch.zeros_like(out)
    eng.rx_decode_device(rx.data_ptr(), rl.data_ptr(), n, stride, int(rl.max()), out2.data_ptr(), ostride, ol.data_ptr(), stt.data_ptr(), st)
    torch.cuda.synchronize()
    assert (out2 == out).all()
    # spot parity against the CPU oracle on 3 of the full-size streams
    from oracle import oracle as oo
    ocfg = oo.make_cfg(True, oo.QAM64, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, 2048)
    for i in (0, 1777, 4095):
        cap = rx[i, : int(rl[i])].cpu().numpy().view(np.complex64).reshape(-1).astype(np.complex128)
        ref = oo.decode(cap, ocfg, want_points=False, out_cap=ostride)
        assert ref.status == 0 and ref.offset == int(off[i])
        assert ref.data.tobytes() == out[i, :plen].cpu().numpy().tobytes()
    eng.close()


def test_api_mirror_reads_like_the_reference():
    """src/lib.rs:37-57 style: encode!/decode! with optional arguments, anyhow-like error for short input."""
    import ofdm_b200 as ob
    from oracle import oracle as oo
    data = b"alskdjas"
    frame = ob.encode(data, True, None)                                    # encoding_works (src/lib.rs:53-57)
    assert frame.dtype == np.complex128 and frame.size == 14 * 80
    np.testing.assert_allclose(frame, oo.encode(data, True, oo.BPSK), atol=2e-6)
    for scheme in (ob.ModulationScheme.Bpsk, ob.ModulationScheme.Qpsk, ob.ModulationScheme.Qam):
        tx = ob.encode(data, True, scheme)
        rx = oo.channel(tx, 30.0, 0.02, 0, 5)                              # channel!(tx, snr: 30.0, timing_error)
        assert ob.decode(rx, True, scheme) == data
    with pytest.raises(ob.DecodeError, match="Input not long enough"):     # src/receiver.rs:27-29
        ob.decode(np.concatenate([np.zeros(4), frame[:600]]), True, None)
    with pytest.raises(ob.DecodeError):                                    # src/receiver.rs:25 panics
        ob.decode(frame, True, None)
    a = ob.Analysis.new(bytes([1, 0, 0, 0]), bytes([1, 0, 1, 0]))          # src/utils.rs:45-68
    assert (a.num_errs, a.num_block_errs, a.err_rate) == (1, 1, 1 / 32)
    # lab3c-style file boundary: sig_to_bytes -> bytes_to_sig -> decode (examples/lab3c_image.rs:15-42)
    tx = ob.encode(bytes(range(200)), True, ob.ModulationScheme.Qpsk)
    wire = ob.sig_to_bytes(oo.channel(tx, 30.0, 0.01, 0, 6))
    assert ob.decode(ob.bytes_to_sig(wire), True, ob.ModulationScheme.Qpsk) == bytes(range(200))
    outs, status = ob.decode_batch([ob.bytes_to_sig(wire), np.zeros(10)], True, ob.ModulationScheme.Qpsk)
    assert outs[0] == bytes(range(200)) and status[1] != 0


def test_config1_dancing_bytes_loopback(golden):
    """BASELINE.json configs[0]: 64QAM loopback of support/dancing.bytes through the lab channel (lab3c-style)."""
    import ofdm_b200 as ob
    from oracle import oracle as oo
    pay = golden["qam64_guard_fec_sc.payload"].tobytes()
    assert len(pay) == 576
    kw = dict(fec=True, sync_mode=ob.SYNC_SCHMIDL_COX, cfo_mode=ob.CFO_ANGLE_OF_SUM, phase_mode=ob.PHASE_ANGLE_OF_SUM, sync_window=1024)
    tx = ob.encode(pay, True, ob.ModulationScheme.Qam, **kw)
    assert tx.size == 3120
    for snr in (30, 35, 40):
        for cfo in (0.0, 0.01, 0.02, 0.035):
            cap = oo.channel(tx, snr, cfo, 0, 0xD0FD0001 + int(cfo * 1000))
            ocfg = oo.make_cfg(True, oo.QAM64, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, 1024)
            ref = oo.decode(cap.astype(np.complex64).astype(np.complex128), ocfg, want_points=False)
            try:
                got = ob.decode(cap, True, ob.ModulationScheme.Qam, **kw)
                st = 0
            except ob.DecodeError as e:
                got, st = b"", e.status
            assert st == ref.status
            if ref.status == 0 and oo.analysis(pay, ref.data)[0] == 0:          # where the oracle decodes error-free
                assert got == pay


def test_cpp_host_lab3c_file_roundtrip(tmp_path):
    """The C++ host mirror (ofdm_b200/host): lab3c --transmit -> fc32 file -> channel -> lab3c --receive --start/--stop."""
    import subprocess
    from ofdm_b200 import _build
    from oracle import oracle as oo
    import ofdm_b200 as ob
    exe = _build.build_host_example()
    pay = bytes(range(256)) * 2 + b"lab3c"
    (tmp_path / "in.bin").write_bytes(pay)
    subprocess.run([exe, "--transmit", str(tmp_path / "tx.dat"), "--payload", str(tmp_path / "in.bin"), "--qpsk", "--guard"], check=True)
    tx = ob.bytes_to_sig((tmp_path / "tx.dat").read_bytes())
    np.testing.assert_allclose(tx, oo.encode(pay, True, oo.QPSK), atol=2e-6)
    cap = np.concatenate([np.zeros(100), oo.channel(tx, 30.0, 0.02, 0, 3), np.zeros(50)])
    (tmp_path / "rx.dat").write_bytes(ob.sig_to_bytes(cap))
    subprocess.run([exe, "--receive", str(tmp_path / "rx.dat"), "--out", str(tmp_path / "out.bin"), "--start", "100", "--stop", str(100 + tx.size + 63),
                    "--qpsk", "--guard"], check=True)
    assert (tmp_path / "out.bin").read_bytes() == pay
    r = subprocess.run([exe, "--receive", str(tmp_path / "rx.dat"), "--out", str(tmp_path / "o2.bin"), "--stop", "500", "--qpsk", "--guard"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "Input not long enough" in r.stderr


def test_cpp_host_lab3c_ecc_roundtrip(tmp_path):
    """lab3c --ecc: RS(255,223) outer code around the modem like examples/lab3c_image.rs:19-21,33-36, at an SNR where BPSK
    alone leaves byte errors."""
    import subprocess
    from ofdm_b200 import _build
    from oracle import oracle as oo
    import ofdm_b200 as ob
    exe = _build.build_host_example()
    pay = bytes((7 * i + 3) & 255 for i in range(576))
    (tmp_path / "in.bin").write_bytes(pay)
    subprocess.run([exe, "--transmit", str(tmp_path / "tx.dat"), "--payload", str(tmp_path / "in.bin"), "--guard", "--ecc"], check=True)
    tx = ob.bytes_to_sig((tmp_path / "tx.dat").read_bytes())
    assert tx.size == 11280                                            # 765 coded bytes, BPSK
    np.testing.assert_allclose(tx, oo.encode(oo.rs_encode(np.frombuffer(pay, np.uint8)).tobytes(), True, oo.BPSK), atol=2e-6)
    done = 0
    for seed in range(1, 9):
        cap = oo.channel(tx, 8.0, 0.01, 0, seed)
        ref = oo.decode(cap.astype(np.complex64).astype(np.complex128), oo.make_cfg(True, oo.BPSK, False, 0, 0, 0, 0), want_points=False)
        if ref.status != 0 or len(ref.data) != 765:
            continue                                                   # the unprotected header took a hit
        want, _, nf = oo.rs_decode(np.frombuffer(ref.data, np.uint8))
        (tmp_path / "rx.dat").write_bytes(ob.sig_to_bytes(cap))
        r = subprocess.run([exe, "--receive", str(tmp_path / "rx.dat"), "--out", str(tmp_path / "out.bin"), "--guard", "--ecc"],
                           capture_output=True, text=True)
        if nf:
            assert r.returncode == 1 and "beyond repair" in r.stderr
        else:
            assert r.returncode == 0, r.stderr
            got = (tmp_path / "out.bin").read_bytes()
            assert got == want.tobytes() and got[:576] == pay
            done += 1
    assert done >= 3
