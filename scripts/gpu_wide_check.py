"""nfft = 1024 parity vs the oracle over the mode matrix (development aid)."""
import itertools, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ofdm_b200 as ob
from oracle import oracle as oo
rng = np.random.default_rng(0)
fails = 0
for mod, guard, fec in itertools.product((0, 1, 2), (False, True), (False, True)):
    for sync, cfo, phase in ((0, 0, 0), (1, 1, 1)):
        win = 0 if sync == 0 else 4096
        cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec, sync_mode=sync, cfo_mode=cfo, phase_mode=phase, sync_window=win, nfft=1024, cp=256)
        eng = ob.Engine(cfg, 0)
        ocfg = oo.make_cfg(guard, mod, fec, sync, cfo, phase, win, nfft=1024)
        pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (3000, 0, 577, 20000 if mod == 2 else 2500)]
        iq, flen = eng.tx_encode(pays)
        txerr = 0.0
        caps = []
        for i, p in enumerate(pays):
            ref = oo.tx(p, ocfg)
            assert ref.size == flen[i], (ref.size, flen[i])
            txerr = max(txerr, np.abs(iq[i, : flen[i]] - ref).max())
            lead = int(rng.integers(0, 700))
            c = oo.channel(ref, 60.0, 0.0015 + 0.0002 * i, 1, 100 + i)
            caps.append(np.concatenate([1e-4 * (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)), c]))
        n = np.array([c.size for c in caps], np.uint32)
        batch = np.zeros((len(caps), n.max()), np.complex64)
        for i, c in enumerate(caps):
            batch[i, : c.size] = c
        res = eng.rx_decode(batch, n, diag=True, points=True)
        ok, msg, perr = True, "", 0.0
        for i, p in enumerate(pays):
            ref = oo.decode(batch[i, : n[i]].astype(np.complex128), ocfg)
            if ref.status != res.status[i] or ref.offset != res.offset[i]:
                ok = False; msg += f" [s{i} status {ref.status}/{res.status[i]} off {ref.offset}/{res.offset[i]}]"; continue
            if ref.status != 0: continue
            if abs(ref.f_delta - res.f_delta[i]) > 1e-6: ok = False; msg += f" [s{i} f {ref.f_delta} {res.f_delta[i]}]"
            herr = np.abs(ref.h_k - res.h_k[i]).max() / np.abs(ref.h_k).max()
            if herr > 1e-4: ok = False; msg += f" [s{i} h {herr:.2e}]"
            npnt = cfg.frame_data_syms(len(p)) * cfg.data_carriers
            pe = np.abs(ref.points[:npnt] - res.points[i, :npnt]).max() / max(1.0, np.abs(ref.points[:npnt]).max())
            perr = max(perr, pe)
            if pe > 1e-4: ok = False; msg += f" [s{i} pts {pe:.2e}]"
            if res.data[i] != ref.data.tobytes(): ok = False; msg += f" [s{i} bytes differ]"
            if ref.data.tobytes() != p: msg += f" [s{i} oracle!=payload]"
        print(f"mod={mod} guard={int(guard)} fec={int(fec)} modes={sync}{cfo}{phase} tx_err={txerr:.2e} pts_err={perr:.2e} {'OK' if ok else 'FAIL'}{msg}", flush=True)
        fails += (not ok) or txerr > 1e-5
        eng.close()
print("FAILS", fails)
sys.exit(1 if fails else 0)
