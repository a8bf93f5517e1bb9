import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ofdm_b200 as ob
from oracle import oracle as oo
mod, guard, fec = 2, False, True
cfg = ob.Config(modulation=mod, guard_bands=guard, fec=fec, sync_mode=1, cfo_mode=1, phase_mode=1, sync_window=2048)
eng = ob.Engine(cfg, 0)
ocfg = oo.make_cfg(guard, mod, fec, 1, 1, 1, 2048)
rng = np.random.default_rng(mod)
S = 700
lens = [cfg.max_payload(S), cfg.max_payload(S) - 1]
pays = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in lens]
caps = []
for i, p in enumerate(pays):
    lead = int(rng.integers(0, 900))
    c = oo.channel(oo.tx(p, ocfg), 45.0, 0.025, 1, i)
    caps.append(np.concatenate([0.002 * (rng.standard_normal(lead) + 1j * rng.standard_normal(lead)), c]))
n = np.array([c.size for c in caps], np.uint32)
iq = np.zeros((2, n.max()), np.complex64)
for i, c in enumerate(caps): iq[i, :c.size] = c
res = eng.rx_decode(iq, n, points=True)
for i, p in enumerate(pays):
    ref = oo.decode(iq[i, :n[i]].astype(np.complex128), ocfg)
    g = np.frombuffer(res.data[i], np.uint8); r = ref.data
    print(i, "status", res.status[i], ref.status, "len", g.size, r.size, len(p), "ref==p", r.tobytes() == p, "f", res.f_delta[i], ref.f_delta)
    d = np.flatnonzero(g != r[:g.size])
    print("  n diff bytes", d.size, d[:10], [(hex(g[k]), hex(r[k])) for k in d[:5]])
    npts = S * 64
    e = np.abs(res.points[i, :npts] - ref.points[:npts])
    print("  points max err", e.max(), "at", e.argmax(), "sym", e.argmax() // 64, "bin", e.argmax() % 64)
    worst = np.argsort(-e)[:5]
    print("  worst", [(int(k) // 64, int(k) % 64, float(e[k]), complex(ref.points[k])) for k in worst])
    # per-symbol mean error
    es = e.reshape(S, 64).max(axis=1)
    print("  per-symbol err: first", es[:3], "mid", es[350:353], "last", es[-3:])
    # margin of ref points to decision boundary
    fr = np.abs((3.5 * ref.points[:npts].real + 4) - np.round(3.5 * ref.points[:npts].real + 4)) / 3.5
    fi = np.abs((3.5 * ref.points[:npts].imag + 4) - np.round(3.5 * ref.points[:npts].imag + 4)) / 3.5
    print("  min margin", min(fr.min(), fi.min()))
