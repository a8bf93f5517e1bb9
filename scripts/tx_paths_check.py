"""Same-box check of the TX paths: the one-pass tensor-memory-resident kernel against the two-pass kernel (bit-for-bit) on the
bench workload and on ragged batches, and the time of each. Usage: python scripts/tx_paths_check.py [n_streams] [syms]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ofdm_b200 as ob  # noqa: E402


def run(path, cfg, payload, plen, pstride, n, stride, steps):
    os.environ["OFDM_TX_PATH"] = path
    eng = ob.Engine(cfg, 0)
    dev = payload.device
    tx = torch.full((n, stride, 2), 7.0, dtype=torch.float32, device=dev)
    flen = torch.zeros(n, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        eng.tx_encode_device(payload.data_ptr(), plen.data_ptr(), pstride, n, tx.data_ptr(), stride, flen.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.tx_encode_device(payload.data_ptr(), plen.data_ptr(), pstride, n, tx.data_ptr(), stride, flen.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    eng.close()
    return tx, flen, ms


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 2038
    sys.argv = sys.argv[:1]
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    ok = True
    cases = [("bench", dict(), n, S, False), ("ragged", dict(), 300, 700, True), ("bpsk_noguard_nofec", dict(modulation=0, guard_bands=0, fec=0), 200, 300, True),
             ("qpsk_fec", dict(modulation=1, fec=1), 150, 1500, True), ("64qam_noguard", dict(guard_bands=0), 64, 5000, True)]
    if os.environ.get("TXCHECK_QUICK"):
        cases = cases[:1]
    if os.environ.get("TXCHECK_SWEEP"):      # where does the one-pass kernel overtake the two-pass one? frames per group = n / (148 / C)
        cases = [(f"sweep{k}", dict(), k, sy, False) for sy in (2038, 500) for k in (74, 148, 296, 444, 592, 888, 1184, 2368)]
    for name, over, ns, syms, ragged in cases:
        import dataclasses
        cfg = dataclasses.replace(bench.workload_cfg(), **over)
        plen_b = cfg.max_payload(syms)
        stride = cfg.frame_len(plen_b) + (0 if not ragged else 96)
        pstride = (plen_b + 15) // 16 * 16
        payload = torch.randint(0, 256, (ns, pstride), dtype=torch.uint8, device=dev, generator=g)
        if ragged:
            plen = torch.randint(0, plen_b + 1, (ns,), dtype=torch.int32, device=dev, generator=g)
            plen[0] = 0
            plen[1] = plen_b
        else:
            plen = torch.full((ns,), plen_b, dtype=torch.int32, device=dev)
        t0 = time.time()
        a, fa, ms_a = run("twopass", cfg, payload, plen, pstride, ns, stride, 10)
        b, fb, ms_b = run("resident", cfg, payload, plen, pstride, ns, stride, 10)
        same = bool(torch.equal(a, b)) and bool(torch.equal(fa, fb))
        nbad = int((a != b).sum().item())
        ok &= same
        print(f"{name}: n={ns} S={syms} stride={stride} twopass {ms_a:.4f} ms resident {ms_b:.4f} ms identical={same} diff_elems={nbad} ({time.time() - t0:.1f}s)", flush=True)
        if not same:
            d = (a != b).any(dim=2)
            rows = d.any(dim=1).nonzero().flatten()[:5].tolist()
            for r in rows:
                cols = d[r].nonzero().flatten()
                print("  stream", r, "plen", int(plen[r]), "first/last diff sample", int(cols[0]), int(cols[-1]), "count", len(cols), a[r, cols[0]].tolist(), b[r, cols[0]].tolist())
        del a, b
    print("TX_PATHS_OK" if ok else "TX_PATHS_MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
