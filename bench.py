#!/usr/bin/env python
"""bench.py -- RX hot-path throughput on B200 (BASELINE.json metric: RX Msamples/s & decoded Gbit/s, % HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload "4096x64QAM_S2038", BASELINE.json configs[1]): per GPU 4096 independent streams, each one a
capture holding a 64QAM + guard-band + Hamming(7,4) frame of S=2038 data symbols (163 840 samples) behind a random
noise-only lead-in, through the 12-tap multipath + CFO + AWGN channel at 40 dB (every frame decodes; the 30 dB operating
point of SURVEY.md 8(d), where frames with a corrupted header are dropped, is reported beside it as secondary.snr30);
sliding Schmidl-Cox sync + the full RX chain. The IQ is synthesised ON THE DEVICE by the engine's own TX + channel kernels
before the timed region. A "step" = one ofdm_rx_decode_batch over all streams of the rank. Weak scaling: every rank owns
its own 4096 streams, no data-path collective; the BER counters are sum-reduced once over NCCL after the timed region
(--scaling strong: 4096 streams in total, split evenly over the ranks).

  value : whole-job Msamples/s, inputs resident in HBM, CUDA events on the launch stream, max over ranks.
  e2e   : same metric through the C ABI with HOST (pinned) buffers: H2D of the IQ and D2H of the payload inside;
          e2e.h2d_ceiling_gb_per_s = a bare cudaMemcpyAsync loop over the same pinned buffer, all ranks at once.
  roofline : decode kernel, algorithmic bytes (8 B/sample in + payload bytes out) / its CUDA-event duration.
  cpu_baseline : the CPU oracle (C port of the reference algorithm; the Rust reference cannot be built here).
  secondary (N=1, after the headline's timed region): the other BASELINE.json configs on the same box --
          snr30 (config 2 at 30 dB), capture (config 3), wide (config 4, nfft 1024), capture_wide (config 3 for the nfft 1024
          layout), tx, rs (the RS(255,223) outer code) -- each with value, kernel_ms, roofline and its checks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def host_threads() -> int:
    """All the host threads this process may use (torchrun pins OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--syms", type=int, default=2038, help="data OFDM symbols per frame")
    ap.add_argument("--snr", type=float, default=40.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the snr30 / wide / capture / capture_wide / tx / rs legs after the headline")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --streams per GPU; strong: --streams in total, split evenly over the ranks")
    ap.add_argument("--workload", default="streams", choices=["streams", "capture", "rs", "ingest", "tx"],
                    help="streams: BASELINE configs[1] (the headline line); capture: configs[2], preamble search over one long capture")
    ap.add_argument("--capture-samples", type=float, default=1e9)
    ap.add_argument("--nfft", type=int, default=64, choices=[64, 1024],
                    help="64: the reference layout (headline); 1024: wideband variant, BASELINE configs[3] (use --syms 128)")
    return ap.parse_args()


LEAD_MIN, LEAD_MAX = 8, 1031
SYNC_WINDOW = 2048
SEED = 0x0FD64


NFFT = 64


def sync_window():
    return SYNC_WINDOW if NFFT == 64 else 4096


def workload_cfg():
    import ofdm_b200 as ob
    return ob.Config(modulation=ob.MOD_QAM64, guard_bands=True, fec=True, sync_mode=ob.SYNC_SCHMIDL_COX,
                     cfo_mode=ob.CFO_ANGLE_OF_SUM, phase_mode=ob.PHASE_ANGLE_OF_SUM, sync_window=sync_window(), nfft=NFFT, cp=NFFT // 4)


def oracle_cfg():
    from oracle import oracle as oo
    return oo.make_cfg(True, oo.QAM64, True, oo.SYNC_SCHMIDL_COX, oo.CFO_ANGLE_OF_SUM, oo.PHASE_ANGLE_OF_SUM, sync_window(), nfft=NFFT)


def bind_host_to_gpu_node(gpu_index: int):
    """Run this process on the CPUs NVML reports as local to the GPU, so that the pinned host buffers of the e2e leg are
    first-touched on the GPU's own NUMA node (with 8 ranks on one host the PCIe copies otherwise cross the socket link).
    Returns the number of CPUs bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = gpu_index
        if vis:
            try:
                idx = int(vis.split(",")[gpu_index])
            except Exception:           # noqa: BLE001
                idx = gpu_index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:                   # noqa: BLE001
        return None


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every ~5 ms DURING the timed region (the recipe's clocks line)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.sm, self.pw, self.reasons = [], [], set()
        self.sm_max = None
        self._stop = threading.Event()
        self.t = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv = nv
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:           # noqa: BLE001
            self.err = repr(e)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception as e:       # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.005)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "error": self.err}
        self._stop.set()
        self.t.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                "power_w_max": max(self.pw) if self.pw else None, "samples": len(self.sm), "reasons": sorted(self.reasons)}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def read_traffic(workload: str):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        e = t.get(workload)
        if e:
            return e.get("dram_bytes_per_launch")
        if workload.startswith("capture_") and "capture" in t:               # capture_<n samples>: per-sample figure x n
            return int(t["capture"]["dram_bytes_per_sample"] * int(workload.split("_")[1]))
    return None


def cpu_sample(iq_host: np.ndarray, n_samples: np.ndarray, out_stride: int, target_s: float, threads: int):
    """Time the oracle on a bounded sample of the same workload: the k streams are decoded repeatedly until about
    `target_s` seconds of wall time are spent. Returns (msamples_per_s, k, seconds, reps, last results)."""
    from oracle import oracle as oo
    cfg = oracle_cfg()
    k = iq_host.shape[0]
    f32 = iq_host.view(np.float32).reshape(k, iq_host.shape[1], 2)
    res = oo.decode_batch_fc32(f32, n_samples, cfg, out_stride, threads)          # warm-up (FFT plan, page faults)
    reps, t0 = 0, time.perf_counter()
    while True:
        res = oo.decode_batch_fc32(f32, n_samples, cfg, out_stride, threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= target_s:
            break
    return float(n_samples.sum()) * reps / dt / 1e6, k, dt, reps, res[:3]


def headline_sizes(args, nfft):
    """Frame / stride arithmetic of the streams workload, shared by both arms (same numbers as ofdm_max_payload etc.)."""
    S = args.syms
    bits_per_sym = 288 if nfft == 64 else 4608                    # 64QAM x 48 (768) data carriers
    coded = (S * bits_per_sym - 128) // 8
    payload_len = (8 * coded) // 14                               # Hamming(7,4): 14 coded bits per payload byte
    frame_len = (10 + S) * (nfft + nfft // 4)
    iq_stride = (frame_len + LEAD_MAX + 63 + 31) // 32 * 32
    out_stride = (payload_len + 15) // 16 * 16
    return payload_len, frame_len, iq_stride, out_stride


def headline_config(args, nfft, n_streams, snr):
    """The `config` object of the JSON line: one function for both arms, so their key sets and values are identical."""
    payload_len, frame_len, iq_stride, _ = headline_sizes(args, nfft)
    return {"workload": f"{args.streams}x64QAM_S{args.syms}" + ("" if nfft == 64 else f"_N{nfft}"), "streams_per_gpu": n_streams,
            "data_syms_per_frame": args.syms, "frame_samples": frame_len, "payload_bytes": payload_len, "modulation": "64QAM",
            "guard_bands": True, "fec": "hamming(7,4)", "sync": "schmidl_cox(window=%d)" % (SYNC_WINDOW if nfft == 64 else 4096),
            "cfo": "angle_of_sum", "snr_db": snr, "nfft": nfft, "lead_in": [LEAD_MIN, LEAD_MAX],
            "l2": "inputs (%.2f GB/GPU) larger than L2" % (n_streams * iq_stride * 8 / 1e9)}


class StreamsWorkload:
    """BASELINE.json configs[1] (nfft 64) / configs[3] (nfft 1024) on one rank: synthesised on the device, untimed."""

    def __init__(self, args, nfft, snr, n_streams, rank, local_rank):
        import torch
        import ofdm_b200 as ob
        global NFFT
        NFFT = nfft
        self.torch, self.ob, self.nfft, self.snr, self.n = torch, ob, nfft, snr, n_streams
        self.cfg = workload_cfg()
        self.payload_len, self.frame_len, self.iq_stride, self.out_stride = headline_sizes(args, nfft)
        assert self.cfg.max_payload(args.syms) == self.payload_len and self.cfg.frame_len(self.payload_len) == self.frame_len
        self.dev = torch.device("cuda", local_rank)
        self.eng = ob.Engine(self.cfg, local_rank)
        dev, eng, n = self.dev, self.eng, n_streams
        g = torch.Generator(device=dev)
        g.manual_seed(SEED + rank)
        self.stream = torch.cuda.current_stream().cuda_stream
        self.payload = torch.randint(0, 256, (n, self.out_stride), dtype=torch.uint8, device=dev, generator=g)
        self.plen = torch.full((n,), self.payload_len, dtype=torch.int32, device=dev)
        tx = torch.empty((n, self.frame_len, 2), dtype=torch.float32, device=dev)
        flen = torch.zeros(n, dtype=torch.int32, device=dev)
        eng.tx_encode_device(self.payload.data_ptr(), self.plen.data_ptr(), self.out_stride, n, tx.data_ptr(), self.frame_len, flen.data_ptr(), self.stream)
        self.rx = torch.empty((n, self.iq_stride, 2), dtype=torch.float32, device=dev)
        self.rx_len = torch.zeros(n, dtype=torch.int32, device=dev)
        chan = ob.ChannelParams(snr_db=snr, cfo_max=0.9 * np.pi / (nfft + nfft // 4), lead_min=LEAD_MIN, lead_max=LEAD_MAX, multipath=True,
                                noise_mode=1, seed=SEED + 1000 * rank)
        eng.channel_device(tx.data_ptr(), flen.data_ptr(), self.frame_len, n, chan, self.rx.data_ptr(), self.iq_stride, self.rx_len.data_ptr(), 0, 0, self.stream)
        torch.cuda.synchronize()
        del tx
        torch.cuda.empty_cache()
        self.total_samples = int(self.rx_len.sum().item())
        self.max_n = int(self.rx_len.max().item())
        self.out = torch.zeros((n, self.out_stride), dtype=torch.uint8, device=dev)
        self.out_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.status = torch.zeros(n, dtype=torch.int32, device=dev)
        eng.reserve(n)

    def step(self):
        self.eng.rx_decode_device(self.rx.data_ptr(), self.rx_len.data_ptr(), self.n, self.iq_stride, self.max_n, self.out.data_ptr(),
                                  self.out_stride, self.out_len.data_ptr(), self.status.data_ptr(), self.stream)

    def timed(self, steps, warmup, barrier, local_rank):
        """W warm-up steps, then exactly K steps between CUDA events on the launch stream. Returns a dict of raw timings."""
        torch = self.torch
        for _ in range(max(warmup, 3)):
            self.step()
        barrier()
        launches0 = self.eng.kernel_launches
        self.eng.profile_begin(steps)
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            self.step()
        ev1.record()
        barrier()
        clocks = sampler.stop()
        ms_total = ev0.elapsed_time(ev1)
        acq_ms, dec_ms = self.eng.profile_read(steps)
        return {"ms_total": ms_total, "launches": self.eng.kernel_launches - launches0, "acq_ms": acq_ms, "dec_ms": dec_ms, "clocks": clocks}

    def ber_counters(self):
        torch = self.torch
        counters = torch.zeros(4, dtype=torch.int64, device=self.dev)
        self.eng.ber_device(self.payload.data_ptr(), self.plen.data_ptr(), self.out_stride, self.out.data_ptr(), self.out_len.data_ptr(),
                            self.out_stride, self.status.data_ptr(), self.n, counters.data_ptr(), self.stream)
        return counters

    def roofline(self, t, steps):
        peak, peak_src = read_peaks()
        alg_bytes = 8 * self.total_samples + int((self.status == 0).sum().item()) * self.payload_len
        dec_avg_ms = float(np.mean(t["dec_ms"])) if len(t["dec_ms"]) else float("nan")
        achieved = alg_bytes / (dec_avg_ms * 1e-3) / 1e9
        wl = f"{self.n}x64QAM_S{(self.frame_len // (self.nfft + self.nfft // 4)) - 10}" + ("" if self.nfft == 64 else f"_N{self.nfft}")
        return {"bound": "hbm", "kernel": "rx_decode_kernel" if self.nfft == 64 else "wide_decode_kernel", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": read_traffic(wl), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": round(dec_avg_ms, 4),
                "acquire_kernel_ms": round(float(np.mean(t["acq_ms"])), 4) if len(t["acq_ms"]) else None,
                "kernel_share_of_step": round(dec_avg_ms / (t["ms_total"] / steps), 4)}

    def close(self):
        self.eng.close()
        del self.rx, self.out, self.payload
        self.torch.cuda.empty_cache()


def h2d_ceiling(torch, host_tensor, dev, barrier, reps=3):
    """A bare cudaMemcpyAsync loop over the e2e leg's pinned buffer (all ranks at the same time): what the PCIe link + host
    memory can feed this GPU at most. GB/s."""
    dst = torch.empty_like(host_tensor, device=dev)
    dst.copy_(host_tensor, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(host_tensor, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    del dst
    return host_tensor.numel() * host_tensor.element_size() / dt / 1e9


def main():
    global NFFT
    args = parse_args()
    NFFT = args.nfft
    if NFFT == 1024 and args.snr == 40.0:
        args.snr = 50.0        # data symbols sit ~12 dB below the frame head at N = 1024 (docs/SPEC.md 9): keep every header decodable
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return 0                                     # rank 0 alone runs the CPU arm
        return reference_arm(args)

    import torch
    import ofdm_b200 as ob

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; the engine has no CPU fallback"}))
        return 1
    if args.workload in ("rs", "tx", "ingest", "capture"):
        fn = {"rs": rs_bench, "tx": tx_bench, "ingest": ingest_bench, "capture": capture_bench}[args.workload]
        line = fn(args, rank, local_rank, world)
        if rank == 0 and line is not None:
            print(json.dumps(line))
        return 0

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(t, op):
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return t

    # weak scaling: every rank owns args.streams streams; strong: args.streams in total, split evenly
    n_streams = args.streams
    if args.scaling == "strong":
        from ofdm_b200 import dist as od
        lo, hi = od.stream_shard(args.streams, rank, world)
        n_streams = hi - lo
    wl = StreamsWorkload(args, NFFT, args.snr, n_streams, rank, local_rank)
    config = headline_config(args, NFFT, n_streams, args.snr)
    payload_len, out_stride, iq_stride = wl.payload_len, wl.out_stride, wl.iq_stride

    # ---- timed region: K steps, inputs resident in HBM -------------------------------------------------------
    t = wl.timed(args.steps, args.warmup, barrier, local_rank)
    ms_step = float(allreduce(torch.tensor([t["ms_total"]], dtype=torch.float64, device=dev), "MAX").item()) / args.steps

    # ---- correctness of what was timed: BER vs the transmitted payload, counters sum-reduced over NCCL ---------
    counters = allreduce(wl.ber_counters(), "SUM")              # the path's only collective (4 x u64)
    tot = allreduce(torch.tensor([wl.total_samples, n_streams], dtype=torch.int64, device=dev), "SUM")
    torch.cuda.synchronize()
    c = [int(x) for x in counters.tolist()]
    job_samples, job_streams = int(tot[0].item()), int(tot[1].item())
    value = job_samples / (ms_step * 1e-3) / 1e6
    decoded_gbit = job_streams * payload_len * 8 / (ms_step * 1e-3) / 1e9
    roofline = wl.roofline(t, args.steps)

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region ---------------------------
    e2e = None
    if not args.no_e2e:
        all_cpus = os.sched_getaffinity(0)
        local_cpus = bind_host_to_gpu_node(local_rank)
        rx_host = torch.empty((n_streams, iq_stride, 2), dtype=torch.float32, pin_memory=True)
        rx_host.copy_(wl.rx)
        n_host = wl.rx_len.cpu().numpy().astype(np.uint32)
        out_host = torch.zeros((n_streams, out_stride), dtype=torch.uint8, pin_memory=True)
        ol_host = np.zeros(n_streams, np.uint32)
        st_host = np.zeros(n_streams, np.int32)
        eng = wl.eng

        def e2e_step():
            eng._check(eng.lib.ofdm_rx_decode_batch(eng._h, rx_host.data_ptr(), n_host.ctypes.data, n_streams, iq_stride, wl.max_n,
                                                    out_host.data_ptr(), out_stride, ol_host.ctypes.data, st_host.ctypes.data,
                                                    None, ob.engine.MEM_HOST, None), "ofdm_rx_decode_batch(host)")
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        dt = float(allreduce(torch.tensor([dt], dtype=torch.float64, device=dev), "MAX").item())
        same = bool((out_host.to(dev) == wl.out).all().item())
        ceil = h2d_ceiling(torch, rx_host, dev, barrier)
        ceil = float(allreduce(torch.tensor([ceil], dtype=torch.float64, device=dev), "MIN").item())
        h2d_moved = int(eng.last_h2d_bytes)                    # what the engine really copied: the cyclic prefixes stay on the host
        h2d_rate = h2d_moved / dt / 1e9
        e2e = {"value": round(job_samples / dt / 1e6, 1), "unit": "Msamples/s", "h2d_bytes_per_step": int(h2d_moved + n_streams * 4),
               "host_capture_bytes_per_step": int(n_streams * iq_stride * 8),
               "h2d_note": ("whole capture copied (chunks of streams, double-buffered)" if h2d_moved >= n_streams * iq_stride * 8 else
                            "head region of every stream, then -- once the acquisition has located the frame -- only the useful nfft samples of "
                            "each wanted data symbol, pulled from the pinned capture by a gather kernel: the cyclic prefixes stay on the host"),
               "d2h_bytes_per_step": int(n_streams * out_stride + n_streams * 8), "ms_per_step": round(dt * 1e3, 3),
               "decoded_gbit_per_s": round(job_streams * payload_len * 8 / dt / 1e9, 2),
               "h2d_gb_per_s_per_gpu": round(h2d_rate, 1), "h2d_ceiling_gb_per_s": round(ceil, 1),
               "h2d_ceiling_note": "bare cudaMemcpyAsync of the same pinned buffer, all ranks concurrently, slowest rank",
               "fraction_of_h2d_ceiling": round(h2d_rate / ceil, 3), "steps": args.e2e_steps,
               "matches_device_path": same, "host_cpus_local_to_gpu": local_cpus}
        del rx_host, out_host
        os.sched_setaffinity(0, all_cpus)                      # the CPU baseline below uses every core again

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample of the same workload ------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        k = min(n_streams, 2 * threads)
        iq_h = wl.rx[:k].cpu().numpy().view(np.complex64).reshape(k, iq_stride)
        ns_h = wl.rx_len[:k].cpu().numpy().astype(np.uint32)
        v1, k1, dt1, r1, _ = cpu_sample(iq_h[:2], ns_h[:2], out_stride, args.cpu_seconds / 2, 1)
        vN, kN, dtN, rN, res = cpu_sample(iq_h, ns_h, out_stride, args.cpu_seconds / 2, threads)
        o_out, o_len, o_st = res
        gpu_out = wl.out[:kN].cpu().numpy()
        gpu_st = wl.status[:kN].cpu().numpy()
        differ = int((o_out[:, :payload_len] != gpu_out[:, :payload_len]).sum()) + int((o_st != gpu_st).sum())
        cpu_baseline = {"value": round(vN, 2), "unit": "Msamples/s", "cores": threads, "kind": "port",
                        "sample": f"{kN} of the {n_streams} streams x {rN} passes ({dtN:.1f} s on {threads} threads; f64 C port of the "
                                  "reference algorithm, FFT plans reused = upper bound on the Rust crate's speed)",
                        "single_core": {"value": round(v1, 2), "sample": f"{k1} streams x {r1} passes, {dt1:.1f} s on 1 thread"},
                        "bytes_differing_from_gpu_on_sample": differ}
    wl.close()

    # ---- the other BASELINE.json configs on the same box (N=1 only, after the headline's timed region) ----------
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary and NFFT == 64:
        secondary = {}
        sec_steps = max(3, min(args.steps, 20))
        for name, fn in (("snr30", lambda: snr30_leg(args, sec_steps, barrier, local_rank)),
                         ("wide", lambda: wide_leg(args, sec_steps, barrier, local_rank)),
                         ("capture", lambda: capture_bench(args, 0, local_rank, 1, steps=sec_steps)),
                         ("capture_wide", lambda: capture_wide_leg(args, sec_steps, local_rank)),
                         ("tx", lambda: tx_bench(args, 0, local_rank, 1, steps=sec_steps)),
                         ("tx_wide", lambda: tx_wide_leg(args, sec_steps, local_rank)),
                         ("rs", lambda: rs_bench(args, 0, local_rank, 1))):
            try:
                full = fn()
                secondary[name] = {k: full[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "config", "roofline", "clocks",
                                                        "gpu_launches", "ber", "all_offsets_exact", "max_cfo_abs_err", "peaks_found",
                                                        "frames_ok", "decoded_gbit_per_s", "credited", "each_frame_found_once",
                                                        "frames_match_oracle_on_sample", "exact_one_pass", "encode_ms", "decode_clean_ms",
                                                        "decode_8_errors_per_block_ms", "decode_2pct_blocks_with_errors_ms", "checks") if k in full}
            except Exception as e:       # noqa: BLE001 -- a secondary leg must never take the headline line down
                secondary[name] = {"error": repr(e)[:300]}
            NFFT = args.nfft
            torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": "rx_msamples_per_s", "value": round(value, 1), "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "decoded_gbit_per_s": round(decoded_gbit, 1), "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
                "gpu_launches": int(t["launches"]), "clocks": t["clocks"],
                "ber": {"bit_errs": c[0], "byte_errs": c[1], "bits_compared": c[2], "frames_failed": c[3]}, "secondary": secondary}
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def _streams_leg(args, nfft, snr, syms, steps, barrier, local_rank):
    """One more streams workload on this GPU (secondary legs): timed like the headline, kernel-only."""
    import copy
    a = copy.copy(args)
    a.syms = syms
    wl = StreamsWorkload(a, nfft, snr, args.streams, 0, local_rank)
    t = wl.timed(steps, args.warmup, barrier, local_rank)
    ms_step = t["ms_total"] / steps
    c = [int(x) for x in wl.ber_counters().tolist()]
    ok = wl.status == 0
    ok_samples = int(wl.rx_len[ok].sum().item())
    n_ok = int(ok.sum().item())
    line = {"metric": "rx_msamples_per_s", "unit": "Msamples/s", "steps": steps, "ms_per_step": round(ms_step, 4),
            "value": round(ok_samples / (ms_step * 1e-3) / 1e6, 1),
            "credited": f"samples of the {n_ok} of {wl.n} streams that decoded (status OK); all {wl.n} are searched and header-decoded",
            "decoded_gbit_per_s": round(n_ok * wl.payload_len * 8 / (ms_step * 1e-3) / 1e9, 1),
            "config": headline_config(a, nfft, wl.n, snr), "roofline": wl.roofline(t, steps), "clocks": t["clocks"],
            "gpu_launches": int(t["launches"]),
            "ber": {"bit_errs": c[0], "byte_errs": c[1], "bits_compared": c[2], "frames_failed": c[3]}}
    wl.close()
    return line


def snr30_leg(args, steps, barrier, local_rank):
    """SURVEY.md 8(d) config 2 at its stated 30 dB: 64QAM has a raw BER of ~1.7e-3 there, so ~17 % of the (unprotected,
    reference-design) 128-bit headers are corrupt; those frames end as BAD_HEADER after sync + header decode and are NOT credited."""
    return _streams_leg(args, 64, 30.0, args.syms, steps, barrier, local_rank)


def wide_leg(args, steps, barrier, local_rank):
    """BASELINE.json configs[3]: the 1024-subcarrier variant, 4096 streams x 128 data symbols."""
    return _streams_leg(args, 1024, 50.0, 128, steps, barrier, local_rank)


def tx_wide_leg(args, steps, local_rank):
    """TX half of BASELINE.json configs[3]: 4096 frames of 128 data symbols x 1024 subcarriers (the frames the wide leg receives)."""
    global NFFT
    import copy
    a = copy.copy(args)
    a.syms = 128
    NFFT = 1024
    try:
        return tx_bench(a, 0, local_rank, 1, steps=steps)
    finally:
        NFFT = args.nfft


def capture_wide_leg(args, steps, local_rank):
    """The capture search for the 1024-subcarrier layout (docs/SPEC.md 9): 4e8 samples, one 128-symbol frame every 1 000 003."""
    global NFFT
    import copy
    a = copy.copy(args)
    a.syms, a.capture_samples, a.no_cpu = 128, 4e8, True
    NFFT = 1024
    try:
        return capture_bench(a, 0, local_rank, 1, steps=steps)
    finally:
        NFFT = args.nfft


def capture_bench(args, rank, local_rank, world, steps=None):
    """BASELINE.json configs[2]: Schmidl-Cox preamble search + CFO estimation over one long capture (8 B/sample, one pass).
    A noise floor (sigma 0.01) with one 64QAM frame (S=2038) every 1 000 003 samples (prime stride), each with its own CFO.
    Under torchrun ONE capture of --capture-samples x N samples is split over the N ranks (SURVEY.md 8e partition 2:
    contiguous ranges, 2 L + frame_len overlap, duplicates removed by ownership of the offset, peak counts gathered); every rank
    synthesises just the samples it reads. Only CPU baseline this config has: the reference's own method on a 2 M-sample
    window (examples/jetson_rx.rs:16), timed on the host in the same run (cpu_baseline)."""
    import torch
    import ofdm_b200 as ob
    from ofdm_b200 import dist as od
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_cfg()
    eng = ob.Engine(cfg, local_rank)
    n_total = int(args.capture_samples) * world
    stride, S = 1_000_003, args.syms
    payload_len = cfg.max_payload(S)
    frame_len = cfg.frame_len(payload_len)
    shard = od.capture_shards(n_total, world, frame_len, sym_len=cfg.sym_len, guard=od.CAPTURE_GUARD * (cfg.sym_len // 80))[rank]
    n = shard.read_hi - shard.read_lo                                          # samples this rank reads
    g = torch.Generator(device=dev)
    g.manual_seed(0x0FD3 + rank)
    st = torch.cuda.current_stream().cuda_stream
    # one transmitted frame (same on every rank), reused with a different CFO at every position
    gp = torch.Generator(device=dev)
    gp.manual_seed(0x0FD3)
    pay = torch.randint(0, 256, (1, payload_len), dtype=torch.uint8, device=dev, generator=gp)
    pl = torch.full((1,), payload_len, dtype=torch.int32, device=dev)
    tx = torch.empty((1, frame_len, 2), dtype=torch.float32, device=dev)
    fl = torch.zeros(1, dtype=torch.int32, device=dev)
    eng.tx_encode_device(pay.data_ptr(), pl.data_ptr(), payload_len, 1, tx.data_ptr(), frame_len, fl.data_ptr(), st)
    torch.cuda.synchronize()
    frame = torch.view_as_complex(tx[0])
    cap = torch.empty((n, 2), dtype=torch.float32, device=dev)
    cap.normal_(0.0, 0.01, generator=g)
    capc = torch.view_as_complex(cap)
    positions = np.arange(5000, n_total - frame_len - 200, stride, dtype=np.int64)              # global frame starts
    cfo_all = ((positions * 2654435761 % (1 << 32)) / float(1 << 32) * 2 - 1) * (0.9 * np.pi / cfg.sym_len)   # per frame, the same on every rank
    t = torch.arange(frame_len, device=dev, dtype=torch.float32)
    touching = np.flatnonzero((positions + frame_len > shard.read_lo) & (positions < shard.read_hi))
    for i in touching:                                                         # frames cut by the range's ends are added in part
        p, f = int(positions[i]), float(cfo_all[i])
        lo, hi = max(p, shard.read_lo), min(p + frame_len, shard.read_hi)
        rot = torch.polar(torch.ones(hi - lo, device=dev), f * t[lo - p: hi - p])
        capc[lo - shard.read_lo: hi - shard.read_lo] += frame[lo - p: hi - p] * rot
    max_peaks = 16384
    peaks = torch.zeros((max_peaks, 2), dtype=torch.int64, device=dev)          # 16-byte ofdm_peak records
    n_peaks = torch.zeros(1, dtype=torch.int32, device=dev)
    eng.reserve(0, n)
    torch.cuda.synchronize()

    def step():
        eng.sync_search_device(cap.data_ptr(), n, peaks.data_ptr(), max_peaks, n_peaks.data_ptr(), st)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    steps = steps or max(1, min(args.steps, 50))
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = eng.kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / steps
    k = int(n_peaks.item())
    rec = peaks[:k].cpu().numpy().view(ob.engine.PEAK_DTYPE).reshape(-1)
    rec = rec[rec["metric"] >= 0]                                              # frame heads cut by the range's end
    found, mine = od.owned_peaks(rec["offset"].astype(np.int64), shard)        # de-duplication: keep what this rank owns
    own = (positions - 1 >= shard.own_lo) & (positions - 1 < shard.own_hi)     # lag - 1 rule, no channel delay
    want = positions[own] - 1
    exact = bool(int(mine.sum()) == len(want) and (found[mine] == want).all())
    cfo_err = float(np.abs(rec["f_delta"][mine] - cfo_all[own]).max()) if exact and len(want) else (0.0 if exact else float("nan"))
    red = torch.tensor([ms, 0.0 if exact else 1.0, cfo_err if exact else 1e9], dtype=torch.float64, device=dev)
    cnt = torch.tensor([int(mine.sum()), k], dtype=torch.int64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)                              # the gather of the (owned) peak counts
    ms, exact_all, cfo_err = float(red[0].item()), float(red[1].item()) == 0.0, float(red[2].item())
    frames_once, detections_raw = int(cnt[0].item()), int(cnt[1].item())
    peak, src = read_peaks()
    achieved = 8.0 * n / (ms * 1e-3) / 1e9                                     # per-GPU kernel rate on the samples a rank reads
    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = capture_cpu_baseline(cap, args)
        line = ({"metric": "sync_search_msamples_per_s", "value": round(n_total / (ms * 1e-3) / 1e6, 1), "unit": "Msamples/s",
                 "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4),
                 "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                 "config": {"workload": f"capture_{n_total}" + ("" if NFFT == 64 else f"_N{NFFT}"), "samples": n_total, "samples_read_per_gpu": n, "frames": int(len(positions)),
                            "frame_stride": stride, "frame_samples": frame_len, "noise_sigma": 0.01,
                            "sharding": "one capture, contiguous ranges, overlap 2L + frame_len, de-duplicated by offset ownership" if world > 1 else "none",
                            "l2": "input (%.1f GB/GPU) larger than L2" % (8 * n / 1e9)},
                 "roofline": {"bound": "hbm", "kernel": "sync_scan_kernel (+select, refine)" if NFFT == 64 else "wide_scan_kernel (+select, refine)", "achieved": round(achieved, 1), "peak": peak,
                              "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": read_traffic(f"capture_{n}"), "peak_source": src,
                              "algorithmic_bytes_per_launch": 8 * n, "kernel_ms": round(ms, 4)},
                 "all_offsets_exact": exact_all, "max_cfo_abs_err": cfo_err, "peaks_found": frames_once,
                 "each_frame_found_once": bool(exact_all and frames_once == len(positions)), "detections_before_dedup": detections_raw,
                 "cpu_baseline": cpu, "gpu_launches": int(eng.kernel_launches - l0), "clocks": clocks})
    eng.close()
    del cap, capc
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return line


def capture_cpu_baseline(cap, args):
    """The reference's own method for this config (src/signals/mod.rs:186-217 via src/receiver.rs:20-21): one whole-buffer
    cross-correlation against the locking ramp + arg-max, on the radio example's 2 000 000-sample buffer
    (examples/jetson_rx.rs:16) -- three length-(2M - 1) complex f64 transforms. Timed with the oracle's FFT-based restatement
    on one host thread (the reference is single-threaded), windows taken from the start of this run's capture."""
    from oracle import oracle as oo
    M = 2_000_000
    lock = oo.locking_signal(80 if NFFT == 64 else 1280)
    win = cap[:M].cpu().numpy().view(np.complex64).reshape(-1).astype(np.complex128)
    oo.xcorr_fft(win[:4096], lock)                                             # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        idx, _ = oo.xcorr_fft(win, lock)                                         # correlation + first strict arg-max of |c|^2
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min(args.cpu_seconds, 6.0):
            break
    return {"value": round(M * reps / dt / 1e6, 3), "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": f"{reps} x one 2 000 000-sample window ({dt:.1f} s): xcorr_fft against the locking ramp + arg-max, f64, three power-of-two "
                      "transforms (the reference's are of length 2M - 1: this is an upper bound on its speed); the reference's whole-buffer "
                      "search yields ONE frame per buffer, the engine every frame",
            "argmax_index": idx}


def tx_kernel_label(launches, steps):
    """Which TX kernel ran, from OFDM_TX_PATH (the engine's A/B switch) and the launches per step."""
    path = os.environ.get("OFDM_TX_PATH", "")
    wide = NFFT == 1024
    if path == "twopass" or (launches == 2 * steps and path not in ("", "spec")):
        return ("wide_tx_kernel" if wide else "tx_tile_kernel") + " (max pass + store pass)"
    if launches == 2 * steps:
        return ("wide_tx_spec_kernel" if wide else "tx_spec_kernel") + (" (one pass: every symbol scaled with the head maximum and stored at once; frames whose "
                                                                          "data beat it are redone) + redo-check launch")
    if path == "resident" or wide:
        return ("wide_tx_resident_kernel" if wide else "tx_resident_kernel") + " (one pass, frames resident in tensor memory)"
    return "tx_warp_kernel (one pass, frames resident in tensor memory, barrier-free frame loop)"


def tx_bench(args, rank, local_rank, world, steps=None):
    """TX side of the path: payload bytes -> Hamming -> 64QAM -> IFFT -> CP -> head -> normalise, for the
    headline workload's frames; the round trip against the RX path and the oracle is in tests/. 8 B/sample written once."""
    import torch
    import ofdm_b200 as ob
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = workload_cfg()
    eng = ob.Engine(cfg, local_rank)
    n, S = args.streams, args.syms
    plen_b = cfg.max_payload(S)
    frame_len = cfg.frame_len(plen_b)
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + rank)
    st = torch.cuda.current_stream().cuda_stream
    pstride = (plen_b + 15) // 16 * 16
    payload = torch.randint(0, 256, (n, pstride), dtype=torch.uint8, device=dev, generator=g)
    plen = torch.full((n,), plen_b, dtype=torch.int32, device=dev)
    tx = torch.empty((n, frame_len, 2), dtype=torch.float32, device=dev)
    flen = torch.zeros(n, dtype=torch.int32, device=dev)

    def step():
        eng.tx_encode_device(payload.data_ptr(), plen.data_ptr(), pstride, n, tx.data_ptr(), frame_len, flen.data_ptr(), st)

    steps = steps or max(1, min(args.steps, 50))
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = eng.kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    samples = n * frame_len
    by = 8 * samples + n * plen_b
    peak, src = read_peaks()
    mx = float(tx.max().item())
    frames_ok = bool((flen == frame_len).all().item())
    # two frames of the batch against the CPU oracle's `encode` (the checker, outside the timed region): fp32 IFFT vs f64
    oracle_ok = None
    if rank == 0:
        try:
            import numpy as np
            from oracle import oracle as oo
            ocfg = oracle_cfg()
            oracle_ok = True
            for i in (0, n - 1):
                ref = oo.tx(payload[i, :plen_b].cpu().numpy().tobytes(), ocfg)
                got = tx[i].cpu().numpy().view(np.complex64).reshape(-1)
                oracle_ok &= bool(ref.size == frame_len and np.allclose(got, ref, atol=2e-6))
        except Exception as e:          # noqa: BLE001
            oracle_ok = f"unchecked: {e}"
    launches = int(eng.kernel_launches - l0)
    label = tx_kernel_label(launches, steps)
    eng.close()
    # beside it: the exact one-pass kernel, whose cost does not depend on the payloads (the speculative default redoes a frame
    # whose data symbols beat the head maximum -- none in this workload, like any scrambled payload)
    exact = None
    if not os.environ.get("OFDM_TX_PATH"):
        os.environ["OFDM_TX_PATH"] = "warp" if NFFT == 64 else "resident"
        try:
            eng2 = ob.Engine(cfg, local_rank)
            for _ in range(3):
                eng2.tx_encode_device(payload.data_ptr(), plen.data_ptr(), pstride, n, tx.data_ptr(), frame_len, flen.data_ptr(), st)
            torch.cuda.synchronize()
            l2 = eng2.kernel_launches
            e0.record()
            for _ in range(steps):
                eng2.tx_encode_device(payload.data_ptr(), plen.data_ptr(), pstride, n, tx.data_ptr(), frame_len, flen.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / steps
            exact = {"kernel": tx_kernel_label(int(eng2.kernel_launches - l2), steps), "ms_per_step": round(ms2, 4),
                     "frac": round(by / (ms2 * 1e-3) / 1e9 / peak, 4)}
            eng2.close()
        finally:
            del os.environ["OFDM_TX_PATH"]
    del tx
    return ({"metric": "tx_msamples_per_s", "value": round(samples / (ms * 1e-3) / 1e6, 1), "unit": "Msamples/s", "n_gpus": 1,
                      "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"tx_{n}x64QAM_S{S}" + ("" if NFFT == 64 else f"_N{NFFT}"), "streams_per_gpu": n, "nfft": NFFT, "data_syms_per_frame": S, "frame_samples": frame_len,
                                 "payload_bytes": plen_b, "l2": "output (%.2f GB) larger than L2" % (8 * samples / 1e9)},
                      "roofline": {"bound": "hbm", "kernel": label, "achieved": round(by / (ms * 1e-3) / 1e9, 1),
                                   "peak": peak, "unit": "GB/s", "frac": round(by / (ms * 1e-3) / 1e9 / peak, 4), "traffic": read_traffic(f"tx_{n}x64QAM_S{S}" + ("" if NFFT == 64 else f"_N{NFFT}")),
                                   "peak_source": src, "algorithmic_bytes_per_launch": by, "kernel_ms": round(ms, 4)},
                      "max_component": round(mx, 6), "frames_ok": frames_ok, "frames_match_oracle_on_sample": oracle_ok, "exact_one_pass": exact,
                      "gpu_launches": launches, "clocks": clocks})


def rs_bench(args, rank, local_rank, world):
    """the RS(255,223) outer code on the headline workload's payloads (n streams x 41 915 B): encode,
    decode of clean codewords, decode with 8 symbol errors in every block; checked against the CPU oracle on two streams."""
    import torch
    import ofdm_b200 as ob
    from oracle import oracle as oo
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = ob.Engine(workload_cfg(), local_rank)
    n, plen = args.streams, workload_cfg().max_payload(args.syms)
    clen_max = 255 * (plen // 223 + 1)
    dlen_max = 223 * (clen_max // 255 + 1)
    g = torch.Generator(device=dev)
    g.manual_seed(0x255223 + rank)
    st = torch.cuda.current_stream().cuda_stream
    pay = torch.randint(0, 256, (n, plen), dtype=torch.uint8, device=dev, generator=g)
    pl = torch.full((n,), plen, dtype=torch.int32, device=dev)
    coded = torch.zeros((n, clen_max), dtype=torch.uint8, device=dev)
    cl = torch.zeros(n, dtype=torch.int32, device=dev)
    out = torch.zeros((n, dlen_max), dtype=torch.uint8, device=dev)
    ol = torch.zeros(n, dtype=torch.int32, device=dev)
    nc = torch.zeros(n, dtype=torch.int32, device=dev)
    nf = torch.zeros(n, dtype=torch.int32, device=dev)

    def enc():
        eng.rs_encode_device(pay.data_ptr(), pl.data_ptr(), n, plen, coded.data_ptr(), clen_max, cl.data_ptr(), st)

    def dec(src):
        eng.rs_decode_device(src.data_ptr(), cl.data_ptr(), n, clen_max, out.data_ptr(), dlen_max, ol.data_ptr(), nc.data_ptr(),
                             nf.data_ptr(), st)

    enc()
    torch.cuda.synchronize()
    nblk = clen_max // 255
    bad = coded.clone().view(n, nblk, 255)
    pos = torch.rand((n, nblk, 255), device=dev, generator=g).argsort(dim=2)[:, :, :8]          # 8 distinct positions per block
    flip = torch.randint(1, 256, (n, nblk, 8), dtype=torch.uint8, device=dev, generator=g)
    bad.scatter_(2, pos, bad.gather(2, pos) ^ flip)
    bad = bad.view(n, clen_max)
    # a realistic operating point: 2 % of the blocks carry 1..8 symbol errors, the others are clean
    hit = torch.rand((n, nblk, 1), device=dev, generator=g) < 0.02
    ne_blk = torch.randint(1, 9, (n, nblk, 1), device=dev, generator=g)
    keep = hit & (torch.arange(8, device=dev).view(1, 1, 8) < ne_blk)
    sparse = coded.clone().view(n, nblk, 255)
    sparse.scatter_(2, pos, sparse.gather(2, pos) ^ torch.where(keep, flip, torch.zeros_like(flip)))
    sparse = sparse.view(n, clen_max)
    sparse_errs = int(keep.sum().item())

    def timed(fn, steps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    steps = max(1, min(args.steps, 50))
    l0 = eng.kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_enc = timed(enc, steps)
    ms_clean = timed(lambda: dec(coded), steps)
    clean_ok = bool((out[:, :plen] == pay).all().item() and int(nf.sum().item()) == 0 and int(nc.sum().item()) == 0)
    ms_bad = timed(lambda: dec(bad), steps)
    fixed_ok = bool((out[:, :plen] == pay).all().item() and int(nf.sum().item()) == 0 and int(nc.sum().item()) == 8 * n * nblk)
    ms_sparse = timed(lambda: dec(sparse), steps)
    clocks = sampler.stop()
    sparse_ok = bool((out[:, :plen] == pay).all().item() and int(nf.sum().item()) == 0 and int(nc.sum().item()) == sparse_errs)
    oracle_ok = True
    for i in (0, n - 1):
        oracle_ok &= bool((oo.rs_encode(pay[i].cpu().numpy()) == coded[i].cpu().numpy()).all())
    peak, src = read_peaks()
    by = n * (plen + clen_max)
    if True:
        return ({"metric": "rs255_223_decode_gbyte_per_s", "value": round(n * clen_max / (ms_clean * 1e-3) / 1e9, 1), "unit": "GB/s",
                          "n_gpus": 1, "steps": steps, "warmup": 3, "ms_per_step": round(ms_clean, 4), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": f"rs255_223_{n}x{plen}B", "streams_per_gpu": n, "payload_bytes": plen, "coded_bytes": clen_max,
                                     "blocks_per_stream": nblk},
                          "encode_ms": round(ms_enc, 4), "decode_clean_ms": round(ms_clean, 4), "decode_8_errors_per_block_ms": round(ms_bad, 4),
                          "decode_2pct_blocks_with_errors_ms": round(ms_sparse, 4),
                          "roofline": {"bound": "hbm", "kernel": "rs_decode_kernel", "achieved": round(by / (ms_clean * 1e-3) / 1e9, 1), "peak": peak,
                                       "unit": "GB/s", "frac": round(by / (ms_clean * 1e-3) / 1e9 / peak, 4), "traffic": None, "peak_source": src,
                                       "algorithmic_bytes_per_launch": by},
                          "checks": {"clean_round_trip": clean_ok, "all_8_error_blocks_repaired": fixed_ok, "sparse_error_blocks_repaired": sparse_ok,
                                     "encode_matches_oracle": oracle_ok},
                          "gpu_launches": int(eng.kernel_launches - l0), "clocks": clocks})


def ingest_bench(args, rank, local_rank, world):
    """an fc32 capture FILE (page cache) -> pinned chunks -> PCIe -> preamble search -> every frame decoded.
    2^27 samples (1 GiB) of noise floor with a 64QAM frame (S=254) every 200 003 samples; end-to-end wall clock."""
    import tempfile
    import time
    import torch
    import ofdm_b200 as ob
    from ofdm_b200 import ingest
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = workload_cfg()
    eng = ob.Engine(cfg, local_rank)
    n, stride, S = 1 << 27, 200_003, 254
    payload_len = cfg.max_payload(S)
    frame_len = cfg.frame_len(payload_len)
    g = torch.Generator(device=dev)
    g.manual_seed(0x1F11E)
    st = torch.cuda.current_stream().cuda_stream
    pay = torch.randint(0, 256, (1, payload_len), dtype=torch.uint8, device=dev, generator=g)
    pl = torch.full((1,), payload_len, dtype=torch.int32, device=dev)
    tx = torch.empty((1, frame_len, 2), dtype=torch.float32, device=dev)
    fl = torch.zeros(1, dtype=torch.int32, device=dev)
    eng.tx_encode_device(pay.data_ptr(), pl.data_ptr(), payload_len, 1, tx.data_ptr(), frame_len, fl.data_ptr(), st)
    torch.cuda.synchronize()
    frame = torch.view_as_complex(tx[0])
    cap = torch.empty((n, 2), dtype=torch.float32, device=dev)
    cap.normal_(0.0, 0.001, generator=g)
    capc = torch.view_as_complex(cap)
    positions = list(range(5000, n - frame_len - 200, stride))
    for p in positions:
        capc[p: p + frame_len] += frame
    want = pay[0].cpu().numpy().tobytes()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        path = os.path.join(d, "capture.fc32")
        cap.cpu().numpy().tofile(path)
        del cap, capc
        kw = dict(chunk_samples=1 << 25, max_frame_samples=1 << 15, out_stride=payload_len + 64, max_peaks=1024)
        rx = ingest.StreamReceiver(cfg, local_rank, **kw)
        ingest.decode_file(path, cfg, stop=1 << 26, receiver=rx)                       # warm-up (page cache, allocations)
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            frames = ingest.decode_file(path, cfg, receiver=rx)
            runs.append(time.perf_counter() - t0)
        rx.close()
        # the same file through the C entry a Rust / C caller binds (ofdm_rx_decode_file: parallel pread into two pinned
        # buffers one chunk ahead of the GPU)
        eng.decode_file(path, stop=1 << 26, chunk_samples=kw["chunk_samples"], max_frame_samples=kw["max_frame_samples"],
                        out_stride=kw["out_stride"], max_frames=2048)
        c_runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            rec, data = eng.decode_file(path, chunk_samples=kw["chunk_samples"], max_frame_samples=kw["max_frame_samples"],
                                        out_stride=kw["out_stride"], max_frames=2048)
            c_runs.append(time.perf_counter() - t0)
        c_sec = sorted(c_runs)[1]
        c_exact = [int(x) for x in rec["offset"]] == [p - 1 for p in positions] and all(d == want for d in data)
    eng.close()
    sec = sorted(runs)[1]
    ok = [f for f in frames if f.status == 0 and f.data == want]
    exact = [f.offset for f in frames] == [p - 1 for p in positions]
    return ({"metric": "file_ingest_msamples_per_s", "value": round(n / sec / 1e6, 1), "unit": "Msamples/s", "n_gpus": 1,
                      "steps": 3, "warmup": 1, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "fc32_file_2^27_samples", "file_bytes": 8 * n, "frames": len(positions), "frame_samples": frame_len,
                                 "chunk_samples": kw["chunk_samples"], "source": "page cache (/dev/shm)", "timing": "host wall clock, median of 3"},
                      "file_gbyte_per_s": round(8 * n / sec / 1e9, 2), "frames_found": len(frames), "frames_decoded_exact": len(ok),
                      "c_abi": {"entry": "ofdm_rx_decode_file", "ms": round(c_sec * 1e3, 2), "msamples_per_s": round(n / c_sec / 1e6, 1),
                                "file_gbyte_per_s": round(8 * n / c_sec / 1e9, 2), "all_frames_exact": bool(c_exact)},
                      "all_offsets_exact": exact})


def reference_arm(args):
    """--impl reference: the reference's CPU algorithm with all host threads, no GPU and none of the engine's code.

    The Rust crate cannot be built here (no rustc/cargo, out-of-tree dependencies), so this times the f64 C oracle port
    of the same algorithm. The workload is the headline's (same `config`), synthesised on the CPU by the oracle's own TX +
    channel; a step is a bounded sample of it: 2 streams per host thread, decoded `passes` times, with `passes` calibrated so
    that the --steps K timed steps take about 3 s in total. Exactly --warmup W untimed and --steps K timed steps are run.
    """
    global NFFT
    NFFT = args.nfft
    from oracle import oracle as oo
    threads = host_threads()
    cfg = oracle_cfg()
    S = args.syms
    if NFFT == 1024 and args.snr == 40.0:
        args.snr = 50.0
    payload_len, frame_len, iq_stride, out_stride = headline_sizes(args, NFFT)
    k = max(threads, 1) * 2
    rng = np.random.default_rng(SEED)
    iq = np.zeros((k, iq_stride), np.complex64)
    ns = np.zeros(k, np.uint32)
    pays = []
    for i in range(k):
        pay = rng.integers(0, 256, payload_len, dtype=np.uint8)
        pays.append(pay)
        tx = oo.tx(pay, cfg)
        assert tx.size == frame_len
        ch = oo.channel(tx, args.snr, float(rng.uniform(0, 0.9 * np.pi / (NFFT + NFFT // 4))), 1, SEED + i)
        lead = int(rng.integers(LEAD_MIN, LEAD_MAX + 1))
        sigma = np.sqrt(0.5 * np.mean(np.abs(ch) ** 2) / 10 ** (args.snr / 10))
        noise = sigma * (rng.standard_normal(lead) + 1j * rng.standard_normal(lead))
        cap = np.concatenate([noise, ch])
        iq[i, : cap.size] = cap
        ns[i] = cap.size
    f32 = iq.view(np.float32).reshape(k, iq_stride, 2)
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # calibration (untimed): how many passes over the sample make K steps last ~3 s
    oo.decode_batch_fc32(f32, ns, cfg, out_stride, threads)
    t0 = time.perf_counter()
    oo.decode_batch_fc32(f32, ns, cfg, out_stride, threads)
    t_pass = max(time.perf_counter() - t0, 1e-4)
    passes = max(1, int(round(3.0 / (steps * t_pass))))

    def step():
        for _ in range(passes):
            r = oo.decode_batch_fc32(f32, ns, cfg, out_stride, threads)
        return r

    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        out, out_len, status, _ = step()
    dt = (time.perf_counter() - t0) / steps
    ok = all(status[i] == 0 and out_len[i] == payload_len and (out[i, :payload_len] == pays[i]).all() for i in range(k))
    v = float(ns.sum()) * passes / dt / 1e6
    gbit = float(out_len[status == 0].sum()) * 8 * passes / dt / 1e9
    config = headline_config(args, NFFT, args.streams, args.snr)
    cpu = {"value": round(v, 2), "unit": "Msamples/s", "cores": threads, "kind": "port",
           "sample": f"{k} streams of the workload x {passes} passes per step (f64 C port of the reference algorithm, {threads} host threads, "
                     "FFT plans reused = upper bound on the Rust crate's speed)",
           "payloads_recovered": bool(ok)}
    line = {"impl": "reference", "metric": "rx_msamples_per_s", "value": round(v, 2), "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "decoded_gbit_per_s": round(gbit, 4),
            "cpu_baseline": cpu, "e2e": {"value": round(v, 2), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
