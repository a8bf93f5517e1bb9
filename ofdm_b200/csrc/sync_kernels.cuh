// sync_kernels.cuh -- preamble search over one long capture (BASELINE.json configs[2]: 1e9-sample capture).
//
// The reference handles a long capture with one whole-capture FFT cross-correlation (src/receiver.rs:20-21,
// src/signals/mod.rs:186-217: three length-(2M-1) f64 transforms, ~96 GB of scratch at M = 1e9 -- not runnable). Here
// the capture is scanned once (8 B/sample) with the sliding Schmidl-Cox metric of docs/SPEC.md 4:
//
//   sync_scan_kernel   : one CTA per 3928 lags. Coalesced IQ tile -> padded smem rows (one 8-sample row per thread);
//                        q[n] = conj(a[n]) a[n+80], e[n] = |a[n]|^2; thread-serial prefix sums inside a row plus
//                        10-row window sums of the row totals; P(d) = Q[d+80]-Q[d], R1(d) = E[d+80]-E[d], R2(d) = E[d+160]-E[d+80]; rising
//                        edges of |P|^2 > 0.5 R1 R2 are appended to a candidate list.
//   sync_select_kernel : one CTA: bitonic sort of the candidates, 800-sample hold-off (one detection per frame).
//   sync_refine_kernel : one CTA per detection: ramp-correlation arg-max around it (lag - 1 rule), CFO estimate.
#pragma once

#include "common.cuh"
#include "rx_kernels.cuh"

namespace ofdm {

constexpr int kScanT = 8;                           // samples per thread = one smem row
constexpr int kScanRows = 512;                      // rows per CTA (= threads)
constexpr int kScanR80 = kSym / kScanT;             // rows per 80 samples (10)
constexpr int kScanEvalRows = kScanRows - 2 * kScanR80 - 1;   // rows whose lags are evaluated here (row 0 only feeds above(d-1))
constexpr int kScanD = kScanEvalRows * kScanT;      // 3928 lags per CTA
constexpr int kScanRowStride = kScanT + 1;          // padded row length: conflict-free row-per-thread LDS.64
constexpr int kSyncCandCap = 8192;                  // candidates the select kernel can sort
constexpr int kSyncHoldoff = 800;                   // lock + preamble + training: one detection per frame

struct SyncPeak {            // = ofdm_peak in include/ofdm_engine.h
    uint64_t offset;         // src/receiver.rs:21 rule: ramp lag - 1
    float f_delta;           // rad/sample, angle of the preamble sum / 80
    float metric;            // |P|^2 / R^2 at the detection lag
};

struct SyncArgs {
    const float2 *iq;
    uint64_t n;
    uint32_t *cand;          // [kSyncCandCap] detection lags (unsorted), then sorted + filtered in place
    uint32_t *counters;      // [0] candidates found, [1] detections accepted (= entries of peaks[])
    const RxTables *tables;
    SyncPeak *peaks;
    uint32_t max_peaks;
};

constexpr size_t sync_scan_smem_bytes()
{
    return (size_t)kScanRows * kScanRowStride * (sizeof(float2) * 2 + sizeof(float)) + sizeof(float) * 4 * (kScanRows + 32) + sizeof(uint32_t) * (kScanRows + 8);
}

__device__ __forceinline__ cpx c_conj_mul(cpx a, cpx b)     // conj(a) * b
{
    float ar, ai, br, bi;
    c_split(a, ar, ai); c_split(b, br, bi);
    return c_fma2(c_make(bi, -br), c_make(ai, ai), c_mul2(b, c_make(ar, ar)));
}

__global__ void __launch_bounds__(kScanRows, 2) sync_scan_kernel(const SyncArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    unsigned long long *s_iq = reinterpret_cast<unsigned long long *>(smem_raw);              // [row][17] packed complex
    unsigned long long *s_q = s_iq + kScanRows * kScanRowStride;                              // thread-local exclusive prefix of q
    float *s_e = reinterpret_cast<float *>(s_q + kScanRows * kScanRowStride);                 // thread-local exclusive prefix of e
    float *s_off = s_e + kScanRows * kScanRowStride;                                          // [3][rows + 8] row offsets (block scan)
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_off + 4 * (kScanRows + 32));

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long long n = (long long)a.n;
    const long long d_base = (long long)blockIdx.x * kScanD;          // first lag evaluated by this CTA (row 1)
    const long long origin = d_base - kScanT;                         // sample index of row 0, column 0
    const long long d_last = n - 2 * kSym;                            // largest valid lag

    // ---- coalesced tile load: 2 samples (16 B) per thread per step -> padded rows ---------------------------------
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.iq) & 15) == 0) && ((origin & 1) == 0);
#pragma unroll
    for (int i = 0; i < kScanT / 2; i++) {
        const int idx = i * kScanRows + t;                             // float4 index inside the tile
        const long long s0 = origin + 2 * idx;
        unsigned long long v0 = 0, v1 = 0;
        if (aligned && s0 >= 0 && s0 + 1 < n) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(a.iq + s0));
            v0 = v.x; v1 = v.y;
        } else {
            if (s0 >= 0 && s0 < n) v0 = __ldg(reinterpret_cast<const unsigned long long *>(a.iq + s0));
            if (s0 + 1 >= 0 && s0 + 1 < n) v1 = __ldg(reinterpret_cast<const unsigned long long *>(a.iq + s0 + 1));
        }
        const int row = idx / (kScanT / 2), col = (idx % (kScanT / 2)) * 2;
        s_iq[row * kScanRowStride + col] = v0;
        s_iq[row * kScanRowStride + col + 1] = v1;
    }
    __syncthreads();

    // ---- q, e and their thread-local exclusive prefix sums (one row per thread) -------------------------------------
    cpx tq = c_make(0.0f, 0.0f);
    float te = 0.0f;
    {
        const unsigned long long *own = s_iq + t * kScanRowStride;
        const bool has5 = t + kScanR80 < kScanRows;
        const unsigned long long *nxt = s_iq + (has5 ? t + kScanR80 : t) * kScanRowStride;
#pragma unroll
        for (int j = 0; j < kScanT; j++) {
            cpx x, y;
            x.v = own[j]; y.v = has5 ? nxt[j] : 0ull;
            s_q[t * kScanRowStride + j] = tq.v;
            s_e[t * kScanRowStride + j] = te;
            float xr, xi;
            c_split(x, xr, xi);
            tq = c_add(tq, c_conj_mul(x, y));
            te = fmaf(xr, xr, fmaf(xi, xi, te));
        }
    }
    // ---- 80-sample window sums at row granularity: W[t] = sum of the row totals of rows t .. t+9 ---------------------
    // (summing the ten small row totals directly, instead of differencing a tile-wide prefix sum, keeps fp32 exact enough in
    // a quiet stretch that follows a loud frame inside the same tile)
    float qr, qi;
    c_split(tq, qr, qi);
    s_off[t] = qr; s_off[kScanRows + t] = qi; s_off[2 * kScanRows + t] = te;
    __syncthreads();
    float wqr = 0.0f, wqi = 0.0f, we = 0.0f;
#pragma unroll
    for (int k = 0; k < kScanR80; k++) {
        const int r = t + k < kScanRows ? t + k : kScanRows - 1;          // rows past the tile are never used by an evaluated lag
        wqr += s_off[r]; wqi += s_off[kScanRows + r]; we += s_off[2 * kScanRows + r];
    }
    float *s_we = s_off + 3 * kScanRows;                                   // [rows]
    s_we[t] = we;
    __syncthreads();

    // ---- metric per lag: rows 0 .. kScanEvalRows (row 0 only provides above(d_base - 1)) ------------------------------
    uint32_t mask = 0;
    if (t <= kScanEvalRows) {
        constexpr int A = kScanR80, B = 2 * kScanR80;
        const float de1 = we, de2 = s_we[t + A];
        const cpx dq = c_make(wqr, wqi);
        const unsigned long long *q0 = s_q + t * kScanRowStride, *q5 = s_q + (t + A) * kScanRowStride;
        const float *e0 = s_e + t * kScanRowStride, *e5 = s_e + (t + A) * kScanRowStride, *e10 = s_e + (t + B) * kScanRowStride;
        // lags of this row that exist: 0 <= d <= d_last (hoisted out of the loop as a bit mask)
        const long long dr = origin + (long long)t * kScanT;
        uint32_t live = (1u << kScanT) - 1u;
        if (dr < 0) live = dr <= -kScanT ? 0u : live & ~((1u << (int)(-dr)) - 1u);
        if (dr + kScanT - 1 > d_last) live = dr > d_last ? 0u : live & ((1u << (int)(d_last - dr + 1)) - 1u);
#pragma unroll
        for (int j = 0; j < kScanT; j++) {
            cpx a0, a5;
            a0.v = q0[j]; a5.v = q5[j];
            float pr, pi;
            c_split(c_add(c_sub(a5, a0), dq), pr, pi);
            const float e5j = e5[j];
            const float r1 = (e5j - e0[j]) + de1, r2 = (e10[j] - e5j) + de2;
            if (pr * pr + pi * pi > 0.5f * r1 * r2) mask |= 1u << j;
        }
        mask &= live;
    }
    s_mask[t] = mask;
    __syncthreads();
    if (t >= 1 && t <= kScanEvalRows && mask) {
        const uint32_t prev = s_mask[t - 1] >> (kScanT - 1);
        uint32_t edges = mask & ~((mask << 1) | prev) & ((1u << kScanT) - 1u);       // above(d) && !above(d - 1)
        while (edges) {
            const int j = __ffs(edges) - 1;
            edges &= edges - 1;
            const uint32_t slot = atomicAdd(a.counters, 1u);
            if (slot < kSyncCandCap) a.cand[slot] = (uint32_t)(origin + (long long)t * kScanT + j);
        }
    }
}

// sort candidates (bitonic, one CTA), then keep the first of every frame (hold-off)
__global__ void __launch_bounds__(1024) sync_select_kernel(const SyncArgs a)
{
    __shared__ uint32_t s_key[kSyncCandCap];
    uint32_t n = a.counters[0];
    if (n > kSyncCandCap) n = kSyncCandCap;
    for (int i = threadIdx.x; i < kSyncCandCap; i += blockDim.x) s_key[i] = i < (int)n ? a.cand[i] : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= kSyncCandCap; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < kSyncCandCap; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t x = s_key[i], y = s_key[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { s_key[i] = y; s_key[p] = x; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        long long last = -(long long)kSyncHoldoff;
        for (uint32_t i = 0; i < n; i++) {
            const long long d = s_key[i];
            if (d >= last + kSyncHoldoff) { a.cand[m++] = (uint32_t)d; last = d; }
        }
        a.counters[1] = m;
    }
    __syncthreads();
}

// one CTA per accepted detection: refinement + CFO (docs/SPEC.md 4, 5)
__global__ void __launch_bounds__(kAcqThreads) sync_refine_kernel(const SyncArgs a)
{
    __shared__ float s_lock[kSym];
    __shared__ float s_val[kAcqThreads / 32];
    __shared__ int s_idx[kAcqThreads / 32];
    const uint32_t i = blockIdx.x;
    if (i >= a.counters[1] || i >= a.max_peaks) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < kSym) s_lock[tid] = a.tables->lock[tid].x;
    __syncthreads();
    const long long d0 = a.cand[i], n = (long long)a.n;
    long long k_lo = d0 - 176, k_hi = d0 + 16;
    if (k_lo < -(kSym - 1)) k_lo = -(kSym - 1);
    // the ramp correlation works on lags relative to a window origin (exact for captures beyond 2^31 samples)
    const long long org = k_lo < 0 ? 0 : k_lo;
    const long long offset = org + (long long)ramp_argmax<kAcqThreads>(a.iq + org, n - org, k_lo - org, k_hi - org, s_lock, s_val, s_idx) - 1;
    const bool ok = offset >= 0 && offset + 800 <= n;
    // metric at the detection lag and (if the frame head fits) the CFO estimate: f64 sums of 80 terms
    double acc[6] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    if (tid < kSym) {
        const float2 u = a.iq[d0 + tid], v = a.iq[d0 + kSym + tid];
        acc[0] = (double)u.x * v.x + (double)u.y * v.y;              // P(d0)
        acc[1] = (double)u.x * v.y - (double)u.y * v.x;
        acc[2] = (double)v.x * v.x + (double)v.y * v.y;              // R2(d0)
        acc[5] = (double)u.x * u.x + (double)u.y * u.y;              // R1(d0)
        if (ok) {
            const float2 *x0 = a.iq + offset;
            const float2 r2 = x0[2 * kSym + tid], r3 = x0[3 * kSym + tid], r4 = x0[4 * kSym + tid];
            acc[3] = (double)r2.x * r3.x + (double)r2.y * r3.y + (double)r3.x * r4.x + (double)r3.y * r4.y;
            acc[4] = (double)r2.x * r3.y - (double)r2.y * r3.x + (double)r3.x * r4.y - (double)r3.y * r4.x;
        }
    }
    __shared__ double s_acc[6 * (kAcqThreads / 32)];
#pragma unroll
    for (int q = 0; q < 6; q++) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], m);
        if (lane == 0) s_acc[q * (kAcqThreads / 32) + warp] = acc[q];
    }
    __syncthreads();
    if (tid == 0) {
        double tt[6];
        for (int q = 0; q < 6; q++) { tt[q] = 0.0; for (int w = 0; w < kAcqThreads / 32; w++) tt[q] += s_acc[q * (kAcqThreads / 32) + w]; }
        SyncPeak p;
        p.offset = ok ? (uint64_t)offset : ~0ull;
        p.f_delta = ok ? (float)(atan2(tt[4], tt[3]) / 80.0) : 0.0f;
        p.metric = ok ? (float)((tt[0] * tt[0] + tt[1] * tt[1]) / (tt[5] * tt[2])) : -1.0f;      // < 0 marks an unusable detection
        a.peaks[i] = p;
    }
}

// per detected frame: where its capture starts and how many samples belong to it (up to the next frame / max_frame)
__global__ void capture_prep_kernel(const SyncPeak *__restrict__ peaks, uint32_t n_frames, uint64_t n, uint32_t max_frame,
                                    uint64_t *__restrict__ base, uint32_t *__restrict__ n_samples)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames) return;
    const SyncPeak p = peaks[i];
    if (p.metric < 0.0f || p.offset >= n) { base[i] = 0; n_samples[i] = 0; return; }      // unusable detection -> TOO_SHORT
    uint64_t end = n;
    for (uint32_t k = i + 1; k < n_frames; k++)
        if (peaks[k].metric >= 0.0f && peaks[k].offset > p.offset) { end = peaks[k].offset; break; }
    uint64_t len = end - p.offset;
    if (max_frame && len > max_frame) len = max_frame;
    if (len > 0xFFFFFFFFull) len = 0xFFFFFFFFull;
    base[i] = p.offset;
    n_samples[i] = (uint32_t)len;
}

}  // namespace ofdm
