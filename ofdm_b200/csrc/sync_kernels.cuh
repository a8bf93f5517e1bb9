// sync_kernels.cuh -- preamble search over one long capture (BASELINE.json configs[2]: 1e9-sample capture).
//
// The reference handles a long capture with one whole-capture FFT cross-correlation (src/receiver.rs:20-21,
// src/signals/mod.rs:186-217: three length-(2M-1) f64 transforms, ~96 GB of scratch at M = 1e9 -- not runnable). Here
// the capture is scanned once (8 B/sample) with the sliding Schmidl-Cox metric of docs/SPEC.md 4:
//
//   sync_scan_kernel   : one CTA per 3928 lags. Coalesced IQ tile -> padded smem rows (one 8-sample row per thread);
//                        q[n] = conj(a[n]) a[n+80], e[n] = |a[n]|^2; per-row totals, 10-row window sums of the totals, and
//                        for the 8 lags of a row P(d) = W_q + pre_q(row+10) - pre_q(row), R1, R2 likewise from the
//                        thread-serial prefix sums inside rows t, t+10, t+20. A row whose bound
//                        (|W_q| + sum|q_t| + sum|q_t+10|)^2 stays below 0.5 min R1 min R2 cannot hold a lag above the
//                        threshold and skips the per-lag work (all but the preamble plateaus of a capture). Rising
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
    uint32_t *counters;      // [0] threshold crossings found, [1] entries of peaks[] = min(detections, max_peaks), [2] detections
    const RxTables *tables;
    SyncPeak *peaks;
    uint32_t max_peaks;
    uint32_t n_tiles;        // 3928-lag tiles of the capture; the scan grid is persistent and strides over them
};

constexpr size_t sync_scan_smem_bytes()
{
    return (size_t)kScanRows * kScanRowStride * sizeof(float2) + sizeof(float) * 5 * (kScanRows + 32) + sizeof(uint32_t) * (kScanRows + 8);
}

__device__ __forceinline__ cpx c_conj_mul(cpx a, cpx b)     // conj(a) * b
{
    float ar, ai, br, bi;
    c_split(a, ar, ai); c_split(b, br, bi);
    return c_fma2(c_make(bi, -br), c_make(ai, ai), c_mul2(b, c_make(ar, ar)));
}

// one tile's samples in flight: 4 x 16 B per thread
struct ScanTileRegs { unsigned long long v[kScanT]; };

__device__ __forceinline__ void scan_tile_load(const SyncArgs &a, long long tile, int t, ScanTileRegs &r)
{
    const long long n = (long long)a.n;
    const long long origin = tile * kScanD - kScanT;                   // sample index of row 0, column 0
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.iq) & 15) == 0) && ((origin & 1) == 0);
#pragma unroll
    for (int i = 0; i < kScanT / 2; i++) {
        const long long s0 = origin + 2 * (i * kScanRows + t);         // coalesced: 2 samples (16 B) per thread per step
        unsigned long long v0 = 0, v1 = 0;
        if (aligned && s0 >= 0 && s0 + 1 < n) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(a.iq + s0));
            v0 = v.x; v1 = v.y;
        } else {
            if (s0 >= 0 && s0 < n) v0 = __ldg(reinterpret_cast<const unsigned long long *>(a.iq + s0));
            if (s0 + 1 >= 0 && s0 + 1 < n) v1 = __ldg(reinterpret_cast<const unsigned long long *>(a.iq + s0 + 1));
        }
        r.v[2 * i] = v0; r.v[2 * i + 1] = v1;
    }
}

template <int = 0>
__global__ void __launch_bounds__(kScanRows, 2) sync_scan_kernel(const SyncArgs a)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    unsigned long long *s_iq = reinterpret_cast<unsigned long long *>(smem_raw);              // [row][9] packed complex
    float *s_off = reinterpret_cast<float *>(s_iq + kScanRows * kScanRowStride);              // row totals: q.re | q.im | e | sum |q|
    float *s_we = s_off + 4 * (kScanRows + 32);                                               // 80-sample energy windows at row granularity
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_we + (kScanRows + 32));
    constexpr int TQR = 0, TQI = kScanRows, TE = 2 * kScanRows, TA = 3 * kScanRows;

    const int t = threadIdx.x;
    const long long n = (long long)a.n;
    const long long d_last = n - 2 * kSym;                            // largest valid lag

    // persistent CTA: the next tile's samples are fetched into registers while this tile is processed from shared memory
    ScanTileRegs pre;
    if (blockIdx.x < a.n_tiles) scan_tile_load(a, blockIdx.x, t, pre);
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < (long long)a.n_tiles; tile += gridDim.x) {
    const long long d_base = tile * kScanD;                           // first lag evaluated from this tile (row 1)
    const long long origin = d_base - kScanT;                         // sample index of row 0, column 0
#pragma unroll
    for (int i = 0; i < kScanT / 2; i++) {
        const int idx = i * kScanRows + t;                             // float4 index inside the tile
        const int row = idx / (kScanT / 2), col = (idx % (kScanT / 2)) * 2;
        s_iq[row * kScanRowStride + col] = pre.v[2 * i];
        s_iq[row * kScanRowStride + col + 1] = pre.v[2 * i + 1];
    }
    if (tile + gridDim.x < (long long)a.n_tiles) scan_tile_load(a, tile + gridDim.x, t, pre);
    __syncthreads();

    // ---- row totals of q, e and of |q| (L1 norm, for the bound) -- one row per thread ---------------------------------
    // The running sums go through exactly the operations the per-lag prefix sums below use, so a total equals the prefix
    // "after column 7" bit for bit.
    cpx tq = c_make(0.0f, 0.0f);
    float te = 0.0f, ta = 0.0f;
    {
        const unsigned long long *own = s_iq + t * kScanRowStride;
        const bool has5 = t + kScanR80 < kScanRows;
        const unsigned long long *nxt = s_iq + (has5 ? t + kScanR80 : t) * kScanRowStride;
#pragma unroll
        for (int j = 0; j < kScanT; j++) {
            cpx x, y;
            x.v = own[j]; y.v = has5 ? nxt[j] : 0ull;
            float xr, xi, pr, pi;
            c_split(x, xr, xi);
            const cpx q = c_conj_mul(x, y);
            c_split(q, pr, pi);
            tq = c_add(tq, q);
            te = fmaf(xr, xr, fmaf(xi, xi, te));
            ta += fabsf(pr) + fabsf(pi);
        }
    }
    // ---- 80-sample window sums at row granularity: W[t] = sum of the row totals of rows t .. t+9 ---------------------
    // (summing the ten small row totals directly, instead of differencing a tile-wide prefix sum, keeps fp32 exact enough in
    // a quiet stretch that follows a loud frame inside the same tile)
    float qr, qi;
    c_split(tq, qr, qi);
    s_off[TQR + t] = qr; s_off[TQI + t] = qi; s_off[TE + t] = te; s_off[TA + t] = ta;
    __syncthreads();
    float wqr = 0.0f, wqi = 0.0f, we = 0.0f;
#pragma unroll
    for (int k = 0; k < kScanR80; k++) {
        const int r = t + k < kScanRows ? t + k : kScanRows - 1;          // rows past the tile are never used by an evaluated lag
        wqr += s_off[TQR + r]; wqi += s_off[TQI + r]; we += s_off[TE + r];
    }
    s_we[t] = we;
    __syncthreads();

    // ---- metric per lag: rows 0 .. kScanEvalRows (row 0 only provides above(d_base - 1)) ------------------------------
    uint32_t mask = 0;
    if (t <= kScanEvalRows) {
        constexpr int A = kScanR80, B = 2 * kScanR80;
        const float de1 = we, de2 = s_we[t + A];
        // Can any of the 8 lags of this row be above the threshold? For lag column j:
        //   P  = W_q + pre_q(t+A)[j] - pre_q(t)[j]            =>  |P| <= |W_q| + sum|q_t| + sum|q_t+A|
        //   R1 = W_e(t) + pre_e(t+A)[j] - pre_e(t)[j]         >=  W_e(t) - E_t       (likewise R2 one window later)
        // with slack for fp32 rounding of either side. Products that underflow compare 0 < 0 = false and take the exact path.
        const float bp = (fabsf(wqr) + fabsf(wqi) + ta + s_off[TA + t + A]) * 1.0001f;
        const float r1m = fmaxf(0.0f, (de1 - te) - 8e-6f * de1);
        const float r2m = fmaxf(0.0f, (de2 - s_off[TE + t + A]) - 8e-6f * de2);
        if (!(bp * bp < 0.499f * r1m * r2m)) {
            const cpx dq = c_make(wqr, wqi);
            const unsigned long long *x0 = s_iq + t * kScanRowStride, *x5 = x0 + A * kScanRowStride, *x10 = x0 + B * kScanRowStride;
            // lags of this row that exist: 0 <= d <= d_last (hoisted out of the loop as a bit mask)
            const long long dr = origin + (long long)t * kScanT;
            uint32_t live = (1u << kScanT) - 1u;
            if (dr < 0) live = dr <= -kScanT ? 0u : live & ~((1u << (int)(-dr)) - 1u);
            if (dr + kScanT - 1 > d_last) live = dr > d_last ? 0u : live & ((1u << (int)(d_last - dr + 1)) - 1u);
            cpx q0 = c_make(0.0f, 0.0f), q5 = q0;                          // exclusive prefix sums inside rows t, t+A (q) and t, t+A, t+B (e)
            float e0 = 0.0f, e5 = 0.0f, e10 = 0.0f;
#pragma unroll
            for (int j = 0; j < kScanT; j++) {
                float pr, pi;
                c_split(c_add(c_sub(q5, q0), dq), pr, pi);
                const float r1 = (e5 - e0) + de1, r2 = (e10 - e5) + de2;
                if (pr * pr + pi * pi > 0.5f * r1 * r2) mask |= 1u << j;
                cpx u, v, w;
                u.v = x0[j]; v.v = x5[j]; w.v = x10[j];
                float ur, ui, vr, vi, wr, wi;
                c_split(u, ur, ui); c_split(v, vr, vi); c_split(w, wr, wi);
                q0 = c_add(q0, c_conj_mul(u, v));
                q5 = c_add(q5, c_conj_mul(v, w));
                e0 = fmaf(ur, ur, fmaf(ui, ui, e0));
                e5 = fmaf(vr, vr, fmaf(vi, vi, e5));
                e10 = fmaf(wr, wr, fmaf(wi, wi, e10));
            }
            mask &= live;
        }
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
    __syncthreads();                                                   // s_iq / s_mask are rewritten by the next tile
    }
}

// sort candidates (bitonic, one CTA), then keep the first of every frame (hold-off)
template <int = 0>
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
        a.counters[2] = m;                                             // detections before truncation to max_peaks
        a.counters[1] = m < a.max_peaks ? m : a.max_peaks;             // entries of peaks[] (what *n_peaks reports)
    }
    __syncthreads();
}

// one CTA per accepted detection: refinement + CFO (docs/SPEC.md 4, 5)
template <int = 0>
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
template <int = 0>
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
