// wide_sync_kernels.cuh -- preamble search over one long capture for the 1024-subcarrier layout (docs/SPEC.md 4 and 9: the
// search of sync_kernels.cuh with every length scaled by 16 -- lag and window L = 1280, hold-off and minimum frame head 10 L,
// refinement window [d - 11 L / 5, d + L / 5]). Same candidate slots, hold-off kernel and peak format as the nfft = 64 search.
//
//   wide_scan_kernel        : one CTA per 4096 lags. The 4096 + 2 L + 8 samples the lags need are staged in shared memory;
//                             q[n] = conj(a[n]) a[n + L], e[n] = |a[n]|^2 are summed per 8-sample row, the row totals are
//                             prefix-summed in f64 (a 160-row window of fp32 totals would otherwise lose the quiet stretch
//                             that follows a loud frame), and a thread then slides P, R1, R2 over the 8 lags of its row.
//                             Rising edges of |P|^2 > 0.5 R1 R2 go to the tile's slots.
//   wide_sync_refine_kernel : one CTA per detection: ramp-correlation arg-max around it (closed form for the built-in ramp,
//                             lag - 1 rule), f64 CFO estimate and metric.
#pragma once

#include "sync_kernels.cuh"
#include "wide_kernels.cuh"

namespace ofdm {

constexpr int kWScanD = 4096;                                   // lags per tile
constexpr int kWScanThreads = 512;                              // one 8-lag row per thread
constexpr int kWScanR = wide::kL / 8;                           // rows per symbol length (160)
constexpr int kWScanRows = kWScanD / 8 + 2 * kWScanR + 1;       // 577 rows of 8 samples; row 0 only feeds above(d - 1)
constexpr int kWScanSamples = kWScanRows * 8;                   // 4616 samples staged per tile
constexpr int kWScanPitch = 9;                                  // samples per padded shared-memory row: a thread walks its own row (stride 72 B:
                                                                // the 16 lanes of an LDS.64 phase hit 16 distinct bank pairs)
constexpr int kWScanPre = kWScanRows + 3;                       // entries of a prefix array
constexpr size_t wide_scan_smem_bytes() { return sizeof(float2) * kWScanRows * kWScanPitch + 3 * sizeof(double) * kWScanPre + 2 * sizeof(float) * kWScanPre + 16; }
static_assert(kWScanD <= 65536, "tile-local lag offsets are 16 bits");

template <int = 0>
__global__ void __launch_bounds__(kWScanThreads) wide_scan_kernel(const SyncArgs a)
{
    constexpr int L = wide::kL;
    extern __shared__ __align__(16) uint8_t wscan_smem[];
    float2 *s_a = reinterpret_cast<float2 *>(wscan_smem);
    double *s_pq_re = reinterpret_cast<double *>(s_a + kWScanRows * kWScanPitch);   // exclusive prefix sums over rows: [0 .. kWScanRows]
    double *s_pq_im = s_pq_re + kWScanPre;
    double *s_pe = s_pq_im + kWScanPre;
    float *s_absq = reinterpret_cast<float *>(s_pe + kWScanPre);                  // per row: sum of |Re q| + |Im q| (>= sum |q|)
    float *s_etot = s_absq + kWScanPre;                                           // per row: sum of e
    auto at = [&](int i) -> float2 & { return s_a[i + (i >> 3)]; };               // sample i of the staged span (padded rows)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tile = (long long)a.tile_first + blockIdx.x;
    const long long n = (long long)a.n;
    const long long origin = tile * kWScanD - 8;                                  // sample index of row 0, column 0
    const long long d_last = n - 2 * L;                                           // last lag whose windows lie inside the capture

    if ((reinterpret_cast<uintptr_t>(a.iq) & 15) == 0 && origin >= 0 && origin + kWScanSamples <= n) {
        // (origin is a multiple of 8 samples: 16-byte loads, two samples each)
        const float4 *src = reinterpret_cast<const float4 *>(a.iq + origin);
        for (int i = tid; i < kWScanSamples / 2; i += kWScanThreads) {
            const float4 v = __ldg(src + i);
            at(2 * i) = make_float2(v.x, v.y); at(2 * i + 1) = make_float2(v.z, v.w);
        }
    } else {
        for (int i = tid; i < kWScanSamples; i += kWScanThreads) {
            const long long s = origin + i;
            at(i) = (s >= 0 && s < n) ? __ldg(a.iq + s) : make_float2(0.0f, 0.0f);
        }
    }
    __syncthreads();
    // ---- row totals (fp32 inside a row, f64 from there on) ----------------------------------------------------------------
    for (int r = tid; r < kWScanRows; r += kWScanThreads) {
        float qr = 0.0f, qi = 0.0f, e = 0.0f, aq = 0.0f;
        const bool has_q = r + kWScanR < kWScanRows;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 u = s_a[kWScanPitch * r + j];
            e += u.x * u.x + u.y * u.y;
            if (has_q) {
                const float2 v = s_a[kWScanPitch * (r + kWScanR) + j];
                const float a0 = u.x * v.x + u.y * v.y, a1 = u.x * v.y - u.y * v.x;   // conj(u) v
                qr += a0; qi += a1;
                aq += fabsf(a0) + fabsf(a1);
            }
        }
        s_pq_re[r + 1] = (double)qr; s_pq_im[r + 1] = (double)qi; s_pe[r + 1] = (double)e;
        s_absq[r] = aq; s_etot[r] = e;
    }
    __syncthreads();
    // ---- inclusive scan of the three arrays, one warp each: entry r + 1 becomes the sum of rows 0 .. r. A lane sums its own
    // run of 19 consecutive rows serially, the 32 run totals are scanned with shuffles, and the lane adds its offset.
    if (warp < 3) {
        double *p = warp == 0 ? s_pq_re : warp == 1 ? s_pq_im : s_pe;
        constexpr int RUN = (kWScanRows + 31) / 32;                                // 19
        const int i0 = RUN * lane;
        double run[RUN], tot = 0.0;
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            const int i = i0 + j;
            tot += i < kWScanRows ? p[i + 1] : 0.0;
            run[j] = tot;
        }
        double incl = tot;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, incl, m);
            if (lane >= m) incl += up;
        }
        const double off = incl - tot;
        if (lane == 0) p[0] = 0.0;
#pragma unroll
        for (int j = 0; j < RUN; j++) {
            const int i = i0 + j;
            if (i < kWScanRows) p[i + 1] = off + run[j];
        }
    }
    __syncthreads();
    // ---- the 8 lags of row r = tid + 1: d = tile * 4096 + 8 tid + k ---------------------------------------------------------
    const int r = tid + 1;
    const long long d0 = tile * kWScanD + 8 * tid;
    if (d0 > d_last) return;
    float pr = (float)(s_pq_re[r + kWScanR] - s_pq_re[r]), pi = (float)(s_pq_im[r + kWScanR] - s_pq_im[r]);
    float r1 = (float)(s_pe[r + kWScanR] - s_pe[r]), r2 = (float)(s_pe[r + 2 * kWScanR] - s_pe[r + kWScanR]);
    {
        // Most rows cannot hold a lag above the threshold -- noise, payload symbols: everything but the preamble plateaus --
        // and skip the per-lag work without changing a decision: over the row's 8 lags |P| <= |W_q| + sum |q| of the rows
        // that slide out and in, R1 >= W_e - (e of the row that slides out), R2 likewise (0.1 % slack for the fp32 sums).
        const float bq = sqrtf(pr * pr + pi * pi) + s_absq[r] + s_absq[r + kWScanR];
        const float r1min = fmaxf(r1 - s_etot[r], 0.0f), r2min = fmaxf(r2 - s_etot[r + kWScanR], 0.0f);
        if (bq * bq * 1.001f < 0.5f * r1min * r2min * 0.999f) return;
    }
    auto xs = [&](int i) -> float2 { return at(8 * r + i); };                     // a[d0 + i]
    auto qv = [&](int i, float &re, float &im) {                                  // q at sample d0 + i
        const float2 u = xs(i), v = xs(i + L);
        re = u.x * v.x + u.y * v.y; im = u.x * v.y - u.y * v.x;
    };
    auto ev = [&](int i) -> float { const float2 u = xs(i); return u.x * u.x + u.y * u.y; };
    bool prev;
    {   // the lag before this row's first: windows slid back by one sample
        float qa, qb, qc, qd;
        qv(-1, qa, qb); qv(L - 1, qc, qd);
        const float mr = pr + qa - qc, mi = pi + qb - qd;
        const float m1 = r1 + ev(-1) - ev(L - 1), m2 = r2 + ev(L - 1) - ev(2 * L - 1);
        prev = d0 > 0 && mr * mr + mi * mi > 0.5f * m1 * m2;                      // the search starts at lag 0 with "below"
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const long long d = d0 + k;
        const bool above = pr * pr + pi * pi > 0.5f * r1 * r2;
        if (above && !prev && d <= d_last) {
            const uint32_t slot = atomicAdd(a.tile_cnt + tile, 1u);
            if (slot < (uint32_t)kTileCand) a.tile_cand[(size_t)tile * kTileCand + slot] = (uint16_t)(8 * tid + k);
        }
        prev = above;
        float qa, qb, qc, qd;
        qv(k, qa, qb); qv(k + L, qc, qd);
        pr += qc - qa; pi += qd - qb;
        const float e0 = ev(k), e1 = ev(k + L), e2 = ev(k + 2 * L);
        r1 += e1 - e0; r2 += e2 - e1;
    }
}

// one CTA per accepted detection: refinement + CFO (docs/SPEC.md 4, 5, 9)
template <int = 0>
__global__ void __launch_bounds__(wide::kThreads) wide_sync_refine_kernel(const SyncArgs a)
{
    constexpr int L = wide::kL, NT = wide::kThreads;
    __shared__ float s_lock[L];
    __shared__ float s_val[NT / 32];
    __shared__ int s_idx[NT / 32];
    __shared__ double s_acc[6 * (NT / 32)];
    const uint32_t i = blockIdx.x;
    if (i >= a.counters[1] || i >= a.max_peaks) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const wide::WideTables *tb = reinterpret_cast<const wide::WideTables *>(a.wtables);
    for (int j = tid; j < L; j += NT) s_lock[j] = tb->lock[j].x;
    __syncthreads();
    const long long d0 = (long long)a.sel[i], n = (long long)a.n;
    long long k_lo = d0 - (11 * L) / 5, k_hi = d0 + L / 5;
    if (k_lo < -(L - 1)) k_lo = -(L - 1);
    const long long org = k_lo < 0 ? 0 : k_lo;                        // lags relative to a window origin (captures beyond 2^31 samples)
    const long rel = a.lock_is_ramp ? wide::ramp_argmax_closed_w(a.iq + org, (long)(n - org), (long)(k_lo - org), (long)(k_hi - org), s_val, s_idx)
                                    : wide::ramp_argmax_w(a.iq + org, (long)(n - org), (long)(k_lo - org), (long)(k_hi - org), s_lock, s_val, s_idx);
    const long long offset = org + rel - 1;
    const bool ok = offset >= 0 && offset + 10 * L <= n;
    double acc[6] = { 0.0, 0.0, 0.0, 0.0, 0.0, 0.0 };
    for (int j = tid; j < L; j += NT) {
        const float2 u = a.iq[d0 + j], v = a.iq[d0 + L + j];
        acc[0] += (double)u.x * v.x + (double)u.y * v.y;             // P(d0)
        acc[1] += (double)u.x * v.y - (double)u.y * v.x;
        acc[2] += (double)v.x * v.x + (double)v.y * v.y;             // R2(d0)
        acc[5] += (double)u.x * u.x + (double)u.y * u.y;             // R1(d0)
        if (ok) {
            const float2 *x0 = a.iq + offset;
            const float2 r2 = x0[2 * L + j], r3 = x0[3 * L + j], r4 = x0[4 * L + j];
            acc[3] += (double)r2.x * r3.x + (double)r2.y * r3.y + (double)r3.x * r4.x + (double)r3.y * r4.y;
            acc[4] += (double)r2.x * r3.y - (double)r2.y * r3.x + (double)r3.x * r4.y - (double)r3.y * r4.x;
        }
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], m);
        if (lane == 0) s_acc[q * (NT / 32) + warp] = acc[q];
    }
    __syncthreads();
    if (tid == 0) {
        double tt[6];
        for (int q = 0; q < 6; q++) { tt[q] = 0.0; for (int w = 0; w < NT / 32; w++) tt[q] += s_acc[q * (NT / 32) + w]; }
        SyncPeak p;
        p.offset = ok ? (uint64_t)offset : ~0ull;
        p.f_delta = ok ? (float)(atan2(tt[4], tt[3]) / (double)L) : 0.0f;
        p.metric = ok ? (float)((tt[0] * tt[0] + tt[1] * tt[1]) / (tt[5] * tt[2])) : -1.0f;      // < 0 marks an unusable detection
        a.peaks[i] = p;
    }
}

}  // namespace ofdm
