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

// Inclusive prefix sums over the rows of the three f64 arrays (entry r + 1 := sum of rows 0 .. r, entry 0 := 0), by the whole
// CTA: a thread owns RPT consecutive rows, the thread totals are scanned with warp shuffles, the warp totals through shared
// memory. (With three warps scanning one array each, the other warps of the CTA waited ~2 000 cycles at the next barrier.)
template <int ROWS, int THREADS>
__device__ __forceinline__ void wide_scan_prefix3(double *p0, double *p1, double *p2, double *s_wsum /* [3][THREADS / 32] */, int tid)
{
    constexpr int RPT = (ROWS + THREADS - 1) / THREADS, NW = THREADS / 32;
    const int lane = tid & 31, warp = tid >> 5, i0 = RPT * tid;
    double *p[3] = { p0, p1, p2 };
    double loc[3][RPT], tot[3], incl[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < RPT; j++) { t += i0 + j < ROWS ? p[c][i0 + j + 1] : 0.0; loc[c][j] = t; }
        tot[c] = t; incl[c] = t;
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double up = __shfl_up_sync(0xffffffffu, incl[c], m);
            if (lane >= m) incl[c] += up;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < 3; c++) s_wsum[c * NW + warp] = incl[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double off = incl[c] - tot[c];
        for (int w = 0; w < warp; w++) off += s_wsum[c * NW + w];
#pragma unroll
        for (int j = 0; j < RPT; j++) if (i0 + j < ROWS) p[c][i0 + j + 1] = off + loc[c][j];
        if (tid == 0) p[c][0] = 0.0;
    }
    __syncthreads();
}

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
    __shared__ double s_wsum[3 * (kWScanThreads / 32)];
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
    wide_scan_prefix3<kWScanRows, kWScanThreads>(s_pq_re, s_pq_im, s_pe, s_wsum, tid);
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

// ---- the same scan with the tile staged by the TMA engine (16-byte aligned captures) ----------------------------------------
// The capture is the [rows][64 B] tensor of the nfft = 64 scan (8 fc32 samples per row, 64-byte swizzle); a tile is three
// 256-row boxes = 768 rows: 447 rows of lags + 2 x 160 halo rows + the row that feeds above(d - 1). No register round trip,
// no shared stores, and a thread reads a row as 4 conflict-free LDS.128 (scan_row<true>). Three CTAs per SM: while one waits
// for its boxes the others reduce theirs.
#ifndef WT_BOXES
#define WT_BOXES 3
#endif
#ifndef WT_CTAS
#define WT_CTAS 3
#endif
constexpr int kWTBoxes = WT_BOXES, kWTCtasPerSm = WT_CTAS;
constexpr int kWTRows = 256 * kWTBoxes;
constexpr int kWTLagRows = kWTRows - 2 * kWScanR - 1;           // 447
constexpr int kWTD = kWTLagRows * 8;                            // 3576 lags per tile
constexpr int kWTThreads = (kWTLagRows + 31) / 32 * 32;         // >= kWTLagRows, whole warps (448 for three boxes)
constexpr int kWTTileBytes = kWTRows * 64;
constexpr int kWTPre = kWTRows + 3;
constexpr size_t wide_scan_tma_smem_bytes() { return 1024 + (size_t)kWTTileBytes + 3 * sizeof(double) * kWTPre + 2 * sizeof(float) * kWTPre + 32; }

template <int = 0>
__global__ void __launch_bounds__(kWTThreads, WT_CTAS) wide_scan_tma_kernel(const SyncArgs a, const __grid_constant__ ScanTensorMap tmap)
{
    constexpr int L = wide::kL;
    extern __shared__ __align__(1024) uint8_t wscan_tma_smem[];
    uint8_t *tile_p = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(wscan_tma_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(tile_p + kWTTileBytes);
    double *s_pq_re = reinterpret_cast<double *>(s_bar + 2);
    double *s_pq_im = s_pq_re + kWTPre;
    double *s_pe = s_pq_im + kWTPre;
    float *s_absq = reinterpret_cast<float *>(s_pe + kWTPre);
    float *s_etot = s_absq + kWTPre;
    __shared__ double s_wsum[3 * (kWTThreads / 32)];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tile = (long long)a.tile_first + blockIdx.x;
    const long long n = (long long)a.n;
    const long long origin = tile * kWTD - 8;                                      // sample index of row 0, column 0 (a multiple of 8)
    const long long d_last = n - 2 * L;
    const uint32_t tile_sa = smem_addr(tile_p);
    if (tid == 0) { mbar_init(s_bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bar_sa = smem_addr(s_bar);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_sa), "r"((uint32_t)kWTTileBytes) : "memory");
#pragma unroll
        for (int h = 0; h < kWTBoxes; h++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(tile_sa + (uint32_t)h * (256 * 64)), "l"(&tmap), "r"(0), "r"((int)(origin / 8 + 256 * h)), "r"(bar_sa) : "memory");
    }
    // One tile per CTA and no second buffer: what hides the next tile's trip to DRAM is an L2 prefetch of the tile this SM slot
    // will most likely run next (CTAs are dispatched in order, 3 per SM), issued while this one's boxes are still in flight.
    if (tid == 32 && a.prefetch_ahead) {
        const long long nxt = tile + a.prefetch_ahead;
        if (nxt < (long long)a.tile_first + a.tile_count) {
#pragma unroll
            for (int h = 0; h < kWTBoxes; h++)
                asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
                             :: "l"(&tmap), "r"(0), "r"((int)(nxt * kWTLagRows - 1 + 256 * h)) : "memory");
        }
    }
    mbar_wait(s_bar, 0);
    {   // the capture's last n % 8 samples lie past the tensor map's last full row: patch them into the staged tile
        const long long n_rows_full = n / 8, row_tail = n_rows_full - origin / 8;
        if ((n % 8) != 0 && row_tail >= 0 && row_tail < kWTRows) {
            if (tid < (int)(n % 8)) {
                const uint32_t sw = ((uint32_t)(row_tail >> 1) & 3u) << 4;
                const uint32_t addr = tile_sa + (uint32_t)row_tail * 64u + ((((uint32_t)tid >> 1) << 4) ^ sw) + ((uint32_t)tid & 1u) * 8u;
                const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long *>(a.iq + n_rows_full * 8 + tid));
                asm volatile("st.shared.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
            }
            __syncthreads();
        }
    }
    auto smp = [&](int i) -> float2 {                                             // sample i of the staged span (swizzled rows)
        const uint32_t row = (uint32_t)i >> 3, c = (uint32_t)i & 7u;
        unsigned long long v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(tile_sa + row * 64u + (((c >> 1) << 4) ^ (((row >> 1) & 3u) << 4)) + (c & 1u) * 8u));
        cpx t; t.v = v;
        return c_to(t);
    };
    // ---- row totals ------------------------------------------------------------------------------------------------------------
    for (int r = tid; r < kWTRows; r += kWTThreads) {
        cpx own[8], nxt[8];
        const bool has_q = r + kWScanR < kWTRows;
        scan_row<true>(tile_sa, r, own);
        scan_row<true>(tile_sa, has_q ? r + kWScanR : r, nxt);
        float qr = 0.0f, qi = 0.0f, e = 0.0f, aq = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 u = c_to(own[j]), v = c_to(nxt[j]);
            e += u.x * u.x + u.y * u.y;
            const float a0 = u.x * v.x + u.y * v.y, a1 = u.x * v.y - u.y * v.x;   // conj(u) v
            qr += a0; qi += a1;
            aq += fabsf(a0) + fabsf(a1);
        }
        if (!has_q) { qr = 0.0f; qi = 0.0f; aq = 0.0f; }
        s_pq_re[r + 1] = (double)qr; s_pq_im[r + 1] = (double)qi; s_pe[r + 1] = (double)e;
        s_absq[r] = aq; s_etot[r] = e;
    }
    __syncthreads();
    wide_scan_prefix3<kWTRows, kWTThreads>(s_pq_re, s_pq_im, s_pe, s_wsum, tid);
    // ---- the 8 lags of row r = tid + 1 -------------------------------------------------------------------------------------------
    if (tid >= kWTLagRows) return;
    const int r = tid + 1;
    const long long d0 = tile * kWTD + 8 * tid;
    if (d0 > d_last) return;
    float pr = (float)(s_pq_re[r + kWScanR] - s_pq_re[r]), pi = (float)(s_pq_im[r + kWScanR] - s_pq_im[r]);
    float r1 = (float)(s_pe[r + kWScanR] - s_pe[r]), r2 = (float)(s_pe[r + 2 * kWScanR] - s_pe[r + kWScanR]);
    {
        const float bq = sqrtf(pr * pr + pi * pi) + s_absq[r] + s_absq[r + kWScanR];
        const float r1min = fmaxf(r1 - s_etot[r], 0.0f), r2min = fmaxf(r2 - s_etot[r + kWScanR], 0.0f);
        if (bq * bq * 1.001f < 0.5f * r1min * r2min * 0.999f) return;             // no lag of this row can cross (see wide_scan_kernel)
    }
    auto xs = [&](int i) -> float2 { return smp(8 * r + i); };
    auto qv = [&](int i, float &re, float &im) {
        const float2 u = xs(i), v = xs(i + L);
        re = u.x * v.x + u.y * v.y; im = u.x * v.y - u.y * v.x;
    };
    auto ev = [&](int i) -> float { const float2 u = xs(i); return u.x * u.x + u.y * u.y; };
    bool prev;
    {
        float qa, qb, qc, qd;
        qv(-1, qa, qb); qv(L - 1, qc, qd);
        const float mr = pr + qa - qc, mi = pi + qb - qd;
        const float m1 = r1 + ev(-1) - ev(L - 1), m2 = r2 + ev(L - 1) - ev(2 * L - 1);
        prev = d0 > 0 && mr * mr + mi * mi > 0.5f * m1 * m2;
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
