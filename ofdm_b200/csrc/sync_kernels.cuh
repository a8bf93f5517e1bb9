// sync_kernels.cuh -- preamble search over one long capture (BASELINE.json configs[2]: 1e9-sample capture).
//
// The reference handles a long capture with one whole-capture FFT cross-correlation (src/receiver.rs:20-21,
// src/signals/mod.rs:186-217: three length-(2M-1) f64 transforms, ~96 GB of scratch at M = 1e9 -- not runnable). Here
// the capture is scanned once (8 B/sample) with the sliding Schmidl-Cox metric of docs/SPEC.md 4:
//
//   sync_scan_kernel   : one CTA per 3928 lags, persistent. The tile (512 rows of 8 samples) is staged by the TMA engine
//                        (cp.async.bulk.tensor, two 256-row boxes of a [rows][64 B] view of the capture, 64-byte swizzle, double
//                        buffered on mbarriers) -- no register round trip, and a thread reads a row as 4 conflict-free LDS.128;
//                        captures that are not 16-byte aligned take a generic path (coalesced loads -> padded rows).
//                        q[n] = conj(a[n]) a[n+80], e[n] = |a[n]|^2; per-row totals, 10-row window sums of the totals, and
//                        for the 8 lags of a row P(d) = W_q + pre_q(row+10) - pre_q(row), R1, R2 likewise from the
//                        thread-serial prefix sums inside rows t, t+10, t+20. A row whose bound
//                        (|W_q| + sum|q_t| + sum|q_t+10|)^2 stays below 0.5 min R1 min R2 cannot hold a lag above the
//                        threshold and skips the per-lag work (all but the preamble plateaus of a capture). Rising
//                        edges of |P|^2 > 0.5 R1 R2 go to the tile's own slots (16-bit lag offsets: no 2^32 limit).
//   sync_select_kernel : one CTA: edges concatenated in tile order (= ascending), 800-sample hold-off by chains, compaction.
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
constexpr int kScanRowStride = kScanT + 1;          // padded row length of the generic path: conflict-free row-per-thread LDS.64
constexpr int kScanTileBytes = kScanRows * kScanT * 8;       // 32 KB: one TMA-staged tile (two 256-row boxes)
constexpr int kTileCand = 12;                       // rising edges a 3928-lag tile can record (a frame is >= 880 samples: <= 5 starts per tile)
constexpr int kSyncHoldoff = 800;                   // lock + preamble + training: one detection per frame

struct SyncPeak {            // = ofdm_peak in include/ofdm_engine.h
    uint64_t offset;         // src/receiver.rs:21 rule: ramp lag - 1
    float f_delta;           // rad/sample, angle of the preamble sum / 80
    float metric;            // |P|^2 / R^2 at the detection lag
};

struct SyncArgs {
    const float2 *iq;
    uint64_t n;
    uint16_t *tile_cand;     // [n_tiles][kTileCand] rising edges of a tile as lag offsets inside the tile (unsorted)
    uint32_t *tile_cnt;      // [n_tiles] rising edges the tile found (may exceed kTileCand: overflow)
    uint64_t *ordered;       // [n_tiles * kTileCand] all recorded edges in ascending order
    uint8_t  *keep;          // [n_tiles * kTileCand] 1: edge survives the hold-off
    uint64_t *sel;           // [n_tiles * kTileCand] detections (kept edges) in ascending order
    uint32_t *counters;      // [0] threshold crossings found, [1] entries of peaks[] = min(detections, max_peaks), [2] detections,
                             // [3] tiles whose edges did not fit their kTileCand slots
    const RxTables *tables;
    SyncPeak *peaks;
    uint32_t max_peaks;
    uint32_t n_tiles;        // 3928-lag tiles of the capture
    uint32_t tile_first, tile_count;   // tiles this scan launch covers (the grid is persistent and strides over them)
    uint32_t tile_lags;      // lags per tile: kScanD (nfft = 64) or kWScanD (nfft = 1024, wide_sync_kernels.cuh)
    uint32_t holdoff;        // lock + preamble + training = 10 symbol lengths: one detection per frame
    const void *wtables;     // wide::WideTables (nfft = 1024)
    int32_t  lock_is_ramp;   // nfft = 1024: the built-in locking ramp (closed-form ramp correlation)
    uint32_t prefetch_ahead; // wide_scan_tma_kernel: L2-prefetch the tile this many tiles ahead (0: off)
};

// generic path: padded rows | row totals | energy windows | masks.  TMA path: 2 x 32 KB swizzled tiles (1024-byte aligned) + the same
constexpr size_t sync_scan_smem_bytes(bool tma)
{
    return 1024 + (tma ? 2 * (size_t)kScanTileBytes : (size_t)kScanRows * kScanRowStride * sizeof(float2)) +
           sizeof(float) * 5 * (kScanRows + 32) + sizeof(uint32_t) * (kScanRows + 8) + 64;        // totals + windows | masks | mbarriers
}

__device__ __forceinline__ cpx c_conj_mul(cpx a, cpx b)     // conj(a) * b
{
    float ar, ai, br, bi;
    c_split(a, ar, ai); c_split(b, br, bi);
    return c_fma2(c_make(bi, -br), c_make(ai, ai), c_mul2(b, c_make(ar, ar)));
}

// ---- generic staging (any 8-byte aligned capture): coalesced global loads -> registers -> padded smem rows -----------------
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

// one 8-sample row of the staged tile -> registers.
//   TMA path : rows are 64 dense bytes, the tensor map's 64-byte swizzle XORs the 16-byte chunk index with bits 1..2 of the row
//              (address bits 4..5 ^= bits 7..8), so the 4 LDS.128 of a thread's row are conflict-free across the lanes of a warp.
//   generic  : rows padded to 9 samples, 8 LDS.64.
template <bool TMA>
__device__ __forceinline__ void scan_row(uint32_t tile_saddr, int row, cpx (&x)[kScanT])
{
    if (TMA) {
        const uint32_t base = tile_saddr + (uint32_t)row * 64u, sw = ((uint32_t)(row >> 1) & 3u) << 4;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            unsigned long long v0, v1;
            asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "r"(base + (((uint32_t)c << 4) ^ sw)));
            x[2 * c].v = v0; x[2 * c + 1].v = v1;
        }
    } else {
        const uint32_t base = tile_saddr + (uint32_t)row * (kScanRowStride * 8);
#pragma unroll
        for (int j = 0; j < kScanT; j++) asm volatile("ld.shared.u64 %0, [%1];" : "=l"(x[j].v) : "r"(base + 8u * j));
    }
}

// TMA tile load: two 256-row boxes of the [rows][16 floats] view of the capture (rows outside the capture arrive as zeros)
__device__ __forceinline__ void scan_tile_tma(const void *tmap, uint32_t dst_saddr, uint32_t bar_saddr, long long row0)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_saddr), "r"((uint32_t)kScanTileBytes) : "memory");
#pragma unroll
    for (int h = 0; h < 2; h++)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(dst_saddr + (uint32_t)h * (kScanTileBytes / 2)), "l"(tmap), "r"(0), "r"((int)(row0 + 256 * h)), "r"(bar_saddr) : "memory");
}

struct __align__(64) ScanTensorMap { unsigned long long opaque[16]; };      // CUtensorMap (128 bytes), encoded by the host

template <bool TMA>
__global__ void __launch_bounds__(kScanRows, 2) sync_scan_kernel(const SyncArgs a, const __grid_constant__ ScanTensorMap tmap)
{
    extern __shared__ __align__(1024) uint8_t scan_smem[];
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(scan_smem) + 1023) & ~(uintptr_t)1023);
    float *s_off = reinterpret_cast<float *>(tiles + (TMA ? 2 * kScanTileBytes : kScanRows * kScanRowStride * 8));   // row totals: q.re | q.im | e | sum |q|
    float *s_we = s_off + 4 * (kScanRows + 32);                                               // 80-sample energy windows at row granularity
    constexpr int TQR = 0, TQI = kScanRows + 32, TE = 2 * (kScanRows + 32), TA = 3 * (kScanRows + 32);
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_we + (kScanRows + 32));
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_mask + kScanRows + 8);                   // TMA path: one mbarrier per tile buffer

    const int t = threadIdx.x;
    const long long n = (long long)a.n;
    const long long d_last = n - 2 * kSym;                            // largest valid lag
    const long long n_rows_full = n / kScanT;                         // rows the tensor map covers; the last n % 8 samples are patched in
    const uint32_t tiles_sa = smem_addr(tiles);

    // persistent CTA: the next tile is staged (TMA into the other buffer / global loads into registers) while this one is processed
    ScanTileRegs pre;
    uint32_t ph = 0;                                                  // bit b: parity of tile buffer b's mbarrier
    if (TMA) {
        if (t == 0) { mbar_init(s_bar, 1); mbar_init(s_bar + 1, 1); mbar_fence_init(); }
        __syncthreads();
        if (t == 0 && blockIdx.x < a.tile_count) scan_tile_tma(&tmap, tiles_sa, smem_addr(s_bar), ((long long)a.tile_first + blockIdx.x) * kScanEvalRows - 1);
    } else if (blockIdx.x < a.tile_count) scan_tile_load(a, (long long)a.tile_first + blockIdx.x, t, pre);
    int buf = 0;
    const long long tile_stop = (long long)a.tile_first + a.tile_count;
#pragma unroll 1
    for (long long tile = (long long)a.tile_first + blockIdx.x; tile < tile_stop; tile += gridDim.x, buf ^= 1) {
    const long long d_base = tile * kScanD;                           // first lag evaluated from this tile (row 1)
    const long long origin = d_base - kScanT;                         // sample index of row 0, column 0
    const uint32_t cur_sa = tiles_sa + (TMA ? (uint32_t)buf * kScanTileBytes : 0u);
    if (TMA) {
        const long long nxt = tile + gridDim.x;
        if (t == 0 && nxt < tile_stop)
            scan_tile_tma(&tmap, tiles_sa + (uint32_t)(buf ^ 1) * kScanTileBytes, smem_addr(s_bar + (buf ^ 1)), nxt * kScanEvalRows - 1);
        mbar_wait(s_bar + buf, (ph >> buf) & 1u);
        ph ^= 1u << buf;
        // the capture's last n % 8 samples lie past the tensor map's last full row: patch them into the staged tile
        const long long row_tail = n_rows_full - (origin / kScanT);   // tile row that holds them (origin is a multiple of 8 here)
        if ((n % kScanT) != 0 && row_tail >= 0 && row_tail < kScanRows) {
            if (t < (int)(n % kScanT)) {
                const uint32_t sw = ((uint32_t)(row_tail >> 1) & 3u) << 4;
                const uint32_t addr = cur_sa + (uint32_t)row_tail * 64u + ((((uint32_t)t >> 1) << 4) ^ sw) + ((uint32_t)t & 1u) * 8u;
                const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long *>(a.iq + n_rows_full * kScanT + t));
                asm volatile("st.shared.u64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
            }
            __syncthreads();
        }
    } else {
        unsigned long long *s_iq = reinterpret_cast<unsigned long long *>(tiles);
#pragma unroll
        for (int i = 0; i < kScanT / 2; i++) {
            const int idx = i * kScanRows + t;                             // float4 index inside the tile
            const int row = idx / (kScanT / 2), col = (idx % (kScanT / 2)) * 2;
            s_iq[row * kScanRowStride + col] = pre.v[2 * i];
            s_iq[row * kScanRowStride + col + 1] = pre.v[2 * i + 1];
        }
        if (tile + gridDim.x < tile_stop) scan_tile_load(a, tile + gridDim.x, t, pre);
        __syncthreads();
    }

    // ---- row totals of q, e and of |q| (L1 norm, for the bound) -- one row per thread ---------------------------------
    // The running sums go through exactly the operations the per-lag prefix sums below use, so a total equals the prefix
    // "after column 7" bit for bit. Rows >= 502 have no partner row inside the tile; their q totals are never used (an
    // evaluated row t <= 491 sums the totals of rows t .. t+9 and looks at row t+10), so they just read a clamped row.
    cpx tq = c_make(0.0f, 0.0f), te2 = c_make(0.0f, 0.0f);
    float ta = 0.0f;
    {
        cpx own[kScanT], nxt[kScanT];
        scan_row<TMA>(cur_sa, t, own);
        scan_row<TMA>(cur_sa, t + kScanR80 < kScanRows ? t + kScanR80 : t, nxt);
#pragma unroll
        for (int j = 0; j < kScanT; j++) {
            float pr, pi;
            const cpx q = c_conj_mul(own[j], nxt[j]);
            c_split(q, pr, pi);
            tq = c_add(tq, q);
            te2 = c_fma2(own[j], own[j], te2);                             // (sum re^2, sum im^2)
            ta += fabsf(pr) + fabsf(pi);
        }
    }
    float qr, qi, te;
    {
        float er, ei;
        c_split(tq, qr, qi);
        c_split(te2, er, ei);
        te = er + ei;
    }
    // ---- 80-sample window sums at row granularity: W[t] = sum of the row totals of rows t .. t+9 ---------------------
    // (summing the ten small row totals directly, instead of differencing a tile-wide prefix sum, keeps fp32 exact enough in
    // a quiet stretch that follows a loud frame inside the same tile). Rows past the tile are never used by an evaluated lag
    // (each array has 32 spare entries behind it).
    // (one float array per quantity: thirty LDS.32 measured 27 % faster than ten LDS.128 of a float4 per row)
    float wqr = 0.0f, wqi = 0.0f, we = 0.0f;
    s_off[TQR + t] = qr; s_off[TQI + t] = qi; s_off[TE + t] = te; s_off[TA + t] = ta;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kScanR80; k++) { wqr += s_off[TQR + t + k]; wqi += s_off[TQI + t + k]; we += s_off[TE + t + k]; }
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
            cpx x0[kScanT], x5[kScanT], x10[kScanT];
            scan_row<TMA>(cur_sa, t, x0);
            scan_row<TMA>(cur_sa, t + A, x5);
            scan_row<TMA>(cur_sa, t + B, x10);
            // lags of this row that exist: 0 <= d <= d_last (hoisted out of the loop as a bit mask)
            const long long dr = origin + (long long)t * kScanT;
            uint32_t live = (1u << kScanT) - 1u;
            if (dr < 0) live = dr <= -kScanT ? 0u : live & ~((1u << (int)(-dr)) - 1u);
            if (dr + kScanT - 1 > d_last) live = dr > d_last ? 0u : live & ((1u << (int)(d_last - dr + 1)) - 1u);
            cpx q0 = c_make(0.0f, 0.0f), q5 = q0;                          // exclusive prefix sums inside rows t, t+A (q) and t, t+A, t+B (e)
            cpx e0 = q0, e5 = q0, e10 = q0;                                // (sum re^2, sum im^2), as in the row totals
#pragma unroll
            for (int j = 0; j < kScanT; j++) {
                float pr, pi, a0, b0, a5, b5, a10, b10;
                c_split(c_add(c_sub(q5, q0), dq), pr, pi);
                c_split(e0, a0, b0); c_split(e5, a5, b5); c_split(e10, a10, b10);
                const float s0 = a0 + b0, s5 = a5 + b5, s10 = a10 + b10;
                const float r1 = (s5 - s0) + de1, r2 = (s10 - s5) + de2;
                if (pr * pr + pi * pi > 0.5f * r1 * r2) mask |= 1u << j;
                q0 = c_add(q0, c_conj_mul(x0[j], x5[j]));
                q5 = c_add(q5, c_conj_mul(x5[j], x10[j]));
                e0 = c_fma2(x0[j], x0[j], e0);
                e5 = c_fma2(x5[j], x5[j], e5);
                e10 = c_fma2(x10[j], x10[j], e10);
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
            const uint32_t slot = atomicAdd(a.tile_cnt + tile, 1u);                 // this tile's own slots: order across tiles is implicit
            if (slot < (uint32_t)kTileCand) a.tile_cand[(size_t)tile * kTileCand + slot] = (uint16_t)((t - 1) * kScanT + j);
        }
    }
    // no barrier here: the tile buffer and s_off / s_we were last read before the barrier above, and the next iteration's
    // s_mask stores come after two more barriers that every thread still reading s_mask[t - 1] has yet to reach
    }
}

// One CTA: (1) the tiles' edges, sorted inside every tile, are concatenated in tile order = ascending lag order; (2) an edge
// that follows its predecessor by >= 800 samples starts a new chain and is always kept, the chains (short) are walked
// greedily -- exactly the sequential "keep d if d >= last kept + 800" rule; (3) the kept edges are compacted in order.
// Every phase is a strided loop over tiles or edges, so there is no capacity limit beyond the per-tile slots.
constexpr int kSelThreads = 1024;
template <int = 0>
__global__ void __launch_bounds__(kSelThreads) sync_select_kernel(const SyncArgs a)
{
    __shared__ unsigned long long s_scan[kSelThreads];
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x;
    // exclusive block scan of one value per thread (Hillis-Steele in shared memory; 1024 threads)
    auto block_exclusive_scan = [&](unsigned long long v) -> unsigned long long {
        s_scan[tid] = v;
        __syncthreads();
        for (int off = 1; off < kSelThreads; off <<= 1) {
            const unsigned long long add = tid >= off ? s_scan[tid - off] : 0ull;
            __syncthreads();
            s_scan[tid] += add;
            __syncthreads();
        }
        const unsigned long long incl = s_scan[tid];
        if (tid == kSelThreads - 1) s_total = incl;
        __syncthreads();
        return incl - v;
    };
    // ---- (1) ordered list ---------------------------------------------------------------------------------------------------
    const uint32_t n_tiles = a.n_tiles;
    const uint32_t chunk = (n_tiles + kSelThreads - 1) / kSelThreads;                 // consecutive tiles per thread
    const uint32_t t_lo = min((uint32_t)tid * chunk, n_tiles), t_hi = min(t_lo + chunk, n_tiles);
    unsigned long long mine = 0, crossings = 0, overflowed = 0;
    for (uint32_t tile = t_lo; tile < t_hi; tile++) {
        const uint32_t c = a.tile_cnt[tile];
        crossings += c;
        overflowed += c > (uint32_t)kTileCand;
        mine += c < (uint32_t)kTileCand ? c : (uint32_t)kTileCand;
    }
    unsigned long long pos = block_exclusive_scan(mine);
    const unsigned long long M = s_total;
    for (uint32_t tile = t_lo; tile < t_hi; tile++) {
        uint32_t c = a.tile_cnt[tile];
        if (c > (uint32_t)kTileCand) c = kTileCand;
        uint16_t v[kTileCand];
        for (uint32_t k = 0; k < c; k++) {                                            // insertion sort of <= 12 offsets
            const uint16_t x = a.tile_cand[(size_t)tile * kTileCand + k];
            uint32_t j = k;
            while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; j--; }
            v[j] = x;
        }
        for (uint32_t k = 0; k < c; k++) a.ordered[pos++] = (unsigned long long)tile * a.tile_lags + v[k];
    }
    const unsigned long long cr_before = block_exclusive_scan(crossings);
    (void)cr_before;
    const unsigned long long crossings_total = s_total;
    const unsigned long long ov_before = block_exclusive_scan(overflowed);
    (void)ov_before;
    const unsigned long long overflow_total = s_total;
    __threadfence_block();
    __syncthreads();
    // ---- (2) hold-off: chains ---------------------------------------------------------------------------------------------
    const unsigned long long holdoff = a.holdoff;
    for (unsigned long long i = tid; i < M; i += kSelThreads) {
        const bool head = i == 0 || a.ordered[i] - a.ordered[i - 1] >= holdoff;
        if (!head) continue;
        unsigned long long last = a.ordered[i];
        a.keep[i] = 1;
        for (unsigned long long j = i + 1; j < M; j++) {
            const unsigned long long d = a.ordered[j];
            if (d - a.ordered[j - 1] >= holdoff) break;                                // the next chain's head
            const bool k = d >= last + holdoff;
            a.keep[j] = k ? 1 : 0;
            if (k) last = d;
        }
    }
    __threadfence_block();
    __syncthreads();
    // ---- (3) compaction in order ------------------------------------------------------------------------------------------------
    const unsigned long long echunk = (M + kSelThreads - 1) / kSelThreads;
    const unsigned long long e_lo = min((unsigned long long)tid * echunk, M), e_hi = min(e_lo + echunk, M);
    unsigned long long kept = 0;
    for (unsigned long long i = e_lo; i < e_hi; i++) kept += a.keep[i];
    unsigned long long out = block_exclusive_scan(kept);
    const unsigned long long m = s_total;
    for (unsigned long long i = e_lo; i < e_hi; i++) if (a.keep[i]) a.sel[out++] = a.ordered[i];
    if (tid == 0) {
        a.counters[0] = crossings_total > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)crossings_total;
        a.counters[2] = m > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)m;                 // detections before truncation to max_peaks
        a.counters[1] = m < a.max_peaks ? (uint32_t)m : a.max_peaks;                   // entries of peaks[] (what *n_peaks reports)
        a.counters[3] = (uint32_t)overflow_total;
    }
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
    const long long d0 = (long long)a.sel[i], n = (long long)a.n;
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
