// tx_kernels.cuh -- transmit path (replaces `encode`, src/transmitter.rs:11-58), the synthetic channel harness
// (src/channel.rs:33-74) and the BER counters (utils::Analysis, src/utils.rs:45-68).
#pragma once

#include "common.cuh"
#include "rx_kernels.cuh"

namespace ofdm {

struct TxArgs {
    const uint8_t  *payload;
    const uint32_t *payload_len;
    uint32_t        payload_stride;
    uint32_t        n_streams;
    float2         *iq;
    uint32_t        iq_stride;
    uint32_t       *frame_len;      // optional
    int            *stream_max;     // per stream: max positive component as float bits (atomicMax on int)
    int32_t         tile_shift;     // Hamming byte alignment of the tile boundaries (same as the decode kernel)
    int32_t         tiles_per_cta;  // consecutive tiles of one frame handled by one CTA
    const RxTables *tables;
    uint32_t        stream0;        // first stream of this launch (gridDim.y <= 65535 streams per launch)
    uint32_t       *stream_cnt;     // tx_resident_kernel: per stream, compute warps that have published their maximum
    int32_t         group_ctas;     // tx_resident_kernel: CTAs sharing one frame
    int32_t         n_groups;       // tx_resident_kernel: groups of the (persistent) grid
    int32_t         redo_only;      // tx_tile_kernel<WRITE>: redo pass behind tx_spec_kernel (frames whose stream_max is set, if stream_cnt[0] != 0)
};

// byte `B` of the frame byte stream [header 16 B | (Hamming-coded) payload ...] (src/transmitter.rs:37-47, docs/SPEC.md 3)
template <bool FEC>
__device__ __forceinline__ uint32_t frame_byte(const uint8_t *__restrict__ pay, uint32_t n, uint64_t coded_len, uint32_t B, const uint8_t *s_enc)
{
    if (B < 16) return B < 8 ? (uint32_t)(coded_len >> (8 * B)) & 255u : 0u;     // bincode u128 LE: low 64 bits = length
    const uint32_t c = B - 16;
    if (!FEC) return c < n ? pay[c] : 0u;
    // bits [8c, 8c + 8) of the packed 7-bit codeword stream; codeword m encodes nibble m (byte m >> 1, low nibble first)
    const uint32_t m0 = (8u * c) / 7u, r = 8u * c - 7u * m0;
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const uint32_t m = m0 + q, b = m >> 1;
        const uint32_t nib = b < n ? ((uint32_t)pay[b] >> (4 * (m & 1))) & 15u : 0u;
        w |= (b < n ? (uint32_t)s_enc[nib] : 0u) << (7 * q);
    }
    return (w >> r) & 255u;
}

constexpr int kTxWarps = 8;
constexpr int kTxThreads = kTxWarps * 32;
constexpr int kTxIters = 7;
constexpr int kTxTileSyms = kTxWarps * 4 * kTxIters;      // 224, same tiling as the decode kernel

// dynamic shared memory of tx_tile_kernel: transpose scratch (the tile's packed bit stream lives in it until it has been
// unpacked) | one byte per data carrier | per-lane constellation table | Hamming encode tables
template <int MOD, bool GUARD> constexpr size_t tx_smem_bytes()
{
    return sizeof(float2) * kTxWarps * kTrWarp + (size_t)kTxTileSyms * (GUARD ? 48 : 64) + 64 +
           sizeof(float2) * 16 * ((1 << ModTraits<MOD>::kBpc) + 2) + 16 + 512;
}

// One 224-symbol tile of one frame: (Hamming) coded bit stream of the tile -> one byte per data carrier in smem, then per OFDM
// symbol (8 lanes): carrier byte -> constellation point -> inverse FFT (packed FFMA2 transform on re/im-swapped data) -> CP.
// The constellation table holds every entry once per lane of a half-warp ([entry][lane & 15], LDS.64 at bank 2 (lane & 15):
// conflict-free whatever the lanes look up); null and pilot bins are two more entries. A CTA owns `tiles_per_cta` consecutive
// tiles of one frame (tables are built once).
// WRITE = false: only the per-stream maximum positive component is produced (`normalize`, src/transmitter.rs:183-194);
// WRITE = true : the symbols are recomputed and stored once, already normalised, together with the frame head and
// the zero fill -- 8 B/sample of HBM traffic in total instead of write + read-modify-write.
// the tiles [bx * tiles_per_cta, ...) of one frame; (bx, nbx) = the kernel's block index and grid extent in x, (0, 1) in the redo pass
template <int MOD, bool GUARD, bool FEC, bool WRITE>
__device__ __forceinline__ void tx_tile_body(const TxArgs &a, const uint32_t stream, const uint32_t bx, const uint32_t nbx)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 48 : 64;
    constexpr int BPS = BPC * D;
    constexpr int NE = 1 << BPC;                        // constellation entries; NE = null carrier, NE + 1 = pilot
    extern __shared__ __align__(128) uint8_t tx_smem[];
    float2 *s_tr = reinterpret_cast<float2 *>(tx_smem);
    uint8_t *s_bits = tx_smem;                                                   // aliases s_tr: consumed before the first transform
    uint8_t *s_car = reinterpret_cast<uint8_t *>(s_tr + kTxWarps * kTrWarp);    // one byte (BPC valid bits) per data carrier of the tile
    float2 *s_lut = reinterpret_cast<float2 *>(s_car + kTxTileSyms * D + 64);   // [NE + 2][16 lanes], stored re/im swapped
    uint8_t *s_enc = reinterpret_cast<uint8_t *>(s_lut + 16 * (NE + 2));
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(s_enc + 16);               // payload byte -> its two 7-bit codewords (low nibble first)
    static_assert(kTxTileSyms * BPS / 8 + 32 <= (int)sizeof(float2) * kTxWarps * kTrWarp, "the packed bit stream must fit the transpose scratch");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    const uint32_t n = a.payload_len[stream];
    const uint64_t coded_len = FEC ? (14ull * n + 7) / 8 : n;
    const uint64_t nbits = kHeaderBits + 8 * coded_len;
    const uint64_t ncar = (nbits + BPC - 1) / BPC;                 // constellation symbols (src/transmitter.rs:108-140)
    const int S = (int)((ncar + D - 1) / D);                       // OFDM data symbols (src/transmitter.rs:49-54)
    const uint32_t frame_len = (kHeadSyms + (uint32_t)S) * kSym;
    if (a.frame_len && bx == 0 && tid == 0) a.frame_len[stream] = frame_len;
    const bool fits = frame_len <= a.iq_stride;
    float2 *out = a.iq + (size_t)stream * a.iq_stride;

    const int n_tiles = (S + a.tile_shift + kTxTileSyms - 1) / kTxTileSyms;
    const int tile_first = (int)bx * a.tiles_per_cta;
    int tile_end = tile_first + a.tiles_per_cta;
    if (tile_end > n_tiles) tile_end = n_tiles;
    float scale = 1.0f / 64.0f;
    if (WRITE) {
        const float mx = fmaxf(__int_as_float(a.stream_max[stream]), a.tables->head_max);
        const float inv = 1.0f / mx;
        scale *= inv;
        // frame head (lock | preamble x4 | training x5) and zero fill past the frame
        if (bx == 0)
            for (uint32_t i = tid; i < (uint32_t)(kHeadSyms * kSym) && i < a.iq_stride; i += kTxThreads) {
                float2 v = make_float2(0.0f, 0.0f);
                if (fits) { v = a.tables->head[i]; v.x = v.x / mx; v.y = v.y / mx; }
                out[i] = v;
            }
        if (bx == nbx - 1) {
            const uint32_t z0 = fits ? frame_len : (uint32_t)(kHeadSyms * kSym);
            for (uint32_t i = z0 + tid; i < a.iq_stride; i += kTxThreads) out[i] = make_float2(0.0f, 0.0f);
        }
    }
    if (!fits || tile_first >= tile_end) return;

    // constellation table, stored re/im swapped (the inverse FFT runs as swap . FFT . swap), one copy per lane of a half-warp
    for (int e = tid; e < 16 * (NE + 2); e += kTxThreads) {
        const int idx = e >> 4;
        float re = 0.0f, im = 0.0f;
        if (idx == NE + 1) re = 1.0f;                                                 // pilot 1 + 0j (src/transmitter.rs:150-161)
        else if (idx == NE) { }                                                       // null carrier / padding
        else if (MOD == 0) { re = (idx & 1) ? 1.0f : -1.0f; }                         // src/transmitter.rs:112-118
        else if (MOD == 1) { re = (idx & 1) ? 1.0f : -1.0f; im = (idx & 2) ? 1.0f : -1.0f; }   // src/transmitter.rs:122-132
        else {
            const uint32_t ci = idx & 7u, cq = (uint32_t)idx >> 3;                     // Gray code -> level (docs/SPEC.md 2)
            const uint32_t li = ci ^ (ci >> 1) ^ (ci >> 2), lq = cq ^ (cq >> 1) ^ (cq >> 2);
            re = (2.0f * (float)li - 7.0f) * (1.0f / 7.0f);
            im = (2.0f * (float)lq - 7.0f) * (1.0f / 7.0f);
        }
        s_lut[e] = make_float2(im, re);
    }
    if (tid < 16) s_enc[tid] = (uint8_t)ham74_encode_nibble(tid);
    if (FEC) s_enc14[tid] = (uint16_t)(ham74_encode_nibble(tid & 15) | (ham74_encode_nibble(tid >> 4) << 7));

    cpx tw[8];
#pragma unroll
    for (int ka = 0; ka < 8; ka++) tw[ka] = c_from(__ldg(a.tables->w64 + ((l * ka) & 63)));
    // bins l, l + 24, l + 32, l + 56 are null / pilot carriers on some lanes: their table entry replaces the looked-up byte.
    // Carrier rank of bin l + 8 j is affine in j except at the pilot / DC crossings (j = 3, 4), as in the decode kernel.
    int d3 = 24 - (l >= 2), d4 = 31 - (l >= 1);                     // rank(l + 24) - (l - 7), rank(l + 32) - (l - 7)
    uint32_t fix0 = 0xFFu, fix3 = 0xFFu, fix4 = 0xFFu, fix7 = 0xFFu;      // 0xFF: data carrier, else the fixed table entry
    if (GUARD) {
        if (data_rank<GUARD>(l) < 0) fix0 = is_pilot_bin(l) ? NE + 1 : NE;
        if (data_rank<GUARD>(l + 24) < 0) fix3 = is_pilot_bin(l + 24) ? NE + 1 : NE;
        if (data_rank<GUARD>(l + 32) < 0) fix4 = is_pilot_bin(l + 32) ? NE + 1 : NE;
        if (data_rank<GUARD>(l + 56) < 0) fix7 = is_pilot_bin(l + 56) ? NE + 1 : NE;
    }
    asm volatile("" : "+r"(d3), "+r"(d4), "+r"(fix0), "+r"(fix3), "+r"(fix4), "+r"(fix7));
    float2 *tr = s_tr + warp * kTrWarp + g * kTrGroup;
    const unsigned long long *lut = reinterpret_cast<const unsigned long long *>(s_lut) + (lane & 15);
    const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
    float mx = 0.0f;

#pragma unroll 1
    for (int tile = tile_first; tile < tile_end; tile++) {
    int t0 = tile * kTxTileSyms - a.tile_shift, t1 = t0 + kTxTileSyms;
    if (t0 < 0) t0 = 0;
    if (t1 > S) t1 = S;
    __syncthreads();                                               // tables ready / the previous tile's transforms are done with s_tr

    // ---- tile bit stream --------------------------------------------------------------------------------------------
    const uint32_t byte0 = (uint32_t)((long)t0 * BPS / 8), nbyte = (uint32_t)((long)(t1 - t0) * BPS / 8);
    if (FEC) {
        // header bytes (tile 0), then groups of 16 payload bytes -> 32 codewords -> 28 coded bytes = 7 aligned words. Tile
        // boundaries fall on 4-byte group boundaries (tile_shift: (288 t0 - 128) is a multiple of 56), so every group of
        // four starts on a coded-byte boundary that is a multiple of 4 inside the tile.
        const uint32_t hdr = byte0 < 16 ? 16 - byte0 : 0;              // header bytes inside this tile (16 or 0)
        if (tid < hdr) s_bits[tid] = (uint8_t)frame_byte<FEC>(pay, n, coded_len, byte0 + tid, s_enc);
        const uint32_t c0 = byte0 + hdr - 16;                          // first coded byte of the tile: multiple of 7
        const uint32_t ngrp = ((nbyte + 2 - hdr + 6) / 7 + 3) / 4;     // groups of 16 payload bytes
        const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
        for (uint32_t u = tid; u < ngrp; u += kTxThreads) {
            const uint32_t pb = (c0 / 7) * 4 + 16 * u;                 // first payload byte of the group
            uint32_t v[4] = { 0, 0, 0, 0 };                            // bytes past the payload encode to zero codewords
            if (pay_aligned && pb + 16 <= n) {
#pragma unroll
                for (int q = 0; q < 4; q++) v[q] = __ldg(reinterpret_cast<const uint32_t *>(pay + pb) + q);
            } else {
#pragma unroll
                for (int q = 0; q < 16; q++) if (pb + q < n) v[q >> 2] |= (uint32_t)pay[pb + q] << (8 * (q & 3));
            }
            uint64_t w[4];                                             // 4 x 56 coded bits
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t lo = (uint32_t)s_enc14[v[q] & 255u] | ((uint32_t)s_enc14[(v[q] >> 8) & 255u] << 14);      // 28 bits
                const uint32_t hi = (uint32_t)s_enc14[(v[q] >> 16) & 255u] | ((uint32_t)s_enc14[v[q] >> 24] << 14);
                w[q] = (uint64_t)lo | ((uint64_t)hi << 28);
            }
            uint32_t *dst = reinterpret_cast<uint32_t *>(s_bits + hdr + 28 * u);
            dst[0] = (uint32_t)w[0];
            dst[1] = (uint32_t)(w[0] >> 32) | ((uint32_t)w[1] << 24);
            dst[2] = (uint32_t)(w[1] >> 8);
            dst[3] = (uint32_t)(w[1] >> 40) | ((uint32_t)w[2] << 16);
            dst[4] = (uint32_t)(w[2] >> 16);
            dst[5] = (uint32_t)(w[2] >> 48) | ((uint32_t)w[3] << 8);
            dst[6] = (uint32_t)(w[3] >> 24);
        }
    } else {
        for (uint32_t b = tid; b < nbyte + 2; b += kTxThreads) s_bits[b] = (uint8_t)frame_byte<FEC>(pay, n, coded_len, byte0 + b, s_enc);
    }
    __syncthreads();

    // ---- bit stream -> one byte per data carrier (modulate, src/transmitter.rs:108-140); carriers past the frame's last
    // constellation symbol are padding (encode_block's exhausted iterator, src/transmitter.rs:144-165) and get the null entry
    {
        const long ncar_local = (long)ncar - (long)t0 * D;            // carriers of the frame that exist from this tile on
        long have = (long)(t1 - t0) * D;                              // ... and inside this tile; rows past t1 are all padding
        if (ncar_local < have) have = ncar_local;
        const int ncar_have = (int)have;
        const uint32_t *bits32 = reinterpret_cast<const uint32_t *>(s_bits);
        for (int c4 = 4 * tid; c4 < kTxTileSyms * D; c4 += 4 * kTxThreads) {  // 4 carriers = 4 BPC bits from bit 4 BPC (c4 / 4)
            const uint32_t bit = (uint32_t)c4 * BPC, wi = bit >> 5, sh = bit & 31;
            const uint32_t v = __funnelshift_r(bits32[wi], bits32[wi + 1], sh);
            constexpr uint32_t M = (uint32_t)(NE - 1);
            uint32_t packed = (v & M) | (((v >> BPC) & M) << 8) | (((v >> (2 * BPC)) & M) << 16) | (((v >> (3 * BPC)) & M) << 24);
            if (c4 + 4 > ncar_have) {                                  // the frame's last carriers / padding rows (rare)
#pragma unroll
                for (int q = 0; q < 4; q++) if (c4 + q >= ncar_have) packed = (packed & ~(0xFFu << (8 * q))) | ((uint32_t)NE << (8 * q));
            }
            *reinterpret_cast<uint32_t *>(s_car + c4) = packed;
        }
    }
    __syncthreads();                                               // s_bits (= s_tr) is free for the transforms from here on

    const uint8_t *rowp = s_car + (warp * (4 * kTxIters) + g) * D + (GUARD ? l - 7 : l);      // this group's first symbol, biased by the lane
#pragma unroll 1
    for (int it = 0; it < kTxIters; it++) {
        const int sl = warp * (4 * kTxIters) + 4 * it + g;        // symbol index inside the tile
        const int s = t0 + sl;
        const bool valid = s < t1;
        cpx x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {                              // encode_block, src/transmitter.rs:144-165
            uint32_t idx;
            if (!GUARD) idx = rowp[8 * j];
            else if (j == 1 || j == 2) idx = rowp[8 * j];
            else if (j == 5 || j == 6) idx = rowp[8 * j - 3];
            else if (j == 3) idx = fix3 != 0xFFu ? fix3 : rowp[d3];
            else if (j == 4) idx = fix4 != 0xFFu ? fix4 : rowp[d4];
            else if (j == 0) idx = fix0 != 0xFFu ? fix0 : rowp[0];
            else idx = fix7 != 0xFFu ? fix7 : rowp[53];
            x[j].v = lut[idx * 16];
        }
        rowp += 4 * D;
        fft64_group_p(x, tw, tr, l);                               // prefix_block, src/transmitter.rs:168-181 (IFFT part)
        if (WRITE) {
            if (valid) {
                float2 *sym = out + (size_t)(kHeadSyms + s) * kSym;
#pragma unroll
                for (int kb = 0; kb < 8; kb++) {
                    float re, im;
                    c_split(x[kb], im, re);                        // un-swap
                    const float2 v = make_float2(re * scale, im * scale);
                    const int t = l + 8 * kb;                      // time index inside the symbol
                    sym[kCp + t] = v;
                    if (kb >= 6) sym[t - (kNfft - kCp)] = v;       // cyclic prefix = last 16 samples
                }
            }
        } else if (valid) {
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                float re, im;
                c_split(x[kb], im, re);
                mx = fmaxf(mx, fmaxf(re, im));
            }
        }
    }
    }
    if (!WRITE) {
        mx *= scale;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0 && mx > 0.0f) atomicMax(a.stream_max + stream, __float_as_int(mx));
    }
}

template <int MOD, bool GUARD, bool FEC, bool WRITE>
__global__ void __launch_bounds__(kTxThreads, 4) tx_tile_kernel(const TxArgs a)
{
    if (WRITE && a.redo_only) {
        // redo pass behind tx_spec_kernel: nothing to do unless the speculative kernel counted a frame whose data beat the head
        // maximum (stream_cnt[0]; for scrambled payloads: never) -- then the CTAs stride over the frames and rewrite, whole, those
        // whose stream_max is set
        if (a.stream_cnt[0] == 0) return;
        for (uint32_t s = blockIdx.y; s < a.n_streams; s += gridDim.y) {
            if (a.stream_max[s] == 0) continue;
            __syncthreads();                                        // (the previous frame's readers of the shared tables are done)
            tx_tile_body<MOD, GUARD, FEC, WRITE>(a, s, 0u, 1u);
        }
        return;
    }
    tx_tile_body<MOD, GUARD, FEC, WRITE>(a, blockIdx.y + a.stream0, blockIdx.x, gridDim.x);
}

// ------------------------------------------------------------------------------------------------------------------
// One-pass TX: a frame per thread-block cluster
// ------------------------------------------------------------------------------------------------------------------
// `normalize` (src/transmitter.rs:183-194) needs the frame's maximum before a single sample can be written, which is why
// tx_tile_kernel runs twice. Here a frame belongs to a CLUSTER of up to 16 CTAs: CTA r of the cluster transforms symbols
// [133 r - tile_shift, 133 (r + 1) - tile_shift) of the frame ONCE, keeps their 64 time-domain samples each (68 kB, un-normalised,
// cyclic prefixes not duplicated) in its own shared memory, publishes its maximum with one atomicMax, the cluster meets at
// ONE hardware cluster barrier, and every CTA then scales its symbols and streams them out (CP included) with coalesced
// 16-byte stores -- 8 B/sample of HBM traffic and one transform per symbol. Used for small batches (one wave of clusters,
// e.g. the reference's one-frame `encode`: 17 us instead of 22 us for a 163 840-sample frame); with thousands of frames the
// cluster barrier keeps the co-resident CTAs of an SM in lock step (all transform, then all store) and the two-pass kernel,
// whose store pass overlaps the two inside every SM, is faster (ofdm_engine.cu: tx_device). Frames longer than 16 x 133
// symbols always take the two-pass kernel.
constexpr int kTxfWarps = 7;
constexpr int kTxfThreads = kTxfWarps * 32;
constexpr int kTxfChunk = 133;                      // symbols per CTA (multiple of 7: Hamming byte alignment of the chunk boundaries)
constexpr int kTxfIters = (kTxfChunk + 4 * kTxfWarps - 1) / (4 * kTxfWarps);      // 5 warp iterations of 4 symbols
constexpr int kTxfMaxCluster = 16;
template <int MOD, bool GUARD> constexpr size_t txf_smem_bytes()
{
    return sizeof(float2) * (kTxfChunk * kNfft + kTxfWarps * kTrWarp) + (size_t)kTxfIters * 4 * kTxfWarps * (GUARD ? 48 : 64) + 64 +
           sizeof(float2) * 16 * ((1 << ModTraits<MOD>::kBpc) + 2) + 16 + 512;
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MOD, bool GUARD, bool FEC>
__global__ void __launch_bounds__(kTxfThreads, 2) tx_frame_kernel(const TxArgs a)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 48 : 64;
    constexpr int BPS = BPC * D;
    constexpr int NE = 1 << BPC;
    constexpr int SLOTS = kTxfIters * 4 * kTxfWarps;                             // symbol slots of a CTA (>= kTxfChunk)
    extern __shared__ __align__(128) uint8_t txf_smem[];
    float2 *s_out = reinterpret_cast<float2 *>(txf_smem);                        // [kTxfChunk][64]: un-normalised time-domain symbols
    float2 *s_tr = s_out + kTxfChunk * kNfft;
    uint8_t *s_bits = reinterpret_cast<uint8_t *>(s_tr);                         // aliases s_tr: consumed before the first transform
    uint8_t *s_car = reinterpret_cast<uint8_t *>(s_tr + kTxfWarps * kTrWarp);
    float2 *s_lut = reinterpret_cast<float2 *>(s_car + SLOTS * D + 64);
    uint8_t *s_enc = reinterpret_cast<uint8_t *>(s_lut + 16 * (NE + 2));
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(s_enc + 16);
    static_assert(SLOTS * BPS / 8 + 32 <= (int)sizeof(float2) * kTxfWarps * kTrWarp, "the packed bit stream must fit the transpose scratch");

    const uint32_t stream = blockIdx.y + a.stream0;
    const uint32_t rank = cluster_ctarank();                                     // = blockIdx.x: the grid's x extent is the cluster size
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    const uint32_t n = a.payload_len[stream];
    const uint64_t coded_len = FEC ? (14ull * n + 7) / 8 : n;
    const uint64_t nbits = kHeaderBits + 8 * coded_len;
    const uint64_t ncar = (nbits + BPC - 1) / BPC;
    const int S = (int)((ncar + D - 1) / D);
    const uint32_t frame_len = (kHeadSyms + (uint32_t)S) * kSym;
    if (a.frame_len && rank == 0 && tid == 0) a.frame_len[stream] = frame_len;
    const bool fits = frame_len <= a.iq_stride;
    float2 *out = a.iq + (size_t)stream * a.iq_stride;
    int t0 = (int)rank * kTxfChunk - a.tile_shift, t1 = t0 + kTxfChunk;
    if (t0 < 0) t0 = 0;
    if (t1 > S) t1 = S;
    const bool work = fits && t0 < t1;
    float mx = 0.0f;

    if (work) {
        for (int e = tid; e < 16 * (NE + 2); e += kTxfThreads) {                 // constellation table, see tx_tile_kernel
            const int idx = e >> 4;
            float re = 0.0f, im = 0.0f;
            if (idx == NE + 1) re = 1.0f;
            else if (idx == NE) { }
            else if (MOD == 0) { re = (idx & 1) ? 1.0f : -1.0f; }
            else if (MOD == 1) { re = (idx & 1) ? 1.0f : -1.0f; im = (idx & 2) ? 1.0f : -1.0f; }
            else {
                const uint32_t ci = idx & 7u, cq = (uint32_t)idx >> 3;
                const uint32_t li = ci ^ (ci >> 1) ^ (ci >> 2), lq = cq ^ (cq >> 1) ^ (cq >> 2);
                re = (2.0f * (float)li - 7.0f) * (1.0f / 7.0f);
                im = (2.0f * (float)lq - 7.0f) * (1.0f / 7.0f);
            }
            s_lut[e] = make_float2(im, re);
        }
        if (tid < 16) s_enc[tid] = (uint8_t)ham74_encode_nibble(tid);
        if (FEC) for (int e = tid; e < 256; e += kTxfThreads) s_enc14[e] = (uint16_t)(ham74_encode_nibble(e & 15) | (ham74_encode_nibble(e >> 4) << 7));
        __syncthreads();

        // ---- chunk bit stream (same construction as a tile of tx_tile_kernel) -------------------------------------------
        const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
        const uint32_t byte0 = (uint32_t)((long)t0 * BPS / 8), nbyte = (uint32_t)((long)(t1 - t0) * BPS / 8);
        if (FEC) {
            const uint32_t hdr = byte0 < 16 ? 16 - byte0 : 0;
            if (tid < (int)hdr) s_bits[tid] = (uint8_t)frame_byte<FEC>(pay, n, coded_len, byte0 + tid, s_enc);
            const uint32_t c0 = byte0 + hdr - 16;
            const uint32_t ngrp = ((nbyte + 2 - hdr + 6) / 7 + 3) / 4;
            const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
            for (uint32_t u = tid; u < ngrp; u += kTxfThreads) {
                const uint32_t pb = (c0 / 7) * 4 + 16 * u;
                uint32_t v[4] = { 0, 0, 0, 0 };
                if (pay_aligned && pb + 16 <= n) {
#pragma unroll
                    for (int q = 0; q < 4; q++) v[q] = __ldg(reinterpret_cast<const uint32_t *>(pay + pb) + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 16; q++) if (pb + q < n) v[q >> 2] |= (uint32_t)pay[pb + q] << (8 * (q & 3));
                }
                uint64_t w[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t lo = (uint32_t)s_enc14[v[q] & 255u] | ((uint32_t)s_enc14[(v[q] >> 8) & 255u] << 14);
                    const uint32_t hi = (uint32_t)s_enc14[(v[q] >> 16) & 255u] | ((uint32_t)s_enc14[v[q] >> 24] << 14);
                    w[q] = (uint64_t)lo | ((uint64_t)hi << 28);
                }
                uint32_t *dst = reinterpret_cast<uint32_t *>(s_bits + hdr + 28 * u);
                dst[0] = (uint32_t)w[0];
                dst[1] = (uint32_t)(w[0] >> 32) | ((uint32_t)w[1] << 24);
                dst[2] = (uint32_t)(w[1] >> 8);
                dst[3] = (uint32_t)(w[1] >> 40) | ((uint32_t)w[2] << 16);
                dst[4] = (uint32_t)(w[2] >> 16);
                dst[5] = (uint32_t)(w[2] >> 48) | ((uint32_t)w[3] << 8);
                dst[6] = (uint32_t)(w[3] >> 24);
            }
        } else {
            for (uint32_t b = tid; b < nbyte + 2; b += kTxfThreads) s_bits[b] = (uint8_t)frame_byte<FEC>(pay, n, coded_len, byte0 + b, s_enc);
        }
        __syncthreads();
        {
            const long ncar_local = (long)ncar - (long)t0 * D;
            long have = (long)(t1 - t0) * D;
            if (ncar_local < have) have = ncar_local;
            const int ncar_have = (int)have;
            const uint32_t *bits32 = reinterpret_cast<const uint32_t *>(s_bits);
            for (int c4 = 4 * tid; c4 < SLOTS * D; c4 += 4 * kTxfThreads) {
                const uint32_t bit = (uint32_t)c4 * BPC, wi = bit >> 5, sh = bit & 31;
                const uint32_t v = __funnelshift_r(bits32[wi], bits32[wi + 1], sh);
                constexpr uint32_t M = (uint32_t)(NE - 1);
                uint32_t packed = (v & M) | (((v >> BPC) & M) << 8) | (((v >> (2 * BPC)) & M) << 16) | (((v >> (3 * BPC)) & M) << 24);
                if (c4 + 4 > ncar_have) {
#pragma unroll
                    for (int q = 0; q < 4; q++) if (c4 + q >= ncar_have) packed = (packed & ~(0xFFu << (8 * q))) | ((uint32_t)NE << (8 * q));
                }
                *reinterpret_cast<uint32_t *>(s_car + c4) = packed;
            }
        }
        __syncthreads();

        // ---- transforms: each symbol once, results stay in shared memory --------------------------------------------------
        cpx tw[8];
#pragma unroll
        for (int ka = 0; ka < 8; ka++) tw[ka] = c_from(__ldg(a.tables->w64 + ((l * ka) & 63)));
        int d3 = 24 - (l >= 2), d4 = 31 - (l >= 1);
        uint32_t fix0 = 0xFFu, fix3 = 0xFFu, fix4 = 0xFFu, fix7 = 0xFFu;
        if (GUARD) {
            if (data_rank<GUARD>(l) < 0) fix0 = is_pilot_bin(l) ? NE + 1 : NE;
            if (data_rank<GUARD>(l + 24) < 0) fix3 = is_pilot_bin(l + 24) ? NE + 1 : NE;
            if (data_rank<GUARD>(l + 32) < 0) fix4 = is_pilot_bin(l + 32) ? NE + 1 : NE;
            if (data_rank<GUARD>(l + 56) < 0) fix7 = is_pilot_bin(l + 56) ? NE + 1 : NE;
        }
        asm volatile("" : "+r"(d3), "+r"(d4), "+r"(fix0), "+r"(fix3), "+r"(fix4), "+r"(fix7));
        float2 *tr = s_tr + warp * kTrWarp + g * kTrGroup;
        const unsigned long long *lut = reinterpret_cast<const unsigned long long *>(s_lut) + (lane & 15);
#pragma unroll 1
        for (int it = 0; it < kTxfIters; it++) {
            const int sl = it * (4 * kTxfWarps) + 4 * warp + g;                   // symbol slot inside the chunk (the tail slots are unused)
            const bool valid = t0 + sl < t1;
            const uint8_t *rowp = s_car + sl * D + (GUARD ? l - 7 : l);
            cpx x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t idx;
                if (!GUARD) idx = rowp[8 * j];
                else if (j == 1 || j == 2) idx = rowp[8 * j];
                else if (j == 5 || j == 6) idx = rowp[8 * j - 3];
                else if (j == 3) idx = fix3 != 0xFFu ? fix3 : rowp[d3];
                else if (j == 4) idx = fix4 != 0xFFu ? fix4 : rowp[d4];
                else if (j == 0) idx = fix0 != 0xFFu ? fix0 : rowp[0];
                else idx = fix7 != 0xFFu ? fix7 : rowp[53];
                x[j].v = lut[idx * 16];
            }
            fft64_group_p(x, tw, tr, l);
            if (valid) {
                unsigned long long *o = reinterpret_cast<unsigned long long *>(s_out + sl * kNfft + l);
#pragma unroll
                for (int kb = 0; kb < 8; kb++) {
                    float re, im;
                    c_split(x[kb], im, re);                                       // un-swap
                    mx = fmaxf(mx, fmaxf(re, im));
                    o[8 * kb] = c_make(re, im).v;                                 // time index l + 8 kb
                }
            }
        }
        mx *= 1.0f / 64.0f;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0 && mx > 0.0f) atomicMax(a.stream_max + stream, __float_as_int(mx));
        __threadfence();
    }
    // ---- the frame's maximum: every CTA of the cluster has published its own ---------------------------------------------
    cluster_barrier();
    const float fmx = fmaxf(__int_as_float(*reinterpret_cast<volatile int *>(a.stream_max + stream)), a.tables->head_max);
    const float scale = (1.0f / 64.0f) / fmx;
    if (rank == 0)
        for (uint32_t i = tid; i < (uint32_t)(kHeadSyms * kSym) && i < a.iq_stride; i += kTxfThreads) {
            float2 v = make_float2(0.0f, 0.0f);
            if (fits) { v = a.tables->head[i]; v.x = v.x / fmx; v.y = v.y / fmx; }
            out[i] = v;
        }
    if (rank == gridDim.x - 1) {
        const uint32_t z0 = fits ? frame_len : (uint32_t)(kHeadSyms * kSym);
        for (uint32_t i = z0 + tid; i < a.iq_stride; i += kTxfThreads) out[i] = make_float2(0.0f, 0.0f);
    }
    if (!work) return;
    // ---- scale + copy out, cyclic prefix included (prefix_block, src/transmitter.rs:168-181: x[48..64] ++ x[0..64]) ------------
    float2 *dst = out + (size_t)(kHeadSyms + t0) * kSym;
    const int n_out = (t1 - t0) * kSym;
    const cpx sc = c_make(scale, scale);
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        for (int i = 2 * tid; i < n_out; i += 2 * kTxfThreads) {                  // two samples per thread: t even, so both come from one 16-byte read
            const int sym = i / kSym, t = i - sym * kSym;
            const int src = sym * kNfft + (t < kCp ? kNfft - kCp + t : t - kCp);
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(s_out + src);
            cpx p0, p1;
            p0.v = v.x; p1.v = v.y;
            ulonglong2 w;
            w.x = c_mul2(p0, sc).v; w.y = c_mul2(p1, sc).v;
            *reinterpret_cast<ulonglong2 *>(dst + i) = w;
        }
    } else {
        for (int i = tid; i < n_out; i += kTxfThreads) {
            const int sym = i / kSym, t = i - sym * kSym;
            cpx p0;
            p0.v = *reinterpret_cast<const unsigned long long *>(s_out + sym * kNfft + (t < kCp ? kNfft - kCp + t : t - kCp));
            *reinterpret_cast<unsigned long long *>(dst + i) = c_mul2(p0, sc).v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// channel harness
// ------------------------------------------------------------------------------------------------------------------
struct ChanArgs {
    const float2   *tx;
    const uint32_t *tx_len;
    uint32_t        tx_stride;
    uint32_t        n_streams;
    float2         *rx;
    uint32_t        rx_stride;
    uint32_t       *rx_len;
    uint32_t       *lead_out;
    float          *cfo_out;
    double         *accum;          // per stream (f64): sum re, sum im, sum (y^2).re, sum (y^2).im, sum |y|^2
    float           snr_lin;
    float           cfo_max;
    uint32_t        lead_min, lead_max;
    uint32_t        multipath;
    uint32_t        noise_mode;
    uint32_t        seed_lo, seed_hi;
    uint32_t        stream0;        // first stream of this launch
};

__device__ __forceinline__ void chan_stream_draws(const ChanArgs &a, uint32_t stream, uint32_t &lead, float &cfo)
{
    uint32_t c[4] = { stream, 0u, 0u, 0xC0FFEEu };
    philox4x32_10(c, a.seed_lo, a.seed_hi);
    uint32_t span = a.lead_max >= a.lead_min ? a.lead_max - a.lead_min + 1 : 1;
    lead = a.lead_min + (uint32_t)(((uint64_t)c[0] * span) >> 32);
    cfo = a.cfo_max >= 0.0f ? u01_from_u32(c[1]) * a.cfo_max : -1.0f;
}

static __constant__ float kChanTaps[12] = { -0.0f, -0.1912f, 0.9316f, 0.2821f, -0.1990f, 0.1630f, -0.1017f, 0.0544f, -0.0261f, 0.0090f, 0.0f, -0.0034f };

// multipath (src/channel.rs:26-31,45) + CFO (src/channel.rs:54-62) + lead-in; accumulates the signal statistics
template <int = 0>
__global__ void __launch_bounds__(256) channel_conv_kernel(const ChanArgs a)
{
    const uint32_t stream = blockIdx.y + a.stream0;
    uint32_t lead; float cfo;
    chan_stream_draws(a, stream, lead, cfo);
    const uint32_t n_tx = a.tx_len[stream];
    const uint32_t n_rx = lead + n_tx + 63;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.rx_len[stream] = n_rx <= a.rx_stride ? n_rx : 0;
        if (a.lead_out) a.lead_out[stream] = lead;
        if (a.cfo_out) a.cfo_out[stream] = cfo;
    }
    const float2 *tx = a.tx + (size_t)stream * a.tx_stride;
    float2 *rx = a.rx + (size_t)stream * a.rx_stride;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
    const bool fits = n_rx <= a.rx_stride;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < a.rx_stride; n += gridDim.x * blockDim.x) {
        float yr = 0.0f, yi = 0.0f;
        if (fits && n >= lead && n < n_rx) {
            const long i = (long)n - lead;                 // index into the convolved output
            if (a.multipath) {
#pragma unroll
                for (int k = 0; k < 12; k++) {
                    long m = i - 7 - k;
                    if (m >= 0 && m < (long)n_tx) { float2 v = tx[m]; yr = fmaf(kChanTaps[k], v.x, yr); yi = fmaf(kChanTaps[k], v.y, yi); }
                }
            } else if (i < (long)n_tx) { float2 v = tx[i]; yr = v.x; yi = v.y; }
            if (cfo >= 0.0f) {
                // exp(+j f (i+1)); phase reduced in f64 so long captures keep fp32-exact phasors
                double ph = (double)cfo * (double)(i + 1);
                ph -= 6.283185307179586476925 * floor(ph * 0.15915494309189533577);
                float s, c;
                sincosf((float)ph, &s, &c);
                cmul(yr, yi, c, s);
            }
            s0 += yr; s1 += yi; s2 += yr * yr - yi * yi; s3 += 2.0f * yr * yi; s4 += yr * yr + yi * yi;
        }
        rx[n] = make_float2(yr, yi);
    }
    // a thread's few samples are summed in fp32, everything above in f64: the statistics of a 1.6e5-sample frame keep
    // 12+ digits, so sum (mean - y)^2 = sum y^2 - n mean^2 can be formed without cancellation trouble in the noise kernel
    double d0 = s0, d1 = s1, d2 = s2, d3 = s3, d4 = s4;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, m); d1 += __shfl_xor_sync(0xffffffffu, d1, m);
        d2 += __shfl_xor_sync(0xffffffffu, d2, m); d3 += __shfl_xor_sync(0xffffffffu, d3, m);
        d4 += __shfl_xor_sync(0xffffffffu, d4, m);
    }
    if ((threadIdx.x & 31) == 0) {
        double *acc = a.accum + 5 * (size_t)stream;
        atomicAdd(acc + 0, d0); atomicAdd(acc + 1, d1); atomicAdd(acc + 2, d2); atomicAdd(acc + 3, d3); atomicAdd(acc + 4, d4);
    }
}

// noise (src/channel.rs:66-71): mode 0 = reference-faithful (complex "variance", uniform draws); mode 1 = Gaussian AWGN.
// The lead-in carries noise too (a receiver never sees an exactly silent channel).
template <int = 0>
__global__ void __launch_bounds__(256) channel_noise_kernel(const ChanArgs a)
{
    const uint32_t stream = blockIdx.y + a.stream0;
    const uint32_t n_rx = a.rx_len[stream];
    if (n_rx == 0) return;
    uint32_t lead; float cfo;
    chan_stream_draws(a, stream, lead, cfo);
    const double *acc = a.accum + 5 * (size_t)stream;
    const double cnt = (double)(n_rx - lead);
    float ar, ai;                                         // complex noise amplitude
    if (a.noise_mode == 0) {
        // var = sum (mean - y)^2 / len with NO conjugate -> complex (src/signals/mod.rs:239-249), formed in f64 as
        // sum y^2 / len - mean^2; amp = sqrt(0.5 var / snr), the principal complex square root (src/channel.rs:66-71)
        const double mr = acc[0] / cnt, mi = acc[1] / cnt;
        double vr = acc[2] / cnt - (mr * mr - mi * mi), vi = acc[3] / cnt - 2.0 * mr * mi;
        vr = 0.5 * vr / (double)a.snr_lin; vi = 0.5 * vi / (double)a.snr_lin;
        const double r = sqrt(sqrt(vr * vr + vi * vi)), th = 0.5 * atan2(vi, vr);
        ar = (float)(r * cos(th)); ai = (float)(r * sin(th));
    } else {
        ar = (float)sqrt(0.5 * (acc[4] / cnt) / (double)a.snr_lin); ai = 0.0f;
    }
    float2 *rx = a.rx + (size_t)stream * a.rx_stride;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < n_rx; n += gridDim.x * blockDim.x) {
        uint32_t c[4] = { n, stream, 1u, 0xC0FFEEu };
        philox4x32_10(c, a.seed_lo, a.seed_hi);
        float nr, ni;
        if (a.noise_mode == 0) {
            float u = 2.0f * u01_from_u32(c[0]) - 1.0f, v = 2.0f * u01_from_u32(c[1]) - 1.0f;
            nr = ar * u - ai * v; ni = ar * v + ai * u;
        } else {
            float u1 = u01_from_u32(c[0]), u2 = u01_from_u32(c[1]);
            float r = sqrtf(-2.0f * logf(u1)), s, co;
            sincospif(2.0f * u2, &s, &co);
            nr = ar * r * co; ni = ar * r * s;
        }
        float2 v = rx[n];
        rx[n] = make_float2(v.x + nr, v.y + ni);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// BER counters (src/utils.rs:45-68): one CTA per stream
// ------------------------------------------------------------------------------------------------------------------
struct BerArgs {
    const uint8_t  *ref;
    const uint32_t *ref_len;
    uint32_t        ref_stride;
    const uint8_t  *got;
    const uint32_t *got_len;
    uint32_t        got_stride;
    const int32_t  *status;
    uint32_t        n_streams;
    unsigned long long *counters;   // bit_errs, byte_errs, bits_compared, frames_failed
};

template <int = 0>
__global__ void __launch_bounds__(256) ber_kernel(const BerArgs a)
{
    const uint32_t stream = blockIdx.x;
    const uint32_t n = a.ref_len[stream];
    const bool failed = a.status[stream] != ST_OK || a.got_len[stream] != n;
    unsigned int bit_errs = 0, byte_errs = 0;
    if (!failed) {
        const uint8_t *r = a.ref + (size_t)stream * a.ref_stride, *g = a.got + (size_t)stream * a.got_stride;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            unsigned int d = (unsigned int)(r[i] ^ g[i]);
            if (d) { bit_errs += __popc(d); byte_errs++; }
        }
    }
    __shared__ unsigned int s_b[8], s_B[8];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        bit_errs += __shfl_xor_sync(0xffffffffu, bit_errs, m);
        byte_errs += __shfl_xor_sync(0xffffffffu, byte_errs, m);
    }
    if ((threadIdx.x & 31) == 0) { s_b[threadIdx.x >> 5] = bit_errs; s_B[threadIdx.x >> 5] = byte_errs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long be = 0, By = 0;
        for (int w = 0; w < 8; w++) { be += s_b[w]; By += s_B[w]; }
        if (failed) { be = 8ull * n; By = n; atomicAdd(a.counters + 3, 1ull); }
        atomicAdd(a.counters + 0, be);
        atomicAdd(a.counters + 1, By);
        atomicAdd(a.counters + 2, 8ull * n);
    }
}

}  // namespace ofdm
