// tx_warp.cuh -- nfft = 64 one-pass transmit kernels for large batches: tx_warp_kernel (exact; below) and, at the end of the file,
// tx_spec_kernel (speculative on the constant head maximum, with a redo pass: the default). tx_warp_kernel is the second version: tx_resident_kernel's transform, tensor-
// memory residency and groups of persistent CTAs (tx_resident.cuh), with the frame loop of wide_tx_resident_kernel -- NOTHING in
// it is a CTA barrier.
//
// tx_resident_kernel builds the carrier bytes of a frame with the whole CTA (phase B: coded bit stream, barrier, one byte per
// carrier, barrier) and exchanges the frame maximum through an atomic maximum plus an arrival counter. Here a warp owns 16
// CONSECUTIVE symbols of a frame (4 tensor-memory slots of 4 symbols), whose bits are one contiguous span of the frame's byte
// stream -- 16 x 36 = 576 coded bytes with 64QAM / guard bands / Hamming, exactly one nfft = 1024 symbol's worth -- so the warp
// prepares them itself with the wideband kernel's helpers (lane u: 16 payload bytes -> 32 codewords -> 28 coded bytes; then one
// byte per data carrier), and nobody waits for anybody inside an SM. The frame maximum travels through one flagged 32-bit word
// per warp of the group (plain store to publish, one L2 round trip to collect). Per frame k a warp runs:
//   collect the maximum of frame k-1 | 4 x { drain slot i of frame k-1 | transform 4 symbols of frame k into slot i } |
//   publish its maximum of frame k | carrier bytes of its symbols of frame k+1, its share of the head / zero fill of frame k-1.
// Same values as tx_resident_kernel and the two-pass kernel (same transform, same scaling).
#pragma once

#include "tx_resident.cuh"
#include "wide_tx_resident.cuh"

namespace ofdm {

constexpr int kTwWarps = 32;
constexpr int kTwThreads = 32 * kTwWarps;
constexpr int kTwWarpSyms = 16;                                  // symbols per warp and frame: 4 slots x 4 symbols (8 lanes each)
constexpr int kTwSyms = kTwWarps * kTwWarpSyms;                  // 512 symbols per CTA = all of the SM's tensor memory
constexpr int kTwCarBuf = 16 * 64 + 16;                          // per warp: one byte per data carrier of its 16 symbols (+ slack)

template <int MOD> struct TwSmem {
    static constexpr int NE = 1 << ModTraits<MOD>::kBpc;
    static constexpr size_t kTr = 0;                                                         // [warp][kTrWarp]: transpose scratch
    static constexpr size_t kLut = kTr + sizeof(float2) * kTwWarps * kTrWarp;                // [entry][lane & 15] conjugated constellation, null, pilot
    static constexpr size_t kEnc = kLut + sizeof(float2) * 16 * (NE + 2);                    // Hamming byte table (256 x u16), tensor-memory base address
    static constexpr size_t kBits = kEnc + 512 + 16 + 256 + 48;                              // (tensor-memory address, arrival counters, warp maxima) | [warp][kWTrsBitsBuf]: coded bit stream being prepared
    static constexpr size_t kCar = kBits + (size_t)kTwWarps * wide::kWTrsBitsBuf;            // [warp][kTwCarBuf]
    static constexpr size_t kTotal = kCar + (size_t)kTwWarps * kTwCarBuf;
};

struct TwGeom {
    uint32_t n;
    uint64_t coded_len, ncar;
    int      S, t0, t1;
    uint32_t frame_len;
    bool     fits;
};
template <int BPC, int D, bool FEC>
__device__ __forceinline__ TwGeom tw_geometry(const TxArgs &a, uint32_t stream, int rank)
{
    TwGeom q;
    q.n = __ldg(a.payload_len + stream);
    q.coded_len = FEC ? (14ull * q.n + 7) / 8 : q.n;
    const uint64_t nbits = kHeaderBits + 8 * q.coded_len;
    q.ncar = (nbits + BPC - 1) / BPC;                               // constellation symbols (src/transmitter.rs:108-140)
    const uint64_t S64 = (q.ncar + D - 1) / D;                      // OFDM data symbols (src/transmitter.rs:49-54)
    q.S = (int)S64;
    q.frame_len = (kHeadSyms + (uint32_t)q.S) * kSym;
    q.fits = ((uint64_t)kHeadSyms + S64) * kSym <= (uint64_t)a.iq_stride;
    const int C = a.group_ctas;
    const int chunk = (q.S + C - 1) / C;
    q.t0 = rank * chunk;
    q.t1 = q.t0 + chunk < q.S ? q.t0 + chunk : q.S;
    if (!q.fits || q.t1 < q.t0 || chunk > kTwSyms) q.t1 = q.t0;    // (the launcher sizes C so that a fitting frame's chunk never exceeds the slots)
    return q;
}

template <int MOD, bool GUARD, bool FEC>
__global__ void __launch_bounds__(kTwThreads, 1) tx_warp_kernel(const TxArgs a)
{
    typedef TwSmem<MOD> L;
    constexpr int BPC = ModTraits<MOD>::kBpc, NE = 1 << BPC, D = GUARD ? 48 : 64;
    constexpr int DW = kTwWarpSyms * D, BPSB = BPC * DW / 8;                    // a warp's span: carriers, stream bytes
    static_assert(16 + 3 + 28 * (((BPSB + 2 + 6 + 6) / 7 + 3) / 4) <= wide::kWTrsBitsBuf - 8 && DW + 8 <= kTwCarBuf, "per-warp buffers");
    extern __shared__ __align__(128) uint8_t tw_smem[];
    float2 *s_tr = reinterpret_cast<float2 *>(tw_smem + L::kTr);
    float2 *s_lut = reinterpret_cast<float2 *>(tw_smem + L::kLut);
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(tw_smem + L::kEnc);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(tw_smem + L::kEnc + 512);

    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    const int C = a.group_ctas, G = a.n_groups;
    const int group = (int)blockIdx.x / C, rank = (int)blockIdx.x - group * C;

    // ---- set-up: tables, tensor memory ------------------------------------------------------------------------------------
    for (int e = tid; e < 16 * (NE + 2); e += kTwThreads) {                     // conjugated constellation (conj . FFT . conj), null, pilot
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
        s_lut[e] = make_float2(re, -im);
    }
    if (FEC && tid < 256) s_enc14[tid] = (uint16_t)(ham74_encode_nibble(tid & 15) | (ham74_encode_nibble(tid >> 4) << 7));
    if (tid < 2) reinterpret_cast<uint32_t *>(tw_smem + L::kEnc + 512 + 8)[tid] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_addr(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(s_tmem);
    const float head_max = a.tables->head_max;

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
    // tensor-memory slots of this warp: lane quadrant warp % 4, columns 64 (warp / 4) + 16 slot
    uint32_t taddr_w = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (warp >> 2));
    asm volatile("" : "+r"(taddr_w));
    uint8_t *mybits = tw_smem + L::kBits + (size_t)warp * wide::kWTrsBitsBuf;
    uint8_t *mycar = tw_smem + L::kCar + (size_t)warp * kTwCarBuf;
    const uint8_t *car = mycar + (GUARD ? l - 7 : l);

    // drain slot `it` of the previous frame: scale, store with the cyclic prefix (prefix_block, src/transmitter.rs:168-181)
    auto drain = [&](int it, int p_first, int p_t1, float2 *p_out, float p_scale) {
        cpx y[8];
        tmem_ld16(taddr_w + 16u * (uint32_t)it, y);
        const int s = p_first + 4 * it + g;
        if (s < p_t1) {
            unsigned long long *sym = reinterpret_cast<unsigned long long *>(p_out + (size_t)(kHeadSyms + s) * kSym + l);
            const cpx sc = c_make(p_scale, -p_scale);                          // conj and scale in one
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                const unsigned long long v = c_mul2(y[kb], sc).v;              // time index l + 8 kb
                sym[kCp + 8 * kb] = v;
                if (kb >= 6) sym[8 * kb - (kNfft - kCp)] = v;                  // cyclic prefix = last 16 samples
            }
        }
    };
    const uint32_t gthreads = (uint32_t)(C * kTwThreads), gt = (uint32_t)(rank * kTwThreads + tid);
    auto write_head = [&](uint32_t stream, bool fits, uint32_t frame_len, float fmx) {
        float2 *out = a.iq + (size_t)stream * a.iq_stride;
        const uint32_t hw = a.iq_stride < (uint32_t)(kHeadSyms * kSym) ? a.iq_stride : (uint32_t)(kHeadSyms * kSym);
        for (uint32_t i = gt; i < hw; i += gthreads) {
            float2 v = make_float2(0.0f, 0.0f);
            if (fits) { v = a.tables->head[i]; v.x = v.x / fmx; v.y = v.y / fmx; }
            out[i] = v;
        }
        const uint32_t z0 = fits ? frame_len : (uint32_t)(kHeadSyms * kSym);
        for (uint32_t i = z0 + gt; i < a.iq_stride; i += gthreads) out[i] = make_float2(0.0f, 0.0f);
    };
    // Frame maximum (normalize, src/transmitter.rs:183-194), two levels. Inside the CTA: every warp drops its maximum into shared
    // memory and counts itself (shared-memory atomic); the LAST warp of a frame reduces the 32 values and publishes ONE flagged
    // word for the CTA (float bits | 0x80000000; the array is zeroed before the launch). Between the CTAs of the group: a poll is
    // one load by lane r < C of CTA r's word, a vote on the flags and a warp maximum -- a quarter of the instructions and of the
    // L2 requests of polling one word per warp, which matters here: up to 31 spinning warps share the schedulers with the working ones.
    float *s_wmax = reinterpret_cast<float *>(tw_smem + L::kEnc + 512 + 16);     // [2][32] (frame parity, warp)
    uint32_t *s_arrived = reinterpret_cast<uint32_t *>(tw_smem + L::kEnc + 512 + 8);   // [2]
    const uint32_t n_words = (uint32_t)C;
    auto publish_max = [&](uint32_t stream, float m, int parity) {
        if (lane == 0) s_wmax[32 * parity + warp] = fmaxf(m, 0.0f);
        __syncwarp();
        uint32_t seen = 0;
        if (lane == 0) { __threadfence_block(); seen = atomicAdd(s_arrived + parity, 1u); }
        seen = __shfl_sync(0xffffffffu, seen, 0);
        if (seen == (uint32_t)(kTwWarps - 1)) {                                  // the CTA's last warp of this frame
            __threadfence_block();
            float v = s_wmax[32 * parity + lane];
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, sft));
            if (lane == 0) {
                s_arrived[parity] = 0;                                           // (parity is reused two frames later, after everybody has collected this one)
                asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(a.stream_cnt + (size_t)stream * n_words + (uint32_t)rank),
                             "r"(__float_as_uint(v) | 0x80000000u) : "memory");
            }
        }
    };
    auto frame_max = [&](uint32_t stream) -> float {
        const uint32_t *sl = a.stream_cnt + (size_t)stream * n_words;
        float m;
        for (;;) {
            uint32_t all = 0x80000000u;
            m = 0.0f;
            for (uint32_t w = lane; w < n_words; w += 32) {
                const uint32_t v = ld_relaxed_gpu(sl + w);
                all &= v;
                m = fmaxf(m, __uint_as_float(v & 0x7FFFFFFFu));
            }
            if (__all_sync(0xffffffffu, (all >> 31) != 0u)) break;
            __nanosleep(64);
        }
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
        return fmaxf(m, head_max);
    };
    // carrier bytes of this warp's 16 symbols of frame `stream` (modulate, src/transmitter.rs:108-140)
    auto build = [&](const TwGeom &q, uint32_t stream) {
        const int first = q.t0 + kTwWarpSyms * warp;
        if (first >= q.t1) return;
        const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
        const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
        const uint32_t bit0 = wide::wtrs_build_bits<BPSB, FEC>(mybits, pay, pay_aligned, q.n, q.coded_len, (uint32_t)first * (BPC * D / 8), s_enc14, lane);
        __syncwarp();
        wide::wtrs_unpack_carriers<BPC, DW>(mycar, mybits, bit0, (long)q.ncar - (long)first * D, lane);
        __syncwarp();
    };
    auto prefetch = [&](const TwGeom &q, uint32_t stream) {
        const int first = q.t0 + kTwWarpSyms * warp;
        if (first >= q.t1) return;
        const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
        const uint32_t B0 = (uint32_t)first * (BPC * D / 8), c0 = B0 < 16 ? 0 : B0 - 16;
        const uint32_t pb = FEC ? 4 * (c0 / 7) + 16 * lane : c0 + 32 * lane;
        constexpr uint32_t span = FEC ? (BPSB * 4) / 7 + 32 : BPSB + 32;
        if ((FEC ? 16u : 32u) * lane < span && pb < q.n) asm volatile("prefetch.global.L1 [%0];" :: "l"(pay + pb));
    };

    // Only (first symbol of the warp, end of the CTA's chunk, frame length, fits) of a frame live across the loop: with 1024
    // threads a thread has 64 registers, and the rest of the geometry is a few integer operations away when it is needed.
    uint32_t stream = (uint32_t)group;                                         // (the launcher guarantees group < n_streams)
    int first, t1;
    uint32_t flen;
    bool fits;
    {
        const TwGeom q0 = tw_geometry<BPC, D, FEC>(a, stream, rank);
        build(q0, stream);
        first = q0.t0 + kTwWarpSyms * warp; t1 = q0.t1; flen = q0.frame_len; fits = q0.fits;
    }
    bool have_prev = false, p_fits = false;
    int p_first = 0, p_t1 = 0;
    uint32_t p_stream = 0, p_flen = 0;

    for (int k = 0; ; k++) {
        if (rank == 0 && tid == 0 && a.frame_len) a.frame_len[stream] = flen;
        const uint32_t next = stream + (uint32_t)G;
        const bool more = next < a.n_streams;
        if (more) prefetch(tw_geometry<BPC, D, FEC>(a, next, rank), next);
        // ---- drain frame k-1, transform frame k -------------------------------------------------------------------------------
        float p_fmx = 1.0f, p_scale = 0.0f;
        if (have_prev) { p_fmx = frame_max(p_stream); p_scale = (1.0f / 64.0f) * (1.0f / p_fmx); }
        tmem_wait_st();                                                        // the slots of frame k-1 were written an iteration ago
        float2 *p_out = a.iq + (size_t)p_stream * a.iq_stride;
        float mx = 0.0f;
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (have_prev && p_first + 4 * it < p_t1) drain(it, p_first, p_t1, p_out, p_scale);
            if (first + 4 * it < t1) {
                const int sl = 4 * it + g;                                     // symbol inside the warp's span
                const bool valid = first + sl < t1;
                const uint8_t *rowp = car + (valid ? sl : 0) * D;
                uint32_t idx[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {                                  // encode_block, src/transmitter.rs:144-165
                    if (!GUARD) idx[j] = rowp[8 * j];
                    else if (j == 1 || j == 2) idx[j] = rowp[8 * j];
                    else if (j == 5 || j == 6) idx[j] = rowp[8 * j - 3];
                    else if (j == 3) idx[j] = fix3 != 0xFFu ? fix3 : rowp[d3];
                    else if (j == 4) idx[j] = fix4 != 0xFFu ? fix4 : rowp[d4];
                    else if (j == 0) idx[j] = fix0 != 0xFFu ? fix0 : rowp[0];
                    else idx[j] = fix7 != 0xFFu ? fix7 : rowp[53];
                }
                cpx x[8];
#pragma unroll
                for (int j = 0; j < 8; j++) x[j].v = lut[idx[j] * 16];
                fft64_group_p(x, tw, tr, l);                                   // prefix_block, src/transmitter.rs:168-181 (IFFT part)
                if (valid) {
#pragma unroll
                    for (int kb = 0; kb < 8; kb++) {
                        float re, im;
                        c_split(x[kb], re, im);                                // the frame's sample is (re, -im)
                        mx = fmaxf(mx, fmaxf(re, -im));
                    }
                }
                tmem_st16(taddr_w + 16u * (uint32_t)it, x);
            }
        }
        // ---- publish this warp's maximum of frame k ---------------------------------------------------------------------------
        mx *= 1.0f / 64.0f;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        publish_max(stream, mx, k & 1);
        // ---- carrier bytes of frame k+1, head / zero fill of frame k-1 (also the time the maximum of frame k needs to travel) -------
        __syncwarp();                                                          // every lane has read its carriers of frame k
        if (have_prev) write_head(p_stream, p_fits, p_flen, p_fmx);
        have_prev = true; p_first = first; p_t1 = t1; p_stream = stream; p_flen = flen; p_fits = fits;
        if (!more) break;
        {
            const TwGeom qn = tw_geometry<BPC, D, FEC>(a, next, rank);
            build(qn, next);
            first = qn.t0 + kTwWarpSyms * warp; t1 = qn.t1; flen = qn.frame_len; fits = qn.fits;
        }
        stream = next;
    }
    // ---- the group's last frame ---------------------------------------------------------------------------------------------
    {
        const float p_fmx = frame_max(p_stream), p_scale = (1.0f / 64.0f) * (1.0f / p_fmx);
        tmem_wait_st();
        float2 *p_out = a.iq + (size_t)p_stream * a.iq_stride;
#pragma unroll 1
        for (int it = 0; it < 4; it++)
            if (p_first + 4 * it < p_t1) drain(it, p_first, p_t1, p_out, p_scale);
        write_head(p_stream, p_fits, p_flen, p_fmx);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}


// ---- speculative one-pass kernel: no residency at all ---------------------------------------------------------------------------
// The frame maximum `normalize` divides by is max(head_max, maximum over the data symbols), and head_max -- the maximum of the
// constant frame head (lock | preamble | training) -- is known before anything is transformed. For scrambled / random payloads the
// data symbols stay well below it (64QAM, nfft 64: 0.68-0.77 of it over 2038 symbols), so the frame maximum IS head_max. This
// kernel bets on that: every symbol is transformed once, scaled with head_max and stored AT ONCE -- no tensor-memory residency,
// no exchange between SMs, no cooperative launch; a warp takes 16 consecutive symbols, prepares their carrier bytes itself and
// streams them out. It also keeps the maximum it saw, and a warp whose data beat head_max records it (atomicMax on
// stream_max[stream], which stays 0 otherwise). Those frames -- a payload of equal bytes does it -- are then redone by the
// two-pass kernel's store pass with the recorded maximum (tx_tile_kernel<WRITE> with redo_only: it exits at once for every other
// frame), which writes the very same values the exact kernels write. Output is identical either way; the bet only decides the cost.
// Work unit = (frame, span of 512 symbols), persistent CTAs stride over the units.
template <int MOD, bool GUARD, bool FEC>
__global__ void __launch_bounds__(kTwThreads, 1) tx_spec_kernel(const TxArgs a)
{
    typedef TwSmem<MOD> L;
    constexpr int BPC = ModTraits<MOD>::kBpc, NE = 1 << BPC, D = GUARD ? 48 : 64;
    constexpr int DW = kTwWarpSyms * D, BPSB = BPC * DW / 8;
    extern __shared__ __align__(128) uint8_t tw_smem[];
    float2 *s_tr = reinterpret_cast<float2 *>(tw_smem + L::kTr);
    float2 *s_lut = reinterpret_cast<float2 *>(tw_smem + L::kLut);
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(tw_smem + L::kEnc);

    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    for (int e = tid; e < 16 * (NE + 2); e += kTwThreads) {                     // conjugated constellation (conj . FFT . conj), null, pilot
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
        s_lut[e] = make_float2(re, -im);
    }
    if (FEC && tid < 256) s_enc14[tid] = (uint16_t)(ham74_encode_nibble(tid & 15) | (ham74_encode_nibble(tid >> 4) << 7));
    __syncthreads();
    const float head_max = a.tables->head_max;
    const float scale = (1.0f / 64.0f) * (1.0f / head_max);

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
    uint8_t *mybits = tw_smem + L::kBits + (size_t)warp * wide::kWTrsBitsBuf;
    uint8_t *mycar = tw_smem + L::kCar + (size_t)warp * kTwCarBuf;
    const uint8_t *car = mycar + (GUARD ? l - 7 : l);

    const uint32_t U = (uint32_t)a.group_ctas;                                  // spans of 512 symbols per frame (from iq_stride)
    const uint64_t n_units = (uint64_t)a.n_streams * U;
    for (uint64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint32_t stream = n_units >> 32 ? (uint32_t)(unit / U) : (uint32_t)unit / U, j = (uint32_t)(unit - (uint64_t)stream * U);
        // geometry of the frame (src/transmitter.rs:37-54, 108-140)
        const uint32_t n = __ldg(a.payload_len + stream);
        const uint64_t coded_len = FEC ? (14ull * n + 7) / 8 : n;
        const uint64_t ncar = (kHeaderBits + 8 * coded_len + BPC - 1) / BPC;
        const uint64_t S64 = (ncar + D - 1) / D;
        const uint32_t flen = (kHeadSyms + (uint32_t)S64) * kSym;
        const bool fits = ((uint64_t)kHeadSyms + S64) * kSym <= (uint64_t)a.iq_stride;
        const int S = (int)S64;
        const int t0 = (int)j * kTwSyms;
        int t1 = t0 + kTwSyms < S ? t0 + kTwSyms : S;
        if (!fits || t1 < t0) t1 = t0;
        float2 *out = a.iq + (size_t)stream * a.iq_stride;
        if (j == 0) {                                                          // frame head, already divided by the bet
            if (tid == 0 && a.frame_len) a.frame_len[stream] = flen;
            for (uint32_t i = tid; i < (uint32_t)(kHeadSyms * kSym) && i < a.iq_stride; i += kTwThreads) {
                float2 v = make_float2(0.0f, 0.0f);
                if (fits) { v = a.tables->head[i]; v.x = v.x / head_max; v.y = v.y / head_max; }
                out[i] = v;
            }
        }
        {                                                                      // zero fill past the frame inside this unit's share of the row
            const uint64_t u_lo = (uint64_t)(kHeadSyms + (uint64_t)j * kTwSyms) * kSym;
            const uint64_t u_hi = j + 1 == U ? (uint64_t)a.iq_stride : (uint64_t)(kHeadSyms + (uint64_t)(j + 1) * kTwSyms) * kSym;
            const uint64_t f_end = fits ? (uint64_t)flen : (uint64_t)(kHeadSyms * kSym);
            uint64_t z = f_end > u_lo ? f_end : u_lo;
            const uint64_t hi = u_hi < (uint64_t)a.iq_stride ? u_hi : (uint64_t)a.iq_stride;
            for (z += tid; z < hi; z += kTwThreads) out[z] = make_float2(0.0f, 0.0f);
        }
        const int first = t0 + kTwWarpSyms * warp;
        if (first >= t1) continue;
        // carrier bytes of this warp's 16 symbols (modulate, src/transmitter.rs:108-140)
        {
            const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
            const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
            __syncwarp();
            const uint32_t bit0 = wide::wtrs_build_bits<BPSB, FEC>(mybits, pay, pay_aligned, n, coded_len, (uint32_t)first * (BPC * D / 8), s_enc14, lane);
            __syncwarp();
            wide::wtrs_unpack_carriers<BPC, DW>(mycar, mybits, bit0, (long)ncar - (long)first * D, lane);
            __syncwarp();
        }
        float mx = 0.0f;
        const cpx sc = c_make(scale, -scale);                                  // conj and scale in one
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            if (first + 4 * it >= t1) break;
            const int sl = 4 * it + g;
            const bool valid = first + sl < t1;
            const uint8_t *rowp = car + (valid ? sl : 0) * D;
            uint32_t idx[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {                                   // encode_block, src/transmitter.rs:144-165
                if (!GUARD) idx[jj] = rowp[8 * jj];
                else if (jj == 1 || jj == 2) idx[jj] = rowp[8 * jj];
                else if (jj == 5 || jj == 6) idx[jj] = rowp[8 * jj - 3];
                else if (jj == 3) idx[jj] = fix3 != 0xFFu ? fix3 : rowp[d3];
                else if (jj == 4) idx[jj] = fix4 != 0xFFu ? fix4 : rowp[d4];
                else if (jj == 0) idx[jj] = fix0 != 0xFFu ? fix0 : rowp[0];
                else idx[jj] = fix7 != 0xFFu ? fix7 : rowp[53];
            }
            cpx x[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) x[jj].v = lut[idx[jj] * 16];
            fft64_group_p(x, tw, tr, l);                                       // prefix_block, src/transmitter.rs:168-181 (IFFT part)
            if (valid) {
                unsigned long long *sym = reinterpret_cast<unsigned long long *>(out + (size_t)(kHeadSyms + first + sl) * kSym + l);
#pragma unroll
                for (int kb = 0; kb < 8; kb++) {
                    float re, im;
                    c_split(x[kb], re, im);                                    // the frame's sample is (re, -im)
                    mx = fmaxf(mx, fmaxf(re, -im));
                    const unsigned long long v = c_mul2(x[kb], sc).v;          // time index l + 8 kb
                    sym[kCp + 8 * kb] = v;
                    if (kb >= 6) sym[8 * kb - (kNfft - kCp)] = v;              // cyclic prefix = last 16 samples
                }
            }
        }
        mx *= 1.0f / 64.0f;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0 && mx > head_max) {                                      // the bet is lost for this frame: it will be redone
            atomicMax(a.stream_max + stream, __float_as_int(mx));
            atomicAdd(a.stream_cnt, 1u);                                       // (what the redo pass looks at first)
        }
    }
}

}  // namespace ofdm
