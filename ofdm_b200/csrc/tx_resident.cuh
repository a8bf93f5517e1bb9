// tx_resident.cuh -- one-pass transmit kernel for large batches: a frame stays ON CHIP between its transform and its store.
//
// `normalize` (src/transmitter.rs:183-194) divides the whole frame by its maximum positive component, so no sample can be
// written before every symbol has been transformed. tx_tile_kernel therefore transforms every symbol twice (maximum pass,
// store pass). Here every symbol is transformed ONCE: a frame belongs to a GROUP of C persistent CTAs (one per SM, any C --
// the group is formed by block index, not by a hardware cluster, so all 148 SMs are used whatever C is), and the
// un-normalised time-domain symbols of CTA r's share of the frame wait in that SM's TENSOR MEMORY (tcgen05.st / tcgen05.ld,
// SASS STTM / LDTM): 256 kB per SM that no other stage of this path uses, exactly 512 symbols, and storage that costs no
// shared-memory bandwidth -- which is what bounds the transform. The frame maximum is exchanged through global memory
// (atomicMax + an arrival counter per frame, release / acquire at gpu scope).
//
// A warp owns 4 tensor-memory slots (its lane quadrant x 64 columns); a slot = one warp iteration = 4 symbols x 64 samples in
// the registers' own layout, so a slot is only ever touched by the warp that wrote it: no inter-warp synchronisation on the
// data. Per frame k a CTA (1024 threads) runs: (A) per warp, 4 x { drain slot i of frame k-1: scale, store with the cyclic
// prefix | transform 4 symbols of frame k into slot i } -- stores and transforms interleave in every warp; publish the
// warp's maximum of frame k; (B) all warps together build the carrier bytes of frame k+1 (coded bit stream -> one byte per
// data carrier) and write the frame head / zero fill of frame k-1 -- which is also the time the maximum of frame k needs
// to travel between the SMs of the group.
#pragma once

#include "tx_kernels.cuh"

namespace ofdm {

// W warps per CTA, 32 / W CTAs per SM (each allocates 16 W of the SM's 512 tensor-memory columns)
constexpr int kTrsWarpsPerCta = 32;                                           // the instantiated W
constexpr int kTrsSlots = 4;                                                   // slots per warp, 16 columns each
template <int W> struct TrsShape {
    static constexpr int kThreads = 32 * W;
    static constexpr int kIterSyms = 4 * W;                                    // symbols per iteration of the CTA
    static constexpr int kSyms = kIterSyms * kTrsSlots;                        // symbols the CTA's slots hold
    static constexpr int kChunkMax = 7 * (kSyms / 7);                          // chunks are multiples of 7 symbols
    static constexpr int kCols = 16 * W;                                       // power of two >= 32 for W = 8, 16, 32
};

template <int MOD, bool GUARD, int W> struct TrsSmem {
    static constexpr int BPC = ModTraits<MOD>::kBpc;
    static constexpr int D = GUARD ? 48 : 64;
    static constexpr int NE = 1 << BPC;
    static constexpr size_t kTr = 0;                                                           // transpose scratch
    static constexpr size_t kCar = kTr + sizeof(float2) * W * kTrWarp;                         // carrier bytes of the chunk
    static constexpr size_t kBits = kCar + (size_t)TrsShape<W>::kSyms * D + 64;                // packed bit stream of the chunk
    static constexpr size_t kPayBytes = (size_t)TrsShape<W>::kSyms * BPC * D / 8 + 64;         // prefetched payload bytes of the next chunk
    static constexpr size_t kPay = kBits + (size_t)TrsShape<W>::kSyms * BPC * D / 8 + 64;
    static constexpr size_t kLut = kPay + kPayBytes;
    static constexpr size_t kEnc = kLut + sizeof(float2) * 16 * (NE + 2);
    static constexpr size_t kMisc = kEnc + 16 + 512;                                           // tensor-memory base address, payload mbarrier
    static constexpr size_t kGeom = kMisc + 64;                                                // TrsGeom[2]
    static constexpr size_t kTotal = kGeom + 256;
};

// Strong (L2) load without acquire semantics: an acquire load is followed by an invalidation of the whole L1 (CCTL.IVALL), once
// per poll and warp. Nothing weak is read on the strength of the counter -- the only dependent read is the frame maximum,
// itself a strong load issued after the polling loop has exited -- so the invalidation buys nothing.
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Publish a warp's maximum and count its arrival WITHOUT a fence (a release would wait for every output store the thread
// still has in flight -- 5 % of the warp time, measured): the maximum is an atomic that RETURNS its old value, so it has
// been performed at the L2 when the value comes back, and the counter increment depends on that value.
__device__ __forceinline__ void publish_max_and_arrive(int *mx_addr, int mx_bits, uint32_t *cnt_addr)
{
    int old;
    asm volatile("atom.relaxed.gpu.global.max.s32 %0, [%1], %2;" : "=r"(old) : "l"(mx_addr), "r"(mx_bits) : "memory");
    const uint32_t one = 1u + ((uint32_t)old & 0x80000000u);                   // old >= 0 (float bits of a maximum >= 0): always 1
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l"(cnt_addr), "r"(one) : "memory");
}
// tensor memory: a warp reads / writes 32 lanes x 16 columns = its 16 registers of one slot (32x32b shape: thread i <-> lane i
// of the warp's quadrant)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const cpx (&x)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr),
                    "r"((uint32_t)x[0].v), "r"((uint32_t)(x[0].v >> 32)), "r"((uint32_t)x[1].v), "r"((uint32_t)(x[1].v >> 32)),
                    "r"((uint32_t)x[2].v), "r"((uint32_t)(x[2].v >> 32)), "r"((uint32_t)x[3].v), "r"((uint32_t)(x[3].v >> 32)),
                    "r"((uint32_t)x[4].v), "r"((uint32_t)(x[4].v >> 32)), "r"((uint32_t)x[5].v), "r"((uint32_t)(x[5].v >> 32)),
                    "r"((uint32_t)x[6].v), "r"((uint32_t)(x[6].v >> 32)), "r"((uint32_t)x[7].v), "r"((uint32_t)(x[7].v >> 32))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, cpx (&x)[8])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) x[i].v = (unsigned long long)r[2 * i] | ((unsigned long long)r[2 * i + 1] << 32);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// this CTA's share of frame `stream`: symbols [t0, t1) of the frame's S data symbols (chunks of a multiple of 7 symbols,
// shifted by tile_shift so that every chunk starts on a Hamming byte boundary, like the tiles of tx_tile_kernel)
struct TrsGeom {
    uint32_t n;            // payload bytes
    uint64_t coded_len, ncar;
    int      S, lo, chunk, t0, t1;
    uint32_t frame_len;
    bool     fits;
    // payload bytes of the chunk, prefetched into shared memory by one TMA bulk copy while the previous frame is transformed
    int      pf_on, pf_off;             // pf_off: payload byte index of the buffer's first byte (16-byte aligned superset)
};
template <int BPC, int D, bool FEC, int CHUNK_MAX>
__device__ __forceinline__ TrsGeom trs_geometry(const TxArgs &a, uint32_t stream, int rank)
{
    TrsGeom q;
    q.n = a.payload_len[stream];
    q.coded_len = FEC ? (14ull * q.n + 7) / 8 : q.n;
    const uint64_t nbits = kHeaderBits + 8 * q.coded_len;
    q.ncar = (nbits + BPC - 1) / BPC;                               // constellation symbols (src/transmitter.rs:108-140)
    q.S = (int)((q.ncar + D - 1) / D);                              // OFDM data symbols (src/transmitter.rs:49-54)
    q.frame_len = (kHeadSyms + (uint32_t)q.S) * kSym;
    q.fits = q.frame_len <= a.iq_stride;
    q.pf_on = 0; q.pf_off = 0;
    const int C = a.group_ctas;
    q.chunk = 7 * ((q.S + a.tile_shift + 7 * C - 1) / (7 * C));
    q.lo = rank * q.chunk - a.tile_shift;
    q.t0 = q.lo < 0 ? 0 : q.lo;
    q.t1 = q.lo + q.chunk < q.S ? q.lo + q.chunk : q.S;
    if (!q.fits || q.t1 < q.t0 || q.chunk > CHUNK_MAX) q.t1 = q.t0;      // (the launcher sizes C so that a fitting frame's chunk never exceeds the ring)
    return q;
}

template <int MOD, bool GUARD, bool FEC, int W>
__global__ void __launch_bounds__(32 * W, 32 / W) tx_resident_kernel(const TxArgs a)
{
    typedef TrsSmem<MOD, GUARD, W> L;
    constexpr int kTrsThreads = TrsShape<W>::kThreads, kTrsIterSyms = TrsShape<W>::kIterSyms, kTrsWarps = W;
    constexpr int BPC = L::BPC, D = L::D, NE = L::NE, BPS = BPC * D;
    extern __shared__ __align__(128) uint8_t trs_smem[];
    float2 *s_tr = reinterpret_cast<float2 *>(trs_smem + L::kTr);
    uint8_t *s_car = trs_smem + L::kCar;
    uint8_t *s_bits = trs_smem + L::kBits;
    float2 *s_lut = reinterpret_cast<float2 *>(trs_smem + L::kLut);
    uint8_t *s_enc = trs_smem + L::kEnc;
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(s_enc + 16);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(trs_smem + L::kMisc);
    uint64_t *s_paybar = reinterpret_cast<uint64_t *>(trs_smem + L::kMisc + 8);
    uint8_t *s_pay = trs_smem + L::kPay;
    TrsGeom *s_geom = reinterpret_cast<TrsGeom *>(trs_smem + L::kGeom);         // [2]: this CTA's share of the current / the next frame (thread 0 works it out)
    static_assert(2 * sizeof(TrsGeom) <= 256, "TrsGeom[2] must fit its slot");

    int tid = threadIdx.x;
    asm volatile("" : "+r"(tid));                                               // keep it in a register: under pressure ptxas re-reads SR_TID.X (S2R, an MIO instruction) inside the loops
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    const int C = a.group_ctas, G = a.n_groups;
    const int group = (int)blockIdx.x / C, rank = (int)blockIdx.x - group * C;
    const uint32_t arrivals = (uint32_t)(C * kTrsWarps);                        // per frame: every warp of the group, once

    // ---- set-up: tables, tensor memory ------------------------------------------------------------------------------------
    // Constellation table as in tx_tile_kernel, but CONJUGATED: the inverse transform runs as conj . FFT . conj here instead
    // of swap . FFT . swap. The two give the same numbers, bit for bit except the sign of an exact zero (swap(x) = j conj(x),
    // and the packed FFT commutes with a multiplication by j exactly: every add / multiply / fma is sign-symmetric; only an
    // exact cancellation, which rounds to +0 on either side, breaks the symmetry), but the closing conj is free -- the
    // scaling multiplies by (s, -s) -- whereas the closing swap costs a register move per value around the tensor-memory slot.
    for (int e = tid; e < 16 * (NE + 2); e += kTrsThreads) {
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
    if (tid < 16) s_enc[tid] = (uint8_t)ham74_encode_nibble(tid);
    if (FEC && tid < 256) s_enc14[tid] = (uint16_t)(ham74_encode_nibble(tid & 15) | (ham74_encode_nibble(tid >> 4) << 7));
    if (tid == 0) { mbar_init(s_paybar, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_addr(s_tmem)), "n"(TrsShape<W>::kCols) : "memory");
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
    const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(64 * (warp >> 2));
    uint32_t taddr_w = taddr0;
    int sym_in_iter = 4 * warp + g;
    asm volatile("" : "+r"(taddr_w), "+r"(sym_in_iter));                        // otherwise re-derived from tid in every iteration (7 instructions per symbol, measured)
    const uint8_t *car = s_car + (GUARD ? l - 7 : l);

    // (B1) coded bit stream of this CTA's chunk of a frame (same construction as a tile of tx_tile_kernel, the chunk is the tile)
    auto build_bits = [&](const TrsGeom &q, uint32_t stream) {
        if (q.t0 >= q.t1) return;                                              // (also when the frame does not fit iq_stride)
        const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
        const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
        const bool pf = q.pf_on != 0;
        const int pf_off = q.pf_off;                                           // payload byte i of the chunk's span is s_pay[i - pf_off]
        auto pay_byte = [&](uint32_t i) -> uint32_t { return pf ? (uint32_t)s_pay[(int)i - pf_off] : (uint32_t)pay[i]; };
        const uint32_t n = q.n;
        const uint32_t byte0 = (uint32_t)((long)q.t0 * BPS / 8), nbyte = (uint32_t)((long)(q.t1 - q.t0) * BPS / 8);
        if (FEC) {
            const uint32_t hdr = byte0 < 16 ? 16 - byte0 : 0;
            if (tid < (int)hdr) s_bits[tid] = (uint8_t)frame_byte<FEC>(pay, n, q.coded_len, byte0 + tid, s_enc);
            const uint32_t c0 = byte0 + hdr - 16;
            const uint32_t ngrp = ((nbyte + 2 - hdr + 6) / 7 + 3) / 4;
            for (uint32_t u = tid; u < ngrp; u += kTrsThreads) {
                const uint32_t pb = (c0 / 7) * 4 + 16 * u;
                uint32_t v[4] = { 0, 0, 0, 0 };
                if (pay_aligned && pb + 16 <= n) {
                    if (pf) {
#pragma unroll
                        for (int w = 0; w < 4; w++) v[w] = *reinterpret_cast<const uint32_t *>(s_pay + ((int)pb - pf_off) + 4 * w);
                    } else {
#pragma unroll
                        for (int w = 0; w < 4; w++) v[w] = __ldg(reinterpret_cast<const uint32_t *>(pay + pb) + w);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < 16; w++) if (pb + w < n) v[w >> 2] |= pay_byte(pb + w) << (8 * (w & 3));
                }
                uint64_t w[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t lo = (uint32_t)s_enc14[v[i] & 255u] | ((uint32_t)s_enc14[(v[i] >> 8) & 255u] << 14);
                    const uint32_t hi = (uint32_t)s_enc14[(v[i] >> 16) & 255u] | ((uint32_t)s_enc14[v[i] >> 24] << 14);
                    w[i] = (uint64_t)lo | ((uint64_t)hi << 28);
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
            for (uint32_t b = tid; b < nbyte + 2; b += kTrsThreads) {          // header, then the payload bytes as they are (src/transmitter.rs:37-47)
                const uint32_t B = byte0 + b;
                s_bits[b] = (uint8_t)(B < 16 ? (B < 8 ? (uint32_t)(q.coded_len >> (8 * B)) & 255u : 0u) : (B - 16 < n ? pay_byte(B - 16) : 0u));
            }
        }
    };
    // (B2) bit stream -> one byte per data carrier (modulate, src/transmitter.rs:108-140); carriers past the frame's last
    // constellation symbol are padding (encode_block's exhausted iterator, src/transmitter.rs:144-165): the null entry
    auto unpack_carriers = [&](const TrsGeom &q) {
        const long ncar_local = (long)q.ncar - (long)q.t0 * D;
        long have = (long)(q.t1 - q.t0) * D;
        const int total = (int)have;
        if (ncar_local < have) have = ncar_local;
        const int ncar_have = (int)have;
        const uint32_t *bits32 = reinterpret_cast<const uint32_t *>(s_bits);
        constexpr uint32_t M = (uint32_t)(NE - 1);
        auto spread4 = [&](uint32_t v) -> uint32_t {                           // 4 x BPC bits -> 4 carrier bytes
            return (v & M) | (((v >> BPC) & M) << 8) | (((v >> (2 * BPC)) & M) << 16) | (((v >> (3 * BPC)) & M) << 24);
        };
        auto pad4 = [&](uint32_t packed, int c4) -> uint32_t {                 // the frame's last carriers / padding rows (rare)
            if (c4 + 4 > ncar_have) {
#pragma unroll
                for (int w = 0; w < 4; w++) if (c4 + w >= ncar_have) packed = (packed & ~(0xFFu << (8 * w))) | ((uint32_t)NE << (8 * w));
            }
            return packed;
        };
        if (BPC == 6) {
            // 16 carriers = 96 bits = 3 aligned words in, 4 words (one 16-byte store) out
            for (int c16 = 16 * tid; c16 < total; c16 += 16 * kTrsThreads) {
                const uint32_t *bw = bits32 + (c16 >> 4) * 3;
                const uint32_t b0 = bw[0], b1 = bw[1], b2 = bw[2];
                uint4 o;
                o.x = pad4(spread4(b0), c16);
                o.y = pad4(spread4(__funnelshift_r(b0, b1, 24)), c16 + 4);
                o.z = pad4(spread4(__funnelshift_r(b1, b2, 16)), c16 + 8);
                o.w = pad4(spread4(b2 >> 8), c16 + 12);
                *reinterpret_cast<uint4 *>(s_car + c16) = o;                   // (total is a multiple of 16: D = 48 or 64 carriers per symbol)
            }
        } else {
            for (int c4 = 4 * tid; c4 < total; c4 += 4 * kTrsThreads) {          // 4 carriers = 4 BPC bits from bit 4 BPC (c4 / 4)
                const uint32_t bit = (uint32_t)c4 * BPC, wi = bit >> 5, sh = bit & 31;
                const uint32_t v = __funnelshift_r(bits32[wi], bits32[wi + 1], sh);
                *reinterpret_cast<uint32_t *>(s_car + c4) = pad4(spread4(v), c4);
            }
        }
    };
    // frame head (lock | preamble x4 | training x5) and zero fill past the frame, once the frame maximum is known
    auto write_head = [&](uint32_t stream, bool fits, uint32_t frame_len, float fmx) {
        float2 *out = a.iq + (size_t)stream * a.iq_stride;
        if (rank == 0)
            for (uint32_t i = tid; i < (uint32_t)(kHeadSyms * kSym) && i < a.iq_stride; i += kTrsThreads) {
                float2 v = make_float2(0.0f, 0.0f);
                if (fits) { v = a.tables->head[i]; v.x = v.x / fmx; v.y = v.y / fmx; }
                out[i] = v;
            }
        if (rank == C - 1) {
            const uint32_t z0 = fits ? frame_len : (uint32_t)(kHeadSyms * kSym);
            for (uint32_t i = z0 + tid; i < a.iq_stride; i += kTrsThreads) out[i] = make_float2(0.0f, 0.0f);
        }
    };
    // the frame's maximum, once every warp of the group has published its own (normalize, src/transmitter.rs:183-194)
    auto frame_max = [&](uint32_t stream) -> float {
        while (ld_relaxed_gpu(a.stream_cnt + stream) < arrivals) __nanosleep(64);
        return fmaxf(__int_as_float((int)ld_relaxed_gpu(reinterpret_cast<const uint32_t *>(a.stream_max) + stream)), head_max);
    };
    // drain slot `it` of the previous frame: scale, store with the cyclic prefix (prefix_block, src/transmitter.rs:168-181)
    auto drain = [&](int it, int p_nsym, float2 *p_out, float p_scale) {
        cpx y[8];
        tmem_ld16(taddr_w + 16u * (uint32_t)it, y);
        const int sl = it * kTrsIterSyms + sym_in_iter;
        if (sl < p_nsym) {
            unsigned long long *sym = reinterpret_cast<unsigned long long *>(p_out + (size_t)sl * kSym + l);
            const cpx sc = c_make(p_scale, -p_scale);                          // conj and scale in one
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                const unsigned long long v = c_mul2(y[kb], sc).v;              // time index l + 8 kb
                sym[kCp + 8 * kb] = v;
                if (kb >= 6) sym[8 * kb - (kNfft - kCp)] = v;                  // cyclic prefix = last 16 samples
            }
        }
    };

    // previous frame: iterations of this warp still in tensor memory, symbols of the chunk, where they go
    int p_nit = 0, p_nsym = 0;
    float2 *p_out = nullptr;
    uint32_t p_stream = 0, p_flen = 0;
    bool p_fits = false, have_prev = false;

    uint32_t stream = (uint32_t)group;                                         // (the launcher guarantees group < n_streams)
    if (tid == 0) s_geom[0] = trs_geometry<BPC, D, FEC, TrsShape<W>::kChunkMax>(a, stream, rank);
    __syncthreads();
    build_bits(s_geom[0], stream);
    __syncthreads();
    unpack_carriers(s_geom[0]);
    __syncthreads();

    for (int k = 0; ; k++) {
        // Frame k's geometry was published at least one CTA barrier ago; everything this iteration needs of it goes into
        // registers now, because thread 0 is about to put frame k+1's geometry into the other slot, and the slot after that
        // -- this one -- is rewritten an iteration later, possibly while a slow warp is still at the end of this iteration.
        const TrsGeom &q = s_geom[k & 1];
        TrsGeom &qn = s_geom[(k + 1) & 1];
        const int q_t0 = q.t0, nsym = q.t1 - q.t0;
        const uint32_t q_flen = q.frame_len;
        const bool q_fits = q.fits;
        if (rank == 0 && tid == 0 && a.frame_len) a.frame_len[stream] = q_flen;
        int n_it = nsym - 4 * warp;                                            // iterations in which this warp has at least one symbol
        n_it = n_it > 0 ? (n_it + kTrsIterSyms - 1) / kTrsIterSyms : 0;
        // ---- the next frame of the group: its geometry, and its payload bytes on their way into shared memory -----------------
        const uint32_t next = stream + (uint32_t)G;
        const bool more = next < a.n_streams;
        if (more && tid == 0) {
            uint32_t pf_bytes = 0;
            uintptr_t pf_src = 0;
            TrsGeom g2 = trs_geometry<BPC, D, FEC, TrsShape<W>::kChunkMax>(a, next, rank);
            if (g2.t0 < g2.t1) {
                const long byte0 = (long)g2.t0 * BPS / 8, nbyte = (long)(g2.t1 - g2.t0) * BPS / 8;
                long lo, hi;
                if (FEC) {
                    const long hdr = byte0 < 16 ? 16 - byte0 : 0, c0 = byte0 + hdr - 16;
                    lo = (c0 / 7) * 4;
                    hi = lo + 16 * (((nbyte + 2 - hdr + 6) / 7 + 3) / 4);
                } else {
                    lo = byte0 < 16 ? 0 : byte0 - 16;
                    hi = byte0 + nbyte + 2 - 16;
                }
                if (hi > (long)g2.n) hi = (long)g2.n;
                const uint8_t *pay = a.payload + (size_t)next * a.payload_stride;
                const uintptr_t s0 = reinterpret_cast<uintptr_t>(pay + lo) & ~(uintptr_t)15, s1 = (reinterpret_cast<uintptr_t>(pay + hi) + 15) & ~(uintptr_t)15;
                if (lo < hi && s0 >= reinterpret_cast<uintptr_t>(a.payload) && s1 - s0 <= L::kPayBytes &&
                    s1 <= reinterpret_cast<uintptr_t>(a.payload + (size_t)a.n_streams * a.payload_stride)) {
                    g2.pf_on = 1;
                    g2.pf_off = (int)((long)s0 - (long)reinterpret_cast<uintptr_t>(pay));
                    pf_bytes = (uint32_t)(s1 - s0);
                    pf_src = s0;
                }
            }
            qn = g2;
            // one mbarrier phase per frame: its completion publishes the geometry (release / acquire) and, when the payload
            // bytes are prefetched, also means that they have landed
            mbar_arrive_expect_tx(s_paybar, pf_bytes);
            if (pf_bytes) tma_bulk_g2s(s_pay, reinterpret_cast<const void *>(pf_src), pf_bytes, s_paybar);
        }
        // ---- (A) drain frame k-1, transform frame k ------------------------------------------------------------------------
        float p_fmx = 1.0f, p_scale = 0.0f;
        if (have_prev) { p_fmx = frame_max(p_stream); p_scale = (1.0f / 64.0f) * (1.0f / p_fmx); }
        tmem_wait_st();                                                        // the slots of frame k-1 were written a whole phase ago
        float mx = 0.0f;
#pragma unroll 1
        for (int it = 0; it < kTrsSlots; it++) {
            const int sl = it * kTrsIterSyms + sym_in_iter;
            const bool valid = sl < nsym;
            uint32_t idx[8];
            if (it < p_nit) drain(it, p_nsym, p_out, p_scale);
            if (it < n_it) {
                const uint8_t *rowp = car + (valid ? sl : 0) * D;
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
        // ---- publish this warp's maximum of frame k, count the arrival -------------------------------------------------------
        mx *= 1.0f / 64.0f;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0) publish_max_and_arrive(a.stream_max + stream, __float_as_int(fmaxf(mx, 0.0f)), a.stream_cnt + stream);
        // ---- (B) carrier bytes of frame k+1, head / zero fill of frame k-1 ----------------------------------------------------
        // (a warp starts on the bit stream as soon as its own transforms are done: s_bits and s_pay are not touched by phase A)
        if (more) { mbar_wait(s_paybar, (uint32_t)(k & 1)); build_bits(qn, next); }
        if (have_prev) write_head(p_stream, p_fits, p_flen, p_fmx);
        __syncthreads();                                                       // every warp is done with the carrier bytes of frame k; the bit stream is complete
        if (more) unpack_carriers(qn);
        __syncthreads();
        have_prev = true; p_nit = n_it; p_nsym = nsym; p_stream = stream; p_flen = q_flen; p_fits = q_fits;
        p_out = a.iq + (size_t)stream * a.iq_stride + (size_t)(kHeadSyms + q_t0) * kSym;
        if (!more) break;
        stream = next;
    }
    // ---- the group's last frame ---------------------------------------------------------------------------------------------
    {
        const float p_fmx = frame_max(p_stream), p_scale = (1.0f / 64.0f) * (1.0f / p_fmx);
        tmem_wait_st();
#pragma unroll 1
        for (int it = 0; it < p_nit; it++) drain(it, p_nsym, p_out, p_scale);
        write_head(p_stream, p_fits, p_flen, p_fmx);
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TrsShape<W>::kCols) : "memory");
}

}  // namespace ofdm
