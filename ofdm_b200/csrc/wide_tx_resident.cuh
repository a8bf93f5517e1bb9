// wide_tx_resident.cuh -- one-pass transmit kernel of the 1024-subcarrier variant for large batches (BASELINE.json configs[3],
// docs/SPEC.md 9): `encode` (src/transmitter.rs:11-58) with every OFDM symbol transformed ONCE.
//
// `normalize` (src/transmitter.rs:183-194) divides the frame by its maximum positive component, so nothing can be stored
// before the whole frame has been transformed; wide_tx_kernel therefore runs every inverse FFT twice (maximum pass, store
// pass). Here a frame belongs to a GROUP of C persistent CTAs (one per SM; formed by block index like tx_resident_kernel's,
// launched cooperatively so that the whole grid is resident), and the un-normalised symbols wait in the SMs' TENSOR MEMORY
// (tcgen05.st / tcgen05.ld, SASS STTM / LDTM): 512 columns x 128 lanes x 4 B = 256 kB = exactly 32 symbols of 1024 complex
// samples per SM, so the 128-symbol frames of the bench workload take C = 4.
//
// One OFDM symbol per WARP, as in wide_decode_kernel: the inverse transform is conj . FFT . conj with the register-resident
// 32 x 32 split (lane l holds bins l + 32 j; 32-point DFT in registers, twiddle W1024^(l k1), one 32 x 32 exchange through a
// warp-private shared-memory buffer, second 32-point DFT -> lane t1 holds the samples t1 + 32 t2). The constellation table
// is stored conjugated, the closing conj rides on the scaling (multiply by (s, -s)). A warp's 64 result registers go to its
// own tensor-memory slot (its lane quadrant x 64 columns): a slot is only ever touched by the warp that wrote it, and NOTHING
// in the frame loop needs a CTA barrier -- a warp builds the coded bit stream of its own symbols (Hamming(7,4) over 16-byte
// payload groups, docs/SPEC.md 3), looks the carriers up, transforms, and later drains its slot with coalesced 256-byte
// row stores (cyclic prefix = the last 8 rows once more). Per frame k a warp runs:
//   wait for the maximum of frame k-1 | 2 x { drain slot i of frame k-1 | transform its symbol i of frame k into slot i } |
//   publish its maximum of frame k (one flagged word per warp, collected by the others in one L2 round trip) |
//   bit streams of its symbols of frame k+1, its share of the head / zero fill of frame k-1
// -- the last step is also the time the maximum of frame k needs to travel between the SMs of the group.
#pragma once

#include "wide_kernels.cuh"
#include "tx_resident.cuh"

namespace ofdm {
namespace wide {

constexpr int kWTrsWarps = 16;                                   // 128 registers per thread x 512 threads = the register file
constexpr int kWTrsThreads = 32 * kWTrsWarps;
constexpr int kWTrsSlots = 2;                                    // symbols per warp and frame (64 tensor-memory columns each)
constexpr int kWTrsSyms = kWTrsWarps * kWTrsSlots;               // 32 symbols per CTA = all 512 columns
constexpr int kWTrsBitsBuf = 832;                                // per warp: header 16 + 28 groups x 28 coded bytes (+ slack), multiple of 64
constexpr int kWTrsHeadPer = 6;                                  // head samples per thread held in shared memory (a group of 4 CTAs: 12 800 / 2048 = 6.25; the rest is read through the cache)
constexpr int kWTrsCarBuf = 1024 + 16;                           // per slot: one byte per data carrier | null entry | pilot entry (+ slack)

template <int MOD> struct WTrsSmem {
    static constexpr int NE = 1 << ModTraits<MOD>::kBpc;
    static constexpr size_t kBuf = 0;                                                      // [warp][kWBuf]: 32 x 32 exchange
    static constexpr size_t kTw = kBuf + sizeof(float2) * kWTrsWarps * kWBuf;              // [l][kWPitch]: W1024^(l k1)
    static constexpr size_t kOff = kTw + sizeof(float2) * kWBuf;                           // [j][l]: index of bin l + 32 j in the carrier bytes (D: null, D + 1: pilot)
    static constexpr size_t kLut = kOff + sizeof(uint16_t) * kN;                           // [entry][lane & 15] conjugated constellation, null, pilot
    static constexpr size_t kEnc = kLut + sizeof(float2) * 16 * (NE + 2);                  // Hamming byte table (256 x u16), tensor-memory base address
    static constexpr size_t kBits = kEnc + 512 + 64;                                       // [warp][kWTrsBitsBuf]: coded bit stream of the symbol being prepared
    static constexpr size_t kCar = kBits + (size_t)kWTrsWarps * kWTrsBitsBuf;              // [warp][slot][kWTrsCarBuf]
    static constexpr size_t kHead = kCar + (size_t)kWTrsWarps * kWTrsSlots * kWTrsCarBuf;  // [kWTrsHeadPer][thread]: this CTA's share of the frame head
    static constexpr size_t kTotal = kHead + sizeof(float2) * kWTrsHeadPer * kWTrsThreads;
};

// a warp reads / writes one whole symbol slot: 32 lanes x 64 columns = its 64 registers (ONE tensor-memory access and one wait
// instead of four of each)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, cpx (&x)[32])
{
    uint32_t r[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) x[i].v = (unsigned long long)r[2 * i] | ((unsigned long long)r[2 * i + 1] << 32);
}
__device__ __forceinline__ void tmem_st64(uint32_t taddr, const cpx (&x)[32])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};"
                 :: "r"(taddr), "r"((uint32_t)x[0].v), "r"((uint32_t)(x[0].v >> 32)), "r"((uint32_t)x[1].v), "r"((uint32_t)(x[1].v >> 32)), "r"((uint32_t)x[2].v), "r"((uint32_t)(x[2].v >> 32)), "r"((uint32_t)x[3].v), "r"((uint32_t)(x[3].v >> 32)), "r"((uint32_t)x[4].v), "r"((uint32_t)(x[4].v >> 32)), "r"((uint32_t)x[5].v), "r"((uint32_t)(x[5].v >> 32)), "r"((uint32_t)x[6].v), "r"((uint32_t)(x[6].v >> 32)), "r"((uint32_t)x[7].v), "r"((uint32_t)(x[7].v >> 32)), "r"((uint32_t)x[8].v), "r"((uint32_t)(x[8].v >> 32)), "r"((uint32_t)x[9].v), "r"((uint32_t)(x[9].v >> 32)), "r"((uint32_t)x[10].v), "r"((uint32_t)(x[10].v >> 32)), "r"((uint32_t)x[11].v), "r"((uint32_t)(x[11].v >> 32)), "r"((uint32_t)x[12].v), "r"((uint32_t)(x[12].v >> 32)), "r"((uint32_t)x[13].v), "r"((uint32_t)(x[13].v >> 32)), "r"((uint32_t)x[14].v), "r"((uint32_t)(x[14].v >> 32)), "r"((uint32_t)x[15].v), "r"((uint32_t)(x[15].v >> 32)), "r"((uint32_t)x[16].v), "r"((uint32_t)(x[16].v >> 32)), "r"((uint32_t)x[17].v), "r"((uint32_t)(x[17].v >> 32)), "r"((uint32_t)x[18].v), "r"((uint32_t)(x[18].v >> 32)), "r"((uint32_t)x[19].v), "r"((uint32_t)(x[19].v >> 32)), "r"((uint32_t)x[20].v), "r"((uint32_t)(x[20].v >> 32)), "r"((uint32_t)x[21].v), "r"((uint32_t)(x[21].v >> 32)), "r"((uint32_t)x[22].v), "r"((uint32_t)(x[22].v >> 32)), "r"((uint32_t)x[23].v), "r"((uint32_t)(x[23].v >> 32)), "r"((uint32_t)x[24].v), "r"((uint32_t)(x[24].v >> 32)), "r"((uint32_t)x[25].v), "r"((uint32_t)(x[25].v >> 32)), "r"((uint32_t)x[26].v), "r"((uint32_t)(x[26].v >> 32)), "r"((uint32_t)x[27].v), "r"((uint32_t)(x[27].v >> 32)), "r"((uint32_t)x[28].v), "r"((uint32_t)(x[28].v >> 32)), "r"((uint32_t)x[29].v), "r"((uint32_t)(x[29].v >> 32)), "r"((uint32_t)x[30].v), "r"((uint32_t)(x[30].v >> 32)), "r"((uint32_t)x[31].v), "r"((uint32_t)(x[31].v >> 32))
                 : "memory");
}

// this CTA's share of frame `stream`: symbols [t0, t1) of its S data symbols (an even split over the group)
struct WTrsGeom {
    uint32_t n;
    uint64_t coded_len, ncar;
    int      S, t0, t1;
    uint32_t frame_len;
    bool     fits;
};
template <int BPC, int D, bool FEC, int PER>
__device__ __forceinline__ WTrsGeom wtrs_geometry(const WideTxArgs &a, uint32_t stream, int rank)
{
    WTrsGeom q;
    q.n = __ldg(a.payload_len + stream);
    q.coded_len = FEC ? (14ull * q.n + 7) / 8 : q.n;
    const uint64_t nbits = kHeaderBits + 8 * q.coded_len;
    q.ncar = (nbits + BPC - 1) / BPC;                               // constellation symbols (src/transmitter.rs:108-140)
    const uint64_t S64 = (q.ncar + D - 1) / D;                      // OFDM data symbols (src/transmitter.rs:49-54)
    q.S = (int)S64;
    q.frame_len = (10u + (uint32_t)q.S) * kL;
    q.fits = (10ull + S64) * kL <= (uint64_t)a.iq_stride;
    const int C = a.group_ctas;
    const int chunk = (q.S + C - 1) / C;
    q.t0 = rank * chunk;
    q.t1 = q.t0 + chunk < q.S ? q.t0 + chunk : q.S;
    if (!q.fits || q.t1 < q.t0 || chunk > PER) q.t1 = q.t0;         // (the launcher sizes C so that a fitting frame's chunk never exceeds the slots)
    return q;
}

// Stream bytes [s BPSB, (s + 1) BPSB + 2) of the frame byte stream [header 16 B | (Hamming-coded) payload] into a warp-private
// buffer; returns the buffer's BIT index of the symbol's first bit. With FEC the buffer starts on a 7-byte unit of the coded
// stream (7 coded bytes = 8 nibbles = 4 payload bytes); lane u encodes payload bytes [16 u, 16 u + 16) into 28 coded bytes.
template <int BPSB, bool FEC>
__device__ __forceinline__ uint32_t wtrs_build_bits(uint8_t *bits, const uint8_t *__restrict__ pay, bool pay_aligned, uint32_t n, uint64_t coded_len,
                                                    uint32_t B0, const uint16_t *s_enc14, int lane)
{
    // B0 = first stream byte wanted
    constexpr uint32_t nbyte = BPSB + 2;
    if (FEC) {
        const uint32_t hdr = B0 < 16 ? 16 - B0 : 0;                  // header bytes inside this span
        const uint32_t pad = (0u - hdr) & 3u;                        // keeps the 28-byte groups behind them on 4-byte boundaries
        bits += pad;
        if ((uint32_t)lane < hdr) bits[lane] = (uint8_t)(B0 + lane < 8 ? (uint32_t)(coded_len >> (8 * (B0 + lane))) & 255u : 0u);
        const uint32_t c0 = B0 + hdr - 16;                           // first coded byte of the symbol
        const uint32_t u0 = c0 / 7, skip = c0 - 7 * u0;
        const uint32_t ngrp = ((nbyte - hdr + skip + 6) / 7 + 3) / 4;
        for (uint32_t u = lane; u < ngrp; u += 32) {
            const uint32_t pb = 4 * u0 + 16 * u;
            uint32_t v[4] = { 0, 0, 0, 0 };
            if (pay_aligned && pb + 16 <= n) {
#pragma unroll
                for (int w = 0; w < 4; w++) v[w] = __ldg(reinterpret_cast<const uint32_t *>(pay + pb) + w);
            } else {
#pragma unroll
                for (int w = 0; w < 16; w++) if (pb + w < n) v[w >> 2] |= (uint32_t)pay[pb + w] << (8 * (w & 3));
            }
            uint64_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t lo = (uint32_t)s_enc14[v[i] & 255u] | ((uint32_t)s_enc14[(v[i] >> 8) & 255u] << 14);
                const uint32_t hi = (uint32_t)s_enc14[(v[i] >> 16) & 255u] | ((uint32_t)s_enc14[v[i] >> 24] << 14);
                w[i] = (uint64_t)lo | ((uint64_t)hi << 28);
            }
            uint32_t *dst = reinterpret_cast<uint32_t *>(bits + hdr + 28 * u);
            dst[0] = (uint32_t)w[0];
            dst[1] = (uint32_t)(w[0] >> 32) | ((uint32_t)w[1] << 24);
            dst[2] = (uint32_t)(w[1] >> 8);
            dst[3] = (uint32_t)(w[1] >> 40) | ((uint32_t)w[2] << 16);
            dst[4] = (uint32_t)(w[2] >> 16);
            dst[5] = (uint32_t)(w[2] >> 48) | ((uint32_t)w[3] << 8);
            dst[6] = (uint32_t)(w[3] >> 24);
        }
        return 8u * (hdr ? pad : skip);
    } else {
        for (uint32_t b = lane; b < nbyte; b += 32) {                // header, then the payload bytes as they are (src/transmitter.rs:37-47)
            const uint32_t B = B0 + b;
            bits[b] = (uint8_t)(B < 16 ? (B < 8 ? (uint32_t)(coded_len >> (8 * B)) & 255u : 0u) : (B - 16 < n ? (uint32_t)pay[B - 16] : 0u));
        }
        return 0u;
    }
}

// bit stream -> one byte per data carrier (modulate, src/transmitter.rs:108-140); carriers past the frame's last constellation
// symbol are padding (encode_block's exhausted iterator, src/transmitter.rs:144-165): the null entry, which is also car[D];
// car[D + 1] is the pilot entry -- the bin table points there, so the transform needs no special cases
template <int BPC, int D>
__device__ __forceinline__ void wtrs_unpack_carriers(uint8_t *car, const uint8_t *bits, uint32_t bit0, long left, int lane)
{
    constexpr uint32_t NE = 1u << BPC, M = NE - 1u;
    const int have = left >= D ? D : (left > 0 ? (int)left : 0);
    const uint32_t *bits32 = reinterpret_cast<const uint32_t *>(bits);
#pragma unroll 2
    for (int c4 = 4 * lane; c4 < D; c4 += 128) {                               // 4 carriers = 4 BPC bits
        const uint32_t bit = bit0 + (uint32_t)c4 * BPC, wi = bit >> 5, sh = bit & 31u;
        const uint32_t v = __funnelshift_r(bits32[wi], bits32[wi + 1], sh);
        uint32_t packed = (v & M) | (((v >> BPC) & M) << 8) | (((v >> (2 * BPC)) & M) << 16) | (((v >> (3 * BPC)) & M) << 24);
        if (c4 + 4 > have) {
#pragma unroll
            for (int w = 0; w < 4; w++) if (c4 + w >= have) packed = (packed & ~(0xFFu << (8 * w))) | (NE << (8 * w));
        }
        *reinterpret_cast<uint32_t *>(car + c4) = packed;
    }
    if (lane == 0) *reinterpret_cast<uint16_t *>(car + D) = (uint16_t)(NE | ((NE + 1u) << 8));
}

// DB = false: a warp holds TWO symbols of a frame (32 symbols per CTA); a frame's slots are drained while the next frame is
//             transformed into them, so every warp waits for the frame maximum right after the group's last transform.
// DB = true : a warp holds ONE symbol of each of TWO consecutive frames (16 symbols per CTA and frame, twice the CTAs per
//             group): frame k is transformed into slot k & 1 BEFORE frame k-1 is drained from the other slot -- the maximum of
//             frame k-1 has had a whole transform's time to arrive, and the warps of a group drift apart by up to a frame
//             instead of marching in phase (all storing, then all transforming).
template <int MOD, bool GUARD, bool FEC, bool DB>
__global__ void __launch_bounds__(kWTrsThreads, 1) wide_tx_resident_kernel(const WideTxArgs a)
{
    typedef WTrsSmem<MOD> L;
    constexpr int BPC = ModTraits<MOD>::kBpc, NE = 1 << BPC, D = GUARD ? 768 : 1024, BPSB = BPC * D / 8;
    constexpr int SPW = DB ? 1 : kWTrsSlots;                                    // symbols per warp and frame
    static_assert(16 + 28 * (((BPSB + 2 + 6 + 6) / 7 + 3) / 4) <= kWTrsBitsBuf - 8, "bit-stream buffer of one symbol");
    extern __shared__ __align__(128) uint8_t wtrs_smem[];
    float2 *s_buf = reinterpret_cast<float2 *>(wtrs_smem + L::kBuf);
    float2 *s_tw = reinterpret_cast<float2 *>(wtrs_smem + L::kTw);
    uint16_t *s_off = reinterpret_cast<uint16_t *>(wtrs_smem + L::kOff);
    float2 *s_lut = reinterpret_cast<float2 *>(wtrs_smem + L::kLut);
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(wtrs_smem + L::kEnc);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(wtrs_smem + L::kEnc + 512);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = a.group_ctas, G = a.n_groups;
    const int group = (int)blockIdx.x / C, rank = (int)blockIdx.x - group * C;
    const uint32_t arrivals = (uint32_t)(C * kWTrsWarps);                       // per frame: every warp of the group, once

    // ---- set-up: tables, tensor memory ------------------------------------------------------------------------------------
    for (int e = tid; e < 16 * (NE + 2); e += kWTrsThreads) {                   // conjugated constellation (conj . FFT . conj), null, pilot
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
    for (int e = tid; e < kN; e += kWTrsThreads) {                              // e = 32 j + l <-> bin l + 32 j = e (encode_block, src/transmitter.rs:144-165)
        const int rk = w_rank<GUARD>(e);
        s_off[e] = (uint16_t)(rk >= 0 ? rk : ((GUARD && w_is_pilot(e)) ? D + 1 : D));
        const int r = e >> 5, c = e & 31;
        s_tw[r * kWPitch + c] = __ldg(a.tables->w1024 + ((r * c) & (kN - 1)));
    }
    if (FEC && tid < 256) s_enc14[tid] = (uint16_t)(ham74_encode_nibble(tid & 15) | (ham74_encode_nibble(tid >> 4) << 7));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_addr(s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(s_tmem);
    const float head_max = a.tables->head_max;

    // this warp's tensor-memory slots: lane quadrant warp % 4, columns 128 (warp / 4) + 64 slot
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 * (warp >> 2));
    float2 *buf = s_buf + warp * kWBuf;
    unsigned long long *wr = reinterpret_cast<unsigned long long *>(buf + lane);
    const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(buf + lane * kWPitch);
    const ulonglong2 *tw_row = reinterpret_cast<const ulonglong2 *>(s_tw + lane * kWPitch);
    const unsigned long long *lut = reinterpret_cast<const unsigned long long *>(s_lut) + (lane & 15);
    const uint16_t *off_col = s_off + lane;
    uint8_t *mybits = wtrs_smem + L::kBits + (size_t)warp * kWTrsBitsBuf;
    uint8_t *mycar = wtrs_smem + L::kCar + (size_t)warp * kWTrsSlots * kWTrsCarBuf;

    float mx = 0.0f;
    // carriers -> inverse FFT into tensor-memory slot `slot` (prefix_block's IFFT, src/transmitter.rs:168-181)
    auto transform = [&](int slot) {
        const uint8_t *car = mycar + slot * kWTrsCarBuf;
        cpx x[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
            if (!w_row_used<GUARD>(j)) { x[j] = c_make(0.0f, 0.0f); continue; }
            x[j].v = lut[(uint32_t)car[off_col[32 * j]] * 16u];
        }
        dft32_p(x);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const ulonglong2 ww = tw_row[q];
            cpx w0, w1;
            w0.v = ww.x; w1.v = ww.y;
            x[2 * q] = c_mul(x[2 * q], w0); x[2 * q + 1] = c_mul(x[2 * q + 1], w1);
        }
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) wr[k1 * kWPitch] = x[k1].v;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; q++) { const ulonglong2 v = row[q]; x[2 * q].v = v.x; x[2 * q + 1].v = v.y; }
        __syncwarp();
        dft32_p(x);
#pragma unroll
        for (int t2 = 0; t2 < 32; t2++) {
            float re, im;
            c_split(x[t2], re, im);                                            // the frame's sample is (re, -im) / 1024
            mx = fmaxf(mx, fmaxf(re, -im));
        }
        tmem_st64(taddr + (uint32_t)(64 * slot), x);
    };
    // drain a slot: scale, store with the cyclic prefix (prefix_block, src/transmitter.rs:168-181)
    auto drain = [&](int slot, int s, float2 *fout, float scale) {
        unsigned long long *sym = reinterpret_cast<unsigned long long *>(fout + (size_t)(10 + s) * kL) + lane;
        const cpx sc = c_make(scale, -scale);                                  // conj and scale in one
        cpx y[32];
        tmem_ld64(taddr + (uint32_t)(64 * slot), y);
#pragma unroll
        for (int t2 = 0; t2 < 32; t2++) {
            const unsigned long long v = c_mul2(y[t2], sc).v;                  // time index lane + 32 t2
            sym[kCpW + 32 * t2] = v;
            if (t2 >= 24) sym[32 * (t2 - 24)] = v;                             // cyclic prefix = last 256 samples
        }
    };
    // frame head (lock | preamble x4 | training x5) and zero fill past the frame, spread over all threads of the group
    // A thread writes the same head samples of every frame: groups of >= 4 CTAs keep their share of the (un-normalised) head
    // table in shared memory (no L2 round trip per frame), smaller groups read the table through the cache.
    const uint32_t gthreads = (uint32_t)(C * kWTrsThreads), gt = (uint32_t)(rank * kWTrsThreads + tid);
    const uint32_t hw = a.iq_stride < (uint32_t)kHeadW ? a.iq_stride : (uint32_t)kHeadW;
    const bool head_cached = C >= 4;
    float2 *s_head = reinterpret_cast<float2 *>(wtrs_smem + L::kHead) + tid;
    if (head_cached) {
#pragma unroll
        for (int u = 0; u < kWTrsHeadPer; u++) {
            const uint32_t i = gt + u * gthreads;
            s_head[u * kWTrsThreads] = i < (uint32_t)kHeadW ? __ldg(a.tables->head + i) : make_float2(0.0f, 0.0f);
        }
    }
    auto write_head = [&](uint32_t stream, bool fits, uint32_t frame_len, float fmx) {
        float2 *out = a.iq + (size_t)stream * a.iq_stride;
        const float rfmx = fits ? 1.0f / fmx : 0.0f;                           // a frame that does not fit iq_stride comes back zeroed
        uint32_t i0 = gt;
        if (head_cached) {
#pragma unroll
            for (int u = 0; u < kWTrsHeadPer; u++) {
                const uint32_t i = gt + u * gthreads;
                const float2 v = s_head[u * kWTrsThreads];
                if (i < hw) out[i] = make_float2(v.x * rfmx, v.y * rfmx);
            }
            i0 = gt + kWTrsHeadPer * gthreads;
        }
        for (; i0 < hw; i0 += 4 * gthreads) {                                  // four table loads in flight per thread
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * gthreads;
                v[u] = make_float2(0.0f, 0.0f);
                if (i < hw) v[u] = __ldg(a.tables->head + i);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + u * gthreads;
                if (i < hw) out[i] = make_float2(v[u].x * rfmx, v[u].y * rfmx);
            }
        }
        const uint32_t z0 = fits ? frame_len : (uint32_t)kHeadW;
        for (uint32_t i = z0 + gt; i < a.iq_stride; i += gthreads) out[i] = make_float2(0.0f, 0.0f);
    };
    // The frame maximum (normalize, src/transmitter.rs:183-194) travels through one 32-bit word per warp of the group,
    // stream_cnt[stream][rank * 16 + warp] = float bits of the warp's maximum (>= 0) | 0x80000000 as the "written" flag
    // (the array is zeroed before the launch): a plain store to publish, ONE round trip to the L2 to collect -- every lane
    // loads two of the words, the warp votes on the flags and reduces the values (no atomics, no second dependent load).
    const uint32_t n_words = arrivals;
    auto publish_max = [&](uint32_t stream, float m) {
        if (lane == 0)
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(a.stream_cnt + (size_t)stream * n_words + (uint32_t)(rank * kWTrsWarps + warp)),
                         "r"(__float_as_uint(fmaxf(m, 0.0f)) | 0x80000000u) : "memory");
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
#ifdef WTRS_EXPERIMENT_NO_WAIT
            break;
#endif
            __nanosleep(32);
        }
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
        return fmaxf(m, head_max);
    };
    // the carrier bytes of this warp's symbols of frame `stream` (geometry q), symbol i -> carrier buffer slot0 + i
    auto build = [&](const WTrsGeom &q, uint32_t stream, int slot0) {
        const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
        const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
#pragma unroll
        for (int i = 0; i < SPW; i++) {
            const int s = q.t0 + warp + kWTrsWarps * i;
            if (s < q.t1) {
                const uint32_t bit0 = wtrs_build_bits<BPSB, FEC>(mybits, pay, pay_aligned, q.n, q.coded_len, (uint32_t)s * BPSB, s_enc14, lane);
                __syncwarp();
                wtrs_unpack_carriers<BPC, D>(mycar + (slot0 + i) * kWTrsCarBuf, mybits, bit0, (long)q.ncar - (long)s * D, lane);
                __syncwarp();
            }
        }
    };
    // bring the payload bytes the next build() will read towards the SM (the build comes an iteration's work later)
    auto prefetch = [&](const WTrsGeom &q, uint32_t stream) {
        const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
#pragma unroll
        for (int i = 0; i < SPW; i++) {
            const int s = q.t0 + warp + kWTrsWarps * i;
            if (s < q.t1) {
                const uint32_t B0 = (uint32_t)s * BPSB, c0 = B0 < 16 ? 0 : B0 - 16;
                const uint32_t pb = FEC ? 4 * (c0 / 7) + 16 * lane : c0 + 32 * lane;
                constexpr uint32_t span = FEC ? (BPSB * 4) / 7 + 32 : BPSB + 32;
                if ((FEC ? 16u : 32u) * lane < span && pb < q.n)
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(pay + pb));
            }
        }
    };

    uint32_t stream = (uint32_t)group;                                         // (the launcher guarantees group < n_streams)
    WTrsGeom q = wtrs_geometry<BPC, D, FEC, kWTrsWarps * SPW>(a, stream, rank);
    build(q, stream, 0);
    bool have_prev = false, p_fits = false;
    int p_t0 = 0, p_t1 = 0;
    uint32_t p_stream = 0, p_flen = 0;

    for (int k = 0; ; k++) {
        if (rank == 0 && tid == 0 && a.frame_len) a.frame_len[stream] = q.frame_len;
        const uint32_t next = stream + (uint32_t)G;
        const bool more = next < a.n_streams;
        WTrsGeom qn = q;
        if (more) { qn = wtrs_geometry<BPC, D, FEC, kWTrsWarps * SPW>(a, next, rank); prefetch(qn, next); }
        float p_fmx = 1.0f;
        mx = 0.0f;
        if (DB) {
            // ---- transform frame k into slot k & 1, publish, prepare frame k+1, THEN drain frame k-1 from the other slot -----
            const int s = q.t0 + warp;
            if (s < q.t1) transform(k & 1);
            mx *= 1.0f / (float)kN;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
            publish_max(stream, mx);
            if (more) build(qn, next, (k + 1) & 1);
            if (have_prev) {
                p_fmx = frame_max(p_stream);
                tmem_wait_st();
                const int ps = p_t0 + warp;
                if (ps < p_t1) drain((k + 1) & 1, ps, a.iq + (size_t)p_stream * a.iq_stride, (1.0f / (float)kN) * (1.0f / p_fmx));
                write_head(p_stream, p_fits, p_flen, p_fmx);
            }
        } else {
            // ---- drain frame k-1, transform frame k ---------------------------------------------------------------------------
            float p_scale = 0.0f;
            if (have_prev) { p_fmx = frame_max(p_stream); p_scale = (1.0f / (float)kN) * (1.0f / p_fmx); }
            tmem_wait_st();                                                    // the slots of frame k-1 were written an iteration ago
            float2 *p_out = a.iq + (size_t)p_stream * a.iq_stride;
#pragma unroll 1
            for (int i = 0; i < SPW; i++) {
                const int ps = p_t0 + warp + kWTrsWarps * i, s = q.t0 + warp + kWTrsWarps * i;
                if (have_prev && ps < p_t1) drain(i, ps, p_out, p_scale);
                if (s < q.t1) transform(i);
            }
            mx *= 1.0f / (float)kN;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
            publish_max(stream, mx);
            // ---- carrier bytes of frame k+1, head / zero fill of frame k-1 (also the time the maximum of frame k needs to travel) ----
            __syncwarp();                                                      // every lane has read its carriers of frame k
            if (more) build(qn, next, 0);
            if (have_prev) write_head(p_stream, p_fits, p_flen, p_fmx);
        }
        have_prev = true; p_t0 = q.t0; p_t1 = q.t1; p_stream = stream; p_flen = q.frame_len; p_fits = q.fits;
        if (!more) break;
        stream = next;
        q = qn;
    }
    // ---- the group's last frame ---------------------------------------------------------------------------------------------
    {
        const float p_fmx = frame_max(p_stream), p_scale = (1.0f / (float)kN) * (1.0f / p_fmx);
        tmem_wait_st();
        float2 *p_out = a.iq + (size_t)p_stream * a.iq_stride;
        if (DB) {
            // p_stream was frame number (streams of this group so far) - 1; its slot parity is tracked by have_prev's loop index:
            // recomputed from the stream index
            const int kk = (int)((p_stream - (uint32_t)group) / (uint32_t)G);
            if (p_t0 + warp < p_t1) drain(kk & 1, p_t0 + warp, p_out, p_scale);
        } else {
#pragma unroll 1
            for (int i = 0; i < SPW; i++) {
                const int ps = p_t0 + warp + kWTrsWarps * i;
                if (ps < p_t1) drain(i, ps, p_out, p_scale);
            }
        }
        write_head(p_stream, p_fits, p_flen, p_fmx);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}


// ---- speculative one-pass kernel: no residency ----------------------------------------------------------------------------------
// Same bet as tx_spec_kernel (tx_warp.cuh): the frame maximum `normalize` divides by is max(head_max, data maximum), head_max is
// a constant, and at nfft = 1024 the data symbols of a scrambled payload reach about a fifth of it. Every symbol is transformed
// once, scaled with head_max and stored at once -- no tensor memory, no exchange between SMs, no cooperative launch -- and a warp
// whose data beat head_max records its maximum (atomicMax on stream_max[stream], 0 otherwise). Those frames are redone by the
// two-pass kernel's store pass with the recorded maximum (wide_tx_kernel<WRITE>, redo_only: every other frame's CTAs exit at
// once). Work unit = (frame, span of 16 symbols: one per warp), persistent CTAs stride over the units.
template <int MOD, bool GUARD, bool FEC>
__global__ void __launch_bounds__(kWTrsThreads, 1) wide_tx_spec_kernel(const WideTxArgs a)
{
    typedef WTrsSmem<MOD> L;
    constexpr int BPC = ModTraits<MOD>::kBpc, NE = 1 << BPC, D = GUARD ? 768 : 1024, BPSB = BPC * D / 8;
    extern __shared__ __align__(128) uint8_t wtrs_smem[];
    float2 *s_buf = reinterpret_cast<float2 *>(wtrs_smem + L::kBuf);
    float2 *s_tw = reinterpret_cast<float2 *>(wtrs_smem + L::kTw);
    uint16_t *s_off = reinterpret_cast<uint16_t *>(wtrs_smem + L::kOff);
    float2 *s_lut = reinterpret_cast<float2 *>(wtrs_smem + L::kLut);
    uint16_t *s_enc14 = reinterpret_cast<uint16_t *>(wtrs_smem + L::kEnc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 16 * (NE + 2); e += kWTrsThreads) {                   // conjugated constellation (conj . FFT . conj), null, pilot
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
    for (int e = tid; e < kN; e += kWTrsThreads) {                              // e = 32 j + l <-> bin l + 32 j = e (encode_block, src/transmitter.rs:144-165)
        const int rk = w_rank<GUARD>(e);
        s_off[e] = (uint16_t)(rk >= 0 ? rk : ((GUARD && w_is_pilot(e)) ? D + 1 : D));
        const int r = e >> 5, c = e & 31;
        s_tw[r * kWPitch + c] = __ldg(a.tables->w1024 + ((r * c) & (kN - 1)));
    }
    if (FEC && tid < 256) s_enc14[tid] = (uint16_t)(ham74_encode_nibble(tid & 15) | (ham74_encode_nibble(tid >> 4) << 7));
    __syncthreads();
    const float head_max = a.tables->head_max;
    const float scale = (1.0f / (float)kN) * (1.0f / head_max);

    float2 *buf = s_buf + warp * kWBuf;
    unsigned long long *wr = reinterpret_cast<unsigned long long *>(buf + lane);
    const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(buf + lane * kWPitch);
    const ulonglong2 *tw_row = reinterpret_cast<const ulonglong2 *>(s_tw + lane * kWPitch);
    const unsigned long long *lut = reinterpret_cast<const unsigned long long *>(s_lut) + (lane & 15);
    const uint16_t *off_col = s_off + lane;
    uint8_t *mybits = wtrs_smem + L::kBits + (size_t)warp * kWTrsBitsBuf;
    uint8_t *car = wtrs_smem + L::kCar + (size_t)warp * kWTrsSlots * kWTrsCarBuf;

    const uint32_t U = (uint32_t)a.group_ctas;                                  // spans of 16 symbols per frame (from iq_stride)
    const uint64_t n_units = (uint64_t)a.n_streams * U;
    auto prefetch_unit = [&](uint64_t u) {
        if (u >= n_units) return;
        const uint32_t st2 = n_units >> 32 ? (uint32_t)(u / U) : (uint32_t)u / U, j2 = (uint32_t)(u - (uint64_t)st2 * U);
        const uint32_t B0 = (j2 * kWTrsWarps + (uint32_t)warp) * BPSB, c0 = B0 < 16 ? 0 : B0 - 16;
        const uint32_t pb = FEC ? 4 * (c0 / 7) + 16 * lane : c0 + 32 * lane;
        constexpr uint32_t span = FEC ? (BPSB * 4) / 7 + 32 : BPSB + 32;
        if ((FEC ? 16u : 32u) * lane < span && pb < __ldg(a.payload_len + st2))
            asm volatile("prefetch.global.L1 [%0];" :: "l"(a.payload + (size_t)st2 * a.payload_stride + pb));
    };
    for (uint64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint32_t stream = n_units >> 32 ? (uint32_t)(unit / U) : (uint32_t)unit / U, j = (uint32_t)(unit - (uint64_t)stream * U);
        prefetch_unit(unit + gridDim.x);                                       // the payload bytes of this warp's share of the CTA's next unit
        const uint32_t n = __ldg(a.payload_len + stream);
        const uint64_t coded_len = FEC ? (14ull * n + 7) / 8 : n;
        const uint64_t ncar = (kHeaderBits + 8 * coded_len + BPC - 1) / BPC;    // constellation symbols (src/transmitter.rs:108-140)
        const uint64_t S64 = (ncar + D - 1) / D;                                // OFDM data symbols (src/transmitter.rs:49-54)
        const uint32_t flen = (10u + (uint32_t)S64) * kL;
        const bool fits = (10ull + S64) * kL <= (uint64_t)a.iq_stride;
        float2 *out = a.iq + (size_t)stream * a.iq_stride;
        if (j == 0 && tid == 0 && a.frame_len) a.frame_len[stream] = flen;
        {                                                                      // this unit's share of the frame head (12 800 samples over the U units), already divided by the bet
            const float rh = fits ? 1.0f / head_max : 0.0f;
            const uint32_t per = ((uint32_t)kHeadW + U - 1) / U, h_lo = j * per;
            uint32_t h_hi = h_lo + per < (uint32_t)kHeadW ? h_lo + per : (uint32_t)kHeadW;
            if (h_hi > a.iq_stride) h_hi = a.iq_stride;
            for (uint32_t i0 = h_lo + tid; i0 < h_hi; i0 += 4 * kWTrsThreads) {  // four table loads in flight per thread
                float2 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + u * kWTrsThreads;
                    v[u] = make_float2(0.0f, 0.0f);
                    if (i < h_hi) v[u] = __ldg(a.tables->head + i);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + u * kWTrsThreads;
                    if (i < h_hi) out[i] = make_float2(v[u].x * rh, v[u].y * rh);
                }
            }
        }
        {                                                                      // zero fill past the frame inside this unit's share of the row
            const uint64_t u_lo = (uint64_t)(10 + (uint64_t)j * kWTrsWarps) * kL;
            const uint64_t u_hi = j + 1 == U ? (uint64_t)a.iq_stride : (uint64_t)(10 + (uint64_t)(j + 1) * kWTrsWarps) * kL;
            const uint64_t f_end = fits ? (uint64_t)flen : (uint64_t)kHeadW;
            uint64_t z = f_end > u_lo ? f_end : u_lo;
            const uint64_t hi = u_hi < (uint64_t)a.iq_stride ? u_hi : (uint64_t)a.iq_stride;
            for (z += tid; z < hi; z += kWTrsThreads) out[z] = make_float2(0.0f, 0.0f);
        }
        const long s = (long)j * kWTrsWarps + warp;                            // this warp's symbol
        if (!fits || s >= (long)S64) continue;
        {
            const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
            const bool pay_aligned = (reinterpret_cast<uintptr_t>(pay) & 3) == 0;
            __syncwarp();
            const uint32_t bit0 = wtrs_build_bits<BPSB, FEC>(mybits, pay, pay_aligned, n, coded_len, (uint32_t)s * BPSB, s_enc14, lane);
            __syncwarp();
            wtrs_unpack_carriers<BPC, D>(car, mybits, bit0, (long)ncar - s * D, lane);
            __syncwarp();
        }
        cpx x[32];
#pragma unroll
        for (int jj = 0; jj < 32; jj++) {
            if (!w_row_used<GUARD>(jj)) { x[jj] = c_make(0.0f, 0.0f); continue; }
            x[jj].v = lut[(uint32_t)car[off_col[32 * jj]] * 16u];
        }
        dft32_p(x);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const ulonglong2 ww = tw_row[q];
            cpx w0, w1;
            w0.v = ww.x; w1.v = ww.y;
            x[2 * q] = c_mul(x[2 * q], w0); x[2 * q + 1] = c_mul(x[2 * q + 1], w1);
        }
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) wr[k1 * kWPitch] = x[k1].v;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 16; q++) { const ulonglong2 v = row[q]; x[2 * q].v = v.x; x[2 * q + 1].v = v.y; }
        __syncwarp();
        dft32_p(x);
        float mx = 0.0f;
        unsigned long long *sym = reinterpret_cast<unsigned long long *>(out + (size_t)(10 + s) * kL) + lane;
        const cpx sc = c_make(scale, -scale);                                  // conj and scale in one
#pragma unroll
        for (int t2 = 0; t2 < 32; t2++) {
            float re, im;
            c_split(x[t2], re, im);                                            // the frame's sample is (re, -im) / 1024
            mx = fmaxf(mx, fmaxf(re, -im));
            const unsigned long long v = c_mul2(x[t2], sc).v;                  // time index lane + 32 t2
            sym[kCpW + 32 * t2] = v;
            if (t2 >= 24) sym[32 * (t2 - 24)] = v;                             // cyclic prefix = last 256 samples
        }
        mx *= 1.0f / (float)kN;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0 && mx > head_max) {                                      // the bet is lost for this frame: it will be redone
            atomicMax(a.stream_max + stream, __float_as_int(mx));
            atomicAdd(a.stream_cnt, 1u);                                       // (what the redo pass looks at first)
        }
    }
}

}  // namespace wide
}  // namespace ofdm
