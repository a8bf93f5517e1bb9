// rs_kernels.cuh -- Reed-Solomon RS(255,223) outer code with the reference's stream framing:
// create_transmission_bytes (src/utils.rs:97-137) and decipher_transmission_bytes (src/utils.rs:152-180), which call the
// `reed-solomon` 0.2.1 crate (GF(2^8), polynomial 0x11d, alpha = 2, g(x) = prod_{i<32} (x - alpha^i), systematic).
//
// One thread owns one 255-byte block; a warp owns a chunk of 32 consecutive blocks of one stream (the grid is flat over
// streams x chunks, so ragged batches fill the CTAs) and keeps their codewords in a
// private shared-memory image (255-byte pitch, i.e. exactly the coded stream), so after the table load warps never wait for
// each other. The coded side of the image moves with 16-byte accesses (the image is placed at the same address modulo 16 as
// the global bytes); the data side (223-byte blocks) is re-blocked through aligned 4-byte global accesses.
// Every thread runs the byte-serial generator LFSR on its block: the 32-byte parity register lives in 8 registers and one
// step is a 32-byte row fetch from a 256-row table (feedback byte x generator), an 8-register byte shift and 8 XORs.
// The table is laid out for conflict-free LDS.128: four copies of every row side by side (128 B per feedback value, copy c
// in banks 8c..8c+7), lane l uses copy l & 3, and lanes 4..7 of each quarter-warp fetch the two 16-byte halves in the
// opposite order -- they keep the parity register rotated by four words so that no data movement is needed -- which puts
// the 8 lanes of a quarter-warp on 8 different 16-byte bank groups whatever their feedback bytes are.
// Decoding re-encodes the 223 received data bytes; parity register XOR received parity = r(x) mod g(x). A zero remainder
// (the common case) finishes the block; otherwise the thread computes the 32 syndromes from the remainder and runs
// Berlekamp-Massey, a Chien search and Forney's formula on its own.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ofdm {

constexpr int kRsN = 255, kRsK = 223, kRsT2 = 32;
constexpr int kRsWarps = 8;
constexpr int kRsThreads = 32 * kRsWarps;                         // blocks of one stream per CTA
constexpr int kRsImgWarp = 32 * kRsN + 16;                        // one warp's image + room for the alignment phase
constexpr size_t kRsSmemBytes = 256 * 128 + kRsWarps * kRsImgWarp + 512 + 256;

struct RsTables {
    uint8_t lfsr[256][32];      // lfsr[f][j] = f * g_{31-j}: what feedback byte f adds to the parity register
    uint8_t exp[512];           // alpha^i, doubled so that log a + log b needs no reduction
    uint8_t log[256];
};

struct RsArgs {
    const uint8_t  *in;
    const uint32_t *in_len;
    uint32_t        in_stride;
    uint8_t        *out;
    uint32_t        out_stride;
    uint32_t       *out_len;
    uint32_t       *n_corrected;    // decode only, per stream (atomic)
    uint32_t       *n_failed;       // decode only, per stream (atomic)
    uint32_t        n_streams;
    uint32_t        chunks_per_stream;   // 32-block chunks per stream the grid provides (from the input stride)
    const RsTables *tables;
};

// ---- warp-level staging ------------------------------------------------------------------------------------------
// coded side, global -> image: `have` bytes exist, the image is zero filled up to `total` (scratch_buf.fill(0), src/utils.rs:167)
__device__ __forceinline__ void rs_warp_copy_in(uint8_t *img, const uint8_t *__restrict__ src, uint32_t have, uint32_t total, int lane)
{
    const uint32_t head = min(have, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(src) & 15)) & 15));
    const uint32_t body = (have - head) >> 4;
    for (uint32_t i = lane; i < head; i += 32) img[i] = src[i];
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    uint4 *d4 = reinterpret_cast<uint4 *>(img + head);
#pragma unroll 8
    for (uint32_t i = lane; i < body; i += 32) d4[i] = __ldg(s4 + i);
    for (uint32_t i = head + (body << 4) + lane; i < have; i += 32) img[i] = src[i];
    for (uint32_t i = have + lane; i < total; i += 32) img[i] = 0;
}

// coded side, image -> global
__device__ __forceinline__ void rs_warp_copy_out(uint8_t *__restrict__ dst, const uint8_t *img, uint32_t n, int lane)
{
    const uint32_t head = min(n, (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    const uint32_t body = (n - head) >> 4;
    for (uint32_t i = lane; i < head; i += 32) dst[i] = img[i];
    const uint4 *s4 = reinterpret_cast<const uint4 *>(img + head);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
#pragma unroll 4
    for (uint32_t i = lane; i < body; i += 32) d4[i] = s4[i];
    for (uint32_t i = head + (body << 4) + lane; i < n; i += 32) dst[i] = img[i];
}

// data side, global -> image: `total` = 223 x blocks contiguous bytes at `src` (the first `have` exist, the rest read as
// zero: scratch_buf.fill(0), src/utils.rs:120) go to the first 223 bytes of each 255-byte image row
__device__ __forceinline__ void rs_warp_scatter_in(uint8_t *img, const uint8_t *__restrict__ src, uint32_t total, uint32_t have, int lane)
{
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src - a);        // aligned words; word k holds bytes 4k - a ..
    const uint32_t nw = (total + a + 3) >> 2;
    for (uint32_t k0 = 0; k0 < nw; k0 += 32 * 8) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t k = k0 + 32 * u + lane;
            v[u] = (4 * k + 4 > a && 4 * k < have + a) ? __ldg(w + k) : 0u;     // only words that overlap [0, have)
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const uint32_t k = k0 + 32 * u + lane;
            if (k >= nw) continue;
            const int i = (int)(4 * k) - (int)a;
            const uint32_t i0 = i < 0 ? 0u : (uint32_t)i;
            uint32_t b = i0 / kRsK, off = i0 - b * kRsK;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int iq = i + q;
                if (iq >= 0 && (uint32_t)iq < total) {
                    img[b * kRsN + off] = (uint32_t)iq < have ? (uint8_t)(v[u] >> (8 * q)) : (uint8_t)0;
                    if (++off == kRsK) { off = 0; b++; }
                }
            }
        }
    }
}

// data side, image -> global: the first 223 bytes of each image row, contiguous at `dst`
__device__ __forceinline__ void rs_warp_gather_out(uint8_t *__restrict__ dst, const uint8_t *img, uint32_t total, int lane)
{
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3);
    uint32_t *w = reinterpret_cast<uint32_t *>(dst - a);
    const uint32_t nw = (total + a + 3) >> 2;
#pragma unroll 2
    for (uint32_t k = lane; k < nw; k += 32) {
        const int i = (int)(4 * k) - (int)a;
        const uint32_t i0 = i < 0 ? 0u : (uint32_t)i;
        uint32_t b = i0 / kRsK, off = i0 - b * kRsK;
        uint32_t word = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int iq = i + q;
            if (iq >= 0 && (uint32_t)iq < total) {
                word |= (uint32_t)img[b * kRsN + off] << (8 * q);
                if (++off == kRsK) { off = 0; b++; }
            }
        }
        if (i >= 0 && (uint32_t)i + 4 <= total) w[k] = word;
        else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int iq = i + q;
                if (iq >= 0 && (uint32_t)iq < total) dst[iq] = (uint8_t)(word >> (8 * q));
            }
        }
    }
}

// The parity register of one thread: r[0..7]; logical word k (bytes 4k..4k+3 of the register, byte 0 = highest-degree
// coefficient) sits in r[k] on lanes 0..3 of a quarter-warp and in r[(k + 4) & 7] on lanes 4..7 (`rot`).
struct RsLfsr {
    uint32_t r[8];
    uint32_t tab;           // shared-memory byte address of this lane's first row half: copy (lane & 3), half `rot`
    uint32_t m3, m7;        // all-ones where r[4] / r[0] is the successor word of r[3] / r[7]
    bool rot;

    __device__ __forceinline__ void init(uint32_t s_lfsr_addr, int lane)
    {
        rot = (lane >> 2) & 1;
        tab = s_lfsr_addr + (lane & 3) * 32 + (rot ? 16 : 0);
        m3 = rot ? 0u : 0xffffffffu;
        m7 = rot ? 0xffffffffu : 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = 0;
    }
    __device__ __forceinline__ void step(uint32_t d)
    {
        const uint32_t f = (d ^ (rot ? r[4] : r[0])) & 255u;
        uint4 a, b;
        const uint32_t addr = tab + f * 128;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr ^ 16u));
        const uint32_t r0 = r[0];
        r[0] = __funnelshift_r(r[0], r[1], 8) ^ a.x;
        r[1] = __funnelshift_r(r[1], r[2], 8) ^ a.y;
        r[2] = __funnelshift_r(r[2], r[3], 8) ^ a.z;
        r[3] = __funnelshift_r(r[3], r[4] & m3, 8) ^ a.w;
        r[4] = __funnelshift_r(r[4], r[5], 8) ^ b.x;
        r[5] = __funnelshift_r(r[5], r[6], 8) ^ b.y;
        r[6] = __funnelshift_r(r[6], r[7], 8) ^ b.z;
        r[7] = __funnelshift_r(r[7], r0 & m7, 8) ^ b.w;
    }
    // clock the 223 message bytes of a shared-memory block through the register
    __device__ __forceinline__ void run223(const uint8_t *msg)
    {
#pragma unroll 1
        for (int j = 0; j + 4 <= kRsK; j += 4) { step(msg[j]); step(msg[j + 1]); step(msg[j + 2]); step(msg[j + 3]); }
        step(msg[kRsK - 3]); step(msg[kRsK - 2]); step(msg[kRsK - 1]);
    }
    __device__ __forceinline__ uint32_t word(int k) const { return rot ? r[(k + 4) & 7] : r[k]; }
};

// table rows -> 4 side-by-side copies; GF tables for the decoder
__device__ __forceinline__ void rs_load_tables(const RsTables *t, uint4 *s_lfsr, uint8_t *s_exp, uint8_t *s_log, int tid, int nt)
{
    const uint4 *g = reinterpret_cast<const uint4 *>(&t->lfsr[0][0]);
    for (int i = tid; i < 512; i += nt) {                             // i = 2 f + half
        const uint4 v = __ldg(g + i);
        uint4 *row = s_lfsr + (i >> 1) * 8 + (i & 1);
        row[0] = v; row[2] = v; row[4] = v; row[6] = v;
    }
    if (s_exp) {
        for (int i = tid; i < 512; i += nt) s_exp[i] = t->exp[i];
        for (int i = tid; i < 256; i += nt) s_log[i] = t->log[i];
    }
}

// ---- encode: data bytes -> [223 data | 32 parity] blocks; the partial (possibly empty) tail block is always emitted -------
__global__ void __launch_bounds__(kRsThreads) rs_encode_kernel(const RsArgs a)
{
    extern __shared__ __align__(128) uint8_t rs_smem[];
    uint4 *s_lfsr = reinterpret_cast<uint4 *>(rs_smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    rs_load_tables(a.tables, s_lfsr, nullptr, nullptr, tid, kRsThreads);
    __syncthreads();
    // warp task = one 32-block chunk of one stream; warps are independent from here on
    const uint32_t task = blockIdx.x * kRsWarps + warp;
    const uint32_t stream = task / a.chunks_per_stream, chunk = task - stream * a.chunks_per_stream;
    if (stream >= a.n_streams) return;
    const uint32_t n = a.in_len[stream];
    const uint32_t nb = n / kRsK + 1;                                      // src/utils.rs:113-134
    const uint32_t need = nb * kRsN;
    if (chunk == 0 && lane == 0) a.out_len[stream] = need;
    const uint32_t b0 = chunk * 32;
    if (b0 >= nb || need > a.out_stride) return;
    const uint32_t cnt = min(32u, nb - b0);
    uint8_t *dst = a.out + (size_t)stream * a.out_stride + (size_t)b0 * kRsN;
    uint8_t *img = rs_smem + 256 * 128 + warp * kRsImgWarp + (reinterpret_cast<uintptr_t>(dst) & 15);
    const uint8_t *src = a.in + (size_t)stream * a.in_stride + (size_t)b0 * kRsK;
    const uint32_t have = n > b0 * kRsK ? min(cnt * kRsK, n - b0 * kRsK) : 0;
    rs_warp_scatter_in(img, src, cnt * kRsK, have, lane);
    __syncwarp();
    if ((uint32_t)lane < cnt) {
        RsLfsr L;
        L.init((uint32_t)__cvta_generic_to_shared(s_lfsr), lane);
        uint8_t *blk = img + lane * kRsN;
        L.run223(blk);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t w = L.word(k);
            blk[kRsK + 4 * k] = (uint8_t)w; blk[kRsK + 4 * k + 1] = (uint8_t)(w >> 8);
            blk[kRsK + 4 * k + 2] = (uint8_t)(w >> 16); blk[kRsK + 4 * k + 3] = (uint8_t)(w >> 24);
        }
    }
    __syncwarp();
    rs_warp_copy_out(dst, img, cnt * kRsN, lane);
}

// ---- decode ------------------------------------------------------------------------------------------------------
struct RsGf {
    const uint8_t *ex, *lg;
    __device__ __forceinline__ uint32_t mul(uint32_t x, uint32_t y) const { return (x && y) ? ex[lg[x] + lg[y]] : 0u; }
    __device__ __forceinline__ uint32_t mul_pow(uint32_t x, uint32_t e) const { return x ? ex[lg[x] + e] : 0u; }   // x * alpha^e, e < 256
    __device__ __forceinline__ uint32_t div(uint32_t x, uint32_t y) const { return x ? ex[lg[x] + 255u - lg[y]] : 0u; }
};

// Corrects the 255-byte word in place from its remainder rem[0..32) (rem[0] = coefficient of x^31).
// Returns the number of corrected symbols or -1 (more than 16 symbol errors: `correct` fails, src/utils.rs:165).
__device__ __noinline__ int rs_correct_from_remainder(uint8_t *word, const uint8_t *rem, const RsGf gf)
{
    uint8_t S[kRsT2];
    for (uint32_t i = 0; i < kRsT2; i++) {                          // S_i = r(alpha^i) = rem(alpha^i)
        uint32_t y = 0;
        for (int j = 0; j < kRsT2; j++) y = gf.mul_pow(y, i) ^ rem[j];
        S[i] = (uint8_t)y;
    }
    // Berlekamp-Massey, Lambda lowest degree first
    uint8_t C[kRsT2 + 2], B[kRsT2 + 2], T[kRsT2 + 2];
    for (int i = 0; i < kRsT2 + 2; i++) { C[i] = 0; B[i] = 0; }
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    uint32_t b = 1;
    for (int r = 0; r < kRsT2; r++) {
        uint32_t d = S[r];
        for (int i = 1; i <= L; i++) d ^= gf.mul(C[i], S[r - i]);
        if (d == 0) { m++; continue; }
        const uint32_t coef = gf.div(d, b);
        const bool grow = 2 * L <= r;
        if (grow) for (int i = 0; i < kRsT2 + 2; i++) T[i] = C[i];
        for (int i = 0; i + m < kRsT2 + 2; i++) C[i + m] ^= (uint8_t)gf.mul(coef, B[i]);
        if (grow) {
            L = r + 1 - L;
            for (int i = 0; i < kRsT2 + 2; i++) B[i] = T[i];
            b = d;
            m = 1;
        } else m++;
    }
    if (2 * L > kRsT2) return -1;

    // Chien search: position p has locator X = alpha^(254-p); Lambda(X^-1) = sum_i C_i alpha^(-i e), e = 254 - p
    uint8_t lt[kRsT2 / 2 + 1];                                      // log of the i-th term, stepped by -i per position
    uint8_t pos[kRsT2 / 2];
    for (int i = 1; i <= L; i++) lt[i] = gf.lg[C[i]];
    int nerr = 0;
    for (int e = 0; e < kRsN; e++) {
        uint32_t y = 1;                                             // C_0
        for (int i = 1; i <= L; i++) {
            if (C[i]) y ^= gf.ex[lt[i]];
            int v = (int)lt[i] - i;
            lt[i] = (uint8_t)(v < 0 ? v + 255 : v);
        }
        if (y == 0) { if (nerr < kRsT2 / 2) pos[nerr] = (uint8_t)(kRsN - 1 - e); nerr++; }
    }
    if (nerr != L) return -1;

    // Forney: Omega = S Lambda mod x^32; e = X Omega(X^-1) / Lambda'(X^-1)
    uint8_t Om[kRsT2];
    for (int i = 0; i < kRsT2; i++) {
        uint32_t v = 0;
        for (int j = 0; j <= i && j <= L; j++) v ^= gf.mul(C[j], S[i - j]);
        Om[i] = (uint8_t)v;
    }
    for (int k = 0; k < nerr; k++) {
        const uint32_t p = pos[k], e = kRsN - 1 - p, einv = (255u - e) % 255u;
        uint32_t om = 0;
        for (int i = kRsT2 - 1; i >= 0; i--) om = gf.mul_pow(om, einv) ^ Om[i];
        uint32_t dl = 0;
        for (int i = 1; i <= L; i += 2) dl ^= gf.mul_pow(C[i], (einv * (uint32_t)(i - 1)) % 255u);
        if (dl == 0) return -1;
        word[p] ^= (uint8_t)gf.mul_pow(gf.div(om, dl), e);
    }
    return nerr;
}

__global__ void __launch_bounds__(kRsThreads) rs_decode_kernel(const RsArgs a)
{
    extern __shared__ __align__(128) uint8_t rs_smem[];
    uint4 *s_lfsr = reinterpret_cast<uint4 *>(rs_smem);
    uint8_t *s_exp = rs_smem + 256 * 128 + kRsWarps * kRsImgWarp, *s_log = s_exp + 512;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    rs_load_tables(a.tables, s_lfsr, s_exp, s_log, tid, kRsThreads);
    __syncthreads();
    const uint32_t task = blockIdx.x * kRsWarps + warp;                    // one 32-block chunk of one stream
    const uint32_t stream = task / a.chunks_per_stream, chunk = task - stream * a.chunks_per_stream;
    if (stream >= a.n_streams) return;
    const uint32_t n = a.in_len[stream];
    const uint32_t nb = n / kRsN + 1;                                      // src/utils.rs:160-176
    const uint32_t need = nb * kRsK;
    if (chunk == 0 && lane == 0) a.out_len[stream] = need;
    const uint32_t b0 = chunk * 32;
    if (b0 >= nb || need > a.out_stride) return;
    const uint32_t cnt = min(32u, nb - b0);
    const uint8_t *src = a.in + (size_t)stream * a.in_stride + (size_t)b0 * kRsN;
    uint8_t *img = rs_smem + 256 * 128 + warp * kRsImgWarp + (reinterpret_cast<uintptr_t>(src) & 15);
    const uint32_t have = n > b0 * kRsN ? min(cnt * kRsN, n - b0 * kRsN) : 0;
    rs_warp_copy_in(img, src, have, cnt * kRsN, lane);
    __syncwarp();
    if ((uint32_t)lane < cnt) {
        uint8_t *word = img + lane * kRsN;
        RsLfsr L;
        L.init((uint32_t)__cvta_generic_to_shared(s_lfsr), lane);
        L.run223(word);
        uint32_t any = 0;
        uint8_t rem[kRsT2];
#pragma unroll
        for (int i = 0; i < kRsT2; i++) {
            rem[i] = (uint8_t)((L.word(i >> 2) >> (8 * (i & 3))) ^ word[kRsK + i]);
            any |= rem[i];
        }
        if (any) {
            const int r = rs_correct_from_remainder(word, rem, RsGf{ s_exp, s_log });
            if (r < 0) atomicAdd(a.n_failed + stream, 1u);
            else atomicAdd(a.n_corrected + stream, (uint32_t)r);
        }
    }
    __syncwarp();
    rs_warp_gather_out(a.out + (size_t)stream * a.out_stride + (size_t)b0 * kRsK, img, cnt * kRsK, lane);
}

}  // namespace ofdm
