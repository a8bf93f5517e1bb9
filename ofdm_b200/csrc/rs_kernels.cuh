// rs_kernels.cuh -- Reed-Solomon RS(255,223) outer code with the reference's stream framing:
// create_transmission_bytes (src/utils.rs:97-137) and decipher_transmission_bytes (src/utils.rs:152-180), which call the
// `reed-solomon` 0.2.1 crate (GF(2^8), polynomial 0x11d, alpha = 2, g(x) = prod_{i<32} (x - alpha^i), systematic).
//
// One thread owns one 255-byte block and streams it through registers: aligned 32-bit words of its own bytes are fetched
// straight from global memory (the 8 loads a sector serves hit L1), re-aligned with a funnel shift, clocked through the
// generator LFSR four bytes at a time and written back -- re-aligned again for the output block -- as aligned words; only
// the ragged first / last bytes of a block are stored byte-wise. No shared-memory image, so the CTA needs the tables only
// and the SM runs as many blocks as registers allow. The grid is flat over streams x blocks.
// LFSR: the 32-byte parity register lives in 8 registers; one step is a 32-byte row fetch from a 256-row table (feedback
// byte x generator), an 8-register byte shift and 8 XORs. The table is laid out for conflict-free LDS.128: four copies of
// every row side by side (128 B per feedback value, copy c in banks 8c..8c+7), lane l uses copy l & 3, and lanes 4..7 of each
// quarter-warp fetch the two 16-byte halves in the opposite order -- they keep the parity register rotated by four words
// so that no data movement is needed -- which puts the 8 lanes of a quarter-warp on 8 different 16-byte bank groups
// whatever their feedback bytes are.
// Decoding re-encodes the 223 received data bytes; parity register XOR received parity = r(x) mod g(x). A zero remainder
// (the common case) finishes the block; otherwise the thread computes the 32 syndromes from the remainder, runs
// Berlekamp-Massey, a Chien search and Forney's formula on its own and patches the data bytes it has just written.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ofdm {

constexpr int kRsN = 255, kRsK = 223, kRsT2 = 32;
constexpr int kRsThreads = 256;
constexpr size_t kRsSmemBytes = 256 * 128 + 512 + 256;

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
    uint32_t        blocks_per_stream;   // blocks per stream the grid provides (from the input stride)
    const RsTables *tables;
};

// ---- one block's bytes as a stream of 32-bit words ---------------------------------------------------------------------
// word j = bytes 4j .. 4j+3 of the block that starts at `p` (any alignment); bytes at or beyond `have` read as zero
// (scratch_buf.fill(0), src/utils.rs:120,167). The bytes come in as aligned 16-byte chunks -- one L2 access per 16 bytes
// and lane, fetched one chunk ahead of their use -- and are re-aligned in registers (word rotation by selects, byte
// shift by funnel shifts). Only chunks that hold at least one existing byte are touched.
struct RsReader {
    const uint4 *c;         // aligned chunk that holds byte 0
    uint32_t off;           // p & 15
    uint32_t have;          // bytes that exist
    uint4 a, b, n;          // chunks q, q + 1 (in use) and q + 2 (in flight), q = the group fetched next
    __device__ __forceinline__ uint4 fetch(uint32_t q) const               // chunk q, or 0 when none of its bytes exists
    {
        return (16 * q < have + off) ? __ldg(c + q) : make_uint4(0u, 0u, 0u, 0u);
    }
    __device__ __forceinline__ void init(const uint8_t *p, uint32_t have_)
    {
        off = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15);
        c = reinterpret_cast<const uint4 *>(p - off);
        have = have_;
        a = fetch(0); b = fetch(1); n = fetch(2);
    }
    // words 4g .. 4g+3 of the block; call with g = 0, 1, 2, ...
    __device__ __forceinline__ void next4(uint32_t g, uint32_t (&d)[4])
    {
        uint32_t t[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
        const bool r1 = (off & 4) != 0, r2 = (off & 8) != 0;
        uint32_t u[7];
#pragma unroll
        for (int i = 0; i < 7; i++) u[i] = r1 ? t[i + 1] : t[i];            // rotate by one word
        uint32_t v[5];
#pragma unroll
        for (int i = 0; i < 5; i++) v[i] = r2 ? u[i + 2] : u[i];            // rotate by two words
        const uint32_t sh = 8 * (off & 3);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t x = __funnelshift_r(v[i], v[i + 1], sh);
            const uint32_t j = 4 * g + i;
            if (4 * j + 4 > have) x = 4 * j >= have ? 0u : x & (0xffffffffu >> (8 * (4 * j + 4 - have)));
            d[i] = x;
        }
        a = b; b = n; n = fetch(g + 3);
    }
};

// the block at `p` (any alignment) written as a stream of 32-bit words: the aligned words are gathered four at a time and
// leave as one 16-byte store when they complete an aligned chunk (a lane's scattered 4-byte stores would cost four times
// the L1/L2 transactions); the ragged ends of a block go out as single words and bytes
struct RsWriter {
    uint8_t *p;
    uint32_t a;             // p & 3
    uint32_t prev;          // word j - 1 of the block
    uint32_t *A0;           // aligned word that holds byte 0 of the block
    uint32_t c0;            // position of A0 inside its 16-byte chunk (0..3)
    uint32_t q0, q1, q2, q3, filled;    // the last aligned words, oldest first
    __device__ __forceinline__ void init(uint8_t *p_)
    {
        p = p_; a = (uint32_t)(reinterpret_cast<uintptr_t>(p_) & 3); prev = 0;
        A0 = reinterpret_cast<uint32_t *>(p_ - a);
        c0 = (uint32_t)((reinterpret_cast<uintptr_t>(A0) >> 2) & 3);
        q0 = q1 = q2 = q3 = 0; filled = 0;
    }
    __device__ __forceinline__ void drain(uint32_t k)                      // the `filled` words that end at aligned word k, one by one
    {
        if (filled >= 3) A0[k - 2] = q1;
        if (filled >= 2) A0[k - 1] = q2;
        if (filled >= 1) A0[k] = q3;
        filled = 0;
    }
    __device__ __forceinline__ void emit(uint32_t k, uint32_t v)           // aligned word k (address A0 + k), all four bytes ours
    {
        q0 = q1; q1 = q2; q2 = q3; q3 = v; filled++;
        if (((c0 + k) & 3) == 3) {                                         // last word of its 16-byte chunk
            if (filled >= 4) { *reinterpret_cast<uint4 *>(A0 + k - 3) = make_uint4(q0, q1, q2, q3); filled = 0; }
            else drain(k);
        }
    }
    // word j (bytes 4j .. 4j+3 of the block), all four bytes valid; call with j = 0, 1, 2, ...
    __device__ __forceinline__ void put(uint32_t j, uint32_t v)
    {
        if (a == 0) emit(j, v);
        else {
            if (j == 0) { for (uint32_t b = 0; b < 4 - a; b++) p[b] = (uint8_t)(v >> (8 * b)); }
            else emit(j, __funnelshift_r(prev, v, 8 * (4 - a)));
            prev = v;
        }
    }
    // the last word: `nv` (1..4) valid low bytes, then the bytes of the previous word still pending
    __device__ __forceinline__ void put_last(uint32_t j, uint32_t v, uint32_t nv)
    {
        drain(a == 0 ? j - 1 : (j > 0 ? j - 1 : 0));                       // aligned words gathered so far end at word j - 1
        if (a != 0 && j > 0) for (uint32_t b = 0; b < a; b++) p[4 * j - a + b] = (uint8_t)(prev >> (8 * (4 - a + b)));
        for (uint32_t b = 0; b < nv; b++) p[4 * j + b] = (uint8_t)(v >> (8 * b));
    }
};

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
    __device__ __forceinline__ void step4(uint32_t v) { step(v); step(v >> 8); step(v >> 16); step(v >> 24); }
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
template <int = 0>
__global__ void __launch_bounds__(kRsThreads) rs_encode_kernel(const RsArgs a)
{
    extern __shared__ __align__(128) uint8_t rs_smem[];
    uint4 *s_lfsr = reinterpret_cast<uint4 *>(rs_smem);
    const int tid = threadIdx.x, lane = tid & 31;
    rs_load_tables(a.tables, s_lfsr, nullptr, nullptr, tid, kRsThreads);
    __syncthreads();
    const uint64_t task = (uint64_t)blockIdx.x * kRsThreads + tid;         // one block of one stream
    const uint32_t stream = (uint32_t)(task / a.blocks_per_stream), b = (uint32_t)(task - (uint64_t)stream * a.blocks_per_stream);
    if (stream >= a.n_streams) return;
    const uint32_t n = a.in_len[stream];
    const uint32_t nb = n / kRsK + 1;                                      // src/utils.rs:113-134
    const uint32_t need = nb * kRsN;
    if (b == 0) a.out_len[stream] = need;
    if (b >= nb || need > a.out_stride) return;

    RsReader rd;
    rd.init(a.in + (size_t)stream * a.in_stride + (size_t)b * kRsK, n > b * kRsK ? min((uint32_t)kRsK, n - b * kRsK) : 0u);
    RsWriter wr;
    wr.init(a.out + (size_t)stream * a.out_stride + (size_t)b * kRsN);
    RsLfsr L;
    L.init((uint32_t)__cvta_generic_to_shared(s_lfsr), lane);
    uint32_t dw[4];
#pragma unroll 1
    for (uint32_t g = 0; g < 13; g++) {                                    // data words 0..51
        rd.next4(g, dw);
#pragma unroll
        for (int i = 0; i < 4; i++) { L.step4(dw[i]); wr.put(4 * g + i, dw[i]); }
    }
    rd.next4(13, dw);                                                      // words 52..54 and bytes 220..222
#pragma unroll
    for (int i = 0; i < 3; i++) { L.step4(dw[i]); wr.put(52 + i, dw[i]); }
    const uint32_t d = dw[3];
    L.step(d); L.step(d >> 8); L.step(d >> 16);
    // codeword words 55..63: [d220 d221 d222 p0] [p1..p4] ... [p29 p30 p31 -]
    uint32_t pw = L.word(0);
    wr.put(55, (d & 0x00ffffffu) | (pw << 24));
#pragma unroll
    for (int k = 1; k < 8; k++) {
        const uint32_t nx = L.word(k);
        wr.put(55 + k, __funnelshift_r(pw, nx, 8));
        pw = nx;
    }
    wr.put_last(63, pw >> 8, 3);
}

// ---- decode ------------------------------------------------------------------------------------------------------
struct RsGf {
    const uint8_t *ex, *lg;
    __device__ __forceinline__ uint32_t mul(uint32_t x, uint32_t y) const { return (x && y) ? ex[lg[x] + lg[y]] : 0u; }
    __device__ __forceinline__ uint32_t mul_pow(uint32_t x, uint32_t e) const { return x ? ex[lg[x] + e] : 0u; }   // x * alpha^e, e < 256
    __device__ __forceinline__ uint32_t div(uint32_t x, uint32_t y) const { return x ? ex[lg[x] + 255u - lg[y]] : 0u; }
};

// Corrects the 223 data bytes at `data` (already written, uncorrected) from the remainder rem[0..32) of the 255-byte word
// (rem[0] = coefficient of x^31); errors in the parity bytes are counted but need no patch.
// Returns the number of corrected symbols or -1 (more than 16 symbol errors: `correct` fails, src/utils.rs:165).
static __device__ __noinline__ int rs_correct_from_remainder(uint8_t *data, const uint8_t *rem, const RsGf gf)
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
        if (p < (uint32_t)kRsK) data[p] ^= (uint8_t)gf.mul_pow(gf.div(om, dl), e);
    }
    return nerr;
}

// The same correction with the 32 lanes of a warp working on ONE block (all lanes call it together; `w` holds the block's
// remainder, the same in every lane). One thread per block is the right shape while every lane has a block to repair (the
// SIMT lanes are the parallelism); at an operating point where most blocks are clean, a warp would instead wait for its one
// or two erroneous lanes to run ~45 000 instructions each on their own. Here lane i holds syndrome S_i, coefficient i of
// the Berlekamp-Massey polynomials and Omega_i, takes 8 of the 255 Chien positions, and the sums over i are warp
// reductions: ~2 000 warp instructions per block. Same bounded-distance semantics and counts as rs_correct_from_remainder.
static __device__ __noinline__ int rs_correct_warp(uint8_t *data, const uint32_t (&w)[8], const RsGf gf, const int lane)
{
    constexpr unsigned kFull = 0xffffffffu;
    uint32_t S = 0;                                                 // S_lane = rem(alpha^lane), Horner over the 32 remainder bytes
#pragma unroll
    for (int j = 0; j < kRsT2; j++) S = gf.mul_pow(S, (uint32_t)lane) ^ ((w[j >> 2] >> (8 * (j & 3))) & 255u);
    // Berlekamp-Massey: C_lane, B_lane (coefficients beyond 31 only ever feed a locator of degree > 16, which fails anyway)
    uint32_t C = lane == 0 ? 1u : 0u, B = C, b = 1;
    int L = 0, m = 1;
    for (int r = 0; r < kRsT2; r++) {
        const uint32_t s = __shfl_sync(kFull, S, (r - lane) & 31);
        uint32_t d = (lane <= r && lane <= L) ? gf.mul(C, s) : 0u;
#pragma unroll
        for (int k = 16; k >= 1; k >>= 1) d ^= __shfl_xor_sync(kFull, d, k);
        if (d == 0) { m++; continue; }
        const uint32_t coef = gf.div(d, b);
        const uint32_t Bs = __shfl_sync(kFull, B, (lane - m) & 31);
        const uint32_t Cn = C ^ (lane >= m ? gf.mul(coef, Bs) : 0u);
        if (2 * L <= r) { B = C; L = r + 1 - L; b = d; m = 1; } else m++;
        C = Cn;
    }
    if (2 * L > kRsT2) return -1;
    // Chien search: this lane evaluates Lambda at the 8 positions e = lane + 32 k (locator X = alpha^e, byte 254 - e)
    uint32_t y[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { y[k] = 1; acc[k] = 0; }          // C_0 = 1; acc = i e mod 255
    for (int i = 1; i <= L; i++) {
        const uint32_t ci = __shfl_sync(kFull, C, i);
        const int lgc = ci ? (int)gf.lg[ci] : 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            acc[k] += (uint32_t)(lane + 32 * k);
            if (acc[k] >= 255u) acc[k] -= 255u;
            int idx = lgc - (int)acc[k];
            if (idx < 0) idx += 255;
            if (ci) y[k] ^= gf.ex[idx];
        }
    }
    int nerr = 0;
    unsigned roots[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        roots[k] = __ballot_sync(kFull, y[k] == 0 && lane + 32 * k < kRsN);
        nerr += __popc(roots[k]);
    }
    if (nerr != L) return -1;
    // Forney: Omega = S Lambda mod x^32 (lane i: Omega_i); e = X Omega(X^-1) / Lambda'(X^-1), one root at a time, sums by reduction
    uint32_t Om = 0;
    for (int j = 0; j <= L; j++) {
        const uint32_t cj = __shfl_sync(kFull, C, j), s = __shfl_sync(kFull, S, (lane - j) & 31);
        if (j <= lane) Om ^= gf.mul(cj, s);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned mask = roots[k];
        while (mask) {
            const uint32_t e = (uint32_t)(__ffs(mask) - 1 + 32 * k), einv = (255u - e) % 255u;
            mask &= mask - 1;
            uint32_t t = gf.mul_pow(Om, (einv * (uint32_t)lane) % 255u);
            if ((lane & 1) && lane <= L) t |= gf.mul_pow(C, (einv * (uint32_t)(lane - 1)) % 255u) << 8;
#pragma unroll
            for (int q = 16; q >= 1; q >>= 1) t ^= __shfl_xor_sync(kFull, t, q);
            const uint32_t om = t & 255u, dl = t >> 8;
            if (dl == 0) return -1;
            const uint32_t p = (uint32_t)(kRsN - 1) - e;
            if (p < (uint32_t)kRsK && lane == 0) data[p] ^= (uint8_t)gf.mul_pow(gf.div(om, dl), e);
        }
    }
    return nerr;
}

#ifndef RS_COOP_MAX_LANES
#define RS_COOP_MAX_LANES 16
#endif
constexpr int kRsCoopMaxLanes = RS_COOP_MAX_LANES;     // erroneous lanes per warp up to which the warp repairs them one by one together

template <int = 0>
__global__ void __launch_bounds__(kRsThreads) rs_decode_kernel(const RsArgs a)
{
    extern __shared__ __align__(128) uint8_t rs_smem[];
    uint4 *s_lfsr = reinterpret_cast<uint4 *>(rs_smem);
    uint8_t *s_exp = rs_smem + 256 * 128, *s_log = s_exp + 512;
    const int tid = threadIdx.x, lane = tid & 31;
    rs_load_tables(a.tables, s_lfsr, s_exp, s_log, tid, kRsThreads);
    __syncthreads();
    const uint64_t task = (uint64_t)blockIdx.x * kRsThreads + tid;         // one block of one stream
    const uint32_t stream = (uint32_t)(task / a.blocks_per_stream), b = (uint32_t)(task - (uint64_t)stream * a.blocks_per_stream);
    // (no early return: all 32 lanes of a warp stay for the cooperative repair below)
    bool live = stream < a.n_streams;
    uint32_t n = 0, nb = 0;
    if (live) {
        n = a.in_len[stream];
        nb = n / kRsN + 1;                                                 // src/utils.rs:160-176
        const uint32_t need = nb * kRsK;
        if (b == 0) a.out_len[stream] = need;
        live = b < nb && need <= a.out_stride;
    }
    uint8_t *data = nullptr;
    uint32_t any = 0, rw[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    if (live) {
        RsReader rd;
        rd.init(a.in + (size_t)stream * a.in_stride + (size_t)b * kRsN, n > b * kRsN ? min((uint32_t)kRsN, n - b * kRsN) : 0u);
        data = a.out + (size_t)stream * a.out_stride + (size_t)b * kRsK;
        RsWriter wr;
        wr.init(data);
        RsLfsr L;
        L.init((uint32_t)__cvta_generic_to_shared(s_lfsr), lane);
        uint32_t dw[4];
#pragma unroll 1
        for (uint32_t g = 0; g < 13; g++) {                                // data words 0..51
            rd.next4(g, dw);
#pragma unroll
            for (int i = 0; i < 4; i++) { L.step4(dw[i]); wr.put(4 * g + i, dw[i]); }
        }
        rd.next4(13, dw);                                                  // words 52..54, bytes 220..222 | received parity byte 0
#pragma unroll
        for (int i = 0; i < 3; i++) { L.step4(dw[i]); wr.put(52 + i, dw[i]); }
        uint32_t d = dw[3];
        L.step(d); L.step(d >> 8); L.step(d >> 16);
        wr.put_last(55, d, 3);
        // remainder = computed parity ^ received parity (bytes 223..254 = byte 3 of word 55 and words 56..63)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            rd.next4(14 + h, dw);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                rw[4 * h + i] = L.word(4 * h + i) ^ __funnelshift_r(d, dw[i], 24);
                any |= rw[4 * h + i];
                d = dw[i];
            }
        }
    }
    __syncwarp();
    unsigned pend = __ballot_sync(0xffffffffu, any != 0);
    if (pend == 0) return;
    const RsGf gf{ s_exp, s_log };
    if (__popc(pend) > kRsCoopMaxLanes) {
        // most lanes have a block to repair: each on its own
        if (any) {
            uint8_t rem[kRsT2];
#pragma unroll
            for (int i = 0; i < kRsT2; i++) rem[i] = (uint8_t)(rw[i >> 2] >> (8 * (i & 3)));
            const int r = rs_correct_from_remainder(data, rem, gf);
            if (r < 0) atomicAdd(a.n_failed + stream, 1u);
            else atomicAdd(a.n_corrected + stream, (uint32_t)r);
        }
        return;
    }
    // few lanes: the warp repairs their blocks one after the other, all lanes on one block (the data bytes a lane has just
    // written are patched by lane 0: the warp barrier above orders them)
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = __shfl_sync(0xffffffffu, rw[i], src);
        const unsigned long long dp = __shfl_sync(0xffffffffu, (unsigned long long)reinterpret_cast<uintptr_t>(data), src);
        const int r = rs_correct_warp(reinterpret_cast<uint8_t *>((uintptr_t)dp), w, gf, lane);
        if (lane == src) {
            if (r < 0) atomicAdd(a.n_failed + stream, 1u);
            else atomicAdd(a.n_corrected + stream, (uint32_t)r);
        }
    }
}

}  // namespace ofdm
