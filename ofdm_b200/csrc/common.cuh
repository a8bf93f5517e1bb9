// common.cuh -- device helpers shared by the OFDM kernels (sm_100a).
//
// FFT design: one 64-point transform is done by 8 lanes x 8 points. Lane l holds x[l + 8j] (j = 0..7),
// which is exactly what a coalesced 8-byte-per-lane load of 64 consecutive fc32 samples delivers.
//   1. radix-8 DFT over j in registers            -> Y_l[ka]
//   2. twiddle W64^(l*ka) (lane constants)
//   3. 8x8 transpose across the 8 lanes through a padded, conflict-free shared-memory tile
//   4. radix-8 DFT over l in registers            -> lane ka holds X[ka + 8 kb] (kb = 0..7)
// A warp therefore transforms 4 OFDM symbols at a time. No cuFFT anywhere.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ofdm {

constexpr int kNfft = 64;
constexpr int kCp = 16;
constexpr int kSym = 80;           // samples per OFDM symbol
constexpr int kHeadSyms = 10;      // lock + 4 preamble + 5 training rows
constexpr int kHeaderBits = 128;   // bincode u128

// transpose scratch: 8 rows (ka) x 10 float2 (8 used + 2 pad => row stride 20 words: LDS.128 of the 8 lanes of a
// group hit 8 disjoint 4-bank groups), + 8 float2 pad per group (group stride 176 words = 16 mod 32: the two
// groups of a half-warp STS.64 into disjoint bank halves).
constexpr int kTrRow = 10;
constexpr int kTrGroup = 8 * kTrRow + 8;   // 88 float2 per 8-lane group
constexpr int kTrWarp = 4 * kTrGroup;      // 352 float2 per warp

__device__ __forceinline__ void cmul(float &ar, float &ai, float br, float bi)
{
    float r = ar * br - ai * bi;
    float i = ar * bi + ai * br;
    ar = r; ai = i;
}

// forward radix-8 DFT (W8 = exp(-j pi/4)), in place, natural order in and out.
// The inverse transform is obtained by calling it with re/im swapped (ifft(x) = swap(fft(swap(x))) / N).
__device__ __forceinline__ void dft8(float (&re)[8], float (&im)[8])
{
    constexpr float kR = 0.70710678118654752440f;
    // DIF stage 1: e[n] = x[n] + x[n+4]; o[n] = (x[n] - x[n+4]) * W8^n
    float er[4], ei[4], orr[4], oi[4];
#pragma unroll
    for (int n = 0; n < 4; n++) {
        er[n] = re[n] + re[n + 4]; ei[n] = im[n] + im[n + 4];
        orr[n] = re[n] - re[n + 4]; oi[n] = im[n] - im[n + 4];
    }
    {   // W8^1 = (1 - j)/sqrt2 : (a + jb) -> ((a + b) + j(b - a))/sqrt2
        float a = orr[1], b = oi[1];
        orr[1] = (a + b) * kR; oi[1] = (b - a) * kR;
        // W8^2 = -j : (a + jb) -> b - ja
        a = orr[2]; b = oi[2];
        orr[2] = b; oi[2] = -a;
        // W8^3 = (-1 - j)/sqrt2 : (a + jb) -> ((b - a) + j(-a - b))/sqrt2
        a = orr[3]; b = oi[3];
        orr[3] = (b - a) * kR; oi[3] = -(a + b) * kR;
    }
    // two 4-point DFTs: Y0 = (y0+y2)+(y1+y3), Y2 = (y0+y2)-(y1+y3), Y1 = (y0-y2) - j(y1-y3), Y3 = (y0-y2) + j(y1-y3)
    {
        float s0r = er[0] + er[2], s0i = ei[0] + ei[2], d0r = er[0] - er[2], d0i = ei[0] - ei[2];
        float s1r = er[1] + er[3], s1i = ei[1] + ei[3], d1r = er[1] - er[3], d1i = ei[1] - ei[3];
        re[0] = s0r + s1r; im[0] = s0i + s1i;
        re[4] = s0r - s1r; im[4] = s0i - s1i;
        re[2] = d0r + d1i; im[2] = d0i - d1r;
        re[6] = d0r - d1i; im[6] = d0i + d1r;
    }
    {
        float s0r = orr[0] + orr[2], s0i = oi[0] + oi[2], d0r = orr[0] - orr[2], d0i = oi[0] - oi[2];
        float s1r = orr[1] + orr[3], s1i = oi[1] + oi[3], d1r = orr[1] - orr[3], d1i = oi[1] - oi[3];
        re[1] = s0r + s1r; im[1] = s0i + s1i;
        re[5] = s0r - s1r; im[5] = s0i - s1i;
        re[3] = d0r + d1i; im[3] = d0i - d1r;
        re[7] = d0r - d1i; im[7] = d0i + d1r;
    }
}

// ---- packed complex arithmetic on Blackwell's 2-wide fp32 pipe ------------------------------------------------------
// A complex number lives in one 64-bit register pair (lo = re, hi = im), i.e. exactly the fc32 memory layout.
// add/sub/mul/fma.f32x2 map to FADD2 / FMUL2 / FFMA2 (sm_100+); ptxas folds the half swaps, per-half negations and
// scalar broadcasts written below into operand modifiers (.LO_HI, .NP, .F32), so a complex multiply is 2 instructions and a
// complex add 1 -- half the issue slots of scalar code.
struct cpx { unsigned long long v; };
__device__ __forceinline__ cpx c_make(float re, float im) { cpx r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(re), "f"(im)); return r; }
__device__ __forceinline__ void c_split(cpx a, float &re, float &im) { asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(a.v)); }
__device__ __forceinline__ cpx c_from(float2 a) { return c_make(a.x, a.y); }
__device__ __forceinline__ float2 c_to(cpx a) { float2 r; c_split(a, r.x, r.y); return r; }
__device__ __forceinline__ cpx c_add(cpx a, cpx b) { cpx r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ cpx c_sub(cpx a, cpx b) { cpx r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ cpx c_mul2(cpx a, cpx b) { cpx r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ cpx c_fma2(cpx a, cpx b, cpx c) { cpx r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ cpx c_scale(cpx a, float k) { return c_mul2(a, c_make(k, k)); }
__device__ __forceinline__ cpx c_mul_mj(cpx a) { float re, im; c_split(a, re, im); return c_make(im, -re); }    // -j a
__device__ __forceinline__ cpx c_mul_pj(cpx a) { float re, im; c_split(a, re, im); return c_make(-im, re); }    // +j a
// a * b = (ar, ai) * (br, br) + (-ai, ar) * (bi, bi): both multiplicands of b are scalar broadcasts and (-ai, ar) is a
// swapped, half-negated view of a -- ptxas folds all three into FMUL2 / FFMA2 operand modifiers, 2 instructions in total
// (the equivalent (ai, ar) * (-bi, bi) costs an extra negate and a register move whenever b is not a computed scalar)
__device__ __forceinline__ cpx c_mul(cpx a, cpx b)
{
    float ar, ai, br, bi;
    c_split(a, ar, ai); c_split(b, br, bi);
    return c_fma2(c_make(-ai, ar), c_make(bi, bi), c_mul2(a, c_make(br, br)));
}

// forward radix-8 DFT on packed complex values (same flow graph as dft8): 28 packed instructions
__device__ __forceinline__ void dft8_p(cpx (&x)[8])
{
    constexpr float kR = 0.70710678118654752440f;
    cpx e0 = c_add(x[0], x[4]), e1 = c_add(x[1], x[5]), e2 = c_add(x[2], x[6]), e3 = c_add(x[3], x[7]);
    cpx o0 = c_sub(x[0], x[4]), o1 = c_sub(x[1], x[5]), o2 = c_sub(x[2], x[6]), o3 = c_sub(x[3], x[7]);
    o1 = c_scale(c_add(o1, c_mul_mj(o1)), kR);            // W8^1 = (1 - j)/sqrt2
    o2 = c_mul_mj(o2);                                    // W8^2 = -j
    o3 = c_scale(c_sub(c_mul_mj(o3), o3), kR);            // W8^3 = (-1 - j)/sqrt2
    {
        cpx s0 = c_add(e0, e2), d0 = c_sub(e0, e2), s1 = c_add(e1, e3), d1 = c_mul_mj(c_sub(e1, e3));
        x[0] = c_add(s0, s1); x[4] = c_sub(s0, s1); x[2] = c_add(d0, d1); x[6] = c_sub(d0, d1);
    }
    {
        cpx s0 = c_add(o0, o2), d0 = c_sub(o0, o2), s1 = c_add(o1, o3), d1 = c_mul_mj(c_sub(o1, o3));
        x[1] = c_add(s0, s1); x[5] = c_sub(s0, s1); x[3] = c_add(d0, d1); x[7] = c_sub(d0, d1);
    }
}

// W64^k = exp(-2 pi j k / 64), k = 0..63 (f64 values rounded to f32)
static const float2 h_w64[64] = {
    {1.0f, 0.0f}, {0.9951847266721969f, -0.0980171403295606f}, {0.9807852804032304f, -0.19509032201612825f}, {0.9569403357322088f, -0.29028467725446233f},
    {0.9238795325112867f, -0.3826834323650898f}, {0.881921264348355f, -0.47139673682599764f}, {0.8314696123025452f, -0.5555702330196022f}, {0.773010453362737f, -0.6343932841636455f},
    {0.7071067811865476f, -0.7071067811865475f}, {0.6343932841636455f, -0.773010453362737f}, {0.5555702330196023f, -0.8314696123025452f}, {0.4713967368259978f, -0.8819212643483549f},
    {0.38268343236508984f, -0.9238795325112867f}, {0.29028467725446233f, -0.9569403357322089f}, {0.19509032201612833f, -0.9807852804032304f}, {0.09801714032956077f, -0.9951847266721968f},
    {0.0f, -1.0f}, {-0.09801714032956065f, -0.9951847266721969f}, {-0.1950903220161282f, -0.9807852804032304f}, {-0.29028467725446216f, -0.9569403357322089f},
    {-0.3826834323650897f, -0.9238795325112867f}, {-0.4713967368259977f, -0.881921264348355f}, {-0.555570233019602f, -0.8314696123025455f}, {-0.6343932841636454f, -0.7730104533627371f},
    {-0.7071067811865475f, -0.7071067811865476f}, {-0.773010453362737f, -0.6343932841636455f}, {-0.8314696123025453f, -0.5555702330196022f}, {-0.8819212643483549f, -0.47139673682599786f},
    {-0.9238795325112867f, -0.3826834323650899f}, {-0.9569403357322088f, -0.2902846772544624f}, {-0.9807852804032304f, -0.1950903220161286f}, {-0.9951847266721968f, -0.09801714032956083f},
    {-1.0f, 0.0f}, {-0.9951847266721969f, 0.09801714032956059f}, {-0.9807852804032304f, 0.19509032201612836f}, {-0.9569403357322089f, 0.2902846772544621f},
    {-0.9238795325112868f, 0.38268343236508967f}, {-0.881921264348355f, 0.47139673682599764f}, {-0.8314696123025455f, 0.555570233019602f}, {-0.7730104533627371f, 0.6343932841636453f},
    {-0.7071067811865477f, 0.7071067811865475f}, {-0.6343932841636459f, 0.7730104533627367f}, {-0.5555702330196022f, 0.8314696123025452f}, {-0.47139673682599786f, 0.8819212643483549f},
    {-0.38268343236509034f, 0.9238795325112865f}, {-0.29028467725446244f, 0.9569403357322088f}, {-0.19509032201612866f, 0.9807852804032303f}, {-0.09801714032956045f, 0.9951847266721969f},
    {0.0f, 1.0f}, {0.09801714032956009f, 0.9951847266721969f}, {0.1950903220161283f, 0.9807852804032304f}, {0.29028467725446205f, 0.9569403357322089f},
    {0.38268343236509f, 0.9238795325112866f}, {0.4713967368259976f, 0.881921264348355f}, {0.5555702330196018f, 0.8314696123025455f}, {0.6343932841636456f, 0.7730104533627369f},
    {0.7071067811865474f, 0.7071067811865477f}, {0.7730104533627367f, 0.6343932841636459f}, {0.8314696123025452f, 0.5555702330196022f}, {0.8819212643483548f, 0.4713967368259979f},
    {0.9238795325112865f, 0.3826834323650904f}, {0.9569403357322088f, 0.2902846772544625f}, {0.9807852804032303f, 0.19509032201612872f}, {0.9951847266721969f, 0.0980171403295605f},
};

// global-memory copy of the table (filled once per engine): lane-dependent indices on __constant__ memory would be
// serialised by the constant cache, a plain cached global load is not
struct W64Table { float2 w[64]; };

// lane twiddles W64^(l*ka), ka = 0..7
__device__ __forceinline__ void fft64_lane_twiddles(const float2 *__restrict__ w64, int l, float (&twr)[8], float (&twi)[8])
{
#pragma unroll
    for (int ka = 0; ka < 8; ka++) {
        float2 w = __ldg(w64 + ((l * ka) & 63));
        twr[ka] = w.x; twi[ka] = w.y;
    }
}

// 64-point forward FFT over an 8-lane group. In: lane l holds x[l + 8j] at index j.
// Out: lane ka holds X[ka + 8kb] at index kb. `tr` points at this group's kTrGroup-float2 scratch.
// All 32 lanes of the warp must call it together (two __syncwarp inside).
__device__ __forceinline__ void fft64_group(float (&re)[8], float (&im)[8], const float (&twr)[8], const float (&twi)[8],
                                            float2 *tr, int l)
{
    dft8(re, im);
#pragma unroll
    for (int ka = 1; ka < 8; ka++) cmul(re[ka], im[ka], twr[ka], twi[ka]);
    __syncwarp();                                     // previous readers of the scratch are done
#pragma unroll
    for (int ka = 0; ka < 8; ka++) tr[ka * kTrRow + l] = make_float2(re[ka], im[ka]);
    __syncwarp();
    const float4 *row = reinterpret_cast<const float4 *>(tr + l * kTrRow);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float4 v = row[q];
        re[2 * q] = v.x; im[2 * q] = v.y; re[2 * q + 1] = v.z; im[2 * q + 1] = v.w;
    }
    dft8(re, im);
}

// packed-complex version of fft64_group (the hot decode kernel uses this one)
__device__ __forceinline__ void fft64_group_p(cpx (&x)[8], const cpx (&tw)[8], float2 *tr, int l)
{
    dft8_p(x);
#pragma unroll
    for (int ka = 1; ka < 8; ka++) x[ka] = c_mul(x[ka], tw[ka]);
    __syncwarp();                                     // previous readers of the scratch are done
    unsigned long long *t64 = reinterpret_cast<unsigned long long *>(tr);
#pragma unroll
    for (int ka = 0; ka < 8; ka++) t64[ka * kTrRow + l] = x[ka].v;
    __syncwarp();
    const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(tr + l * kTrRow);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        ulonglong2 v = row[q];
        x[2 * q].v = v.x; x[2 * q + 1].v = v.y;
    }
    dft8_p(x);
}

// 1/sqrt(x) for x known to be a normal number: one MUFU.RSQ, without the denormal pre/post-scaling rsqrtf() adds when
// flush-to-zero is off (the results are identical on normal inputs)
__device__ __forceinline__ float rsqrt_normal(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    // the thread is suspended by the hardware until the phase completes or the time hint (ns) expires
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) { }                      // try_wait itself suspends the thread (time hint); an extra
                                                                 // nanosleep back-off measured 0.4 % slower
}
// global -> shared bulk copy through the TMA engine; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// ---- carrier map (src/transmitter.rs:150-161, src/receiver.rs:119-134) ------------------------------------------
__host__ __device__ __forceinline__ bool is_null_bin(int k) { return k >= 59 || k <= 5 || k == 32; }
__host__ __device__ __forceinline__ bool is_pilot_bin(int k) { return k == 6 || k == 25 || k == 39 || k == 58; }
// rank of data bin k among the data bins in ascending order, or -1
template <bool GUARD>
__host__ __device__ __forceinline__ int data_rank(int k)
{
    if (!GUARD) return k;
    if (is_null_bin(k) || is_pilot_bin(k)) return -1;
    return k - 7 - (k > 25) - (k > 32) - (k > 39);
}

// ---- Hamming(7,4), docs/SPEC.md section 3 ------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned ham74_encode_nibble(unsigned d)
{
    unsigned d1 = d & 1, d2 = (d >> 1) & 1, d3 = (d >> 2) & 1, d4 = (d >> 3) & 1;
    unsigned p1 = d1 ^ d2 ^ d4, p2 = d1 ^ d3 ^ d4, p3 = d2 ^ d3 ^ d4;
    return p1 | (p2 << 1) | (d1 << 2) | (p3 << 3) | (d2 << 4) | (d3 << 5) | (d4 << 6);
}
__host__ __device__ __forceinline__ unsigned ham74_decode_word(unsigned c)
{
    unsigned s1 = (c ^ (c >> 2) ^ (c >> 4) ^ (c >> 6)) & 1;
    unsigned s2 = ((c >> 1) ^ (c >> 2) ^ (c >> 5) ^ (c >> 6)) & 1;
    unsigned s3 = ((c >> 3) ^ (c >> 4) ^ (c >> 5) ^ (c >> 6)) & 1;
    unsigned s = s1 | (s2 << 1) | (s3 << 2);
    c ^= (1u << s) >> 1;
    return ((c >> 2) & 1) | ((c >> 3) & 0xEu);
}

// ---- Philox4x32-10 (counter-based RNG for the synthetic channel) -------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__host__ __device__ __forceinline__ float u01_from_u32(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// angle (in units of pi, [-1, 1)) of a 0.64 fixed-point turn count
__device__ __forceinline__ float turns_to_pi_units(uint64_t turns) { return (float)(int32_t)(turns >> 32) * (1.0f / 2147483648.0f); }

}  // namespace ofdm
