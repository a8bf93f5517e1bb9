// rx_kernels.cuh -- receive path kernels (replace `decode`, src/receiver.rs:9-96).
//
//   rx_acquire_kernel : one CTA per stream. Frame sync (reference ramp correlation or sliding Schmidl-Cox on
//                       warp-shuffle prefix sums over smem-staged IQ), CFO estimate, channel estimate from the 5
//                       training symbols, header symbol decode -> StreamState (offset, phase step, 1/h_k, lengths).
//   rx_decode_kernel  : the HBM-bound hot kernel. One CTA = 7 warps x 32 OFDM symbols of one stream; per symbol
//                       (8 lanes): coalesced CP-stripped load -> CFO derotation -> register/smem radix-8x8 FFT ->
//                       one-tap equalise -> pilot phase -> hard demap -> (smem) -> Hamming(7,4) LUT decode -> bytes.
//                       The IQ is read exactly once; nothing but the decoded payload is written.
#pragma once

#include "common.cuh"

namespace ofdm {

enum { ST_OK = 0, ST_TOO_SHORT = 1, ST_NO_SYNC = 2, ST_BAD_HEADER = 3, ST_NEG_OFFSET = 4 };

struct __align__(16) StreamState {
    int32_t  status;
    int32_t  offset;      // src/receiver.rs:21
    uint32_t n_syms;      // data symbols the header asks for (<= symbols received)
    uint32_t out_len;     // decoded payload bytes
    uint64_t fstep;       // -f_delta / 2pi as a 0.64 fixed-point turn count (wrapping)
    float    f_delta;     // src/receiver.rs:39
    uint32_t plen;        // Header.packet_length (coded bytes)
    uint32_t n_syms_rx;   // data symbols present in the capture
    uint32_t pad[3];
    float2   g[64];       // 1 / h_k
    float2   h[64];       // h_k, src/receiver.rs:56
};

struct RxTables {          // device copy of the engine's tables
    float2 lock[80];       // locking_signal (src/transmitter.rs:60-72)
    float2 inv_training[64];  // 1 / training_signals (src/transmitter.rs:88-96)
    float2 head[800];      // un-normalised lock | preamble x4 | (CP + IFFT(training)) x5
    float2 w64[64];        // W64^k twiddles
    float  head_max;       // max positive component of head
};

struct RxArgs {
    const float2   *iq;
    const uint32_t *n_samples;
    uint32_t        iq_stride;
    uint32_t        n_streams;
    const uint64_t *stream_base;    // optional: sample index of every stream's capture inside iq (NULL: stream * iq_stride)
    StreamState    *state;
    const RxTables *tables;
    uint8_t        *out;
    uint32_t        out_stride;
    uint32_t       *out_len;
    int32_t        *status;
    uint32_t        sync_window;
    int32_t         tile_shift;     // symbols by which tile boundaries are shifted so they fall on Hamming byte boundaries
    int32_t         sync_mode, cfo_mode, fec;   // run-time switches of the (cold) acquisition kernel
    int32_t         tiles_per_cta;  // consecutive tiles of one stream handled by one CTA of the decode kernel
    uint32_t        stream0;        // decode kernel: first stream of this launch (gridDim.y <= 65535 streams per launch)
    // diag (optional)
    int32_t  *d_offset;
    float    *d_f_delta;
    float2   *d_h;
    uint32_t *d_nsyms;
    float2   *d_points;
    uint32_t  points_stride;
};

template <int MOD> struct ModTraits;
template <> struct ModTraits<0> { static constexpr int kBpc = 1; };
template <> struct ModTraits<1> { static constexpr int kBpc = 2; };
template <> struct ModTraits<2> { static constexpr int kBpc = 6; };

// per-lane constants of the symbol pipeline (constant for all symbols of one stream)
struct RxLane {
    float twr[8], twi[8];   // FFT twiddles W64^(l*ka)
    float wr[8], wi[8];     // derotation inside a symbol: exp(-j f (l + 8j))
    float gr[8], gi[8];     // equaliser 1/h at bins l + 8kb
};

__device__ __forceinline__ void phasor_from_turns(uint64_t turns, float &c, float &s)
{
    sincospif(turns_to_pi_units(turns), &s, &c);
}

__device__ __forceinline__ void rx_lane_init(RxLane &L, const StreamState *st, const float2 *__restrict__ w64, int l)
{
    fft64_lane_twiddles(w64, l, L.twr, L.twi);
    const uint64_t fstep = st->fstep;
    // exp(-j f (l + 8j)) = exp(-j f l) * exp(-j f 8)^j : two exact phasors + a 7-step recurrence (error ~ 7 ulp)
    float sr, si;
    phasor_from_turns(fstep * (uint64_t)l, L.wr[0], L.wi[0]);
    phasor_from_turns(fstep * 8ull, sr, si);
#pragma unroll
    for (int j = 1; j < 8; j++) {
        L.wr[j] = L.wr[j - 1]; L.wi[j] = L.wi[j - 1];
        cmul(L.wr[j], L.wi[j], sr, si);
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        float2 g = st->g[l + 8 * j];
        L.gr[j] = g.x; L.gi[j] = g.y;
    }
}

// predicated byte store to shared memory (keeps the producing arithmetic branch-free: null / pilot bins are computed
// like data bins and simply not stored)
__device__ __forceinline__ void st_shared_u8_if_nonneg(uint8_t *p, uint32_t v, int cond)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ge.s32 q, %2, 0;\n\t@q st.shared.u8 [%0], %1;\n\t}"
                 :: "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v), "r"(cond) : "memory");
}

// predicated byte store to global memory (keeps the look-ups that produce the byte out of a branch)
__device__ __forceinline__ void st_global_u8_if(uint8_t *p, uint32_t v, bool cond)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u8 [%0], %1;\n\t}" :: "l"(p), "r"(v), "r"((uint32_t)cond) : "memory");
}

// the same with the condition held as one bit of a lane-constant mask (one LOP3 with a predicate result per store)
template <uint32_t BIT>
__device__ __forceinline__ void st_shared_u8_if_bit(uint8_t *p, uint32_t v, uint32_t mask)
{
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 q, t, 0;\n\t@q st.shared.u8 [%0], %1;\n\t}"
                 :: "r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v), "r"(mask), "n"(BIT) : "memory");
}

// hard decision of one data point -> BPC bits (src/receiver.rs:155-178; 64QAM docs/SPEC.md 2)
template <int MOD>
__device__ __forceinline__ uint32_t demap_point(float re, float im)
{
    if (MOD == 0) return re > 0.0f ? 1u : 0u;
    if (MOD == 1) {
        bool lb = re >= 0.0f;
        bool rb = lb ? (im >= 0.0f) : (re < 0.0f && im > 0.0f);
        return (lb ? 1u : 0u) | (rb ? 2u : 0u);
    }
    // floor(3.5 v + 4) lands in the mantissa of 2^23 + i (round-down FMA), clamp to [0, 7]
    float ti = __fmaf_rd(re, 3.5f, 8388612.0f);
    float tq = __fmaf_rd(im, 3.5f, 8388612.0f);
    ti = fminf(fmaxf(ti, 8388608.0f), 8388615.0f);
    tq = fminf(fmaxf(tq, 8388608.0f), 8388615.0f);
    uint32_t x = (__float_as_uint(ti) & 7u) | ((__float_as_uint(tq) & 7u) << 3);
    return x ^ ((x >> 1) & 0x1Bu);      // per-axis Gray: i ^ (i >> 1)
}

// 64QAM demap for the hot kernel: 4 FMA-pipe ops + 1 shift-add + 1 smem table look-up.
//   t = sat((3.5 v + 4) / 8) in [0, 1]. A round-down FMA whose result is a DENORMAL has the integer floor(.) as its bit
//   pattern: bits(fma_rd(tq, 2^-146, 0)) = floor(8 tq) = iq (0..8), and with the addend c = lut + 16 iq taken as the
//   denormal c 2^-149, bits(fma_rd(ti, 2^-146, c)) = lut + 16 iq + ii -- the shared-memory address of the table entry
//   gray(min(ii,7)) | gray(min(iq,7)) << 3 (256-byte table, rows of 16). Denormal FMAs run at full rate (no -ftz).
__device__ __forceinline__ uint8_t qam64_lut_entry(int t)
{
    int ii = t & 15, iq = t >> 4;
    ii = ii > 7 ? 7 : ii; iq = iq > 7 ? 7 : iq;
    return (uint8_t)((ii ^ (ii >> 1)) | ((iq ^ (iq >> 1)) << 3));
}
// four carrier bytes (6 valid bits each) -> 24 contiguous bits
__device__ __forceinline__ uint32_t pack4x6(uint32_t x)
{
    const uint32_t t = (x & 0x003F003Fu) | ((x >> 2) & 0x0FC00FC0u);
    return (t & 0xFFFu) | ((t >> 4) & 0xFFF000u);
}
__device__ __forceinline__ uint32_t demap_qam64_lut(float re, float im, uint32_t lut_saddr, float k = 0.4375f)
{
    const float ti = __saturatef(fmaf(re, k, 0.5f));                // k = 3.5 / 8 x (1 / scale of the point)
    const float tq = __saturatef(fmaf(im, k, 0.5f));
    const uint32_t uq = __float_as_uint(__fmul_rd(tq, 0x1p-146f));
    const uint32_t addr = __float_as_uint(__fmaf_rd(ti, 0x1p-146f, __uint_as_float(lut_saddr + (uq << 4))));
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// CP-stripped samples of one OFDM symbol, lane l gets x[l + 8j]: direct (coalesced 8-byte) global loads.
// x0 = pointer to the first post-offset sample of the stream; n_avail = samples from x0 to the end of the capture.
__device__ __forceinline__ void rx_load_symbol(const float2 *__restrict__ x0, uint32_t n_avail, uint32_t sym, bool valid, int l,
                                               float (&zr)[8], float (&zi)[8])
{
    const uint32_t n0 = (kHeadSyms + sym) * kSym + kCp + l;
    const float2 *p = x0 + n0;
    if (valid && n0 + 56 < n_avail) {                    // whole symbol inside the capture: 8 loads, immediate offsets
#pragma unroll
        for (int j = 0; j < 8; j++) { float2 v = __ldg(p + 8 * j); zr[j] = v.x; zi[j] = v.y; }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float2 v = make_float2(0.0f, 0.0f);
            if (valid && n0 + 8 * j < n_avail) v = __ldg(p + 8 * j);     // zero-padded tail row, src/receiver.rs:206-210
            zr[j] = v.x; zi[j] = v.y;
        }
    }
}

// One OFDM symbol per 8-lane group, samples already in (zr, zi): derotate, FFT, equalise, pilot phase.
// On return (zr, zi)[kb] is the equalised + phase-corrected value of bin l + 8kb.
template <bool GUARD, int PHASE>
__device__ __forceinline__ void rx_symbol(const RxLane &L, float br, float bi, float2 *tr, int l, float (&zr)[8], float (&zi)[8])
{
    fft64_group(zr, zi, L.twr, L.twi, tr, l);                              // src/receiver.rs:99-104
#pragma unroll
    for (int kb = 0; kb < 8; kb++) cmul(zr[kb], zi[kb], L.gr[kb], L.gi[kb]);   // src/receiver.rs:67-70
    float rr = br, ri = bi;                                                // common rotation of the data bins
    if (GUARD) {
        // pilots: bins 6, 25, 39, 58 = (lane, kb) (6,0) (1,3) (7,4) (2,7)   src/receiver.rs:125-128
        float pr = 0.0f, pi = 0.0f;
        if (l == 6) { pr = zr[0]; pi = zi[0]; }
        if (l == 1) { pr = zr[3]; pi = zi[3]; }
        if (l == 7) { pr = zr[4]; pi = zi[4]; }
        if (l == 2) { pr = zr[7]; pi = zi[7]; }
        if (PHASE == 1) {
            // angle of the pilot sum: the per-symbol base phasor cancels, rot = conj(sum)/|sum|
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
                pr += __shfl_xor_sync(0xffffffffu, pr, m);
                pi += __shfl_xor_sync(0xffffffffu, pi, m);
            }
            float inv = rsqrt_normal(fmaxf(pr * pr + pi * pi, 1e-30f));
            rr = pr * inv; ri = -pi * inv;
        } else {
            // reference: mean of the four pilot angles (after the full derotation), src/receiver.rs:126,137
            cmul(pr, pi, br, bi);
            const bool pilot_lane = (l == 6) | (l == 1) | (l == 7) | (l == 2);
            float ang = pilot_lane ? atan2f(pi, pr) : 0.0f;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) ang += __shfl_xor_sync(0xffffffffu, ang, m);
            float s, c;
            sincosf(-0.25f * ang, &s, &c);
            rr = c; ri = s;
            cmul(rr, ri, br, bi);
        }
    }
#pragma unroll
    for (int kb = 0; kb < 8; kb++) cmul(zr[kb], zi[kb], rr, ri);           // src/receiver.rs:140-144
}

// ---- packed-complex (FFMA2 / FADD2) versions used by the hot kernel ---------------------------------------------------
struct RxLaneP {
    cpx tw[8];   // FFT twiddles W64^(l*ka)
    cpx w[8];    // derotation inside a symbol: exp(-j f (l + 8j))
};

__device__ __forceinline__ cpx phasor_from_turns_p(uint64_t turns)
{
    float c, s;
    phasor_from_turns(turns, c, s);
    return c_make(c, s);
}

// `rot` (0 or 1): register j of this lane holds sample chunk (j + rot) & 7, i.e. x[l + 8((j + rot) & 7)]. A cyclic shift of
// the radix-8 input only multiplies output ka by W8^(rot*ka) (shift theorem), which is folded into the lane twiddles:
// W64^(l*ka) * W8^(rot*ka) = W64^((l + 8 rot) ka). Odd 8-lane groups use rot = 1 so that the two groups of a half-warp read
// their staged symbols (640 B apart = same banks) from different banks.
__device__ __forceinline__ void rx_lane_init_p(RxLaneP &L, const StreamState *st, const float2 *__restrict__ w64, int l, int rot)
{
#pragma unroll
    for (int ka = 0; ka < 8; ka++) L.tw[ka] = c_from(__ldg(w64 + (((l + 8 * rot) * ka) & 63)));
    const uint64_t fstep = st->fstep;
    // exp(-j f (l + 8m)) = exp(-j f l) * exp(-j f 8)^m : exact phasors for m = rot and the step, then a recurrence (error ~ 7 ulp)
    const cpx step = phasor_from_turns_p(fstep * 8ull);
    cpx w = phasor_from_turns_p(fstep * (uint64_t)(l + 8 * rot));
#pragma unroll
    for (int j = 0; j < 8; j++) {
        L.w[j] = w;                                        // chunk (j + rot) & 7
        if (j + 1 < 8) w = (rot && j == 6) ? phasor_from_turns_p(fstep * (uint64_t)l) : c_mul(w, step);
    }
}

// One OFDM symbol per 8-lane group, derotated samples in z: FFT, equalise, pilot phase.
// On return z[kb] is the equalised + phase-corrected value of bin l + 8kb.
template <bool GUARD, int PHASE>
__device__ __forceinline__ void rx_symbol_p(const RxLaneP &L, const ulonglong2 *__restrict__ g_row, cpx base, float2 *tr, int l, cpx (&z)[8],
                                            uint32_t pmask /* bit q: this lane holds pilot bin 6, 25, 39, 58 in z[0], z[3], z[4], z[7] */,
                                            float &pscale /* out: multiply the returned points by this to normalise them */)
{
    fft64_group_p(z, L.tw, tr, l);                                         // src/receiver.rs:99-104
#pragma unroll
    for (int q = 0; q < 4; q++) {                                          // src/receiver.rs:67-70; 1/h_k of this lane's bins from smem
        const ulonglong2 gg = g_row[q];
        cpx g0, g1;
        g0.v = gg.x; g1.v = gg.y;
        z[2 * q] = c_mul(z[2 * q], g0); z[2 * q + 1] = c_mul(z[2 * q + 1], g1);
    }
    cpx rot = base;                                                        // common rotation of the data bins
    pscale = 1.0f;
    if (GUARD) {
        // pilots: bins 6, 25, 39, 58 = (lane, kb) (6,0) (1,3) (7,4) (2,7)   src/receiver.rs:125-128
        if (PHASE == 1) {
            // angle of the pilot sum: the per-symbol base phasor cancels, rot = conj(sum)/|sum|
            // branch-free pick of this lane's pilot (the compiler turns `if (l == ..) p = z[..]` into divergent branches)
            float pr = 0.0f, pi = 0.0f, ar, ai;
            c_split(z[0], ar, ai); pr = (pmask & 1u) ? ar : pr; pi = (pmask & 1u) ? ai : pi;
            c_split(z[3], ar, ai); pr = (pmask & 2u) ? ar : pr; pi = (pmask & 2u) ? ai : pi;
            c_split(z[4], ar, ai); pr = (pmask & 4u) ? ar : pr; pi = (pmask & 4u) ? ai : pi;
            c_split(z[7], ar, ai); pr = (pmask & 8u) ? ar : pr; pi = (pmask & 8u) ? ai : pi;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) {
                pr += __shfl_xor_sync(0xffffffffu, pr, m);
                pi += __shfl_xor_sync(0xffffffffu, pi, m);
            }
            // rotate by conj(sum) right away; 1/|sum| (MUFU latency) leaves the dependency chain and is applied by the
            // consumer: as the scale of the demap threshold FMA and of the stored points
            pscale = rsqrt_normal(fmaxf(pr * pr + pi * pi, 1e-30f));
            rot = c_make(pr, -pi);
        } else {
            // reference: mean of the four pilot angles (after the full derotation), src/receiver.rs:126,137
            float qr = 0.0f, qi = 0.0f, ar, ai;
            c_split(z[0], ar, ai); qr = (pmask & 1u) ? ar : qr; qi = (pmask & 1u) ? ai : qi;
            c_split(z[3], ar, ai); qr = (pmask & 2u) ? ar : qr; qi = (pmask & 2u) ? ai : qi;
            c_split(z[4], ar, ai); qr = (pmask & 4u) ? ar : qr; qi = (pmask & 4u) ? ai : qi;
            c_split(z[7], ar, ai); qr = (pmask & 8u) ? ar : qr; qi = (pmask & 8u) ? ai : qi;
            float pr, pi;
            c_split(c_mul(c_make(qr, qi), base), pr, pi);
            const bool pilot_lane = pmask != 0;
            float ang = pilot_lane ? atan2f(pi, pr) : 0.0f;
#pragma unroll
            for (int m = 1; m < 8; m <<= 1) ang += __shfl_xor_sync(0xffffffffu, ang, m);
            float sn, cs;
            sincosf(-0.25f * ang, &sn, &cs);
            rot = c_mul(c_make(cs, sn), base);
        }
    }
#pragma unroll
    for (int kb = 0; kb < 8; kb++) z[kb] = c_mul(z[kb], rot);              // src/receiver.rs:140-144
}

// ------------------------------------------------------------------------------------------------------------------
// hot kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDecWarps = 7;
constexpr int kDecThreads = kDecWarps * 32;      // 256
#ifndef DEC_ITERS
#define DEC_ITERS 8
#endif
#ifndef DEC_WARP_OUT
#define DEC_WARP_OUT 0
#endif
constexpr int kDecIters = DEC_ITERS;             // 4 symbols per warp iteration -> 32 symbols per warp (28 with 7 iterations: a warp's span is then Hamming-aligned)
constexpr bool kDecWarpOut = DEC_WARP_OUT != 0 && (4 * DEC_ITERS) % 7 == 0;   // warps turn their own carrier bytes into payload bytes: no CTA barrier (tiles > 0)
constexpr int kDecBaseTiles = 8;                 // tiles per CTA whose start phasors are tabulated (tiles_per_cta <= 8, kDecBaseTiles * 4 * kDecWarps <= threads)
constexpr int kTileSyms = kDecWarps * 4 * kDecIters;   // 224 OFDM symbols per CTA (multiple of 7: Hamming byte alignment)
constexpr int kStageBytes = 4 * kSym * 8 + 16;       // 4 consecutive OFDM symbols (CPs included) + 1 leading / 1 trailing alignment sample
constexpr int kStageGroup = (kStageBytes + 15) / 16 * 2 / 4 + 1;   // float2 per warp staging slot / 4
template <bool GUARD> constexpr size_t rx_decode_smem_bytes()
{
    return sizeof(float2) * (kDecWarps * 4 * kStageGroup + kDecWarps * kTrWarp) + (kDecWarps * 4 * kDecIters * (GUARD ? 48 : 64) + 64) + 128 + 256 + 8 * 8 + 8 * kTrRow * sizeof(float2) + sizeof(float2) * kDecBaseTiles * kDecWarps * 4;
}

template <int MOD, bool GUARD, bool FEC, int PHASE, bool POINTS>
__global__ void __launch_bounds__(kDecThreads, GUARD ? 4 : 3) rx_decode_kernel(const RxArgs a)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 48 : 64;
    constexpr int BPS = BPC * D;                 // bits per OFDM symbol
    constexpr int NB = FEC ? 14 : 8;             // stream bits per output byte

    // dynamic shared memory (> 48 KB for the 64-carrier layout): staging | transpose scratch | carrier bytes | LUTs | mbarriers
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2 *s_stage = reinterpret_cast<float2 *>(smem_raw);
    float2 *s_tr = s_stage + kDecWarps * 4 * kStageGroup;
    uint8_t *s_car = reinterpret_cast<uint8_t *>(s_tr + kDecWarps * kTrWarp);       // one byte (BPC valid bits) per data carrier
    uint8_t *s_ham = s_car + (kTileSyms * D + 64);
    uint8_t *s_qam = s_ham + 128;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_qam + 256);
    float2 *s_g = reinterpret_cast<float2 *>(s_bar + 8);                            // [lane l][kTrRow]: 1/h at bins l + 8kb, padded rows

    // A CTA owns `tiles_per_cta` consecutive 224-symbol tiles of one stream. Tile k covers symbols
    // [k*224 - tile_shift, (k+1)*224 - tile_shift) n [0, S): every inner boundary falls on a Hamming byte boundary.
    const uint32_t stream = blockIdx.y + a.stream0;
    const StreamState *st = a.state + stream;
    if (st->status != ST_OK) return;
    const int S = (int)st->n_syms;
    const int tile_first = (int)blockIdx.x * a.tiles_per_cta;
    int tile_end = tile_first + a.tiles_per_cta;
    {
        const int n_tiles = (S + a.tile_shift + kTileSyms - 1) / kTileSyms;
        if (tile_end > n_tiles) tile_end = n_tiles;
    }
    if (tile_first >= tile_end) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    if (FEC && tid < 128) s_ham[tid] = (uint8_t)ham74_decode_word(tid);
    if (MOD == 2) s_qam[tid] = qam64_lut_entry(tid);
    if (tid < 64) s_g[(tid & 7) * kTrRow + (tid >> 3)] = st->g[tid];
    const ulonglong2 *g_row = reinterpret_cast<const ulonglong2 *>(s_g + (threadIdx.x & 7) * kTrRow);

    const uint32_t n_samples = a.n_samples[stream];
    const uint32_t offset = (uint32_t)st->offset;
    const float2 *x0 = a.iq + (a.stream_base ? (size_t)a.stream_base[stream] : (size_t)stream * a.iq_stride) + offset;
    const uint32_t n_avail = n_samples - offset;
    const uint64_t fstep = st->fstep;
    float2 *tr = s_tr + warp * kTrWarp + g * kTrGroup;
    const uint32_t qam_biased = (uint32_t)__cvta_generic_to_shared(s_qam);      // shared-memory address of the demap table

    // ---- TMA prefetch: one bulk copy per OFDM symbol (its 64 CP-stripped samples, widened to 16-byte alignment) lands in
    // this warp's staging slot one warp-iteration ahead -- also across tile boundaries, so only the very first iteration of
    // a CTA sees the DRAM latency. Completion is signalled on the warp's mbarrier. Symbols below `s_fast` lie, with their
    // two alignment samples, inside the capture; the (rare) others take the bounds-checked direct-load path.
    const uintptr_t xaddr = reinterpret_cast<uintptr_t>(x0 + (kHeadSyms * kSym + kCp));   // sample 0 of data symbol 0
    const int shift = (int)((xaddr >> 3) & 1);      // symbol starts are 8-byte aligned; the 80-sample stride keeps the parity
    int s_fast = n_avail >= (uint32_t)(kHeadSyms * kSym + kCp + kNfft + 2) ? (int)((n_avail - (kHeadSyms * kSym + kCp + kNfft + 2)) / kSym) + 1 : 0;
    if (s_fast > S) s_fast = S;
    float2 *stage_w = s_stage + warp * (4 * kStageGroup);          // [group][kStageGroup]
    uint64_t *bar = s_bar + warp;
    const unsigned long long *stage_rd = reinterpret_cast<const unsigned long long *>(stage_w + shift + g * kSym + kCp + l + 8 * (g & 1));
    const unsigned long long *stage_rd7 = stage_rd + ((g & 1) ? -8 : 56);                       // chunk (7 + rot) & 7
    auto tile_t0 = [&](int tile) { int t = tile * kTileSyms - a.tile_shift; return t < 0 ? 0 : t; };
    auto tile_t1 = [&](int tile) { int t = (tile + 1) * kTileSyms - a.tile_shift; return t > S ? S : t; };
    // one copy for the 4 symbols sb .. sb+3 of a warp iteration when all of them are below lim; lane 0 only. The shared
    // addresses of the slot and its barrier are kept in registers (opaque) instead of being re-derived from tid per issue.
    uint32_t stage_sa = smem_addr(stage_w), bar_sa = smem_addr(bar);
    asm volatile("" : "+r"(stage_sa), "+r"(bar_sa));
    auto issue = [&](int sb, int lim) {
        if (lane == 0) {
            const bool full = sb + 4 <= lim;
            if (full) {
                const uintptr_t src = (xaddr - kCp * 8 + (uintptr_t)sb * (kSym * 8)) & ~(uintptr_t)15;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(stage_sa), "l"(src), "r"((uint32_t)kStageBytes), "r"(bar_sa) : "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_sa), "r"(full ? (uint32_t)kStageBytes : 0u) : "memory");
        }
    };
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    {
        const int t1 = tile_t1(tile_first);
        issue(tile_t0(tile_first) + warp * (4 * kDecIters), s_fast < t1 ? s_fast : t1);
    }

    const int rot = g & 1;                                          // see rx_lane_init_p
    // derotation phasor at the first symbol of every (tile, warp, group) of this CTA: one exact evaluation per thread here
    // instead of one per thread and tile in the loop (the 8 lanes of a group need the same value)
    unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_g + 8 * kTrRow);
    if (tid < (tile_end - tile_first) * (4 * kDecWarps)) {
        const int bt = tid / (4 * kDecWarps), wg = tid - bt * (4 * kDecWarps);
        const int sf = tile_t0(tile_first + bt) + (wg >> 2) * (4 * kDecIters) + (wg & 3);
        s_base[tid] = phasor_from_turns_p(fstep * (uint64_t)((kHeadSyms + sf) * kSym + kCp)).v;
    }
    RxLaneP L;
    rx_lane_init_p(L, st, a.tables->w64, l, rot);
    // byte offset of each of this lane's 8 bins inside a symbol's carrier row (null / pilot bins are not stored)
    int off[8];
#pragma unroll
    for (int kb = 0; kb < 8; kb++) off[kb] = data_rank<GUARD>(l + 8 * kb);
    int d3 = 24 - (l >= 2), d4 = 31 - (l >= 1);                     // rank(l + 24) - (l - 7), rank(l + 32) - (l - 7)
    // bins l, l + 24, l + 32, l + 56 are data carriers on some lanes only: one mask bit each. The mask and the two
    // lane-dependent offsets are made opaque so that they stay in registers instead of being re-derived from l per iteration.
    uint32_t dmask = (off[0] >= 0 ? 1u : 0u) | (off[3] >= 0 ? 2u : 0u) | (off[4] >= 0 ? 4u : 0u) | (off[7] >= 0 ? 8u : 0u);
    uint32_t pmask = (l == 6 ? 1u : 0u) | (l == 1 ? 2u : 0u) | (l == 7 ? 4u : 0u) | (l == 2 ? 8u : 0u);     // pilot bins 6, 25, 39, 58
    asm volatile("" : "+r"(dmask), "+r"(d3), "+r"(d4), "+r"(pmask));
    const cpx dbase = phasor_from_turns_p(fstep * (uint64_t)(4 * kSym));
    uint8_t *out = a.out + (size_t)stream * a.out_stride;
    uint32_t phase = 0;                                             // mbarrier phase parity
    __syncthreads();                                                // LUTs visible

#pragma unroll 1
    for (int tile = tile_first; tile < tile_end; tile++) {
        const int t0 = tile_t0(tile), t1 = tile_t1(tile);
        const int lim = s_fast < t1 ? s_fast : t1;                  // symbols of this tile served by the TMA path
        const int s_warp = t0 + warp * (4 * kDecIters);             // first symbol of this warp
        const int s_first = s_warp + g;
        // base phasor of this group's first symbol, then a x4-symbol recurrence (7 steps: negligible drift)
        cpx base;
        base.v = s_base[(tile - tile_first) * (4 * kDecWarps) + 4 * warp + g];
        uint8_t *rowp = s_car + (s_first - t0) * D + (GUARD ? l - 7 : l);       // row of this group's symbol, biased by the lane

#pragma unroll 1
        for (int it = 0; it < kDecIters; it++) {
            const int s = s_first + 4 * it;
            cpx z[8];
            mbar_wait(bar, phase);
            phase ^= 1;
            if (s_warp + 4 * it + 4 <= lim) {
#pragma unroll
                for (int j = 0; j < 7; j++) z[j].v = stage_rd[8 * j];          // chunk j + rot
                z[7].v = stage_rd7[0];                                          // chunk (7 + rot) & 7
            } else {
                float zr[8], zi[8];
                rx_load_symbol(x0, n_avail, (uint32_t)s, s < t1, l, zr, zi);
#pragma unroll
                for (int j = 0; j < 8; j++) z[j] = rot ? c_make(zr[(j + 1) & 7], zi[(j + 1) & 7]) : c_make(zr[j], zi[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) z[j] = c_mul(z[j], L.w[j]);            // src/receiver.rs:44-50 (intra-symbol part)
            __syncwarp();                                                      // staging slot consumed by every lane
            if (it + 1 < kDecIters) {
                issue(s_warp + 4 * (it + 1), lim);
            } else if (tile + 1 < tile_end) {                                  // first iteration of the next tile
                const int n1 = tile_t1(tile + 1);
                issue(t1 + warp * (4 * kDecIters), s_fast < n1 ? s_fast : n1);
            }
            float pscale;
            rx_symbol_p<GUARD, PHASE>(L, g_row, base, tr, l, z, pmask, pscale);
            const float dk = 0.4375f * pscale;
            base = c_mul(base, dbase);
            // rows of symbols past t1 exist in s_car but are never read: no `valid` predicate needed on the stores
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                float zr, zi;
                c_split(z[kb], zr, zi);
                uint32_t v = MOD == 2 ? demap_qam64_lut(zr, zi, qam_biased, dk) : demap_point<MOD>(zr, zi);
                // carrier rank of bin l + 8kb is affine in kb except at the pilot / DC crossings (kb = 3, 4): immediate
                // offsets from one row pointer; kb in {1, 2, 5, 6} are data carriers on every lane
                if (!GUARD) rowp[8 * kb] = (uint8_t)v;
                else if (kb == 1 || kb == 2) rowp[8 * kb] = (uint8_t)v;
                else if (kb == 5 || kb == 6) rowp[8 * kb - 3] = (uint8_t)v;
                else if (kb == 3) st_shared_u8_if_bit<2>(rowp + d3, v, dmask);
                else if (kb == 4) st_shared_u8_if_bit<4>(rowp + d4, v, dmask);
                else if (kb == 0) st_shared_u8_if_bit<1>(rowp, v, dmask);
                else st_shared_u8_if_bit<8>(rowp + 53, v, dmask);
                if (POINTS && s < t1 && (!GUARD || off[kb] >= 0)) {
                    size_t p = (size_t)s * D + off[kb];
                    if (p < a.points_stride) a.d_points[(size_t)stream * a.points_stride + p] = make_float2(zr * pscale, zi * pscale);
                }
            }
            rowp += 4 * D;
        }
        if (kDecWarpOut && tile > 0) {
            // ---- a warp's 28 symbols start on a payload-byte boundary (28 x BPS bits = whole bytes, Hamming or not, and the tile
            // does): the warp converts its own carrier bytes, nobody waits for anybody. (Tile 0 starts at stream bit 0, its
            // payload bytes at bit 128: there the boundaries between the warps cut through bytes -- CTA-wide conversion below.)
            __syncwarp();
            const int w_s1 = s_warp + 4 * kDecIters < t1 ? s_warp + 4 * kDecIters : t1;
            if (s_warp < w_s1) {
                const long wbit0 = (long)s_warp * BPS, wbit1 = (long)w_s1 * BPS;
                const long wj0 = (wbit0 - kHeaderBits + NB - 1) / NB;
                long wj1 = (wbit1 - kHeaderBits) / NB;
                if (wj1 > (long)st->out_len) wj1 = (long)st->out_len;
                const int wpbase = (int)(kHeaderBits + wj0 * NB - wbit0);
                const int wnbytes = (int)(wj1 - wj0);
                const uint8_t *wcar = s_car + (s_warp - t0) * D;
                if (MOD == 2) {
                    constexpr int NC = (3 * NB + 5) / 6 + 1;
                    const int c0 = wpbase / 6, sh = wpbase - 6 * c0;
                    const uint8_t *cp = wcar + c0 + (NB / 2) * lane;
                    const uint32_t cp_s = (uint32_t)__cvta_generic_to_shared(cp);
                    uint32_t wa = cp_s & ~3u;
                    const uint32_t sel = 0x3210u + 0x1111u * (cp_s & 3u);
                    static_assert(((NB / 2) * 32) % 4 == 0, "carrier stride per iteration must keep the word alignment");
                    for (int u = 3 * lane; u < wnbytes; u += 3 * 32, cp += (NB / 2) * 32, wa += (NB / 2) * 32) {
                        uint32_t lo, hi = 0;
                        if (NC > 5) {
                            uint32_t w0, w1, w2;
                            asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                                         : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(wa));
                            const uint32_t p0 = pack4x6(__byte_perm(w0, w1, sel)), p1 = pack4x6(__byte_perm(w1, w2, sel));
                            lo = p0 | (p1 << 24);
                            hi = p1 >> 8;
                        } else {
                            lo = cp[0] | (cp[1] << 6) | (cp[2] << 12) | (cp[3] << 18) | (cp[4] << 24);
                        }
                        lo = __funnelshift_r(lo, hi, sh);
                        hi >>= sh;
                        uint32_t w0, w1, w2;
                        if (FEC) {
                            w0 = lo & 0x3FFFu; w1 = (lo >> 14) & 0x3FFFu; w2 = __funnelshift_r(lo, hi, 28) & 0x3FFFu;
                            w0 = s_ham[w0 & 127u] | (s_ham[w0 >> 7] << 4);
                            w1 = s_ham[w1 & 127u] | (s_ham[w1 >> 7] << 4);
                            w2 = s_ham[w2 & 127u] | (s_ham[w2 >> 7] << 4);
                        } else {
                            w0 = lo & 255u; w1 = (lo >> 8) & 255u; w2 = (lo >> 16) & 255u;
                        }
                        uint8_t *o = out + wj0 + u;
                        o[0] = (uint8_t)w0;
                        st_global_u8_if(o + 1, w1, u + 1 < wnbytes);
                        st_global_u8_if(o + 2, w2, u + 2 < wnbytes);
                    }
                } else {
                    for (int u = lane; u < wnbytes; u += 32) {
                        const int p = wpbase + u * NB;
                        const int c = p / BPC, sh = p - c * BPC;
                        constexpr int NC = (NB + BPC - 1) / BPC + (BPC > 1 ? 1 : 0);
                        uint32_t v = 0;
#pragma unroll
                        for (int i = 0; i < NC; i++) v |= (uint32_t)wcar[c + i] << (BPC * i);
                        v >>= sh;
                        uint32_t byte;
                        if (FEC) byte = (uint32_t)s_ham[v & 127u] | ((uint32_t)s_ham[(v >> 7) & 127u] << 4);
                        else byte = v & 255u;
                        out[wj0 + u] = (uint8_t)byte;
                    }
                }
            }
            __syncwarp();                                       // this warp's rows of s_car are rewritten by its next tile
            continue;
        }
        __syncthreads();

        // ---- tile bits -> output bytes (header strip src/receiver.rs:86-93, Hamming docs/SPEC.md 3) --------------
        const long bit0 = (long)t0 * BPS;                       // first stream bit held by this tile
        const long bit1 = (long)t1 * BPS;
        long j0 = bit0 <= kHeaderBits ? 0 : (bit0 - kHeaderBits + NB - 1) / NB;
        long j1 = (bit1 - kHeaderBits) / NB;                    // bytes fully inside the tile
        if (j1 > (long)st->out_len) j1 = (long)st->out_len;
        const int pbase = (int)(kHeaderBits + j0 * NB - bit0);  // tile-local bit position of byte j0
        const int nbytes = (int)(j1 - j0);
        if (MOD == 2) {
            // 3 output bytes per thread: 42 (Hamming) or 24 stream bits = 7 or 4 six-bit carriers (+1 for the bit shift)
            // 3 NB = 42 or 24 bits is a whole number of carriers, so the sub-carrier bit shift is the same for every thread
            constexpr int NC = (3 * NB + 5) / 6 + 1;
            static_assert((3 * NB) % 6 == 0, "3 output bytes must span whole carriers");
            const int c0 = pbase / 6, sh = pbase - 6 * c0;
            const uint8_t *cp = s_car + c0 + (NB / 2) * tid;
            // FEC: the 8 carrier bytes come from 3 aligned words + 2 byte permutes (the per-iteration stride is a multiple
            // of 4, so a thread's byte alignment -- the permute selector -- never changes) instead of 8 byte loads
            const uint32_t cp_s = (uint32_t)__cvta_generic_to_shared(cp);
            uint32_t wa = cp_s & ~3u;
            const uint32_t sel = 0x3210u + 0x1111u * (cp_s & 3u);
            static_assert(((NB / 2) * kDecThreads) % 4 == 0, "carrier stride per iteration must keep the word alignment");
            for (int u = 3 * tid; u < nbytes; u += 3 * kDecThreads, cp += (NB / 2) * kDecThreads, wa += (NB / 2) * kDecThreads) {
                uint32_t lo, hi = 0;
                if (NC > 5) {
                    uint32_t w0, w1, w2;
                    asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                                 : "=r"(w0), "=r"(w1), "=r"(w2) : "r"(wa));
                    const uint32_t p0 = pack4x6(__byte_perm(w0, w1, sel)), p1 = pack4x6(__byte_perm(w1, w2, sel));
                    lo = p0 | (p1 << 24);
                    hi = p1 >> 8;
                } else {
                    lo = cp[0] | (cp[1] << 6) | (cp[2] << 12) | (cp[3] << 18) | (cp[4] << 24);
                }
                lo = __funnelshift_r(lo, hi, sh);
                hi >>= sh;
                uint32_t w0, w1, w2;
                if (FEC) {
                    w0 = lo & 0x3FFFu; w1 = (lo >> 14) & 0x3FFFu; w2 = __funnelshift_r(lo, hi, 28) & 0x3FFFu;
                    w0 = s_ham[w0 & 127u] | (s_ham[w0 >> 7] << 4);
                    w1 = s_ham[w1 & 127u] | (s_ham[w1 >> 7] << 4);
                    w2 = s_ham[w2 & 127u] | (s_ham[w2 >> 7] << 4);
                } else {
                    w0 = lo & 255u; w1 = (lo >> 8) & 255u; w2 = (lo >> 16) & 255u;
                }
                uint8_t *o = out + j0 + u;
                o[0] = (uint8_t)w0;
                st_global_u8_if(o + 1, w1, u + 1 < nbytes);
                st_global_u8_if(o + 2, w2, u + 2 < nbytes);
            }
        } else {
            for (int u = tid; u < nbytes; u += kDecThreads) {
                const int p = pbase + u * NB;
                const int c = p / BPC, sh = p - c * BPC;
                constexpr int NC = (NB + BPC - 1) / BPC + (BPC > 1 ? 1 : 0);
                uint32_t v = 0;
#pragma unroll
                for (int i = 0; i < NC; i++) v |= (uint32_t)s_car[c + i] << (BPC * i);
                v >>= sh;
                uint32_t byte;
                if (FEC) byte = (uint32_t)s_ham[v & 127u] | ((uint32_t)s_ham[(v >> 7) & 127u] << 4);
                else byte = v & 255u;
                out[j0 + u] = (uint8_t)byte;
            }
        }
        if (tile + 1 < tile_end) __syncthreads();               // s_car is rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------------------------
// acquisition kernel
// ------------------------------------------------------------------------------------------------------------------
constexpr int kAcqThreads = 256;
constexpr int kAcq64Threads = 128;               // CTA size of the N=64 acquisition kernel (small per-stream work, many barriers)
constexpr int kAcqChunk = 1024;                  // Schmidl-Cox lags per smem-staged chunk
constexpr int kAcq64Ctas = 8;                    // CTAs per SM the N=64 acquisition kernel is compiled for (64 registers; it is
                                                 // latency / barrier bound, so resident warps are what buys time)

__device__ __forceinline__ float2 ld_sample(const float2 *__restrict__ x, long n, long n_samples)
{
    return (n >= 0 && n < n_samples) ? __ldg(x + n) : make_float2(0.0f, 0.0f);
}

// block-wide arg-max with "first strict maximum" semantics (larger value wins, ties -> smaller index)
template <int NT>
__device__ __forceinline__ void block_argmax(float &val, int &idx, float *s_val, int *s_idx)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, val, m);
        int oi = __shfl_xor_sync(0xffffffffu, idx, m);
        if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { s_val[warp] = val; s_idx[warp] = idx; }
    __syncthreads();
    if (warp == 0) {
        float v = lane < (NT / 32) ? s_val[lane] : -1.0f;
        int i = lane < (NT / 32) ? s_idx[lane] : 0x7fffffff;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, m);
            int oi = __shfl_xor_sync(0xffffffffu, i, m);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        if (lane == 0) { s_val[0] = v; s_idx[0] = i; }
    }
    __syncthreads();
    val = s_val[0]; idx = s_idx[0];
    __syncthreads();
}

// |c[k]|^2 of the ramp correlation c[k] = sum_{n<80} a[n+k] lock[n] (lock is real) -- src/signals/mod.rs:186-217 direct form
__device__ __forceinline__ float ramp_corr_sq(const float2 *__restrict__ x, long k, long n_samples, const float *s_lock)
{
    float cr = 0.0f, ci = 0.0f;
    if (k >= 0 && k + kSym <= n_samples) {
#pragma unroll 8
        for (int n = 0; n < kSym; n++) { float2 v = __ldg(x + k + n); cr = fmaf(v.x, s_lock[n], cr); ci = fmaf(v.y, s_lock[n], ci); }
    } else {
        for (int n = 0; n < kSym; n++) { float2 v = ld_sample(x, k + n, n_samples); cr = fmaf(v.x, s_lock[n], cr); ci = fmaf(v.y, s_lock[n], ci); }
    }
    return cr * cr + ci * ci;
}

// arg-max of the ramp correlation over lags [k_lo, k_hi]
template <int NT>
__device__ __forceinline__ int ramp_argmax(const float2 *__restrict__ x, long n_samples, long k_lo, long k_hi,
                                           const float *s_lock, float *s_val, int *s_idx)
{
    float best = 0.0f;
    int bidx = 0x7fffffff;
    for (long k = k_lo + threadIdx.x; k <= k_hi; k += NT) {
        float v = ramp_corr_sq(x, k, n_samples, s_lock);
        if (v > best) { best = v; bidx = (int)k; }         // per-thread lags ascend: strict > keeps the first
    }
    block_argmax<NT>(best, bidx, s_val, s_idx);
    return best > 0.0f ? bidx : (int)k_lo;
}

// The same arg-max for a short lag range (<= 2 NT - 1 lags) well inside the capture: a thread takes two adjacent lags K, K + 1
// with K on a 16-byte boundary and slides over 16-byte sample pairs -- a quarter of the load requests of ramp_argmax.
// Same products, same summation order (n ascending), same tie rule -> the same lag. Falls back to ramp_argmax otherwise.
template <int NT>
__device__ __forceinline__ int ramp_argmax_pairs(const float2 *__restrict__ x, long n_samples, long k_lo, long k_hi,
                                                 const float *s_lock, float *s_val, int *s_idx)
{
    const long par = (long)((reinterpret_cast<uintptr_t>(x) >> 3) & 1);          // x + k is 16-byte aligned iff (k + par) is even
    const long k_start = k_lo - ((k_lo + par) & 1);
    if ((reinterpret_cast<uintptr_t>(x) & 7) != 0 || k_start < 0 || k_hi - k_start + 1 > 2 * NT || k_start + 2 * NT + kSym + 2 > n_samples)
        return ramp_argmax<NT>(x, n_samples, k_lo, k_hi, s_lock, s_val, s_idx);
    const long K = k_start + 2 * (long)threadIdx.x;
    float best = 0.0f;
    int bidx = 0x7fffffff;
    if (K <= k_hi) {
        const float4 *p = reinterpret_cast<const float4 *>(x + K);
        float ar = 0.0f, ai = 0.0f, br = 0.0f, bi = 0.0f;
        float4 cur = __ldg(p);
#pragma unroll 4
        for (int m = 0; m < kSym / 2; m++) {
            const float4 nxt = __ldg(p + m + 1);
            const float l0 = s_lock[2 * m], l1 = s_lock[2 * m + 1];
            ar = fmaf(cur.x, l0, ar); ai = fmaf(cur.y, l0, ai);                 // lag K:     x[K + 2m], x[K + 2m + 1]
            ar = fmaf(cur.z, l1, ar); ai = fmaf(cur.w, l1, ai);
            br = fmaf(cur.z, l0, br); bi = fmaf(cur.w, l0, bi);                 // lag K + 1: x[K + 2m + 1], x[K + 2m + 2]
            br = fmaf(nxt.x, l1, br); bi = fmaf(nxt.y, l1, bi);
            cur = nxt;
        }
        if (K >= k_lo) { const float v = ar * ar + ai * ai; if (v > best) { best = v; bidx = (int)K; } }
        if (K + 1 <= k_hi) { const float v = br * br + bi * bi; if (v > best) { best = v; bidx = (int)K + 1; } }
    }
    block_argmax<NT>(best, bidx, s_val, s_idx);
    return best > 0.0f ? bidx : (int)k_lo;
}

template <int MOD, bool GUARD, int PHASE>
__global__ void __launch_bounds__(kAcq64Threads, kAcq64Ctas) rx_acquire_kernel(const RxArgs a)
{
    const int SYNC = a.sync_mode, CFO = a.cfo_mode;
    const bool FEC = a.fec != 0;
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 48 : 64;
    constexpr int BPS = BPC * D;
    constexpr int HDR_SYMS = (kHeaderBits + BPS - 1) / BPS;         // 1..3
    constexpr int NW = kAcq64Threads / 32;                          // warps per CTA

    __shared__ float s_lock[kSym];
    __shared__ float s_val[kAcq64Threads / 32];
    __shared__ int s_idx[kAcq64Threads / 32];
    __shared__ int s_d0;
    __shared__ double s_red[2 * (kAcq64Threads / 32)];
    __shared__ __align__(16) float2 s_tr[kTrWarp];
    __shared__ uint8_t s_car[HDR_SYMS * D + 32];
    // Schmidl-Cox staging: prefix sums of q[n] = conj(a[n]) a[n+80] and e[n] = |a[n]|^2
    __shared__ float2 s_q[kAcqChunk + kSym + 1];
    __shared__ float s_e[kAcqChunk + 2 * kSym + 1];

    const uint32_t stream = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    StreamState *st = a.state + stream;
    const long M = (long)a.n_samples[stream];
    const float2 *x = a.iq + (a.stream_base ? (size_t)a.stream_base[stream] : (size_t)stream * a.iq_stride);

    if (tid < kSym) s_lock[tid] = a.tables->lock[tid].x;
    if (tid == 0) s_d0 = 0x7fffffff;
    __syncthreads();

    long W = a.sync_window > 0 ? (long)a.sync_window : M;
    if (W > M) W = M;
    int status = ST_OK;
    long offset = 0;

    if (SYNC == 2) {
        offset = 0;                                    // frame start already known (ofdm_rx_decode_capture)
    } else if (SYNC == 0) {
        offset = (long)ramp_argmax<kAcq64Threads>(x, M, -(kSym - 1), W - 1, s_lock, s_val, s_idx) - 1;     // src/receiver.rs:21: lag - 1
    } else {
        // sliding Schmidl-Cox (docs/SPEC.md 4): P(d) = Q[d+80] - Q[d], R1(d) = E[d+80] - E[d], R2(d) = E[d+160] - E[d+80]
        long d_end = W;
        if (d_end > M - 2 * kSym + 1) d_end = M - 2 * kSym + 1;
        for (long base = 0; base < d_end; base += kAcqChunk) {
            // exclusive prefix sums over this chunk (+ halo), built from per-thread serial runs + a warp-shuffle scan
            constexpr int NQ = kAcqChunk + kSym;          // q needed for n in [base, base + chunk + 80)
            constexpr int NE = kAcqChunk + 2 * kSym;      // e needed for n in [base, base + chunk + 160)
            constexpr int RUN = (NE + kAcq64Threads - 1) / kAcq64Threads;   // consecutive samples per thread
            float qr[RUN], qi[RUN], ee[RUN];
            float tqr = 0.0f, tqi = 0.0f, te = 0.0f;
            // The global-load path (L1 tag stage) is what limits this kernel: a thread's run is fetched with 16-byte loads
            // when the capture allows it, and a[n + 80] is lane + 8's a[n] (RUN = 10 samples per lane) -- only lanes 24..31
            // load it themselves.
            static_assert(RUN * 8 == kSym && RUN % 2 == 0, "lane + 8 holds the sample 80 further on");
            float2 v0s[RUN];
            const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (base + (long)kAcq64Threads * RUN + kSym <= M);   // uniform
            if (vec) {
                const float4 *x4 = reinterpret_cast<const float4 *>(x + base + (long)tid * RUN);
#pragma unroll
                for (int q = 0; q < RUN / 2; q++) {
                    const float4 pq = __ldg(x4 + q);
                    v0s[2 * q] = make_float2(pq.x, pq.y); v0s[2 * q + 1] = make_float2(pq.z, pq.w);
                }
            } else {
#pragma unroll
                for (int r = 0; r < RUN; r++) v0s[r] = ld_sample(x, base + (long)tid * RUN + r, M);
            }
#pragma unroll
            for (int r = 0; r < RUN; r++) {
                long n = base + (long)tid * RUN + r;
                const float2 v0 = v0s[r];
                float2 v1 = make_float2(__shfl_down_sync(0xffffffffu, v0.x, 8), __shfl_down_sync(0xffffffffu, v0.y, 8));
                if (lane >= 24) v1 = ld_sample(x, n + kSym, M);
                // conj(v0) * v1
                float pr = v0.x * v1.x + v0.y * v1.y, pi = v0.x * v1.y - v0.y * v1.x;
                qr[r] = tqr; qi[r] = tqi; ee[r] = te;                   // exclusive within the run
                tqr += pr; tqi += pi; te += v0.x * v0.x + v0.y * v0.y;
            }
            // scan of the per-thread totals: warp shuffle, then across warps through smem
            float sqr = tqr, sqi = tqi, se = te;
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) {
                float a0 = __shfl_up_sync(0xffffffffu, sqr, m), a1 = __shfl_up_sync(0xffffffffu, sqi, m), a2 = __shfl_up_sync(0xffffffffu, se, m);
                if (lane >= m) { sqr += a0; sqi += a1; se += a2; }
            }
            __shared__ float s_wtot[3 * (kAcq64Threads / 32)];
            __syncthreads();
            if (lane == 31) { s_wtot[warp] = sqr; s_wtot[NW + warp] = sqi; s_wtot[2 * NW + warp] = se; }
            __syncthreads();
            float oqr = sqr - tqr, oqi = sqi - tqi, oe = se - te;       // exclusive offset of this thread inside its warp
            for (int w2 = 0; w2 < warp; w2++) { oqr += s_wtot[w2]; oqi += s_wtot[NW + w2]; oe += s_wtot[2 * NW + w2]; }
#pragma unroll
            for (int r = 0; r < RUN; r++) {
                int i = tid * RUN + r;
                if (i <= NQ) s_q[i] = make_float2(qr[r] + oqr, qi[r] + oqi);
                if (i <= NE) s_e[i] = ee[r] + oe;
            }
            __syncthreads();
            int found = 0x7fffffff;
            for (int i = tid; i < kAcqChunk && base + i < d_end; i += kAcq64Threads) {
                float2 qa = s_q[i], qb = s_q[i + kSym];
                float pr = qb.x - qa.x, pi = qb.y - qa.y;
                float r1 = s_e[i + kSym] - s_e[i], r2 = s_e[i + 2 * kSym] - s_e[i + kSym];
                if (pr * pr + pi * pi > 0.5f * r1 * r2) { found = i; break; }
            }
            if (found != 0x7fffffff) atomicMin(&s_d0, (int)(base + found));
            __syncthreads();
            if (s_d0 != 0x7fffffff) break;
        }
        if (s_d0 == 0x7fffffff) {
            status = ST_NO_SYNC;
        } else {
            long d0 = s_d0, k_lo = d0 - 176, k_hi = d0 + 16;
            if (k_lo < -(kSym - 1)) k_lo = -(kSym - 1);
            offset = (long)ramp_argmax_pairs<kAcq64Threads>(x, M, k_lo, k_hi, s_lock, s_val, s_idx) - 1;
        }
    }
    if (status == ST_OK && offset < 0) status = ST_NEG_OFFSET;                       // src/receiver.rs:25
    if (status == ST_OK && (offset > M || M - offset < 800)) status = ST_TOO_SHORT;  // src/receiver.rs:27-29

    if (status != ST_OK) {
        if (tid == 0) {
            st->status = status; st->offset = (int32_t)offset; st->n_syms = 0; st->out_len = 0; st->f_delta = 0.0f; st->fstep = 0;
            a.status[stream] = status; a.out_len[stream] = 0;
            if (a.d_offset) a.d_offset[stream] = (int32_t)offset;
            if (a.d_f_delta) a.d_f_delta[stream] = 0.0f;
            if (a.d_nsyms) a.d_nsyms[stream] = 0;
        }
        return;
    }

    const float2 *x0 = x + offset;
    const long n_avail = M - offset;

    // ---- CFO estimate in f64 (80 or 160 terms per stream; src/receiver.rs:231-240) -------------------------------
    double acc0 = 0.0, acc1 = 0.0;
    if (tid < kSym) {
        float2 r2 = x0[2 * kSym + tid], r3 = x0[3 * kSym + tid], r4 = x0[4 * kSym + tid];
        if (CFO == 0) {
            // angle(r / l), naive complex division like num::Complex
            double lr = r3.x, li = r3.y, rr = r4.x, ri = r4.y, nn = lr * lr + li * li;
            acc0 = atan2((ri * lr - rr * li) / nn, (rr * lr + ri * li) / nn);
        } else {
            acc0 = (double)r2.x * r3.x + (double)r2.y * r3.y + (double)r3.x * r4.x + (double)r3.y * r4.y;
            acc1 = (double)r2.x * r3.y - (double)r2.y * r3.x + (double)r3.x * r4.y - (double)r3.y * r4.x;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, m);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, m);
    }
    if (lane == 0) { s_red[warp] = acc0; s_red[NW + warp] = acc1; }
    __syncthreads();
    if (warp != 0) return;                             // the rest (f64 atan2, channel estimate, header) is one warp's work
    double f_delta;
    {
        double t0 = 0.0, t1 = 0.0;
        for (int w2 = 0; w2 < kAcq64Threads / 32; w2++) { t0 += s_red[w2]; t1 += s_red[NW + w2]; }
        if (CFO == 0) f_delta = fabs((t0 / 80.0) / 80.0);
        else f_delta = atan2(t1, t0) / 80.0;
    }
    const uint64_t fstep = (uint64_t)(int64_t)llrint(-f_delta * (0.15915494309189533577 * 18446744073709551616.0));
    if (tid == 0) { st->fstep = fstep; st->f_delta = (float)f_delta; st->offset = (int32_t)offset; }

    // ---- channel estimate (src/receiver.rs:212-229) and header symbols: warp 0 -----------------------------------
    {
        const int g = lane >> 3, l = lane & 7;
        float2 *tr = s_tr + g * kTrGroup;
        float twr[8], twi[8];
        fft64_lane_twiddles(a.tables->w64, l, twr, twi);
        float hr[8], hi[8];
#pragma unroll
        for (int kb = 0; kb < 8; kb++) { hr[kb] = 0.0f; hi[kb] = 0.0f; }
#pragma unroll 1
        for (int pass = 0; pass < 2; pass++) {
            const int row = 5 + 4 * pass + g;                 // training rows 5..9
            const bool valid = row < 10;
            float zr[8], zi[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t n = row * kSym + kCp + l + 8 * j;
                float2 v = valid ? x0[n] : make_float2(0.0f, 0.0f);
                float c, s;
                phasor_from_turns(fstep * (uint64_t)n, c, s);
                cmul(v.x, v.y, c, s);
                zr[j] = v.x; zi[j] = v.y;
            }
            fft64_group(zr, zi, twr, twi, tr, l);
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                float2 it = a.tables->inv_training[l + 8 * kb];
                cmul(zr[kb], zi[kb], it.x, it.y);
                hr[kb] += zr[kb]; hi[kb] += zi[kb];
            }
        }
#pragma unroll
        for (int kb = 0; kb < 8; kb++) {
            hr[kb] += __shfl_xor_sync(0xffffffffu, hr[kb], 8);  hi[kb] += __shfl_xor_sync(0xffffffffu, hi[kb], 8);
            hr[kb] += __shfl_xor_sync(0xffffffffu, hr[kb], 16); hi[kb] += __shfl_xor_sync(0xffffffffu, hi[kb], 16);
            hr[kb] *= 0.2f; hi[kb] *= 0.2f;
            if (g == 0) {
                float inv = 1.0f / (hr[kb] * hr[kb] + hi[kb] * hi[kb]);
                st->h[l + 8 * kb] = make_float2(hr[kb], hi[kb]);
                st->g[l + 8 * kb] = make_float2(hr[kb] * inv, -hi[kb] * inv);
                if (a.d_h) a.d_h[(size_t)stream * 64 + l + 8 * kb] = make_float2(hr[kb], hi[kb]);
            }
        }
        __syncwarp();
        __threadfence_block();

        // header symbols (src/receiver.rs:86-89): the first HDR_SYMS data symbols, one per 8-lane group
        RxLane L;
        rx_lane_init(L, st, a.tables->w64, l);
        const bool valid = g < HDR_SYMS;
        float br, bi;
        phasor_from_turns(fstep * (uint64_t)((kHeadSyms + g) * kSym + kCp), br, bi);
        float zr[8], zi[8];
        rx_load_symbol(x0, (uint32_t)(n_avail > 0xffffffffL ? 0xffffffffL : n_avail), (uint32_t)g, valid, l, zr, zi);
#pragma unroll
        for (int j = 0; j < 8; j++) cmul(zr[j], zi[j], L.wr[j], L.wi[j]);      // src/receiver.rs:44-50 (intra-symbol part)
        rx_symbol<GUARD, PHASE>(L, br, bi, tr, l, zr, zi);
        if (valid) {
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                int rk = data_rank<GUARD>(l + 8 * kb);
                if (rk >= 0) s_car[g * D + rk] = (uint8_t)demap_point<MOD>(zr[kb], zi[kb]);
            }
        }
        __syncwarp();
        // the 128 header bits (u128 little-endian length, src/packets/mod.rs:20-32): lane b of ballot j holds bit 32j + b
        uint32_t hw[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b = 32 * j + lane, c = b / BPC, sh = b - c * BPC;
            hw[j] = __ballot_sync(0xffffffffu, (s_car[c] >> sh) & 1u);
        }
        if (lane == 0) {
            const uint64_t lo = (uint64_t)hw[0] | ((uint64_t)hw[1] << 32), hi64 = (uint64_t)hw[2] | ((uint64_t)hw[3] << 32);
            const long rows = (n_avail + kSym - 1) / kSym;                // src/receiver.rs:192-203
            const long s_rx = rows - kHeadSyms;
            const long avail_bytes = (s_rx * BPS) / 8 - 16;
            int stt = ST_OK;
            uint32_t n_syms = 0, out_len = 0;
            if (hi64 != 0 || avail_bytes < 0 || lo > (uint64_t)avail_bytes) {        // unsigned compare: bit 63 of a corrupt length must not pass
                stt = ST_BAD_HEADER;
            } else {
                const uint64_t plen = lo;
                const uint64_t nbits = kHeaderBits + 8 * plen;
                const uint64_t ncar = (nbits + BPC - 1) / BPC;
                n_syms = (uint32_t)((ncar + D - 1) / D);                             // <= s_rx because plen <= avail_bytes
                out_len = (uint32_t)(FEC ? (8 * plen) / 14 : plen);
                if (out_len > a.out_stride) { stt = ST_BAD_HEADER; n_syms = 0; out_len = 0; }
            }
            st->status = stt; st->n_syms = n_syms; st->out_len = out_len; st->plen = (uint32_t)lo; st->n_syms_rx = (uint32_t)s_rx;
            a.status[stream] = stt; a.out_len[stream] = out_len;
            if (a.d_offset) a.d_offset[stream] = (int32_t)offset;
            if (a.d_f_delta) a.d_f_delta[stream] = (float)f_delta;
            if (a.d_nsyms) a.d_nsyms[stream] = (uint32_t)s_rx;
        }
    }
}

}  // namespace ofdm
