// wide_kernels.cuh -- the 1024-subcarrier variant (BASELINE.json configs[3]; docs/SPEC.md section 9; no reference
// implementation exists, the layout is the reference's N=64 construction scaled by 16: CP = 256, L = 1280).
//
// First (correct, not yet tuned) implementation: one 256-thread CTA transforms one OFDM symbol at a time with a radix-4
// Stockham FFT in shared memory (5 stages, 4 points per thread); the stages around it are the same as in the N=64 kernels:
//   wide_acquire_kernel : sync (ramp correlation / sliding Schmidl-Cox on prefix sums), f64 CFO estimate, channel estimate
//                         from the 5 training symbols, header symbol decode -> per-stream state.
//   wide_decode_kernel  : per 28-symbol tile: load (CP stripped) -> derotate -> FFT-1024 -> equalise -> pilot phase (64
//                         pilots, block reduction) -> demap -> carrier bytes in smem -> Hamming / header strip -> bytes.
//   wide_tx_kernel      : bits -> constellation -> IFFT-1024 -> CP, two passes (max, then store once) like tx_tile_kernel.
#pragma once

#include "common.cuh"
#include "rx_kernels.cuh"
#include "tx_kernels.cuh"

namespace ofdm {
namespace wide {

constexpr int kN = 1024, kCpW = 256, kL = kN + kCpW, kHeadW = 10 * kL;
constexpr int kThreads = 256;
constexpr int kTileSymsW = 14;                   // symbols per CTA tile (multiple of 7: Hamming byte alignment)

__host__ __device__ __forceinline__ bool w_is_null(int k) { return k <= 95 || k == 512 || k >= 929; }
__host__ __device__ __forceinline__ int w_rem(int k) { return k < 512 ? k - 96 : k - 97; }          // index among the 832 used bins
__host__ __device__ __forceinline__ bool w_is_pilot(int k) { return !w_is_null(k) && w_rem(k) % 13 == 0; }
template <bool GUARD>
__host__ __device__ __forceinline__ int w_rank(int k)
{
    if (!GUARD) return k;
    if (w_is_null(k)) return -1;
    const int r = w_rem(k);
    return r % 13 == 0 ? -1 : r - r / 13 - 1;
}

struct WideTables {
    float2 lock[kL];            // locking_signal::<1280>
    float2 inv_training[kN];    // 1 / training_signals::<1024>
    float2 w1024[kN];           // W1024^k
    float2 head[kHeadW];        // un-normalised lock | preamble x4 | (CP + IFFT(training)) x5
    float  head_max;
};

struct __align__(16) StreamStateW {
    int32_t  status;
    int32_t  offset;
    uint32_t n_syms;
    uint32_t out_len;
    uint64_t fstep;
    float    f_delta;
    uint32_t plen;
    uint32_t n_syms_rx;
    uint32_t pad[3];
    float2   g[kN];
    float2   h[kN];
};

struct WideRxArgs {
    const float2   *iq;
    const uint32_t *n_samples;
    uint32_t        iq_stride;
    uint32_t        n_streams;
    StreamStateW   *state;
    const WideTables *tables;
    uint8_t        *out;
    uint32_t        out_stride;
    uint32_t       *out_len;
    int32_t        *status;
    uint32_t        sync_window;
    int32_t         tile_shift;
    int32_t         sync_mode, cfo_mode, fec;
    int32_t         lock_is_ramp;   // the locking table is the built-in ramp: closed-form ramp correlation
    int32_t         tiles_per_cta;  // consecutive tiles of one stream handled by one CTA of the decode kernel
    int32_t  *d_offset;
    float    *d_f_delta;
    float2   *d_h;              // [n_streams][1024]
    uint32_t *d_nsyms;
    float2   *d_points;
    uint32_t  points_stride;
    uint32_t  stream0;          // decode kernel: first stream of this launch (gridDim.y <= 65535)
    const uint64_t *stream_base;    // optional: sample index of every stream's capture inside iq (NULL: stream * iq_stride)
};

// ---- block FFT: 1024 points, 256 threads, radix-4 Stockham autosort -------------------------------------------------
// in : v[r] = x[tid + 256 r].  out: natural-order spectrum in `bufA`. `s_w` = W1024 table in smem. 5 barriers.
__device__ __forceinline__ void radix4(cpx (&v)[4])
{
    const cpx a = c_add(v[0], v[2]), b = c_sub(v[0], v[2]), c = c_add(v[1], v[3]), d = c_mul_mj(c_sub(v[1], v[3]));
    v[0] = c_add(a, c); v[2] = c_sub(a, c); v[1] = c_add(b, d); v[3] = c_sub(b, d);
}
// Padded layouts keep every stage free of bank conflicts (64-bit accesses, 16 lanes per phase): buffer A inserts one
// element per 16 (stage-0 writes A[4t + r]), buffer B four per 16 (stage-1 writes B[16m + k + 4r], t = 4m + k).
constexpr int kBufA = kN + kN / 16, kBufB = kN + kN / 4;
__device__ __forceinline__ int padA(int i) { return i + (i >> 4); }
__device__ __forceinline__ int padB(int i) { return i + ((i >> 4) << 2); }

// per-thread twiddles of stages 1..4 (W_{4 Ns}^{k}, k = tid mod Ns): constant for the thread, kept in registers; the powers
// 2 and 3 are formed by two complex multiplies per stage instead of (bank-conflicting) table look-ups
struct FftTw { cpx w[4]; };
__device__ __forceinline__ void fft1024_tw_init(FftTw &T, const float2 *__restrict__ w1024, int tid)
{
#pragma unroll
    for (int st = 1; st < 5; st++) {
        const int Ns = 1 << (2 * st);
        T.w[st - 1] = c_from(__ldg(w1024 + (tid & (Ns - 1)) * (256 >> (2 * st))));
    }
}

__device__ __forceinline__ void fft1024_block(cpx (&v)[4], float2 *bufA, float2 *bufB, const FftTw &T, int tid)
{
    unsigned long long *A = reinterpret_cast<unsigned long long *>(bufA), *B = reinterpret_cast<unsigned long long *>(bufB);
    // stage 0 (Ns = 1): no twiddles, out[4 j + r]
    radix4(v);
#pragma unroll
    for (int r = 0; r < 4; r++) A[padA(4 * tid + r)] = v[r].v;
    __syncthreads();
#pragma unroll
    for (int st = 1; st < 5; st++) {
        const int Ns = 1 << (2 * st);
        const int k = tid & (Ns - 1);
        const bool fromA = (st & 1) != 0;                          // stages 1, 3 read A and write B; 2, 4 read B and write A
#pragma unroll
        for (int r = 0; r < 4; r++) v[r].v = fromA ? A[padA(tid + 256 * r)] : B[padB(tid + 256 * r)];
        const cpx w1 = T.w[st - 1], w2 = c_mul(w1, w1), w3 = c_mul(w2, w1);
        v[1] = c_mul(v[1], w1); v[2] = c_mul(v[2], w2); v[3] = c_mul(v[3], w3);
        radix4(v);
        const int j0 = ((tid - k) << 2) + k;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (fromA) B[padB(j0 + r * Ns)] = v[r].v;
            else A[padA(j0 + r * Ns)] = v[r].v;
        }
        __syncthreads();
    }
    // stage 4 wrote A: natural-order spectrum at A[padA(k)]
}

// sum over the block of a packed complex (for the pilot sum) or a float (angles); s_red: 2 * 8 floats
__device__ __forceinline__ void block_sum2(float &x, float &y, float *s_red, int tid)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { x += __shfl_xor_sync(0xffffffffu, x, m); y += __shfl_xor_sync(0xffffffffu, y, m); }
    if ((tid & 31) == 0) { s_red[tid >> 5] = x; s_red[8 + (tid >> 5)] = y; }
    __syncthreads();
    float sx = 0.0f, sy = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; w++) { sx += s_red[w]; sy += s_red[8 + w]; }
    x = sx; y = sy;
}

// per-thread constants of the symbol pipeline
struct WideLane {
    cpx w[4];        // derotation exp(-j f (tid + 256 r))
    cpx g[4];        // equaliser 1/h at bins tid + 256 r
    int rank[4];     // data-carrier rank of those bins or -1
    uint32_t pilot;  // bit r: bin tid + 256 r is a pilot
};

template <bool GUARD>
__device__ __forceinline__ void wide_lane_init(WideLane &L, const StreamStateW *st, int tid)
{
    L.pilot = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int k = tid + 256 * r;
        L.w[r] = phasor_from_turns_p(st->fstep * (uint64_t)k);
        L.g[r] = c_from(st->g[k]);
        L.rank[r] = w_rank<GUARD>(k);
        if (GUARD && w_is_pilot(k)) L.pilot |= 1u << r;
    }
}

__device__ __forceinline__ void wide_load_symbol(const float2 *__restrict__ x0, uint32_t n_avail, uint32_t sym, bool valid, int tid, cpx (&v)[4])
{
    const uint32_t n0 = (10u + sym) * kL + kCpW + tid;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint32_t n = n0 + 256u * r;
        v[r].v = 0ull;
        if (valid && n < n_avail) v[r].v = __ldg(reinterpret_cast<const unsigned long long *>(x0 + n));    // zero-padded tail row
    }
}

// One OFDM symbol by the whole CTA: derotate, FFT, equalise, pilot phase. On return z[r] = corrected bin tid + 256 r.
template <bool GUARD, int PHASE>
__device__ __forceinline__ void wide_symbol(const WideLane &L, cpx base, cpx (&z)[4], float2 *bufA, float2 *bufB, const FftTw &T,
                                            float *s_red, int tid)
{
#pragma unroll
    for (int r = 0; r < 4; r++) z[r] = c_mul(z[r], L.w[r]);
    fft1024_block(z, bufA, bufB, T, tid);
    const unsigned long long *A = reinterpret_cast<const unsigned long long *>(bufA);
#pragma unroll
    for (int r = 0; r < 4; r++) { cpx x; x.v = A[padA(tid + 256 * r)]; z[r] = c_mul(x, L.g[r]); }
    cpx rot = base;
    if (GUARD) {
        float px = 0.0f, py = 0.0f;
        if (PHASE == 1) {
            cpx p = c_make(0.0f, 0.0f);
#pragma unroll
            for (int r = 0; r < 4; r++) if (L.pilot & (1u << r)) p = c_add(p, z[r]);
            c_split(p, px, py);
            block_sum2(px, py, s_red, tid);
            const float inv = rsqrt_normal(fmaxf(px * px + py * py, 1e-30f));
            rot = c_make(px * inv, -py * inv);
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (L.pilot & (1u << r)) { float a, b; c_split(c_mul(z[r], base), a, b); px += atan2f(b, a); }
            block_sum2(px, py, s_red, tid);
            float sn, cs;
            sincosf(-px * (1.0f / 64.0f), &sn, &cs);
            rot = c_mul(c_make(cs, sn), base);
        }
    } else {
        __syncthreads();                                          // same barrier count on every path (bufA reuse)
    }
#pragma unroll
    for (int r = 0; r < 4; r++) z[r] = c_mul(z[r], rot);
}

// ---- tile carrier bytes -> payload bytes (shared with nothing else: same scheme as rx_decode_kernel's second phase) ----
template <int MOD, bool FEC, int D, int NT>
__device__ __forceinline__ void wide_tile_bytes(const uint8_t *s_car, const uint8_t *s_ham, int t0, int t1, uint32_t out_len, uint8_t *out,
                                                int tid)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int BPS = BPC * D;
    constexpr int NB = FEC ? 14 : 8;
    const long bit0 = (long)t0 * BPS, bit1 = (long)t1 * BPS;
    long j0 = bit0 <= kHeaderBits ? 0 : (bit0 - kHeaderBits + NB - 1) / NB;
    long j1 = (bit1 - kHeaderBits) / NB;
    if (j1 > (long)out_len) j1 = (long)out_len;
    const int pbase = (int)(kHeaderBits + j0 * NB - bit0);
    const int nbytes = (int)(j1 - j0);
    if (MOD == 2) {
        // 3 output bytes per thread: 42 (Hamming) or 24 stream bits = 7 or 4 six-bit carriers (+1 for the bit shift)
        constexpr int NC = (3 * NB + 5) / 6 + 1;
        uint8_t *o0 = out + j0;
        // as in rx_decode_kernel: 3 NB bits = whole carriers -> one bit shift per tile; with FEC the 8 carrier bytes come
        // from 3 aligned shared words + 2 byte permutes (a thread's byte alignment never changes: stride % 4 == 0)
        static_assert((3 * NB) % 6 == 0 && ((NB / 2) * NT) % 4 == 0, "whole carriers per 3 bytes, word-aligned stride");
        const int c0 = pbase / 6, sh = pbase - 6 * c0;
        const uint8_t *cp = s_car + c0 + (NB / 2) * tid;
        const uint32_t cp_s = (uint32_t)__cvta_generic_to_shared(cp);
        uint32_t wa = cp_s & ~3u;
        const uint32_t sel = 0x3210u + 0x1111u * (cp_s & 3u);
        for (int u = 3 * tid; u < nbytes; u += 3 * NT, cp += (NB / 2) * NT, wa += (NB / 2) * NT) {
            uint32_t lo, hi = 0;
            if (NC > 5) {
                uint32_t x0, x1, x2;
                asm volatile("ld.shared.u32 %0, [%3];\n\tld.shared.u32 %1, [%3+4];\n\tld.shared.u32 %2, [%3+8];"
                             : "=r"(x0), "=r"(x1), "=r"(x2) : "r"(wa));
                const uint32_t p0 = pack4x6(__byte_perm(x0, x1, sel)), p1 = pack4x6(__byte_perm(x1, x2, sel));
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
            uint8_t *o = o0 + u;
            o[0] = (uint8_t)w0;
            st_global_u8_if(o + 1, w1, u + 1 < nbytes);
            st_global_u8_if(o + 2, w2, u + 2 < nbytes);
        }
    } else {
    for (int u = tid; u < nbytes; u += NT) {
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
}

// ---- decode kernel: register-resident 1024-point FFT as 32 x 32, one OFDM symbol per WARP ---------------------------------
// With n = l + 32 j (l = lane, j = register) and k = k1 + 32 k2:
//     X[k1 + 32 k2] = sum_l W32^(l k2) { W1024^(l k1) sum_j x[l + 32 j] W32^(j k1) }
//   A. lane l reads its 32 samples x[l + 32 j] from the warp's staging buffer (the CP-stripped symbol, copied there by ONE TMA
//      bulk copy, cp.async.bulk + mbarrier, issued while the previous symbol was still being processed), applies the part of the
//      CFO derotation that is uniform over the lanes, exp(-j f 32 j), runs a 32-point DFT in registers and multiplies by
//      W1024^(l k1) exp(-j f l) (inter-stage twiddle and the rest of the derotation in one per-stream table);
//   B. 32 x 32 exchange through the same buffer (rows of 34 complex: STS.64 by column and LDS.128 by row are conflict-free);
//      as soon as the rows are back in registers the buffer is handed to the TMA engine for the warp's next symbol;
//   C. second 32-point DFT in registers: lane k1 holds X[k1 + 32 k2], k2 = 0..31;
//   D. equalise (1/h rows in the same 34-pitch layout), pilot sum (predicated adds + 5 xor-shuffles), rotate, LUT demap,
//      carrier bytes of the tile (double-buffered).
//   E. warp 7 does nothing but turn the carrier bytes of finished tiles into payload bytes (Hamming / header strip) while
//      the seven compute warps are already on the next tile: full[2] / empty[2] mbarriers, producer / consumer.
// No CTA or named barrier after the table set-up: a compute warp only needs __syncwarp and never waits for its siblings.
// 7 warps x 2 symbols = one 14-symbol tile; a CTA owns `tiles_per_cta` consecutive tiles of one stream (tables and the
// prefetch pipeline are reused across them).
constexpr int kWDecWarps = 7;                                   // compute warps; one more warp turns carrier bytes into payload bytes
constexpr int kWDecThreads = (kWDecWarps + 1) * 32;
constexpr int kWSymsPerWarp = kTileSymsW / kWDecWarps;          // 2
constexpr int kWPitch = 34;                                     // complex per row of the 32 x 32 layouts
constexpr int kWBuf = 32 * kWPitch;                             // complex per warp buffer: 8704 B >= staged symbol + alignment sample
constexpr int kWStageBytes = kN * 8 + 16;
constexpr int kWPw = 36;                                        // floats per row of the pilot weight table (conflict-free LDS.128 by row)
static_assert(kWBuf * 8 >= kWStageBytes && kTileSymsW % kWDecWarps == 0, "wide decode layout");
constexpr size_t wide_decode_smem(bool guard)
{
    return sizeof(float2) * (kWDecWarps * kWBuf + 2 * kWBuf + 32 + 32) + sizeof(float) * 32 * kWPw + sizeof(int16_t) * kN +
           2 * ((size_t)kTileSymsW * (guard ? 768 : 1024) + 64) + 128 + 256 + 8 * 16;
}

// W32^m = exp(-2 pi j m / 32), m = 0 .. 21 (the inter-stage twiddles of the 8 x 4 split below)
__device__ constexpr float kW32r[22] = { 1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654757f, 0.55557023301960229f,
    0.38268343236508984f, 0.19509032201612833f, 0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f, -0.70710678118654746f,
    -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f, -1.0f, -0.98078528040323043f, -0.92387953251128685f, -0.83146961230254546f,
    -0.70710678118654768f, -0.55557023301960218f };
__device__ constexpr float kW32i[22] = { -0.0f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f, -0.70710678118654746f, -0.83146961230254524f,
    -0.92387953251128674f, -0.98078528040323043f, -1.0f, -0.98078528040323043f, -0.92387953251128674f, -0.83146961230254546f, -0.70710678118654757f,
    -0.55557023301960218f, -0.38268343236508989f, -0.19509032201612861f, 0.0f, 0.19509032201612836f, 0.38268343236508967f, 0.55557023301960196f,
    0.70710678118654746f, 0.83146961230254524f };

// 32-point forward DFT in registers, natural order in and out: n = n2 + 4 n1, k = k1 + 8 k2 -> four DFT8 over n1, twiddles
// W32^(n2 k1), eight DFT4 over n2. Everything is unrolled, so the index maps are register renaming.
__device__ __forceinline__ void dft32_p(cpx (&x)[32])
{
    cpx y[4][8];
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) {
        cpx u[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; n1++) u[n1] = x[n2 + 4 * n1];
        dft8_p(u);
#pragma unroll
        for (int k1 = 0; k1 < 8; k1++) {
            const int m = n2 * k1;
            if (m == 0) y[n2][k1] = u[k1];
            else if (m == 8) y[n2][k1] = c_mul_mj(u[k1]);                          // W32^8 = -j
            else y[n2][k1] = c_mul(u[k1], c_make(kW32r[m], kW32i[m]));
        }
    }
#pragma unroll
    for (int k1 = 0; k1 < 8; k1++) {
        cpx u[4] = { y[0][k1], y[1][k1], y[2][k1], y[3][k1] };
        radix4(u);
#pragma unroll
        for (int k2 = 0; k2 < 4; k2++) x[k1 + 8 * k2] = u[k2];
    }
}

// With guard bands the bins k1 + 32 k2 with k2 in {0, 1, 2, 30, 31} are null carriers on every lane (k <= 95 or k >= 960):
// stage D skips them at compile time, and the unrolled second DFT drops whatever only feeds them.
template <bool GUARD> __device__ __forceinline__ constexpr bool w_row_used(int k2) { return !GUARD || (k2 >= 3 && k2 <= 29); }

template <int MOD, bool GUARD, bool FEC, int PHASE, bool POINTS>
__global__ void __launch_bounds__(kWDecThreads, 2) wide_decode_kernel(const WideRxArgs a)
{
    constexpr int D = GUARD ? 768 : 1024;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float2 *s_buf = reinterpret_cast<float2 *>(smem_raw);                        // [warp][kWBuf]: staged symbol, then the 32 x 32 exchange
    float2 *s_tw = s_buf + kWDecWarps * kWBuf;                                   // [l][kWPitch]: W1024^(l k1) exp(-j f l)
    float2 *s_g = s_tw + kWBuf;                                                  // [k1][kWPitch]: 1/h at bin k1 + 32 k2
    float2 *s_step = s_g + kWBuf;                                                // [32]: exp(-j f 32 j)
    float2 *s_lanew = s_step + 32;                                               // [32]: exp(-j f l) (table set-up only)
    float *s_pw = reinterpret_cast<float *>(s_lanew + 32);                       // [k1][kWPw]: 1 where bin k1 + 32 k2 is a pilot, else 0
    int16_t *s_rk = reinterpret_cast<int16_t *>(s_pw + 32 * kWPw);               // [k2][k1]: data-carrier rank of bin k1 + 32 k2, or -1
    constexpr int kCarBuf = kTileSymsW * D + 64;
    uint8_t *s_car = reinterpret_cast<uint8_t *>(s_rk + kN);                     // [2][kCarBuf]: carrier bytes of the tile in flight / being converted
    uint8_t *s_ham = s_car + 2 * kCarBuf;
    uint8_t *s_qam = s_ham + 128;                                                // 64QAM demap table (see demap_qam64_lut)
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_qam + 256);                 // [0..6] staging, [8..9] tile full, [10..11] tile empty

    const uint32_t stream = blockIdx.y + a.stream0;
    const StreamStateW *st = a.state + stream;
    if (st->status != ST_OK) return;
    const int S = (int)st->n_syms;
    const int tile_first = (int)blockIdx.x * a.tiles_per_cta;
    int tile_end = tile_first + a.tiles_per_cta;
    {
        const int n_tiles = (S + a.tile_shift + kTileSymsW - 1) / kTileSymsW;
        if (tile_end > n_tiles) tile_end = n_tiles;
    }
    if (tile_first >= tile_end) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t fstep = st->fstep;
    const uint32_t offset = (uint32_t)st->offset;
    const float2 *x0 = a.iq + (a.stream_base ? (size_t)a.stream_base[stream] : (size_t)stream * a.iq_stride) + offset;
    const uint32_t n_avail = a.n_samples[stream] - offset;
    auto tile_t0 = [&](int tile) { int t = tile * kTileSymsW - a.tile_shift; return t < 0 ? 0 : t; };
    auto tile_t1 = [&](int tile) { int t = (tile + 1) * kTileSymsW - a.tile_shift; return t > S ? S : t; };

    // ---- TMA prefetch pipeline: started before the tables are built, so the first symbol's DRAM latency hides behind them ----
    // Symbols below s_fast lie, with the two alignment samples of their 16-byte aligned superset, inside the capture; the
    // (rare) others take the bounds-checked direct-load path.
    const uintptr_t xaddr = reinterpret_cast<uintptr_t>(x0 + (10 * kL + kCpW));  // sample 0 of data symbol 0 (CP skipped)
    const int shift = (int)((xaddr >> 3) & 1);                                   // symbol pitch 10 240 B keeps the 16-byte phase
    int s_fast = n_avail >= (uint32_t)(10 * kL + kCpW + kN + 2) ? (int)((n_avail - (10 * kL + kCpW + kN + 2)) / kL) + 1 : 0;
    if (s_fast > S) s_fast = S;
    float2 *buf = s_buf + warp * kWBuf;
    uint64_t *bar = s_bar + warp;
    uint32_t buf_sa = smem_addr(buf), bar_sa = smem_addr(bar);
    asm volatile("" : "+r"(buf_sa), "+r"(bar_sa));
    auto issue = [&](int sym, int lim) {
        if (lane == 0) {
            const bool full = sym < lim;
            if (full) {
                const uintptr_t src = (xaddr + (uintptr_t)sym * (kL * 8)) & ~(uintptr_t)15;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(buf_sa), "l"(src), "r"((uint32_t)kWStageBytes), "r"(bar_sa) : "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_sa), "r"(full ? (uint32_t)kWStageBytes : 0u) : "memory");
        }
    };
    const bool is_out_warp = warp == kWDecWarps;
    uint64_t *bar_full = s_bar + 8, *bar_empty = s_bar + 10;
    if (tid == 0) { mbar_init(bar_full, kWDecWarps); mbar_init(bar_full + 1, kWDecWarps); mbar_init(bar_empty, 1); mbar_init(bar_empty + 1, 1); }
    if (lane == 0 && !is_out_warp) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (!is_out_warp) {
        const int t1 = tile_t1(tile_first);
        issue(tile_t0(tile_first) + warp * kWSymsPerWarp, s_fast < t1 ? s_fast : t1);
    }

    // ---- per-stream tables ------------------------------------------------------------------------------------------------
    if (FEC && tid < 128) s_ham[tid] = (uint8_t)ham74_decode_word(tid);
    if (MOD == 2) for (int i = tid; i < 256; i += kWDecThreads) s_qam[i] = qam64_lut_entry(i);
    if (tid < 32) s_step[tid] = c_to(phasor_from_turns_p(fstep * (uint64_t)(32 * tid)));
    else if (tid < 64) s_lanew[tid - 32] = c_to(phasor_from_turns_p(fstep * (uint64_t)(tid - 32)));
    for (int e = tid; e < kN; e += kWDecThreads) {
        const int r = e >> 5, c = e & 31;
        s_g[c * kWPitch + r] = st->g[e];                                         // bin e = k1 + 32 k2 with k1 = c, k2 = r
        const int rank = w_rank<GUARD>(e);
        s_rk[e] = (int16_t)rank;
        s_pw[c * kWPw + r] = (GUARD && rank < 0 && !w_is_null(e)) ? 1.0f : 0.0f;
    }
    __syncthreads();
    for (int e = tid; e < kN; e += kWDecThreads) {
        const int r = e >> 5, c = e & 31;
        s_tw[r * kWPitch + c] = c_to(c_mul(c_from(__ldg(a.tables->w1024 + ((r * c) & (kN - 1)))), c_from(s_lanew[r])));
    }
    if (tid == 0) mbar_fence_init();
    const uint32_t qam_saddr = (uint32_t)__cvta_generic_to_shared(s_qam);
    const unsigned long long *rd = reinterpret_cast<const unsigned long long *>(buf + shift + lane);
    unsigned long long *wr = reinterpret_cast<unsigned long long *>(buf + lane);
    const ulonglong2 *row = reinterpret_cast<const ulonglong2 *>(buf + lane * kWPitch);
    const ulonglong2 *tw_row = reinterpret_cast<const ulonglong2 *>(s_tw + lane * kWPitch);
    const ulonglong2 *g_row = reinterpret_cast<const ulonglong2 *>(s_g + lane * kWPitch);
    const float4 *pw_row = reinterpret_cast<const float4 *>(s_pw + lane * kWPw);
    const unsigned long long *step64 = reinterpret_cast<const unsigned long long *>(s_step);
    const int16_t *rk = s_rk + lane;
    uint8_t *out = a.out + (size_t)stream * a.out_stride;
    uint32_t phase = 0;
    __syncthreads();

    if (is_out_warp) {
        // ---- E: consumer warp: carrier bytes of every finished tile -> payload bytes ----------------------------------------
#pragma unroll 1
        for (int tile = tile_first; tile < tile_end; tile++) {
            const int u = tile - tile_first, p = u & 1;
            mbar_wait(bar_full + p, (uint32_t)(u >> 1) & 1u);
            wide_tile_bytes<MOD, FEC, D, 32>(s_car + p * kCarBuf, s_ham, tile_t0(tile), tile_t1(tile), st->out_len, out, lane);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar_empty + p)) : "memory");
        }
        return;
    }

#pragma unroll 1
    for (int tile = tile_first; tile < tile_end; tile++) {
        const int t0 = tile_t0(tile), t1 = tile_t1(tile);
        const int lim = s_fast < t1 ? s_fast : t1;
        const int u = tile - tile_first, p = u & 1;
        if (u >= 2) mbar_wait(bar_empty + p, (uint32_t)((u >> 1) - 1) & 1u);      // the consumer is done with this buffer (tile u - 2)
        uint8_t *car = s_car + p * kCarBuf;
#pragma unroll 1
        for (int i = 0; i < kWSymsPerWarp; i++) {
            const int s = t0 + warp * kWSymsPerWarp + i;
            cpx x[32];
            mbar_wait(bar, phase);
            phase ^= 1;
            auto issue_next = [&]() {                                            // prefetch this warp's next symbol (possibly of the next tile)
                if (i + 1 < kWSymsPerWarp) {
                    issue(s + 1, lim);
                } else if (tile + 1 < tile_end) {
                    const int n1 = tile_t1(tile + 1);
                    issue(t1 + warp * kWSymsPerWarp, s_fast < n1 ? s_fast : n1);
                }
            };
            if (s >= t1) { __syncwarp(); issue_next(); continue; }               // past the frame's last symbol (every lane is past the wait before the barrier is re-armed)
            // ---- A: samples, uniform derotation, DFT32, twiddle ------------------------------------------------------------
            if (s < lim) {
#pragma unroll
                for (int j = 0; j < 32; j++) x[j].v = rd[32 * j];
            } else {
                const uint32_t n0 = (10u + (uint32_t)s) * kL + kCpW + (uint32_t)lane;
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t n = n0 + 32u * j;
                    x[j].v = (s < t1 && n < n_avail) ? __ldg(reinterpret_cast<const unsigned long long *>(x0 + n)) : 0ull;   // zero-padded tail row
                }
            }
#pragma unroll
            for (int j = 1; j < 32; j++) { cpx w; w.v = step64[j]; x[j] = c_mul(x[j], w); }
            dft32_p(x);
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const ulonglong2 ww = tw_row[q];
                cpx w0, w1;
                w0.v = ww.x; w1.v = ww.y;
                x[2 * q] = c_mul(x[2 * q], w0); x[2 * q + 1] = c_mul(x[2 * q + 1], w1);
            }
            __syncwarp();                                                        // every lane has read its staged samples
            // ---- B: 32 x 32 exchange --------------------------------------------------------------------------------------
#pragma unroll
            for (int k1 = 0; k1 < 32; k1++) wr[k1 * kWPitch] = x[k1].v;
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 16; q++) { const ulonglong2 v = row[q]; x[2 * q].v = v.x; x[2 * q + 1].v = v.y; }
            __syncwarp();                                                        // buffer free: prefetch this warp's next symbol
            issue_next();
            // ---- C: second DFT32 -> lane k1 holds X[k1 + 32 k2] --------------------------------------------------------------
            dft32_p(x);
            // ---- D: equalise, pilot phase, demap ----------------------------------------------------------------------------
#pragma unroll
            for (int q = 0; q < 16; q++) {
                if (!w_row_used<GUARD>(2 * q) && !w_row_used<GUARD>(2 * q + 1)) continue;
                const ulonglong2 gg = g_row[q];
                cpx g0, g1;
                g0.v = gg.x; g1.v = gg.y;
                if (w_row_used<GUARD>(2 * q)) x[2 * q] = c_mul(x[2 * q], g0);
                if (w_row_used<GUARD>(2 * q + 1)) x[2 * q + 1] = c_mul(x[2 * q + 1], g1);
            }
            cpx base = c_make(1.0f, 0.0f);
            if (!GUARD || PHASE == 0) base = phasor_from_turns_p(fstep * (uint64_t)((10 + s) * kL + kCpW));
            cpx rot = base;
            if (GUARD) {
                float px, py;
                if (PHASE == 1) {
                    // pilot sum as a multiply-accumulate with the 0 / 1 weights of this lane's bins: 8 LDS.128 + 27 FFMA2
                    cpx psum = c_make(0.0f, 0.0f);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 w4 = pw_row[q];
                        if (w_row_used<GUARD>(4 * q)) psum = c_fma2(x[4 * q], c_make(w4.x, w4.x), psum);
                        if (w_row_used<GUARD>(4 * q + 1)) psum = c_fma2(x[4 * q + 1], c_make(w4.y, w4.y), psum);
                        if (w_row_used<GUARD>(4 * q + 2)) psum = c_fma2(x[4 * q + 2], c_make(w4.z, w4.z), psum);
                        if (w_row_used<GUARD>(4 * q + 3)) psum = c_fma2(x[4 * q + 3], c_make(w4.w, w4.w), psum);
                    }
                    c_split(psum, px, py);
                } else {
                    // reference: mean of the 64 pilot angles (src/receiver.rs:126,137 scaled to 64 pilots, docs/SPEC.md 9)
                    px = 0.0f; py = 0.0f;
#pragma unroll
                    for (int k2 = 0; k2 < 32; k2++) {
                        if (!w_row_used<GUARD>(k2)) continue;
                        if (s_pw[lane * kWPw + k2] != 0.0f) { float pa, pb; c_split(c_mul(x[k2], base), pa, pb); px += atan2f(pb, pa); }
                    }
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) { px += __shfl_xor_sync(0xffffffffu, px, m); py += __shfl_xor_sync(0xffffffffu, py, m); }
                if (PHASE == 1) {
                    const float inv = rsqrt_normal(fmaxf(px * px + py * py, 1e-30f));
                    rot = c_make(px * inv, -py * inv);
                } else {
                    float sn, cs;
                    sincosf(-px * (1.0f / 64.0f), &sn, &cs);
                    rot = c_mul(c_make(cs, sn), base);
                }
            }
            uint8_t *crow = car + (s - t0) * D;                                  // rows of symbols past t1 exist but are never read
#pragma unroll
            for (int k2 = 0; k2 < 32; k2++) {
                if (!w_row_used<GUARD>(k2)) continue;
                float zr, zi;
                c_split(c_mul(x[k2], rot), zr, zi);
                const int r = rk[32 * k2];
                // branch-free: null / pilot bins are demapped like data bins and simply not stored
                const uint32_t sym6 = MOD == 2 ? demap_qam64_lut(zr, zi, qam_saddr) : demap_point<MOD>(zr, zi);
                st_shared_u8_if_nonneg(crow + r, sym6, r);
                if (POINTS && r >= 0 && s < t1) {
                    size_t p = (size_t)s * D + r;
                    if (p < a.points_stride) a.d_points[(size_t)stream * a.points_stride + p] = make_float2(zr, zi);
                }
            }
        }
        __syncwarp();                                                            // this warp's carrier bytes of the tile are written
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar_full + p)) : "memory");
    }
}

// ---- acquisition ---------------------------------------------------------------------------------------------------------
constexpr int kAcqChunkW = 1024;
constexpr int kAcqNQ = kAcqChunkW + kL, kAcqNE = kAcqChunkW + 2 * kL;
constexpr size_t wide_acquire_smem()
{
    // lock ramp | union { S&C prefix arrays , FFT buffers + twiddles + header carrier bytes }
    return sizeof(float) * kL + sizeof(float2) * (kAcqNQ + 1) + sizeof(float) * (kAcqNE + 1) + 256;
}

// ramp correlation |c[k]|^2, c[k] = sum_{n<1280} a[n+k] lock[n], for kRampTile consecutive lags per thread: every loaded
// sample feeds kRampTile accumulators (register tiling; 1 load per 2*kRampTile FMAs instead of 1 per 2)
constexpr int kRampTile = 8;
__device__ __forceinline__ void ramp_corr_tile_w(const float2 *__restrict__ x, long k0, long n_samples, const float *s_lock, float (&out)[kRampTile])
{
    float cr[kRampTile], ci[kRampTile];
#pragma unroll
    for (int j = 0; j < kRampTile; j++) { cr[j] = 0.0f; ci[j] = 0.0f; }
    const bool inside = k0 >= 0 && k0 + kL + kRampTile <= n_samples;
    // sample m (relative to k0) contributes to lag j with tap m - j
    float tap[kRampTile];                                     // tap[j] = lock[m - j]
#pragma unroll
    for (int j = 0; j < kRampTile; j++) tap[j] = 0.0f;
    for (int m = 0; m < kL + kRampTile - 1; m++) {
#pragma unroll
        for (int j = kRampTile - 1; j > 0; j--) tap[j] = tap[j - 1];
        tap[0] = m < kL ? s_lock[m] : 0.0f;
        const float2 v = inside ? __ldg(x + k0 + m) : ld_sample(x, k0 + m, n_samples);
#pragma unroll
        for (int j = 0; j < kRampTile; j++) { cr[j] = fmaf(v.x, tap[j], cr[j]); ci[j] = fmaf(v.y, tap[j], ci[j]); }
    }
#pragma unroll
    for (int j = 0; j < kRampTile; j++) out[j] = cr[j] * cr[j] + ci[j] * ci[j];
}
__device__ __forceinline__ long ramp_argmax_w(const float2 *__restrict__ x, long n_samples, long k_lo, long k_hi, const float *s_lock,
                                              float *s_val, int *s_idx)
{
    float best = 0.0f;
    int bidx = 0x7fffffff;
    for (long k = k_lo + (long)threadIdx.x * kRampTile; k <= k_hi; k += (long)kThreads * kRampTile) {
        float v[kRampTile];
        ramp_corr_tile_w(x, k, n_samples, s_lock, v);
#pragma unroll
        for (int j = 0; j < kRampTile; j++)
            if (k + j <= k_hi && v[j] > best) { best = v[j]; bidx = (int)(k + j - k_lo); }     // ascending lags: strict > keeps the first
    }
    block_argmax<kThreads>(best, bidx, s_val, s_idx);
    return best > 0.0f ? k_lo + bidx : k_lo;
}

// Closed form of the same arg-max for the built-in locking ramp: lock[n] = 3/8 + n/(4L) for n < L/2 and 1/8 + n/(4L) above
// (locking_signal::<1280>, src/transmitter.rs:60-72 after fft_shift), so with i counting samples from lag k_lo
//   c[q] = 3/8 S0a(q) + 1/8 S0b(q) + (M(q) - q S0(q)) / (4L),
//   S0a/S0b = sums of a over the two half windows of lag q, S0 = S0a + S0b, M(q) = sum_{i=q}^{q+L-1} i a[i].
// S0a, S0b and M slide by data-only increments, so the start values of every thread's run of lags come from one block
// reduction (lag k_lo: L samples over the CTA) plus a block prefix scan of the per-run increments; each thread then slides
// through its run. All in f64. Samples read per stream: about 7 (L + lags) instead of L per thread.
__device__ __forceinline__ long ramp_argmax_closed_w(const float2 *__restrict__ x, long n_samples, long k_lo, long k_hi, float *s_val, int *s_idx)
{
    __shared__ double s_ramp[2][kThreads / 32][6];
    const long n_lags = k_hi - k_lo + 1;
    const int run = (int)((n_lags + kThreads - 1) / kThreads);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int H = kL / 2;
    const float2 *xl = x + k_lo;                                       // local index 0 = lag k_lo
    const long nl = n_samples - k_lo;                                  // local indices [-k_lo, nl) exist
    auto ld = [&](long i) -> float2 { return (i >= -k_lo && i < nl) ? __ldg(xl + i) : make_float2(0.0f, 0.0f); };

    // (1) partial sums of lag k_lo's window, (2) this thread's increments over its run of lags [q0, q0 + run)
    double base[6] = { 0, 0, 0, 0, 0, 0 };                             // S0a, S0b, M (re, im each)
    for (int i = tid; i < kL; i += kThreads) {
        const float2 v = ld(i);
        if (i < H) { base[0] += v.x; base[1] += v.y; } else { base[2] += v.x; base[3] += v.y; }
        base[4] += (double)i * v.x; base[5] += (double)i * v.y;
    }
    const long q0 = (long)tid * run;
    double inc[6] = { 0, 0, 0, 0, 0, 0 };
    for (int j = 0; j < run; j++) {
        const long q = q0 + j;
        const float2 a0 = ld(q), ah = ld(q + H), al = ld(q + kL);
        inc[0] += (double)ah.x - a0.x; inc[1] += (double)ah.y - a0.y;
        inc[2] += (double)al.x - ah.x; inc[3] += (double)al.y - ah.y;
        inc[4] += (double)(q + kL) * al.x - (double)q * a0.x; inc[5] += (double)(q + kL) * al.y - (double)q * a0.y;
    }
    // block sum of `base`, block exclusive scan of `inc`
    double mine[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        mine[c] = inc[c];
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, inc[c], m);
            if (lane >= m) inc[c] += up;
            base[c] += __shfl_xor_sync(0xffffffffu, base[c], m);
        }
    }
    __syncthreads();                                                   // s_ramp may still be read by a previous call
    if (lane == 31) {
#pragma unroll
        for (int c = 0; c < 6; c++) s_ramp[0][warp][c] = inc[c];
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < 6; c++) s_ramp[1][warp][c] = base[c];
    }
    __syncthreads();
    double cur[6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        double v = inc[c] - mine[c];                                   // exclusive inside the warp
        for (int w = 0; w < kThreads / 32; w++) {
            if (w < warp) v += s_ramp[0][w][c];
            v += s_ramp[1][w][c];
        }
        cur[c] = v;
    }

    float best = 0.0f;
    int bidx = 0x7fffffff;
    const double beta = 1.0 / (4.0 * kL);
    for (int j = 0; j < run && q0 + j < n_lags; j++) {
        const long q = q0 + j;
        const double s0r = cur[0] + cur[2], s0i = cur[1] + cur[3];
        const double cr = 0.375 * cur[0] + 0.125 * cur[2] + beta * (cur[4] - (double)q * s0r);
        const double ci = 0.375 * cur[1] + 0.125 * cur[3] + beta * (cur[5] - (double)q * s0i);
        const float v = (float)(cr * cr + ci * ci);
        if (v > best) { best = v; bidx = (int)q; }
        const float2 a0 = ld(q), ah = ld(q + H), al = ld(q + kL);     // slide the window by one sample
        cur[0] += (double)ah.x - a0.x; cur[1] += (double)ah.y - a0.y;
        cur[2] += (double)al.x - ah.x; cur[3] += (double)al.y - ah.y;
        cur[4] += (double)(q + kL) * al.x - (double)q * a0.x; cur[5] += (double)(q + kL) * al.y - (double)q * a0.y;
    }
    block_argmax<kThreads>(best, bidx, s_val, s_idx);
    return best > 0.0f ? k_lo + bidx : k_lo;
}
template <int MOD, bool GUARD, int PHASE>
__global__ void __launch_bounds__(kThreads, 3) wide_acquire_kernel(const WideRxArgs a)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 768 : 1024;
    constexpr int BPS = BPC * D;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *s_lock = reinterpret_cast<float *>(smem_raw);
    uint8_t *u = reinterpret_cast<uint8_t *>(s_lock + kL);
    float2 *s_q = reinterpret_cast<float2 *>(u);                       // S&C phase
    float *s_e = reinterpret_cast<float *>(s_q + kAcqNQ + 1);
    float2 *bufA = reinterpret_cast<float2 *>(u), *bufB = bufA + kBufA;                     // FFT phase (aliases the S&C arrays)
    uint8_t *s_car = reinterpret_cast<uint8_t *>(bufB + kBufB);        // 1024 carrier bytes
    __shared__ float s_val[kThreads / 32];
    __shared__ int s_idx[kThreads / 32];
    __shared__ int s_d0;
    __shared__ double s_red[2 * (kThreads / 32)];
    __shared__ float s_redf[16];
    __shared__ float s_wtot[3 * (kThreads / 32)];

    const uint32_t stream = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int SYNC = a.sync_mode, CFO = a.cfo_mode;
    const bool FEC = a.fec != 0;
    StreamStateW *st = a.state + stream;
    const long M = (long)a.n_samples[stream];
    const float2 *x = a.iq + (a.stream_base ? (size_t)a.stream_base[stream] : (size_t)stream * a.iq_stride);
    for (int i = tid; i < kL; i += kThreads) s_lock[i] = a.tables->lock[i].x;
    if (tid == 0) s_d0 = 0x7fffffff;
    __syncthreads();

    long W = a.sync_window > 0 ? (long)a.sync_window : M;
    if (W > M) W = M;
    int status = ST_OK;
    long offset = 0;
    if (SYNC == 2) {
        offset = 0;                                    // frame start already known (ofdm_rx_decode_capture)
    } else if (SYNC == 0) {
        offset = (a.lock_is_ramp ? ramp_argmax_closed_w(x, M, -(kL - 1), W - 1, s_val, s_idx) : ramp_argmax_w(x, M, -(kL - 1), W - 1, s_lock, s_val, s_idx)) - 1;
    } else {
        long d_end = W;
        if (d_end > M - 2 * kL + 1) d_end = M - 2 * kL + 1;
        for (long base = 0; base < d_end; base += kAcqChunkW) {
            constexpr int RUN = (kAcqNE + kThreads - 1) / kThreads;           // 14 consecutive samples per thread
            float tqr = 0.0f, tqi = 0.0f, te = 0.0f;
            float qr[RUN], qi[RUN], ee[RUN];
#pragma unroll
            for (int r = 0; r < RUN; r++) {
                long n = base + (long)tid * RUN + r;
                float2 v0 = ld_sample(x, n, M), v1 = ld_sample(x, n + kL, M);
                float pr = v0.x * v1.x + v0.y * v1.y, pi = v0.x * v1.y - v0.y * v1.x;
                qr[r] = tqr; qi[r] = tqi; ee[r] = te;
                tqr += pr; tqi += pi; te += v0.x * v0.x + v0.y * v0.y;
            }
            float sqr = tqr, sqi = tqi, se = te;
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) {
                float a0 = __shfl_up_sync(0xffffffffu, sqr, m), a1 = __shfl_up_sync(0xffffffffu, sqi, m), a2 = __shfl_up_sync(0xffffffffu, se, m);
                if (lane >= m) { sqr += a0; sqi += a1; se += a2; }
            }
            __syncthreads();
            if (lane == 31) { s_wtot[warp] = sqr; s_wtot[8 + warp] = sqi; s_wtot[16 + warp] = se; }
            __syncthreads();
            float oqr = sqr - tqr, oqi = sqi - tqi, oe = se - te;
            for (int w2 = 0; w2 < warp; w2++) { oqr += s_wtot[w2]; oqi += s_wtot[8 + w2]; oe += s_wtot[16 + w2]; }
#pragma unroll
            for (int r = 0; r < RUN; r++) {
                int i = tid * RUN + r;
                if (i <= kAcqNQ) s_q[i] = make_float2(qr[r] + oqr, qi[r] + oqi);
                if (i <= kAcqNE) s_e[i] = ee[r] + oe;
            }
            __syncthreads();
            int found = 0x7fffffff;
            for (int i = tid; i < kAcqChunkW && base + i < d_end; i += kThreads) {
                float2 qa = s_q[i], qb = s_q[i + kL];
                float pr = qb.x - qa.x, pi = qb.y - qa.y;
                float r1 = s_e[i + kL] - s_e[i], r2 = s_e[i + 2 * kL] - s_e[i + kL];
                if (pr * pr + pi * pi > 0.5f * r1 * r2) { found = i; break; }
            }
            if (found != 0x7fffffff) atomicMin(&s_d0, (int)(base + found));
            __syncthreads();
            if (s_d0 != 0x7fffffff) break;
        }
        if (s_d0 == 0x7fffffff) {
            status = ST_NO_SYNC;
        } else {
            long d0 = s_d0, k_lo = d0 - (11 * kL) / 5, k_hi = d0 + kL / 5;
            if (k_lo < -(kL - 1)) k_lo = -(kL - 1);
            // The kernel is latency bound (one CTA per stream, a dozen dependent trips to DRAM). Whatever the refinement
            // decides, the CFO rows, the training rows and the header symbol lie between k_lo + 2 L and k_hi + 11 L: ask
            // for those lines now (L2 prefetch, no registers, no shared memory), while the ramp search runs.
            {
                long p0 = k_lo + 2 * kL, p1 = k_hi + 11 * kL;
                if (p0 < 0) p0 = 0;
                if (p1 > M) p1 = M;
                const char *b = reinterpret_cast<const char *>(x + p0), *e = reinterpret_cast<const char *>(x + p1);
                for (const char *q = b + 128 * (long)tid; q < e; q += 128 * kThreads) asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
            }
            offset = (a.lock_is_ramp ? ramp_argmax_closed_w(x, M, k_lo, k_hi, s_val, s_idx) : ramp_argmax_w(x, M, k_lo, k_hi, s_lock, s_val, s_idx)) - 1;
        }
    }
    if (status == ST_OK && offset < 0) status = ST_NEG_OFFSET;
    if (status == ST_OK && (offset > M || M - offset < kHeadW)) status = ST_TOO_SHORT;
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

    // ---- CFO estimate (f64) ----------------------------------------------------------------------------------------------
    double acc0 = 0.0, acc1 = 0.0;
    for (int i = tid; i < kL; i += kThreads) {
        float2 r2 = x0[2 * kL + i], r3 = x0[3 * kL + i], r4 = x0[4 * kL + i];
        if (CFO == 0) {
            double lr = r3.x, li = r3.y, rr = r4.x, ri = r4.y, nn = lr * lr + li * li;
            acc0 += atan2((ri * lr - rr * li) / nn, (rr * lr + ri * li) / nn);
        } else {
            acc0 += (double)r2.x * r3.x + (double)r2.y * r3.y + (double)r3.x * r4.x + (double)r3.y * r4.y;
            acc1 += (double)r2.x * r3.y - (double)r2.y * r3.x + (double)r3.x * r4.y - (double)r3.y * r4.x;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) { acc0 += __shfl_xor_sync(0xffffffffu, acc0, m); acc1 += __shfl_xor_sync(0xffffffffu, acc1, m); }
    if (lane == 0) { s_red[warp] = acc0; s_red[8 + warp] = acc1; }
    __syncthreads();
    double f_delta;
    {
        double t0 = 0.0, t1 = 0.0;
        for (int w2 = 0; w2 < kThreads / 32; w2++) { t0 += s_red[w2]; t1 += s_red[8 + w2]; }
        if (CFO == 0) f_delta = fabs((t0 / (double)kL) / (double)kL);
        else f_delta = atan2(t1, t0) / (double)kL;
    }
    const uint64_t fstep = (uint64_t)(int64_t)llrint(-f_delta * (0.15915494309189533577 * 18446744073709551616.0));
    if (tid == 0) { st->fstep = fstep; st->f_delta = (float)f_delta; st->offset = (int32_t)offset; }

    // ---- channel estimate: 5 training symbols ------------------------------------------------------------------------------
    FftTw T;
    fft1024_tw_init(T, a.tables->w1024, tid);
    __syncthreads();
    cpx hsum[4], wt[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        hsum[r] = c_make(0.0f, 0.0f);
        wt[r] = phasor_from_turns_p(fstep * (uint64_t)(5 * kL + kCpW + tid + 256 * r));      // derotation at this thread's samples of row 5
    }
    const cpx row_step = phasor_from_turns_p(fstep * (uint64_t)kL);                         // ... advanced by one symbol per row (4 steps: no drift)
#pragma unroll 1
    for (int row = 5; row < 10; row++) {
        cpx v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t n = row * kL + kCpW + tid + 256 * r;
            v[r] = c_mul(c_from(x0[n]), wt[r]);
            wt[r] = c_mul(wt[r], row_step);
        }
        fft1024_block(v, bufA, bufB, T, tid);
        const unsigned long long *A = reinterpret_cast<const unsigned long long *>(bufA);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            cpx xx; xx.v = A[padA(tid + 256 * r)];
            hsum[r] = c_add(hsum[r], c_mul(xx, c_from(a.tables->inv_training[tid + 256 * r])));
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        float hr, hi;
        c_split(c_scale(hsum[r], 0.2f), hr, hi);
        const float inv = 1.0f / (hr * hr + hi * hi);
        const int k = tid + 256 * r;
        st->h[k] = make_float2(hr, hi);
        st->g[k] = make_float2(hr * inv, -hi * inv);
        if (a.d_h) a.d_h[(size_t)stream * kN + k] = make_float2(hr, hi);
    }
    __threadfence_block();
    __syncthreads();

    // ---- header symbol (data symbol 0 holds >= 128 bits in every mode) --------------------------------------------------------
    WideLane L;
    wide_lane_init<GUARD>(L, st, tid);
    cpx z[4];
    wide_load_symbol(x0, (uint32_t)(n_avail > 0xffffffffL ? 0xffffffffL : n_avail), 0u, true, tid, z);
    const cpx base = phasor_from_turns_p(fstep * (uint64_t)(10 * kL + kCpW));
    wide_symbol<GUARD, PHASE>(L, base, z, bufA, bufB, T, s_redf, tid);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        float zr, zi;
        c_split(z[r], zr, zi);
        if (L.rank[r] >= 0) s_car[L.rank[r]] = (uint8_t)demap_point<MOD>(zr, zi);
    }
    __syncthreads();
    if (tid == 0) {
        uint64_t lo = 0, hi64 = 0;
        for (int b = 0; b < 128; b++) {
            int c = b / BPC, sh = b - c * BPC;
            uint64_t bit = (s_car[c] >> sh) & 1u;
            if (b < 64) lo |= bit << b; else hi64 |= bit << (b - 64);
        }
        const long rows = (n_avail + kL - 1) / kL;
        const long s_rx = rows - 10;
        const long avail_bytes = (s_rx * BPS) / 8 - 16;
        int stt = ST_OK;
        uint32_t n_syms = 0, out_len = 0;
        if (hi64 != 0 || avail_bytes < 0 || lo > (uint64_t)avail_bytes) {            // unsigned compare (see rx_acquire_kernel)
            stt = ST_BAD_HEADER;
        } else {
            const uint64_t nbits = kHeaderBits + 8 * lo;
            const uint64_t ncar = (nbits + BPC - 1) / BPC;
            n_syms = (uint32_t)((ncar + D - 1) / D);
            out_len = (uint32_t)(FEC ? (8 * lo) / 14 : lo);
            if (out_len > a.out_stride) { stt = ST_BAD_HEADER; n_syms = 0; out_len = 0; }
        }
        st->status = stt; st->n_syms = n_syms; st->out_len = out_len; st->plen = (uint32_t)lo; st->n_syms_rx = (uint32_t)s_rx;
        a.status[stream] = stt; a.out_len[stream] = out_len;
        if (a.d_offset) a.d_offset[stream] = (int32_t)offset;
        if (a.d_f_delta) a.d_f_delta[stream] = (float)f_delta;
        if (a.d_nsyms) a.d_nsyms[stream] = (uint32_t)s_rx;
    }
}

// ---- TX ------------------------------------------------------------------------------------------------------------------
struct WideTxArgs {
    const uint8_t  *payload;
    const uint32_t *payload_len;
    uint32_t        payload_stride;
    uint32_t        n_streams;
    float2         *iq;
    uint32_t        iq_stride;
    uint32_t       *frame_len;
    int            *stream_max;
    const WideTables *tables;
    uint32_t        stream0;        // first stream of this launch
    uint32_t       *stream_cnt;     // wide_tx_resident_kernel: per stream, warps that have published their maximum
    int32_t         group_ctas;     // wide_tx_resident_kernel: CTAs sharing one frame
    int32_t         n_groups;       // wide_tx_resident_kernel: groups of the (persistent) grid
    int32_t         redo_only;      // wide_tx_kernel<WRITE>: redo pass behind wide_tx_spec_kernel (frames whose stream_max is set, if stream_cnt[0] != 0)
    int32_t         redo_tiles;     // ... walking this many tiles of 8 symbols each
};

// 8 symbols of one frame: tile `bx` of `nbx` (the kernel's blockIdx.x / gridDim.x, or the redo loop's counter)
template <int MOD, bool GUARD, bool FEC, bool WRITE>
__device__ __forceinline__ void wide_tx_tile(const WideTxArgs &a, const uint32_t stream, const uint32_t bx, const uint32_t nbx)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 768 : 1024;
    constexpr int BPS = BPC * D;
    constexpr int TS = 8;                                            // symbols per CTA
    __shared__ __align__(16) float2 bufA[kBufA], bufB[kBufB];
    __shared__ __align__(16) uint8_t s_bits[TS * BPS / 8 + 16];
    __shared__ __align__(8) float2 s_map[64];
    __shared__ uint8_t s_enc[16];

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t n = a.payload_len[stream];
    const uint64_t coded_len = FEC ? (14ull * n + 7) / 8 : n;
    const uint64_t nbits = kHeaderBits + 8 * coded_len;
    const uint64_t ncar = (nbits + BPC - 1) / BPC;
    const int S = (int)((ncar + D - 1) / D);
    const uint32_t frame_len = (10u + (uint32_t)S) * kL;
    if (a.frame_len && bx == 0 && tid == 0) a.frame_len[stream] = frame_len;
    const bool fits = frame_len <= a.iq_stride;
    float2 *out = a.iq + (size_t)stream * a.iq_stride;
    int t0 = (int)bx * TS, t1 = t0 + TS;
    if (t1 > S) t1 = S;
    float scale = 1.0f / (float)kN;
    if (WRITE) {
        const float mx = fmaxf(__int_as_float(a.stream_max[stream]), a.tables->head_max);
        scale *= 1.0f / mx;
        if (bx == 0)
            for (uint32_t i = tid; i < (uint32_t)kHeadW && i < a.iq_stride; i += kThreads) {
                float2 v = make_float2(0.0f, 0.0f);
                if (fits) { v = a.tables->head[i]; v.x = v.x / mx; v.y = v.y / mx; }
                out[i] = v;
            }
        if (bx == nbx - 1) {
            const uint32_t z0 = fits ? frame_len : (uint32_t)kHeadW;
            for (uint32_t i = z0 + tid; i < a.iq_stride; i += kThreads) out[i] = make_float2(0.0f, 0.0f);
        }
    }
    if (!fits || t0 >= t1) return;
    if (tid < 64) {
        float re = 0.0f, im = 0.0f;
        if (MOD == 0) { re = (tid & 1) ? 1.0f : -1.0f; }
        else if (MOD == 1) { re = (tid & 1) ? 1.0f : -1.0f; im = (tid & 2) ? 1.0f : -1.0f; }
        else {
            const uint32_t ci = tid & 7u, cq = tid >> 3;
            const uint32_t li = ci ^ (ci >> 1) ^ (ci >> 2), lq = cq ^ (cq >> 1) ^ (cq >> 2);
            re = (2.0f * (float)li - 7.0f) * (1.0f / 7.0f);
            im = (2.0f * (float)lq - 7.0f) * (1.0f / 7.0f);
        }
        s_map[tid] = make_float2(im, re);                            // swapped: the IFFT runs as swap . FFT . swap
    }
    if (tid < 16) s_enc[tid] = (uint8_t)ham74_encode_nibble(tid);
    FftTw T;
    fft1024_tw_init(T, a.tables->w1024, tid);
    __syncthreads();
    const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
    const uint32_t byte0 = (uint32_t)((long)t0 * BPS / 8), nbyte = (uint32_t)((long)(t1 - t0) * BPS / 8);
    for (uint32_t b = tid; b < nbyte + 2; b += kThreads) s_bits[b] = (uint8_t)frame_byte<FEC>(pay, n, coded_len, byte0 + b, s_enc);
    __syncthreads();

    const long ncar_local = (long)ncar - (long)t0 * D;
    float mx = 0.0f;
#pragma unroll 1
    for (int s = t0; s < t1; s++) {
        cpx v[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int k = tid + 256 * r;
            cpx val = c_make(0.0f, 0.0f);
            const int rk = w_rank<GUARD>(k);
            if (GUARD && w_is_pilot(k)) val = c_make(0.0f, 1.0f);
            else if (rk >= 0) {
                const long c = (long)(s - t0) * D + rk;
                if (c < ncar_local) {
                    const uint32_t bit = (uint32_t)c * BPC, bb = bit >> 3;
                    const uint32_t w = ((uint32_t)s_bits[bb] | ((uint32_t)s_bits[bb + 1] << 8)) >> (bit & 7);
                    val = c_from(s_map[w & ((1u << BPC) - 1u)]);
                }
            }
            v[r] = val;
        }
        fft1024_block(v, bufA, bufB, T, tid);
        const unsigned long long *A = reinterpret_cast<const unsigned long long *>(bufA);
        float2 *sym = out + (size_t)(10 + s) * kL;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            cpx x; x.v = A[padA(tid + 256 * r)];
            float re, im;
            c_split(x, im, re);                                      // un-swap
            const int t = tid + 256 * r;
            if (WRITE) {
                const float2 o = make_float2(re * scale, im * scale);
                sym[kCpW + t] = o;
                if (r == 3) sym[t - (kN - kCpW)] = o;                // cyclic prefix = last 256 samples
            } else {
                mx = fmaxf(mx, fmaxf(re, im));
            }
        }
        __syncthreads();                                            // bufA is rewritten by the next symbol
    }
    if (!WRITE) {
        mx *= scale;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
        if (lane == 0 && mx > 0.0f) atomicMax(a.stream_max + stream, __float_as_int(mx));
    }
}

template <int MOD, bool GUARD, bool FEC, bool WRITE>
__global__ void __launch_bounds__(kThreads) wide_tx_kernel(const WideTxArgs a)
{
    if (WRITE && a.redo_only) {
        // redo pass behind wide_tx_spec_kernel: nothing to do unless the speculative kernel counted a frame whose data beat the head
        // maximum (stream_cnt[0]) -- then the CTAs stride over the frames and rewrite, tile by tile, those whose stream_max is set
        if (a.stream_cnt[0] == 0) return;
        for (uint32_t s = blockIdx.y; s < a.n_streams; s += gridDim.y) {
            if (a.stream_max[s] == 0) continue;
            for (uint32_t bx = 0; bx < (uint32_t)a.redo_tiles; bx++) {
                __syncthreads();
                wide_tx_tile<MOD, GUARD, FEC, WRITE>(a, s, bx, (uint32_t)a.redo_tiles);
            }
        }
        return;
    }
    const uint32_t stream = blockIdx.y + a.stream0;
    wide_tx_tile<MOD, GUARD, FEC, WRITE>(a, stream, blockIdx.x, gridDim.x);
}

}  // namespace wide
}  // namespace ofdm
