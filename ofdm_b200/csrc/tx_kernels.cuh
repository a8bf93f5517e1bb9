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
    const RxTables *tables;
};

// 14-bit Hamming word pair of payload byte b: low nibble codeword | high nibble codeword << 7 (docs/SPEC.md 3)
__device__ __forceinline__ uint32_t ham74_encode_byte(uint32_t b)
{
    return ham74_encode_nibble(b & 15u) | (ham74_encode_nibble(b >> 4) << 7);
}

// up to 8 bits [q, q+8) of the (optionally Hamming-coded) payload bit stream; zeros past the end
template <bool FEC>
__device__ __forceinline__ uint32_t payload_bits(const uint8_t *__restrict__ pay, uint32_t n, uint64_t q)
{
    if (!FEC) {
        uint64_t b = q >> 3;
        uint32_t v = (b < n ? pay[b] : 0u) | ((b + 1 < n ? pay[b + 1] : 0u) << 8);
        return (v >> (q & 7)) & 255u;
    }
    uint64_t m = q / 7;                 // codeword index; byte m>>1, nibble m&1
    uint32_t r = (uint32_t)(q - 7 * m);
    uint64_t b = m >> 1;
    uint64_t w = (b < n ? ham74_encode_byte(pay[b]) : 0u) | ((uint64_t)(b + 1 < n ? ham74_encode_byte(pay[b + 1]) : 0u) << 14);
    return (uint32_t)(w >> (7 * (m & 1) + r)) & 255u;
}

// `nb` (<= 8) bits of the frame bit stream [header 128 | payload...] starting at bit p (src/transmitter.rs:37-47)
template <bool FEC>
__device__ __forceinline__ uint32_t frame_bits(const uint8_t *__restrict__ pay, uint32_t n, uint64_t coded_len, uint64_t p, int nb)
{
    uint32_t v;
    if (p >= kHeaderBits) {
        v = payload_bits<FEC>(pay, n, p - kHeaderBits);
    } else {
        v = p < 64 ? (uint32_t)((coded_len >> p) & 255u) : 0u;      // bincode u128 LE: low 64 bits = length
        int nh = (int)(kHeaderBits - p);
        if (nh < nb) v = (v & ((1u << nh) - 1u)) | (payload_bits<FEC>(pay, n, 0) << nh);
    }
    return v & ((1u << nb) - 1u);
}

// One data OFDM symbol per 8-lane group: bits -> constellation -> IFFT -> CP. Writes un-normalised samples and
// tracks the per-stream maximum positive component for `normalize` (src/transmitter.rs:183-194).
template <int MOD, bool GUARD, bool FEC>
__global__ void __launch_bounds__(256) tx_symbols_kernel(const TxArgs a)
{
    constexpr int BPC = ModTraits<MOD>::kBpc;
    constexpr int D = GUARD ? 48 : 64;
    __shared__ __align__(16) float2 s_tr[8 * kTrWarp];

    const uint32_t stream = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 3, l = lane & 7;
    const uint32_t n = a.payload_len[stream];
    const uint64_t coded_len = FEC ? (14ull * n + 7) / 8 : n;
    const uint64_t nbits = kHeaderBits + 8 * coded_len;
    const uint64_t ncar = (nbits + BPC - 1) / BPC;                 // constellation symbols (src/transmitter.rs:108-140)
    const uint32_t S = (uint32_t)((ncar + D - 1) / D);             // OFDM data symbols (src/transmitter.rs:49-54)
    const uint32_t frame_len = (kHeadSyms + S) * kSym;
    if (a.frame_len && blockIdx.x == 0 && tid == 0) a.frame_len[stream] = frame_len;
    if (frame_len > a.iq_stride) return;                           // does not fit: host reports the error

    const uint8_t *pay = a.payload + (size_t)stream * a.payload_stride;
    float2 *out = a.iq + (size_t)stream * a.iq_stride;
    float twr[8], twi[8];
    fft64_lane_twiddles(a.tables->w64, l, twr, twi);
    float2 *tr = s_tr + warp * kTrWarp + g * kTrGroup;

    float mx = 0.0f;
    const uint32_t sym_per_block = 8 * 4;
    for (uint32_t s = blockIdx.x * sym_per_block + warp * 4 + g; s < ((S + 3) & ~3u) + 0; s += gridDim.x * sym_per_block) {
        const bool valid = s < S;
        float xr[8], xi[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = l + 8 * j;                               // encode_block, src/transmitter.rs:144-165
            float re = 0.0f, im = 0.0f;
            const int rk = data_rank<GUARD>(k);
            if (GUARD && is_pilot_bin(k)) { re = 1.0f; }
            else if (rk >= 0 && valid) {
                uint64_t c = (uint64_t)s * D + rk;
                if (c < ncar) {
                    uint32_t v = frame_bits<FEC>(pay, n, coded_len, c * BPC, BPC);
                    if (MOD == 0) { re = v ? 1.0f : -1.0f; }
                    else if (MOD == 1) { re = (v & 1) ? 1.0f : -1.0f; im = (v & 2) ? 1.0f : -1.0f; }
                    else {
                        // Gray code -> level index: i = c ^ (c>>1) ^ (c>>2) per axis, amplitude (2i-7)/7
                        uint32_t ci = v & 7u, cq = v >> 3;
                        uint32_t li = ci ^ (ci >> 1) ^ (ci >> 2), lq = cq ^ (cq >> 1) ^ (cq >> 2);
                        re = (2.0f * (float)li - 7.0f) * (1.0f / 7.0f);
                        im = (2.0f * (float)lq - 7.0f) * (1.0f / 7.0f);
                    }
                }
            }
            xr[j] = re; xi[j] = im;
        }
        // inverse FFT = swap(re, im) around the forward transform, scaled 1/64 (src/signals/mod.rs:49-58)
        fft64_group(xi, xr, twr, twi, tr, l);
        if (valid) {
            float2 *sym = out + (size_t)(kHeadSyms + s) * kSym;
#pragma unroll
            for (int kb = 0; kb < 8; kb++) {
                const int t = l + 8 * kb;                          // time index within the symbol
                float2 v = make_float2(xr[kb] * (1.0f / 64.0f), xi[kb] * (1.0f / 64.0f));
                mx = fmaxf(mx, fmaxf(v.x, v.y));
                sym[kCp + t] = v;                                  // prefix_block, src/transmitter.rs:168-181
                if (t >= kNfft - kCp) sym[t - (kNfft - kCp)] = v;
            }
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));
    if (lane == 0 && mx > 0.0f) atomicMax(a.stream_max + stream, __float_as_int(mx));
}

// normalize (src/transmitter.rs:183-194) + head (lock | preamble | training) + zero fill up to iq_stride
__global__ void __launch_bounds__(256) tx_finalize_kernel(const TxArgs a, const uint32_t *__restrict__ frame_len)
{
    const uint32_t stream = blockIdx.y;
    const uint32_t flen = frame_len[stream];
    float2 *out = a.iq + (size_t)stream * a.iq_stride;
    const bool fits = flen <= a.iq_stride;
    const float mx = fmaxf(__int_as_float(a.stream_max[stream]), a.tables->head_max);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.iq_stride; i += gridDim.x * blockDim.x) {
        float2 v = make_float2(0.0f, 0.0f);
        if (fits && i < flen) {
            v = i < kHeadSyms * kSym ? a.tables->head[i] : out[i];
            v.x = v.x / mx; v.y = v.y / mx;
        }
        out[i] = v;
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
    float          *accum;          // per stream: sum re, sum im, sum (y^2).re, sum (y^2).im, sum |y|^2
    float           snr_lin;
    float           cfo_max;
    uint32_t        lead_min, lead_max;
    uint32_t        multipath;
    uint32_t        noise_mode;
    uint32_t        seed_lo, seed_hi;
};

__device__ __forceinline__ void chan_stream_draws(const ChanArgs &a, uint32_t stream, uint32_t &lead, float &cfo)
{
    uint32_t c[4] = { stream, 0u, 0u, 0xC0FFEEu };
    philox4x32_10(c, a.seed_lo, a.seed_hi);
    uint32_t span = a.lead_max >= a.lead_min ? a.lead_max - a.lead_min + 1 : 1;
    lead = a.lead_min + (uint32_t)(((uint64_t)c[0] * span) >> 32);
    cfo = a.cfo_max >= 0.0f ? u01_from_u32(c[1]) * a.cfo_max : -1.0f;
}

__constant__ float kChanTaps[12] = { -0.0f, -0.1912f, 0.9316f, 0.2821f, -0.1990f, 0.1630f, -0.1017f, 0.0544f, -0.0261f, 0.0090f, 0.0f, -0.0034f };

// multipath (src/channel.rs:26-31,45) + CFO (src/channel.rs:54-62) + lead-in; accumulates the signal statistics
__global__ void __launch_bounds__(256) channel_conv_kernel(const ChanArgs a)
{
    const uint32_t stream = blockIdx.y;
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
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, m); s1 += __shfl_xor_sync(0xffffffffu, s1, m);
        s2 += __shfl_xor_sync(0xffffffffu, s2, m); s3 += __shfl_xor_sync(0xffffffffu, s3, m);
        s4 += __shfl_xor_sync(0xffffffffu, s4, m);
    }
    if ((threadIdx.x & 31) == 0) {
        float *acc = a.accum + 5 * (size_t)stream;
        atomicAdd(acc + 0, s0); atomicAdd(acc + 1, s1); atomicAdd(acc + 2, s2); atomicAdd(acc + 3, s3); atomicAdd(acc + 4, s4);
    }
}

// noise (src/channel.rs:66-71): mode 0 = reference-faithful (complex "variance", uniform draws); mode 1 = Gaussian AWGN.
// The lead-in carries noise too (a receiver never sees an exactly silent channel).
__global__ void __launch_bounds__(256) channel_noise_kernel(const ChanArgs a)
{
    const uint32_t stream = blockIdx.y;
    const uint32_t n_rx = a.rx_len[stream];
    if (n_rx == 0) return;
    uint32_t lead; float cfo;
    chan_stream_draws(a, stream, lead, cfo);
    const float *acc = a.accum + 5 * (size_t)stream;
    const float cnt = (float)(n_rx - lead);
    float ar, ai;                                         // complex noise amplitude
    if (a.noise_mode == 0) {
        // var = E[y^2] - mean^2 (no conjugate, src/signals/mod.rs:239-249); amp = sqrt(0.5 var / snr) (complex sqrt)
        float mr = acc[0] / cnt, mi = acc[1] / cnt;
        float vr = acc[2] / cnt - (mr * mr - mi * mi), vi = acc[3] / cnt - 2.0f * mr * mi;
        vr = 0.5f * vr / a.snr_lin; vi = 0.5f * vi / a.snr_lin;
        float r = sqrtf(sqrtf(vr * vr + vi * vi)), th = 0.5f * atan2f(vi, vr);
        ar = r * cosf(th); ai = r * sinf(th);
    } else {
        ar = sqrtf(0.5f * (acc[4] / cnt) / a.snr_lin); ai = 0.0f;
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
