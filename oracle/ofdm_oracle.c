/*
 * ofdm_oracle.c -- CPU restatement (f64) of the jkelleyrtp/ofdm modem hot path.
 * TEST INFRASTRUCTURE ONLY (see ofdm_oracle.h for who may load it and for the parity status).
 *
 * Every function cites the reference file:line under /root/reference it follows. Nothing here is
 * copied from the reference (which is Rust); it is a from-scratch C restatement of its arithmetic.
 * Stages with no reference (64QAM, Hamming, Schmidl-Cox, robust CFO/phase) follow docs/SPEC.md.
 */
#include "ofdm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NFFT 64
#define NCP 16
#define NSYM 80

/* ------------------------------------------------------------------------------------------------
 * complex helpers -- num::Complex64 formulas (naive mul/div, polar exp/sqrt)
 * ---------------------------------------------------------------------------------------------- */
static inline oo_c64 c_make(double re, double im) { oo_c64 z = { re, im }; return z; }
static inline oo_c64 c_add(oo_c64 a, oo_c64 b) { return c_make(a.re + b.re, a.im + b.im); }
static inline oo_c64 c_sub(oo_c64 a, oo_c64 b) { return c_make(a.re - b.re, a.im - b.im); }
static inline oo_c64 c_mul(oo_c64 a, oo_c64 b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline oo_c64 c_conj(oo_c64 a) { return c_make(a.re, -a.im); }
static inline double c_norm_sqr(oo_c64 a) { return a.re * a.re + a.im * a.im; }
/* num-complex Div: (a * conj(b)) / |b|^2 component-wise */
static inline oo_c64 c_div(oo_c64 a, oo_c64 b)
{
    double n = c_norm_sqr(b);
    return c_make((a.re * b.re + a.im * b.im) / n, (a.im * b.re - a.re * b.im) / n);
}
/* exp(j * theta) = from_polar(1, theta) */
static inline oo_c64 c_expj(double theta) { return c_make(cos(theta), sin(theta)); }
/* num-complex sqrt: polar form with the pure-real special case */
static oo_c64 c_sqrt(oo_c64 z)
{
    if (z.im == 0.0) {
        if (!signbit(z.re)) return c_make(sqrt(z.re), z.im);
        double im = sqrt(-z.re);
        return c_make(0.0, signbit(z.im) ? -im : im);
    }
    if (z.re == 0.0) {
        double x = sqrt(fabs(z.im) / 2.0);
        return c_make(x, signbit(z.im) ? -x : x);
    }
    double r = sqrt(c_norm_sqr(z)), th = atan2(z.im, z.re);
    double sr = sqrt(r);
    return c_make(sr * cos(th / 2.0), sr * sin(th / 2.0));
}

/* src/receiver.rs:242-246 */
double oo_angle(oo_c64 z) { return atan2(z.im, z.re); }

/* ------------------------------------------------------------------------------------------------
 * rand 0.8 StdRng restatement (third-party, absent from /root/reference: rand = "0.8.3",
 * StdRng = rand_chacha 0.3 ChaCha12Rng, seed_from_u64 = rand_core 0.6 PCG32 expansion).
 * Call sites: src/transmitter.rs:76,80,89,93. No reference test pins the values -> unverifiable here.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t key[8];
    uint64_t counter;
    uint32_t buf[16];
    int idx;
} stdrng;

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static inline uint32_t rotr32(uint32_t x, unsigned r) { r &= 31; return r ? ((x >> r) | (x << (32 - r))) : x; }

#define QR(a, b, c, d)                      \
    a += b; d ^= a; d = rotl32(d, 16);      \
    c += d; b ^= c; b = rotl32(b, 12);      \
    a += b; d ^= a; d = rotl32(d, 8);       \
    c += d; b ^= c; b = rotl32(b, 7);

static void chacha12_block(stdrng *g)
{
    uint32_t s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) s[4 + i] = g->key[i];
    s[12] = (uint32_t)g->counter; s[13] = (uint32_t)(g->counter >> 32);
    s[14] = 0; s[15] = 0;                      /* stream id 0 */
    memcpy(x, s, sizeof x);
    for (int r = 0; r < 6; r++) {              /* 12 rounds = 6 double rounds */
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) g->buf[i] = x[i] + s[i];
    g->counter++;
    g->idx = 0;
}

static void stdrng_seed_from_u64(stdrng *g, uint64_t state)
{
    for (int i = 0; i < 8; i++) {              /* PCG32 output per 4 seed bytes */
        state = state * 6364136223846793005ULL + 11634580027462260723ULL;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        g->key[i] = rotr32(xorshifted, rot);   /* to_le_bytes -> LE word */
    }
    g->counter = 0;
    g->idx = 16;
}

static uint32_t stdrng_next_u32(stdrng *g)
{
    if (g->idx >= 16) chacha12_block(g);
    return g->buf[g->idx++];
}

static uint64_t stdrng_next_u64(stdrng *g)
{
    uint64_t lo = stdrng_next_u32(g);
    uint64_t hi = stdrng_next_u32(g);
    return (hi << 32) | lo;
}

/* gen_range(-1.0..1.0): UniformFloat<f64>::sample_single */
static double stdrng_range_pm1(stdrng *g)
{
    for (;;) {
        uint64_t bits = (stdrng_next_u64(g) >> 12) | 0x3FF0000000000000ULL;
        double v12;
        memcpy(&v12, &bits, 8);
        double res = (v12 - 1.0) * 2.0 + (-1.0);
        if (res < 1.0) return res;
    }
}

/* known-answer hooks: the ChaCha12 block of a given key / block counter, and StdRng::from_seed(seed).next_u64() x n
 * (pinned on the published zero-key vector and on rand 0.8's own `test_stdrng_construction`, tests/test_oracle_kats.py) */
void oo_chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16])
{
    stdrng g;
    for (int i = 0; i < 8; i++) g.key[i] = key[i];
    g.counter = counter;
    chacha12_block(&g);
    for (int i = 0; i < 16; i++) out[i] = g.buf[i];
}

void oo_stdrng_from_seed_u64(const uint8_t seed[32], uint64_t *out, int n)
{
    stdrng g;
    for (int i = 0; i < 8; i++)
        g.key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    g.counter = 0;
    g.idx = 16;
    for (int i = 0; i < n; i++) out[i] = stdrng_next_u64(&g);
}

void oo_stdrng_uniform_pm1(uint64_t seed, double *out, int n)
{
    stdrng g;
    stdrng_seed_from_u64(&g, seed);
    for (int i = 0; i < n; i++) out[i] = stdrng_range_pm1(&g);
}

/* ------------------------------------------------------------------------------------------------
 * FFT -- stands in for rustfft 6 (third-party, semver range "6.0.0"; call sites
 * src/signals/mod.rs:42-44,50-52). Forward unscaled e^{-j2pi kn/N}; inverse scaled by 1/N
 * (src/signals/mod.rs:49-58). Plain DFT definition; agreement with rustfft is to rounding.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { size_t n; oo_c64 *tw; uint32_t *rev; } fft_plan;
static fft_plan g_plans[40];

static const fft_plan *plan_get(size_t n)
{
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    fft_plan *p = &g_plans[lg];
    if (p->n == n) return p;
#pragma omp critical(oo_fft_plan)
    {
        if (p->n != n) {
            oo_c64 *tw = (oo_c64 *)malloc(sizeof(oo_c64) * (n / 2 + 1));
            uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * n);
            for (size_t i = 0; i < n / 2; i++) {
                double a = -2.0 * M_PI * (double)i / (double)n;
                tw[i] = c_make(cos(a), sin(a));
            }
            for (size_t i = 0; i < n; i++) {
                uint32_t r = 0;
                for (int b = 0; b < lg; b++) if (i & ((size_t)1 << b)) r |= 1u << (lg - 1 - b);
                rev[i] = r;
            }
            p->tw = tw; p->rev = rev;
#pragma omp flush
            p->n = n;
        }
    }
    return p;
}

static void fft_pow2(oo_c64 *x, size_t n, int inverse)
{
    if (n <= 1) return;
    const fft_plan *p = plan_get(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = p->rev[i];
        if (j > i) { oo_c64 t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < half; k++) {
                oo_c64 w = p->tw[k * step];
                if (inverse) w.im = -w.im;
                oo_c64 u = x[i + k], v = c_mul(x[i + k + half], w);
                x[i + k] = c_add(u, v);
                x[i + k + half] = c_sub(u, v);
            }
        }
    }
}

/* Bluestein chirp-z for lengths that are not powers of two (the reference's 2M-1 transforms) */
static void fft_bluestein(oo_c64 *x, size_t n, int inverse)
{
    size_t m = 1;
    while (m < 2 * n - 1) m <<= 1;
    oo_c64 *w = (oo_c64 *)malloc(sizeof(oo_c64) * n);
    oo_c64 *a = (oo_c64 *)calloc(m, sizeof(oo_c64));
    oo_c64 *b = (oo_c64 *)calloc(m, sizeof(oo_c64));
    for (size_t k = 0; k < n; k++) {
        unsigned long long k2 = ((unsigned long long)k * k) % (2ULL * n);
        double ang = (inverse ? 1.0 : -1.0) * M_PI * (double)k2 / (double)n;
        w[k] = c_make(cos(ang), sin(ang));
    }
    for (size_t k = 0; k < n; k++) a[k] = c_mul(x[k], w[k]);
    b[0] = c_conj(w[0]);
    for (size_t k = 1; k < n; k++) b[k] = b[m - k] = c_conj(w[k]);
    fft_pow2(a, m, 0);
    fft_pow2(b, m, 0);
    for (size_t k = 0; k < m; k++) a[k] = c_mul(a[k], b[k]);
    fft_pow2(a, m, 1);
    for (size_t k = 0; k < n; k++) {
        oo_c64 v = c_make(a[k].re / (double)m, a[k].im / (double)m);
        x[k] = c_mul(v, w[k]);
    }
    free(w); free(a); free(b);
}

/* src/signals/mod.rs:27-58: fft() unscaled forward; ifft() inverse then normalize_by(1/len) */
void oo_fft(oo_c64 *x, size_t n, int inverse_scaled)
{
    if (n == 0) return;
    if ((n & (n - 1)) == 0) fft_pow2(x, n, inverse_scaled);
    else fft_bluestein(x, n, inverse_scaled);
    if (inverse_scaled) {
        double s = 1.0 / (double)n;
        for (size_t i = 0; i < n; i++) { x[i].re *= s; x[i].im *= s; }
    }
}

/* src/signals/mod.rs:61-77: mid = floor((len+1)/2); out = x[mid..] ++ x[..mid] */
void oo_fft_shift(oo_c64 *x, size_t n)
{
    size_t mid = (n + 1) / 2;
    oo_c64 *t = (oo_c64 *)malloc(sizeof(oo_c64) * n);
    memcpy(t, x, sizeof(oo_c64) * n);
    for (size_t i = 0; i < n; i++) x[i] = t[(i + mid) % n];
    free(t);
}

/* src/signals/mod.rs:80-95: mid = floor(len/2) */
void oo_ifft_shift(oo_c64 *x, size_t n)
{
    size_t mid = n / 2;
    oo_c64 *t = (oo_c64 *)malloc(sizeof(oo_c64) * n);
    memcpy(t, x, sizeof(oo_c64) * n);
    for (size_t i = 0; i < n; i++) x[i] = t[(i + mid) % n];
    free(t);
}

/* first strict maximum of |.|^2 starting from 0 (src/signals/mod.rs:205-214) */
static size_t argmax_norm_sqr(const oo_c64 *x, size_t n)
{
    double best = 0.0;
    size_t idx = 0;
    for (size_t i = 0; i < n; i++) {
        double v = c_norm_sqr(x[i]);
        if (v > best) { best = v; idx = i; }
    }
    return idx;
}

/*
 * src/signals/mod.rs:186-217: zero-pad both to P = 2*a_len-1, ifft(fft(a) * conj(fft(b))), fft_shift,
 * arg-max. With b_len <= a_len the circular correlation of length P holds every linear lag without
 * aliasing, so it is evaluated here on a power-of-two grid (same values to rounding, exact zeros at
 * lags < -(b_len-1)) and laid out exactly like the reference: lag k >= 0 at index a_len-1+k.
 */
size_t oo_xcorr_fft(const oo_c64 *a, size_t a_len, const oo_c64 *b, size_t b_len, oo_c64 *out)
{
    size_t P = 2 * a_len - 1;
    if (b_len > a_len) {                      /* general (tiny test) case: literal length-P transforms */
        oo_c64 *fa = (oo_c64 *)calloc(P > b_len ? P : b_len, sizeof(oo_c64));
        oo_c64 *fb = (oo_c64 *)calloc(P > b_len ? P : b_len, sizeof(oo_c64));
        size_t Q = P > b_len ? P : b_len;
        memcpy(fa, a, sizeof(oo_c64) * a_len);
        memcpy(fb, b, sizeof(oo_c64) * b_len);
        oo_fft(fa, Q, 0); oo_fft(fb, Q, 0);
        for (size_t i = 0; i < Q; i++) fa[i] = c_mul(fa[i], c_conj(fb[i]));
        oo_fft(fa, Q, 1);
        oo_fft_shift(fa, Q);
        memcpy(out, fa, sizeof(oo_c64) * P);
        free(fa); free(fb);
        return argmax_norm_sqr(out, P);
    }
    size_t m = 1;
    while (m < a_len + b_len) m <<= 1;
    oo_c64 *fa = (oo_c64 *)calloc(m, sizeof(oo_c64));
    oo_c64 *fb = (oo_c64 *)calloc(m, sizeof(oo_c64));
    memcpy(fa, a, sizeof(oo_c64) * a_len);
    memcpy(fb, b, sizeof(oo_c64) * b_len);
    fft_pow2(fa, m, 0); fft_pow2(fb, m, 0);
    for (size_t i = 0; i < m; i++) fa[i] = c_mul(fa[i], c_conj(fb[i]));
    fft_pow2(fa, m, 1);
    double s = 1.0 / (double)m;
    for (size_t i = 0; i < P; i++) out[i] = c_make(0.0, 0.0);
    for (size_t k = 0; k < a_len; k++)                       /* lags 0..a_len-1 */
        out[a_len - 1 + k] = c_make(fa[k].re * s, fa[k].im * s);
    for (size_t k = 1; k < b_len; k++)                       /* lags -1..-(b_len-1) */
        out[a_len - 1 - k] = c_make(fa[m - k].re * s, fa[m - k].im * s);
    free(fa); free(fb);
    return argmax_norm_sqr(out, P);
}

/* src/signals/mod.rs:219-237: ifft(fft(a,P) * fft(b,P)), P = a_len + b_len - 1 (direct form here) */
void oo_convolve(const oo_c64 *a, size_t a_len, const oo_c64 *b, size_t b_len, oo_c64 *out)
{
    size_t P = a_len + b_len - 1;
    for (size_t i = 0; i < P; i++) out[i] = c_make(0.0, 0.0);
    for (size_t j = 0; j < b_len; j++) {
        if (b[j].re == 0.0 && b[j].im == 0.0) continue;
        for (size_t i = 0; i < a_len; i++) out[i + j] = c_add(out[i + j], c_mul(a[i], b[j]));
    }
}

/* src/signals/mod.rs:251-259 */
oo_c64 oo_mean(const oo_c64 *x, size_t n)
{
    oo_c64 s = c_make(0.0, 0.0);
    for (size_t i = 0; i < n; i++) s = c_add(s, x[i]);
    s.re /= (double)n; s.im /= (double)n;
    return s;
}

/* src/signals/mod.rs:239-249: sum((mean - x)^2) / n, NO conjugate -> complex-valued */
oo_c64 oo_variance(const oo_c64 *x, size_t n)
{
    oo_c64 m = oo_mean(x, n), s = c_make(0.0, 0.0);
    for (size_t i = 0; i < n; i++) { oo_c64 d = c_sub(m, x[i]); s = c_add(s, c_mul(d, d)); }
    return c_make(s.re / (double)n, s.im / (double)n);
}

/* ------------------------------------------------------------------------------------------------
 * bits / BER / wire format -- src/utils.rs
 * ---------------------------------------------------------------------------------------------- */
/* src/utils.rs:21-27 (LSB first) */
void oo_to_bools(uint8_t byte, uint8_t out[8]) { for (int b = 0; b < 8; b++) out[b] = (byte >> b) & 1; }
/* src/utils.rs:30-36 */
uint8_t oo_bools_to_u8(const uint8_t b[8]) { uint8_t o = 0; for (int i = 0; i < 8; i++) o |= (uint8_t)((b[i] & 1) << i); return o; }

/* src/utils.rs:45-68 */
void oo_analysis(const uint8_t *l, const uint8_t *r, size_t n, uint32_t *num_errs, uint32_t *num_block_errs, double *err_rate)
{
    uint32_t e = 0, be = 0;
    for (size_t i = 0; i < n; i++) if (l[i] != r[i]) { e += (uint32_t)__builtin_popcount((unsigned)(l[i] ^ r[i])); be++; }
    *num_errs = e; *num_block_errs = be; *err_rate = (double)e / ((double)n * 8.0);
}

/* src/utils.rs:228-236 */
void oo_sig_to_fc32(const oo_c64 *x, size_t n, float *out) { for (size_t i = 0; i < n; i++) { out[2 * i] = (float)x[i].re; out[2 * i + 1] = (float)x[i].im; } }
/* src/utils.rs:239-254 */
void oo_fc32_to_sig(const float *in, size_t n, oo_c64 *out) { for (size_t i = 0; i < n; i++) out[i] = c_make((double)in[2 * i], (double)in[2 * i + 1]); }

/* ------------------------------------------------------------------------------------------------
 * Hamming(7,4) -- docs/SPEC.md section 3 (no reference implementation)
 * ---------------------------------------------------------------------------------------------- */
static inline unsigned ham_enc_nibble(unsigned d)
{
    unsigned d1 = d & 1, d2 = (d >> 1) & 1, d3 = (d >> 2) & 1, d4 = (d >> 3) & 1;
    unsigned p1 = d1 ^ d2 ^ d4, p2 = d1 ^ d3 ^ d4, p3 = d2 ^ d3 ^ d4;
    return p1 | (p2 << 1) | (d1 << 2) | (p3 << 3) | (d2 << 4) | (d3 << 5) | (d4 << 6);
}
static inline unsigned ham_dec_word(unsigned c)
{
    unsigned s1 = (c ^ (c >> 2) ^ (c >> 4) ^ (c >> 6)) & 1;
    unsigned s2 = ((c >> 1) ^ (c >> 2) ^ (c >> 5) ^ (c >> 6)) & 1;
    unsigned s3 = ((c >> 3) ^ (c >> 4) ^ (c >> 5) ^ (c >> 6)) & 1;
    unsigned s = s1 | (s2 << 1) | (s3 << 2);
    if (s) c ^= 1u << (s - 1);
    return ((c >> 2) & 1) | (((c >> 4) & 1) << 1) | (((c >> 5) & 1) << 2) | (((c >> 6) & 1) << 3);
}
size_t oo_hamming74_encoded_len(size_t n) { return (14 * n + 7) / 8; }
size_t oo_hamming74_decoded_len(size_t n_coded) { return (8 * n_coded) / 14; }
void oo_hamming74_encode(const uint8_t *in, size_t n, uint8_t *out)
{
    size_t nout = oo_hamming74_encoded_len(n);
    memset(out, 0, nout);
    size_t bit = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned w = ham_enc_nibble(in[i] & 15u) | (ham_enc_nibble(in[i] >> 4) << 7);   /* 14 bits */
        for (int b = 0; b < 14; b++, bit++) if ((w >> b) & 1) out[bit >> 3] |= (uint8_t)(1u << (bit & 7));
    }
}
void oo_hamming74_decode(const uint8_t *in, size_t n_coded, uint8_t *out)
{
    size_t n = oo_hamming74_decoded_len(n_coded);
    size_t bit = 0;
    for (size_t i = 0; i < n; i++) {
        unsigned w = 0;
        for (int b = 0; b < 14; b++, bit++) w |= (unsigned)((in[bit >> 3] >> (bit & 7)) & 1) << b;
        out[i] = (uint8_t)(ham_dec_word(w & 127u) | (ham_dec_word(w >> 7) << 4));
    }
}

/* ------------------------------------------------------------------------------------------------
 * tables -- src/transmitter.rs:60-96
 * ---------------------------------------------------------------------------------------------- */
/* src/transmitter.rs:60-72 */
void oo_locking_signal(oo_c64 *out, int len)
{
    for (int i = 0; i < len; i++) out[i] = c_make(0.5 * ((double)i / (2.0 * (double)len) + 0.5), 0.0);
    oo_fft_shift(out, (size_t)len);
}
/* src/transmitter.rs:75-84 */
void oo_preamble(oo_c64 *out, int len)
{
    stdrng g;
    stdrng_seed_from_u64(&g, 100);
    for (int i = 0; i < len; i++) {
        double re = stdrng_range_pm1(&g), im = stdrng_range_pm1(&g);
        out[i] = c_make(re * 0.25, im * 0.25);
    }
}
/* src/transmitter.rs:88-96 */
void oo_training_signals(oo_c64 *out, int len)
{
    stdrng g;
    stdrng_seed_from_u64(&g, 50);
    for (int i = 0; i < len; i++) {
        double re = stdrng_range_pm1(&g), im = stdrng_range_pm1(&g);
        out[i] = c_make(re * 1.0, im * 1.0);
    }
}

/* carrier classes, src/transmitter.rs:150-161 / src/receiver.rs:119-134 */
static inline int is_null_bin(int i) { return i >= 59 || i <= 5 || i == 32; }
static inline int is_pilot_bin(int i) { return i == 6 || i == 25 || i == 39 || i == 58; }

/* layout: N = 64 is the reference's (above); N = 1024 is the wideband variant of docs/SPEC.md section 9 (no reference):
 * nulls {0..95, 512, 929..1023}, of the remaining 832 bins (ascending) every 13th starting with bin 96 is a pilot
 * (64 pilots, 1+0j), the other 768 carry data. */
#define NMAX 1024
static inline int lay_nfft(const oo_cfg *cfg) { return cfg->nfft == 1024 ? 1024 : 64; }
static inline int lay_is_null(int nf, int k) { return nf == 64 ? is_null_bin(k) : (k <= 95 || k == 512 || k >= 929); }
static inline int lay_is_pilot(int nf, int k)
{
    if (nf == 64) return is_pilot_bin(k);
    if (lay_is_null(nf, k)) return 0;
    int r = k < 512 ? k - 96 : k - 97;
    return r % 13 == 0;
}
static inline int lay_data_carriers(int nf, int guard) { return guard ? (nf == 64 ? 48 : 768) : nf; }
static inline int lay_pilots(int nf) { return nf == 64 ? 4 : 64; }

/* ------------------------------------------------------------------------------------------------
 * TX -- src/transmitter.rs
 * ---------------------------------------------------------------------------------------------- */
static const double QAM64_LEVEL_OF_CODE[8] = { -7.0, -5.0, -1.0, -3.0, 7.0, 5.0, 1.0, 3.0 };  /* SPEC section 2 */

/* src/transmitter.rs:108-140 (Bpsk, Qpsk); Qam arm per docs/SPEC.md section 2 */
size_t oo_modulate(const uint8_t *bytes, size_t n, int scheme, oo_c64 *out)
{
    size_t k = 0;
    if (scheme == OO_BPSK) {
        for (size_t i = 0; i < n; i++)
            for (int b = 0; b < 8; b++) out[k++] = c_make(((bytes[i] >> b) & 1) ? 1.0 : -1.0, 0.0);
    } else if (scheme == OO_QPSK) {
        for (size_t i = 0; i < n; i++)
            for (int b = 0; b < 8; b += 2)
                out[k++] = c_make(((bytes[i] >> b) & 1) ? 1.0 : -1.0, ((bytes[i] >> (b + 1)) & 1) ? 1.0 : -1.0);
    } else {
        size_t nbits = 8 * n, nsym = (nbits + 5) / 6;
        for (size_t s = 0; s < nsym; s++) {
            unsigned v = 0;
            for (int b = 0; b < 6; b++) {
                size_t bit = 6 * s + (size_t)b;
                if (bit < nbits) v |= (unsigned)((bytes[bit >> 3] >> (bit & 7)) & 1) << b;
            }
            out[k++] = c_make(QAM64_LEVEL_OF_CODE[v & 7] / 7.0, QAM64_LEVEL_OF_CODE[v >> 3] / 7.0);
        }
    }
    return k;
}

static inline unsigned qam64_axis_code(double v)
{
    double f = floor(3.5 * v + 4.0);
    int i = f < 0.0 ? 0 : (f > 7.0 ? 7 : (int)f);
    return (unsigned)(i ^ (i >> 1));
}

/* src/receiver.rs:147-190; Qam arm per docs/SPEC.md */
size_t oo_demodulate(const oo_c64 *syms, size_t n, int scheme, uint8_t *out)
{
    size_t k = 0;
    if (n % 8 != 0) return (size_t)-1;                       /* assert_eq!(remainder.len(), 0) */
    for (size_t g = 0; g < n; g += 8) {
        const oo_c64 *c = syms + g;
        if (scheme == OO_BPSK) {
            uint8_t v = 0;
            for (int i = 0; i < 8; i++) if (c[i].re > 0.0) v |= (uint8_t)(1u << i);
            out[k++] = v;
        } else if (scheme == OO_QPSK) {
            unsigned v = 0;
            for (int i = 0; i < 8; i++) {
                double re = c[i].re, im = c[i].im;
                int l, r;
                if (re >= 0.0 && im >= 0.0) { l = 1; r = 1; }
                else if (re >= 0.0 && im <= 0.0) { l = 1; r = 0; }
                else if (re < 0.0 && im > 0.0) { l = 0; r = 1; }
                else { l = 0; r = 0; }
                v |= (unsigned)l << (2 * i);
                v |= (unsigned)r << (2 * i + 1);
            }
            out[k++] = (uint8_t)(v & 255u);
            out[k++] = (uint8_t)(v >> 8);
        } else {
            uint64_t v = 0;
            for (int i = 0; i < 8; i++) {
                uint64_t s6 = qam64_axis_code(c[i].re) | (qam64_axis_code(c[i].im) << 3);
                v |= s6 << (6 * i);
            }
            for (int b = 0; b < 6; b++) out[k++] = (uint8_t)(v >> (8 * b));
        }
    }
    return k;
}

static size_t n_symbols_for_bytes(size_t n_bytes_with_header, int scheme)
{
    if (scheme == OO_BPSK) return 8 * n_bytes_with_header;
    if (scheme == OO_QPSK) return 4 * n_bytes_with_header;
    return (8 * n_bytes_with_header + 5) / 6;
}

size_t oo_frame_data_syms_n(size_t n_bytes, int guard_bands, int scheme, int nfft)
{
    size_t d = (size_t)lay_data_carriers(nfft, guard_bands);
    size_t ns = n_symbols_for_bytes(n_bytes + 16, scheme);
    return (ns + d - 1) / d;
}
size_t oo_frame_data_syms(size_t n_bytes, int guard_bands, int scheme) { return oo_frame_data_syms_n(n_bytes, guard_bands, scheme, 64); }
size_t oo_frame_len_n(size_t n_bytes, int guard_bands, int scheme, int nfft) { return (10 + oo_frame_data_syms_n(n_bytes, guard_bands, scheme, nfft)) * (size_t)(nfft + nfft / 4); }
size_t oo_frame_len(size_t n_bytes, int guard_bands, int scheme) { return oo_frame_len_n(n_bytes, guard_bands, scheme, 64); }

/* src/transmitter.rs:168-181: in-place scaled IFFT then out = x[N-CP..N] ++ x[0..N] */
static void prefix_block(oo_c64 *freq /* nf, clobbered */, oo_c64 *out /* nf + nf/4 */, int nf)
{
    int cp = nf / 4;
    oo_fft(freq, (size_t)nf, 1);
    memcpy(out, freq + (nf - cp), sizeof(oo_c64) * (size_t)cp);
    memcpy(out + cp, freq, sizeof(oo_c64) * (size_t)nf);
}

/* src/transmitter.rs:11-58 (nfft = 64); the same construction scaled to nfft = 1024 (docs/SPEC.md section 9) */
size_t oo_encode_n(const uint8_t *data, size_t n, int guard_bands, int scheme, int nfft, oo_c64 *out)
{
    const int NF = nfft == 1024 ? 1024 : 64, LS = NF + NF / 4;
    size_t pos = 0;
    oo_locking_signal(out + pos, LS); pos += LS;                            /* :22-24 */
    for (int i = 0; i < 4; i++) { oo_preamble(out + pos, LS); pos += LS; }       /* :27-29 */
    for (int i = 0; i < 5; i++) {                                           /* :32-34 */
        oo_c64 t[NMAX];
        oo_training_signals(t, NF);
        prefix_block(t, out + pos, NF); pos += LS;
    }
    /* :37-47 header (bincode u128 LE, src/packets/mod.rs:20-32) ++ data, modulated as one stream */
    uint8_t *bytes = (uint8_t *)calloc(n + 16, 1);
    uint64_t len64 = (uint64_t)n;
    for (int i = 0; i < 8; i++) bytes[i] = (uint8_t)(len64 >> (8 * i));
    if (n) memcpy(bytes + 16, data, n);
    size_t nsym_cap = n_symbols_for_bytes(n + 16, scheme);
    oo_c64 *syms = (oo_c64 *)malloc(sizeof(oo_c64) * (nsym_cap + 1));
    size_t nsym = oo_modulate(bytes, n + 16, scheme, syms);
    free(bytes);
    /* :49-54 blocks; encode_block :144-165 */
    size_t k = 0;
    while (k < nsym) {
        oo_c64 blk[NMAX];
        for (int i = 0; i < NF; i++) {
            if (guard_bands && lay_is_null(NF, i)) blk[i] = c_make(0.0, 0.0);
            else if (guard_bands && lay_is_pilot(NF, i)) blk[i] = c_make(1.0, 0.0);
            else blk[i] = (k < nsym) ? syms[k++] : c_make(0.0, 0.0);
        }
        prefix_block(blk, out + pos, NF); pos += LS;
    }
    free(syms);
    /* normalize :183-194: max over signed re/im values starting from 0 */
    double mx = 0.0;
    for (size_t i = 0; i < pos; i++) { mx = fmax(out[i].re, mx); mx = fmax(out[i].im, mx); }
    for (size_t i = 0; i < pos; i++) { out[i].re = out[i].re / mx; out[i].im = out[i].im / mx; }
    return pos;
}

size_t oo_encode(const uint8_t *data, size_t n, int guard_bands, int scheme, oo_c64 *out) { return oo_encode_n(data, n, guard_bands, scheme, 64, out); }

size_t oo_tx_len(size_t n_payload, const oo_cfg *cfg)
{
    size_t n = cfg->fec ? oo_hamming74_encoded_len(n_payload) : n_payload;
    return oo_frame_len_n(n, cfg->guard_bands, cfg->modulation, lay_nfft(cfg));
}

size_t oo_tx(const uint8_t *payload, size_t n, const oo_cfg *cfg, oo_c64 *out)
{
    if (!cfg->fec) return oo_encode_n(payload, n, cfg->guard_bands, cfg->modulation, lay_nfft(cfg), out);
    size_t nc = oo_hamming74_encoded_len(n);
    uint8_t *coded = (uint8_t *)malloc(nc + 1);
    oo_hamming74_encode(payload, n, coded);
    size_t r = oo_encode_n(coded, nc, cfg->guard_bands, cfg->modulation, lay_nfft(cfg), out);
    free(coded);
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * channel -- src/channel.rs:26-74, with a seeded generator instead of thread_rng
 * ---------------------------------------------------------------------------------------------- */
static const double CHANNEL_TAPS[12] = { -0.0, -0.1912, 0.9316, 0.2821, -0.1990, 0.1630, -0.1017, 0.0544, -0.0261, 0.0090, 0.0, -0.0034 };

typedef struct { uint64_t s[4]; } xo256;
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix64(uint64_t *x) { uint64_t z = (*x += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
static void xo_seed(xo256 *g, uint64_t seed) { for (int i = 0; i < 4; i++) g->s[i] = splitmix64(&seed); }
static uint64_t xo_next(xo256 *g)
{
    uint64_t *s = g->s, r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static double xo_u01(xo256 *g) { return (double)(xo_next(g) >> 11) * (1.0 / 9007199254740992.0); }

void oo_channel(const oo_c64 *tx, size_t n, double snr_db, double f_delta, int noise_mode, uint64_t seed, oo_c64 *out)
{
    size_t P = n + 63;
    double snr = pow(10.0, snr_db / 10.0);                                  /* :40 */
    oo_c64 h[64];
    for (int i = 0; i < 64; i++) h[i] = c_make(0.0, 0.0);
    for (int i = 0; i < 12; i++) h[7 + i] = c_make(CHANNEL_TAPS[i], 0.0);   /* :26-31 */
    oo_convolve(tx, n, h, 64, out);                                         /* :45 */
    if (f_delta >= 0.0)                                                     /* :54-62, 1-based index */
        for (size_t i = 0; i < P; i++) out[i] = c_mul(out[i], c_expj(f_delta * (double)(i + 1)));
    xo256 g;
    xo_seed(&g, seed);
    if (noise_mode == 0) {                                                  /* :66-71 */
        oo_c64 var = oo_variance(out, P);
        oo_c64 nv = c_make(var.re / snr, var.im / snr);
        oo_c64 amp = c_sqrt(c_make(0.5 * nv.re, 0.5 * nv.im));
        for (size_t i = 0; i < P; i++) {
            double a = 2.0 * xo_u01(&g) - 1.0, b = 2.0 * xo_u01(&g) - 1.0;
            out[i] = c_add(out[i], c_mul(amp, c_make(a, b)));
        }
    } else {                                                                /* proper complex AWGN */
        double pw = 0.0;
        for (size_t i = 0; i < P; i++) pw += c_norm_sqr(out[i]);
        pw /= (double)P;
        double sigma = sqrt(0.5 * pw / snr);
        for (size_t i = 0; i < P; i++) {
            double u1 = xo_u01(&g), u2 = xo_u01(&g);
            if (u1 < 1e-300) u1 = 1e-300;
            double r = sqrt(-2.0 * log(u1));
            out[i] = c_add(out[i], c_make(sigma * r * cos(2.0 * M_PI * u2), sigma * r * sin(2.0 * M_PI * u2)));
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * RX -- src/receiver.rs
 * ---------------------------------------------------------------------------------------------- */
static inline oo_c64 sample_or_zero(const oo_c64 *a, size_t n, long i)
{
    return (i >= 0 && (size_t)i < n) ? a[i] : c_make(0.0, 0.0);
}

/* c[k] = sum_{n<L} a[n+k] b[n] over k in [k_lo, k_hi]; first strict max of |c|^2 (equivalent direct form of
 * src/signals/mod.rs:186-217 restricted to a lag window) */
static long ramp_argmax_n(const oo_c64 *a, size_t n, long k_lo, long k_hi, const oo_c64 *lock, int LS)
{
    double best = 0.0;
    long kbest = k_lo;
    for (long k = k_lo; k <= k_hi; k++) {
        oo_c64 c = c_make(0.0, 0.0);
        for (int j = 0; j < LS; j++) c = c_add(c, c_mul(sample_or_zero(a, n, k + j), c_conj(lock[j])));
        double v = c_norm_sqr(c);
        if (v > best) { best = v; kbest = k; }
    }
    return kbest;
}

static int find_offset(const oo_c64 *a, size_t n, const oo_cfg *cfg, long *offset)
{
    const int NF = lay_nfft(cfg), LS = NF + NF / 4;
    oo_c64 lock[NMAX + NMAX / 4];
    oo_locking_signal(lock, LS);
    long W = cfg->sync_window > 0 ? (long)cfg->sync_window : (long)n;
    if (W > (long)n) W = (long)n;
    if (cfg->sync_mode == OO_SYNC_REFERENCE) {
        if (cfg->xcorr_fft && cfg->sync_window <= 0) {
            /* src/receiver.rs:20-21 */
            oo_c64 *cross = (oo_c64 *)malloc(sizeof(oo_c64) * (2 * n - 1));
            size_t idxmax = oo_xcorr_fft(a, n, lock, (size_t)LS, cross);
            free(cross);
            *offset = (long)idxmax - (long)(((2 * n - 1) - 1) / 2 + 1);
        } else {
            *offset = ramp_argmax_n(a, n, -(LS - 1), W - 1, lock, LS) - 1;
        }
        return OO_OK;
    }
    /* Schmidl-Cox, docs/SPEC.md section 4 (window and lag = one preamble period L) */
    long d0 = -1;
    oo_c64 P = c_make(0.0, 0.0);
    double R1 = 0.0, R2 = 0.0;
    for (long d = 0; d < W && (size_t)(d + 2 * LS) <= n; d++) {
        if (d == 0 || (d & 1023) == 0) {           /* exact re-sum periodically: no drift */
            P = c_make(0.0, 0.0); R1 = 0.0; R2 = 0.0;
            for (int m = 0; m < LS; m++) {
                P = c_add(P, c_mul(c_conj(a[d + m]), a[d + m + LS]));
                R1 += c_norm_sqr(a[d + m]);
                R2 += c_norm_sqr(a[d + m + LS]);
            }
        } else {
            P = c_sub(P, c_mul(c_conj(a[d - 1]), a[d - 1 + LS]));
            P = c_add(P, c_mul(c_conj(a[d - 1 + LS]), a[d - 1 + 2 * LS]));
            R1 += c_norm_sqr(a[d - 1 + LS]) - c_norm_sqr(a[d - 1]);
            R2 += c_norm_sqr(a[d - 1 + 2 * LS]) - c_norm_sqr(a[d - 1 + LS]);
        }
        if (c_norm_sqr(P) > 0.5 * R1 * R2) { d0 = d; break; }
    }
    if (d0 < 0) return OO_NO_SYNC;
    long k_lo = d0 - (11 * LS) / 5, k_hi = d0 + LS / 5;            /* 176 / 16 at L = 80 */
    if (k_lo < -(LS - 1)) k_lo = -(LS - 1);
    *offset = ramp_argmax_n(a, n, k_lo, k_hi, lock, LS) - 1;
    return OO_OK;
}

/* src/receiver.rs:99-104 */
static void unprefix_block(const oo_c64 *row /* L */, oo_c64 *out /* N */, int NF)
{
    memcpy(out, row + NF / 4, sizeof(oo_c64) * (size_t)NF);
    oo_fft(out, (size_t)NF, 0);
}

int oo_decode(const oo_c64 *samples, size_t n, const oo_cfg *cfg,
              uint8_t *out, size_t out_cap, size_t *out_len,
              oo_c64 *points, size_t points_cap, oo_diag *diag)
{
    const int NF = lay_nfft(cfg), LS = NF + NF / 4;
    oo_diag local;
    if (!diag) { diag = &local; local.h_full = NULL; }
    oo_c64 *h_full = diag->h_full;
    memset(diag, 0, sizeof *diag);
    diag->h_full = h_full;
    *out_len = 0;

    long offset = 0;
    int st = find_offset(samples, n, cfg, &offset);                          /* :20-21 */
    diag->offset = (int32_t)offset;
    if (st != OO_OK) { diag->status = st; return st; }
    if (offset < 0) { diag->status = OO_NEG_OFFSET; return OO_NEG_OFFSET; }  /* :25 panics in the reference */
    if ((size_t)offset > n || n - (size_t)offset < (size_t)(10 * LS)) { diag->status = OO_TOO_SHORT; return OO_TOO_SHORT; } /* :27-29 */

    size_t len = n - (size_t)offset;
    size_t rows = (len + LS - 1) / LS;                                   /* :36, :192-210 zero-padded tail row */
    oo_c64 *x = (oo_c64 *)calloc(rows * LS, sizeof(oo_c64));
    memcpy(x, samples + offset, sizeof(oo_c64) * len);

    /* :39, :231-240 */
    double f_delta;
    if (cfg->cfo_mode == OO_CFO_REFERENCE) {
        double acc = 0.0;
        for (int i = 0; i < LS; i++) acc += oo_angle(c_div(x[4 * LS + i], x[3 * LS + i]));
        f_delta = fabs((acc / (double)LS) / (double)LS);
    } else {
        oo_c64 s = c_make(0.0, 0.0);
        for (int i = 0; i < LS; i++) {
            s = c_add(s, c_mul(c_conj(x[2 * LS + i]), x[3 * LS + i]));
            s = c_add(s, c_mul(c_conj(x[3 * LS + i]), x[4 * LS + i]));
        }
        f_delta = oo_angle(s) / (double)LS;
    }
    diag->f_delta = f_delta;

    /* :44-50 */
    for (size_t i = 0; i < rows * LS; i++) x[i] = c_mul(x[i], c_expj(-f_delta * (double)i));

    /* :56, :212-229 */
    oo_c64 hk[NMAX], training[NMAX];
    oo_training_signals(training, NF);
    for (int i = 0; i < NF; i++) hk[i] = c_make(0.0, 0.0);
    for (int b = 5; b < 10; b++) {
        oo_c64 blk[NMAX];
        unprefix_block(x + (size_t)b * LS, blk, NF);
        for (int i = 0; i < NF; i++) hk[i] = c_add(hk[i], c_div(blk[i], training[i]));
    }
    for (int i = 0; i < NF; i++) { hk[i].re /= 5.0; hk[i].im /= 5.0; }
    memcpy(diag->h_k, hk, sizeof(oo_c64) * 64);            /* diag carries the first 64 bins */
    if (diag->h_full) memcpy(diag->h_full, hk, sizeof(oo_c64) * (size_t)NF);

    /* :63-74 */
    size_t S = rows - 10, D = (size_t)lay_data_carriers(NF, cfg->guard_bands);
    oo_c64 *stream = (oo_c64 *)malloc(sizeof(oo_c64) * (S * D + 8));
    size_t np = 0;
    for (size_t s = 0; s < S; s++) {
        oo_c64 Y[NMAX];
        unprefix_block(x + (10 + s) * LS, Y, NF);
        for (int i = 0; i < NF; i++) Y[i] = c_div(Y[i], hk[i]);            /* :68-70 */
        /* decode_block :106-145 */
        double phase = 0.0;
        oo_c64 psum = c_make(0.0, 0.0);
        size_t first = np;
        for (int i = 0; i < NF; i++) {
            if (cfg->guard_bands && lay_is_null(NF, i)) continue;
            if (cfg->guard_bands && lay_is_pilot(NF, i)) {
                phase += oo_angle(c_div(Y[i], c_make(1.0, 0.0)));            /* :126 */
                psum = c_add(psum, Y[i]);
                continue;
            }
            stream[np++] = Y[i];
        }
        phase /= (double)lay_pilots(NF);                                     /* :137 (pilot_count = 4.0 at N = 64) */
        if (cfg->guard_bands && cfg->phase_mode == OO_PHASE_ANGLE_OF_SUM) phase = oo_angle(psum);
        oo_c64 rot = c_expj(-phase);                                         /* :140-144 */
        for (size_t i = first; i < np; i++) stream[i] = c_mul(stream[i], rot);
    }
    free(x);
    diag->n_data_syms = (int64_t)S;
    diag->n_points = (int64_t)np;
    if (points) memcpy(points, stream, sizeof(oo_c64) * (np < points_cap ? np : points_cap));

    /* :83 */
    size_t bps = cfg->modulation == OO_BPSK ? 1 : (cfg->modulation == OO_QPSK ? 2 : 6);
    uint8_t *bytes = (uint8_t *)malloc(np * bps / 8 + 8);
    size_t nb = oo_demodulate(stream, np, cfg->modulation, bytes);
    free(stream);

    /* :86-93 */
    if (nb == (size_t)-1 || nb < 16) { free(bytes); diag->status = OO_BAD_HEADER; return OO_BAD_HEADER; }
    uint64_t lo = 0, hi = 0;
    for (int i = 0; i < 8; i++) { lo |= (uint64_t)bytes[i] << (8 * i); hi |= (uint64_t)bytes[8 + i] << (8 * i); }
    diag->packet_length = lo;
    size_t avail = nb - 16;
    if (hi != 0 || lo > avail) { free(bytes); diag->status = OO_BAD_HEADER; return OO_BAD_HEADER; }
    size_t plen = (size_t)lo;
    size_t olen = cfg->fec ? oo_hamming74_decoded_len(plen) : plen;
    if (olen > out_cap) { free(bytes); diag->status = OO_BAD_HEADER; return OO_BAD_HEADER; }
    if (cfg->fec) oo_hamming74_decode(bytes + 16, plen, out);
    else memcpy(out, bytes + 16, plen);
    *out_len = olen;
    free(bytes);
    diag->status = OO_OK;
    return OO_OK;
}

/* docs/SPEC.md section 4 (capture search). The reference's equivalent is the whole-capture xcorr_fft of
 * src/receiver.rs:20-21, which finds only the single strongest frame of a buffer (examples/jetson_rx.rs:84). */
/* L = symbol length (80 for nfft = 64, 1280 for nfft = 1024): every length of section 4 scales with it -- correlation lag and
 * window L, hold-off and minimum post-offset length 10 L (lock + preamble + training), refinement window [d - 11 L / 5, d + L / 5]. */
static size_t sync_search_len(const oo_c64 *a, size_t n, oo_peak *peaks, size_t max_peaks, int LS)
{
    if (n < 2 * (size_t)LS) return 0;
    oo_c64 *lock = (oo_c64 *)malloc(sizeof(oo_c64) * (size_t)LS);
    oo_locking_signal(lock, LS);
    const long HEAD = 10L * LS;
    size_t np = 0;
    long d_last = (long)n - 2 * LS;
    long last_acc = -HEAD;
    int prev_above = 0;
    oo_c64 P = c_make(0.0, 0.0);
    double R1 = 0.0, R2 = 0.0;
    for (long d = 0; d <= d_last; d++) {
        if ((d & 1023) == 0) {                    /* exact re-sum periodically: no drift */
            P = c_make(0.0, 0.0); R1 = 0.0; R2 = 0.0;
            for (int m = 0; m < LS; m++) {
                P = c_add(P, c_mul(c_conj(a[d + m]), a[d + m + LS]));
                R1 += c_norm_sqr(a[d + m]);
                R2 += c_norm_sqr(a[d + m + LS]);
            }
        } else {
            P = c_sub(P, c_mul(c_conj(a[d - 1]), a[d - 1 + LS]));
            P = c_add(P, c_mul(c_conj(a[d - 1 + LS]), a[d - 1 + 2 * LS]));
            R1 += c_norm_sqr(a[d - 1 + LS]) - c_norm_sqr(a[d - 1]);
            R2 += c_norm_sqr(a[d - 1 + 2 * LS]) - c_norm_sqr(a[d - 1 + LS]);
        }
        int above = c_norm_sqr(P) > 0.5 * R1 * R2;
        if (above && !prev_above && d >= last_acc + HEAD) {
            last_acc = d;
            long k_lo = d - (11L * LS) / 5, k_hi = d + LS / 5;
            if (k_lo < -(LS - 1)) k_lo = -(LS - 1);
            long offset = ramp_argmax_n(a, n, k_lo, k_hi, lock, LS) - 1;
            if (offset >= 0 && (size_t)offset + (size_t)HEAD <= n && np < max_peaks) {
                const oo_c64 *x = a + offset;
                oo_c64 s = c_make(0.0, 0.0);
                for (int i = 0; i < LS; i++) {
                    s = c_add(s, c_mul(c_conj(x[2 * LS + i]), x[3 * LS + i]));
                    s = c_add(s, c_mul(c_conj(x[3 * LS + i]), x[4 * LS + i]));
                }
                oo_c64 p0 = c_make(0.0, 0.0);
                double r1 = 0.0, r2 = 0.0;
                for (int m = 0; m < LS; m++) {
                    p0 = c_add(p0, c_mul(c_conj(a[d + m]), a[d + m + LS]));
                    r1 += c_norm_sqr(a[d + m]);
                    r2 += c_norm_sqr(a[d + m + LS]);
                }
                peaks[np].offset = (uint64_t)offset;
                peaks[np].f_delta = oo_angle(s) / (double)LS;
                peaks[np].metric = c_norm_sqr(p0) / (r1 * r2);
                np++;
            }
        }
        prev_above = above;
    }
    free(lock);
    return np;
}

size_t oo_sync_search(const oo_c64 *a, size_t n, oo_peak *peaks, size_t max_peaks) { return sync_search_len(a, n, peaks, max_peaks, NSYM); }

size_t oo_sync_search_fc32(const float *iq, size_t n, oo_peak *peaks, size_t max_peaks)
{
    return oo_sync_search_fc32_n(iq, n, peaks, max_peaks, NFFT);
}

/* the same for the layout with `nfft` subcarriers (64 or 1024; docs/SPEC.md 9) */
size_t oo_sync_search_fc32_n(const float *iq, size_t n, oo_peak *peaks, size_t max_peaks, int nfft)
{
    oo_c64 *x = (oo_c64 *)malloc(sizeof(oo_c64) * (n + 1));
    oo_fc32_to_sig(iq, n, x);
    size_t r = sync_search_len(x, n, peaks, max_peaks, nfft == 1024 ? 1280 : NSYM);
    free(x);
    return r;
}

int oo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int oo_decode_batch_fc32(const float *iq, const uint32_t *n_samples, uint32_t n_streams, size_t iq_stride,
                         const oo_cfg *cfg, uint8_t *out, size_t out_stride, uint32_t *out_len,
                         int32_t *status, int32_t *offsets, int threads)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    plan_get(NFFT);
#pragma omp parallel for schedule(dynamic, 1)
    for (long s = 0; s < (long)n_streams; s++) {
        size_t n = n_samples[s];
        oo_c64 *x = (oo_c64 *)malloc(sizeof(oo_c64) * (n + 1));
        oo_fc32_to_sig(iq + 2 * iq_stride * (size_t)s, n, x);               /* src/utils.rs:239-254 */
        size_t ol = 0;
        oo_diag d;
        d.h_full = NULL;
        int st = oo_decode(x, n, cfg, out + out_stride * (size_t)s, out_stride, &ol, NULL, 0, &d);
        out_len[s] = (uint32_t)ol;
        status[s] = st;
        if (offsets) offsets[s] = d.offset;
        free(x);
    }
    return 0;
}
