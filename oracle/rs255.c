/* rs255.c -- CPU oracle for the reference's Reed-Solomon outer code. TEST INFRASTRUCTURE ONLY (see ofdm_oracle.h).
 *
 * The reference calls the third-party crate `reed-solomon = "0.2.1"` (Cargo.toml:35), which is NOT under /root/reference.
 * Call sites: `Encoder::new(32)` / `encoder.encode(&[u8;223])` in create_transmission_bytes (src/utils.rs:97-137) and
 * `Decoder::new(32)` / `decoder.correct(&[u8;255], None)` in decipher_transmission_bytes (src/utils.rs:152-180).
 * The crate is a port of the "Reed-Solomon codes for coders" construction; its published algorithm is restated here:
 *   field GF(2^8), primitive polynomial x^8+x^4+x^3+x^2+1 (0x11d), alpha = 2;
 *   generator g(x) = prod_{i=0}^{nsym-1} (x - alpha^i)   (first consecutive root alpha^0);
 *   systematic codeword = message bytes (highest-degree coefficient first) followed by the nsym remainder bytes of
 *   message(x) * x^nsym mod g(x);
 *   correct(): bounded-distance decoding -- succeeds iff the word is within nsym/2 symbols of a codeword.
 * The reference holds no known-answer test for it (`ecc_packets`, src/utils.rs:358-367, only prints): parity unpinned
 * beyond the construction above. The oracle is pinned on the one published vector of that construction (the QR-code
 * example, nsym = 10) in tests/test_oracle_kats.py, and on the code's defining properties.
 */
#include "ofdm_oracle.h"
#include <string.h>

static uint8_t gf_exp[512];
static uint8_t gf_log[256];
static int gf_ready = 0;

static void gf_init(void)
{
    if (gf_ready) return;
    unsigned x = 1;
    for (int i = 0; i < 255; i++) {
        gf_exp[i] = (uint8_t)x;
        gf_log[x] = (uint8_t)i;
        x <<= 1;
        if (x & 0x100) x ^= 0x11d;
    }
    for (int i = 255; i < 512; i++) gf_exp[i] = gf_exp[i - 255];
    gf_log[0] = 0;
    gf_ready = 1;
}

static inline uint8_t gf_mul(uint8_t a, uint8_t b) { return (a && b) ? gf_exp[gf_log[a] + gf_log[b]] : 0; }
static inline uint8_t gf_div(uint8_t a, uint8_t b) { return a ? gf_exp[gf_log[a] + 255 - gf_log[b]] : 0; }   /* b != 0 */
static inline uint8_t gf_pow2(int e) { e %= 255; if (e < 0) e += 255; return gf_exp[e]; }

/* g(x), coefficients highest degree first, nsym + 1 entries, g[0] = 1 */
static void rs_generator(int nsym, uint8_t *g)
{
    memset(g, 0, (size_t)nsym + 1);
    g[0] = 1;
    for (int i = 0; i < nsym; i++) {
        /* multiply by (x - alpha^i) = (x + alpha^i) */
        const uint8_t r = gf_pow2(i);
        for (int j = i + 1; j >= 1; j--) g[j] = (uint8_t)(g[j] ^ gf_mul(g[j - 1], r));
    }
}

/* nsym parity bytes of msg[0..k): remainder of msg(x) x^nsym by g(x), schoolbook long division */
void oo_rs_encode_block(const uint8_t *msg, int k, int nsym, uint8_t *parity)
{
    gf_init();
    uint8_t g[256 + 1], buf[255 + 256];
    rs_generator(nsym, g);
    memcpy(buf, msg, (size_t)k);
    memset(buf + k, 0, (size_t)nsym);
    for (int i = 0; i < k; i++) {
        const uint8_t c = buf[i];
        if (!c) continue;
        for (int j = 1; j <= nsym; j++) buf[i + j] = (uint8_t)(buf[i + j] ^ gf_mul(g[j], c));
    }
    memcpy(parity, buf + k, (size_t)nsym);
}

/* polynomial evaluation, coefficients highest degree first */
static uint8_t poly_eval(const uint8_t *p, int n, uint8_t x)
{
    uint8_t y = 0;
    for (int i = 0; i < n; i++) y = (uint8_t)(gf_mul(y, x) ^ p[i]);
    return y;
}

/* Corrects word[0..n) (n <= 255, the last nsym bytes are parity) in place.
 * Returns the number of corrected symbols, or -1 when the word is not within nsym/2 symbols of a codeword. */
int oo_rs_correct_block(uint8_t *word, int n, int nsym)
{
    gf_init();
    uint8_t S[256];
    int any = 0;
    for (int i = 0; i < nsym; i++) { S[i] = poly_eval(word, n, gf_pow2(i)); any |= S[i]; }
    if (!any) return 0;

    /* Berlekamp-Massey: error locator Lambda(x) = 1 + L1 x + ..., lowest degree first */
    uint8_t C[260], B[260], T[260];
    memset(C, 0, sizeof C); memset(B, 0, sizeof B);
    C[0] = 1; B[0] = 1;
    int L = 0, m = 1;
    uint8_t b = 1;
    for (int r = 0; r < nsym; r++) {
        uint8_t d = S[r];
        for (int i = 1; i <= L; i++) d = (uint8_t)(d ^ gf_mul(C[i], S[r - i]));
        if (d == 0) { m++; continue; }
        const uint8_t coef = gf_div(d, b);
        if (2 * L <= r) {
            memcpy(T, C, sizeof T);
            for (int i = 0; i + m < 260; i++) C[i + m] = (uint8_t)(C[i + m] ^ gf_mul(coef, B[i]));
            L = r + 1 - L;
            memcpy(B, T, sizeof B);
            b = d;
            m = 1;
        } else {
            for (int i = 0; i + m < 260; i++) C[i + m] = (uint8_t)(C[i + m] ^ gf_mul(coef, B[i]));
            m++;
        }
    }
    if (2 * L > nsym) return -1;

    /* Chien search over the n positions: position p (0 = first byte) has locator X = alpha^(n-1-p); root at X^-1 */
    int pos[128], nerr = 0;
    for (int p = 0; p < n; p++) {
        const uint8_t xinv = gf_pow2(255 - (n - 1 - p));
        uint8_t y = 0;
        for (int i = L; i >= 0; i--) y = (uint8_t)(gf_mul(y, xinv) ^ C[i]);
        if (y == 0) { if (nerr < 128) pos[nerr] = p; nerr++; }
    }
    if (nerr != L) return -1;

    /* Forney: Omega(x) = S(x) Lambda(x) mod x^nsym; e = X^(1-fcr) Omega(X^-1) / Lambda'(X^-1), fcr = 0 */
    uint8_t Om[256];
    for (int i = 0; i < nsym; i++) {
        uint8_t v = 0;
        for (int j = 0; j <= i && j <= L; j++) v = (uint8_t)(v ^ gf_mul(C[j], S[i - j]));
        Om[i] = v;
    }
    for (int e = 0; e < nerr; e++) {
        const int p = pos[e];
        const uint8_t X = gf_pow2(n - 1 - p), xinv = gf_pow2(255 - (n - 1 - p));
        uint8_t om = 0;
        for (int i = nsym - 1; i >= 0; i--) om = (uint8_t)(gf_mul(om, xinv) ^ Om[i]);
        uint8_t dl = 0;                                        /* formal derivative: odd-degree terms */
        for (int i = 1; i <= L; i += 2) dl = (uint8_t)(dl ^ gf_mul(C[i], gf_pow2((255 - (n - 1 - p)) * (i - 1))));
        if (dl == 0) return -1;
        word[p] = (uint8_t)(word[p] ^ gf_mul(X, gf_div(om, dl)));
    }
    for (int i = 0; i < nsym; i++) if (poly_eval(word, n, gf_pow2(i))) return -1;
    return nerr;
}

/* ---- the reference's stream framing ----------------------------------------------------------------------------- */

/* src/utils.rs:97-137: 223-byte blocks; the partially filled (possibly empty) last block is always emitted, zero filled */
size_t oo_rs_encoded_len(size_t n) { return 255 * (n / 223 + 1); }
/* src/utils.rs:152-180: 255-byte blocks; the partially filled (possibly empty) last block is always decoded */
size_t oo_rs_decoded_len(size_t n) { return 223 * (n / 255 + 1); }

void oo_rs_encode(const uint8_t *in, size_t n, uint8_t *out)
{
    const size_t nb = n / 223 + 1;
    for (size_t b = 0; b < nb; b++) {
        uint8_t blk[223];
        memset(blk, 0, sizeof blk);
        const size_t have = b * 223 < n ? (n - b * 223 < 223 ? n - b * 223 : 223) : 0;
        memcpy(blk, in + b * 223, have);
        memcpy(out + b * 255, blk, 223);
        oo_rs_encode_block(blk, 223, 32, out + b * 255 + 223);
    }
}

/* Returns 0 and the corrected-symbol total in *n_corrected, or -1 when any block fails (the reference returns None). out is
 * written for every block that decodes; a failed block's data bytes are copied uncorrected. */
int oo_rs_decode(const uint8_t *in, size_t n, uint8_t *out, uint32_t *n_corrected, uint32_t *n_failed)
{
    const size_t nb = n / 255 + 1;
    uint32_t corr = 0, fail = 0;
    for (size_t b = 0; b < nb; b++) {
        uint8_t blk[255];
        memset(blk, 0, sizeof blk);
        const size_t have = b * 255 < n ? (n - b * 255 < 255 ? n - b * 255 : 255) : 0;
        memcpy(blk, in + b * 255, have);
        uint8_t fixed[255];
        memcpy(fixed, blk, 255);
        const int r = oo_rs_correct_block(fixed, 255, 32);
        if (r < 0) { fail++; memcpy(out + b * 223, blk, 223); }
        else { corr += (uint32_t)r; memcpy(out + b * 223, fixed, 223); }
    }
    if (n_corrected) *n_corrected = corr;
    if (n_failed) *n_failed = fail;
    return fail ? -1 : 0;
}
