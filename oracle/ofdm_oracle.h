/*
 * ofdm_oracle.h -- CPU restatement of the jkelleyrtp/ofdm modem hot path (TEST INFRASTRUCTURE).
 *
 * This is the checker, not the product. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product path
 * (ofdm_b200/, include/ofdm_engine.h) never links, imports or calls anything here.
 *
 * Parity status: PINNED at stage level on the reference's own known-answer tests
 * (QPSK mod/demod round trip src/lib.rs:37-51, bit order src/utils.rs:281-327,
 * xcorr lags src/signals/mod.rs:420-441, channel response listing src/channel.rs:99-177,
 * angle src/receiver.rs:253-256, fft_shift src/signals/mod.rs:61-77).
 * UNPINNED end to end: the reference holds no golden IQ / decoded-bytes fixture and cannot
 * be compiled here (no rustc/cargo; out-of-tree and git dependencies). The StdRng tables
 * (preamble / training) are a best-effort ChaCha12 restatement that cannot be verified offline.
 * 64QAM / Hamming / Schmidl-Cox have no reference at all (docs/SPEC.md).
 *
 * All arithmetic is f64 like the reference (num::Complex64).
 */
#ifndef OFDM_ORACLE_H
#define OFDM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } oo_c64;

enum { OO_BPSK = 0, OO_QPSK = 1, OO_QAM64 = 2 };
enum { OO_SYNC_REFERENCE = 0, OO_SYNC_SCHMIDL_COX = 1 };
enum { OO_CFO_REFERENCE = 0, OO_CFO_ANGLE_OF_SUM = 1 };
enum { OO_PHASE_REFERENCE = 0, OO_PHASE_ANGLE_OF_SUM = 1 };
enum { OO_OK = 0, OO_TOO_SHORT = 1, OO_NO_SYNC = 2, OO_BAD_HEADER = 3, OO_NEG_OFFSET = 4 };

typedef struct {
    int32_t guard_bands;   /* 0/1 */
    int32_t modulation;    /* OO_BPSK.. */
    int32_t fec;           /* 0 none, 1 Hamming(7,4) fused */
    int32_t sync_mode;
    int32_t cfo_mode;
    int32_t phase_mode;
    int32_t sync_window;   /* 0 = whole capture */
    int32_t xcorr_fft;     /* 1: SYNC_REFERENCE uses the FFT cross-correlation like the reference; 0: direct form */
    int32_t nfft;          /* 0 or 64: the reference's layout; 1024: wideband variant (docs/SPEC.md section 9, CP = 256) */
} oo_cfg;

typedef struct {
    int32_t status;
    int32_t offset;
    double  f_delta;
    oo_c64  h_k[64];
    int64_t n_data_syms;     /* data OFDM symbols demodulated */
    int64_t n_points;        /* equalised data points written to `points` */
    uint64_t packet_length;  /* low 64 bits of header */
    oo_c64  *h_full;         /* optional in: receives all nfft bins of h_k */
} oo_diag;

/* ---- tables (src/transmitter.rs:60-96) ---- */
void oo_locking_signal(oo_c64 *out, int len);
void oo_preamble(oo_c64 *out, int len);
void oo_training_signals(oo_c64 *out, int len);
/* rand 0.8 StdRng::seed_from_u64 + gen_range(-1.0..1.0) stream, for tests */
void oo_stdrng_uniform_pm1(uint64_t seed, double *out, int n);
void oo_chacha12_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]);        /* KAT hook */
void oo_stdrng_from_seed_u64(const uint8_t seed[32], uint64_t *out, int n);                   /* KAT hook: StdRng::from_seed */

/* ---- signal primitives (src/signals/mod.rs) ---- */
void oo_fft(oo_c64 *x, size_t n, int inverse_scaled);            /* any n (pow2 fast path, else Bluestein) */
void oo_fft_shift(oo_c64 *x, size_t n);
void oo_ifft_shift(oo_c64 *x, size_t n);
size_t oo_xcorr_fft(const oo_c64 *a, size_t a_len, const oo_c64 *b, size_t b_len, oo_c64 *out /* 2*a_len-1 */);
void oo_convolve(const oo_c64 *a, size_t a_len, const oo_c64 *b, size_t b_len, oo_c64 *out /* a_len+b_len-1 */);
oo_c64 oo_mean(const oo_c64 *x, size_t n);
oo_c64 oo_variance(const oo_c64 *x, size_t n);
double oo_angle(oo_c64 z);

/* ---- bits / BER / wire format (src/utils.rs) ---- */
void oo_to_bools(uint8_t byte, uint8_t out[8]);
uint8_t oo_bools_to_u8(const uint8_t b[8]);
void oo_analysis(const uint8_t *l, const uint8_t *r, size_t n, uint32_t *num_errs, uint32_t *num_block_errs, double *err_rate);
void oo_sig_to_fc32(const oo_c64 *x, size_t n, float *out /* 2n */);
void oo_fc32_to_sig(const float *in, size_t n, oo_c64 *out);

/* ---- FEC (docs/SPEC.md section 3) ---- */
size_t oo_hamming74_encoded_len(size_t n);
size_t oo_hamming74_decoded_len(size_t n_coded);
void oo_hamming74_encode(const uint8_t *in, size_t n, uint8_t *out);
void oo_hamming74_decode(const uint8_t *in, size_t n_coded, uint8_t *out);

/* Reed-Solomon outer code (rs255.c): the reference's `reed-solomon` 0.2.1 call sites, src/utils.rs:97-137,152-180 */
void oo_rs_encode_block(const uint8_t *msg, int k, int nsym, uint8_t *parity);
int oo_rs_correct_block(uint8_t *word, int n, int nsym);
size_t oo_rs_encoded_len(size_t n);
size_t oo_rs_decoded_len(size_t n);
void oo_rs_encode(const uint8_t *in, size_t n, uint8_t *out);
int oo_rs_decode(const uint8_t *in, size_t n, uint8_t *out, uint32_t *n_corrected, uint32_t *n_failed);

/* ---- TX (src/transmitter.rs) ---- */
size_t oo_modulate(const uint8_t *bytes, size_t n, int scheme, oo_c64 *out);      /* returns #symbols */
size_t oo_demodulate(const oo_c64 *syms, size_t n, int scheme, uint8_t *out);      /* returns #bytes */
size_t oo_frame_data_syms(size_t n_bytes, int guard_bands, int scheme);            /* S for n_bytes given to encode */
size_t oo_frame_len(size_t n_bytes, int guard_bands, int scheme);                  /* (10+S)*80 */
size_t oo_encode(const uint8_t *data, size_t n, int guard_bands, int scheme, oo_c64 *out);
size_t oo_encode_n(const uint8_t *data, size_t n, int guard_bands, int scheme, int nfft, oo_c64 *out);
size_t oo_frame_data_syms_n(size_t n_bytes, int guard_bands, int scheme, int nfft);
size_t oo_frame_len_n(size_t n_bytes, int guard_bands, int scheme, int nfft);
/* payload -> (optional Hamming) -> encode; returns frame length in samples */
size_t oo_tx(const uint8_t *payload, size_t n, const oo_cfg *cfg, oo_c64 *out);
size_t oo_tx_len(size_t n_payload, const oo_cfg *cfg);

/* ---- channel (src/channel.rs), seeded ---- */
/* noise_mode 0: reference-faithful (uniform, complex "variance"); 1: proper complex Gaussian AWGN.
 * f_delta < 0 : no CFO. out has n + 63 samples. */
void oo_channel(const oo_c64 *tx, size_t n, double snr_db, double f_delta, int noise_mode, uint64_t seed, oo_c64 *out);

/* ---- RX (src/receiver.rs) ---- */
/* out_cap bytes available in out; points (optional, may be NULL) receives equalised+phase-corrected data points. */
int oo_decode(const oo_c64 *samples, size_t n, const oo_cfg *cfg,
              uint8_t *out, size_t out_cap, size_t *out_len,
              oo_c64 *points, size_t points_cap, oo_diag *diag);

/* preamble search over one long capture (docs/SPEC.md section 4: sliding Schmidl-Cox rising edges, 800-sample hold-off,
 * ramp-correlation refinement with the lag-1 rule, angle-of-sum CFO). Returns the number of peaks written. */
typedef struct { uint64_t offset; double f_delta; double metric; } oo_peak;
size_t oo_sync_search(const oo_c64 *a, size_t n, oo_peak *peaks, size_t max_peaks);
size_t oo_sync_search_fc32(const float *iq, size_t n, oo_peak *peaks, size_t max_peaks);
size_t oo_sync_search_fc32_n(const float *iq, size_t n, oo_peak *peaks, size_t max_peaks, int nfft);   /* nfft = 64 or 1024: every length scales with the symbol length */

/* fc32 batch front end used by the CPU baseline: threads = OpenMP threads (0 = default). */
int oo_decode_batch_fc32(const float *iq, const uint32_t *n_samples, uint32_t n_streams, size_t iq_stride,
                         const oo_cfg *cfg, uint8_t *out, size_t out_stride, uint32_t *out_len,
                         int32_t *status, int32_t *offsets, int threads);
int oo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
