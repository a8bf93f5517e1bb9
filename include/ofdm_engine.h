/*
 * ofdm_engine.h -- C ABI of the B200-native OFDM baseband engine (libofdm_b200.so).
 *
 * Drop-in boundary for the modem hot path of jkelleyrtp/ofdm. The reference has NO existing FFI
 * (no extern, no build.rs); each entry point below names the reference function(s) it replaces.
 * A thin Rust `ofdm-sys` shim (rust/ofdm-sys) binds exactly these symbols and keeps
 *   pub fn encode(data:&[u8], guard_bands:Option<bool>, modulation:Option<ModulationScheme>) -> Vec<Complex64>
 *                                                                        (src/transmitter.rs:10-15)
 *   pub fn decode(samples:Vec<Complex64>, guard_bands:Option<bool>, modulation:Option<ModulationScheme>)
 *                                                      -> anyhow::Result<Vec<u8>>   (src/receiver.rs:8-13)
 * as the host surface.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every buffer; the engine owns its handle + scratch.
 *  - IQ is fc32: interleaved float re, im (the reference's wire format, src/utils.rs:228-254).
 *  - `mem` says where ALL buffer arguments of that call live: OFDM_MEM_HOST or OFDM_MEM_DEVICE.
 *    With OFDM_MEM_HOST the call copies in/out itself (pinned memory recommended: ofdm_host_alloc) and is
 *    synchronous. With OFDM_MEM_DEVICE the call only enqueues kernels on `stream` (a cudaStream_t, may be
 *    NULL) and returns; results are valid after the stream is synchronised.
 *  - return value: 0 ok, <0 engine/CUDA error (text via ofdm_last_error). Per-stream problems never fail
 *    the call; they are reported in `status[]` (the reference panics / returns Err instead:
 *    src/receiver.rs:25,27-29,87-89).
 *  - There is no CPU fallback: without a CUDA device ofdm_engine_create fails.
 *  - A handle is not thread-safe; use one handle per GPU / host thread. No global state.
 *  - A handle owns ONE set of device scratch (per-stream state, staging, candidate lists): it supports one call in flight
 *    at a time. OFDM_MEM_DEVICE calls issued on different CUDA streams with the same handle must be ordered by the caller
 *    (events) -- or use one handle per stream. Scratch grows on demand with cudaFree + cudaMalloc, which synchronises the
 *    device; ofdm_engine_reserve sizes it once so that steady-state calls only enqueue work.
 */
#ifndef OFDM_ENGINE_H
#define OFDM_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFDM_ABI_VERSION 1

typedef struct ofdm_engine ofdm_engine;
typedef struct { float re, im; } ofdm_fc32;

/* ModulationScheme, src/transmitter.rs:98-104 (Qam is an empty arm there; 64QAM here, docs/SPEC.md) */
enum { OFDM_MOD_BPSK = 0, OFDM_MOD_QPSK = 1, OFDM_MOD_QAM64 = 2 };
enum { OFDM_SYNC_REFERENCE = 0, OFDM_SYNC_SCHMIDL_COX = 1 };      /* src/receiver.rs:20-21 | SPEC 4 */
enum { OFDM_CFO_REFERENCE = 0, OFDM_CFO_ANGLE_OF_SUM = 1 };       /* src/receiver.rs:231-240 | SPEC 5 */
enum { OFDM_PHASE_REFERENCE = 0, OFDM_PHASE_ANGLE_OF_SUM = 1 };   /* src/receiver.rs:126,137 | SPEC 6 */
enum { OFDM_MEM_HOST = 0, OFDM_MEM_DEVICE = 1 };

/* per-stream status */
enum {
    OFDM_OK = 0,
    OFDM_TOO_SHORT = 1,    /* src/receiver.rs:27-29 "Input not long enough" */
    OFDM_NO_SYNC = 2,      /* Schmidl-Cox metric never crossed the threshold */
    OFDM_BAD_HEADER = 3,   /* header length does not fit the received symbols / out_stride (src/receiver.rs:86-93 has no check) */
    OFDM_NEG_OFFSET = 4    /* src/receiver.rs:25 panics on a negative offset */
};

/* call-level errors */
enum {
    OFDM_E_INVALID = -1,   /* bad argument / unsupported configuration */
    OFDM_E_CUDA = -2,      /* a CUDA runtime call failed */
    OFDM_E_NOMEM = -3,
    OFDM_E_NODEVICE = -4   /* no CUDA device: there is no CPU fallback */
};

typedef struct {
    uint32_t struct_size;  /* = sizeof(ofdm_cfg) */
    uint32_t nfft;         /* 64 (src/receiver.rs:99), or 1024: wideband variant (docs/SPEC.md 9; no reference) */
    uint32_t cp;           /* nfft / 4: 16 or 256 */
    uint32_t modulation;   /* decode!/encode! `modulation`, default Bpsk (src/transmitter.rs:17) */
    uint32_t guard_bands;  /* decode!/encode! `guard_bands`, default false (src/transmitter.rs:16) */
    uint32_t fec;          /* 0 none; 1 Hamming(7,4) fused where the reference applies RS (examples/lab3c_image.rs:19-21) */
    uint32_t sync_mode;
    uint32_t cfo_mode;
    uint32_t phase_mode;
    uint32_t sync_window;  /* lags searched for the frame start; 0 = whole capture like the reference */
    /* optional table overrides (NULL = the reference's generators, src/transmitter.rs:60-96) */
    const ofdm_fc32 *locking;   /* nfft + cp entries (80) */
    const ofdm_fc32 *preamble;  /* nfft + cp entries (80) */
    const ofdm_fc32 *training;  /* nfft entries (64), frequency domain */
} ofdm_cfg;

/* optional per-stream diagnostics of ofdm_rx_decode_batch; every pointer may be NULL.
 * These are the reference's npy tap points (src/receiver.rs:41,52,58,76). Same `mem` as the call. */
typedef struct {
    int32_t   *offset;        /* [n_streams]  src/receiver.rs:21 */
    float     *f_delta;       /* [n_streams]  src/receiver.rs:39 */
    ofdm_fc32 *h_k;           /* [n_streams][nfft]  src/receiver.rs:56 */
    uint32_t  *n_data_syms;   /* [n_streams]  data OFDM symbols demodulated */
    ofdm_fc32 *points;        /* [n_streams][points_stride] equalised + phase-corrected data points (src/receiver.rs:76) */
    uint32_t   points_stride;
} ofdm_rx_diag;

/* parameters of the synthetic channel harness (src/channel.rs:33-74) */
typedef struct {
    float    snr_db;          /* src/channel.rs:40, default 30 */
    float    cfo_max;         /* per-stream f ~ U(0, cfo_max) rad/sample; <0: none (src/channel.rs:54) */
    uint32_t lead_min;        /* per-stream noise-only lead-in, U{lead_min..lead_max} samples */
    uint32_t lead_max;
    uint32_t multipath;       /* 1: the 12 taps of src/channel.rs:26-31 */
    uint32_t noise_mode;      /* 0: reference-faithful uniform noise (src/channel.rs:66-71); 1: Gaussian AWGN */
    uint64_t seed;
} ofdm_channel_params;

uint32_t ofdm_abi_version(void);
void     ofdm_cfg_default(ofdm_cfg *cfg);
const char *ofdm_status_name(int32_t status);

/* engine life cycle */
int  ofdm_engine_create(const ofdm_cfg *cfg, int device, ofdm_engine **out);
void ofdm_engine_destroy(ofdm_engine *h);
const char *ofdm_last_error(const ofdm_engine *h);        /* h may be NULL: error of the last failed create */
/* Pre-size the handle's device scratch for batches of up to max_streams streams / frames and captures of up to
 * max_capture_samples samples (0 = leave as is), so that later OFDM_MEM_DEVICE calls never allocate. */
int  ofdm_engine_reserve(ofdm_engine *h, uint32_t max_streams, uint64_t max_capture_samples);
int  ofdm_get_tables(const ofdm_engine *h, ofdm_fc32 *locking80, ofdm_fc32 *preamble80, ofdm_fc32 *training64);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) */
int  ofdm_host_alloc(size_t bytes, void **out);
void ofdm_host_free(void *p);

/* sizes -- src/transmitter.rs:49-54 (block count), docs/SPEC.md 3 (Hamming lengths) */
uint32_t ofdm_coded_len(const ofdm_cfg *cfg, uint32_t payload_len);       /* bytes handed to `encode` */
uint32_t ofdm_frame_data_syms(const ofdm_cfg *cfg, uint32_t payload_len); /* S */
uint32_t ofdm_frame_len(const ofdm_cfg *cfg, uint32_t payload_len);       /* (10+S)*80 samples */
uint32_t ofdm_max_payload(const ofdm_cfg *cfg, uint32_t n_data_syms);     /* largest payload that fits S symbols */

/*
 * TX: replaces `encode` (src/transmitter.rs:11-58) for a batch of frames.
 * payload[s*payload_stride ..][payload_len[s]] -> iq_out[s*iq_stride ..][frame_len]; samples past the
 * frame up to iq_stride are zero-filled. frame_len_out (optional) receives each frame's length.
 * Every frame is normalised by its own maximum positive component (`normalize`, src/transmitter.rs:183-194). Batches take a
 * speculative one-pass kernel: that maximum is max(maximum of the constant frame head, maximum of the data symbols), and for
 * scrambled payloads it is the head's -- so every symbol is scaled with the head maximum and stored at once, and the (rare)
 * frames whose data symbols beat it are redone with their own maximum by a second pass. The output does not depend on the bet,
 * only the cost does (a frame that loses it is written twice). The exact one-pass kernels keep a frame on chip (tensor memory)
 * until its maximum is known (cooperative launches: they need the device's SMs to themselves for the ~ms they run); a handful
 * of frames, or frames too long for them, take the cluster / two-pass kernels. The results are the same values whichever kernel
 * runs. (Measurement aid, read once at ofdm_engine_create: the environment variable
 * OFDM_TX_PATH = twopass | cluster | resident | warp | spec pins one of them.)
 * nfft = 1024 has the same two large-batch choices: wide_tx_resident_kernel (one pass, a symbol per warp, frames resident in
 * tensor memory; from two frames per group of CTAs on) and the two-pass wide_tx_kernel; they agree to fp32 rounding (different
 * FFT factorisations), each within 2e-6 of the oracle's `encode`.
 */
int ofdm_tx_encode_batch(ofdm_engine *h, const uint8_t *payload, const uint32_t *payload_len,
                         uint32_t payload_stride, uint32_t n_streams,
                         ofdm_fc32 *iq_out, uint32_t iq_stride, uint32_t *frame_len_out,
                         int mem, void *stream);

/*
 * RX: replaces `decode` (src/receiver.rs:9-96) for a batch of captures: sync, CFO estimate + derotation,
 * channel estimate, CP strip, FFT, equalise, pilot phase, demap, header parse, (Hamming decode).
 * iq[s*iq_stride ..][n_samples[s]] -> out[s*out_stride ..][out_len[s]], status[s].
 * max_n_samples: upper bound of n_samples[] (0 = iq_stride); only used to size the launch.
 */
int ofdm_rx_decode_batch(ofdm_engine *h, const ofdm_fc32 *iq, const uint32_t *n_samples,
                         uint32_t n_streams, uint32_t iq_stride, uint32_t max_n_samples,
                         uint8_t *out, uint32_t out_stride, uint32_t *out_len, int32_t *status,
                         const ofdm_rx_diag *diag, int mem, void *stream);

/*
 * Channel harness: replaces `channel` (src/channel.rs:33-74) with a seeded, batched device version.
 * tx[s*tx_stride ..][tx_len[s]] -> rx[s*rx_stride ..][rx_len[s]], rx_len = lead + tx_len + 63.
 * lead_out / cfo_out (optional) receive the per-stream draws.
 */
int ofdm_channel_apply_batch(ofdm_engine *h, const ofdm_fc32 *tx, const uint32_t *tx_len, uint32_t tx_stride,
                             uint32_t n_streams, const ofdm_channel_params *p,
                             ofdm_fc32 *rx, uint32_t rx_stride, uint32_t *rx_len,
                             uint32_t *lead_out, float *cfo_out, int mem, void *stream);

/*
 * Preamble search over ONE long capture: replaces `samples.xcorr_fft(locking_signal)` on a whole radio buffer
 * (src/receiver.rs:20-21, src/signals/mod.rs:186-217; examples/jetson_rx.rs:46-57,84 is the 2 M-sample case) with a single
 * 8 B/sample pass of the sliding Schmidl-Cox metric (docs/SPEC.md 4): every frame start is reported once as
 * (offset by the lag-1 rule, CFO estimate, metric). Captures of any length that fits the device (64-bit offsets; with
 * OFDM_MEM_HOST the capture is copied in chunks that overlap the scan). peaks[0 .. *n_peaks) is in ascending offset order;
 * with OFDM_MEM_DEVICE entries whose metric < 0 are unusable detections (frame head cut by the capture end) and
 * *n_peaks counts them too; with OFDM_MEM_HOST they are removed. More than max_peaks detections are truncated:
 * *n_peaks = min(detections, max_peaks) in both modes, so it can be handed to ofdm_rx_decode_capture as it is.
 * ofdm_sync_counts (waits for `stream`) tells a device-mode caller whether that happened: counts[0] = threshold crossings the
 * scan recorded, counts[1] = frames detected after the hold-off, counts[2] = entries written to peaks[], counts[3] = 3928-lag
 * tiles that held more than 12 threshold crossings (their surplus is dropped; a capture made of frames never does that --
 * OFDM_MEM_HOST turns it into an error). counts[1] > counts[2] means peaks[] was too small.
 * nfft = 1024 engines run the same search with every length scaled by 16 (docs/SPEC.md 9; 4096-lag tiles), and so do
 * ofdm_rx_decode_capture and ofdm_rx_decode_file below.
 */
typedef struct {
    uint64_t offset;
    float    f_delta;
    float    metric;
} ofdm_peak;
int ofdm_sync_search(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n_samples, ofdm_peak *peaks, uint32_t max_peaks,
                     uint32_t *n_peaks, int mem, void *stream);
int ofdm_sync_counts(ofdm_engine *h, uint32_t counts[4], void *stream);

/*
 * Streaming receiver: decode every frame ofdm_sync_search found in one long capture -- the loop of
 * examples/jetson_rx.rs:46-57,83-108 (capture buffer -> decode! -> drop on failure) for all frames of the buffer at once.
 * Frame i starts at peaks[i].offset and owns the samples up to the next detection (capped by max_frame_samples, 0 = no cap);
 * out / out_len / status are indexed by frame. peaks may come straight from ofdm_sync_search (same `mem`).
 */
int ofdm_rx_decode_capture(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n_samples, const ofdm_peak *peaks, uint32_t n_frames,
                           uint32_t max_frame_samples, uint8_t *out, uint32_t out_stride, uint32_t *out_len, int32_t *status,
                           int mem, void *stream);

/*
 * fc32 capture FILE -> every frame decoded: `decode!(bytes_to_sig(read(path))[start..stop])` of examples/lab3c.rs:57-74
 * (src/utils.rs:238-254 is the file format: interleaved native-endian f32 pairs, what `rx_samples_to_file --type float` writes,
 * data/receive.sh:1) for captures of any length holding any number of frames -- the radio loop of examples/jetson_rx.rs:46-57
 * decodes one frame per buffer. The file is read in chunks of chunk_samples (0 = 32 Mi samples) straight into two pinned
 * buffers, one chunk ahead of the GPU; a chunk is one PCIe copy + ofdm_sync_search + ofdm_rx_decode_capture on the device;
 * chunks overlap by max_frame_samples (the longest frame the link carries, head included), so every frame is decoded exactly
 * once from a chunk that holds all of it. start / stop are sample indices into the file (stop = 0: end of file); frame
 * offsets are relative to start. Frame i's payload is at out[i * out_stride .. + frames[i].out_len).
 * Fails (OFDM_E_INVALID) when the file holds more than max_frames frames or a chunk more than 16 384.
 */
typedef struct {
    uint64_t offset;     /* frame start, samples from `start` (lag - 1 rule, src/receiver.rs:21) */
    float    f_delta;    /* CFO estimate, rad/sample */
    float    metric;     /* Schmidl-Cox metric at the detection */
    int32_t  status;     /* OFDM_OK ... */
    uint32_t out_len;    /* decoded payload bytes */
} ofdm_frame_info;
int ofdm_rx_decode_file(ofdm_engine *h, const char *path, uint64_t start, uint64_t stop, uint32_t chunk_samples,
                        uint32_t max_frame_samples, uint8_t *out, uint32_t out_stride, ofdm_frame_info *frames,
                        uint32_t max_frames, uint32_t *n_frames);

/*
 * BER: replaces utils::Analysis::new (src/utils.rs:45-68) for a batch and accumulates into
 * counters[4] = { bit_errs, byte_errs, bits_compared, frames_failed }. A stream whose status != OK or
 * whose length differs from ref_len counts as failed with all its reference bits in error.
 * The counters are what a multi-GPU run sum-reduces (NCCL, 4 x u64) -- the only collective of the path.
 */
int ofdm_ber_accumulate(ofdm_engine *h, const uint8_t *ref, const uint32_t *ref_len, uint32_t ref_stride,
                        const uint8_t *got, const uint32_t *got_len, uint32_t got_stride,
                        const int32_t *status, uint32_t n_streams, uint64_t *counters,
                        int mem, void *stream);

/*
 * The path's only collective: sum-reduce counters[4] over the ranks of an NCCL communicator (one process per GPU; the
 * fields of utils::Analysis, src/utils.rs:39-43, plus the failure count). nccl_comm is the caller's ncclComm_t. The library
 * has no link-time NCCL dependency: ncclAllReduce is resolved at the first call from the NCCL already loaded in the
 * process (else libnccl.so.2). OFDM_MEM_DEVICE: in place on `stream`; OFDM_MEM_HOST: copied to the device, reduced, copied
 * back, synchronous. Callers without NCCL sum the four integers themselves -- that is all this does.
 */
int ofdm_stats_allreduce(ofdm_engine *h, uint64_t *counters, void *nccl_comm, int mem, void *stream);

/*
 * Reed-Solomon RS(255,223) outer code with the reference's stream framing.
 * ofdm_rs_encode_batch replaces create_transmission_bytes (src/utils.rs:97-137): 223-byte blocks, each followed by its
 * 32 parity bytes; the partially filled -- possibly empty -- last block is always emitted, zero filled:
 * coded_len = 255 * (data_len / 223 + 1). ofdm_rs_decode_batch replaces decipher_transmission_bytes
 * (src/utils.rs:152-180): 255-byte blocks, the partial -- possibly empty -- tail zero filled and decoded as well:
 * data_len = 223 * (coded_len / 255 + 1). Code: the `reed-solomon` 0.2.1 crate's (Cargo.toml:35), i.e. GF(2^8) with
 * polynomial 0x11d, alpha = 2, roots alpha^0..alpha^31, message first, parity last; up to 16 symbol errors per block
 * are corrected. n_corrected[s] = symbols corrected in stream s; n_failed[s] = blocks beyond repair (the reference
 * returns None for the whole stream when it is non-zero; their data bytes are passed through uncorrected).
 * The *_len outputs are always the required lengths; a stream whose output does not fit its stride is not written.
 */
size_t ofdm_rs_encoded_len(size_t data_len);
size_t ofdm_rs_decoded_len(size_t coded_len);
int ofdm_rs_encode_batch(ofdm_engine *h, const uint8_t *data, const uint32_t *data_len, uint32_t n_streams, uint32_t data_stride,
                         uint8_t *coded, uint32_t coded_stride, uint32_t *coded_len, int mem, void *stream);
int ofdm_rs_decode_batch(ofdm_engine *h, const uint8_t *coded, const uint32_t *coded_len, uint32_t n_streams, uint32_t coded_stride,
                         uint8_t *data, uint32_t data_stride, uint32_t *data_len, uint32_t *n_corrected, uint32_t *n_failed,
                         int mem, void *stream);

/*
 * Per-kernel device timing of ofdm_rx_decode_batch(OFDM_MEM_DEVICE) for the roofline report: after
 * ofdm_profile_begin(h, n) the next n calls record CUDA events on their stream around the acquisition and the
 * decode kernel; ofdm_profile_read waits for them and returns the durations (ms) of each call.
 */
int ofdm_profile_begin(ofdm_engine *h, uint32_t max_calls);
int ofdm_profile_read(ofdm_engine *h, float *acquire_ms, float *decode_ms, uint32_t *n_calls);

/* how many kernels this handle has launched so far (bench.py's gpu_launches) */
uint64_t ofdm_kernel_launches(const ofdm_engine *h);

/*
 * Host -> device bytes the last ofdm_rx_decode_batch(OFDM_MEM_HOST) call moved (bench.py's e2e.h2d_bytes_per_step). With
 * sync_window > 0 and captures much longer than the window the host path fetches the frame head region of every stream,
 * lets the acquisition locate the frame, and then copies only the useful nfft samples of the data symbols the header asks
 * for: the cyclic prefixes that `unprefix_block` (src/receiver.rs:104-118) discards never cross PCIe. Otherwise the whole
 * capture is copied. OFDM_RX_FEED=full|skipcp in the environment pins one of the two for A/B measurements.
 */
uint64_t ofdm_last_h2d_bytes(const ofdm_engine *h);

#ifdef __cplusplus
}
#endif
#endif
