// ofdm_engine.cu -- C ABI (include/ofdm_engine.h) over the sm_100a kernels. No torch types, no CPU fallback.
#include "../../include/ofdm_engine.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <future>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <set>
#include <chrono>
#include <string>
#include <vector>

#include "kernels.h"
#include "tables.h"

using namespace ofdm;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct ofdm_engine {
    ofdm_cfg cfg;
    int device = 0;
    std::string err;
    uint64_t launches = 0;
    RxTables *d_tables = nullptr;
    wide::WideTables *d_wtables = nullptr;   // nfft = 1024
    bool wide = false;
    bool lock_is_ramp = true;           // false when cfg.locking overrides the built-in table
    int nfft = 64, sym_len = 80;
    std::vector<ofdm_fc32> lock, pre, train;
    DevBuf state;                       // StreamState[n_streams] (StreamStateW for nfft = 1024)
    DevBuf scratch_u32;                 // frame_len / stream_max
    DevBuf scratch_f32;                 // channel accumulators (5 x f64 per stream)
    DevBuf counters;                    // 4 x u64
    DevBuf sync_scratch;                // candidate list + counters of ofdm_sync_search
    DevBuf cap_base;                    // per-frame base / length of ofdm_rx_decode_capture
    DevBuf rs_tables;                   // RsTables, built on first use
    DevBuf wtx_slots;                   // wide_tx_resident_kernel: one word per warp of a group and frame (frame maximum exchange)
    // host-mode staging
    DevBuf s_iq3, s_plan, s_moved;
    DevBuf s_iq, s_iq2, s_bytes, s_bytes2, s_len, s_len2, s_status, s_aux, s_points, s_h;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_copied[2] = { nullptr, nullptr }, ev_done[2] = { nullptr, nullptr };
    // host path of ofdm_rx_decode_batch that leaves the cyclic prefixes on the host (rx_host_skip_cp)
    cudaEvent_t ev_feed[3][3] = {};                 // per staging buffer: head copied, plan on the host, decode done
    cudaStream_t feed_streams[4] = {};              // the per-stream 2-D copies of a chunk are spread over several streams (copy engines)
    cudaEvent_t ev_lane[4] = {};
    void *pin_plan = nullptr;                       // pinned: per stream (offset, data symbols to fetch), 3 chunks
    size_t pin_plan_cap = 0;
    uint64_t h2d_bytes_last = 0;                    // host -> device bytes of the last host-mode RX call
    int rx_feed = 0;                                // 0 = automatic; OFDM_RX_FEED=full|skipcp|gather pins one (A/B measurements)
    int bps_sym = 0, bpc = 0, dcar = 0, tile_shift = 0;
    int n_sm = 0;                       // multiprocessors of the device (grid of the persistent TX kernel)
    int tx_path = 0;                    // 0 = automatic; OFDM_TX_PATH=twopass|cluster|resident|warp|spec pins one (A/B measurements)
    // per-kernel timing of ofdm_rx_decode_batch(OFDM_MEM_DEVICE): 3 events per call (start, after acquire, end)
    std::vector<cudaEvent_t> prof_ev;
    uint32_t prof_cap = 0, prof_n = 0;
    std::set<const void *> smem_configured;   // kernels whose dynamic shared memory limit has been raised
    void *pin[2] = { nullptr, nullptr };      // pinned chunk buffers of ofdm_rx_decode_file
    size_t pin_bytes = 0;
};

#define ENG_FAIL(h, code, ...)                                   \
    do {                                                         \
        char _b[512];                                            \
        snprintf(_b, sizeof _b, __VA_ARGS__);                    \
        (h)->err = _b;                                           \
        return (code);                                           \
    } while (0)

#define CU(h, call)                                                                              \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) ENG_FAIL(h, OFDM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

// gridDim.y carries the stream index (<= 65535): larger batches are launched in chunks, the kernels add a.stream0
template <class K, class A>
static void launch_streams(K k, A a, uint32_t grid_x, uint32_t n_streams, unsigned threads, size_t smem, cudaStream_t st, uint64_t &launches)
{
    for (uint32_t s0 = 0; s0 < n_streams; s0 += 65535u) {
        a.stream0 = s0;
        const uint32_t ns = n_streams - s0 < 65535u ? n_streams - s0 : 65535u;
        k<<<dim3(grid_x, ns), threads, smem, st>>>(a);
        launches++;
    }
}

// ---- sizes -------------------------------------------------------------------------------------------------------
static int cfg_bpc(const ofdm_cfg *c) { return c->modulation == 0 ? 1 : (c->modulation == 1 ? 2 : 6); }
static int cfg_dcar(const ofdm_cfg *c) { return c->nfft == 1024 ? (c->guard_bands ? 768 : 1024) : (c->guard_bands ? 48 : 64); }
static int cfg_symlen(const ofdm_cfg *c) { return c->nfft == 1024 ? 1280 : 80; }

extern "C" uint32_t ofdm_abi_version(void) { return OFDM_ABI_VERSION; }

extern "C" void ofdm_cfg_default(ofdm_cfg *cfg)
{
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = sizeof *cfg;
    cfg->nfft = 64; cfg->cp = 16;
    cfg->modulation = OFDM_MOD_BPSK;        // src/transmitter.rs:17
    cfg->guard_bands = 0;                   // src/transmitter.rs:16
}

extern "C" const char *ofdm_status_name(int32_t s)
{
    switch (s) {
    case OFDM_OK: return "OK";
    case OFDM_TOO_SHORT: return "TOO_SHORT";
    case OFDM_NO_SYNC: return "NO_SYNC";
    case OFDM_BAD_HEADER: return "BAD_HEADER";
    case OFDM_NEG_OFFSET: return "NEG_OFFSET";
    default: return "UNKNOWN";
    }
}

extern "C" uint32_t ofdm_coded_len(const ofdm_cfg *cfg, uint32_t payload_len)
{
    return cfg->fec ? (uint32_t)((14ull * payload_len + 7) / 8) : payload_len;
}
extern "C" uint32_t ofdm_frame_data_syms(const ofdm_cfg *cfg, uint32_t payload_len)
{
    uint64_t nbits = 128 + 8ull * ofdm_coded_len(cfg, payload_len);
    uint64_t ncar = (nbits + cfg_bpc(cfg) - 1) / cfg_bpc(cfg);
    return (uint32_t)((ncar + cfg_dcar(cfg) - 1) / cfg_dcar(cfg));
}
extern "C" uint32_t ofdm_frame_len(const ofdm_cfg *cfg, uint32_t payload_len)
{
    return (10 + ofdm_frame_data_syms(cfg, payload_len)) * (uint32_t)cfg_symlen(cfg);
}
extern "C" uint32_t ofdm_max_payload(const ofdm_cfg *cfg, uint32_t n_data_syms)
{
    uint64_t bits = (uint64_t)n_data_syms * cfg_bpc(cfg) * cfg_dcar(cfg);
    if (bits < 128) return 0;
    uint64_t coded = (bits - 128) / 8;
    return (uint32_t)(cfg->fec ? (8 * coded) / 14 : coded);
}

// ---- life cycle --------------------------------------------------------------------------------------------------
static int validate_cfg(const ofdm_cfg *c, std::string &err)
{
    if (!c || c->struct_size != sizeof(ofdm_cfg)) { err = "ofdm_cfg.struct_size mismatch"; return OFDM_E_INVALID; }
    if (!((c->nfft == 64 && c->cp == 16) || (c->nfft == 1024 && c->cp == 256))) {
        err = "supported layouts: nfft=64/cp=16 (the reference's) and nfft=1024/cp=256 (wideband variant)";
        return OFDM_E_INVALID;
    }
    if (c->locking)                 // every sync kernel correlates with the real part only (the reference's table is a real ramp)
        for (uint32_t i = 0; i < c->nfft + c->cp; i++)
            if (c->locking[i].im != 0.0f) { err = "ofdm_cfg.locking must be real (imaginary parts are not supported by the sync kernels)"; return OFDM_E_INVALID; }
    if (c->modulation > 2 || c->guard_bands > 1 || c->fec > 1 || c->sync_mode > 1 || c->cfo_mode > 1 || c->phase_mode > 1) {
        err = "ofdm_cfg field out of range";
        return OFDM_E_INVALID;
    }
    return 0;
}

extern "C" int ofdm_engine_create(const ofdm_cfg *cfg, int device, ofdm_engine **out)
{
    if (!out) return OFDM_E_INVALID;
    *out = nullptr;
    int rc = validate_cfg(cfg, g_create_error);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return OFDM_E_NODEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return OFDM_E_INVALID; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return OFDM_E_CUDA; }

    ofdm_engine *h = new ofdm_engine();
    h->cfg = *cfg;
    h->device = device;
    h->bpc = cfg_bpc(cfg);
    h->dcar = cfg_dcar(cfg);
    cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (const char *rf = getenv("OFDM_RX_FEED")) h->rx_feed = !strcmp(rf, "full") ? 1 : !strcmp(rf, "skipcp") ? 2 : !strcmp(rf, "gather") ? 3 : 0;
    if (const char *tp = getenv("OFDM_TX_PATH"))
        h->tx_path = !strcmp(tp, "twopass") ? 1 : !strcmp(tp, "cluster") ? 2 : !strcmp(tp, "resident") ? 3 : !strcmp(tp, "warp") ? 4 : !strcmp(tp, "spec") ? 5 : 0;
    h->bps_sym = h->bpc * h->dcar;
    h->tile_shift = 0;
    if (cfg->fec) {                          // tile boundaries on Hamming byte boundaries: BPS*s == 128 (mod 14)
        int s0 = -1;
        for (int s = 0; s < 7; s++) if (((h->bps_sym * s - 128) % 14 + 14) % 14 == 0) { s0 = s; break; }
        if (s0 < 0) { g_create_error = "no Hamming-aligned tiling for this carrier layout"; delete h; return OFDM_E_INVALID; }
        h->tile_shift = (7 - s0) % 7;
    }

    // tables (src/transmitter.rs:60-96) in f64, then fc32; nfft = 1024 uses the same generators with LEN = 1280 / 1024
    h->wide = cfg->nfft == 1024;
    h->nfft = (int)cfg->nfft;
    h->sym_len = cfg_symlen(cfg);
    const int NF = h->nfft, LS = h->sym_len, CPL = NF / 4;
    std::vector<ofdm_host::cd> lock(LS), pre(LS), train(NF), tsym(NF);
    ofdm_host::locking_signal(lock.data(), LS);
    ofdm_host::preamble(pre.data(), LS);
    ofdm_host::training_signals(train.data(), NF);
    if (cfg->locking) { h->lock_is_ramp = false; for (int i = 0; i < LS; i++) lock[i] = { cfg->locking[i].re, cfg->locking[i].im }; }
    if (cfg->preamble) for (int i = 0; i < LS; i++) pre[i] = { cfg->preamble[i].re, cfg->preamble[i].im };
    if (cfg->training) for (int i = 0; i < NF; i++) train[i] = { cfg->training[i].re, cfg->training[i].im };
    h->cfg.locking = h->cfg.preamble = h->cfg.training = nullptr;
    ofdm_host::idft(train.data(), tsym.data(), NF);
    h->lock.resize(LS); h->pre.resize(LS); h->train.resize(NF);
    std::vector<float2> head((size_t)10 * LS), invt(NF), lockf(LS);
    for (int i = 0; i < LS; i++) {
        h->lock[i] = { (float)lock[i].re, (float)lock[i].im };
        h->pre[i] = { (float)pre[i].re, (float)pre[i].im };
        lockf[i] = make_float2((float)lock[i].re, (float)lock[i].im);
        head[i] = lockf[i];
        for (int r = 0; r < 4; r++) head[(size_t)LS * (1 + r) + i] = make_float2((float)pre[i].re, (float)pre[i].im);
    }
    for (int i = 0; i < NF; i++) {
        h->train[i] = { (float)train[i].re, (float)train[i].im };
        double n2 = train[i].re * train[i].re + train[i].im * train[i].im;
        invt[i] = make_float2((float)(train[i].re / n2), (float)(-train[i].im / n2));
    }
    for (int r = 0; r < 5; r++)
        for (int i = 0; i < LS; i++) {
            const ofdm_host::cd &v = tsym[i < CPL ? NF - CPL + i : i - CPL];        // prefix_block, src/transmitter.rs:168-181
            head[(size_t)LS * (5 + r) + i] = make_float2((float)v.re, (float)v.im);
        }
    float mx = 0.0f;
    for (size_t i = 0; i < head.size(); i++) { mx = fmaxf(mx, head[i].x); mx = fmaxf(mx, head[i].y); }
    RxTables *t = new RxTables();
    wide::WideTables *wt = nullptr;
    if (!h->wide) {
        for (int i = 0; i < 80; i++) t->lock[i] = lockf[i];
        for (int i = 0; i < 64; i++) t->inv_training[i] = invt[i];
        for (int i = 0; i < 800; i++) t->head[i] = head[i];
        t->head_max = mx;
        for (int i = 0; i < 64; i++) t->w64[i] = h_w64[i];
    } else {
        wt = new wide::WideTables();
        for (int i = 0; i < LS; i++) wt->lock[i] = lockf[i];
        for (int i = 0; i < NF; i++) {
            wt->inv_training[i] = invt[i];
            const double ang = -2.0 * 3.14159265358979323846 * (double)i / (double)NF;
            wt->w1024[i] = make_float2((float)cos(ang), (float)sin(ang));
        }
        for (size_t i = 0; i < head.size(); i++) wt->head[i] = head[i];
        wt->head_max = mx;
        for (int i = 0; i < 64; i++) t->w64[i] = h_w64[i];
    }

    bool ok = cudaMalloc(&h->d_tables, sizeof(RxTables)) == cudaSuccess &&
              cudaMemcpy(h->d_tables, t, sizeof(RxTables), cudaMemcpyHostToDevice) == cudaSuccess &&
              (!wt || (cudaMalloc(&h->d_wtables, sizeof(wide::WideTables)) == cudaSuccess &&
                       cudaMemcpy(h->d_wtables, wt, sizeof(wide::WideTables), cudaMemcpyHostToDevice) == cudaSuccess)) &&
              cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_copied[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_copied[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_done[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&h->ev_done[1], cudaEventDisableTiming) == cudaSuccess &&
              h->counters.ensure(4 * sizeof(uint64_t)) == cudaSuccess;
    delete t;
    delete wt;
    if (!ok) {
        g_create_error = std::string("engine allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        ofdm_engine_destroy(h);
        return OFDM_E_CUDA;
    }
    *out = h;
    return 0;
}

extern "C" void ofdm_engine_destroy(ofdm_engine *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_tables) cudaFree(h->d_tables);
    if (h->d_wtables) cudaFree(h->d_wtables);
    DevBuf *bufs[] = { &h->state, &h->scratch_u32, &h->scratch_f32, &h->counters, &h->s_iq, &h->s_iq2, &h->s_bytes, &h->s_bytes2,
                       &h->s_len, &h->s_len2, &h->s_status, &h->s_aux, &h->s_points, &h->s_h, &h->sync_scratch, &h->cap_base, &h->rs_tables, &h->wtx_slots, &h->s_iq3, &h->s_plan, &h->s_moved };
    for (DevBuf *b : bufs) b->release();
    for (void *p : h->pin) if (p) cudaFreeHost(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    for (int i = 0; i < 2; i++) { if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]); }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) if (h->ev_feed[i][j]) cudaEventDestroy(h->ev_feed[i][j]);
    if (h->pin_plan) cudaFreeHost(h->pin_plan);
    for (int i = 0; i < 4; i++) { if (h->feed_streams[i]) cudaStreamDestroy(h->feed_streams[i]); if (h->ev_lane[i]) cudaEventDestroy(h->ev_lane[i]); }
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    delete h;
}

extern "C" const char *ofdm_last_error(const ofdm_engine *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int ofdm_engine_reserve(ofdm_engine *h, uint32_t max_streams, uint64_t max_capture_samples)
{
    if (!h) return OFDM_E_INVALID;
    CU(h, cudaSetDevice(h->device));
    if (max_streams) {
        CU(h, h->state.ensure((h->wide ? sizeof(wide::StreamStateW) : sizeof(StreamState)) * (size_t)max_streams));
        CU(h, h->scratch_u32.ensure(3 * sizeof(uint32_t) * (size_t)max_streams));
        CU(h, h->scratch_f32.ensure(5 * sizeof(double) * (size_t)max_streams));
        CU(h, h->cap_base.ensure((sizeof(uint64_t) + sizeof(uint32_t)) * (size_t)max_streams));
    }
    if (max_capture_samples >= 2 * (uint64_t)h->sym_len) {                 // same layout as sync_plan
        const uint64_t tl = h->wide ? (uint64_t)kWTD : (uint64_t)kScanD;          // the smaller of the two wide tile sizes
        const uint64_t T = (max_capture_samples - 2 * (uint64_t)h->sym_len + 1 + tl - 1) / tl;
        CU(h, h->sync_scratch.ensure(((32 + 4 * T + 7) & ~(size_t)7) + (size_t)T * kTileCand * 19 + 16));
    }
    return 0;
}

extern "C" int ofdm_get_tables(const ofdm_engine *h, ofdm_fc32 *locking80, ofdm_fc32 *preamble80, ofdm_fc32 *training64)
{
    if (!h) return OFDM_E_INVALID;
    if (locking80) memcpy(locking80, h->lock.data(), sizeof(ofdm_fc32) * h->lock.size());
    if (preamble80) memcpy(preamble80, h->pre.data(), sizeof(ofdm_fc32) * h->pre.size());
    if (training64) memcpy(training64, h->train.data(), sizeof(ofdm_fc32) * h->train.size());
    return 0;
}

extern "C" int ofdm_host_alloc(size_t bytes, void **out)
{
    if (!out) return OFDM_E_INVALID;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? 0 : OFDM_E_NOMEM;
}
extern "C" void ofdm_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" uint64_t ofdm_kernel_launches(const ofdm_engine *h) { return h ? h->launches : 0; }

// ---- TX ----------------------------------------------------------------------------------------------------------
static int tx_device(ofdm_engine *h, const uint8_t *payload, const uint32_t *payload_len, uint32_t payload_stride,
                     uint32_t n_streams, ofdm_fc32 *iq, uint32_t iq_stride, uint32_t *frame_len_out, cudaStream_t st)
{
    CU(h, h->scratch_u32.ensure(3 * sizeof(uint32_t) * (size_t)n_streams));
    uint32_t *d_flen = h->scratch_u32.as<uint32_t>();
    int *d_max = reinterpret_cast<int *>(d_flen + n_streams);
    uint32_t *d_cnt = d_flen + 2 * (size_t)n_streams;
    CU(h, cudaMemsetAsync(d_flen, 0, 3 * sizeof(uint32_t) * (size_t)n_streams, st));
    if (h->wide) {
        wide::WideTxArgs w{};
        w.payload = payload; w.payload_len = payload_len; w.payload_stride = payload_stride; w.n_streams = n_streams;
        w.iq = reinterpret_cast<float2 *>(iq); w.iq_stride = iq_stride; w.frame_len = d_flen; w.stream_max = d_max; w.tables = h->d_wtables;
        const long max_syms = (long)iq_stride / wide::kL - 10;
        // Large batches: ONE pass, frames resident in tensor memory (wide_tx_resident.cuh): groups of C persistent CTAs (one per SM,
        // 32 symbols each) per frame, launched cooperatively because the CTAs of a group wait for one another's frame maximum.
        // Speculative one pass (wide_tx_spec_kernel): bet that the frame maximum is the head's, store every symbol at once, redo the
        // frames that lost the bet with the two-pass kernel (its store pass with redo_only). Large-batch default.
        const int spw = wide_tx_resident_threads() / 32;                       // symbols per unit: one per warp
        const uint64_t wspec_units = max_syms > 0 ? (uint64_t)n_streams * (uint64_t)((max_syms + spw - 1) / spw) : 0;
        if (max_syms > 0 && h->n_sm > 0 && (h->tx_path == 5 || (h->tx_path == 0 && wspec_units >= 2ull * (uint64_t)h->n_sm))) {
            WTxKernel k = wpick_tx_spec(h->cfg);
            const size_t smem = wide_tx_resident_smem(h->cfg);
            if (h->smem_configured.insert((const void *)k).second)
                CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            w.group_ctas = (int32_t)((max_syms + spw - 1) / spw);
            w.stream_cnt = d_cnt;                                                // [0]: warps whose data beat the head maximum (zeroed above)
            k<<<(unsigned)(wspec_units < (uint64_t)h->n_sm ? wspec_units : (uint64_t)h->n_sm), wide_tx_resident_threads(), smem, st>>>(w);
            h->launches++;
            w.redo_only = 1;
            w.redo_tiles = (int32_t)((max_syms + 7) / 8);                        // tiles of 8 symbols a redone frame walks
            w.stream0 = 0;
            wpick_tx(h->cfg, true)<<<dim3(1, std::min<uint32_t>(n_streams, 2u * (uint32_t)h->n_sm)), wide::kThreads, 0, st>>>(w);
            h->launches++;
            CU(h, cudaGetLastError());
            if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
            return 0;
        }
        if (max_syms > 0 && h->n_sm > 0 && h->tx_path == 3) {
            bool db = false;
            if (const char *e = getenv("OFDM_WTX_DB")) db = atoi(e) != 0;
            const int per = wide_tx_resident_syms_per_cta(db);
            const int C = (int)((max_syms + per - 1) / per);
            int G = C <= h->n_sm ? h->n_sm / C : 0;
            if (G > 0 && (uint32_t)G > n_streams) G = (int)n_streams;
            if (G > 0 && (h->tx_path == 3 || n_streams >= 2u * (uint32_t)(h->n_sm / C))) {
                WTxKernel k = wpick_tx_resident(h->cfg, db);
                const size_t smem = wide_tx_resident_smem(h->cfg);
                if (h->smem_configured.insert((const void *)k).second)
                    CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                // one flagged word per warp of a group and frame (the frame maximum travels through them), zeroed per call
                const size_t slot_words = (size_t)n_streams * (size_t)C * (size_t)(wide_tx_resident_threads() / 32);
                CU(h, h->wtx_slots.ensure(sizeof(uint32_t) * slot_words));
                CU(h, cudaMemsetAsync(h->wtx_slots.p, 0, sizeof(uint32_t) * slot_words, st));
                w.stream_cnt = h->wtx_slots.as<uint32_t>(); w.group_ctas = C; w.n_groups = G;
                void *kargs[] = { (void *)&w };
                CU(h, cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)(G * C)), dim3((unsigned)wide_tx_resident_threads()), kargs, smem, st));
                h->launches++;
                CU(h, cudaGetLastError());
                if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
                return 0;
            }
        }
        const uint32_t tiles = max_syms > 0 ? (uint32_t)((max_syms + 7) / 8) : 1;
        launch_streams(wpick_tx(h->cfg, false), w, tiles, n_streams, wide::kThreads, 0, st, h->launches);
        launch_streams(wpick_tx(h->cfg, true), w, tiles, n_streams, wide::kThreads, 0, st, h->launches);
        CU(h, cudaGetLastError());
        if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    TxArgs a{};
    a.payload = payload; a.payload_len = payload_len; a.payload_stride = payload_stride; a.n_streams = n_streams;
    a.iq = reinterpret_cast<float2 *>(iq); a.iq_stride = iq_stride; a.frame_len = d_flen; a.stream_max = d_max; a.tables = h->d_tables;
    a.tile_shift = h->tile_shift;
    const long max_syms = (long)iq_stride / 80 - 10;
    // Large batches: ONE pass with the frames resident on chip (tx_resident.cuh). A frame is shared by a group of C persistent
    // CTAs (one per SM), its un-normalised symbols wait in tensor memory for the frame maximum, and the stores of frame k-1
    // overlap the transforms of frame k in every warp. C = the smallest group whose rings hold the longest frame iq_stride admits.
    // Speculative one pass (tx_spec_kernel, tx_warp.cuh): bet that the frame maximum is the head's (true for scrambled payloads),
    // scale and store every symbol at once, then redo the frames that lost the bet with the store pass of the two-pass kernel.
    // Large-batch default (from two units of 512 symbols per SM on): 4096 frames 1.15 ms against 1.29 ms for the exact one-pass
    // kernel below, which costs the same whatever the payloads are (OFDM_TX_PATH=warp) -- a frame that loses the bet costs a
    // second store pass.
    const uint64_t spec_units = max_syms > 0 ? (uint64_t)n_streams * (uint64_t)((max_syms + tx_warp_syms_per_cta() - 1) / tx_warp_syms_per_cta()) : 0;
    if (max_syms > 0 && h->n_sm > 0 && (h->tx_path == 5 || (h->tx_path == 0 && spec_units >= 2ull * (uint64_t)h->n_sm))) {
        const int per = tx_warp_syms_per_cta();
        const uint32_t U = (uint32_t)((max_syms + per - 1) / per);
        const uint64_t units = (uint64_t)n_streams * U;
        TxKernel k = pick_tx_spec(h->cfg);
        const size_t smem = tx_warp_smem(h->cfg);
        if (h->smem_configured.insert((const void *)k).second)
            CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        a.group_ctas = (int32_t)U;
        a.stream_cnt = d_cnt;                                                  // [0]: warps whose data beat the head maximum (zeroed above)
        k<<<(unsigned)(units < (uint64_t)h->n_sm ? units : (uint64_t)h->n_sm), tx_warp_threads(), smem, st>>>(a);
        h->launches++;
        uint32_t tiles = (uint32_t)((max_syms + h->tile_shift + kTxTileSyms - 1) / kTxTileSyms);
        a.tiles_per_cta = (int)tiles;                                          // one CTA per frame: it exits at once unless the frame has to be redone
        a.redo_only = 1;
        const size_t tx_smem = sizeof(float2) * kTxWarps * kTrWarp + (size_t)kTxTileSyms * h->dcar + 64 + sizeof(float2) * 16 * ((1u << h->bpc) + 2) + 16 + 512;
        TxKernel k2 = pick_tx(h->cfg, true);
        if (h->smem_configured.insert((const void *)k2).second)
            CU(h, cudaFuncSetAttribute((const void *)k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tx_smem));
        a.stream0 = 0;
        k2<<<dim3(1, std::min<uint32_t>(n_streams, 4u * (uint32_t)h->n_sm)), kTxThreads, tx_smem, st>>>(a);
        h->launches++;
        CU(h, cudaGetLastError());
        if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    // ... second version of it (tx_warp.cuh): a warp owns 16 consecutive symbols and prepares their carrier bytes itself, the
    // frame maximum travels through one flagged word per warp -- no CTA barrier in the frame loop.
    if (max_syms > 0 && h->n_sm > 0 && h->tx_path == 4) {
        const int per = tx_warp_syms_per_cta();
        const int C = (int)((max_syms + per - 1) / per);
        int G = C <= h->n_sm ? h->n_sm / C : 0;
        if (G > 0 && (uint32_t)G > n_streams) G = (int)n_streams;
        if (G > 0 && (h->tx_path == 4 || n_streams >= 2u * (uint32_t)(h->n_sm / C))) {
            TxKernel k = pick_tx_warp(h->cfg);
            const size_t smem = tx_warp_smem(h->cfg);
            if (h->smem_configured.insert((const void *)k).second)
                CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const size_t slot_words = (size_t)n_streams * (size_t)C;              // one flagged word per CTA of a group and frame
            CU(h, h->wtx_slots.ensure(sizeof(uint32_t) * slot_words));
            CU(h, cudaMemsetAsync(h->wtx_slots.p, 0, sizeof(uint32_t) * slot_words, st));
            a.stream_cnt = h->wtx_slots.as<uint32_t>(); a.group_ctas = C; a.n_groups = G;
            void *kargs[] = { (void *)&a };
            CU(h, cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)(G * C)), dim3((unsigned)tx_warp_threads()), kargs, smem, st));
            h->launches++;
            CU(h, cudaGetLastError());
            if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
            return 0;
        }
    }
    if (max_syms > 0 && h->n_sm > 0 && h->tx_path == 3) {
        const int W = kTrsWarpsPerCta;                                        // warps per CTA; 32 / W CTAs per SM
        const int chunk_max = 7 * ((16 * W) / 7);
        const int C = (int)((max_syms + h->tile_shift + chunk_max - 1) / chunk_max);
        const int ctas = (32 / W) * h->n_sm;
        int G = C <= ctas ? ctas / C : 0;
        if (G > 0 && (uint32_t)G > n_streams) G = (int)n_streams;
        // worth it once every group has two frames to pipeline (measured: ahead of the two-pass kernel from 2 frames per group on,
        // 0.050 vs 0.053 ms for 74 frames of 163 840 samples, 0.87 vs 1.02 ms for 2368); smaller batches take the cluster kernel
        if (G > 0 && (h->tx_path == 3 || n_streams >= 2u * (uint32_t)(ctas / C))) {
            TxKernel k = pick_tx_resident(h->cfg);
            const size_t smem = tx_resident_smem(h->cfg);
            if (h->smem_configured.insert((const void *)k).second)
                CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            a.stream_cnt = d_cnt; a.group_ctas = C; a.n_groups = G;
            // The CTAs of a group wait for one another (frame maximum through global memory), so the whole grid has to be
            // resident at once: a cooperative launch is what guarantees that -- the grid (<= one CTA per SM) is scheduled only
            // when all of it fits, also next to other work and next to a second kernel of this kind from another handle.
            void *kargs[] = { (void *)&a };
            CU(h, cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)(G * C)), dim3(32 * W), kargs, smem, st));
            h->launches++;
            CU(h, cudaGetLastError());
            if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
            return 0;
        }
    }
    // Small batches (the reference's one-frame `encode` call): one pass, a frame per thread-block cluster (<= 16 CTAs x 133
    // symbols held in shared memory, every symbol transformed once and written once), as long as all frames of the call are
    // in flight as ONE wave of clusters -- measured 17 vs 22 us for 1..8 frames of 163 840 samples. For larger batches the
    // cluster barrier keeps the two CTAs of an SM in the same phase (transform, then store) and the two-pass kernel wins
    // (1.77 vs 2.69 ms for 4096 frames): its store pass overlaps transforms and stores inside every SM.
    if (h->tx_path != 1 && max_syms > 0 && max_syms + h->tile_shift <= (long)kTxfMaxCluster * kTxfChunk) {
        unsigned cs = 1;
        while ((long)cs * kTxfChunk < max_syms + h->tile_shift) cs <<= 1;
        TxKernel k = pick_tx_frame(h->cfg);
        const size_t smem = sizeof(float2) * (kTxfChunk * kNfft + kTxfWarps * kTrWarp) + (size_t)kTxfIters * 4 * kTxfWarps * h->dcar + 64 +
                            sizeof(float2) * 16 * ((1u << h->bpc) + 2) + 16 + 512;
        if (h->smem_configured.insert((const void *)k).second) {
            CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        cudaLaunchConfig_t lc{};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.blockDim = dim3(kTxfThreads); lc.dynamicSmemBytes = smem; lc.stream = st; lc.attrs = at; lc.numAttrs = 1;
        lc.gridDim = dim3(cs, n_streams < 65535u ? n_streams : 65535u);
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, (const void *)k, &lc) == cudaSuccess && max_clusters > 0 && (n_streams <= (uint32_t)max_clusters || h->tx_path == 2)) {
            for (uint32_t s0 = 0; s0 < n_streams; s0 += 65535u) {
                a.stream0 = s0;
                lc.gridDim = dim3(cs, n_streams - s0 < 65535u ? n_streams - s0 : 65535u);
                CU(h, cudaLaunchKernelEx(&lc, k, a));
                h->launches++;
            }
            CU(h, cudaGetLastError());
            if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
            return 0;
        }
        (void)cudaGetLastError();                                          // more frames than one wave of clusters (or the shape cannot be scheduled): two passes
    }
    // two passes over the same tiles: maximum for `normalize`, then recompute + store once (8 B/sample written)
    uint32_t tiles = max_syms > 0 ? (uint32_t)((max_syms + h->tile_shift + kTxTileSyms - 1) / kTxTileSyms) : 1;
    // several consecutive tiles per CTA (tables are built once per CTA), but keep >= ~16 waves of CTAs (148 SMs x 4 CTAs)
    {
        uint32_t tpc = (uint32_t)(((uint64_t)tiles * n_streams) / (16u * 148u * 4u));
        if (tpc < 1) tpc = 1;
        if (tpc > tiles) tpc = tiles;
        tpc = (tiles + (tiles + tpc - 1) / tpc - 1) / ((tiles + tpc - 1) / tpc);
        a.tiles_per_cta = (int)tpc;
        tiles = (tiles + tpc - 1) / tpc;
    }
    const size_t tx_smem = sizeof(float2) * kTxWarps * kTrWarp + (size_t)kTxTileSyms * h->dcar + 64 + sizeof(float2) * 16 * ((1u << h->bpc) + 2) + 16 + 512;
    // (Running the maximum pass of frame batch i+1 concurrently with the store pass of batch i on two streams, each kernel held
    // to two CTAs per SM, was measured: 1.82 vs 1.77 ms -- together the passes are bound by the shared-memory pipe and the issue
    // slots, which both of them use for the same transforms, so overlapping them buys nothing.)
    for (int pass = 0; pass < 2; pass++) {
        TxKernel k = pick_tx(h->cfg, pass == 1);
        if (h->smem_configured.insert((const void *)k).second)
            CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tx_smem));
        launch_streams(k, a, tiles, n_streams, kTxThreads, tx_smem, st, h->launches);
    }
    CU(h, cudaGetLastError());
    if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int ofdm_tx_encode_batch(ofdm_engine *h, const uint8_t *payload, const uint32_t *payload_len,
                                    uint32_t payload_stride, uint32_t n_streams,
                                    ofdm_fc32 *iq_out, uint32_t iq_stride, uint32_t *frame_len_out,
                                    int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!payload || !payload_len || !iq_out || n_streams == 0 || iq_stride < 11u * (uint32_t)h->sym_len) ENG_FAIL(h, OFDM_E_INVALID, "tx: bad arguments");
    CU(h, cudaSetDevice(h->device));
    if (mem == OFDM_MEM_DEVICE) return tx_device(h, payload, payload_len, payload_stride, n_streams, iq_out, iq_stride, frame_len_out, (cudaStream_t)stream);

    for (uint32_t s = 0; s < n_streams; s++) {
        if (payload_len[s] > payload_stride) ENG_FAIL(h, OFDM_E_INVALID, "tx: payload_len[%u] exceeds payload_stride", s);
        if (ofdm_frame_len(&h->cfg, payload_len[s]) > iq_stride) ENG_FAIL(h, OFDM_E_INVALID, "tx: frame %u does not fit iq_stride", s);
    }
    cudaStream_t st = h->own_stream;
    size_t pb = (size_t)n_streams * payload_stride, ib = (size_t)n_streams * iq_stride * sizeof(float2);
    CU(h, h->s_bytes.ensure(pb));
    CU(h, h->s_len.ensure(2 * sizeof(uint32_t) * (size_t)n_streams));
    CU(h, h->s_iq.ensure(ib));
    uint32_t *d_len = h->s_len.as<uint32_t>(), *d_flen = d_len + n_streams;
    CU(h, cudaMemcpyAsync(h->s_bytes.p, payload, pb, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_len, payload_len, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyHostToDevice, st));
    int rc = tx_device(h, h->s_bytes.as<uint8_t>(), d_len, payload_stride, n_streams, h->s_iq.as<ofdm_fc32>(), iq_stride, d_flen, st);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(iq_out, h->s_iq.p, ib, cudaMemcpyDeviceToHost, st));
    if (frame_len_out) CU(h, cudaMemcpyAsync(frame_len_out, d_flen, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

// ---- RX ----------------------------------------------------------------------------------------------------------
static int rx_device(ofdm_engine *h, const ofdm_fc32 *iq, const uint32_t *n_samples, uint32_t n_streams, uint32_t iq_stride,
                     uint32_t max_n_samples, uint8_t *out, uint32_t out_stride, uint32_t *out_len, int32_t *status,
                     const ofdm_rx_diag *diag, cudaStream_t st, size_t state_offset = 0, size_t state_total = 0,
                     const uint64_t *stream_base = nullptr, int phase = 0)
{
    // phase: 0 = acquisition + decode; 1 = acquisition only, 2 = decode only (the host path runs them apart: the data symbols
    // are fetched from the host only once the acquisition has said where they are)
    if (state_total < n_streams) state_total = n_streams;
    const size_t state_sz = h->wide ? sizeof(wide::StreamStateW) : sizeof(StreamState);
    if (state_offset == 0) CU(h, h->state.ensure(state_sz * state_total));
    const bool prof = h->prof_n < h->prof_cap;
    cudaEvent_t *pe = prof ? &h->prof_ev[3 * (size_t)h->prof_n] : nullptr;
    if (prof) { h->prof_n++; CU(h, cudaEventRecord(pe[0], st)); }
    if (h->wide) {
        wide::WideRxArgs w{};
        w.iq = reinterpret_cast<const float2 *>(iq); w.n_samples = n_samples; w.iq_stride = iq_stride; w.n_streams = n_streams;
        w.state = h->state.as<wide::StreamStateW>() + state_offset; w.tables = h->d_wtables;
        w.out = out; w.out_stride = out_stride; w.out_len = out_len; w.status = status;
        w.sync_window = h->cfg.sync_window; w.tile_shift = h->tile_shift;
        w.sync_mode = stream_base ? 2 : (int)h->cfg.sync_mode; w.cfo_mode = (int)h->cfg.cfo_mode; w.fec = (int)h->cfg.fec;
        w.lock_is_ramp = h->lock_is_ramp ? 1 : 0;
        w.stream_base = stream_base;
        bool wpoints = false;
        if (diag) {
            w.d_offset = diag->offset; w.d_f_delta = diag->f_delta; w.d_h = reinterpret_cast<float2 *>(diag->h_k);
            w.d_nsyms = diag->n_data_syms; w.d_points = reinterpret_cast<float2 *>(diag->points); w.points_stride = diag->points_stride;
            wpoints = diag->points != nullptr && diag->points_stride > 0;
        }
        WDecodeKernel ka = wpick_acquire(h->cfg);
        if (h->smem_configured.insert((const void *)ka).second)
            CU(h, cudaFuncSetAttribute((const void *)ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide::wide_acquire_smem()));
        if (phase != 2) { ka<<<n_streams, wide::kThreads, wide::wide_acquire_smem(), st>>>(w); h->launches += 1; }
        if (prof) CU(h, cudaEventRecord(pe[1], st));
        uint32_t mxs = max_n_samples ? max_n_samples : iq_stride;
        if (mxs > iq_stride) mxs = iq_stride;
        const long S = ((long)mxs + wide::kL - 1) / wide::kL - 10;
        if (S > 0 && phase != 1) {
            uint32_t tiles = (uint32_t)((S + h->tile_shift + wide::kTileSymsW - 1) / wide::kTileSymsW);
            // several consecutive tiles per CTA (per-stream tables and the prefetch pipeline are reused), but keep >= ~8 waves
            // of CTAs (148 SMs x 2 CTAs) and split a stream's tiles evenly over its CTAs
            uint32_t tpc = (uint32_t)(((uint64_t)tiles * n_streams) / (8u * 148u * 2u));
            if (tpc < 1) tpc = 1;
            if (tpc > tiles) tpc = tiles;
            tpc = (tiles + (tiles + tpc - 1) / tpc - 1) / ((tiles + tpc - 1) / tpc);
            w.tiles_per_cta = (int)tpc;
            tiles = (tiles + tpc - 1) / tpc;
            WDecodeKernel kd = wpick_decode(h->cfg, wpoints);
            const size_t smem = wide::wide_decode_smem(h->cfg.guard_bands != 0);
            if (h->smem_configured.insert((const void *)kd).second)
                CU(h, cudaFuncSetAttribute((const void *)kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            launch_streams(kd, w, tiles, n_streams, wide::kWDecThreads, smem, st, h->launches);
        }
        if (prof) CU(h, cudaEventRecord(pe[2], st));
        CU(h, cudaGetLastError());
        return 0;
    }
    RxArgs a{};
    a.iq = reinterpret_cast<const float2 *>(iq); a.n_samples = n_samples; a.iq_stride = iq_stride; a.n_streams = n_streams;
    a.state = h->state.as<StreamState>() + state_offset; a.tables = h->d_tables;
    a.out = out; a.out_stride = out_stride; a.out_len = out_len; a.status = status;
    a.sync_window = h->cfg.sync_window; a.tile_shift = h->tile_shift;
    a.sync_mode = stream_base ? 2 : (int)h->cfg.sync_mode; a.cfo_mode = (int)h->cfg.cfo_mode; a.fec = (int)h->cfg.fec;
    a.stream_base = stream_base;
    bool points = false;
    if (diag) {
        a.d_offset = diag->offset; a.d_f_delta = diag->f_delta; a.d_h = reinterpret_cast<float2 *>(diag->h_k);
        a.d_nsyms = diag->n_data_syms; a.d_points = reinterpret_cast<float2 *>(diag->points); a.points_stride = diag->points_stride;
        points = diag->points != nullptr && diag->points_stride > 0;
    }
    AcquireKernel kacq = pick_acquire(h->cfg);
    if (h->smem_configured.insert((const void *)kacq).second)       // kAcq64Ctas CTAs x ~19 kB static smem must fit the carve-out
        CU(h, cudaFuncSetAttribute((const void *)kacq, cudaFuncAttributePreferredSharedMemoryCarveout, 80));
    if (phase != 2) { kacq<<<n_streams, kAcq64Threads, 0, st>>>(a); h->launches += 1; }
    if (prof) CU(h, cudaEventRecord(pe[1], st));
    uint32_t mx = max_n_samples ? max_n_samples : iq_stride;
    if (mx > iq_stride) mx = iq_stride;
    long rows = ((long)mx + 79) / 80, S = rows - 10;
    if (S > 0 && phase != 1) {
        uint32_t tiles = (uint32_t)((S + h->tile_shift + kTileSyms - 1) / kTileSyms);
        // several consecutive tiles per CTA (lane constants and the prefetch pipeline are reused across them), but keep
        // >= ~32 waves of CTAs (148 SMs x 4 CTAs) so the tail stays small, and split a stream's tiles evenly over its CTAs
        uint32_t tpc = (uint32_t)(((uint64_t)tiles * n_streams) / (32u * 148u * 4u));
        if (tpc < 1) tpc = 1;
        if (tpc > tiles) tpc = tiles;
        if (tpc > (uint32_t)kDecBaseTiles) tpc = kDecBaseTiles;        // start phasors of a CTA's tiles are tabulated in smem
        tpc = (tiles + (tiles + tpc - 1) / tpc - 1) / ((tiles + tpc - 1) / tpc);      // ceil(tiles / ceil(tiles / tpc)): balanced
        a.tiles_per_cta = (int)tpc;
        tiles = (tiles + tpc - 1) / tpc;
        DecodeKernel k = pick_decode(h->cfg, points);
        const size_t smem = h->cfg.guard_bands ? rx_decode_smem_bytes<true>() : rx_decode_smem_bytes<false>();
        if (h->smem_configured.insert((const void *)k).second)
            CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_streams(k, a, tiles, n_streams, kDecThreads, smem, st, h->launches);
    }
    if (prof) CU(h, cudaEventRecord(pe[2], st));
    CU(h, cudaGetLastError());
    return 0;
}

extern "C" int ofdm_profile_begin(ofdm_engine *h, uint32_t max_calls)
{
    if (!h) return OFDM_E_INVALID;
    CU(h, cudaSetDevice(h->device));
    while (h->prof_ev.size() < 3 * (size_t)max_calls) {
        cudaEvent_t e;
        CU(h, cudaEventCreate(&e));
        h->prof_ev.push_back(e);
    }
    h->prof_cap = max_calls;
    h->prof_n = 0;
    return 0;
}

extern "C" int ofdm_profile_read(ofdm_engine *h, float *acquire_ms, float *decode_ms, uint32_t *n_calls)
{
    if (!h || !n_calls) return OFDM_E_INVALID;
    CU(h, cudaSetDevice(h->device));
    uint32_t n = h->prof_n;
    for (uint32_t i = 0; i < n; i++) {
        cudaEvent_t *pe = &h->prof_ev[3 * (size_t)i];
        CU(h, cudaEventSynchronize(pe[2]));
        float a = 0, d = 0;
        CU(h, cudaEventElapsedTime(&a, pe[0], pe[1]));
        CU(h, cudaEventElapsedTime(&d, pe[1], pe[2]));
        if (acquire_ms) acquire_ms[i] = a;
        if (decode_ms) decode_ms[i] = d;
    }
    *n_calls = n;
    h->prof_cap = 0;
    h->prof_n = 0;
    return 0;
}

// per stream: where the acquisition put the frame and how many data symbols the header asks for (0 unless status OK)
template <class State>
__global__ void rx_plan_kernel(const State *__restrict__ state, uint32_t n, uint32_t *__restrict__ plan)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = state[i].status == ST_OK;
    plan[2 * i] = ok ? (uint32_t)state[i].offset : 0u;
    plan[2 * i + 1] = ok ? state[i].n_syms : 0u;
}

// Pull the useful samples of the wanted data symbols of a chunk of streams straight out of (pinned, device-mapped) HOST memory:
// the SMs issue the PCIe reads themselves -- thousands of 8-byte loads in flight, one 512-byte (nfft 64) row per two warp
// requests -- so neither a copy-engine descriptor per stream nor a trip of the plan to the host is needed. Only samples
// below n_samples exist (a last, cut-off symbol is fetched as far as it goes). blockIdx.y = stream, blockIdx.x strides its rows.
template <int N>
__global__ void __launch_bounds__(256) rx_gather_kernel(const float2 *__restrict__ host_iq, float2 *__restrict__ dev, const uint32_t *__restrict__ plan,
                                                        const uint32_t *__restrict__ n_samples, uint32_t iq_stride, uint32_t head_len,
                                                        unsigned long long *__restrict__ moved)
{
    constexpr uint32_t L = N + N / 4, CP = N / 4;
    const uint32_t i = blockIdx.y;
    const uint64_t off = plan[2 * i], want = plan[2 * i + 1], M = n_samples[i];
    if (want == 0) return;
    uint64_t r0 = head_len > off ? (head_len - off) / L : 0;                     // data symbols below r0 lie wholly inside the head copy
    r0 = r0 > 10 ? r0 - 10 : 0;
    if (r0 >= want) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                   // bytes this stream's CTAs pull (for e2e.h2d_bytes_per_step)
        uint64_t r_full = M >= off + L ? (M - off) / L : 0;
        r_full = r_full > 10 ? r_full - 10 : 0;
        if (r_full > want) r_full = want;
        uint64_t n = r_full > r0 ? (r_full - r0) * N : 0;
        const uint64_t rp = r_full > r0 ? r_full : r0;
        if (rp < want) { const uint64_t t = off + (10 + rp) * L + CP; if (t < M) n += M - t < N ? M - t : N; }
        atomicAdd(moved, (unsigned long long)(n * sizeof(float2)));
    }
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(host_iq + (size_t)i * iq_stride);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(dev + (size_t)i * iq_stride);
    const uint64_t total = (want - r0) * N;
    constexpr int U = 8;
    for (uint64_t e0 = (uint64_t)blockIdx.x * (256 * U) + threadIdx.x; e0 < total; e0 += (uint64_t)gridDim.x * (256 * U)) {
        unsigned long long v[U];
        uint64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint64_t e = e0 + (uint64_t)u * 256;
            idx[u] = off + (10 + r0 + e / N) * L + CP + (e % N);
            v[u] = 0ull;
            if (e < total && idx[u] < M) v[u] = src[idx[u]];
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (e0 + (uint64_t)u * 256 < total && idx[u] < M) dst[idx[u]] = v[u];
    }
}

// Host path of ofdm_rx_decode_batch that never moves a cyclic prefix over PCIe. `unprefix_block` (src/receiver.rs:104-118)
// throws the prefix of every data symbol away, and past the frame head the prefixes are a fifth of the capture -- but where
// they are is only known after synchronisation. So per chunk of streams: (1) ONE 2-D copy brings the first
// sync_window + 13 symbols of every stream (everything the acquisition can touch: search window, refinement, CFO rows,
// training rows, header symbol); (2) the acquisition kernel runs on that, and (offset, data symbols wanted) of every stream
// goes back to the host; (3) per stream ONE 2-D copy (width nfft, pitch nfft + cp samples) fetches the useful part of exactly
// the data symbols the header asks for, to the very place they have in the capture -- the device image is the capture with
// holes where the prefixes (and whatever follows the frame) would be, so the decode kernel runs on it unchanged; (4) decode,
// payload bytes back on the third stream. The head copy of chunk c + 1 is queued before the symbol copies of chunk c, so the
// copy engine is never idle while the host waits for an acquisition; three staging buffers rotate.
static int rx_host_skip_cp(ofdm_engine *h, const ofdm_fc32 *iq, const uint32_t *n_samples, uint32_t n_streams, uint32_t iq_stride, uint32_t mx,
                           uint32_t head_len, uint8_t *out, uint32_t out_stride, const ofdm_rx_diag *dd, bool have_diag,
                           uint32_t *d_ns, uint32_t *d_ol, const ofdm_fc32 *iq_mapped)
{
    // iq_mapped: device-side address of the (pinned) host capture, or NULL -- then step (3) is one 2-D copy per stream issued by
    // the host after it has read the plan, instead of the gather kernel
    cudaStream_t st = h->own_stream, cs = h->copy_stream, os = h->d2h_stream;
    const size_t stream_bytes = (size_t)iq_stride * sizeof(float2);
    const uint32_t L = (uint32_t)h->sym_len, N = (uint32_t)h->nfft, CP = L - N;
    uint32_t chunk = (uint32_t)((128u << 20) / stream_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > n_streams) chunk = n_streams;
    DevBuf *stage[3] = { &h->s_iq, &h->s_iq2, &h->s_iq3 };
    for (DevBuf *b : stage) CU(h, b->ensure(chunk * stream_bytes));
    CU(h, h->s_plan.ensure(3 * 2 * sizeof(uint32_t) * (size_t)chunk));
    CU(h, h->s_moved.ensure(16));
    CU(h, cudaMemsetAsync(h->s_moved.p, 0, 16, cs));
    if (h->pin_plan_cap < 3 * 2 * sizeof(uint32_t) * (size_t)chunk) {
        if (h->pin_plan) cudaFreeHost(h->pin_plan);
    for (int i = 0; i < 4; i++) { if (h->feed_streams[i]) cudaStreamDestroy(h->feed_streams[i]); if (h->ev_lane[i]) cudaEventDestroy(h->ev_lane[i]); }
        h->pin_plan = nullptr; h->pin_plan_cap = 0;
        CU(h, cudaHostAlloc(&h->pin_plan, 3 * 2 * sizeof(uint32_t) * (size_t)chunk, cudaHostAllocDefault));
        h->pin_plan_cap = 3 * 2 * sizeof(uint32_t) * (size_t)chunk;
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        if (!h->ev_feed[i][j]) CU(h, cudaEventCreateWithFlags(&h->ev_feed[i][j], cudaEventDisableTiming));
    const uint32_t n_chunks = (n_streams + chunk - 1) / chunk;
    uint64_t moved = 0;
    double submit_s = 0.0;
    int n_lanes = 4;
    if (const char *e = getenv("OFDM_RX_FEED_LANES")) { n_lanes = atoi(e); n_lanes = n_lanes < 1 ? 1 : (n_lanes > 4 ? 4 : n_lanes); }
    for (int i = 0; i < n_lanes && !iq_mapped; i++) {
        if (!h->feed_streams[i]) CU(h, cudaStreamCreateWithFlags(&h->feed_streams[i], cudaStreamNonBlocking));
        if (!h->ev_lane[i]) CU(h, cudaEventCreateWithFlags(&h->ev_lane[i], cudaEventDisableTiming));
    }

    auto diag_at = [&](uint32_t s0) {
        ofdm_rx_diag dc = *dd;
        if (dc.offset) dc.offset += s0;
        if (dc.f_delta) dc.f_delta += s0;
        if (dc.n_data_syms) dc.n_data_syms += s0;
        if (dc.h_k) dc.h_k += (size_t)s0 * h->nfft;
        if (dc.points) dc.points += (size_t)s0 * dc.points_stride;
        return dc;
    };
    // (1) + (2) of chunk c
    auto feed_head = [&](uint32_t c) -> int {
        const uint32_t s0 = c * chunk, ns = n_streams - s0 < chunk ? n_streams - s0 : chunk;
        const int b = (int)(c % 3);
        if (c >= 3) CU(h, cudaStreamWaitEvent(cs, h->ev_feed[b][2], 0));         // staging buffer free again (decode of chunk c - 3)
        CU(h, cudaMemcpy2DAsync(stage[b]->p, stream_bytes, iq + (size_t)s0 * iq_stride, stream_bytes, (size_t)head_len * sizeof(float2), ns,
                                cudaMemcpyHostToDevice, cs));
        moved += (uint64_t)ns * head_len * sizeof(float2);
        CU(h, cudaEventRecord(h->ev_feed[b][0], cs));
        CU(h, cudaStreamWaitEvent(st, h->ev_feed[b][0], 0));
        ofdm_rx_diag dc = diag_at(s0);
        int rc = rx_device(h, stage[b]->as<ofdm_fc32>(), d_ns + s0, ns, iq_stride, mx, h->s_bytes.as<uint8_t>() + (size_t)s0 * out_stride,
                           out_stride, d_ol + s0, h->s_status.as<int32_t>() + s0, have_diag ? &dc : nullptr, st, s0, n_streams, nullptr, 1);
        if (rc) return rc;
        uint32_t *d_plan = h->s_plan.as<uint32_t>() + (size_t)b * 2 * chunk;
        if (h->wide) rx_plan_kernel<<<(ns + 127) / 128, 128, 0, st>>>(h->state.as<wide::StreamStateW>() + s0, ns, d_plan);
        else rx_plan_kernel<<<(ns + 127) / 128, 128, 0, st>>>(h->state.as<StreamState>() + s0, ns, d_plan);
        h->launches++;
        CU(h, cudaMemcpyAsync(reinterpret_cast<uint32_t *>(h->pin_plan) + (size_t)b * 2 * chunk, d_plan, 2 * sizeof(uint32_t) * (size_t)ns, cudaMemcpyDeviceToHost, st));
        CU(h, cudaEventRecord(h->ev_feed[b][1], st));
        return 0;
    };
    if (int rc = feed_head(0)) return rc;
    for (uint32_t c = 0; c < n_chunks; c++) {
        const uint32_t s0 = c * chunk, ns = n_streams - s0 < chunk ? n_streams - s0 : chunk;
        const int b = (int)(c % 3);
        if (c + 1 < n_chunks) { if (int rc = feed_head(c + 1)) return rc; }
        if (iq_mapped) {
            CU(h, cudaStreamWaitEvent(cs, h->ev_feed[b][1], 0));
            const uint32_t *d_plan = h->s_plan.as<uint32_t>() + (size_t)b * 2 * chunk;
            const float2 *src = reinterpret_cast<const float2 *>(iq_mapped) + (size_t)s0 * iq_stride;
            const dim3 grid(8, ns);
            if (h->wide) rx_gather_kernel<1024><<<grid, 256, 0, cs>>>(src, stage[b]->as<float2>(), d_plan, d_ns + s0, iq_stride, head_len, h->s_moved.as<unsigned long long>());
            else rx_gather_kernel<64><<<grid, 256, 0, cs>>>(src, stage[b]->as<float2>(), d_plan, d_ns + s0, iq_stride, head_len, h->s_moved.as<unsigned long long>());
            h->launches++;
        } else
            CU(h, cudaEventSynchronize(h->ev_feed[b][1]));
        // (3) the useful samples of the data symbols that are not already inside the head copy
        const uint32_t *plan = reinterpret_cast<const uint32_t *>(h->pin_plan) + (size_t)b * 2 * chunk;
        if (!iq_mapped) {                                                        // the lanes start once the head copy (and the buffer) is theirs
            CU(h, cudaEventRecord(h->ev_copied[b & 1], cs));
            for (int l = 0; l < n_lanes; l++) CU(h, cudaStreamWaitEvent(h->feed_streams[l], h->ev_copied[b & 1], 0));
        }
        const auto t_sub0 = std::chrono::steady_clock::now();
        for (uint32_t i = 0; i < ns && !iq_mapped; i++) {
            cudaStream_t cs = h->feed_streams[i % (uint32_t)n_lanes];
            const uint64_t off = plan[2 * i], want = plan[2 * i + 1], M = n_samples[s0 + i];
            if (want == 0) continue;
            uint64_t r0 = head_len > off ? (head_len - off) / L : 0;             // data symbols r < r0 - 10 lie wholly inside the head copy
            r0 = r0 > 10 ? r0 - 10 : 0;
            // symbol r is whole in the capture when off + (10 + r) L + L <= M ... only its useful part [CP, L) matters: off + (10 + r) L + L <= M
            uint64_t r_full = M >= off + L ? (M - off) / L : 0;                  // rows (of L samples, counted from the offset) that are whole
            r_full = r_full > 10 ? r_full - 10 : 0;
            if (r_full > want) r_full = want;
            const float2 *src = reinterpret_cast<const float2 *>(iq) + (size_t)(s0 + i) * iq_stride;
            float2 *dst = stage[b]->as<float2>() + (size_t)i * iq_stride;
            if (r_full > r0) {
                const uint64_t t = off + (10 + r0) * L + CP;
                CU(h, cudaMemcpy2DAsync(dst + t, (size_t)L * sizeof(float2), src + t, (size_t)L * sizeof(float2), (size_t)N * sizeof(float2), (size_t)(r_full - r0),
                                        cudaMemcpyHostToDevice, cs));
                moved += (r_full - r0) * N * sizeof(float2);
            }
            const uint64_t rp = r_full > r0 ? r_full : r0;                       // a last, cut-off symbol (zero-padded tail row, src/receiver.rs:206-210)
            if (rp < want) {
                const uint64_t t = off + (10 + rp) * L + CP;
                if (t < M) {
                    CU(h, cudaMemcpyAsync(dst + t, src + t, (size_t)(M - t) * sizeof(float2), cudaMemcpyHostToDevice, cs));
                    moved += (M - t) * sizeof(float2);
                }
            }
        }
        submit_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_sub0).count();
        if (!iq_mapped)
            for (int l = 0; l < n_lanes; l++) {
                CU(h, cudaEventRecord(h->ev_lane[l], h->feed_streams[l]));
                CU(h, cudaStreamWaitEvent(st, h->ev_lane[l], 0));
                CU(h, cudaStreamWaitEvent(cs, h->ev_lane[l], 0));                // later head copies into this buffer come after its symbol copies
            }
        CU(h, cudaEventRecord(h->ev_copied[b & 1], cs));
        CU(h, cudaStreamWaitEvent(st, h->ev_copied[b & 1], 0));
        // (4) decode, payload bytes back
        ofdm_rx_diag dc = diag_at(s0);
        int rc = rx_device(h, stage[b]->as<ofdm_fc32>(), d_ns + s0, ns, iq_stride, mx, h->s_bytes.as<uint8_t>() + (size_t)s0 * out_stride,
                           out_stride, d_ol + s0, h->s_status.as<int32_t>() + s0, have_diag ? &dc : nullptr, st, s0, n_streams, nullptr, 2);
        if (rc) return rc;
        CU(h, cudaEventRecord(h->ev_feed[b][2], st));
        CU(h, cudaStreamWaitEvent(os, h->ev_feed[b][2], 0));
        CU(h, cudaMemcpyAsync(out + (size_t)s0 * out_stride, h->s_bytes.as<uint8_t>() + (size_t)s0 * out_stride, (size_t)ns * out_stride, cudaMemcpyDeviceToHost, os));
    }
    if (getenv("OFDM_DEBUG_FEED")) fprintf(stderr, "[ofdm] skip-cp feed: %u chunks, host time in the 2-D copy submissions %.2f ms\n", n_chunks, submit_s * 1e3);
    if (iq_mapped) {                                                           // what the gather kernels pulled (counted on the device)
        unsigned long long pulled = 0;
        CU(h, cudaMemcpyAsync(&pulled, h->s_moved.p, sizeof(pulled), cudaMemcpyDeviceToHost, cs));
        CU(h, cudaStreamSynchronize(cs));
        moved += pulled;
    }
    h->h2d_bytes_last = moved;
    return 0;
}

extern "C" uint64_t ofdm_last_h2d_bytes(const ofdm_engine *h) { return h ? h->h2d_bytes_last : 0; }

extern "C" int ofdm_rx_decode_batch(ofdm_engine *h, const ofdm_fc32 *iq, const uint32_t *n_samples,
                                    uint32_t n_streams, uint32_t iq_stride, uint32_t max_n_samples,
                                    uint8_t *out, uint32_t out_stride, uint32_t *out_len, int32_t *status,
                                    const ofdm_rx_diag *diag, int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!iq || !n_samples || !out || !out_len || !status || n_streams == 0 || iq_stride == 0)
        ENG_FAIL(h, OFDM_E_INVALID, "rx: bad arguments");
    CU(h, cudaSetDevice(h->device));
    if (mem == OFDM_MEM_DEVICE)
        return rx_device(h, iq, n_samples, n_streams, iq_stride, max_n_samples, out, out_stride, out_len, status, diag, (cudaStream_t)stream);

    uint32_t mx = 0;
    for (uint32_t s = 0; s < n_streams; s++) {
        if (n_samples[s] > iq_stride) ENG_FAIL(h, OFDM_E_INVALID, "rx: n_samples[%u] exceeds iq_stride", s);
        if (n_samples[s] > mx) mx = n_samples[s];
    }
    // Host path: the capture is streamed through two device staging buffers in chunks of whole streams; the H2D copy of
    // chunk c+1 (copy_stream) overlaps the kernels of chunk c (own_stream), and the payloads of chunk c go back on a third
    // stream as soon as its kernels are done (PCIe is full duplex: the D2H traffic hides behind the H2D feed).
    cudaStream_t st = h->own_stream, cs = h->copy_stream, os = h->d2h_stream;
    const size_t stream_bytes = (size_t)iq_stride * sizeof(float2);
    uint32_t chunk = (uint32_t)((256u << 20) / stream_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > n_streams) chunk = n_streams;
    const size_t ob = (size_t)n_streams * out_stride;
    CU(h, h->s_iq.ensure(chunk * stream_bytes));
    CU(h, h->s_iq2.ensure(chunk * stream_bytes));
    CU(h, h->s_bytes.ensure(ob));
    CU(h, h->s_len.ensure(2 * sizeof(uint32_t) * (size_t)n_streams));
    CU(h, h->s_status.ensure(sizeof(int32_t) * (size_t)n_streams));
    CU(h, h->state.ensure((h->wide ? sizeof(wide::StreamStateW) : sizeof(StreamState)) * (size_t)n_streams));
    uint32_t *d_ns = h->s_len.as<uint32_t>(), *d_ol = d_ns + n_streams;
    ofdm_rx_diag dd{};
    if (diag) {
        CU(h, h->s_aux.ensure(3 * sizeof(uint32_t) * (size_t)n_streams));
        uint32_t *aux = h->s_aux.as<uint32_t>();
        if (diag->offset) dd.offset = reinterpret_cast<int32_t *>(aux);
        if (diag->f_delta) dd.f_delta = reinterpret_cast<float *>(aux + n_streams);
        if (diag->n_data_syms) dd.n_data_syms = aux + 2 * (size_t)n_streams;
        if (diag->h_k) { CU(h, h->s_h.ensure(sizeof(float2) * (size_t)h->nfft * n_streams)); dd.h_k = h->s_h.as<ofdm_fc32>(); }
        if (diag->points && diag->points_stride) {
            CU(h, h->s_points.ensure(sizeof(float2) * (size_t)diag->points_stride * n_streams));
            CU(h, cudaMemsetAsync(h->s_points.p, 0, sizeof(float2) * (size_t)diag->points_stride * n_streams, st));
            dd.points = h->s_points.as<ofdm_fc32>(); dd.points_stride = diag->points_stride;
        }
    }
    CU(h, cudaMemcpyAsync(d_ns, n_samples, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyHostToDevice, st));
    // Everything the acquisition can touch lies in the first sync_window + 13 symbols of a stream (docs/SPEC.md 4-5: lags
    // below the window, refinement up to a fifth of a symbol later, 10 head rows, the header symbol). When that is a small part
    // of the capture the data symbols are fetched WITHOUT their cyclic prefixes once the acquisition has located them.
    const uint64_t head_len64 = (uint64_t)h->cfg.sync_window + 13ull * (uint64_t)h->sym_len;
    // Measured (B200, PCIe 5 x16, 55 GB/s copy ceiling; profiles/r2_feed_ab.txt): the gather kernel pulls 46 GB/s, and the GPU
    // fetches host memory in 128-byte lines -- an nfft = 64 symbol's 512 useful bytes at an arbitrary 8-byte alignment touch the
    // same five lines as the whole 640-byte symbol, so nothing is saved there (6 753 vs 6 857 Msamples/s); 2-D copies per stream
    // run at 40 GB/s (a copy-engine cost per copy, not host time: 1.1 us per submission). With nfft = 1024 (8 KB useful of
    // 10 KB) the gather path wins: 7 022 vs 6 695 Msamples/s. So: automatic = gather for symbols of >= 4 KB, else the whole copy.
    bool skip_cp = h->rx_feed != 1 && h->cfg.sync_window > 0 &&
                   (h->rx_feed >= 2 ? head_len64 <= iq_stride : (4 * head_len64 <= iq_stride && (size_t)h->nfft * sizeof(float2) >= 4096));
    // the gather kernel reads the capture in place: only pinned (device-mapped) host memory qualifies; a pageable capture is
    // copied whole (automatic mode) or fetched with 2-D copies (OFDM_RX_FEED=skipcp)
    const ofdm_fc32 *iq_mapped = nullptr;
    if (skip_cp && h->rx_feed != 2) {
        cudaPointerAttributes pa{};
        if (cudaPointerGetAttributes(&pa, iq) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
            iq_mapped = reinterpret_cast<const ofdm_fc32 *>(pa.devicePointer);
        else {
            (void)cudaGetLastError();
            skip_cp = false;
        }
    }
    h->h2d_bytes_last = (uint64_t)n_streams * stream_bytes;
    if (skip_cp) {
        int rc = rx_host_skip_cp(h, iq, n_samples, n_streams, iq_stride, mx, (uint32_t)head_len64, out, out_stride, &dd, diag != nullptr, d_ns, d_ol, iq_mapped);
        if (rc) return rc;
    }
    DevBuf *stage[2] = { &h->s_iq, &h->s_iq2 };
    uint32_t ci = 0;
    for (uint32_t s0 = 0; s0 < n_streams && !skip_cp; s0 += chunk, ci++) {
        const uint32_t ns = n_streams - s0 < chunk ? n_streams - s0 : chunk;
        const int b = ci & 1;
        if (ci >= 2) CU(h, cudaStreamWaitEvent(cs, h->ev_done[b], 0));           // staging buffer free again
        CU(h, cudaMemcpyAsync(stage[b]->p, iq + (size_t)s0 * iq_stride, ns * stream_bytes, cudaMemcpyHostToDevice, cs));
        CU(h, cudaEventRecord(h->ev_copied[b], cs));
        CU(h, cudaStreamWaitEvent(st, h->ev_copied[b], 0));
        ofdm_rx_diag dc = dd;
        if (dc.offset) dc.offset += s0;
        if (dc.f_delta) dc.f_delta += s0;
        if (dc.n_data_syms) dc.n_data_syms += s0;
        if (dc.h_k) dc.h_k += (size_t)s0 * h->nfft;
        if (dc.points) dc.points += (size_t)s0 * dc.points_stride;
        int rc = rx_device(h, stage[b]->as<ofdm_fc32>(), d_ns + s0, ns, iq_stride, mx, h->s_bytes.as<uint8_t>() + (size_t)s0 * out_stride,
                           out_stride, d_ol + s0, h->s_status.as<int32_t>() + s0, diag ? &dc : nullptr, st, s0, n_streams);
        if (rc) return rc;
        CU(h, cudaEventRecord(h->ev_done[b], st));
        CU(h, cudaStreamWaitEvent(os, h->ev_done[b], 0));
        CU(h, cudaMemcpyAsync(out + (size_t)s0 * out_stride, h->s_bytes.as<uint8_t>() + (size_t)s0 * out_stride, (size_t)ns * out_stride, cudaMemcpyDeviceToHost, os));
    }
    CU(h, cudaMemcpyAsync(out_len, d_ol, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(status, h->s_status.p, sizeof(int32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
    if (diag) {
        if (dd.offset) CU(h, cudaMemcpyAsync(diag->offset, dd.offset, sizeof(int32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
        if (dd.f_delta) CU(h, cudaMemcpyAsync(diag->f_delta, dd.f_delta, sizeof(float) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
        if (dd.n_data_syms) CU(h, cudaMemcpyAsync(diag->n_data_syms, dd.n_data_syms, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
        if (dd.h_k) CU(h, cudaMemcpyAsync(diag->h_k, dd.h_k, sizeof(float2) * (size_t)h->nfft * n_streams, cudaMemcpyDeviceToHost, st));
        if (dd.points) CU(h, cudaMemcpyAsync(diag->points, dd.points, sizeof(float2) * (size_t)diag->points_stride * n_streams, cudaMemcpyDeviceToHost, st));
    }
    CU(h, cudaStreamSynchronize(st));
    CU(h, cudaStreamSynchronize(os));
    return 0;
}

// ---- channel harness ---------------------------------------------------------------------------------------------
static int channel_device(ofdm_engine *h, const ofdm_fc32 *tx, const uint32_t *tx_len, uint32_t tx_stride, uint32_t n_streams,
                          const ofdm_channel_params *p, ofdm_fc32 *rx, uint32_t rx_stride, uint32_t *rx_len,
                          uint32_t *lead_out, float *cfo_out, cudaStream_t st)
{
    CU(h, h->scratch_f32.ensure(5 * sizeof(double) * (size_t)n_streams));
    CU(h, cudaMemsetAsync(h->scratch_f32.p, 0, 5 * sizeof(double) * (size_t)n_streams, st));
    ChanArgs a{};
    a.tx = reinterpret_cast<const float2 *>(tx); a.tx_len = tx_len; a.tx_stride = tx_stride; a.n_streams = n_streams;
    a.rx = reinterpret_cast<float2 *>(rx); a.rx_stride = rx_stride; a.rx_len = rx_len; a.lead_out = lead_out; a.cfo_out = cfo_out;
    a.accum = h->scratch_f32.as<double>();
    a.snr_lin = powf(10.0f, p->snr_db / 10.0f);
    a.cfo_max = p->cfo_max; a.lead_min = p->lead_min; a.lead_max = p->lead_max; a.multipath = p->multipath; a.noise_mode = p->noise_mode;
    a.seed_lo = (uint32_t)p->seed; a.seed_hi = (uint32_t)(p->seed >> 32);
    uint32_t gx = (rx_stride + 256 * 8 - 1) / (256 * 8);
    if (gx < 1) gx = 1;
    launch_streams(channel_conv_fn(), a, gx, n_streams, 256, 0, st, h->launches);
    launch_streams(channel_noise_fn(), a, gx, n_streams, 256, 0, st, h->launches);
    CU(h, cudaGetLastError());
    return 0;
}

extern "C" int ofdm_channel_apply_batch(ofdm_engine *h, const ofdm_fc32 *tx, const uint32_t *tx_len, uint32_t tx_stride,
                                        uint32_t n_streams, const ofdm_channel_params *p,
                                        ofdm_fc32 *rx, uint32_t rx_stride, uint32_t *rx_len,
                                        uint32_t *lead_out, float *cfo_out, int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!tx || !tx_len || !p || !rx || !rx_len || n_streams == 0) ENG_FAIL(h, OFDM_E_INVALID, "channel: bad arguments");
    CU(h, cudaSetDevice(h->device));
    if (mem == OFDM_MEM_DEVICE)
        return channel_device(h, tx, tx_len, tx_stride, n_streams, p, rx, rx_stride, rx_len, lead_out, cfo_out, (cudaStream_t)stream);
    cudaStream_t st = h->own_stream;
    size_t tb = (size_t)n_streams * tx_stride * sizeof(float2), rb = (size_t)n_streams * rx_stride * sizeof(float2);
    CU(h, h->s_iq.ensure(tb));
    CU(h, h->s_iq2.ensure(rb));
    CU(h, h->s_len.ensure(4 * sizeof(uint32_t) * (size_t)n_streams));
    uint32_t *d_tl = h->s_len.as<uint32_t>(), *d_rl = d_tl + n_streams, *d_lead = d_rl + n_streams;
    float *d_cfo = reinterpret_cast<float *>(d_lead + n_streams);
    CU(h, cudaMemcpyAsync(h->s_iq.p, tx, tb, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_tl, tx_len, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyHostToDevice, st));
    int rc = channel_device(h, h->s_iq.as<ofdm_fc32>(), d_tl, tx_stride, n_streams, p, h->s_iq2.as<ofdm_fc32>(), rx_stride, d_rl, d_lead, d_cfo, st);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(rx, h->s_iq2.p, rb, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(rx_len, d_rl, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
    if (lead_out) CU(h, cudaMemcpyAsync(lead_out, d_lead, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
    if (cfo_out) CU(h, cudaMemcpyAsync(cfo_out, d_cfo, sizeof(float) * (size_t)n_streams, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

// ---- capture search ----------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct SyncPlan {
    SyncArgs a;
    ScanTensorMap tmap;
    bool tma;
    uint32_t grid;
};

// scratch + arguments of one capture search; the per-tile slots mean there is no candidate capacity to overflow globally
static int sync_plan(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n, ofdm_peak *peaks, uint32_t max_peaks, cudaStream_t st, SyncPlan &p)
{
    static_assert(sizeof(SyncPeak) == sizeof(ofdm_peak), "peak layout");
    static_assert(sizeof(ScanTensorMap) == sizeof(CUtensorMap), "tensor map size");
    memset(&p, 0, sizeof p);
    SyncArgs &a = p.a;
    a.iq = reinterpret_cast<const float2 *>(iq); a.n = n;
    a.tables = h->d_tables; a.peaks = reinterpret_cast<SyncPeak *>(peaks); a.max_peaks = max_peaks;
    a.wtables = h->d_wtables; a.lock_is_ramp = h->lock_is_ramp ? 1 : 0;
    p.tma = false;
    // TMA staging needs a 16-byte aligned capture and at least one full 8-sample row
    if ((reinterpret_cast<uintptr_t>(iq) & 15) == 0 && n / kScanT >= 1 && n / kScanT < 0x7FFFFF00ull && encode_tiled_fn() &&
        !(h->wide && getenv("OFDM_WIDE_SCAN_GENERIC"))) {
        const cuuint64_t dims[2] = { 16, (cuuint64_t)(n / kScanT) };        // [rows][16 floats = 8 fc32 samples]
        const cuuint64_t strides[1] = { 64 };
        const cuuint32_t box[2] = { 16, 256 }, estr[2] = { 1, 1 };
        const CUresult r = encode_tiled_fn()(reinterpret_cast<CUtensorMap *>(&p.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<ofdm_fc32 *>(iq),
                                             dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        p.tma = r == CUDA_SUCCESS;
    }
    const uint64_t LS = (uint64_t)h->sym_len;                              // 80 or 1280
    a.tile_lags = h->wide ? (p.tma ? (uint32_t)kWTD : (uint32_t)kWScanD) : (uint32_t)kScanD;
    a.holdoff = (uint32_t)(10 * LS);
    a.prefetch_ahead = 0;
    if (h->wide && p.tma) a.prefetch_ahead = (uint32_t)kWTCtasPerSm * (uint32_t)h->n_sm;   // the tile that follows in the same SM slot
    const uint64_t lags = n >= 2 * LS ? n - 2 * LS + 1 : 0;
    const uint64_t T = (lags + a.tile_lags - 1) / a.tile_lags;
    if (T > 0xFFFFFFFFull) ENG_FAIL(h, OFDM_E_INVALID, "sync: capture too long");
    a.n_tiles = (uint32_t)T;
    const size_t C = (size_t)T * kTileCand;
    const size_t off_cnt = 32, off_ord = (off_cnt + 4 * T + 7) & ~(size_t)7, off_sel = off_ord + 8 * C, off_cand = off_sel + 8 * C, off_keep = off_cand + 2 * C;
    CU(h, h->sync_scratch.ensure(off_keep + C + 16));
    uint8_t *base = h->sync_scratch.as<uint8_t>();
    a.counters = reinterpret_cast<uint32_t *>(base);
    a.tile_cnt = reinterpret_cast<uint32_t *>(base + off_cnt);
    a.ordered = reinterpret_cast<uint64_t *>(base + off_ord);
    a.sel = reinterpret_cast<uint64_t *>(base + off_sel);
    a.tile_cand = reinterpret_cast<uint16_t *>(base + off_cand);
    a.keep = base + off_keep;
    CU(h, cudaMemsetAsync(base, 0, off_cnt + 4 * T, st));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    p.grid = (uint32_t)(2 * sms);                                          // persistent: 2 CTAs per SM
    if (h->wide) {
        if (p.tma) {
            SyncScanKernel kw = wide_scan_tma_fn();
            if (h->smem_configured.insert((const void *)kw).second)
                CU(h, cudaFuncSetAttribute((const void *)kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_scan_tma_smem_bytes()));
        } else {
            SyncKernel kw = wide_scan_fn();
            if (h->smem_configured.insert((const void *)kw).second)
                CU(h, cudaFuncSetAttribute((const void *)kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_scan_smem_bytes()));
        }
        return 0;
    }
    SyncScanKernel k = sync_scan_fn(p.tma);
    if (h->smem_configured.insert((const void *)k).second)
        CU(h, cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sync_scan_smem_bytes(p.tma)));
    return 0;
}

// scan tiles [tile_first, tile_first + tile_count)
static int sync_scan_range(ofdm_engine *h, SyncPlan &p, uint32_t tile_first, uint32_t tile_count, cudaStream_t st)
{
    if (tile_count == 0) return 0;
    p.a.tile_first = tile_first; p.a.tile_count = tile_count;
    if (h->wide) {                                                         // one CTA per tile (3576 lags staged by TMA, or 4096)
        if (p.tma) wide_scan_tma_fn()<<<tile_count, kWTThreads, wide_scan_tma_smem_bytes(), st>>>(p.a, p.tmap);
        else wide_scan_fn()<<<tile_count, kWScanThreads, wide_scan_smem_bytes(), st>>>(p.a);
        h->launches += 1;
        CU(h, cudaGetLastError());
        return 0;
    }
    const uint32_t grid = tile_count < p.grid ? tile_count : p.grid;
    sync_scan_fn(p.tma)<<<grid, kScanRows, sync_scan_smem_bytes(p.tma), st>>>(p.a, p.tmap);
    h->launches += 1;
    CU(h, cudaGetLastError());
    return 0;
}

// hold-off + refinement of everything the scan launches recorded; *n_peaks = min(detections, max_peaks), on the device
static int sync_finish(ofdm_engine *h, SyncPlan &p, uint32_t *n_peaks, cudaStream_t st)
{
    if (p.a.n_tiles) {
        sync_select_fn()<<<1, kSelThreads, 0, st>>>(p.a);
        // one CTA per detection; the launch is sized by max_peaks (in chunks), CTAs beyond the detection count exit at once
        const uint32_t worst = p.a.n_tiles > 0xFFFFFFFFu / kTileCand ? 0xFFFFFFFFu : p.a.n_tiles * kTileCand;
        const uint32_t rg = p.a.max_peaks < worst ? p.a.max_peaks : worst;
        if (rg && h->wide) wide_sync_refine_fn()<<<rg, wide::kThreads, 0, st>>>(p.a);
        else if (rg) sync_refine_fn()<<<rg, kAcqThreads, 0, st>>>(p.a);
        h->launches += 2;
    }
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(n_peaks, p.a.counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int sync_device(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n, ofdm_peak *peaks, uint32_t max_peaks, uint32_t *n_peaks, cudaStream_t st)
{
    SyncPlan p;
    if (int rc = sync_plan(h, iq, n, peaks, max_peaks, st, p)) return rc;
    if (int rc = sync_scan_range(h, p, 0, p.a.n_tiles, st)) return rc;
    return sync_finish(h, p, n_peaks, st);
}

extern "C" int ofdm_sync_search(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n_samples, ofdm_peak *peaks, uint32_t max_peaks,
                                uint32_t *n_peaks, int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!iq || !peaks || !n_peaks || max_peaks == 0) ENG_FAIL(h, OFDM_E_INVALID, "sync: bad arguments");
    CU(h, cudaSetDevice(h->device));
    if (mem == OFDM_MEM_DEVICE) return sync_device(h, iq, n_samples, peaks, max_peaks, n_peaks, (cudaStream_t)stream);

    // Host capture: copied in chunks on the copy stream; the scan of the tiles a chunk completes runs while the next chunk
    // is on the PCIe link. Hold-off, refinement and the peak list come once at the end.
    cudaStream_t st = h->own_stream, cs = h->copy_stream;
    CU(h, h->s_iq.ensure(n_samples * sizeof(float2) + 16));
    CU(h, h->s_points.ensure(sizeof(ofdm_peak) * (size_t)max_peaks + 16));
    CU(h, h->s_len.ensure(16));
    SyncPlan p;
    if (int rc = sync_plan(h, h->s_iq.as<ofdm_fc32>(), n_samples, h->s_points.as<ofdm_peak>(), max_peaks, st, p)) return rc;
    const uint64_t chunk = 8ull << 20;                                     // samples per copy (64 MB)
    uint32_t tiles_done = 0;
    int ev = 0;
    for (uint64_t s0 = 0; s0 < n_samples || s0 == 0; s0 += chunk) {
        const uint64_t s1 = s0 + chunk < n_samples ? s0 + chunk : n_samples;
        if (s1 > s0) CU(h, cudaMemcpyAsync(h->s_iq.as<ofdm_fc32>() + s0, iq + s0, (s1 - s0) * sizeof(float2), cudaMemcpyHostToDevice, cs));
        CU(h, cudaEventRecord(h->ev_copied[ev], cs));
        CU(h, cudaStreamWaitEvent(st, h->ev_copied[ev], 0));
        ev ^= 1;
        // tile k reads samples [3928 k - 8, 3928 k - 8 + 4096): complete once s1 covers them (or the capture ends)
        // (nfft = 1024: tile k reads samples [4096 k - 8, 4096 k - 8 + kWScanSamples))
        const uint64_t t_span = h->wide ? (p.tma ? (uint64_t)kWTRows * 8 : (uint64_t)kWScanSamples) : (uint64_t)kScanRows * kScanT, t_lags = p.a.tile_lags;
        uint64_t ready = s1 >= n_samples ? p.a.n_tiles : (s1 + kScanT >= t_span ? (s1 + kScanT - t_span) / t_lags + 1 : 0);
        if (ready > p.a.n_tiles) ready = p.a.n_tiles;
        if (ready > tiles_done) {
            if (int rc = sync_scan_range(h, p, tiles_done, (uint32_t)ready - tiles_done, st)) return rc;
            tiles_done = (uint32_t)ready;
        }
        if (s1 >= n_samples) break;
    }
    if (int rc = sync_finish(h, p, h->s_len.as<uint32_t>(), st)) return rc;
    uint32_t cnt[4] = { 0, 0, 0, 0 };
    CU(h, cudaMemcpyAsync(cnt, p.a.counters, sizeof cnt, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    if (cnt[3]) ENG_FAIL(h, OFDM_E_INVALID, "sync: %u tile(s) of %u lags hold more than %d threshold crossings (the capture is not a sequence of frames)", cnt[3], p.a.tile_lags, kTileCand);
    const uint32_t m = cnt[1];
    std::vector<ofdm_peak> tmp(m);
    if (m) CU(h, cudaMemcpy(tmp.data(), h->s_points.p, sizeof(ofdm_peak) * m, cudaMemcpyDeviceToHost));
    uint32_t k = 0;
    for (uint32_t i = 0; i < m; i++) if (tmp[i].metric >= 0.0f) peaks[k++] = tmp[i];
    *n_peaks = k;
    return 0;
}

extern "C" int ofdm_sync_counts(ofdm_engine *h, uint32_t counts[4], void *stream)
{
    if (!h || !counts) return OFDM_E_INVALID;
    if (!h->sync_scratch.p) ENG_FAIL(h, OFDM_E_INVALID, "sync counts: no ofdm_sync_search has run on this handle");
    CU(h, cudaSetDevice(h->device));
    uint32_t c[4] = { 0, 0, 0, 0 };
    CU(h, cudaMemcpyAsync(c, h->sync_scratch.p, sizeof c, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(h, cudaStreamSynchronize((cudaStream_t)stream));
    counts[0] = c[0]; counts[1] = c[2]; counts[2] = c[1]; counts[3] = c[3];
    return 0;
}

// ---- streaming receiver ------------------------------------------------------------------------------------------
static int capture_decode_device(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n, const ofdm_peak *peaks, uint32_t n_frames,
                                 uint32_t max_frame, uint8_t *out, uint32_t out_stride, uint32_t *out_len, int32_t *status, cudaStream_t st)
{
    CU(h, h->cap_base.ensure((sizeof(uint64_t) + sizeof(uint32_t)) * (size_t)n_frames));
    uint64_t *base = h->cap_base.as<uint64_t>();
    uint32_t *ns = reinterpret_cast<uint32_t *>(base + n_frames);
    capture_prep_fn()<<<(n_frames + 255) / 256, 256, 0, st>>>(reinterpret_cast<const SyncPeak *>(peaks), n_frames, n, max_frame, base, ns);
    h->launches += 1;
    uint64_t cap = max_frame ? max_frame : n;
    if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
    return rx_device(h, iq, ns, n_frames, (uint32_t)cap, (uint32_t)cap, out, out_stride, out_len, status, nullptr, st, 0, 0, base);
}

extern "C" int ofdm_rx_decode_capture(ofdm_engine *h, const ofdm_fc32 *iq, uint64_t n_samples, const ofdm_peak *peaks, uint32_t n_frames,
                                      uint32_t max_frame_samples, uint8_t *out, uint32_t out_stride, uint32_t *out_len, int32_t *status,
                                      int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!iq || !peaks || !out || !out_len || !status) ENG_FAIL(h, OFDM_E_INVALID, "capture decode: bad arguments");
    if (n_frames == 0) return 0;
    CU(h, cudaSetDevice(h->device));
    if (mem == OFDM_MEM_DEVICE)
        return capture_decode_device(h, iq, n_samples, peaks, n_frames, max_frame_samples, out, out_stride, out_len, status, (cudaStream_t)stream);
    cudaStream_t st = h->own_stream;
    const size_t ob = (size_t)n_frames * out_stride;
    CU(h, h->s_iq.ensure(n_samples * sizeof(float2) + 16));
    CU(h, h->s_points.ensure(sizeof(ofdm_peak) * (size_t)n_frames));
    CU(h, h->s_bytes.ensure(ob));
    CU(h, h->s_len.ensure(sizeof(uint32_t) * (size_t)n_frames));
    CU(h, h->s_status.ensure(sizeof(int32_t) * (size_t)n_frames));
    CU(h, cudaMemcpyAsync(h->s_iq.p, iq, n_samples * sizeof(float2), cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->s_points.p, peaks, sizeof(ofdm_peak) * (size_t)n_frames, cudaMemcpyHostToDevice, st));
    int rc = capture_decode_device(h, h->s_iq.as<ofdm_fc32>(), n_samples, h->s_points.as<ofdm_peak>(), n_frames, max_frame_samples,
                                   h->s_bytes.as<uint8_t>(), out_stride, h->s_len.as<uint32_t>(), h->s_status.as<int32_t>(), st);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(out, h->s_bytes.p, ob, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(out_len, h->s_len.p, sizeof(uint32_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(status, h->s_status.p, sizeof(int32_t) * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

// ---- fc32 file ingest ------------------------------------------------------------------------------------------------
// pread `bytes` at `off` into dst, split over `pieces` threads (page-cache / NVMe reads scale with the queue depth)
static bool pread_parallel(int fd, uint64_t off, uint8_t *dst, size_t bytes, int pieces)
{
    std::vector<std::thread> th;
    std::vector<int> ok((size_t)pieces, 1);
    const size_t step = (bytes + pieces - 1) / pieces;
    for (int i = 0; i < pieces; i++) {
        const size_t a = (size_t)i * step;
        if (a >= bytes) break;
        const size_t b = std::min(bytes, a + step);
        th.emplace_back([=, &ok] {
            size_t done = a;
            while (done < b) {
                const ssize_t got = pread(fd, dst + done, b - done, (off_t)(off + done));
                if (got <= 0) { ok[(size_t)i] = 0; return; }
                done += (size_t)got;
            }
        });
    }
    for (auto &t : th) t.join();
    return std::all_of(ok.begin(), ok.end(), [](int v) { return v != 0; });
}

extern "C" int ofdm_rx_decode_file(ofdm_engine *h, const char *path, uint64_t start, uint64_t stop, uint32_t chunk_samples,
                                   uint32_t max_frame_samples, uint8_t *out, uint32_t out_stride, ofdm_frame_info *frames,
                                   uint32_t max_frames, uint32_t *n_frames)
{
    if (!h) return OFDM_E_INVALID;
    if (!path || !out || !frames || !n_frames || max_frames == 0 || out_stride == 0 || max_frame_samples == 0)
        ENG_FAIL(h, OFDM_E_INVALID, "decode file: bad arguments");
    *n_frames = 0;
    const uint64_t chunk = chunk_samples ? chunk_samples : (1u << 25);
    const uint64_t overlap = max_frame_samples;
    if (chunk <= overlap) ENG_FAIL(h, OFDM_E_INVALID, "decode file: chunk_samples must exceed max_frame_samples");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) ENG_FAIL(h, OFDM_E_INVALID, "decode file: cannot open %s", path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); ENG_FAIL(h, OFDM_E_INVALID, "decode file: cannot stat %s", path); }
    const uint64_t n_file = (uint64_t)sb.st_size / 8;                      // bytes_to_sig: 8 bytes per sample (src/utils.rs:238-254)
    if (stop == 0 || stop > n_file) stop = n_file;
    if (start > stop) start = stop;
    const uint64_t n = stop - start;
    struct Closer { int fd; ~Closer() { close(fd); } } closer{ fd };
    if (n == 0) return 0;
    CU(h, cudaSetDevice(h->device));
    const uint64_t buf_samples = std::min(chunk, n);
    if (h->pin_bytes < buf_samples * 8) {
        for (void *&p : h->pin) { if (p) cudaFreeHost(p); p = nullptr; }
        h->pin_bytes = 0;
        for (void *&p : h->pin) CU(h, cudaHostAlloc(&p, buf_samples * 8, cudaHostAllocDefault));
        h->pin_bytes = buf_samples * 8;
    }
    const uint32_t max_peaks = std::min<uint32_t>(max_frames, 16384u);     // detections one chunk may hold
    cudaStream_t st = h->own_stream;
    CU(h, h->s_iq.ensure(buf_samples * 8 + 16));
    CU(h, h->s_points.ensure(sizeof(ofdm_peak) * (size_t)max_peaks + 16));
    CU(h, h->s_bytes.ensure((size_t)max_peaks * out_stride));
    CU(h, h->s_len.ensure(sizeof(uint32_t) * (size_t)max_peaks + 16));
    CU(h, h->s_status.ensure(sizeof(int32_t) * (size_t)max_peaks));
    CU(h, h->s_aux.ensure(16));
    uint32_t *d_npk = h->s_aux.as<uint32_t>();

    // chunk plan: chunks advance by chunk - overlap; a chunk owns the frames that start before its last `overlap` samples
    std::vector<std::pair<uint64_t, uint64_t>> plan;
    for (uint64_t a = 0;;) {
        const uint64_t b = std::min(n, a + chunk);
        plan.emplace_back(a, b);
        if (b >= n) break;
        a = b - overlap;
    }
    // page cache -> pinned memory is a kernel-side copy, a few GB/s per thread: as many readers as the host has threads (<= 32)
    int readers = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("OFDM_FILE_READERS")) readers = atoi(e);
    readers = readers < 1 ? 8 : (readers > 32 ? 32 : readers);
    auto load = [&](size_t i) { return pread_parallel(fd, 8 * (start + plan[i].first), (uint8_t *)h->pin[i & 1], 8 * (plan[i].second - plan[i].first), readers); };
    std::future<bool> ahead = std::async(std::launch::async, load, (size_t)0);
    std::vector<ofdm_peak> pk(max_peaks);
    std::vector<uint32_t> lens(max_peaks);
    std::vector<int32_t> stat(max_peaks);
    std::vector<uint8_t> rows;
    uint32_t count = 0;
    bool have_last = false;
    uint64_t last = 0;
    for (size_t i = 0; i < plan.size(); i++) {
        if (!ahead.get()) ENG_FAIL(h, OFDM_E_INVALID, "decode file: short read from %s", path);
        if (i + 1 < plan.size()) ahead = std::async(std::launch::async, load, i + 1);       // overlaps this chunk's PCIe copy and kernels
        const uint64_t cn = plan[i].second - plan[i].first;
        const bool final = i + 1 == plan.size();
        auto drain = [&]() { if (ahead.valid()) ahead.wait(); };                            // never leave a reader running on an error path
        if (cn < (uint64_t)kHeadSyms * (uint64_t)h->sym_len) continue;
#define FCU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { drain(); ENG_FAIL(h, OFDM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); } } while (0)
        FCU(cudaMemcpyAsync(h->s_iq.p, h->pin[i & 1], cn * 8, cudaMemcpyHostToDevice, st));
        if (int rc = sync_device(h, h->s_iq.as<ofdm_fc32>(), cn, h->s_points.as<ofdm_peak>(), max_peaks, d_npk, st)) { drain(); return rc; }
        uint32_t cnt[4] = { 0, 0, 0, 0 };
        FCU(cudaMemcpyAsync(cnt, h->sync_scratch.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        if (cnt[3] || cnt[2] > cnt[1]) { drain(); ENG_FAIL(h, OFDM_E_INVALID, "decode file: a chunk holds more than %u detections; use a smaller chunk_samples", max_peaks); }
        const uint32_t k = cnt[1];
        if (k == 0) continue;
        if (int rc = capture_decode_device(h, h->s_iq.as<ofdm_fc32>(), cn, h->s_points.as<ofdm_peak>(), k, (uint32_t)overlap, h->s_bytes.as<uint8_t>(),
                                           out_stride, h->s_len.as<uint32_t>(), h->s_status.as<int32_t>(), st)) { drain(); return rc; }
        FCU(cudaMemcpyAsync(pk.data(), h->s_points.p, sizeof(ofdm_peak) * k, cudaMemcpyDeviceToHost, st));
        FCU(cudaMemcpyAsync(lens.data(), h->s_len.p, sizeof(uint32_t) * k, cudaMemcpyDeviceToHost, st));
        FCU(cudaMemcpyAsync(stat.data(), h->s_status.p, sizeof(int32_t) * k, cudaMemcpyDeviceToHost, st));
        FCU(cudaStreamSynchronize(st));
        uint32_t width = 0;
        for (uint32_t j = 0; j < k; j++) if (stat[j] == OFDM_OK && lens[j] > width) width = lens[j];
        if (width) {                                                        // one strided device -> host copy of the payload rows
            rows.resize((size_t)k * width);
            FCU(cudaMemcpy2DAsync(rows.data(), width, h->s_bytes.p, out_stride, width, k, cudaMemcpyDeviceToHost, st));
            FCU(cudaStreamSynchronize(st));
        }
#undef FCU
        for (uint32_t j = 0; j < k; j++) {
            if (pk[j].metric < 0.0f) continue;                             // frame head cut by the chunk end
            if (!final && pk[j].offset >= cn - overlap) continue;          // owned by the next chunk
            const uint64_t absolute = plan[i].first + pk[j].offset;
            if (have_last && absolute < last + 10ull * (uint64_t)h->sym_len) continue;    // already reported from the previous chunk's body
            have_last = true; last = absolute;
            if (count >= max_frames) { drain(); ENG_FAIL(h, OFDM_E_INVALID, "decode file: more than max_frames = %u frames", max_frames); }
            ofdm_frame_info &f = frames[count];
            f.offset = absolute; f.f_delta = pk[j].f_delta; f.metric = pk[j].metric; f.status = stat[j];
            f.out_len = stat[j] == OFDM_OK ? lens[j] : 0;
            if (f.out_len) memcpy(out + (size_t)count * out_stride, rows.data() + (size_t)j * width, f.out_len);
            count++;
        }
    }
    *n_frames = count;
    return 0;
}

// ---- statistics reduction (the path's only collective) ---------------------------------------------------------------------
typedef int (*NcclAllReduceFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
static NcclAllReduceFn nccl_allreduce_fn()
{
    static NcclAllReduceFn fn = [] {
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");                  // an NCCL the process already exposes globally
        if (!sym) {
            void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);     // the NCCL already loaded (e.g. by the caller's framework)
            if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
            if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
            if (lib) sym = dlsym(lib, "ncclAllReduce");
        }
        return (NcclAllReduceFn)sym;
    }();
    return fn;
}

extern "C" int ofdm_stats_allreduce(ofdm_engine *h, uint64_t *counters, void *nccl_comm, int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!counters || !nccl_comm) ENG_FAIL(h, OFDM_E_INVALID, "stats allreduce: bad arguments");
    NcclAllReduceFn ar = nccl_allreduce_fn();
    if (!ar) ENG_FAIL(h, OFDM_E_INVALID, "stats allreduce: no NCCL library in this process (libnccl.so.2 not found)");
    CU(h, cudaSetDevice(h->device));
    constexpr int kNcclUint64 = 5, kNcclSum = 0;                           // ncclDataType_t / ncclRedOp_t (nccl.h)
    if (mem == OFDM_MEM_DEVICE) {
        const int rc = ar(counters, counters, 4, kNcclUint64, kNcclSum, nccl_comm, (cudaStream_t)stream);
        if (rc != 0) ENG_FAIL(h, OFDM_E_CUDA, "ncclAllReduce failed (ncclResult_t %d)", rc);
        return 0;
    }
    cudaStream_t st = h->own_stream;
    CU(h, cudaMemcpyAsync(h->counters.p, counters, 4 * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    const int rc = ar(h->counters.p, h->counters.p, 4, kNcclUint64, kNcclSum, nccl_comm, st);
    if (rc != 0) ENG_FAIL(h, OFDM_E_CUDA, "ncclAllReduce failed (ncclResult_t %d)", rc);
    CU(h, cudaMemcpyAsync(counters, h->counters.p, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

// ---- Reed-Solomon outer code -------------------------------------------------------------------------------------
extern "C" size_t ofdm_rs_encoded_len(size_t n) { return (size_t)kRsN * (n / kRsK + 1); }
extern "C" size_t ofdm_rs_decoded_len(size_t n) { return (size_t)kRsK * (n / kRsN + 1); }

static int rs_tables_ready(ofdm_engine *h)
{
    if (h->rs_tables.p) return 0;
    std::vector<uint8_t> host(sizeof(RsTables));
    RsTables *t = reinterpret_cast<RsTables *>(host.data());
    unsigned x = 1;
    for (int i = 0; i < 255; i++) {
        t->exp[i] = (uint8_t)x;
        t->log[x] = (uint8_t)i;
        x <<= 1;
        if (x & 0x100) x ^= 0x11d;
    }
    for (int i = 255; i < 512; i++) t->exp[i] = t->exp[i - 255];
    t->log[0] = 0;
    auto mul = [&](unsigned a, unsigned b) -> uint8_t { return (a && b) ? t->exp[t->log[a] + t->log[b]] : (uint8_t)0; };
    uint8_t g[kRsT2 + 1] = { 1 };                                   // highest degree first
    for (int i = 0; i < kRsT2; i++)
        for (int j = i + 1; j >= 1; j--) g[j] = (uint8_t)(g[j] ^ mul(g[j - 1], t->exp[i]));
    for (int f = 0; f < 256; f++)
        for (int j = 0; j < kRsT2; j++) t->lfsr[f][j] = mul((unsigned)f, g[j + 1]);
    CU(h, h->rs_tables.ensure(sizeof(RsTables)));
    CU(h, cudaMemcpy(h->rs_tables.p, host.data(), sizeof(RsTables), cudaMemcpyHostToDevice));
    CU(h, cudaFuncSetAttribute((const void *)rs_encode_fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRsSmemBytes));
    CU(h, cudaFuncSetAttribute((const void *)rs_decode_fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRsSmemBytes));
    return 0;
}

// encode = true: `blk_in` = 223, `blk_out` = 255
static int rs_run(ofdm_engine *h, bool encode, const uint8_t *in, const uint32_t *in_len, uint32_t n_streams, uint32_t in_stride,
                  uint8_t *out, uint32_t out_stride, uint32_t *out_len, uint32_t *n_corrected, uint32_t *n_failed, int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!in || !in_len || !out || !out_len || n_streams == 0 || (!encode && (!n_corrected || !n_failed)))
        ENG_FAIL(h, OFDM_E_INVALID, "rs: bad arguments");
    CU(h, cudaSetDevice(h->device));
    if (int r = rs_tables_ready(h)) return r;
    const uint32_t blk_in = encode ? kRsK : kRsN;
    const uint32_t max_blocks = in_stride / blk_in + 1;
    const uint64_t tasks = (uint64_t)max_blocks * n_streams;                 // one thread per (stream, block)
    if (tasks > 0x7fffffffull * kRsThreads) ENG_FAIL(h, OFDM_E_INVALID, "rs: batch too large");
    const dim3 grid((unsigned)((tasks + kRsThreads - 1) / kRsThreads));
    RsArgs a{};
    a.in_stride = in_stride; a.out_stride = out_stride; a.tables = h->rs_tables.as<RsTables>();
    a.n_streams = n_streams; a.blocks_per_stream = max_blocks;
    if (mem == OFDM_MEM_DEVICE) {
        cudaStream_t st = (cudaStream_t)stream;
        a.in = in; a.in_len = in_len; a.out = out; a.out_len = out_len; a.n_corrected = n_corrected; a.n_failed = n_failed;
        if (!encode) {
            CU(h, cudaMemsetAsync(n_corrected, 0, sizeof(uint32_t) * (size_t)n_streams, st));
            CU(h, cudaMemsetAsync(n_failed, 0, sizeof(uint32_t) * (size_t)n_streams, st));
            rs_decode_fn()<<<grid, kRsThreads, kRsSmemBytes, st>>>(a);
        } else rs_encode_fn()<<<grid, kRsThreads, kRsSmemBytes, st>>>(a);
        h->launches += 1;
        CU(h, cudaGetLastError());
        return 0;
    }
    for (uint32_t s = 0; s < n_streams; s++)
        if (in_len[s] > in_stride) ENG_FAIL(h, OFDM_E_INVALID, "rs: in_len[%u] exceeds the input stride", s);
    cudaStream_t st = h->own_stream;
    const size_t ib = (size_t)n_streams * in_stride, ob = (size_t)n_streams * out_stride, lb = sizeof(uint32_t) * (size_t)n_streams;
    CU(h, h->s_bytes.ensure(ib));
    CU(h, h->s_bytes2.ensure(ob));
    CU(h, h->s_len2.ensure(4 * lb));
    uint32_t *d_il = h->s_len2.as<uint32_t>(), *d_ol = d_il + n_streams, *d_nc = d_ol + n_streams, *d_nf = d_nc + n_streams;
    CU(h, cudaMemcpyAsync(h->s_bytes.p, in, ib, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_il, in_len, lb, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemsetAsync(d_nc, 0, 2 * lb, st));
    a.in = h->s_bytes.as<uint8_t>(); a.in_len = d_il; a.out = h->s_bytes2.as<uint8_t>(); a.out_len = d_ol; a.n_corrected = d_nc; a.n_failed = d_nf;
    if (encode) rs_encode_fn()<<<grid, kRsThreads, kRsSmemBytes, st>>>(a);
    else rs_decode_fn()<<<grid, kRsThreads, kRsSmemBytes, st>>>(a);
    h->launches += 1;
    CU(h, cudaGetLastError());
    CU(h, cudaMemcpyAsync(out, h->s_bytes2.p, ob, cudaMemcpyDeviceToHost, st));
    CU(h, cudaMemcpyAsync(out_len, d_ol, lb, cudaMemcpyDeviceToHost, st));
    if (!encode) {
        CU(h, cudaMemcpyAsync(n_corrected, d_nc, lb, cudaMemcpyDeviceToHost, st));
        CU(h, cudaMemcpyAsync(n_failed, d_nf, lb, cudaMemcpyDeviceToHost, st));
    }
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

extern "C" int ofdm_rs_encode_batch(ofdm_engine *h, const uint8_t *data, const uint32_t *data_len, uint32_t n_streams, uint32_t data_stride,
                                    uint8_t *coded, uint32_t coded_stride, uint32_t *coded_len, int mem, void *stream)
{
    return rs_run(h, true, data, data_len, n_streams, data_stride, coded, coded_stride, coded_len, nullptr, nullptr, mem, stream);
}

extern "C" int ofdm_rs_decode_batch(ofdm_engine *h, const uint8_t *coded, const uint32_t *coded_len, uint32_t n_streams, uint32_t coded_stride,
                                    uint8_t *data, uint32_t data_stride, uint32_t *data_len, uint32_t *n_corrected, uint32_t *n_failed,
                                    int mem, void *stream)
{
    return rs_run(h, false, coded, coded_len, n_streams, coded_stride, data, data_stride, data_len, n_corrected, n_failed, mem, stream);
}

// ---- BER ---------------------------------------------------------------------------------------------------------
extern "C" int ofdm_ber_accumulate(ofdm_engine *h, const uint8_t *ref, const uint32_t *ref_len, uint32_t ref_stride,
                                   const uint8_t *got, const uint32_t *got_len, uint32_t got_stride,
                                   const int32_t *status, uint32_t n_streams, uint64_t *counters,
                                   int mem, void *stream)
{
    if (!h) return OFDM_E_INVALID;
    if (!ref || !ref_len || !got || !got_len || !status || !counters || n_streams == 0) ENG_FAIL(h, OFDM_E_INVALID, "ber: bad arguments");
    CU(h, cudaSetDevice(h->device));
    BerArgs a{};
    a.ref_stride = ref_stride; a.got_stride = got_stride; a.n_streams = n_streams;
    if (mem == OFDM_MEM_DEVICE) {
        a.ref = ref; a.ref_len = ref_len; a.got = got; a.got_len = got_len; a.status = status;
        a.counters = reinterpret_cast<unsigned long long *>(counters);
        ber_fn()<<<n_streams, 256, 0, (cudaStream_t)stream>>>(a);
        h->launches += 1;
        CU(h, cudaGetLastError());
        return 0;
    }
    cudaStream_t st = h->own_stream;
    size_t rb = (size_t)n_streams * ref_stride, gb = (size_t)n_streams * got_stride;
    CU(h, h->s_bytes.ensure(rb));
    CU(h, h->s_bytes2.ensure(gb));
    CU(h, h->s_len2.ensure(3 * sizeof(uint32_t) * (size_t)n_streams));
    uint32_t *d_rl = h->s_len2.as<uint32_t>(), *d_gl = d_rl + n_streams;
    int32_t *d_st = reinterpret_cast<int32_t *>(d_gl + n_streams);
    CU(h, cudaMemcpyAsync(h->s_bytes.p, ref, rb, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->s_bytes2.p, got, gb, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_rl, ref_len, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_gl, got_len, sizeof(uint32_t) * (size_t)n_streams, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(d_st, status, sizeof(int32_t) * (size_t)n_streams, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemsetAsync(h->counters.p, 0, 4 * sizeof(uint64_t), st));
    a.ref = h->s_bytes.as<uint8_t>(); a.ref_len = d_rl; a.got = h->s_bytes2.as<uint8_t>(); a.got_len = d_gl; a.status = d_st;
    a.counters = h->counters.as<unsigned long long>();
    ber_fn()<<<n_streams, 256, 0, st>>>(a);
    h->launches += 1;
    CU(h, cudaGetLastError());
    uint64_t tmp[4];
    CU(h, cudaMemcpyAsync(tmp, h->counters.p, sizeof tmp, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    for (int i = 0; i < 4; i++) counters[i] += tmp[i];
    return 0;
}
