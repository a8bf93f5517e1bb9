// kernels.h -- launch interface between the C ABI (ofdm_engine.cu) and the per-family kernel translation units
// (rx64*.cu, tx64.cu, wide_rx*.cu, wide_tx.cu, sync.cu, rs.cu). Every getter returns the host handle of a __global__ kernel
// that is instantiated in exactly one translation unit; the engine launches it with <<<>>> on the pointer. Splitting the
// kernels over several files only parallelises the build (one nvcc per family) -- there is no run-time dispatch beyond
// the reference's own run-time options (modulation, guard bands, FEC, phase mode).
#pragma once

#include "../../include/ofdm_engine.h"
#include "common.cuh"
#include "rx_kernels.cuh"
#include "tx_kernels.cuh"
#include "tx_resident.cuh"
#include "sync_kernels.cuh"
#include "wide_kernels.cuh"
#include "wide_sync_kernels.cuh"
#include "rs_kernels.cuh"

namespace ofdm {

typedef void (*DecodeKernel)(const RxArgs);
typedef void (*AcquireKernel)(const RxArgs);
typedef void (*TxKernel)(const TxArgs);
typedef void (*ChanKernel)(const ChanArgs);
typedef void (*BerKernel)(const BerArgs);
typedef void (*SyncKernel)(const SyncArgs);
typedef void (*SyncScanKernel)(const SyncArgs, const ScanTensorMap);
typedef void (*CapturePrepKernel)(const SyncPeak *, uint32_t, uint64_t, uint32_t, uint64_t *, uint32_t *);
typedef void (*RsKernel)(const RsArgs);
typedef void (*WDecodeKernel)(const wide::WideRxArgs);
typedef void (*WTxKernel)(const wide::WideTxArgs);

// rx64.cu (dispatch, acquisition), rx64_m{0,1,2}.cu (decode kernels of one modulation each)
template <int MOD> DecodeKernel pick_decode_mod(bool guard, bool fec, int phase, bool points);
template <> DecodeKernel pick_decode_mod<0>(bool, bool, int, bool);
template <> DecodeKernel pick_decode_mod<1>(bool, bool, int, bool);
template <> DecodeKernel pick_decode_mod<2>(bool, bool, int, bool);
DecodeKernel pick_decode(const ofdm_cfg &c, bool points);
AcquireKernel pick_acquire(const ofdm_cfg &c);
// tx64.cu
TxKernel pick_tx(const ofdm_cfg &c, bool write);
TxKernel pick_tx_frame(const ofdm_cfg &c);      // one-pass cluster kernel
// tx64r.cu
TxKernel pick_tx_resident(const ofdm_cfg &c);   // one-pass persistent kernel, frames resident in tensor memory
size_t tx_resident_smem(const ofdm_cfg &c);
// tx64w.cu
TxKernel pick_tx_warp(const ofdm_cfg &c);       // one-pass persistent kernel without CTA barriers in the frame loop (tx_warp.cuh)
TxKernel pick_tx_spec(const ofdm_cfg &c);       // speculative one-pass kernel: scales with the head maximum at once, records the frames that beat it
size_t tx_warp_smem(const ofdm_cfg &c);
int tx_warp_syms_per_cta();
int tx_warp_threads();
ChanKernel channel_conv_fn();
ChanKernel channel_noise_fn();
BerKernel ber_fn();
// sync.cu
SyncScanKernel sync_scan_fn(bool tma);
SyncKernel sync_select_fn();
SyncKernel sync_refine_fn();
CapturePrepKernel capture_prep_fn();
SyncKernel wide_scan_fn();             // nfft = 1024
SyncScanKernel wide_scan_tma_fn();     // ... tile staged by TMA tensor copies (16-byte aligned captures)
SyncKernel wide_sync_refine_fn();
// rs.cu
RsKernel rs_encode_fn();
RsKernel rs_decode_fn();
// wide_rx.cu (dispatch, acquisition), wide_rx_m{0,1,2}.cu, wide_tx.cu
template <int MOD> WDecodeKernel wpick_decode_mod(bool guard, bool fec, int phase, bool points);
template <> WDecodeKernel wpick_decode_mod<0>(bool, bool, int, bool);
template <> WDecodeKernel wpick_decode_mod<1>(bool, bool, int, bool);
template <> WDecodeKernel wpick_decode_mod<2>(bool, bool, int, bool);
WDecodeKernel wpick_decode(const ofdm_cfg &c, bool points);
WDecodeKernel wpick_acquire(const ofdm_cfg &c);
WTxKernel wpick_tx(const ofdm_cfg &c, bool write);
// wide_txr.cu
WTxKernel wpick_tx_resident(const ofdm_cfg &c, bool double_buffered);  // one-pass persistent kernel, frames resident in tensor memory (wide_tx_resident.cuh)
WTxKernel wpick_tx_spec(const ofdm_cfg &c);      // speculative one-pass kernel (scales with the head maximum at once, records the frames that beat it)
size_t wide_tx_resident_smem(const ofdm_cfg &c);
int wide_tx_resident_syms_per_cta(bool double_buffered);
int wide_tx_resident_threads();

}  // namespace ofdm
