// sync.cu -- instantiates the capture search / streaming receiver kernels (sync_kernels.cuh, wide_sync_kernels.cuh).
#include "kernels.h"

namespace ofdm {

SyncScanKernel sync_scan_fn(bool tma) { return tma ? (SyncScanKernel)sync_scan_kernel<true> : (SyncScanKernel)sync_scan_kernel<false>; }
SyncKernel sync_select_fn() { return sync_select_kernel<>; }
SyncKernel sync_refine_fn() { return sync_refine_kernel<>; }
CapturePrepKernel capture_prep_fn() { return capture_prep_kernel<>; }
SyncKernel wide_scan_fn() { return wide_scan_kernel<>; }
SyncScanKernel wide_scan_tma_fn() { return wide_scan_tma_kernel<>; }
SyncKernel wide_sync_refine_fn() { return wide_sync_refine_kernel<>; }

}  // namespace ofdm
