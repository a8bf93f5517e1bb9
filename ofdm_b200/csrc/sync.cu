// sync.cu -- instantiates the capture search / streaming receiver kernels (sync_kernels.cuh).
#include "kernels.h"

namespace ofdm {

SyncKernel sync_scan_fn() { return sync_scan_kernel<>; }
SyncKernel sync_select_fn() { return sync_select_kernel<>; }
SyncKernel sync_refine_fn() { return sync_refine_kernel<>; }
CapturePrepKernel capture_prep_fn() { return capture_prep_kernel<>; }

}  // namespace ofdm
