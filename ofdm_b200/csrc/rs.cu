// rs.cu -- instantiates the Reed-Solomon RS(255,223) kernels (rs_kernels.cuh).
#include "kernels.h"

namespace ofdm {

RsKernel rs_encode_fn() { return rs_encode_kernel<>; }
RsKernel rs_decode_fn() { return rs_decode_kernel<>; }

}  // namespace ofdm
