// rx64_m0.cu -- rx_decode_kernel<MOD = 0, ...> instantiations (see rx64_mod.inc)
#define RX64_MOD 0
#include "rx64_mod.inc"
