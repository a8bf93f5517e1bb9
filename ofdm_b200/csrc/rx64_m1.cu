// rx64_m1.cu -- rx_decode_kernel<MOD = 1, ...> instantiations (see rx64_mod.inc)
#define RX64_MOD 1
#include "rx64_mod.inc"
