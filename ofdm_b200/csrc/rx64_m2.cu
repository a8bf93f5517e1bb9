// rx64_m2.cu -- rx_decode_kernel<MOD = 2, ...> instantiations (see rx64_mod.inc)
#define RX64_MOD 2
#include "rx64_mod.inc"
