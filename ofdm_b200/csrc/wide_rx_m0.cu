// wide_rx_m0.cu -- wide_decode_kernel<MOD = 0, ...> instantiations (see wide_rx_mod.inc)
#define WIDE_MOD 0
#include "wide_rx_mod.inc"
