// wide_rx_m1.cu -- wide_decode_kernel<MOD = 1, ...> instantiations (see wide_rx_mod.inc)
#define WIDE_MOD 1
#include "wide_rx_mod.inc"
