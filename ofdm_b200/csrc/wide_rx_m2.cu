// wide_rx_m2.cu -- wide_decode_kernel<MOD = 2, ...> instantiations (see wide_rx_mod.inc)
#define WIDE_MOD 2
#include "wide_rx_mod.inc"
