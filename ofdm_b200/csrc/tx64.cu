// tx64.cu -- instantiates the nfft = 64 transmit kernels, the channel harness and the BER counters (tx_kernels.cuh).
#include "kernels.h"

namespace ofdm {

template <int MOD, bool WRITE>
static TxKernel pick_tx_mod(bool guard, bool fec)
{
    if (guard) return fec ? (TxKernel)tx_tile_kernel<MOD, true, true, WRITE> : (TxKernel)tx_tile_kernel<MOD, true, false, WRITE>;
    return fec ? (TxKernel)tx_tile_kernel<MOD, false, true, WRITE> : (TxKernel)tx_tile_kernel<MOD, false, false, WRITE>;
}
template <bool WRITE>
static TxKernel pick_tx_w(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return pick_tx_mod<0, WRITE>(c.guard_bands, c.fec);
    case 1: return pick_tx_mod<1, WRITE>(c.guard_bands, c.fec);
    default: return pick_tx_mod<2, WRITE>(c.guard_bands, c.fec);
    }
}
TxKernel pick_tx(const ofdm_cfg &c, bool write) { return write ? pick_tx_w<true>(c) : pick_tx_w<false>(c); }

template <int MOD>
static TxKernel pick_txf_mod(bool guard, bool fec)
{
    if (guard) return fec ? (TxKernel)tx_frame_kernel<MOD, true, true> : (TxKernel)tx_frame_kernel<MOD, true, false>;
    return fec ? (TxKernel)tx_frame_kernel<MOD, false, true> : (TxKernel)tx_frame_kernel<MOD, false, false>;
}
TxKernel pick_tx_frame(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return pick_txf_mod<0>(c.guard_bands, c.fec);
    case 1: return pick_txf_mod<1>(c.guard_bands, c.fec);
    default: return pick_txf_mod<2>(c.guard_bands, c.fec);
    }
}

ChanKernel channel_conv_fn() { return channel_conv_kernel<>; }
ChanKernel channel_noise_fn() { return channel_noise_kernel<>; }
BerKernel ber_fn() { return ber_kernel<>; }

}  // namespace ofdm
