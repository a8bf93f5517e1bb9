// tx64w.cu -- instantiates the barrier-free one-pass transmit kernel (tx_warp.cuh).
#include "kernels.h"
#include "tx_warp.cuh"

namespace ofdm {

template <int MOD>
static TxKernel pick_txw_mod(bool guard, bool fec)
{
    if (guard) return fec ? (TxKernel)tx_warp_kernel<MOD, true, true> : (TxKernel)tx_warp_kernel<MOD, true, false>;
    return fec ? (TxKernel)tx_warp_kernel<MOD, false, true> : (TxKernel)tx_warp_kernel<MOD, false, false>;
}
TxKernel pick_tx_warp(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return pick_txw_mod<0>(c.guard_bands, c.fec);
    case 1: return pick_txw_mod<1>(c.guard_bands, c.fec);
    default: return pick_txw_mod<2>(c.guard_bands, c.fec);
    }
}
template <int MOD>
static TxKernel pick_txs_mod(bool guard, bool fec)
{
    if (guard) return fec ? (TxKernel)tx_spec_kernel<MOD, true, true> : (TxKernel)tx_spec_kernel<MOD, true, false>;
    return fec ? (TxKernel)tx_spec_kernel<MOD, false, true> : (TxKernel)tx_spec_kernel<MOD, false, false>;
}
TxKernel pick_tx_spec(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return pick_txs_mod<0>(c.guard_bands, c.fec);
    case 1: return pick_txs_mod<1>(c.guard_bands, c.fec);
    default: return pick_txs_mod<2>(c.guard_bands, c.fec);
    }
}
size_t tx_warp_smem(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return TwSmem<0>::kTotal;
    case 1: return TwSmem<1>::kTotal;
    default: return TwSmem<2>::kTotal;
    }
}
int tx_warp_syms_per_cta() { return kTwSyms; }
int tx_warp_threads() { return kTwThreads; }

}  // namespace ofdm
