// wide_tx.cu -- instantiates the nfft = 1024 transmit kernels (wide_kernels.cuh).
#include "kernels.h"

namespace ofdm {

template <int MOD, bool WRITE>
static WTxKernel wpick_tx_mod(bool guard, bool fec)
{
    if (guard) return fec ? (WTxKernel)wide::wide_tx_kernel<MOD, true, true, WRITE> : (WTxKernel)wide::wide_tx_kernel<MOD, true, false, WRITE>;
    return fec ? (WTxKernel)wide::wide_tx_kernel<MOD, false, true, WRITE> : (WTxKernel)wide::wide_tx_kernel<MOD, false, false, WRITE>;
}
template <bool WRITE>
static WTxKernel wpick_tx_w(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return wpick_tx_mod<0, WRITE>(c.guard_bands, c.fec);
    case 1: return wpick_tx_mod<1, WRITE>(c.guard_bands, c.fec);
    default: return wpick_tx_mod<2, WRITE>(c.guard_bands, c.fec);
    }
}
WTxKernel wpick_tx(const ofdm_cfg &c, bool write) { return write ? wpick_tx_w<true>(c) : wpick_tx_w<false>(c); }

}  // namespace ofdm
