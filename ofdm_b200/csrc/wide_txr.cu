// wide_txr.cu -- instantiates the one-pass, tensor-memory-resident transmit kernel of the nfft = 1024 variant (wide_tx_resident.cuh).
#include "kernels.h"
#include "wide_tx_resident.cuh"

namespace ofdm {

template <int MOD>
static WTxKernel wpick_txr_mod(bool guard, bool fec)
{
    if (guard) return fec ? (WTxKernel)wide::wide_tx_resident_kernel<MOD, true, true> : (WTxKernel)wide::wide_tx_resident_kernel<MOD, true, false>;
    return fec ? (WTxKernel)wide::wide_tx_resident_kernel<MOD, false, true> : (WTxKernel)wide::wide_tx_resident_kernel<MOD, false, false>;
}
WTxKernel wpick_tx_resident(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return wpick_txr_mod<0>(c.guard_bands, c.fec);
    case 1: return wpick_txr_mod<1>(c.guard_bands, c.fec);
    default: return wpick_txr_mod<2>(c.guard_bands, c.fec);
    }
}
size_t wide_tx_resident_smem(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return wide::WTrsSmem<0>::kTotal;
    case 1: return wide::WTrsSmem<1>::kTotal;
    default: return wide::WTrsSmem<2>::kTotal;
    }
}
int wide_tx_resident_syms_per_cta() { return wide::kWTrsSyms; }
int wide_tx_resident_threads() { return wide::kWTrsThreads; }

}  // namespace ofdm
