// tx64r.cu -- instantiates the one-pass, tensor-memory-resident transmit kernel (tx_resident.cuh).
#include "kernels.h"

namespace ofdm {

template <int MOD, int W>
static TxKernel pick_txr_mod(bool guard, bool fec)
{
    if (guard) return fec ? (TxKernel)tx_resident_kernel<MOD, true, true, W> : (TxKernel)tx_resident_kernel<MOD, true, false, W>;
    return fec ? (TxKernel)tx_resident_kernel<MOD, false, true, W> : (TxKernel)tx_resident_kernel<MOD, false, false, W>;
}
template <int W>
static TxKernel pick_txr_w(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return pick_txr_mod<0, W>(c.guard_bands, c.fec);
    case 1: return pick_txr_mod<1, W>(c.guard_bands, c.fec);
    default: return pick_txr_mod<2, W>(c.guard_bands, c.fec);
    }
}
template <int W>
static size_t txr_smem_w(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return c.guard_bands ? TrsSmem<0, true, W>::kTotal : TrsSmem<0, false, W>::kTotal;
    case 1: return c.guard_bands ? TrsSmem<1, true, W>::kTotal : TrsSmem<1, false, W>::kTotal;
    default: return c.guard_bands ? TrsSmem<2, true, W>::kTotal : TrsSmem<2, false, W>::kTotal;
    }
}
// One CTA of 32 warps per SM. (The kernel is templated on the warps per CTA W, with 32 / W CTAs per SM sharing the SM's tensor
// memory; W = 8 and W = 16 were measured on the bench workload: 1.50 and 1.64 ms against 1.48 ms for W = 32.)
TxKernel pick_tx_resident(const ofdm_cfg &c) { return pick_txr_w<kTrsWarpsPerCta>(c); }
size_t tx_resident_smem(const ofdm_cfg &c) { return txr_smem_w<kTrsWarpsPerCta>(c); }

}  // namespace ofdm
