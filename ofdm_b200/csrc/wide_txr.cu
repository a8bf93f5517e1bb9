// wide_txr.cu -- instantiates the one-pass, tensor-memory-resident transmit kernel of the nfft = 1024 variant (wide_tx_resident.cuh).
#include "kernels.h"
#include "wide_tx_resident.cuh"

namespace ofdm {

template <int MOD, bool DB>
static WTxKernel wpick_txr_mod(bool guard, bool fec)
{
    if (guard) return fec ? (WTxKernel)wide::wide_tx_resident_kernel<MOD, true, true, DB> : (WTxKernel)wide::wide_tx_resident_kernel<MOD, true, false, DB>;
    return fec ? (WTxKernel)wide::wide_tx_resident_kernel<MOD, false, true, DB> : (WTxKernel)wide::wide_tx_resident_kernel<MOD, false, false, DB>;
}
template <bool DB>
static WTxKernel wpick_txr_db(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return wpick_txr_mod<0, DB>(c.guard_bands, c.fec);
    case 1: return wpick_txr_mod<1, DB>(c.guard_bands, c.fec);
    default: return wpick_txr_mod<2, DB>(c.guard_bands, c.fec);
    }
}
// double_buffered: one symbol per warp and frame, two frames in flight per CTA (16 symbols per CTA and frame); else 32
WTxKernel wpick_tx_resident(const ofdm_cfg &c, bool double_buffered) { return double_buffered ? wpick_txr_db<true>(c) : wpick_txr_db<false>(c); }
template <int MOD>
static WTxKernel wpick_txs_mod(bool guard, bool fec)
{
    if (guard) return fec ? (WTxKernel)wide::wide_tx_spec_kernel<MOD, true, true> : (WTxKernel)wide::wide_tx_spec_kernel<MOD, true, false>;
    return fec ? (WTxKernel)wide::wide_tx_spec_kernel<MOD, false, true> : (WTxKernel)wide::wide_tx_spec_kernel<MOD, false, false>;
}
WTxKernel wpick_tx_spec(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return wpick_txs_mod<0>(c.guard_bands, c.fec);
    case 1: return wpick_txs_mod<1>(c.guard_bands, c.fec);
    default: return wpick_txs_mod<2>(c.guard_bands, c.fec);
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
int wide_tx_resident_syms_per_cta(bool double_buffered) { return double_buffered ? wide::kWTrsWarps : wide::kWTrsSyms; }
int wide_tx_resident_threads() { return wide::kWTrsThreads; }

}  // namespace ofdm
static_assert(ofdm::wide::WTrsSmem<2>::kTotal <= 232448, "wide_tx_resident_kernel: dynamic shared memory of one CTA");
