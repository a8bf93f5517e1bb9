// wide_rx.cu -- maps an engine configuration to the nfft = 1024 receive kernels (wide_kernels.cuh); the acquisition
// kernels are instantiated here, the decode kernels in wide_rx_m{0,1,2}.cu.
#include "kernels.h"

namespace ofdm {

WDecodeKernel wpick_decode(const ofdm_cfg &c, bool points)
{
    switch (c.modulation) {
    case 0: return wpick_decode_mod<0>(c.guard_bands, c.fec, c.phase_mode, points);
    case 1: return wpick_decode_mod<1>(c.guard_bands, c.fec, c.phase_mode, points);
    default: return wpick_decode_mod<2>(c.guard_bands, c.fec, c.phase_mode, points);
    }
}
template <int MOD, bool GUARD>
static WDecodeKernel wpick_acquire_phase(int phase)
{
    return phase ? (WDecodeKernel)wide::wide_acquire_kernel<MOD, GUARD, 1> : (WDecodeKernel)wide::wide_acquire_kernel<MOD, GUARD, 0>;
}
WDecodeKernel wpick_acquire(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return c.guard_bands ? wpick_acquire_phase<0, true>(c.phase_mode) : wpick_acquire_phase<0, false>(c.phase_mode);
    case 1: return c.guard_bands ? wpick_acquire_phase<1, true>(c.phase_mode) : wpick_acquire_phase<1, false>(c.phase_mode);
    default: return c.guard_bands ? wpick_acquire_phase<2, true>(c.phase_mode) : wpick_acquire_phase<2, false>(c.phase_mode);
    }
}

}  // namespace ofdm
