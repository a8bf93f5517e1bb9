// rx64.cu -- maps an engine configuration to the nfft = 64 receive kernels (rx_kernels.cuh); the acquisition kernels are
// instantiated here, the decode kernels in rx64_m{0,1,2}.cu (one translation unit per modulation).
#include "kernels.h"

namespace ofdm {

DecodeKernel pick_decode(const ofdm_cfg &c, bool points)
{
    switch (c.modulation) {
    case 0: return pick_decode_mod<0>(c.guard_bands, c.fec, c.phase_mode, points);
    case 1: return pick_decode_mod<1>(c.guard_bands, c.fec, c.phase_mode, points);
    default: return pick_decode_mod<2>(c.guard_bands, c.fec, c.phase_mode, points);
    }
}

template <int MOD, bool GUARD>
static AcquireKernel pick_acquire_phase(int phase)
{
    return phase ? (AcquireKernel)rx_acquire_kernel<MOD, GUARD, 1> : (AcquireKernel)rx_acquire_kernel<MOD, GUARD, 0>;
}
AcquireKernel pick_acquire(const ofdm_cfg &c)
{
    switch (c.modulation) {
    case 0: return c.guard_bands ? pick_acquire_phase<0, true>(c.phase_mode) : pick_acquire_phase<0, false>(c.phase_mode);
    case 1: return c.guard_bands ? pick_acquire_phase<1, true>(c.phase_mode) : pick_acquire_phase<1, false>(c.phase_mode);
    default: return c.guard_bands ? pick_acquire_phase<2, true>(c.phase_mode) : pick_acquire_phase<2, false>(c.phase_mode);
    }
}

}  // namespace ofdm
