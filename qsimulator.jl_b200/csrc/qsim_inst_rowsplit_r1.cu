// Instantiates the row-split Tsit5 kernels with ONE row per lane for sixteen-warp teams (512 threads, 128 registers):
// the component passes of large propagators (cfg5: d^2 = 729 in 13 components).  A pass then holds 16 chunks of 32 rows,
// one per warp, instead of 24 on eight warps with three each: the SM runs twice the warps and a warp's share of a stage is
// a third as long.  Opt-in (QSIM_RS_R1=1): parity-green, 2.5 % slower than the eight-warp teams on cfg5.
#include "qsim_registry.h"
namespace qsim {
#ifdef QSIM_FAST_BUILD
const KernelVariant *variants_rowsplit_r1(int *count) {
    *count = 0;
    return nullptr;
}
#else
const KernelVariant *variants_rowsplit_r1(int *count) {
    static const KernelVariant v[] = {
        {1, 1, 3, 512, 1, 0, 1, "tsit5_rowsplit_r1c1_t512_vm0", launch_solver<1, 1, 3, 512, 1, 0, true>, occupancy_solver<1, 1, 3, 512, 1, 0, true>},
        {1, 1, 3, 512, 1, 4, 1, "tsit5_rowsplit_r1c1_t512_vm4", launch_solver<1, 1, 3, 512, 1, 4, true>, occupancy_solver<1, 1, 3, 512, 1, 4, true>},
        {1, 1, 3, 512, 1, 1, 1, "tsit5_rowsplit_r1c1_t512_vm1", launch_solver<1, 1, 3, 512, 1, 1, true>, occupancy_solver<1, 1, 3, 512, 1, 1, true>},
    };
    *count = 3;
    return v;
}
#endif
}  // namespace qsim
