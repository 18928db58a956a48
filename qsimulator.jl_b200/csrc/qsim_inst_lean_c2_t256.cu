// Instantiates the lean Tsit5 kernels for multi-warp resident teams of component forms at two CTAs per SM: one row x
// three columns per lane in 168 registers, teams of up to six warps (the 3-transmon gates of cfg3 as five warps: 14 + 13
// rows twice per warp).  Against the four-columns-per-lane kernel (one CTA of two four-warp teams per SM) a step is
// shorter and the SM still holds two teams: 441 ms against 470 ms per 1184 x 50 ns.  (One row x two columns per lane in
// 128 registers, seven-warp teams, measured 484 ms and was dropped again.)
#include "qsim_registry.h"
namespace qsim {
const KernelVariant *variants_lean_c2_t256(int *count) {
    static const KernelVariant v[] = {
#ifndef QSIM_FAST_BUILD
        {1, 3, 9, 192, 2, 1, 0, "tsit5_lean_r1c3_nnz9_t192b2_vm1", launch_solver<1, 3, 9, 192, 2, 1, false, true>,
         occupancy_solver<1, 3, 9, 192, 2, 1, false, true>, 1},
        {1, 3, 9, 192, 2, 2, 0, "tsit5_lean_r1c3_nnz9_t192b2_vm2", launch_solver<1, 3, 9, 192, 2, 2, false, true>,
         occupancy_solver<1, 3, 9, 192, 2, 2, false, true>, 1},
#endif
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace qsim
