// Instantiates the lean Tsit5 kernels (Kern<..., LEAN = true>): one-warp resident teams, four per CTA, three CTAs =
// 12 warps per SM at no more than 168 registers per thread.
#include "qsim_registry.h"
namespace qsim {
const KernelVariant *variants_lean(int *count) {
    static const KernelVariant v[] = {
        {1, 3, 5, 128, 3, 2, 0, "tsit5_lean_r1c3_nnz5_t128_vm2", launch_solver<1, 3, 5, 128, 3, 2, false, true>,
         occupancy_solver<1, 3, 5, 128, 3, 2, false, true>, 1},
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace qsim
