// Instantiates the lean Tsit5 kernels (Kern<..., LEAN = true>): one-warp resident teams, four per CTA, three CTAs =
// 12 warps per SM at no more than 168 registers per thread.  Row-per-lane (value mode 2), row-per-lane with split
// rows (value mode 6) and tiered (value mode 5: three rows of one column per lane, tier capacities 5 / 3 / 2 packed
// as 0532) forms.
#include "qsim_registry.h"
namespace qsim {
const KernelVariant *variants_lean(int *count) {
    static const KernelVariant v[] = {
        {1, 3, 5, 128, 3, 2, 0, "tsit5_lean_r1c3_nnz5_t128_vm2", launch_solver<1, 3, 5, 128, 3, 2, false, true>,
         occupancy_solver<1, 3, 5, 128, 3, 2, false, true>, 1},
        // nine-warp teams (25..27-level propagators, three columns per lane): one team per SM, up to 224 registers
        {1, 3, 9, 288, 1, 1, 0, "tsit5_lean_r1c3_nnz9_t288_vm1", launch_solver<1, 3, 9, 288, 1, 1, false, true>,
         occupancy_solver<1, 3, 9, 288, 1, 1, false, true>, 1},
        {1, 3, 9, 288, 1, 2, 0, "tsit5_lean_r1c3_nnz9_t288_vm2", launch_solver<1, 3, 9, 288, 1, 2, false, true>,
         occupancy_solver<1, 3, 9, 288, 1, 2, false, true>, 1},
        // state evolution of small systems (unitary_state / me_state, one column): 168 registers without spills
#define QSIM_LEAN_R1C1(M, I)                                                                                              \
        {1, 1, M, 128, 3, I, 0, "tsit5_lean_r1c1_nnz" #M "_t128_vm" #I, launch_solver<1, 1, M, 128, 3, I, false, true>,   \
         occupancy_solver<1, 1, M, 128, 3, I, false, true>, 1},
        QSIM_LEAN_R1C1(5, 0) QSIM_LEAN_R1C1(5, 1) QSIM_LEAN_R1C1(5, 2)
        QSIM_LEAN_R1C1(9, 0) QSIM_LEAN_R1C1(9, 1) QSIM_LEAN_R1C1(9, 2) QSIM_LEAN_R1C1(9, 4)
        // additively split rows (value mode 7): the 2-transmon gates as 10 table rows of 1 + 2 entries on 30 lanes
        {1, 3, 3, 128, 3, 7, 0, "tsit5_lean_addsplit_r1c3_nnz3_t128_vm7", launch_solver<1, 3, 3, 128, 3, 7, false, true>,
         occupancy_solver<1, 3, 3, 128, 3, 7, false, true>, 1},
        {1, 3, 3, 128, 3, 6, 0, "tsit5_lean_split_r1c3_nnz3_t128_vm6", launch_solver<1, 3, 3, 128, 3, 6, false, true>,
         occupancy_solver<1, 3, 3, 128, 3, 6, false, true>, 1},
        {3, 1, 0532, 128, 3, 5, 0, "tsit5_lean_tiered_r3c1_nnz532_t128_vm5", launch_solver<3, 1, 0532, 128, 3, 5, false, true>,
         occupancy_solver<3, 1, 0532, 128, 3, 5, false, true>, 1},
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace qsim
