// Instantiates the lean Tsit5 kernels for multi-warp resident teams (one team per SM, up to 255 registers): there is no
// occupancy to gain here -- the shorter live ranges of the lean step remove the spills of the 255-register kernel and
// leave room for the static off-diagonal values (value mode 2) next to four entries per lane, and operands are
// requested four row entries at a time.  25..32-level propagators (cfg3).
#ifndef QSIM_LEAN_CH_EARLY
#define QSIM_LEAN_CH_EARLY 4
#endif
#ifndef QSIM_LEAN_CH_MID
#define QSIM_LEAN_CH_MID 4
#endif
#ifndef QSIM_LEAN_CH
#define QSIM_LEAN_CH 4
#endif
#include "qsim_registry.h"
namespace qsim {
const KernelVariant *variants_lean_t256(int *count) {
    static const KernelVariant v[] = {
        {1, 4, 9, 256, 1, 1, 0, "tsit5_lean_r1c4_nnz9_t256_vm1", launch_solver<1, 4, 9, 256, 1, 1, false, true>,
         occupancy_solver<1, 4, 9, 256, 1, 1, false, true>, 1},
        {1, 4, 9, 256, 1, 2, 0, "tsit5_lean_r1c4_nnz9_t256_vm2", launch_solver<1, 4, 9, 256, 1, 2, false, true>,
         occupancy_solver<1, 4, 9, 256, 1, 2, false, true>, 1},
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace qsim
