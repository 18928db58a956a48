// Instantiates the lean Tsit5 kernels (Kern<..., LEAN = true>) for one row x two columns per lane: one-warp resident
// teams of component forms (qsim_components.cpp) -- the 2-transmon gates of cfg1/cfg2 are two blocks of 5 x 5 and 4 x 4
// entries: 15 + 8 lanes with two entries each.  Four teams per CTA; MINB = 4 asks for 128 registers (four CTAs = 16
// warps per SM), MINB = 3 for 168 (12 warps per SM).
#include "qsim_registry.h"
namespace qsim {
const KernelVariant *variants_lean_c2(int *count) {
    static const KernelVariant v[] = {
#define QSIM_LEAN_R1C2(M, I, B)                                                                                                \
        {1, 2, M, 128, B, I, 0, "tsit5_lean_r1c2_nnz" #M "_t128b" #B "_vm" #I, launch_solver<1, 2, M, 128, B, I, false, true>, \
         occupancy_solver<1, 2, M, 128, B, I, false, true>, 1},
        QSIM_LEAN_R1C2(5, 2, 3) QSIM_LEAN_R1C2(5, 2, 4)
#ifndef QSIM_FAST_BUILD
        QSIM_LEAN_R1C2(5, 0, 3) QSIM_LEAN_R1C2(5, 1, 3)
        QSIM_LEAN_R1C2(9, 0, 3) QSIM_LEAN_R1C2(9, 1, 3) QSIM_LEAN_R1C2(9, 2, 3) QSIM_LEAN_R1C2(9, 4, 3)
#endif
    };
    *count = (int)(sizeof(v) / sizeof(v[0]));
    return v;
}
}  // namespace qsim
