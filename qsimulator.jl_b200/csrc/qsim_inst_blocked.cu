// Instantiates the blocked Tsit5 kernels (value mode 3): 3 consecutive rows x 1 column per lane, the off-diagonal
// part of the generator as dense 3 x 3 blocks with static values in registers; capacity 1 + 2 or 1 + 4 blocks.
#include "qsim_registry.h"
namespace qsim {
#ifdef QSIM_FAST_BUILD
const KernelVariant *variants_blocked(int *count) {
    *count = 0;
    return nullptr;
}
#else
const KernelVariant *variants_blocked(int *count) {
    static const KernelVariant v[] = {QSIM_VARIANT(3, 1, 3, 128, 2, 3), QSIM_VARIANT(3, 1, 5, 128, 2, 3)};
    *count = 2;
    return v;
}
#endif
}
