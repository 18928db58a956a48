// Instantiates the column-per-lane Tsit5 kernels (propagators with 17..32 rows, e.g. the d = 27 three-transmon gate):
// the row-split machinery with lane = column and four whole rows per warp.
#include "qsim_registry.h"
namespace qsim {
#ifdef QSIM_FAST_BUILD
const KernelVariant *variants_colperlane(int *count) {
    *count = 0;
    return nullptr;
}
#else
const KernelVariant *variants_colperlane(int *count) {
    static const KernelVariant v[] = {QSIM_VARIANT_CU(256, 0), QSIM_VARIANT_CU(256, 1), QSIM_VARIANT_CU(256, 4)};
    *count = 3;
    return v;
}
#endif
}
