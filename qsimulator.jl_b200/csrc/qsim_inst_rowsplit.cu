// Instantiates the row-split Tsit5 kernels (state vectors longer than 96 entries, e.g. d^2 = 729).
#include "qsim_registry.h"
namespace qsim {
#ifdef QSIM_FAST_BUILD
const KernelVariant *variants_rowsplit(int *count) {
    static const KernelVariant v[] = {QSIM_VARIANT_RS(256, 0)};
    *count = 1;
    return v;
}
#else
const KernelVariant *variants_rowsplit(int *count) {
    static const KernelVariant v[] = {QSIM_VARIANT_RS(256, 0), QSIM_VARIANT_RS(256, 1), QSIM_VARIANT_RS(256, 4),
                                      QSIM_VARIANT_RS(288, 0), QSIM_VARIANT_RS(288, 1), QSIM_VARIANT_RS(288, 4)};
    *count = 6;
    return v;
}
#endif
}
