// Instantiates the Tsit5 solver kernels for 1 row(s) x 1 column(s) per lane, blocks of up to 128 threads
// (one-warp teams, four per CTA, two CTAs = 8 warps per SM, no register spills).
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c1_t128, 1, 1, 128, 2)
}
