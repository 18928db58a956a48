// Instantiates the Tsit5 solver kernels for 2 row(s) x 2 column(s) per lane, blocks of up to 128 threads
// (one-warp teams, four per CTA, two CTAs = 8 warps per SM, no register spills).
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r2c2_t128, 2, 2, 128, 2)
}
