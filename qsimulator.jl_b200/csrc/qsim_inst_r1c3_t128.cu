// Instantiates the Tsit5 solver kernels for 1 row(s) x 3 column(s) per lane, blocks of up to 128 threads
// (one-warp teams, four per CTA, two CTAs = 8 warps per SM, no register spills): the 255-register forms, used when
// the shape has no lean variant (qsim_inst_lean.cu) or with QSIM_NO_LEAN=1.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c3_t128, 1, 3, 128, 2)
}
