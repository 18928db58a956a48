// Instantiates the Tsit5 solver kernels for 1 row(s) x 3 column(s) per lane, blocks of up to 160 threads
// (one-warp teams, five per CTA, two CTAs per SM).
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c3_t160, 1, 3, 160, 2)
}
