// Instantiates the Tsit5 solver kernels for 3 row(s) x 1 column(s) per lane, blocks of up to 160 threads
// (one-warp teams, five per CTA, two CTAs per SM).
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r3c1_t160, 3, 1, 160, 2)
}
