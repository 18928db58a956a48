// Instantiates the Tsit5 solver kernels for 2 row(s) x 2 column(s) per lane, blocks of up to 160 threads
// (one-warp teams, five per CTA, two CTAs per SM).
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r2c2_t160, 2, 2, 160, 2)
}
