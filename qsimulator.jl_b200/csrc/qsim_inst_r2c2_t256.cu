// Instantiates the Tsit5 solver kernels for 2 row(s) x 2 column(s) per lane, blocks of up to 256 threads.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r2c2_t256, 2, 2, 256, 1)
}
