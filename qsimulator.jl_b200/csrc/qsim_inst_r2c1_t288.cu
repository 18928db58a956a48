// Instantiates the Tsit5 solver kernels for 2 row(s) x 1 column(s) per lane, blocks of up to 288 threads.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r2c1_t288, 2, 1, 288, 1)
}
