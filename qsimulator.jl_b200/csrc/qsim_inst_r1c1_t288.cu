// Instantiates the Tsit5 solver kernels for 1 row(s) x 1 column(s) per lane, blocks of up to 288 threads.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c1_t288, 1, 1, 288, 1)
}
