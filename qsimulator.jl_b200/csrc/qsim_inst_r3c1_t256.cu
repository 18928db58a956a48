// Instantiates the Tsit5 solver kernels for 3 row(s) x 1 column(s) per lane, blocks of up to 256 threads.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r3c1_t256, 3, 1, 256, 1)
}
