// Instantiates the Tsit5 solver kernels for 1 row(s) x 3 column(s) per lane, blocks of up to 256 threads.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c3_t256, 1, 3, 256, 1)
}
