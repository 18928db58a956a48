// Instantiates the Tsit5 solver kernels for 2 row(s) x 1 column(s) per lane.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r2c1, 2, 1)
}
