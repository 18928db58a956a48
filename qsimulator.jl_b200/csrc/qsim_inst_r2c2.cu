// Instantiates the Tsit5 solver kernels for 2 row(s) x 2 column(s) per lane.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r2c2, 2, 2)
}
