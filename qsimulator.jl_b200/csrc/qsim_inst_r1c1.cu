// Instantiates the Tsit5 solver kernels for 1 row(s) x 1 column(s) per lane.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c1, 1, 1)
}
