// Instantiates the Tsit5 solver kernels for 1 row(s) x 3 column(s) per lane.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c3, 1, 3)
}
