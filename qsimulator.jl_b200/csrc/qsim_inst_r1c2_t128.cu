// Instantiates the Tsit5 solver kernels for 1 row(s) x 2 column(s) per lane, blocks of up to 128 threads
// (one-warp teams, four per CTA): the 255-register forms, used by component forms (qsim_components.cpp) whose shape
// has no lean variant (qsim_inst_lean_c2.cu) or with QSIM_NO_LEAN=1.
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c2_t128, 1, 2, 128, 2)
}
