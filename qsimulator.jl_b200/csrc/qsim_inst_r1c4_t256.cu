// Instantiates the Tsit5 solver kernels for 1 row(s) x 4 column(s) per lane, blocks of up to 256 threads
// (25- to 32-level propagators: seven or eight warps per trajectory).
#include "qsim_registry.h"
namespace qsim {
QSIM_DEFINE_VARIANTS(variants_r1c4_t256, 1, 4, 256, 1)
}
