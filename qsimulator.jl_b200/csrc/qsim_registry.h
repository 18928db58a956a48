// Kernel registry: one entry per compiled (rows-per-lane, columns-per-lane, max row length) variant.
#pragma once
#include "qsim_solver.cuh"

namespace qsim {

struct KernelVariant {
    int rpl, cpl, maxnnz, maxt, minb, vm, rs;  // vm: value mode (0 complex, 1 imaginary, 2 imaginary + static off-diagonals,
                                               // 3 blocked, 4 mixed real/imaginary entries; opt-in experiments: 5 tiered
                                               // rows (maxnnz packs the tier capacities), 6 split rows, 7 additively split rows)
    const char *name;
    launch_fn launch;
    cudaError_t (*occupancy)(int block, size_t smem, int *ctas_per_sm);
    int lean;  // 1: Kern<..., LEAN = true> (resident teams only; see qsim_solver.cuh)
};

// defined one per translation unit (qsim_inst_*.cu) so the variants compile in parallel
#define QSIM_DECLARE(FN) const KernelVariant *FN(int *count);
// t128: four one-warp teams per CTA, two CTAs per SM.  t256: teams of 2-8 warps, one CTA per SM.  Both leave
// 255 registers per thread (two warps per scheduler).  The lean kernels (variants_lean*, qsim_inst_lean*.cu) fit the
// one-warp teams into 168 registers: three CTAs, 12 warps per SM.
QSIM_DECLARE(variants_r1c1_t128)
QSIM_DECLARE(variants_r1c3_t128) QSIM_DECLARE(variants_r1c3_t256)
QSIM_DECLARE(variants_r1c4_t256)
QSIM_DECLARE(variants_r2c1_t128)
QSIM_DECLARE(variants_r2c2_t256)  // 33-64 level propagators always stream their columns through an 8-warp team
QSIM_DECLARE(variants_r3c1_t128) QSIM_DECLARE(variants_r3c1_t256)
QSIM_DECLARE(variants_rowsplit)
QSIM_DECLARE(variants_colperlane)
QSIM_DECLARE(variants_blocked)
QSIM_DECLARE(variants_lean)
QSIM_DECLARE(variants_lean_t256)
QSIM_DECLARE(variants_lean_c2)
QSIM_DECLARE(variants_lean_c2_t256)

#define QSIM_VARIANT(R, C, M, T, B, I)                                                                    \
    {                                                                                                     \
        R, C, M, T, B, I, 0, "tsit5_r" #R "c" #C "_nnz" #M "_t" #T "_vm" #I,                                \
            launch_solver<R, C, M, T, B, I, false>, occupancy_solver<R, C, M, T, B, I, false>               \
    }
// row-split variants: rows spread over the team's warps, sparse rows from a table in shared (or global) memory;
// the per-lane `ell` registers hold {table offset, row length, unit position (component passes)} of each of the lane's three chunks
#define QSIM_VARIANT_RS(T, I)                                                                             \
    {                                                                                                     \
        3, 1, 3, T, 1, I, 1, "tsit5_rowsplit_r3c1_t" #T "_vm" #I, launch_solver<3, 1, 3, T, 1, I, true>,      \
            occupancy_solver<3, 1, 3, T, 1, I, true>                                                      \
    }
// column-per-lane variants (propagators with 17..32 rows): four whole rows per warp, lane = column
#define QSIM_VARIANT_CU(T, I)                                                                             \
    {                                                                                                     \
        4, 1, 2, T, 1, I, 1, "tsit5_colperlane_r4c1_t" #T "_vm" #I, launch_solver<4, 1, 2, T, 1, I, true>,    \
            occupancy_solver<4, 1, 2, T, 1, I, true>                                                      \
    }

#ifdef QSIM_FAST_BUILD
// development builds (make EXTRA=-DQSIM_FAST_BUILD): only the capacity-5 value-mode-2 kernels (cfg1/cfg2)
#define QSIM_DEFINE_VARIANTS(FN, R, C, T, B)                                                              \
    const KernelVariant *FN(int *count) {                                                                 \
        static const KernelVariant v[] = {QSIM_VARIANT(R, C, 5, T, B, 2)};                                \
        *count = 1;                                                                                       \
        return v;                                                                                         \
    }
#else
// Row capacities 5 / 9 / 12 / 16: the generator rows are padded to the variant's capacity and applied
// without per-entry branches (5 and 9 are the exact row lengths of the 2- and 3-transmon gate Hamiltonians).
#define QSIM_DEFINE_VARIANTS(FN, R, C, T, B)                                                              \
    const KernelVariant *FN(int *count) {                                                                 \
        static const KernelVariant v[] = {                                                                \
            QSIM_VARIANT(R, C, 5, T, B, 0),  QSIM_VARIANT(R, C, 5, T, B, 1),  QSIM_VARIANT(R, C, 5, T, B, 2),   \
            QSIM_VARIANT(R, C, 9, T, B, 0),  QSIM_VARIANT(R, C, 9, T, B, 1),  QSIM_VARIANT(R, C, 9, T, B, 2),   \
            QSIM_VARIANT(R, C, 9, T, B, 4),                                                               \
            QSIM_VARIANT(R, C, 12, T, B, 0), QSIM_VARIANT(R, C, 12, T, B, 1), QSIM_VARIANT(R, C, 12, T, B, 4), \
            QSIM_VARIANT(R, C, 16, T, B, 0), QSIM_VARIANT(R, C, 16, T, B, 1), QSIM_VARIANT(R, C, 16, T, B, 4)}; \
        *count = (int)(sizeof(v) / sizeof(v[0]));                                                         \
        return v;                                                                                         \
    }
#endif

}  // namespace qsim
