// Kernel registry: one entry per compiled (rows-per-lane, columns-per-lane, max row length) variant.
#pragma once
#include "qsim_solver.cuh"

namespace qsim {

struct KernelVariant {
    int rpl, cpl, maxnnz, maxt;
    const char *name;
    launch_fn launch;
    cudaError_t (*occupancy)(int block, size_t smem, int *ctas_per_sm);
};

// defined one per translation unit (qsim_inst_*.cu) so the variants compile in parallel
const KernelVariant *variants_r1c1(int *count);
const KernelVariant *variants_r1c3(int *count);
const KernelVariant *variants_r2c1(int *count);
const KernelVariant *variants_r2c2(int *count);
const KernelVariant *variants_r3c1(int *count);

#define QSIM_VARIANT(R, C, M, T) \
    { R, C, M, T, "tsit5_r" #R "c" #C "_nnz" #M "_t" #T, launch_solver<R, C, M, T>, occupancy_solver<R, C, M, T> }

#define QSIM_DEFINE_VARIANTS(FN, R, C)                                                                     \
    const KernelVariant *FN(int *count) {                                                                  \
        static const KernelVariant v[] = {QSIM_VARIANT(R, C, 8, 128), QSIM_VARIANT(R, C, 16, 128),         \
                                          QSIM_VARIANT(R, C, 8, 288), QSIM_VARIANT(R, C, 16, 288)};        \
        *count = 4;                                                                                        \
        return v;                                                                                          \
    }

}  // namespace qsim
