// Kernel registry: one entry per compiled (rows-per-lane, columns-per-lane, max row length) variant.
#pragma once
#include "qsim_solver.cuh"

namespace qsim {

struct KernelVariant {
    int rpl, cpl, maxnnz;
    const char *name;
    launch_fn launch;
    cudaError_t (*occupancy)(int block, size_t smem, int *ctas_per_sm);
    cudaError_t (*attributes)(cudaFuncAttributes *attr);
};

// defined one per translation unit (qsim_inst_*.cu) so the variants compile in parallel
const KernelVariant *variants_r1c1(int *count);
const KernelVariant *variants_r1c3(int *count);
const KernelVariant *variants_r2c1(int *count);
const KernelVariant *variants_r2c2(int *count);
const KernelVariant *variants_r3c1(int *count);

#define QSIM_DEFINE_VARIANTS(FN, R, C)                                                                          \
    template <int M>                                                                                            \
    static cudaError_t attr_##FN(cudaFuncAttributes *attr) {                                                    \
        return cudaFuncGetAttributes(attr, solve_kernel<R, C, M>);                                              \
    }                                                                                                           \
    const KernelVariant *FN(int *count) {                                                                       \
        static const KernelVariant v[] = {                                                                      \
            {R, C, 8, "tsit5_r" #R "c" #C "_nnz8", launch_solver<R, C, 8>, occupancy_solver<R, C, 8>, attr_##FN<8>},    \
            {R, C, 16, "tsit5_r" #R "c" #C "_nnz16", launch_solver<R, C, 16>, occupancy_solver<R, C, 16>, attr_##FN<16>}, \
        };                                                                                                      \
        *count = 2;                                                                                             \
        return v;                                                                                               \
    }

}  // namespace qsim
