// one kernel per translation unit: see pqp_kernels.h
#include "pqp_kernels.h"
#include "pqp_gen_core.cuh"

// Generic banded-QP kernel ("K" and "KPC" formulations): one warp per path, sparse data from the host.
__global__ void __launch_bounds__(32)
pqp_gen_solve_kernel(const __grid_constant__ pqp::DevParams prm, const __grid_constant__ pqp::GenView gv, int smem_doubles) {
    extern __shared__ double pqp_smem[];
    pqp::Warp w;
    pqp::gen_solve_qp(w, prm, gv, blockIdx.x, pqp_smem, (size_t)smem_doubles);
}

const void *pqp_gen_kernel_fn() { return (const void *)pqp_gen_solve_kernel; }
