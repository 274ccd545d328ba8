// pqp_kernel_tu.cuh -- kernel entry templates + the macros that turn a translation unit into exactly one
// compiled kernel and its PqpVariant record (see pqp_kernels.h for why there is one kernel per file).
#pragma once
#include "pqp_kernels.h"
#include "pqp_kp_core3.cuh"

#ifndef PQP_KP3_MINBLOCKS
#define PQP_KP3_MINBLOCKS 2
#endif
// Thread-per-station kernels (pqp_kp_core3.cuh): NW warps per path, N <= 32*NW stations.
#ifndef PQP_KP3_MAXNREG
#define PQP_KP3_MAXNREG 0
#endif
template <int IMAX, int BW, int NW, int MMAX, int FORM = 0>
__global__ void
#if PQP_KP3_MAXNREG
__maxnreg__(PQP_KP3_MAXNREG)   // (diagnostics: registers are per SM sub-partition, 16 K for the warps that share one)
#else
__launch_bounds__(NW * 32, NW <= 4 ? PQP_KP3_MINBLOCKS : 1)
#endif
pqp_kp3_solve_kernel(const __grid_constant__ pqp::DevParams prm, const __grid_constant__ pqp::BatchView bv,
                     const int32_t *__restrict__ order, int smem_doubles) {
    extern __shared__ double pqp_smem[];
    int prob = blockIdx.x;
    if (order) prob = order[prob];
    pqp::Cta c{pqp::Warp(), pqp::CtaSync(), (int)(threadIdx.x >> 5), NW, pqp_smem};
    constexpr int kRes = pqp::Kp3<IMAX, BW, NW, MMAX, FORM>::kCtaScratch;
    pqp::Kp3<IMAX, BW, NW, MMAX, FORM>::solve_path(c, prm, bv, prob, pqp_smem + kRes, (size_t)smem_doubles - kRes);
}

#define PQP_KP3_TU(I, B, W, MM)                                                                                   \
    static size_t tu_smem(int n, int keep) {                                                                      \
        return (pqp::Kp3<I, B, W, MM>::kCtaScratch + pqp::Kp3<I, B, W, MM>::smem_doubles(pqp::Kp3<I, B, W, MM>::dims(n, keep))) * sizeof(double); \
    }                                                                                                             \
    static bool tu_fits(int n, int keep) { return pqp::Kp3<I, B, W, MM>::fits(n, keep); }                         \
    void pqp_variant_k3_##I##_##B##_##W##_##MM(PqpVariant *out) {                                                 \
        *out = PqpVariant{I, B, W * 32, (const void *)pqp_kp3_solve_kernel<I, B, W, MM>, tu_smem, tu_fits,        \
                          "pqp_kp3_solve_kernel<" #I "," #B "," #W "," #MM ">"};                                  \
    }

// The same for the "KPC" formulation (SolverKpAsInputConstrained): Kp3<..., FORM = 2>, keep_control_steps fixed at 4.
#define PQP_KP3C_TU(I, B, W, MM)                                                                                  \
    typedef pqp::Kp3<I, B, W, MM, 2> TuK;                                                                         \
    static size_t tu_smem(int n, int keep) { return (TuK::kCtaScratch + TuK::smem_doubles(TuK::dims(n, keep))) * sizeof(double); } \
    static bool tu_fits(int n, int keep) { return keep == 4 && TuK::fits(n, keep); }                              \
    void pqp_variant_k3c_##I##_##B##_##W##_##MM(PqpVariant *out) {                                                \
        *out = PqpVariant{I, B, W * 32, (const void *)pqp_kp3_solve_kernel<I, B, W, MM, 2>, tu_smem, tu_fits,     \
                          "pqp_kp3_solve_kernel<" #I "," #B "," #W "," #MM ",KPC>"};                              \
    }
