// pqp_kk_tu.cuh -- kernel entry + variant record of one "K" class (pqp_kk_core.cuh): NW warps per path, N <= 32*NW stations.
#pragma once
#include "pqp_kernels.h"
#include "pqp_kk_core.cuh"

#ifndef PQP_KK_MINBLOCKS4
#define PQP_KK_MINBLOCKS4 2      // resident CTAs per SM the four-warp class is compiled for (register cap 255 / 168 / 128)
#endif
#define PQP_KK_MINBLOCKS(NW) ((NW) <= 4 ? PQP_KK_MINBLOCKS4 : 1)
template <int NW>
__global__ void __launch_bounds__(NW * 32, PQP_KK_MINBLOCKS(NW))
pqp_kk_solve_kernel(const __grid_constant__ pqp::DevParams prm, const __grid_constant__ pqp::BatchView bv,
                    const int32_t *__restrict__ order, int smem_doubles) {
    extern __shared__ double pqp_smem[];
    int prob = blockIdx.x;
    if (order) prob = order[prob];
    pqp::Cta c{pqp::Warp(), pqp::CtaSync(), (int)(threadIdx.x >> 5), NW, pqp_smem};
    constexpr int kRes = pqp::Kk<NW>::kCtaScratch;
    pqp::Kk<NW>::solve_path(c, prm, bv, prob, pqp_smem + kRes, (size_t)smem_doubles - kRes);
}

#define PQP_KK_TU(W)                                                                                              \
    static size_t tu_smem(int n, int) { return (pqp::Kk<W>::kCtaScratch + pqp::Kk<W>::smem_doubles(n)) * sizeof(double); } \
    static bool tu_fits(int n, int keep) { return pqp::Kk<W>::fits(n, keep); }                                    \
    void pqp_variant_kk_##W(PqpVariant *out) {                                                                    \
        *out = PqpVariant{3, 3, W * 32, (const void *)pqp_kk_solve_kernel<W>, tu_smem, tu_fits,                   \
                          "pqp_kk_solve_kernel<" #W ">"};                                                         \
    }
