// one kernel per translation unit: see pqp_kernels.h
// "K" (SolverKAsInput) on the thread-per-station block-cyclic-reduction kernel: up to 416 stations, 13 warps.
#include "pqp_kk_tu.cuh"
PQP_KK_TU(13)
