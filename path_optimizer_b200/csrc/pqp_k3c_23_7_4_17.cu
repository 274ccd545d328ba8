// one kernel per translation unit: see pqp_kernels.h
// "KPC" (SolverKpAsInputConstrained, keep_control_steps 4) on the thread-per-station skeleton: up to 128 stations, four warps, two CTAs per SM.
#include "pqp_kernel_tu.cuh"
PQP_KP3C_TU(23, 7, 4, 17)
