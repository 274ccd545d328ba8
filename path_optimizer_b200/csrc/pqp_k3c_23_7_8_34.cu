// one kernel per translation unit: see pqp_kernels.h
// "KPC" (SolverKpAsInputConstrained, keep_control_steps 4) on the thread-per-station skeleton: 129..256 stations, eight warps, 34 separators.
#include "pqp_kernel_tu.cuh"
PQP_KP3C_TU(23, 7, 8, 34)
