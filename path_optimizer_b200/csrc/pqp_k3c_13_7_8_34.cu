// one kernel per translation unit: see pqp_kernels.h
// "KPC" (SolverKpAsInputConstrained, keep_control_steps 4) on the thread-per-station skeleton: up to 136 stations with a
// separator every 4 stations (10-unknown interiors, dense inverses), 34 separators, two-level separator system.
#include "pqp_kernel_tu.cuh"
PQP_KP3C_TU(13, 7, 8, 34)
