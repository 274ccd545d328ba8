// one kernel per translation unit: see pqp_kernels.h
// 257..306 stations at keep_control_steps 3: ten warps, 34 separators with 27-unknown interiors (a class costs what its
// interiors cost: the 37-unknown classes take the same time at 300 stations as at 384).
#include "pqp_kernel_tu.cuh"
PQP_KP3_TU(27, 7, 10, 34)
