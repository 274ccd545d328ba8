// one kernel per translation unit: see pqp_kernels.h
#include "pqp_kernel_tu.cuh"
PQP_KP2_TU(49, 7)
