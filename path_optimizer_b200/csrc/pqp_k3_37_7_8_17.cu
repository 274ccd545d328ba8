// one kernel per translation unit: see pqp_kernels.h
#include "pqp_kernel_tu.cuh"
PQP_KP3_TU(37, 7, 8, 17)
