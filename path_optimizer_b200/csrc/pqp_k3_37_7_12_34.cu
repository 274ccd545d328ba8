// one kernel per translation unit: see pqp_kernels.h
// Long paths up to 384 stations: twelve warps (three per SM sub-partition: 168 registers per thread instead of the 128
// the thirteen-warp class is held to), otherwise as pqp_k3_37_7_13_34.cu.
#include "pqp_kernel_tu.cuh"
PQP_KP3_TU(37, 7, 12, 34)
