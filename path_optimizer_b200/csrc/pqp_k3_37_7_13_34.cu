// one kernel per translation unit: see pqp_kernels.h
// Long paths (257..408 stations at keep_control_steps 3 or 4): thirteen warps, one station per thread, 34 separators
// with 37-unknown interiors, two-level separator system; one CTA per SM with the whole 227 KB of shared memory.  (128 registers per thread: warps 0, 4, 8, 12 share
// one SM sub-partition and its 16 K registers.)
#include "pqp_kernel_tu.cuh"
PQP_KP3_TU(37, 7, 13, 34)
