// pqp_kp_core2.cuh -- lane-private banded LDL' of one interior: the unrolled factor / solve routines the
// thread-per-station kernel (pqp_kp_core3.cuh) runs on its solver threads.
//
// An interior (the unknowns between two separator stations, in station order: half-bandwidth <= BW) keeps its factor
// in a compact, lane-private, bank-conflict-free column of shared memory: fac[(k * (BW + 1) + d) * Mst + lane] holds
// L[k][k-d] (d >= 1) and 1 / D[k] (d = 0).  Everything is fully unrolled over the compile-time (IMAX, BW): the vector
// lives in registers, factor entries stream from shared memory with no dependent address arithmetic, and the serial
// chain of a substitution is ONE fused multiply-add per pivot.
// (Round 1's chunked warp-per-path kernel that used to live in this file was retired in round 2: every path it took
// now runs on a thread-per-station class.)
//
// Reference being replaced: the QDLDL solve inside OSQP, src/solver/solver.cpp:66-74.
#pragma once
#include "pqp_kp_core.cuh"

#if defined(__CUDACC__) && !defined(PQP_HOST_EMU)
#define PQP_NOINLINE __device__ __noinline__
#else
#define PQP_NOINLINE
#endif

namespace pqp {

constexpr int kRed2 = 30;  // scratch per separator during the separator-system factorisation: Sinv[9] | Off[9] | G[9] | g[3]

// TAG only separates the out-of-line helpers of different users: each kernel gets its own copies, so the register
// needs of one kernel do not leak into another through a shared callee.
template <int IMAX, int BW, int TAG = 0>
struct Kp2 {
#define PQP_F(k, dd) fcol[((k) * (BW + 1) + (dd)) * Mst]

    // LDL' of this lane's interior in place (compact column).  Rows are processed top down; row k
    // needs rows k-BW..k-1 (already final) -> every address is a compile-time multiple of Mst.
    // (kept out of line: fully unrolled, it would otherwise drag the whole kernel's register allocation)
    PQP_NOINLINE static int local_factor(double *fcol, int Mst) {
        int ok = 1;
#pragma unroll
        for (int k = 0; k < IMAX; ++k) {
            double u[BW + 1];
#pragma unroll
            for (int dd = 0; dd <= BW; ++dd) u[dd] = (k - dd >= 0) ? PQP_F(k, dd) : 0.0;
            // u[dd] becomes L[k][k-dd] * D[k-dd]
#pragma unroll
            for (int dd = BW; dd >= 1; --dd) {
                const int j = k - dd;
                if (j < 0) continue;
#pragma unroll
                for (int e = BW; e > dd; --e) {
                    const int m = k - e;  // m < j
                    if (m < 0 || j - m > BW) continue;
                    u[dd] -= u[e] * PQP_F(j, j - m);
                }
            }
            double dk = u[0];
#pragma unroll
            for (int dd = 1; dd <= BW; ++dd) {
                if (k - dd < 0) continue;
                const double l = u[dd] * PQP_F(k - dd, 0);
                dk -= u[dd] * l;
                PQP_F(k, dd) = l;
            }
            if (!(dk > 0.0)) ok = 0;
            PQP_F(k, 0) = 1.0 / dk;
        }
        return ok;
    }

    // y = K_I^-1 r with separate input / output vectors (unit stride).
    PQP_NOINLINE static void local_solve2(const double *in, double *outv, const double *fcol, int Mst) {
        double y[IMAX];
#pragma unroll
        for (int k = 0; k < IMAX; ++k) y[k] = in[k];
        // forward: y[k] -= sum_d L[k][k-d] y[k-d].  Only the d = 1 term depends on the pivot just
        // computed, so the other terms are accumulated in two independent partial sums first and the
        // serial chain is ONE DFMA per pivot (instead of BW).
#pragma unroll
        for (int k = 1; k < IMAX; ++k) {
            double p0 = y[k], p1 = 0.0;
#pragma unroll
            for (int dd = BW; dd >= 2; --dd)
                if (k - dd >= 0) {
                    if (dd & 1) p1 -= PQP_F(k, dd) * y[k - dd];
                    else p0 -= PQP_F(k, dd) * y[k - dd];
                }
            y[k] = (p0 + p1) - PQP_F(k, 1) * y[k - 1];
        }
#pragma unroll
        for (int k = IMAX - 1; k >= 0; --k) {
            double p0 = y[k] * PQP_F(k, 0), p1 = 0.0;
#pragma unroll
            for (int dd = BW; dd >= 2; --dd)
                if (k + dd < IMAX) {
                    if (dd & 1) p1 -= PQP_F(k + dd, dd) * y[k + dd];
                    else p0 -= PQP_F(k + dd, dd) * y[k + dd];
                }
            double r = p0 + p1;
            if (k + 1 < IMAX) r -= PQP_F(k + 1, 1) * y[k + 1];
            y[k] = r;
        }
#pragma unroll
        for (int k = 0; k < IMAX; ++k) outv[k] = y[k];
    }

    // Column j of K_I^-1 into registers: x = K_I^-1 e_j (the forward sweep starts at row j).  Used by
    // the thread-per-station kernel to build the dense interior inverses at a refactorisation.
    PQP_NOINLINE static void local_solve_unit(int j, double *x, const double *fcol, int Mst) {
        double y[IMAX];
#pragma unroll
        for (int k = 0; k < IMAX; ++k) y[k] = (k == j) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 1; k < IMAX; ++k) {
            double p0 = y[k], p1 = 0.0;
#pragma unroll
            for (int dd = BW; dd >= 2; --dd)
                if (k - dd >= 0) {
                    if (dd & 1) p1 -= PQP_F(k, dd) * y[k - dd];
                    else p0 -= PQP_F(k, dd) * y[k - dd];
                }
            y[k] = (p0 + p1) - PQP_F(k, 1) * y[k - 1];
        }
#pragma unroll
        for (int k = IMAX - 1; k >= 0; --k) {
            double p0 = y[k] * PQP_F(k, 0), p1 = 0.0;
#pragma unroll
            for (int dd = BW; dd >= 2; --dd)
                if (k + dd < IMAX) {
                    if (dd & 1) p1 -= PQP_F(k + dd, dd) * y[k + dd];
                    else p0 -= PQP_F(k + dd, dd) * y[k + dd];
                }
            double r = p0 + p1;
            if (k + 1 < IMAX) r -= PQP_F(k + 1, 1) * y[k + 1];
            y[k] = r;
        }
#pragma unroll
        for (int k = 0; k < IMAX; ++k) x[k] = y[k];
    }

    // The same two solves IN PLACE in shared memory (vector at v[0..IMAX), unit stride), for the refactorisation: the
    // vector never sits in registers, so these out-of-line calls need a handful of registers and the call does not push
    // the caller's (255 live registers) onto the local-memory stack.  (With the register forms above, 80 % of the
    // kernel's local-memory stores -- and through L2 write-back most of its DRAM writes -- came from exactly these calls.)
    PQP_NOINLINE static void local_solve_mem(double *v, const double *fcol, int Mst) {
        // (same association order as local_solve: even / odd partial sums over d >= 2, the d = 1 term last)
#pragma unroll
        for (int k = 1; k < IMAX; ++k) {
            double p0 = v[k], p1 = 0.0;
#pragma unroll
            for (int dd = BW; dd >= 2; --dd)
                if (k - dd >= 0) {
                    if (dd & 1) p1 -= PQP_F(k, dd) * v[k - dd];
                    else p0 -= PQP_F(k, dd) * v[k - dd];
                }
            v[k] = (p0 + p1) - PQP_F(k, 1) * v[k - 1];
        }
#pragma unroll
        for (int k = IMAX - 1; k >= 0; --k) {
            double p0 = v[k] * PQP_F(k, 0), p1 = 0.0;
#pragma unroll
            for (int dd = BW; dd >= 2; --dd)
                if (k + dd < IMAX) {
                    if (dd & 1) p1 -= PQP_F(k + dd, dd) * v[k + dd];
                    else p0 -= PQP_F(k + dd, dd) * v[k + dd];
                }
            double r = p0 + p1;
            if (k + 1 < IMAX) r -= PQP_F(k + 1, 1) * v[k + 1];
            v[k] = r;
        }
    }
    PQP_NOINLINE static void local_solve_unit_mem(int j, double *v, const double *fcol, int Mst) {
#pragma unroll
        for (int k = 0; k < IMAX; ++k) v[k] = (k == j) ? 1.0 : 0.0;
        local_solve_mem(v, fcol, Mst);
    }

    // ---- row weights for the current rho, from the workspace E ----------------------------------

#undef PQP_F
};

}  // namespace pqp
