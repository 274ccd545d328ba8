// pqp_kp_core2.cuh -- latency-optimised warp-per-path "KP" solver (the production kernel).
//
// Same mathematics as pqp_kp_core.cuh (read its header first: unscaled-weighted form of OSQP's ADMM,
// station-ordered reduced KKT, separators every L stations, lane p owns the interior after
// separator p).  What changes is how the linear algebra is laid out on the SM:
//   * no global band matrix: every lane keeps the LDL' factor of ITS interior in a compact,
//     lane-private, bank-conflict-free column of shared memory (fac[(k*(BW+1)+d)*M + lane]); the
//     couplings between an interior and its two separators are closed-form functions of the
//     per-station weights and are applied as stencils;
//   * interior solves are fully unrolled over compile-time (IMAX, BW): the vector lives in registers,
//     factor entries stream from shared memory with no dependent address arithmetic;
//   * per-row weights W = rho_row E^2 are precomputed at every (re)factorisation; Ruiz E and D
//     themselves are only needed at residual checks and live in a global-memory workspace (L2);
//   * vectors use a padded chunk layout (chunk p = [sep a,b,c | interior | pad], odd stride CS).
//
// Reference being replaced: src/solver/solver_kp_as_input.cpp:26-203 + the OSQP solve at
// src/solver/solver.cpp:66-74.
#pragma once
#include "pqp_kp_core.cuh"

#if defined(__CUDACC__) && !defined(PQP_HOST_EMU)
#define PQP_NOINLINE __device__ __noinline__
#else
#define PQP_NOINLINE
#endif

namespace pqp {

struct Kp2Dims {
    int N, keep, ch, h, L, M, I, CS, nv, bw;
};

PQP_HD Kp2Dims kp2_dims(int N, int keep) {
    const KpDims a = kp_dims(N, keep);
    Kp2Dims d;
    d.N = N; d.keep = keep; d.ch = a.ch; d.h = a.h; d.L = a.L; d.M = a.M; d.bw = a.bw;
    d.I = 3 * (a.L - 1) + a.L / keep;   // unknowns of a full interior
    int cs = d.I + 3;
    if ((cs & 1) == 0) ++cs;            // odd chunk stride: 64-bit accesses of consecutive lanes spread over banks
    d.CS = cs;
    d.nv = d.M * cs;
    return d;
}

constexpr int kRed2 = 30;  // per separator: Sinv[9] | Off[9] | G[9] | g[3]

// doubles of workspace (global memory) a path needs: E (9N + ch + 2), Dr (nv), Dsl (N)
// (nv <= 32 chunks of at most 52 slots); laid out at workspace + 16*offset + kWsPerPath*path.
// (workspace layout helpers: kWsPerPath, kp2_ws_doubles, kp_ws_base, kp_ws_wold -- pqp_device.cuh)

// TAG only separates the out-of-line helpers of different users: each kernel gets its own copies, so the register
// needs of one kernel do not leak into another through a shared callee.
template <int IMAX, int BW, int TAG = 0>
struct Kp2 {
    // ---- shared-memory layout -----------------------------------------------------------------
    // Per-station fields are tiled: station i = 32*r + l lives at st[(r*kNF + f)*32 + l].  For a lane
    // walking its stations (l = lane) every field is then a compile-time offset from ONE base
    // register, and consecutive lanes hit consecutive banks.
    struct Fld {
        double *p;
        PQP_DEV double &operator[](int i) const { return p[(i >> 5) * (kNF * 32) + (i & 31)]; }
    };
    struct FldI {
        double *p;
        PQP_DEV int &operator[](int i) const { return *(int *)&p[(i >> 5) * (kNF * 32) + (i & 31)]; }
    };
    static constexpr int kNF = 39;
    struct Smem {
        double *base;
        int N, ch, nv, M, R;   // R = number of 32-station tiles
        // region A (persistent)
        PQP_DEV double *xr() const { return base; }
        PQP_DEV double *tr() const { return base + nv; }
        PQP_DEV double *sgr() const { return base + 2 * nv; }
        PQP_DEV double *st() const { return base + 3 * nv; }
        PQP_DEV Fld fld(int f) const { return Fld{st() + f * 32}; }
        PQP_DEV Fld xs() const { return fld(0); }
        PQP_DEV Fld ts() const { return fld(1); }
        PQP_DEV Fld sgs() const { return fld(2); }
        PQP_DEV Fld ksinv() const { return fld(3); }
        PQP_DEV Fld ds() const { return fld(4); }
        PQP_DEV Fld q10() const { return fld(5); }
        PQP_DEV Fld kds() const { return fld(6); }
        PQP_DEV Fld lH1() const { return fld(7); }
        PQP_DEV Fld uH1() const { return fld(8); }
        PQP_DEV Fld lH3() const { return fld(9); }
        PQP_DEV Fld uH3() const { return fld(10); }
        PQP_DEV Fld uS4m() const { return fld(11); }
        PQP_DEV Fld lS4p() const { return fld(12); }
        PQP_DEV Fld uS2m() const { return fld(13); }
        PQP_DEV Fld lS2p() const { return fld(14); }
        PQP_DEV Fld WD(int k) const { return fld(15 + k); }   // k = 0..2
        PQP_DEV Fld WKB() const { return fld(18); }
        PQP_DEV Fld WSB() const { return fld(19); }
        PQP_DEV Fld WH1() const { return fld(20); }
        PQP_DEV Fld WH3() const { return fld(21); }
        PQP_DEV Fld WS4() const { return fld(22); }
        PQP_DEV Fld WS2() const { return fld(23); }
        PQP_DEV Fld vD(int k) const { return fld(24 + k); }   // k = 0..2
        PQP_DEV Fld vKB() const { return fld(27); }
        PQP_DEV Fld vSB() const { return fld(28); }
        PQP_DEV Fld vH1() const { return fld(29); }
        PQP_DEV Fld vH3() const { return fld(30); }
        PQP_DEV Fld vS4m() const { return fld(31); }
        PQP_DEV Fld vS4p() const { return fld(32); }
        PQP_DEV Fld vS2m() const { return fld(33); }
        PQP_DEV Fld vS2p() const { return fld(34); }
        PQP_DEV Fld gD(int k) const { return fld(35 + k); }   // k = 0..2
        PQP_DEV FldI gxi() const { return FldI{st() + 38 * 32}; }
        PQP_DEV double *tailA() const { return st() + R * kNF * 32; }
        PQP_DEV double *WUB() const { return tailA(); }
        PQP_DEV double *vUB() const { return tailA() + ch; }
        PQP_DEV double *WEnd() const { return tailA() + 2 * ch; }
        PQP_DEV double *vEnd() const { return tailA() + 2 * ch + 2; }
        PQP_DEV int *gui() const { return (int *)(tailA() + 2 * ch + 4); }
        PQP_DEV double *regB() const { return tailA() + 2 * ch + 4 + (ch + 1) / 2; }
        // region B: factor, separator system, lane scratch (aliased by the scaling scratch)
        PQP_DEV double *fac() const { return regB(); }
        PQP_DEV double *red() const { return fac() + IMAX * (BW + 1) * M; }
        PQP_DEV double *lsc() const { return red() + kRed2 * M; }   // IMAX*M lane scratch
        // scaling scratch: region B plus the (not yet live) v / gD fields is too scattered, so the
        // scratch simply extends past region B; smem_doubles() accounts for it.
        PQP_DEV double *sDr() const { return regB(); }
        PQP_DEV double *sDsl() const { return regB() + nv; }
        PQP_DEV double *sE() const { return regB() + nv + N; }                    // 9N + ch + 2
        PQP_DEV double *sfDr() const { return sE() + 9 * N + ch + 2; }
        PQP_DEV double *sfDs() const { return sfDr() + nv; }
        PQP_DEV double *sfE() const { return sfDs() + N; }
    };
    PQP_HD static size_t smem_doubles(const Kp2Dims &d) {
        const size_t N = (size_t)d.N, ch = (size_t)d.ch, nv = (size_t)d.nv, M = (size_t)d.M;
        const size_t R = (N + 31) / 32;
        const size_t A = 3 * nv + R * kNF * 32 + 2 * ch + 4 + (ch + 1) / 2;
        size_t B = (size_t)IMAX * (BW + 1) * M + (size_t)kRed2 * M + (size_t)IMAX * M;
        const size_t S = 2 * nv + 20 * N + 2 * ch + 4;
        if (B < S) B = S;
        return A + B;
    }
    PQP_HD static bool fits(const Kp2Dims &d) { return d.I <= IMAX && d.bw <= BW; }
    // dims with the chunk stride of THIS instantiation: every lane touches IMAX interior slots, so a
    // chunk is [3 separator | IMAX interior (zero / identity padded)] rounded up to an odd stride.
    PQP_HD static Kp2Dims dims(int N, int keep) {
        Kp2Dims d = kp2_dims(N, keep);
        d.CS = (IMAX + 3) | 1;
        d.nv = d.M * d.CS;
        return d;
    }

    struct Ctx {
        Kp2Dims d;
        Smem s;
        const DevParams *pm;
        double x0[3];
        double lEH, uEH;
        double c, Dt, rho;
        int lo, cnt;        // this lane's interior: padded start index and size (lane < M)
        double *ws;         // global workspace: E (9N+ch+2) | Dr (nv) | Dsl (N)
    };
    PQP_DEV static double *wsE(const Ctx &cx) { return cx.ws; }
    PQP_DEV static double *wsDr(const Ctx &cx) { return cx.ws + 9 * cx.d.N + cx.d.ch + 2; }
    PQP_DEV static double *wsDsl(const Ctx &cx) { return wsDr(cx) + cx.d.nv; }

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

    // x = K_I^-1 y.  The vector is read from / written back to shared memory (element k at
    // vec[k * vst]) and lives in registers in between.  Out of line on purpose (register allocation).
    PQP_NOINLINE static void local_solve(double *vec, int vst, const double *fcol, int Mst) {
        double y[IMAX];
#pragma unroll
        for (int k = 0; k < IMAX; ++k) y[k] = vec[k * vst];
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
        for (int k = 0; k < IMAX; ++k) vec[k * vst] = y[k];
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

    // ---- row weights for the current rho, from the workspace E ----------------------------------
    PQP_DEV static void weights(const Cta &c, Ctx &cx) {
        const Smem &s = cx.s;
        const Kp2Dims &d = cx.d;
        const DevParams &pm = *cx.pm;
        const int N = d.N, ch = d.ch, tid = c.tid(), nt = c.nthreads();
        const double rho = cx.rho;
        const double *E = wsE(cx);
        for (int i = tid; i < N; i += nt) {
            s.WD(0)[i] = kp_w_eq(E[i], rho);
            s.WD(1)[i] = kp_w_eq(E[N + i], rho);
            s.WD(2)[i] = kp_w_eq(E[2 * N + i], rho);
            s.WKB()[i] = kp_w_box(E[3 * N + i], -pm.kmax, pm.kmax, rho);
            s.WSB()[i] = kp_w_box(E[4 * N + i], 0.0, pm.margin, rho);
            s.WH1()[i] = kp_w_box(E[5 * N + i], s.lH1()[i], s.uH1()[i], rho);
            s.WH3()[i] = kp_w_box(E[6 * N + i], s.lH3()[i], s.uH3()[i], rho);
            s.WS4()[i] = kp_w_box(E[7 * N + i], -kOsqpInfty, s.uS4m()[i], rho);
            s.WS2()[i] = kp_w_box(E[8 * N + i], -kOsqpInfty, s.uS2m()[i], rho);
        }
        for (int j = tid; j < ch; j += nt) s.WUB()[j] = kp_w_box(E[9 * N + j], -kOsqpInfty, kOsqpInfty, rho);
        if (tid == 0) {
            s.WEnd()[0] = kp_w_box(E[9 * N + ch], -1.0, 1.0, rho);
            s.WEnd()[1] = kp_w_box(E[9 * N + ch + 1], cx.lEH, cx.uEH, rho);
        }
        c.sync();
    }

    // ---- Ruiz equilibration (same arithmetic as kp_scale in pqp_kp_core.cuh, new layout) ---------
    PQP_DEV static void scale(const Cta &c, Ctx &cx) {
        const Smem &s = cx.s;
        const Kp2Dims &d = cx.d;
        const int N = d.N, ch = d.ch, keep = d.keep, tid = c.tid(), nt = c.nthreads();
        const DevParams &pm = *cx.pm;
        double *Dr = s.sDr(), *Dsl = s.sDsl(), *E = s.sE(), *fDr = s.sfDr(), *fDs = s.sfDs(), *fE = s.sfE();
        const auto gxi = s.gxi();
        const int *gui = s.gui();
        for (int g = tid; g < d.nv; g += nt) Dr[g] = 1.0;
        for (int i = tid; i < N; i += nt) Dsl[i] = 1.0;
        for (int k = tid; k < 9 * N + ch + 2; k += nt) E[k] = 1.0;
        cx.c = 1.0;
        cx.Dt = 1.0;
        c.sync();
        const double ad1 = fabs(pm.d1), ad2 = fabs(pm.d2), ad3 = fabs(pm.d3), ad4 = fabs(pm.d4);
        for (int sweep = 0; sweep < pm.scaling; ++sweep) {
            const double cst = cx.c;
            for (int i = tid; i < N; i += nt) {
                const int ga = gxi[i];
                const double Da = Dr[ga], Db = Dr[ga + 1], Dc = Dr[ga + 2], Dsv = Dsl[i];
                const double e0 = E[i], e1 = E[N + i], e2 = E[2 * N + i];
                const double eKB = E[3 * N + i], eSB = E[4 * N + i], eH1 = E[5 * N + i], eH3 = E[6 * N + i];
                const double eS4 = E[7 * N + i], eS2 = E[8 * N + i];
                const bool last = (i == N - 1);
                double Aa = fmax(fmax(e0, eH1), fmax(eH3, fmax(eS4, eS2)));
                double Ab = fmax(fmax(e1, eH1 * ad1), fmax(eH3 * ad3, fmax(eS4 * ad4, eS2 * ad2)));
                double Ac = fmax(e2, eKB);
                if (!last) {
                    const double e0n = E[i + 1], e1n = E[N + i + 1], e2n = E[2 * N + i + 1];
                    const double dsi = s.ds()[i], aq = fabs(s.q10()[i]);
                    Aa = fmax(Aa, fmax(e0n, e1n * aq));
                    Ab = fmax(Ab, fmax(e0n * dsi, e1n));
                    Ac = fmax(Ac, fmax(e1n * dsi, e2n));
                } else {
                    Aa = fmax(Aa, E[9 * N + ch]);
                    Ab = fmax(Ab, E[9 * N + ch + 1]);
                }
                const double As = fmax(eSB, fmax(eS4, eS2));
                fDr[ga] = 1.0 / sqrt(limit_scaling(fmax(cst * pm.w_pq * Da * Da, Aa * Da)));
                fDr[ga + 1] = 1.0 / sqrt(limit_scaling(Ab * Db));
                fDr[ga + 2] = 1.0 / sqrt(limit_scaling(fmax(cst * pm.w_c * Dc * Dc, Ac * Dc)));
                fDs[i] = 1.0 / sqrt(limit_scaling(fmax(cst * pm.w_s * Dsv * Dsv, As * Dsv)));
                double r0, r1, r2;
                if (i == 0) {
                    r0 = e0 * Da; r1 = e1 * Db; r2 = e2 * Dc;
                } else {
                    const int t = i - 1, gt = gxi[t];
                    const double Dat = Dr[gt], Dbt = Dr[gt + 1], Dct = Dr[gt + 2], Dut = Dr[gui[t / keep]];
                    const double dst = s.ds()[t], aqt = fabs(s.q10()[t]);
                    r0 = e0 * fmax(Da, fmax(Dat, dst * Dbt));
                    r1 = e1 * fmax(fmax(Db, aqt * Dat), fmax(Dbt, dst * Dct));
                    r2 = e2 * fmax(Dc, fmax(Dct, dst * Dut));
                }
                fE[i] = 1.0 / sqrt(limit_scaling(r0));
                fE[N + i] = 1.0 / sqrt(limit_scaling(r1));
                fE[2 * N + i] = 1.0 / sqrt(limit_scaling(r2));
                fE[3 * N + i] = 1.0 / sqrt(limit_scaling(eKB * Dc));
                fE[4 * N + i] = 1.0 / sqrt(limit_scaling(eSB * Dsv));
                fE[5 * N + i] = 1.0 / sqrt(limit_scaling(eH1 * fmax(Da, ad1 * Db)));
                fE[6 * N + i] = 1.0 / sqrt(limit_scaling(eH3 * fmax(Da, ad3 * Db)));
                fE[7 * N + i] = 1.0 / sqrt(limit_scaling(eS4 * fmax(Da, fmax(ad4 * Db, Dsv))));
                fE[8 * N + i] = 1.0 / sqrt(limit_scaling(eS2 * fmax(Da, fmax(ad2 * Db, Dsv))));
                if (last) {
                    fE[9 * N + ch] = 1.0 / sqrt(limit_scaling(E[9 * N + ch] * Da));
                    fE[9 * N + ch + 1] = 1.0 / sqrt(limit_scaling(E[9 * N + ch + 1] * Db));
                }
            }
            for (int j = tid; j < ch; j += nt) {
                const int gu = gui[j];
                const double Du = Dr[gu];
                double Au = E[9 * N + j];
                int t1 = j * keep + keep - 1;
                if (t1 > N - 2) t1 = N - 2;
                for (int t = j * keep; t <= t1; ++t) Au = fmax(Au, E[2 * N + t + 1] * s.ds()[t]);
                fDr[gu] = 1.0 / sqrt(limit_scaling(fmax(cst * (keep * pm.w_cr) * Du * Du, Au * Du)));
                fE[9 * N + j] = 1.0 / sqrt(limit_scaling(E[9 * N + j] * Du));
            }
            const double fDt = 1.0 / sqrt(limit_scaling(cst * pm.w_s * cx.Dt * cx.Dt));
            c.sync();
            for (int i = tid; i < N; i += nt) {
                const int ga = gxi[i];
                Dr[ga] *= fDr[ga]; Dr[ga + 1] *= fDr[ga + 1]; Dr[ga + 2] *= fDr[ga + 2];
                Dsl[i] *= fDs[i];
            }
            for (int j = tid; j < ch; j += nt) Dr[gui[j]] *= fDr[gui[j]];
            for (int k = tid; k < 9 * N + ch + 2; k += nt) E[k] *= fE[k];
            cx.Dt *= fDt;
            c.sync();
            double part = 0.0;
            for (int i = tid; i < N; i += nt) {
                const int ga = gxi[i];
                const double Da = Dr[ga], Dc = Dr[ga + 2], Dsv = Dsl[i];
                part += cst * pm.w_pq * Da * Da + cst * pm.w_c * Dc * Dc + cst * pm.w_s * Dsv * Dsv +
                        cst * pm.w_s * cx.Dt * cx.Dt;
            }
            for (int j = tid; j < ch; j += nt) {
                const double Du = Dr[gui[j]];
                part += cst * (keep * pm.w_cr) * Du * Du;
            }
            const double mean = c.sum(part) / (double)(5 * N + ch);
            double ct = fmax(mean, 1.0);
            ct = limit_scaling(ct);
            cx.c = cst * (1.0 / ct);
            c.sync();
        }
        // publish: sigma_v = sigma / D_v^2 into shared memory; E, D into the global workspace
        for (int g = tid; g < d.nv; g += nt) {
            const double D = Dr[g];
            s.sgr()[g] = pm.sigma / (D * D);
            wsDr(cx)[g] = D;
        }
        for (int i = tid; i < N; i += nt) {
            const double D = Dsl[i];
            s.sgs()[i] = pm.sigma / (D * D);
            wsDsl(cx)[i] = D;
        }
        for (int k = tid; k < 9 * N + ch + 2; k += nt) wsE(cx)[k] = E[k];
        c.sync();
    }

    // ---- assembly + factorisation ---------------------------------------------------------------
    PQP_DEV static int factor(const Cta &c, Ctx &cx) {
        const Smem &s = cx.s;
        const DevParams &pm = *cx.pm;
        // K_ss^-1 of the (exactly decoupled) slack unknowns: all threads
        for (int i = c.tid(); i < cx.d.N; i += c.nthreads())
            s.ksinv()[i] = 1.0 / (cx.c * pm.w_s + s.sgs()[i] + s.WSB()[i] + 2.0 * s.WS4()[i] + 2.0 * s.WS2()[i]);
        int ok = 1;
        if (c.wid == 0) ok = factor_w0(c.w, cx);   // partitioned KKT factorisation: warp 0
        return !c.any(!ok);
    }
    PQP_DEV static int factor_w0(const Warp &w, Ctx &cx) {
        const Smem &s = cx.s;
        const Kp2Dims &d = cx.d;
        const DevParams &pm = *cx.pm;
        const int N = d.N, keep = d.keep, L = d.L, M = d.M, lane = w.lane();
        const double c = cx.c;
        const auto gxi = s.gxi();
        const int *gui = s.gui();
        const bool act = lane < M;
        const int Mst = M;
        double *fcol = s.fac() + lane;
        int ok = 1;
        const int e = lane * L;                  // separator station of this lane
        const int lo = cx.lo, cnt = cx.cnt;
        if (act) {
#pragma unroll 1
            for (int k = 0; k < IMAX; ++k) {
#pragma unroll
                for (int dd = 0; dd <= BW; ++dd) PQP_F(k, dd) = (dd == 0 && k >= cnt) ? 1.0 : 0.0;
            }
            int i1 = e + L - 1;
            if (i1 > N - 1) i1 = N - 1;
            const double d1 = pm.d1, d2 = pm.d2, d3 = pm.d3, d4 = pm.d4;
            for (int i = e + 1; i <= i1; ++i) {
                const int ka = gxi[i] - lo;
                const bool last = (i == N - 1);
                const double W0 = s.WD(0)[i], W1 = s.WD(1)[i], W2 = s.WD(2)[i];
                double N0 = 0, N1 = 0, N2 = 0, dsi = 0, q = 0;
                if (!last) {
                    N0 = s.WD(0)[i + 1]; N1 = s.WD(1)[i + 1]; N2 = s.WD(2)[i + 1];
                    dsi = s.ds()[i]; q = s.q10()[i];
                }
                const double wH1 = s.WH1()[i], wH3 = s.WH3()[i], w4 = s.WS4()[i], w2 = s.WS2()[i];
                const int ga = gxi[i];
                double da = c * pm.w_pq + s.sgr()[ga] + W0 + N0 + N1 * q * q + wH1 + wH3 + 2.0 * w4 + 2.0 * w2;
                double db = s.sgr()[ga + 1] + W1 + N0 * dsi * dsi + N1 + wH1 * d1 * d1 + wH3 * d3 * d3 +
                            2.0 * w4 * d4 * d4 + 2.0 * w2 * d2 * d2;
                const double dc = c * pm.w_c + s.sgr()[ga + 2] + W2 + N1 * dsi * dsi + N2 + s.WKB()[i];
                if (last) {
                    da += s.WEnd()[0];
                    db += s.WEnd()[1];
                }
                PQP_F(ka, 0) = da;
                PQP_F(ka + 1, 0) = db;
                PQP_F(ka + 2, 0) = dc;
                PQP_F(ka + 1, 1) = N0 * dsi + N1 * q + wH1 * d1 + wH3 * d3 + 2.0 * w4 * d4 + 2.0 * w2 * d2;
                PQP_F(ka + 2, 2) = N1 * q * dsi;
                PQP_F(ka + 2, 1) = N1 * dsi;
                if (i - 1 > e) {  // previous station is interior too
                    const int t = i - 1;
                    const int off = ga - gxi[t];
                    const double dst = s.ds()[t], qt = s.q10()[t];
                    PQP_F(ka, off) = -W0;
                    PQP_F(ka, off - 1) = -W0 * dst;
                    PQP_F(ka + 1, off + 1) = -W1 * qt;
                    PQP_F(ka + 1, off) = -W1;
                    PQP_F(ka + 1, off - 1) = -W1 * dst;
                    PQP_F(ka + 2, off) = -W2;
                }
            }
            // held controls homed in this chunk
            const int j0 = e / keep;
            for (int j = j0; j < d.ch; ++j) {
                const int gu = gui[j];
                if (gu < lo || gu >= lo + cnt) break;
                const int ku = gu - lo;
                double du = c * (keep * pm.w_cr) + s.sgr()[gu] + s.WUB()[j];
                int ii1 = j * keep + keep;
                if (ii1 > N - 1) ii1 = N - 1;
                for (int ii = j * keep; ii <= ii1; ++ii) {
                    double val = 0.0;
                    if (ii >= 1 && (ii - 1) / keep == j) {
                        const double wv = s.WD(2)[ii], dst = s.ds()[ii - 1];
                        val -= wv * dst;
                        du += wv * dst * dst;
                    }
                    if (ii <= N - 2 && ii / keep == j) val += s.WD(2)[ii + 1] * s.ds()[ii];
                    if (ii % L == 0) continue;           // c of a separator station: closed-form coupling
                    const int kc = gxi[ii] + 2 - lo;
                    if (kc < ku) PQP_F(ku, ku - kc) = val;
                    else PQP_F(kc, kc - ku) = val;
                }
                PQP_F(ku, 0) = du;
            }
            ok = local_factor(fcol, Mst);
        }
        w.sync();
        // ---- Schur complement of the interiors onto the separators
        double Ap[9], Cp[9], Of[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Ap[k] = Cp[k] = Of[k] = 0.0;
        const bool has_right = act && (lane + 1 < M);
        double *lcol = s.lsc() + lane;
        if (act) {
            // coupling data: left = transition e -> e+1, right = transition e2-1 -> e2
            const int e2 = e + L;
            const bool has_int = cnt > 0 && e < N - 1;
            double lN0 = 0, lN1 = 0, lN2 = 0, lds = 0, lq = 0;
            int ka1 = 0, kul = -1;
            if (has_int) {
                lN0 = s.WD(0)[e + 1]; lN1 = s.WD(1)[e + 1]; lN2 = s.WD(2)[e + 1];
                lds = s.ds()[e]; lq = s.q10()[e];
                ka1 = gxi[e + 1] - lo;
                kul = gui[e / keep] - lo;
            }
            double rW0 = 0, rW1 = 0, rW2 = 0, rds = 0, rq = 0;
            int kat = 0, kur = -1;
            if (has_right) {
                rW0 = s.WD(0)[e2]; rW1 = s.WD(1)[e2]; rW2 = s.WD(2)[e2];
                rds = s.ds()[e2 - 1]; rq = s.q10()[e2 - 1];
                kat = gxi[e2 - 1] - lo;
                kur = gui[(e2 - 1) / keep] - lo;
            }
            for (int col = 0; col < 6; ++col) {
                if (col < 3 && !has_int) continue;
                if (col >= 3 && !has_right) break;
#pragma unroll 1
                for (int k = 0; k < IMAX; ++k) lcol[k * Mst] = 0.0;
                // column of K[I, s]
                if (col == 0) { lcol[ka1 * Mst] += -lN0; lcol[(ka1 + 1) * Mst] += -lN1 * lq; }
                else if (col == 1) { lcol[ka1 * Mst] += -lN0 * lds; lcol[(ka1 + 1) * Mst] += -lN1; }
                else if (col == 2) { lcol[(ka1 + 1) * Mst] += -lN1 * lds; lcol[(ka1 + 2) * Mst] += -lN2; lcol[kul * Mst] += lN2 * lds; }
                else if (col == 3) { lcol[kat * Mst] += -rW0; lcol[(kat + 1) * Mst] += -rW0 * rds; }
                else if (col == 4) { lcol[kat * Mst] += -rW1 * rq; lcol[(kat + 1) * Mst] += -rW1; lcol[(kat + 2) * Mst] += -rW1 * rds; }
                else { lcol[(kat + 2) * Mst] += -rW2; lcol[kur * Mst] += -rW2 * rds; }
                local_solve(lcol, Mst, fcol, Mst);
                // rows of K[S_p, I] . w  and  K[S_q, I] . w
                double tl[3] = {0, 0, 0}, trr[3] = {0, 0, 0};
                if (has_int) {
                    const double wa = lcol[ka1 * Mst], wb = lcol[(ka1 + 1) * Mst], wc = lcol[(ka1 + 2) * Mst];
                    const double wu = lcol[kul * Mst];
                    tl[0] = -lN0 * wa - lN1 * lq * wb;
                    tl[1] = -lN0 * lds * wa - lN1 * wb;
                    tl[2] = -lN1 * lds * wb - lN2 * wc + lN2 * lds * wu;
                }
                if (has_right) {
                    const double wa = lcol[kat * Mst], wb = lcol[(kat + 1) * Mst], wc = lcol[(kat + 2) * Mst];
                    const double wu = lcol[kur * Mst];
                    trr[0] = -rW0 * wa - rW0 * rds * wb;
                    trr[1] = -rW1 * rq * wa - rW1 * wb - rW1 * rds * wc;
                    trr[2] = -rW2 * wc - rW2 * rds * wu;
                }
                if (col < 3) {
                    for (int r = 0; r < 3; ++r) { Ap[r * 3 + col] = tl[r]; Of[col * 3 + r] = -trr[r]; }
                } else {
                    for (int r = 0; r < 3; ++r) Cp[r * 3 + (col - 3)] = trr[r];
                }
            }
        }
        // Dg_p = K[S_p,S_p] - A_p - C_{p-1}; C_{p-1} comes from the neighbouring lane
        double Dg[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            double cprev = w.shfl(Cp[k], (lane + 31) & 31);
            if (lane == 0) cprev = 0.0;
            Dg[k] = -Ap[k] - cprev;
        }
        if (act) {
            const int i = e;
            const bool last = (i == N - 1);
            const double W0 = s.WD(0)[i], W1 = s.WD(1)[i], W2 = s.WD(2)[i];
            double N0 = 0, N1 = 0, N2 = 0, dsi = 0, q = 0;
            if (!last) {
                N0 = s.WD(0)[i + 1]; N1 = s.WD(1)[i + 1]; N2 = s.WD(2)[i + 1];
                dsi = s.ds()[i]; q = s.q10()[i];
            }
            const double wH1 = s.WH1()[i], wH3 = s.WH3()[i], w4 = s.WS4()[i], w2 = s.WS2()[i];
            const double d1 = pm.d1, d2 = pm.d2, d3 = pm.d3, d4 = pm.d4;
            const int ga = gxi[i];
            double da = c * pm.w_pq + s.sgr()[ga] + W0 + N0 + N1 * q * q + wH1 + wH3 + 2.0 * w4 + 2.0 * w2;
            double db = s.sgr()[ga + 1] + W1 + N0 * dsi * dsi + N1 + wH1 * d1 * d1 + wH3 * d3 * d3 +
                        2.0 * w4 * d4 * d4 + 2.0 * w2 * d2 * d2;
            const double dc = c * pm.w_c + s.sgr()[ga + 2] + W2 + N1 * dsi * dsi + N2 + s.WKB()[i];
            if (last) { da += s.WEnd()[0]; db += s.WEnd()[1]; }
            const double kba = N0 * dsi + N1 * q + wH1 * d1 + wH3 * d3 + 2.0 * w4 * d4 + 2.0 * w2 * d2;
            const double kca = N1 * q * dsi, kcb = N1 * dsi;
            Dg[0] += da; Dg[4] += db; Dg[8] += dc;
            Dg[1] += kba; Dg[3] += kba; Dg[2] += kca; Dg[6] += kca; Dg[5] += kcb; Dg[7] += kcb;
            double *R = s.red() + kRed2 * lane;
#pragma unroll
            for (int k = 0; k < 9; ++k) { R[k] = Dg[k]; R[9 + k] = Of[k]; }
        }
        w.sync();
        // ---- block LDL' of the separator system (sequential, lane 0).  Stored per separator:
        //      Sinv_p | H_p = Sinv_p Off_p | G_p = Off_{p-1}' Sinv_{p-1} | g_p
        if (lane == 0) {
            double Sch[9], Sinv[9];
            for (int k = 0; k < 9; ++k) Sch[k] = s.red()[k];
            for (int p = 0; p < M; ++p) {
                double *Rp = s.red() + kRed2 * p;
                if (!(Sch[0] > 0.0)) ok = 0;
                inv3_spd(Sch, Sinv);
                if (p + 1 < M) {
                    double *Rn = Rp + kRed2;
                    double Off[9];
                    for (int k = 0; k < 9; ++k) Off[k] = Rp[9 + k];
                    for (int r = 0; r < 3; ++r)
                        for (int cc = 0; cc < 3; ++cc) {
                            double a = 0.0, hh = 0.0;
                            for (int k = 0; k < 3; ++k) {
                                a += Off[k * 3 + r] * Sinv[k * 3 + cc];
                                hh += Sinv[r * 3 + k] * Off[k * 3 + cc];
                            }
                            Rn[18 + r * 3 + cc] = a;
                            Rp[9 + r * 3 + cc] = hh;
                        }
                    for (int r = 0; r < 3; ++r)
                        for (int cc = 0; cc < 3; ++cc) {
                            double a = Rn[r * 3 + cc];
                            for (int k = 0; k < 3; ++k) a -= Rn[18 + r * 3 + k] * Off[k * 3 + cc];
                            Sch[r * 3 + cc] = a;
                        }
                }
                for (int k = 0; k < 9; ++k) Rp[k] = Sinv[k];
            }
        }
        ok = !w.any(!ok);
        w.sync();
        return ok;
    }

    // ---- K xt = rhs.  rhs in tr (padded order) on entry, xt on exit.  (ts is already final.)
    PQP_DEV static void solve(const Cta &c, Ctx &cx) {
        if (c.wid == 0) solve_w0(c.w, cx);
        c.sync();
    }
    PQP_DEV static void solve_w0(const Warp &w, Ctx &cx) {
        const Smem &s = cx.s;
        const Kp2Dims &d = cx.d;
        const int N = d.N, keep = d.keep, L = d.L, M = d.M, lane = w.lane();
        const bool act = lane < M;
        const int Mst = M;
        double *fcol = s.fac() + lane;
        double *tr = s.tr();
        const auto gxi = s.gxi();
        const int *gui = s.gui();
        const int e = lane * L, lo = cx.lo, sp = lo - 3;
        const bool has_int = act && cx.cnt > 0 && e < N - 1;
        const bool has_right = act && (lane + 1 < M);
        if (act) local_solve(tr + lo, 1, fcol, Mst);
        w.sync();
        double lN0 = 0, lN1 = 0, lN2 = 0, lds = 0, lq = 0;
        int pa1 = 0, pul = 0;
        if (has_int) {
            lN0 = s.WD(0)[e + 1]; lN1 = s.WD(1)[e + 1]; lN2 = s.WD(2)[e + 1];
            lds = s.ds()[e]; lq = s.q10()[e];
            pa1 = gxi[e + 1]; pul = gui[e / keep];
        }
        if (act) {
            double ga_ = tr[sp], gb_ = tr[sp + 1], gc_ = tr[sp + 2];
            if (lane > 0) {
                const int t = e - 1;
                const double W0 = s.WD(0)[e], W1 = s.WD(1)[e], W2 = s.WD(2)[e];
                const double dst = s.ds()[t], qt = s.q10()[t];
                const int pt = gxi[t];
                const double ya = tr[pt], yb = tr[pt + 1], yc = tr[pt + 2], yu = tr[gui[t / keep]];
                ga_ += W0 * (ya + dst * yb);
                gb_ += W1 * (qt * ya + yb + dst * yc);
                gc_ += W2 * (yc + dst * yu);
            }
            if (has_int) {
                const double ya = tr[pa1], yb = tr[pa1 + 1], yc = tr[pa1 + 2], yu = tr[pul];
                ga_ += lN0 * ya + lN1 * lq * yb;
                gb_ += lN0 * lds * ya + lN1 * yb;
                gc_ += lN1 * lds * yb + lN2 * yc - lN2 * lds * yu;
            }
            double *R = s.red() + kRed2 * lane;
            R[27] = ga_; R[28] = gb_; R[29] = gc_;
        }
        w.sync();
        // separator system: forward sweep (lane 0), g^ = Sinv g' (all lanes), backward sweep (lane 0)
        if (lane == 0) {
            double gp0 = s.red()[27], gp1 = s.red()[28], gp2 = s.red()[29];
            const double *Rp = s.red() + kRed2;
#pragma unroll 2
            for (int p = 1; p < M; ++p, Rp += kRed2) {
                const double G0 = Rp[18], G1 = Rp[19], G2 = Rp[20], G3 = Rp[21], G4 = Rp[22], G5 = Rp[23],
                             G6 = Rp[24], G7 = Rp[25], G8 = Rp[26];
                const double g0 = Rp[27] - (G0 * gp0 + G1 * gp1 + G2 * gp2);
                const double g1 = Rp[28] - (G3 * gp0 + G4 * gp1 + G5 * gp2);
                const double g2 = Rp[29] - (G6 * gp0 + G7 * gp1 + G8 * gp2);
                double *Rw = s.red() + kRed2 * p;
                Rw[27] = g0; Rw[28] = g1; Rw[29] = g2;
                gp0 = g0; gp1 = g1; gp2 = g2;
            }
        }
        w.sync();
        if (act) {
            double *R = s.red() + kRed2 * lane;
            const double t0 = R[27], t1 = R[28], t2 = R[29];
            const double h0 = R[0] * t0 + R[1] * t1 + R[2] * t2;
            const double h1 = R[3] * t0 + R[4] * t1 + R[5] * t2;
            const double h2 = R[6] * t0 + R[7] * t1 + R[8] * t2;
            R[27] = h0; R[28] = h1; R[29] = h2;
        }
        w.sync();
        if (lane == 0) {
            const double *Rp = s.red() + kRed2 * (M - 1);
            double x0 = Rp[27], x1 = Rp[28], x2 = Rp[29];
            int q = (M - 1) * d.CS;
            tr[q] = x0; tr[q + 1] = x1; tr[q + 2] = x2;
#pragma unroll 2
            for (int p = M - 2; p >= 0; --p) {
                Rp -= kRed2;
                q -= d.CS;
                const double H0 = Rp[9], H1 = Rp[10], H2 = Rp[11], H3 = Rp[12], H4 = Rp[13], H5 = Rp[14],
                             H6 = Rp[15], H7 = Rp[16], H8 = Rp[17];
                const double y0 = Rp[27] - (H0 * x0 + H1 * x1 + H2 * x2);
                const double y1 = Rp[28] - (H3 * x0 + H4 * x1 + H5 * x2);
                const double y2 = Rp[29] - (H6 * x0 + H7 * x1 + H8 * x2);
                tr[q] = y0; tr[q + 1] = y1; tr[q + 2] = y2;
                x0 = y0; x1 = y1; x2 = y2;
            }
        }
        w.sync();
        if (act) {
            double *lcol = s.lsc() + lane;
#pragma unroll
            for (int k = 0; k < IMAX; ++k) lcol[k * Mst] = 0.0;
            if (has_int) {
                const double xa = tr[sp], xb = tr[sp + 1], xc = tr[sp + 2];
                const int ka1 = pa1 - lo;
                lcol[ka1 * Mst] += lN0 * (xa + lds * xb);
                lcol[(ka1 + 1) * Mst] += lN1 * (lq * xa + xb + lds * xc);
                lcol[(ka1 + 2) * Mst] += lN2 * xc;
                lcol[(pul - lo) * Mst] += -lN2 * lds * xc;
            }
            if (has_right) {
                const int e2 = e + L, t = e2 - 1;
                const double W0 = s.WD(0)[e2], W1 = s.WD(1)[e2], W2 = s.WD(2)[e2];
                const double dst = s.ds()[t], qt = s.q10()[t];
                const int sq = sp + d.CS;
                const double xa = tr[sq], xb = tr[sq + 1], xc = tr[sq + 2];
                const int kat = gxi[t] - lo, kur = gui[t / keep] - lo;
                lcol[kat * Mst] += W0 * xa + W1 * qt * xb;
                lcol[(kat + 1) * Mst] += W0 * dst * xa + W1 * xb;
                lcol[(kat + 2) * Mst] += W1 * dst * xb + W2 * xc;
                lcol[kur * Mst] += W2 * dst * xc;
            }
            local_solve(lcol, Mst, fcol, Mst);
#pragma unroll
            for (int k = 0; k < IMAX; ++k) tr[lo + k] += lcol[k * Mst];
        }
        w.sync();
    }
#undef PQP_F

    // The 11 row values (A x)_r of station i for a vector in padded order (vr) + slack (vs).
    PQP_DEV static KpRows apply_A(const Ctx &cx, int i, const double *vr, const Fld vs) {
        const Smem &s = cx.s;
        const DevParams &pm = *cx.pm;
        const auto gxi = s.gxi();
        const int ga = gxi[i];
        const double a = vr[ga], b = vr[ga + 1], cc = vr[ga + 2], sl = vs[i];
        KpRows r;
        r.D0 = -a; r.D1 = -b; r.D2 = -cc;
        if (i >= 1) {
            const int t = i - 1, gt = gxi[t];
            const double at = vr[gt], bt = vr[gt + 1], ct = vr[gt + 2], ut = vr[s.gui()[t / cx.d.keep]];
            const double dst = s.ds()[t];
            r.D0 += at + dst * bt;
            r.D1 += s.q10()[t] * at + bt + dst * ct;
            r.D2 += ct + dst * ut;
        }
        r.KB = cc;
        r.SB = sl;
        r.H1 = a + pm.d1 * b;
        r.H3 = a + pm.d3 * b;
        const double e4 = a + pm.d4 * b, e2 = a + pm.d2 * b;
        r.S4m = e4 - sl; r.S4p = e4 + sl;
        r.S2m = e2 - sl; r.S2p = e2 + sl;
        return r;
    }
    PQP_DEV static void apply_At(const Ctx &cx, int i, const KpRows &o, double gEY, double gEH, double &ra,
                                 double &rb, double &rc, double &rs) {
        const Smem &s = cx.s;
        const DevParams &pm = *cx.pm;
        const int N = cx.d.N;
        const double s4 = o.S4m + o.S4p, s2 = o.S2m + o.S2p;
        ra = -o.D0 + o.H1 + o.H3 + s4 + s2;
        rb = -o.D1 + pm.d1 * o.H1 + pm.d3 * o.H3 + pm.d4 * s4 + pm.d2 * s2;
        rc = -o.D2 + o.KB;
        rs = o.SB - o.S4m + o.S4p - o.S2m + o.S2p;
        if (i < N - 1) {
            const double n0 = s.gD(0)[i + 1], n1 = s.gD(1)[i + 1], n2 = s.gD(2)[i + 1];
            const double dsi = s.ds()[i];
            ra += n0 + s.q10()[i] * n1;
            rb += dsi * n0 + n1;
            rc += dsi * n1 + n2;
        } else {
            ra += gEY;
            rb += gEH;
        }
    }
    PQP_DEV static void dyn_bounds(const Ctx &cx, int i, double &b0, double &b1, double &b2) {
        if (i == 0) { b0 = -cx.x0[0]; b1 = -cx.x0[1]; b2 = -cx.x0[2]; }
        else { b0 = 0.0; b1 = cx.s.kds()[i - 1]; b2 = 0.0; }
    }

    // ---- the whole per-path solve ----------------------------------------------------------------
    PQP_DEV static void solve_path(const Cta &c, const DevParams &prm, const BatchView &bv, int prob, double *smem,
                                   size_t smem_cap) {
        const int lane = c.lane(), tid = c.tid(), nt = c.nthreads();
        const int N = bv.n_points[prob];
        const int off = bv.offsets[prob];
        const pqp_state *ref = bv.ref + off;
        const pqp_station_bounds *bnd = bv.bounds + off;
        pqp_state *out = bv.out_states + off;
        int keep = 1;
        {
            double interval = 0.0;  // solver.cpp:21-27, solver_kp_as_input.cpp:17
            for (int i = 1; i < N && i < 10; ++i) {
                const double dd = ref[i].s - ref[i - 1].s;
                interval = interval > dd ? interval : dd;
            }
            const double q = 1.2 / interval;
            keep = (q < 2147483647.0) ? (int)q : 2147483647;
            if (!(q == q)) keep = 0;
            if (keep < 1) keep = 1;
        }
        int bad = (N < 2) || (keep > 10);
        Ctx cx;
        cx.pm = &prm;
        cx.d = dims(N < 2 ? 2 : N, keep > 10 ? 10 : keep);
        const Kp2Dims &d = cx.d;
        if (!bad && (!fits(d) || smem_doubles(d) > smem_cap || !bv.workspace)) bad = 1;
        const double qnan = nan("");
        if (bad) {
            if (tid == 0) {
                bv.status[prob] = PQP_INVALID_PROBLEM;
                if (bv.iters) bv.iters[prob] = 0;
            }
            for (int i = tid; i < N; i += nt) {
                out[i].x = out[i].y = out[i].z = out[i].k = out[i].s = qnan;
                out[i].v = out[i].a = 0.0;
                if (bv.out_frenet) {
                    double *f = bv.out_frenet + 3 * (size_t)(off + i);
                    f[0] = f[1] = f[2] = qnan;
                }
            }
            return;
        }
        Smem &s = cx.s;
        s.base = smem; s.N = N; s.ch = d.ch; s.nv = d.nv; s.M = d.M; s.R = (N + 31) / 32;
        cx.ws = kp_ws_base(bv.workspace, off, prob);
        const int ch = d.ch;
        const DevParams &pm = prm;
        cx.x0[0] = bv.x0[3 * (size_t)prob];
        cx.x0[1] = bv.x0[3 * (size_t)prob + 1];
        cx.x0[2] = bv.x0[3 * (size_t)prob + 2];
        cx.lEH = -kOsqpInfty;
        cx.uEH = kOsqpInfty;
        if (pm.constraint_end_heading) {  // solver_kp_as_input.cpp:193-201
            const double pi = 3.14159265358979323846;
            const double end_psi = constraint_angle(bv.end_heading[prob] - ref[N - 1].z);
            if (end_psi < 70 * pi / 180) {
                cx.lEH = end_psi - 5 * pi / 180;
                cx.uEH = end_psi + 5 * pi / 180;
            }
        }
        // ---- index tables (padded chunk order) and this lane's interior
        {
            const KpDims a = kp_dims(N, keep);
            for (int i = tid; i < N; i += nt) {
                const int p = i / d.L;
                s.gxi()[i] = p * d.CS + (kp_gx(a, i) - kp_gx(a, p * d.L));
            }
            for (int j = tid; j < ch; j += nt) {
                int home = j * keep + d.h;
                if (home > N - 1) home = N - 1;
                const int p = home / d.L;
                s.gui()[j] = p * d.CS + (kp_gu(a, j) - kp_gx(a, p * d.L));
            }
            cx.lo = 3;
            cx.cnt = 0;
            if (c.wid == 0 && lane < d.M) {
                const int g0 = kp_gx(a, lane * d.L);
                const int g1 = (lane + 1 < d.M) ? kp_gx(a, (lane + 1) * d.L) : a.nred;
                cx.lo = lane * d.CS + 3;
                cx.cnt = g1 - g0 - 3;
            }
        }
        int invalid = (cx.cnt > IMAX);
        // ---- per-station coefficients (setConstraintMatrix :84-98, :166-187)
        for (int i = tid; i < N; i += nt) {
            const double kap = ref[i].k;
            if (i < N - 1) {
                const double dsv = ref[i + 1].s - ref[i].s;
                s.ds()[i] = dsv;
                s.q10()[i] = -(kap * kap) * dsv;
                s.kds()[i] = dsv * kap;
            } else {
                s.ds()[i] = 0.0; s.q10()[i] = 0.0; s.kds()[i] = 0.0;
            }
            const pqp_station_bounds bb = bnd[i];
            s.lH1()[i] = bb.c0_lb; s.uH1()[i] = bb.c0_ub;
            s.lH3()[i] = bb.c2_lb; s.uH3()[i] = bb.c2_ub;
            s.uS4m()[i] = bb.c3_ub - pm.margin; s.lS4p()[i] = bb.c3_lb + pm.margin;
            s.uS2m()[i] = bb.c1_ub - pm.margin; s.lS2p()[i] = bb.c1_lb + pm.margin;
            if (!(bb.c0_lb <= bb.c0_ub) || !(bb.c2_lb <= bb.c2_ub)) invalid = 1;
            if (!(-kOsqpInfty <= s.uS4m()[i]) || !(s.lS4p()[i] <= kOsqpInfty) || !(-kOsqpInfty <= s.uS2m()[i]) ||
                !(s.lS2p()[i] <= kOsqpInfty))
                invalid = 1;
        }
        if (!(0.0 <= pm.margin) || !(-pm.kmax <= pm.kmax) || !(cx.lEH <= cx.uEH)) invalid = 1;
        invalid = c.any(invalid);
        c.sync();

        int status = PQP_UNSOLVED;
        int iter = 0;
        if (invalid) {
            status = PQP_INVALID_PROBLEM;
        } else {
            scale(c, cx);
            // cold start: OSQP's first iteration from zero leaves x = 0, v = 0 (see pqp_kp_core.cuh)
            for (int g = tid; g < d.nv; g += nt) { s.xr()[g] = 0.0; s.tr()[g] = 0.0; }
            for (int i = tid; i < N; i += nt) {
                s.xs()[i] = 0.0;
                s.vD(0)[i] = s.vD(1)[i] = s.vD(2)[i] = 0.0;
                s.vKB()[i] = s.vSB()[i] = s.vH1()[i] = s.vH3()[i] = 0.0;
                s.vS4m()[i] = s.vS4p()[i] = s.vS2m()[i] = s.vS2p()[i] = 0.0;
            }
            for (int j = tid; j < ch; j += nt) s.vUB()[j] = 0.0;
            if (tid == 0) s.vEnd()[0] = s.vEnd()[1] = 0.0;
            cx.rho = fmin(fmax(pm.rho, kRhoMin), kRhoMax);
            c.sync();
            weights(c, cx);
            if (!factor(c, cx)) status = PQP_NON_CVX;
            const double alpha = pm.alpha;
            double pri_res = 0, dua_res = 0, pri_nrm = 0, dua_nrm = 0;
            double inf_nrm = 0, inf_lhs = 0, inf_cert = 0;   // primal-infeasibility certificate of the last check
            double *wold = kp_ws_wold(cx.ws, N);
            const auto gxi = s.gxi();
        const int *gui = s.gui();
            iter = 1;
            while (status == PQP_UNSOLVED && iter < pm.max_iter) {
                ++iter;
                // ---- (a) rhs = sigma_v x + A' W (2 clamp(v) - v)
                for (int i = tid; i < N; i += nt) {
                    double b0, b1, b2;
                    dyn_bounds(cx, i, b0, b1, b2);
                    s.gD(0)[i] = s.WD(0)[i] * (2.0 * b0 - s.vD(0)[i]);
                    s.gD(1)[i] = s.WD(1)[i] * (2.0 * b1 - s.vD(1)[i]);
                    s.gD(2)[i] = s.WD(2)[i] * (2.0 * b2 - s.vD(2)[i]);
                }
                c.sync();
                for (int i = tid; i < N; i += nt) {
                    KpRows g;
                    g.D0 = s.gD(0)[i]; g.D1 = s.gD(1)[i]; g.D2 = s.gD(2)[i];
                    double v;
                    v = s.vKB()[i]; g.KB = s.WKB()[i] * (2.0 * clampd(v, -pm.kmax, pm.kmax) - v);
                    v = s.vSB()[i]; g.SB = s.WSB()[i] * (2.0 * clampd(v, 0.0, pm.margin) - v);
                    v = s.vH1()[i]; g.H1 = s.WH1()[i] * (2.0 * clampd(v, s.lH1()[i], s.uH1()[i]) - v);
                    v = s.vH3()[i]; g.H3 = s.WH3()[i] * (2.0 * clampd(v, s.lH3()[i], s.uH3()[i]) - v);
                    const double w4 = s.WS4()[i], w2 = s.WS2()[i];
                    v = s.vS4m()[i]; g.S4m = w4 * (2.0 * clampd(v, -kOsqpInfty, s.uS4m()[i]) - v);
                    v = s.vS4p()[i]; g.S4p = w4 * (2.0 * clampd(v, s.lS4p()[i], kOsqpInfty) - v);
                    v = s.vS2m()[i]; g.S2m = w2 * (2.0 * clampd(v, -kOsqpInfty, s.uS2m()[i]) - v);
                    v = s.vS2p()[i]; g.S2p = w2 * (2.0 * clampd(v, s.lS2p()[i], kOsqpInfty) - v);
                    double gEY = 0, gEH = 0;
                    if (i == N - 1) {
                        v = s.vEnd()[0]; gEY = s.WEnd()[0] * (2.0 * clampd(v, -1.0, 1.0) - v);
                        v = s.vEnd()[1]; gEH = s.WEnd()[1] * (2.0 * clampd(v, cx.lEH, cx.uEH) - v);
                    }
                    double ra, rb, rc, rs;
                    apply_At(cx, i, g, gEY, gEH, ra, rb, rc, rs);
                    const int ga = gxi[i];
                    s.tr()[ga] = s.sgr()[ga] * s.xr()[ga] + ra;
                    s.tr()[ga + 1] = s.sgr()[ga + 1] * s.xr()[ga + 1] + rb;
                    s.tr()[ga + 2] = s.sgr()[ga + 2] * s.xr()[ga + 2] + rc;
                    s.ts()[i] = (s.sgs()[i] * s.xs()[i] + rs) * s.ksinv()[i];   // slack decouples: final
                }
                for (int j = tid; j < ch; j += nt) {
                    const int gu = gui[j];
                    const double v = s.vUB()[j];
                    double acc = s.sgr()[gu] * s.xr()[gu] + s.WUB()[j] * (2.0 * clampd(v, -kOsqpInfty, kOsqpInfty) - v);
                    int t1 = j * keep + keep - 1;
                    if (t1 > N - 2) t1 = N - 2;
                    for (int t = j * keep; t <= t1; ++t) acc += s.ds()[t] * s.gD(2)[t + 1];
                    s.tr()[gu] = acc;
                }
                c.sync();
                // ---- (b) reduced KKT solve
                solve(c, cx);
                // iterations that end in a termination check first park w = v - clamp(v): the check needs
                // delta_y = W (w_new - w_old) (OSQP update_y / is_primal_infeasible)
                const bool chk = wold && ((pm.check_termination && (iter % pm.check_termination == 0)) || iter == pm.max_iter);
                if (chk) {
                    for (int i = tid; i < N; i += nt) {
                        double b0, b1, b2, v;
                        dyn_bounds(cx, i, b0, b1, b2);
                        double *wo = wold + i;
                        wo[0] = s.vD(0)[i] - b0; wo[N] = s.vD(1)[i] - b1; wo[2 * N] = s.vD(2)[i] - b2;
                        v = s.vKB()[i]; wo[3 * N] = v - clampd(v, -pm.kmax, pm.kmax);
                        v = s.vSB()[i]; wo[4 * N] = v - clampd(v, 0.0, pm.margin);
                        v = s.vH1()[i]; wo[5 * N] = v - clampd(v, s.lH1()[i], s.uH1()[i]);
                        v = s.vH3()[i]; wo[6 * N] = v - clampd(v, s.lH3()[i], s.uH3()[i]);
                        v = s.vS4m()[i]; wo[7 * N] = v - clampd(v, -kOsqpInfty, s.uS4m()[i]);
                        v = s.vS4p()[i]; wo[8 * N] = v - clampd(v, s.lS4p()[i], kOsqpInfty);
                        v = s.vS2m()[i]; wo[9 * N] = v - clampd(v, -kOsqpInfty, s.uS2m()[i]);
                        v = s.vS2p()[i]; wo[10 * N] = v - clampd(v, s.lS2p()[i], kOsqpInfty);
                        if (i == N - 1) {
                            v = s.vEnd()[0]; wold[11 * N] = v - clampd(v, -1.0, 1.0);
                            v = s.vEnd()[1]; wold[11 * N + 1] = v - clampd(v, cx.lEH, cx.uEH);
                        }
                    }
                }
                // ---- (c) v += alpha (A xt - clamp(v)),  x = alpha xt + (1 - alpha) x
                for (int i = tid; i < N; i += nt) {
                    const KpRows zt = apply_A(cx, i, s.tr(), s.ts());
                    double b0, b1, b2, v;
                    dyn_bounds(cx, i, b0, b1, b2);
                    s.vD(0)[i] += alpha * (zt.D0 - b0);
                    s.vD(1)[i] += alpha * (zt.D1 - b1);
                    s.vD(2)[i] += alpha * (zt.D2 - b2);
                    v = s.vKB()[i]; s.vKB()[i] = v + alpha * (zt.KB - clampd(v, -pm.kmax, pm.kmax));
                    v = s.vSB()[i]; s.vSB()[i] = v + alpha * (zt.SB - clampd(v, 0.0, pm.margin));
                    v = s.vH1()[i]; s.vH1()[i] = v + alpha * (zt.H1 - clampd(v, s.lH1()[i], s.uH1()[i]));
                    v = s.vH3()[i]; s.vH3()[i] = v + alpha * (zt.H3 - clampd(v, s.lH3()[i], s.uH3()[i]));
                    v = s.vS4m()[i]; s.vS4m()[i] = v + alpha * (zt.S4m - clampd(v, -kOsqpInfty, s.uS4m()[i]));
                    v = s.vS4p()[i]; s.vS4p()[i] = v + alpha * (zt.S4p - clampd(v, s.lS4p()[i], kOsqpInfty));
                    v = s.vS2m()[i]; s.vS2m()[i] = v + alpha * (zt.S2m - clampd(v, -kOsqpInfty, s.uS2m()[i]));
                    v = s.vS2p()[i]; s.vS2p()[i] = v + alpha * (zt.S2p - clampd(v, s.lS2p()[i], kOsqpInfty));
                    if (i == N - 1) {
                        const int ga = gxi[i];
                        v = s.vEnd()[0]; s.vEnd()[0] = v + alpha * (s.tr()[ga] - clampd(v, -1.0, 1.0));
                        v = s.vEnd()[1]; s.vEnd()[1] = v + alpha * (s.tr()[ga + 1] - clampd(v, cx.lEH, cx.uEH));
                    }
                }
                for (int j = tid; j < ch; j += nt) {
                    const double v = s.vUB()[j];
                    s.vUB()[j] = v + alpha * (s.tr()[gui[j]] - clampd(v, -kOsqpInfty, kOsqpInfty));
                }
                c.sync();
                for (int g = tid; g < d.nv; g += nt) s.xr()[g] = alpha * s.tr()[g] + (1.0 - alpha) * s.xr()[g];
                for (int i = tid; i < N; i += nt) s.xs()[i] = alpha * s.ts()[i] + (1.0 - alpha) * s.xs()[i];
                c.sync();
                // ---- (d) residuals, termination, adaptive rho
                const bool can_check = pm.check_termination && (iter % pm.check_termination == 0);
                const bool can_adapt = pm.adaptive_rho && pm.adaptive_rho_interval &&
                                       (iter % pm.adaptive_rho_interval == 0);
                if (can_check || can_adapt || iter == pm.max_iter) {
                    const double *E = wsE(cx), *Dr = wsDr(cx), *Dsl = wsDsl(cx);
                    double pr = 0, nz = 0, nax = 0, prs = 0, nzs = 0, naxs = 0;
                    const double cinv = 1.0 / cx.c;
#define PQP_ROW(AX, V, LO, HI, EE)                                                   \
    {                                                                                \
        const double ax_ = (AX), v_ = (V), z_ = clampd(v_, (LO), (HI)), r_ = ax_ - z_; \
        const double e_ = (EE);                                                      \
        pr = fmax(pr, fabs(r_)); nz = fmax(nz, fabs(z_)); nax = fmax(nax, fabs(ax_)); \
        prs = fmax(prs, e_ * fabs(r_)); nzs = fmax(nzs, e_ * fabs(z_));              \
        naxs = fmax(naxs, e_ * fabs(ax_));                                           \
    }
#define PQP_DUAL(V, LO, HI, WW) ((WW) * ((V) - clampd((V), (LO), (HI))) * cinv)
                    for (int i = tid; i < N; i += nt) {
                        const KpRows ax = apply_A(cx, i, s.xr(), s.xs());
                        double b0, b1, b2;
                        dyn_bounds(cx, i, b0, b1, b2);
                        PQP_ROW(ax.D0, s.vD(0)[i], b0, b0, E[i])
                        PQP_ROW(ax.D1, s.vD(1)[i], b1, b1, E[N + i])
                        PQP_ROW(ax.D2, s.vD(2)[i], b2, b2, E[2 * N + i])
                        PQP_ROW(ax.KB, s.vKB()[i], -pm.kmax, pm.kmax, E[3 * N + i])
                        PQP_ROW(ax.SB, s.vSB()[i], 0.0, pm.margin, E[4 * N + i])
                        PQP_ROW(ax.H1, s.vH1()[i], s.lH1()[i], s.uH1()[i], E[5 * N + i])
                        PQP_ROW(ax.H3, s.vH3()[i], s.lH3()[i], s.uH3()[i], E[6 * N + i])
                        PQP_ROW(ax.S4m, s.vS4m()[i], -kOsqpInfty, s.uS4m()[i], E[7 * N + i])
                        PQP_ROW(ax.S4p, s.vS4p()[i], s.lS4p()[i], kOsqpInfty, E[7 * N + i])
                        PQP_ROW(ax.S2m, s.vS2m()[i], -kOsqpInfty, s.uS2m()[i], E[8 * N + i])
                        PQP_ROW(ax.S2p, s.vS2p()[i], s.lS2p()[i], kOsqpInfty, E[8 * N + i])
                        s.gD(0)[i] = PQP_DUAL(s.vD(0)[i], b0, b0, s.WD(0)[i]);
                        s.gD(1)[i] = PQP_DUAL(s.vD(1)[i], b1, b1, s.WD(1)[i]);
                        s.gD(2)[i] = PQP_DUAL(s.vD(2)[i], b2, b2, s.WD(2)[i]);
                        if (i == N - 1) {
                            const int ga = gxi[i];
                            PQP_ROW(s.xr()[ga], s.vEnd()[0], -1.0, 1.0, E[9 * N + ch])
                            PQP_ROW(s.xr()[ga + 1], s.vEnd()[1], cx.lEH, cx.uEH, E[9 * N + ch + 1])
                        }
                    }
                    for (int j = tid; j < ch; j += nt)
                        PQP_ROW(s.xr()[gui[j]], s.vUB()[j], -kOsqpInfty, kOsqpInfty, E[9 * N + j])
                    c.sync();
                    double dr = 0, npx = 0, naty = 0, drs = 0, npxs = 0, natys = 0;
                    const double cc = cx.c;
#define PQP_VAR(PX, ATY, DD)                                                          \
    {                                                                                 \
        const double px_ = (PX), aty_ = (ATY), r_ = px_ + aty_, cd_ = cc * (DD);      \
        dr = fmax(dr, fabs(r_)); npx = fmax(npx, fabs(px_)); naty = fmax(naty, fabs(aty_)); \
        drs = fmax(drs, cd_ * fabs(r_)); npxs = fmax(npxs, cd_ * fabs(px_));          \
        natys = fmax(natys, cd_ * fabs(aty_));                                        \
    }
                    for (int i = tid; i < N; i += nt) {
                        const int ga = gxi[i];
                        KpRows y;
                        y.D0 = s.gD(0)[i]; y.D1 = s.gD(1)[i]; y.D2 = s.gD(2)[i];
                        y.KB = PQP_DUAL(s.vKB()[i], -pm.kmax, pm.kmax, s.WKB()[i]);
                        y.SB = PQP_DUAL(s.vSB()[i], 0.0, pm.margin, s.WSB()[i]);
                        y.H1 = PQP_DUAL(s.vH1()[i], s.lH1()[i], s.uH1()[i], s.WH1()[i]);
                        y.H3 = PQP_DUAL(s.vH3()[i], s.lH3()[i], s.uH3()[i], s.WH3()[i]);
                        y.S4m = PQP_DUAL(s.vS4m()[i], -kOsqpInfty, s.uS4m()[i], s.WS4()[i]);
                        y.S4p = PQP_DUAL(s.vS4p()[i], s.lS4p()[i], kOsqpInfty, s.WS4()[i]);
                        y.S2m = PQP_DUAL(s.vS2m()[i], -kOsqpInfty, s.uS2m()[i], s.WS2()[i]);
                        y.S2p = PQP_DUAL(s.vS2p()[i], s.lS2p()[i], kOsqpInfty, s.WS2()[i]);
                        double yEY = 0, yEH = 0;
                        if (i == N - 1) {
                            yEY = PQP_DUAL(s.vEnd()[0], -1.0, 1.0, s.WEnd()[0]);
                            yEH = PQP_DUAL(s.vEnd()[1], cx.lEH, cx.uEH, s.WEnd()[1]);
                        }
                        double ra, rb, rc, rs;
                        apply_At(cx, i, y, yEY, yEH, ra, rb, rc, rs);
                        PQP_VAR(pm.w_pq * s.xr()[ga], ra, Dr[ga])
                        PQP_VAR(0.0, rb, Dr[ga + 1])
                        PQP_VAR(pm.w_c * s.xr()[ga + 2], rc, Dr[ga + 2])
                        PQP_VAR(pm.w_s * s.xs()[i], rs, Dsl[i])
                    }
                    for (int j = tid; j < ch; j += nt) {
                        const int gu = gui[j];
                        double aty = PQP_DUAL(s.vUB()[j], -kOsqpInfty, kOsqpInfty, s.WUB()[j]);
                        int t1 = j * keep + keep - 1;
                        if (t1 > N - 2) t1 = N - 2;
                        for (int t = j * keep; t <= t1; ++t) aty += s.ds()[t] * s.gD(2)[t + 1];
                        PQP_VAR((keep * pm.w_cr) * s.xr()[gu], aty, Dr[gu])
                    }
#undef PQP_ROW
#undef PQP_DUAL
#undef PQP_VAR
                    // ---- primal-infeasibility certificate (OSQP is_primal_infeasible) in unscaled terms:
                    // g = W (w_new - w_old) = E delta_y projected on the cone of the finite bounds;
                    // ||g||_inf, u'g+ + l'g-, ||A'g||_inf.  The control rows are free (g = 0).
                    if (chk) {
                        double c_nrm = 0, c_lhs = 0, c_cert = 0;
#define PQP_G(V, LO, HI, WW, WO) ((WW) * (((V) - clampd((V), (LO), (HI))) - (WO)))
#define PQP_ACC(G, LO, HI) { const double g_ = (G); c_nrm = fmax(c_nrm, fabs(g_)); c_lhs += (HI) * fmax(g_, 0.0) + (LO) * fmin(g_, 0.0); }
                        c.sync();   // the dual-residual pass has consumed gD
                        for (int i = tid; i < N; i += nt) {
                            double b0, b1, b2;
                            dyn_bounds(cx, i, b0, b1, b2);
                            const double *wo = wold + i;
                            s.gD(0)[i] = s.WD(0)[i] * ((s.vD(0)[i] - b0) - wo[0]);
                            s.gD(1)[i] = s.WD(1)[i] * ((s.vD(1)[i] - b1) - wo[N]);
                            s.gD(2)[i] = s.WD(2)[i] * ((s.vD(2)[i] - b2) - wo[2 * N]);
                        }
                        c.sync();
                        for (int i = tid; i < N; i += nt) {
                            double b0, b1, b2;
                            dyn_bounds(cx, i, b0, b1, b2);
                            const double *wo = wold + i;
                            KpRows g;
                            g.D0 = s.gD(0)[i]; g.D1 = s.gD(1)[i]; g.D2 = s.gD(2)[i];
                            g.KB = PQP_G(s.vKB()[i], -pm.kmax, pm.kmax, s.WKB()[i], wo[3 * N]);
                            g.SB = PQP_G(s.vSB()[i], 0.0, pm.margin, s.WSB()[i], wo[4 * N]);
                            g.H1 = PQP_G(s.vH1()[i], s.lH1()[i], s.uH1()[i], s.WH1()[i], wo[5 * N]);
                            g.H3 = PQP_G(s.vH3()[i], s.lH3()[i], s.uH3()[i], s.WH3()[i], wo[6 * N]);
                            // one-sided rows: l = -inf keeps the positive part, u = +inf the negative part
                            g.S4m = fmax(PQP_G(s.vS4m()[i], -kOsqpInfty, s.uS4m()[i], s.WS4()[i], wo[7 * N]), 0.0);
                            g.S4p = fmin(PQP_G(s.vS4p()[i], s.lS4p()[i], kOsqpInfty, s.WS4()[i], wo[8 * N]), 0.0);
                            g.S2m = fmax(PQP_G(s.vS2m()[i], -kOsqpInfty, s.uS2m()[i], s.WS2()[i], wo[9 * N]), 0.0);
                            g.S2p = fmin(PQP_G(s.vS2p()[i], s.lS2p()[i], kOsqpInfty, s.WS2()[i], wo[10 * N]), 0.0);
                            PQP_ACC(g.D0, b0, b0) PQP_ACC(g.D1, b1, b1) PQP_ACC(g.D2, b2, b2)
                            PQP_ACC(g.KB, -pm.kmax, pm.kmax) PQP_ACC(g.SB, 0.0, pm.margin)
                            PQP_ACC(g.H1, s.lH1()[i], s.uH1()[i]) PQP_ACC(g.H3, s.lH3()[i], s.uH3()[i])
                            PQP_ACC(g.S4m, 0.0, s.uS4m()[i]) PQP_ACC(g.S4p, s.lS4p()[i], 0.0)
                            PQP_ACC(g.S2m, 0.0, s.uS2m()[i]) PQP_ACC(g.S2p, s.lS2p()[i], 0.0)
                            double gEY = 0, gEH = 0;
                            if (i == N - 1) {
                                gEY = PQP_G(s.vEnd()[0], -1.0, 1.0, s.WEnd()[0], wold[11 * N]);
                                gEH = PQP_G(s.vEnd()[1], cx.lEH, cx.uEH, s.WEnd()[1], wold[11 * N + 1]);
                                if (cx.uEH >= kOsqpInfty) gEH = (cx.lEH <= -kOsqpInfty) ? 0.0 : fmin(gEH, 0.0);
                                else if (cx.lEH <= -kOsqpInfty) gEH = fmax(gEH, 0.0);
                                PQP_ACC(gEY, -1.0, 1.0)
                                PQP_ACC(gEH, (cx.lEH <= -kOsqpInfty ? 0.0 : cx.lEH), (cx.uEH >= kOsqpInfty ? 0.0 : cx.uEH))
                            }
                            double ra, rb, rc, rs;
                            apply_At(cx, i, g, gEY, gEH, ra, rb, rc, rs);
                            c_cert = fmax(c_cert, fmax(fmax(fabs(ra), fabs(rb)), fmax(fabs(rc), fabs(rs))));
                        }
                        for (int j = tid; j < ch; j += nt) {
                            double aty = 0.0;
                            int t1 = j * keep + keep - 1;
                            if (t1 > N - 2) t1 = N - 2;
                            for (int t = j * keep; t <= t1; ++t) aty += s.ds()[t] * s.gD(2)[t + 1];
                            c_cert = fmax(c_cert, fabs(aty));
                        }
#undef PQP_G
#undef PQP_ACC
                        inf_nrm = c.max(c_nrm); inf_cert = c.max(c_cert); inf_lhs = c.sum(c_lhs);
                    }
                    pr = c.max(pr); nz = c.max(nz); nax = c.max(nax);
                    prs = c.max(prs); nzs = c.max(nzs); naxs = c.max(naxs);
                    dr = c.max(dr); npx = c.max(npx); naty = c.max(naty);
                    drs = c.max(drs); npxs = c.max(npxs); natys = c.max(natys);
                    c.sync();
                    pri_res = pr; dua_res = dr;
                    pri_nrm = fmax(nz, nax); dua_nrm = fmax(npx, naty);
                    if (can_check || iter == pm.max_iter) {
                        // OSQP check_termination; q = 0, so the dual-infeasibility test never fires
                        const bool prim_ok = pri_res < pm.eps_abs + pm.eps_rel * pri_nrm;
                        if (pri_res > kOsqpInfty || dua_res > kOsqpInfty) status = PQP_NON_CVX;
                        else if (prim_ok && dua_res < pm.eps_abs + pm.eps_rel * dua_nrm) status = PQP_SOLVED;
                        else if (!prim_ok && primal_infeasible(inf_nrm, inf_lhs, inf_cert, pm.eps_prim_inf))
                            status = PQP_PRIMAL_INFEASIBLE;
                    }
                    if (status == PQP_UNSOLVED && can_adapt) {
                        const double rho = cx.rho;
                        const double pn = prs / (fmax(nzs, naxs) + 1e-10);
                        const double dn = drs / (fmax(npxs, natys) + 1e-10);
                        double rho_new = rho * sqrt(pn / (dn + 1e-10));
                        rho_new = fmin(fmax(rho_new, kRhoMin), kRhoMax);
                        if (rho_new > rho * pm.adaptive_rho_tolerance || rho_new < rho / pm.adaptive_rho_tolerance) {
                            const double ratio = rho / rho_new;
                            for (int i = tid; i < N; i += nt) {
                                double b0, b1, b2, v, z;
                                dyn_bounds(cx, i, b0, b1, b2);
                                s.vD(0)[i] = b0 + (s.vD(0)[i] - b0) * ratio;
                                s.vD(1)[i] = b1 + (s.vD(1)[i] - b1) * ratio;
                                s.vD(2)[i] = b2 + (s.vD(2)[i] - b2) * ratio;
#define PQP_RESC(V, LO, HI) v = (V); z = clampd(v, (LO), (HI)); (V) = z + (v - z) * ratio;
                                PQP_RESC(s.vKB()[i], -pm.kmax, pm.kmax)
                                PQP_RESC(s.vSB()[i], 0.0, pm.margin)
                                PQP_RESC(s.vH1()[i], s.lH1()[i], s.uH1()[i])
                                PQP_RESC(s.vH3()[i], s.lH3()[i], s.uH3()[i])
                                PQP_RESC(s.vS4m()[i], -kOsqpInfty, s.uS4m()[i])
                                PQP_RESC(s.vS4p()[i], s.lS4p()[i], kOsqpInfty)
                                PQP_RESC(s.vS2m()[i], -kOsqpInfty, s.uS2m()[i])
                                PQP_RESC(s.vS2p()[i], s.lS2p()[i], kOsqpInfty)
                                if (i == N - 1) {
                                    PQP_RESC(s.vEnd()[0], -1.0, 1.0)
                                    PQP_RESC(s.vEnd()[1], cx.lEH, cx.uEH)
                                }
#undef PQP_RESC
                            }
                            cx.rho = rho_new;
                            c.sync();
                            weights(c, cx);
                            if (!factor(c, cx)) status = PQP_NON_CVX;
                        }
                    }
                }
            }
            if (status == PQP_UNSOLVED) {
                const bool prim_ok = pri_res < 10 * pm.eps_abs + 10 * pm.eps_rel * pri_nrm;
                if (prim_ok && dua_res < 10 * pm.eps_abs + 10 * pm.eps_rel * dua_nrm) status = PQP_SOLVED_INACCURATE;
                else if (!prim_ok && primal_infeasible(inf_nrm, inf_lhs, inf_cert, 10 * pm.eps_prim_inf))
                    status = PQP_PRIMAL_INFEASIBLE;
                else status = PQP_MAX_ITER_REACHED;
            }
        }
        // ---- epilogue: getOptimizedPath, solver_kp_as_input.cpp:26-43
        const bool has_sol = (status == PQP_SOLVED || status == PQP_SOLVED_INACCURATE || status == PQP_MAX_ITER_REACHED);
        double *px = s.tr(), *py = s.tr() + N;   // nv >= 3N: room for x and y of every station
        const Fld seg = s.ts();
        c.sync();
        for (int i = tid; i < N; i += nt) {
            double ey = qnan, ephi = qnan, kk = qnan;
            if (has_sol) {
                const int ga = s.gxi()[i];
                ey = s.xr()[ga]; ephi = s.xr()[ga + 1]; kk = s.xr()[ga + 2];
            }
            const double angle = ref[i].z;
            const double new_angle = constraint_angle(angle + 1.57079632679489661923);
            const double tx = ref[i].x + ey * cos(new_angle);
            const double ty = ref[i].y + ey * sin(new_angle);
            out[i].x = tx; out[i].y = ty; out[i].z = angle + ephi; out[i].k = kk;
            out[i].v = 0.0; out[i].a = 0.0;
            px[i] = tx;
            py[i] = ty;
            if (bv.out_frenet) {
                double *f = bv.out_frenet + 3 * (size_t)(off + i);
                f[0] = ey; f[1] = ephi; f[2] = kk;
            }
        }
        c.sync();
        for (int i = tid; i < N; i += nt) {
            double sg = 0.0;
            if (i > 0) {
                const double dx = px[i] - px[i - 1], dy = py[i] - py[i - 1];
                sg = sqrt(dx * dx + dy * dy);
            }
            seg[i] = sg;
        }
        c.sync();
        if (tid == 0) {
            double acc = 0.0;  // sequential: same association order as the reference's running sum
            for (int i = 0; i < N; ++i) {
                acc += seg[i];
                out[i].s = acc;
            }
            bv.status[prob] = status;
            if (bv.iters) bv.iters[prob] = iter;
        }
        c.sync();
    }
};

}  // namespace pqp
