// pqp_forms.h -- host-side (C++) sparse assembly of the formulations that run on the generic
// banded kernel (pqp_gen_core.cuh):
//   "K"   SolverKAsInput::setHessianMatrix / setConstraintMatrix      solver_k_as_input.cpp:46-207
//   "KPC" SolverKpAsInputConstrained::setHessianMatrix / ...          solver_kp_as_input_constrained.cpp:45-221
// The reference fills dense Eigen matrices (O(N^2) zeros) and calls sparseView(); here the same rows
// are emitted directly as <= 4-entry sparse rows, with the unknowns renumbered station by station
// ("band order") so that the reduced KKT is banded, plus the separator table the kernel partitions on.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "pqp_device.cuh"

namespace pqp {

struct GenProblem {
    int n = 0, m = 0, n_den = 0, bw = 0, M = 0, N = 0;
    std::vector<int32_t> A_col;   // [m][4]
    std::vector<double> A_val;    // [m][4]
    std::vector<double> l, u;     // [m]
    std::vector<double> Pd;       // [n]
    std::vector<int32_t> Po_idx;  // [n][2]
    std::vector<double> Po_val;   // [n][2]
    std::vector<int32_t> csc_ptr, csc_row;
    std::vector<double> csc_val;
    int32_t sep[32];
    std::vector<int32_t> out_idx; // [N][3]
};

namespace forms_detail {
constexpr double kInf = 1e30;  // OsqpEigen::INFTY

struct Builder {
    GenProblem &g;
    explicit Builder(GenProblem &gp) : g(gp) {}
    void init(int n, int n_den, int N) {
        g.n = n; g.n_den = n_den; g.N = N; g.m = 0;
        g.Pd.assign(n, 0.0);
        g.Po_idx.assign(2 * (size_t)n, -1);
        g.Po_val.assign(2 * (size_t)n, 0.0);
        g.A_col.clear(); g.A_val.clear(); g.l.clear(); g.u.clear();
        g.out_idx.assign(3 * (size_t)N, 0);
    }
    // one constraint row with up to 4 (column, value) pairs; pass col < 0 to skip an entry
    void row(double lo, double hi, int c0, double v0, int c1 = -1, double v1 = 0, int c2 = -1, double v2 = 0, int c3 = -1,
             double v3 = 0) {
        const int cs[4] = {c0, c1, c2, c3};
        const double vs[4] = {v0, v1, v2, v3};
        for (int t = 0; t < 4; ++t) { g.A_col.push_back(cs[t]); g.A_val.push_back(cs[t] >= 0 ? vs[t] : 0.0); }
        g.l.push_back(lo); g.u.push_back(hi);
        g.m++;
    }
    void p_offdiag(int i, int j, double v) {   // symmetric pair
        for (int q = 0; q < 2; ++q) if (g.Po_idx[2 * i + q] < 0) { g.Po_idx[2 * i + q] = j; g.Po_val[2 * i + q] = v; break; }
        for (int q = 0; q < 2; ++q) if (g.Po_idx[2 * j + q] < 0) { g.Po_idx[2 * j + q] = i; g.Po_val[2 * j + q] = v; break; }
    }
    void finish() {   // CSC of A and the half-bandwidth of the reduced KKT
        const int n = g.n, m = g.m;
        g.csc_ptr.assign(n + 1, 0);
        for (int r = 0; r < m; ++r)
            for (int t = 0; t < 4; ++t) if (g.A_col[4 * r + t] >= 0) g.csc_ptr[g.A_col[4 * r + t] + 1]++;
        for (int j = 0; j < n; ++j) g.csc_ptr[j + 1] += g.csc_ptr[j];
        g.csc_row.assign(g.csc_ptr[n], 0);
        g.csc_val.assign(g.csc_ptr[n], 0.0);
        std::vector<int> next(g.csc_ptr.begin(), g.csc_ptr.end() - 1);
        int bw = 1;
        for (int r = 0; r < m; ++r) {
            int lo = 1 << 30, hi = -1;
            for (int t = 0; t < 4; ++t) {
                const int c = g.A_col[4 * r + t];
                if (c < 0) continue;
                g.csc_row[next[c]] = r;
                g.csc_val[next[c]++] = g.A_val[4 * r + t];
                lo = std::min(lo, c); hi = std::max(hi, c);
            }
            if (hi >= 0) bw = std::max(bw, hi - lo);
        }
        for (int j = 0; j < n; ++j)
            for (int q = 0; q < 2; ++q) if (g.Po_idx[2 * j + q] >= 0) bw = std::max(bw, std::abs(g.Po_idx[2 * j + q] - j));
        g.bw = bw;
    }
};

inline double constraint_angle_h(double a) {   // tools.hpp:24-35
    while (a > M_PI) a -= 2 * M_PI;
    while (a < -M_PI) a += 2 * M_PI;
    return a;
}
inline void end_window(const pqp_params &p, double end_heading, double ref_back_z, double *lo, double *hi) {
    *lo = -kInf; *hi = kInf;   // e.g. solver_k_as_input.cpp:172-178
    if (p.constraint_end_heading) {
        const double end_psi = constraint_angle_h(end_heading - ref_back_z);
        if (end_psi < 70 * M_PI / 180) { *lo = end_psi - 5 * M_PI / 180; *hi = end_psi + 5 * M_PI / 180; }
    }
}
}  // namespace forms_detail

// ---- "K": solver_k_as_input.cpp ------------------------------------------------------------------
// band order per station i: e_phi (4i), e_y (4i+1), steering control (4i+2, i < N-1), slack (4i+3;
// 4(N-1)+2 for the last station).  Separator = (e_phi, e_y, control) of stations 0, L, 2L, ... < N-1.
inline bool assemble_k(const pqp_params &p, int N, const pqp_state *ref, const pqp_station_bounds *b, const double x0[3],
                       double end_heading, GenProblem &g) {
    using namespace forms_detail;
    if (N < 2) return false;
    Builder B(g);
    const int n = 4 * N - 1;                      // :18
    B.init(n, n, N);
    auto pb = [&](int i) { return 4 * i; };
    auto pa = [&](int i) { return 4 * i + 1; };
    auto pd = [&](int i) { return 4 * i + 2; };   // i < N-1
    auto ps = [&](int i) { return i < N - 1 ? 4 * i + 3 : 4 * i + 2; };
    const double w_c = p.K_curvature_weight, w_cr = p.K_curvature_rate_weight, w_pq = p.K_deviation_weight,
                 w_e = p.KP_slack_weight;          // :50-53
    const int nc = N - 1;
    for (int i = 0; i < N; ++i) { g.Pd[pa(i)] = w_pq; g.Pd[ps(i)] = w_e; }                 // Q, S :57-59,78
    for (int i = 0; i < nc; ++i) {                                                       // R :61-76
        g.Pd[pd(i)] = (i == 0 || i == nc - 1) ? (w_c + w_cr) : (w_cr * 2 + w_c);
        if (i + 1 < nc) B.p_offdiag(pd(i), pd(i + 1), -w_cr);
    }
    // transition rows, original order :112-121 with bounds :156-167
    B.row(-x0[1], -x0[1], pb(0), -1.0);            // x0 << err[1], err[0]
    B.row(-x0[0], -x0[0], pa(0), -1.0);
    for (int i = 0; i + 1 < N; ++i) {
        const double ref_k = ref[i].k, ref_s = ref[i + 1].s - ref[i].s;
        const double ref_delta = atan(ref_k * p.wheel_base);
        const double b0 = ref_s / p.wheel_base / pow(cos(ref_delta), 2);       // setDynamicMatrix :89-103
        const double steer = atan(ref[i].k * p.wheel_base);
        const double c0 = ref_s * steer / p.wheel_base / pow(cos(steer), 2);   // :163-166
        B.row(c0, c0, pb(i + 1), -1.0, pb(i), 1.0, pa(i), -ref_s * pow(ref_k, 2), pd(i), b0);
        B.row(0.0, 0.0, pa(i + 1), -1.0, pb(i), ref_s, pa(i), 1.0);
    }
    // "variable constraint part": identity rows over every variable :124-126, bounds :169-187
    double lo, hi;
    end_window(p, end_heading, ref[N - 1].z, &lo, &hi);
    for (int i = 0; i < N; ++i) {
        const bool endrow = (i == N - 1) && hi < kInf;
        B.row(endrow ? lo : -kInf, endrow ? hi : kInf, pb(i), 1.0);
        B.row(-kInf, kInf, pa(i), 1.0);
    }
    for (int i = 0; i < nc; ++i) B.row(-p.max_steering_angle, p.max_steering_angle, pd(i), 1.0);
    for (int i = 0; i < N; ++i) B.row(0.0, p.expected_safety_margin, ps(i), 1.0);
    const double mg = p.expected_safety_margin;
    for (int i = 0; i < N; ++i) {                  // collision part 1 :129-137,189-199
        B.row(b[i].c0_lb, b[i].c0_ub, pb(i), p.d1, pa(i), 1.0);
        B.row(b[i].c2_lb, b[i].c2_ub, pb(i), p.d3, pa(i), 1.0);
        B.row(b[i].c3_lb, b[i].c3_ub, pb(i), p.d4, pa(i), 1.0);
    }
    for (int i = 0; i < N; ++i) {                  // part 2 (second circle, soft) :141-147,200-207
        B.row(-kInf, b[i].c1_ub - mg, pb(i), p.d2, pa(i), 1.0, ps(i), -1.0);
        B.row(b[i].c1_lb + mg, kInf, pb(i), p.d2, pa(i), 1.0, ps(i), 1.0);
    }
    for (int i = 0; i < N; ++i) {                  // getOptimizedPath :22-44
        g.out_idx[3 * i] = pa(i);
        g.out_idx[3 * i + 1] = pb(i);
        g.out_idx[3 * i + 2] = (i != N - 1) ? pd(i) : pd(N - 2);
    }
    B.finish();
    int L = (N - 2) / 31 + 1;
    if (L < 2) L = 2;
    g.M = 0;
    for (int e = 0; e < N - 1 && g.M < 32; e += L) g.sep[g.M++] = pb(e);
    return true;
}

// ---- "KPC": solver_kp_as_input_constrained.cpp ---------------------------------------------------
// band order per station i: e_y, e_phi, kappa, collision slack, curvature slack (5 unknowns), with the
// held control u_j and its rate slack placed after station 4j+2.  The reference also allocates
// N - ch slack variables that appear in no row and have zero cost (slack_size_ = 3N, :21): they stay 0
// and only count in OSQP's cost-scaling mean (n_den).
inline bool assemble_kpc(const pqp_params &p, int N, const pqp_state *ref, const pqp_station_bounds *b, const double x0[3],
                         double end_heading, const double *max_k, const double *max_kp, GenProblem &g) {
    using namespace forms_detail;
    if (N < 2 || !max_k || !max_kp) return false;
    const int keep = 4, h = 2;                     // :17
    const int ch = (N + keep - 2) / keep;          // :18
    Builder B(g);
    const int n = 5 * N + 2 * ch;
    B.init(n, 6 * N + ch, N);                      // :22 num_of_variables_ = 3N + ch + 3N
    auto cnt_before = [&](int i) { int r = i - h; if (r <= 0) return 0; int c = (r - 1) / keep + 1; return c < ch ? c : ch; };
    auto base = [&](int i) { return 5 * i + 2 * cnt_before(i); };
    auto home = [&](int j) { int hm = j * keep + h; return hm > N - 1 ? N - 1 : hm; };
    auto pu = [&](int j) { return 5 * (home(j) + 1) + 2 * j; };
    auto pkp = [&](int j) { return pu(j) + 1; };
    auto pa = [&](int i) { return base(i); };
    auto pbb = [&](int i) { return base(i) + 1; };
    auto pc = [&](int i) { return base(i) + 2; };
    auto ps = [&](int i) { return base(i) + 3; };
    auto psk = [&](int i) { return base(i) + 4; };
    const double w_c = p.KP_curvature_weight, w_cr = p.KP_curvature_rate_weight, w_pq = p.KP_deviation_weight,
                 w_s = p.KP_slack_weight, w_k_slack = 500, w_kp_slack = 25000;   // :48-53
    for (int i = 0; i < N; ++i) { g.Pd[pa(i)] = w_pq; g.Pd[pc(i)] = w_c; g.Pd[ps(i)] = w_s; g.Pd[psk(i)] = w_k_slack; }
    for (int j = 0; j < ch; ++j) { g.Pd[pu(j)] = keep * w_cr; g.Pd[pkp(j)] = w_kp_slack * keep; }
    // transition part :81-104, bounds :161-169
    B.row(-x0[0], -x0[0], pa(0), -1.0);
    B.row(-x0[1], -x0[1], pbb(0), -1.0);
    B.row(-x0[2], -x0[2], pc(0), -1.0);
    for (int i = 0; i + 1 < N; ++i) {
        const double ref_k = ref[i].k, ds = ref[i + 1].s - ref[i].s;
        const double ref_kp = (ref[i + 1].k - ref_k) / ds;
        const double c0 = ds * ((0.0 - 0.0) - 0.0 * ref_kp), c1 = ds * ((0.0 - ref_k) - 0.0 * ref_kp),
                     c2 = ds * ((ref_kp - 0.0) - 1.0 * ref_kp);
        B.row(-c0, -c0, pa(i + 1), -1.0, pa(i), 1.0, pbb(i), 1.0 * ds);
        B.row(-c1, -c1, pbb(i + 1), -1.0, pa(i), -pow(ref_k, 2) * ds, pbb(i), 1.0, pc(i), 1.0 * ds);
        B.row(-c2, -c2, pc(i + 1), -1.0, pc(i), 1.0, pu(i / keep), 1.0 * ds);
    }
    const double kmax = tan(p.max_steering_angle) / p.wheel_base;
    for (int i = 0; i < N; ++i) B.row(-max_k[i], kInf, pc(i), 1.0, psk(i), 1.0);          // kl :108-109,175-176
    for (int i = 0; i < N; ++i) B.row(-kInf, max_k[i], pc(i), 1.0, psk(i), -1.0);         // ku :110-111,177-178
    for (int j = 0; j < ch; ++j) B.row(-max_kp[j], kInf, pu(j), 1.0, pkp(j), 1.0);        // kpl :117-118,188-189
    for (int j = 0; j < ch; ++j) B.row(-kInf, max_kp[j], pu(j), 1.0, pkp(j), -1.0);       // kpu :119-120,190-191
    for (int i = 0; i < N; ++i) B.row(0.0, p.expected_safety_margin, ps(i), 1.0);         // slack boxes :112-113,180-184
    for (int i = 0; i < N; ++i) B.row(0.0, std::max(kmax - max_k[i], 0.0), psk(i), 1.0);
    for (int j = 0; j < ch; ++j) B.row(0.0, kInf, pkp(j), 1.0);                           // :121,193-194
    const double mg = p.expected_safety_margin;
    for (int i = 0; i < N; ++i) {                  // collision :126-142,198-214
        B.row(b[i].c0_lb, b[i].c0_ub, pa(i), 1.0, pbb(i), p.d1);
        B.row(b[i].c1_lb, b[i].c1_ub, pa(i), 1.0, pbb(i), p.d2);
        B.row(b[i].c3_lb, b[i].c3_ub, pa(i), 1.0, pbb(i), p.d4);
    }
    for (int i = 0; i < N; ++i) B.row(-kInf, b[i].c2_ub - mg, pa(i), 1.0, pbb(i), p.d3, ps(i), -1.0);
    for (int i = 0; i < N; ++i) B.row(b[i].c2_lb + mg, kInf, pa(i), 1.0, pbb(i), p.d3, ps(i), 1.0);
    double lo, hi;                                 // end state :145-146,216-228
    end_window(p, end_heading, ref[N - 1].z, &lo, &hi);
    B.row(-kInf, kInf, pa(N - 1), 1.0);
    B.row(lo, hi, pbb(N - 1), 1.0);
    for (int i = 0; i < N; ++i) { g.out_idx[3 * i] = pa(i); g.out_idx[3 * i + 1] = pbb(i); g.out_idx[3 * i + 2] = pc(i); }
    B.finish();
    int L = keep * ((N - 1) / (31 * keep) + 1);
    g.M = 0;
    for (int e = 0; e <= N - 1 && g.M < 32; e += L) g.sep[g.M++] = pa(e);
    return true;
}

}  // namespace pqp
