/*
 * pqp_oracle_assemble.c -- CPU ORACLE (test infrastructure, not product code; see pqp_oracle.h).
 *
 * Restates, in sparse-triplet form, the QP the reference builds as dense Eigen matrices:
 *   KP : src/solver/solver_kp_as_input.cpp:13-203
 *   K  : src/solver/solver_k_as_input.cpp:14-207
 *   KPC: src/solver/solver_kp_as_input_constrained.cpp:13-221
 * plus flag defaults (src/config/planning_flags.cpp) and getOptimizedPath.
 * PARITY UNPINNED (no reference golden vectors exist for this path) -- see pqp_oracle.h.
 */
#include "pqp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "pqp_oracle_arena.h"   /* malloc / calloc / free of this file go through the per-thread arena when it is on */

#define OQP_INFTY 1e30 /* OsqpEigen::INFTY == OSQP_INFTY [upstream OSQP 0.6.x] */

/* ---- flags: planning_flags.cpp:18-119, updateConfig :8-14 ---------------------------------- */
void oracle_params_default(pqp_params *p) {
    memset(p, 0, sizeof(*p));
    p->car_width = 2.0;                       /* :18 */
    p->car_length = 4.9;                      /* :20 */
    p->safety_margin = 0.0;                   /* :22 */
    p->wheel_base = 2.85;                     /* :27 */
    p->rear_axle_to_center = 1.45;            /* :29 */
    p->max_steering_angle = 30.0 * M_PI / 180.0; /* :39 */
    p->mu = 0.4;                              /* :41 */
    p->max_curvature_rate = 0.1;              /* :43 */
    /* updateConfig(), :8-14 */
    p->circle_radius = sqrt(pow(p->car_length / 8, 2) + pow(p->car_width / 2, 2)) + p->safety_margin;
    p->d1 = -3.0 / 8.0 * p->car_length + p->rear_axle_to_center;
    p->d2 = -1.0 / 8.0 * p->car_length + p->rear_axle_to_center;
    p->d3 = 1.0 / 8.0 * p->car_length + p->rear_axle_to_center;
    p->d4 = 3.0 / 8.0 * p->car_length + p->rear_axle_to_center;
    p->K_curvature_weight = 50;               /* :102 */
    p->K_curvature_rate_weight = 200;         /* :104 */
    p->K_deviation_weight = 0;                /* :106 */
    p->KP_curvature_weight = 10;              /* :108 */
    p->KP_curvature_rate_weight = 200;        /* :110 */
    p->KP_deviation_weight = 0;               /* :112 */
    p->KP_slack_weight = 3;                   /* :114 */
    p->expected_safety_margin = 1.3;          /* :116 */
    p->constraint_end_heading = 1;            /* :119 */
    /* OSQP 0.6.x defaults [upstream]; the reference only sets verbosity and warm start
     * (solver.cpp:48-49). */
    p->rho = 0.1;
    p->sigma = 1e-6;
    p->alpha = 1.6;
    p->eps_abs = 1e-3;
    p->eps_rel = 1e-3;
    p->eps_prim_inf = 1e-4;
    p->eps_dual_inf = 1e-4;
    p->max_iter = 4000;
    p->scaling = 10;
    p->check_termination = 25;
    p->adaptive_rho = 1;
    p->adaptive_rho_interval = 25; /* deterministic stand-in for OSQP's timing-based choice, see pqp.h */
    p->adaptive_rho_tolerance = 5;
}

/* tools.hpp:24-35 */
static double constraint_angle(double angle) {
    if (angle > M_PI) {
        angle -= 2 * M_PI;
        return constraint_angle(angle);
    } else if (angle < -M_PI) {
        angle += 2 * M_PI;
        return constraint_angle(angle);
    }
    return angle;
}

/* solver.cpp:21-27 then solver_kp_as_input.cpp:17 / ..._constrained.cpp:17 */
int oracle_keep_control_steps(int formulation, const pqp_state *ref, int n) {
    if (formulation == PQP_FORM_K) return 1;
    if (formulation == PQP_FORM_KPC) return 4;
    double interval = 0;
    const int check_num = 10;
    for (int i = 1; i < n && i < check_num; ++i) {
        double d = ref[i].s - ref[i - 1].s;
        interval = interval > d ? interval : d; /* std::max(a, b): a if !(a < b) */
    }
    int keep = (int)(1.2 / interval);
    return keep > 1 ? keep : 1;
}

/* ---- triplet builder ------------------------------------------------------------------------ */

static oqp_problem *qp_alloc(int n, int m, int p_cap, int a_cap) {
    oqp_problem *qp = (oqp_problem *)calloc(1, sizeof(*qp));
    qp->n = n;
    qp->m = m;
    qp->p_i = (int *)malloc(sizeof(int) * (size_t)p_cap);
    qp->p_j = (int *)malloc(sizeof(int) * (size_t)p_cap);
    qp->p_v = (double *)malloc(sizeof(double) * (size_t)p_cap);
    qp->a_i = (int *)malloc(sizeof(int) * (size_t)a_cap);
    qp->a_j = (int *)malloc(sizeof(int) * (size_t)a_cap);
    qp->a_v = (double *)malloc(sizeof(double) * (size_t)a_cap);
    qp->q = (double *)calloc((size_t)n, sizeof(double)); /* gradient = 0, solver.cpp:54 */
    qp->l = (double *)calloc((size_t)m, sizeof(double));
    qp->u = (double *)calloc((size_t)m, sizeof(double));
    return qp;
}

void oracle_problem_free(oqp_problem *qp) {
    if (!qp) return;
    free(qp->p_i); free(qp->p_j); free(qp->p_v);
    free(qp->a_i); free(qp->a_j); free(qp->a_v);
    free(qp->q); free(qp->l); free(qp->u);
    free(qp);
}

static void addP(oqp_problem *qp, int i, int j, double v) { /* upper triangle, i <= j */
    int k = qp->p_nnz++;
    qp->p_i[k] = i; qp->p_j[k] = j; qp->p_v[k] = v;
}
static void addA(oqp_problem *qp, int i, int j, double v) {
    int k = qp->a_nnz++;
    qp->a_i[k] = i; qp->a_j[k] = j; qp->a_v[k] = v;
}

/* end-heading window shared by the three formulations, e.g. solver_kp_as_input.cpp:195-201.
 * NB no fabs on end_psi in the reference: large negative differences are constrained too. */
static void end_heading_window(const pqp_params *prm, double end_heading, double ref_back_z,
                               double *lo, double *hi) {
    *lo = -OQP_INFTY;
    *hi = OQP_INFTY;
    if (prm->constraint_end_heading) {
        double end_psi = constraint_angle(end_heading - ref_back_z);
        if (end_psi < 70 * M_PI / 180) {
            *lo = end_psi - 5 * M_PI / 180;
            *hi = end_psi + 5 * M_PI / 180;
        }
    }
}

/* ---- KP: solver_kp_as_input.cpp ------------------------------------------------------------- */
static oqp_problem *assemble_kp(const pqp_params *prm, int N, const pqp_state *ref,
                                const pqp_station_bounds *b, const double x0[3],
                                double end_heading) {
    const int keep = oracle_keep_control_steps(PQP_FORM_KP, ref, N);   /* :17 */
    const int ch = (N + keep - 2) / keep;                               /* :18 */
    const int state_size = 3 * N, control_size = ch, slack_size = 2 * N; /* :19-21 */
    const int n = state_size + control_size + slack_size;               /* :22 */
    const int m = 11 * N + ch + 2;                                      /* :23 */
    oqp_problem *qp = qp_alloc(n, m, n, 30 * N + ch + 8);

    /* Hessian :45-63 (diagonal) */
    const double w_c = prm->KP_curvature_weight, w_cr = prm->KP_curvature_rate_weight;
    const double w_pq = prm->KP_deviation_weight, w_s = prm->KP_slack_weight;
    for (int i = 0; i < N; ++i) {
        addP(qp, 3 * i, 3 * i, w_pq);
        addP(qp, 3 * i + 2, 3 * i + 2, w_c);
        addP(qp, state_size + control_size + i, state_size + control_size + i, w_s);
        addP(qp, state_size + control_size + N + i, state_size + control_size + N + i, w_s);
    }
    for (int j = 0; j < ch; ++j) addP(qp, state_size + j, state_size + j, keep * w_cr);

    /* Constraints :65-203 */
    const int vars_begin = 3 * N;                       /* :70 */
    const int coll_begin = vars_begin + 2 * N + ch;     /* :71 */
    const int end_begin = coll_begin + 6 * N;           /* :72 */
    for (int i = 0; i < state_size; ++i) addA(qp, i, i, -1);            /* :75-77 */
    for (int i = 0; i + 1 < N; ++i) {                                   /* :84-98 */
        const double ref_k = ref[i].k;
        const double ds = ref[i + 1].s - ref[i].s;
        /* A = a*ds + I with a = [[0,1,0],[-k^2,0,1],[0,0,0]] */
        const int r = 3 * (i + 1), c = 3 * i;
        addA(qp, r + 0, c + 0, 1.0);
        addA(qp, r + 0, c + 1, 1.0 * ds);
        addA(qp, r + 1, c + 0, -pow(ref_k, 2) * ds);
        addA(qp, r + 1, c + 1, 1.0);
        addA(qp, r + 1, c + 2, 1.0 * ds);
        addA(qp, r + 2, c + 2, 1.0);
        addA(qp, r + 2, state_size + i / keep, 1.0 * ds);               /* B = b*ds :94 */
        /* bounds = -c_list[i], c_list = ds*(c - a*ref_state - b*ref_kp) = ds*(0,-k,0)  :95-97,149 */
        const double ref_kp = (ref[i + 1].k - ref_k) / ds;
        const double c0 = ds * ((0.0 - 0.0) - 0.0 * ref_kp);
        const double c1 = ds * ((0.0 - ref_k) - 0.0 * ref_kp);
        const double c2 = ds * ((ref_kp - 0.0) - 1.0 * ref_kp);
        qp->l[r + 0] = qp->u[r + 0] = -c0;
        qp->l[r + 1] = qp->u[r + 1] = -c1;
        qp->l[r + 2] = qp->u[r + 2] = -c2;
    }
    qp->l[0] = qp->u[0] = -x0[0];                                       /* :143-147 */
    qp->l[1] = qp->u[1] = -x0[1];
    qp->l[2] = qp->u[2] = -x0[2];
    const double kmax = tan(prm->max_steering_angle) / prm->wheel_base;
    for (int i = 0; i < N; ++i) {                                       /* :101-104,154-159 */
        addA(qp, vars_begin + i, 3 * i + 2, 1);
        qp->l[vars_begin + i] = -kmax;
        qp->u[vars_begin + i] = kmax;
        addA(qp, vars_begin + N + ch + i, state_size + control_size + i, 1);
        qp->l[vars_begin + N + ch + i] = 0;
        qp->u[vars_begin + N + ch + i] = prm->expected_safety_margin;
    }
    for (int j = 0; j < ch; ++j) {                                      /* :105-107,160-163 */
        addA(qp, vars_begin + N + j, state_size + j, 1);
        qp->l[vars_begin + N + j] = -OQP_INFTY;
        qp->u[vars_begin + N + j] = OQP_INFTY;
    }
    const double margin = prm->expected_safety_margin;
    const int sl = state_size + control_size;
    for (int i = 0; i < N; ++i) {                                       /* :110-133,166-187 */
        addA(qp, coll_begin + 2 * i, 3 * i, 1);
        addA(qp, coll_begin + 2 * i, 3 * i + 1, prm->d1);
        addA(qp, coll_begin + 2 * i + 1, 3 * i, 1);
        addA(qp, coll_begin + 2 * i + 1, 3 * i + 1, prm->d3);
        qp->l[coll_begin + 2 * i] = b[i].c0_lb;
        qp->u[coll_begin + 2 * i] = b[i].c0_ub;
        qp->l[coll_begin + 2 * i + 1] = b[i].c2_lb;
        qp->u[coll_begin + 2 * i + 1] = b[i].c2_ub;
        int r;
        r = coll_begin + 2 * N + i;
        addA(qp, r, 3 * i, 1); addA(qp, r, 3 * i + 1, prm->d4); addA(qp, r, sl + i, -1);
        qp->l[r] = -OQP_INFTY; qp->u[r] = b[i].c3_ub - margin;
        r = coll_begin + 3 * N + i;
        addA(qp, r, 3 * i, 1); addA(qp, r, 3 * i + 1, prm->d4); addA(qp, r, sl + i, 1);
        qp->l[r] = b[i].c3_lb + margin; qp->u[r] = OQP_INFTY;
        r = coll_begin + 4 * N + i;
        addA(qp, r, 3 * i, 1); addA(qp, r, 3 * i + 1, prm->d2); addA(qp, r, sl + i, -1);
        qp->l[r] = -OQP_INFTY; qp->u[r] = b[i].c1_ub - margin;
        r = coll_begin + 5 * N + i;
        addA(qp, r, 3 * i, 1); addA(qp, r, 3 * i + 1, prm->d2); addA(qp, r, sl + i, 1);
        qp->l[r] = b[i].c1_lb + margin; qp->u[r] = OQP_INFTY;
    }
    addA(qp, end_begin, state_size - 3, 1);                             /* :136-137 */
    addA(qp, end_begin + 1, state_size - 2, 1);
    qp->l[end_begin] = -1;                                              /* :191-192 */
    qp->u[end_begin] = 1;
    end_heading_window(prm, end_heading, ref[N - 1].z, &qp->l[end_begin + 1], &qp->u[end_begin + 1]);
    return qp;
}

/* ---- K: solver_k_as_input.cpp --------------------------------------------------------------- */
static oqp_problem *assemble_k(const pqp_params *prm, int N, const pqp_state *ref,
                               const pqp_station_bounds *b, const double x0[3],
                               double end_heading) {
    const int n = 4 * N - 1, m = 11 * N - 1;                            /* :18-19 */
    const int state_size = 2 * N, control_size = N - 1;
    oqp_problem *qp = qp_alloc(n, m, n + N, 24 * N + 8);
    const double w_c = prm->K_curvature_weight, w_cr = prm->K_curvature_rate_weight;
    const double w_pq = prm->K_deviation_weight, w_e = prm->KP_slack_weight; /* :50-53 */
    for (int i = 0; i < N; ++i) {                                       /* Q :57-59,79-81 */
        addP(qp, 2 * i, 2 * i, 0.0);
        addP(qp, 2 * i + 1, 2 * i + 1, w_pq);
    }
    for (int i = 0; i < control_size; ++i) {                            /* R :61-76 */
        double d = (i == 0 || i == control_size - 1) ? (w_c + w_cr) : (w_cr * 2 + w_c);
        addP(qp, 2 * N + i, 2 * N + i, d);
        if (i + 1 < control_size) addP(qp, 2 * N + i, 2 * N + i + 1, -w_cr);
    }
    for (int i = 0; i < N; ++i) addP(qp, 3 * N - 1 + i, 3 * N - 1 + i, w_e); /* S :78,84 */

    for (int i = 0; i < state_size; ++i) addA(qp, i, i, -1);            /* :112-114 */
    for (int i = 0; i + 1 < N; ++i) {                                   /* :117-121, 89-103 */
        const double ref_k = ref[i].k;
        const double ref_s = ref[i + 1].s - ref[i].s;
        const double ref_delta = atan(ref_k * prm->wheel_base);
        const int r = 2 * (i + 1), c = 2 * i;
        addA(qp, r + 0, c + 0, 1);
        addA(qp, r + 0, c + 1, -ref_s * pow(ref_k, 2));
        addA(qp, r + 1, c + 0, ref_s);
        addA(qp, r + 1, c + 1, 1);
        addA(qp, r + 0, 2 * N + i, ref_s / prm->wheel_base / pow(cos(ref_delta), 2));
        addA(qp, r + 1, 2 * N + i, 0.0);
        /* bounds :160-167 */
        const double ds = ref[i + 1].s - ref[i].s;
        const double steer = atan(ref[i].k * prm->wheel_base);
        qp->l[2 + 2 * i] = qp->u[2 + 2 * i] = ds * steer / prm->wheel_base / pow(cos(steer), 2);
        qp->l[2 + 2 * i + 1] = qp->u[2 + 2 * i + 1] = 0;
    }
    qp->l[0] = qp->u[0] = -x0[1];                                       /* x0 << err[1], err[0] :156-159 */
    qp->l[1] = qp->u[1] = -x0[0];
    for (int i = 0; i < 4 * N - 1; ++i) addA(qp, 2 * N + i, i, 1);      /* :124-126 */
    for (int i = 0; i < 2 * N; ++i) {                                   /* :169-170 */
        qp->l[2 * N + i] = -OQP_INFTY;
        qp->u[2 * N + i] = OQP_INFTY;
    }
    {                                                                   /* :172-178 */
        double lo, hi;
        end_heading_window(prm, end_heading, ref[N - 1].z, &lo, &hi);
        if (hi < OQP_INFTY) {
            qp->l[2 * N + 2 * N - 2] = lo;
            qp->u[2 * N + 2 * N - 2] = hi;
        }
    }
    for (int i = 0; i < N - 1; ++i) {                                   /* :180-183 */
        qp->l[4 * N + i] = -prm->max_steering_angle;
        qp->u[4 * N + i] = prm->max_steering_angle;
    }
    for (int i = 0; i < N; ++i) {                                       /* :185-187 */
        qp->l[5 * N - 1 + i] = 0;
        qp->u[5 * N - 1 + i] = prm->expected_safety_margin;
    }
    const double margin = prm->expected_safety_margin;
    for (int i = 0; i < N; ++i) {                                       /* :129-147,189-207 */
        const int r = 6 * N - 1 + 3 * i;
        addA(qp, r + 0, 2 * i, prm->d1); addA(qp, r + 0, 2 * i + 1, 1);
        addA(qp, r + 1, 2 * i, prm->d3); addA(qp, r + 1, 2 * i + 1, 1);
        addA(qp, r + 2, 2 * i, prm->d4); addA(qp, r + 2, 2 * i + 1, 1);
        qp->l[r + 0] = b[i].c0_lb; qp->u[r + 0] = b[i].c0_ub;
        qp->l[r + 1] = b[i].c2_lb; qp->u[r + 1] = b[i].c2_ub;
        qp->l[r + 2] = b[i].c3_lb; qp->u[r + 2] = b[i].c3_ub;
        const int r1 = 9 * N - 1 + i, r2 = 10 * N - 1 + i;
        addA(qp, r1, 2 * i, prm->d2); addA(qp, r1, 2 * i + 1, 1); addA(qp, r1, 3 * N - 1 + i, -1);
        addA(qp, r2, 2 * i, prm->d2); addA(qp, r2, 2 * i + 1, 1); addA(qp, r2, 3 * N - 1 + i, 1);
        qp->l[r1] = -OQP_INFTY; qp->u[r1] = b[i].c1_ub - margin;
        qp->l[r2] = b[i].c1_lb + margin; qp->u[r2] = OQP_INFTY;
    }
    return qp;
}

/* ---- KPC: solver_kp_as_input_constrained.cpp ------------------------------------------------ */
static oqp_problem *assemble_kpc(const pqp_params *prm, int N, const pqp_state *ref,
                                 const pqp_station_bounds *b, const double x0[3],
                                 double end_heading, const double *max_k, const double *max_kp) {
    const int keep = 4;                                                 /* :17 */
    const int ch = (N + keep - 2) / keep;                               /* :18 */
    const int state_size = 3 * N, control_size = ch, slack_size = 3 * N;
    const int n = state_size + control_size + slack_size;               /* :22 */
    const int m = 12 * N + 3 * ch + 2;                                  /* :23 */
    oqp_problem *qp = qp_alloc(n, m, n, 32 * N + 8 * ch + 8);
    const double w_c = prm->KP_curvature_weight, w_cr = prm->KP_curvature_rate_weight;
    const double w_pq = prm->KP_deviation_weight, w_s = prm->KP_slack_weight;
    const double w_k_slack = 500, w_kp_slack = 25000;                   /* :52-53 */
    const int sl = state_size + control_size;
    for (int i = 0; i < N; ++i) {                                       /* :54-59 */
        addP(qp, 3 * i, 3 * i, w_pq);
        addP(qp, 3 * i + 2, 3 * i + 2, w_c);
        addP(qp, sl + i, sl + i, w_s);
        addP(qp, sl + N + i, sl + N + i, w_k_slack);
    }
    for (int j = 0; j < ch; ++j) {                                      /* :60-64 */
        addP(qp, state_size + j, state_size + j, keep * w_cr);
        addP(qp, sl + 2 * N + j, sl + 2 * N + j, w_kp_slack * keep);
    }
    const int kl = 3 * N, ku = kl + N, kpl = ku + N, kpu = kpl + ch;    /* :71-77 */
    const int slack_begin = kpu + ch;
    const int coll_begin = slack_begin + 2 * N + ch;
    const int end_begin = coll_begin + 5 * N;
    for (int i = 0; i < state_size; ++i) addA(qp, i, i, -1);            /* :81-83 */
    for (int i = 0; i + 1 < N; ++i) {                                   /* :90-104 */
        const double ref_k = ref[i].k;
        const double ds = ref[i + 1].s - ref[i].s;
        const int r = 3 * (i + 1), c = 3 * i;
        addA(qp, r + 0, c + 0, 1.0);
        addA(qp, r + 0, c + 1, 1.0 * ds);
        addA(qp, r + 1, c + 0, -pow(ref_k, 2) * ds);
        addA(qp, r + 1, c + 1, 1.0);
        addA(qp, r + 1, c + 2, 1.0 * ds);
        addA(qp, r + 2, c + 2, 1.0);
        addA(qp, r + 2, state_size + i / keep, 1.0 * ds);
        const double ref_kp = (ref[i + 1].k - ref_k) / ds;
        const double c0 = ds * ((0.0 - 0.0) - 0.0 * ref_kp);
        const double c1 = ds * ((0.0 - ref_k) - 0.0 * ref_kp);
        const double c2 = ds * ((ref_kp - 0.0) - 1.0 * ref_kp);
        qp->l[r + 0] = qp->u[r + 0] = -c0;
        qp->l[r + 1] = qp->u[r + 1] = -c1;
        qp->l[r + 2] = qp->u[r + 2] = -c2;
    }
    qp->l[0] = qp->u[0] = -x0[0];                                       /* :161-165 */
    qp->l[1] = qp->u[1] = -x0[1];
    qp->l[2] = qp->u[2] = -x0[2];
    const double kmax = tan(prm->max_steering_angle) / prm->wheel_base;
    for (int i = 0; i < N; ++i) {                                       /* :108-115,173-185 */
        addA(qp, kl + i, 3 * i + 2, 1); addA(qp, kl + i, sl + N + i, 1);
        addA(qp, ku + i, 3 * i + 2, 1); addA(qp, ku + i, sl + N + i, -1);
        addA(qp, slack_begin + i, sl + i, 1);
        addA(qp, slack_begin + N + i, sl + N + i, 1);
        qp->l[kl + i] = -max_k[i]; qp->u[kl + i] = OQP_INFTY;
        qp->l[ku + i] = -OQP_INFTY; qp->u[ku + i] = max_k[i];
        qp->l[slack_begin + i] = 0; qp->u[slack_begin + i] = prm->expected_safety_margin;
        qp->l[slack_begin + N + i] = 0;
        double v = kmax - max_k[i];
        qp->u[slack_begin + N + i] = v > 0.0 ? v : 0.0;
    }
    for (int j = 0; j < ch; ++j) {                                      /* :117-123,186-195 */
        addA(qp, kpl + j, state_size + j, 1); addA(qp, kpl + j, sl + 2 * N + j, 1);
        addA(qp, kpu + j, state_size + j, 1); addA(qp, kpu + j, sl + 2 * N + j, -1);
        addA(qp, slack_begin + 2 * N + j, sl + 2 * N + j, 1);
        qp->l[kpl + j] = -max_kp[j]; qp->u[kpl + j] = OQP_INFTY;
        qp->l[kpu + j] = -OQP_INFTY; qp->u[kpu + j] = max_kp[j];
        qp->l[slack_begin + 2 * N + j] = 0; qp->u[slack_begin + 2 * N + j] = OQP_INFTY;
    }
    const double margin = prm->expected_safety_margin;
    for (int i = 0; i < N; ++i) {                                       /* :126-142,198-214 */
        const int r = coll_begin + 3 * i;
        addA(qp, r + 0, 3 * i, 1); addA(qp, r + 0, 3 * i + 1, prm->d1);
        addA(qp, r + 1, 3 * i, 1); addA(qp, r + 1, 3 * i + 1, prm->d2);
        addA(qp, r + 2, 3 * i, 1); addA(qp, r + 2, 3 * i + 1, prm->d4);
        qp->l[r + 0] = b[i].c0_lb; qp->u[r + 0] = b[i].c0_ub;
        qp->l[r + 1] = b[i].c1_lb; qp->u[r + 1] = b[i].c1_ub;
        qp->l[r + 2] = b[i].c3_lb; qp->u[r + 2] = b[i].c3_ub;
        const int r1 = coll_begin + 3 * N + i, r2 = coll_begin + 4 * N + i;
        addA(qp, r1, 3 * i, 1); addA(qp, r1, 3 * i + 1, prm->d3); addA(qp, r1, sl + i, -1);
        addA(qp, r2, 3 * i, 1); addA(qp, r2, 3 * i + 1, prm->d3); addA(qp, r2, sl + i, 1);
        qp->l[r1] = -OQP_INFTY; qp->u[r1] = b[i].c2_ub - margin;
        qp->l[r2] = b[i].c2_lb + margin; qp->u[r2] = OQP_INFTY;
    }
    addA(qp, end_begin, state_size - 3, 1);                             /* :145-146 */
    addA(qp, end_begin + 1, state_size - 2, 1);
    qp->l[end_begin] = -OQP_INFTY;                                      /* :218-219 */
    qp->u[end_begin] = OQP_INFTY;
    end_heading_window(prm, end_heading, ref[N - 1].z, &qp->l[end_begin + 1], &qp->u[end_begin + 1]);
    return qp;
}

oqp_problem *oracle_assemble(const pqp_params *prm, int formulation, int n, const pqp_state *ref,
                             const pqp_station_bounds *bounds, const double x0[3],
                             double end_heading, const double *max_k, const double *max_kp) {
    if (!prm || !ref || !bounds || !x0 || n < 2) return NULL;
    switch (formulation) {
    case PQP_FORM_KP: return assemble_kp(prm, n, ref, bounds, x0, end_heading);
    case PQP_FORM_K: return assemble_k(prm, n, ref, bounds, x0, end_heading);
    case PQP_FORM_KPC:
        if (!max_k || !max_kp) return NULL;
        return assemble_kpc(prm, n, ref, bounds, x0, end_heading, max_k, max_kp);
    default: return NULL;
    }
}

/* ---- getOptimizedPath ----------------------------------------------------------------------- */
void oracle_extract(int formulation, int N, const pqp_state *ref, const double *x,
                    pqp_state *out, double *frenet) {
    double tmp_s = 0;
    for (int i = 0; i < N; ++i) {
        double ey, ephi, k;
        if (formulation == PQP_FORM_K) {               /* solver_k_as_input.cpp:22-44 */
            ey = x[2 * i + 1];
            ephi = x[2 * i];
            k = (i != N - 1) ? x[2 * N + i] : x[3 * N - 2];
        } else {                                       /* solver_kp_as_input.cpp:26-43 */
            ey = x[3 * i];
            ephi = x[3 * i + 1];
            k = x[3 * i + 2];
        }
        double angle = ref[i].z;
        double new_angle = constraint_angle(angle + M_PI_2);
        double tmp_x = ref[i].x + ey * cos(new_angle);
        double tmp_y = ref[i].y + ey * sin(new_angle);
        if (i != 0) {
            tmp_s += sqrt(pow(tmp_x - out[i - 1].x, 2) + pow(tmp_y - out[i - 1].y, 2));
        }
        out[i].x = tmp_x; out[i].y = tmp_y; out[i].z = angle + ephi;
        out[i].k = k; out[i].s = tmp_s; out[i].v = 0; out[i].a = 0;
        if (frenet) { frenet[3 * i] = ey; frenet[3 * i + 1] = ephi; frenet[3 * i + 2] = k; }
    }
}
