/*
 * pqp_oracle_api.c -- CPU ORACLE drivers (test infrastructure, not product code; see
 * pqp_oracle.h): one whole hot-path call per path = OsqpSolver::solve (solver.cpp:46-77), and a
 * pthread batch loop used as the CPU baseline.  PARITY UNPINNED -- see pqp_oracle.h.
 */
#include "pqp_oracle.h"
#define PQP_ORACLE_ARENA_IMPL   /* this file keeps the C library allocator; it only drives the arena */
#include "pqp_oracle_arena.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

int oracle_solve_path(const pqp_params *prm, int formulation, int n, const pqp_state *ref,
                      const pqp_station_bounds *bounds, const double x0[3], double end_heading,
                      const double *max_k, const double *max_kp, pqp_state *out, double *frenet,
                      oqp_info *info) {
    oqp_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof(*info));
    oqp_problem *qp = oracle_assemble(prm, formulation, n, ref, bounds, x0, end_heading, max_k, max_kp);
    if (!qp) {
        info->status = PQP_INVALID_PROBLEM;
        return info->status;
    }
    double *x = (double *)malloc(sizeof(double) * (size_t)qp->n);
    int st = oracle_osqp_solve(prm, qp, x, NULL, info, NULL, 0);
    /* the reference only calls getOptimizedPath when solve() returned true (solver.cpp:73-75);
     * for any other status the output path is left untouched -> we fill NaN via x. */
    oracle_extract(formulation, n, ref, x, out, frenet);
    free(x);
    oracle_problem_free(qp);
    return st;
}

typedef struct {
    const pqp_params *prm;
    int formulation, begin, end;
    const int32_t *n_points;
    const int *offsets, *ch_offsets;
    const pqp_state *ref;
    const pqp_station_bounds *bounds;
    const double *x0, *end_heading, *max_k, *max_kp;
    pqp_state *out;
    double *frenet;
    int32_t *status, *iters;
    int *next;   /* shared work counter */
} job;

/* Paths are handed out one at a time from a shared counter (iteration counts differ by 2x between paths, so a
 * static split leaves threads idle), and every thread solves out of its own scratch arena (pqp_oracle_arena.h). */
static void *worker(void *arg) {
    job *j = (job *)arg;
    oa_begin((size_t)64 << 20);
    for (;;) {
        const int b = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
        if (b >= j->end) break;
        oa_reset();
        const int off = j->offsets[b], n = j->n_points[b];
        oqp_info info;
        int st = oracle_solve_path(j->prm, j->formulation, n, j->ref + off, j->bounds + off,
                                   j->x0 + 3 * (size_t)b, j->end_heading[b],
                                   j->max_k ? j->max_k + off : NULL,
                                   j->max_kp ? j->max_kp + off : NULL,
                                   j->out + off, j->frenet ? j->frenet + 3 * (size_t)off : NULL, &info);
        if (j->status) j->status[b] = st;
        if (j->iters) j->iters[b] = info.iters;
    }
    oa_end();
    return NULL;
}

double oracle_solve_batch(const pqp_params *prm, int formulation, int batch, const int32_t *n_points,
                          const pqp_state *ref, const pqp_station_bounds *bounds, const double *x0,
                          const double *end_heading, const double *max_k, const double *max_kp,
                          pqp_state *out, double *frenet, int32_t *status, int32_t *iters,
                          int threads) {
    if (threads < 1) threads = 1;
    if (threads > batch) threads = batch > 0 ? batch : 1;
    int *offsets = (int *)malloc(sizeof(int) * (size_t)(batch + 1));
    int *ch_offsets = (int *)malloc(sizeof(int) * (size_t)(batch + 1));
    offsets[0] = 0;
    ch_offsets[0] = 0;
    for (int b = 0; b < batch; ++b) {
        offsets[b + 1] = offsets[b] + n_points[b];
        ch_offsets[b + 1] = ch_offsets[b] + (n_points[b] + 4 - 2) / 4; /* KPC: keep = 4 */
    }
    int next = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    job *jobs = (job *)malloc(sizeof(job) * (size_t)threads);
    for (int t = 0; t < threads; ++t) {
        job *j = &jobs[t];
        j->prm = prm; j->formulation = formulation;
        j->begin = 0;
        j->end = batch;
        j->next = &next;
        j->n_points = n_points; j->offsets = offsets; j->ch_offsets = ch_offsets;
        j->ref = ref; j->bounds = bounds; j->x0 = x0; j->end_heading = end_heading;
        j->max_k = max_k; j->max_kp = max_kp; j->out = out; j->frenet = frenet;
        j->status = status; j->iters = iters;
        if (threads == 1) worker(j);
        else pthread_create(&th[t], NULL, worker, j);
    }
    if (threads > 1)
        for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(jobs); free(offsets); free(ch_offsets);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
