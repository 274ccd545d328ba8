// kp_emu.cpp -- TEST HARNESS ONLY: runs the product's warp-synchronous solver source
// (path_optimizer_b200/csrc/pqp_kp_core.cuh) on the CPU by backing each of the 32 lanes with a host
// thread (see pqp_warp.cuh, PQP_HOST_EMU).  It exists so that the kernel logic can be checked against
// the oracle on a machine without a GPU.  It is NOT part of libpqp.so and the product never calls it.
#define PQP_HOST_EMU 1
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../../path_optimizer_b200/csrc/pqp_kp_core2.cuh"

namespace {
struct LaneArgs {
    pqp::EmuShared *sh;
    int lane;
    const pqp::DevParams *prm;
    const pqp::BatchView *bv;
    int prob;
    double *smem;
    size_t smem_doubles;
    int variant;  // 0: generic v1 core; 1: v2 <17,6>; 2: v2 <10,7>; 3: v2 <27,7>; 4: v2 <48,7>
};
void *lane_main(void *p) {
    LaneArgs *a = (LaneArgs *)p;
    pqp::Warp w{a->lane, a->sh};
    switch (a->variant) {
    case 1: pqp::Kp2<17, 6>::solve_path(w, *a->prm, *a->bv, a->prob, a->smem, a->smem_doubles); break;
    case 2: pqp::Kp2<10, 7>::solve_path(w, *a->prm, *a->bv, a->prob, a->smem, a->smem_doubles); break;
    case 3: pqp::Kp2<27, 7>::solve_path(w, *a->prm, *a->bv, a->prob, a->smem, a->smem_doubles); break;
    case 4: pqp::Kp2<48, 7>::solve_path(w, *a->prm, *a->bv, a->prob, a->smem, a->smem_doubles); break;
    default: pqp::kp_solve_path(w, *a->prm, *a->bv, a->prob, a->smem, a->smem_doubles);
    }
    return nullptr;
}
}  // namespace

extern "C" int kp_emu_solve_batch(const pqp_params *params, int batch, const int32_t *n_points,
                                  const int32_t *offsets, const pqp_state *ref,
                                  const pqp_station_bounds *bounds, const double *x0,
                                  const double *end_heading, pqp_state *out_states, double *out_frenet,
                                  int32_t *status, int32_t *iters, int smem_bytes, int variant) {
    pqp::DevParams prm = pqp::dev_params_from(*params);
    pqp::BatchView bv;
    bv.batch = batch; bv.n_points = n_points; bv.offsets = offsets; bv.ref = ref; bv.bounds = bounds;
    bv.x0 = x0; bv.end_heading = end_heading; bv.out_states = out_states; bv.out_frenet = out_frenet;
    bv.status = status; bv.iters = iters;
    const size_t ws_n = pqp::kp2_ws_doubles((size_t)offsets[batch], (size_t)batch);
    bv.workspace = (double *)malloc(ws_n * sizeof(double));
    for (size_t k = 0; k < ws_n; ++k) bv.workspace[k] = nan("");
    const size_t smem_doubles = (size_t)smem_bytes / sizeof(double);
    double *smem = (double *)malloc(smem_doubles * sizeof(double));
    for (int prob = 0; prob < batch; ++prob) {
        // poison the scratch so that reads of uninitialised shared memory show up as NaN
        for (size_t k = 0; k < smem_doubles; ++k) smem[k] = nan("");
        pqp::EmuShared sh;
        pthread_barrier_init(&sh.bar, nullptr, 32);
        pthread_t th[32];
        LaneArgs args[32];
        for (int l = 0; l < 32; ++l) {
            args[l] = LaneArgs{&sh, l, &prm, &bv, prob, smem, smem_doubles, variant};
            pthread_create(&th[l], nullptr, lane_main, &args[l]);
        }
        for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
        pthread_barrier_destroy(&sh.bar);
    }
    free(smem);
    free(bv.workspace);
    return 0;
}
