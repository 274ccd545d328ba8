// env_emu.cpp -- TEST HARNESS: the per-thread device functions of pqp_env_core.cuh compiled with
// plain g++ (-DPQP_HOST_EMU) and driven by loops that mirror the kernels of pqp_env.cu
// (thread = (station, circle) for the bounds; first-failure cut for the tails).  Lets the CPU test
// suite check the kernel source against the oracle without a GPU; never linked into libpqp.so.
#define PQP_HOST_EMU 1
#include <string.h>

#include <vector>

#include "../../path_optimizer_b200/csrc/pqp_env_core.cuh"

using namespace pqp;

static MapView view_of(const pqp_distance_map *m) {
    return make_map_view(m->distance, m->rows, m->cols, m->resolution, m->center_x, m->center_y);
}

extern "C" {

void env_emu_map_distance(const pqp_distance_map *m, int n, const double *xy, double *out) {
    const MapView mv = view_of(m);
    for (int i = 0; i < n; ++i) out[i] = map_distance(mv, xy[2 * i], xy[2 * i + 1]);
}

void env_emu_update_bounds(const pqp_params *prm, const pqp_distance_map *m, int mode, int batch,
                           const int32_t *n_points, const pqp_state *ref, const int32_t *n_knots, const double *knots,
                           const double *xc, const double *yc, pqp_station_bounds *out, int32_t *n_valid) {
    const MapView mv = view_of(m);
    const double d[4] = {prm->d1, prm->d2, prm->d3, prm->d4};
    int off = 0, koff = 0;
    for (int b = 0; b < batch; ++b) {
        const int n = n_points[b];
        SplineView xs{0, nullptr, nullptr}, ys{0, nullptr, nullptr};
        if (mode == PQP_BOUNDS_IMPROVED) {
            xs = SplineView{n_knots[b], knots + koff, xc + 4 * (size_t)koff};
            ys = SplineView{n_knots[b], knots + koff, yc + 4 * (size_t)koff};
            koff += n_knots[b];
        }
        int first = n;
        for (int t = 0; t < 4 * n; ++t) {            // one kernel thread each
            const int i = t >> 2, j = t & 3;
            double ub, lb;
            const bool blocked = circle_bounds(mv, prm->circle_radius, mode, ref[off + i], d[j], xs, ys, ub, lb);
            double *o = &out[off + i].c0_ub + 2 * j;
            o[0] = ub; o[1] = lb;
            if (blocked && i < first) first = i;     // atomicMin
        }
        n_valid[b] = first;
        off += n;
    }
}

void env_emu_check_states(const pqp_params *prm, const pqp_distance_map *m, int n, const pqp_state *s, int32_t *ok) {
    const MapView mv = view_of(m);
    const CarCircles car = make_car_circles(*prm);
    for (int i = 0; i < n; ++i) ok[i] = state_collision_free(mv, car, s[i].x, s[i].y, s[i].z) ? 1 : 0;
}

void env_emu_finish_raw(const pqp_params *prm, const pqp_distance_map *m, int batch, const int32_t *n_points,
                        pqp_state *paths, int collision_check, int32_t *n_kept, int32_t *ok) {
    const MapView mv = view_of(m);
    const CarCircles car = make_car_circles(*prm);
    int off = 0;
    for (int b = 0; b < batch; ++b) {
        const int n = n_points[b];
        pqp_state *p = paths + off;
        int first = n;
        if (collision_check)
            for (int i = 0; i < n; ++i)
                if (!state_collision_free(mv, car, p[i].x, p[i].y, p[i].z) && i < first) first = i;
        accumulate_s(n, p);
        n_kept[b] = first;
        ok[b] = (first >= n) ? 1 : (first > 0 ? (p[first - 1].s >= 20 ? 1 : 0) : 0);
        off += n;
    }
}

void env_emu_densify(const pqp_params *prm, const pqp_distance_map *m, int batch, const int32_t *n_points,
                     const pqp_state *paths, double spacing, int collision_check, int max_out, pqp_state *out_all,
                     int32_t *n_out, int32_t *ok) {
    const MapView mv = view_of(m);
    const CarCircles car = make_car_circles(*prm);
    int off = 0;
    for (int b = 0; b < batch; ++b) {
        const int n = n_points[b];
        const pqp_state *p = paths + off;
        pqp_state *out = out_all + (size_t)b * max_out;
        off += n;
        if (n < 3) { n_out[b] = 0; ok[b] = 0; continue; }
        std::vector<double> ws((size_t)13 * n);
        double *t = ws.data(), *xc = t + n, *yc = t + 5 * (size_t)n, *scr = t + 9 * (size_t)n;
        for (int i = 0; i < n; ++i) t[i] = p[i].s;
        spline_fit(n, t, [&](int i) { return p[i].x; }, xc, scr, scr + n);
        spline_fit(n, t, [&](int i) { return p[i].y; }, yc, scr + 2 * (size_t)n, scr + 3 * (size_t)n);
        const double s_end = t[n - 1];
        long long cnt = 0;
        if (s_end >= 0.0 && spacing > 0.0) {
            cnt = (long long)(s_end / spacing) + 1;
            while (cnt > 0 && !(mul((double)(cnt - 1), spacing) <= s_end)) --cnt;
            while (mul((double)cnt, spacing) <= s_end) ++cnt;
        }
        const SplineView xs{n, t, xc}, ys{n, t, yc};
        const long long lim = cnt < (long long)max_out + 1 ? cnt : (long long)max_out + 1;
        long long first = 0x7fffffff;
        for (long long i = 0; i < lim; ++i) {
            const pqp_state st = densify_sample(xs, ys, mul((double)i, spacing));
            if (collision_check && !state_collision_free(mv, car, st.x, st.y, st.z) && i < first) first = i;
            if (i < max_out) out[i] = st;
        }
        if (first < lim) {
            n_out[b] = (int)first;
            ok[b] = first > 0 ? (mul((double)(first - 1), spacing) >= 20 ? 1 : 0) : 0;
        } else if (cnt > max_out) {
            n_out[b] = max_out; ok[b] = 0;
        } else {
            n_out[b] = (int)cnt; ok[b] = 1;
        }
    }
}

}  // extern "C"
