/*
 * pqp_oracle_env.c -- CPU ORACLE (test infrastructure, not product code; see pqp_oracle.h) for the
 * stages either side of the QP: distance-map lookup, clearance-bounds generation, the
 * post-solve collision check / s re-accumulation, spline densification, and the
 * solveWithoutSmoothing-shaped chain.  Plain C, fp64, one path at a time, every function citing
 * the reference lines it restates.
 *
 * PARITY UNPINNED (see pqp_oracle.h): the reference has no tests or fixtures for these stages, and
 * the grid_map_core library its Map wraps (ANYbotics grid_map, apt package ros-<distro>-grid-map,
 * version not pinned by scripts/install_deps.sh:130-132) is not vendored.  The lookup below
 * restates grid_map_core's documented geometry (getIndexFromPosition / getPositionFromIndex /
 * checkIfPositionWithinMap / atPositionLinearInterpolated of the 1.6.x sources) with ONE stated
 * deviation: grid_map decides "a neighbour is outside" through a linear-buffer-index range test
 * that lets some out-of-range neighbours wrap onto another row; here a neighbour is outside
 * exactly when its (i, j) is outside [0,rows)x[0,cols).  The two differ only within half a cell
 * of the map border.
 */
#include "pqp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* Map::getObstacleDistance, src/tools/Map.cpp:16-22                                            */
/* ------------------------------------------------------------------------------------------- */

/* grid_map::GridMap::isInside -> checkIfPositionWithinMap */
int oracle_map_inside(const pqp_distance_map *m, double x, double y) {
    const double lx = m->rows * m->resolution, ly = m->cols * m->resolution;
    const double tx = -((x - m->center_x) - 0.5 * lx);
    const double ty = -((y - m->center_y) - 0.5 * ly);
    return tx >= 0.0 && ty >= 0.0 && tx < lx && ty < ly;
}

static double cell_center(double c, double len, double res, int idx) {
    /* getPositionFromIndex: mapPosition + (0.5*length - 0.5*res) + res * (-index) */
    return (c + (0.5 * len - 0.5 * res)) + res * (double)(-idx);
}

double oracle_map_distance(const pqp_distance_map *m, double x, double y) {
    if (!oracle_map_inside(m, x, y)) return 0.0;                                   /* Map.cpp:20 */
    const double res = m->resolution;
    const double lx = m->rows * res, ly = m->cols * res;
    /* getIndexFromPosition: index = (int)(-((position - 0.5*length - mapPosition) / res)) */
    const int i0 = (int)(-(((x - 0.5 * lx) - m->center_x) / res));
    const int j0 = (int)(-(((y - 0.5 * ly) - m->center_y) / res));
    /* atPositionLinearInterpolated: the four cell centres surrounding the position */
    const double px = cell_center(m->center_x, lx, res, i0);
    const double py = cell_center(m->center_y, ly, res, j0);
    int i_lo, i_hi, j_lo, j_hi;  /* lo/hi in x resp. y; a larger coordinate is a SMALLER index */
    if (x >= px) { i_lo = i0; i_hi = i0 - 1; } else { i_lo = i0 + 1; i_hi = i0; }
    if (y >= py) { j_lo = j0; j_hi = j0 - 1; } else { j_lo = j0 + 1; j_hi = j0; }
    const int ok = i_lo >= 0 && i_lo < m->rows && i_hi >= 0 && i_hi < m->rows &&
                   j_lo >= 0 && j_lo < m->cols && j_hi >= 0 && j_hi < m->cols;
    if (!ok) {
        /* INTER_LINEAR falls back to INTER_NEAREST (GridMap::atPosition) */
        if (i0 < 0 || i0 >= m->rows || j0 < 0 || j0 >= m->cols) return 0.0;
        return (double)m->distance[(size_t)i0 * m->cols + j0];
    }
    const double f0 = m->distance[(size_t)i_lo * m->cols + j_lo];
    const double f1 = m->distance[(size_t)i_hi * m->cols + j_lo];
    const double f2 = m->distance[(size_t)i_lo * m->cols + j_hi];
    const double f3 = m->distance[(size_t)i_hi * m->cols + j_hi];
    const double rx = (x - cell_center(m->center_x, lx, res, i_lo)) / res;
    const double ry = (y - cell_center(m->center_y, ly, res, j_lo)) / res;
    const double fx = 1.0 - rx, fy = 1.0 - ry;
    const float value = (float)(f0 * fx * fy + f1 * rx * fy + f2 * fx * ry + f3 * rx * ry);
    return (double)value;
}

/* ------------------------------------------------------------------------------------------- */
/* tk::spline (src/tools/spline.cpp) -- natural cubic spline, re-derived                        */
/* ------------------------------------------------------------------------------------------- */

/* set_points, spline.cpp:161-249 with the default boundary (second derivative 0 at both ends,
 * :146-159): unknowns b_i = f''(t_i)/2 from the tridiagonal continuity system
 *   h_{i-1}/3 b_{i-1} + 2(h_{i-1}+h_i)/3 b_i + h_i/3 b_{i+1} = (y_{i+1}-y_i)/h_i - (y_i-y_{i-1})/h_{i-1},
 * then a_i = (b_{i+1}-b_i)/(3 h_i), c_i = (y_{i+1}-y_i)/h_i - (2 b_i + b_{i+1}) h_i / 3.
 * The last knot carries the right-extrapolation polynomial: a = 0, b = b_{n-1} (= 0 here),
 * c = f'_{n-2}(t_{n-1}).  coef[i] = {a, b, c, y}. */
int oracle_spline_fit(int n, const double *t, const double *y, double *coef) {
    if (n < 3) return -1;
    double *diag = (double *)malloc(sizeof(double) * (size_t)n * 3);
    double *rhs = diag + n, *b = diag + 2 * n;
    /* Thomas algorithm; row 0 and n-1 are  2 b = 0 */
    diag[0] = 2.0; rhs[0] = 0.0;
    double upper_prev = 0.0;   /* A(0,1) = 0 */
    for (int i = 1; i < n - 1; ++i) {
        const double hl = t[i] - t[i - 1], hr = t[i + 1] - t[i];
        if (!(hl > 0.0) || !(hr > 0.0)) { free(diag); return -1; }
        const double lo = hl / 3.0, di = 2.0 * (t[i + 1] - t[i - 1]) / 3.0, up = hr / 3.0;
        const double r = (y[i + 1] - y[i]) / hr - (y[i] - y[i - 1]) / hl;
        const double w = lo / diag[i - 1];
        diag[i] = di - w * upper_prev;
        rhs[i] = r - w * rhs[i - 1];
        upper_prev = up;
    }
    /* last row: 2 b_{n-1} = 0, no coupling to b_{n-2} */
    b[n - 1] = 0.0;
    for (int i = n - 2; i >= 1; --i) {
        const double up = (t[i + 1] - t[i]) / 3.0;
        b[i] = (rhs[i] - up * b[i + 1]) / diag[i];
    }
    b[0] = 0.0;
    for (int i = 0; i < n - 1; ++i) {
        const double h = t[i + 1] - t[i];
        coef[4 * i + 0] = (b[i + 1] - b[i]) / (3.0 * h);
        coef[4 * i + 1] = b[i];
        coef[4 * i + 2] = (y[i + 1] - y[i]) / h - (2.0 * b[i] + b[i + 1]) * h / 3.0;
        coef[4 * i + 3] = y[i];
    }
    {
        const int i = n - 2;
        const double h = t[n - 1] - t[n - 2];
        coef[4 * (n - 1) + 0] = 0.0;
        coef[4 * (n - 1) + 1] = b[n - 1];
        coef[4 * (n - 1) + 2] = 3.0 * coef[4 * i] * h * h + 2.0 * coef[4 * i + 1] * h + coef[4 * i + 2];
        coef[4 * (n - 1) + 3] = y[n - 1];
    }
    free(diag);
    return 0;
}

/* operator() and deriv, spline.cpp:250-318.  idx = max(lower_bound(t, at) - 1, 0); left of the
 * first knot the quadratic (b0 = b[0], c0 = c[0]) is used, right of the last one the last knot's
 * (b, c); inside, the cubic of segment idx. */
double oracle_spline_eval(int n, const double *t, const double *coef, int order, double at) {
    int lo = 0, hi = n;                      /* lower_bound: first index with t[idx] >= at */
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (t[mid] < at) lo = mid + 1; else hi = mid;
    }
    int idx = lo - 1;
    if (idx < 0) idx = 0;
    const double h = at - t[idx];
    const double *c = coef + 4 * idx;
    if (at < t[0]) {
        const double b0 = coef[1], c0 = coef[2];
        if (order == 0) return (b0 * h + c0) * h + coef[3];
        if (order == 1) return 2.0 * b0 * h + c0;
        if (order == 2) return 2.0 * b0 * h;      /* sic, spline.cpp:288 */
        return 0.0;
    }
    if (at > t[n - 1]) {
        const double *e = coef + 4 * (n - 1);
        if (order == 0) return (e[1] * h + e[2]) * h + e[3];
        if (order == 1) return 2.0 * e[1] * h + e[2];
        if (order == 2) return 2.0 * e[1];
        return 0.0;
    }
    if (order == 0) return ((c[0] * h + c[1]) * h + c[2]) * h + c[3];
    if (order == 1) return (3.0 * c[0] * h + 2.0 * c[1]) * h + c[2];
    if (order == 2) return 6.0 * c[0] * h + 2.0 * c[1];
    return 0.0;
}

/* ------------------------------------------------------------------------------------------- */
/* Clearance bounds, src/data_struct/reference_path_impl.cpp                                    */
/* ------------------------------------------------------------------------------------------- */

static double wrap_angle(double a) {             /* constraintAngle, tools.hpp:24-35 */
    while (a > M_PI) a -= 2 * M_PI;
    while (a < -M_PI) a += 2 * M_PI;
    return a;
}

/* getClearanceWithDirectionStrict, reference_path_impl.cpp:283-472.  out = {left, right}.
 * The branch at :322-389 (direction from the ORIGINAL spline) needs
 * FLAGS_enable_simple_boundary_decision = false (default true, planning_flags.cpp:84) and is not
 * restated. */
void oracle_clearance_strict(const pqp_params *prm, const pqp_distance_map *map, double sx,
                             double sy, double sz, double out[2]) {
    double left_bound = 0, right_bound = 0;
    const double delta_s = 0.5;
    const double left_angle = wrap_angle(sz + M_PI_2);
    const double right_angle = wrap_angle(sz - M_PI_2);
    const int n = (int)(5.0 / delta_s);
    const double radius = prm->circle_radius;
    const double original_clearance = oracle_map_distance(map, sx, sy);
    if (original_clearance > radius) {
        /* :296-321 free: march out on both sides until a sample is closer than the radius */
        double right_s = 0;
        for (int j = 0; j != n; ++j) {
            right_s += delta_s;
            const double x = sx + right_s * cos(right_angle), y = sy + right_s * sin(right_angle);
            if (oracle_map_distance(map, x, y) < radius) break;
        }
        double left_s = 0;
        for (int j = 0; j != n; ++j) {
            left_s += delta_s;
            const double x = sx + left_s * cos(left_angle), y = sy + left_s * sin(left_angle);
            if (oracle_map_distance(map, x, y) < radius) break;
        }
        right_bound = -(right_s - delta_s);
        left_bound = left_s - delta_s;
    } else {
        /* :390-441 in collision: find the nearer free side, then that side's far edge */
        double right_s = 0;
        for (int j = 0; j != n; ++j) {
            right_s += delta_s;
            const double x = sx + right_s * cos(right_angle), y = sy + right_s * sin(right_angle);
            if (oracle_map_distance(map, x, y) > radius) break;
        }
        double left_s = 0;
        for (int j = 0; j != n; ++j) {
            left_s += delta_s;
            const double x = sx + left_s * cos(left_angle), y = sy + left_s * sin(left_angle);
            if (oracle_map_distance(map, x, y) > radius) break;
        }
        if (left_s < right_s) {
            right_bound = left_s;
            for (int j = 0; j != n; ++j) {
                left_s += delta_s;
                const double x = sx + left_s * cos(left_angle), y = sy + left_s * sin(left_angle);
                if (oracle_map_distance(map, x, y) < radius) break;
            }
            left_bound = left_s - delta_s;
        } else {
            left_bound = -right_s;
            for (int j = 0; j != n; ++j) {
                right_s += delta_s;
                const double x = sx + right_s * cos(right_angle), y = sy + right_s * sin(right_angle);
                if (oracle_map_distance(map, x, y) < radius) break;
            }
            right_bound = -(right_s - delta_s);
        }
    }
    /* :442-465 refine both edges outwards in 0.1 m steps */
    const double smaller_ds = 0.1;
    const int fine = (int)(delta_s / smaller_ds);
    for (int i = 1; i != fine; ++i) {
        left_bound += smaller_ds;
        const double x = sx + left_bound * cos(left_angle), y = sy + left_bound * sin(left_angle);
        if (oracle_map_distance(map, x, y) < radius) { left_bound -= smaller_ds; break; }
    }
    for (int i = 1; i != fine; ++i) {
        right_bound -= smaller_ds;
        const double x = sx + right_bound * cos(right_angle), y = sy + right_bound * sin(right_angle);
        if (oracle_map_distance(map, x, y) < radius) { right_bound += smaller_ds; break; }
    }
    out[0] = left_bound;
    out[1] = right_bound;
}

/* getApproxState, reference_path_impl.cpp:120-140 */
static void approx_state(int nk, const double *knots, const double *xc, const double *yc,
                         const pqp_state *orig, double ax, double ay, double len, double out[3]) {
    const double x = oracle_spline_eval(nk, knots, xc, 0, orig->s + len);
    const double y = oracle_spline_eval(nk, knots, yc, 0, orig->s + len);
    const double v1x = ax - orig->x, v1y = ay - orig->y;
    const double v2x = x - orig->x, v2y = y - orig->y;
    const double proj = (v1x * v2x + v1y * v2y) / fmax(0.001, sqrt(v1x * v1x + v1y * v1y));
    const double move_dis = fabs(len) - proj;
    const int sign = len >= 0 ? 1 : -1;
    out[0] = x + sign * move_dis * cos(orig->z);
    out[1] = y + sign * move_dis * sin(orig->z);
    out[2] = orig->z;
}

/* updateBoundsImproved (:142-201) / updateBounds (:237-281).  Returns the number of leading
 * stations with bounds (the size bounds_ ends up with). */
int oracle_update_bounds(const pqp_params *prm, const pqp_distance_map *map, int mode, int n,
                         const pqp_state *ref, int n_knots, const double *knots,
                         const double *x_coef, const double *y_coef, pqp_station_bounds *out) {
    const double d[4] = {prm->d1, prm->d2, prm->d3, prm->d4};
    const double eps = 1e-6;                       /* FLAGS_epsilon, planning_flags.cpp:135 */
    for (int i = 0; i < n; ++i) {
        const pqp_state *st = &ref[i];
        double cl[4][2];
        int blocked = 0;
        for (int j = 0; j < 4; ++j) {
            const double cx = st->x + d[j] * cos(st->z), cy = st->y + d[j] * sin(st->z);
            if (mode == PQP_BOUNDS_IMPROVED) {
                double cc[3];
                approx_state(n_knots, knots, x_coef, y_coef, st, cx, cy, d[j], cc);
                oracle_clearance_strict(prm, map, cc[0], cc[1], cc[2], cl[j]);
                /* global2Local(c_j, c_jj).y, tools.cpp:61-68 */
                const double dx = cc[0] - cx, dy = cc[1] - cy;
                const double offset = -dx * sin(st->z) + dy * cos(st->z);
                cl[j][0] += offset;
                cl[j][1] += offset;
                if (fabs(cl[j][0] - cl[j][1]) < eps) blocked = 1;       /* isEqual, tools.cpp:30-32 */
            } else {
                oracle_clearance_strict(prm, map, cx, cy, st->z, cl[j]);
                if (cl[j][0] == cl[j][1]) blocked = 1;
            }
        }
        if (blocked) return i;
        out[i].c0_ub = cl[0][0]; out[i].c0_lb = cl[0][1];
        out[i].c1_ub = cl[1][0]; out[i].c1_lb = cl[1][1];
        out[i].c2_ub = cl[2][0]; out[i].c2_lb = cl[2][1];
        out[i].c3_ub = cl[3][0]; out[i].c3_lb = cl[3][1];
    }
    return n;
}

/* ------------------------------------------------------------------------------------------- */
/* Collision check, src/tools/collision_checker.cpp + car_geometry.cpp                          */
/* ------------------------------------------------------------------------------------------- */

/* CarGeometry::setCircles, car_geometry.cpp:38-57, with the constructor arguments of
 * collision_checker.cpp:9-15.  c[0] = bounding circle, c[1..6] = rr, rl, fr, fl, fm, rm; each
 * {x, y, r} in the vehicle frame. */
void oracle_car_circles(const pqp_params *prm, double c[7][3]) {
    const double width = prm->car_width;
    const double back = prm->car_length / 2.0 - prm->rear_axle_to_center;
    const double front = prm->car_length / 2.0 + prm->rear_axle_to_center;
    const double length = front + back;
    c[0][0] = (front - back) / 2.0; c[0][1] = 0;
    c[0][2] = sqrt(pow(length / 2, 2) + pow(width / 2, 2));
    const double shift = width / 4.0;
    const double small_r = sqrt(2 * pow(shift, 2));
    const double large_r = sqrt(pow(width, 2) + pow((length - width) / 2.0, 2)) / 2;
    c[1][0] = -back + shift;  c[1][1] = -width / 2.0 + shift; c[1][2] = small_r;   /* rr */
    c[2][0] = -back + shift;  c[2][1] = width / 2.0 - shift;  c[2][2] = small_r;   /* rl */
    c[3][0] = front - shift;  c[3][1] = -width / 2.0 + shift; c[3][2] = small_r;   /* fr */
    c[4][0] = front - shift;  c[4][1] = width / 2.0 - shift;  c[4][2] = small_r;   /* fl */
    c[5][0] = c[0][0] + (length - width) / 4; c[5][1] = 0; c[5][2] = large_r;      /* fm */
    c[6][0] = c[0][0] - (length - width) / 4; c[6][1] = 0; c[6][2] = large_r;      /* rm */
}

/* isSingleStateCollisionFreeImproved (:41-59) -> isSingleStateCollisionFree (:17-39) */
int oracle_state_collision_free(const pqp_params *prm, const pqp_distance_map *map,
                                const pqp_state *s) {
    double c[7][3];
    oracle_car_circles(prm, c);
    const double cz = cos(s->z), sz = sin(s->z);
    /* local2Global, tools.cpp:54-59 */
    const double bx = c[0][0] * cz - c[0][1] * sz + s->x;
    const double by = c[0][0] * sz + c[0][1] * cz + s->y;
    if (!oracle_map_inside(map, bx, by)) return 0;
    if (!(oracle_map_distance(map, bx, by) < c[0][2])) return 1;
    for (int k = 1; k < 7; ++k) {
        const double x = c[k][0] * cz - c[k][1] * sz + s->x;
        const double y = c[k][0] * sz + c[k][1] * cz + s->y;
        if (!oracle_map_inside(map, x, y)) return 0;
        if (oracle_map_distance(map, x, y) < c[k][2]) return 0;
    }
    return 1;
}

/* Raw tail of optimizePath, path_optimizer.cpp:191-202.  In place; *n_kept = states left.
 * Returns the reference's bool (0 when the very first state collides: the reference would read
 * back() of an empty vector there). */
int oracle_finish_raw(const pqp_params *prm, const pqp_distance_map *map, int n, pqp_state *path,
                      int collision_check, int *n_kept) {
    double s = 0;
    for (int i = 0; i < n; ++i) {
        if (i != 0) {
            const double dx = path[i - 1].x - path[i].x, dy = path[i - 1].y - path[i].y;
            s += sqrt(dx * dx + dy * dy);                                 /* distance(), tools.cpp:50-52 */
        }
        path[i].s = s;
        if (collision_check && !oracle_state_collision_free(prm, map, &path[i])) {
            *n_kept = i;
            return i > 0 ? path[i - 1].s >= 20 : 0;
        }
    }
    *n_kept = n;
    return 1;
}

/* Densifying tail, path_optimizer.cpp:203-230.  out has room for max_out states.  The reference
 * reads final_path->back() when the FIRST sample collides (empty vector); 0 is returned there. */
int oracle_densify(const pqp_params *prm, const pqp_distance_map *map, int n, const pqp_state *path,
                   double spacing, int collision_check, int max_out, pqp_state *out, int *n_out) {
    double *t = (double *)malloc(sizeof(double) * (size_t)n * 11);
    double *xs = t + n, *ys = t + 2 * n, *xc = t + 3 * n, *yc = t + 7 * n;
    for (int i = 0; i < n; ++i) { t[i] = path[i].s; xs[i] = path[i].x; ys[i] = path[i].y; }
    *n_out = 0;
    if (oracle_spline_fit(n, t, xs, xc) || oracle_spline_fit(n, t, ys, yc)) { free(t); return 0; }
    int ok = 1;
    for (int i = 0; i * spacing <= t[n - 1]; ++i) {
        const double ts = i * spacing;
        pqp_state st;
        memset(&st, 0, sizeof(st));
        st.x = oracle_spline_eval(n, t, xc, 0, ts);
        st.y = oracle_spline_eval(n, t, yc, 0, ts);
        const double x1 = oracle_spline_eval(n, t, xc, 1, ts), y1 = oracle_spline_eval(n, t, yc, 1, ts);
        const double x2 = oracle_spline_eval(n, t, xc, 2, ts), y2 = oracle_spline_eval(n, t, yc, 2, ts);
        st.z = atan2(y1, x1);                                              /* getHeading, tools.cpp:34-38 */
        st.k = (x1 * y2 - y1 * x2) / pow(pow(x1, 2) + pow(y1, 2), 1.5);    /* getCurvature, :40-46 */
        st.s = ts;
        if (collision_check && !oracle_state_collision_free(prm, map, &st)) {
            ok = *n_out > 0 ? out[*n_out - 1].s >= 20 : 0;
            break;
        }
        if (*n_out >= max_out) { ok = 0; break; }
        out[(*n_out)++] = st;
    }
    free(t);
    return ok;
}

/* PathOptimizer::solveWithoutSmoothing, path_optimizer.cpp:87-117 (+ optimizePath :180-231) for one
 * path: bounds -> (trim to the unblocked prefix) -> QP -> tail.  Returns the reference's bool. */
int oracle_plan_path(const pqp_params *prm, const pqp_distance_map *map, int formulation,
                     int bounds_mode, int output_mode, int n, const pqp_state *ref, int n_knots,
                     const double *knots, const double *x_coef, const double *y_coef,
                     const double x0[3], double end_heading, double spacing, int collision_check,
                     int max_out, pqp_state *out, int *n_out, int *status, int *iters,
                     pqp_station_bounds *bounds_out) {
    pqp_station_bounds *bounds = (pqp_station_bounds *)malloc(sizeof(pqp_station_bounds) * (size_t)(n > 0 ? n : 1));
    const int nv = oracle_update_bounds(prm, map, bounds_mode, n, ref, n_knots, knots, x_coef, y_coef, bounds);
    if (bounds_out) memcpy(bounds_out, bounds, sizeof(pqp_station_bounds) * (size_t)nv);
    *n_out = 0;
    int ok = 0;
    oqp_info info;
    memset(&info, 0, sizeof(info));
    info.status = PQP_INVALID_PROBLEM;
    if (nv >= 2) {
        pqp_state *sol = (pqp_state *)malloc(sizeof(pqp_state) * (size_t)nv);
        /* reference_path_->updateLimits() (path_optimizer.cpp:102): only KPC uses the limits */
        double *mk = NULL, *mkp = NULL;
        if (formulation == PQP_FORM_KPC) {
            mk = (double *)malloc(sizeof(double) * (size_t)nv);
            mkp = (double *)malloc(sizeof(double) * (size_t)nv);
            oracle_update_limits(prm, 0, nv, ref, mk, mkp);
        }
        oracle_solve_path(prm, formulation, nv, ref, bounds, x0, end_heading, mk, mkp, sol, NULL, &info);
        free(mk); free(mkp);
        if (info.status == PQP_SOLVED) {
            if (output_mode == PQP_OUTPUT_RAW) {
                ok = oracle_finish_raw(prm, map, nv, sol, collision_check, n_out);
                memcpy(out, sol, sizeof(pqp_state) * (size_t)*n_out);
            } else {
                ok = oracle_densify(prm, map, nv, sol, spacing, collision_check, max_out, out, n_out);
            }
        }
        free(sol);
    }
    if (status) *status = info.status;
    if (iters) *iters = info.iters;
    free(bounds);
    return ok;
}

/* ReferencePathImpl::updateLimits, reference_path_impl.cpp:203-235: curvature and curvature-rate limits of the KPC
 * formulation from the speed profile (v, a) of the reference states; from_spline = the use_spline_ branch (:214-221). */
void oracle_update_limits(const pqp_params *prm, int from_spline, int n, const pqp_state *ref, double *max_k, double *max_kp) {
    for (int i = 0; i < n; ++i) {
        if (from_spline) {
            max_k[i] = tan(prm->max_steering_angle) / prm->wheel_base;
            max_kp[i] = 1.7976931348623157e308; /* DBL_MAX */
            continue;
        }
        /* Friction circle limit. */
        double ref_v = ref[i].v;
        double ref_ax = ref[i].a;
        double ay_allowed = sqrt(pow(prm->mu * 9.8, 2) - pow(ref_ax, 2));
        if (ref_v > 0.0001) max_k[i] = ay_allowed / pow(ref_v, 2);
        else max_k[i] = 1.7976931348623157e308;
        /* Control rate limit. */
        if (ref_v > 0.0001) max_kp[i] = prm->max_curvature_rate / ref_v;
        else max_kp[i] = 1.7976931348623157e308;
    }
}
