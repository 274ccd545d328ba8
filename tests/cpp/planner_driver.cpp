// planner_driver.cpp -- C++ driver over include/pqp_planner.hpp (needs a GPU): reads a distance map
// and a batch of reference lines written by tests/test_gpu_env.py, runs the batched
// solveWithoutSmoothing (improved bounds with splines fitted here, raw output) plus the single-path
// overload for path 0, and writes the results back for comparison with the oracle.
//   planner_driver <in.bin> <out.bin>
// in : int32 rows, cols; double res, cx, cy; float dist[rows*cols]; int32 B; int32 n[B]; State ref[sumN]; double veh[B][4]
// out: int32 n_out[B]; int32 ok[B]; int32 status[B]; State paths (concatenated, n_out each); int32 ok0; int32 n0; State path0[n0]
#include <cstdio>
#include <vector>

#include "../../include/pqp_planner.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    pqp::DistanceMap map;
    int32_t dims[2];
    double geo[3];
    if (fread(dims, 4, 2, f) != 2 || fread(geo, 8, 3, f) != 3) return 2;
    map.rows = dims[0]; map.cols = dims[1]; map.resolution = geo[0]; map.center_x = geo[1]; map.center_y = geo[2];
    map.distance.resize((size_t)map.rows * map.cols);
    if (fread(map.distance.data(), 4, map.distance.size(), f) != map.distance.size()) return 2;
    int32_t B = 0;
    if (fread(&B, 4, 1, f) != 1) return 2;
    std::vector<int32_t> n(B);
    if (fread(n.data(), 4, B, f) != (size_t)B) return 2;
    std::vector<std::vector<pqp::State>> refs(B);
    size_t total = 0;
    for (int b = 0; b < B; ++b) {
        refs[b].resize(n[b]);
        if (fread(refs[b].data(), sizeof(pqp::State), n[b], f) != (size_t)n[b]) return 2;
        total += n[b];
    }
    std::vector<pqp::VehicleStateView> veh(B);
    if (fread(veh.data(), sizeof(pqp::VehicleStateView), B, f) != (size_t)B) return 2;
    fclose(f);

    auto po = pqp::PathOptimizerGpu::create(map, B, (int)total);
    if (!po) return 4;
    // x_s_, y_s_: natural splines through the stations (what buildReferenceFromSpline samples from)
    std::vector<pqp::Spline> xs(B), ys(B);
    for (int b = 0; b < B; ++b) {
        std::vector<double> s, x, y;
        for (const auto &st : refs[b]) { s.push_back(st.s); x.push_back(st.x); y.push_back(st.y); }
        if (!xs[b].set_points(s, x) || !ys[b].set_points(s, y)) return 5;
    }
    po->bounds_mode = PQP_BOUNDS_IMPROVED;
    std::vector<std::vector<pqp::State>> paths;
    std::vector<char> ok;
    std::vector<int32_t> status;
    if (!po->solveWithoutSmoothing(refs, veh, &paths, &ok, &xs, &ys, &status)) return 6;

    po->bounds_mode = PQP_BOUNDS_SIMPLE;
    std::vector<pqp::State> path0;
    const int32_t ok0 = po->solveWithoutSmoothing(refs[0], veh[0], &path0) ? 1 : 0;
    if (po->getObstacleDistance(1e9, 1e9) != 0.0) return 7;   // outside the map

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 8;
    std::vector<int32_t> n_out(B), okv(B);
    for (int b = 0; b < B; ++b) { n_out[b] = (int32_t)paths[b].size(); okv[b] = ok[b]; }
    fwrite(n_out.data(), 4, B, o);
    fwrite(okv.data(), 4, B, o);
    fwrite(status.data(), 4, B, o);
    for (int b = 0; b < B; ++b) fwrite(paths[b].data(), sizeof(pqp::State), paths[b].size(), o);
    const int32_t n0 = (int32_t)path0.size();
    fwrite(&ok0, 4, 1, o);
    fwrite(&n0, 4, 1, o);
    fwrite(path0.data(), sizeof(pqp::State), path0.size(), o);
    fclose(o);
    int solved = 0;
    for (int b = 0; b < B; ++b) solved += ok[b];
    std::printf("planner_driver: %d reference lines, %lld stations, %d ok\n", B, (long long)total, solved);
    return 0;
}
