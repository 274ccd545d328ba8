// host_driver.cpp -- small C++ driver over include/pqp_solver.hpp (needs a GPU).  It reads a batch
// from a flat binary file written by tests/test_gpu_parity.py, solves it through the C++ host mirror
// (BatchPathSolver and the single-path GpuOsqpSolver) and writes the result back, so the Python test
// can compare the C++ path with the oracle.
//   host_driver <in.bin> <out.bin>
// in : int32 B, int32 n[B], State ref[sumN], Bounds b[sumN], double veh[B][4]
// out: double frenet[sumN][3], int32 status[B], int32 iters[B], then for path 0 via GpuOsqpSolver:
//      int32 ok, State path0[n0]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/pqp_solver.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t B = 0;
    if (fread(&B, 4, 1, f) != 1) return 2;
    std::vector<int32_t> n(B);
    if (fread(n.data(), 4, B, f) != (size_t)B) return 2;
    size_t total = 0;
    for (int v : n) total += v;
    std::vector<pqp::State> ref(total);
    std::vector<pqp::CoveringCircleBounds> bounds(total);
    std::vector<pqp::VehicleStateView> veh(B);
    if (fread(ref.data(), sizeof(pqp::State), total, f) != total) return 2;
    if (fread(bounds.data(), sizeof(pqp::CoveringCircleBounds), total, f) != total) return 2;
    if (fread(veh.data(), sizeof(pqp::VehicleStateView), B, f) != (size_t)B) return 2;
    fclose(f);

    if (pqp::BatchPathSolver::create("NOPE", 1, 10)) return 3;  // unknown type -> nullptr
    auto solver = pqp::BatchPathSolver::create("KP", B, (int)total);
    if (!solver) return 4;
    pqp::BatchPathSolver::Result r;
    if (!solver->solve(n, ref, bounds, veh, &r)) return 5;

    std::vector<pqp::State> ref0(ref.begin(), ref.begin() + n[0]);
    std::vector<pqp::CoveringCircleBounds> b0(bounds.begin(), bounds.begin() + n[0]);
    auto single = pqp::GpuOsqpSolver::create("KP", ref0, b0, veh[0], (size_t)n[0]);
    if (!single) return 6;
    std::vector<pqp::State> path0;
    int32_t ok = single->solve(&path0) ? 1 : 0;
    path0.resize(n[0]);

    FILE *o = fopen(argv[2], "wb");
    if (!o) return 7;
    fwrite(r.frenet.data(), sizeof(double), 3 * total, o);
    fwrite(r.status.data(), 4, B, o);
    fwrite(r.iters.data(), 4, B, o);
    fwrite(&ok, 4, 1, o);
    fwrite(path0.data(), sizeof(pqp::State), n[0], o);
    fclose(o);
    std::printf("host_driver: %d paths, %lld stations, kernel %.3f ms, solved %d\n", B, (long long)total,
                r.stats.kernel_ms, r.stats.n_solved);
    return 0;
}
