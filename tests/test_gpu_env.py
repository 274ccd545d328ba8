"""GPU parity tests of the stages either side of the QP (scope rows N2-N4), through the C ABI of
include/pqp_env.h, against the CPU oracle on the same seeded inputs.

Tolerances.  Map lookups, clearance bounds, collision flags, cuts and return flags are decisions on
float32 map values / a 0.1 m lattice: compared EXACTLY, with at most FLIP_BUDGET stations allowed
to differ where CUDA's sin/cos and glibc's differ in the last ulp and a sample lands within that
ulp of the circle radius (none observed).  Offsets added by the improved variant, re-accumulated s
and resampled states are floating point: 1e-9 absolute.  The chained planner iteration inherits the
QP bar (FRENET_TOL = 1e-8, identical status and iteration counts)."""
import numpy as np
import pytest

from oracle import oracle
from path_optimizer_b200 import planner, synth
from path_optimizer_b200.abi import SOLVED, STATE_DTYPE

pytestmark = pytest.mark.gpu

FP_TOL = 1e-9
FRENET_TOL = 1e-8
FLIP_BUDGET = 0
WILD = dict(y_range=(-3.0, 3.0), heading_range=0.05, curvature_amp=0.02)


@pytest.fixture(scope="module")
def field():
    return synth.disc_field_map()


@pytest.fixture(scope="module")
def pl(field):
    p = planner.PathPlanner(max_batch=2048, max_total_points=2048 * 200)
    p.set_map(field)
    yield p
    p.close()


def test_needs_a_map():
    p = planner.PathPlanner(max_batch=4, max_total_points=400)
    with pytest.raises(Exception, match="map"):
        p.map_distance([[0.0, 0.0]])
    p.close()


def test_map_distance(pl, field):
    rng = np.random.default_rng(7)
    xy = np.stack([rng.uniform(-112, 112, 20000), rng.uniform(-26, 26, 20000)], 1)
    # exact border / outside cases
    lx, ly = field["rows"] * field["resolution"], field["cols"] * field["resolution"]
    xy[:6] = [[lx / 2, 0], [-lx / 2, 0], [0, ly / 2], [0, -ly / 2], [lx / 2 - 0.05, ly / 2 - 0.05], [1e9, -1e9]]
    got = pl.map_distance(xy)
    want = oracle.map_distance(field, xy)
    assert (got == want).all()


@pytest.mark.parametrize("mode", [planner.BOUNDS_SIMPLE, planner.BOUNDS_IMPROVED])
def test_bounds_match_oracle(pl, field, mode):
    prm = oracle.default_params()
    n_points = np.r_[np.full(56, 150), [3, 4, 7, 50, 199, 64, 33, 90]].astype(np.int32)
    b = synth.map_reference_paths(64, 150, n_points=n_points, **WILD)
    spl = planner.reference_splines(b) if mode == planner.BOUNDS_IMPROVED else None
    want = oracle.update_bounds(prm, field, b, mode=mode, splines=spl)
    got = pl.update_bounds(b, mode=mode, splines=spl)
    assert (got["n_valid"] == want["n_valid"]).all()
    assert (want["n_valid"] < n_points).any() and (want["n_valid"] == n_points).any()
    W = want["bounds"].view(np.float64).reshape(-1, 8)
    G = got["bounds"].view(np.float64).reshape(-1, 8)
    flips = 0
    for i in range(len(n_points)):
        sl = slice(b["offsets"][i], b["offsets"][i] + want["n_valid"][i])
        d = np.abs(G[sl] - W[sl])
        flips += int((d > FP_TOL).any(axis=1).sum())
    assert flips <= FLIP_BUDGET
    assert got["stats"].kernel_launches == 1


def test_bounds_full_size_properties(pl, field):
    """Config-3 sized batch (2048 x 200 here): left >= right everywhere that is not blocked, bounds on
    the 0.1 m lattice for the simple variant, determinism, and a sampled oracle comparison."""
    b = synth.map_reference_paths(2048, 200)
    r1 = pl.update_bounds(b, mode=planner.BOUNDS_SIMPLE)
    r2 = pl.update_bounds(b, mode=planner.BOUNDS_SIMPLE)
    assert (r1["n_valid"] == r2["n_valid"]).all()
    assert r1["bounds"].tobytes() == r2["bounds"].tobytes()
    B8 = r1["bounds"].view(np.float64).reshape(-1, 8)
    valid = np.zeros(len(B8), bool)
    for i in range(2048):
        valid[b["offsets"][i]:b["offsets"][i] + r1["n_valid"][i]] = True
    V = B8[valid]
    assert (V[:, 0::2] > V[:, 1::2]).all()                      # ub > lb on every circle
    assert np.abs(V * 10 - np.round(V * 10)).max() < 1e-9       # 0.5 m + 0.1 m lattice
    assert np.abs(V).max() <= 9.9 + 1e-9
    prm = oracle.default_params()
    sub = synth.slice_batch(b, 100, 164)
    want = oracle.update_bounds(prm, field, sub, mode=planner.BOUNDS_SIMPLE)
    assert (want["n_valid"] == r1["n_valid"][100:164]).all()
    lo, hi = b["offsets"][100], b["offsets"][164]
    Wv = want["bounds"].view(np.float64).reshape(-1, 8)
    assert (Wv[valid[lo:hi]] == B8[lo:hi][valid[lo:hi]]).all()


def test_collision_check(pl, field):
    prm = oracle.default_params()
    rng = np.random.default_rng(11)
    st = np.zeros(20000, dtype=STATE_DTYPE)
    st["x"], st["y"], st["z"] = rng.uniform(-112, 112, 20000), rng.uniform(-26, 26, 20000), rng.uniform(-np.pi, np.pi, 20000)
    got = pl.check_states(st)
    want = oracle.check_states(prm, field, st)
    assert int((got != want).sum()) <= FLIP_BUDGET
    assert 0.05 < want.mean() < 0.95


def _solved_like_paths(B=64, n=150):
    tame = synth.map_reference_paths(B // 2, n)
    wild = synth.map_reference_paths(B - B // 2, n, first_path=1000, y_range=(-2.0, 2.0), heading_range=0.06,
                                     curvature_amp=0.01)
    paths = np.concatenate([tame["ref"], wild["ref"]])
    paths["s"] = 0.0
    return np.concatenate([tame["n_points"], wild["n_points"]]), paths


def test_raw_tail(pl, field):
    prm = oracle.default_params()
    n_points, paths = _solved_like_paths()
    for cc in (True, False):
        want = oracle.finish_raw(prm, field, n_points, paths, collision_check=cc)
        got = pl.finish_raw(n_points, paths, collision_check=cc)
        assert (got["n_kept"] == want["n_kept"]).all() and (got["ok"] == want["ok"]).all()
        off = np.concatenate([[0], np.cumsum(n_points)])
        for i in range(len(n_points)):
            sl = slice(off[i], off[i] + want["n_kept"][i])
            assert np.abs(got["states"]["s"][sl] - want["states"]["s"][sl]).max(initial=0.0) <= FP_TOL
            assert (got["states"]["x"][sl] == paths["x"][sl]).all()
    assert (want["n_kept"] == n_points).all()


def test_densify_tail(pl, field):
    prm = oracle.default_params()
    n_points, paths = _solved_like_paths()
    src = oracle.finish_raw(prm, field, n_points, paths, collision_check=False)["states"]
    for spacing, cc, max_out in ((0.3, True, 200), (0.25, False, 400), (0.3, False, 20), (1.0, True, 64)):
        want = oracle.densify(prm, field, n_points, src, spacing, cc, max_out)
        got = pl.densify(n_points, src, spacing, cc, max_out)
        assert (got["n_out"] == want["n_out"]).all(), (spacing, cc, max_out)
        assert (got["ok"] == want["ok"]).all()
        for i in range(len(n_points)):
            k = want["n_out"][i]
            for f in ("x", "y", "z", "k", "s"):
                assert np.abs(got["states"][f][i, :k] - want["states"][f][i, :k]).max(initial=0.0) <= FP_TOL


@pytest.mark.parametrize("bounds_mode", [planner.BOUNDS_SIMPLE, planner.BOUNDS_IMPROVED])
@pytest.mark.parametrize("output_mode", [planner.OUTPUT_RAW, planner.OUTPUT_DENSIFY])
def test_plan_chain(pl, field, bounds_mode, output_mode):
    """solveWithoutSmoothing for a batch: bounds -> QP -> tail on the device vs the oracle chain."""
    prm = oracle.default_params()
    tame = synth.map_reference_paths(40, 120)
    wild = synth.map_reference_paths(24, 120, first_path=500, y_range=(-1.5, 1.5), heading_range=0.03, curvature_amp=0.01)
    b = dict(n_points=np.concatenate([tame["n_points"], wild["n_points"]]),
             ref=np.concatenate([tame["ref"], wild["ref"]]), x0=np.concatenate([tame["x0"], wild["x0"]]),
             end_heading=np.concatenate([tame["end_heading"], wild["end_heading"]]))
    b["offsets"] = np.concatenate([[0], np.cumsum(b["n_points"])]).astype(np.int32)
    spl = planner.reference_splines(b) if bounds_mode == planner.BOUNDS_IMPROVED else None
    want = oracle.plan(prm, field, b, bounds_mode=bounds_mode, splines=spl, output_mode=output_mode, max_out=256)
    got = pl.plan(b, bounds_mode=bounds_mode, splines=spl, output_mode=output_mode, max_out=256, want_bounds=True)
    assert (got["status"] == want["status"]).all()
    assert (got["iters"] == want["iters"]).all()
    assert (got["n_out"] == want["n_out"]).all()
    assert (got["ok"] == want["ok"]).all()
    assert (want["status"] == SOLVED).sum() >= 20 and (want["status"] != SOLVED).any()
    for i in range(len(b["n_points"])):
        k = want["n_out"][i]
        if output_mode == planner.OUTPUT_RAW:
            lo = b["offsets"][i]
            g, w = got["states"][lo:lo + k], want["states"][lo:lo + k]
        else:
            g, w = got["states"][i, :k], want["states"][i, :k]
        for f in ("x", "y", "z", "k", "s"):
            assert np.abs(g[f] - w[f]).max(initial=0.0) <= FRENET_TOL
    assert got["stats"].kernel_launches == 3


def test_plan_full_size(pl, field):
    """2048 x 200 planner iterations in one call: determinism + sampled oracle agreement."""
    b = synth.map_reference_paths(2048, 200)
    r1 = pl.plan(b)
    r2 = pl.plan(b)
    assert (r1["status"] == r2["status"]).all() and (r1["iters"] == r2["iters"]).all()
    assert r1["states"].tobytes() == r2["states"].tobytes()
    assert (r1["status"] == SOLVED).mean() > 0.7
    prm = oracle.default_params()
    sub = synth.slice_batch(b, 300, 332)
    want = oracle.plan(prm, field, sub)
    assert (want["status"] == r1["status"][300:332]).all()
    assert (want["iters"] == r1["iters"][300:332]).all()
    assert (want["n_out"] == r1["n_out"][300:332]).all() and (want["ok"] == r1["ok"][300:332]).all()
    lo = b["offsets"][300]
    for i in range(32):
        k = want["n_out"][i]
        o = sub["offsets"][i]
        for f in ("x", "y", "z", "k", "s"):
            assert np.abs(r1["states"][f][lo + o:lo + o + k] - want["states"][f][o:o + k]).max(initial=0.0) <= FRENET_TOL
    print(f"plan 2048x200: h2d {r1['stats'].h2d_ms:.3f} ms, kernels {r1['stats'].kernel_ms:.3f} ms, "
          f"d2h {r1['stats'].d2h_ms:.3f} ms, solved {int((r1['status'] == SOLVED).sum())}")


def test_cpp_planner_mirror(field, tmp_path):
    """The C++ host side of the chain (include/pqp_planner.hpp: PathOptimizerGpu::solveWithoutSmoothing batched and
    with the reference's single-path signature, pqp::Spline = tk::spline's interface) through a compiled driver."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "planner_driver")
    src = exe + ".cpp"
    if not os.path.exists(exe) or os.path.getmtime(src) > os.path.getmtime(exe):
        subprocess.run(["g++", "-O2", "-std=c++17", src, "-o", exe, "-L" + os.path.join(root, "path_optimizer_b200"),
                        "-lpqp", "-Wl,-rpath," + os.path.join(root, "path_optimizer_b200")], check=True)
    prm = oracle.default_params()
    b = synth.map_reference_paths(12, 90, y_range=(-1.0, 1.0), heading_range=0.03, curvature_amp=0.008)
    B = 12
    veh = np.column_stack([b["x0"], b["end_heading"]])
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.array([field["rows"], field["cols"]], dtype=np.int32).tobytes())
        f.write(np.array([field["resolution"], field["center_x"], field["center_y"]], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(field["distance"], dtype=np.float32).tobytes())
        f.write(np.int32(B).tobytes()); f.write(b["n_points"].tobytes()); f.write(b["ref"].tobytes())
        f.write(np.ascontiguousarray(veh).tobytes())
    subprocess.run([exe, str(fin), str(fout)], check=True)
    raw = open(fout, "rb").read()
    n_out = np.frombuffer(raw, dtype=np.int32, count=B)
    ok = np.frombuffer(raw, dtype=np.int32, count=B, offset=4 * B)
    status = np.frombuffer(raw, dtype=np.int32, count=B, offset=8 * B)
    o = 12 * B
    spl = planner.reference_splines(b)
    want = oracle.plan(prm, field, b, bounds_mode=planner.BOUNDS_IMPROVED, splines=spl)
    assert (n_out == want["n_out"]).all() and (ok == want["ok"]).all() and (status == want["status"]).all()
    for i in range(B):
        p = np.frombuffer(raw, dtype=STATE_DTYPE, count=n_out[i], offset=o)
        o += STATE_DTYPE.itemsize * int(n_out[i])
        lo = b["offsets"][i]
        for f_ in ("x", "y", "z", "k", "s"):
            assert np.abs(p[f_] - want["states"][f_][lo:lo + n_out[i]]).max(initial=0.0) <= FRENET_TOL
    ok0, n0 = np.frombuffer(raw, dtype=np.int32, count=2, offset=o)
    one = oracle.plan(prm, field, synth.slice_batch(b, 0, 1), bounds_mode=planner.BOUNDS_SIMPLE)
    assert ok0 == one["ok"][0] and n0 == one["n_out"][0]
    p0 = np.frombuffer(raw, dtype=STATE_DTYPE, count=n0, offset=o + 8)
    assert np.abs(p0["x"] - one["states"]["x"][:n0]).max(initial=0.0) <= FRENET_TOL


def test_golden_env_fixture(tmp_path):
    """The committed fixture tests/golden/env_small.npz (oracle outputs, generator committed) without the oracle in
    the loop: lookups, both bounds variants, collision flags and all four chain variants."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "env_small.npz"))
    field = dict(distance=g["map_distance"], rows=g["map_distance"].shape[0], cols=g["map_distance"].shape[1],
                 resolution=float(g["map_geo"][0]), center_x=float(g["map_geo"][1]), center_y=float(g["map_geo"][2]))
    b = dict(n_points=g["n_points"], ref=g["ref"], x0=g["x0"], end_heading=g["end_heading"])
    b["offsets"] = np.concatenate([[0], np.cumsum(b["n_points"])]).astype(np.int32)
    spl = dict(n_knots=g["n_points"], knots=g["knots"], x_coef=g["x_coef"], y_coef=g["y_coef"])
    p = planner.PathPlanner(max_batch=16, max_total_points=16 * 128)
    p.set_map(field)
    assert (p.map_distance(g["xy"]) == g["xy_distance"]).all()
    assert (p.check_states(b["ref"]) == g["collision_free"]).all()
    for mode, tag in ((planner.BOUNDS_SIMPLE, "simple"), (planner.BOUNDS_IMPROVED, "improved")):
        sp = spl if mode == planner.BOUNDS_IMPROVED else None
        r = p.update_bounds(b, mode=mode, splines=sp)
        assert (r["n_valid"] == g[f"n_valid_{tag}"]).all()
        G, W = r["bounds"].view(np.float64).reshape(-1, 8), g[f"bounds_{tag}"].view(np.float64).reshape(-1, 8)
        assert np.abs(G - W).max() <= FP_TOL
        for om, otag in ((planner.OUTPUT_RAW, "raw"), (planner.OUTPUT_DENSIFY, "dense")):
            r = p.plan(b, bounds_mode=mode, splines=sp, output_mode=om, max_out=128)
            k = f"plan_{tag}_{otag}_"
            assert (r["status"] == g[k + "status"]).all() and (r["iters"] == g[k + "iters"]).all()
            assert (r["n_out"] == g[k + "n_out"]).all() and (r["ok"] == g[k + "ok"]).all()
            for i in range(len(b["n_points"])):
                n = r["n_out"][i]
                if om == planner.OUTPUT_RAW:
                    lo = b["offsets"][i]
                    a, w = r["states"][lo:lo + n], g[k + "states"][lo:lo + n]
                else:
                    a, w = r["states"][i, :n], g[k + "states"][i, :n]
                for f in ("x", "y", "z", "k", "s"):
                    assert np.abs(a[f] - w[f]).max(initial=0.0) <= FRENET_TOL
    p.close()


def test_map_survives_generic_solves(field):
    """Regression: a K / KPC solve grows the handle's generic staging buffers; the uploaded map and the scratch of the
    map-based stages belong to the same handle and must be untouched by that."""
    prm = oracle.default_params()
    p = planner.PathPlanner(max_batch=64, max_total_points=64 * 100)
    p.set_map(field)
    xy = np.array([[0.0, 0.0], [-20.0, 1.0], [35.0, -3.0]])
    before = p.map_distance(xy)
    b = synth.curvy_corridors(4, 40)
    assert (p.solve(b, "K")["status"] == oracle.solve_batch(prm, 1, b)["status"]).all()
    b2 = synth.curvy_corridors(48, 100)              # larger batch: the staging buffers grow again
    p.solve(b2, "K")
    assert (p.map_distance(xy) == before).all()
    assert (before == oracle.map_distance(field, xy)).all()
    r = p.plan(synth.map_reference_paths(4, 60))
    assert len(r["status"]) == 4
    p.close()


def test_config1_reference_benchmark_case():
    """BASELINE config 1 (the reference's own benchmark inputs, tests/golden/config1_benchmark_map.npz): bounds on the
    reference's obstacle map -> KP QP -> collision-checked output through pqp_plan_batch, against the committed
    oracle outputs."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_benchmark_map.npz"))
    field = dict(distance=g["map_distance"], rows=int(g["image_shape"][0]), cols=int(g["image_shape"][1]),
                 resolution=float(g["map_geo"][0]), center_x=float(g["map_geo"][1]), center_y=float(g["map_geo"][2]))
    b = dict(n_points=g["n_points"], ref=g["ref"], x0=g["x0"], end_heading=g["end_heading"])
    b["offsets"] = np.array([0, int(g["n_points"][0])], dtype=np.int32)
    spl = dict(n_knots=np.array([len(g["knots"])], dtype=np.int32), knots=g["knots"], x_coef=g["x_coef"], y_coef=g["y_coef"])
    p = planner.PathPlanner(max_batch=1, max_total_points=256)
    p.set_map(field)
    for mode, tag in ((planner.BOUNDS_IMPROVED, "improved"), (planner.BOUNDS_SIMPLE, "simple")):
        sp = spl if mode == planner.BOUNDS_IMPROVED else None
        r = p.plan(b, bounds_mode=mode, splines=sp, want_bounds=True)
        assert r["status"][0] == g[f"plan_{tag}_status"][0] == SOLVED and r["iters"][0] == g[f"plan_{tag}_iters"][0]
        assert r["n_out"][0] == g[f"plan_{tag}_n_out"][0] and r["ok"][0] == 1
        assert np.abs(r["bounds"].view(np.float64) - g[f"bounds_{tag}"].view(np.float64)).max() <= FP_TOL
        for f in ("x", "y", "z", "k", "s"):
            assert np.abs(r["states"][f] - g[f"plan_{tag}_states"][f]).max() <= FRENET_TOL
    p.close()
