"""CPU tests of the stages either side of the QP (scope rows N2-N4): the oracle restatement
(oracle/pqp_oracle_env.c) against independent checks, and the product's per-thread device source
(pqp_env_core.cuh, compiled for the host by tests/emu/env_emu.cpp) against the oracle.

Tolerances: bounds / cuts / flags are decisions on a 0.1 m lattice -> compared exactly; spline
coefficients and resampled states are floating point -> 1e-9 absolute (stated per assert)."""
import numpy as np
import pytest
from scipy.interpolate import CubicSpline

from oracle import oracle
from path_optimizer_b200 import planner, synth
from path_optimizer_b200.abi import BOUNDS_DTYPE, STATE_DTYPE
from tests.emu import emu


# reference lines that leave the obstacle-free corridor: blocked stations, infeasible QPs, collisions
WILD = dict(y_range=(-3.0, 3.0), heading_range=0.05, curvature_amp=0.02)


@pytest.fixture(scope="module")
def field():
    return synth.disc_field_map()


def _numpy_bilinear(m, x, y):
    """Independent statement of the lookup from the geometry alone: cell centres on a lattice,
    bilinear weights, float32 result.  Interior points only."""
    res, rows, cols = m["resolution"], m["rows"], m["cols"]
    gx = (m["center_x"] + rows * res / 2 - x) / res - 0.5     # fractional row index
    gy = (m["center_y"] + cols * res / 2 - y) / res - 0.5
    i0, j0 = int(np.floor(gx)), int(np.floor(gy))
    fx, fy = gx - i0, gy - j0
    d = m["distance"].astype(np.float64)
    v = (d[i0, j0] * (1 - fx) * (1 - fy) + d[i0 + 1, j0] * fx * (1 - fy) + d[i0, j0 + 1] * (1 - fx) * fy
         + d[i0 + 1, j0 + 1] * fx * fy)
    return v


def test_map_distance_matches_geometry(field):
    rng = np.random.default_rng(1)
    xy = np.stack([rng.uniform(-105, 105, 3000), rng.uniform(-24, 24, 3000)], 1)
    got = oracle.map_distance(field, xy)
    want = np.array([_numpy_bilinear(field, x, y) for x, y in xy])
    assert np.abs(got - want).max() <= 5e-6          # float32 rounding of a <= 20 m value
    # outside the map -> 0 (Map.cpp:20); the +x/+y edge is inside, the -x/-y edge is not
    lx, ly = field["rows"] * field["resolution"], field["cols"] * field["resolution"]
    assert oracle.map_distance(field, [[lx / 2 + 0.01, 0.0]])[0] == 0.0
    assert oracle.map_distance(field, [[0.0, -ly / 2 - 0.01]])[0] == 0.0
    assert oracle.map_distance(field, [[1e9, 1e9]])[0] == 0.0
    # half a cell from the border a neighbour is missing -> nearest-cell value
    i, j = 0, 17
    x = field["center_x"] + lx / 2 - 0.02
    y = field["center_y"] + ly / 2 - (j + 0.5) * field["resolution"]
    assert oracle.map_distance(field, [[x, y]])[0] == float(field["distance"][i, j])


def test_spline_matches_scipy_natural():
    rng = np.random.default_rng(2)
    t = np.cumsum(rng.uniform(0.2, 0.5, 60))
    y = np.cos(0.3 * t) * 5 + rng.standard_normal(60) * 0.05
    c = oracle.spline_fit(t, y)
    cs = CubicSpline(t, y, bc_type="natural")
    # scipy stores c[k, i] for (x - x_i)^(3-k)
    assert np.abs(c[:-1, 0] - cs.c[0]).max() <= 1e-9
    assert np.abs(c[:-1, 1] - cs.c[1]).max() <= 1e-9
    assert np.abs(c[:-1, 2] - cs.c[2]).max() <= 1e-9
    at = rng.uniform(t[0], t[-1], 200)
    for order in (0, 1, 2):
        assert np.abs(oracle.spline_eval(t, c, at, order) - cs(at, order)).max() <= 1e-8
    # extrapolation: natural boundary -> b = 0 at both ends -> straight lines (spline.cpp:259-265)
    lo = oracle.spline_eval(t, c, [t[0] - 2.0])[0]
    assert abs(lo - (y[0] + c[0, 2] * -2.0)) <= 1e-12
    hi = oracle.spline_eval(t, c, [t[-1] + 3.0])[0]
    assert abs(hi - (y[-1] + cs(t[-1], 1) * 3.0)) <= 1e-9
    assert oracle.spline_eval(t, c, [t[-1] + 3.0], 2)[0] == 0.0


def _wall_map(half_width, rows=600, cols=200, res=0.2):
    """Two walls at |y| = half_width: distance = half_width - |y| (>= 0)."""
    ys = cols * res / 2 - (np.arange(cols) + 0.5) * res
    d = np.maximum(half_width - np.abs(ys), 0.0).astype(np.float32)
    return dict(distance=np.ascontiguousarray(np.broadcast_to(d, (rows, cols))), rows=rows, cols=cols, resolution=res,
                center_x=0.0, center_y=0.0)


def test_clearance_known_answers():
    """Hand-computed ray marches (reference_path_impl.cpp:283-472) between two walls."""
    p = oracle.default_params()
    m = _wall_map(3.05)
    # free at the centre: coarse march stops at 2.0 (3.05-2.0 < r=1.1727) -> 1.5, fine steps 1.6,1.7,1.8 pass, 1.9 fails
    lb = oracle.clearance_strict(p, m, 0.0, 0.0, 0.0)
    assert np.allclose(lb, [1.8, -1.8], atol=1e-12)
    # off-centre by +0.5: left wall nearer.  left: 1.0 -> d=1.55 ok, 1.5 -> 1.05 fail -> 1.0; fine 1.1,1.2,1.3 ok,
    # 1.4 -> 1.15 fail.  right: coarse 2.5 fails -> -2.0.  The reference's fine search on the right evaluates
    # state + right_bound * (cos, sin)(right_angle) with right_bound NEGATIVE (reference_path_impl.cpp:455-465),
    # i.e. it samples the LEFT side at |right_bound|: -2.1 -> y = 2.6, d = 0.45 -> rejected -> stays -2.0.
    lb = oracle.clearance_strict(p, m, 3.0, 0.5, 0.0)
    assert np.allclose(lb, [1.3, -2.0], atol=1e-12)
    # heading pi: left now faces -y (coarse 2.0, fine to 2.3); right coarse -1.0, and its fine search samples the
    # free far side four times -> -1.4
    lb = oracle.clearance_strict(p, m, 3.0, 0.5, np.pi)
    assert np.allclose(lb, [2.3, -1.4], atol=1e-12)
    # wide open: the march gives up after 10 coarse steps -> 4.5 + 4 fine steps = 4.9
    lb = oracle.clearance_strict(p, _wall_map(15.0), 0.0, 0.0, 0.3)
    assert np.allclose(lb, [4.9, -4.9], atol=1e-12)
    # start in collision (distance 0.55 < r) next to the left wall: expand to the right only
    lb = oracle.clearance_strict(p, m, 0.0, 2.5, 0.0)
    assert lb[0] < 0 and lb[1] < lb[0]
    # a corridor narrower than the vehicle circle: nowhere free -> left == right (blocked)
    lb = oracle.clearance_strict(p, _wall_map(1.0), 0.0, 0.0, 0.0)
    blocked = oracle.update_bounds(p, _wall_map(1.0), synth.map_reference_paths(1, 20, x_range=(-10, -10)), mode=1)
    assert blocked["n_valid"][0] < 20


def test_car_circles_known_values():
    """CarGeometry::setCircles with the default vehicle (car_geometry.cpp:38-57)."""
    c = oracle.car_circles(oracle.default_params())
    assert np.allclose(c[0], [1.45, 0.0, np.hypot(2.45, 1.0)])
    assert np.allclose(c[1], [-1.0 + 0.5, -0.5, np.sqrt(0.5)])       # rr
    assert np.allclose(c[4], [3.9 - 0.5, 0.5, np.sqrt(0.5)])         # fl
    assert np.allclose(c[5], [1.45 + 0.725, 0.0, np.hypot(2.0, 1.45) / 2])


@pytest.mark.parametrize("mode", [planner.BOUNDS_SIMPLE, planner.BOUNDS_IMPROVED])
def test_emu_bounds_match_oracle(field, mode):
    p = oracle.default_params()
    b = synth.map_reference_paths(48, 120, n_points=np.r_[np.full(40, 120), [2, 3, 7, 50, 199, 64, 33, 90]].astype(np.int32),
                                  **WILD)
    spl = planner.reference_splines(synth.slice_batch(b, 0, 40)) if mode == planner.BOUNDS_IMPROVED else None
    if mode == planner.BOUNDS_IMPROVED:
        b = synth.slice_batch(b, 0, 40)
    want = oracle.update_bounds(p, field, b, mode=mode, splines=spl)
    got = emu.update_bounds(p, field, b, mode=mode, splines=spl)
    assert (got["n_valid"] == want["n_valid"]).all()
    assert (want["n_valid"] < b["n_points"]).any() and (want["n_valid"] == b["n_points"]).any()
    W = want["bounds"].view(np.float64).reshape(-1, 8)
    G = got["bounds"].view(np.float64).reshape(-1, 8)
    for i in range(len(b["n_points"])):
        lo = b["offsets"][i]
        assert (G[lo:lo + want["n_valid"][i]] == W[lo:lo + want["n_valid"][i]]).all()
    # bounds carry structure: not all at the 4.9 m march limit
    assert (np.abs(W[:, 0]) < 4.0).any()


def test_emu_map_and_collision_match_oracle(field):
    p = oracle.default_params()
    rng = np.random.default_rng(5)
    xy = np.stack([rng.uniform(-112, 112, 5000), rng.uniform(-26, 26, 5000)], 1)
    assert (emu.map_distance(field, xy) == oracle.map_distance(field, xy)).all()
    st = np.zeros(4000, dtype=STATE_DTYPE)
    st["x"], st["y"], st["z"] = rng.uniform(-112, 112, 4000), rng.uniform(-26, 26, 4000), rng.uniform(-np.pi, np.pi, 4000)
    want = oracle.check_states(p, field, st)
    assert (emu.check_states(p, field, st) == want).all()
    assert 0.05 < want.mean() < 0.95


def _solved_like_paths(field, B=24, n=150):
    """Paths shaped like QP output (x, y, heading, k; s left at zero): half stay in the corridor, half wander
    into obstacles."""
    tame = synth.map_reference_paths(B // 2, n)
    wild = synth.map_reference_paths(B - B // 2, n, first_path=1000, y_range=(-2.0, 2.0), heading_range=0.06,
                                     curvature_amp=0.01)
    paths = np.concatenate([tame["ref"], wild["ref"]])
    paths["s"] = 0.0
    return np.concatenate([tame["n_points"], wild["n_points"]]), paths


def test_emu_tails_match_oracle(field):
    p = oracle.default_params()
    n_points, paths = _solved_like_paths(field)
    want = oracle.finish_raw(p, field, n_points, paths)
    got = emu.finish_raw(p, field, n_points, paths)
    assert (got["n_kept"] == want["n_kept"]).all() and (got["ok"] == want["ok"]).all()
    off = np.concatenate([[0], np.cumsum(n_points)])
    for i in range(len(n_points)):                                    # same serial sum, bit for bit, on what is kept
        sl = slice(off[i], off[i] + want["n_kept"][i])
        assert (got["states"]["s"][sl] == want["states"]["s"][sl]).all()
    assert (want["n_kept"] < n_points).any() and (want["n_kept"] == n_points).any()
    assert set(np.unique(want["ok"])) == {0, 1}
    # the raw tail returns true iff nothing collided or the cut is at s >= 20 m (path_optimizer.cpp:199)
    for i in range(len(n_points)):
        k = want["n_kept"][i]
        exp = 1 if k == n_points[i] else int(k > 0 and want["states"]["s"][off[i] + k - 1] >= 20)
        assert want["ok"][i] == exp
    # no collision check -> nothing is cut
    free = oracle.finish_raw(p, field, n_points, paths, collision_check=False)
    assert (free["n_kept"] == n_points).all() and free["ok"].all()

    # densify: needs s, take the re-accumulated one
    src = free["states"]
    want = oracle.densify(p, field, n_points, src, 0.3, True, 200)
    got = emu.densify(p, field, n_points, src, 0.3, True, 200)
    assert (got["n_out"] == want["n_out"]).all() and (got["ok"] == want["ok"]).all()
    for i in range(len(n_points)):
        k = want["n_out"][i]
        for f in ("x", "y", "z", "k", "s"):
            assert np.abs(got["states"][f][i, :k] - want["states"][f][i, :k]).max(initial=0.0) <= 1e-9
    assert (want["n_out"] == 200).any() or (want["n_out"] < 150).any()
    # sample count rule: i * spacing <= s_end (path_optimizer.cpp:213)
    nc = oracle.densify(p, field, n_points, src, 0.25, False, 400)
    for i in range(len(n_points)):
        s_end = src["s"][off[i + 1] - 1]
        assert nc["n_out"][i] == int(np.floor(s_end / 0.25 + 1e-12)) + 1 and nc["ok"][i] == 1
    # overflow of the caller's buffer is reported, not silently cut
    small = oracle.densify(p, field, n_points, src, 0.3, False, 20)
    assert (small["n_out"] == 20).all() and not small["ok"].any()
    assert (emu.densify(p, field, n_points, src, 0.3, False, 20)["ok"] == 0).all()


def test_densified_states_follow_the_path():
    """Resampled positions interpolate the input, heading/curvature match a circle's."""
    p = oracle.default_params()
    n = 80
    s = np.arange(n) * 0.4
    R = 25.0
    path = np.zeros(n, dtype=STATE_DTYPE)
    path["x"], path["y"], path["s"] = R * np.sin(s / R), R * (1 - np.cos(s / R)), s
    m = _wall_map(90.0, rows=400, cols=1000)
    out = oracle.densify(p, m, np.array([n], dtype=np.int32), path, 0.3, False, 200)
    k = out["n_out"][0]
    st = out["states"][0, :k]
    assert k == int(s[-1] / 0.3) + 1
    mid = slice(5, k - 5)
    assert np.abs(st["x"][mid] - R * np.sin(st["s"][mid] / R)).max() < 1e-5
    assert np.abs(st["z"][mid] - st["s"][mid] / R).max() < 1e-4
    assert np.abs(st["k"][mid] - 1 / R).max() < 1e-3


def test_plan_chain_oracle(field):
    """solveWithoutSmoothing shape: blocked paths are trimmed before the QP, unsolved QPs return
    false, solved ones go through the raw tail."""
    p = oracle.default_params()
    b = synth.map_reference_paths(12, 120, y_range=(-1.5, 1.5), heading_range=0.03, curvature_amp=0.01)
    r = oracle.plan(p, field, b, bounds_mode=planner.BOUNDS_SIMPLE)
    nv = oracle.update_bounds(p, field, b, mode=planner.BOUNDS_SIMPLE)["n_valid"]
    solved = r["status"] == 1
    assert solved.any()
    assert (r["n_out"][solved] <= nv[solved]).all()
    assert (r["ok"][~solved] == 0).all() and (r["n_out"][~solved] == 0).all()
    # the QP inside the chain is the plain hot path on the trimmed reference
    i = int(np.flatnonzero(solved)[0])
    one = synth.slice_batch(b, i, i + 1)
    one["ref"] = one["ref"][:nv[i]]
    one["bounds"] = r["bounds"][b["offsets"][i]:b["offsets"][i] + nv[i]]
    one["n_points"] = np.array([nv[i]], dtype=np.int32)
    one["offsets"] = np.array([0, nv[i]], dtype=np.int32)
    q = oracle.solve_batch(p, 0, one)
    lo = b["offsets"][i]
    k = r["n_out"][i]
    assert np.abs(q["states"]["x"][:k] - r["states"]["x"][lo:lo + k]).max() == 0.0


def _golden_env():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "env_small.npz"))
    field = dict(distance=g["map_distance"], rows=g["map_distance"].shape[0], cols=g["map_distance"].shape[1],
                 resolution=float(g["map_geo"][0]), center_x=float(g["map_geo"][1]), center_y=float(g["map_geo"][2]))
    b = dict(n_points=g["n_points"], ref=g["ref"], x0=g["x0"], end_heading=g["end_heading"])
    b["offsets"] = np.concatenate([[0], np.cumsum(b["n_points"])]).astype(np.int32)
    spl = dict(n_knots=g["n_points"], knots=g["knots"], x_coef=g["x_coef"], y_coef=g["y_coef"])
    return g, field, b, spl


def test_golden_env_fixture_is_reproduced():
    """The committed fixture (tests/golden/make_golden_env.py) is what the oracle computes today, and the product's
    per-thread device source (host build) reproduces its map lookups, bounds and collision flags."""
    g, field, b, spl = _golden_env()
    p = oracle.default_params()
    assert (oracle.map_distance(field, g["xy"]) == g["xy_distance"]).all()
    assert (emu.map_distance(field, g["xy"]) == g["xy_distance"]).all()
    for mode, tag in ((planner.BOUNDS_SIMPLE, "simple"), (planner.BOUNDS_IMPROVED, "improved")):
        sp = spl if mode == planner.BOUNDS_IMPROVED else None
        r = oracle.update_bounds(p, field, b, mode=mode, splines=sp)
        e = emu.update_bounds(p, field, b, mode=mode, splines=sp)
        assert (r["n_valid"] == g[f"n_valid_{tag}"]).all() and (e["n_valid"] == g[f"n_valid_{tag}"]).all()
        assert r["bounds"].tobytes() == g[f"bounds_{tag}"].tobytes()
        assert e["bounds"].tobytes() == g[f"bounds_{tag}"].tobytes()
    assert (oracle.check_states(p, field, b["ref"]) == g["collision_free"]).all()
    assert (emu.check_states(p, field, b["ref"]) == g["collision_free"]).all()
    pl = oracle.plan(p, field, b, bounds_mode=planner.BOUNDS_SIMPLE)
    assert (pl["status"] == g["plan_simple_raw_status"]).all() and (pl["iters"] == g["plan_simple_raw_iters"]).all()
    assert set(g["plan_simple_raw_status"].tolist()) >= {1, -3}


def _config1():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "config1_benchmark_map.npz"))
    field = dict(distance=g["map_distance"], rows=int(g["image_shape"][0]), cols=int(g["image_shape"][1]),
                 resolution=float(g["map_geo"][0]), center_x=float(g["map_geo"][1]), center_y=float(g["map_geo"][2]))
    b = dict(n_points=g["n_points"], ref=g["ref"], x0=g["x0"], end_heading=g["end_heading"])
    b["offsets"] = np.array([0, int(g["n_points"][0])], dtype=np.int32)
    spl = dict(n_knots=np.array([len(g["knots"])], dtype=np.int32), knots=g["knots"], x_coef=g["x_coef"], y_coef=g["y_coef"])
    return g, field, b, spl


def test_config1_reference_benchmark_case():
    """BASELINE config 1: the reference's own benchmark inputs (obstacles_for_benchmark.png, the 100-point polyline and
    poses of path_optimizer_benchmark.cpp:47-82; fixture generated by tests/golden/make_golden_config1.py).  The map
    is the reference's: 495 x 497 px at 0.2 m, 7 % occupied; the oracle reproduces the committed outputs,
    and the product's kernel sources (host builds) reproduce bounds and QP."""
    g, field, b, spl = _config1()
    assert tuple(g["image_shape"]) == (495, 497) and abs(float(g["occupied_fraction"]) - 0.07) < 0.005
    assert 125 <= int(g["n_points"][0]) <= 135                       # 39.56 m at 0.3 m
    p = oracle.default_params()
    for mode, tag in ((planner.BOUNDS_IMPROVED, "improved"), (planner.BOUNDS_SIMPLE, "simple")):
        sp = spl if mode == planner.BOUNDS_IMPROVED else None
        r = oracle.update_bounds(p, field, b, mode=mode, splines=sp)
        assert r["bounds"].tobytes() == g[f"bounds_{tag}"].tobytes() and (r["n_valid"] == g[f"n_valid_{tag}"]).all()
        e = emu.update_bounds(p, field, b, mode=mode, splines=sp)
        assert e["bounds"].tobytes() == g[f"bounds_{tag}"].tobytes()
    pl = oracle.plan(p, field, b, bounds_mode=planner.BOUNDS_IMPROVED, splines=spl)
    assert pl["status"][0] == g["plan_improved_status"][0] == 1 and pl["iters"][0] == g["plan_improved_iters"][0]
    assert pl["ok"][0] == 1 and pl["n_out"][0] == g["n_points"][0]
    assert np.abs(pl["states"]["x"] - g["plan_improved_states"]["x"]).max() == 0.0
    # the QP of the chain on the kernel source (eight-warp class: 132 stations)
    qb = dict(b)
    qb["bounds"] = g["bounds_improved"]
    k = emu.solve_batch(p, qb, variant=8)
    assert k["status"][0] == 1 and k["iters"][0] == g["plan_improved_iters"][0]
    assert np.abs(k["states"]["x"] - g["plan_improved_states"]["x"]).max() <= 1e-9
    # the optimized path stays clear of the obstacles and inside the corridor it was given
    assert oracle.check_states(p, field, pl["states"]).all()


def test_update_limits_host_matches_oracle_bit_for_bit(oracle_params):
    """ReferencePathImpl::updateLimits (reference_path_impl.cpp:203-235): friction circle and rate limit from (v, a),
    DBL_MAX at standstill, the use_spline_ branch, and a hand-computed value.  The library's host helper needs no GPU."""
    from path_optimizer_b200 import planner
    from path_optimizer_b200.abi import STATE_DTYPE
    rng = np.random.default_rng(5)
    ref = np.zeros(400, dtype=STATE_DTYPE)
    ref["v"] = rng.uniform(0.0, 12.0, 400)
    ref["a"] = rng.uniform(-3.0, 3.0, 400)
    ref["v"][:4] = [0.0, 1e-4, 1.0000001e-4, 5.0]
    ref["a"][3] = 1.0
    ref["a"][10] = 4.5      # beyond mu g = 3.92: sqrt of a negative number, NaN like the reference
    mk, mkp = planner.update_limits(oracle_params, ref)
    omk, omkp = oracle.update_limits(oracle_params, ref)
    assert np.array_equal(mk, omk, equal_nan=True) and np.array_equal(mkp, omkp)
    assert mk[0] == mk[1] == np.finfo(np.float64).max and mkp[0] == np.finfo(np.float64).max
    assert np.isnan(mk[10]) and np.isfinite(mk[2])
    assert mk[3] == np.sqrt((0.4 * 9.8) ** 2 - 1.0) / 25.0 and mkp[3] == 0.1 / 5.0
    smk, smkp = planner.update_limits(oracle_params, ref, from_spline=True)
    osk, oskp = oracle.update_limits(oracle_params, ref, from_spline=True)
    assert np.array_equal(smk, osk) and np.array_equal(smkp, oskp)
    assert (smk == np.tan(oracle_params.max_steering_angle) / oracle_params.wheel_base).all()
    assert (smkp == np.finfo(np.float64).max).all()
