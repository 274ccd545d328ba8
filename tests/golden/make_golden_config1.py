"""Generates tests/golden/config1_benchmark_map.npz -- BASELINE config 1, the reference's own CPU-runnable case:
one path on obstacles_for_benchmark.png with the 100-point polyline and the start / goal poses hard-coded in
src/test/path_optimizer_benchmark.cpp:47-82.  Run HERE (it reads /root/reference; the GPU box only sees the
fixture):  python tests/golden/make_golden_config1.py

What is taken from the reference: the image, the polyline and the poses (its benchmark's inputs), and the recipe
that turns the image into the "distance" layer (path_optimizer_benchmark.cpp:28-44: cv::distanceTransform(L2,
MASK_PRECISE) * 0.2 m on a grid_map centred at the origin).  The reference's smoothing / DP search stage
(PathOptimizer::solve) cannot be rebuilt here (tinyspline, IPOPT, OSQP absent), so -- as SURVEY.md section 8d
specifies for this plumbing case -- the polyline itself is the reference line: a natural cubic spline through it
(chord-length parameter), resampled every 0.3 m, heading and curvature from the spline derivatives, then
solveWithoutSmoothing's chain: updateBoundsImproved -> KP QP -> raw output with collision check.  Outputs are the
C oracle's; PARITY UNPINNED with respect to a run of the reference itself.
"""
import os
import re
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from path_optimizer_b200 import planner  # noqa: E402
from path_optimizer_b200.abi import STATE_DTYPE  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def parse_benchmark_inputs():
    src = open(os.path.join(REF, "src", "test", "path_optimizer_benchmark.cpp")).read()
    lists = re.findall(r"std::vector<double>\s+([xy])_list_\s*=\s*\{([^}]*)\}", src)
    x = np.array([float(v) for v in lists[0][1].replace("\n", " ").split(",")])
    y = np.array([float(v) for v in lists[1][1].replace("\n", " ").split(",")])
    assert lists[0][0] == "x" and lists[1][0] == "y" and len(x) == len(y)

    def pose(name):
        return [float(re.search(rf"{name}_state\.{f}\s*=\s*([-0-9.eE]+);", src).group(1)) for f in ("x", "y", "z", "k")]
    return x, y, pose("start"), pose("goal")


def main():
    x, y, start, goal = parse_benchmark_inputs()
    img = cv2.imread(os.path.join(REF, "obstacles_for_benchmark.png"), cv2.IMREAD_GRAYSCALE)
    res = 0.2
    dist = (cv2.distanceTransform(img, cv2.DIST_L2, cv2.DIST_MASK_PRECISE) * np.float32(res)).astype(np.float32)
    field = dict(distance=np.ascontiguousarray(dist), rows=img.shape[0], cols=img.shape[1], resolution=res,
                 center_x=0.0, center_y=0.0)
    # reference line: natural spline through the polyline, resampled at 0.3 m
    s_poly = np.concatenate([[0.0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))])
    xc, yc = oracle.spline_fit(s_poly, x), oracle.spline_fit(s_poly, y)
    s = []
    acc = 0.0
    while acc <= s_poly[-1]:
        s.append(acc)
        acc += 0.3
    s = np.array(s)
    ref = np.zeros(len(s), dtype=STATE_DTYPE)
    ref["s"] = s
    ref["x"], ref["y"] = oracle.spline_eval(s_poly, xc, s, 0), oracle.spline_eval(s_poly, yc, s, 0)
    x1, y1 = oracle.spline_eval(s_poly, xc, s, 1), oracle.spline_eval(s_poly, yc, s, 1)
    x2, y2 = oracle.spline_eval(s_poly, xc, s, 2), oracle.spline_eval(s_poly, yc, s, 2)
    ref["z"] = np.arctan2(y1, x1)
    ref["k"] = (x1 * y2 - y1 * x2) / np.power(x1 ** 2 + y1 ** 2, 1.5)
    batch = dict(n_points=np.array([len(s)], dtype=np.int32), offsets=np.array([0, len(s)], dtype=np.int32), ref=ref,
                 x0=np.array([[0.0, 0.0, start[3]]]), end_heading=np.array([goal[2]]))
    spl = dict(n_knots=np.array([len(s_poly)], dtype=np.int32), knots=s_poly, x_coef=xc, y_coef=yc)
    p = oracle.default_params()
    out = dict(image_shape=np.array(img.shape), occupied_fraction=float((img == 0).mean()), map_distance=dist,
               map_geo=np.array([res, 0.0, 0.0]), poly_x=x, poly_y=y, start=np.array(start), goal=np.array(goal),
               n_points=batch["n_points"], ref=ref, x0=batch["x0"], end_heading=batch["end_heading"],
               knots=s_poly, x_coef=xc, y_coef=yc)
    for mode, tag in ((planner.BOUNDS_IMPROVED, "improved"), (planner.BOUNDS_SIMPLE, "simple")):
        sp = spl if mode == planner.BOUNDS_IMPROVED else None
        b = oracle.update_bounds(p, field, batch, mode=mode, splines=sp)
        r = oracle.plan(p, field, batch, bounds_mode=mode, splines=sp)
        out[f"bounds_{tag}"], out[f"n_valid_{tag}"] = b["bounds"], b["n_valid"]
        for k in ("states", "n_out", "ok", "status", "iters"):
            out[f"plan_{tag}_{k}"] = r[k]
        print(tag, "stations", len(s), "n_valid", b["n_valid"], "status", r["status"], "iters", r["iters"], "n_out", r["n_out"],
              "ok", r["ok"], "ub range", float(b["bounds"]["c0_ub"].min()), float(b["bounds"]["c0_ub"].max()))
    np.savez_compressed(os.path.join(OUT, "config1_benchmark_map.npz"), **out)
    print("image", img.shape, "occupied", out["occupied_fraction"], "polyline length", s_poly[-1])


if __name__ == "__main__":
    main()
