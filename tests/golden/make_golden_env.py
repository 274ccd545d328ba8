"""Generates tests/golden/env_small.npz: a small distance map, six reference lines and the C oracle's outputs
for every stage around the QP (map lookups, both bounds variants, collision flags, the raw and the densifying
tail, the whole solveWithoutSmoothing chain).  Run from the repo root:  python tests/golden/make_golden_env.py

The reference ships no fixtures for these stages either: the file pins the ORACLE against
regressions and pins the CUDA path against the oracle without needing the oracle at test time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from path_optimizer_b200 import planner, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    p = oracle.default_params()
    field = synth.disc_field_map(rows=420, cols=160, n_discs=110, keep_clear_halfwidth=1.6)
    tame = synth.map_reference_paths(3, 70, x_range=(-38.0, 10.0))
    wild = synth.map_reference_paths(5, 70, first_path=900, x_range=(-38.0, 10.0), y_range=(-3.0, 3.0), heading_range=0.08,
                                     curvature_amp=0.02)
    b = dict(n_points=np.concatenate([tame["n_points"], wild["n_points"]]), ref=np.concatenate([tame["ref"], wild["ref"]]),
             x0=np.concatenate([tame["x0"], wild["x0"]]), end_heading=np.concatenate([tame["end_heading"], wild["end_heading"]]))
    b["offsets"] = np.concatenate([[0], np.cumsum(b["n_points"])]).astype(np.int32)
    spl = planner.reference_splines(b)
    rng = np.random.default_rng(5)
    xy = np.stack([rng.uniform(-43, 43, 400), rng.uniform(-17, 17, 400)], 1)
    out = dict(map_distance=field["distance"], map_geo=np.array([field["resolution"], field["center_x"], field["center_y"]]),
               n_points=b["n_points"], ref=b["ref"], x0=b["x0"], end_heading=b["end_heading"],
               knots=spl["knots"], x_coef=spl["x_coef"], y_coef=spl["y_coef"], xy=xy,
               xy_distance=oracle.map_distance(field, xy))
    for mode, tag in ((planner.BOUNDS_SIMPLE, "simple"), (planner.BOUNDS_IMPROVED, "improved")):
        r = oracle.update_bounds(p, field, b, mode=mode, splines=spl if mode == planner.BOUNDS_IMPROVED else None)
        out[f"bounds_{tag}"], out[f"n_valid_{tag}"] = r["bounds"], r["n_valid"]
        for om, otag in ((planner.OUTPUT_RAW, "raw"), (planner.OUTPUT_DENSIFY, "dense")):
            pl = oracle.plan(p, field, b, bounds_mode=mode, splines=spl if mode == planner.BOUNDS_IMPROVED else None,
                             output_mode=om, max_out=128)
            for k in ("states", "n_out", "ok", "status", "iters"):
                out[f"plan_{tag}_{otag}_{k}"] = pl[k]
    st = np.array(b["ref"])
    out["collision_free"] = oracle.check_states(p, field, st)
    np.savez_compressed(os.path.join(OUT, "env_small.npz"), **out)
    print("n_valid", out["n_valid_simple"], out["n_valid_improved"], "status", out["plan_simple_raw_status"],
          "ok", out["plan_simple_raw_ok"], "free", int(out["collision_free"].sum()), "/", len(st))


if __name__ == "__main__":
    main()
