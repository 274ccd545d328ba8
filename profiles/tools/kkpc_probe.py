"""Diagnostic: K / KPC throughput (1024 x 100 and other shapes); PQP_GENERIC_K=1 / PQP_GENERIC_KPC=1 select round 1's generic kernel."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from path_optimizer_b200 import synth
from path_optimizer_b200.solver import BatchPathSolver
b = synth.curvy_corridors(1024, 100)
s = BatchPathSolver(max_batch=1024, max_total_points=1024 * 100)
total = 1024 * 100
for form, mk, mkp in (("K", None, None), ("KPC", np.full(total, 0.2), np.full(total, 0.05))):
    s.solve(b, form, max_k=mk, max_kp=mkp)
    r = s.solve(b, form, max_k=mk, max_kp=mkp)
    print(form, "kernel_ms", round(r["stats"].kernel_ms, 2), "solves/s", round(1024 / (r["stats"].kernel_ms * 1e-3)),
          "iters", r["iters"].mean(), "solved", (r["status"] == 1).mean(), flush=True)
# "K" at other shapes: straight corridors (config 2 / 4 shapes) and long paths
for name, bb in (("K straight 1024x100", synth.straight_corridors(1024, 100)), ("K curvy 1024x200", synth.curvy_corridors(1024, 200)),
                 ("K curvy 512x400", synth.curvy_corridors(512, 400))):
    B = len(bb["n_points"])
    s2 = BatchPathSolver(max_batch=B, max_total_points=int(bb["offsets"][-1]))
    s2.solve(bb, "K")
    r = s2.solve(bb, "K")
    st = int((bb["n_points"] * r["iters"]).sum())
    print(name, "kernel_ms", round(r["stats"].kernel_ms, 2), "solves/s", round(B / (r["stats"].kernel_ms * 1e-3)),
          "iters", r["iters"].mean(), "ns/station-it", round(r["stats"].kernel_ms * 1e6 / st, 4), flush=True)
    s2.close()
