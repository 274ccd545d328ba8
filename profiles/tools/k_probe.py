"""Diagnostic: "K" kernel time per shape (kernel span of pqp_solve_batch; PQP_LIB selects an A/B build)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from path_optimizer_b200 import synth
from path_optimizer_b200.solver import BatchPathSolver
shapes = [("curvy 1024x100", synth.curvy_corridors(1024, 100)), ("straight 1024x100", synth.straight_corridors(1024, 100)),
          ("straight 8192x100", synth.straight_corridors(8192, 100))]
if "--long" in sys.argv:
    shapes += [("curvy 1024x200", synth.curvy_corridors(1024, 200)), ("curvy 512x400", synth.curvy_corridors(512, 400))]
out = []
for name, bb in shapes:
    B = len(bb["n_points"])
    s = BatchPathSolver(max_batch=B, max_total_points=int(bb["offsets"][-1]))
    s.solve(bb, "K")
    ms = min(s.solve(bb, "K")["stats"].kernel_ms for _ in range(3))
    r = s.solve(bb, "K")
    st = int((bb["n_points"] * r["iters"]).sum())
    out.append(f"{name}: {ms:.2f} ms, {B / ms:.1f} k solves/s, {ms * 1e6 / st:.3f} ns/station-it")
    s.close()
print(os.environ.get("PQP_LIB", "libpqp.so"), "|", " | ".join(out), flush=True)
