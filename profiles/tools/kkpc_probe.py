"""Diagnostic: K / KPC throughput on the generic kernel (1024 x 100)."""
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
