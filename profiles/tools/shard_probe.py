"""Diagnostic: device time of the solve kernel for the eight 1024-path shards of the multi-GPU bench."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from path_optimizer_b200 import synth
from path_optimizer_b200.solver import BatchPathSolver
s = BatchPathSolver(max_batch=1024, max_total_points=1024 * 100)
for r in range(8):
    b = synth.straight_corridors(1024, 100, first_path=r * 1024)
    s.solve(b)
    k = min(s.solve(b)["stats"].kernel_ms for _ in range(3))
    res = s.solve(b)
    print("shard", r, "pipeline-middle ms", round(k, 3), "iters mean", res["iters"].mean(), "max", res["iters"].max(), flush=True)
