"""Diagnostic: per-phase cycle breakdown of the production kernel (needs a -DPQP_PHASE_TIMING build of
libpqp.so selected with PQP_LIB=...; prints to stderr from inside pqp_solve_batch)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from path_optimizer_b200 import synth
from path_optimizer_b200.solver import BatchPathSolver
b = synth.straight_corridors(int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 100)
s = BatchPathSolver(max_batch=1024, max_total_points=1024 * 100)
s.solve(b)
r = s.solve(b)
print("kernel_ms", r["stats"].kernel_ms, "iters", r["iters"].mean())
