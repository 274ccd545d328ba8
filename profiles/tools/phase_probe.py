"""Diagnostic: per-phase cycle breakdown of the production kernels (needs a -DPQP_PHASE_TIMING build of
libpqp.so selected with PQP_LIB=...; prints to stderr from inside pqp_solve_batch)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from path_optimizer_b200 import synth
from path_optimizer_b200.solver import BatchPathSolver
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
b = synth.straight_corridors(batch, n) if n == 100 else synth.curvy_corridors(batch, n)
s = BatchPathSolver(max_batch=batch, max_total_points=batch * n)
s.solve(b)
r = s.solve(b)
print("kernel_ms", r["stats"].kernel_ms, "iters", r["iters"].mean())
