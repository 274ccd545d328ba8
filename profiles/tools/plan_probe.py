"""Diagnostic: where a config-3 planner iteration (pqp_plan_batch, 8192 x 200, host buffers) spends its time."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from path_optimizer_b200 import synth, workloads, planner
field = synth.disc_field_map()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pl = planner.PathPlanner(max_batch=B + 1024, max_total_points=(B + 1024) * 200)
pl.set_map(field)
b = workloads.config3_candidates(B)
pl.plan(b)
for _ in range(3):
    t0 = time.perf_counter()
    r = pl.plan(b)
    ms = (time.perf_counter() - t0) * 1e3
    st = r["stats"]
    print(f"wall {ms:.1f} ms | h2d {st.h2d_ms:.1f} ({st.h2d_bytes / 1e6:.0f} MB) kernels {st.kernel_ms:.1f} d2h {st.d2h_ms:.1f} ({st.d2h_bytes / 1e6:.0f} MB) "
          f"launches {st.kernel_launches} solved {int(r['solved'].sum())}", flush=True)
rb = pl.update_bounds(b)
t0 = time.perf_counter(); rb = pl.update_bounds(b); print(f"update_bounds alone: wall {(time.perf_counter() - t0) * 1e3:.1f} ms, kernel {rb['stats'].kernel_ms:.2f}")
