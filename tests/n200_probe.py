import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from path_optimizer_b200 import synth
from path_optimizer_b200.solver import BatchPathSolver
for n in (200, 150):
    b = synth.curvy_corridors(1024, n)
    s = BatchPathSolver(max_batch=1024, max_total_points=1024*n)
    s.solve(b)
    r = s.solve(b)
    print('N', n, 'kernel_ms', round(r['stats'].kernel_ms,3), 'total', round(r['stats'].h2d_ms+r['stats'].kernel_ms+r['stats'].d2h_ms,3), 'iters', r['iters'].mean(), 'solved', (r['status']==1).mean(), flush=True)
    s.close()
