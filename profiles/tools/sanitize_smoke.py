import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from path_optimizer_b200 import synth, planner
pl = planner.PathPlanner(max_batch=64, max_total_points=64*300)
b = synth.curvy_corridors(6, n_points=[40, 100, 7, 128, 60, 2])
r = pl.solve(b); print('KP', r['status'], r['iters'])
b2 = synth.infeasible_corridors(4, 50)
r = pl.solve(b2); print('KP infeasible', r['status'], r['iters'])
b3 = synth.curvy_corridors(2, 150); r = pl.solve(b3); print('KP 150', r['status'], r['iters'])
b3 = synth.curvy_corridors(2, n_points=[200, 193]); r = pl.solve(b3); print('KP 200/193 (two-level)', r['status'], r['iters'])
b3 = synth.curvy_corridors(1, 250); r = pl.solve(b3); print('KP 250', r['status'], r['iters'])
b3 = synth.curvy_corridors(1, 180); b3['ref']['s'] = np.arange(180) * 0.25; r = pl.solve(b3); print('KP 180 keep4', r['status'], r['iters'])
b4 = synth.curvy_corridors(1, 300); r = pl.solve(b4); print('KP 300', r['status'], r['iters'])
b5 = synth.curvy_corridors(2, 40); b5['ref']['s'] = np.tile(np.arange(40)*0.2, 2); r = pl.solve(b5); print('KP keep6', r['status'], r['iters'])
r = pl.solve(synth.curvy_corridors(2, 30), 'K'); print('K', r['status'], r['iters'])
field = synth.disc_field_map(rows=400, cols=150)
pl.set_map(field)
bm = synth.map_reference_paths(6, 60, x_range=(-30, 10))
spl = planner.reference_splines(bm)
for mode in (0, 1):
    for om in (0, 1):
        r = pl.plan(bm, bounds_mode=mode, splines=spl if mode == 0 else None, output_mode=om, max_out=128)
        print('plan', mode, om, r['status'], r['n_out'], r['ok'])
pl.close()
