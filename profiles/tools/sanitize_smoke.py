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
# round 2: "K" classes (SPIKE form 4 / 8 warps, cyclic reduction 13 warps), "KPC" classes, long KP classes, the order kernel
for nn in ([30, 2, 100, 128], [129, 200], [256], [300]):
    r = pl.solve(synth.curvy_corridors(len(nn), n_points=nn), 'K'); print('K', nn, r['status'], r['iters'])
r = pl.solve(synth.infeasible_corridors(2, 40), 'K'); print('K infeasible', r['status'], r['iters'])
for nn in ([60, 100], [200]):
    bk = synth.curvy_corridors(len(nn), n_points=nn)
    tot = int(bk['offsets'][-1])
    bk['ref']['v'] = 4.0 + 3.0 * np.sin(np.arange(tot) * 0.05); bk['ref']['a'] = 0.5 * np.cos(np.arange(tot) * 0.05)
    mk, mkp = planner.update_limits(pl.params, bk['ref'])
    r = pl.solve(bk, 'KPC', max_k=mk, max_kp=mkp); print('KPC', nn, r['status'], r['iters'])
r = pl.solve(synth.curvy_corridors(1, 384)); print('KP 384', r['status'], r['iters'])
import ctypes as C, torch
from path_optimizer_b200 import _lib
from path_optimizer_b200.abi import STATE_DTYPE
L = _lib.load()
pb = planner.PathPlanner(max_batch=256, max_total_points=256 * 40)
bo = synth.curvy_corridors(200, 40)
def up(a): return torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype=np.uint8).copy()).cuda()
d = [up(bo[k]) for k in ('n_points', 'offsets', 'ref', 'bounds', 'x0', 'end_heading')]
d_out = torch.zeros(200 * 40 * STATE_DTYPE.itemsize, dtype=torch.uint8, device='cuda'); d_st = torch.zeros(200, dtype=torch.int32, device='cuda')
torch.cuda.synchronize()
rc = L.pqp_solve_batch_device(pb._h, 0, 200, 8000, 40, 3, 3, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(),
                              d[5].data_ptr(), None, None, d_out.data_ptr(), None, d_st.data_ptr(), None, None, None)
torch.cuda.synchronize(); print('device entry with the order kernel', rc, int((d_st.cpu().numpy() == 1).sum()), 'of 200 solved')
pb.close()
field = synth.disc_field_map(rows=400, cols=150)
pl.set_map(field)
bm = synth.map_reference_paths(6, 60, x_range=(-30, 10))
spl = planner.reference_splines(bm)
for mode in (0, 1):
    for om in (0, 1):
        r = pl.plan(bm, bounds_mode=mode, splines=spl if mode == 0 else None, output_mode=om, max_out=128)
        print('plan', mode, om, r['status'], r['n_out'], r['ok'])
pl.close()
