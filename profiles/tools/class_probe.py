"""Diagnostic: device-resident kernel time per shape class (CUDA events, as bench.py's device arm).

  python profiles/tools/class_probe.py [--paths B] [--kind curvy|straight] N [N ...]

Prints one JSON line per N: kernel class, ms per launch, iterations, ns per station-iteration.
PQP_FORCE_SMEM=<bytes> (library diagnostics hook) raises the dynamic shared memory of every launch, e.g. to
hold a two-CTA-per-SM class at one CTA per SM."""
import argparse
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from path_optimizer_b200 import _lib, synth
from path_optimizer_b200.abi import STATE_DTYPE
from path_optimizer_b200.solver import BatchPathSolver

ap = argparse.ArgumentParser()
ap.add_argument("--paths", type=int, default=1024)
ap.add_argument("--kind", default="curvy")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("n", type=int, nargs="+")
args = ap.parse_args()
L = _lib.load()
dev = torch.device("cuda", 0)
B = args.paths
for n in args.n:
    b = synth.curvy_corridors(B, n) if args.kind == "curvy" else synth.straight_corridors(B, n)
    total = B * n
    s = BatchPathSolver(max_batch=B, max_total_points=total)
    keep = int(L.pqp_keep_control_steps(0, np.ascontiguousarray(b["ref"][:n]).ctypes.data_as(C.c_void_p), n))
    up = lambda a: torch.from_numpy(np.frombuffer(np.ascontiguousarray(a).tobytes(), dtype=np.uint8).copy()).to(dev)  # noqa: E731
    d = {k: up(b[k]) for k in ("n_points", "offsets", "ref", "bounds", "x0", "end_heading")}
    d_out = torch.zeros(total * STATE_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_status = torch.zeros(B, dtype=torch.int32, device=dev)
    d_iters = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)

    def step():
        rc = L.pqp_solve_batch_device(s._h, 0, B, total, n, keep, keep, d["n_points"].data_ptr(), d["offsets"].data_ptr(),
                                      d["ref"].data_ptr(), d["bounds"].data_ptr(), d["x0"].data_ptr(),
                                      d["end_heading"].data_ptr(), None, None, d_out.data_ptr(), None,
                                      d_status.data_ptr(), d_iters.data_ptr(), C.c_void_p(st.cuda_stream), None)
        assert rc == 0, _lib.last_error()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.reps):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    it = d_iters.cpu().numpy().astype(np.float64)
    ms = e0.elapsed_time(e1) / args.reps
    v, t, sm = C.c_int(), C.c_int(), C.c_int64()
    L.pqp_device_class_info(n, keep, keep, 0, C.byref(v), C.byref(t), C.byref(sm))
    # balanced-work estimate: with S resident CTA slots the launch cannot end before sum(iters) / S slot-iterations
    print(json.dumps({"N": n, "keep": keep, "paths": B, "kind": args.kind, "class": L.pqp_class_name(v.value).decode(),
                      "threads": t.value, "smem": sm.value, "force_smem": os.environ.get("PQP_FORCE_SMEM"),
                      "ms": round(ms, 3), "iters_mean": it.mean(), "iters_max": it.max(),
                      "ns_per_station_iteration": round(ms * 1e6 / (it.sum() * n), 4),
                      "solved": float((d_status.cpu().numpy() == 1).mean())}), flush=True)
    s.close()
