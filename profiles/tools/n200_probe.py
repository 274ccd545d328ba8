"""Diagnostic: device-resident kernel time (CUDA events, as bench.py's device arm) for 1024 paths of N stations."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from path_optimizer_b200 import _lib, synth
from path_optimizer_b200.abi import STATE_DTYPE
from path_optimizer_b200.solver import BatchPathSolver

L = _lib.load()
dev = torch.device("cuda", 0)
B = 1024
for n in [int(a) for a in sys.argv[1:]] or [200, 150]:
    b = synth.curvy_corridors(B, n)
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
    for _ in range(3):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    it = d_iters.cpu().numpy()
    print("N", n, "keep", keep, "kernel ms", round(e0.elapsed_time(e1) / 3, 3), "iters mean", it.mean(), "max", it.max(),
          "solved", float((d_status.cpu().numpy() == 1).mean()), flush=True)
    s.close()
