"""Host-side mirror of include/pqp_multi.h: the library's own multi-GPU entry points (C++ orchestration + NCCL inside
libpqp.so).  This module only marshals buffers."""
import ctypes as C

import numpy as np

from . import _lib
from .abi import BOUNDS_DTYPE, OK, SOLVED, STATE_DTYPE, Stats, ptr
from .solver import PqpError, default_params


class MultiGpuSolver:
    """One process, several GPUs (pqp_multi_create / pqp_multi_solve_batch)."""

    def __init__(self, devices, params=None, max_batch_per_device=8192, max_total_points_per_device=8192 * 200):
        self._L = _lib.load()
        self.params = params if params is not None else default_params()
        self.devices = np.ascontiguousarray(devices, dtype=np.int32)
        self._m = C.c_void_p()
        rc = self._L.pqp_multi_create(C.byref(self._m), C.byref(self.params), len(self.devices), ptr(self.devices),
                                      int(max_batch_per_device), int(max_total_points_per_device))
        if rc != OK:
            raise PqpError(f"pqp_multi_create failed (rc={rc}): {_lib.last_error()}")

    def close(self):
        if self._m:
            self._L.pqp_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, batch, gather=True, formulation="KP"):
        """Returns dict(states, frenet, status, iters, ok, stats, shards, gather_rows); the gathered device buffers stay
        inside the library (gathered_ptr(k))."""
        n_points = np.ascontiguousarray(batch["n_points"], dtype=np.int32)
        B, total = len(n_points), int(n_points.sum())
        ref = np.ascontiguousarray(batch["ref"], dtype=STATE_DTYPE)
        bounds = np.ascontiguousarray(batch["bounds"], dtype=BOUNDS_DTYPE)
        x0 = np.ascontiguousarray(batch["x0"], dtype=np.float64)
        end_heading = np.ascontiguousarray(batch["end_heading"], dtype=np.float64)
        states = np.zeros(total, dtype=STATE_DTYPE)
        frenet = np.zeros((total, 3))
        status = np.zeros(B, dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        stats = Stats()
        rc = self._L.pqp_multi_solve_batch(self._m, {"KP": 0, "K": 1}[formulation], B, ptr(n_points), ptr(ref), ptr(bounds), ptr(x0), ptr(end_heading),
                                           ptr(states), ptr(frenet), ptr(status), ptr(iters), int(bool(gather)), C.byref(stats))
        if rc != OK:
            raise PqpError(f"pqp_multi_solve_batch failed (rc={rc}): {_lib.last_error()}")
        shards = []
        for k in range(len(self.devices)):
            fp, npth, fs = C.c_int(), C.c_int(), C.c_int64()
            self._L.pqp_multi_shard(self._m, k, C.byref(fp), C.byref(npth), C.byref(fs))
            shards.append((fp.value, npth.value, fs.value))
        return dict(states=states, frenet=frenet, status=status, iters=iters, ok=(status == SOLVED), stats=stats,
                    shards=shards, gather_rows=int(self._L.pqp_multi_gather_rows(self._m)))

    def gathered_ptr(self, k):
        return self._L.pqp_multi_gathered(self._m, int(k))
