"""ctypes loader for the CPU oracle (``oracle/_build/libpqp_oracle.so``).

TEST INFRASTRUCTURE, not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` leg may import this.  PARITY UNPINNED (see pqp_oracle.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from path_optimizer_b200.abi import BOUNDS_DTYPE, STATE_DTYPE, Params, ptr  # noqa: E402

_LIB_PATH = os.path.join(_HERE, "_build", "libpqp_oracle.so")


class OqpInfo(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("rho_updates", C.c_int),
                ("rho_final", C.c_double), ("pri_res", C.c_double), ("dua_res", C.c_double),
                ("obj_val", C.c_double), ("kkt_n", C.c_int), ("kkt_lnz", C.c_int)]


class OqpProblem(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("p_nnz", C.c_int), ("a_nnz", C.c_int),
                ("p_i", C.POINTER(C.c_int)), ("p_j", C.POINTER(C.c_int)), ("p_v", C.POINTER(C.c_double)),
                ("a_i", C.POINTER(C.c_int)), ("a_j", C.POINTER(C.c_int)), ("a_v", C.POINTER(C.c_double)),
                ("q", C.POINTER(C.c_double)), ("l", C.POINTER(C.c_double)), ("u", C.POINTER(C.c_double))]


def build(force=False):
    """(Re)build the oracle library with the committed Makefile when missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "pqp.h"))
    stale = (not os.path.exists(_LIB_PATH)
             or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs))
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_params_default.argtypes = [C.POINTER(Params)]
        L.oracle_params_default.restype = None
        L.oracle_keep_control_steps.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.oracle_keep_control_steps.restype = C.c_int
        L.oracle_assemble.restype = C.POINTER(OqpProblem)
        L.oracle_assemble.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_problem_free.argtypes = [C.POINTER(OqpProblem)]
        L.oracle_problem_free.restype = None
        L.oracle_osqp_solve.argtypes = [C.POINTER(Params), C.POINTER(OqpProblem), C.c_void_p,
                                        C.c_void_p, C.POINTER(OqpInfo), C.c_void_p, C.c_int]
        L.oracle_osqp_solve.restype = C.c_int
        L.oracle_extract.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_extract.restype = None
        L.oracle_solve_path.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.POINTER(OqpInfo)]
        L.oracle_solve_path.restype = C.c_int
        L.oracle_solve_batch.argtypes = [C.POINTER(Params), C.c_int, C.c_int] + [C.c_void_p] * 11 + [C.c_int]
        L.oracle_solve_batch.restype = C.c_double
        _lib = L
    return _lib


def default_params():
    p = Params()
    lib().oracle_params_default(C.byref(p))
    return p


def keep_control_steps(formulation, ref):
    ref = np.ascontiguousarray(ref, dtype=STATE_DTYPE)
    return lib().oracle_keep_control_steps(formulation, ptr(ref), len(ref))


def assemble(params, formulation, ref, bounds, x0, end_heading, max_k=None, max_kp=None):
    """Assembled QP of one path as numpy arrays (copies): dict(n, m, P=(i,j,v) upper, A=(i,j,v), q, l, u)."""
    ref = np.ascontiguousarray(ref, dtype=STATE_DTYPE)
    bounds = np.ascontiguousarray(bounds, dtype=BOUNDS_DTYPE)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    qp = lib().oracle_assemble(C.byref(params), formulation, len(ref), ptr(ref), ptr(bounds), ptr(x0),
                               float(end_heading), ptr(max_k), ptr(max_kp))
    if not qp:
        raise ValueError("oracle_assemble refused the problem")
    q = qp.contents
    arr = lambda p, n, t: np.ctypeslib.as_array(p, shape=(n,)).astype(t).copy()  # noqa: E731
    out = dict(n=q.n, m=q.m,
               P=(arr(q.p_i, q.p_nnz, np.int64), arr(q.p_j, q.p_nnz, np.int64), arr(q.p_v, q.p_nnz, np.float64)),
               A=(arr(q.a_i, q.a_nnz, np.int64), arr(q.a_j, q.a_nnz, np.int64), arr(q.a_v, q.a_nnz, np.float64)),
               q=arr(q.q, q.n, np.float64), l=arr(q.l, q.m, np.float64), u=arr(q.u, q.m, np.float64))
    lib().oracle_problem_free(qp)
    return out


def solve_qp(params, formulation, ref, bounds, x0, end_heading, max_k=None, max_kp=None, trace_rows=0):
    """Assemble + OSQP-restatement solve of one path; returns dict(x, y, info, trace)."""
    ref = np.ascontiguousarray(ref, dtype=STATE_DTYPE)
    bounds = np.ascontiguousarray(bounds, dtype=BOUNDS_DTYPE)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    qp = lib().oracle_assemble(C.byref(params), formulation, len(ref), ptr(ref), ptr(bounds), ptr(x0),
                               float(end_heading), ptr(max_k), ptr(max_kp))
    if not qp:
        raise ValueError("oracle_assemble refused the problem")
    n, m = qp.contents.n, qp.contents.m
    x = np.zeros(n)
    y = np.zeros(m)
    info = OqpInfo()
    trace = np.zeros((trace_rows, n)) if trace_rows else None
    lib().oracle_osqp_solve(C.byref(params), qp, ptr(x), ptr(y), C.byref(info), ptr(trace), trace_rows)
    lib().oracle_problem_free(qp)
    return dict(x=x, y=y, info=info, trace=trace, n=n, m=m)


def solve_batch(params, formulation, batch, threads=1, max_k=None, max_kp=None):
    """Batch driver (same layout as pqp_solve_batch).  Returns dict(states, frenet, status, iters, seconds)."""
    B = len(batch["n_points"])
    total = int(batch["offsets"][-1])
    ref = np.ascontiguousarray(batch["ref"], dtype=STATE_DTYPE)
    bounds = np.ascontiguousarray(batch["bounds"], dtype=BOUNDS_DTYPE)
    out = np.zeros(total, dtype=STATE_DTYPE)
    frenet = np.zeros((total, 3))
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    secs = lib().oracle_solve_batch(C.byref(params), formulation, B, ptr(batch["n_points"]), ptr(ref),
                                    ptr(bounds), ptr(batch["x0"]), ptr(batch["end_heading"]),
                                    ptr(max_k), ptr(max_kp), ptr(out), ptr(frenet), ptr(status),
                                    ptr(iters), int(threads))
    return dict(states=out, frenet=frenet, status=status, iters=iters, seconds=secs)
