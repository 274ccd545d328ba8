"""CPU tests of the C-ABI library (no compute calls: there is no GPU here): it loads, exports every
symbol include/pqp.h declares, its host helpers agree with the oracle, and pqp_create fails loudly
without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle
from path_optimizer_b200 import _lib, build as pbuild, synth
from path_optimizer_b200.abi import Params, Stats, STATE_DTYPE, BOUNDS_DTYPE, ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    pbuild.build()
    return _lib.load()


def test_exports_every_declared_symbol(L):
    header = "".join(open(os.path.join(ROOT, "include", f)).read() for f in ("pqp.h", "pqp_env.h", "pqp_multi.h"))
    declared = set(re.findall(r"\b(pqp_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym)


def test_record_sizes():
    assert STATE_DTYPE.itemsize == 56 and BOUNDS_DTYPE.itemsize == 64
    # sizeof(pqp_params): 21 doubles + int (+pad) + 7 doubles + 5 ints (+pad) + double + 3 ints (+pad)
    assert C.sizeof(Params) == 21 * 8 + 8 + 7 * 8 + 24 + 8 + 16
    assert C.sizeof(Stats) == 56
    from path_optimizer_b200.abi import DistanceMap
    assert C.sizeof(DistanceMap) == 8 + 8 + 3 * 8


def test_host_spline_helpers_match_oracle(L):
    """pqp_spline_fit / pqp_spline_eval are host helpers (no GPU needed): same coefficients and values as the
    oracle's restatement of tk::spline, including both extrapolation sides."""
    from path_optimizer_b200 import planner
    rng = np.random.default_rng(3)
    t = np.cumsum(rng.uniform(0.2, 0.6, 40))
    y = np.sin(t) + 0.1 * rng.standard_normal(40)
    c = planner.spline_fit(t, y)
    co = oracle.spline_fit(t, y)
    assert np.abs(c - co).max() <= 1e-12
    at = np.concatenate([[t[0] - 1.0, t[0], t[-1], t[-1] + 2.0], rng.uniform(t[0], t[-1], 50), t[5:8]])
    for order in (0, 1, 2):
        assert np.abs(planner.spline_eval(t, c, at, order) - oracle.spline_eval(t, co, at, order)).max() <= 1e-12
    with pytest.raises(Exception):
        planner.spline_fit(t[:2], y[:2])


def test_env_calls_fail_without_gpu(L):
    """Every map-based entry point needs a handle; creating one without a GPU fails, there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from path_optimizer_b200 import planner
    with pytest.raises(Exception):
        planner.PathPlanner(max_batch=1, max_total_points=64)


def test_params_default_matches_oracle_bytes(L):
    p = Params()
    assert L.pqp_params_default(C.byref(p)) == 0
    o = oracle.default_params()
    assert bytes(p) == bytes(o)
    assert abs(p.d1 - (-0.3875)) < 1e-12 and abs(p.d4 - 3.2875) < 1e-12  # SURVEY 8 conventions
    assert abs(p.circle_radius - 1.1727) < 1e-4


def test_keep_and_sizes(L):
    b = synth.straight_corridors(1, 100)
    ref = np.ascontiguousarray(b["ref"], dtype=STATE_DTYPE)
    assert L.pqp_keep_control_steps(0, ptr(ref), 100) == oracle.keep_control_steps(0, ref) == 3
    assert L.pqp_keep_control_steps(2, ptr(ref), 100) == 4
    nv, nc = C.c_int(), C.c_int()
    assert L.pqp_problem_size(0, 100, 3, C.byref(nv), C.byref(nc)) == 0
    assert (nv.value, nc.value) == (533, 1135)
    assert L.pqp_problem_size(1, 100, 1, C.byref(nv), C.byref(nc)) == 0
    assert (nv.value, nc.value) == (399, 1099)
    assert L.pqp_problem_size(2, 100, 4, C.byref(nv), C.byref(nc)) == 0
    assert (nv.value, nc.value) == (625, 1277)
    assert L.pqp_problem_size(7, 100, 4, C.byref(nv), C.byref(nc)) != 0


def test_create_fails_loudly_without_gpu(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    p = Params()
    L.pqp_params_default(C.byref(p))
    rc = L.pqp_create(C.byref(h), C.byref(p), 0, 4, 400)
    assert rc != 0 and not h.value
    assert _lib.last_error()


def test_python_wrapper_raises_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from path_optimizer_b200.solver import BatchPathSolver, PqpError, OsqpSolver
    with pytest.raises(PqpError):
        BatchPathSolver(max_batch=1, max_total_points=10)
    assert OsqpSolver.create("NOPE", None, None, (0, 0), 0, 0, 0) is None  # solver.cpp:41-43
