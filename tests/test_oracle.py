"""CPU tests of the ORACLE itself: C restatement vs the independent numpy/scipy twin, the committed
golden fixtures, a KKT optimality certificate at tight tolerance and a brute-force check on a tiny
problem.  (The reference ships no golden vectors: PARITY UNPINNED, see oracle/pqp_oracle.h.)"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle, twin
from path_optimizer_b200 import synth
from path_optimizer_b200.abi import SOLVED

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "kp_*.npz")))


def _dense(qp):
    P = sp.coo_matrix((qp["P"][2], (qp["P"][0], qp["P"][1])), shape=(qp["n"], qp["n"])).toarray()
    A = sp.coo_matrix((qp["A"][2], (qp["A"][0], qp["A"][1])), shape=(qp["m"], qp["n"])).toarray()
    return P, A


@pytest.mark.parametrize("gen,n", [(synth.straight_corridors, 12), (synth.curvy_corridors, 37)])
def test_kp_assembly_matches_dense_block_twin(oracle_params, gen, n):
    b = gen(2, n)
    for k in range(2):
        o0, o1 = b["offsets"][k], b["offsets"][k + 1]
        args = (b["ref"][o0:o1], b["bounds"][o0:o1], b["x0"][k], b["end_heading"][k])
        qp = oracle.assemble(oracle_params, 0, *args)
        H, q, A, l, u = twin.assemble_kp(oracle_params, *args)
        P, Ac = _dense(qp)
        keep = oracle.keep_control_steps(0, b["ref"][o0:o1])
        ch = (n + keep - 2) // keep
        assert qp["n"] == 5 * n + ch and qp["m"] == 11 * n + ch + 2  # solver_kp_as_input.cpp:18-23
        assert np.array_equal(np.triu(H), P)
        assert np.array_equal(A, Ac)
        assert np.array_equal(l, qp["l"]) and np.array_equal(u, qp["u"])
        assert np.all(qp["q"] == 0)  # solver.cpp:54


@pytest.mark.parametrize("form", [1, 2])
@pytest.mark.parametrize("gen,n", [(synth.straight_corridors, 9), (synth.curvy_corridors, 38)])
def test_k_and_kpc_assembly_match_dense_block_twins(oracle_params, form, gen, n):
    """SolverKAsInput / SolverKpAsInputConstrained: the oracle's sparse triplets against an independent numpy restatement
    written in the reference's own dense-block style (sparseView drops the exact zeros of the dense matrix)."""
    b = gen(2, n)
    total = 2 * n
    v = 3.0 + 2.0 * np.sin(np.arange(total) * 0.07)
    a = 0.8 * np.cos(np.arange(total) * 0.05)
    v[3] = 0.0                                   # standstill: DBL_MAX limits
    ref_all = b["ref"].copy()
    ref_all["v"], ref_all["a"] = v, a
    mk_all, mkp_all = oracle.update_limits(oracle_params, ref_all)
    for k in range(2):
        o0, o1 = b["offsets"][k], b["offsets"][k + 1]
        args = (b["ref"][o0:o1], b["bounds"][o0:o1], b["x0"][k], b["end_heading"][k])
        if form == 1:
            qp = oracle.assemble(oracle_params, 1, *args)
            H, q, A, l, u = twin.assemble_k(oracle_params, *args)
            assert qp["n"] == 4 * n - 1 and qp["m"] == 11 * n - 1           # solver_k_as_input.cpp:18-19
        else:
            mk, mkp = np.ascontiguousarray(mk_all[o0:o1]), np.ascontiguousarray(mkp_all[o0:o1])
            qp = oracle.assemble(oracle_params, 2, *args, max_k=mk, max_kp=mkp)
            H, q, A, l, u = twin.assemble_kpc(oracle_params, *args, mk, mkp)
            ch = (n + 2) // 4
            assert qp["n"] == 6 * n + ch and qp["m"] == 12 * n + 3 * ch + 2  # solver_kp_as_input_constrained.cpp:18-24
        P, Ac = _dense(qp)
        assert np.array_equal(np.triu(H), P)
        assert np.array_equal(A, Ac)
        # bounds: OSQP clips nothing at assembly, DBL_MAX limits pass through as they are
        assert np.array_equal(l, qp["l"]) and np.array_equal(u, qp["u"])
        assert np.all(qp["q"] == 0)


@pytest.mark.parametrize("form", [1, 2])
def test_k_and_kpc_oracle_vs_twin_solve(oracle_params, form):
    """Same iteration counts and iterates from the C oracle and the scipy twin on the K / KPC problems."""
    b = synth.curvy_corridors(1, 30)
    args = (b["ref"], b["bounds"], b["x0"][0], b["end_heading"][0])
    if form == 1:
        r = oracle.solve_qp(oracle_params, 1, *args)
        H, q, A, l, u = twin.assemble_k(oracle_params, *args)
    else:
        ref = b["ref"].copy()
        ref["v"] = 5.0
        mk, mkp = oracle.update_limits(oracle_params, ref)
        r = oracle.solve_qp(oracle_params, 2, *args, max_k=mk, max_kp=mkp)
        H, q, A, l, u = twin.assemble_kpc(oracle_params, *args, mk, mkp)
    t = twin.osqp_twin(oracle_params, H, q, A, l, u)
    assert r["info"].status == t["status"] and r["info"].iters == t["iters"]
    assert r["info"].rho_updates == t["rho_updates"]
    np.testing.assert_allclose(r["x"], t["x"], rtol=0, atol=1e-9)


def test_keep_control_steps_is_3_for_accumulated_03(oracle_params):
    # stations accumulated by += 0.3 give int(1.2 / 0.30000000000000004) = 3
    b = synth.straight_corridors(1, 30)
    assert oracle.keep_control_steps(0, b["ref"]) == 3
    assert twin.keep_control_steps(b["ref"]) == 3
    assert oracle.keep_control_steps(2, b["ref"]) == 4  # KPC fixes 4
    assert oracle.keep_control_steps(1, b["ref"]) == 1


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(oracle_params, path):
    g = np.load(path)
    batch = dict(n_points=g["n_points"], ref=g["ref"], bounds=g["bounds"], x0=g["x0"],
                 end_heading=g["end_heading"])
    batch["offsets"] = np.concatenate([[0], np.cumsum(g["n_points"])]).astype(np.int32)
    res = oracle.solve_batch(oracle_params, 0, batch)
    assert np.array_equal(res["status"], g["status"])
    assert np.array_equal(res["iters"], g["iters"])
    np.testing.assert_allclose(res["frenet"], g["frenet"], rtol=0, atol=1e-12)
    # the twin (different linear algebra, different assembly style) agrees too
    assert np.array_equal(g["twin_iters"], g["iters"])
    np.testing.assert_allclose(g["twin_frenet"], g["frenet"], rtol=0, atol=1e-9)
    for f in "xyzks":
        np.testing.assert_allclose(res["states"][f], g["states"][f], rtol=0, atol=1e-12)


def test_oracle_vs_twin_iterates(oracle_params):
    b = synth.curvy_corridors(1, 40)
    args = (b["ref"], b["bounds"], b["x0"][0], b["end_heading"][0])
    r = oracle.solve_qp(oracle_params, 0, *args, trace_rows=6)
    H, q, A, l, u = twin.assemble_kp(oracle_params, *args)
    t = twin.osqp_twin(oracle_params, H, q, A, l, u, trace_every=1)
    assert r["info"].status == SOLVED and t["status"] == SOLVED
    assert r["info"].iters == t["iters"] and r["info"].rho_updates == t["rho_updates"]
    k = min(6, len(t["trace"]))
    np.testing.assert_allclose(r["trace"][:k], t["trace"][:k], rtol=0, atol=1e-10)
    np.testing.assert_allclose(r["x"], t["x"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(r["y"], t["y"], rtol=0, atol=1e-8)


def test_kkt_certificate_at_tight_tolerance(oracle_params):
    """Run the recurrence to eps = 1e-9: the result must satisfy the QP's optimality conditions
    (this validates the recurrence itself without a second solver)."""
    p = oracle_params.copy()
    p.eps_abs = p.eps_rel = 1e-9
    p.max_iter = 200000
    b = synth.straight_corridors(1, 24)
    args = (b["ref"], b["bounds"], b["x0"][0], b["end_heading"][0])
    r = oracle.solve_qp(p, 0, *args)
    assert r["info"].status == SOLVED
    H, q, A, l, u = twin.assemble_kp(p, *args)
    cert = twin.kkt_certificate(H, q, A, l, u, r["x"], r["y"])
    assert cert["primal"] < 1e-7 and cert["dual"] < 1e-6 and cert["comp"] < 1e-6, cert
    # second (dead) slack block of the reference stays at 0 (solver_kp_as_input.cpp:21,56-57)
    n = 24
    keep = oracle.keep_control_steps(0, b["ref"])
    ch = (n + keep - 2) // keep
    assert np.all(np.abs(r["x"][3 * n + ch + n:]) < 1e-12)


def test_tiny_problem_against_scipy_trust_constr(oracle_params):
    """N = 3: an independent general-purpose solver finds the same optimum (Tier-2 cross-check).
    The KP objective does not penalise e_y (KP_deviation_weight = 0), so compare the objective and
    feasibility rather than the argmin."""
    from scipy.optimize import LinearConstraint, minimize
    p = oracle_params.copy()
    p.eps_abs = p.eps_rel = 1e-9
    p.max_iter = 400000
    b = synth.straight_corridors(1, 3)
    args = (b["ref"], b["bounds"], b["x0"][0], b["end_heading"][0])
    r = oracle.solve_qp(p, 0, *args)
    assert r["info"].status == SOLVED
    H, q, A, l, u = twin.assemble_kp(p, *args)
    lo = np.where(l < -1e29, -np.inf, l)
    hi = np.where(u > 1e29, np.inf, u)
    res = minimize(lambda x: 0.5 * x @ H @ x, np.zeros(len(q)), jac=lambda x: H @ x, hess=lambda x: H,
                   constraints=[LinearConstraint(A, lo, hi)], method="trust-constr",
                   options=dict(gtol=1e-12, xtol=1e-14, maxiter=3000))
    obj_or = 0.5 * r["x"] @ H @ r["x"]
    # trust-constr is an interior-point method that stops at a barrier parameter of ~1e-6, so its
    # objective sits slightly ABOVE the optimum; the oracle must match it to ~1e-5 and not exceed it.
    assert abs(obj_or - res.fun) <= 1e-5 * max(1.0, abs(res.fun))
    assert obj_or <= res.fun + 1e-9
    Ax = A @ r["x"]
    assert np.all(Ax <= hi + 1e-7) and np.all(Ax >= lo - 1e-7)


def test_invalid_bounds_status(oracle_params):
    b = synth.straight_corridors(1, 10)
    b["bounds"]["c0_lb"][3] = 2.0  # lb > ub: osqp_setup refuses -> reference solve() returns false
    res = oracle.solve_batch(oracle_params, 0, b)
    assert res["status"][0] == -100
    assert np.all(np.isnan(res["frenet"]))


def test_k_and_kpc_formulations_solve(oracle_params):
    b = synth.curvy_corridors(1, 30)
    args = (b["ref"], b["bounds"], b["x0"][0], b["end_heading"][0])
    r = oracle.solve_qp(oracle_params, 1, *args)
    assert r["n"] == 4 * 30 - 1 and r["m"] == 11 * 30 - 1  # solver_k_as_input.cpp:18-19
    assert r["info"].status in (SOLVED, -2)
    ch = (30 + 4 - 2) // 4
    max_k = np.full(30, 0.15)
    max_kp = np.full(ch, 0.1)
    r = oracle.solve_qp(oracle_params, 2, *args, max_k=max_k, max_kp=max_kp)
    assert r["n"] == 6 * 30 + ch and r["m"] == 12 * 30 + 3 * ch + 2  # ..._constrained.cpp:18-24
