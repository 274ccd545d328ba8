"""Generates tests/golden/*.npz: small KP/K/KPC problems with the C oracle's outputs and the numpy
twin's outputs side by side.  Run from the repo root:  python tests/golden/make_golden.py

The reference ships no golden vectors for this path and no OSQP binary exists here,
so these fixtures pin the ORACLE (C restatement) against the independent twin, and pin the CUDA path
against both.  PARITY UNPINNED with respect to a real OSQP build.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, twin  # noqa: E402
from path_optimizer_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def kp_case(name, batch):
    p = oracle.default_params()
    res = oracle.solve_batch(p, 0, batch)
    tw_x, tw_it = [], []
    for b in range(len(batch["n_points"])):
        o0, o1 = batch["offsets"][b], batch["offsets"][b + 1]
        H, q, A, l, u = twin.assemble_kp(p, batch["ref"][o0:o1], batch["bounds"][o0:o1], batch["x0"][b],
                                         batch["end_heading"][b])
        t = twin.osqp_twin(p, H, q, A, l, u)
        n = o1 - o0
        tw_x.append(t["x"][:3 * n].reshape(n, 3))
        tw_it.append(t["iters"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), n_points=batch["n_points"], ref=batch["ref"],
                        bounds=batch["bounds"], x0=batch["x0"], end_heading=batch["end_heading"],
                        frenet=res["frenet"], states=res["states"], status=res["status"], iters=res["iters"],
                        twin_frenet=np.concatenate(tw_x), twin_iters=np.array(tw_it))
    print(name, "iters", res["iters"].tolist(), "twin", tw_it,
          "max|oracle-twin|", float(np.abs(res["frenet"] - np.concatenate(tw_x)).max()))


if __name__ == "__main__":
    kp_case("kp_straight_n16", synth.straight_corridors(4, 16))
    kp_case("kp_straight_n100", synth.straight_corridors(3, 100))
    kp_case("kp_curvy_n50", synth.curvy_corridors(4, 50))
    kp_case("kp_mixed", synth.curvy_corridors(6, n_points=[2, 3, 8, 33, 64, 97]))
