"""ctypes loader for the CPU warp-emulator build of the product's solver source (TEST HARNESS).

Builds tests/emu/libkp_emu.so from tests/emu/kp_emu.cpp with plain g++ (-DPQP_HOST_EMU) and runs
the SAME pqp_kp_core.cuh that nvcc compiles for sm_100a, one host thread per lane.  Used by the
CPU test-suite to check the kernel logic against the oracle without a GPU; never used by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from path_optimizer_b200.abi import BOUNDS_DTYPE, STATE_DTYPE, Params, ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libkp_emu.so")
_CSRC = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "path_optimizer_b200", "csrc")


def build():
    srcs = [os.path.join(_HERE, "kp_emu.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)
                                                  if f.endswith((".cuh", ".h"))]
    if (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                        os.path.join(_HERE, "kp_emu.cpp"), "-o", _LIB, "-lpthread"], check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.kp_emu_solve_batch.argtypes = [C.POINTER(Params), C.c_int] + [C.c_void_p] * 10 + [C.c_int, C.c_int, C.c_int]
        _lib.kp_emu_solve_batch.restype = C.c_int
        _lib.gen_emu_solve_batch.argtypes = [C.POINTER(Params), C.c_int, C.c_int] + [C.c_void_p] * 12
        _lib.gen_emu_solve_batch.restype = C.c_int
    return _lib


def solve_batch(params, batch, smem_bytes=227 * 1024, variant=0, nwarps=4):
    """variant 0 = generic core (pqp_kp_core.cuh); 1..4 = Kp2<17,6>, <10,7>, <27,7>, <49,7> run by a CTA of `nwarps` warps."""
    B = len(batch["n_points"])
    total = int(batch["offsets"][-1])
    ref = np.ascontiguousarray(batch["ref"], dtype=STATE_DTYPE)
    bounds = np.ascontiguousarray(batch["bounds"], dtype=BOUNDS_DTYPE)
    out = np.zeros(total, dtype=STATE_DTYPE)
    frenet = np.zeros((total, 3))
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    lib().kp_emu_solve_batch(C.byref(params), B, ptr(batch["n_points"]), ptr(batch["offsets"]), ptr(ref),
                             ptr(bounds), ptr(batch["x0"]), ptr(batch["end_heading"]), ptr(out),
                             ptr(frenet), ptr(status), ptr(iters), smem_bytes, variant, nwarps)
    return dict(states=out, frenet=frenet, status=status, iters=iters)


def solve_batch_generic(params, formulation, batch, max_k=None, max_kp=None):
    """The generic banded kernel source (pqp_gen_core.cuh) + the host assembly (pqp_forms.h) for the
    "K" (formulation 1) and "KPC" (2) formulations, under the warp emulator."""
    B = len(batch["n_points"])
    total = int(batch["offsets"][-1])
    ref = np.ascontiguousarray(batch["ref"], dtype=STATE_DTYPE)
    bounds = np.ascontiguousarray(batch["bounds"], dtype=BOUNDS_DTYPE)
    out = np.zeros(total, dtype=STATE_DTYPE)
    frenet = np.zeros((total, 3))
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    lib().gen_emu_solve_batch(C.byref(params), int(formulation), B, ptr(batch["n_points"]), ptr(batch["offsets"]), ptr(ref),
                              ptr(bounds), ptr(batch["x0"]), ptr(batch["end_heading"]), ptr(max_k), ptr(max_kp), ptr(out),
                              ptr(frenet), ptr(status), ptr(iters))
    return dict(states=out, frenet=frenet, status=status, iters=iters)
