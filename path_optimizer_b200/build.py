"""Builds libpqp.so (the C-ABI shared library with the sm_100a kernels) in-tree with nvcc."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libpqp.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "550"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(_HERE), "include", "pqp.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into path_optimizer_b200/libpqp.so."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", LIB_PATH]
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
