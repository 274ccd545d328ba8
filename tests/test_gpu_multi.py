"""GPU tests of the library-level multi-GPU entry points (include/pqp_multi.h): shards + one NCCL all-gather inside
libpqp.so.  The single-device form runs on any GPU box; the two-device form skips unless two GPUs are visible."""
import ctypes as C

import numpy as np
import pytest

from path_optimizer_b200 import synth

pytestmark = pytest.mark.gpu


def _gathered(ms, k, n_dev, rows, device):
    import torch
    buf = torch.zeros(n_dev * rows * 3, dtype=torch.float64, device=f"cuda:{device}")
    torch.cuda.synchronize()
    # device-to-device copy of the library's gather buffer into a torch tensor
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaSetDevice(int(device)) == 0
    assert rt.cudaMemcpy(buf.data_ptr(), ms.gathered_ptr(k), n_dev * rows * 3 * 8, 4) == 0   # cudaMemcpyDefault
    return buf.cpu().numpy().reshape(n_dev, rows, 3)


@pytest.mark.parametrize("n_dev", [1, 2])
def test_multi_solver_matches_single_device(n_dev):
    import torch
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    from path_optimizer_b200.multi import MultiGpuSolver
    from path_optimizer_b200.solver import BatchPathSolver
    rng = np.random.default_rng(2)
    n_points = rng.integers(40, 260, size=300)
    batch = synth.curvy_corridors(300, n_points=n_points)
    single = BatchPathSolver(max_batch=300, max_total_points=int(n_points.sum()))
    ref = single.solve(batch)
    single.close()
    ms = MultiGpuSolver(list(range(n_dev)), max_batch_per_device=300, max_total_points_per_device=int(n_points.sum()))
    res = ms.solve(batch, gather=True)
    assert np.array_equal(res["status"], ref["status"]) and np.array_equal(res["iters"], ref["iters"])
    assert np.array_equal(res["frenet"], ref["frenet"])
    for f in "xyzks":
        assert np.array_equal(res["states"][f], ref["states"][f])
    assert res["stats"].kernel_launches >= n_dev and res["stats"].n_solved == int((ref["status"] == 1).sum())
    # shards are contiguous, cover the batch, and are balanced by station count
    shards = res["shards"]
    assert shards[0][0] == 0 and sum(s[1] for s in shards) == 300
    tot = int(n_points.sum())
    off = np.concatenate([[0], np.cumsum(n_points)])
    for k, (fp, npth, fs) in enumerate(shards):
        assert fs == off[fp]
        assert abs((off[fp + npth] - off[fp]) - tot / n_dev) <= 260
    # every device holds the Frenet states of the WHOLE batch after the one all-gather
    rows = res["gather_rows"]
    for k in range(n_dev):
        g = _gathered(ms, k, n_dev, rows, k)
        for j, (fp, npth, fs) in enumerate(shards):
            ns = int(off[fp + npth] - off[fp])
            assert np.array_equal(g[j, :ns], ref["frenet"][fs:fs + ns])
    # second call reuses the buffers
    again = ms.solve(batch, gather=True)
    assert np.array_equal(again["frenet"], ref["frenet"])
    # "K" shards the same way
    single = BatchPathSolver(max_batch=300, max_total_points=int(n_points.sum()))
    refk = single.solve(batch, "K")
    single.close()
    resk = ms.solve(batch, gather=True, formulation="K")
    assert np.array_equal(resk["status"], refk["status"]) and np.array_equal(resk["iters"], refk["iters"])
    assert np.array_equal(resk["frenet"], refk["frenet"])
    ms.close()


def test_per_rank_communicator_single_rank():
    """pqp_nccl_unique_id / pqp_comm_init_rank / pqp_allgather with one rank (the per-process form bench.py uses)."""
    import torch
    from path_optimizer_b200 import _lib
    from path_optimizer_b200.solver import BatchPathSolver
    L = _lib.load()
    s = BatchPathSolver(max_batch=4, max_total_points=400)
    uid = (C.c_char * 128)()
    assert L.pqp_nccl_unique_id(uid) == 0
    assert L.pqp_comm_init_rank(s._h, 1, 0, uid) == 0, _lib.last_error()
    a = torch.arange(30, dtype=torch.float64, device="cuda")
    b = torch.zeros(30, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    assert L.pqp_allgather(s._h, a.data_ptr(), b.data_ptr(), 30, None) == 0, _lib.last_error()
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert L.pqp_comm_destroy(s._h) == 0
    s.close()
