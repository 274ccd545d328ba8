"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, work-balanced sharding,
and the single all-gather of solved states.  The per-rank solve is stood in for by the oracle on a
tiny shard (the GPU kernels cannot run here); what is under test is that shards regenerate their
inputs independently (counter-based RNG) and that the gathered tensor equals the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from path_optimizer_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_paths, n_points, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    begin, end = parallel.shard_range(n_paths, world, rank)
    shard = synth.straight_corridors(end - begin, n_points, first_path=begin)   # regenerated locally
    res = oracle.solve_batch(oracle.default_params(), 0, shard)
    local = torch.from_numpy(res["frenet"].reshape(end - begin, n_points, 3).copy())
    full = parallel.gather_frenet(local, world)
    pad, lens = parallel.gather_padded(torch.from_numpy(res["frenet"].copy()), [n_points] * (end - begin), n_points + 3, world)
    if rank == 0:
        q.put((full.numpy(), pad.numpy(), lens.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_gather():
    from oracle import oracle
    n_paths, n_points, world = 6, 12, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_paths, n_points, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, pad, lens = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = oracle.solve_batch(oracle.default_params(), 0, synth.straight_corridors(n_paths, n_points))
    np.testing.assert_array_equal(full.reshape(-1, 3), single["frenet"])      # bit-identical: same inputs per path
    np.testing.assert_array_equal(pad[:, :n_points].reshape(-1, 3), single["frenet"])
    assert np.all(pad[:, n_points:] == 0) and np.all(lens == n_points)


def test_shard_range_covers_everything():
    for n, w in [(1024, 1), (1024, 8), (65536, 8), (7, 4), (3, 8)]:
        got = []
        for r in range(w):
            b, e = parallel.shard_range(n, w, r)
            got += list(range(b, e))
        assert got == list(range(n))


def test_shard_by_work_balances_mixed_lengths():
    rng = np.random.default_rng(0)
    n_points = rng.integers(50, 401, size=16384)           # BASELINE config 5
    parts = parallel.shard_by_work(n_points, 4)
    assert sorted(np.concatenate(parts).tolist()) == list(range(16384))
    loads = np.array([n_points[p].sum() for p in parts])
    assert loads.max() - loads.min() <= 400                # within one path of perfect balance


def test_counter_based_rng_is_shard_independent():
    a = synth.straight_corridors(8, 10)
    b = synth.straight_corridors(3, 10, first_path=5)
    np.testing.assert_array_equal(a["x0"][5:], b["x0"])
    np.testing.assert_array_equal(a["bounds"][50:], b["bounds"])
