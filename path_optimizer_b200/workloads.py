"""BASELINE.json configs 2..5 as named workloads (shared by bench.py and the tests).

Every config is defined per GPU (weak scaling: a rank's shard has the same shape at every world size):

  2  1 024 paths x 100 stations, straight corridors                    (synth.straight_corridors)
  3  8 192 paths x 200 stations, clearance bounds from a random-obstacle distance map
     (synth.disc_field_map + synth.map_reference_paths; bounds by the clearance stage; blocked paths dropped,
     reference_path_impl.cpp:263-269)
  4  8 192 paths x 100 stations per GPU (65 536 over 8), straight corridors
  5  4 096 paths per GPU (16 384 over 4), N ~ U{50..400}, analytic curved corridors, ranks split by equal
     station count (parallel.shard_by_work)

Pure host code (numpy); the clearance bounds of config 3 are computed by whoever calls `config3_bounds`
with a bounds function (the GPU clearance kernel for the product arm, the oracle for the CPU arm -- the two
agree bit for bit, tests/test_gpu_env.py).
"""
import numpy as np

from . import synth
from .parallel import shard_by_work

CONFIGS = {
    2: dict(paths_per_gpu=1024, n_points=100, kind="straight",
            text="BASELINE config 2: 1024 paths x 100 stations per GPU, straight corridors"),
    3: dict(paths_per_gpu=8192, n_points=200, kind="map",
            text="BASELINE config 3: 8192 paths x 200 stations per GPU, clearance bounds from a random-obstacle "
                 "distance map (220 m x 50 m, 0.2 m cells, 300 discs), blocked paths dropped"),
    4: dict(paths_per_gpu=8192, n_points=100, kind="straight",
            text="BASELINE config 4: 8192 paths x 100 stations per GPU (65 536 over 8 GPUs), straight corridors"),
    5: dict(paths_per_gpu=4096, n_points=(50, 400), kind="mixed",
            text="BASELINE config 5: 4096 paths per GPU (16 384 over 4 GPUs), N ~ U{50..400} stations, analytic "
                 "curved corridors, ranks split by equal station count"),
}


FORMULATION_IDS = {"KP": 0, "K": 1, "KPC": 2}


def speed_profile(total):
    """Synthetic (v, a) per station for the "KPC" limits (ReferencePathImpl::updateLimits reads them from the reference states)."""
    t = np.arange(total)
    return 4.0 + 3.0 * np.sin(t * 0.05), 0.5 * np.cos(t * 0.05)


def describe(config, world=1, formulation="KP"):
    """The `config` object of a bench line: identical for the product arm and the CPU reference arm."""
    c = CONFIGS[config]
    d = {
        "workload": c["text"] + f"; {formulation} formulation, OSQP defaults (eps 1e-3, check every 25 it, adaptive rho every 25 it)"
                    + ("; curvature limits from a synthetic speed profile (v = 4 + 3 sin(0.05 i), a = 0.5 cos(0.05 i))" if formulation == "KPC" else ""),
        "formulation": formulation,
        "baseline_config": config,
        "paths_per_gpu": c["paths_per_gpu"],
        "n_points": c["n_points"] if isinstance(c["n_points"], int) else f"U{{{c['n_points'][0]}..{c['n_points'][1]}}}",
        "n_gpus": world,
        "seed": synth.BASE_SEED,
        "l2": "GPU arm: L2 flushed between timed steps (256 MiB fill); CPU arm: not applicable",
        "multi_gpu": "independent shards, one NCCL all-gather of the Frenet states per step (pqp_allgather: NCCL inside libpqp.so)" if world > 1 else "single GPU",
    }
    return d


def config3_candidates(count, first_path=0):
    """Reference lines for config 3 (no bounds yet)."""
    return synth.map_reference_paths(count, 200, first_path=first_path, config=3)


def config3_from_bounds(cand, bounds, n_valid, want):
    """Attach clearance bounds to the candidates and keep the first `want` unblocked paths."""
    keep = np.nonzero(np.asarray(n_valid) == cand["n_points"])[0][:want]
    b = dict(cand)
    b["bounds"] = np.ascontiguousarray(bounds)
    return synth.take_paths(b, keep), len(keep)


def build(config, rank=0, world=1, paths=None, bounds_fn=None):
    """This rank's shard of a config as a host batch (synth dict).  `paths` overrides the per-GPU path count
    (tests, bounded CPU samples).  Config 3 needs `bounds_fn(batch) -> (bounds, n_valid)`."""
    c = CONFIGS[config]
    B = int(paths or c["paths_per_gpu"])
    if c["kind"] == "straight":
        return synth.straight_corridors(B, c["n_points"], first_path=rank * B, config=config)
    if c["kind"] == "mixed":
        lo, hi = c["n_points"]
        lengths = synth.mixed_lengths(B * world, lo, hi, config=config)
        ids = shard_by_work(lengths, world)[rank]
        return synth.curvy_corridors(len(ids), n_points=lengths[ids], path_ids=ids, config=config)
    if c["kind"] == "map":
        assert bounds_fn is not None, "config 3 needs a clearance-bounds function"
        # a few per cent of the random reference lines run into a disc: draw spares, keep the first B unblocked
        spare = B + max(64, B // 16)
        cand = config3_candidates(spare, first_path=rank * 2 * c["paths_per_gpu"])
        bounds, n_valid = bounds_fn(cand)
        batch, got = config3_from_bounds(cand, bounds, n_valid, B)
        assert got == B, f"config 3: only {got} of {spare} candidate paths are unblocked"
        return batch
    raise ValueError(config)
