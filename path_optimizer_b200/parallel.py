"""Multi-GPU plumbing: paths are independent, so the batch shards with no data-path collective; the
only exchange is ONE all-gather of the solved Frenet states at the end (north_star / SURVEY 8e).

Pure host logic + torch.distributed calls (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_paths, world, rank):
    """Contiguous block [begin, end) of a uniform-length batch for `rank`."""
    begin = (n_paths * rank) // world
    end = (n_paths * (rank + 1)) // world
    return begin, end


def shard_by_work(n_points, world):
    """Mixed lengths (BASELINE config 5): work ~ N per path.  Longest-first greedy onto the least
    loaded rank, so ranks finish together.  Returns a list of index arrays (original path ids)."""
    n_points = np.asarray(n_points)
    order = np.argsort(-n_points, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    parts = [[] for _ in range(world)]
    for idx in order:
        r = int(np.argmin(load))
        parts[r].append(int(idx))
        load[r] += int(n_points[idx])
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def gather_frenet(local, world, out=None):
    """One all-gather of the fixed-stride result tensor [B_local, N, 3] -> [world * B_local, N, 3]."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1))
    return out


def gather_padded(local, n_points_local, n_max, world):
    """Mixed lengths: pad each path's [N_i, 3] block to [n_max, 3] and all-gather with the lengths."""
    import torch
    import torch.distributed as dist
    B = len(n_points_local)
    padded = torch.zeros((B, n_max, 3), dtype=local.dtype, device=local.device)
    off = 0
    for b, n in enumerate(n_points_local):
        padded[b, :n] = local[off:off + n]
        off += n
    lens = torch.tensor(list(n_points_local), dtype=torch.int32, device=local.device)
    if world == 1:
        return padded, lens
    out = torch.empty((world * B, n_max, 3), dtype=local.dtype, device=local.device)
    out_l = torch.empty(world * B, dtype=torch.int32, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), padded.view(-1))
    dist.all_gather_into_tensor(out_l, lens)
    return out, out_l
