"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* pair sharding (config 5): pair i -> rank i % world, no collective on the data path.
* disparity-range split of one pair: rank r evaluates slices [r*D/N, (r+1)*D/N), the per-pixel 64-bit WTA keys
  (48-bit orderable cost << 16 | d, see csrc/asw_common.cuh) are MIN-all-reduced and converted to a map.
  torch.distributed has no unsigned 64-bit MIN, so the keys travel as int64 with the top bit flipped
  (order-preserving map u64 -> i64); NCCL on GPUs, gloo in the CPU tests.
"""
import numpy as np

_FLIP = np.uint64(1 << 63)


def shard_pairs(n_pairs, rank, world):
    """indices of the pairs owned by `rank` (round-robin, as SURVEY 8e-1)"""
    return list(range(rank, n_pairs, world))


def split_range(num_disparity, rank, world):
    """[d_begin, d_end) of rank; ranges tile [0, D) exactly, early ranks take the remainder"""
    base, rem = divmod(num_disparity, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def keys_to_i64(keys_u64):
    return (np.asarray(keys_u64, dtype=np.uint64) ^ _FLIP).view(np.int64)


def keys_from_i64(keys_i64):
    return np.asarray(keys_i64, dtype=np.int64).view(np.uint64) ^ _FLIP


def allreduce_min_keys(keys_u64, device=None):
    """MIN all-reduce of a [H][W] uint64 key map over the default process group; returns uint64 numpy"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(keys_to_i64(keys_u64).copy())
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return keys_from_i64(t.cpu().numpy())


def split_stereo_matching(ctx, L, R, algorithm, disp_type, win, min_d, num_d, rank, world, device=None):
    """disparity-split of one pair across the process group; every rank returns the full disparity map"""
    from .api import method_candidates
    # the candidates the method really scans: numDisparity + 1 for traditional / geodesic / grid (loop `<=`)
    lo, hi = split_range(method_candidates(algorithm, num_d), rank, world)
    keys, _ = ctx.split_local_keys(L, R, algorithm, disp_type, win, min_d, num_d, lo, hi)
    merged = allreduce_min_keys(keys, device)
    return ctx.keys_to_disparity(merged)
