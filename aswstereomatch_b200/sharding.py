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


class _DevKeys:
    """a device key buffer as a torch-readable CUDA array (int64 view of the u64 keys)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def allreduce_min_keys_device(ctx, dk, H, W, device):
    """MIN all-reduce of the ctx's device-resident u64 keys over the default process group (NCCL), IN PLACE: the keys never
    leave device memory.  torch.distributed has no unsigned 64-bit MIN, so the sign bit is flipped on the device before
    and after (asw_keys_flip_sign: order-preserving u64 <-> i64) and NCCL reduces the buffer through an int64 alias."""
    import torch
    import torch.distributed as dist
    ctx.keys_flip_sign(dk, H, W)
    ctx.sync()                                   # the library's stream -> torch's stream
    t = torch.as_tensor(_DevKeys(dk.value, H * W), device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    e1.record()
    torch.cuda.synchronize(device)
    ctx.keys_flip_sign(dk, H, W)
    return e0.elapsed_time(e1)               # device time of the collective on this rank, ms


def split_stereo_matching(ctx, L, R, algorithm, disp_type, win, min_d, num_d, rank, world, device=None):
    """disparity-split of one pair across the process group; every rank returns the full disparity map.
    With a CUDA `device` the keys are reduced in device memory (NCCL, no host bounce); without one (gloo, CPU tests)
    they travel through the host."""
    from .api import method_candidates
    # the candidates the method really scans: numDisparity + 1 for traditional / geodesic / grid (loop `<=`)
    lo, hi = split_range(method_candidates(algorithm, num_d), rank, world)
    if device is not None:
        dk, (H, W) = ctx.split_local_keys_device(L, R, algorithm, disp_type, win, min_d, num_d, lo, hi)
        split_stereo_matching.last_allreduce_ms = allreduce_min_keys_device(ctx, dk, H, W, device)
        return ctx.device_keys_to_disparity(dk, H, W)
    keys, _ = ctx.split_local_keys(L, R, algorithm, disp_type, win, min_d, num_d, lo, hi)
    merged = allreduce_min_keys(keys, None)
    return ctx.keys_to_disparity(merged)


split_stereo_matching.last_allreduce_ms = None
