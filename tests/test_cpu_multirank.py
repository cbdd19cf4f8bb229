"""World-size-2 gloo tests (CPU): the host-side multi-GPU logic -- pair sharding and the disparity-split
key exchange -- exactly as it runs under NCCL, with the CPU oracle standing in for the per-rank kernels."""
import os
import socket
import struct

import numpy as np
import torch.multiprocessing as mp

from aswstereomatch_b200 import sharding

INF_TOP = 0xFFF0000000000000


def ref_keys(vol, d_first):
    """numpy restatement of wta_key_d (csrc/asw_common.cuh): 48-bit orderable double cost << 16 | d"""
    D, H, W = vol.shape
    best = np.full((H, W), np.uint64(0xFFFFFFFFFFFFFFFF))
    for d in range(D):
        c = vol[d].astype(np.float64)
        b = c.view(np.uint64)
        o = np.where(b >> np.uint64(63) == 1, ~b, b | np.uint64(1 << 63))
        o = np.where(np.isnan(c), np.uint64(0xFFFFFFFFFFFFFFFF), o)
        k = (o & np.uint64(0xFFFFFFFFFFFF0000)) | np.uint64(d_first + d)
        best = np.minimum(best, k)
    return best


def keys_to_disp(keys):
    return np.where(keys >= np.uint64(INF_TOP), 0.0, (keys & np.uint64(0xFFFF)).astype(np.float32)).astype(np.float32)


def test_shard_and_split_ranges():
    for world in (1, 2, 3, 4, 8):
        owned = sorted(i for r in range(world) for i in sharding.shard_pairs(64, r, world))
        assert owned == list(range(64))
        for D in (1, 5, 64, 129, 256):
            rng = [sharding.split_range(D, r, world) for r in range(world)]
            assert rng[0][0] == 0 and rng[-1][1] == D
            assert all(rng[i][1] == rng[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rng) - min(b - a for a, b in rng) <= 1


def test_key_i64_round_trip_preserves_order():
    rng = np.random.default_rng(0)
    k = rng.integers(0, 2**64 - 1, 4096, dtype=np.uint64)
    k[:4] = [0, 2**63 - 1, 2**63, 2**64 - 1]
    i = sharding.keys_to_i64(k)
    assert np.array_equal(sharding.keys_from_i64(i), k)
    assert np.array_equal(np.argsort(i, kind="stable"), np.argsort(k, kind="stable"))


def test_keys_reproduce_reference_wta():
    from oracle import orc
    rng = np.random.default_rng(1)
    vol = rng.integers(0, 5, (11, 17, 23)).astype(np.float32)
    vol[:, 0, 0] = np.nan
    vol[:, 1, 1] = np.inf
    vol[4, 2, 2] = -3.5
    assert np.array_equal(keys_to_disp(ref_keys(vol, 7)), orc.wta(vol, 7))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from aswstereomatch_b200.synth import make_pair
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, W, D = 40, 56, 9
    L, R, _ = make_pair(H, W, D, 5)
    full_d, q = orc.asw_guidedf2(L, R, 0, 1e-4, 5, 0, D, agg=True)        # every slice is independent (A.cpp:2775)
    lo, hi = sharding.split_range(D, rank, world)
    local = ref_keys(q[lo:hi], lo)                                         # what asw_split_local_keys yields per rank
    merged = sharding.allreduce_min_keys(local)
    disp = keys_to_disp(merged)
    ok = bool(np.array_equal(disp, full_d))
    # pair sharding: each rank processes its own pairs, no collective; gather only the check
    mine = sharding.shard_pairs(5, rank, world)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([int(ok), len(mine)]))
    dist.barrier()
    dist.destroy_process_group()


def test_disparity_split_exchange_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"r{r}.npy") for r in range(world)]
    assert all(r[0] == 1 for r in res)
    assert sum(int(r[1]) for r in res) == 5
