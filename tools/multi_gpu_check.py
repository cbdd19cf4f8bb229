"""N-GPU check of the multi-GPU paths on real devices (run under torchrun, one rank per GPU, NCCL):
  1. disparity-range split of one pair: per-rank keys from the CUDA path, MIN all-reduce over NCCL, map == the unsplit map
  2. pair sharding: every rank runs its own pairs with no collective; results equal the single-GPU results
Prints one JSON line on rank 0; exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aswstereomatch_b200 as asw
from aswstereomatch_b200 import sharding
from aswstereomatch_b200.synth import make_pair

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = asw.Context(local)
H, W, D = 375, 450, 64
L, R, _ = make_pair(H, W, D, 2)
full = ctx.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 0, D, strict=True)
split = sharding.split_stereo_matching(ctx, L, R, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 0, 9, 0, D, rank, world,
                                       device=torch.device("cuda", local))
ok_split = bool(np.array_equal(split, full))
# pair sharding: 6 pairs round-robin; every rank checks its pairs against a fresh single-context run
n_pairs = 6
mine = sharding.shard_pairs(n_pairs, rank, world)
b = asw.Batch(ctx, len(mine), 96, 128)
pairs = {i: make_pair(96, 128, 16, 100 + i)[:2] for i in mine}
for k, i in enumerate(mine):
    b.upload(k, *pairs[i])
b.run_guidedf2_lr_refine(1e-4, 9, 0, 16)
ok_pairs = all(np.array_equal(b.download(k), ctx.guidedf2_lr_refine(*pairs[i], 1e-4, 9, 0, 16)) for k, i in enumerate(mine))
b.close()
ok_split_dplus1 = True
for alg, win in ((asw.ADAPTIVE_WEIGHT, 9), (asw.ADAPTIVE_WEIGHT_BILATERAL_GRID, 9), (asw.ADAPTIVE_WEIGHT_BLO1, 9)):
    # D + 1 candidates for traditional / grid: the last one must be evaluated by some rank
    Ls, Rs, _ = make_pair(96, 128, 16, 5)
    f2 = ctx.stereoMatching(Ls, Rs, 0, alg, win, 0, 16, strict=True)
    s2 = sharding.split_stereo_matching(ctx, Ls, Rs, alg, 0, win, 0, 16, rank, world, device=torch.device("cuda", local))
    ok_split_dplus1 = ok_split_dplus1 and bool(np.array_equal(s2, f2))
ok_split = ok_split and ok_split_dplus1
t = torch.tensor([int(ok_split), int(ok_pairs), len(mine)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.SUM)
if rank == 0:
    res = {"world": world, "split_equals_unsplit_on_all_ranks": int(t[0]) == world, "pair_shards_ok_on_all_ranks": int(t[1]) == world,
           "pairs_covered": int(t[2]), "backend": "nccl"}
    print(json.dumps(res), flush=True)
ctx.close()
dist.destroy_process_group()
pool_ok = True
if rank == 0 and os.environ.get("ASW_CHECK_POOL", "1") == "1":
    # the single-process face of the same thing: asw_pool_* (one host thread per device, NCCL bound inside the library)
    pool = asw.Pool(world)
    c0 = asw.Context(0)
    ps = pool.stereoMatchingSplit(L, R, 0, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 0, D)
    pool_ok = bool(np.array_equal(ps, c0.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 0, D, strict=True)))
    Lb = [make_pair(96, 128, 16, 300 + i)[0] for i in range(5)]
    Rb = [make_pair(96, 128, 16, 300 + i)[1] for i in range(5)]
    outs = pool.guidedf2_lr_refine_batch(Lb, Rb, 1e-4, 9, 0, 16)
    pool_ok = pool_ok and all(np.array_equal(o, c0.guidedf2_lr_refine(a, b, 1e-4, 9, 0, 16)) for o, a, b in zip(outs, Lb, Rb))
    print(json.dumps({"pool_devices": pool.size, "pool_split_and_batch_ok": pool_ok, "pool_allreduce_ms": pool.last_allreduce_ms()}), flush=True)
    pool.close(); c0.close()
    if not pool_ok:
        sys.exit(2)
sys.exit(0 if (int(t[0]) == world and int(t[1]) == world and int(t[2]) == n_pairs) else 1)
