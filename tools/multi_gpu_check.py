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
t = torch.tensor([int(ok_split), int(ok_pairs), len(mine)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.SUM)
if rank == 0:
    res = {"world": world, "split_equals_unsplit_on_all_ranks": int(t[0]) == world, "pair_shards_ok_on_all_ranks": int(t[1]) == world,
           "pairs_covered": int(t[2]), "backend": "nccl"}
    print(json.dumps(res), flush=True)
ctx.close()
dist.destroy_process_group()
sys.exit(0 if (int(t[0]) == world and int(t[1]) == world and int(t[2]) == n_pairs) else 1)
