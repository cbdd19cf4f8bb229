"""Developer timing probe (not the bench): per-kernel CUDA-event times for a method on a synthetic pair."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=375)
ap.add_argument("--W", type=int, default=450)
ap.add_argument("--D", type=int, default=64)
ap.add_argument("--win", type=int, default=9)
ap.add_argument("--alg", type=int, default=8)
ap.add_argument("--pairs", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lr", action="store_true")
a = ap.parse_args()

ctx = asw.Context(0)
pairs = [make_pair(a.H, a.W, a.D, 1000 + i)[:2] for i in range(a.pairs)]
b = asw.Batch(ctx, a.pairs, a.H, a.W)
for i, (L, R) in enumerate(pairs):
    b.upload(i, L, R)


def run():
    if a.lr:
        b.run_guidedf2_lr_refine(1e-4, a.win, 0, a.D)
    else:
        b.run_method(a.alg, 0, a.win, 0, a.D)


run(); ctx.sync()
views = 2 if a.lr else 1
for rep in range(a.reps):
    ctx.flush_l2()
    ctx.timer_start()
    run()
    ms = ctx.timer_stop()
    mde = a.H * a.W * a.D * views * a.pairs / 1e6
    print(f"rep {rep}: {ms:.3f} ms  -> {mde / ms * 1e3:.0f} MDE/s  ({ms / a.pairs:.3f} ms/frame)")
ctx.profile_enable(True); ctx.profile_reset()
run(); ctx.sync()
tot = sum(v[0] for v in ctx.profile().values())
for k, (ms, n) in sorted(ctx.profile().items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:20s} {ms:9.3f} ms  {n:5d} launches  {100 * ms / tot:5.1f}%")
print("  total", tot)
b.close()
ctx.close()
