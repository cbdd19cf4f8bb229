"""Developer fuzz run: random small geometries for every method of the hot path, CUDA path (through the C ABI) against the
oracle.  Prints one line per failure and a summary; exits non-zero on any failure.
usage: fuzz_parity.py [n_cases] [seed] [big]   (big: images up to 200 x 320, windows up to 35 for the methods that take them)"""
import os
import sys
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair
from oracle import orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
BIG = len(sys.argv) > 3
ctx = asw.Context(0)
fails = 0
counts = {}


def agree(a, b):
    return float((a == b).mean())


for case in range(N):
    H, W = int(rng.integers(12, 200 if BIG else 90)), int(rng.integers(12, 320 if BIG else 150))
    D = int(rng.integers(2, min(40 if BIG else 24, W - 2)))
    dt = int(rng.integers(0, 2))
    seed = int(rng.integers(0, 10000))
    L, R, _ = make_pair(H, W, D, seed)
    kind = ["gf2", "gf2lr", "trad", "d8", "geo", "grid", "blo1", "gf1", "gf3", "ncc", "wm"][case % 11]
    win = int(rng.choice([3, 5, 7, 9, 11, 13, 15]))
    if BIG and kind in ("trad", "d8", "geo", "blo1", "ncc"):
        win = int(rng.choice([9, 15, 21, 27, 35]))
    desc = f"{kind} H={H} W={W} D={D} win={win} dt={dt} seed={seed}"
    try:
        if kind == "gf2":
            d = ctx.computeAdaptiveWeight_GuidedF_2(L, R, dt, 1e-4, win, 0, D, strict=True)
            ok = agree(d, orc.asw_guidedf2(L, R, dt, 1e-4, win, 0, D)) >= 0.995
        elif kind == "gf2lr":
            out = ctx.guidedf2_lr_refine(L, R, 1e-4, win, 0, D)
            ok = agree(out, orc.guidedf2_lr_refine(L, R, 1e-4, win, 0, D)[0]) >= 0.98
        elif kind == "trad":
            d = ctx.computeAdaptiveWeight(L, R, 30, 20, dt, win, 0, D, strict=True)
            ok = agree(d, orc.asw_traditional(L, R, 30, 20, dt, win, 0, D)) >= 0.995
        elif kind == "d8":
            d = ctx.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_8DIRECT, win, 0, D, strict=True)
            ok = agree(d, orc.stereo_matching(L, R, 0, asw.ADAPTIVE_WEIGHT_8DIRECT, win, 0, D)) >= 0.995
        elif kind == "geo":
            d = ctx.computeAdaptiveWeight_geodesic(L, R, dt, win, 0, D, strict=True)
            ok = agree(d, orc.asw_geodesic(L, R, dt, win, 0, D)) >= 0.995
        elif kind == "grid":
            if min(H, W) < 64:
                continue
            d = ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, 10, 10, 0, D, strict=True)
            ok = agree(d, orc.asw_bilateral_grid(L, R, 0, 10, 10, 0, D)) >= 0.995
        elif kind == "blo1":
            d = ctx.computeAdaptiveWeight_BLO1(L, R, dt, 0.015, win, 0, D, strict=True)
            ok = agree(d, orc.asw_blo1(L, R, dt, 0.015, win, 0, D)) >= 0.995
        elif kind == "gf1":
            d = ctx.computeAdaptiveWeight_GuidedF(L, R, dt, 1e-6, win, 0, D, strict=True)
            ok = agree(d, orc.asw_guidedf(L, R, dt, 1e-6, win, 0, D)) >= 0.995
        elif kind == "gf3":
            d = ctx.computeAdaptiveWeight_GuidedF_3(L, R, dt, 1e-6, win, 0, D, strict=True)
            ok = agree(d, orc.asw_guidedf3(L, R, dt, 1e-6, win, 0, D)) >= 0.99
        elif kind == "ncc":
            d = ctx.computeNCC(L, R, dt, win, 0, D, strict=True)
            ok = agree(d, orc.asw_ncc(L, R, dt, win, 0, D)) >= 0.995
        else:
            d = ctx.computeAdaptiveWeight_WeightedMedian(L, R, 0, win, 10, 10, 0, D, strict=True)
            ok = agree(d, orc.asw_weighted_median(L, R, 0, win, 10, 10, 0, D)) >= 0.995
        counts[kind] = counts.get(kind, 0) + 1
        if not ok:
            fails += 1
            print("MISMATCH", desc, flush=True)
    except Exception as e:      # noqa: BLE001
        msg = str(e).splitlines()[0][:120]
        if "UNSUPPORTED" in msg or "BAD_ARG" in msg or "oracle" in msg:
            print("skip", desc, "->", msg, flush=True)
            continue
        fails += 1
        print("ERROR", desc, "->", msg, flush=True)
        traceback.print_exc()
print("cases per method:", counts, "failures:", fails)
sys.exit(1 if fails else 0)
