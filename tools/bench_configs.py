"""Per-config measurement of BASELINE.json's configs 1-4 (config 5 is bench.py's headline line).

For every config: kernel-only time with the pair resident in HBM (CUDA events on the library's stream, L2 flushed
before every repetition, >= 3 warm-ups), named-D MDE/s, the roofline figure of SURVEY.md section 8(d) for that config
(FP32 issue rate or HBM bytes) and the per-kernel split.  One JSON line per config; `--out` also writes them to a file.
Developer/measurement tool: bench.py stays the driver's contract.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair

SM, LANES = 148, 128


def peaks():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        hbm = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    return hbm


# name, (W, H), D named, D evaluated, runner, roofline (kind, work per DE_eval)
CONFIGS = [
    ("cfg1 traditional ASW 384x288 D=16 win=35", (384, 288), 16, 17, ("method", asw.ADAPTIVE_WEIGHT, 35), ("fp32", 4896.0)),
    ("cfg2 GuidedF_2 450x375 D=64 r=9 eps=1e-4 + right view + LR + refine", (450, 375), 64, 64, ("gf2lr", 9, 1e-4), ("hbm", 52.0)),
    ("cfg3a bilateral grid 1280x720 D=128", (1280, 720), 128, 129, ("method", asw.ADAPTIVE_WEIGHT_BILATERAL_GRID, 1), ("hbm", 358.0)),
    ("cfg3b BLO(1) 1280x720 D=128 win=35", (1280, 720), 128, 128, ("method", asw.ADAPTIVE_WEIGHT_BLO1, 35), ("fp32", 516.0)),
    ("cfg4 geodesic ASW 1280x720 D=128 win=35", (1280, 720), 128, 129, ("method", asw.ADAPTIVE_WEIGHT_GEODESIC, 35), ("fp32", 3675.0)),
]


def run_configs(ctx, reps=3, warmup=3, only="", brief=False):
    """measure configs 1-4 on `ctx`; returns one dict per config (brief: the bench.py `configs` entries)"""
    hbm = peaks()
    lines = []
    for idx, (name, (W, H), D, Dev, run, (kind, work)) in enumerate(CONFIGS):
        if only and only not in name:
            continue
        L, R, _ = make_pair(H, W, D, idx + 1)
        b = asw.Batch(ctx, 1, H, W)
        b.upload(0, L, R)
        views = 2 if run[0] == "gf2lr" else 1

        def step():
            if run[0] == "gf2lr":
                b.run_guidedf2_lr_refine(run[2], run[1], 0, D)
            else:
                b.run_method(run[1], 0, run[2], 0, D)

        for _ in range(warmup):
            step()
        ctx.sync()
        times = []
        for _ in range(reps):
            ctx.flush_l2()
            ctx.timer_start()
            step()
            times.append(ctx.timer_stop())
        ms = float(np.median(times))
        ctx.profile_enable(True); ctx.profile_reset()
        step(); ctx.sync()
        prof = ctx.profile()
        ctx.profile_enable(False)
        de_named = W * H * D * views
        de_eval = W * H * Dev * views
        if kind == "fp32":
            peak = SM * LANES * 1.965e9
            unit = "lane-instr/s"
        else:
            peak = hbm * 1e9
            unit = "B/s"
        achieved = work * de_eval / (ms * 1e-3)
        kern = {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]}
        if brief:
            line = {"cfg": name.split()[0], "config": name, "ms": round(ms, 4), "mde_s": round(de_named / ms / 1e3, 1),
                    "frac": round(achieved / peak, 4), "bound": kind, "work_per_de": work, "kernels_ms": kern}
        else:
            line = {"config": name, "ms": ms, "mde_per_s": de_named / ms / 1e3, "roofline": {
                "bound": kind, "work_per_de": work, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak},
                "kernels_ms": kern, "l2": "flushed before every repetition", "reps": reps, "warmup": warmup}
        lines.append(line)
        b.close()
    return lines


def run_driver_methods(ctx, reps=3, warmup=2):
    """every value of the reference's dispatcher that belongs to the hot path (aswMethods.cpp:46-88), called the way the
    reference's own driver calls it (aswStereoMatch.cpp:30, 94): 640x360, DISPARITY_LEFT, winSize 15, minDisparity 0,
    numDisparity 64.  Kernel-only time with the pair resident, L2 flushed before every repetition."""
    H, W, D, WIN = 360, 640, 64, 15
    L, R, _ = make_pair(H, W, D, 9)
    b = asw.Batch(ctx, 1, H, W)
    b.upload(0, L, R)
    names = {getattr(asw, k): k for k in ("ADAPTIVE_WEIGHT", "ADAPTIVE_WEIGHT_8DIRECT", "ADAPTIVE_WEIGHT_GEODESIC",
                                          "ADAPTIVE_WEIGHT_BILATERAL_GRID", "ADAPTIVE_WEIGHT_BLO1", "ADAPTIVE_WEIGHT_GUIDED_FILTER",
                                          "ADAPTIVE_WEIGHT_GUIDED_FILTER_2", "ADAPTIVE_WEIGHT_GUIDED_FILTER_3",
                                          "ADAPTIVE_WEIGHT_MEDIAN", "NCC")}
    out = []
    for alg in sorted(names):
        for _ in range(warmup):
            b.run_method(alg, 0, WIN, 0, D)
        ctx.sync()
        ts = []
        for _ in range(reps):
            ctx.flush_l2()
            ctx.timer_start()
            b.run_method(alg, 0, WIN, 0, D)
            ts.append(ctx.timer_stop())
        n0 = ctx.launch_count()
        b.run_method(alg, 0, WIN, 0, D)
        ctx.sync()
        out.append({"algorithm": names[alg], "ms": round(float(np.median(ts)), 4), "launches": int(ctx.launch_count() - n0)})
    b.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    ctx = asw.Context(0)
    lines = run_configs(ctx, a.reps, a.warmup, a.only)
    for ln in lines:
        print(json.dumps(ln), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
