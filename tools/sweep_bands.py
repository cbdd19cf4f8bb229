import sys, numpy as np
sys.path.insert(0, "/root/repo")
import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair
ctx = asw.Context(0)
H, W, D = 375, 450, 64
L, R, _ = make_pair(H, W, D, 2)
b = asw.Batch(ctx, 1, H, W); b.upload(0, L, R)
for bands in (0, 2, 3, 4, 5, 6, 8):
    ctx.set_tuning(ctx.TUNE_GFS_BANDS, bands)
    for _ in range(3): b.run_guidedf2_lr_refine(1e-4, 9, 0, D)
    ctx.sync(); ts = []
    for _ in range(5):
        ctx.flush_l2(); ctx.timer_start(); b.run_guidedf2_lr_refine(1e-4, 9, 0, D); ts.append(ctx.timer_stop())
    print("bands", bands, "ms", round(float(np.median(ts)), 4))
