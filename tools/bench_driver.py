"""One iteration of the reference DRIVER's loop (aswStereoMatch.cpp:27-100) on the device, timed with CUDA events:
raw pair (any size) -> upload + resize to 640x360 + V-channel bilateral detail boost (asw_batch_upload_raw) ->
stereoMatching(..., DISPARITY_LEFT, GuidedF_2, 15, 0, 64) (the driver's own literals, :94) -> download ->
convertTo(CV_8U) + normalize(0, 255) (asw_disparity_to_u8).  Developer measurement tool; bench.py stays the contract."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair

ctx = asw.Context(0)
H0, W0, H, W, D, WIN = 720, 1280, 360, 640, 64, 15
L, R, _ = make_pair(H0, W0, 2 * D, 77)
Lp = asw.pinned_empty(L.shape, np.uint8); Rp = asw.pinned_empty(R.shape, np.uint8)
Lp[...] = L; Rp[...] = R
b = asw.Batch(ctx, 1, H, W)


def step():
    b.upload_raw(0, Lp, Rp)
    b.run_method(asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 0, WIN, 0, D)
    return b.download(0)


for _ in range(3):
    step()
ts = []
for _ in range(7):
    ctx.flush_l2(); ctx.timer_start(); d = step(); ts.append(ctx.timer_stop())
ctx.profile_enable(True); ctx.profile_reset(); step(); ctx.sync()
prof = {k: round(v[0], 4) for k, v in sorted(ctx.profile().items(), key=lambda kv: -kv[1][0])}
ctx.profile_enable(False)
u8 = ctx.disparity_to_u8(d)
print(json.dumps({"what": "driver loop iteration: 1280x720 raw pair -> 640x360, GuidedF_2 win 15, D 64 (aswStereoMatch.cpp:30-31, 67-89, 94)",
                  "ms_host_to_host": round(float(np.median(ts)), 4), "mde_s": round(H * W * D / 1e6 / (float(np.median(ts)) * 1e-3), 1),
                  "kernels_ms": prof, "u8_range": [int(u8.min()), int(u8.max())]}))
b.close(); ctx.close()
