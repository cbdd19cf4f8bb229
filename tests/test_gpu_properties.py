"""Size-independent properties at BASELINE.json's full sizes (configs 1 - 4; config 5: test_gpu_full_size.py).

The oracle covers whole images where it finishes in seconds and bands of rows elsewhere (test_gpu_full_size.py); these
properties cover EVERY pixel of the full-size maps without the oracle's cost:

  * winner-take-all is an integer stage: the disparity map must be the reference's scan (strict `<`, ascending d, NaN
    never wins -- A.cpp:1144-1150 and the other WTA blocks) of the method's OWN aggregated volume;
  * labels are integral and inside [minDisparity, minDisparity + candidates - 1];
  * a second run is bit-identical (the per-pixel WTA keys are merged with atomicMin: the result must not depend on the
    order in which CTAs finish), and the capture of the volume does not change the map.
"""
import numpy as np
import pytest

import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair
from oracle import orc

pytestmark = pytest.mark.gpu


def _run(ctx, name, L, R, D, agg):
    if name == "traditional":
        return ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 35, 0, D, agg=agg, strict=True)
    if name == "geodesic":
        return ctx.computeAdaptiveWeight_geodesic(L, R, 0, 35, 0, D, agg=agg, strict=True)
    if name == "grid":
        return ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, 10, 10, 0, D, agg=agg, strict=True)
    if name == "blo1":
        return ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.015, 35, 0, D, agg=agg, strict=True)
    if name == "guidedf2":
        return ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-4, 9, 0, D, agg=agg, strict=True)
    raise ValueError(name)


# (method, H, W, D, candidates, seed, exact): `exact` = the kernel compares the float values it stores; the loop-only methods
# compare E = num / den in double (as the reference does) and store it as float, so two candidates whose doubles differ
# but round to one float may resolve differently in a float re-scan: equality then holds on all but near-tie pixels
CASES = [
    ("traditional", 288, 384, 16, 17, 1, False),
    ("guidedf2", 375, 450, 64, 64, 2, True),
    ("grid", 720, 1280, 128, 129, 3, False),
    ("blo1", 720, 1280, 128, 128, 4, True),
    ("geodesic", 720, 1280, 128, 129, 5, False),
]


@pytest.mark.parametrize("name,H,W,D,ncand,seed,exact", CASES, ids=[c[0] for c in CASES])
def test_full_size_map_is_the_wta_of_its_own_volume(ctx, name, H, W, D, ncand, seed, exact):
    L, R, _ = make_pair(H, W, D, seed)
    d, vol = _run(ctx, name, L, R, D, True)
    assert vol.shape == (ncand, H, W) and d.shape == (H, W) and d.dtype == np.float32
    # integral labels inside the candidate range
    assert np.array_equal(d, np.rint(d))
    assert d.min() >= 0 and d.max() <= ncand - 1
    scan = orc.wta(vol, 0)                                           # the reference's scan, on the GPU's own volume
    if exact:
        assert np.array_equal(d, scan)
    else:
        # near-ties only: wherever the two scans pick different labels, the two labels' stored costs agree to float rounding
        # (geodesic: where x + h < d every operand of candidate d is clamped onto the same pixels as for d - 1, the reference's
        # costs are one number and strict < keeps the lower d; the kernel that sums the last candidate in another order is
        # not allowed to win there -- k_geo_aggregate_q -- although its stored float is a few ulp lower)
        assert (d == scan).mean() >= 0.998
        yy, xx = np.nonzero(d != scan)
        a = vol[d[yy, xx].astype(int), yy, xx].astype(np.float64)
        b = vol[scan[yy, xx].astype(int), yy, xx].astype(np.float64)
        assert np.all(np.abs(a - b) <= 1e-5 * np.maximum(np.abs(b), 1e-30))
    # the map of a run without the capture is the same map, and a repeated run is bit-identical
    d2 = _run(ctx, name, L, R, D, False)
    assert np.array_equal(d, d2)
    assert np.array_equal(d2, _run(ctx, name, L, R, D, False))

