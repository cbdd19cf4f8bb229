"""GPU parity of the streaming guided-filter kernel (k_guided_stream.cuh) at the shapes that stress its geometry:
strips that fold at the left / right image edge, bands whose first / last block carries the REFLECT_101 closed
forms, slice groups that are not a multiple of 4, every tuned window size, the right view, and many row bands.
Bars as everywhere: aggregated costs within 1e-4 of the slice max, WTA bit-exact given the GPU volume, maps >= 99.9 %.
The tiled kernels (k_guided_fast.cuh) stay covered as the fallback for images lower than 4 windows."""
import os

import numpy as np
import pytest

import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair
from oracle import orc

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
AGREE = 0.999


def rel_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.float64)
    b = b.reshape(b.shape[0], -1).astype(np.float64)
    s = np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-30)
    return float((np.abs(a - b) / s).max())


def check_left(ctx, H, W, D, win, eps, seed):
    L, R, _ = make_pair(H, W, D, seed)
    d, q = ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, eps, win, 0, D, agg=True, strict=True)
    d_ref, q_ref = orc.asw_guidedf2(L, R, 0, eps, win, 0, D, agg=True)
    assert rel_err(q, q_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE
    assert np.array_equal(d, orc.wta(q, 0))


# widths around the 48 / 52 / 56-column strip pitch, heights around the 2-band minimum (4 windows) and around
# multiples of the block height, disparity counts around the 4-slice group
@pytest.mark.parametrize("H,W,D,win,seed", [
    (36, 48, 4, 9, 1), (37, 49, 5, 9, 2), (45, 47, 3, 9, 3), (54, 96, 1, 9, 4), (63, 97, 9, 9, 5),
    (81, 143, 6, 9, 6), (100, 33, 7, 9, 7), (121, 200, 2, 9, 8),
    (28, 52, 5, 7, 9), (50, 105, 4, 7, 10), (20, 56, 3, 5, 11), (47, 113, 6, 5, 12),
])
def test_stream_geometry(ctx, H, W, D, win, seed):
    check_left(ctx, H, W, D, win, 1e-4, seed)


@pytest.mark.parametrize("bands", [2, 3, 5, 8])
def test_stream_many_bands(ctx, bands):
    """middle bands (no image edge), top and bottom bands, forced band counts"""
    ctx.set_tuning(ctx.TUNE_GFS_BANDS, bands)
    try:
        check_left(ctx, 150, 120, 8, 9, 1e-4, 20 + bands)
    finally:
        ctx.set_tuning(ctx.TUNE_GFS_BANDS, 0)


def test_stream_small_eps_large_range(ctx):
    """eps = 1e-6 (the dispatcher's literal, A.cpp:76): the a = cov / (var + eps) amplification at its worst"""
    check_left(ctx, 90, 130, 12, 9, 1e-6, 31)


def test_stream_small_eps_half_hd(ctx):
    """eps = 1e-6 at 960 x 540 x 24: the fp32 window sums of the streaming kernel against the oracle's double accumulation
    where a = cov / (var + eps) amplifies most, on a frame large enough for every band / strip variant"""
    orc.set_num_threads(os.cpu_count() or 1)
    check_left(ctx, 540, 960, 24, 9, 1e-6, 33)


def test_stream_right_view_and_refine(ctx):
    L, R, _ = make_pair(77, 131, 14, 41)
    out, parts = ctx.guidedf2_lr_refine(L, R, 1e-4, 9, 0, 14, parts=True)
    ref, rparts = orc.guidedf2_lr_refine(L, R, 1e-4, 9, 0, 14)
    assert (parts["dl"] == rparts["dl"]).mean() >= AGREE
    assert (parts["dr"] == rparts["dr"]).mean() >= AGREE
    v = orc.lr_check(parts["dl"], parts["dr"], 0.0)
    assert np.array_equal(parts["valid"], v)
    assert np.array_equal(out, orc.wmedian_refine(L, orc.fill_invalid(parts["dl"], v), v, 9, 10, 10))


def test_stream_nonzero_min_disparity(ctx):
    L, R, _ = make_pair(60, 100, 12, 51)
    d, q = ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-4, 9, 3, 7, agg=True, strict=True)
    d_ref, q_ref = orc.asw_guidedf2(L, R, 0, 1e-4, 9, 3, 7, agg=True)
    assert rel_err(q, q_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


def test_stream_equals_tiled_fallback(ctx):
    """development builds: the two CUDA paths agree with each other far inside the oracle tolerance"""
    if not ctx.has_dev_kernels():
        pytest.skip("the tiled pair is compiled only with -DASW_DEV_KERNELS")
    L, R, _ = make_pair(96, 160, 10, 61)
    d1, q1 = ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-4, 9, 0, 10, agg=True, strict=True)
    os.environ["ASW_GF_TILED"] = "1"
    try:
        d2, q2 = ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-4, 9, 0, 10, agg=True, strict=True)
    finally:
        del os.environ["ASW_GF_TILED"]
    assert rel_err(q1, q2) <= REL_TOL
    assert (d1 == d2).mean() >= AGREE


def test_low_images_take_the_generic_kernels(ctx):
    """H < 4 windows: below the streaming kernel's two-band minimum -> the size-generic two-pass guided filter"""
    check_left(ctx, 30, 64, 6, 9, 1e-4, 71)


def test_full_hd_row_properties(ctx):
    """config-5 size: size-independent properties instead of the (slow) oracle: the disparity-range split of the
    streaming path reproduces the unsplit map bit-exactly, and the map recovers the synthetic ground truth"""
    from aswstereomatch_b200 import sharding
    H, W, D = 1080, 1920, 64
    L, R, gt = make_pair(H, W, D, 1005)
    full = ctx.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 0, D, strict=True)
    merged = None
    for r in range(2):
        lo, hi = sharding.split_range(D, r, 2)
        keys, _ = ctx.split_local_keys(L, R, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 0, 9, 0, D, lo, hi)
        merged = keys if merged is None else np.minimum(merged, keys)
    assert np.array_equal(ctx.keys_to_disparity(merged), full)
    assert np.mean(np.abs(full - gt) <= 1) > 0.85
