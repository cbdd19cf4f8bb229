"""GPU parity of the four loop-heavy ASW variants (traditional, geodesic, bilateral grid, BLO(1)) and the
dispatcher, through the C ABI, against the CPU oracle.  Same tolerances as test_gpu_guided.py."""
import numpy as np
import pytest

import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair
from oracle import orc

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
AGREE = 0.999


def rel_err(a, b, mask=None):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    ok = np.isfinite(b) if mask is None else mask
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    if not ok.any():
        return 0.0
    s = np.maximum(np.abs(np.where(ok, b, 0)).reshape(b.shape[0], -1).max(axis=1), 1e-30)[:, None, None]
    return float((np.abs(np.where(ok, a - b, 0)) / s).max())


@pytest.mark.parametrize("H,W,D,win,disp_type,seed", [(40, 56, 6, 5, 0, 1), (48, 64, 8, 9, 0, 2), (36, 50, 5, 7, 1, 3),
                                                       (64, 80, 16, 35, 0, 4)])
def test_traditional(ctx, H, W, D, win, disp_type, seed):
    L, R, _ = make_pair(H, W, D, seed)
    d, e = ctx.computeAdaptiveWeight(L, R, 30, 20, disp_type, win, 0, D, agg=True, strict=True)
    d_ref, e_ref = orc.asw_traditional(L, R, 30, 20, disp_type, win, 0, D, agg=True)
    assert e.shape == (D + 1, H, W)                       # D+1 candidates (A.cpp:1021, 1074)
    assert rel_err(e, e_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE
    assert np.array_equal(d, orc.wta(e, 0)) or (d == orc.wta(e, 0)).mean() >= 0.9999


@pytest.mark.parametrize("H,W,D,win,disp_type,seed", [(14, 300, 33, 7, 0, 21), (12, 290, 40, 9, 1, 22), (10, 200, 16, 35, 0, 23),
                                                       (10, 420, 70, 5, 1, 24), (9, 130, 3, 5, 0, 25)])
@pytest.mark.parametrize("diag", [False, True])
def test_traditional_wide_many_candidates(ctx, H, W, D, win, disp_type, seed, diag):
    """rows longer than one tile / 128-pixel segment and >= 32 candidates, both views, for the default tiled kernel
    and for the selectable diagonal-blocked kernel (interior and edge segments, full chunks plus remainder)"""
    import os
    if diag and not ctx.has_dev_kernels():
        pytest.skip("the diagonal-blocked kernel is compiled only with -DASW_DEV_KERNELS")
    L, R, _ = make_pair(H, W, D, seed)
    if diag:
        os.environ["ASW_TRAD_DIAG"] = "1"
    try:
        d, e = ctx.computeAdaptiveWeight(L, R, 30, 20, disp_type, win, 0, D, agg=True, strict=True)
    finally:
        os.environ.pop("ASW_TRAD_DIAG", None)
    d_ref, e_ref = orc.asw_traditional(L, R, 30, 20, disp_type, win, 0, D, agg=True)
    assert rel_err(e, e_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


def needs_dev(ctx):
    if not ctx.has_dev_kernels():
        pytest.skip("superseded kernels are compiled only with -DASW_DEV_KERNELS (python __graft_entry__.py --dev)")


@pytest.mark.parametrize("env", ["ASW_TRAD_DIAG", "ASW_TRAD_EXACT"])
def test_traditional_other_kernels(ctx, env):
    """development builds: the diagonal-blocked kernel and the exact-table kernel agree with the oracle"""
    import os
    needs_dev(ctx)
    L, R, _ = make_pair(40, 72, 8, 26)
    os.environ[env] = "1"
    try:
        d, e = ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 9, 0, 8, agg=True, strict=True)
    finally:
        del os.environ[env]
    d_ref, e_ref = orc.asw_traditional(L, R, 30, 20, 0, 9, 0, 8, agg=True)
    assert rel_err(e, e_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


def test_traditional_nonzero_min_disparity(ctx):
    L, R, _ = make_pair(40, 64, 12, 8)
    d = ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 7, 3, 6, strict=True)
    d_ref = orc.asw_traditional(L, R, 30, 20, 0, 7, 3, 6)
    assert (d == d_ref).mean() >= AGREE
    assert d.min() >= 3 and d.max() <= 9


@pytest.mark.parametrize("H,W,win,seed", [(24, 30, 5, 1), (30, 41, 9, 2), (20, 26, 35, 3)])
def test_geodesic_distance_bit_exact(ctx, H, W, win, seed):
    L, _, _ = make_pair(H, W, 4, seed)
    got = ctx.getGeodesicDist(L, win)
    ref = orc.geodesic_dist(L, win)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("H,W,D,win,disp_type,seed", [(40, 56, 6, 5, 0, 1), (44, 60, 8, 9, 0, 2), (36, 50, 5, 7, 1, 3),
                                                       (48, 64, 8, 35, 0, 4)])
def test_geodesic(ctx, H, W, D, win, disp_type, seed):
    L, R, _ = make_pair(H, W, D, seed)
    d, e = ctx.computeAdaptiveWeight_geodesic(L, R, disp_type, win, 0, D, agg=True, strict=True)
    d_ref, e_ref = orc.asw_geodesic(L, R, disp_type, win, 0, D, agg=True)
    assert rel_err(e, e_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


@pytest.mark.parametrize("H,W,D,win,disp_type,seed", [(20, 300, 33, 7, 0, 11), (18, 290, 40, 9, 1, 12), (14, 210, 32, 35, 0, 13),
                                                       (16, 400, 70, 5, 1, 14)])
def test_geodesic_wide_many_candidates(ctx, H, W, D, win, disp_type, seed):
    """>= 32 candidates and rows longer than one 128-pixel segment: the diagonal-blocked kernel on interior and
    edge segments (pre-shift clamp), plus the tile kernel on the candidate remainder"""
    L, R, _ = make_pair(H, W, D, seed)
    d, e = ctx.computeAdaptiveWeight_geodesic(L, R, disp_type, win, 0, D, agg=True, strict=True)
    d_ref, e_ref = orc.asw_geodesic(L, R, disp_type, win, 0, D, agg=True)
    assert rel_err(e, e_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


@pytest.mark.parametrize("H,W,D,sS,sR,seed", [(48, 64, 8, 10, 10, 1), (120, 160, 8, 10, 10, 2), (60, 90, 6, 7, 12, 3)])
def test_bilateral_grid(ctx, H, W, D, sS, sR, seed):
    L, R, _ = make_pair(H, W, D, seed)
    d, e = ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, sS, sR, 0, D, agg=True, strict=True)
    d_ref, e_ref = orc.asw_bilateral_grid(L, R, 0, sS, sR, 0, D, agg=True)
    # fp64 with the reference's expression order: NaN / inf pattern identical, finite costs equal to float rounding
    assert np.array_equal(np.isnan(e), np.isnan(e_ref))
    fin = np.isfinite(e_ref)
    assert np.array_equal(np.isfinite(e), fin)
    assert np.array_equal(e[fin], e_ref[fin])
    assert np.array_equal(d, d_ref)


@pytest.mark.parametrize("H,W,D,win,seed", [(40, 56, 8, 7, 7), (64, 96, 12, 9, 2), (70, 100, 8, 35, 3), (50, 300, 20, 15, 4),
                                            (150, 200, 9, 25, 5), (48, 64, 8, 11, 6)])
def test_blo1(ctx, H, W, D, win, seed):
    L, R, _ = make_pair(H, W, D, seed)
    d, q = ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.015, win, 0, D, agg=True, strict=True)
    d_ref, q_ref = orc.asw_blo1(L, R, 0, 0.015, win, 0, D, agg=True)
    assert rel_err(q, q_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


@pytest.mark.parametrize("H,W,D,win,seed", [(40, 56, 8, 7, 17), (70, 100, 8, 35, 13), (50, 300, 20, 15, 14), (48, 64, 8, 11, 16)])
def test_blo1_right(ctx, H, W, D, win, seed):
    """DISPARITY_RIGHT (A.cpp:2538-2546, 2600-2631, 2685-2722): the reference defines it; templated and tiled kernels"""
    L, R, _ = make_pair(H, W, D, seed)
    d, q = ctx.computeAdaptiveWeight_BLO1(L, R, 1, 0.015, win, 0, D, agg=True, strict=True)
    d_ref, q_ref = orc.asw_blo1(L, R, 1, 0.015, win, 0, D, agg=True)
    assert rel_err(q, q_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE
    assert np.array_equal(ctx.stereoMatching(L, R, 1, asw.ADAPTIVE_WEIGHT_BLO1, win, 0, D, strict=True), d)


def test_blo1_golden(ctx):
    g = np.load("tests/golden/cv2_stages_40x56_d8.npz")
    d, q = ctx.computeAdaptiveWeight_BLO1(g["L"], g["R"], 0, 0.015, 7, 0, 8, agg=True, strict=True)
    assert rel_err(q, g["blo1_q"]) <= REL_TOL
    assert (d == g["blo1_disp"]).mean() >= AGREE


def test_blo1_coarse_levels(ctx):
    """a level step that does not divide 255 exercises the appended 255 level (A.cpp:2556-2559)"""
    L, R, _ = make_pair(40, 56, 6, 5)
    d, q = ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.05, 5, 0, 6, agg=True, strict=True)
    d_ref, q_ref = orc.asw_blo1(L, R, 0, 0.05, 5, 0, 6, agg=True)
    assert rel_err(q, q_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE


@pytest.mark.parametrize("alg", [asw.ADAPTIVE_WEIGHT, asw.ADAPTIVE_WEIGHT_GEODESIC, asw.ADAPTIVE_WEIGHT_BILATERAL_GRID,
                                 asw.ADAPTIVE_WEIGHT_BLO1, asw.ADAPTIVE_WEIGHT_GUIDED_FILTER,
                                 asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2])
def test_dispatcher_literals(ctx, alg):
    """stereoMatching (A.cpp:46-88) with its hard-coded hyper-parameters"""
    L, R, _ = make_pair(48, 64, 8, 21)
    d = ctx.stereoMatching(L, R, asw.DISPARITY_LEFT, alg, 9, 0, 8, strict=True)
    d_ref = orc.stereo_matching(L, R, 0, alg, 9, 0, 8)
    assert (d == d_ref).mean() >= AGREE


def test_size_generic_fallbacks_of_the_product_build(ctx):
    """the product build keeps one tuned path per method plus size-generic kernels; these sizes reach the generic ones:
    a 101 x 101 traditional window (tile > shared memory -> k_trad_aggregate over the exact table), a fine range grid
    (sR = 2.8: the (z, w) plane does not fit -> global splat + 4 pass launches), an 11 x 11 geodesic window (generic DP),
    a 21 x 21 guided / refine window (two-pass guided filter, dense weighted-median refine)"""
    L, R, _ = make_pair(40, 56, 3, 51)
    d, e = ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 101, 0, 3, agg=True, strict=True)
    d_ref, e_ref = orc.asw_traditional(L, R, 30, 20, 0, 101, 0, 3, agg=True)
    assert rel_err(e, e_ref) <= REL_TOL and (d == d_ref).mean() >= AGREE
    d, e = ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, 10, 2.8, 0, 3, agg=True, strict=True)
    d_ref, e_ref = orc.asw_bilateral_grid(L, R, 0, 10, 2.8, 0, 3, agg=True)
    fin = np.isfinite(e_ref)
    assert np.array_equal(np.isfinite(e), fin) and np.array_equal(e[fin], e_ref[fin]) and np.array_equal(d, d_ref)
    d, e = ctx.computeAdaptiveWeight_geodesic(L, R, 0, 11, 0, 3, agg=True, strict=True)
    d_ref, e_ref = orc.asw_geodesic(L, R, 0, 11, 0, 3, agg=True)
    assert rel_err(e, e_ref) <= REL_TOL and (d == d_ref).mean() >= AGREE
    L, R, _ = make_pair(90, 110, 6, 52)
    out, parts = ctx.guidedf2_lr_refine(L, R, 1e-4, 21, 0, 6, parts=True)
    ref, rparts = orc.guidedf2_lr_refine(L, R, 1e-4, 21, 0, 6)
    assert (parts["dl"] == rparts["dl"]).mean() >= AGREE and (parts["dr"] == rparts["dr"]).mean() >= AGREE
    v = orc.lr_check(parts["dl"], parts["dr"], 0.0)
    assert np.array_equal(parts["valid"], v)
    assert np.array_equal(out, orc.wmedian_refine(L, orc.fill_invalid(parts["dl"], v), v, 21, 10, 10))


@pytest.mark.parametrize("H,W,D,win,seed", [(44, 60, 6, 9, 21), (40, 50, 5, 35, 3), (30, 41, 4, 3, 5), (33, 147, 17, 15, 8)])
def test_direct8(ctx, H, W, D, win, seed):
    """computeAdaptiveWeight_direct8 (A.cpp:1167-1319): 3 (win - 1) taps (diagonal, row, column), gamma_g = win * 2 / 3 in
    integers, D + 1 candidates; LEFT only (the RIGHT branch indexes out of bounds, A.cpp:1291)"""
    L, R, _ = make_pair(H, W, D, seed)
    d, e = ctx.computeAdaptiveWeight_direct8(L, R, 0, win, 0, D, agg=True, strict=True)
    d_ref, e_ref = orc.asw_direct8(L, R, 0, win, 0, D, agg=True)
    assert e.shape == (D + 1, H, W)
    assert rel_err(e, e_ref) <= REL_TOL
    assert (d == d_ref).mean() >= AGREE
    assert np.array_equal(ctx.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_8DIRECT, win, 0, D, strict=True), d)
    assert ctx.computeAdaptiveWeight_direct8(L, R, 1, win, 0, D).size == 0


def test_direct8_reference_golden(ctx):
    g = np.load("tests/golden/ref_methods_44x60_d6.npz")
    d = ctx.computeAdaptiveWeight_direct8(g["L"], g["R"], 0, 9, 0, 6, strict=True)
    assert (d == g["direct8_w9"]).mean() >= AGREE


@pytest.mark.parametrize("H,W,D,win,seed", [(40, 56, 6, 9, 3), (33, 47, 5, 5, 8), (64, 90, 9, 7, 4)])
def test_ncc_and_guidedf3(ctx, H, W, D, win, seed):
    """row f-4: computeNCC (both overloads, A.cpp:812-1013) and GuidedF_3 (A.cpp:3063-3137), both views"""
    L, R, _ = make_pair(H, W, D, seed)
    for dt in (0, 1):
        v = ctx.computeNCC_volume(L, R, dt, win, 0, D)
        v_ref = orc.cost_ncc(L, R, 0, D, win, dt)
        assert np.abs(v - v_ref).max() <= 1e-6
        d, q = ctx.computeAdaptiveWeight_GuidedF_3(L, R, dt, 1e-6, win, 0, D, agg=True, strict=True)
        d_ref, q_ref = orc.asw_guidedf3(L, R, dt, 1e-6, win, 0, D, agg=True)
        assert rel_err(q, q_ref) <= REL_TOL
        assert (d == d_ref).mean() >= AGREE
        assert (ctx.computeNCC(L, R, dt, win, 0, D, strict=True) == orc.asw_ncc(L, R, dt, win, 0, D)).mean() >= AGREE
    assert np.array_equal(ctx.stereoMatching(L, R, 0, asw.NCC, win, 0, D, strict=True), ctx.computeNCC(L, R, 0, win, 0, D))
    assert ctx.computeNCC(L, R, 0, 8, 0, D).size == 0                # even window: empty Mat (A.cpp:824-827)


def test_dispatcher_out_of_scope(ctx):
    L, R, _ = make_pair(32, 40, 4, 1)
    for alg in (asw.BM, asw.SGBM):
        assert ctx.stereoMatching(L, R, 0, alg, 9, 0, 4).size == 0
        with pytest.raises(asw.AswError):
            ctx.stereoMatching(L, R, 0, alg, 9, 0, 4, strict=True)


def test_method_error_behaviour(ctx):
    L, R, _ = make_pair(32, 40, 4, 1)
    assert ctx.computeAdaptiveWeight_geodesic(L, R, 0, 8, 0, 4).size == 0          # even window (A.cpp:1440-1443)
    assert ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.015, 8, 0, 4).size == 0       # even window (A.cpp:2458-2462)
    assert ctx.computeAdaptiveWeight_bilateralGrid(L, R, 1, 10, 10, 0, 4).size == 0  # RIGHT is out of bounds in the reference
    assert ctx.computeAdaptiveWeight(L, R[:-1], 30, 20, 0, 7, 0, 4).size == 0      # size mismatch


@pytest.mark.parametrize("H,W,D,win,seed", [(32, 44, 6, 5, 1), (40, 56, 8, 9, 2)])
def test_weighted_median_method(ctx, H, W, D, win, seed):
    """computeAdaptiveWeight_WeightedMedian (A.cpp:3228-3383): medians are selected values of the cost volume"""
    L, R, _ = make_pair(H, W, D, seed)
    d, q = ctx.computeAdaptiveWeight_WeightedMedian(L, R, 0, win, 10, 10, 0, D, agg=True, strict=True)
    d_ref, q_ref = orc.asw_weighted_median(L, R, 0, win, 10, 10, 0, D, agg=True)
    assert (q == q_ref).mean() >= 0.999           # a selection: equal except where a crossing flips on a weight ulp
    assert rel_err(q, q_ref) <= 0.05
    assert (d == d_ref).mean() >= AGREE


@pytest.mark.parametrize("H,W,D,win,seed", [(33, 47, 5, 15, 4), (21, 30, 3, 13, 5), (17, 9, 2, 3, 6)])
def test_weighted_median_warp_kernel_edges(ctx, H, W, D, win, seed):
    """the warp-per-pixel selection kernel: the largest window (225 of its 256 key slots), windows larger than the image
    in one direction, widths that are not a multiple of the 4 pixels of a CTA, a 9-element window"""
    L, R, _ = make_pair(H, W, D, seed)
    d, q = ctx.computeAdaptiveWeight_WeightedMedian(L, R, 0, win, 10, 10, 0, D, agg=True, strict=True)
    d_ref, q_ref = orc.asw_weighted_median(L, R, 0, win, 10, 10, 0, D, agg=True)
    assert (q == q_ref).mean() >= 0.999
    assert (d == d_ref).mean() >= AGREE


@pytest.mark.parametrize("H,W,D,win,seed", [(30, 150, 11, 9, 7), (40, 131, 3, 21, 8)])
def test_ncc_tile_kernel_edges(ctx, H, W, D, win, seed):
    """k_ncc_cost_tile: candidate counts that are not a multiple of its 8-candidate chunks, a second 128-column strip that
    is partly outside the image, a window wider than its candidate chunk; bit-compatible with the direct kernel's order"""
    L, R, _ = make_pair(H, W, D, seed)
    for dt in (0, 1):
        v = ctx.computeNCC_volume(L, R, dt, win, 0, D)
        assert np.abs(v - orc.cost_ncc(L, R, 0, D, win, dt)).max() <= 1e-6
        assert (ctx.computeNCC(L, R, dt, win, 0, D, strict=True) == orc.asw_ncc(L, R, dt, win, 0, D)).mean() >= AGREE


@pytest.mark.parametrize("D", [2, 5])
def test_guidedf_batched_slices(ctx, D):
    """GuidedF (6-channel guidance per slice) runs every slice of a chunk in one set of launches: two slices, several slices,
    both views, against the per-slice oracle"""
    L, R, _ = make_pair(37, 52, max(D, 2), 9)
    for dt in (0, 1):
        d, q = ctx.computeAdaptiveWeight_GuidedF(L, R, dt, 1e-6, 7, 0, D, agg=True, strict=True)
        d_ref, q_ref = orc.asw_guidedf(L, R, dt, 1e-6, 7, 0, D, agg=True)
        assert rel_err(q, q_ref) <= REL_TOL
        assert (d == d_ref).mean() >= AGREE


def test_weighted_median_dispatcher_and_limits(ctx):
    L, R, _ = make_pair(32, 44, 6, 3)
    d = ctx.stereoMatching(L, R, 0, asw.ADAPTIVE_WEIGHT_MEDIAN, 7, 0, 6, strict=True)
    assert (d == orc.stereo_matching(L, R, 0, asw.ADAPTIVE_WEIGHT_MEDIAN, 7, 0, 6)).mean() >= AGREE
    assert ctx.computeAdaptiveWeight_WeightedMedian(L, R, 0, 8, 10, 10, 0, 6).size == 0      # even window (A.cpp:3238-3241)
    assert ctx.computeAdaptiveWeight_WeightedMedian(L, R, 1, 7, 10, 10, 0, 6).size == 0      # RIGHT: UB in the reference


@pytest.mark.parametrize("env,alg", [("ASW_GRID_C32", "grid"), ("ASW_GRID_UNFUSED", "grid"), ("ASW_WM_SCAN", "wm"),
                                     ("ASW_TRAD_FAST", "trad"), ("ASW_BLO_TILED", "blo1"), ("ASW_GEO_DIAG_REM", "geo"),
                                     ("ASW_GEO_GENERIC", "geo"), ("ASW_REFINE_DENSE", "refine")])
def test_selectable_fallback_paths(ctx, env, alg):
    """development builds (-DASW_DEV_KERNELS): every kernel the default path replaced stays selectable through an
    environment switch and keeps agreeing with the oracle (32-bit grid counts / global splat, scan-based weighted median,
    clamped traditional kernel, tiled BLO(1), diagonal-kernel remainder and thread-per-pixel geodesic, dense refine)"""
    import os
    needs_dev(ctx)
    L, R, _ = make_pair(48, 72, 33 if alg == "geo" else 8, 41)
    os.environ[env] = "1"
    try:
        if alg == "grid":
            d, e = ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, 10, 10, 0, 8, agg=True, strict=True)
            d_ref, e_ref = orc.asw_bilateral_grid(L, R, 0, 10, 10, 0, 8, agg=True)
            fin = np.isfinite(e_ref)
            assert np.array_equal(np.isfinite(e), fin) and np.array_equal(e[fin], e_ref[fin]) and np.array_equal(d, d_ref)
        elif alg == "wm":
            d, q = ctx.computeAdaptiveWeight_WeightedMedian(L, R, 0, 5, 10, 10, 0, 8, agg=True, strict=True)
            d_ref, q_ref = orc.asw_weighted_median(L, R, 0, 5, 10, 10, 0, 8, agg=True)
            assert np.array_equal(q, q_ref) and np.array_equal(d, d_ref)
        elif alg == "trad":
            d, e = ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 9, 0, 8, agg=True, strict=True)
            d_ref, e_ref = orc.asw_traditional(L, R, 30, 20, 0, 9, 0, 8, agg=True)
            assert rel_err(e, e_ref) <= REL_TOL and (d == d_ref).mean() >= AGREE
        elif alg == "blo1":
            d, q = ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.015, 9, 0, 8, agg=True, strict=True)
            d_ref, q_ref = orc.asw_blo1(L, R, 0, 0.015, 9, 0, 8, agg=True)
            assert rel_err(q, q_ref) <= REL_TOL and (d == d_ref).mean() >= AGREE
        elif alg == "geo":
            d, e = ctx.computeAdaptiveWeight_geodesic(L, R, 0, 7, 0, 33, agg=True, strict=True)
            d_ref, e_ref = orc.asw_geodesic(L, R, 0, 7, 0, 33, agg=True)
            assert rel_err(e, e_ref) <= REL_TOL and (d == d_ref).mean() >= AGREE
        else:
            out = ctx.guidedf2_lr_refine(L, R, 1e-4, 9, 0, 8)
            ref, _ = orc.guidedf2_lr_refine(L, R, 1e-4, 9, 0, 8)
            assert (out == ref).mean() >= 0.995
    finally:
        del os.environ[env]
