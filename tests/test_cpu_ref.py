"""The oracle pinned to the reference's OWN code.

oracle/_ref/libasw_ref.so is /root/reference's unmodified aswStereoMatch/methods/aswMethods.cpp compiled with g++
against an OpenCV stand-in (oracle/refshim/: cv::Mat, the lazy cv::MatExpr algebra with OpenCV's fusion rules, and the
primitives with OpenCV-4.13 arithmetic).  Three legs:
  1. the stand-in's primitives and MatExpr lowerings against the REAL OpenCV (Python cv2 4.13), bit-exact;
  2. every method of the C restatement (oracle/asw_oracle.c, what the GPU tests compare with) against the reference's
     own code on the same seeded inputs, bit-exact (maps) -- including the behaviour-defining quirks;
  3. the committed golden fixture tests/golden/ref_methods_44x60_d6.npz (tools/make_ref_golden.py, made by the
     reference's own code) against the C restatement -- this leg needs neither /root/reference nor the built library.
"""
import numpy as np
import pytest

from aswstereomatch_b200.synth import make_pair
from oracle import orc, ref

have_ref = ref.available()
needs_ref = pytest.mark.skipif(not have_ref, reason="oracle/_ref/libasw_ref.so not built and /root/reference absent")
try:
    import cv2
    cv2.setNumThreads(1)     # OpenCV's boxFilter restarts its column sums per stripe; one stripe = the stated order
except Exception:            # pragma: no cover
    cv2 = None
needs_cv2 = pytest.mark.skipif(cv2 is None, reason="cv2 not importable")


@pytest.fixture(scope="module")
def gold():
    return np.load("tests/golden/ref_methods_44x60_d6.npz")


# ---------------------------------------------------------------------------------------------------------------
# 1. the OpenCV stand-in against the real OpenCV
# ---------------------------------------------------------------------------------------------------------------
@needs_ref
@needs_cv2
def test_shim_primitives_match_cv2():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    assert np.array_equal(ref.shim_bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    k = np.array([[-3, 0, 3], [-10, 0, 10], [-3, 0, 3]], np.int8)
    assert np.array_equal(ref.shim_scharr_x(img), cv2.filter2D(img, cv2.CV_32F, k))
    assert np.array_equal(ref.shim_normalize_u8c3(img), cv2.normalize(img, None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F))
    f = (rng.random((37, 53)) * 13000 + 5100).astype(np.float32)
    assert np.array_equal(ref.shim_normalize_f32(f), cv2.normalize(f, None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F))
    for ksz in (3, 5, 7, 9, 15, 35):
        assert np.array_equal(ref.shim_box_filter(f, ksz), cv2.boxFilter(f, cv2.CV_32F, (ksz, ksz))), ksz
    e = (-rng.random((35, 35)) * 80).astype(np.float32)
    got, want = ref.shim_exp(e), cv2.exp(e)
    assert np.abs(got - want).max() <= 1.2e-7 * want.max()          # cv::exp is a 1e-7-relative polynomial


@needs_ref
@needs_cv2
def test_shim_matexpr_lowerings_match_cv2_primitives():
    """(a + b + c) / 3, alpha A + beta B and the inverted truncation of A.cpp:459-484 lower to the OpenCV calls of SURVEY
    Appendix B; the stand-in's MatExpr engine must produce what those real calls produce"""
    rng = np.random.default_rng(6)
    a, b, c = (rng.integers(0, 256, (31, 47), dtype=np.uint8) for _ in range(3))
    assert np.array_equal(ref.shim_mean3_u8(a, b, c), cv2.addWeighted(cv2.add(a, b), 1 / 3, c, 1 / 3, 0))
    fa, fb, fc = ((rng.random((31, 47)) * 4000).astype(np.float32) for _ in range(3))
    assert np.array_equal(ref.shim_mean3_f32(fa, fb, fc), cv2.addWeighted(cv2.add(fa, fb), 1 / 3, fc, 1 / 3, 0))
    assert np.array_equal(ref.shim_blend_f32(fa, 0.6, fb, 0.4), cv2.addWeighted(fa, 0.6, fb, 0.4, 0))
    mask = cv2.compare(a, 10.0, cv2.CMP_GT)
    want = cv2.scaleAdd(mask, 10.0 / 255, cv2.multiply(a, mask, scale=1 / 255))
    assert np.array_equal(ref.shim_trunc_u8(a, 10.0), want)
    assert np.array_equal(want, np.where(a > 10, np.minimum(a.astype(int) + 10, 255), 0))     # Appendix A-1


# ---------------------------------------------------------------------------------------------------------------
# 2. the C restatement against the reference's own code
# ---------------------------------------------------------------------------------------------------------------
CASES = [(48, 64, 8, 7), (35, 35, 3, 2), (41, 57, 5, 9)]


@needs_ref
@pytest.mark.parametrize("H,W,D,seed", CASES)
def test_stage_level_equals_reference(H, W, D, seed):
    L, R, _ = make_pair(H, W, D, seed)
    assert np.array_equal(orc.cost_tad_cg(L, R, 0, D, 0), ref.cost_tad_cg(L, R, 0, D, 0))
    assert np.array_equal(orc.cost_tad_cg(L, R, 2, D, 0), ref.cost_tad_cg(L, R, 2, D, 0))
    assert np.array_equal(orc.cost_tad_cg_padded(L, R, 0, D, 7, 0), ref.cost_tad_cg_padded(L, R, 0, D, 7, 0))
    for dt in (0, 1):
        assert np.array_equal(orc.cost_sad_box(L, R, 0, D, 9, dt), ref.cost_sad_box(L, R, 0, D, 9, dt))
    cost = orc.cost_tad_cg(L, R, 0, D, 0)
    for r, eps in ((5, 1e-4), (9, 1e-6)):
        assert np.array_equal(orc.guided_filter(L, cost[D // 2], r, eps), ref.guided_filter(L, cost[D // 2], r, eps))
    assert np.array_equal(orc.geodesic_dist(L, 7), ref.geodesic_dist(L, 7))


@needs_ref
@pytest.mark.parametrize("H,W,D,seed", CASES)
def test_methods_equal_reference(H, W, D, seed):
    L, R, _ = make_pair(H, W, D, seed)
    for dt in (0, 1):          # the loop-only methods define DISPARITY_RIGHT too
        assert np.array_equal(orc.asw_traditional(L, R, 30, 20, dt, 9, 0, D), ref.asw_traditional(L, R, 30, 20, dt, 9, 0, D))
        assert np.array_equal(orc.asw_geodesic(L, R, dt, 7, 0, D), ref.asw_geodesic(L, R, dt, 7, 0, D))
        assert np.array_equal(orc.asw_guidedf(L, R, dt, 1e-6, 9, 0, D), ref.asw_guidedf(L, R, dt, 1e-6, 9, 0, D))
        assert np.array_equal(orc.asw_blo1(L, R, dt, 0.015, 9, 0, D), ref.asw_blo1(L, R, dt, 0.015, 9, 0, D))
    assert np.array_equal(orc.asw_bilateral_grid(L, R, 0, 10, 10, 0, D), ref.asw_bilateral_grid(L, R, 0, 10, 10, 0, D))
    assert np.array_equal(orc.asw_guidedf2(L, R, 0, 1e-4, 9, 0, D), ref.asw_guidedf2(L, R, 0, 1e-4, 9, 0, D))
    assert np.array_equal(orc.asw_guidedf2(L, R, 0, 1e-6, 5, 0, D), ref.asw_guidedf2(L, R, 0, 1e-6, 5, 0, D))
    assert np.array_equal(orc.asw_weighted_median(L, R, 0, 5, 10, 10, 0, D), ref.asw_weighted_median(L, R, 0, 5, 10, 10, 0, D))


@needs_ref
def test_config1_window_equals_reference():
    """35 x 35 window (config 1 / 3b / 4's): image smaller than the window in one direction, every border clamp active"""
    L, R, _ = make_pair(30, 52, 4, 13)
    assert np.array_equal(orc.asw_traditional(L, R, 30, 20, 0, 35, 0, 4), ref.asw_traditional(L, R, 30, 20, 0, 35, 0, 4))
    assert np.array_equal(orc.asw_blo1(L, R, 0, 0.015, 35, 0, 4), ref.asw_blo1(L, R, 0, 0.015, 35, 0, 4))
    L, R, _ = make_pair(24, 40, 3, 14)
    assert np.array_equal(orc.asw_geodesic(L, R, 0, 35, 0, 3), ref.asw_geodesic(L, R, 0, 35, 0, 3))


@needs_ref
def test_dispatcher_equals_reference():
    """stereoMatching with its own literals (A.cpp:46-88)"""
    L, R, _ = make_pair(40, 56, 6, 17)
    for alg in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11):
        assert np.array_equal(orc.stereo_matching(L, R, 0, alg, 9, 0, 6), ref.stereo_matching(L, R, 0, alg, 9, 0, 6)), alg


@needs_ref
@pytest.mark.parametrize("H,W,D,win,seed", [(40, 56, 6, 9, 3), (33, 47, 5, 5, 8), (30, 40, 4, 15, 2)])
def test_direct8_ncc_guidedf3_equal_reference(H, W, D, win, seed):
    """rows f-3 / f-4: the 8-direction method (LEFT; its RIGHT branch indexes out of bounds), the NCC cost in both overloads
    and GuidedF_3, both views"""
    L, R, _ = make_pair(H, W, D, seed)
    assert np.array_equal(orc.asw_direct8(L, R, 0, win, 0, D), ref.asw_direct8(L, R, 0, win, 0, D))
    for dt in (0, 1):
        assert np.array_equal(orc.cost_ncc(L, R, 0, D, win, dt), ref.cost_ncc(L, R, 0, D, win, dt))
        assert np.array_equal(orc.asw_guidedf3(L, R, dt, 1e-6, win, 0, D), ref.asw_guidedf3(L, R, dt, 1e-6, win, 0, D))
        assert np.array_equal(orc.asw_ncc(L, R, dt, win, 0, D), ref.asw_ncc(L, R, dt, win, 0, D))
    assert not ref.asw_ncc(L, R, 1, win, 0, D).any()                # the RIGHT branch of the Mat overload never writes


@needs_ref
def test_reference_right_branch_of_tad_cost_throws():
    """SURVEY Appendix A-3, now observed on the reference's own code: CV_32F.mul(CV_8U / 255) without a dtype raises, so
    GuidedF_2 and the weighted-median method are LEFT-only; the C restatement's RIGHT view is the mirrored LEFT formula"""
    L, R, _ = make_pair(40, 56, 6, 3)
    with pytest.raises(ref.RefError, match="different types"):
        ref.asw_guidedf2(L, R, 1, 1e-4, 9, 0, 6)
    with pytest.raises(ref.RefError, match="different types"):
        ref.asw_weighted_median(L, R, 1, 5, 10, 10, 0, 6)
    with pytest.raises(ref.RefError, match="different types"):
        ref.cost_tad_cg(L, R, 0, 6, 1)


@needs_ref
def test_reference_argument_errors_are_empty_mats():
    L, R, _ = make_pair(32, 40, 4, 3)
    assert ref.asw_geodesic(L, R, 0, 8, 0, 4) is None               # even window (A.cpp:1440-1443)
    assert ref.asw_weighted_median(L, R, 0, 8, 10, 10, 0, 4) is None   # A.cpp:3238-3241
    with pytest.raises(ValueError):
        orc.asw_geodesic(L, R, 0, 8, 0, 4)


# ---------------------------------------------------------------------------------------------------------------
# 3. the committed fixture made by the reference's own code (tools/make_ref_golden.py)
# ---------------------------------------------------------------------------------------------------------------
def test_golden_from_reference_own_code(gold):
    L, R, D = gold["L"], gold["R"], int(gold["D"])
    eq = np.array_equal
    assert eq(orc.cost_tad_cg(L, R, 0, D, 0), gold["cost_tad_cg"])
    assert eq(orc.cost_tad_cg_padded(L, R, 0, D, 5, 0), gold["cost_tad_cg_padded_w5"])
    assert eq(orc.cost_sad_box(L, R, 0, D, 7, 0), gold["cost_sad_box_w7"])
    assert eq(orc.cost_sad_box(L, R, 0, D, 7, 1), gold["cost_sad_box_w7_right"])
    assert eq(orc.guided_filter(L, gold["cost_tad_cg"][2], 9, 1e-4), gold["gf_slice2_r9"])
    assert eq(orc.geodesic_dist(L, 7), gold["geodesic_dist_w7"])
    assert eq(orc.asw_traditional(L, R, 30, 20, 0, 35, 0, D), gold["traditional_w35"])
    assert eq(orc.asw_traditional(L, R, 30, 20, 1, 9, 0, D), gold["traditional_w9_right"])
    assert eq(orc.asw_geodesic(L, R, 0, 35, 0, D), gold["geodesic_w35"])
    assert eq(orc.asw_geodesic(L, R, 1, 7, 0, D), gold["geodesic_w7_right"])
    assert eq(orc.asw_bilateral_grid(L, R, 0, 10, 10, 0, D), gold["grid_s10_r10"])
    assert eq(orc.asw_blo1(L, R, 0, 0.015, 35, 0, D), gold["blo1_w35"])
    assert eq(orc.asw_blo1(L, R, 1, 0.015, 9, 0, D), gold["blo1_w9_right"])
    assert eq(orc.asw_guidedf(L, R, 0, 1e-6, 9, 0, D), gold["guidedf_w9"])
    assert eq(orc.asw_guidedf(L, R, 1, 1e-6, 9, 0, D), gold["guidedf_w9_right"])
    assert eq(orc.asw_guidedf2(L, R, 0, 1e-4, 9, 0, D), gold["guidedf2_w9_eps1e-4"])
    assert eq(orc.asw_guidedf2(L, R, 0, 1e-6, 15, 0, D), gold["guidedf2_w15_eps1e-6"])
    assert eq(orc.asw_weighted_median(L, R, 0, 7, 10, 10, 0, D), gold["wmedian_w7"])
    assert eq(orc.asw_direct8(L, R, 0, 9, 0, D), gold["direct8_w9"])
    assert eq(orc.cost_ncc(L, R, 0, D, 7, 0), gold["cost_ncc_w7"])
    assert eq(orc.asw_guidedf3(L, R, 0, 1e-6, 9, 0, D), gold["guidedf3_w9"])
    assert eq(orc.asw_guidedf3(L, R, 1, 1e-6, 9, 0, D), gold["guidedf3_w9_right"])
    assert eq(orc.asw_ncc(L, R, 0, 9, 0, D), gold["ncc_w9"])
    for alg in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11):
        assert eq(orc.stereo_matching(L, R, 0, alg, 9, 0, D), gold[f"dispatch_alg{alg}_w9"]), alg
