"""The CUDA path (through the C ABI) against the reference's OWN code.

Two sources of reference output: the committed fixture tests/golden/ref_methods_44x60_d6.npz, produced by
oracle/_ref/libasw_ref.so (= /root/reference's unmodified aswMethods.cpp over the OpenCV stand-in; script
tools/make_ref_golden.py), and -- when that library travelled to this box -- fresh runs of it on other seeds.
Bars: integer stages bit-exact, float costs within 1e-4 of the slice maximum, disparity maps >= 99.9 %
(the loop-only methods accumulate in double on the CPU and in float on the GPU: ties may flip, nothing else).
"""
import numpy as np
import pytest

import aswstereomatch_b200 as asw
from aswstereomatch_b200.synth import make_pair
from oracle import ref

pytestmark = pytest.mark.gpu
AGREE = 0.999


@pytest.fixture(scope="module")
def gold():
    return np.load("tests/golden/ref_methods_44x60_d6.npz")


def agree(a, b):
    assert a.shape == b.shape
    return float((a == b).mean())


def rel_err(a, b):
    a = a.reshape(a.shape[0], -1).astype(np.float64)
    b = b.reshape(b.shape[0], -1).astype(np.float64)
    return float((np.abs(a - b) / np.maximum(np.abs(b).max(axis=1, keepdims=True), 1e-30)).max())


def test_stages_against_reference_golden(ctx, gold):
    L, R, D = gold["L"], gold["R"], int(gold["D"])
    assert np.array_equal(ctx.computeSimilarity(L, R, 0.4, 10, 50, 0, 0, D), gold["cost_tad_cg"])          # bit-exact
    assert rel_err(ctx.getCostSAD(L, R, 0, 7, 0, D), gold["cost_sad_box_w7"]) <= 1e-6
    assert rel_err(ctx.getCostSAD(L, R, 1, 7, 0, D), gold["cost_sad_box_w7_right"]) <= 1e-6
    q = ctx.getGuidedFilter(L, gold["cost_tad_cg"][2], 9, 1e-4)
    assert np.abs(q - gold["gf_slice2_r9"]).max() <= 1e-4
    assert np.array_equal(ctx.getGeodesicDist(L, 7), gold["geodesic_dist_w7"])                              # integer DP


def test_methods_against_reference_golden(ctx, gold):
    L, R, D = gold["L"], gold["R"], int(gold["D"])
    assert agree(ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 35, 0, D, strict=True), gold["traditional_w35"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight(L, R, 30, 20, 1, 9, 0, D, strict=True), gold["traditional_w9_right"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_geodesic(L, R, 0, 35, 0, D, strict=True), gold["geodesic_w35"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_geodesic(L, R, 1, 7, 0, D, strict=True), gold["geodesic_w7_right"]) >= AGREE
    assert np.array_equal(ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, 10, 10, 0, D, strict=True), gold["grid_s10_r10"])
    assert agree(ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.015, 35, 0, D, strict=True), gold["blo1_w35"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_BLO1(L, R, 1, 0.015, 9, 0, D, strict=True), gold["blo1_w9_right"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_GuidedF(L, R, 0, 1e-6, 9, 0, D, strict=True), gold["guidedf_w9"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_GuidedF(L, R, 1, 1e-6, 9, 0, D, strict=True), gold["guidedf_w9_right"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-4, 9, 0, D, strict=True), gold["guidedf2_w9_eps1e-4"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-6, 15, 0, D, strict=True), gold["guidedf2_w15_eps1e-6"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_WeightedMedian(L, R, 0, 7, 10, 10, 0, D, strict=True), gold["wmedian_w7"]) >= AGREE


def test_ncc_family_against_reference_golden(ctx, gold):
    L, R, D = gold["L"], gold["R"], int(gold["D"])
    assert np.abs(ctx.computeNCC_volume(L, R, 0, 7, 0, D) - gold["cost_ncc_w7"]).max() <= 1e-6      # same double sums: ~exact
    assert agree(ctx.computeAdaptiveWeight_GuidedF_3(L, R, 0, 1e-6, 9, 0, D, strict=True), gold["guidedf3_w9"]) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_GuidedF_3(L, R, 1, 1e-6, 9, 0, D, strict=True), gold["guidedf3_w9_right"]) >= AGREE
    assert agree(ctx.computeNCC(L, R, 0, 9, 0, D, strict=True), gold["ncc_w9"]) >= AGREE
    assert not ctx.computeNCC(L, R, 1, 9, 0, D, strict=True).any()
    assert agree(ctx.computeAdaptiveWeight_direct8(L, R, 0, 9, 0, D, strict=True), gold["direct8_w9"]) >= AGREE


@pytest.mark.parametrize("alg", [2, 3, 4, 5, 6, 7, 8, 9, 10, 11])
def test_dispatcher_against_reference_golden(ctx, gold, alg):
    L, R, D = gold["L"], gold["R"], int(gold["D"])
    assert agree(ctx.stereoMatching(L, R, 0, alg, 9, 0, D, strict=True), gold[f"dispatch_alg{alg}_w9"]) >= AGREE


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libasw_ref.so did not travel to this box")
@pytest.mark.parametrize("H,W,D,seed", [(64, 96, 8, 31), (52, 70, 5, 32)])
def test_fresh_reference_runs(ctx, H, W, D, seed):
    """the reference's own code, run here on other seeds and sizes"""
    L, R, _ = make_pair(H, W, D, seed)
    assert np.array_equal(ctx.computeSimilarity(L, R, 0.4, 10, 50, 0, 0, D), ref.cost_tad_cg(L, R, 0, D, 0))
    assert agree(ctx.computeAdaptiveWeight_GuidedF_2(L, R, 0, 1e-4, 9, 0, D, strict=True), ref.asw_guidedf2(L, R, 0, 1e-4, 9, 0, D)) >= AGREE
    assert agree(ctx.computeAdaptiveWeight(L, R, 30, 20, 0, 15, 0, D, strict=True), ref.asw_traditional(L, R, 30, 20, 0, 15, 0, D)) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_geodesic(L, R, 0, 9, 0, D, strict=True), ref.asw_geodesic(L, R, 0, 9, 0, D)) >= AGREE
    assert agree(ctx.computeAdaptiveWeight_BLO1(L, R, 0, 0.015, 15, 0, D, strict=True), ref.asw_blo1(L, R, 0, 0.015, 15, 0, D)) >= AGREE
    assert np.array_equal(ctx.computeAdaptiveWeight_bilateralGrid(L, R, 0, 10, 10, 0, D, strict=True), ref.asw_bilateral_grid(L, R, 0, 10, 10, 0, D))
    # the reference's DISPARITY_RIGHT of the TAD-cost methods throws; the library reports UNSUPPORTED (empty Mat in the shim)
    with pytest.raises(ref.RefError):
        ref.asw_guidedf2(L, R, 1, 1e-4, 9, 0, D)
    assert ctx.computeAdaptiveWeight_GuidedF_2(L, R, 1, 1e-4, 9, 0, D).size == 0
