"""CPU-only tests (no GPU): the oracle against the committed cv2 golden fixtures and its own invariants,
the C-ABI library's symbol table, and the host-side argument handling that needs no device."""
import ctypes
import os
import re

import numpy as np
import pytest

from aswstereomatch_b200.synth import make_batch, make_pair
from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cv2_stages_40x56_d8.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


# ---------------------------------------------------------------------------------------------
# oracle == cv2 4.13 on the OpenCV-primitive stages (fixtures written by oracle/cv2_restatement.py)
# ---------------------------------------------------------------------------------------------
def test_gray_golden(gold):
    assert np.array_equal(orc.bgr2gray(gold["L"]), gold["gray_L"])


def test_primitives_golden(gold):
    assert np.array_equal(orc.box_filter(gold["cost_tad_cg"][2], 9), gold["box9"])
    assert np.array_equal(orc.normalize_f32(gold["cost_tad_cg"][2]), gold["norm_cost2"])
    assert np.array_equal(orc.normalize_u8(gold["L"]), gold["norm_L"])


def test_cost_tad_cg_golden(gold):
    assert np.array_equal(orc.cost_tad_cg(gold["L"], gold["R"], 0, 8), gold["cost_tad_cg"])


def test_cost_sad_box_golden(gold):
    assert np.array_equal(orc.cost_sad_box(gold["L"], gold["R"], 0, 8, 5), gold["cost_sad_box_w5"])


def test_guided_filter_golden(gold):
    q = orc.guided_filter(gold["L"], gold["cost_tad_cg"][3], 5, 1e-4)
    assert np.array_equal(q, gold["gf_slice3_r5"])


def test_guidedf2_golden(gold):
    d, q = orc.asw_guidedf2(gold["L"], gold["R"], 0, 1e-4, 5, 0, 8, agg=True)
    assert np.array_equal(q, gold["guidedf2_q"])
    assert np.array_equal(d, gold["guidedf2_disp"])


def test_guidedf_six_channel_golden(gold):
    d, q = orc.asw_guidedf(gold["L"], gold["R"], 0, 1e-4, 5, 0, 8, agg=True)
    assert np.array_equal(q, gold["guidedf_q"])
    assert np.array_equal(d, gold["guidedf_disp"])


def test_blo1_golden(gold):
    d, q = orc.asw_blo1(gold["L"], gold["R"], 0, 0.015, 7, 0, 8, agg=True)
    assert np.array_equal(q, gold["blo1_q"])
    assert np.array_equal(d, gold["blo1_disp"])


# ---------------------------------------------------------------------------------------------
# behaviour-defining quirks of the reference hold in the oracle
# ---------------------------------------------------------------------------------------------
def test_quirk_inverted_colour_truncation_and_gradient_offset():
    """identical images: colour cost 0, gradient cost 255*T_G -> 0.6*0 + 0.4*12750 = 5100 (A.cpp:461-482)"""
    L, _, _ = make_pair(24, 32, 4, 0)
    c = orc.cost_tad_cg(L, L, 0, 1)
    assert np.all(c == 5100.0)


def test_quirk_colour_cost_jumps_past_threshold():
    L = np.zeros((8, 8, 3), np.uint8)
    R = np.zeros((8, 8, 3), np.uint8)
    R[:] = 10                       # mean AD = 10 -> not > 10 -> colour cost 0
    assert np.all(orc.cost_tad_cg(L, R, 0, 1) == 5100.0)
    R[:] = 11                       # mean AD = 11 > 10 -> 11 + 10 = 21 -> 0.6*21 + 5100
    assert np.allclose(orc.cost_tad_cg(L, R, 0, 1), 0.6 * 21 + 5100, rtol=0, atol=1e-3)


def test_quirk_candidate_counts():
    L, R, _ = make_pair(24, 32, 4, 1)
    _, e = orc.asw_traditional(L, R, 30, 20, 0, 5, 0, 4, agg=True)
    assert e.shape[0] == 5          # D+1 (A.cpp:1021, 1074)
    _, e = orc.asw_guidedf2(L, R, 0, 1e-4, 5, 0, 4, agg=True)
    assert e.shape[0] == 4          # D   (A.cpp:3036)


def test_quirk_geodesic_weights_are_raw_distances():
    L, _, _ = make_pair(12, 14, 2, 2)
    d = orc.geodesic_dist(L, 5)
    assert np.all(d[:, :, 2, 2] == 0)                    # centre distance 0 -> centre weight 0 (A.cpp:1488)
    assert np.all(d >= 0) and np.all(d == np.round(d))   # exact integers


def test_wta_rules():
    vol = np.array([[[1.0, np.nan, 2.0, np.inf]], [[1.0, np.nan, 1.0, np.inf]], [[0.5, 3.0, 1.0, np.inf]]], np.float32)
    d = orc.wta(vol, 10)
    assert d.tolist() == [[12.0, 12.0, 11.0, 0.0]]       # lowest d on ties, NaN/inf never win, untouched = 0


def test_weighted_median_selection_rule():
    """A.cpp:3276-3304: the element BEFORE the one whose partial sum crosses total/2"""
    img = np.full((5, 5, 3), 100, np.uint8)              # uniform colour -> colour weight 1 everywhere
    filled = np.zeros((5, 5), np.float32)
    filled[:, 3:] = 7.0
    valid = np.ones((5, 5), np.uint8)
    valid[2, 2] = 0
    out = orc.wmedian_refine(img, filled, valid, 3, 1e9, 10)   # flat spatial weight: 6 zeros then 3 sevens
    assert out[2, 2] == 0.0
    filled[:, 1:] = 7.0                                  # 3 zeros, 6 sevens: crossing at the 5th element -> previous = 7
    assert orc.wmedian_refine(img, filled, valid, 3, 1e9, 10)[2, 2] == 7.0


def test_lr_check_and_fill_spec():
    dl = np.array([[2, 2, 2, 5, 2]], np.float32)
    dr = np.array([[2, 2, 2, 2, 2]], np.float32)
    v = orc.lr_check(dl, dr, 0.0)
    assert v.tolist() == [[1, 1, 1, 0, 1]]
    assert orc.fill_invalid(dl, v).tolist() == [[2, 2, 2, 2, 2]]
    v0 = np.zeros_like(v)
    assert np.array_equal(orc.fill_invalid(dl, v0), dl)  # no valid pixel on the row: keep


def test_oracle_thread_count_independent():
    L, R, _ = make_pair(40, 56, 6, 3)
    n0 = orc.num_threads()
    orc.set_num_threads(1)
    a = orc.asw_guidedf2(L, R, 0, 1e-4, 9, 0, 6)
    b1 = orc.asw_bilateral_grid(L, R, 0, 10, 10, 0, 6)
    orc.set_num_threads(max(2, n0))
    assert np.array_equal(a, orc.asw_guidedf2(L, R, 0, 1e-4, 9, 0, 6))
    assert np.array_equal(b1, orc.asw_bilateral_grid(L, R, 0, 10, 10, 0, 6))
    orc.set_num_threads(n0)


def test_all_methods_beat_chance_on_synthetic():
    L, R, gt = make_pair(48, 64, 8, 11)
    acc = lambda d: float(np.mean(np.abs(d - gt) <= 1))
    assert acc(orc.asw_traditional(L, R, 30, 20, 0, 9, 0, 8)) > 0.9
    assert acc(orc.asw_geodesic(L, R, 0, 9, 0, 8)) > 0.9
    assert acc(orc.asw_blo1(L, R, 0, 0.015, 9, 0, 8)) > 0.9
    assert acc(orc.asw_guidedf2(L, R, 0, 1e-4, 9, 0, 8)) > 0.9
    assert acc(orc.asw_guidedf(L, R, 0, 1e-4, 9, 0, 8)) > 0.85
    assert acc(orc.asw_weighted_median(L, R, 0, 9, 10, 10, 0, 8)) > 0.9
    out, _ = orc.guidedf2_lr_refine(L, R, 1e-4, 9, 0, 8)
    assert acc(out) > 0.93


def test_oracle_error_codes():
    L, R, _ = make_pair(16, 20, 2, 0)
    with pytest.raises(ValueError):
        orc.asw_geodesic(L, R, 0, 4, 0, 2)               # even window (A.cpp:1440-1443)
    with pytest.raises(ValueError):
        orc.cost_sad_box(L, R, 0, 2, 4)                  # even window (A.cpp:2458-2462)
    with pytest.raises(ValueError):
        orc.asw_bilateral_grid(L, R, 1, 10, 10, 0, 2)    # RIGHT out of bounds in the reference


def test_synth_is_deterministic_and_textured():
    a = make_pair(32, 48, 8, 5)
    b = make_pair(32, 48, 8, 5)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert a[0].std() > 20
    Ls, Rs = make_batch(5, 32, 48, 8, distinct=2)
    assert len(Ls) == 5 and not np.array_equal(Ls[0], Ls[2])


# ---------------------------------------------------------------------------------------------
# the C-ABI library: loads without a GPU and exports every symbol include/asw/asw.h declares
# ---------------------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol(built):
    import aswstereomatch_b200 as asw
    hdr = open(os.path.join(ROOT, "include", "asw", "asw.h")).read()
    declared = set(re.findall(r"\b(asw_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"asw_status"}
    lib = ctypes.CDLL(asw.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(asw.EXPORTS), declared ^ set(asw.EXPORTS)
    assert b"sm_100a" in asw.load_library().asw_version()


def test_no_cpu_fallback_without_device(built):
    """the product fails loudly when no CUDA device is usable"""
    import aswstereomatch_b200 as asw
    lib = asw.load_library()
    if lib.asw_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(asw.AswError):
        asw.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "aswstereomatch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in txt and "import oracle" not in txt and "liborc" not in txt, f


def test_driver_post_processing_matches_cv2():
    """aswStereoMatch.cpp:97-98 (convertTo(CV_8UC1) + normalize(0, 255, NORM_MINMAX)) against the real OpenCV"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for lo, hi in ((0, 63), (3, 200), (17, 17), (0, 255), (40, 41)):
        d = rng.integers(lo, hi + 1, (45, 61)).astype(np.float32)
        want = cv2.normalize(d.astype(np.uint8), None, 0, 255, cv2.NORM_MINMAX)
        assert np.array_equal(orc.disparity_to_u8(d), want), (lo, hi)
