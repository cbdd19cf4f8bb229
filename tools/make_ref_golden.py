"""Golden vectors produced by the reference's OWN code: oracle/_ref/libasw_ref.so is /root/reference's unmodified
aswMethods.cpp compiled against the OpenCV stand-in (oracle/refshim/).  Run in the build container (where
/root/reference exists); the fixture travels to the GPU box, the reference sources do not.

    python tools/make_ref_golden.py        ->  tests/golden/ref_methods_44x60_d6.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aswstereomatch_b200.synth import make_pair   # noqa: E402
from oracle import ref                            # noqa: E402

H, W, D, SEED = 44, 60, 6, 21


def main():
    assert ref.sources_present(), "needs /root/reference"
    ref.build(force=True)
    L, R, gt = make_pair(H, W, D, SEED)
    out = dict(L=L, R=R, gt=gt, D=np.int32(D))
    out["cost_tad_cg"] = ref.cost_tad_cg(L, R, 0, D, 0)                       # computeSimilarity 7-arg, LEFT
    out["cost_tad_cg_padded_w5"] = ref.cost_tad_cg_padded(L, R, 0, D, 5, 0)   # computeSimilarity 8-arg, LEFT
    out["cost_sad_box_w7"] = ref.cost_sad_box(L, R, 0, D, 7, 0)               # getCostSAD_d as BLO1 calls it
    out["cost_sad_box_w7_right"] = ref.cost_sad_box(L, R, 0, D, 7, 1)
    out["gf_slice2_r9"] = ref.guided_filter(L, out["cost_tad_cg"][2], 9, 1e-4)
    out["geodesic_dist_w7"] = ref.geodesic_dist(L, 7)
    out["traditional_w35"] = ref.asw_traditional(L, R, 30, 20, 0, 35, 0, D)
    out["traditional_w9_right"] = ref.asw_traditional(L, R, 30, 20, 1, 9, 0, D)
    out["direct8_w9"] = ref.asw_direct8(L, R, 0, 9, 0, D)
    out["geodesic_w35"] = ref.asw_geodesic(L, R, 0, 35, 0, D)
    out["geodesic_w7_right"] = ref.asw_geodesic(L, R, 1, 7, 0, D)
    out["grid_s10_r10"] = ref.asw_bilateral_grid(L, R, 0, 10, 10, 0, D)
    out["blo1_w35"] = ref.asw_blo1(L, R, 0, 0.015, 35, 0, D)
    out["blo1_w9_right"] = ref.asw_blo1(L, R, 1, 0.015, 9, 0, D)
    out["guidedf_w9"] = ref.asw_guidedf(L, R, 0, 1e-6, 9, 0, D)
    out["guidedf_w9_right"] = ref.asw_guidedf(L, R, 1, 1e-6, 9, 0, D)
    out["guidedf2_w9_eps1e-4"] = ref.asw_guidedf2(L, R, 0, 1e-4, 9, 0, D)
    out["guidedf2_w15_eps1e-6"] = ref.asw_guidedf2(L, R, 0, 1e-6, 15, 0, D)
    out["wmedian_w7"] = ref.asw_weighted_median(L, R, 0, 7, 10, 10, 0, D)
    out["cost_ncc_w7"] = ref.cost_ncc(L, R, 0, D, 7, 0)                       # computeNCC vector overload
    out["guidedf3_w9"] = ref.asw_guidedf3(L, R, 0, 1e-6, 9, 0, D)
    out["guidedf3_w9_right"] = ref.asw_guidedf3(L, R, 1, 1e-6, 9, 0, D)
    out["ncc_w9"] = ref.asw_ncc(L, R, 0, 9, 0, D)                             # computeNCC Mat overload
    for alg in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11):                              # the dispatcher with its literals, win 9
        out[f"dispatch_alg{alg}_w9"] = ref.stereo_matching(L, R, 0, alg, 9, 0, D)
    for k, v in out.items():
        assert v is not None, k
    path = os.path.join(ROOT, "tests", "golden", f"ref_methods_{H}x{W}_d{D}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
