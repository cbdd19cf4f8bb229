"""ctypes binding of oracle/_ref/libasw_ref.so: the reference's OWN aswMethods.cpp, compiled unmodified
against the OpenCV stand-in in oracle/refshim/ (recipe: oracle/Makefile, target `ref`).

TEST INFRASTRUCTURE ONLY, like oracle/orc.py: imported by tests/ and by bench.py's CPU legs; the product
package never imports it.  /root/reference exists only in the build container: there the library is (re)built
from the reference's sources where they lie; on the GPU box the prebuilt file (git-ignored, shipped by gpurun)
is loaded as is.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libasw_ref.so")
REFERENCE = os.environ.get("ASW_REFERENCE", "/root/reference")
_LIB = None

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


class RefError(RuntimeError):
    """the reference threw (cv::Exception) -- code -2 -- or misbehaved"""


def sources_present():
    return os.path.exists(os.path.join(REFERENCE, "aswStereoMatch", "methods", "aswMethods.cpp"))


def build(force=False):
    """(re)build from /root/reference when it is there; otherwise keep the prebuilt library"""
    if sources_present():
        if force and os.path.exists(SO):
            os.remove(SO)
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", f"REFERENCE={REFERENCE}"], stdout=subprocess.DEVNULL)
    return SO if os.path.exists(SO) else None


def available():
    return build() is not None


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RefError("oracle/_ref/libasw_ref.so is missing and /root/reference is not here to build it")
        _LIB = C.CDLL(so)
        _LIB.ref_last_error.restype = C.c_char_p
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _chk(rc, what):
    """0 ok; -1 = empty Mat (returned as None by callers); anything else raises"""
    if rc == -1:
        return False
    if rc != 0:
        raise RefError(f"reference {what}: rc={rc}: {lib().ref_last_error().decode(errors='replace')}")
    return True


def _method(name, L, R, args):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    disp = np.zeros((H, W), np.float32)
    fn = getattr(lib(), name)
    return disp if _chk(fn(pl, pr, H, W, *args, disp.ctypes.data_as(f32p)), name) else None


def stereo_matching(L, R, disp_type, algorithm, win=15, min_d=0, num_d=64):
    return _method("ref_stereo_matching", L, R, (int(disp_type), int(algorithm), int(win), int(min_d), int(num_d)))


def asw_traditional(L, R, gamma_c=30.0, gamma_g=20.0, disp_type=0, win=35, min_d=0, num_d=16):
    return _method("ref_adaptive_weight", L, R, (C.c_double(gamma_c), C.c_double(gamma_g), int(disp_type), int(win),
                                                  int(min_d), int(num_d)))


def asw_direct8(L, R, disp_type=0, win=35, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_direct8", L, R, (int(disp_type), int(win), int(min_d), int(num_d)))


def asw_geodesic(L, R, disp_type=0, win=35, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_geodesic", L, R, (int(disp_type), int(win), int(min_d), int(num_d)))


def asw_bilateral_grid(L, R, disp_type=0, rate_s=10.0, rate_r=10.0, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_bilateral_grid", L, R, (int(disp_type), C.c_double(rate_s), C.c_double(rate_r),
                                                                 int(min_d), int(num_d)))


def asw_blo1(L, R, disp_type=0, rate_r=0.015, win=35, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_blo1", L, R, (int(disp_type), C.c_double(rate_r), int(win), int(min_d), int(num_d)))


def asw_guidedf(L, R, disp_type=0, eps=1e-6, win=9, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_guidedf", L, R, (int(disp_type), C.c_double(eps), int(win), int(min_d), int(num_d)))


def asw_guidedf2(L, R, disp_type=0, eps=1e-6, win=9, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_guidedf_2", L, R, (int(disp_type), C.c_double(eps), int(win), int(min_d), int(num_d)))


def asw_guidedf3(L, R, disp_type=0, eps=1e-6, win=9, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_guidedf_3", L, R, (int(disp_type), C.c_double(eps), int(win), int(min_d), int(num_d)))


def asw_ncc(L, R, disp_type=0, win=9, min_d=0, num_d=16):
    return _method("ref_ncc", L, R, (int(disp_type), int(win), int(min_d), int(num_d)))


def cost_ncc(L, R, min_d, num_d, win, disp_type=0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    vol = np.empty((num_d, H, W), np.float32)
    ok = _chk(lib().ref_cost_ncc(pl, pr, H, W, int(min_d), int(num_d), int(disp_type), int(win), vol.ctypes.data_as(f32p)), "cost_ncc")
    return vol if ok else None


def asw_weighted_median(L, R, disp_type=0, win=9, rate_s=10.0, rate_r=10.0, min_d=0, num_d=16):
    return _method("ref_adaptive_weight_weighted_median", L, R, (int(disp_type), int(win), C.c_double(rate_s),
                                                                  C.c_double(rate_r), int(min_d), int(num_d)))


def geodesic_dist(img, win):
    a, pa = _u8(img)
    H, W = a.shape[:2]
    out = np.empty((H, W, win, win), np.float32)
    return out if _chk(lib().ref_geodesic_dist(pa, H, W, int(win), out.ctypes.data_as(f32p)), "geodesic_dist") else None


def cost_tad_cg(L, R, min_d, num_d, disp_type=0, regularity=0.4, thres_c=10.0, thres_g=50.0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    vol = np.empty((num_d, H, W), np.float32)
    ok = _chk(lib().ref_cost_tad_cg(pl, pr, H, W, int(min_d), int(num_d), int(disp_type), C.c_double(regularity),
                                    C.c_double(thres_c), C.c_double(thres_g), vol.ctypes.data_as(f32p)), "cost_tad_cg")
    return vol if ok else None


def cost_tad_cg_padded(L, R, min_d, num_d, win, disp_type=0, regularity=0.4, thres_c=10.0, thres_g=50.0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    h = win // 2
    vol = np.empty((num_d, H + 2 * h, W + 2 * h), np.float32)
    ok = _chk(lib().ref_cost_tad_cg_padded(pl, pr, H, W, int(min_d), int(num_d), int(disp_type), C.c_double(regularity),
                                           C.c_double(thres_c), C.c_double(thres_g), int(win), vol.ctypes.data_as(f32p)),
              "cost_tad_cg_padded")
    return vol if ok else None


def cost_sad_box(L, R, min_d, num_d, win, disp_type=0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    vol = np.empty((num_d, H, W), np.float32)
    ok = _chk(lib().ref_cost_sad_box(pl, pr, H, W, int(min_d), int(num_d), int(disp_type), int(win),
                                     vol.ctypes.data_as(f32p)), "cost_sad_box")
    return vol if ok else None


def guided_filter(guide, p, r, eps):
    g, pg = _u8(guide)
    p, pp = _f32(p)
    H, W = p.shape
    cn = 1 if g.ndim == 2 else g.shape[2]
    out = np.empty((H, W), np.float32)
    ok = _chk(lib().ref_guided_filter(pg, cn, pp, H, W, int(r), C.c_double(eps), out.ctypes.data_as(f32p)), "guided_filter")
    return out if ok else None


# ---- the OpenCV stand-in's primitives (pinned against cv2 in tests/test_cpu_ref.py) ----
def shim_bgr2gray(img):
    a, pa = _u8(img)
    out = np.empty(a.shape[:2], np.uint8)
    _chk(lib().shim_bgr2gray(pa, a.shape[0], a.shape[1], out.ctypes.data_as(u8p)), "shim_bgr2gray")
    return out


def _f32_unary(name, src, *args):
    a, pa = _f32(src)
    out = np.empty_like(a)
    _chk(getattr(lib(), name)(pa, a.shape[0], a.shape[1], *args, out.ctypes.data_as(f32p)), name)
    return out


def shim_box_filter(src, k):
    return _f32_unary("shim_box_filter_f32", src, int(k))


def shim_normalize_f32(src):
    return _f32_unary("shim_normalize_f32", src)


def shim_exp(src):
    return _f32_unary("shim_exp_f32", src)


def shim_normalize_u8c3(img):
    a, pa = _u8(img)
    out = np.empty(a.shape, np.float32)
    _chk(lib().shim_normalize_u8c3(pa, a.shape[0], a.shape[1], out.ctypes.data_as(f32p)), "shim_normalize_u8c3")
    return out


def shim_scharr_x(img):
    a, pa = _u8(img)
    out = np.empty(a.shape, np.float32)
    _chk(lib().shim_scharr_x_u8c3(pa, a.shape[0], a.shape[1], out.ctypes.data_as(f32p)), "shim_scharr_x_u8c3")
    return out


def shim_mean3_u8(a, b, c):
    a, pa = _u8(a); b, pb = _u8(b); c, pc = _u8(c)
    out = np.empty_like(a)
    _chk(lib().shim_mean3_u8(pa, pb, pc, a.shape[0], a.shape[1], out.ctypes.data_as(u8p)), "shim_mean3_u8")
    return out


def shim_mean3_f32(a, b, c):
    a, pa = _f32(a); b, pb = _f32(b); c, pc = _f32(c)
    out = np.empty_like(a)
    _chk(lib().shim_mean3_f32(pa, pb, pc, a.shape[0], a.shape[1], out.ctypes.data_as(f32p)), "shim_mean3_f32")
    return out


def shim_blend_f32(a, alpha, b, beta):
    a, pa = _f32(a); b, pb = _f32(b)
    out = np.empty_like(a)
    _chk(lib().shim_blend_f32(pa, C.c_double(alpha), pb, C.c_double(beta), a.shape[0], a.shape[1], out.ctypes.data_as(f32p)),
         "shim_blend_f32")
    return out


def shim_trunc_u8(color, T):
    a, pa = _u8(color)
    out = np.empty_like(a)
    _chk(lib().shim_trunc_u8(pa, C.c_double(T), a.shape[0], a.shape[1], out.ctypes.data_as(u8p)), "shim_trunc_u8")
    return out
