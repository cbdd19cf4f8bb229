"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(aswstereomatch_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = [os.path.join(_HERE, f) for f in ("asw_oracle.c", "asw_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(u8p)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(f32p)


def _chk(rc, what):
    if rc != 0:
        raise ValueError(f"oracle {what} failed with code {rc}")


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def num_threads():
    return int(lib().orc_num_threads())


def bgr2gray(img):
    a, pa = _u8(img)
    out = np.empty(a.shape[:2], np.uint8)
    lib().orc_bgr2gray(pa, C.c_int(out.size), out.ctypes.data_as(u8p))
    return out


def box_filter(src, k):
    a, pa = _f32(src)
    out = np.empty_like(a)
    lib().orc_box_filter_f32(pa, a.shape[0], a.shape[1], int(k), out.ctypes.data_as(f32p))
    return out


def normalize_f32(src):
    a, pa = _f32(src)
    out = np.empty_like(a)
    lib().orc_normalize_minmax_f32(pa, C.c_long(a.size), out.ctypes.data_as(f32p))
    return out


def normalize_u8(src):
    a, pa = _u8(src)
    out = np.empty(a.shape, np.float32)
    lib().orc_normalize_minmax_u8(pa, C.c_long(a.size), out.ctypes.data_as(f32p))
    return out


def scharr_x(img):
    a, pa = _u8(img)
    out = np.empty(a.shape, np.float32)
    lib().orc_scharr_x_u8c3(pa, a.shape[0], a.shape[1], out.ctypes.data_as(f32p))
    return out


def cost_tad_cg(L, R, min_d, num_d, disp_type=0, regularity=0.4, thres_c=10.0, thres_g=50.0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    vol = np.empty((num_d, H, W), np.float32)
    _chk(lib().orc_cost_tad_cg(pl, pr, H, W, int(min_d), int(num_d), int(disp_type),
                               C.c_double(regularity), C.c_double(thres_c), C.c_double(thres_g),
                               vol.ctypes.data_as(f32p)), "cost_tad_cg")
    return vol


def cost_tad_cg_padded(L, R, min_d, num_d, win, disp_type=0, regularity=0.4, thres_c=10.0, thres_g=50.0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    h = win // 2
    vol = np.empty((num_d, H + 2 * h, W + 2 * h), np.float32)
    _chk(lib().orc_cost_tad_cg_padded(pl, pr, H, W, int(min_d), int(num_d), int(disp_type),
                                      C.c_double(regularity), C.c_double(thres_c), C.c_double(thres_g),
                                      int(win), vol.ctypes.data_as(f32p)), "cost_tad_cg_padded")
    return vol


def cost_sad_box(L, R, min_d, num_d, win, disp_type=0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    vol = np.empty((num_d, H, W), np.float32)
    _chk(lib().orc_cost_sad_box(pl, pr, H, W, int(min_d), int(num_d), int(disp_type), int(win),
                                vol.ctypes.data_as(f32p)), "cost_sad_box")
    return vol


def wta(vol, min_d=0):
    v, pv = _f32(vol)
    D, H, W = v.shape
    out = np.empty((H, W), np.float32)
    lib().orc_wta(pv, D, H, W, int(min_d), out.ctypes.data_as(f32p))
    return out


def guided_filter(guide, p, r, eps):
    g, pg = _u8(guide)
    p, pp = _f32(p)
    H, W = p.shape
    Cn = 1 if g.ndim == 2 else g.shape[2]
    out = np.empty((H, W), np.float32)
    _chk(lib().orc_guided_filter(pg, Cn, pp, H, W, int(r), C.c_double(eps), out.ctypes.data_as(f32p)),
         "guided_filter")
    return out


def _method(fn, L, R, n_eval, args, want_agg):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    disp = np.empty((H, W), np.float32)
    agg = np.empty((n_eval, H, W), np.float32) if want_agg else None
    pagg = agg.ctypes.data_as(f32p) if want_agg else f32p()
    rc = fn(pl, pr, H, W, *args, disp.ctypes.data_as(f32p), pagg)
    _chk(rc, fn.__name__)
    return (disp, agg) if want_agg else disp


def asw_traditional(L, R, gamma_c=30.0, gamma_g=20.0, disp_type=0, win=35, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_traditional, L, R, num_d + 1,
                   (C.c_double(gamma_c), C.c_double(gamma_g), int(disp_type), int(win), int(min_d), int(num_d)), agg)


def asw_direct8(L, R, disp_type=0, win=35, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_direct8, L, R, num_d + 1, (int(disp_type), int(win), int(min_d), int(num_d)), agg)


def asw_geodesic(L, R, disp_type=0, win=35, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_geodesic, L, R, num_d + 1,
                   (int(disp_type), int(win), int(min_d), int(num_d)), agg)


def geodesic_dist(img, win):
    a, pa = _u8(img)
    H, W = a.shape[:2]
    out = np.empty((H, W, win, win), np.float32)
    _chk(lib().orc_geodesic_dist(pa, H, W, int(win), out.ctypes.data_as(f32p)), "geodesic_dist")
    return out


def asw_bilateral_grid(L, R, disp_type=0, rate_s=10.0, rate_r=10.0, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_bilateral_grid, L, R, num_d + 1,
                   (int(disp_type), C.c_double(rate_s), C.c_double(rate_r), int(min_d), int(num_d)), agg)


def asw_blo1(L, R, disp_type=0, rate_r=0.015, win=35, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_blo1, L, R, num_d,
                   (int(disp_type), C.c_double(rate_r), int(win), int(min_d), int(num_d)), agg)


def asw_guidedf(L, R, disp_type=0, eps=1e-6, win=9, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_guidedf, L, R, num_d,
                   (int(disp_type), C.c_double(eps), int(win), int(min_d), int(num_d)), agg)


def asw_guidedf3(L, R, disp_type=0, eps=1e-6, win=9, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_guidedf3, L, R, num_d,
                   (int(disp_type), C.c_double(eps), int(win), int(min_d), int(num_d)), agg)


def asw_ncc(L, R, disp_type=0, win=9, min_d=0, num_d=16):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    disp = np.empty((H, W), np.float32)
    _chk(lib().orc_asw_ncc(pl, pr, H, W, int(disp_type), int(win), int(min_d), int(num_d), disp.ctypes.data_as(f32p)), "asw_ncc")
    return disp


def cost_ncc(L, R, min_d, num_d, win, disp_type=0):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    vol = np.empty((num_d, H, W), np.float32)
    _chk(lib().orc_cost_ncc(pl, pr, H, W, int(min_d), int(num_d), int(disp_type), int(win), vol.ctypes.data_as(f32p)), "cost_ncc")
    return vol


def asw_guidedf2(L, R, disp_type=0, eps=1e-6, win=9, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_guidedf2, L, R, num_d,
                   (int(disp_type), C.c_double(eps), int(win), int(min_d), int(num_d)), agg)


def asw_weighted_median(L, R, disp_type=0, win=9, rate_s=10.0, rate_r=10.0, min_d=0, num_d=16, agg=False):
    return _method(lib().orc_asw_weighted_median, L, R, num_d,
                   (int(disp_type), int(win), C.c_double(rate_s), C.c_double(rate_r), int(min_d), int(num_d)), agg)


def lr_check(dl, dr, tol=0.0):
    dl, pl = _f32(dl)
    dr, pr = _f32(dr)
    H, W = dl.shape
    out = np.empty((H, W), np.uint8)
    lib().orc_lr_check(pl, pr, H, W, C.c_float(tol), out.ctypes.data_as(u8p))
    return out


def fill_invalid(d, valid):
    d, pd = _f32(d)
    v, pv = _u8(valid)
    H, W = d.shape
    out = np.empty((H, W), np.float32)
    lib().orc_fill_invalid(pd, pv, H, W, out.ctypes.data_as(f32p))
    return out


def wmedian_refine(img, filled, valid, win=9, rate_s=10.0, rate_r=10.0):
    img, pi = _u8(img)
    f, pf = _f32(filled)
    v, pv = _u8(valid)
    H, W = f.shape
    out = np.empty((H, W), np.float32)
    _chk(lib().orc_wmedian_refine(pi, pf, pv, H, W, int(win), C.c_double(rate_s), C.c_double(rate_r),
                                  out.ctypes.data_as(f32p)), "wmedian_refine")
    return out


def disparity_to_u8(disp):
    d, pd = _f32(disp)
    out = np.empty(d.shape, np.uint8)
    lib().orc_disparity_to_u8(pd, d.shape[0], d.shape[1], out.ctypes.data_as(u8p))
    return out


def stereo_matching(L, R, disp_type, algorithm, win=15, min_d=0, num_d=64):
    L, pl = _u8(L)
    R, pr = _u8(R)
    H, W = L.shape[:2]
    disp = np.empty((H, W), np.float32)
    _chk(lib().orc_stereo_matching(pl, pr, H, W, int(disp_type), int(algorithm), int(win), int(min_d),
                                   int(num_d), disp.ctypes.data_as(f32p)), "stereo_matching")
    return disp


def guidedf2_lr_refine(L, R, eps=1e-4, win=9, min_d=0, num_d=64, tol=0.0, rate_s=10.0, rate_r=10.0):
    """cfg-2/cfg-5 pipeline: left + right view, LR check, fill, weighted-median refine (spec a-14)."""
    dl = asw_guidedf2(L, R, 0, eps, win, min_d, num_d)
    dr = asw_guidedf2(L, R, 1, eps, win, min_d, num_d)
    valid = lr_check(dl, dr, tol)
    filled = fill_invalid(dl, valid)
    out = wmedian_refine(L, filled, valid, win, rate_s, rate_r)
    return out, dict(dl=dl, dr=dr, valid=valid, filled=filled)
