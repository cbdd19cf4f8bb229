"""ctypes host binding over the C ABI (include/asw/asw.h -> libasw_b200.so).

Mirrors the reference's operator interface (aswStereoMatch/methods/aswMethods.h): the same
function names, argument order, defaults and error behaviour, with numpy arrays in place of
cv::Mat (HxWx3 uint8 BGR in, HxW float32 disparity out; invalid arguments give an empty
array where the reference returns an empty Mat()).  There is NO CPU fallback: importing
works without a GPU (so the symbol table can be checked), but creating a context raises
if the CUDA library or a device is missing.
"""
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ASW_B200_LIB") or os.path.join(_HERE, "libasw_b200.so")   # the override is for A/B builds

# P.h:4-24
DISPARITY_LEFT, DISPARITY_RIGHT = 0, 1
(BM, SGBM, ADAPTIVE_WEIGHT, ADAPTIVE_WEIGHT_8DIRECT, ADAPTIVE_WEIGHT_GEODESIC, ADAPTIVE_WEIGHT_BILATERAL_GRID,
 ADAPTIVE_WEIGHT_BLO1, ADAPTIVE_WEIGHT_GUIDED_FILTER, ADAPTIVE_WEIGHT_GUIDED_FILTER_2,
 ADAPTIVE_WEIGHT_GUIDED_FILTER_3, ADAPTIVE_WEIGHT_MEDIAN, NCC) = range(12)

ASW_OK, ASW_ERR_BAD_ARG, ASW_ERR_SIZE_MISMATCH, ASW_ERR_CUDA, ASW_ERR_UNSUPPORTED, ASW_ERR_NOMEM = range(6)
_STATUS = {0: "ASW_OK", 1: "ASW_ERR_BAD_ARG", 2: "ASW_ERR_SIZE_MISMATCH", 3: "ASW_ERR_CUDA",
           4: "ASW_ERR_UNSUPPORTED", 5: "ASW_ERR_NOMEM"}


class AswError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{_STATUS.get(status, status)}: {msg}")
        self.status = status


class U8Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("channels", C.c_int),
                ("step", C.c_size_t)]


class F32Image(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("step", C.c_size_t)]


class MaskImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("step", C.c_size_t)]


# every symbol include/asw/asw.h declares (tests check the library exports all of them)
EXPORTS = [
    "asw_device_count", "asw_create", "asw_destroy", "asw_last_error", "asw_version", "asw_sync", "asw_stream",
    "asw_set_tuning", "asw_host_alloc", "asw_host_free", "asw_stereo_matching", "asw_method_candidates", "asw_adaptive_weight",
    "asw_adaptive_weight_direct8", "asw_adaptive_weight_geodesic", "asw_adaptive_weight_bilateral_grid", "asw_adaptive_weight_blo1",
    "asw_adaptive_weight_guidedf", "asw_adaptive_weight_guidedf_2", "asw_adaptive_weight_guidedf_3", "asw_ncc", "asw_cost_ncc", "asw_adaptive_weight_weighted_median",
    "asw_capture_aggregated", "asw_cost_tad_cg", "asw_cost_tad_cg_padded", "asw_cost_sad_box", "asw_wta", "asw_guided_filter",
    "asw_geodesic_dist", "asw_lr_check", "asw_fill_invalid", "asw_wmedian_refine", "asw_guidedf2_lr_refine", "asw_disparity_to_u8", "asw_preprocess", "asw_batch_upload_raw",
    "asw_batch_create", "asw_batch_destroy", "asw_batch_set_active", "asw_batch_upload", "asw_batch_run_guidedf2_lr_refine",
    "asw_batch_run_method", "asw_batch_download", "asw_split_local_keys", "asw_keys_alloc", "asw_keys_download",
    "asw_keys_upload", "asw_keys_min_merge", "asw_keys_to_disparity", "asw_keys_flip_sign", "asw_pool_create",
    "asw_pool_destroy", "asw_pool_size", "asw_pool_last_error", "asw_pool_ctx", "asw_stereo_matching_batch",
    "asw_guidedf2_lr_refine_batch", "asw_stereo_matching_split", "asw_pool_last_allreduce_ms", "asw_timer_start", "asw_timer_stop",
    "asw_profile_enable", "asw_profile_reset", "asw_profile_count", "asw_profile_entry", "asw_launch_count",
    "asw_flush_l2",
]

_lib = None


def load_library():
    """Load libasw_b200.so.  Raises (never falls back) when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, ci, cd, cf = C.c_void_p, C.c_int, C.c_double, C.c_float
    pu8, pf32, pmask = C.POINTER(U8Image), C.POINTER(F32Image), C.POINTER(MaskImage)
    sig = {
        "asw_device_count": (ci, []),
        "asw_create": (ci, [ci, C.POINTER(vp)]),
        "asw_destroy": (None, [vp]),
        "asw_last_error": (C.c_char_p, [vp]),
        "asw_version": (C.c_char_p, []),
        "asw_sync": (ci, [vp]),
        "asw_stream": (vp, [vp]),
        "asw_set_tuning": (ci, [vp, ci, ci]),
        "asw_host_alloc": (vp, [C.c_size_t]),
        "asw_host_free": (None, [vp]),
        "asw_stereo_matching": (ci, [vp, pu8, pu8, pf32, ci, ci, ci, ci, ci]),
        "asw_method_candidates": (ci, [ci, ci]),
        "asw_adaptive_weight": (ci, [vp, pu8, pu8, pf32, cd, cd, ci, ci, ci, ci]),
        "asw_adaptive_weight_direct8": (ci, [vp, pu8, pu8, pf32, ci, ci, ci, ci]),
        "asw_adaptive_weight_geodesic": (ci, [vp, pu8, pu8, pf32, ci, ci, ci, ci]),
        "asw_adaptive_weight_bilateral_grid": (ci, [vp, pu8, pu8, pf32, ci, cd, cd, ci, ci]),
        "asw_adaptive_weight_blo1": (ci, [vp, pu8, pu8, pf32, ci, cd, ci, ci, ci]),
        "asw_adaptive_weight_guidedf": (ci, [vp, pu8, pu8, pf32, ci, cd, ci, ci, ci]),
        "asw_adaptive_weight_guidedf_2": (ci, [vp, pu8, pu8, pf32, ci, cd, ci, ci, ci]),
        "asw_adaptive_weight_guidedf_3": (ci, [vp, pu8, pu8, pf32, ci, cd, ci, ci, ci]),
        "asw_ncc": (ci, [vp, pu8, pu8, pf32, ci, ci, ci, ci]),
        "asw_cost_ncc": (ci, [vp, pu8, pu8, vp, ci, ci, ci, ci]),
        "asw_adaptive_weight_weighted_median": (ci, [vp, pu8, pu8, pf32, ci, ci, cd, cd, ci, ci]),
        "asw_capture_aggregated": (ci, [vp, vp, C.c_size_t]),
        "asw_cost_tad_cg": (ci, [vp, pu8, pu8, vp, cd, cd, cd, ci, ci, ci]),
        "asw_cost_sad_box": (ci, [vp, pu8, pu8, vp, ci, ci, ci, ci]),
        "asw_wta": (ci, [vp, vp, ci, ci, ci, ci, pf32]),
        "asw_guided_filter": (ci, [vp, pu8, pf32, ci, cd, pf32]),
        "asw_geodesic_dist": (ci, [vp, pu8, ci, vp]),
        "asw_lr_check": (ci, [vp, pf32, pf32, cf, pmask]),
        "asw_fill_invalid": (ci, [vp, pf32, pmask, pf32]),
        "asw_wmedian_refine": (ci, [vp, pu8, pf32, pmask, ci, cd, cd, pf32]),
        "asw_guidedf2_lr_refine": (ci, [vp, pu8, pu8, pf32, cd, ci, ci, ci, cf, cd, cd, pf32, pf32, pmask]),
        "asw_cost_tad_cg_padded": (ci, [vp, pu8, pu8, vp, cd, cd, cd, ci, ci, ci, ci]),
        "asw_disparity_to_u8": (ci, [vp, pf32, pmask]),
        "asw_preprocess": (ci, [vp, pu8, pu8]),
        "asw_batch_upload_raw": (ci, [vp, ci, pu8, pu8]),
        "asw_batch_create": (ci, [vp, ci, ci, ci, C.POINTER(vp)]),
        "asw_batch_destroy": (None, [vp]),
        "asw_batch_set_active": (ci, [vp, ci]),
        "asw_batch_upload": (ci, [vp, ci, pu8, pu8]),
        "asw_batch_run_guidedf2_lr_refine": (ci, [vp, cd, ci, ci, ci, cf, cd, cd]),
        "asw_batch_run_method": (ci, [vp, ci, ci, ci, ci, ci]),
        "asw_batch_download": (ci, [vp, ci, pf32]),
        "asw_split_local_keys": (ci, [vp, pu8, pu8, ci, ci, ci, ci, ci, ci, ci, C.POINTER(vp)]),
        "asw_keys_alloc": (ci, [vp, ci, ci, C.POINTER(vp)]),
        "asw_keys_download": (ci, [vp, vp, ci, ci, vp]),
        "asw_keys_upload": (ci, [vp, vp, ci, ci, vp]),
        "asw_keys_min_merge": (ci, [vp, vp, vp, ci, ci]),
        "asw_keys_to_disparity": (ci, [vp, vp, pf32]),
        "asw_keys_flip_sign": (ci, [vp, vp, ci, ci]),
        "asw_pool_create": (ci, [ci, C.POINTER(vp)]),
        "asw_pool_destroy": (None, [vp]),
        "asw_pool_size": (ci, [vp]),
        "asw_pool_last_error": (C.c_char_p, [vp]),
        "asw_pool_ctx": (vp, [vp, ci]),
        "asw_stereo_matching_batch": (ci, [vp, ci, pu8, pu8, pf32, ci, ci, ci, ci, ci]),
        "asw_guidedf2_lr_refine_batch": (ci, [vp, ci, pu8, pu8, pf32, cd, ci, ci, ci, cf, cd, cd]),
        "asw_stereo_matching_split": (ci, [vp, pu8, pu8, pf32, ci, ci, ci, ci, ci]),
        "asw_pool_last_allreduce_ms": (cf, [vp]),
        "asw_timer_start": (ci, [vp]),
        "asw_timer_stop": (ci, [vp, C.POINTER(cf)]),
        "asw_profile_enable": (ci, [vp, ci]),
        "asw_profile_reset": (ci, [vp]),
        "asw_profile_count": (ci, [vp]),
        "asw_profile_entry": (ci, [vp, ci, C.POINTER(C.c_char_p), C.POINTER(cd), C.POINTER(C.c_longlong)]),
        "asw_launch_count": (C.c_longlong, [vp]),
        "asw_flush_l2": (ci, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def method_candidates(algorithm, num_disparity):
    """candidates the dispatcher's method scans (numDisparity + 1 for traditional / geodesic / grid); needs no device"""
    n = load_library().asw_method_candidates(int(algorithm), int(num_disparity))
    if n < 0:
        raise ValueError("algorithm outside the dense-matching hot path")
    return n


def _u8(a, channels=3):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError("image must be uint8")
    if a.ndim == 2:
        a = a[:, :, None]
    if a.strides[2] != 1 or a.strides[1] != a.shape[2]:
        a = np.ascontiguousarray(a)
    return a, U8Image(a.ctypes.data, a.shape[0], a.shape[1], a.shape[2], a.strides[0])


def _f32_in(a):
    a = np.asarray(a, dtype=np.float32)
    if a.ndim != 2 or a.strides[1] != 4:
        a = np.ascontiguousarray(a)
    return a, F32Image(a.ctypes.data, a.shape[0], a.shape[1], a.strides[0])


def _f32_out(H, W):
    a = np.empty((H, W), np.float32)
    return a, F32Image(a.ctypes.data, H, W, a.strides[0])


def _mask_in(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, MaskImage(a.ctypes.data, a.shape[0], a.shape[1], a.strides[0])


def _mask_out(H, W):
    a = np.empty((H, W), np.uint8)
    return a, MaskImage(a.ctypes.data, H, W, a.strides[0])


class Context:
    """One CUDA device + stream + workspaces (asw_ctx).  Not thread-safe."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.asw_create(int(device), C.byref(h))
        if st != ASW_OK:
            raise AswError(st, "asw_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = device
        self._children = weakref.WeakSet()     # batches must be destroyed before their context

    def close(self):
        if getattr(self, "h", None):
            for b in list(self._children):
                b.close()
            self.lib.asw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st != ASW_OK:
            raise AswError(st, self.lib.asw_last_error(self.h).decode())

    # ---- per-method entry points; strict=False mirrors the reference (empty Mat on invalid input) ----
    def _method(self, fn, L, R, args, n_eval, agg, strict):
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        H, W = La.shape[:2]
        out, outs = _f32_out(H, W)
        vol = None
        if agg:
            vol = np.empty((n_eval, H, W), np.float32)
            self._chk(self.lib.asw_capture_aggregated(self.h, vol.ctypes.data, vol.size))
        st = fn(self.h, C.byref(Ls), C.byref(Rs), C.byref(outs), *args)
        if st != ASW_OK:
            self.lib.asw_capture_aggregated(self.h, None, 0)
            if strict or st in (ASW_ERR_CUDA, ASW_ERR_NOMEM):
                self._chk(st)
            empty = np.empty((0, 0), np.float32)
            return (empty, None) if agg else empty
        return (out, vol) if agg else out

    def stereoMatching(self, srcLeft, srcRight, disparityType, algorithmType, winSize=15, minDisparity=0,
                       numDisparity=64, strict=False):
        """stereoMatching (A.h:91-92); returns the disparity map the reference writes to disparityMap."""
        return self._method(self.lib.asw_stereo_matching, srcLeft, srcRight,
                            (int(disparityType), int(algorithmType), int(winSize), int(minDisparity),
                             int(numDisparity)), 0, False, strict)

    def computeAdaptiveWeight(self, leftImg, rightImg, gamma_c=30.0, gamma_g=2.0, dispType=DISPARITY_LEFT,
                              winSize=7, minDisparity=186, numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight, leftImg, rightImg,
                            (float(gamma_c), float(gamma_g), int(dispType), int(winSize), int(minDisparity),
                             int(numDisparity)), numDisparity + 1, agg, strict)

    def computeAdaptiveWeight_direct8(self, leftImg, rightImg, dispType=DISPARITY_LEFT, winSize=7, minDisparity=186,
                                      numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_direct8, leftImg, rightImg,
                            (int(dispType), int(winSize), int(minDisparity), int(numDisparity)),
                            numDisparity + 1, agg, strict)

    def computeAdaptiveWeight_geodesic(self, leftImg, rightImg, dispType=DISPARITY_LEFT, winSize=7,
                                       minDisparity=186, numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_geodesic, leftImg, rightImg,
                            (int(dispType), int(winSize), int(minDisparity), int(numDisparity)),
                            numDisparity + 1, agg, strict)

    def computeAdaptiveWeight_bilateralGrid(self, leftImg, rightImg, dispType=DISPARITY_LEFT, sampleRateS=10.0,
                                            sampleRateR=10.0, minDisparity=186, numDisparity=144, agg=False,
                                            strict=False):
        return self._method(self.lib.asw_adaptive_weight_bilateral_grid, leftImg, rightImg,
                            (int(dispType), float(sampleRateS), float(sampleRateR), int(minDisparity),
                             int(numDisparity)), numDisparity + 1, agg, strict)

    def computeAdaptiveWeight_BLO1(self, leftImg, rightImg, dispType=DISPARITY_LEFT, sampleRateR=10.0, winSize=35,
                                   minDisparity=186, numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_blo1, leftImg, rightImg,
                            (int(dispType), float(sampleRateR), int(winSize), int(minDisparity), int(numDisparity)),
                            numDisparity, agg, strict)

    def computeAdaptiveWeight_GuidedF(self, leftImg, rightImg, dispType=DISPARITY_LEFT, eps=1e-8, winSize=35,
                                      minDisparity=186, numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_guidedf, leftImg, rightImg,
                            (int(dispType), float(eps), int(winSize), int(minDisparity), int(numDisparity)),
                            numDisparity, agg, strict)

    def computeAdaptiveWeight_GuidedF_2(self, leftImg, rightImg, dispType=DISPARITY_LEFT, eps=1e-8, winSize=35,
                                        minDisparity=186, numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_guidedf_2, leftImg, rightImg,
                            (int(dispType), float(eps), int(winSize), int(minDisparity), int(numDisparity)),
                            numDisparity, agg, strict)

    def computeAdaptiveWeight_GuidedF_3(self, leftImg, rightImg, dispType=DISPARITY_LEFT, eps=1e-6, winSize=35,
                                        minDisparity=186, numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_guidedf_3, leftImg, rightImg,
                            (int(dispType), float(eps), int(winSize), int(minDisparity), int(numDisparity)),
                            numDisparity, agg, strict)

    def computeNCC(self, leftImg, rightImg, dispType=DISPARITY_LEFT, winSize=7, minDisparity=0, numDisparity=30, strict=False):
        """computeNCC, Mat overload (A.h:124-125)"""
        return self._method(self.lib.asw_ncc, leftImg, rightImg,
                            (int(dispType), int(winSize), int(minDisparity), int(numDisparity)), 0, False, strict)

    def computeNCC_volume(self, leftImg, rightImg, dispType=DISPARITY_LEFT, winSize=7, minDisparity=0, numDisparity=30):
        """computeNCC, vector overload (A.h:126-128): the normalised [D][H][W] cost volume"""
        La, Ls = _u8(leftImg)
        Ra, Rs = _u8(rightImg)
        H, W = La.shape[:2]
        vol = np.empty((numDisparity, H, W), np.float32)
        self._chk(self.lib.asw_cost_ncc(self.h, C.byref(Ls), C.byref(Rs), vol.ctypes.data, int(dispType), int(winSize),
                                        int(minDisparity), int(numDisparity)))
        return vol

    def computeAdaptiveWeight_WeightedMedian(self, leftImg, rightImg, dispType=DISPARITY_LEFT, winSize=35,
                                             sampleRateS=10.0, sampleRateR=10.0, minDisparity=186,
                                             numDisparity=144, agg=False, strict=False):
        return self._method(self.lib.asw_adaptive_weight_weighted_median, leftImg, rightImg,
                            (int(dispType), int(winSize), float(sampleRateS), float(sampleRateR),
                             int(minDisparity), int(numDisparity)), numDisparity, agg, strict)

    # ---- stage level ----
    def computeSimilarity_padded(self, leftImg, rightImg, regularity, thresC, thresG, dispType, winSize, minDisparity,
                                 numDisparity):
        """computeSimilarity 8-arg (A.h:115-117): the [D][H + 2h][W + 2h] float volume, every slice REFLECT-padded by winSize / 2."""
        La, Ls = _u8(leftImg)
        Ra, Rs = _u8(rightImg)
        H, W = La.shape[:2]
        h = int(winSize) // 2
        vol = np.empty((numDisparity, H + 2 * h, W + 2 * h), np.float32)
        self._chk(self.lib.asw_cost_tad_cg_padded(self.h, C.byref(Ls), C.byref(Rs), vol.ctypes.data, float(regularity),
                                                  float(thresC), float(thresG), int(dispType), int(winSize),
                                                  int(minDisparity), int(numDisparity)))
        return vol

    def computeSimilarity(self, leftImg, rightImg, regularity, thresC, thresG, dispType, minDisparity,
                          numDisparity):
        """computeSimilarity 7-arg (A.h:112-114): returns the [D][H][W] float volume (cost_d_imgs)."""
        La, Ls = _u8(leftImg)
        Ra, Rs = _u8(rightImg)
        H, W = La.shape[:2]
        vol = np.empty((numDisparity, H, W), np.float32)
        self._chk(self.lib.asw_cost_tad_cg(self.h, C.byref(Ls), C.byref(Rs), vol.ctypes.data, float(regularity),
                                           float(thresC), float(thresG), int(dispType), int(minDisparity),
                                           int(numDisparity)))
        return vol

    def getCostSAD(self, leftImg, rightImg, dispType, winSize, minDisparity, numDisparity):
        """getCostSAD_d for every d (A.h:157 as called at A.cpp:2524-2536)."""
        La, Ls = _u8(leftImg)
        Ra, Rs = _u8(rightImg)
        H, W = La.shape[:2]
        vol = np.empty((numDisparity, H, W), np.float32)
        self._chk(self.lib.asw_cost_sad_box(self.h, C.byref(Ls), C.byref(Rs), vol.ctypes.data, int(dispType),
                                            int(winSize), int(minDisparity), int(numDisparity)))
        return vol

    def wta(self, volume, minDisparity=0):
        v = np.ascontiguousarray(volume, dtype=np.float32)
        D, H, W = v.shape
        out, outs = _f32_out(H, W)
        self._chk(self.lib.asw_wta(self.h, v.ctypes.data, D, H, W, int(minDisparity), C.byref(outs)))
        return out

    def getGuidedFilter(self, guidedImg, inputP, r, eps):
        ga, gs = _u8(guidedImg)
        pa, ps = _f32_in(inputP)
        out, outs = _f32_out(pa.shape[0], pa.shape[1])
        self._chk(self.lib.asw_guided_filter(self.h, C.byref(gs), C.byref(ps), int(r), float(eps), C.byref(outs)))
        return out

    def getGeodesicDist(self, img, winSize):
        ia, is_ = _u8(img)
        H, W = ia.shape[:2]
        out = np.empty((H, W, winSize, winSize), np.float32)
        self._chk(self.lib.asw_geodesic_dist(self.h, C.byref(is_), int(winSize), out.ctypes.data))
        return out

    def lr_check(self, dl, dr, tol=0.0):
        a, as_ = _f32_in(dl)
        b, bs = _f32_in(dr)
        m, ms = _mask_out(a.shape[0], a.shape[1])
        self._chk(self.lib.asw_lr_check(self.h, C.byref(as_), C.byref(bs), float(tol), C.byref(ms)))
        return m

    def fill_invalid(self, d, valid):
        a, as_ = _f32_in(d)
        m, ms = _mask_in(valid)
        out, outs = _f32_out(a.shape[0], a.shape[1])
        self._chk(self.lib.asw_fill_invalid(self.h, C.byref(as_), C.byref(ms), C.byref(outs)))
        return out

    def wmedian_refine(self, img, filled, valid, win=9, rate_s=10.0, rate_r=10.0):
        ia, is_ = _u8(img)
        f, fs = _f32_in(filled)
        m, ms = _mask_in(valid)
        out, outs = _f32_out(f.shape[0], f.shape[1])
        self._chk(self.lib.asw_wmedian_refine(self.h, C.byref(is_), C.byref(fs), C.byref(ms), int(win),
                                              float(rate_s), float(rate_r), C.byref(outs)))
        return out

    def guidedf2_lr_refine(self, L, R, eps=1e-4, win=9, min_d=0, num_d=64, tol=0.0, rate_s=10.0, rate_r=10.0,
                           parts=False):
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        H, W = La.shape[:2]
        out, outs = _f32_out(H, W)
        if parts:
            dl, dls = _f32_out(H, W)
            dr, drs = _f32_out(H, W)
            m, ms = _mask_out(H, W)
            self._chk(self.lib.asw_guidedf2_lr_refine(self.h, C.byref(Ls), C.byref(Rs), C.byref(outs), float(eps),
                                                      int(win), int(min_d), int(num_d), float(tol), float(rate_s),
                                                      float(rate_r), C.byref(dls), C.byref(drs), C.byref(ms)))
            return out, dict(dl=dl, dr=dr, valid=m)
        self._chk(self.lib.asw_guidedf2_lr_refine(self.h, C.byref(Ls), C.byref(Rs), C.byref(outs), float(eps),
                                                  int(win), int(min_d), int(num_d), float(tol), float(rate_s),
                                                  float(rate_r), None, None, None))
        return out

    def disparity_to_u8(self, disp):
        """the driver's 8-bit output (aswStereoMatch.cpp:97-98): convertTo(CV_8UC1) + normalize(0, 255, NORM_MINMAX)"""
        a, as_ = _f32_in(disp)
        m, ms = _mask_out(a.shape[0], a.shape[1])
        self._chk(self.lib.asw_disparity_to_u8(self.h, C.byref(as_), C.byref(ms)))
        return m

    def preprocess(self, img, width=640, height=360):
        """the driver's per-image pre-processing (aswStereoMatch.cpp:30-31, 67-89): resize to width x height + V-channel
        bilateral detail boost, CV_8UC3 -> CV_8UC3"""
        a, as_ = _u8(img)
        out = np.empty((int(height), int(width), 3), np.uint8)
        _, os_ = _u8(out)
        self._chk(self.lib.asw_preprocess(self.h, C.byref(as_), C.byref(os_)))
        return out

    # ---- disparity split ----
    def split_local_keys(self, L, R, algorithm, disp_type, win, min_d, num_d, d_begin, d_end):
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        H, W = La.shape[:2]
        dk = C.c_void_p()
        self._chk(self.lib.asw_split_local_keys(self.h, C.byref(Ls), C.byref(Rs), int(algorithm), int(disp_type),
                                                int(win), int(min_d), int(num_d), int(d_begin), int(d_end),
                                                C.byref(dk)))
        keys = np.empty((H, W), np.uint64)
        self._chk(self.lib.asw_keys_download(self.h, dk, H, W, keys.ctypes.data))
        return keys, dk

    def split_local_keys_device(self, L, R, algorithm, disp_type, win, min_d, num_d, d_begin, d_end):
        """as split_local_keys, but the keys stay in device memory: returns the device pointer (valid until the next call
        that touches the ctx's key workspace)"""
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        dk = C.c_void_p()
        self._chk(self.lib.asw_split_local_keys(self.h, C.byref(Ls), C.byref(Rs), int(algorithm), int(disp_type),
                                                int(win), int(min_d), int(num_d), int(d_begin), int(d_end),
                                                C.byref(dk)))
        return dk, La.shape[:2]

    def keys_flip_sign(self, dk, H, W):
        self._chk(self.lib.asw_keys_flip_sign(self.h, dk, int(H), int(W)))

    def device_keys_to_disparity(self, dk, H, W):
        out, outs = _f32_out(H, W)
        self._chk(self.lib.asw_keys_to_disparity(self.h, dk, C.byref(outs)))
        return out

    def keys_to_disparity(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        H, W = keys.shape
        dk = C.c_void_p()
        self._chk(self.lib.asw_keys_alloc(self.h, H, W, C.byref(dk)))
        self._chk(self.lib.asw_keys_upload(self.h, keys.ctypes.data, H, W, dk))
        out, outs = _f32_out(H, W)
        self._chk(self.lib.asw_keys_to_disparity(self.h, dk, C.byref(outs)))
        return out

    def keys_device_merge(self, keys_a_host, dk_other):
        """upload keys_a into a fresh device key buffer, MIN-merge the device buffer dk_other into it on the GPU
        (asw_keys_min_merge, the peer-buffer variant of the exchange) and return the resulting disparity map"""
        keys_a_host = np.ascontiguousarray(keys_a_host, dtype=np.uint64)
        H, W = keys_a_host.shape
        dk = C.c_void_p()
        self._chk(self.lib.asw_keys_alloc(self.h, H, W, C.byref(dk)))
        self._chk(self.lib.asw_keys_upload(self.h, keys_a_host.ctypes.data, H, W, dk))
        self._chk(self.lib.asw_keys_min_merge(self.h, dk, dk_other, H, W))
        out, outs = _f32_out(H, W)
        self._chk(self.lib.asw_keys_to_disparity(self.h, dk, C.byref(outs)))
        return out

    # ---- tuning knobs (asw_set_tuning): 0 restores the built-in choice ----
    TUNE_GFS_BANDS, TUNE_BLO_DCH, TUNE_GF_CHUNK_MB = 0, 1, 2

    def set_tuning(self, key, value):
        self._chk(self.lib.asw_set_tuning(self.h, int(key), int(value)))

    def has_dev_kernels(self):
        """True when the library was built with -DASW_DEV_KERNELS (superseded kernels selectable through ASW_* variables)"""
        return b"+dev-kernels" in self.lib.asw_version()

    # ---- measurement ----
    def sync(self):
        self._chk(self.lib.asw_sync(self.h))

    def timer_start(self):
        self._chk(self.lib.asw_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._chk(self.lib.asw_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        self._chk(self.lib.asw_profile_enable(self.h, 1 if on else 0))

    def profile_reset(self):
        self._chk(self.lib.asw_profile_reset(self.h))

    def profile(self):
        out = {}
        for i in range(self.lib.asw_profile_count(self.h)):
            name = C.c_char_p()
            ms = C.c_double()
            n = C.c_longlong()
            self._chk(self.lib.asw_profile_entry(self.h, i, C.byref(name), C.byref(ms), C.byref(n)))
            out[name.value.decode()] = (ms.value, n.value)
        return out

    def launch_count(self):
        return int(self.lib.asw_launch_count(self.h))

    def flush_l2(self):
        self._chk(self.lib.asw_flush_l2(self.h))


class Batch:
    """Device-resident batch of stereo pairs (config 5)."""

    def __init__(self, ctx, n_pairs, H, W):
        self.ctx = ctx
        self.n, self.H, self.W = n_pairs, H, W
        h = C.c_void_p()
        ctx._chk(ctx.lib.asw_batch_create(ctx.h, n_pairs, H, W, C.byref(h)))
        self.h = h
        ctx._children.add(self)

    def close(self):
        # finalisers of a garbage cycle run in arbitrary order: never touch a batch whose context is gone
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):
                self.ctx.lib.asw_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, i, L, R):
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        self.ctx._chk(self.ctx.lib.asw_batch_upload(self.h, i, C.byref(Ls), C.byref(Rs)))

    def upload_raw(self, i, L, R):
        """raw frames of any size: uploaded once, pre-processed on the device (Context.preprocess) into slot i"""
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        self.ctx._chk(self.ctx.lib.asw_batch_upload_raw(self.h, i, C.byref(Ls), C.byref(Rs)))

    def set_active(self, n_pairs=None):
        """the run calls process pairs [0, n_pairs) (None: the whole batch)"""
        self.ctx._chk(self.ctx.lib.asw_batch_set_active(self.h, self.n if n_pairs is None else int(n_pairs)))

    def run_guidedf2_lr_refine(self, eps=1e-4, win=9, min_d=0, num_d=64, tol=0.0, rate_s=10.0, rate_r=10.0,
                               n_pairs=None):
        self.set_active(n_pairs)
        self.ctx._chk(self.ctx.lib.asw_batch_run_guidedf2_lr_refine(self.h, float(eps), int(win), int(min_d),
                                                                    int(num_d), float(tol), float(rate_s),
                                                                    float(rate_r)))

    def run_method(self, algorithm, disp_type=DISPARITY_LEFT, win=15, min_d=0, num_d=64, n_pairs=None):
        self.set_active(n_pairs)
        self.ctx._chk(self.ctx.lib.asw_batch_run_method(self.h, int(algorithm), int(disp_type), int(win),
                                                        int(min_d), int(num_d)))

    def download(self, i, out=None, sync=True):
        """D2H of pair i's map (asynchronous on the ctx stream unless sync)."""
        if out is None:
            out = np.empty((self.H, self.W), np.float32)
        s = F32Image(out.ctypes.data, self.H, self.W, out.strides[0])
        self.ctx._chk(self.ctx.lib.asw_batch_download(self.h, i, C.byref(s)))
        if sync:
            self.ctx.sync()
        return out


class Pool:
    """Every device of the box behind one handle (asw_pool): pair-sharded batches and the disparity-range split of one
    pair with an NCCL MIN all-reduce of the device-resident keys.  Single process, one host thread per device inside
    the library (the torchrun path of bench.py uses one Context per process instead)."""

    def __init__(self, n_devices=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.asw_pool_create(int(n_devices), C.byref(h))
        if st != ASW_OK:
            raise AswError(st, "asw_pool_create failed: not enough usable CUDA devices (there is no CPU fallback)")
        self.h = h
        self.size = int(self.lib.asw_pool_size(h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.asw_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st != ASW_OK:
            raise AswError(st, self.lib.asw_pool_last_error(self.h).decode())

    @staticmethod
    def _arrays(Ls, Rs):
        n = len(Ls)
        keep, la, ra = [], (U8Image * n)(), (U8Image * n)()
        for i in range(n):
            a, la[i] = _u8(Ls[i])
            b, ra[i] = _u8(Rs[i])
            keep += [a, b]
        H, W = keep[0].shape[:2]
        outs = [np.empty((H, W), np.float32) for _ in range(n)]
        da = (F32Image * n)()
        for i in range(n):
            da[i] = F32Image(outs[i].ctypes.data, H, W, outs[i].strides[0])
        return n, keep, la, ra, outs, da

    def stereoMatchingBatch(self, Ls, Rs, disparityType, algorithmType, winSize=15, minDisparity=0, numDisparity=64):
        """stereoMatching (A.h:91-92) for a list of pairs, pair i on device i % n"""
        n, keep, la, ra, outs, da = self._arrays(Ls, Rs)
        self._chk(self.lib.asw_stereo_matching_batch(self.h, n, la, ra, da, int(disparityType), int(algorithmType),
                                                     int(winSize), int(minDisparity), int(numDisparity)))
        return outs

    def guidedf2_lr_refine_batch(self, Ls, Rs, eps=1e-4, win=9, min_d=0, num_d=64, tol=0.0, rate_s=10.0, rate_r=10.0):
        n, keep, la, ra, outs, da = self._arrays(Ls, Rs)
        self._chk(self.lib.asw_guidedf2_lr_refine_batch(self.h, n, la, ra, da, float(eps), int(win), int(min_d),
                                                        int(num_d), float(tol), float(rate_s), float(rate_r)))
        return outs

    def stereoMatchingSplit(self, L, R, disparityType, algorithmType, winSize=15, minDisparity=0, numDisparity=64):
        """one pair, its candidate range split over the devices, NCCL MIN all-reduce of the keys"""
        La, Ls = _u8(L)
        Ra, Rs = _u8(R)
        out, outs = _f32_out(*La.shape[:2])
        self._chk(self.lib.asw_stereo_matching_split(self.h, C.byref(Ls), C.byref(Rs), C.byref(outs), int(disparityType),
                                                     int(algorithmType), int(winSize), int(minDisparity),
                                                     int(numDisparity)))
        return out

    def last_allreduce_ms(self):
        return float(self.lib.asw_pool_last_allreduce_ms(self.h))


def pinned_empty(shape, dtype):
    """numpy array over cudaHostAlloc'ed (pinned) memory; the allocation lives as long as any view of it
    (the ctypes block is the views' base) and is released with asw_host_free when the last one dies."""
    lib = load_library()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = lib.asw_host_alloc(max(nbytes, 1))
    if not p:
        raise MemoryError("asw_host_alloc failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
    weakref.finalize(buf, lib.asw_host_free, p)     # the ctypes block is every view's base: freed with the last view
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr
