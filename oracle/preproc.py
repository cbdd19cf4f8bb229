"""CPU restatement (numpy) of the reference DRIVER's pre-processing -- TEST INFRASTRUCTURE ONLY.

aswStereoMatch/aswStereoMatch.cpp ("main.cpp"):
    :30-31   resize(img, img, Size(640, 360))                           (INTER_LINEAR)
    :67-89   cvtColor(BGR2HSV); split; bilateralFilter(V, blur, 7, 10, 3, BORDER_REFLECT);
             detail = V - blur; V = V + detail * 2; merge; cvtColor(HSV2BGR)

These are OpenCV library calls; what is restated here is OpenCV 4.13's 8-bit arithmetic, pinned against the real `cv2`
of the build container (tests/test_cpu_preproc.py, tools/make_preproc_golden.py):
  * resize: bit-exact (the classic 11-bit fixed-point bilinear path; an exact 2x2 downscale silently takes INTER_AREA)
  * BGR2HSV: bit-exact (12-bit fixed-point division tables, H in 0..179)
  * bilateralFilter: cv2 accumulates `sum += w * val` per tap in the tap order of its offset table; its SIMD body uses FMA,
    its scalar row tail separate multiply + add.  Restated with FMA: equal on > 99.99 % of the pixels, |diff| <= 1 elsewhere
  * V + 2 (V - blur): exact integer saturation (the MatExpr lowers to a saturating subtract and a scaled add)
  * HSV2BGR: cv2's SIMD body TRUNCATES b * 255 where its scalar row tail ROUNDS it (observed: converting one pixel alone
    gives a different byte than converting it inside a row).  cv2's own output therefore depends on the vector width of
    the build and on the pixel's position in the row; `hsv2bgr_u8(mode="trunc")` is the SIMD body (the spec of the CUDA
    path), `mode="round"` the scalar tail.  > 99.9 % of cv2's bytes equal one of the two; the rest sit one below the truncated
    value (the vector code associates v (1 - s + s h) differently: none of the tried associations / FMA contractions is exact).
"""
import math

import numpy as np

F = np.float32
HSV_SHIFT = 12
_SDIV = np.zeros(256, np.int64)
_HDIV = np.zeros(256, np.int64)
for _i in range(1, 256):
    _SDIV[_i] = int(np.rint((255 << HSV_SHIFT) / (1.0 * _i)))
    _HDIV[_i] = int(np.rint((180 << HSV_SHIFT) / (6.0 * _i)))


def _lin_coeffs(dn, sn, scale, clamp):
    """source index and the two 11-bit weights of every destination index (resize.cpp, INTER_LINEAR, 8u)"""
    ofs = np.zeros(dn, np.int64)
    a = np.zeros((dn, 2), np.int64)
    for d in range(dn):
        f = F((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = F(f - F(s))
        if clamp and s < 0:
            s, f = 0, F(0)
        if clamp and s >= sn - 1:
            s, f = sn - 1, F(0)
        ofs[d] = s
        a[d, 0] = int(np.rint(F(F(1.0) - f) * F(2048)))
        a[d, 1] = int(np.rint(F(f * F(2048))))
    return ofs, a


def resize_linear_u8(src, dw, dh):
    """cv::resize(src, dst, Size(dw, dh)) for CV_8UC3 / CV_8UC1, default INTER_LINEAR"""
    squeeze = src.ndim == 2
    if squeeze:
        src = src[..., None]
    sh, sw = src.shape[:2]
    scale_x = 1.0 / (dw / sw)
    scale_y = 1.0 / (dh / sh)
    if sw == 2 * dw and sh == 2 * dh:                      # INTER_LINEAR with an exact 2x2 downscale runs INTER_AREA
        s = src.astype(np.int64)
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
    else:
        xo, xa = _lin_coeffs(dw, sw, scale_x, True)        # columns: index and weights clamped at the borders
        yo, ya = _lin_coeffs(dh, sh, scale_y, False)       # rows: weights unclamped, the ROW INDEX is clipped
        s = src.astype(np.int64)
        x1 = np.minimum(xo + 1, sw - 1)
        rows = s[:, xo, :] * xa[:, 0][None, :, None] + s[:, x1, :] * xa[:, 1][None, :, None]
        y0 = np.clip(yo, 0, sh - 1)
        y1 = np.clip(yo + 1, 0, sh - 1)
        b0 = ya[:, 0][:, None, None]
        b1 = ya[:, 1][:, None, None]
        out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


def bgr2hsv_u8(img):
    """cvtColor(COLOR_BGR2HSV) on CV_8UC3: H in 0..179"""
    b = img[..., 0].astype(np.int64)
    g = img[..., 1].astype(np.int64)
    r = img[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(b, g), r)
    diff = v - np.minimum(np.minimum(b, g), r)
    s = (diff * _SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV[diff] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


_SECTOR = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])


def hsv2bgr_u8(hsv, mode="trunc"):
    """cvtColor(COLOR_HSV2BGR) on CV_8UC3 (H in 0..179).  mode: "trunc" = cv2's SIMD body, "round" = its scalar tail"""
    h = hsv[..., 0].astype(F) * F(6.0 / 180.0)
    s = hsv[..., 1].astype(F) * F(1.0 / 255.0)
    v = hsv[..., 2].astype(F) * F(1.0 / 255.0)
    sector = np.floor(h).astype(np.int64)
    hf = (h - sector.astype(F)).astype(F)
    sector = np.where(sector >= 6, sector - 6, sector)
    sh = (s * hf).astype(F)
    tabs = np.stack([v, (v * (F(1) - s)).astype(F), (v * (F(1) - sh)).astype(F),
                     (v * ((F(1) - s).astype(F) + sh).astype(F)).astype(F)], -1)
    idx = _SECTOR[sector]                                  # (b, g, r) table slots
    bgr = np.take_along_axis(tabs, idx, -1)
    bgr = np.where((hsv[..., 1] == 0)[..., None], v[..., None], bgr)
    out = (bgr * F(255.0)).astype(F)
    out = np.floor(out) if mode == "trunc" else np.rint(out)
    return np.clip(out, 0, 255).astype(np.uint8)


def bilateral_taps(d, sigma_color, sigma_space):
    """offset table and weights of bilateralFilter_8u: taps inside the radius circle, row-major"""
    radius = d // 2 if d > 0 else int(round(sigma_space * 1.5))
    radius = max(radius, 1)
    gcc = -0.5 / (sigma_color * sigma_color)
    gsc = -0.5 / (sigma_space * sigma_space)
    color_w = np.array([F(math.exp(i * i * gcc)) for i in range(256)], F)
    ofs, space_w = [], []
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = math.sqrt(float(i) * i + float(j) * j)
            if r > radius:
                continue
            ofs.append((i, j))
            space_w.append(F(math.exp(r * r * gsc)))
    return radius, ofs, np.array(space_w, F), color_w


def reflect_idx(p, n):
    """BORDER_REFLECT (edge pixel repeated): fedcba|abcdefgh|hgfedcb"""
    p = np.asarray(p)
    period = 2 * n
    q = np.mod(p, period)
    return np.where(q >= n, period - 1 - q, q)


def bilateral_u8(src, d=7, sigma_color=10.0, sigma_space=3.0, fma=True):
    """bilateralFilter(src, dst, d, sigma_color, sigma_space, BORDER_REFLECT) on CV_8UC1"""
    radius, ofs, space_w, color_w = bilateral_taps(d, sigma_color, sigma_space)
    H, W = src.shape
    yy = reflect_idx(np.arange(-radius, H + radius), H)
    xx = reflect_idx(np.arange(-radius, W + radius), W)
    tmp = src[yy][:, xx]
    val0 = src.astype(np.int64)
    acc = np.zeros((H, W), F)
    wsum = np.zeros((H, W), F)
    for (i, j), ws in zip(ofs, space_w):
        val = tmp[radius + i:radius + i + H, radius + j:radius + j + W].astype(np.int64)
        w = (ws * color_w[np.abs(val - val0)]).astype(F)
        if fma:
            acc = (acc.astype(np.float64) + val.astype(np.float64) * w.astype(np.float64)).astype(F)   # one rounding
        else:
            acc = (acc + (val.astype(F) * w).astype(F)).astype(F)
        wsum = (wsum + w).astype(F)
    return np.clip(np.rint((acc / wsum).astype(F)), 0, 255).astype(np.uint8)


def detail_boost_v(v, blur):
    """detail = V - blur (saturating); V = V + detail * 2 (one saturating scaled add)     main.cpp:76-78"""
    detail = np.clip(v.astype(np.int64) - blur.astype(np.int64), 0, 255)
    return np.clip(v.astype(np.int64) + 2 * detail, 0, 255).astype(np.uint8)


def preprocess(img, dw=640, dh=360, mode="trunc", parts=False):
    """the driver's per-image pre-processing: resize + V-channel bilateral detail boost"""
    small = resize_linear_u8(img, dw, dh)
    hsv = bgr2hsv_u8(small)
    blur = bilateral_u8(hsv[..., 2], 7, 10.0, 3.0)
    hsv2 = hsv.copy()
    hsv2[..., 2] = detail_boost_v(hsv[..., 2], blur)
    out = hsv2bgr_u8(hsv2, mode)
    if parts:
        return out, {"resized": small, "hsv": hsv, "blur": blur, "hsv_boosted": hsv2}
    return out
