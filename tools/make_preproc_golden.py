"""Writes tests/golden/preproc_cv2.npz: outputs of the REAL cv2 (4.13, this container) for the reference driver's
pre-processing calls (aswStereoMatch.cpp:30-31, 67-89) on small synthetic frames.  Test infrastructure: the GPU box has
no /root/reference, and its cv2 -- if any -- may dispatch to another vector width (see oracle/preproc.py on HSV2BGR)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cv2.setNumThreads(1)


def frame(h, w, seed):
    """textured colour frame: smooth blobs + noise, so that hue sectors, gray pixels and saturated values all occur"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.int32)
    img += rng.integers(-12, 13, img.shape)
    img = np.clip(img, 0, 255).astype(np.uint8)
    img[:3] = img[:3, :, :1]                      # gray rows (S = 0)
    img[3, :8] = 255
    img[4, :8] = 0
    return img


def cv2_preprocess(img, dw, dh):
    small = cv2.resize(img, (dw, dh))
    hsv = cv2.cvtColor(small, cv2.COLOR_BGR2HSV)
    h, s, v = cv2.split(hsv)
    blur = cv2.bilateralFilter(v, 7, 10, 3, borderType=cv2.BORDER_REFLECT)
    detail = cv2.subtract(v, blur)
    v2 = cv2.addWeighted(v, 1, detail, 2, 0)       # V + detail * 2 (cv::MatExpr lowers A + B*s to one scaled add)
    out = cv2.cvtColor(cv2.merge([h, s, v2]), cv2.COLOR_HSV2BGR)
    return small, hsv, blur, v2, out


if __name__ == "__main__":
    arrays = {}
    # (name, raw H, raw W, target W, target H): generic downscale, exact 2x (INTER_AREA path), upscale, the driver's 640 columns
    for name, h, w, dw, dh in [("a", 150, 233, 128, 72), ("b", 144, 256, 128, 72), ("c", 50, 70, 128, 72), ("d", 27, 900, 640, 20)]:
        img = frame(h, w, len(name) * 17 + h)
        small, hsv, blur, v2, out = cv2_preprocess(img, dw, dh)
        arrays.update({f"{name}_raw": img, f"{name}_resized": small, f"{name}_hsv": hsv, f"{name}_blur": blur,
                       f"{name}_v2": v2, f"{name}_out": out})
    out_path = os.path.join(ROOT, "tests", "golden", "preproc_cv2.npz")
    np.savez_compressed(out_path, **arrays)
    print("wrote", out_path, os.path.getsize(out_path), "bytes; cv2", cv2.__version__)
