"""Deterministic synthetic stereo pairs (numpy only; shared by tests and bench.py).

Textured everywhere (no zero-variance windows), seeded, piecewise-constant
ground-truth disparity: background D//4, centred half-size rectangle D//2
(BASELINE.md section 3).  The same bytes go to the oracle and to the GPU.
"""
import numpy as np


def _binomial_blur(a):
    """5-tap binomial [1 4 6 4 1]/16, separable, edge-replicated; float32 in/out."""
    k = np.array([1, 4, 6, 4, 1], np.float32) / 16.0
    for axis in (0, 1):
        pad = [(0, 0)] * a.ndim
        pad[axis] = (2, 2)
        p = np.pad(a, pad, mode="edge")
        out = np.zeros_like(a)
        n = a.shape[axis]
        for i in range(5):
            sl = [slice(None)] * a.ndim
            sl[axis] = slice(i, i + n)
            out += k[i] * p[tuple(sl)]
        a = out
    return a


def make_pair(H, W, D, seed=0, noise=2):
    """Returns (L, R, gt): L,R uint8 HxWx3 (BGR), gt int32 HxW (left-view disparity)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H, W + D, 3), dtype=np.uint8).astype(np.float32)
    base = _binomial_blur(_binomial_blur(base))
    base = np.clip(127.5 + 3.0 * (base - 127.5), 0, 255).round().astype(np.uint8)
    L = base[:, :W].copy()
    gt = np.full((H, W), D // 4, np.int32)
    y0, y1, x0, x1 = H // 4, H // 4 + H // 2, W // 4, W // 4 + W // 2
    gt[y0:y1, x0:x1] = D // 2
    R = base[:, D:D + W].copy()                       # hole filler
    db = D // 4                                       # far plane first
    if db < W:
        R[:, :W - db] = L[:, db:]
    df = D // 2                                       # near plane overwrites
    lo = max(x0 - df, 0)
    R[y0:y1, lo:x1 - df] = L[y0:y1, lo + df:x1]
    if noise:
        n = rng.integers(-noise, noise + 1, R.shape)
        R = np.clip(R.astype(np.int32) + n, 0, 255).astype(np.uint8)
    return L, R, gt


def make_batch(n_pairs, H, W, D, seed0=1000, distinct=None):
    """Batch of pairs, seeds seed0+i.  With distinct=k only k pairs are generated from
    scratch; the rest are row-rolled copies (still distinct bytes, far cheaper to make)."""
    distinct = n_pairs if distinct is None else max(1, min(distinct, n_pairs))
    base = [make_pair(H, W, D, seed0 + i) for i in range(distinct)]
    Ls, Rs = [], []
    for i in range(n_pairs):
        L, R, _ = base[i % distinct]
        s = (i // distinct) * 7
        Ls.append(np.ascontiguousarray(np.roll(L, s, axis=0)))
        Rs.append(np.ascontiguousarray(np.roll(R, s, axis=0)))
    return Ls, Rs
