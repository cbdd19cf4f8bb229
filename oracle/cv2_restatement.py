"""Independent restatement of the OpenCV-heavy reference stages over Python cv2 4.13.

TEST INFRASTRUCTURE ONLY.  Follows aswStereoMatch/methods/aswMethods.cpp ("A.cpp")
call by call, with every cv::MatExpr lowered to the cv2 primitive OpenCV itself
would call.  Used (a) to pin oracle/asw_oracle.c bit-exactly
on these stages and (b) to write the golden fixtures under tests/golden/
(`python oracle/cv2_restatement.py --write-golden`).  It needs cv2, which exists
in the build container; nothing on the GPU box imports this file.
"""
import os
import sys

import cv2
import numpy as np

cv2.setNumThreads(1)   # single stripe: boxFilter column sums are then deterministic (see asw_oracle.c)
REFLECT = cv2.BORDER_REFLECT


def cost_tad_cg(L, R, min_d, num_d, regularity=0.4, thres_c=10.0, thres_g=50.0):
    """computeSimilarity 7-arg, 3-channel DISPARITY_LEFT branch (A.cpp:437-487)."""
    H, W = L.shape[:2]
    max_off = min_d + num_d - 1
    reg_r = 1 - regularity
    Rb = cv2.copyMakeBorder(R, 0, 0, max_off, 0, REFLECT)                      # :442
    K = np.array([[-3, 0, 3], [-10, 0, 10], [-3, 0, 3]], np.int8)
    gL = cv2.filter2D(L, cv2.CV_32F, K)                                        # :449
    gR = cv2.filter2D(Rb, cv2.CV_32F, K)                                       # :450
    out = []
    for off in range(min_d, max_off + 1):
        x0 = max_off - off
        ct = cv2.absdiff(L, Rb[:, x0:x0 + W])                                  # :455
        c0, c1, c2 = cv2.split(ct)
        color = cv2.addWeighted(cv2.add(c0, c1), 1 / 3, c2, 1 / 3, 0)          # :459 (B-1)
        cmp_c = cv2.compare(color, thres_c, cv2.CMP_GT)                        # :462
        m1 = cv2.multiply(color, cmp_c, scale=1 / 255)                         # color_.mul(cmp/255)
        cc = cv2.addWeighted(cmp_c, thres_c / 255, m1, 1, 0)                   # scaleAdd on u8 (B-2)
        cc = cc.astype(np.float32)                                             # :465
        gt = cv2.absdiff(gL, gR[:, x0:x0 + W])                                 # :470
        g0, g1, g2 = cv2.split(gt)
        G = cv2.addWeighted(cv2.add(g0, g1), 1 / 3, g2, 1 / 3, 0)              # :473
        cmp_g = cv2.compare(G, thres_g, cv2.CMP_GT)                            # :475
        bit = cv2.convertScaleAbs(cmp_g, alpha=1 / 255)                        # Mat = cmp/255 -> {0,1}
        nbit = cv2.bitwise_not(bit)                                            # {255,254}
        bit = bit.astype(np.float32)
        nbit = nbit.astype(np.float32)
        gc = cv2.scaleAdd(nbit, thres_g, cv2.multiply(G, bit))                 # :482
        out.append(cv2.addWeighted(cc, reg_r, gc, regularity, 0))              # :484
    return np.stack(out)


def cost_sad_box(L, R, min_d, num_d, win):
    """getCostSAD_d (A.cpp:2442-2503) as called at A.cpp:2524-2536 (LEFT)."""
    H, W = L.shape[:2]
    max_off = min_d + num_d - 1
    lg = cv2.cvtColor(L, cv2.COLOR_BGR2GRAY)
    Rb = cv2.copyMakeBorder(R, 0, 0, max_off, 0, REFLECT)
    rg = cv2.cvtColor(Rb, cv2.COLOR_BGR2GRAY)
    out = []
    for d in range(min_d, max_off + 1):
        x0 = rg.shape[1] - W - d
        ad = cv2.absdiff(lg, rg[:, x0:x0 + W]).astype(np.float32)
        out.append(cv2.boxFilter(ad, -1, (win, win), normalize=True))
    return np.stack(out)


def vec_dot(a, b):
    """operator*(VecNf, VecNf) (A.cpp:22-31): float, left to right."""
    acc = a[..., 0] * b[..., 0]
    for c in range(1, a.shape[-1]):
        acc = acc + a[..., c] * b[..., c]
    return acc.astype(np.float32)


def guided_filter(guide, p, r, eps):
    """getGuidedFilter (A.cpp:2766-2854)."""
    I = cv2.normalize(guide, None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F)          # :2774
    p = cv2.normalize(p, None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F)              # :2775
    if I.ndim == 2:
        I = I[..., None]
    C = I.shape[2]
    box = lambda m: cv2.boxFilter(m, cv2.CV_32F, (r, r))
    mI = np.stack([box(np.ascontiguousarray(I[..., c])) for c in range(C)], -1)    # :2778
    mP = box(p)                                                                # :2780
    cIp = np.stack([box(cv2.multiply(np.ascontiguousarray(I[..., c]), p)) for c in range(C)], -1)
    cII = np.stack([box(cv2.multiply(np.ascontiguousarray(I[..., c]), np.ascontiguousarray(I[..., c])))
                    for c in range(C)], -1)                                    # :2796
    var = cII - (mI * mI).astype(np.float32)                                   # :2799
    cov = cIp - (mI * mP[..., None]).astype(np.float32)                        # :2805-2815
    den = (np.float32(1.0) * np.float32(eps) + var).astype(np.float32)         # scaleAdd(ones, eps, var)
    a = (cov / den).astype(np.float32)                                         # :2846
    b = (mP - vec_dot(a, mI)).astype(np.float32)                               # :2847
    a = np.stack([box(np.ascontiguousarray(a[..., c])) for c in range(C)], -1) # :2849
    b = box(b)                                                                 # :2850
    return (vec_dot(a, I) + b).astype(np.float32)                              # :2852


def wta(vol, min_d=0):
    D, H, W = vol.shape
    best = np.full((H, W), np.finfo(np.float64).max)
    disp = np.zeros((H, W), np.float32)
    for d in range(D):
        c = vol[d].astype(np.float64)
        m = c < best
        best[m] = c[m]
        disp[m] = d + min_d
    return disp


def guidedf2(L, R, eps, win, min_d, num_d):
    """computeAdaptiveWeight_GuidedF_2, LEFT (A.cpp:2976-3050)."""
    cost = cost_tad_cg(L, R, min_d, num_d)
    q = np.stack([guided_filter(L, cost[d], win, eps) for d in range(num_d)])
    return wta(q, min_d), q


def guidedf(L, R, eps, win, min_d, num_d):
    """computeAdaptiveWeight_GuidedF, LEFT (A.cpp:2867-2963)."""
    H, W = L.shape[:2]
    max_off = min_d + num_d - 1
    cost = cost_sad_box(L, R, min_d, num_d, win)
    Rb = cv2.copyMakeBorder(R, 0, 0, max_off, 0, REFLECT)
    q = []
    for i in range(num_d):
        x0 = num_d - i - 1
        guide = np.concatenate([L, Rb[:, x0:x0 + W]], axis=2)
        q.append(guided_filter(guide, cost[i], win, eps))
    q = np.stack(q)
    return wta(q, min_d), q


def blo1(L, R, rate_r, win, min_d, num_d):
    """computeAdaptiveWeight_BLO1, LEFT (A.cpp:2505-2725)."""
    H, W = L.shape[:2]
    max_off = min_d + num_d - 1
    lg = cv2.cvtColor(L, cv2.COLOR_BGR2GRAY)
    rg = cv2.cvtColor(R, cv2.COLOR_BGR2GRAY)
    rb = cv2.copyMakeBorder(rg, 0, 0, max_off, 0, REFLECT)
    cost = cost_sad_box(L, R, min_d, num_d, win)
    step = int(256 * rate_r)
    levels = list(range(0, 256, step))
    if 255 not in levels:
        levels.append(255)
    JB = {}
    for k in levels:
        ml = cv2.absdiff(lg, np.full_like(lg, k)).astype(np.float32)
        planes = []
        M = None
        for i in range(num_d):
            x0 = max_off - i
            mr = cv2.absdiff(rb[:, x0:x0 + W], np.full_like(lg, k)).astype(np.float32)
            M = cv2.multiply(mr, ml)
            J = cv2.multiply(M, cost[i])
            planes.append(cv2.boxFilter(J, -1, (win, win)))
        N = cv2.boxFilter(M, -1, (win, win))
        JB[k] = [cv2.divide(pl, N) for pl in planes]
    q = np.zeros((num_d, H, W), np.float32)
    Ii = lg.astype(np.int32)
    for d in range(num_d):
        for k in levels:
            m = Ii == k
            q[d][m] = JB[k][d][m]
        lo = Ii // step * step
        hi = np.minimum(lo + step, 255)
        for k in levels:
            pass
        notlev = ~np.isin(Ii, levels)
        jlo = np.zeros((H, W), np.float32)
        jhi = np.zeros((H, W), np.float32)
        for k in levels:
            jlo[lo == k] = JB[k][d][lo == k]
            jhi[hi == k] = JB[k][d][hi == k]
        v = ((Ii - lo).astype(np.float32) * jlo).astype(np.float32) + ((hi - Ii).astype(np.float32) * jhi).astype(np.float32)
        q[d][notlev] = v.astype(np.float32)[notlev]
    return wta(q, min_d), q


def write_golden(outdir):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from aswstereomatch_b200.synth import make_pair
    os.makedirs(outdir, exist_ok=True)
    H, W, D = 40, 56, 8
    L, R, gt = make_pair(H, W, D, seed=7)
    cost = cost_tad_cg(L, R, 0, D)
    sad = cost_sad_box(L, R, 0, D, 5)
    q0 = guided_filter(L, cost[3], 5, 1e-4)
    d2, q2 = guidedf2(L, R, 1e-4, 5, 0, D)
    d1, q1 = guidedf(L, R, 1e-4, 5, 0, D)
    db, qb = blo1(L, R, 0.015, 7, 0, D)
    np.savez_compressed(
        os.path.join(outdir, "cv2_stages_40x56_d8.npz"),
        L=L, R=R, gray_L=cv2.cvtColor(L, cv2.COLOR_BGR2GRAY),
        cost_tad_cg=cost, cost_sad_box_w5=sad, gf_slice3_r5=q0,
        guidedf2_disp=d2, guidedf2_q=q2, guidedf_disp=d1, guidedf_q=q1,
        blo1_disp=db, blo1_q=qb,
        box9=cv2.boxFilter(cost[2], cv2.CV_32F, (9, 9)),
        norm_cost2=cv2.normalize(cost[2], None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F),
        norm_L=cv2.normalize(L, None, 0, 1, cv2.NORM_MINMAX, cv2.CV_32F),
        cv2_version=np.array(cv2.__version__))
    print("wrote", os.path.join(outdir, "cv2_stages_40x56_d8.npz"))


if __name__ == "__main__":
    if "--write-golden" in sys.argv:
        write_golden(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
