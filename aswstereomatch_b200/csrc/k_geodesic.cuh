// k_geodesic.cuh -- geodesic-distance ASW (A.cpp:1321-1534).
//
// K8 k_geo_dist: per pixel, the chamfer DP of getWinGeoDist (A.cpp:1328-1390) on its (win+2)^2 window cut
//   from the BORDER_REFLECT-padded image (A.cpp:1404).  Pass schedule of the reference: iterations 0 and 1
//   are the BACKWARD raster (neighbours R, BR, B, BL), iteration 2 the FORWARD raster (L, UL, U, UR) because
//   the test is iterCount/2 (A.cpp:1341, 1365).  The second backward sweep is provably a no-op (every cell
//   already equals the min over its final predecessors), so two sweeps are run.  The backward raster only looks right
//   and down, so it leaves every window row BELOW the centre row at its initial FLT_MAX: those rows are neither swept,
//   written nor re-read (the forward sweep starts them from "infinity"): 18 instead of 35 backward rows at 35 x 35 and
//   a third less traffic through the distance volume.  Distances are exact
//   integers (L1 colour steps; < 2^24) so integer arithmetic is bit-identical to the reference's floats.
//   One thread per pixel; the window is stored plane-major dist[tap][y][x] so that both the DP's global
//   traffic and the aggregation's loads are coalesced across threads.
// K9 k_geo_aggregate: E = sum(DL*DR*cd) / sum(DL*DR) with the distances themselves as weights
//   (A.cpp:1488-1492), D+1 candidates, clamp addressing, WTA folded into keys.
#pragma once
#include "k_cost.cuh"

#define GEO_MAXW 63
#define GEO_INF 0x3FFFFFFF

// BGR -> packed B | G<<8 | R<<16, padded by `pad` on all sides with BORDER_REFLECT
__global__ void k_pack_bgrx_pad(const uint8_t* __restrict__ img, int H, int W, int pad, uint32_t* __restrict__ out) {
    int Wp = W + 2 * pad;
    int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y;
    if (xp >= Wp) return;
    int x = border_idx(xp - pad, W, 0), y = border_idx(yp - pad, H, 0);
    const uint8_t* s = img + ((size_t)y * W + x) * 3;
    out[(size_t)yp * Wp + xp] = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[2] << 16);
}

// ext: packed image padded by h+1.  dist: [win*win][H][W] float (exact integers).
__global__ void __launch_bounds__(128)
k_geo_dist(const uint32_t* __restrict__ ext, int H, int W, int win, float* __restrict__ dist) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int h = win / 2, Wp = W + 2 * (h + 1);
    const size_t n = (size_t)H * W, p = (size_t)y * W + x;
    // window cell (r, c), r,c in [0, win+1], is ext[(y + r) * Wp + (x + c)]; the centre is (h+1, h+1)
    int prev[GEO_MAXW + 2], cur[GEO_MAXW + 2];
    uint32_t pprev[GEO_MAXW + 2], pcur[GEO_MAXW + 2];
    // ---- backward sweep: rows win..1, cols win..1; neighbours (r,c+1), (r+1,c+1), (r+1,c), (r+1,c-1) ----
    // rows win .. h+2 stay "infinite" (no finite neighbour to the right or below): start at the centre row
    for (int c = 0; c <= win + 1; c++) { prev[c] = GEO_INF; pprev[c] = ext[(size_t)(y + h + 2) * Wp + x + c]; }
    for (int r = h + 1; r >= 1; r--) {
        const uint32_t* row = ext + (size_t)(y + r) * Wp + x;
        for (int c = 0; c <= win + 1; c++) pcur[c] = row[c];
        cur[win + 1] = GEO_INF; cur[0] = GEO_INF;
        for (int c = win; c >= 1; c--) {
            int v = (r == h + 1 && c == h + 1) ? 0 : GEO_INF;            // A.cpp:1416-1417
            uint32_t me = pcur[c];
            v = min(v, cur[c + 1] + (int)__vsadu4(pcur[c + 1], me));      // right
            v = min(v, prev[c + 1] + (int)__vsadu4(pprev[c + 1], me));    // bottom-right
            v = min(v, prev[c] + (int)__vsadu4(pprev[c], me));            // bottom
            v = min(v, prev[c - 1] + (int)__vsadu4(pprev[c - 1], me));    // bottom-left
            cur[c] = min(v, GEO_INF);
        }
        for (int c = 1; c <= win; c++) dist[(size_t)((r - 1) * win + (c - 1)) * n + p] = (float)cur[c];
        for (int c = 0; c <= win + 1; c++) { prev[c] = cur[c]; pprev[c] = pcur[c]; }
    }
    // ---- forward sweep: rows 1..win, cols 1..win; neighbours (r,c-1), (r-1,c-1), (r-1,c), (r-1,c+1) ----
    for (int c = 0; c <= win + 1; c++) { prev[c] = GEO_INF; pprev[c] = ext[(size_t)y * Wp + x + c]; }
    for (int r = 1; r <= win; r++) {
        const uint32_t* row = ext + (size_t)(y + r) * Wp + x;
        for (int c = 0; c <= win + 1; c++) pcur[c] = row[c];
        cur[0] = GEO_INF; cur[win + 1] = GEO_INF;
        for (int c = 1; c <= win; c++) {
            int v = r <= h + 1 ? (int)dist[(size_t)((r - 1) * win + (c - 1)) * n + p] : GEO_INF;  // value left by the backward sweep
            uint32_t me = pcur[c];
            v = min(v, cur[c - 1] + (int)__vsadu4(pcur[c - 1], me));      // left
            v = min(v, prev[c - 1] + (int)__vsadu4(pprev[c - 1], me));    // up-left
            v = min(v, prev[c] + (int)__vsadu4(pprev[c], me));            // up
            v = min(v, prev[c + 1] + (int)__vsadu4(pprev[c + 1], me));    // up-right
            cur[c] = min(v, GEO_INF);
        }
        for (int c = 1; c <= win; c++) dist[(size_t)((r - 1) * win + (c - 1)) * n + p] = (float)cur[c];
        for (int c = 0; c <= win + 1; c++) { prev[c] = cur[c]; pprev[c] = pcur[c]; }
    }
}

// Window size as a template constant: the DP row (distances + the colours of the previous and the current window
// row) lives in registers instead of local memory; the row update runs in place with one saved neighbour.
// 4 resident CTAs per SM (128 registers, 32 bytes of spills at 35 x 35): the DP is a chain of dependent min / add steps,
// the fourth CTA hides more of it than the spills cost (measured at 1280 x 720, two images: 3.83 ms; 3 CTAs / 168 registers
// 4.09, unconstrained 140 registers 4.04, 5 CTAs / 96 registers 5.20 ms)
#ifndef GEO_DIST_MINB
#define GEO_DIST_MINB 4
#endif
template <int WIN>
__global__ void __launch_bounds__(128, GEO_DIST_MINB)
k_geo_dist_t(const uint32_t* __restrict__ ext, int H, int W, float* __restrict__ dist) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    constexpr int h = WIN / 2;
    const int Wp = W + 2 * (h + 1);
    const size_t n = (size_t)H * W, p = (size_t)y * W + x;
    int d[WIN + 2];
    uint32_t pprev[WIN + 2], pcur[WIN + 2];
    // ---- backward sweep: rows WIN..1, cols WIN..1; neighbours (r,c+1), (r+1,c+1), (r+1,c), (r+1,c-1) ----
#pragma unroll
    for (int c = 0; c <= WIN + 1; c++) { d[c] = GEO_INF; pprev[c] = ext[(size_t)(y + h + 2) * Wp + x + c]; }
    // rows WIN .. h+2 stay "infinite" (no finite neighbour to the right or below): the sweep starts at the centre row
#pragma unroll 1
    for (int r = h + 1; r >= 1; r--) {
        const uint32_t* row = ext + (size_t)(y + r) * Wp + x;
#pragma unroll
        for (int c = 0; c <= WIN + 1; c++) pcur[c] = row[c];
        int old_right = d[WIN + 1];                                 // previous row's value right of the current cell
        d[WIN + 1] = GEO_INF;
#pragma unroll
        for (int c = WIN; c >= 1; c--) {
            const int old_c = d[c];
            const uint32_t me = pcur[c];
            int v = (r == h + 1 && c == h + 1) ? 0 : GEO_INF;            // A.cpp:1416-1417
            v = min(v, d[c + 1] + (int)__vsadu4(pcur[c + 1], me));        // right (already updated)
            v = min(v, old_right + (int)__vsadu4(pprev[c + 1], me));      // bottom-right
            v = min(v, old_c + (int)__vsadu4(pprev[c], me));              // bottom
            v = min(v, d[c - 1] + (int)__vsadu4(pprev[c - 1], me));       // bottom-left (not yet updated)
            d[c] = min(v, GEO_INF);
            old_right = old_c;
        }
        d[0] = GEO_INF;
#pragma unroll
        for (int c = 1; c <= WIN; c++) dist[(size_t)((r - 1) * WIN + (c - 1)) * n + p] = (float)d[c];
#pragma unroll
        for (int c = 0; c <= WIN + 1; c++) pprev[c] = pcur[c];
    }
    // ---- forward sweep: rows 1..WIN, cols 1..WIN; neighbours (r,c-1), (r-1,c-1), (r-1,c), (r-1,c+1) ----
#pragma unroll
    for (int c = 0; c <= WIN + 1; c++) { d[c] = GEO_INF; pprev[c] = ext[(size_t)y * Wp + x + c]; }
#pragma unroll 1
    for (int r = 1; r <= WIN; r++) {
        const uint32_t* row = ext + (size_t)(y + r) * Wp + x;
        float back[WIN];                                            // left by the backward sweep (rows up to the centre row)
        if (r <= h + 1) {
#pragma unroll
            for (int c = 0; c < WIN; c++) back[c] = dist[(size_t)((r - 1) * WIN + c) * n + p];
        } else {
#pragma unroll
            for (int c = 0; c < WIN; c++) back[c] = (float)GEO_INF;
        }
#pragma unroll
        for (int c = 0; c <= WIN + 1; c++) pcur[c] = row[c];
        int old_left = d[0];
        d[0] = GEO_INF;
#pragma unroll
        for (int c = 1; c <= WIN; c++) {
            const int old_c = d[c];
            const uint32_t me = pcur[c];
            int v = (int)back[c - 1];
            v = min(v, d[c - 1] + (int)__vsadu4(pcur[c - 1], me));        // left (already updated)
            v = min(v, old_left + (int)__vsadu4(pprev[c - 1], me));       // up-left
            v = min(v, old_c + (int)__vsadu4(pprev[c], me));              // up
            v = min(v, d[c + 1] + (int)__vsadu4(pprev[c + 1], me));       // up-right (not yet updated)
            d[c] = min(v, GEO_INF);
            old_left = old_c;
        }
        d[WIN + 1] = GEO_INF;
#pragma unroll
        for (int c = 1; c <= WIN; c++) dist[(size_t)((r - 1) * WIN + (c - 1)) * n + p] = (float)d[c];
#pragma unroll
        for (int c = 0; c <= WIN + 1; c++) pprev[c] = pcur[c];
    }
}

struct GeoGeom { int H, W, win, h, sign, d_first, n_cand; };

// thread = pixel, blockIdx.z = chunk of GEO_Q candidates.  dref/dtgt: [win*win][H][W]; pref/ptgt: packed BGRx [H][W]
// c_base: index of the launch's first candidate in the aggregated-cost volume (g.d_first is its disparity).
template <int GEO_Q>
__global__ void __launch_bounds__(128)
k_geo_aggregate_q(const float* __restrict__ dref, const float* __restrict__ dtgt, const uint32_t* __restrict__ pref,
                const uint32_t* __restrict__ ptgt, GeoGeom g, int c_base, unsigned long long* __restrict__ keys,
                float* __restrict__ agg) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.W) return;
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const size_t n = (size_t)H * W, p = (size_t)y * W + x;
    const int c0 = blockIdx.z * GEO_Q;
    const int nq = min(GEO_Q, g.n_cand - c0);
    int dq[GEO_Q];
    size_t pt[GEO_Q];
#pragma unroll
    for (int q = 0; q < GEO_Q; q++) {
        dq[q] = g.d_first + c0 + min(q, nq - 1);
        int xs = g.sign > 0 ? max(0, x - dq[q]) : min(x + dq[q], W - 1);
        pt[q] = (size_t)y * W + xs;                                 // weightAllRight[Point(max(0,x-offset), y)]
    }
    double num[GEO_Q], den[GEO_Q];
#pragma unroll
    for (int q = 0; q < GEO_Q; q++) { num[q] = 0; den[q] = 0; }
    const float* dl_p = dref + p;                                   // tap-major planes: + tap * n
    for (int j = 0; j < win; j++) {
        const int ny = clampi(y - h + j, 0, H - 1);
        const uint32_t* prow = pref + (size_t)ny * W;
        const uint32_t* trow = ptgt + (size_t)ny * W;
        float fn[GEO_Q], fd[GEO_Q];
#pragma unroll
        for (int q = 0; q < GEO_Q; q++) { fn[q] = 0; fd[q] = 0; }
        // 7 taps per round: their loads (two distance planes are streamed from HBM) are issued before the first use
        for (int i0 = 0; i0 < win; i0 += 7) {
            float dl[7], dr[7][GEO_Q];
            uint32_t cr[7], ct[7][GEO_Q];
#pragma unroll
            for (int u = 0; u < 7; u++) {
                const int i = min(i0 + u, win - 1);
                const int nx = clampi(x - h + i, 0, W - 1);
                const size_t tap = (size_t)(j * win + i) * n;
                dl[u] = __ldg(dl_p + tap);
                cr[u] = __ldg(prow + nx);
#pragma unroll
                for (int q = 0; q < GEO_Q; q++) {
                    const int nxs = g.sign > 0 ? max(0, nx - dq[q]) : min(nx + dq[q], W - 1);
                    ct[u][q] = __ldg(trow + nxs);
                    dr[u][q] = __ldg(dtgt + tap + pt[q]);
                }
            }
#pragma unroll
            for (int u = 0; u < 7; u++) {
                if (i0 + u < win) {
#pragma unroll
                    for (int q = 0; q < GEO_Q; q++) {
                        const float cd = (float)__vsadu4(cr[u], ct[u][q]);          // getColorDist (A.cpp:1321-1326)
                        const float t = __fmul_rn(dl[u], dr[u][q]);                 // float * float (A.cpp:1488-1489)
                        fn[q] = __fadd_rn(fn[q], __fmul_rn(t, cd));
                        fd[q] = __fadd_rn(fd[q], t);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < GEO_Q; q++) { num[q] += (double)fn[q]; den[q] += (double)fd[q]; }
    }
    unsigned long long best = WTA_KEY_EMPTY;
#pragma unroll
    for (int q = 0; q < GEO_Q; q++) {
        if (q < nq) {
            double E = num[q] / den[q];
            if (agg) agg[(size_t)(c_base + c0 + q) * n + p] = (float)E;
            // c_base > 0: the lower candidates come from another kernel (other summation order).  Where every operand of
            // candidate d is clamped to the same pixels as for d - 1 (x + h <= d - 1 for LEFT, x - h >= W - d for RIGHT)
            // the reference's two costs are the same number and strict < keeps the lower d: d cannot win there.
            const bool same_as_lower = c_base > 0 && (g.sign > 0 ? x + h <= dq[q] - 1 : x - h >= W - dq[q]);
            if (!same_as_lower) best = min(best, wta_key_d(E, dq[q]));
        }
    }
    atomicMin(&keys[p], best);
}

// ------------------------------------------------------------------------------------------------
// tiled aggregation: one CTA = one image row x GT_X pixels x 8*KC candidates.  Every thread owns 4 ADJACENT
// pixels and KC candidates, so one LDS.128 feeds 4 taps.  Operands that are addressed at (x - d) are staged in
// shared memory as 4 copies skewed by 0..3 elements: whatever (d, tap column) is, the 4 consecutive values a
// thread needs start on a 16-byte boundary in one of the copies (the copy index is uniform across the warp).
// Staged cells hold CLAMPED image columns, so max(0, x-d) / min(x+d, W-1) cost nothing; only the segment
// whose window sample columns get clamped BEFORE the disparity shift (right image edge for LEFT, left edge
// for RIGHT) takes the scalar BORDER path.  Per tap: VABSDIFF4.ACC, I2F, FMUL, FFMA, FADD + 2/4 LDS.128.
// All terms are non-negative; accumulation is fp32 (relative error ~1e-6 over the 1225 taps, two orders of
// magnitude inside the 1e-4 budget).  KC = 4 for full chunks of 32 candidates, smaller KC for the remainder
// (D+1 candidates is never a multiple of 32) so that no candidate slot is wasted.
// ------------------------------------------------------------------------------------------------
#define GT_X 128
#define GT_TC 12
#define GT_THREADS 256

template <bool BORDER, int KC>
__global__ void __launch_bounds__(GT_THREADS, 2)
k_geo_agg_tile(const float* __restrict__ dref, const float* __restrict__ dtgt, const uint32_t* __restrict__ pref,
               const uint32_t* __restrict__ ptgt, GeoGeom g, int seg_first, int cand_first,
               unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    extern __shared__ __align__(16) float sm_gt[];
    constexpr int DC = 8 * KC;                                // candidates per CTA
    constexpr int DRW = GT_X + DC + 4;                        // staged target-distance row (cells)
    const int W = g.W, H = g.H, win = g.win, h = g.h, sign = g.sign;
    const int CLW = (GT_X + 2 * h + 7) & ~3;                  // reference colour row
    const int CRW = (GT_X + 2 * h + DC + 7) & ~3;             // target colour row
    float* DLs = sm_gt;                                       // [GT_TC][GT_X]
    float* DRs = DLs + GT_TC * GT_X;                          // [4][GT_TC][DRW]
    uint32_t* CL = (uint32_t*)(DRs + 4 * GT_TC * DRW);        // [4][CLW]
    uint32_t* CR = CL + 4 * CLW;                              // [4][CRW]
    const int tid = threadIdx.x, pg = tid & 31, ds = tid >> 5;
    const int y = blockIdx.y, xb = (seg_first + blockIdx.x) * GT_X;
    const int c0 = cand_first + blockIdx.z * DC;
    const int d_lo = g.d_first + c0, d_hi = d_lo + DC - 1;
    const size_t n = (size_t)H * W;
    const size_t rowoff = (size_t)y * W;
    // candidate k of this thread: c0 + ds + 8k; sh = its shift inside the staged rows (0 .. DC-1)
    int sh[KC];
    const float* drp[KC];                                      // per-candidate base into the skewed DR copies
    const uint32_t* crp[KC];                                   // unaligned start (copy 0) of the target colour quad
#pragma unroll
    for (int k = 0; k < KC; k++) {
        int d = g.d_first + min(c0 + ds + 8 * k, g.n_cand - 1);
        sh[k] = sign > 0 ? d_hi - d : d - d_lo;
        drp[k] = DRs + (sh[k] & 3) * GT_TC * DRW + 4 * pg + (sh[k] & ~3);
        crp[k] = CR + 4 * pg + sh[k];
    }
    const int oDR = sign > 0 ? xb - d_hi : xb + d_lo;         // image column of staged cell 0 (before clamping)
    const int oCL = xb - h;
    const int oCR = sign > 0 ? xb - h - d_hi : xb - h + d_lo;
    float fn[KC * 4], fd[KC * 4];
#pragma unroll
    for (int a = 0; a < KC * 4; a++) { fn[a] = 0.0f; fd[a] = 0.0f; }
    const int x0 = xb + 4 * pg;
    const uint32_t* clp = CL + 4 * pg;

    for (int j = 0; j < win; j++) {
        const int ny = clampi(y - h + j, 0, H - 1);
        __syncthreads();                                       // previous row's colour rows are dead
        for (int e = tid; e < CLW; e += GT_THREADS) {
            uint32_t v = pref[(size_t)ny * W + clampi(oCL + e, 0, W - 1)];
#pragma unroll
            for (int r = 0; r < 4; r++) if (e - r >= 0) CL[r * CLW + e - r] = v;
        }
        for (int e = tid; e < CRW; e += GT_THREADS) {
            uint32_t v = ptgt[(size_t)ny * W + clampi(oCR + e, 0, W - 1)];
#pragma unroll
            for (int r = 0; r < 4; r++) if (e - r >= 0) CR[r * CRW + e - r] = v;
        }
        for (int i0 = 0; i0 < win; i0 += GT_TC) {
            const int i1 = min(i0 + GT_TC, win);
            __syncthreads();                                   // previous chunk consumed
            for (int q = tid; q < (i1 - i0) * GT_X; q += GT_THREADS) {
                int tt = q / GT_X, xx = q - tt * GT_X;
                DLs[q] = dref[(size_t)(j * win + i0 + tt) * n + rowoff + min(xb + xx, W - 1)];
            }
            for (int q = tid; q < (i1 - i0) * DRW; q += GT_THREADS) {
                int tt = q / DRW, e = q - tt * DRW;
                float v = dtgt[(size_t)(j * win + i0 + tt) * n + rowoff + clampi(oDR + e, 0, W - 1)];
#pragma unroll
                for (int r = 0; r < 4; r++) if (e - r >= 0) DRs[(r * GT_TC + tt) * DRW + e - r] = v;
            }
            __syncthreads();
            for (int i = i0; i < i1; i++) {
                const int tt = i - i0;
                const float4 dl = *(const float4*)(DLs + tt * GT_X + 4 * pg);
                uint32_t cl[4];
                if (!BORDER) {
                    // copy r = i & 3 at aligned index i - r: address = clp + i + r * (CLW - 1)
                    const uint4 c4 = *(const uint4*)(clp + i + (i & 3) * (CLW - 1));
                    cl[0] = c4.x; cl[1] = c4.y; cl[2] = c4.z; cl[3] = c4.w;
                } else {
#pragma unroll
                    for (int p = 0; p < 4; p++) cl[p] = CL[clampi(clampi(x0 + p - h + i, 0, W - 1) - oCL, 0, CLW - 1)];
                }
#pragma unroll
                for (int k = 0; k < KC; k++) {
                    const float4 dr = *(const float4*)(drp[k] + tt * DRW);
                    uint32_t cr[4];
                    if (!BORDER) {
                        const uint4 c4 = *(const uint4*)(crp[k] + i + ((sh[k] + i) & 3) * (CRW - 1));
                        cr[0] = c4.x; cr[1] = c4.y; cr[2] = c4.z; cr[3] = c4.w;
                    } else {
                        const int d = sign > 0 ? d_hi - sh[k] : d_lo + sh[k];
#pragma unroll
                        for (int p = 0; p < 4; p++) {
                            int nx = clampi(x0 + p - h + i, 0, W - 1);
                            int col = sign > 0 ? max(0, nx - d) : min(nx + d, W - 1);
                            cr[p] = CR[clampi(col - oCR, 0, CRW - 1)];   // cells past either end hold the clamped column
                        }
                    }
                    const float dlv[4] = {dl.x, dl.y, dl.z, dl.w};
                    const float drv[4] = {dr.x, dr.y, dr.z, dr.w};
#pragma unroll
                    for (int p = 0; p < 4; p++) {
                        float t = __fmul_rn(dlv[p], drv[p]);                              // A.cpp:1488-1489
                        float cd = (float)__vsadu4(cl[p], cr[p]);                         // getColorDist
                        fn[k * 4 + p] = fmaf(t, cd, fn[k * 4 + p]);
                        fd[k * 4 + p] = __fadd_rn(fd[k * 4 + p], t);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int x = x0 + p;
        if (x >= W) continue;
        unsigned long long best = WTA_KEY_EMPTY;
#pragma unroll
        for (int k = 0; k < KC; k++) {
            int c = c0 + ds + 8 * k;
            if (c < g.n_cand) {
                double E = (double)fn[k * 4 + p] / (double)fd[k * 4 + p];
                if (agg) agg[(size_t)c * n + rowoff + x] = (float)E;
                best = min(best, wta_key_d(E, g.d_first + c));
            }
        }
        atomicMin(&keys[rowoff + x], best);
    }
}

template <bool BORDER, int KC>
static asw_status geo_tile_launch(asw_ctx* ctx, const float* dref, const float* dtgt, const uint32_t* cref, const uint32_t* ctgt,
                                  GeoGeom g, int seg_first, int seg_count, int cand_first, int n_chunks,
                                  unsigned long long* keys, float* agg) {
    if (seg_count <= 0 || n_chunks <= 0) return ASW_OK;
    const int h = g.h, DC = 8 * KC;
    const int CLW = (GT_X + 2 * h + 7) & ~3, CRW = (GT_X + 2 * h + DC + 7) & ~3, DRW = GT_X + DC + 4;
    size_t smem = ((size_t)GT_TC * GT_X + (size_t)4 * GT_TC * DRW + (size_t)4 * (CLW + CRW)) * sizeof(float);
    cudaFuncSetAttribute(k_geo_agg_tile<BORDER, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LAUNCH(ctx, BORDER ? "geo_aggregate_border" : "geo_aggregate",
           (k_geo_agg_tile<BORDER, KC><<<dim3(seg_count, g.H, n_chunks), GT_THREADS, smem, ctx->stream>>>(
               dref, dtgt, cref, ctgt, g, seg_first, cand_first, keys, agg)));
    return ASW_OK;
}
// all candidates of a segment range: full chunks of 32 (KC = 4), then one remainder chunk with the smallest KC
template <bool BORDER>
static asw_status geo_tile_segments(asw_ctx* ctx, const float* dref, const float* dtgt, const uint32_t* cref, const uint32_t* ctgt,
                                    GeoGeom g, int seg_first, int seg_count, unsigned long long* keys, float* agg) {
    int full = g.n_cand / 32, rem = g.n_cand - full * 32;
    ASW_TRY((geo_tile_launch<BORDER, 4>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, 0, full, keys, agg)));
    if (rem > 24) ASW_TRY((geo_tile_launch<BORDER, 4>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, full * 32, 1, keys, agg)));
    else if (rem > 16) ASW_TRY((geo_tile_launch<BORDER, 3>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, full * 32, 1, keys, agg)));
    else if (rem > 8) ASW_TRY((geo_tile_launch<BORDER, 2>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, full * 32, 1, keys, agg)));
    else if (rem > 0) ASW_TRY((geo_tile_launch<BORDER, 1>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, full * 32, 1, keys, agg)));
    return ASW_OK;
}

// ------------------------------------------------------------------------------------------------
// diagonal-blocked aggregation (full chunks of 32 candidates on interior segments).  Same CTA tile as
// k_geo_agg_tile (one image row x 128 pixels x 32 candidates) but a thread owns 4 adjacent pixels x 4 ADJACENT
// candidates: the operands addressed at (x - d) then fall on a diagonal, 7 distinct values for the 16
// (pixel, candidate) pairs, fetched as one aligned 8-wide window (2 LDS.128) instead of one quad per candidate
// (4 LDS.128): 6 LDS.128 per tap instead of 10 for the same 16 evaluations.  Because a warp's 4 candidates start
// at a multiple of 4, the window alignment of the target distances is CTA-uniform: one staged copy, no skewed
// replicas.  Staging is asynchronous: the next 12-tap chunk of distance rows (and the next window row's colours)
// are copied global -> shared with cp.async (4-byte, clamped columns) while the current chunk is evaluated.
// ------------------------------------------------------------------------------------------------
#ifndef GEO_FULL_KC
// full candidate chunks: 8 candidates per thread x 8 warps = 64 candidates per CTA, 2 CTAs per SM (128 registers).
// Measured at config 4 (1280 x 720, 129 candidates, 35 x 35): 4 x 8 warps (round 1) 32.0 ms; 8 x 4 warps 29.9 (3 CTAs) /
// 27.6 (4 CTAs); 8 x 8 warps 25.9; 8 x 16 warps 25.6 ms
#define GEO_FULL_KC 8
#define GEO_FULL_NW 8
#define GEO_FULL_MINB 2      // resident CTAs the register budget is cut for
#endif
#ifndef GEO_SAD_I2F
#define GEO_SAD_I2F 1
#endif
#define GEO_FULL_CAND (GEO_FULL_KC * GEO_FULL_NW)
// staged target-distance row: 4*31 + KC (NW - 1) + KC + 4 cells
#define GD_DRW (GEO_FULL_CAND > 32 ? 4 * 31 + GEO_FULL_CAND + 4 + 0 : 160)
__device__ __forceinline__ void gd_cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// sum of the 4 byte-wise absolute differences added onto an accumulator (one VABSDIFF4.ACC)
__device__ __forceinline__ uint32_t gd_sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}
// two SADs straight into the halves of one 64-bit register pair (the operand of the packed FADD2 that follows)
__device__ __forceinline__ float2 gd_sad4_acc_pair(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, uint32_t acc) {
    unsigned long long r;
    asm("{\n\t.reg .b32 lo, hi;\n\tvabsdiff4.u32.u32.u32.add lo, %1, %2, %5;\n\tvabsdiff4.u32.u32.u32.add hi, %3, %4, %5;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}" : "=l"(r) : "r"(a0), "r"(b0), "r"(a1), "r"(b1), "r"(acc));
    float2 f;
    f.x = __uint_as_float((uint32_t)r); f.y = __uint_as_float((uint32_t)(r >> 32));
    return f;
}
// 16-byte shared-memory loads the compiler cannot split: with 8 candidates per thread not every element of the last quad is
// used, and ptxas otherwise narrows that load to LDS.64 + LDS whose 16-byte lane stride is a 2- / 4-way bank conflict
__device__ __forceinline__ float4 gd_lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ uint4 gd_lds128(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void gd_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gd_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// BORDER: the segment touches the image edge where the window's sample column is clamped BEFORE the disparity
// shift (right edge for LEFT, left edge for RIGHT).  Those taps read the target colour at the shifted EDGE column,
// which depends on the candidate only: 4 values per thread and window row, selected per evaluation.
// A thread owns 4 adjacent pixels x KC ADJACENT candidates; NW warps per CTA = KC NW candidates per CTA.  Full chunks of 32
// candidates run KC = 8, NW = 4: the 11 distinct target operands of the 32 evaluations are one aligned 12-wide window
// (3 LDS.128 for distances, 3 for colours) -- 8 LDS.128 per tap and 32 evaluations, against 12 for two 4 x 4 threads.  The
// kernel is bound by the shared-memory pipe, so that ratio is its speed.  The candidate remainder runs KC = 4 with as many
// warps as it needs.
template <int SIGN, bool BORDER, int NW, int KC>
__global__ void __launch_bounds__(32 * NW, KC == 8 ? GEO_FULL_MINB : (NW == 8 ? 2 : 1))
k_geo_agg_diag(const float* __restrict__ dref, const float* __restrict__ dtgt, const uint32_t* __restrict__ pref,
               const uint32_t* __restrict__ ptgt, GeoGeom g, int seg_first, int cand_first,
               unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    constexpr int NT = 32 * NW;
    extern __shared__ __align__(16) float sm_gd[];
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const int CLW = (GT_X + 2 * h + 7) & ~3;                  // reference colour row (cells)
    const int CRW = (GD_DRW + 2 * h + 7) & ~3;                // target colour row
    float* DLs = sm_gd;                                       // [2][GT_TC][GT_X]
    float* DRs = DLs + 2 * GT_TC * GT_X;                      // [2][GT_TC][GD_DRW]
    uint32_t* CL = (uint32_t*)(DRs + 2 * GT_TC * GD_DRW);     // [2][4][CLW]   4 copies skewed by 0..3 cells
    uint32_t* CR = CL + 2 * 4 * CLW;                          // [2][4][CRW]
    const int tid = threadIdx.x, pg = tid & 31, ds = tid >> 5;
    const int y = blockIdx.y, xb = (seg_first + blockIdx.x) * GT_X;
    const int c0 = cand_first + blockIdx.z * (KC * NW);
    const int d_lo = g.d_first + c0;
    const size_t n = (size_t)H * W;
    const size_t rowoff = (size_t)y * W;
    // staged cell 0 <-> image column (before clamping).  LEFT: a thread's (KC + 4)-wide window starts at
    // x - (d_lo + KC ds) - KC = oDR + (4 pg - KC ds + KC (NW - 1)); RIGHT: at x + d_lo + KC ds = oDR + (4 pg + KC ds)
    static_assert(KC % 4 == 0 && 4 * 31 + KC * (NW - 1) + KC + 4 <= GD_DRW, "staged target row too narrow");
    const int oDR = SIGN > 0 ? xb - d_lo - KC * NW : xb + d_lo;
    const int oCL = xb - h;
    const int oCR = oDR - h;
    const int e0 = SIGN > 0 ? 4 * pg - KC * ds + KC * (NW - 1) : 4 * pg + KC * ds;   // window start (cells), multiple of 4
    // accumulators: the 4 KC (pixel, candidate) pairs of a thread are evaluated as 2 (KC - 1) packed pairs that share a
    // target operand -- (p, k) and (p+1, k+SIGN), p even, sit on the same diagonal -- plus 4 singles, so that the FMA-pipe
    // work of two evaluations issues as one FMUL2 / FADD2 / FFMA2
    float2 fnp[2][KC - 1], fdp[2][KC - 1];
    float fns[2][2], fds[2][2];
#pragma unroll
    for (int a = 0; a < 2; a++) {
#pragma unroll
        for (int q = 0; q < KC - 1; q++) { fnp[a][q] = make_float2(0.f, 0.f); fdp[a][q] = make_float2(0.f, 0.f); }
        fns[a][0] = fns[a][1] = fds[a][0] = fds[a][1] = 0.0f;
    }

    // chunk c of the window: row j = c / 3, taps i0 .. i1 of that row (3 chunks per row: 12 + 12 + rest)
    const int cpr = (win + GT_TC - 1) / GT_TC, nchunk = win * cpr;
    // staging map without divisions: thread `tid` copies column `tid` (+ NT, ...) of every tap row; the clamped
    // source column and the smem column are loop invariants, the tap advances by one plane (n floats)
    auto stage_dist = [&](int c, int buf) {
        const int j = c / cpr, i0 = (c - j * cpr) * GT_TC, cnt = min(GT_TC, win - i0);
        const size_t tap0 = (size_t)(j * win + i0) * n + rowoff;
        for (int xx = tid; xx < GT_X; xx += NT) {
            const float* src = dref + tap0 + min(xb + xx, W - 1);
            float* dst = DLs + buf * GT_TC * GT_X + xx;
#pragma unroll 4
            for (int tt = 0; tt < cnt; tt++) gd_cp_async4(dst + tt * GT_X, src + (size_t)tt * n);
        }
        for (int e = tid; e < GD_DRW; e += NT) {
            const float* src = dtgt + tap0 + clampi(oDR + e, 0, W - 1);
            float* dst = DRs + buf * GT_TC * GD_DRW + e;
#pragma unroll 4
            for (int tt = 0; tt < cnt; tt++) gd_cp_async4(dst + tt * GD_DRW, src + (size_t)tt * n);
        }
    };
    auto stage_colour = [&](int j, int buf) {
        const int ny = clampi(y - h + j, 0, H - 1);
        uint32_t* cl = CL + buf * 4 * CLW;
        uint32_t* cr = CR + buf * 4 * CRW;
        for (int e = tid; e < CLW; e += NT) {
            const uint32_t* src = pref + (size_t)ny * W + clampi(oCL + e, 0, W - 1);
#pragma unroll
            for (int r = 0; r < 4; r++) if (e - r >= 0) gd_cp_async4(cl + r * CLW + e - r, src);
        }
        for (int e = tid; e < CRW; e += NT) {
            const uint32_t* src = ptgt + (size_t)ny * W + clampi(oCR + e, 0, W - 1);
#pragma unroll
            for (int r = 0; r < 4; r++) if (e - r >= 0) gd_cp_async4(cr + r * CRW + e - r, src);
        }
    };
    stage_colour(0, 0);
    stage_dist(0, 0);
    gd_cp_commit();
    for (int c = 0; c < nchunk; c++) {
        const int j = c / cpr, i0 = (c - j * cpr) * GT_TC, i1 = min(i0 + GT_TC, win);
        gd_cp_wait_all();
        __syncthreads();                                       // chunk c (and row j's colours) landed; chunk c-1 consumed
        if (c + 1 < nchunk) {
            const int jn = (c + 1) / cpr;
            if (jn != j) stage_colour(jn, jn & 1);
            stage_dist(c + 1, (c + 1) & 1);
            gd_cp_commit();
        }
        const float* dl = DLs + (c & 1) * GT_TC * GT_X + 4 * pg;
        const float* dr = DRs + (c & 1) * GT_TC * GD_DRW + e0;
        const uint32_t* cl = CL + (j & 1) * 4 * CLW + 4 * pg;
        const uint32_t* cr = CR + (j & 1) * 4 * CRW + e0;
        uint32_t edge[KC];
        if (BORDER) {
            const int ny = clampi(y - h + j, 0, H - 1);
#pragma unroll
            for (int k = 0; k < KC; k++) {
                const int d = d_lo + KC * ds + k;
                edge[k] = __ldg(ptgt + (size_t)ny * W + (SIGN > 0 ? max(0, W - 1 - d) : min(d, W - 1)));
            }
        }
        // i0 is a multiple of GT_TC = 12, so the skewed colour copy r = i & 3 = tt & 3 is static after unrolling; the
        // four per-copy bases are chunk invariants, a tap then addresses everything with immediate offsets
        const uint32_t* clb[4];
        const uint32_t* crb4[4];
#pragma unroll
        for (int r = 0; r < 4; r++) { clb[r] = cl + r * CLW + i0 - r; crb4[r] = cr + r * CRW + i0 - r; }
#pragma unroll
        for (int tt = 0; tt < GT_TC; tt++) {
            if (tt < i1 - i0) {
                const int i = i0 + tt, r = tt & 3;
                const float4 dl4 = *(const float4*)(dl + tt * GT_X);
                // colour cell of pixel quad start: oCL + (4 pg + i) -> copy r at aligned index 4 pg + i - r
                const uint4 cl4 = *(const uint4*)(clb[r] + tt);
                const float dlv[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
                const uint32_t clv[4] = {cl4.x, cl4.y, cl4.z, cl4.w};
                float drv[KC + 4];
                uint32_t crv[KC + 4];
#pragma unroll
                for (int m = 0; m < (KC + 4) / 4; m++) {
                    const float4 d4 = gd_lds128(dr + tt * GD_DRW + 4 * m);
                    const uint4 c4 = gd_lds128(crb4[r] + tt + 4 * m);
                    drv[4 * m] = d4.x; drv[4 * m + 1] = d4.y; drv[4 * m + 2] = d4.z; drv[4 * m + 3] = d4.w;
                    crv[4 * m] = c4.x; crv[4 * m + 1] = c4.y; crv[4 * m + 2] = c4.z; crv[4 * m + 3] = c4.w;
                }
                bool clamped[4];
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const int nx = xb + 4 * pg + p - h + i;                   // sample column before clamping
                    clamped[p] = BORDER && (SIGN > 0 ? nx > W - 1 : nx < 0);
                }
                // getColorDist as a float without an integer->float conversion: the byte SAD accumulates onto the bit
                // pattern of 2^23 (VABSDIFF4.ACC), giving the float 2^23 + sad exactly; 2^23 is subtracted on the FMA pipe
                auto sadbits = [&](int p, int k, int w) {
                    const uint32_t cright = BORDER ? (clamped[p] ? edge[k] : crv[w]) : crv[w];
                    return __uint_as_float(gd_sad4_acc(clv[p], cright, 0x4B000000u));
                };
                const float2 m23 = make_float2(-8388608.0f, -8388608.0f);
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const int pe = 2 * a;
                    const float2 dl2 = make_float2(dlv[pe], dlv[pe + 1]);
#pragma unroll
                    for (int q = 0; q < KC - 1; q++) {
                        const int k = SIGN > 0 ? q : q + 1;                   // (pe, k) and (pe+1, k+SIGN): same window cell
                        const int w = SIGN > 0 ? pe - k + KC : pe + k;
                        const float2 t2 = __fmul2_rn(dl2, make_float2(drv[w], drv[w]));          // A.cpp:1488-1489
                        const uint32_t cr0 = BORDER ? (clamped[pe] ? edge[k] : crv[w]) : crv[w];
                        const uint32_t cr1 = BORDER ? (clamped[pe + 1] ? edge[k + SIGN] : crv[w]) : crv[w];
#if GEO_SAD_I2F
                        // byte SADs converted on the ALU side (I2FP): the FMA pipe is the loaded one in this kernel
                        const float2 cd2 = make_float2((float)gd_sad4_acc(clv[pe], cr0, 0u), (float)gd_sad4_acc(clv[pe + 1], cr1, 0u));
#else
                        const float2 cd2 = __fadd2_rn(gd_sad4_acc_pair(clv[pe], cr0, clv[pe + 1], cr1, 0x4B000000u), m23);
#endif
                        fnp[a][q] = __ffma2_rn(t2, cd2, fnp[a][q]);
                        fdp[a][q] = __fadd2_rn(fdp[a][q], t2);
                    }
                    // singles: (pe, ks0) and (pe+1, ks1) have no partner on their diagonal
                    const int ks0 = SIGN > 0 ? KC - 1 : 0, ks1 = SIGN > 0 ? 0 : KC - 1;
                    const int w0 = SIGN > 0 ? pe - ks0 + KC : pe + ks0, w1 = SIGN > 0 ? pe + 1 - ks1 + KC : pe + 1 + ks1;
                    const float t0 = __fmul_rn(dlv[pe], drv[w0]), t1 = __fmul_rn(dlv[pe + 1], drv[w1]);
                    fns[a][0] = fmaf(t0, sadbits(pe, ks0, w0) - 8388608.0f, fns[a][0]);
                    fds[a][0] = __fadd_rn(fds[a][0], t0);
                    fns[a][1] = fmaf(t1, sadbits(pe + 1, ks1, w1) - 8388608.0f, fns[a][1]);
                    fds[a][1] = __fadd_rn(fds[a][1], t1);
                }
            }
        }
    }
    const int x0 = xb + 4 * pg;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int x = x0 + p;
        if (x >= W) continue;
        unsigned long long best = WTA_KEY_EMPTY;
#pragma unroll
        for (int k = 0; k < KC; k++) {
            const int c = c0 + KC * ds + k;
            if (c >= g.n_cand) continue;
            // (p, k) lives in: pair slot q of half a = p/2 (x for even p with k = q (LEFT) / q+1 (RIGHT), y for odd p
            // with the partner's k), or one of the two singles of that half
            const int a = p >> 1, odd = p & 1;
            const int kq = SIGN > 0 ? (odd ? k - 1 : k) : (odd ? k : k - 1);          // pair slot if in 0..KC-2
            const bool single = odd ? (k == (SIGN > 0 ? 0 : KC - 1)) : (k == (SIGN > 0 ? KC - 1 : 0));
            float vn, vd;
            if (single) { vn = fns[a][odd]; vd = fds[a][odd]; }
            else { vn = odd ? fnp[a][kq].y : fnp[a][kq].x; vd = odd ? fdp[a][kq].y : fdp[a][kq].x; }
            const double E = (double)vn / (double)vd;
            if (agg) agg[(size_t)c * n + rowoff + x] = (float)E;
            best = min(best, wta_key_d(E, g.d_first + c));
        }
        atomicMin(&keys[rowoff + x], best);
    }
}

template <int SIGN, bool BORDER, int NW, int KC>
static asw_status geo_diag_launch(asw_ctx* ctx, const float* dref, const float* dtgt, const uint32_t* cref, const uint32_t* ctgt,
                                  GeoGeom g, int seg_first, int seg_count, int cand_first, int n_chunks,
                                  unsigned long long* keys, float* agg) {
    if (seg_count <= 0 || n_chunks <= 0) return ASW_OK;
    const int h = g.h;
    const int CLW = (GT_X + 2 * h + 7) & ~3, CRW = (GD_DRW + 2 * h + 7) & ~3;
    size_t smem = (2 * (size_t)GT_TC * GT_X + 2 * (size_t)GT_TC * GD_DRW + 8 * (size_t)(CLW + CRW)) * sizeof(float);
    cudaFuncSetAttribute(k_geo_agg_diag<SIGN, BORDER, NW, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LAUNCH(ctx, BORDER ? "geo_aggregate_border" : "geo_aggregate",
           (k_geo_agg_diag<SIGN, BORDER, NW, KC><<<dim3(seg_count, g.H, n_chunks), 32 * NW, smem, ctx->stream>>>(
                                     dref, dtgt, cref, ctgt, g, seg_first, cand_first, keys, agg)));
    return ASW_OK;
}
// a range of segments: full chunks of 32 candidates with 4 warps x 8 candidates per thread, the candidate remainder with
// as many warps as it needs at 4 candidates per thread
template <int SIGN, bool BORDER>
static asw_status geo_segments_signed(asw_ctx* ctx, const float* dref, const float* dtgt, const uint32_t* cref,
                                      const uint32_t* ctgt, GeoGeom g, int seg_first, int seg_count,
                                      unsigned long long* keys, float* agg) {
    const int full = g.n_cand / GEO_FULL_CAND, rem = g.n_cand - full * GEO_FULL_CAND;
    ASW_TRY((geo_diag_launch<SIGN, BORDER, GEO_FULL_NW, GEO_FULL_KC>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, 0, full, keys, agg)));
    int nw = (rem + 3) / 4, c1 = full * GEO_FULL_CAND;
    if (nw > 8) {   // more than 32 left (64-candidate chunks): one 32-candidate launch first
        ASW_TRY((geo_diag_launch<SIGN, BORDER, 8, 4>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, c1, 1, keys, agg)));
        c1 += 32; nw -= 8;
    }
#define GEO_REM(NWV) ASW_TRY((geo_diag_launch<SIGN, BORDER, NWV, 4>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, c1, 1, keys, agg)))
    switch (nw) {
        case 1: GEO_REM(1); break;
        case 2: GEO_REM(2); break;
        case 3: case 4: GEO_REM(4); break;
        case 5: case 6: GEO_REM(6); break;
        case 7: case 8: GEO_REM(8); break;
        default: break;
    }
#undef GEO_REM
    return ASW_OK;
}
template <bool BORDER>
static asw_status geo_segments(asw_ctx* ctx, const float* dref, const float* dtgt, const uint32_t* cref,
                               const uint32_t* ctgt, GeoGeom g, int seg_first, int seg_count,
                               unsigned long long* keys, float* agg) {
#ifdef ASW_DEV_KERNELS
    if (asw_dev("ASW_GEO_TILE")) return geo_tile_segments<BORDER>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, keys, agg);
#endif
    if (g.sign > 0) return geo_segments_signed<1, BORDER>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, keys, agg);
    return geo_segments_signed<-1, BORDER>(ctx, dref, dtgt, cref, ctgt, g, seg_first, seg_count, keys, agg);
}

__global__ void k_pack_bgrx(const uint8_t* __restrict__ img, size_t n, uint32_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (uint32_t)img[3 * i] | ((uint32_t)img[3 * i + 1] << 8) | ((uint32_t)img[3 * i + 2] << 16);
}
// [taps][n] -> [n][taps] (oracle / reference layout of one window per pixel)
__global__ void k_geo_transpose(const float* __restrict__ in, size_t n, int taps, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * taps) return;
    size_t p = i / taps; int t = (int)(i - p * taps);
    out[i] = in[(size_t)t * n + p];
}

static asw_status dev_geodesic_dist(asw_ctx* ctx, const uint8_t* img, int H, int W, int win, int ws_slot, float** dist_out) {
    if (win > GEO_MAXW) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "geodesic window larger than 63%s%s");
    size_t n = (size_t)H * W;
    int pad = win / 2 + 1, Wp = W + 2 * pad, Hp = H + 2 * pad;
    uint32_t* ext;
    float* dist;
    ASW_TRY(ws_get(ctx, WS_MISC2, (size_t)Hp * Wp, &ext));
    ASW_TRY(ws_get(ctx, ws_slot, n * win * win, &dist));
    LAUNCH(ctx, "pack_bgrx_pad", (k_pack_bgrx_pad<<<dim3(cdiv(Wp, 128), Hp), 128, 0, ctx->stream>>>(img, H, W, pad, ext)));   // A.cpp:1404
    dim3 dgrid(cdiv(W, 128), H);
    if (asw_dev("ASW_GEO_DIST_GENERIC")) LAUNCH(ctx, "geo_dist", (k_geo_dist<<<dgrid, 128, 0, ctx->stream>>>(ext, H, W, win, dist)));
    else if (win == 35) LAUNCH(ctx, "geo_dist", (k_geo_dist_t<35><<<dgrid, 128, 0, ctx->stream>>>(ext, H, W, dist)));
    else if (win == 9) LAUNCH(ctx, "geo_dist", (k_geo_dist_t<9><<<dgrid, 128, 0, ctx->stream>>>(ext, H, W, dist)));
    else if (win == 7) LAUNCH(ctx, "geo_dist", (k_geo_dist_t<7><<<dgrid, 128, 0, ctx->stream>>>(ext, H, W, dist)));
    else if (win == 5) LAUNCH(ctx, "geo_dist", (k_geo_dist_t<5><<<dgrid, 128, 0, ctx->stream>>>(ext, H, W, dist)));
    else LAUNCH(ctx, "geo_dist", (k_geo_dist<<<dgrid, 128, 0, ctx->stream>>>(ext, H, W, win, dist)));
    *dist_out = dist;
    return ASW_OK;
}

static asw_status dev_geodesic_dist_to_host(asw_ctx* ctx, const uint8_t* img, int H, int W, int win, float* host) {
    size_t n = (size_t)H * W; int taps = win * win;
    float *dist, *tr;
    ASW_TRY(dev_geodesic_dist(ctx, img, H, W, win, WS_GEO_L, &dist));
    ASW_TRY(ws_get(ctx, WS_GEO_R, n * taps, &tr));
    LAUNCH(ctx, "geo_transpose", (k_geo_transpose<<<(unsigned)((n * taps + 255) / 256), 256, 0, ctx->stream>>>(dist, n, taps, tr)));
    ASW_CUDA(ctx, cudaMemcpyAsync(host, tr, n * taps * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ASW_OK;
}

static asw_status dev_geodesic(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, int win,
                               int min_d, int num_d, float* disp_dev, float* agg_dev) {
    size_t n = (size_t)H * W;
    float *distL, *distR;
    ASW_TRY(dev_geodesic_dist(ctx, dL, H, W, win, WS_GEO_L, &distL));     // A.cpp:1464-1465
    ASW_TRY(dev_geodesic_dist(ctx, dR, H, W, win, WS_GEO_R, &distR));
    uint32_t *pl, *pr;
    ASW_TRY(ws_get(ctx, WS_TMP0, n, &pl));
    ASW_TRY(ws_get(ctx, WS_TMP1, n, &pr));
    LAUNCH(ctx, "pack_bgrx", (k_pack_bgrx<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dL, n, pl)));
    LAUNCH(ctx, "pack_bgrx", (k_pack_bgrx<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dR, n, pr)));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    GeoGeom g;
    g.H = H; g.W = W; g.win = win; g.h = win / 2;
    g.sign = disp_type == ASW_DISPARITY_LEFT ? 1 : -1;
    g.d_first = min_d; g.n_cand = num_d + 1;                              // A.cpp:1447, 1467
    bool left = disp_type == ASW_DISPARITY_LEFT;
    const float* dref = left ? distL : distR; const float* dtgt = left ? distR : distL;
    const uint32_t* cref = left ? pl : pr; const uint32_t* ctgt = left ? pr : pl;
    if (asw_dev("ASW_GEO_GENERIC")) {
        dim3 grid(cdiv(W, 128), H, cdiv(g.n_cand, 4));
        LAUNCH(ctx, "geo_aggregate", (k_geo_aggregate_q<4><<<grid, 128, 0, ctx->stream>>>(dref, dtgt, cref, ctgt, g, 0, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    // D+1 candidates is one more than a multiple of 32 for the usual disparity ranges: a remainder of <= 4 candidates
    // would run the diagonal kernel with one warp per CTA (all the staging, 1/8 of the work).  The thread-per-pixel
    // kernel evaluates exactly those candidates over the whole image instead.
    const int rem = g.n_cand % 32;
    if (rem >= 1 && rem <= 4 && g.n_cand > 32 && !asw_dev("ASW_GEO_DIAG_REM")) {
        GeoGeom gr = g;
        const int c1 = g.n_cand - rem;
        gr.d_first = g.d_first + c1; gr.n_cand = rem;
        dim3 grid(cdiv(W, 128), H, 1);
        switch (rem) {
            case 1: LAUNCH(ctx, "geo_aggregate_rem", (k_geo_aggregate_q<1><<<grid, 128, 0, ctx->stream>>>(dref, dtgt, cref, ctgt, gr, c1, keys, agg_dev))); break;
            case 2: LAUNCH(ctx, "geo_aggregate_rem", (k_geo_aggregate_q<2><<<grid, 128, 0, ctx->stream>>>(dref, dtgt, cref, ctgt, gr, c1, keys, agg_dev))); break;
            case 3: LAUNCH(ctx, "geo_aggregate_rem", (k_geo_aggregate_q<3><<<grid, 128, 0, ctx->stream>>>(dref, dtgt, cref, ctgt, gr, c1, keys, agg_dev))); break;
            default: LAUNCH(ctx, "geo_aggregate_rem", (k_geo_aggregate_q<4><<<grid, 128, 0, ctx->stream>>>(dref, dtgt, cref, ctgt, gr, c1, keys, agg_dev))); break;
        }
        g.n_cand = c1;
    }
    const int h = g.h, nseg = cdiv(W, GT_X);
    // segments whose window sample columns are clamped BEFORE the disparity shift take the scalar BORDER path:
    // the right image edge for LEFT (x + h > W-1), the left edge for RIGHT (x - h < 0)
    int int_first, int_count;
    if (left) { int_first = 0; int_count = (W - 1 - h - (GT_X - 1) >= 0) ? (W - 1 - h - (GT_X - 1)) / GT_X + 1 : 0; }
    else { int_first = h > 0 ? 1 : 0; int_count = nseg - int_first; }
    if (int_count < 0) int_count = 0;
    if (int_count > nseg - int_first) int_count = nseg - int_first;
    ASW_TRY(geo_segments<false>(ctx, dref, dtgt, cref, ctgt, g, int_first, int_count, keys, agg_dev));
    ASW_TRY(geo_segments<true>(ctx, dref, dtgt, cref, ctgt, g, 0, int_first, keys, agg_dev));
    ASW_TRY(geo_segments<true>(ctx, dref, dtgt, cref, ctgt, g, int_first + int_count, nseg - int_first - int_count, keys, agg_dev));
    return keys_to_disp(ctx, keys, n, disp_dev);
}
