// k_traditional.cuh -- traditional (Yoon-Kweon) bilateral ASW, computeAdaptiveWeight (A.cpp:1016-1156).
//
// Reference behaviour reproduced:
//   * D+1 candidates (offset = min .. min+num inclusive, A.cpp:1021, 1074)
//   * weight list n = 0..win^2-2 skips the centre at BUILD (pw = n < c ? n : n+1, A.cpp:1050-1053) but the
//     aggregation remaps with i > c (pc = n <= c ? n : n+1, A.cpp:1091) and uses pc/win as the X offset,
//     pc%win as the Y offset (A.cpp:1093-1102): weights are applied to the transposed neighbour
//   * w = (float)(wL_n * wR_n); num += (double)w * |dI| ; den += w  (A.cpp:1104-1108)
//   * raw cost is the plain gray absolute difference, clamp addressing
// Weights are exact: the host tabulates (float)(3*exp(-(delta/gamma_c + dist_n/gamma_g))) in double for all
// 256 gray differences and all window offsets, exactly as A.cpp:1062-1066 evaluates them.
// Accumulation: fp32 inside one window row, double across rows (the reference accumulates in double).
#pragma once
#include "k_cost.cuh"

#include <type_traits>
#include <vector>

struct TradGeom {
    int H, W, win, h, nw, cidx;
    int mode;        // 0: computeAdaptiveWeight (A.cpp:1016-1156), 1: computeAdaptiveWeight_direct8 (A.cpp:1167-1319)
    int sign;        // LEFT: target column = max(0, x - d) ; RIGHT: min(x + d, W-1)
    int d_first;     // first candidate offset evaluated by this launch
    int n_cand;      // candidates in this launch (<= TRAD_Q per thread pass)
};

#define TRAD_Q 6

// tap n of the weight list -> (dy, dx) = neighbour whose gray difference weights the tap, (ky, kx) = cost sample offset
// from the window's top-left corner.
//  mode 0: weight list skips the centre (A.cpp:1050-1053); the sample list skips centre + 1 and is TRANSPOSED
//          (pc / win is used as the column offset, A.cpp:1088-1102).
//  mode 1: 8-direction subset (A.cpp:1191-1199, 1239-1247): row-major (j, i) with i == j || i == 0 || j == 0 (the
//          "anti-diagonal" test i + j == win - 1 only ever matches (h, h)): 2 taps per row j != 0, win - 1 in row 0;
//          weight and sample use the same offset.
__host__ __device__ __forceinline__ void trad_tap(int mode, int n, int win, int cidx, int* dy, int* dx, int* ky, int* kx) {
    const int h = win / 2;
    if (mode == 0) {
        const int pw = n < cidx ? n : n + 1, pc = n <= cidx ? n : n + 1;
        *dy = pw / win - h; *dx = pw % win - h;
        *kx = pc / win; *ky = pc % win;
    } else {
        int j, i;
        if (n < 2 * h) { j = -h + n / 2; i = (n & 1) ? 0 : j; }
        else if (n < 4 * h) { const int m = n - 2 * h; j = 0; i = m < h ? m - h : m - h + 1; }
        else { const int m = n - 4 * h; j = 1 + m / 2; i = (m & 1) ? j : 0; }
        *dy = j; *dx = i; *ky = j + h; *kx = i + h;
    }
}

__device__ __forceinline__ int trad_shift(int x, int d, int sign, int W) {
    return sign > 0 ? max(0, x - d) : min(x + d, W - 1);
}

// thread = pixel; blockIdx.z = candidate chunk of TRAD_Q offsets
__global__ void __launch_bounds__(128)
k_trad_aggregate(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, const float* __restrict__ table,
                 TradGeom g, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.W) return;
    const int c0 = blockIdx.z * TRAD_Q;
    const int nq = min(TRAD_Q, g.n_cand - c0);
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    int xs[TRAD_Q];
    int tc[TRAD_Q];          // target-side centre gray
#pragma unroll
    for (int q = 0; q < TRAD_Q; q++) {
        int d = g.d_first + c0 + min(q, nq - 1);
        xs[q] = trad_shift(x, d, g.sign, W);
        tc[q] = tgt[(size_t)y * W + xs[q]];
    }
    const int rc = ref[(size_t)y * W + x];
    double num[TRAD_Q], den[TRAD_Q];
#pragma unroll
    for (int q = 0; q < TRAD_Q; q++) { num[q] = 0; den[q] = 0; }
    int n = 0;
    while (n < g.nw) {
        int n_end = min(n + win, g.nw);
        float fn[TRAD_Q], fd[TRAD_Q];
#pragma unroll
        for (int q = 0; q < TRAD_Q; q++) { fn[q] = 0; fd[q] = 0; }
        for (; n < n_end; n++) {
            int pw = n < g.cidx ? n : n + 1;                 // weight offset (A.cpp:1044-1053)
            int dy = pw / win - h, dx = pw - (pw / win) * win - h;
            int pc = n <= g.cidx ? n : n + 1;                // sample offset (A.cpp:1088-1102), transposed
            int kx = pc / win, ky = pc - kx * win;
            int nx = clampi(x - h + kx, 0, W - 1), ny = clampi(y - h + ky, 0, H - 1);
            int wy = clampi(y + dy, 0, H - 1);
            const float* trow = table + (size_t)n * 256;
            int rnb = ref[(size_t)wy * W + clampi(x + dx, 0, W - 1)];
            float wl = __ldg(&trow[abs(rnb - rc)]);
            int rs = ref[(size_t)ny * W + nx];
#pragma unroll
            for (int q = 0; q < TRAD_Q; q++) {
                int d = g.d_first + c0 + min(q, nq - 1);
                int tnb = tgt[(size_t)wy * W + clampi(xs[q] + dx, 0, W - 1)];
                float wr = __ldg(&trow[abs(tnb - tc[q])]);
                float w = __fmul_rn(wl, wr);
                int ts = tgt[(size_t)ny * W + trad_shift(nx, d, g.sign, W)];
                fn[q] = fmaf(w, (float)abs(rs - ts), fn[q]);
                fd[q] = __fadd_rn(fd[q], w);
            }
        }
#pragma unroll
        for (int q = 0; q < TRAD_Q; q++) { num[q] += (double)fn[q]; den[q] += (double)fd[q]; }
    }
    unsigned long long best = WTA_KEY_EMPTY;
    size_t p = (size_t)y * W + x;
#pragma unroll
    for (int q = 0; q < TRAD_Q; q++) {
        if (q < nq) {
            double E = num[q] / den[q];
            int ci = c0 + q;
            if (agg) agg[(size_t)ci * H * W + p] = (float)E;
            best = min(best, wta_key_d(E, g.d_first + ci));
        }
    }
    atomicMin(&keys[p], best);
}

// ------------------------------------------------------------------------------------------------
// tiled variant: 32x8 pixel tile, gray tiles (+ window halo, + the chunk's disparity span) in shared memory
// both as u8 (weight-table index) and as float (cost samples), the exact weight-table rows of one window row
// staged in shared memory per group of `win` taps.  Thread = pixel, TR_Q candidates in registers.
// ------------------------------------------------------------------------------------------------
// Tile 32 x 4 pixels, 17 candidates per thread, 4 CTAs of 128 threads per SM (128 registers, no spills).  17 candidates = the
// D + 1 = 17 of config 1 in one chunk, and the per-tap work that does not depend on the candidate (tap record, reference
// exponent, reference sample) is shared by 17 evaluations instead of 9.  Measured (trad_aggregate, ms; 384x288x16 win 35 /
// 640x360x64 win 15 / 1280x720x64 win 35): 32x8, Q 9, 3 CTAs (round 1-2) 0.96 / 1.31 / 26.3; 32x8, Q 17, 2 CTAs 0.92 / 1.19 / 22.8;
// 32x4, Q 17, 4 CTAs 0.88 / 1.17 / 22.8; 32x4, Q 22, 3 CTAs 1.13 / 1.14 / 22.1; 32x4, Q 13 1.50 / 1.20 / 23.7; 32x2, Q 17 1.03 / 1.20 / 25.7
#define TR_TW 32
#ifndef TR_TH
#define TR_TH 4
#endif
#ifndef TR_Q
#define TR_Q 17
#endif
#ifndef TR_MINB
#define TR_MINB 4
#endif

template <bool INTERIOR>
__device__ __forceinline__ void trad_tile_body(const uint8_t* __restrict__ refU, const float* __restrict__ refF,
                                               const uint8_t* __restrict__ tgtU, const float* __restrict__ tgtF,
                                               float* __restrict__ tab, const float* __restrict__ table,
                                               const TradGeom& g, int IWr, int IWt, int oy, int oxr, int oxt,
                                               int x, int y, int nq, int d_lo, double* num, double* den) {
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const int tid = threadIdx.y * TR_TW + threadIdx.x;
    const int yc = min(y, H - 1), xc = min(x, W - 1);             // threads past the image edge compute a clamped copy
    int xs[TR_Q], tc[TR_Q];
#pragma unroll
    for (int q = 0; q < TR_Q; q++) {
        int d = d_lo + min(q, nq - 1);
        xs[q] = trad_shift(xc, d, g.sign, W);
        tc[q] = tgtU[(yc - oy) * IWt + xs[q] - oxt];
    }
    const int rc = refU[(yc - oy) * IWr + xc - oxr];
    for (int n0 = 0; n0 < g.nw; n0 += win) {
        const int n1 = min(n0 + win, g.nw);
        __syncthreads();                                          // previous group's table rows are dead
        for (int i = tid; i < (n1 - n0) * 256; i += TR_TW * TR_TH) tab[i] = __ldg(&table[(size_t)n0 * 256 + i]);
        __syncthreads();
        float fn[TR_Q], fd[TR_Q];
#pragma unroll
        for (int q = 0; q < TR_Q; q++) { fn[q] = 0.0f; fd[q] = 0.0f; }
        for (int n = n0; n < n1; n++) {
            int pw = n < g.cidx ? n : n + 1;                      // weight offset (A.cpp:1044-1053)
            int dy = pw / win - h, dx = pw - (pw / win) * win - h;
            int pc = n <= g.cidx ? n : n + 1;                     // sample offset (A.cpp:1088-1102), transposed
            int kx = pc / win, ky = pc - kx * win;
            const float* trow = tab + (n - n0) * 256;
            int wy, ny, nx, rnb_col;
            if (INTERIOR) { wy = yc + dy; ny = yc - h + ky; nx = xc - h + kx; rnb_col = xc + dx; }
            else {
                wy = clampi(yc + dy, 0, H - 1); ny = clampi(yc - h + ky, 0, H - 1);
                nx = clampi(xc - h + kx, 0, W - 1); rnb_col = clampi(xc + dx, 0, W - 1);
            }
            const float wl = trow[abs((int)refU[(wy - oy) * IWr + rnb_col - oxr] - rc)];
            const float rs = refF[(ny - oy) * IWr + nx - oxr];
            const uint8_t* tu = tgtU + (wy - oy) * IWt - oxt;
            const float* tf = tgtF + (ny - oy) * IWt - oxt;
#pragma unroll
            for (int q = 0; q < TR_Q; q++) {
                int cn, cs;
                if (INTERIOR) { cn = xs[q] + dx; cs = nx - g.sign * (d_lo + min(q, nq - 1)); }
                else { cn = clampi(xs[q] + dx, 0, W - 1); cs = trad_shift(nx, d_lo + min(q, nq - 1), g.sign, W); }
                float wr = trow[abs((int)tu[cn] - tc[q])];
                float w = __fmul_rn(wl, wr);
                fn[q] = fmaf(w, fabsf(rs - tf[cs]), fn[q]);
                fd[q] = __fadd_rn(fd[q], w);
            }
        }
#pragma unroll
        for (int q = 0; q < TR_Q; q++) { num[q] += (double)fn[q]; den[q] += (double)fd[q]; }
    }
}

#ifdef ASW_DEV_KERNELS
// grid: (tiles_x, tiles_y, candidate chunks of TR_Q); block (32, 8)
__global__ void __launch_bounds__(TR_TW * TR_TH)
k_trad_tile(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, const float* __restrict__ table,
            TradGeom g, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    extern __shared__ float sm_tr[];
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const int IH = TR_TH + 2 * h, IWr = TR_TW + 2 * h, IWt = TR_TW + 2 * h + TR_Q - 1;
    float* tab = sm_tr;                                           // [win][256]
    float* refF = tab + win * 256;                                // [IH][IWr]
    float* tgtF = refF + IH * IWr;                                // [IH][IWt]
    uint8_t* refU = (uint8_t*)(tgtF + IH * IWt);                  // [IH][IWr]
    uint8_t* tgtU = refU + IH * IWr;                              // [IH][IWt]
    const int x0t = blockIdx.x * TR_TW, y0t = blockIdx.y * TR_TH;
    const int c0 = blockIdx.z * TR_Q;
    const int nq = min(TR_Q, g.n_cand - c0);
    const int d_lo = g.d_first + c0, d_hi = d_lo + nq - 1;
    // tile origins in absolute (already clamped) image coordinates; cell (r, c) <-> (min(oy+r, H-1), min(ox+c, W-1))
    const int oy = max(0, y0t - h), oxr = max(0, x0t - h);
    const int oxt = g.sign > 0 ? max(0, x0t - d_hi - h) : max(0, min(x0t + d_lo, W - 1) - h);
    const int tid = threadIdx.y * TR_TW + threadIdx.x;
    for (int i = tid; i < IH * IWr; i += TR_TW * TR_TH) {
        int r = i / IWr, c = i - r * IWr;
        uint8_t v = ref[(size_t)min(oy + r, H - 1) * W + min(oxr + c, W - 1)];
        refU[i] = v; refF[i] = (float)v;
    }
    for (int i = tid; i < IH * IWt; i += TR_TW * TR_TH) {
        int r = i / IWt, c = i - r * IWt;
        uint8_t v = tgt[(size_t)min(oy + r, H - 1) * W + min(oxt + c, W - 1)];
        tgtU[i] = v; tgtF[i] = (float)v;
    }
    __syncthreads();
    const int x = x0t + threadIdx.x, y = y0t + threadIdx.y;
    double num[TR_Q], den[TR_Q];
#pragma unroll
    for (int q = 0; q < TR_Q; q++) { num[q] = 0; den[q] = 0; }
    // a tile is interior when no coordinate it touches is clamped: then all clamps are identities
    const bool interior = (y0t - h >= 0) && (y0t + TR_TH - 1 + h <= H - 1) && (x0t - h >= 0) && (x0t + TR_TW - 1 + h <= W - 1) &&
                          (g.sign > 0 ? (x0t - d_hi - h >= 0) : (x0t + TR_TW - 1 + d_hi + h <= W - 1));
    if (interior) trad_tile_body<true>(refU, refF, tgtU, tgtF, tab, table, g, IWr, IWt, oy, oxr, oxt, x, y, nq, d_lo, num, den);
    else trad_tile_body<false>(refU, refF, tgtU, tgtF, tab, table, g, IWr, IWt, oy, oxr, oxt, x, y, nq, d_lo, num, den);
    if (x < W && y < H) {
        unsigned long long best = WTA_KEY_EMPTY;
        size_t p = (size_t)y * W + x;
#pragma unroll
        for (int q = 0; q < TR_Q; q++) {
            if (q < nq) {
                double E = num[q] / den[q];
                if (agg) agg[(size_t)(c0 + q) * H * W + p] = (float)E;
                best = min(best, wta_key_d(E, d_lo + q));
            }
        }
        atomicMin(&keys[p], best);
    }
}
#endif

// ------------------------------------------------------------------------------------------------
// default variant: same tiling, but the weight product is evaluated in one SFU op instead of two
// (bank-conflicted) table look-ups:
//   wL_n * wR_n = 9 * exp(-(dL + dR)/gamma_c - 2 g_n/gamma_g) = 2^(-(dL + dR) * a2 - c2[n])
// with a2 = log2(e)/gamma_c and c2[n] = 2 g_n log2(e)/gamma_g - log2(9) tabulated on the host in double.
// ex2.approx + the fp32 exponent give weights within ~1e-6 relative of the reference's double-precision
// exp rounded to float (two orders of magnitude inside the 1e-4 cost budget); the exact-table kernel above
// stays selectable (ASW_TRAD_EXACT=1).  Only float tiles are needed (|dI| and the sample cost are float ops).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// SIGN = +1 (DISPARITY_LEFT: target column x - d) or -1 (RIGHT: x + d).  In interior tiles no coordinate is
// clamped, so the candidate q of a tap sits at a compile-time offset (-SIGN*q) from a per-tap base pointer:
// the inner loop is 2 LDS (immediate offsets) + 6 FP instructions per (tap, candidate).
template <bool INTERIOR, int SIGN>
__device__ __forceinline__ void trad_fast_body(const float* __restrict__ refF, const float* __restrict__ tgtF,
                                               const float* __restrict__ c2, float a2, const TradGeom& g, int IWr, int IWt,
                                               int oy, int oxr, int oxt, int x, int y, int nq, int d_lo,
                                               double* num, double* den) {
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const int yc = min(y, H - 1), xc = min(x, W - 1);
    int xs[TR_Q];
    float tc[TR_Q];
#pragma unroll
    for (int q = 0; q < TR_Q; q++) {
        xs[q] = INTERIOR ? xc - SIGN * (d_lo + q) : trad_shift(xc, d_lo + min(q, nq - 1), SIGN, W);
        tc[q] = tgtF[(yc - oy) * IWt + xs[q] - oxt];
    }
    const float rc = refF[(yc - oy) * IWr + xc - oxr];
    // running window coordinates: (wj, wi) = weight tap pw = n (+1 past the centre, A.cpp:1050-1053),
    // (sj, si) = sample tap pc = n (+1 past n > centre, A.cpp:1091) -- no integer division in the loop
    int wj = 0, wi = 0, sj = 0, si = 0;
    int n = 0;
    while (n < g.nw) {
        const int n1 = min(n + win, g.nw);
        float fn[TR_Q], fd[TR_Q];
#pragma unroll
        for (int q = 0; q < TR_Q; q++) { fn[q] = 0.0f; fd[q] = 0.0f; }
        for (; n < n1; n++) {
            const int dy = wj - h, dx = wi - h;                   // weight offset
            const int kx = sj, ky = si;                           // sample offset, transposed (A.cpp:1093-1102)
            int wy, ny, nx, rnb_col;
            if (INTERIOR) { wy = yc + dy; ny = yc - h + ky; nx = xc - h + kx; rnb_col = xc + dx; }
            else {
                wy = clampi(yc + dy, 0, H - 1); ny = clampi(yc - h + ky, 0, H - 1);
                nx = clampi(xc - h + kx, 0, W - 1); rnb_col = clampi(xc + dx, 0, W - 1);
            }
            // e0 = -(dL * a2 + c2[n]); per candidate: w = 2^(e0 - dR * a2)
            const float e0 = -fmaf(fabsf(refF[(wy - oy) * IWr + rnb_col - oxr] - rc), a2, __ldg(&c2[n]));
            const float rs = refF[(ny - oy) * IWr + nx - oxr];
            const float* tu = tgtF + (wy - oy) * IWt - oxt;
            const float* tf = tgtF + (ny - oy) * IWt - oxt;
            if (INTERIOR) {
                const float* pu = tu + xc + dx - SIGN * d_lo;
                const float* pf = tf + nx - SIGN * d_lo;
#pragma unroll
                for (int q = 0; q < TR_Q; q++) {
                    float w = fast_ex2(fmaf(-fabsf(pu[-SIGN * q] - tc[q]), a2, e0));
                    fn[q] = fmaf(w, fabsf(rs - pf[-SIGN * q]), fn[q]);
                    fd[q] = __fadd_rn(fd[q], w);
                }
            } else {
#pragma unroll
                for (int q = 0; q < TR_Q; q++) {
                    int cn = clampi(xs[q] + dx, 0, W - 1), cs = trad_shift(nx, d_lo + min(q, nq - 1), SIGN, W);
                    float w = fast_ex2(fmaf(-fabsf(tu[cn] - tc[q]), a2, e0));
                    fn[q] = fmaf(w, fabsf(rs - tf[cs]), fn[q]);
                    fd[q] = __fadd_rn(fd[q], w);
                }
            }
            // advance both tap counters by one, and once more where the reference's remapping skips a position
            int adv_w = (n + 1 == g.cidx) ? 2 : 1;                // pw jumps over the centre
            int adv_s = (n == g.cidx) ? 2 : 1;                    // pc jumps over position centre + 1
            wi += adv_w; if (wi >= win) { wi -= win; wj++; }
            si += adv_s; if (si >= win) { si -= win; sj++; }
        }
#pragma unroll
        for (int q = 0; q < TR_Q; q++) { num[q] += (double)fn[q]; den[q] += (double)fd[q]; }
    }
}

#ifdef ASW_DEV_KERNELS
__global__ void __launch_bounds__(TR_TW * TR_TH)
k_trad_fast(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, const float* __restrict__ c2, float a2,
            TradGeom g, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    extern __shared__ float sm_tr[];
    const int W = g.W, H = g.H, h = g.h;
    const int IH = TR_TH + 2 * h, IWr = TR_TW + 2 * h, IWt = TR_TW + 2 * h + TR_Q - 1;
    float* refF = sm_tr;                                          // [IH][IWr]
    float* tgtF = refF + IH * IWr;                                // [IH][IWt]
    const int x0t = blockIdx.x * TR_TW, y0t = blockIdx.y * TR_TH;
    const int c0 = blockIdx.z * TR_Q;
    const int nq = min(TR_Q, g.n_cand - c0);
    const int d_lo = g.d_first + c0, d_hi = d_lo + nq - 1;
    const int oy = max(0, y0t - h), oxr = max(0, x0t - h);
    const int oxt = g.sign > 0 ? max(0, x0t - d_hi - h) : max(0, min(x0t + d_lo, W - 1) - h);
    const int tid = threadIdx.y * TR_TW + threadIdx.x;
    for (int i = tid; i < IH * IWr; i += TR_TW * TR_TH) {
        int r = i / IWr, c = i - r * IWr;
        refF[i] = (float)ref[(size_t)min(oy + r, H - 1) * W + min(oxr + c, W - 1)];
    }
    for (int i = tid; i < IH * IWt; i += TR_TW * TR_TH) {
        int r = i / IWt, c = i - r * IWt;
        tgtF[i] = (float)tgt[(size_t)min(oy + r, H - 1) * W + min(oxt + c, W - 1)];
    }
    __syncthreads();
    const int x = x0t + threadIdx.x, y = y0t + threadIdx.y;
    double num[TR_Q], den[TR_Q];
#pragma unroll
    for (int q = 0; q < TR_Q; q++) { num[q] = 0; den[q] = 0; }
    const bool interior = (y0t - h >= 0) && (y0t + TR_TH - 1 + h <= H - 1) && (x0t - h >= 0) && (x0t + TR_TW - 1 + h <= W - 1) &&
                          (g.sign > 0 ? (x0t - d_hi - h >= 0) : (x0t + TR_TW - 1 + d_hi + h <= W - 1));
    if (g.sign > 0) {
        if (interior) trad_fast_body<true, 1>(refF, tgtF, c2, a2, g, IWr, IWt, oy, oxr, oxt, x, y, nq, d_lo, num, den);
        else trad_fast_body<false, 1>(refF, tgtF, c2, a2, g, IWr, IWt, oy, oxr, oxt, x, y, nq, d_lo, num, den);
    } else {
        if (interior) trad_fast_body<true, -1>(refF, tgtF, c2, a2, g, IWr, IWt, oy, oxr, oxt, x, y, nq, d_lo, num, den);
        else trad_fast_body<false, -1>(refF, tgtF, c2, a2, g, IWr, IWt, oy, oxr, oxt, x, y, nq, d_lo, num, den);
    }
    if (x < W && y < H) {
        unsigned long long best = WTA_KEY_EMPTY;
        size_t p = (size_t)y * W + x;
#pragma unroll
        for (int q = 0; q < TR_Q; q++) {
            if (q < nq) {
                double E = num[q] / den[q];
                if (agg) agg[(size_t)(c0 + q) * H * W + p] = (float)E;
                best = min(best, wta_key_d(E, d_lo + q));
            }
        }
        atomicMin(&keys[p], best);
    }
}
#endif

// ------------------------------------------------------------------------------------------------
// linear-addressing variant of the single-SFU-op kernel (the default).  Two observations remove every clamp
// from the inner loop of ALL tiles, image borders included:
//   * the reference clamps coordinates (A.cpp:1050-1060, 1093-1102); a tile that is STAGED through the clamp
//     (cell (r, c) <-> pixel (clamp(y0 - h + r), clamp(x0 - h + c))) can be read with unclamped offsets;
//   * the cost sample column is clamped BEFORE the candidate shift (nx = clamp(x - h + kx), then nx -/+ d,
//     A.cpp:1093-1102): that is one min/max per TAP, after which the TR_Q candidates of the tap sit at
//     compile-time offsets -SIGN*q from one pointer; the clamp after the shift is again the staging clamp.
// Only where the candidate shift itself clamps the weighted pixel (x - d < 0 / x + d > W-1: the first or last
// tile column) the target-weight operand needs one base pointer per candidate (XS_CLAMP).
// The tap geometry (centre skip of the weight list, centre+1 skip and transposition of the sample list,
// A.cpp:1050-1053, 1091-1102) is tabulated once per CTA in shared memory: 2 broadcast loads per tap instead of
// ~45 integer instructions.
// ------------------------------------------------------------------------------------------------
template <int SIGN, bool XS_CLAMP>
__device__ __forceinline__ void trad_lin_body(const float* __restrict__ RF, const float* __restrict__ TF,
                                              const int4* __restrict__ taps, const float* __restrict__ c2s, float a2,
                                              const TradGeom& g, int IWr, int IWt, int x0t, int y0t, int oxt, int x, int y,
                                              int d_lo, float* num, float* den) {
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const int yc = min(y, H - 1), xc = min(x, W - 1);             // threads past the image edge compute a clamped copy
    const int rowc = yc - y0t + h, colc = xc - x0t + h;           // centre cell in the reference tile
    const float* cw = RF + rowc * IWr + colc;                     // weight neighbours: cw[dy * IWr + dx]
    const float* cs = RF + (rowc - h) * IWr + (colc - h);         // cost samples:      cs[ky * IWr + kx]
    const float rc = cw[0];
    const float* trow = TF + rowc * IWt - oxt;                    // target tile row of the pixel, indexed by image column
    const float* tq[XS_CLAMP ? TR_Q : 1];                         // weighted target pixel of candidate q
    float tc[TR_Q];
#pragma unroll
    for (int q = 0; q < TR_Q; q++) {
        const int xs = XS_CLAMP ? trad_shift(xc, d_lo + q, SIGN, W) : xc - SIGN * (d_lo + q);
        if (XS_CLAMP) tq[q] = trow + xs; else if (q == 0) tq[0] = trow + xs;
        tc[q] = trow[xs];
    }
    const float* ts = TF + (rowc - h) * IWt - oxt - SIGN * d_lo;  // cost samples of the target: ts[ky * IWt + nx - SIGN * q]
    const int xl = xc - h;
    int n = 0;
    while (n < g.nw) {
        const int n1 = min(n + win, g.nw);
        float fn[TR_Q], fd[TR_Q];
#pragma unroll
        for (int q = 0; q < TR_Q; q++) { fn[q] = 0.0f; fd[q] = 0.0f; }
#pragma unroll 2
        for (; n < n1; n++) {
            const int4 tp = taps[n];                              // {dy*IWr+dx, dy*IWt+dx, ky*IWr+kx, ky*IWt | kx << 20}
            // e0 = -(dL * a2 + c2[n]); per candidate: w = 2^(e0 - dR * a2)
            const float e0 = -fmaf(fabsf(cw[tp.x] - rc), a2, c2s[n]);
            const float rs = cs[tp.z];
            const int kx = tp.w >> 20;
            const int nx = SIGN > 0 ? min(xl + kx, W - 1) : max(xl + kx, 0);      // clamp BEFORE the shift
            const float* pf = ts + (tp.w & 0xFFFFF) + nx;
            if (!XS_CLAMP) {
                const float* pu = tq[0] + tp.y;
#pragma unroll
                for (int q = 0; q < TR_Q; q++) {
                    float w = fast_ex2(fmaf(-fabsf(pu[-SIGN * q] - tc[q]), a2, e0));
                    fn[q] = fmaf(w, fabsf(rs - pf[-SIGN * q]), fn[q]);
                    fd[q] = __fadd_rn(fd[q], w);
                }
            } else {
                // first / last tile columns: the shift clamps, so the sample column is clamped per candidate as well
                const int nxc = clampi(xl + kx, 0, W - 1);
                const float* pr = TF + (rowc - h) * IWt - oxt + (tp.w & 0xFFFFF);
#pragma unroll
                for (int q = 0; q < TR_Q; q++) {
                    float w = fast_ex2(fmaf(-fabsf(tq[XS_CLAMP ? q : 0][tp.y] - tc[q]), a2, e0));
                    fn[q] = fmaf(w, fabsf(rs - pr[trad_shift(nxc, d_lo + q, SIGN, W)]), fn[q]);
                    fd[q] = __fadd_rn(fd[q], w);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < TR_Q; q++) { num[q] = __fadd_rn(num[q], fn[q]); den[q] = __fadd_rn(den[q], fd[q]); }
    }
}

template <int SIGN>
__global__ void __launch_bounds__(TR_TW * TR_TH, TR_MINB)
k_trad_lin(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, const float* __restrict__ c2, float a2,
           TradGeom g, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    extern __shared__ __align__(16) float sm_tr[];
    const int W = g.W, H = g.H, h = g.h, win = g.win;
    const int IH = TR_TH + 2 * h, IWr = TR_TW + 2 * h, IWt = TR_TW + 2 * h + TR_Q - 1;
    int4* taps = (int4*)sm_tr;                                    // [nw]
    float* c2s = sm_tr + 4 * g.nw;                                // [nw]
    float* RF = c2s + g.nw;                                       // [IH][IWr]  Lg(clamp(y0t-h+r), clamp(x0t-h+c))
    float* TF = RF + IH * IWr;                                    // [IH][IWt]  Rg(clamp(y0t-h+r), clamp(oxt+c))
    const int x0t = blockIdx.x * TR_TW, y0t = blockIdx.y * TR_TH;
    const int c0 = blockIdx.z * TR_Q;
    const int nq = min(TR_Q, g.n_cand - c0);
    const int d_lo = g.d_first + c0, d_hi = d_lo + TR_Q - 1;      // the tile always spans TR_Q candidates
    // image column of target cell 0: the leftmost weighted target pixel of the tile (after the shift clamp) minus h
    const int oxt = SIGN > 0 ? max(0, x0t - d_hi) - h : min(x0t + d_lo, W - 1) - h;
    const int tid = threadIdx.y * TR_TW + threadIdx.x;
    for (int i = tid; i < IH * IWr; i += TR_TW * TR_TH) {
        int r = i / IWr, c = i - r * IWr;
        RF[i] = (float)ref[(size_t)clampi(y0t - h + r, 0, H - 1) * W + clampi(x0t - h + c, 0, W - 1)];
    }
    for (int i = tid; i < IH * IWt; i += TR_TW * TR_TH) {
        int r = i / IWt, c = i - r * IWt;
        TF[i] = (float)tgt[(size_t)clampi(y0t - h + r, 0, H - 1) * W + clampi(oxt + c, 0, W - 1)];
    }
    for (int n = tid; n < g.nw; n += TR_TW * TR_TH) {
        int dy, dx, ky, kx;
        trad_tap(g.mode, n, win, g.cidx, &dy, &dx, &ky, &kx);
        taps[n] = make_int4(dy * IWr + dx, dy * IWt + dx, ky * IWr + kx, (ky * IWt) | (kx << 20));
        c2s[n] = __ldg(&c2[n]);
    }
    __syncthreads();
    const int x = x0t + threadIdx.x, y = y0t + threadIdx.y;
    // fp32 over a window row, fp32 across rows (all terms are non-negative: ~1e-6 relative against the reference's double)
    float num[TR_Q], den[TR_Q];
#pragma unroll
    for (int q = 0; q < TR_Q; q++) { num[q] = 0.0f; den[q] = 0.0f; }
    // does the candidate shift clamp the weighted pixel anywhere in this tile?
    const bool xs_clamp = SIGN > 0 ? (x0t - d_hi < 0) : (x0t + TR_TW - 1 + d_hi > W - 1);
    if (xs_clamp) trad_lin_body<SIGN, true>(RF, TF, taps, c2s, a2, g, IWr, IWt, x0t, y0t, oxt, x, y, d_lo, num, den);
    else trad_lin_body<SIGN, false>(RF, TF, taps, c2s, a2, g, IWr, IWt, x0t, y0t, oxt, x, y, d_lo, num, den);
    if (x < W && y < H) {
        unsigned long long best = WTA_KEY_EMPTY;
        size_t p = (size_t)y * W + x;
#pragma unroll
        for (int q = 0; q < TR_Q; q++) {
            if (q < nq) {
                double E = (double)num[q] / (double)den[q];
                if (agg) agg[(size_t)(c0 + q) * H * W + p] = (float)E;
                best = min(best, wta_key_d(E, d_lo + q));
            }
        }
        atomicMin(&keys[p], best);
    }
}

// ------------------------------------------------------------------------------------------------
// diagonal-blocked variant (ASW_TRAD_DIAG=1; not the default -- see dev_traditional): one CTA = one image row x 128 pixels x 4 NW candidates, thread = 4 adjacent
// pixels x 4 ADJACENT candidates, as k_geo_agg_diag (k_geodesic.cuh):
//   * the bilateral weights are evaluated ONCE per pixel and tap -- wL_n for the 128 reference pixels, wR_n for the
//     128 + 4 NW + 4 target pixels the chunk can reach (one ex2 each) -- into shared memory, one 12-tap chunk ahead,
//     and reused by every candidate: 280 SFU operations per tap instead of 128 x (candidates)
//   * operands addressed at (x - d) sit on a diagonal of the thread's 4 x 4 block: the 7 distinct target weights
//     are one aligned 8-wide window (2 LDS.128); the target cost samples 7 scalars
//   * per evaluation: |dI| (FADD), w = wL wR (FMUL2), num += w |dI| (FFMA), den += w (FADD2)
// Float tiles of the gray rows y-h .. y+h hold the cost samples (clamp addressing applied when they are staged).
// BORDER segments (sample column clamped BEFORE the shift: right edge for LEFT, left edge for RIGHT) read the
// shifted edge column, which depends on the candidate only.
// Accumulation is fp32 over the window (all terms are non-negative; ~1e-6 relative against the reference's double).
// ------------------------------------------------------------------------------------------------
#define TD_X 128
#define TD_TC 12

template <int SIGN, bool BORDER, int NW>
__device__ __forceinline__ void trad_diag_body(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt,
                                               const float* __restrict__ c2h, float a2, const TradGeom& g, int xb, int c0,
                                               unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    constexpr int NT = 32 * NW;
    constexpr int WRW = TD_X + 4 * NW + 4;                     // staged target-weight row (cells)
    extern __shared__ __align__(16) float sm_td[];
    const int W = g.W, H = g.H, win = g.win, h = g.h;
    const int RTW = (TD_X + 2 * h + 7) & ~3, TTW = (WRW + 2 * h + 7) & ~3;   // sample tiles (rows 16-byte aligned, 8 / 12-wide windows)
    float* WLs = sm_td;                                        // [2][TD_TC][TD_X]
    float* WRs = WLs + 2 * TD_TC * TD_X;                       // [2][TD_TC][WRW]
    float* RT = WRs + 2 * TD_TC * WRW;                         // [win][RTW]  Lg(clamp(y-h+r), clamp(xb-h+c))
    float* TT = RT + win * RTW;                                // [win][TTW]  Rg(clamp(y-h+r), clamp(oWR-h+c))
    float* C2 = TT + win * TTW;                                // [nw]        spatial exponent of every tap
    int* TAPW = (int*)(C2 + g.nw);                             // [nw]        weight tap: (dy + h) << 16 | (dx + h)
    int* TAPS = TAPW + g.nw;                                   // [nw]        sample tap: ky << 16 | kx   (transposed)
    const int tid = threadIdx.x, pg = tid & 31, ds = tid >> 5;
    const int y = blockIdx.y;
    const int d_lo = g.d_first + c0;
    // staged target cell 0 <-> column (before clamping); a thread's 8-wide window starts at cell e0 (multiple of 4)
    const int oWR = SIGN > 0 ? xb - d_lo - 4 * NW : xb + d_lo;
    const int e0 = SIGN > 0 ? 4 * pg + 4 * (NW - 1 - ds) : 4 * pg + 4 * ds;
    const size_t rowoff = (size_t)y * W;
    // ---- sample tiles, tap tables ----
    for (int i = tid; i < win * RTW; i += NT) {
        const int r = i / RTW, c = i - r * RTW;
        RT[i] = (float)ref[(size_t)clampi(y - h + r, 0, H - 1) * W + clampi(xb - h + c, 0, W - 1)];
    }
    for (int i = tid; i < win * TTW; i += NT) {
        const int r = i / TTW, c = i - r * TTW;
        TT[i] = (float)tgt[(size_t)clampi(y - h + r, 0, H - 1) * W + clampi(oWR - h + c, 0, W - 1)];
    }
    for (int n = tid; n < g.nw; n += NT) {
        const int pw = n < g.cidx ? n : n + 1;                 // weight tap skips the centre (A.cpp:1044-1053)
        const int pc = n <= g.cidx ? n : n + 1;                // sample tap skips centre + 1 (A.cpp:1088-1102)
        C2[n] = __ldg(&c2h[n]);
        TAPW[n] = ((pw / win) << 16) | (pw % win);
        TAPS[n] = ((pc % win) << 16) | (pc / win);             // pc / win is the COLUMN offset, pc % win the ROW offset
    }
    __syncthreads();
    // ---- weights of a 12-tap chunk: thread `tid` owns column `tid` (+ NT ...) of the reference and target rows.
    // The weighted pixel and its neighbour come from the sample tiles (same clamp addressing); target cells that lie
    // outside the image stand for the clamped edge pixel, whose neighbourhood the tile may not hold: global loads.
    auto stage_weights = [&](int n0, int buf) {
        const int cnt = min(TD_TC, g.nw - n0);
        for (int col = tid; col < TD_X + WRW; col += NT) {
            const bool is_ref = col < TD_X;
            const int cell = is_ref ? col : col - TD_X;
            const int xu = is_ref ? xb + cell : oWR + cell;                        // the weighted pixel, before clamping
            float* dst = is_ref ? WLs + buf * TD_TC * TD_X + cell : WRs + buf * TD_TC * WRW + cell;
            const int pitch = is_ref ? TD_X : WRW;
            if (xu >= 0 && xu <= W - 1) {
                const float* tile = (is_ref ? RT : TT) + cell;                     // tile column of (xu - h)
                const int tw = is_ref ? RTW : TTW;
                const float centre = tile[h * tw + h];
                for (int t = 0; t < cnt; t++) {
                    const int tp = TAPW[n0 + t];
                    const float nb = tile[(tp >> 16) * tw + (tp & 0xFFFF)];
                    // 3 exp(-(delta/gamma_c + dist/gamma_g)) = 2^(-(delta a2 + c2h[n]))   (A.cpp:1062-1066)
                    dst[t * pitch] = fast_ex2(-fmaf(fabsf(nb - centre), a2, C2[n0 + t]));
                }
            } else {
                const uint8_t* img = is_ref ? ref : tgt;
                const int xc = clampi(xu, 0, W - 1);
                const float centre = (float)img[rowoff + xc];
                for (int t = 0; t < cnt; t++) {
                    const int tp = TAPW[n0 + t];
                    const int dy = (tp >> 16) - h, dx = (tp & 0xFFFF) - h;
                    const float nb = (float)img[(size_t)clampi(y + dy, 0, H - 1) * W + clampi(xc + dx, 0, W - 1)];
                    dst[t * pitch] = fast_ex2(-fmaf(fabsf(nb - centre), a2, C2[n0 + t]));
                }
            }
        }
    };
    float2 fnp[2][3], fdp[2][3];
    float fns[2][2], fds[2][2];
#pragma unroll
    for (int a = 0; a < 2; a++) {
#pragma unroll
        for (int q = 0; q < 3; q++) { fnp[a][q] = make_float2(0.f, 0.f); fdp[a][q] = make_float2(0.f, 0.f); }
        fns[a][0] = fns[a][1] = fds[a][0] = fds[a][1] = 0.0f;
    }
    stage_weights(0, 0);
    __syncthreads();
    const int nchunk = (g.nw + TD_TC - 1) / TD_TC;
    for (int c = 0; c < nchunk; c++) {
        const int n0 = c * TD_TC, cnt = min(TD_TC, g.nw - n0);
        if (c + 1 < nchunk) stage_weights(n0 + TD_TC, (c + 1) & 1);      // next chunk's weights into the other buffer
        const float* wl = WLs + (c & 1) * TD_TC * TD_X + 4 * pg;
        const float* wr = WRs + (c & 1) * TD_TC * WRW + e0;
        for (int t = 0; t < cnt; t++) {
            const int tp = TAPS[n0 + t];
            const int ky = tp >> 16, kx = tp & 0xFFFF;
            const float4 wl4 = *(const float4*)(wl + t * TD_X);
            const float4 wra = *(const float4*)(wr + t * WRW), wrb = *(const float4*)(wr + t * WRW + 4);
            const float wlv[4] = {wl4.x, wl4.y, wl4.z, wl4.w};
            const float wrv[8] = {wra.x, wra.y, wra.z, wra.w, wrb.x, wrb.y, wrb.z, wrb.w};
            // cost samples: a thread's 4 reference and 7 target samples start at column (4 pg + kx): 16-byte aligned
            // loads from the aligned-down column, the (warp-uniform) remainder kx & 3 picks the registers
            const float4* rs4 = (const float4*)(RT + ky * RTW + 4 * pg + (kx & ~3));
            const float4* ts4 = (const float4*)(TT + ky * TTW + e0 + (kx & ~3));
            const float4 r0 = rs4[0], r1 = rs4[1], t0 = ts4[0], t1 = ts4[1], t2 = ts4[2];
            const float rsw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            const float tsw[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
            float edge[4];
            bool clamped[4];
            if (BORDER) {
                const int ny = clampi(y - h + ky, 0, H - 1);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int d = d_lo + 4 * ds + k;
                    edge[k] = (float)__ldg(tgt + (size_t)ny * W + (SIGN > 0 ? max(0, W - 1 - d) : min(d, W - 1)));
                }
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const int nx = xb + 4 * pg + p - h + kx;              // sample column before clamping
                    clamped[p] = SIGN > 0 ? nx > W - 1 : nx < 0;
                }
            }
            auto tap = [&](auto sc) {
                constexpr int S = decltype(sc)::value;
                auto cost = [&](int p, int k, int w) {
                    const float tv = BORDER ? (clamped[p] ? edge[k] : tsw[S + w]) : tsw[S + w];
                    return fabsf(rsw[S + p] - tv);                        // |Lg - Rg| (A.cpp:1104)
                };
#pragma unroll
                for (int a = 0; a < 2; a++) {
                    const int pe = 2 * a;
                    const float2 wl2 = make_float2(wlv[pe], wlv[pe + 1]);
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        const int k = SIGN > 0 ? q : q + 1;               // (pe, k) and (pe+1, k+SIGN): same target pixel
                        const int w = SIGN > 0 ? pe - k + 4 : pe + k;
                        const float2 w2 = __fmul2_rn(wl2, make_float2(wrv[w], wrv[w]));       // wL * wR
                        fnp[a][q].x = fmaf(w2.x, cost(pe, k, w), fnp[a][q].x);
                        fnp[a][q].y = fmaf(w2.y, cost(pe + 1, k + SIGN, w), fnp[a][q].y);
                        fdp[a][q] = __fadd2_rn(fdp[a][q], w2);
                    }
                    const int ks0 = SIGN > 0 ? 3 : 0, ks1 = SIGN > 0 ? 0 : 3;
                    const int w0 = SIGN > 0 ? pe - ks0 + 4 : pe + ks0, w1 = SIGN > 0 ? pe + 1 - ks1 + 4 : pe + 1 + ks1;
                    const float u0 = __fmul_rn(wlv[pe], wrv[w0]), u1 = __fmul_rn(wlv[pe + 1], wrv[w1]);
                    fns[a][0] = fmaf(u0, cost(pe, ks0, w0), fns[a][0]);
                    fds[a][0] = __fadd_rn(fds[a][0], u0);
                    fns[a][1] = fmaf(u1, cost(pe + 1, ks1, w1), fns[a][1]);
                    fds[a][1] = __fadd_rn(fds[a][1], u1);
                }
            };
            switch (kx & 3) {
                case 0: tap(std::integral_constant<int, 0>{}); break;
                case 1: tap(std::integral_constant<int, 1>{}); break;
                case 2: tap(std::integral_constant<int, 2>{}); break;
                default: tap(std::integral_constant<int, 3>{}); break;
            }
        }
        __syncthreads();          // chunk c consumed by everyone, chunk c+1's weights visible
    }
    const int x0 = xb + 4 * pg;
#pragma unroll
    for (int p = 0; p < 4; p++) {
        const int x = x0 + p;
        if (x >= W) continue;
        unsigned long long best = WTA_KEY_EMPTY;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int ci = c0 + 4 * ds + k;
            if (ci >= g.n_cand) continue;
            const int a = p >> 1, odd = p & 1;
            const int kq = SIGN > 0 ? (odd ? k - 1 : k) : (odd ? k : k - 1);
            const bool single = odd ? (k == (SIGN > 0 ? 0 : 3)) : (k == (SIGN > 0 ? 3 : 0));
            float vn, vd;
            if (single) { vn = fns[a][odd]; vd = fds[a][odd]; }
            else { vn = odd ? fnp[a][kq].y : fnp[a][kq].x; vd = odd ? fdp[a][kq].y : fdp[a][kq].x; }
            const double E = (double)vn / (double)vd;
            if (agg) agg[(size_t)ci * H * W + rowoff + x] = (float)E;
            best = min(best, wta_key_d(E, g.d_first + ci));
        }
        atomicMin(&keys[rowoff + x], best);
    }
}

// grid: (segments, H, candidate chunks).  Interior and edge segments run in ONE launch (the edge variant is a
// CTA-uniform branch) so that small images still fill the GPU.
template <int SIGN, int NW>
__global__ void __launch_bounds__(32 * NW)
k_trad_diag(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, const float* __restrict__ c2h, float a2,
            TradGeom g, int cand_first, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    const int xb = blockIdx.x * TD_X;
    const int c0 = cand_first + blockIdx.z * (4 * NW);
    const bool border = SIGN > 0 ? xb + TD_X - 1 + g.h > g.W - 1 : xb - g.h < 0;
    if (border) trad_diag_body<SIGN, true, NW>(ref, tgt, c2h, a2, g, xb, c0, keys, agg);
    else trad_diag_body<SIGN, false, NW>(ref, tgt, c2h, a2, g, xb, c0, keys, agg);
}

template <int SIGN, int NW>
static asw_status trad_diag_launch(asw_ctx* ctx, const uint8_t* ref, const uint8_t* tgt, const float* c2h, float a2, TradGeom g,
                                   int cand_first, int n_chunks, unsigned long long* keys, float* agg) {
    if (n_chunks <= 0) return ASW_OK;
    constexpr int WRW = TD_X + 4 * NW + 4;
    const int h = g.h;
    size_t smem = (2 * (size_t)TD_TC * (TD_X + WRW) + (size_t)g.win * ((TD_X + 2 * h + 7) & ~3) +
                   (size_t)g.win * ((WRW + 2 * h + 7) & ~3) + 3 * (size_t)g.nw) * sizeof(float);
    if (smem > 220 * 1024) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "window too large for the diagonal kernel%s%s");
    cudaFuncSetAttribute(k_trad_diag<SIGN, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    LAUNCH(ctx, "trad_aggregate", (k_trad_diag<SIGN, NW><<<dim3(cdiv(g.W, TD_X), g.H, n_chunks), 32 * NW, smem, ctx->stream>>>(
                                      ref, tgt, c2h, a2, g, cand_first, keys, agg)));
    return ASW_OK;
}
// full chunks of 32 candidates (8 warps), then the remainder with the warps it needs
template <int SIGN>
static asw_status trad_diag_all(asw_ctx* ctx, const uint8_t* ref, const uint8_t* tgt, const float* c2h, float a2, TradGeom g,
                                unsigned long long* keys, float* agg) {
    const int full = g.n_cand / 32, rem = g.n_cand - full * 32, c1 = full * 32;
    ASW_TRY((trad_diag_launch<SIGN, 8>(ctx, ref, tgt, c2h, a2, g, 0, full, keys, agg)));
#define TRAD_REM(NWV) ASW_TRY((trad_diag_launch<SIGN, NWV>(ctx, ref, tgt, c2h, a2, g, c1, 1, keys, agg)))
    switch ((rem + 3) / 4) {
        case 1: TRAD_REM(1); break;
        case 2: TRAD_REM(2); break;
        case 3: TRAD_REM(3); break;
        case 4: TRAD_REM(4); break;
        case 5: TRAD_REM(5); break;
        case 6: TRAD_REM(6); break;
        case 7: TRAD_REM(7); break;
        case 8: TRAD_REM(8); break;
        default: break;
    }
#undef TRAD_REM
    return ASW_OK;
}

// host: exact weight table [nw][256], (float)(k * exp(-(delta/gamma_c + sqrt(i*i+j*j)/gamma_g))), k = 3
static void trad_build_table(int win, double gamma_c, double gamma_g, std::vector<float>& t) {
    int h = win / 2, nw = win * win - 1, cidx = win * win / 2;
    t.resize((size_t)nw * 256);
    const double k = 3;
    for (int n = 0; n < nw; n++) {
        int pw = n < cidx ? n : n + 1;
        int j = pw / win - h, i = pw % win - h;
        double delta_g = sqrt((double)(i * i + j * j));
        for (int dlt = 0; dlt < 256; dlt++)
            t[(size_t)n * 256 + dlt] = (float)(k * exp(-((double)dlt / gamma_c + delta_g / gamma_g)));
    }
}

// c2[n] = 2 g_n log2(e) / gamma_g - log2(9), g_n = spatial distance of tap n (k = 3 -> k * k = 9): the spatial part of
// w_L w_R = 9 * 2^-(a2 (dL + dR) + c2[n]), a2 = log2(e) / gamma_c.  Cached on the device per parameter set.
static asw_status trad_upload_c2(asw_ctx* ctx, const TradGeom& g, double gamma_g, float** dc2) {
    char key[96];
    snprintf(key, sizeof(key), "c2:%d:%d:%.17g", g.mode, g.win, gamma_g);
    ASW_TRY(ws_get(ctx, WS_TRAD_C2, (size_t)g.nw, dc2));
    if (table_cached(ctx, 0, WS_TRAD_C2, key)) return ASW_OK;
    std::vector<float> c2(g.nw);
    const double log2e = 1.4426950408889634;
    for (int i = 0; i < g.nw; i++) {
        int dy, dx, ky, kx;
        trad_tap(g.mode, i, g.win, g.cidx, &dy, &dx, &ky, &kx);
        c2[i] = (float)(2.0 * sqrt((double)(dx * dx + dy * dy)) * log2e / gamma_g - log2(9.0));
    }
    ASW_CUDA(ctx, cudaMemcpyAsync(*dc2, c2.data(), c2.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // c2 is a host temporary (first call with these parameters only)
    return ASW_OK;
}

// mode 0: computeAdaptiveWeight (A.cpp:1016-1156); mode 1: computeAdaptiveWeight_direct8 (A.cpp:1167-1319, LEFT only)
static asw_status dev_traditional(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, double gamma_c,
                                  double gamma_g, int disp_type, int win, int min_d, int num_d, float* disp_dev,
                                  float* agg_dev, int mode = 0) {
    size_t n = (size_t)H * W;
    uint8_t *gl, *gr;
    ASW_TRY(ws_get(ctx, WS_GRAY_L, n, &gl));
    ASW_TRY(ws_get(ctx, WS_GRAY_R, n, &gr));
    LAUNCH(ctx, "bgr2gray", (k_bgr2gray_pad<<<dim3(cdiv(W, 256), H), 256, 0, ctx->stream>>>(dL, H, W, 0, 0, gl)));   // A.cpp:1030-1033
    LAUNCH(ctx, "bgr2gray", (k_bgr2gray_pad<<<dim3(cdiv(W, 256), H), 256, 0, ctx->stream>>>(dR, H, W, 0, 0, gr)));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    TradGeom g;
    g.H = H; g.W = W; g.win = win; g.h = win / 2; g.nw = mode == 0 ? win * win - 1 : 3 * (win - 1); g.cidx = win * win / 2;
    g.mode = mode;
    g.sign = disp_type == ASW_DISPARITY_LEFT ? 1 : -1;
    g.d_first = min_d;
    g.n_cand = num_d + 1;                                    // A.cpp:1021, 1074: <= max_offset
    const uint8_t* ref = disp_type == ASW_DISPARITY_LEFT ? gl : gr;
    const uint8_t* tgt = disp_type == ASW_DISPARITY_LEFT ? gr : gl;
    const int h = win / 2, IH = TR_TH + 2 * h, IWr = TR_TW + 2 * h, IWt = TR_TW + 2 * h + TR_Q - 1;
    const double log2e = 1.4426950408889634;
    const size_t smem_fast = ((size_t)IH * IWr + (size_t)IH * IWt) * sizeof(float);
    const size_t smem_lin = smem_fast + (size_t)5 * g.nw * sizeof(float);
    const dim3 grid(cdiv(W, TR_TW), cdiv(H, TR_TH), cdiv(g.n_cand, TR_Q));
#ifdef ASW_DEV_KERNELS
    if (mode == 0 && !asw_dev("ASW_TRAD_EXACT") && asw_dev("ASW_TRAD_DIAG") && win <= 41) {
        // diagonal-blocked kernel (measured 15-25 % slower than k_trad_lin): per-side weights 2^(-(delta a2 + c2h[n]))
        std::vector<float> c2h(g.nw);
        for (int i = 0; i < g.nw; i++) {
            int pw = i < g.cidx ? i : i + 1;
            int dj = pw / win - h, di = pw % win - h;
            c2h[i] = (float)(sqrt((double)(di * di + dj * dj)) * log2e / gamma_g - log2(3.0));
        }
        float* dc2;
        ASW_TRY(ws_get(ctx, WS_TRAD_C2, c2h.size(), &dc2));
        ctx->table_key[0].clear();
        ASW_CUDA(ctx, cudaMemcpyAsync(dc2, c2h.data(), c2h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const float a2 = (float)(log2e / gamma_c);
        if (g.sign > 0) ASW_TRY((trad_diag_all<1>(ctx, ref, tgt, dc2, a2, g, keys, agg_dev)));
        else ASW_TRY((trad_diag_all<-1>(ctx, ref, tgt, dc2, a2, g, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    if (mode == 0 && !asw_dev("ASW_TRAD_EXACT") && asw_dev("ASW_TRAD_FAST") && smem_fast <= 200 * 1024) {
        float* dc2;
        ASW_TRY(trad_upload_c2(ctx, g, gamma_g, &dc2));
        cudaFuncSetAttribute(k_trad_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast);
        LAUNCH(ctx, "trad_aggregate", (k_trad_fast<<<grid, dim3(TR_TW, TR_TH), smem_fast, ctx->stream>>>(
                                          ref, tgt, dc2, (float)(log2e / gamma_c), g, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
#endif
    // the product kernel: tiled, clamp-free inner loop, one ex2 per (tap, candidate)
    if (!asw_dev("ASW_TRAD_EXACT") && !asw_dev("ASW_TRAD_GENERIC") && smem_lin <= 200 * 1024 && (size_t)IH * IWt < (1u << 20)) {
        float* dc2;
        ASW_TRY(trad_upload_c2(ctx, g, gamma_g, &dc2));
        if (g.sign > 0) {
            cudaFuncSetAttribute(k_trad_lin<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lin);
            LAUNCH(ctx, "trad_aggregate", (k_trad_lin<1><<<grid, dim3(TR_TW, TR_TH), smem_lin, ctx->stream>>>(
                                              ref, tgt, dc2, (float)(log2e / gamma_c), g, keys, agg_dev)));
        } else {
            cudaFuncSetAttribute(k_trad_lin<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lin);
            LAUNCH(ctx, "trad_aggregate", (k_trad_lin<-1><<<grid, dim3(TR_TW, TR_TH), smem_lin, ctx->stream>>>(
                                              ref, tgt, dc2, (float)(log2e / gamma_c), g, keys, agg_dev)));
        }
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    if (mode != 0) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "8-direction ASW: window too large for the tiled kernel%s%s");
    // windows whose tile does not fit shared memory: the size-generic kernel over the exact weight table
    // (float(3 exp(-(delta / gamma_c + g / gamma_g))) for every tap and gray difference, built in double as A.cpp:1065)
    float* dtable;
    ASW_TRY(ws_get(ctx, WS_TRAD_TABLE, (size_t)g.nw * 256, &dtable));
    {
        char key[96];
        snprintf(key, sizeof(key), "tab:%d:%.17g:%.17g", win, gamma_c, gamma_g);
        if (!table_cached(ctx, 1, WS_TRAD_TABLE, key)) {
            std::vector<float> table;
            trad_build_table(win, gamma_c, gamma_g, table);
            ASW_CUDA(ctx, cudaMemcpyAsync(dtable, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // table is a host temporary (first call only)
        }
    }
#ifdef ASW_DEV_KERNELS
    const size_t smem = ((size_t)win * 256 + (size_t)IH * IWr + (size_t)IH * IWt) * sizeof(float) + (size_t)IH * (IWr + IWt);
    if (smem <= 200 * 1024 && !asw_dev("ASW_TRAD_GENERIC")) {
        cudaFuncSetAttribute(k_trad_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        LAUNCH(ctx, "trad_aggregate", (k_trad_tile<<<grid, dim3(TR_TW, TR_TH), smem, ctx->stream>>>(ref, tgt, dtable, g, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
#endif
    LAUNCH(ctx, "trad_aggregate", (k_trad_aggregate<<<dim3(cdiv(W, 128), H, cdiv(g.n_cand, TRAD_Q)), 128, 0, ctx->stream>>>(ref, tgt, dtable, g, keys, agg_dev)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}
