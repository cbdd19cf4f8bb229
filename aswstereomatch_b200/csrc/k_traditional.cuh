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

#include <vector>

struct TradGeom {
    int H, W, win, h, nw, cidx;
    int sign;        // LEFT: target column = max(0, x - d) ; RIGHT: min(x + d, W-1)
    int d_first;     // first candidate offset evaluated by this launch
    int n_cand;      // candidates in this launch (<= TRAD_Q per thread pass)
};

#define TRAD_Q 6

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
#define TR_TW 32
#define TR_TH 8
#define TR_Q 9

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

static asw_status dev_traditional(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, double gamma_c,
                                  double gamma_g, int disp_type, int win, int min_d, int num_d, float* disp_dev,
                                  float* agg_dev) {
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
    g.H = H; g.W = W; g.win = win; g.h = win / 2; g.nw = win * win - 1; g.cidx = win * win / 2;
    g.sign = disp_type == ASW_DISPARITY_LEFT ? 1 : -1;
    g.d_first = min_d;
    g.n_cand = num_d + 1;                                    // A.cpp:1021, 1074: <= max_offset
    const uint8_t* ref = disp_type == ASW_DISPARITY_LEFT ? gl : gr;
    const uint8_t* tgt = disp_type == ASW_DISPARITY_LEFT ? gr : gl;
    const int h = win / 2, IH = TR_TH + 2 * h, IWr = TR_TW + 2 * h, IWt = TR_TW + 2 * h + TR_Q - 1;
    const bool exact = getenv("ASW_TRAD_EXACT") != nullptr;
    size_t smem_fast = ((size_t)IH * IWr + (size_t)IH * IWt) * sizeof(float);
    if (!exact && smem_fast <= 200 * 1024) {
        // c2[n] = 2 g_n log2(e)/gamma_g - log2(9), a2 = log2(e)/gamma_c  (k = 3 -> k*k = 9)
        std::vector<float> c2(g.nw);
        const double log2e = 1.4426950408889634;
        for (int i = 0; i < g.nw; i++) {
            int pw = i < g.cidx ? i : i + 1;
            int dj = pw / win - h, di = pw % win - h;
            c2[i] = (float)(2.0 * sqrt((double)(di * di + dj * dj)) * log2e / gamma_g - log2(9.0));
        }
        float* dc2;
        ASW_TRY(ws_get(ctx, WS_TABLE1, c2.size(), &dc2));
        ASW_CUDA(ctx, cudaMemcpyAsync(dc2, c2.data(), c2.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // c2 is a host temporary
        cudaFuncSetAttribute(k_trad_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast);
        dim3 grid(cdiv(W, TR_TW), cdiv(H, TR_TH), cdiv(g.n_cand, TR_Q));
        LAUNCH(ctx, "trad_aggregate", (k_trad_fast<<<grid, dim3(TR_TW, TR_TH), smem_fast, ctx->stream>>>(
                                          ref, tgt, dc2, (float)(log2e / gamma_c), g, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    std::vector<float> table;
    trad_build_table(win, gamma_c, gamma_g, table);
    float* dtable;
    ASW_TRY(ws_get(ctx, WS_TABLE0, table.size(), &dtable));
    ASW_CUDA(ctx, cudaMemcpyAsync(dtable, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // table is a host temporary
    size_t smem = ((size_t)win * 256 + (size_t)IH * IWr + (size_t)IH * IWt) * sizeof(float) + (size_t)IH * (IWr + IWt);
    if (smem <= 200 * 1024 && !getenv("ASW_TRAD_GENERIC")) {
        cudaFuncSetAttribute(k_trad_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(cdiv(W, TR_TW), cdiv(H, TR_TH), cdiv(g.n_cand, TR_Q));
        LAUNCH(ctx, "trad_aggregate", (k_trad_tile<<<grid, dim3(TR_TW, TR_TH), smem, ctx->stream>>>(ref, tgt, dtable, g, keys, agg_dev)));
    } else {
        dim3 grid(cdiv(W, 128), H, cdiv(g.n_cand, TRAD_Q));
        LAUNCH(ctx, "trad_aggregate", (k_trad_aggregate<<<grid, 128, 0, ctx->stream>>>(ref, tgt, dtable, g, keys, agg_dev)));
    }
    return keys_to_disp(ctx, keys, n, disp_dev);
}
