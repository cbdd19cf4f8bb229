// k_geodesic.cuh -- geodesic-distance ASW (A.cpp:1321-1534).
//
// K8 k_geo_dist: per pixel, the chamfer DP of getWinGeoDist (A.cpp:1328-1390) on its (win+2)^2 window cut
//   from the BORDER_REFLECT-padded image (A.cpp:1404).  Pass schedule of the reference: iterations 0 and 1
//   are the BACKWARD raster (neighbours R, BR, B, BL), iteration 2 the FORWARD raster (L, UL, U, UR) because
//   the test is iterCount/2 (A.cpp:1341, 1365).  The second backward sweep is provably a no-op (every cell
//   already equals the min over its final predecessors), so two sweeps are run.  Distances are exact
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
    for (int c = 0; c <= win + 1; c++) { prev[c] = GEO_INF; pprev[c] = ext[(size_t)(y + win + 1) * Wp + x + c]; }
    for (int r = win; r >= 1; r--) {
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
            int v = (int)dist[(size_t)((r - 1) * win + (c - 1)) * n + p];  // value left by the backward sweep
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

#define GEO_Q 4
struct GeoGeom { int H, W, win, h, sign, d_first, n_cand; };

// thread = pixel, blockIdx.z = chunk of GEO_Q candidates.  dref/dtgt: [win*win][H][W]; pref/ptgt: packed BGRx [H][W]
__global__ void __launch_bounds__(128)
k_geo_aggregate(const float* __restrict__ dref, const float* __restrict__ dtgt, const uint32_t* __restrict__ pref,
                const uint32_t* __restrict__ ptgt, GeoGeom g, unsigned long long* __restrict__ keys,
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
    for (int j = 0; j < win; j++) {
        int ny = clampi(y - h + j, 0, H - 1);
        float fn[GEO_Q], fd[GEO_Q];
#pragma unroll
        for (int q = 0; q < GEO_Q; q++) { fn[q] = 0; fd[q] = 0; }
        for (int i = 0; i < win; i++) {
            int nx = clampi(x - h + i, 0, W - 1);
            size_t tap = (size_t)(j * win + i) * n;
            float dl = dref[tap + p];
            uint32_t cr = pref[(size_t)ny * W + nx];
#pragma unroll
            for (int q = 0; q < GEO_Q; q++) {
                int nxs = g.sign > 0 ? max(0, nx - dq[q]) : min(nx + dq[q], W - 1);
                float cd = (float)__vsadu4(cr, ptgt[(size_t)ny * W + nxs]);     // getColorDist (A.cpp:1321-1326)
                float t = __fmul_rn(dl, dtgt[tap + pt[q]]);                     // float * float (A.cpp:1488-1489)
                fn[q] = __fadd_rn(fn[q], __fmul_rn(t, cd));
                fd[q] = __fadd_rn(fd[q], t);
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
            if (agg) agg[(size_t)(c0 + q) * n + p] = (float)E;
            best = min(best, wta_key_d(E, dq[q]));
        }
    }
    atomicMin(&keys[p], best);
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
    LAUNCH(ctx, "geo_dist", (k_geo_dist<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(ext, H, W, win, dist)));
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
    dim3 grid(cdiv(W, 128), H, cdiv(g.n_cand, GEO_Q));
    LAUNCH(ctx, "geo_aggregate", (k_geo_aggregate<<<grid, 128, 0, ctx->stream>>>(left ? distL : distR, left ? distR : distL,
                                                                                 left ? pl : pr, left ? pr : pl, g, keys, agg_dev)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}
