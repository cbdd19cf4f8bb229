// k_wmedian.cuh -- weighted-median cost aggregation, computeAdaptiveWeight_WeightedMedian (A.cpp:3228-3383),
// DISPARITY_LEFT (the RIGHT branch lacks the begin() guard and its cost throws).
//
// Per (d, y, x): the weighted median of the win x win window of the REFLECT-padded TAD C+G cost slice
// (A.cpp:651-668) under weights  wL(y,x) .* spatial .* wR(y, x - d + D - 1)  (A.cpp:3273), with the reference's
// selection rule (A.cpp:3276-3304): stable ascending order, double partial sums, the element BEFORE the one that
// crosses half the total.  One thread per evaluation; the window's (value, weight) pairs live in local memory
// and are consumed in exactly the reference's order (repeated "next smallest (value, index)" extraction), so
// the selection is bit-faithful.  O(win^4) per evaluation: like the reference this is a small-window method
// (win <= 15 here; the reference keeps a win^2 Mat per pixel and is 640x360-class only).
#pragma once
#include "k_refine.cuh"

#define WM_MAXWIN 15
#define WM_MAXN (WM_MAXWIN * WM_MAXWIN)

struct WmGeom { int H, W, win, h, D, max_off; };

// lpk: packed BGRx left image [H][W]; rpk: packed right image [H][W]; cost: [D][H][W] (unpadded; the reference's
// copyMakeBorder(REFLECT) is folded into the addressing)
__global__ void __launch_bounds__(64)
k_wm_aggregate(const uint32_t* __restrict__ lpk, const uint32_t* __restrict__ rpk, const float* __restrict__ cost,
               WmGeom g, double alpha_r, float alpha_s, int d_label0, unsigned long long* __restrict__ keys,
               float* __restrict__ agg) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, off = blockIdx.z;
    if (x >= g.W) return;
    const int win = g.win, h = g.h, n = win * win, W = g.W, H = g.H;
    const int Wr = W + g.max_off;                          // rightImg_border width (A.cpp:3246)
    float v[WM_MAXN], w[WM_MAXN];
    const float* cs = cost + (size_t)off * H * W;
    uint32_t cl = lpk[(size_t)y * W + x];
    int xr = x - off + g.D - 1;                            // weightWinsR[y][x - offset + numDisparity - 1]
    uint32_t cr = rpk[(size_t)y * W + border_idx(xr - g.max_off, W, 0)];
    double total = 0;
    for (int wy = 0; wy < win; wy++) {
        int sy = border_idx(y - h + wy, H, 0);
        for (int wx = 0; wx < win; wx++) {
            int t = wy * win + wx;
            int sx = border_idx(x - h + wx, W, 0);
            v[t] = cs[(size_t)sy * W + sx];
            uint32_t ql = lpk[(size_t)sy * W + sx];
            int xb = border_idx(xr - h + wx, Wr, 0);       // REFLECT inside the padded right image (A.cpp:3156)
            uint32_t qr = rpk[(size_t)sy * W + border_idx(xb - g.max_off, W, 0)];
            // colour weights: exp((|dB|+|dG|+|dR|) * (-1/rateR)) with the addWeighted lowering (A.cpp:3177-3179)
            float wl, wr;
            {
                float d0 = (float)abs((int)(ql & 0xFF) - (int)(cl & 0xFF)), d1 = (float)abs((int)((ql >> 8) & 0xFF) - (int)((cl >> 8) & 0xFF));
                float d2 = (float)abs((int)((ql >> 16) & 0xFF) - (int)((cl >> 16) & 0xFF));
                wl = (float)exp((double)(float)fma((double)__fadd_rn(d0, d1), alpha_r, __dmul_rn((double)d2, alpha_r)));
            }
            {
                float d0 = (float)abs((int)(qr & 0xFF) - (int)(cr & 0xFF)), d1 = (float)abs((int)((qr >> 8) & 0xFF) - (int)((cr >> 8) & 0xFF));
                float d2 = (float)abs((int)((qr >> 16) & 0xFF) - (int)((cr >> 16) & 0xFF));
                wr = (float)exp((double)(float)fma((double)__fadd_rn(d0, d1), alpha_r, __dmul_rn((double)d2, alpha_r)));
            }
            float dist2 = __fadd_rn((float)((wx - h) * (wx - h)), (float)((wy - h) * (wy - h)));
            float wd = (float)exp((double)__fmul_rn(dist2, alpha_s));            // A.cpp:3219-3225
            w[t] = __fmul_rn(__fmul_rn(wl, wd), wr);                              // wL.mul(dist).mul(wR)
            total += (double)w[t];                                                // cv::sum, double
        }
    }
    const double half = total / 2;
    double partial = 0;
    float last_v = -INFINITY, prev_v = 0.0f, result = 0.0f;
    int last_i = -1;
    bool first = true;
    for (int step = 0; step < n; step++) {
        // next element in stable ascending (value, index) order after (last_v, last_i)
        float bv = INFINITY; int bi = -1;
        for (int t = 0; t < n; t++) {
            float tv = v[t];
            bool after = (tv > last_v) || (tv == last_v && t > last_i);
            if (after && (bi < 0 || tv < bv)) { bv = tv; bi = t; }
        }
        if (bi < 0) break;                                  // NaN costs are never ordered: stop like an exhausted map
        partial += (double)w[bi];
        if (partial > half) { result = first ? bv : prev_v; break; }
        first = false; prev_v = bv; last_v = bv; last_i = bi;
    }
    size_t p = (size_t)y * W + x;
    if (agg) agg[(size_t)off * H * W + p] = result;
    atomicMin(&keys[p], wta_key(result, d_label0 + off));
}

static asw_status dev_weighted_median(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int win,
                                      double rate_s, double rate_r, int min_d, int num_d, float* disp_dev, float* agg_dev) {
    if (win > WM_MAXWIN)
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "weighted-median aggregation is implemented for windows up to 15%s%s");
    if (min_d != 0)
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "weighted median indexes slices by offset (A.cpp:3264-3274): minDisparity must be 0%s%s");
    size_t n = (size_t)H * W;
    ViewGeom v = make_view(dL, dR, H, W, ASW_DISPARITY_LEFT, min_d, num_d);
    Feat *fref, *ftgt;
    float* cost;
    ASW_TRY(ws_get(ctx, WS_FEAT_REF, n, &fref));
    ASW_TRY(ws_get(ctx, WS_FEAT_TGT, (size_t)H * v.Wp, &ftgt));
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &cost));
    LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(W, 128), H), 128, 0, ctx->stream>>>(v.ref, H, W, 0, 0, fref)));
    LAUNCH(ctx, "features", (k_features<<<dim3(cdiv(v.Wp, 128), H), 128, 0, ctx->stream>>>(v.tgt, H, W, v.pad_l, v.pad_r, ftgt)));
    TadParams tp = make_tad_params(0.4, 10, 50);                                 // A.cpp:3250
    LAUNCH(ctx, "cost_tad_volume", (k_cost_tad_volume<<<dim3(cdiv(W, 128), H, num_d), 128, 0, ctx->stream>>>(
                                       fref, ftgt, H, W, v.Wp, v.x0_base, v.x0_step, tp, cost)));
    uint32_t *pl, *pr;
    ASW_TRY(ws_get(ctx, WS_TMP0, n, &pl));
    ASW_TRY(ws_get(ctx, WS_TMP1, n, &pr));
    LAUNCH(ctx, "pack_bgrx", (k_pack_bgrx<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dL, n, pl)));
    LAUNCH(ctx, "pack_bgrx", (k_pack_bgrx<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dR, n, pr)));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    WmGeom g;
    g.H = H; g.W = W; g.win = win; g.h = win / 2; g.D = num_d; g.max_off = v.max_off;
    double alpha_r = (1.0 / rate_r) * (-1);
    float alpha_s = (float)((1.0 / rate_s) * (-1));
    LAUNCH(ctx, "wm_aggregate", (k_wm_aggregate<<<dim3(cdiv(W, 64), H, num_d), 64, 0, ctx->stream>>>(
                                    pl, pr, cost, g, alpha_r, alpha_s, min_d, keys, agg_dev)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}
