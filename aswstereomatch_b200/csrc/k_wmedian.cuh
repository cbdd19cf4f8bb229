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

// ---------------------------------------------------------------------------------------------
// Fast path.  The reference itself precomputes one colour-weight window per pixel of either image
// (computeColorWeightGau, A.cpp:3139-3205) and re-uses it for every candidate; so do we:
//   k_wm_weights    W[t][y][x] = exp(-L1(I(neighbour t), I(y,x)) / rateR) for both images (the right one on its padded
//                   width, REFLECT inside the padded image as A.cpp:3156), one double-precision exp per entry, once
//   k_wm_aggregate2 per (d, y, x): keys (value bits << 32 | window index) of the win^2 costs go into a binary min-heap
//                   in shared memory (thread-private column: bank = thread, no conflicts) next to their weights (wL * spatial) * wR,
//                   elements are popped in exactly the reference's stable ascending order,
//                   partial sums in double, stop at the first element that crosses half the total (A.cpp:3276-3304).
// O(n + k log n) per evaluation instead of O(k n) extraction scans and 3 exps per element per candidate.
// ---------------------------------------------------------------------------------------------
// img: packed BGRx [H][W]; out: [n][H][Wout]; the centre of output column xo is image column reflect(xo - pad, W),
// neighbour columns reflect inside the padded width Wout first (pad = 0, Wout = W for the left image)
__global__ void k_wm_weights(const uint32_t* __restrict__ img, int H, int W, int Wout, int pad, int win, double alpha_r,
                             float* __restrict__ out) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, t = blockIdx.z;
    if (xo >= Wout) return;
    const int h = win / 2, wy = t / win, wx = t - wy * win;
    const uint32_t c = img[(size_t)y * W + border_idx(xo - pad, W, 0)];
    const int sy = border_idx(y - h + wy, H, 0);
    const int xb = border_idx(xo - h + wx, Wout, 0);
    const uint32_t q = img[(size_t)sy * W + border_idx(xb - pad, W, 0)];
    float d0 = (float)abs((int)(q & 0xFF) - (int)(c & 0xFF)), d1 = (float)abs((int)((q >> 8) & 0xFF) - (int)((c >> 8) & 0xFF));
    float d2 = (float)abs((int)((q >> 16) & 0xFF) - (int)((c >> 16) & 0xFF));
    out[((size_t)t * H + y) * Wout + xo] = (float)exp((double)(float)fma((double)__fadd_rn(d0, d1), alpha_r, __dmul_rn((double)d2, alpha_r)));
}

template <int BT>
__global__ void __launch_bounds__(BT)
k_wm_aggregate2(const float* __restrict__ WL, const float* __restrict__ WR, const float* __restrict__ cost, WmGeom g,
                float alpha_s, int d_label0, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    extern __shared__ unsigned long long sm_wm2[];
    const int win = g.win, h = g.h, n = win * win, W = g.W, H = g.H, Wr = W + g.max_off;
    unsigned long long* heap = sm_wm2 + threadIdx.x;         // heap[i * BT]: thread-private column
    float* wts = (float*)(sm_wm2 + (size_t)n * BT) + threadIdx.x;   // wts[t * BT]: the window's weights, thread-private column
    float* wsp = (float*)(sm_wm2 + (size_t)n * BT) + (size_t)n * BT;   // [n] spatial weights
    for (int e = threadIdx.x; e < n; e += BT) {
        int wy = e / win, wx = e - wy * win;
        float dist2 = __fadd_rn((float)((wx - h) * (wx - h)), (float)((wy - h) * (wy - h)));
        wsp[e] = (float)exp((double)__fmul_rn(dist2, alpha_s));                  // A.cpp:3219-3225
    }
    __syncthreads();
    const int x = blockIdx.x * BT + threadIdx.x, y = blockIdx.y, off = blockIdx.z;
    if (x >= W) return;
    const size_t p = (size_t)y * W + x;
    const int xr = x - off + g.D - 1;                        // weightWinsR[y][x - offset + numDisparity - 1]
    const float* wl = WL + p;                                // + t * H * W
    const float* wr = WR + (size_t)y * Wr + xr;              // + t * H * Wr
    const size_t sl = (size_t)H * W, sr = (size_t)H * Wr;
    const float* cs = cost + (size_t)off * sl;
    double total = 0;
    for (int t0 = 0; t0 < n; t0 += 8) {                      // 8 window elements per round: their 24 loads issued together
        float a[8], b[8], v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = min(t0 + u, n - 1), wy = t / win, wx = t - wy * win;
            a[u] = __ldg(&wl[t * sl]);
            b[u] = __ldg(&wr[t * sr]);
            v[u] = cs[(size_t)border_idx(y - h + wy, H, 0) * W + border_idx(x - h + wx, W, 0)];
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = t0 + u;
            if (t < n) {
                const float w = __fmul_rn(__fmul_rn(a[u], wsp[t]), b[u]);                          // wL.mul(dist).mul(wR)
                wts[t * BT] = w;
                heap[t * BT] = ((unsigned long long)__float_as_uint(v[u]) << 32) | (unsigned)t;   // costs are >= 0: bit order = value order
                total += (double)w;                                                                // cv::sum, double, window order
            }
        }
    }
    // heapify (min-heap): sift down from the last parent
    auto sift = [&](int i, int len, unsigned long long key) {
        for (;;) {
            int c = 2 * i + 1;
            if (c >= len) break;
            unsigned long long kc = heap[c * BT];
            if (c + 1 < len) { unsigned long long k2 = heap[(c + 1) * BT]; if (k2 < kc) { kc = k2; c++; } }
            if (kc >= key) break;
            heap[i * BT] = kc;
            i = c;
        }
        heap[i * BT] = key;
    };
    for (int i = n / 2 - 1; i >= 0; i--) sift(i, n, heap[i * BT]);
    const double half = total / 2;
    double partial = 0;
    float prev_v = 0.0f, result = 0.0f;
    bool first = true;
    for (int len = n; len > 0; len--) {
        const unsigned long long top = heap[0];
        const float bv = __uint_as_float((unsigned)(top >> 32));
        partial += (double)wts[(int)(top & 0xFFFFFFFFu) * BT];
        if (partial > half) { result = first ? bv : prev_v; break; }
        first = false; prev_v = bv;
        if (len > 1) sift(0, len - 1, heap[(len - 1) * BT]);
    }
    if (agg) agg[(size_t)off * sl + p] = result;
    atomicMin(&keys[p], wta_key(result, d_label0 + off));
}

// ---------------------------------------------------------------------------------------------
// Warp-per-evaluation path (the default).  The heap kernel keeps a 225-entry heap per THREAD in shared memory: 86 KB per
// 32-thread CTA, two warps per SM.  Here one warp owns one pixel and walks its candidates; the win^2 (cost bits << 32 | window
// index) keys sit 2 / 4 / 8 per lane in registers (64 / 128 / 256 slots for windows up to 7 / 11 / 15) and are sorted with a bitonic
// network (256 slots: 21 in-lane stages, 15 shuffle stages) -- the index in the low bits makes the keys distinct, so the order is the reference's stable ascending order.  The
// weights (wL * spatial) * wR go to a per-warp shared array and are gathered in sorted order; their running sum is a per-lane
// prefix + a warp scan in double, and the first element whose partial sum exceeds half the total selects its predecessor
// (A.cpp:3276-3304).  Weight planes are pixel-major ([y][x][tap]) so that a warp reads one pixel's window contiguously.
// 301 -> ~35 ms at 640 x 360 x 64, window 15.
// ---------------------------------------------------------------------------------------------
#define WM3_WARPS 4
// out[(y * Wout + xo) * n + t]
__global__ void k_wm_weights_px(const uint32_t* __restrict__ img, int H, int W, int Wout, int pad, int win, double alpha_r,
                                float* __restrict__ out) {
    const int n = win * win;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)H * Wout * n) return;
    const int t = (int)(i % n);
    const size_t pix = i / n;
    const int xo = (int)(pix % Wout), y = (int)(pix / Wout);
    const int h = win / 2, wy = t / win, wx = t - wy * win;
    const uint32_t c = img[(size_t)y * W + border_idx(xo - pad, W, 0)];
    const int sy = border_idx(y - h + wy, H, 0);
    const int xb = border_idx(xo - h + wx, Wout, 0);
    const uint32_t q = img[(size_t)sy * W + border_idx(xb - pad, W, 0)];
    float d0 = (float)abs((int)(q & 0xFF) - (int)(c & 0xFF)), d1 = (float)abs((int)((q >> 8) & 0xFF) - (int)((c >> 8) & 0xFF));
    float d2 = (float)abs((int)((q >> 16) & 0xFF) - (int)((c >> 16) & 0xFF));
    out[i] = (float)exp((double)(float)fma((double)__fadd_rn(d0, d1), alpha_r, __dmul_rn((double)d2, alpha_r)));
}

__device__ __forceinline__ void wm3_cx(unsigned long long& a, unsigned long long& b, bool asc) {   // in-lane compare-exchange
    const unsigned long long lo = min(a, b), hi = max(a, b);
    a = asc ? lo : hi; b = asc ? hi : lo;
}
// 32 * E keys, E per lane (element e = lane * E + r), ascending; E = 2, 4, 8
template <int E>
__device__ __forceinline__ void wm3_sort(unsigned long long (&k)[E], int lane) {
    constexpr int N = 32 * E;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= E) {
                const int lm = stride / E;                                 // partner lane = lane ^ lm
                const bool asc = size == N ? true : ((lane & (size / E)) == 0);
                const bool lower = (lane & lm) == 0;
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, k[r], lm);
                    k[r] = (lower == asc) ? min(k[r], o) : max(k[r], o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < E; r++) {
                    if ((r & stride) == 0) {
                        // direction bit of element e = lane * E + r for this merge size
                        const bool asc = size == N ? true : (size >= E ? ((lane & (size / E)) == 0) : ((r & size) == 0));
                        wm3_cx(k[r], k[r | stride], asc);
                    }
                }
            }
        }
    }
}

// E = key slots per lane: 2 for windows up to 7 (n <= 64), 4 up to 11 (n <= 128), 8 up to 15 (n <= 256)
template <int E>
__global__ void __launch_bounds__(32 * WM3_WARPS)
k_wm_aggregate3(const float* __restrict__ WL, const float* __restrict__ WR, const float* __restrict__ cost, WmGeom g,
                float alpha_s, int d_label0, unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    __shared__ float wsm[WM3_WARPS][32 * E];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int win = g.win, h = g.h, n = win * win, W = g.W, H = g.H, Wr = W + g.max_off;
    const int x = blockIdx.x * WM3_WARPS + wp, y = blockIdx.y;
    if (x >= W) return;
    const size_t p = (size_t)y * W + x, sl = (size_t)H * W;
    // per lane: its E window elements t = lane * E + r: cost offset inside a slice, wL * spatial (candidate-independent)
    int coff[E];
    float wls[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const int t = lane * E + r;
        if (t < n) {
            const int wy = t / win, wx = t - wy * win;
            coff[r] = border_idx(y - h + wy, H, 0) * W + border_idx(x - h + wx, W, 0);
            const float dist2 = __fadd_rn((float)((wx - h) * (wx - h)), (float)((wy - h) * (wy - h)));
            const float wsp = (float)exp((double)__fmul_rn(dist2, alpha_s));                 // A.cpp:3219-3225
            wls[r] = __fmul_rn(__ldg(&WL[p * n + t]), wsp);                                  // wL.mul(dist)
        } else { coff[r] = 0; wls[r] = 0.0f; }
    }
    unsigned long long best = WTA_KEY_EMPTY;
    for (int off = 0; off < g.D; off++) {
        const int xr = x - off + g.D - 1;                         // weightWinsR[y][x - offset + numDisparity - 1]
        const float* wr = WR + ((size_t)y * Wr + xr) * n;
        const float* cs = cost + (size_t)off * sl;
        unsigned long long k[E];
        double lsum = 0.0;
#pragma unroll
        for (int r = 0; r < E; r++) {
            const int t = lane * E + r;
            if (t < n) {
                const float w = __fmul_rn(wls[r], __ldg(&wr[t]));                            // .mul(wR)
                wsm[wp][t] = w;
                lsum += (double)w;
                k[r] = ((unsigned long long)__float_as_uint(cs[coff[r]]) << 32) | (unsigned)t;   // costs >= 0: bit order = value order
            } else {
                wsm[wp][t] = 0.0f;
                k[r] = 0xFFFFFFFFFFFFFFFFull;
            }
        }
        double total = lsum;                                       // cv::sum of the weights, double
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        const double half = total / 2;
        __syncwarp();
        wm3_sort<E>(k, lane);
        // running weight sums in sorted order
        double pre[E];
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < E; r++) { acc += (double)wsm[wp][(int)(k[r] & (unsigned)(32 * E - 1))]; pre[r] = acc; }
        double scan = acc;                                         // inclusive scan of the lane totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(0xffffffffu, scan, o);
            if (lane >= o) scan += up;
        }
        const double base = scan - acc;                            // sum of every element before this lane's
        int cross = E;
#pragma unroll
        for (int r = E - 1; r >= 0; r--) if (base + pre[r] > half) cross = r;
        const unsigned ball = __ballot_sync(0xffffffffu, cross < E);
        const unsigned prev_last = __shfl_up_sync(0xffffffffu, (unsigned)(k[E - 1] >> 32), 1);   // previous lane's largest value
        float result = 0.0f;
        if (ball) {
            const int src = __ffs(ball) - 1;                       // lane that holds the first crossing element
            unsigned vb = 0;
            if (lane == src) {
                // the element BEFORE the crossing one; the very first element of the order selects itself (A.cpp:3296-3302)
                if (cross == 0) {
                    vb = (src == 0) ? (unsigned)(k[0] >> 32) : prev_last;
                } else {
                    vb = (unsigned)(k[0] >> 32);
#pragma unroll
                    for (int r = 1; r < E - 1; r++) if (cross - 1 == r) vb = (unsigned)(k[r] >> 32);
                }
            }
            vb = __shfl_sync(0xffffffffu, vb, src);
            result = __uint_as_float(vb);
        }
        __syncwarp();                                              // wsm is rewritten by the next candidate
        if (lane == 0) {
            if (agg) agg[(size_t)off * sl + p] = result;
            best = min(best, wta_key(result, d_label0 + off));
        }
    }
    if (lane == 0) atomicMin(&keys[p], best);
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
    const int nn = win * win, Wr = W + v.max_off;
    const size_t planes = (size_t)nn * H * ((size_t)W + Wr) * sizeof(float);
    if (!asw_dev("ASW_WM_SCAN") && !asw_dev("ASW_WM_HEAP") && planes <= ((size_t)6 << 30) && nn <= 256) {
        float *WLp, *WRp;
        ASW_TRY(ws_get(ctx, WS_GEO_L, (size_t)nn * H * W, &WLp));
        ASW_TRY(ws_get(ctx, WS_GEO_R, (size_t)nn * H * Wr, &WRp));
        LAUNCH(ctx, "wm_weights", (k_wm_weights_px<<<(unsigned)(((size_t)nn * H * W + 255) / 256), 256, 0, ctx->stream>>>(pl, H, W, W, 0, win, alpha_r, WLp)));
        LAUNCH(ctx, "wm_weights", (k_wm_weights_px<<<(unsigned)(((size_t)nn * H * Wr + 255) / 256), 256, 0, ctx->stream>>>(pr, H, W, Wr, v.max_off, win, alpha_r, WRp)));
        const dim3 wgrid(cdiv(W, WM3_WARPS), H);
        if (nn <= 64) LAUNCH(ctx, "wm_aggregate", (k_wm_aggregate3<2><<<wgrid, 32 * WM3_WARPS, 0, ctx->stream>>>(WLp, WRp, cost, g, alpha_s, min_d, keys, agg_dev)));
        else if (nn <= 128) LAUNCH(ctx, "wm_aggregate", (k_wm_aggregate3<4><<<wgrid, 32 * WM3_WARPS, 0, ctx->stream>>>(WLp, WRp, cost, g, alpha_s, min_d, keys, agg_dev)));
        else LAUNCH(ctx, "wm_aggregate", (k_wm_aggregate3<8><<<wgrid, 32 * WM3_WARPS, 0, ctx->stream>>>(WLp, WRp, cost, g, alpha_s, min_d, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    if (!asw_dev("ASW_WM_SCAN") && planes <= ((size_t)6 << 30)) {
        float *WLp, *WRp;
        ASW_TRY(ws_get(ctx, WS_GEO_L, (size_t)nn * H * W, &WLp));
        ASW_TRY(ws_get(ctx, WS_GEO_R, (size_t)nn * H * Wr, &WRp));
        LAUNCH(ctx, "wm_weights", (k_wm_weights<<<dim3(cdiv(W, 128), H, nn), 128, 0, ctx->stream>>>(pl, H, W, W, 0, win, alpha_r, WLp)));
        LAUNCH(ctx, "wm_weights", (k_wm_weights<<<dim3(cdiv(Wr, 128), H, nn), 128, 0, ctx->stream>>>(pr, H, W, Wr, v.max_off, win, alpha_r, WRp)));
        constexpr int BT = 32;
        const size_t smem = (size_t)nn * BT * (sizeof(unsigned long long) + sizeof(float)) + (size_t)nn * sizeof(float);
        cudaFuncSetAttribute(k_wm_aggregate2<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        LAUNCH(ctx, "wm_aggregate", (k_wm_aggregate2<BT><<<dim3(cdiv(W, BT), H, num_d), BT, smem, ctx->stream>>>(
                                        WLp, WRp, cost, g, alpha_s, min_d, keys, agg_dev)));
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    LAUNCH(ctx, "wm_aggregate", (k_wm_aggregate<<<dim3(cdiv(W, 64), H, num_d), 64, 0, ctx->stream>>>(
                                    pl, pr, cost, g, alpha_r, alpha_s, min_d, keys, agg_dev)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}
