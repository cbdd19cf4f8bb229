// k_refine.cuh -- stage 4: LR consistency check, invalid fill, weighted-median refinement.
// NOT in the reference: our specification, re-using the reference's weight formulas
// (computeColorWeightGau A.cpp:3139-3205, computeSpaceWeightGau A.cpp:3207-3226) and its median
// selection rule (A.cpp:3276-3304).  Mask / fill / selection are integer-exact against the oracle.
#pragma once
#include "k_cost.cuh"

// column of the right map a left-view disparity d points at: x - (int)d clamped to the row (the spec's max(0, .), plus
// the upper clamp so that negative / non-finite user maps given to the stage-level entry never read out of bounds;
// __float2int_rz saturates and maps NaN to 0)
__device__ __forceinline__ int lr_target_col(int x, float d, int W) {
    const long long t = (long long)x - (long long)__float2int_rz(d);
    return (int)min(max(t, 0ll), (long long)(W - 1));
}
// valid = |dL(y,x) - dR(y, clamp(x - (int)dL, 0, W-1))| <= tol
__global__ void k_lr_check(const float* __restrict__ dl, const float* __restrict__ dr, int H, int W, float tol,
                           uint8_t* __restrict__ valid) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    float d = dl[(size_t)y * W + x];
    int xr = lr_target_col(x, d, W);
    valid[(size_t)y * W + x] = fabsf(d - dr[(size_t)y * W + xr]) <= tol ? 1 : 0;
}

// invalid pixel <- min(nearest valid to the left, nearest valid to the right) on the same row
__global__ void k_fill_invalid(const float* __restrict__ d, const uint8_t* __restrict__ valid, int H, int W,
                               float* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const float* row = d + (size_t)y * W;
    const uint8_t* v = valid + (size_t)y * W;
    float r = row[x];
    if (!v[x]) {
        int xl = x - 1; while (xl >= 0 && !v[xl]) xl--;
        int xr = x + 1; while (xr < W && !v[xr]) xr++;
        if (xl >= 0 && xr < W) r = fminf(row[xl], row[xr]);
        else if (xl >= 0) r = row[xl];
        else if (xr < W) r = row[xr];
    }
    out[(size_t)y * W + x] = r;
}

// exp(-(|dB|+|dG|+|dR|)/rateR) with the float lowering of A.cpp:3177-3179:
// (w0 + w1 + w2) / rateR * (-1) -> addWeighted(w0+w1, alpha, w2, alpha), alpha = (1/rateR)*(-1)
__device__ __forceinline__ float wm_color_weight(const uint8_t* c, const uint8_t* q, double alpha) {
    float d0 = (float)abs((int)q[0] - (int)c[0]), d1 = (float)abs((int)q[1] - (int)c[1]);
    float d2 = (float)abs((int)q[2] - (int)c[2]);
    float m1 = __fadd_rn(d0, d1);
    float arg = (float)fma((double)m1, alpha, __dmul_rn((double)d2, alpha));
    return (float)exp((double)arg);
}

// One thread per invalid pixel.  Window on the REFLECT-padded (A.cpp:3156) filled map.
// Selection (A.cpp:3276-3304): stable ascending sort by value (ties keep window row-major order),
// weights accumulated in double in that order; at the first element whose partial sum exceeds
// total/2 return the PREVIOUS element's value (the first element's own value if it is the first).
__global__ void k_wmedian_refine(const uint8_t* __restrict__ img, const float* __restrict__ filled,
                                 const uint8_t* __restrict__ valid, int H, int W, int win, double alpha_r,
                                 float alpha_s, float* __restrict__ out) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    size_t p = (size_t)y * W + x;
    if (valid[p]) { out[p] = filled[p]; return; }
    const int h = win / 2;
    const uint8_t* c = img + p * 3;
    // weight of window element (wy, wx): colour * spatial (float product, A.cpp:3273 order)
    auto weight = [&](int wy, int wx, int sy, int sx) -> float {
        float wc = wm_color_weight(c, img + ((size_t)sy * W + sx) * 3, alpha_r);
        float dist2 = __fadd_rn((float)((wx - h) * (wx - h)), (float)((wy - h) * (wy - h)));
        float wd = (float)exp((double)__fmul_rn(dist2, alpha_s));                 // A.cpp:3219-3225
        return __fmul_rn(wc, wd);
    };
    double total = 0;
    for (int wy = 0; wy < win; wy++) {
        int sy = border_idx(y - h + wy, H, 0);
        for (int wx = 0; wx < win; wx++) {
            int sx = border_idx(x - h + wx, W, 0);
            total += (double)weight(wy, wx, sy, sx);
        }
    }
    const double half = total / 2;
    double partial = 0;
    float cur = -INFINITY, prev_val = 0.0f, result = filled[p];
    bool first = true, done = false;
    while (!done) {
        // next distinct value above cur
        float nxt = INFINITY; bool found = false;
        for (int wy = 0; wy < win; wy++) {
            int sy = border_idx(y - h + wy, H, 0);
            for (int wx = 0; wx < win; wx++) {
                float v = filled[(size_t)sy * W + border_idx(x - h + wx, W, 0)];
                if (v > cur && v <= nxt) { nxt = v; found = true; }
            }
        }
        if (!found) break;
        for (int wy = 0; wy < win && !done; wy++) {
            int sy = border_idx(y - h + wy, H, 0);
            for (int wx = 0; wx < win; wx++) {
                int sx = border_idx(x - h + wx, W, 0);
                if (filled[(size_t)sy * W + sx] != nxt) continue;
                partial += (double)weight(wy, wx, sy, sx);
                if (partial > half) {
                    result = first ? nxt : prev_val;
                    done = true;
                    break;
                }
                first = false;
                prev_val = nxt;
            }
        }
        cur = nxt;
    }
    out[p] = result;
}

// ---------------------------------------------------------------------------------------------
// List-driven variant (the default when the window fits in shared memory).  Only the pixels the LR check rejected
// are refined, and they are few and scattered: k_refine_compact copies the valid pixels through and appends the
// invalid ones to a list, k_wmedian_refine_list walks the list with one thread per pixel (no idle lanes), evaluates
// every window weight ONCE (one double-precision exp per element; the spatial factor comes from a per-CTA table)
// into a thread-private shared-memory column together with the window's values, and runs the selection of
// A.cpp:3276-3304 on those columns.  Same operations in the same order as k_wmedian_refine: identical results.
// ---------------------------------------------------------------------------------------------
__global__ void k_refine_compact(const float* __restrict__ filled, const uint8_t* __restrict__ valid, int n,
                                 float* __restrict__ out, int* __restrict__ list, int* __restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (valid[i]) { out[i] = filled[i]; return; }
    list[atomicAdd(count, 1)] = i;
}

template <int BT>
__global__ void __launch_bounds__(BT)
k_wmedian_refine_list(const uint8_t* __restrict__ img, const float* __restrict__ filled, const int* __restrict__ list,
                      const int* __restrict__ count, int H, int W, int win, double alpha_r, float alpha_s,
                      float* __restrict__ out) {
    extern __shared__ float sm_wm[];
    const int ne = win * win, h = win / 2, tid = threadIdx.x;
    float* wsp = sm_wm;                      // [ne]      exp(-dist^2 / rateS)  (A.cpp:3219-3225)
    float* wts = wsp + ne;                   // [ne][BT]  colour * spatial weight of the thread's pixel
    float* vals = wts + (size_t)ne * BT;     // [ne][BT]  window of the filled map
    for (int e = tid; e < ne; e += BT) {
        int wy = e / win, wx = e - wy * win;
        float dist2 = __fadd_rn((float)((wx - h) * (wx - h)), (float)((wy - h) * (wy - h)));
        wsp[e] = (float)exp((double)__fmul_rn(dist2, alpha_s));
    }
    __syncthreads();
    const int n_inv = *count;
    for (int idx = blockIdx.x * BT + tid; idx < n_inv; idx += gridDim.x * BT) {
        const int p = list[idx], y = p / W, x = p - y * W;
        const uint8_t* c = img + (size_t)p * 3;
        double total = 0;
        for (int wy = 0, e = 0; wy < win; wy++) {
            const int sy = border_idx(y - h + wy, H, 0);
            for (int wx = 0; wx < win; wx++, e++) {
                const int sx = border_idx(x - h + wx, W, 0);
                const size_t q = (size_t)sy * W + sx;
                const float w = __fmul_rn(wm_color_weight(c, img + q * 3, alpha_r), wsp[e]);   // A.cpp:3273 order
                wts[e * BT + tid] = w;
                vals[e * BT + tid] = filled[q];
                total += (double)w;
            }
        }
        const double half = total / 2;
        double partial = 0;
        float cur = -INFINITY, prev_val = 0.0f, result = filled[p];
        bool first = true, done = false;
        while (!done) {
            float nxt = INFINITY; bool found = false;                 // next distinct value above cur
            for (int e = 0; e < ne; e++) {
                const float v = vals[e * BT + tid];
                if (v > cur && v <= nxt) { nxt = v; found = true; }
            }
            if (!found) break;
            for (int e = 0; e < ne; e++) {                            // window row-major = the multimap's insertion order
                if (vals[e * BT + tid] != nxt) continue;
                partial += (double)wts[e * BT + tid];
                if (partial > half) { result = first ? nxt : prev_val; done = true; break; }
                first = false;
                prev_val = nxt;
            }
            cur = nxt;
        }
        out[p] = result;
    }
}

// LR check + fill + compaction of one image row per CTA (the three steps of k_lr_check, k_fill_invalid and
// k_refine_compact in one launch: the row's mask stays in shared memory between them).
__global__ void __launch_bounds__(256)
k_lr_fill_compact(const float* __restrict__ dl, const float* __restrict__ dr, int H, int W, float tol,
                  uint8_t* __restrict__ valid, float* __restrict__ filled, float* __restrict__ out,
                  int* __restrict__ list, int* __restrict__ count) {
    extern __shared__ uint8_t sm_valid[];
    const int y = blockIdx.x;
    const float* row = dl + (size_t)y * W;
    const float* rrow = dr + (size_t)y * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const float d = row[x];
        const uint8_t v = fabsf(d - rrow[lr_target_col(x, d, W)]) <= tol ? 1 : 0;
        sm_valid[x] = v;
        valid[(size_t)y * W + x] = v;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        const size_t p = (size_t)y * W + x;
        float r = row[x];
        if (sm_valid[x]) { filled[p] = r; out[p] = r; continue; }
        int xl = x - 1; while (xl >= 0 && !sm_valid[xl]) xl--;
        int xr = x + 1; while (xr < W && !sm_valid[xr]) xr++;
        if (xl >= 0 && xr < W) r = fminf(row[xl], row[xr]);
        else if (xl >= 0) r = row[xl];
        else if (xr < W) r = row[xr];
        filled[p] = r;
        list[atomicAdd(count, 1)] = (int)p;
    }
}

// driver post-processing (aswStereoMatch.cpp:97-98): convertTo(CV_8UC1) (cvRound + saturate) and its min / max ...
__global__ void k_disp_round_minmax(const float* __restrict__ d, size_t n, uint8_t* __restrict__ out, int* __restrict__ mm) {
    int mn = 255, mx = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int v = min(max(__float2int_rn(d[i]), 0), 255);
        out[i] = (uint8_t)v;
        mn = min(mn, v); mx = max(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], mn); atomicMax(&mm[1], mx); }
}
// ... then normalize(0, 255, NORM_MINMAX) of the u8 map: double scale / shift, applied in float with one fused multiply-add
__global__ void k_disp_normalize_u8(uint8_t* __restrict__ io, size_t n, const int* __restrict__ mm) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double span = (double)(mm[1] - mm[0]);
    const double scale = 255.0 * (span > 2.220446049250313e-16 ? 1.0 / span : 0.0);
    const double shift = 0.0 - (double)mm[0] * scale;
    io[i] = (uint8_t)min(max(__float2int_rn(fmaf((float)io[i], (float)scale, (float)shift)), 0), 255);
}
