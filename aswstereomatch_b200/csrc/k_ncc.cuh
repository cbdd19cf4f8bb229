// k_ncc.cuh -- NCC cost (getInputImgNCC A.cpp:767-800, computeNCC A.cpp:812-1013) and the GuidedF_3 method built on it
// (A.cpp:3063-3137): SURVEY row f-4.  Not a BASELINE config: written for parity first (every quirk of the reference is
// kept, see oracle/asw_oracle.c: RGB2GRAY on BGR bytes, REFLECT_101 mean over REFLECT windows, no square root, float
// products summed in double in window order, statistics of the PADDED target image), with the per-pixel statistics
// (window mean, sum of squared deviations) computed once per image instead of once per (pixel, disparity).
#pragma once
#include "k_guided.cuh"

// COLOR_RGB2GRAY applied to BGR bytes (A.cpp:829-836), REFLECT-padded on the column axis (A.cpp:857 / 886)
__global__ void k_rgb2gray_on_bgr_pad(const uint8_t* __restrict__ img, int H, int W, int pad_l, int pad_r, uint8_t* __restrict__ out) {
    const int Wp = W + pad_l + pad_r;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xp >= Wp) return;
    const uint8_t* px = img + ((size_t)y * W + border_idx(xp - pad_l, W, 0)) * 3;
    out[(size_t)y * Wp + xp] = (uint8_t)((9798 * px[0] + 19235 * px[1] + 3735 * px[2] + (1 << 14)) >> 15);
}

// per pixel: mean = boxFilter(u8 -> 32F, win) (REFLECT_101; integer window sum * (1 / win^2) in double, A.cpp:787-788) and
// s2 = sum over the REFLECT window of fl((p - mean)^2), accumulated in double in window row-major order (cv::sum)
__global__ void k_ncc_stats(const uint8_t* __restrict__ g, int H, int W, int win, float* __restrict__ mean, double* __restrict__ s2) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int h = win / 2;
    int sum = 0;
    for (int j = -h; j <= h; j++) {
        const uint8_t* row = g + (size_t)border_idx(y + j, H, 1) * W;
        for (int i = -h; i <= h; i++) sum += row[border_idx(x + i, W, 1)];
    }
    const float m = (float)((double)sum * (1.0 / ((double)win * win)));
    double acc = 0.0;
    for (int j = -h; j <= h; j++) {
        const uint8_t* row = g + (size_t)border_idx(y + j, H, 0) * W;
        for (int i = -h; i <= h; i++) {
            const float v = __fadd_rn((float)row[border_idx(x + i, W, 0)], -m);
            acc += (double)__fmul_rn(v, v);
        }
    }
    mean[(size_t)y * W + x] = m;
    s2[(size_t)y * W + x] = acc;
}

// raw cost sum(l r) / (sum(l l) sum(r r)) of candidates [c0, c0 + gridDim.z) : KEYS = false writes float(cost) into
// raw [slice][H][W] (vector overload, A.cpp:960-971); KEYS = true keeps the double and min-reduces 64-bit keys (Mat
// overload: the dispatcher's NCC, A.cpp:866-876)
template <bool KEYS>
__global__ void __launch_bounds__(128)
k_ncc_cost(const uint8_t* __restrict__ ref, const float* __restrict__ mr, const double* __restrict__ sr,
           const uint8_t* __restrict__ tgt, const float* __restrict__ mt, const double* __restrict__ st, int H, int W, int Wt,
           int win, int x0_base, int x0_step, int c0, int d_label0, float* __restrict__ raw, unsigned long long* __restrict__ keys) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, ci = c0 + blockIdx.z;
    if (x >= W) return;
    const int h = win / 2;
    const int xt = x + x0_base + x0_step * ci;                    // window centre in the padded target
    const float m0 = mr[(size_t)y * W + x], m1 = mt[(size_t)y * Wt + xt];
    double sxy = 0.0;
    for (int j = -h; j <= h; j++) {
        const int sy = border_idx(y + j, H, 0);                   // the windows come from BORDER_REFLECT copies (A.cpp:783)
        const uint8_t* ra = ref + (size_t)sy * W;
        const uint8_t* rb = tgt + (size_t)sy * Wt;
        for (int i = -h; i <= h; i++) {
            const float u = __fadd_rn((float)ra[border_idx(x + i, W, 0)], -m0);
            const float v = __fadd_rn((float)rb[border_idx(xt + i, Wt, 0)], -m1);
            sxy += (double)__fmul_rn(u, v);
        }
    }
    const double c = sxy / (sr[(size_t)y * W + x] * st[(size_t)y * Wt + xt]);
    const size_t p = (size_t)y * W + x;
    if (KEYS) atomicMin(&keys[p], wta_key_d(c, d_label0 + ci));
    else raw[(size_t)blockIdx.z * H * W + p] = (float)c;
}

// The same cost with the window rows staged in shared memory: one CTA = one image row x 128 columns x NCC_DC candidates.  The
// REFLECT indices are resolved once when a row is staged (the direct kernel evaluates two of them per tap), the taps then read
// floats at immediate offsets.  Identical float / double operations in the identical (window row-major) order: bit-identical
// results; 3.4 -> ~1 ms at 640 x 360 x 64, window 15.
#define NCC_COLS 128
#define NCC_DC 8
template <bool KEYS>
__global__ void __launch_bounds__(NCC_COLS)
k_ncc_cost_tile(const uint8_t* __restrict__ ref, const float* __restrict__ mr, const double* __restrict__ sr,
                const uint8_t* __restrict__ tgt, const float* __restrict__ mt, const double* __restrict__ st, int H, int W, int Wt,
                int win, int x0_base, int x0_step, int n_cand, int d_label0, float* __restrict__ raw, unsigned long long* __restrict__ keys) {
    extern __shared__ float sm_ncc[];
    const int h = win / 2, RW = NCC_COLS + 2 * h, TW = NCC_COLS + 2 * h + NCC_DC - 1;
    float* RA = sm_ncc;                 // [win][RW]  reference rows, cell c <-> column x0 - h + c (REFLECT)
    float* RB = sm_ncc + win * RW;      // [win][TW]  target rows, cell c <-> padded-target column t0 - h + c (REFLECT)
    const int x0 = blockIdx.x * NCC_COLS, y = blockIdx.y, c0 = blockIdx.z * NCC_DC;
    const int nc = min(NCC_DC, n_cand - c0);
    // target window centres of the chunk: x + x0_base + x0_step * ci, ci = c0 .. c0 + nc - 1
    const int ta = x0 + x0_base + x0_step * c0, tb = x0 + x0_base + x0_step * (c0 + nc - 1);
    const int t0 = min(ta, tb);
    for (int i = threadIdx.x; i < win * RW; i += NCC_COLS) {
        const int j = i / RW, c = i - j * RW;
        RA[i] = (float)ref[(size_t)border_idx(y - h + j, H, 0) * W + border_idx(x0 - h + c, W, 0)];
    }
    for (int i = threadIdx.x; i < win * TW; i += NCC_COLS) {
        const int j = i / TW, c = i - j * TW;
        RB[i] = (float)tgt[(size_t)border_idx(y - h + j, H, 0) * Wt + border_idx(t0 - h + c, Wt, 0)];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= W) return;
    const float m0 = mr[(size_t)y * W + x];
    const double s0 = sr[(size_t)y * W + x];
    const size_t p = (size_t)y * W + x;
    unsigned long long best = WTA_KEY_EMPTY;
    for (int k = 0; k < nc; k++) {
        const int ci = c0 + k;
        const int xt = x + x0_base + x0_step * ci;                // window centre in the padded target
        const float m1 = mt[(size_t)y * Wt + xt];
        const float* a = RA + threadIdx.x;                        // cell of column x - h
        const float* b = RB + (xt - t0);                          // cell of column xt - h
        double sxy = 0.0;
        for (int j = 0; j < win; j++) {
#pragma unroll 5
            for (int i = 0; i < win; i++) {
                const float u = __fadd_rn(a[j * RW + i], -m0);
                const float v = __fadd_rn(b[j * TW + i], -m1);
                sxy += (double)__fmul_rn(u, v);
            }
        }
        const double c = sxy / (s0 * st[(size_t)y * Wt + xt]);
        if (KEYS) best = min(best, wta_key_d(c, d_label0 + ci));
        else raw[(size_t)ci * H * W + p] = (float)c;
    }
    if (KEYS) atomicMin(&keys[p], best);
}
static inline size_t ncc_tile_smem(int win) { return (size_t)win * (2 * NCC_COLS + 4 * (win / 2) + NCC_DC - 1) * sizeof(float); }

// normalize(slice, 0, 1, NORM_MINMAX) in place (A.cpp:974-976), one slice per blockIdx.y
__global__ void k_normalize_slices(float* __restrict__ vol, size_t n, const uint32_t* __restrict__ mm) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = blockIdx.y;
    float sf, hf;
    minmax_scale_shift((double)from_orderable(mm[2 * d]), (double)from_orderable(mm[2 * d + 1]), &sf, &hf);
    float* s = vol + (size_t)d * n;
    s[i] = fmaf(s[i], sf, hf);
}
