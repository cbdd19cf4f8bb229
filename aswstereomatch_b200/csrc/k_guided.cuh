// k_guided.cuh -- guided-filter ASW (A.cpp:2766-2854, 2976-3050).
//
// Fast path (GuidedF_2, guidance = reference image, C = 3), two HBM passes per slice chunk:
//   pass 1  k_gf_ab : TAD C+G cost recomputed on chip from 16-byte feature records, box(c'), box(I_c c')
//                     -> a_c, b written as one float4 per disparity evaluation (16 B/DE)
//   pass 2  k_gf_q  : box(a), box(b) -> q = a.I + b, per-slice affine (the reference's NORM_MINMAX),
//                     WTA folded into 64-bit atomicMin keys
// The per-slice min-max normalisation p = c*s + h (A.cpp:2775) is affine and the filter is linear in p,
// so q(p) = s*q(c - c0) + (c0*s + h); the kernels filter c' = c - c0 (c0 = smallest possible cost, which
// keeps the float magnitudes equal to p's) while pass 1 also reduces the slice min/max, and pass 2
// applies (s, c0*s + h) before the WTA.  Loop-invariant guidance moments (mean_I, var_I) are computed
// once per image instead of once per slice (A.cpp:2774, 2778, 2796 recompute them every slice).
//
// Generic path (any C in {1,3,6}, guidance may change per slice): plane kernels + k_box_f32, used by
// GuidedF (6-channel guidance) and the stage-level asw_guided_filter.
#pragma once
#include "k_cost.cuh"

// ------------------------------------------------------------------------------------------------
// guidance preparation (per image)
// ------------------------------------------------------------------------------------------------
// planes[c] = I_c, planes[C + c] = I_c * I_c, Gi[y][x] = {I0, I1, I2, 0} (only when C == 3)
__global__ void k_guide_normalize(const uint8_t* __restrict__ img, size_t n, int C, const int* __restrict__ mm,
                                  float* __restrict__ planes, float4* __restrict__ Gi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // blockIdx.y = slice of a batch (GuidedF / GuidedF_3: one guidance image per slice), every array slice-major
    img += (size_t)blockIdx.y * n * C; mm += 2 * blockIdx.y; planes += (size_t)blockIdx.y * 2 * C * n;
    float sf, hf;
    minmax_scale_shift((double)mm[0], (double)mm[1], &sf, &hf);
    float v[6];
    for (int c = 0; c < C; c++) {
        v[c] = fmaf((float)img[i * C + c], sf, hf);                 // cv::normalize (A.cpp:2774)
        planes[(size_t)c * n + i] = v[c];
        planes[(size_t)(C + c) * n + i] = __fmul_rn(v[c], v[c]);    // guidedImg.mul(guidedImg) (A.cpp:2796)
    }
    if (Gi && C == 3) Gi[i] = make_float4(v[0], v[1], v[2], 0.0f);
}
// in: boxed[c] = mean_I_c, boxed[C+c] = corr_II_c.  out: den planes in boxed[C+c] (var + eps), and for
// C == 3 the packed records Gm = {mI}, Gd = {den}.
__global__ void k_guide_finish(float* __restrict__ boxed, size_t n, int C, float eps, float4* __restrict__ Gm,
                               float4* __restrict__ Gd) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    boxed += (size_t)blockIdx.y * 2 * C * n;                        // slice of a batch
    float m[6], dn[6];
    for (int c = 0; c < C; c++) {
        m[c] = boxed[(size_t)c * n + i];
        float var = __fsub_rn(boxed[(size_t)(C + c) * n + i], __fmul_rn(m[c], m[c]));   // A.cpp:2799
        dn[c] = __fadd_rn(__fmul_rn(1.0f, eps), var);                                   // A.cpp:2846
        boxed[(size_t)(C + c) * n + i] = dn[c];
    }
    if (Gm && C == 3) {
        Gm[i] = make_float4(m[0], m[1], m[2], 0.0f);
        Gd[i] = make_float4(dn[0], dn[1], dn[2], 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// generic (plane) path kernels
// ------------------------------------------------------------------------------------------------
// p = normalize(cost) per slice; planes[0] = p, planes[1+c] = I_c * p          (A.cpp:2775, 2787-2792)
// blockIdx.y = slice of a batch: cost / mm / planes are slice-major, the guidance arrays advance by gstride floats per slice
// (0: one guidance image for every slice)
__global__ void k_gfg_products(const float* __restrict__ cost, const uint32_t* __restrict__ mm_slice,
                               const float* __restrict__ I, size_t n, int C, float* __restrict__ planes, size_t gstride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cost += (size_t)blockIdx.y * n; mm_slice += 2 * blockIdx.y; I += (size_t)blockIdx.y * gstride;
    planes += (size_t)blockIdx.y * (1 + C) * n;
    float sf, hf;
    minmax_scale_shift((double)from_orderable(mm_slice[0]), (double)from_orderable(mm_slice[1]), &sf, &hf);
    float p = fmaf(cost[i], sf, hf);
    planes[i] = p;
    for (int c = 0; c < C; c++) planes[(size_t)(1 + c) * n + i] = __fmul_rn(I[(size_t)c * n + i], p);
}
// in: boxed[0] = mean_p, boxed[1+c] = corr_Ip_c; out: boxed[0] = b, boxed[1+c] = a_c   (A.cpp:2805-2847)
__global__ void k_gfg_ab(float* __restrict__ boxed, const float* __restrict__ mI, const float* __restrict__ den,
                         size_t n, int C, size_t gstride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    boxed += (size_t)blockIdx.y * (1 + C) * n; mI += (size_t)blockIdx.y * gstride; den += (size_t)blockIdx.y * gstride;
    float mP = boxed[i];
    float dot = 0.0f;
    for (int c = 0; c < C; c++) {
        float m = mI[(size_t)c * n + i];
        float cov = __fsub_rn(boxed[(size_t)(1 + c) * n + i], __fmul_rn(m, mP));
        float a = __fdiv_rn(cov, den[(size_t)c * n + i]);
        boxed[(size_t)(1 + c) * n + i] = a;
        float t = __fmul_rn(a, m);
        dot = (c == 0) ? t : __fadd_rn(dot, t);                     // Vec dot, left to right (A.cpp:22-31)
    }
    boxed[i] = __fsub_rn(mP, dot);
}
// q = sum_c abar_c * I_c + bbar                                                 (A.cpp:2852)
__global__ void k_gfg_q(const float* __restrict__ boxed, const float* __restrict__ I, size_t n, int C,
                        float* __restrict__ q, size_t gstride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    boxed += (size_t)blockIdx.y * (1 + C) * n; I += (size_t)blockIdx.y * gstride; q += (size_t)blockIdx.y * n;
    float dot = 0.0f;
    for (int c = 0; c < C; c++) {
        float t = __fmul_rn(boxed[(size_t)(1 + c) * n + i], I[(size_t)c * n + i]);
        dot = (c == 0) ? t : __fadd_rn(dot, t);
    }
    q[i] = __fadd_rn(dot, boxed[i]);
}
// 6-channel guidance for GuidedF: L (+) crop of padded R (A.cpp:2905-2912)
// blockIdx.z = slice of a batch: crop column x0 + blockIdx.z * x0_step, output slice-major
__global__ void k_merge_guide6(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt_pad, int H, int W,
                               int Wp, int x0, int ref_first, uint8_t* __restrict__ out, int x0_step) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    x0 += (int)blockIdx.z * x0_step; out += (size_t)blockIdx.z * H * W * 6;
    const uint8_t* a = ref + ((size_t)y * W + x) * 3;
    const uint8_t* b = tgt_pad + ((size_t)y * Wp + x0 + x) * 3;
    uint8_t* o = out + ((size_t)y * W + x) * 6;
    const uint8_t* first = ref_first ? a : b;
    const uint8_t* second = ref_first ? b : a;
    o[0] = first[0]; o[1] = first[1]; o[2] = first[2];
    o[3] = second[0]; o[4] = second[1]; o[5] = second[2];
}

// ------------------------------------------------------------------------------------------------
// fast path, pass 1: cost + first box level -> a, b
// ------------------------------------------------------------------------------------------------
// Tile geometry: TW x TH outputs per CTA, window k (radius a = k/2, anchor k/2), all DC slices of the
// chunk processed by the same CTA so the guidance records stay in L1.
#define GF_TW 64
#define GF_TH 32
#define GF_THREADS 256
#define GF_HP (GF_TW + 1)      // odd pitch (float4 units): conflict-free row-strided STS.128

struct GfGeom {
    int H, W, Wp;              // image size, padded target width
    int k, a;                  // window side, anchor
    int x0_base, x0_step;      // target crop column = x0_base + x0_step * di
    int D;                     // slices
};

// shared-memory h-sum helper: each thread sums runs of 8 adjacent outputs along a row with a sliding
// window; src row pitch is padded to avoid LDS.128 bank conflicts between threads on different rows.
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

template <int DC>
__global__ void __launch_bounds__(GF_THREADS)
k_gf_ab(const Feat* __restrict__ ref, const Feat* __restrict__ tgt, const float4* __restrict__ Gi,
        const float4* __restrict__ Gm, const float4* __restrict__ Gd, GfGeom g, TadParams tp,
        float4* __restrict__ ab /*[D][H][W]*/, uint32_t* __restrict__ slice_mm /*[D][2] orderable bits of c*/) {
    extern __shared__ float4 sm_gf[];
    const int k = g.k, a = g.a;
    const int IW = GF_TW + k - 1, IH = GF_TH + k - 1;
    const int PP = IW | 1;                              // odd pitch (in float4) -> conflict-free row-strided access
    float4* P = sm_gf;                                  // [IH][PP]   {c', I0c', I1c', I2c'}
    float4* Hs = sm_gf + (size_t)IH * PP;               // [IH][GF_HP] horizontal sums
    const int tid = threadIdx.x;
    const int x0t = blockIdx.x * GF_TW, y0t = blockIdx.y * GF_TH;
    const int d_begin = blockIdx.z * DC;
    const float inv = 1.0f / (float)(k * k);

    for (int dd = 0; dd < DC; dd++) {
        const int di = d_begin + dd;
        if (di >= g.D) break;
        const int xoff = g.x0_base + g.x0_step * di;
        // ---- phase A: products on tile + halo (REFLECT_101 coordinates) ----
        float cmin = 3.0e38f, cmax = -3.0e38f;
        for (int i = tid; i < IH * IW; i += GF_THREADS) {
            int r = i / IW, c = i - r * IW;
            int sy = border_idx(y0t - a + r, g.H, 1), sx = border_idx(x0t - a + c, g.W, 1);
            Feat fa = ref[(size_t)sy * g.W + sx];
            Feat fb = tgt[(size_t)sy * g.Wp + xoff + sx];
            float4 I = __ldg(&Gi[(size_t)sy * g.W + sx]);
            float cst = tad_cost(fa, fb, tp);
            cmin = fminf(cmin, cst); cmax = fmaxf(cmax, cst);
            float cp = __fsub_rn(cst, tp.c0);
            P[r * PP + c] = make_float4(cp, __fmul_rn(I.x, cp), __fmul_rn(I.y, cp), __fmul_rn(I.z, cp));
        }
        // slice min / max of the raw cost (halo positions are reflections of real pixels)
        for (int o = 16; o > 0; o >>= 1) {
            cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        }
        if ((tid & 31) == 0) {
            atomicMin(&slice_mm[2 * di], orderable_u32(cmin));
            atomicMax(&slice_mm[2 * di + 1], orderable_u32(cmax));
        }
        __syncthreads();
        // ---- phase B: horizontal k-sums, runs of 8 outputs; consecutive threads -> consecutive rows ----
        {
            const int runs_per_row = GF_TW / 8;
            for (int i = tid; i < IH * runs_per_row; i += GF_THREADS) {
                int run = i / IH, r = i - run * IH;
                const float4* src = P + r * PP + run * 8;
                float4 s = src[0];
                for (int j = 1; j < k; j++) s = f4add(s, src[j]);
                float4* dst = Hs + r * GF_HP + run * 8;
                dst[0] = s;
#pragma unroll
                for (int o = 1; o < 8; o++) {
                    s = f4add(f4sub(s, src[o - 1]), src[o - 1 + k]);
                    dst[o] = s;
                }
            }
        }
        __syncthreads();
        // ---- phase C: vertical k-sums + a, b; thread = (column, 8-row run) ----
        {
            const int col = tid % GF_TW, rrun = tid / GF_TW;        // 256 threads = 64 cols x 4 runs of 8 rows
            const int x = x0t + col;
            const float4* src = Hs + (rrun * 8) * GF_HP + col;
            float4 s = src[0];
            for (int j = 1; j < k; j++) s = f4add(s, src[j * GF_HP]);
#pragma unroll
            for (int o = 0; o < 8; o++) {
                if (o > 0) s = f4add(f4sub(s, src[(o - 1) * GF_HP]), src[(o - 1 + k) * GF_HP]);
                int y = y0t + rrun * 8 + o;
                if (x < g.W && y < g.H) {
                    size_t pix = (size_t)y * g.W + x;
                    float4 m = __ldg(&Gm[pix]);
                    float4 dn = __ldg(&Gd[pix]);
                    float mP = s.x * inv;
                    float a0 = __fdiv_rn(__fsub_rn(s.y * inv, __fmul_rn(m.x, mP)), dn.x);   // A.cpp:2805-2846
                    float a1 = __fdiv_rn(__fsub_rn(s.z * inv, __fmul_rn(m.y, mP)), dn.y);
                    float a2 = __fdiv_rn(__fsub_rn(s.w * inv, __fmul_rn(m.z, mP)), dn.z);
                    float dot = __fadd_rn(__fadd_rn(__fmul_rn(a0, m.x), __fmul_rn(a1, m.y)), __fmul_rn(a2, m.z));
                    float b = __fsub_rn(mP, dot);                                            // A.cpp:2847
                    ab[(size_t)di * g.H * g.W + pix] = make_float4(a0, a1, a2, b);
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// fast path, pass 2: second box level -> q, affine, WTA keys
// ------------------------------------------------------------------------------------------------
template <int DC>
__global__ void __launch_bounds__(GF_THREADS)
k_gf_q(const float4* __restrict__ ab, const float4* __restrict__ Gi, GfGeom g, float c0,
       const uint32_t* __restrict__ slice_mm, int d_first_label, unsigned long long* __restrict__ keys,
       float* __restrict__ agg /*optional [D][H][W]*/) {
    extern __shared__ float4 sm_gf[];
    const int k = g.k, a = g.a;
    const int IW = GF_TW + k - 1, IH = GF_TH + k - 1;
    const int PP = IW | 1;
    float4* P = sm_gf;
    float4* Hs = sm_gf + (size_t)IH * PP;
    const int tid = threadIdx.x;
    const int x0t = blockIdx.x * GF_TW, y0t = blockIdx.y * GF_TH;
    const int d_begin = blockIdx.z * DC;
    const float inv = 1.0f / (float)(k * k);
    const int col = tid % GF_TW, rrun = tid / GF_TW;
    const int x = x0t + col;
    unsigned long long best[8];
    float4 Ipix[8];
#pragma unroll
    for (int o = 0; o < 8; o++) {
        best[o] = WTA_KEY_EMPTY;
        int y = y0t + rrun * 8 + o;
        Ipix[o] = (x < g.W && y < g.H) ? __ldg(&Gi[(size_t)y * g.W + x]) : make_float4(0, 0, 0, 0);
    }
    for (int dd = 0; dd < DC; dd++) {
        const int di = d_begin + dd;
        if (di >= g.D) break;
        const float4* abd = ab + (size_t)di * g.H * g.W;
        for (int i = tid; i < IH * IW; i += GF_THREADS) {
            int r = i / IW, c = i - r * IW;
            int sy = border_idx(y0t - a + r, g.H, 1), sx = border_idx(x0t - a + c, g.W, 1);
            P[r * PP + c] = abd[(size_t)sy * g.W + sx];
        }
        // per-slice affine of cv::normalize (A.cpp:2775): q = sf * q' + (c0 * sf + hf)
        float sf, hf;
        minmax_scale_shift((double)from_orderable(slice_mm[2 * di]), (double)from_orderable(slice_mm[2 * di + 1]), &sf, &hf);
        float h2 = (float)fma((double)c0, (double)sf, (double)hf);
        __syncthreads();
        {
            const int runs_per_row = GF_TW / 8;
            for (int i = tid; i < IH * runs_per_row; i += GF_THREADS) {
                int run = i / IH, r = i - run * IH;
                const float4* src = P + r * PP + run * 8;
                float4 s = src[0];
                for (int j = 1; j < k; j++) s = f4add(s, src[j]);
                float4* dst = Hs + r * GF_HP + run * 8;
                dst[0] = s;
#pragma unroll
                for (int o = 1; o < 8; o++) {
                    s = f4add(f4sub(s, src[o - 1]), src[o - 1 + k]);
                    dst[o] = s;
                }
            }
        }
        __syncthreads();
        {
            const float4* src = Hs + (rrun * 8) * GF_HP + col;
            float4 s = src[0];
            for (int j = 1; j < k; j++) s = f4add(s, src[j * GF_HP]);
#pragma unroll
            for (int o = 0; o < 8; o++) {
                if (o > 0) s = f4add(f4sub(s, src[(o - 1) * GF_HP]), src[(o - 1 + k) * GF_HP]);
                float4 I = Ipix[o];
                float dot = __fadd_rn(__fadd_rn(__fmul_rn(s.x * inv, I.x), __fmul_rn(s.y * inv, I.y)), __fmul_rn(s.z * inv, I.z));
                float qc = __fadd_rn(dot, s.w * inv);                                         // A.cpp:2852
                float q = fmaf(qc, sf, h2);
                int y = y0t + rrun * 8 + o;
                if (agg && x < g.W && y < g.H) agg[((size_t)di * g.H + y) * g.W + x] = q;
                best[o] = min(best[o], wta_key(q, d_first_label + di));
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 0; o < 8; o++) {
        int y = y0t + rrun * 8 + o;
        if (x < g.W && y < g.H) atomicMin(&keys[(size_t)y * g.W + x], best[o]);
    }
}
