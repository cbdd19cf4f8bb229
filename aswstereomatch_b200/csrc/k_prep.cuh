// k_prep.cuh -- image preparation kernels shared by the methods:
//   * feature records (BGR + Scharr-x gradient) for the TAD C+G cost (A.cpp:442-450)
//   * BGR2GRAY (OpenCV >= 4.2 fixed point), column REFLECT padding
//   * generic double-accumulated box filter (cv::boxFilter 32F semantics, SURVEY B-9)
//   * min/max reductions
#pragma once
#include "asw_common.cuh"

// One 16-byte record per pixel: x = B | G<<8 | R<<16, y = gB (lo16) | gG (hi16), z = gR (lo16).
// Gradients are exact integers in [-4080, 4080] (16 * 255) so int16 holds them.
struct __align__(16) Feat {
    uint32_t bgr;
    uint32_t g01;
    uint32_t g2;
    uint32_t pad;
};

// Build records for an image padded on the column axis with BORDER_REFLECT (pad_l / pad_r columns);
// the gradient is filter2D([-3 0 3; -10 0 10; -3 0 3]) of the PADDED image with BORDER_REFLECT_101
// (A.cpp:446-450 applies filter2D to right_border).
// float form of the record for the streaming guided kernel: exact-integer gradients (optionally negated) + BGR
struct __align__(16) FeatF { float g0, g1, g2; uint32_t bgr; };

// FLOAT_OUT = false: Feat records; true: FeatF records with the gradients multiplied by `sign` (the streaming
// kernel adds the target's negated gradients to the reference's).
template <bool FLOAT_OUT>
__device__ __forceinline__ void features_body(const uint8_t* __restrict__ img, int H, int W, int pad_l, int pad_r,
                                              void* __restrict__ out_v, float sign) {
    int Wp = W + pad_l + pad_r;
    int xp = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (xp >= Wp) return;
    int ys[3] = {border_idx(y - 1, H, 1), y, border_idx(y + 1, H, 1)};
    int xs[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int xq = border_idx(xp - 1 + i, Wp, 1);       // REFLECT_101 inside the padded image
        xs[i] = border_idx(xq - pad_l, W, 0);         // padded column -> source column (REFLECT)
    }
    int g[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int v = 0;
        const int wgt[3] = {3, 10, 3};
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const uint8_t* row = img + (size_t)ys[r] * W * 3;
            v += wgt[r] * ((int)row[xs[2] * 3 + c] - (int)row[xs[0] * 3 + c]);
        }
        g[c] = v;
    }
    const uint8_t* px = img + ((size_t)y * W + xs[1]) * 3;
    const uint32_t bgr = (uint32_t)px[0] | ((uint32_t)px[1] << 8) | ((uint32_t)px[2] << 16);
    if (FLOAT_OUT) {
        FeatF f;
        f.g0 = sign * (float)g[0]; f.g1 = sign * (float)g[1]; f.g2 = sign * (float)g[2];
        f.bgr = bgr;
        ((FeatF*)out_v)[(size_t)y * Wp + xp] = f;
    } else {
        Feat f;
        f.bgr = bgr;
        f.g01 = ((uint32_t)(uint16_t)(int16_t)g[0]) | ((uint32_t)(uint16_t)(int16_t)g[1] << 16);
        f.g2 = (uint32_t)(uint16_t)(int16_t)g[2];
        f.pad = 0;
        ((Feat*)out_v)[(size_t)y * Wp + xp] = f;
    }
}
__global__ void k_features(const uint8_t* __restrict__ img, int H, int W, int pad_l, int pad_r, Feat* __restrict__ out) {
    features_body<false>(img, H, W, pad_l, pad_r, out, 1.0f);
}
__global__ void k_features_f(const uint8_t* __restrict__ img, int H, int W, int pad_l, int pad_r, float sign,
                             FeatF* __restrict__ out) {
    features_body<true>(img, H, W, pad_l, pad_r, out, sign);
}

// cvtColor(BGR2GRAY) u8: (3735 B + 19235 G + 9798 R + 2^14) >> 15, with optional REFLECT column padding
__global__ void k_bgr2gray_pad(const uint8_t* __restrict__ img, int H, int W, int pad_l, int pad_r,
                               uint8_t* __restrict__ out) {
    int Wp = W + pad_l + pad_r;
    int xp = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (xp >= Wp) return;
    int x = border_idx(xp - pad_l, W, 0);
    const uint8_t* px = img + ((size_t)y * W + x) * 3;
    out[(size_t)y * Wp + xp] = (uint8_t)((3735 * px[0] + 19235 * px[1] + 9798 * px[2] + (1 << 14)) >> 15);
}

// min / max of a u8 buffer -> mm[0] = min, mm[1] = max (ints, pre-initialised to 255 / 0)
// blockIdx.y = slice of a batch (n bytes per slice, one min / max pair per slice)
__global__ void k_minmax_u8(const uint8_t* __restrict__ src, size_t n, int* __restrict__ mm) {
    src += (size_t)blockIdx.y * n; mm += 2 * blockIdx.y;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    // 16 bytes per load, byte-wise min / max on packed words; the unaligned head and the tail byte by byte
    size_t head = (16 - ((size_t)(uintptr_t)src & 15)) & 15;
    if (head > n) head = n;
    const size_t nvec = (n - head) / 16;
    const uint4* v = (const uint4*)(src + head);
    unsigned int mn4 = 0xFFFFFFFFu, mx4 = 0u;
    for (size_t i = tid; i < nvec; i += nt) {
        const uint4 q = __ldg(&v[i]);
        mn4 = __vminu4(mn4, __vminu4(__vminu4(q.x, q.y), __vminu4(q.z, q.w)));
        mx4 = __vmaxu4(mx4, __vmaxu4(__vmaxu4(q.x, q.y), __vmaxu4(q.z, q.w)));
    }
    int mn = (int)min(min(mn4 & 0xFF, (mn4 >> 8) & 0xFF), min((mn4 >> 16) & 0xFF, mn4 >> 24));
    int mx = (int)max(max(mx4 & 0xFF, (mx4 >> 8) & 0xFF), max((mx4 >> 16) & 0xFF, mx4 >> 24));
    for (size_t i = tid; i < head; i += nt) { const int b = src[i]; mn = min(mn, b); mx = max(mx, b); }
    for (size_t i = head + nvec * 16 + tid; i < n; i += nt) { const int b = src[i]; mn = min(mn, b); mx = max(mx, b); }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], mn); atomicMax(&mm[1], mx); }
}

// per-slice min / max of a float volume [D][n]; mm[2*d] = min bits, mm[2*d+1] = max bits using the
// orderable-uint encoding (works for any sign).  mm pre-initialised to 0xFFFFFFFF / 0.
__global__ void k_minmax_f32_slices(const float* __restrict__ vol, size_t n, uint32_t* __restrict__ mm) {
    int d = blockIdx.y;
    const float* s = vol + (size_t)d * n;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    auto take = [&](float v) { if (v == v) { uint32_t o = orderable_u32(v); mn = min(mn, o); mx = max(mx, o); } };
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    // 16-byte loads over the aligned body of the slice, the unaligned head and the tail element by element
    size_t head = ((16 - ((size_t)(uintptr_t)s & 15)) & 15) / 4;
    if (head > n) head = n;
    const size_t nvec = (n - head) / 4;
    const float4* v4 = (const float4*)(s + head);
    for (size_t i = tid; i < nvec; i += nt) { const float4 q = __ldg(&v4[i]); take(q.x); take(q.y); take(q.z); take(q.w); }
    for (size_t i = tid; i < head; i += nt) take(s[i]);
    for (size_t i = head + nvec * 4 + tid; i < n; i += nt) take(s[i]);
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * d], mn); atomicMax(&mm[2 * d + 1], mx); }
}
__device__ __forceinline__ float from_orderable(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(b);
}

// cv::normalize(NORM_MINMAX, 0..1, CV_32F) scalars (SURVEY B-10): sf = (float)(1/(max-min)),
// hf = 0f - (float)(min * sf)
__device__ __forceinline__ void minmax_scale_shift(double mn, double mx, float* sf, float* hf) {
    double scale = (mx - mn > 2.220446049250313e-16) ? 1.0 / (mx - mn) : 0.0;
    *sf = (float)scale;
    *hf = 0.0f - (float)(mn * (double)(*sf));
}

// Generic box filter, one plane per blockIdx.z: cv::boxFilter(32F -> 32F, normalised, anchor k/2,
// BORDER_REFLECT_101) with double accumulation (direct separable sums; OpenCV uses running double
// sums, the two agree to the last float ulp except on rare pixels).  Tile 32x8 outputs per block.
template <int TW, int TH>
__global__ void k_box_f32(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int k,
                          size_t plane_stride) {
    extern __shared__ double sm_box[];
    int a = k / 2;
    int IW = TW + k - 1, IH = TH + k - 1;
    double* tile = sm_box;                  // [IH][IW]
    double* hs = sm_box + (size_t)IH * IW;  // [IH][TW]
    const float* s = src + (size_t)blockIdx.z * plane_stride;
    float* o = dst + (size_t)blockIdx.z * plane_stride;
    int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    for (int i = tid; i < IH * IW; i += nt) {
        int r = i / IW, c = i % IW;
        int sy = border_idx(y0 - a + r, H, 1), sx = border_idx(x0 - a + c, W, 1);
        tile[i] = (double)s[(size_t)sy * W + sx];
    }
    __syncthreads();
    for (int i = tid; i < IH * TW; i += nt) {
        int r = i / TW, c = i % TW;
        double acc = 0;
        for (int j = 0; j < k; j++) acc += tile[r * IW + c + j];
        hs[i] = acc;
    }
    __syncthreads();
    double scale = 1.0 / ((double)k * k);
    for (int i = tid; i < TH * TW; i += nt) {
        int r = i / TW, c = i % TW;
        int x = x0 + c, y = y0 + r;
        if (x < W && y < H) {
            double acc = 0;
            for (int j = 0; j < k; j++) acc += hs[(r + j) * TW + c];
            o[(size_t)y * W + x] = (float)(acc * scale);
        }
    }
}

// Streaming form of the same filter (the default): one thread per halo column of a 128-column strip walks down a band of
// rows with the vertical window sum in a double register (one row in, one row out), the horizontal sums run over shared
// memory.  Sums of <= k^2 floats are exact in double for any data whose exponents span < 29 bits, so the summation order
// does not change the rounded float (same argument as k_gfs_guide_moments); 2 global loads and k + 2 double adds per output
// instead of the tiled kernel's ~4 k, and no (k-1)-row halo per 8 output rows.
#define BOXS_COLS 128
#define BOXS_RUN 7
#define BOXS_ROWS 4
__global__ void __launch_bounds__(BOXS_COLS)
k_box_f32_stream(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int k, size_t plane_stride, int band_rows) {
    __shared__ double vs[2][BOXS_ROWS][BOXS_COLS];
    __shared__ float os[2][BOXS_ROWS][BOXS_COLS];
    const int a = k / 2, SW = BOXS_COLS - (k - 1);
    const int cx = threadIdx.x, x0 = blockIdx.x * SW;
    const int sx = border_idx(x0 - a + cx, W, 1);
    const float* s = src + (size_t)blockIdx.z * plane_stride + sx;
    float* o = dst + (size_t)blockIdx.z * plane_stride;
    const int y_begin = blockIdx.y * band_rows, y_end = min(H, y_begin + band_rows);
    double V = 0.0;
    for (int j = -a; j < a; j++) V += (double)s[(size_t)border_idx(y_begin + j, H, 1) * W];
    const double scale = 1.0 / ((double)k * k);
    const int xo = x0 + cx;
    const bool writer = cx < SW && xo < W;
    int buf = 0;
    // BOXS_ROWS = 4 rows per barrier (measured for GuidedF at 640x360x64, box_f32 of 3 launches: 1 row 2.23 ms, 4 rows 1.42, 8 rows 1.50, 12 rows 1.71).  Vertical: a thread advances its column sum over the 4 rows (their 8 loads are issued
    // together).  Horizontal: warp r takes row r; lane l slides over BOXS_RUN = 7 adjacent outputs (k - 1 + 7 loads for 7 outputs;
    // an odd run keeps the 8-byte loads of a half-warp on distinct banks).  Every partial sum is exact in double, so neither the
    // grouping nor the sliding update changes a result.  The results go through shared memory and are written to global memory
    // by the column threads (coalesced) after the next group's barrier.
    const int hr = cx >> 5, c0 = (cx & 31) * BOXS_RUN;
    for (int y = y_begin; y < y_end; y += BOXS_ROWS) {
        float fin[BOXS_ROWS], fout[BOXS_ROWS];
#pragma unroll
        for (int r = 0; r < BOXS_ROWS; r++) {
            fin[r] = s[(size_t)border_idx(y + r + a, H, 1) * W];
            fout[r] = s[(size_t)border_idx(y + r - a, H, 1) * W];
        }
#pragma unroll
        for (int r = 0; r < BOXS_ROWS; r++) {
            V += (double)fin[r];
            vs[buf][r][cx] = V;
            V -= (double)fout[r];
        }
        __syncthreads();
        if (y > y_begin && writer) {
#pragma unroll
            for (int r = 0; r < BOXS_ROWS; r++) o[(size_t)(y - BOXS_ROWS + r) * W + xo] = os[buf ^ 1][r][cx];
        }
        if (hr < BOXS_ROWS && y + hr < y_end && c0 < SW) {
            const double* w = &vs[buf][hr][c0];
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = 0;
            for (; j + 4 <= k; j += 4) { a0 += w[j]; a1 += w[j + 1]; a2 += w[j + 2]; a3 += w[j + 3]; }
            for (; j < k; j++) a0 += w[j];
            double sum = (a0 + a1) + (a2 + a3);
            os[buf][hr][c0] = (float)(sum * scale);
#pragma unroll
            for (int g = 1; g < BOXS_RUN; g++) {
                if (c0 + g < SW) {
                    sum += w[k - 1 + g] - w[g - 1];
                    os[buf][hr][c0 + g] = (float)(sum * scale);
                }
            }
        }
        buf ^= 1;
    }
    __syncthreads();
    if (y_end > y_begin && writer) {
        const int ylast = y_begin + ((y_end - y_begin - 1) / BOXS_ROWS) * BOXS_ROWS;     // first row of the last group
#pragma unroll
        for (int r = 0; r < BOXS_ROWS; r++) if (ylast + r < y_end) o[(size_t)(ylast + r) * W + xo] = os[buf ^ 1][r][cx];
    }
}

static inline asw_status launch_box_f32(asw_ctx* ctx, const float* src, float* dst, int H, int W, int k,
                                        int planes, size_t plane_stride) {
    if (k - 1 <= BOXS_COLS / 2 && !asw_dev("ASW_BOX_TILED")) {
        const int strips = cdiv(W, BOXS_COLS - (k - 1));
        // bands: a band is entered once (k - 1 warm-up rows); as few as fill the GPU ~4 times over
        int bands = std::max(1, std::min(cdiv(H, 4 * k), cdiv(4 * ctx->sm_count * 8, std::max(1, strips * planes))));
        const int band_rows = cdiv(H, bands);
        dim3 grid(strips, cdiv(H, band_rows), planes);
        LAUNCH(ctx, "box_f32", (k_box_f32_stream<<<grid, BOXS_COLS, 0, ctx->stream>>>(src, dst, H, W, k, plane_stride, band_rows)));
        return ASW_OK;
    }
    const int TW = 32, TH = 8;
    size_t smem = ((size_t)(TH + k - 1) * (TW + k - 1) + (size_t)(TH + k - 1) * TW) * sizeof(double);
    if (smem > 200 * 1024) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "box window too large%s%s");
    cudaFuncSetAttribute(k_box_f32<TW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid(cdiv(W, TW), cdiv(H, TH), planes), block(32, 8);
    LAUNCH(ctx, "box_f32", (k_box_f32<TW, TH><<<grid, block, smem, ctx->stream>>>(src, dst, H, W, k, plane_stride)));
    return ASW_OK;
}

// copyMakeBorder(img, 0, 0, pad_l, pad_r, BORDER_REFLECT) for BGR u8
__global__ void k_pad_cols_bgr(const uint8_t* __restrict__ img, int H, int W, int pad_l, int pad_r,
                               uint8_t* __restrict__ out) {
    int Wp = W + pad_l + pad_r;
    int xp = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xp >= Wp) return;
    int x = border_idx(xp - pad_l, W, 0);
    const uint8_t* s = img + ((size_t)y * W + x) * 3;
    uint8_t* o = out + ((size_t)y * Wp + xp) * 3;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
}
