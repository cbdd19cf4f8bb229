// k_preproc.cuh -- the reference DRIVER's per-image pre-processing on the device, so that a raw
// frame goes to HBM once and the matcher's inputs never return to the host:
//     main.cpp:30-31   resize(img, img, Size(640, 360))                                   (INTER_LINEAR, CV_8UC3)
//     main.cpp:67-89   cvtColor(BGR2HSV); bilateralFilter(V, blur, 7, 10, 3, BORDER_REFLECT);
//                      detail = V - blur; V = V + detail * 2; cvtColor(HSV2BGR)
// These are OpenCV library calls; the arithmetic below is OpenCV 4.13's 8-bit arithmetic as pinned against the real cv2 by
// oracle/preproc.py (resize and BGR2HSV bit-exact; the bilateral sum with FMA like cv2's vector body; HSV2BGR with the
// TRUNCATION of cv2's vector body -- cv2's scalar row tail rounds instead, so cv2's own bytes depend on the build's vector
// width; see the header of oracle/preproc.py for the measured agreement).
#pragma once
#include "asw_common.cuh"

struct PpResize { int sh, sw, dh, dw, area2; double scale_x, scale_y; };

// source index and the two 11-bit weights of a destination index (resize.cpp, INTER_LINEAR, 8u): columns clamp the index AND
// the weights at the borders, rows keep the weights and clip the row index
__device__ __forceinline__ void pp_lin_coeff(int d, double scale, int sn, bool clamp, int* s_out, int* a0, int* a1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp && s < 0) { s = 0; f = 0.0f; }
    if (clamp && s >= sn - 1) { s = sn - 1; f = 0.0f; }
    *a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    *a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    *s_out = s;
}

// one thread per destination pixel (3 channels)
__global__ void k_pp_resize(const uint8_t* __restrict__ src, PpResize g, uint8_t* __restrict__ dst) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
    if (dx >= g.dw) return;
    uint8_t* o = dst + ((size_t)dy * g.dw + dx) * 3;
    if (g.area2) {      // INTER_LINEAR with an exact 2 x 2 downscale runs INTER_AREA: rounded mean of the 2 x 2 block
        const uint8_t* p0 = src + ((size_t)(2 * dy) * g.sw + 2 * dx) * 3;
        const uint8_t* p1 = p0 + (size_t)g.sw * 3;
#pragma unroll
        for (int c = 0; c < 3; c++) o[c] = (uint8_t)(((int)p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
        return;
    }
    int xo, a0, a1, yo, b0, b1;
    pp_lin_coeff(dx, g.scale_x, g.sw, true, &xo, &a0, &a1);
    pp_lin_coeff(dy, g.scale_y, g.sh, false, &yo, &b0, &b1);
    const int x1 = min(xo + 1, g.sw - 1);
    const int y0 = min(max(yo, 0), g.sh - 1), y1 = min(max(yo + 1, 0), g.sh - 1);
    const uint8_t* r0 = src + (size_t)y0 * g.sw * 3;
    const uint8_t* r1 = src + (size_t)y1 * g.sw * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int h0 = (int)r0[xo * 3 + c] * a0 + (int)r0[x1 * 3 + c] * a1;      // horizontal pass, 11 fractional bits
        const int h1 = (int)r1[xo * 3 + c] * a0 + (int)r1[x1 * 3 + c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(max(v, 0), 255);
    }
}

// cvtColor(COLOR_BGR2HSV), CV_8UC3: 12-bit fixed-point division tables, H in 0..179
__global__ void k_pp_bgr2hsv(const uint8_t* __restrict__ bgr, size_t n, uint8_t* __restrict__ hsv) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    const int v = max(max(b, g), r), diff = v - min(min(b, g), r);
    const int sdiv = v ? __double2int_rn((double)(255 << 12) / (double)v) : 0;
    const int hdiv = diff ? __double2int_rn((double)(180 << 12) / (6.0 * (double)diff)) : 0;
    const int s = (diff * sdiv + (1 << 11)) >> 12;
    int h = (v == r) ? g - b : ((v == g) ? b - r + 2 * diff : r - g + 4 * diff);
    h = (h * hdiv + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    hsv[3 * i] = (uint8_t)h; hsv[3 * i + 1] = (uint8_t)s; hsv[3 * i + 2] = (uint8_t)v;
}

#define PP_MAX_TAPS 81
struct PpBilateral {
    int ntaps;
    signed char dy[PP_MAX_TAPS], dx[PP_MAX_TAPS];
    float space_w[PP_MAX_TAPS];
    float color_w[256];
};
// BORDER_REFLECT (edge pixel repeated): fedcba|abcdefgh|hgfedcb
__device__ __forceinline__ int pp_reflect(int p, int n) {
    const int period = 2 * n;
    int q = p % period;
    if (q < 0) q += period;
    return q >= n ? period - 1 - q : q;
}

// bilateralFilter on the V plane + detail boost + HSV2BGR, one thread per pixel
__global__ void k_pp_boost(const uint8_t* __restrict__ hsv, int H, int W, const PpBilateral* __restrict__ tb, uint8_t* __restrict__ out) {
    __shared__ float color_w[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) color_w[i] = tb->color_w[i];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t p = (size_t)y * W + x;
    const int hh = hsv[3 * p], ss = hsv[3 * p + 1], v0 = hsv[3 * p + 2];
    float acc = 0.0f, wsum = 0.0f;
    const int nt = tb->ntaps;
    for (int k = 0; k < nt; k++) {
        const int yy = pp_reflect(y + tb->dy[k], H), xx = pp_reflect(x + tb->dx[k], W);
        const int val = hsv[3 * ((size_t)yy * W + xx) + 2];
        const float w = __fmul_rn(tb->space_w[k], color_w[abs(val - v0)]);
        acc = fmaf((float)val, w, acc);                          // cv2's vector body: v_muladd
        wsum = __fadd_rn(wsum, w);
    }
    const int blur = min(max(__float2int_rn(__fdiv_rn(acc, wsum)), 0), 255);
    const int detail = max(v0 - blur, 0);                        // V - blur, saturating                 main.cpp:76
    const int v2 = min(v0 + 2 * detail, 255);                    // V + detail * 2, one saturating scaled add  main.cpp:77
    // HSV2BGR (float, H * 6/180, S/255, V/255, sector tables of color_hsv), bytes truncated like cv2's vector body
    const float h = __fmul_rn((float)hh, 6.0f / 180.0f);
    const float s = __fmul_rn((float)ss, 1.0f / 255.0f), v = __fmul_rn((float)v2, 1.0f / 255.0f);
    float b, g, r;
    if (ss == 0) {
        b = g = r = v;
    } else {
        int sector = (int)floorf(h);
        const float hf = __fsub_rn(h, (float)sector);
        if (sector >= 6) sector -= 6;
        const float sh = __fmul_rn(s, hf), one_s = __fsub_rn(1.0f, s);
        const float t0 = v, t1 = __fmul_rn(v, one_s), t2 = __fmul_rn(v, __fsub_rn(1.0f, sh)), t3 = __fmul_rn(v, __fadd_rn(one_s, sh));
        switch (sector) {                                        // (b, g, r) slots: {1,3,0} {1,0,2} {3,0,1} {0,2,1} {0,1,3} {2,1,0}
            case 0: b = t1; g = t3; r = t0; break;
            case 1: b = t1; g = t0; r = t2; break;
            case 2: b = t3; g = t0; r = t1; break;
            case 3: b = t0; g = t2; r = t1; break;
            case 4: b = t0; g = t1; r = t3; break;
            default: b = t2; g = t1; r = t0; break;
        }
    }
    out[3 * p] = (uint8_t)min(max((int)floorf(__fmul_rn(b, 255.0f)), 0), 255);
    out[3 * p + 1] = (uint8_t)min(max((int)floorf(__fmul_rn(g, 255.0f)), 0), 255);
    out[3 * p + 2] = (uint8_t)min(max((int)floorf(__fmul_rn(r, 255.0f)), 0), 255);
}

// host side of the tap table: taps inside the radius circle, row-major; weights as cv2 builds them ((float)std::exp(double))
static void pp_build_bilateral(int d, double sigma_color, double sigma_space, PpBilateral* t) {
    int radius = d > 0 ? d / 2 : (int)lrint(sigma_space * 1.5);
    radius = std::max(radius, 1);
    radius = std::min(radius, 4);                                // PP_MAX_TAPS = 9 x 9
    const double gcc = -0.5 / (sigma_color * sigma_color), gsc = -0.5 / (sigma_space * sigma_space);
    for (int i = 0; i < 256; i++) t->color_w[i] = (float)std::exp((double)i * i * gcc);
    t->ntaps = 0;
    for (int i = -radius; i <= radius; i++)
        for (int j = -radius; j <= radius; j++) {
            const double r = std::sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            t->dy[t->ntaps] = (signed char)i; t->dx[t->ntaps] = (signed char)j;
            t->space_w[t->ntaps++] = (float)std::exp(r * r * gsc);
        }
}

// src_dev: raw frame [sh][sw][3] in device memory; dst_dev: [dh][dw][3].  Runs on ctx->stream.
static asw_status dev_preprocess(asw_ctx* ctx, const uint8_t* src_dev, int sh, int sw, uint8_t* dst_dev, int dh, int dw) {
    uint8_t *small, *hsv;
    PpBilateral* tb;
    const size_t n = (size_t)dh * dw;
    ASW_TRY(ws_get(ctx, WS_TMP2, n * 3, &small));
    ASW_TRY(ws_get(ctx, WS_TMP3, n * 3, &hsv));
    ASW_TRY(ws_get(ctx, WS_TABLE1, (size_t)1, &tb));
    if (!table_cached(ctx, 4, WS_TABLE1, "bilateral:7:10:3")) {   // the driver's literals (main.cpp:74)
        static PpBilateral host_tb;                               // lives as long as the library: the copy below is asynchronous
        pp_build_bilateral(7, 10.0, 3.0, &host_tb);
        ASW_CUDA(ctx, cudaMemcpyAsync(tb, &host_tb, sizeof(PpBilateral), cudaMemcpyHostToDevice, ctx->stream));
    }
    PpResize g;
    g.sh = sh; g.sw = sw; g.dh = dh; g.dw = dw;
    g.scale_x = 1.0 / ((double)dw / sw); g.scale_y = 1.0 / ((double)dh / sh);
    g.area2 = (sw == 2 * dw && sh == 2 * dh) ? 1 : 0;
    LAUNCH(ctx, "pp_resize", (k_pp_resize<<<dim3(cdiv(dw, 128), dh), 128, 0, ctx->stream>>>(src_dev, g, small)));
    LAUNCH(ctx, "pp_bgr2hsv", (k_pp_bgr2hsv<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(small, n, hsv)));
    LAUNCH(ctx, "pp_boost", (k_pp_boost<<<dim3(cdiv(dw, 128), dh), 128, 0, ctx->stream>>>(hsv, dh, dw, tb, dst_dev)));
    return ASW_OK;
}
