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
    std::vector<float> table;
    trad_build_table(win, gamma_c, gamma_g, table);
    float* dtable;
    ASW_TRY(ws_get(ctx, WS_TABLE0, table.size(), &dtable));
    ASW_CUDA(ctx, cudaMemcpyAsync(dtable, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // table is a host temporary
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
    dim3 grid(cdiv(W, 128), H, cdiv(g.n_cand, TRAD_Q));
    LAUNCH(ctx, "trad_aggregate", (k_trad_aggregate<<<grid, 128, 0, ctx->stream>>>(ref, tgt, dtable, g, keys, agg_dev)));
    return keys_to_disp(ctx, keys, n, disp_dev);
}
