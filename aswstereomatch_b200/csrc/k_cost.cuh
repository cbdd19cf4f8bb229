// k_cost.cuh -- stage 1 raw costs: TAD colour+gradient (A.cpp:415-487) and box-SAD (A.cpp:2442-2503),
// and the stand-alone WTA (A.cpp:3032-3048).
#pragma once
#include "k_prep.cuh"

struct TadParams {
    double reg, reg_r;     // regularity, 1 - regularity            (A.cpp:435)
    float thr_c;           // colour threshold                       (A.cpp:462)
    float add_c;           // 255 * (float)(thresC/255): scaleAdd on u8 (SURVEY B-2)
    float thr_g;           // gradient threshold                     (A.cpp:475)
    float g_hi, g_lo;      // 254*thrG, 255*thrG                     (A.cpp:474-482)
    float c0;              // smallest possible cost: blend(0, 255*thrG)
};

static inline TadParams make_tad_params(double regularity, double thres_c, double thres_g) {
    TadParams p;
    p.reg = regularity;
    p.reg_r = 1 - regularity;
    p.thr_c = (float)thres_c;
    p.add_c = 255.0f * (float)(thres_c / 255.0);
    p.thr_g = (float)thres_g;
    p.g_hi = 254.0f * p.thr_g;
    p.g_lo = 255.0f * p.thr_g;
    p.c0 = (float)__builtin_fma(0.0, p.reg_r, (double)p.g_lo * p.reg);
    return p;
}

// cv::addWeighted on CV_32F (OpenCV 4.13): double scalars, (float)fma(a, alpha, b*beta)
__device__ __forceinline__ float add_weighted_f32(float a, double alpha, float b, double beta) {
    return (float)fma((double)a, alpha, __dmul_rn((double)b, beta));
}

// One disparity evaluation of computeSimilarity's 3-channel LEFT branch (A.cpp:455-484).
__device__ __forceinline__ float tad_cost(const Feat& a, const Feat& b, const TadParams& p) {
    uint32_t ad = __vabsdiffu4(a.bgr, b.bgr);                       // absdiff per channel (A.cpp:455)
    int c0 = ad & 0xFF, c1 = (ad >> 8) & 0xFF, c2 = (ad >> 16) & 0xFF;
    int s = min(c0 + c1, 255) + c2;                                 // (c0+c1) saturates, then +c2
    int color = ((s + 1) * 43691) >> 17;                            // round(s/3): addWeighted(1/3,1/3) u8
    float cc = 0.0f;
    if ((float)color > p.thr_c)                                     // A.cpp:461-465 (inverted truncation)
        cc = fminf(rintf((float)color + p.add_c), 255.0f);
    int ga0 = (int16_t)(a.g01 & 0xFFFF), ga1 = (int16_t)(a.g01 >> 16), ga2 = (int16_t)(a.g2 & 0xFFFF);
    int gb0 = (int16_t)(b.g01 & 0xFFFF), gb1 = (int16_t)(b.g01 >> 16), gb2 = (int16_t)(b.g2 & 0xFFFF);
    float g0 = (float)abs(ga0 - gb0), g1 = (float)abs(ga1 - gb1), g2 = (float)abs(ga2 - gb2);
    float gm1 = __fadd_rn(g0, g1);
    float G = add_weighted_f32(gm1, 1.0 / 3, g2, 1.0 / 3);         // A.cpp:473
    float gc = (G > p.thr_g) ? __fadd_rn(p.g_hi, G) : p.g_lo;       // A.cpp:474-482
    return add_weighted_f32(cc, p.reg_r, gc, p.reg);                // A.cpp:484
}

// fp32-only variant used inside the fused guided-filter kernels.  Returns c' = cost - c0 directly:
//   cost - c0 = regR*Cc + reg*(Gc - 255*T_G) = regR*Cc + reg*max(G - T_G, 0)        (A.cpp:474-484)
// The colour part Cc is the same exact integer; G = (g0+g1+g2)/3 and the blend are evaluated in fp32 instead
// of OpenCV's double (differences <= 1 float ulp of the ~5100-offset cost, i.e. ~1e-7 of the slice range --
// three orders of magnitude inside the 1e-4 budget) and no float<->double conversions hit the XU pipe.
struct TadFast { float thr_c, add_c, thr_g, reg_r, reg; };
static inline TadFast make_tad_fast(const TadParams& p) {
    TadFast f; f.thr_c = p.thr_c; f.add_c = p.add_c; f.thr_g = p.thr_g; f.reg_r = (float)p.reg_r; f.reg = (float)p.reg;
    return f;
}
__device__ __forceinline__ float tad_cost_prime(uint32_t a_bgr, uint32_t a_g01, uint32_t a_g2, const Feat& b, const TadFast& p) {
    uint32_t ad = __vabsdiffu4(a_bgr, b.bgr);
    int c0 = ad & 0xFF, c1 = (ad >> 8) & 0xFF, c2 = (ad >> 16) & 0xFF;
    int s = min(c0 + c1, 255) + c2;
    int color = ((s + 1) * 43691) >> 17;
    float cc = ((float)color > p.thr_c) ? fminf(rintf((float)color + p.add_c), 255.0f) : 0.0f;
    int ga0 = (int16_t)(a_g01 & 0xFFFF), ga1 = (int16_t)(a_g01 >> 16), ga2 = (int16_t)(a_g2 & 0xFFFF);
    int gb0 = (int16_t)(b.g01 & 0xFFFF), gb1 = (int16_t)(b.g01 >> 16), gb2 = (int16_t)(b.g2 & 0xFFFF);
    int S = abs(ga0 - gb0) + abs(ga1 - gb1) + abs(ga2 - gb2);
    float G = (float)S * 0.33333334f;
    return fmaf(p.reg_r, cc, p.reg * fmaxf(G - p.thr_g, 0.0f));
}

// Raw TAD C+G volume [D][H][W]; x0_base + x0_step*di = column offset of the target crop
// (LEFT: max_off - offset, RIGHT: offset).
__global__ void k_cost_tad_volume(const Feat* __restrict__ ref, const Feat* __restrict__ tgt, int H, int W,
                                  int Wp, int x0_base, int x0_step, TadParams p, float* __restrict__ vol) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y, di = blockIdx.z;
    if (x >= W) return;
    int x0 = x0_base + x0_step * di;
    Feat a = ref[(size_t)y * W + x];
    Feat b = tgt[(size_t)y * Wp + x0 + x];
    vol[((size_t)di * H + y) * W + x] = tad_cost(a, b, p);
}

// gray absolute difference as float, one slice per blockIdx.z (A.cpp:2477-2478)
__global__ void k_gray_absdiff(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, int H, int W,
                               int Wp, int x0_base, int x0_step, float* __restrict__ vol) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y, di = blockIdx.z;
    if (x >= W) return;
    int x0 = x0_base + x0_step * di;
    int v = (int)ref[(size_t)y * W + x] - (int)tgt[(size_t)y * Wp + x0 + x];
    vol[((size_t)di * H + y) * W + x] = (float)abs(v);
}

// Fused getCostSAD_d (A.cpp:2477-2480) for every d: gray |L - R_d| -> normalised win x win boxFilter,
// BORDER_REFLECT_101.  The absolute differences are integers, so the window sum is exact in int32 whatever the
// order, and (float)(sum * (1/win^2)) in double is what cv::boxFilter returns.  One thread per halo column walks
// down a band of rows keeping the vertical window sum in a register (one row in, one row out); the
// horizontal sums run over shared memory, RB rows per barrier.  Nothing but the result reaches HBM.
#define SADBOX_COLS 256
#define SADBOX_PITCH 264         // row pitch of the staged vertical sums (8 pad columns for the last quad's loads)
// WIN_T > 0: window side as a template constant (quad path, fully static indexing); WIN_T = 0: any window
template <int RB, int WIN_T>
__global__ void __launch_bounds__(SADBOX_COLS)
k_sad_box_u8(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ tgt, int H, int W, int Wp, int x0_base,
             int x0_step, int win_rt, int band_rows, float* __restrict__ vol) {
    const int win = WIN_T > 0 ? WIN_T : win_rt;
    __shared__ __align__(16) int vs[2][RB][SADBOX_PITCH];
    const int h = win / 2, SW = SADBOX_COLS - (win - 1);
    const int cx = threadIdx.x, x0 = blockIdx.x * SW, di = blockIdx.z;
    const int sx = border_idx(x0 - h + cx, W, 1);
    const uint8_t* rc = ref + sx;
    const uint8_t* tc = tgt + x0_base + x0_step * di + sx;
    const int y_begin = blockIdx.y * band_rows, y_end = min(H, y_begin + band_rows);
    auto ad = [&](int y) {
        int sy = border_idx(y, H, 1);
        return abs((int)rc[(size_t)sy * W] - (int)tc[(size_t)sy * Wp]);
    };
    int s = 0;
    for (int j = -h; j < h; j++) s += ad(y_begin + j);
    const double scale = 1.0 / ((double)win * win);
    // horizontal phase: one thread per (row of the batch, quad of 4 adjacent output columns): the 4 + win - 1 vertical
    // sums it needs arrive as aligned 16-byte loads, the 4 window sums slide over them (templated windows; any other
    // window takes one output per thread)
    constexpr bool quads = WIN_T > 0;
    const int nquad = (SW + 3) / 4;
    const int hr = cx / nquad, hq = cx - hr * nquad;
    float* plane = vol + (size_t)di * H * W;
    int buf = 0;
    // the RB entering and RB leaving differences of a batch are requested one batch ahead: their latency hides behind the
    // horizontal phase instead of stalling the running sum
    int ain[RB], aout[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) { ain[r] = ad(y_begin + r + h); aout[r] = ad(y_begin + r - h); }
    for (int y = y_begin; y < y_end; y += RB) {
#pragma unroll
        for (int r = 0; r < RB; r++) {
            s += ain[r];
            vs[buf][r][cx] = s;
            s -= aout[r];
        }
        if (y + RB < y_end) {
#pragma unroll
            for (int r = 0; r < RB; r++) { ain[r] = ad(y + RB + r + h); aout[r] = ad(y + RB + r - h); }
        }
        __syncthreads();
        if (quads) {
            if (hr < RB && y + hr < y_end) {
                constexpr int NV = ((WIN_T + 3 + 3) / 4) * 4;             // 4 + WIN - 1 values, rounded up to whole int4
                int v[NV > 0 ? NV : 4];
                const int4* src = (const int4*)&vs[buf][hr][4 * hq];
#pragma unroll
                for (int m = 0; m < NV / 4; m++) { const int4 t = src[m]; v[4 * m] = t.x; v[4 * m + 1] = t.y; v[4 * m + 2] = t.z; v[4 * m + 3] = t.w; }
                int acc = 0;
#pragma unroll
                for (int j = 0; j < WIN_T; j++) acc += v[j];
                float* out = plane + (size_t)(y + hr) * W + x0 + 4 * hq;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (i > 0) acc += v[i - 1 + WIN_T] - v[i - 1];
                    if (4 * hq + i < SW && x0 + 4 * hq + i < W) out[i] = (float)((double)acc * scale);
                }
            }
        } else if (cx < SW && x0 + cx < W) {
#pragma unroll
            for (int r = 0; r < RB; r++) {
                if (y + r < y_end) {
                    int acc = 0;
                    for (int j = 0; j < win; j++) acc += vs[buf][r][cx + j];
                    plane[(size_t)(y + r) * W + x0 + cx] = (float)((double)acc * scale);
                }
            }
        }
        buf ^= 1;
    }
}

// WTA over a materialised volume -> 64-bit keys (strict <, ascending d, NaN/inf never win).  One thread = 4 adjacent pixels
// (16-byte loads where the slice pitch allows); the D loads of a thread are independent: 8 slices = 128 bytes in flight per
// thread, streaming (evict-first) loads.
__global__ void __launch_bounds__(256)
k_wta_keys(const float* __restrict__ vol, int D, size_t n, int d_first, unsigned long long* __restrict__ keys) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float INF = __int_as_float(0x7f800000);          // a cost of +inf (or NaN) never wins: the key stays empty
    float best[4] = {INF, INF, INF, INF};
    int bd[4] = {-1, -1, -1, -1};
    const bool vec = (n & 3) == 0;                           // every slice starts 16-byte aligned
    const int cnt = (int)min((size_t)4, n - i);
    for (int d0 = 0; d0 < D; d0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (d0 + u < D) {
                const float* p = vol + (size_t)(d0 + u) * n + i;
                if (vec) v[u] = __ldcs((const float4*)p);
                else { v[u].x = p[0]; v[u].y = cnt > 1 ? p[1] : INF; v[u].z = cnt > 2 ? p[2] : INF; v[u].w = cnt > 3 ? p[3] : INF; }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (d0 + u < D) {
                const float q[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; k++) if (q[k] < best[k]) { best[k] = q[k]; bd[k] = d0 + u; }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < cnt && bd[k] >= 0) keys[i + k] = min(keys[i + k], wta_key(best[k], d_first + bd[k]));
}
__global__ void k_fill_u64(unsigned long long* __restrict__ p, size_t n, unsigned long long v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// keys -> float disparity; a cost of +inf or NaN never beats DBL_MAX in the reference: sentinel 0
__global__ void k_keys_to_disp(const unsigned long long* __restrict__ keys, size_t n, float* __restrict__ disp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = keys[i];
    disp[i] = (k >= WTA_KEY_INF_TOP) ? 0.0f : (float)(uint32_t)(k & 0xFFFFull);
}
__global__ void k_keys_min_merge(unsigned long long* __restrict__ a, const unsigned long long* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = min(a[i], b[i]);
}
