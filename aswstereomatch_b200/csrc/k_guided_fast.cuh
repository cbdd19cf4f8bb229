// k_guided_fast.cuh -- tuned guided-filter kernels, window size K a template constant (K = 5, 7, 9 are
// instantiated; configs 2 and 5 use K = 9; other windows take the generic kernels of k_guided.cuh).
//
// Same algorithm as k_gf_ab / k_gf_q; what changes is the mapping onto the SM:
//   * the input tile is fixed at 32 rows x 64 columns (outputs TH x TW = (33-K) x (65-K)); 8 warps, warp w owns
//     input rows 4w..4w+3, lane l owns input columns l and l+32 -> every thread owns the same 8 tile positions
//     for every disparity: no div/mod or border arithmetic in the loop.  256 threads x 128 registers x 2 CTAs
//     fill the register file exactly (a 288-thread CTA is accounted as 320 threads and wastes a tenth of it)
//   * everything that does not depend on the disparity is loaded ONCE per CTA and reused for all DC slices of
//     the chunk: reference-side feature records + normalised guidance + addresses in registers,
//     (mean_I, 1/(var_I+eps)) of the output pixels in shared memory.  Per disparity evaluation the only global
//     traffic of pass 1 is one 16-byte target feature record in and one 16-byte (a,b) record out
//   * c' = cost - c0 is evaluated in fp32 (tad_cost_prime) -- no float<->double conversions on the XU pipe
//   * the kernels are L1TEX (shared-memory pipe) bound, so every window value is read from shared memory exactly
//     once per pass: with K a compile-time constant the first K-1 values of a sliding run stay in registers
//     for the "subtract" side of the update (the runtime-k version re-loaded them: 23 instead of 16 LDS.128
//     per 8 outputs)
//   * box sums run on Blackwell's packed fp32x2 pipe (FADD2 / FFMA2): a float4 window update = 4 instructions
//   * the slice min/max is reduced per CTA (one atomic pair per CTA and slice)
//   * a = cov * (1/den) with 1/den precomputed per pixel instead of three IEEE divisions per evaluation
//   * odd shared-memory pitches -> conflict-free LDS.128 / STS.128 for both the row- and column-strided phases
// Plane order inside the float4 records of pass 1 is (I0 c', I1 c', I2 c', c') so that channel pairs line up
// with the packed instructions.
#pragma once
#include "k_guided.cuh"

#define GFF_IW 64
#define GFF_IH 32
#define GFF_PP 65              // P pitch (float4), odd
#define GFF_THREADS 256        // 8 warps
#define GFF_NW 8
#define GFF_EPT 8              // input positions per thread (32*64 / 256)
#define GFF_RUN 8              // outputs per sliding run
#define GFF_MAXK 15

struct GffGeom {
    int H, W, Wp, a;
    int x0_base, x0_step, D;
};

// packed fp32x2 helpers (sm_100a FADD2 / FFMA2 / FMUL2)
__device__ __forceinline__ float4 p4add(float4 a, float4 b) {
    float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// s + (nw - old): the difference does not depend on the running sum, so a sliding run is a chain of ONE packed add per
// output (the compiler may not reassociate floats itself)
__device__ __forceinline__ float4 p4slide(float4 s, float4 old, float4 nw) {
    const float2 m1 = make_float2(-1.0f, -1.0f);
    float2 lo = __fadd2_rn(make_float2(s.x, s.y), __ffma2_rn(make_float2(old.x, old.y), m1, make_float2(nw.x, nw.y)));
    float2 hi = __fadd2_rn(make_float2(s.z, s.w), __ffma2_rn(make_float2(old.z, old.w), m1, make_float2(nw.z, nw.w)));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// One sliding run: LEN <= 8 window sums of K consecutive values starting at src (stride `st` float4); every value
// is loaded once.  f(o, sum) consumes output o.
template <int K, typename F>
__device__ __forceinline__ void gff_run(const float4* __restrict__ src, int st, int len, F&& f) {
    float4 v[K - 1];
    float4 s;
#pragma unroll
    for (int j = 0; j < K - 1; j++) {
        v[j] = src[j * st];
        s = (j == 0) ? v[0] : p4add(s, v[j]);
    }
    s = p4add(s, src[(K - 1) * st]);
    f(0, s);
#pragma unroll
    for (int o = 1; o < GFF_RUN; o++) {
        if (o < len) {
            // outputs beyond K-1 subtract values that were "new" earlier in this run; for K-1 >= RUN-1 they are all in v[]
            float4 old = (o - 1 < K - 1) ? v[(o - 1) % (K - 1)] : src[(o - 1) * st];
            s = p4slide(s, old, src[(o - 1 + K) * st]);
            f(o, s);
        }
    }
}

template <int K>
__device__ __forceinline__ void gff_hsum(const float4* __restrict__ P, float4* __restrict__ Hs, int tid) {
    constexpr int TW = GFF_IW + 1 - K, HP = TW + 1, NRUN = (TW + GFF_RUN - 1) / GFF_RUN;
    // item = (row r, run of 8 outputs); consecutive threads -> consecutive rows (odd pitches: no bank conflicts)
    if (tid < GFF_IH * NRUN) {
        int run = tid / GFF_IH, r = tid - run * GFF_IH;
        float4* dst = Hs + r * HP + run * GFF_RUN;
        gff_run<K>(P + r * GFF_PP + run * GFF_RUN, 1, min(GFF_RUN, TW - run * GFF_RUN), [&](int o, float4 s) { dst[o] = s; });
    }
}

// pass 1: cost + first box level -> (a0, a1, a2, b) per disparity evaluation
template <int K, int DC>
__global__ void __launch_bounds__(GFF_THREADS, 2)
k_gff_ab(const Feat* __restrict__ ref, const Feat* __restrict__ tgt, const float4* __restrict__ Gi,
         const float4* __restrict__ Gm, const float4* __restrict__ Grd /* 1/den */, GffGeom g, TadFast tp, float c0,
         float4* __restrict__ ab, uint32_t* __restrict__ slice_mm) {
    constexpr int TW = GFF_IW + 1 - K, TH = GFF_IH + 1 - K, HP = TW + 1, NRR = (TH + GFF_RUN - 1) / GFF_RUN;
    extern __shared__ float4 sm_gff[];
    float4* P = sm_gff;                                  // [32][65]
    float4* Hs = sm_gff + GFF_IH * GFF_PP;               // [32][TW+1]
    float2* GM = (float2*)(Hs + GFF_IH * HP);            // [TH*TW][3] float2: (m0,m1) (m2,rd0) (rd1,rd2)
    __shared__ float red_min[GFF_NW], red_max[GFF_NW];
    __shared__ int rowidx[GFF_IH], colidx[GFF_IW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0t = blockIdx.x * TW, y0t = blockIdx.y * TH;
    if (tid < GFF_IH) rowidx[tid] = border_idx(y0t - g.a + tid, g.H, 1);
    else if (tid < GFF_IH + GFF_IW) colidx[tid - GFF_IH] = border_idx(x0t - g.a + tid - GFF_IH, g.W, 1);
    // guidance moments of the output pixels -> shared memory (disparity independent)
    for (int i = tid; i < TH * TW; i += GFF_THREADS) {
        int r = i / TW, c = i - r * TW;
        int y = min(y0t + r, g.H - 1), x = min(x0t + c, g.W - 1);
        float4 m = __ldg(&Gm[(size_t)y * g.W + x]);
        float4 rd = __ldg(&Grd[(size_t)y * g.W + x]);
        GM[i * 3 + 0] = make_float2(m.x, m.y);
        GM[i * 3 + 1] = make_float2(m.z, rd.x);
        GM[i * 3 + 2] = make_float2(rd.y, rd.z);
    }
    __syncthreads();
    // ---- disparity-independent prologue: 8 owned input positions ----
    uint32_t f_bgr[GFF_EPT], f_g01[GFF_EPT], f_g2[GFF_EPT];
    float I0[GFF_EPT], I1[GFF_EPT], I2[GFF_EPT];
    int toff[GFF_EPT];
#pragma unroll
    for (int e = 0; e < GFF_EPT; e++) {
        int r = warp * 4 + (e >> 1), c = lane + ((e & 1) << 5);
        int sy = rowidx[r], sx = colidx[c];
        Feat fa = ref[(size_t)sy * g.W + sx];
        f_bgr[e] = fa.bgr; f_g01[e] = fa.g01; f_g2[e] = fa.g2;
        float4 I = __ldg(&Gi[(size_t)sy * g.W + sx]);
        I0[e] = I.x; I1[e] = I.y; I2[e] = I.z;
        toff[e] = sy * g.Wp + sx;
    }
    // phase C ownership: column col, rows rrun*8 .. rrun*8+7
    const int col = tid % TW, rrun = tid / TW;
    const int x = x0t + col;
    int c_len = (rrun < NRR) ? min(GFF_RUN, TH - rrun * GFF_RUN) : 0;
    c_len = min(c_len, g.H - (y0t + rrun * GFF_RUN));    // rows below the image are not written
    if (x >= g.W) c_len = 0;
    const int pix0 = (y0t + rrun * GFF_RUN) * g.W + x;
    const float inv = 1.0f / (float)(K * K);
    const int d_begin = blockIdx.z * DC;
    const size_t n = (size_t)g.H * g.W;

    for (int dd = 0; dd < DC; dd++) {
        const int di = d_begin + dd;
        if (di >= g.D) break;
        const int xoff = g.x0_base + g.x0_step * di;
        // ---- phase A: c' and its products with the guidance, pre-scaled by 1/K^2 ----
        float cmin = 3.0e38f, cmax = -3.0e38f;
        Feat fb[GFF_EPT];
#pragma unroll
        for (int e = 0; e < GFF_EPT; e++) fb[e] = tgt[toff[e] + xoff];
#pragma unroll
        for (int e = 0; e < GFF_EPT; e++) {
            float cp = tad_cost_prime(f_bgr[e], f_g01[e], f_g2[e], fb[e], tp);     // c' = cost - c0
            cmin = fminf(cmin, cp); cmax = fmaxf(cmax, cp);
            float cs = cp * inv;
            float2 p01 = __fmul2_rn(make_float2(I0[e], I1[e]), make_float2(cs, cs));
            int r = warp * 4 + (e >> 1), c = lane + ((e & 1) << 5);
            P[r * GFF_PP + c] = make_float4(p01.x, p01.y, I2[e] * cs, cs);
        }
        for (int o = 16; o > 0; o >>= 1) {
            cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        }
        if (lane == 0) { red_min[warp] = cmin; red_max[warp] = cmax; }
        __syncthreads();
        if (tid == 0) {
            float mn = red_min[0], mx = red_max[0];
#pragma unroll
            for (int w = 1; w < GFF_NW; w++) { mn = fminf(mn, red_min[w]); mx = fmaxf(mx, red_max[w]); }
            atomicMin(&slice_mm[2 * di], orderable_u32(__fadd_rn(c0, mn)));      // slice min / max of the raw cost
            atomicMax(&slice_mm[2 * di + 1], orderable_u32(__fadd_rn(c0, mx)));
        }
        // ---- phase B: horizontal window sums ----
        gff_hsum<K>(P, Hs, tid);
        __syncthreads();
        // ---- phase C: vertical window sums -> mean_p, corr_Ip -> a, b ----
        if (c_len > 0) {
            const float2* gm = GM + ((rrun * GFF_RUN) * TW + col) * 3;
            float4* out = ab + (size_t)di * n + pix0;
            gff_run<K>(Hs + (rrun * GFF_RUN) * HP + col, HP, c_len, [&](int o, float4 s) {
                float2 m01 = gm[o * TW * 3 + 0], m2r0 = gm[o * TW * 3 + 1], r12 = gm[o * TW * 3 + 2];
                float mP = s.w;                                                             // box(c')
                const float2 m1 = make_float2(-1.0f, -1.0f);
                // cov = corr_Ip - mean_I * mean_p ; a = cov / (var + eps)                    (A.cpp:2805-2846)
                float2 cov01 = __ffma2_rn(__fmul2_rn(m01, make_float2(mP, mP)), m1, make_float2(s.x, s.y));
                float2 a01 = __fmul2_rn(cov01, make_float2(m2r0.y, r12.x));
                float a2 = __fmul_rn(__fsub_rn(s.z, __fmul_rn(m2r0.x, mP)), r12.y);
                float2 am = __fmul2_rn(a01, m01);
                float dot = __fadd_rn(__fadd_rn(am.x, am.y), __fmul_rn(a2, m2r0.x));
                __stcs(out + o * g.W, make_float4(a01.x, a01.y, a2, __fsub_rn(mP, dot)));      // A.cpp:2847
            });
        }
        // the next slice's phase A only writes P (last read in phase B, before the barrier above); its phase B
        // writes Hs only after the barrier that follows phase A, i.e. after every thread left this phase C
    }
}

// pass 2: second box level -> q, per-slice affine, WTA keys.  dc = slices handled by one CTA (runtime).
template <int K>
__global__ void __launch_bounds__(GFF_THREADS, 2)
k_gff_q(const float4* __restrict__ ab, const float4* __restrict__ Gi, GffGeom g, float c0,
        const uint32_t* __restrict__ slice_mm, int d_first_label, int dc, unsigned long long* __restrict__ keys,
        float* __restrict__ agg) {
    constexpr int TW = GFF_IW + 1 - K, TH = GFF_IH + 1 - K, HP = TW + 1, NRR = (TH + GFF_RUN - 1) / GFF_RUN;
    extern __shared__ float4 sm_gff[];
    float4* P = sm_gff;
    float4* Hs = sm_gff + GFF_IH * GFF_PP;
    __shared__ int rowidx[GFF_IH], colidx[GFF_IW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0t = blockIdx.x * TW, y0t = blockIdx.y * TH;
    if (tid < GFF_IH) rowidx[tid] = border_idx(y0t - g.a + tid, g.H, 1);
    else if (tid < GFF_IH + GFF_IW) colidx[tid - GFF_IH] = border_idx(x0t - g.a + tid - GFF_IH, g.W, 1);
    __syncthreads();
    int off[GFF_EPT];
#pragma unroll
    for (int e = 0; e < GFF_EPT; e++) off[e] = rowidx[warp * 4 + (e >> 1)] * g.W + colidx[lane + ((e & 1) << 5)];
    const int col = tid % TW, rrun = tid / TW;
    const int x = x0t + col;
    int c_len = (rrun < NRR) ? min(GFF_RUN, TH - rrun * GFF_RUN) : 0;
    c_len = min(c_len, g.H - (y0t + rrun * GFF_RUN));
    if (x >= g.W) c_len = 0;
    const int pix0 = (y0t + rrun * GFF_RUN) * g.W + x;
    const float inv = 1.0f / (float)(K * K);
    const size_t n = (size_t)g.H * g.W;
    unsigned long long best[GFF_RUN];
    float Ix[GFF_RUN], Iy[GFF_RUN], Iz[GFF_RUN];        // guidance of the owned output pixels, pre-scaled by 1/K^2
#pragma unroll
    for (int o = 0; o < GFF_RUN; o++) {
        best[o] = WTA_KEY_EMPTY;
        float4 I = (o < c_len) ? __ldg(&Gi[pix0 + o * g.W]) : make_float4(0, 0, 0, 0);
        Ix[o] = I.x * inv; Iy[o] = I.y * inv; Iz[o] = I.z * inv;
    }
    const int d_begin = blockIdx.z * dc;
    for (int dd = 0; dd < dc; dd++) {
        const int di = d_begin + dd;
        if (di >= g.D) break;
        const float4* abd = ab + (size_t)di * n;
        float4 v[GFF_EPT];
#pragma unroll
        for (int e = 0; e < GFF_EPT; e++) v[e] = __ldcs(&abd[off[e]]);
#pragma unroll
        for (int e = 0; e < GFF_EPT; e++) P[(warp * 4 + (e >> 1)) * GFF_PP + lane + ((e & 1) << 5)] = v[e];
        // per-slice affine of cv::normalize (A.cpp:2775): q = sf * q' + (c0 * sf + hf)
        float sf, hf;
        minmax_scale_shift((double)from_orderable(slice_mm[2 * di]), (double)from_orderable(slice_mm[2 * di + 1]), &sf, &hf);
        float h2 = (float)fma((double)c0, (double)sf, (double)hf);
        __syncthreads();
        gff_hsum<K>(P, Hs, tid);
        __syncthreads();
        if (c_len > 0) {
            gff_run<K>(Hs + (rrun * GFF_RUN) * HP + col, HP, c_len, [&](int o, float4 s) {
                float2 t = __fmul2_rn(make_float2(s.x, s.y), make_float2(Ix[o], Iy[o]));
                float dot = __fadd_rn(__fadd_rn(t.x, t.y), __fmul_rn(s.z, Iz[o]));             // abar . I (A.cpp:2852)
                float q = fmaf(__fadd_rn(dot, s.w * inv), sf, h2);                              // + bbar, slice affine
                if (agg) agg[(size_t)di * n + pix0 + o * g.W] = q;
                best[o] = min(best[o], wta_key(q, d_first_label + di));
            });
        }
    }
#pragma unroll
    for (int o = 0; o < GFF_RUN; o++)
        if (o < c_len) atomicMin(&keys[pix0 + o * g.W], best[o]);
}

// host side of the tiled pair (windows 11, 13, 15 in the product build; 5, 7, 9 take the streaming kernel)
__global__ void k_reciprocal4(const float4* __restrict__ in, size_t n, float4* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = in[i];
    out[i] = make_float4(__fdiv_rn(1.0f, v.x), __fdiv_rn(1.0f, v.y), __fdiv_rn(1.0f, v.z), 0.0f);
}

// host-side launcher for one window size
template <int K>
static asw_status gff_launch(asw_ctx* ctx, const Feat* fref, const Feat* ftgt, const float4* Gi, const float4* Gm,
                             const float4* grd, GffGeom g, const TadParams& tp, float4* ab, uint32_t* slice_mm,
                             int cn, int d_label0, unsigned long long* keys, float* agg) {
    constexpr int TW = GFF_IW + 1 - K, TH = GFF_IH + 1 - K, DC1 = 8;
    size_t smem = ((size_t)GFF_IH * GFF_PP + (size_t)GFF_IH * (TW + 1)) * sizeof(float4);
    size_t smem_ab = smem + (size_t)TW * TH * 3 * sizeof(float2);
    cudaFuncSetAttribute(k_gff_ab<K, DC1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ab);
    cudaFuncSetAttribute(k_gff_q<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int tx = cdiv(g.W, TW), ty = cdiv(g.H, TH);
    g.D = cn;
    LAUNCH(ctx, "gf_ab", (k_gff_ab<K, DC1><<<dim3(tx, ty, cdiv(cn, DC1)), GFF_THREADS, smem_ab, ctx->stream>>>(
                             fref, ftgt, Gi, Gm, grd, g, make_tad_fast(tp), tp.c0, ab, slice_mm)));
    // pass 2: as many slices per CTA as still leaves >= 4 CTAs per SM (fewer key atomics per pixel)
    int dc2 = cn;
    while (dc2 > 8 && (long long)tx * ty * cdiv(cn, dc2) < (long long)ctx->sm_count * 4) dc2 = (dc2 + 1) / 2;
    LAUNCH(ctx, "gf_q", (k_gff_q<K><<<dim3(tx, ty, cdiv(cn, dc2)), GFF_THREADS, smem, ctx->stream>>>(
                            ab, Gi, g, tp.c0, slice_mm, d_label0, dc2, keys, agg)));
    return ASW_OK;
}
