// k_guided_fast.cuh -- tuned guided-filter kernels for windows k <= 13 (configs 2 and 5 use k = 9).
//
// Same algorithm and arithmetic as k_gf_ab / k_gf_q in k_guided.cuh (which remain the generic path for
// larger windows); what changes is the mapping onto the SM:
//   * the input tile is fixed at 36 rows x 64 columns (outputs TH x TW = (37-k) x (65-k)); 9 warps, warp w
//     owns input rows 4w..4w+3, lane l owns input columns l and l+32 -> every thread owns the same 8 tile
//     positions for every disparity, no div/mod or border arithmetic in the loop
//   * everything that does not depend on the disparity (reference-side feature record, normalised guidance,
//     reflected addresses) is loaded ONCE per CTA into registers and reused for all DC slices of the chunk;
//     per disparity evaluation the only global load of pass 1 is one 16-byte target feature record
//   * the slice min/max is reduced per CTA (one atomic pair per CTA and slice instead of one per warp)
//   * a = cov * (1/den) with 1/den precomputed per pixel instead of three IEEE divisions per evaluation
//   * sliding k-sums on float4 in shared memory, odd pitches -> conflict-free LDS.128 / STS.128
#pragma once
#include "k_guided.cuh"

#define GFF_IW 64
#define GFF_IH 36
#define GFF_PP 65              // P pitch (float4), odd
#define GFF_THREADS 288        // 9 warps
#define GFF_MAXK 13

struct GffGeom {
    int H, W, Wp, k, a, TW, TH;
    int x0_base, x0_step, D;
};

__device__ __forceinline__ void gff_hsum(const float4* __restrict__ P, float4* __restrict__ Hs, int tid, int k, int TW, int HP) {
    // item = (row r, run of 8 outputs); consecutive threads -> consecutive rows (odd pitches: no bank conflicts)
    const int nrun = (TW + 7) >> 3;
    if (tid < GFF_IH * nrun) {
        int run = tid / GFF_IH, r = tid - run * GFF_IH;
        const float4* src = P + r * GFF_PP + run * 8;
        float4* dst = Hs + r * HP + run * 8;
        int len = min(8, TW - run * 8);
        float4 s = src[0];
        for (int j = 1; j < k; j++) s = f4add(s, src[j]);
        dst[0] = s;
#pragma unroll
        for (int o = 1; o < 8; o++) {
            if (o < len) {
                s = f4add(f4sub(s, src[o - 1]), src[o - 1 + k]);
                dst[o] = s;
            }
        }
    }
}

// pass 1: cost + first box level -> (a0, a1, a2, b) per disparity evaluation
template <int DC>
__global__ void __launch_bounds__(GFF_THREADS, 2)
k_gff_ab(const Feat* __restrict__ ref, const Feat* __restrict__ tgt, const float4* __restrict__ Gi,
         const float4* __restrict__ Gm, const float4* __restrict__ Grd /* 1/den */, GffGeom g, TadParams tp,
         float4* __restrict__ ab, uint32_t* __restrict__ slice_mm) {
    extern __shared__ float4 sm_gff[];
    float4* P = sm_gff;                                  // [36][65]
    float4* Hs = sm_gff + GFF_IH * GFF_PP;               // [36][TW+1]
    __shared__ float red_min[9], red_max[9];
    __shared__ int rowidx[GFF_IH], colidx[GFF_IW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = g.k, TW = g.TW, TH = g.TH, HP = TW + 1;
    const int x0t = blockIdx.x * TW, y0t = blockIdx.y * TH;
    if (tid < GFF_IH) rowidx[tid] = border_idx(y0t - g.a + tid, g.H, 1);
    else if (tid < GFF_IH + GFF_IW) colidx[tid - GFF_IH] = border_idx(x0t - g.a + tid - GFF_IH, g.W, 1);
    __syncthreads();
    // ---- disparity-independent prologue: 8 owned positions ----
    uint32_t f_bgr[8], f_g01[8], f_g2[8];
    float I0[8], I1[8], I2[8];
    int toff[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        int r = warp * 4 + (e >> 1), c = lane + ((e & 1) << 5);
        int sy = rowidx[r], sx = colidx[c];
        Feat fa = ref[(size_t)sy * g.W + sx];
        f_bgr[e] = fa.bgr; f_g01[e] = fa.g01; f_g2[e] = fa.g2;
        float4 I = __ldg(&Gi[(size_t)sy * g.W + sx]);
        I0[e] = I.x; I1[e] = I.y; I2[e] = I.z;
        toff[e] = sy * g.Wp + sx;
    }
    // phase C ownership: column col, rows rrun*8 .. +7
    const int col = tid % TW, rrun = tid / TW;
    const int nrr = (TH + 7) >> 3;
    const bool c_active = rrun < nrr;
    const int c_len = c_active ? min(8, TH - rrun * 8) : 0;
    const int x = x0t + col;
    const float inv = 1.0f / (float)(k * k);
    const int d_begin = blockIdx.z * DC;

    for (int dd = 0; dd < DC; dd++) {
        const int di = d_begin + dd;
        if (di >= g.D) break;
        const int xoff = g.x0_base + g.x0_step * di;
        // ---- phase A ----
        float cmin = 3.0e38f, cmax = -3.0e38f;
        Feat fb[8];
#pragma unroll
        for (int e = 0; e < 8; e++) fb[e] = tgt[toff[e] + xoff];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            Feat fa; fa.bgr = f_bgr[e]; fa.g01 = f_g01[e]; fa.g2 = f_g2[e]; fa.pad = 0;
            float cst = tad_cost(fa, fb[e], tp);
            cmin = fminf(cmin, cst); cmax = fmaxf(cmax, cst);
            float cp = __fsub_rn(cst, tp.c0);
            int r = warp * 4 + (e >> 1), c = lane + ((e & 1) << 5);
            P[r * GFF_PP + c] = make_float4(cp, __fmul_rn(I0[e], cp), __fmul_rn(I1[e], cp), __fmul_rn(I2[e], cp));
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
            for (int w = 1; w < 9; w++) { mn = fminf(mn, red_min[w]); mx = fmaxf(mx, red_max[w]); }
            atomicMin(&slice_mm[2 * di], orderable_u32(mn));
            atomicMax(&slice_mm[2 * di + 1], orderable_u32(mx));
        }
        // ---- phase B ----
        gff_hsum(P, Hs, tid, k, TW, HP);
        __syncthreads();
        // ---- phase C ----
        if (c_active) {
            const float4* src = Hs + (rrun * 8) * HP + col;
            float4 s = src[0];
            for (int j = 1; j < k; j++) s = f4add(s, src[j * HP]);
#pragma unroll
            for (int o = 0; o < 8; o++) {
                if (o < c_len) {
                    if (o > 0) s = f4add(f4sub(s, src[(o - 1) * HP]), src[(o - 1 + k) * HP]);
                    int y = y0t + rrun * 8 + o;
                    if (x < g.W && y < g.H) {
                        size_t pix = (size_t)y * g.W + x;
                        float4 m = __ldg(&Gm[pix]);
                        float4 rd = __ldg(&Grd[pix]);
                        float mP = s.x * inv;
                        float a0 = __fmul_rn(__fsub_rn(s.y * inv, __fmul_rn(m.x, mP)), rd.x);   // A.cpp:2805-2846
                        float a1 = __fmul_rn(__fsub_rn(s.z * inv, __fmul_rn(m.y, mP)), rd.y);
                        float a2 = __fmul_rn(__fsub_rn(s.w * inv, __fmul_rn(m.z, mP)), rd.z);
                        float dot = __fadd_rn(__fadd_rn(__fmul_rn(a0, m.x), __fmul_rn(a1, m.y)), __fmul_rn(a2, m.z));
                        ab[(size_t)di * g.H * g.W + pix] = make_float4(a0, a1, a2, __fsub_rn(mP, dot));   // A.cpp:2847
                    }
                }
            }
        }
        // the next slice's phase A only writes P (last read in phase B, before the barrier above); its phase B
        // writes Hs only after the barrier that follows phase A, i.e. after every thread left this phase C
    }
}

// pass 2: second box level -> q, per-slice affine, WTA keys.  dc = slices handled by one CTA (runtime).
__global__ void __launch_bounds__(GFF_THREADS, 2)
k_gff_q(const float4* __restrict__ ab, const float4* __restrict__ Gi, GffGeom g, float c0,
        const uint32_t* __restrict__ slice_mm, int d_first_label, int dc, unsigned long long* __restrict__ keys,
        float* __restrict__ agg) {
    extern __shared__ float4 sm_gff[];
    float4* P = sm_gff;
    float4* Hs = sm_gff + GFF_IH * GFF_PP;
    __shared__ int rowidx[GFF_IH], colidx[GFF_IW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = g.k, TW = g.TW, TH = g.TH, HP = TW + 1;
    const int x0t = blockIdx.x * TW, y0t = blockIdx.y * TH;
    if (tid < GFF_IH) rowidx[tid] = border_idx(y0t - g.a + tid, g.H, 1);
    else if (tid < GFF_IH + GFF_IW) colidx[tid - GFF_IH] = border_idx(x0t - g.a + tid - GFF_IH, g.W, 1);
    __syncthreads();
    int off[8];
#pragma unroll
    for (int e = 0; e < 8; e++) off[e] = rowidx[warp * 4 + (e >> 1)] * g.W + colidx[lane + ((e & 1) << 5)];
    const int col = tid % TW, rrun = tid / TW;
    const int nrr = (TH + 7) >> 3;
    const bool c_active = rrun < nrr;
    const int c_len = c_active ? min(8, TH - rrun * 8) : 0;
    const int x = x0t + col;
    const float inv = 1.0f / (float)(k * k);
    const size_t n = (size_t)g.H * g.W;
    unsigned long long best[8];
    float4 Ipix[8];
#pragma unroll
    for (int o = 0; o < 8; o++) {
        best[o] = WTA_KEY_EMPTY;
        int y = y0t + rrun * 8 + o;
        Ipix[o] = (c_active && o < c_len && x < g.W && y < g.H) ? __ldg(&Gi[(size_t)y * g.W + x]) : make_float4(0, 0, 0, 0);
    }
    const int d_begin = blockIdx.z * dc;
    for (int dd = 0; dd < dc; dd++) {
        const int di = d_begin + dd;
        if (di >= g.D) break;
        const float4* abd = ab + (size_t)di * n;
        float4 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = abd[off[e]];
#pragma unroll
        for (int e = 0; e < 8; e++) P[(warp * 4 + (e >> 1)) * GFF_PP + lane + ((e & 1) << 5)] = v[e];
        float sf, hf;
        minmax_scale_shift((double)from_orderable(slice_mm[2 * di]), (double)from_orderable(slice_mm[2 * di + 1]), &sf, &hf);
        float h2 = (float)fma((double)c0, (double)sf, (double)hf);
        __syncthreads();
        gff_hsum(P, Hs, tid, k, TW, HP);
        __syncthreads();
        if (c_active) {
            const float4* src = Hs + (rrun * 8) * HP + col;
            float4 s = src[0];
            for (int j = 1; j < k; j++) s = f4add(s, src[j * HP]);
#pragma unroll
            for (int o = 0; o < 8; o++) {
                if (o < c_len) {
                    if (o > 0) s = f4add(f4sub(s, src[(o - 1) * HP]), src[(o - 1 + k) * HP]);
                    float4 I = Ipix[o];
                    float dot = __fadd_rn(__fadd_rn(__fmul_rn(s.x * inv, I.x), __fmul_rn(s.y * inv, I.y)), __fmul_rn(s.z * inv, I.z));
                    float q = fmaf(__fadd_rn(dot, s.w * inv), sf, h2);                    // A.cpp:2852 + slice affine
                    int y = y0t + rrun * 8 + o;
                    if (agg && x < g.W && y < g.H) agg[(size_t)di * n + (size_t)y * g.W + x] = q;
                    best[o] = min(best[o], wta_key(q, d_first_label + di));
                }
            }
        }
    }
    if (c_active) {
#pragma unroll
        for (int o = 0; o < 8; o++) {
            int y = y0t + rrun * 8 + o;
            if (o < c_len && x < g.W && y < g.H) atomicMin(&keys[(size_t)y * g.W + x], best[o]);
        }
    }
}

__global__ void k_reciprocal4(const float4* __restrict__ in, size_t n, float4* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = in[i];
    out[i] = make_float4(__fdiv_rn(1.0f, v.x), __fdiv_rn(1.0f, v.y), __fdiv_rn(1.0f, v.z), 0.0f);
}
