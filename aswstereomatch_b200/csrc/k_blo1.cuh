// k_blo1.cuh -- O(1)-bilateral ASW, computeAdaptiveWeight_BLO1 (A.cpp:2505-2725), minDisparity = 0.  Written for the LEFT
// view; DISPARITY_RIGHT (A.cpp:2538-2546, 2600-2631, 2685-2722) is the same computation with the right gray image as the
// reference side `lg` and the left one, padded on its right, as the target `rpad` (crop column +d): only the geometry differs.
//
// Reference: for every intensity level k (0, step, 2 step, ..., 255; step = int(256*sampleRateR)) and every d:
//   M_d = |L-k| * |R_d-k| ; J_{k,d} = box(M_d * c_d) ; N_k = box(M_{D-1})  (LAST d only, A.cpp:2588) ;
//   JB_{k,d} = J_{k,d} / N_k ; then per pixel with I = L(p): cost = JB_{I,d} if I is a level, otherwise
//   (I-lo)*JB_{lo,d} + (hi-I)*JB_{hi,d} (weights swapped and un-normalised, A.cpp:2666-2667) ; WTA.
// The reference materialises all 86 x D planes (40 GB at 1280x720x128).  Here one CTA owns a 64x32 tile
// of one disparity, keeps the L / R_d / c_d tiles (+ window halo) in shared memory and loops only over the
// levels that some pixel of the tile actually consumes (a pixel consumes <= 2 levels), running the box
// filter as separable sliding sums in shared memory; nothing per-level ever reaches HBM.  N_k is
// d-independent and is computed once per level with exact integer sums.
#pragma once
#include "k_cost.cuh"

#define BLO_TW 64
#define BLO_TH 32
#define BLO_THREADS 256
#define BLO_MAXLEV 257

struct BloGeom {
    int H, W, Wp, win, h, D;
    int x0_base, x0_step;        // padded target crop column for slice di: x0_base + x0_step * di
    int step, nl;                // level step, number of levels
    int last255;                 // 1 if the last level (255) is not a multiple of step
    int di_lo, di_hi;            // slices [di_lo, di_hi) are aggregated by this launch (disparity-range split)
};
__device__ __forceinline__ int blo_level_value(const BloGeom& g, int li) {
    return (li == g.nl - 1 && g.last255) ? 255 : li * g.step;
}

// N_k = box(|L-k| * |R_{D-1}-k|), exact: integer window sums, (float)(sum * (1/win^2)) as cv::boxFilter.
// Only the two levels a pixel consumes are kept: Nk is [H][W][2] = {N at the lower level, N at the next level}
// (7 MB at 1280x720, L2-resident) instead of the nl full-resolution planes.
// grid: (tiles_x, tiles_y, nl)
__global__ void __launch_bounds__(BLO_THREADS)
k_blo1_norm(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rpad, BloGeom g, float* __restrict__ Nk) {
    extern __shared__ int sm_blo_i[];
    const int win = g.win, h = g.h;
    const int IW = BLO_TW + win - 1, IH = BLO_TH + win - 1, PP = IW | 1, HP = BLO_TW + 1;
    int* M = sm_blo_i;                       // [IH][PP]
    int* Hs = sm_blo_i + IH * PP;            // [IH][HP]
    const int tid = threadIdx.x, x0t = blockIdx.x * BLO_TW, y0t = blockIdx.y * BLO_TH;
    const int k = blo_level_value(g, blockIdx.z);
    const int xoff = g.x0_base + g.x0_step * (g.D - 1);        // the last disparity's crop (A.cpp:2588)
    for (int i = tid; i < IH * IW; i += BLO_THREADS) {
        int r = i / IW, c = i - r * IW;
        int sy = border_idx(y0t - h + r, g.H, 1), sx = border_idx(x0t - h + c, g.W, 1);
        int l = lg[(size_t)sy * g.W + sx], rr = rpad[(size_t)sy * g.Wp + xoff + sx];
        M[r * PP + c] = abs(l - k) * abs(rr - k);
    }
    __syncthreads();
    for (int i = tid; i < IH * (BLO_TW / 8); i += BLO_THREADS) {
        int seg = i / IH, r = i - seg * IH;
        const int* src = M + r * PP + seg * 8;
        int s = 0;
        for (int j = 0; j < win; j++) s += src[j];
        int* dst = Hs + r * HP + seg * 8;
        dst[0] = s;
#pragma unroll
        for (int o = 1; o < 8; o++) { s += src[o - 1 + win] - src[o - 1]; dst[o] = s; }
    }
    __syncthreads();
    {
        const int col = tid % BLO_TW, rseg = tid / BLO_TW, x = x0t + col;
        const int* src = Hs + (rseg * 8) * HP + col;
        int s = 0;
        for (int j = 0; j < win; j++) s += src[j * HP];
        const double scale = 1.0 / ((double)win * win);
        const int li = blockIdx.z;
#pragma unroll
        for (int o = 0; o < 8; o++) {
            if (o > 0) s += src[(o - 1 + win) * HP] - src[(o - 1) * HP];
            int y = y0t + rseg * 8 + o;
            if (x < g.W && y < g.H) {
                // a pixel consumes N_k at its lower level (slot 0) and, unless its intensity is a level, at the next one (slot 1)
                const int I = lg[(size_t)y * g.W + x];
                const bool isl = (I % g.step == 0) || I == 255;
                const int key = (I == 255 && g.last255) ? g.nl - 1 : I / g.step;
                if (li == key) Nk[((size_t)y * g.W + x) * 2] = (float)((double)s * scale);
                else if (li == key + 1 && !isl) Nk[((size_t)y * g.W + x) * 2 + 1] = (float)((double)s * scale);
            }
        }
    }
}

// grid: (tiles_x, tiles_y, D)
__global__ void __launch_bounds__(BLO_THREADS)
k_blo1_aggregate(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rpad, const float* __restrict__ cost,
                 const float* __restrict__ Nk, BloGeom g, int d_label0, unsigned long long* __restrict__ keys,
                 float* __restrict__ agg) {
    extern __shared__ float sm_blo_f[];
    const int win = g.win, h = g.h;
    const int IW = BLO_TW + win - 1, IH = BLO_TH + win - 1, PP = IW | 1, HP = BLO_TW + 1;
    float* P = sm_blo_f;                                  // [IH][PP]   M * c for the current level
    float* Hs = P + IH * PP;                              // [IH][HP]
    float* Ct = Hs + IH * HP;                             // [IH][PP]   c_d tile
    uint8_t* Lt = (uint8_t*)(Ct + IH * PP);               // [IH][PP]
    uint8_t* Rt = Lt + IH * PP;                           // [IH][PP]
    __shared__ uint32_t need[(BLO_MAXLEV + 31) / 32];
    const int tid = threadIdx.x, x0t = blockIdx.x * BLO_TW, y0t = blockIdx.y * BLO_TH;
    const int di = g.di_lo + blockIdx.z;
    const int xoff = g.x0_base + g.x0_step * di;
    const size_t n = (size_t)g.H * g.W;
    const float* cd = cost + (size_t)di * n;
    if (tid < (BLO_MAXLEV + 31) / 32) need[tid] = 0;
    for (int i = tid; i < IH * IW; i += BLO_THREADS) {
        int r = i / IW, c = i - r * IW;
        int sy = border_idx(y0t - h + r, g.H, 1), sx = border_idx(x0t - h + c, g.W, 1);   // boxFilter REFLECT_101
        Lt[r * PP + c] = lg[(size_t)sy * g.W + sx];
        Rt[r * PP + c] = rpad[(size_t)sy * g.Wp + xoff + sx];
        Ct[r * PP + c] = cd[(size_t)sy * g.W + sx];
    }
    __syncthreads();
    // the 8 pixels this thread owns: (col, rseg*8 + o); their consumed levels
    const int col = tid % BLO_TW, rseg = tid / BLO_TW, x = x0t + col;
    int I8[8], lo_li[8], hi_li[8];
    float part_lo[8], part_hi[8];
#pragma unroll
    for (int o = 0; o < 8; o++) {
        int y = y0t + rseg * 8 + o;
        bool in = x < g.W && y < g.H;
        int I = Lt[(h + rseg * 8 + o) * PP + h + col];
        I8[o] = I;
        part_lo[o] = 0.0f; part_hi[o] = 0.0f;
        bool is_level = (I % g.step == 0) || I == 255;          // discretInten membership (A.cpp:2656)
        if (is_level) {
            lo_li[o] = (I == 255 && g.last255) ? g.nl - 1 : I / g.step;
            hi_li[o] = -1;
        } else {
            int lo = I / g.step * g.step, hi = lo + g.step;      // A.cpp:2658-2663
            lo_li[o] = I / g.step;
            hi_li[o] = hi > 255 ? g.nl - 1 : lo_li[o] + 1;
        }
        if (!in) { lo_li[o] = -1; hi_li[o] = -1; }
        if (lo_li[o] >= 0) atomicOr(&need[lo_li[o] >> 5], 1u << (lo_li[o] & 31));
        if (hi_li[o] >= 0) atomicOr(&need[hi_li[o] >> 5], 1u << (hi_li[o] & 31));
    }
    __syncthreads();
    const float inv = 1.0f / (float)(win * win);
    for (int li = 0; li < g.nl; li++) {
        if (!((need[li >> 5] >> (li & 31)) & 1u)) continue;      // block-uniform
        const int k = blo_level_value(g, li);
        for (int i = tid; i < IH * IW; i += BLO_THREADS) {
            int r = i / IW, c = i - r * IW, a = r * PP + c;
            float m = (float)(abs((int)Lt[a] - k) * abs((int)Rt[a] - k));     // A.cpp:2571-2580
            P[a] = __fmul_rn(m, Ct[a]);                                        // A.cpp:2583
        }
        __syncthreads();
        for (int i = tid; i < IH * (BLO_TW / 8); i += BLO_THREADS) {
            int seg = i / IH, r = i - seg * IH;
            const float* src = P + r * PP + seg * 8;
            float s = 0.0f;
            for (int j = 0; j < win; j++) s += src[j];
            float* dst = Hs + r * HP + seg * 8;
            dst[0] = s;
#pragma unroll
            for (int o = 1; o < 8; o++) { s += src[o - 1 + win] - src[o - 1]; dst[o] = s; }
        }
        __syncthreads();
        {
            const float* src = Hs + (rseg * 8) * HP + col;
            float s = 0.0f;
            for (int j = 0; j < win; j++) s += src[j * HP];
#pragma unroll
            for (int o = 0; o < 8; o++) {
                if (o > 0) s += src[(o - 1 + win) * HP] - src[(o - 1) * HP];
                if (lo_li[o] == li || hi_li[o] == li) {
                    int y = y0t + rseg * 8 + o;
                    float nk = __ldg(&Nk[((size_t)y * g.W + x) * 2 + (lo_li[o] == li ? 0 : 1)]);
                    float jb = __fdiv_rn(s * inv, nk);                          // A.cpp:2594
                    if (hi_li[o] < 0) part_lo[o] = jb;                          // I is a level: cost = JB_{I,d}
                    else if (lo_li[o] == li) part_lo[o] = __fmul_rn((float)(I8[o] - k), jb);    // (I - lo) * JB_lo
                    else part_hi[o] = __fmul_rn((float)(k - I8[o]), jb);        // (hi - I) * JB_hi
                }
            }
        }
        // the next level's P writes do not touch Hs; its h-sum pass starts only after the next barrier
    }
#pragma unroll
    for (int o = 0; o < 8; o++) {
        int y = y0t + rseg * 8 + o;
        if (x < g.W && y < g.H) {
            float c = hi_li[o] < 0 ? part_lo[o] : __fadd_rn(part_lo[o], part_hi[o]);   // A.cpp:2666-2667
            size_t p = (size_t)y * g.W + x;
            if (agg) agg[(size_t)di * n + p] = c;
            atomicMin(&keys[p], wta_key(c, d_label0 + di));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Register-resident BLO(1) aggregation (window side as template constant).
//
// One CTA = 256 threads owns a tile of WIN+1 rows x (128 - (WIN-1)) columns and walks a chunk of
// disparities.  The 2*WIN halo rows of the tile split into two blocks of WIN rows; thread (g, cx) keeps
// block g of halo column cx in registers as three floats per element,
//     a = L*R_d, b = L+R_d, c = c_d,
// so that for a level k the product |L-k|*|R_d-k|*c_d of A.cpp:2571-2583 is |a - k*b + k^2| * c: one FFMA
// and one FADD (both exact: integers below 2^24) plus the FFMA that accumulates the block's running sum.
// A window of WIN rows is always "suffix of block 0 + prefix of block 1", so the vertical window sums
// of all WIN+1 output rows come out of the two running sums with no halo recomputation.  The horizontal
// window sums go through shared memory as 32-column segment prefix sums; only the pixels that consume
// the level (every pixel consumes <= 2 levels: A.cpp:2656-2667) read them, through a list of the
// tile's pixels sorted by level that is built once per CTA.  Only levels some pixel of the tile
// consumes are evaluated; nothing per-level reaches HBM.
// ---------------------------------------------------------------------------------------------
#define BLO2_COLS 128
#define BLO2_PITCH 132
#define BLO2_THREADS 256

// NORM = true computes the normalisers instead: the same sums with c = 1 over the slice [di_lo, di_lo + 1) (the host
// passes the LAST disparity, A.cpp:2588), written as box means into the per-pixel pair plane Nk_out = {N at the pixel's
// lower level, N at the next level} (7 MB at 1280x720 instead of nl full-resolution planes).
template <int WIN, bool NORM>
__global__ void __launch_bounds__(BLO2_THREADS, 2)
k_blo1_agg2(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rpad, const float* __restrict__ cost,
            const float* __restrict__ Nk, float* __restrict__ Nk_out, BloGeom g, int dch, int d_label0,
            unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    constexpr int TH = WIN + 1, SW = BLO2_COLS - (WIN - 1), NPIX = TH * SW, h = WIN / 2, PITCH = BLO2_PITCH;
    extern __shared__ float sm_blo2[];
    // running sums of the two blocks, each preceded by a row of zeros: [Z][U0: WIN rows][Z][U1: WIN rows]
    float* U0 = sm_blo2 + PITCH;
    float* U1 = U0 + (WIN + 1) * PITCH;
    float* Q = U1 + WIN * PITCH;                                  // [TH][PITCH]     segment prefix sums of the row sums
    float* part = Q + TH * PITCH;                                 // [NPIX]          (I - lo) * JB_lo
    uint32_t* list = (uint32_t*)(part + NPIX);                    // [NPIX]          pixel | I << 16 | is_level << 24
    unsigned long long* bestk = (unsigned long long*)(list + NPIX);   // [NPIX]
    __shared__ int lstart[BLO_MAXLEV + 1];                        // lstart[k] = first list entry of level k; lstart[nl] = total
    __shared__ int cursor[BLO_MAXLEV + 1];
    __shared__ uint32_t need[(BLO_MAXLEV + 31) / 32];
    const int tid = threadIdx.x, grp = tid >> 7, cx = tid & 127;
    const int x0 = blockIdx.x * SW, y0 = blockIdx.y * TH;
    const size_t n = (size_t)g.H * g.W;

    // ---- pixel list sorted by (lower) level ----
    for (int i = tid; i < BLO_MAXLEV + 1; i += BLO2_THREADS) lstart[i] = 0;
    for (int i = tid; i < BLO_MAXLEV + 1; i += BLO2_THREADS) cursor[i] = 0;
    if (tid < (BLO_MAXLEV + 31) / 32) need[tid] = 0;
    __syncthreads();
    for (int p = tid; p < NPIX; p += BLO2_THREADS) {
        int o = p / SW, xo = p - o * SW, y = y0 + o, x = x0 + xo;
        bestk[p] = WTA_KEY_EMPTY;
        if (x < g.W && y < g.H) {
            int I = lg[(size_t)y * g.W + x];
            bool isl = (I % g.step == 0) || I == 255;               // discretInten membership (A.cpp:2656)
            int key = (I == 255 && g.last255) ? g.nl - 1 : I / g.step;
            atomicAdd(&lstart[key], 1);
            atomicOr(&need[key >> 5], 1u << (key & 31));
            if (!isl) atomicOr(&need[(key + 1) >> 5], 1u << ((key + 1) & 31));   // hi level = next level (A.cpp:2658-2663)
        }
    }
    __syncthreads();
    if (tid == 0) {
        int s = 0;                                                 // counts -> exclusive prefix
        for (int i = 0; i <= g.nl; i++) { int c = lstart[i]; lstart[i] = s; s += c; }
    }
    __syncthreads();
    for (int p = tid; p < NPIX; p += BLO2_THREADS) {
        int o = p / SW, xo = p - o * SW, y = y0 + o, x = x0 + xo;
        if (x < g.W && y < g.H) {
            int I = lg[(size_t)y * g.W + x];
            bool isl = (I % g.step == 0) || I == 255;
            int key = (I == 255 && g.last255) ? g.nl - 1 : I / g.step;
            int pos = lstart[key] + atomicAdd(&cursor[key], 1);
            list[pos] = (uint32_t)p | ((uint32_t)I << 16) | ((uint32_t)isl << 24);
        }
    }
    __syncthreads();
    const int sx = border_idx(x0 - h + cx, g.W, 1);                // boxFilter BORDER_REFLECT_101
    float* Ug = (grp ? U1 : U0) + cx;
    if (tid < PITCH) { U0[tid - PITCH] = 0.0f; U1[tid - PITCH] = 0.0f; }
    // source rows of the two blocks (reflected once per CTA, not once per disparity and row)
    __shared__ int srow[2][WIN];
    if (tid < 2 * WIN) {
        const int gq = tid / WIN, i = tid - gq * WIN;
        srow[gq][i] = border_idx(y0 - h + (gq ? WIN + i : WIN - 1 - i), g.H, 1);   // block 0 runs upwards: its running sum is a suffix sum
    }
    __syncthreads();
    const float inv = 1.0f / (float)(WIN * WIN);

    for (int dd = 0; dd < dch; dd++) {
        const int di = g.di_lo + blockIdx.z * dch + dd;
        if (di >= g.di_hi) break;
        const int xoff = g.x0_base + g.x0_step * di;
        const float* cd = cost + (size_t)di * n;
        float a[WIN], b[WIN], c[WIN];
#pragma unroll
        for (int i = 0; i < WIN; i++) {
            const int sy = srow[grp][i];
            int l = lg[(size_t)sy * g.W + sx], rr = rpad[(size_t)sy * g.Wp + xoff + sx];
            a[i] = (float)(l * rr);
            b[i] = (float)(l + rr);
            c[i] = NORM ? 1.0f : __ldg(&cd[(size_t)sy * g.W + sx]);
        }
        for (int li = 0; li < g.nl; li++) {
            if (!((need[li >> 5] >> (li & 31)) & 1u)) continue;   // block-uniform
            const int k = blo_level_value(g, li);
            const float kf = (float)k, k2 = (float)(k * k);
            // this thread's first consumer entry of the level: fetch its normaliser now, use it after the two barriers
            const int e1 = lstart[li], e0 = li > 0 ? lstart[li - 1] : e1, e2 = lstart[li + 1];   // level li-1 | level li
            float nk_pre = 0.0f;
            const int rt = BLO2_THREADS - 1 - tid;                 // consumers are taken by the highest threads first:
            if (!NORM && e0 + rt < e2) {                           // the scan keeps the lowest 4 * TH threads busy
                const uint32_t e = list[e0 + rt];
                const int pix = e & 0xFFFF, o = pix / SW, xo = pix - o * SW;
                nk_pre = __ldg(&Nk[((size_t)(y0 + o) * g.W + (x0 + xo)) * 2 + (e0 + rt < e1 ? 1 : 0)]);
            }
            float run = 0.0f;
#pragma unroll
            for (int i = 0; i < WIN; i++) {
                float t = __fadd_rn(fmaf(b[i], -kf, a[i]), k2);    // (L-k)*(R-k), exact
                run = fmaf(fabsf(t), c[i], run);                   // + |L-k|*|R-k|*c   (A.cpp:2571-2583)
                Ug[i * PITCH] = run;
            }
            __syncthreads();
            if (tid < 4 * TH) {
                const int seg = tid / TH, o = tid - seg * TH;
                // rows [o, o+WIN-1] of the halo = block-0 suffix from row o + block-1 prefix up to row o-1
                // (o = WIN: the suffix is empty -> the zero row before U0; o = 0: the prefix is empty -> the zero row before U1)
                const float4* u0 = (const float4*)(U0 + (WIN - 1 - o) * PITCH + seg * 32);
                const float4* u1 = (const float4*)(U1 + (o - 1) * PITCH + seg * 32);
                float4* q = (float4*)(Q + o * PITCH + seg * 32);
                float s = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float4 v = u0[j];
                    const float4 w = u1[j];
                    float4 r4;
                    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;          // off the dependent chain
                    s += v.x; r4.x = s;
                    s += v.y; r4.y = s;
                    s += v.z; r4.z = s;
                    s += v.w; r4.w = s;
                    q[j] = r4;
                }
            }
            __syncthreads();
            for (int t = e0 + rt; t < e2; t += BLO2_THREADS) {
                const uint32_t e = list[t];
                const bool role_hi = t < e1, isl = (e >> 24) & 1u;
                if (role_hi && isl) continue;
                const int pix = e & 0xFFFF, I = (e >> 16) & 0xFF;
                const int o = pix / SW, xo = pix - o * SW;
                const float* q = Q + o * PITCH;
                const int cr = xo + WIN - 1, segr = cr >> 5;
                float sum = q[cr];
                int s = 0;
                if (xo > 0) { sum -= q[xo - 1]; s = (xo - 1) >> 5; }
                for (; s < segr; s++) sum += q[s * 32 + 31];
                const size_t p = (size_t)(y0 + o) * g.W + (x0 + xo);
                if (NORM) { Nk_out[p * 2 + (role_hi ? 1 : 0)] = sum * inv; continue; }   // N_k = box(M_{D-1})  (A.cpp:2588)
                const float nk = t < e0 + BLO2_THREADS ? nk_pre : __ldg(&Nk[p * 2 + (role_hi ? 1 : 0)]);
                const float jb = __fdiv_rn(sum * inv, nk);                        // A.cpp:2594
                float cst;
                if (!role_hi) {
                    if (!isl) { part[pix] = __fmul_rn((float)(I - k), jb); continue; }   // (I - lo) * JB_lo
                    cst = jb;                                                     // I is a level: cost = JB_{I,d}
                } else {
                    cst = __fadd_rn(part[pix], __fmul_rn((float)(k - I), jb));    // + (hi - I) * JB_hi  (A.cpp:2666-2667)
                }
                if (agg) agg[(size_t)di * n + p] = cst;
                const unsigned long long key = ((unsigned long long)orderable_u32(cst) << 16) | (unsigned)(d_label0 + di);
                if (key < bestk[pix]) bestk[pix] = key;
            }
            // the next level's running sums overwrite U only (the scan above is behind a barrier); Q, part and
            // bestk are next touched after the next barrier
        }
        __syncthreads();
    }
    for (int p = tid; !NORM && p < NPIX; p += BLO2_THREADS) {
        int o = p / SW, xo = p - o * SW, y = y0 + o, x = x0 + xo;
        if (x < g.W && y < g.H && bestk[p] != WTA_KEY_EMPTY) {
            // back to the library's key format (orderable double cost | d)
            const uint32_t ob = (uint32_t)(bestk[p] >> 16);
            if (ob != 0xFFFFFFFFu) {                                                  // NaN never wins
                const float cst = __uint_as_float((ob & 0x80000000u) ? (ob & 0x7FFFFFFFu) : ~ob);
                atomicMin(&keys[(size_t)y * g.W + x], wta_key(cst, (int)(bestk[p] & 0xFFFFu)));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming BLO(1) aggregation for the dispatcher's level set (sampleRateR = 0.015 -> step 3, 86 levels, A.cpp:70).
//
// One thread owns one halo column of one disparity and walks DOWN a band of rows with the vertical window sums of ALL 86
// levels in registers: V_k = sum over the WIN rows of |L-k| |R_d-k| c_d.  Moving one row down is, per level, the product of
// the entering pixel minus the product of the leaving pixel; the pair is evaluated with two packed instructions
// ((L-k)(R-k) = a - k b + k^2 for both pixels: FFMA2 + FADD2, exact integers) and two FFMAs with |.| operands -- 4 issue
// slots per level and row, no vertical halo (a band is entered once), no shared memory and no barrier in this phase.
// The horizontal window needs the neighbouring columns: the 86 sums of a row go to shared memory ([level][column],
// conflict-free both ways), and every output pixel adds the WIN columns of the (at most) two levels it consumes
// (A.cpp:2656-2667).  Costs are written as a [slices][H][W] volume (4 B per evaluation) for the WTA pass.
// Against k_blo1_agg2: no per-level barrier triple (2 barriers per ROW of 86 levels), 2.65 x halo re-evaluation -> 1.15 x
// (columns only), every level of a row in flight at once (86 independent FMA chains per thread).
// NORM = true computes the normalisers N_k = box(M_{D-1}) with unit costs into the per-pixel pair plane.
// ---------------------------------------------------------------------------------------------
#define BLS_COLS 256
#define BLS_NL 86
#define BLS_STEP 3
#ifndef BLS_VARIANT_NR
#define BLS_VARIANT_NR 12         // levels kept in registers (aggregation at 1280x720x128, ms: 4 -> 5.00, 8 -> 4.92, 12 -> 4.87, 16 / 20 -> 5.07, 24 -> 5.13, 28 -> 5.04, 32 -> 5.08)
#define BLS_VARIANT_NBUF 1        // shared-memory row buffers (2: one CTA per SM, one barrier per row; measured slower)
#endif

// one level of the window update: V += |(L-k)(R-k)|_in c_in - |(L-k)(R-k)|_out c_out  (nco = -c_out)
#define BLS_UPDATE(v, li)                                                                                       \
    {                                                                                                           \
        const float k_ = (float)((li) * BLS_STEP), k2_ = (float)((li) * BLS_STEP * (li) * BLS_STEP);            \
        float2 t_ = __ffma2_rn(make_float2(bi, bo), make_float2(-k_, -k_), make_float2(ai, ao));                \
        t_ = __fadd2_rn(t_, make_float2(k2_, k2_));                /* (L-k)(R-k) of both pixels, exact */        \
        v = fmaf(fabsf(t_.x), ci, v);                              /* + |L-k| |R-k| c  (A.cpp:2571-2583) */      \
        v = fmaf(fabsf(t_.y), nco, v);                                                                          \
    }

template <int WIN, bool NORM, int NR, int NBUF>
__global__ void __launch_bounds__(BLS_COLS, NBUF == 2 ? 1 : 2)
k_blo1_stream(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rpad, const float* __restrict__ cost,
              const float* __restrict__ Nk, float* __restrict__ Nk_out, BloGeom g, int band_rows, int slice0,
              float* __restrict__ vol) {
    constexpr int h = WIN / 2, SW = BLS_COLS - (WIN - 1);
    // levels [0, NR) keep their window sum in a register; the others live in the thread's own cells of the shared row
    // buffer (it receives every sum once per row anyway) and make one extra shared-memory round trip per row: 86 sums
    // plus the working set of the horizontal phase do not fit 128 registers, and a spill would go through L2
    // NBUF = 2 (one CTA per SM, up to 255 registers: NR = 86) double-buffers the row buffer and needs one barrier per row.
    static_assert(NBUF == 1 || NR == BLS_NL, "the double-buffered variant keeps every level in registers");
    extern __shared__ float sm_bls[];                              // [NBUF][BLS_NL][BLS_COLS]
    const int cx = threadIdx.x;
    const int x0 = blockIdx.x * SW;
    const int yb0 = blockIdx.y * band_rows, yb1 = min(g.H, yb0 + band_rows);
    const int di = g.di_lo + blockIdx.z;
    const int H = g.H, W = g.W;
    const size_t n = (size_t)H * W;
    const int sx = border_idx(x0 - h + cx, W, 1);                  // boxFilter BORDER_REFLECT_101
    const uint8_t* lcol = lg + sx;
    const uint8_t* rcol = rpad + (g.x0_base + g.x0_step * di) + sx;
    const float* ccol = cost + (size_t)di * n + sx;
    // (a, b, c) = (L R, L + R, c_d) of one halo pixel
    auto load_px = [&](int r, float& a, float& b, float& c) {
        const int sy = border_idx(r, H, 1);
        const int l = lcol[(size_t)sy * W], rr = rcol[(size_t)sy * g.Wp];
        a = (float)(l * rr); b = (float)(l + rr);
        c = NORM ? 1.0f : __ldg(&ccol[(size_t)sy * W]);
    };
    float* scol = sm_bls + cx;
    float V[NR];
    int buf = 0;
    // window of the band's first row: rows yb0 - h .. yb0 + h, two rows per step (an "update" whose leaving pixel has
    // the sign of an entering one), the last row alone; register levels first, then the shared-memory levels
#pragma unroll
    for (int li = 0; li < NR; li++) V[li] = 0.0f;
    for (int r = yb0 - h; r <= yb0 + h; r += 2) {
        float ai, bi, ci, ao, bo, co;
        load_px(r, ai, bi, ci);
        load_px(r + 1, ao, bo, co);
        const float nco = r + 1 <= yb0 + h ? co : 0.0f;
#pragma unroll
        for (int li = 0; li < NR; li++) BLS_UPDATE(V[li], li)
    }
    if (NR < BLS_NL) {
        float T[BLS_NL - NR + (NR == BLS_NL ? 1 : 0)];
#pragma unroll
        for (int li = NR; li < BLS_NL; li++) T[li - NR] = 0.0f;
        for (int r = yb0 - h; r <= yb0 + h; r += 2) {
            float ai, bi, ci, ao, bo, co;
            load_px(r, ai, bi, ci);
            load_px(r + 1, ao, bo, co);
            const float nco = r + 1 <= yb0 + h ? co : 0.0f;
#pragma unroll
            for (int li = NR; li < BLS_NL; li++) BLS_UPDATE(T[li - NR], li)
        }
#pragma unroll
        for (int li = NR; li < BLS_NL; li++) scol[li * BLS_COLS] = T[li - NR];
    }
    const float inv = 1.0f / (float)(WIN * WIN);
    const int x = x0 + cx - h;                                      // output column of this thread
    const bool consumer = cx >= h && cx < BLS_COLS - h && x < W;
    // raw operands of a halo pixel (consumed one phase later: see the loop)
    auto load_raw = [&](int r, int& l, int& rr, float& c) {
        const int sy = border_idx(r, H, 1);
        l = lcol[(size_t)sy * W]; rr = rcol[(size_t)sy * g.Wp];
        c = NORM ? 1.0f : __ldg(&ccol[(size_t)sy * W]);
    };
    // what the horizontal phase needs from global memory is requested one row ahead
    int I_nx = 0;
    float2 nk_nx = make_float2(1.0f, 1.0f);
    if (consumer) {
        const size_t p0 = (size_t)yb0 * W + x;
        I_nx = lg[p0];
        if (!NORM) nk_nx = __ldg((const float2*)&Nk[p0 * 2]);
    }
    for (int y = yb0; y < yb1; y++) {
        const size_t p = (size_t)y * W + (consumer ? x : 0);
        const int I = I_nx;
        const float2 nk = nk_nx;
        float* const srow = sm_bls + buf * (BLS_NL * BLS_COLS);
#pragma unroll
        for (int li = 0; li < NR; li++) srow[cx + li * BLS_COLS] = V[li];
        __syncthreads();
        // requests that stay in flight during the horizontal phase: the pixels entering / leaving the window of the next
        // row, and the next row's intensity / normaliser pair
        int li_, ri_, lo_, ro_;
        float ci, co;
        load_raw(y + 1 + h, li_, ri_, ci);
        load_raw(y - h, lo_, ro_, co);
        if (consumer && y + 1 < yb1) {
            I_nx = lg[p + W];
            if (!NORM) nk_nx = __ldg((const float2*)&Nk[(p + W) * 2]);
        }
        if (consumer) {
            const int lo = I / BLS_STEP;                           // level index of the lower key (A.cpp:2658)
            const bool isl = (I - lo * BLS_STEP) == 0;             // discretInten membership (A.cpp:2656); 255 = 85 * 3
            const int hi = min(lo + 1, BLS_NL - 1);
            const float* slo = srow + lo * BLS_COLS + (cx - h);
            const float* shi = srow + hi * BLS_COLS + (cx - h);
            // WIN columns of both levels; batches of HB loads bound the registers the loads in flight take
            constexpr int HB = WIN % 7 == 0 ? 7 : (WIN % 5 == 0 ? 5 : 3);
            float s0 = 0.0f, s1 = 0.0f, u0 = 0.0f, u1 = 0.0f;
#pragma unroll 1
            for (int j0 = 0; j0 < WIN / HB * HB; j0 += HB) {
                float a_[HB], b_[HB];
#pragma unroll
                for (int j = 0; j < HB; j++) { a_[j] = slo[j0 + j]; b_[j] = shi[j0 + j]; }
#pragma unroll
                for (int j = 0; j < HB; j++) {
                    if (j & 1) { s1 += a_[j]; u1 += b_[j]; } else { s0 += a_[j]; u0 += b_[j]; }
                }
            }
#pragma unroll
            for (int j = WIN / HB * HB; j < WIN; j++) { s0 += slo[j]; u0 += shi[j]; }
            const float jl = (s0 + s1) * inv, jh = (u0 + u1) * inv;
            if (NORM) {                                            // N_k = box(M_{D-1})  (A.cpp:2588)
                Nk_out[p * 2] = jl;
                if (!isl) Nk_out[p * 2 + 1] = jh;
            } else {
                const float jbl = __fdiv_rn(jl, nk.x);            // A.cpp:2594
                float cst = jbl;                                   // I is a level: cost = JB_{I,d}
                if (!isl) {
                    const int klo = lo * BLS_STEP, khi = min(klo + BLS_STEP, 255);
                    cst = __fadd_rn(__fmul_rn((float)(I - klo), jbl), __fmul_rn((float)(khi - I), __fdiv_rn(jh, nk.y)));   // A.cpp:2666-2667
                }
                vol[(size_t)(di - slice0) * n + p] = cst;
            }
        }
        if (NBUF == 1) __syncthreads(); else buf ^= 1;              // two buffers: the next row's stores go to the other one
        // the raw operands are consumed only here: the compiler may not hoist their first use (and with it the wait for
        // the loads) above the horizontal phase
        asm volatile("" : "+r"(li_), "+r"(ri_), "+r"(lo_), "+r"(ro_));
        if (y + 1 < yb1) {
            const float ai = (float)(li_ * ri_), bi = (float)(li_ + ri_), ao = (float)(lo_ * ro_), bo = (float)(lo_ + ro_);
            const float nco = -co;
#pragma unroll
            for (int li = 0; li < NR; li++) BLS_UPDATE(V[li], li)
            // shared-memory levels in groups of 4: the fence keeps the compiler from hoisting every load of the row
            // (one register each) above the arithmetic
#pragma unroll
            for (int l0 = NR; l0 < BLS_NL; l0 += 4) {
                float v4[4];
#pragma unroll
                for (int q = 0; q < 4; q++) if (l0 + q < BLS_NL) v4[q] = scol[(l0 + q) * BLS_COLS];
#pragma unroll
                for (int q = 0; q < 4; q++) if (l0 + q < BLS_NL) { BLS_UPDATE(v4[q], l0 + q) scol[(l0 + q) * BLS_COLS] = v4[q]; }
                asm volatile("" ::: "memory");
            }
        }
    }
}
#undef BLS_UPDATE

template <int WIN>
static asw_status launch_blo1_agg2(asw_ctx* ctx, const uint8_t* gref, const uint8_t* gtgt, const float* cost, float* Nk,
                                   const BloGeom& g, int min_d, unsigned long long* keys, float* agg_dev) {
    constexpr int TH = WIN + 1, SW = BLO2_COLS - (WIN - 1), NPIX = TH * SW;
    size_t smem = ((size_t)(2 * WIN + 2) * BLO2_PITCH + (size_t)TH * BLO2_PITCH + NPIX) * sizeof(float) +
                  (size_t)NPIX * sizeof(uint32_t) + (size_t)NPIX * sizeof(unsigned long long);
    cudaFuncSetAttribute(k_blo1_agg2<WIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_blo1_agg2<WIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int dch = ctx->tune[ASW_TUNE_BLO_DCH] > 0 ? ctx->tune[ASW_TUNE_BLO_DCH] : 8;
    {   // normalisers: the same kernel on the last disparity with unit costs
        BloGeom gn = g;
        gn.di_lo = g.D - 1; gn.di_hi = g.D;
        dim3 grid(cdiv(g.W, SW), cdiv(g.H, TH), 1);
        LAUNCH(ctx, "blo1_norm", (k_blo1_agg2<WIN, true><<<grid, BLO2_THREADS, smem, ctx->stream>>>(gref, gtgt, cost, nullptr, Nk, gn, 1, min_d, keys, nullptr)));
    }
    dim3 grid(cdiv(g.W, SW), cdiv(g.H, TH), cdiv(g.di_hi - g.di_lo, dch));
    LAUNCH(ctx, "blo1_aggregate", (k_blo1_agg2<WIN, false><<<grid, BLO2_THREADS, smem, ctx->stream>>>(gref, gtgt, cost, Nk, nullptr, g, dch, min_d, keys, agg_dev)));
    return ASW_OK;
}

// streaming kernel (step 3 / 86 levels): normalisers, then the cost volume of slices [di_lo, di_hi), then the WTA pass
template <int WIN>
static asw_status launch_blo1_stream(asw_ctx* ctx, const uint8_t* gref, const uint8_t* gtgt, const float* cost, float* Nk,
                                     const BloGeom& g, int min_d, unsigned long long* keys, float* agg_dev) {
    constexpr int SW = BLS_COLS - (WIN - 1);
    constexpr int NR = BLS_VARIANT_NR, NBUF = BLS_VARIANT_NBUF;
    const size_t smem = (size_t)NBUF * BLS_NL * BLS_COLS * sizeof(float);
    const size_t n = (size_t)g.H * g.W;
    const int cnt = g.di_hi - g.di_lo, strips = cdiv(g.W, SW);
    cudaFuncSetAttribute(k_blo1_stream<WIN, false, NR, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_blo1_stream<WIN, true, NR, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // bands: a band is entered once (WIN - 1 warm-up rows), so they are as tall as the CTA count allows: ~4 waves of
    // 2 CTAs per SM for the aggregation, one wave for the single-slice normaliser launch
    // The band count minimises (waves of resident CTAs) x (rows a CTA walks, warm-up included).
    const int slots = ctx->sm_count * (NBUF == 2 ? 1 : 2);
    auto rows_for = [&](int slices, int) {
        int best_rows = g.H; long long best_cost = -1;
        for (int bands = 1; bands <= std::max(1, g.H / (2 * WIN)); bands++) {
            const int rows = cdiv(g.H, bands);
            const long long cost = (long long)cdiv(strips * slices * cdiv(g.H, rows), slots) * (rows + WIN - 1);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_rows = rows; }
        }
        return best_rows;
    };
    {
        BloGeom gn = g;
        gn.di_lo = g.D - 1; gn.di_hi = g.D;                          // the LAST disparity's M (A.cpp:2588)
        const int br = rows_for(1, 2 * ctx->sm_count);
        LAUNCH(ctx, "blo1_norm", (k_blo1_stream<WIN, true, NR, NBUF><<<dim3(strips, cdiv(g.H, br), 1), BLS_COLS, smem, ctx->stream>>>(
                                     gref, gtgt, cost, nullptr, Nk, gn, br, 0, nullptr)));
    }
    float* vol = agg_dev;                                             // the capture volume is [num_d][H][W]: slice di at di
    int slice0 = 0;
    if (!vol) { ASW_TRY(ws_get(ctx, WS_VOL1, n * (size_t)cnt, &vol)); slice0 = g.di_lo; }
    const int br = rows_for(cnt, 8 * ctx->sm_count);
    LAUNCH(ctx, "blo1_aggregate", (k_blo1_stream<WIN, false, NR, NBUF><<<dim3(strips, cdiv(g.H, br), cnt), BLS_COLS, smem, ctx->stream>>>(
                                      gref, gtgt, cost, Nk, nullptr, g, br, slice0, vol)));
    LAUNCH(ctx, "wta_keys", (k_wta_keys<<<(unsigned)((n / 4 + 256) / 256), 256, 0, ctx->stream>>>(
                                vol + (size_t)(g.di_lo - slice0) * n, cnt, n, min_d + g.di_lo, keys)));
    return ASW_OK;
}

// slices [di_lo, di_hi) of the num_d-slice problem (the whole range: dev_blo1).  agg_dev, if given, is the full
// [num_d][H][W] volume.
static asw_status dev_blo1_range(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, double rate_r,
                                 int win, int min_d, int num_d, int di_lo, int di_hi, float* disp_dev, float* agg_dev) {
    size_t n = (size_t)H * W;
    BloGeom g;
    g.step = (int)(256 * rate_r);                                  // A.cpp:2549
    if (g.step <= 0) return asw_fail(ctx, ASW_ERR_BAD_ARG, "sampleRateR too small: the level step is 0%s%s");
    g.nl = 0;
    for (int i = 0; i < 256; i += g.step) g.nl++;                  // A.cpp:2550-2555
    g.last255 = ((g.nl - 1) * g.step != 255) ? 1 : 0;              // A.cpp:2556-2559
    if (g.last255) g.nl++;
    float* cost;
    uint8_t *gref, *gtgt;
    ASW_TRY(ws_get(ctx, WS_VOL0, n * num_d, &cost));
    if (agg_dev == cost) return asw_fail(ctx, ASW_ERR_BAD_ARG, "internal: capture buffer aliases the cost volume%s%s");
    ASW_TRY(dev_cost_sad_box(ctx, dL, dR, H, W, disp_type, win, min_d, num_d, cost, &gref, &gtgt));   // A.cpp:2531-2546
    ViewGeom v = make_view(dL, dR, H, W, disp_type, min_d, num_d);
    g.H = H; g.W = W; g.Wp = v.Wp; g.win = win; g.h = win / 2; g.D = num_d;
    g.x0_base = v.x0_base; g.x0_step = v.x0_step;
    g.di_lo = di_lo; g.di_hi = di_hi;
    float* Nk;
    ASW_TRY(ws_get(ctx, WS_TMP0, n * 2, &Nk));
    int IW = BLO_TW + win - 1, IH = BLO_TH + win - 1, PP = IW | 1, HP = BLO_TW + 1;
    size_t smem_n = ((size_t)IH * PP + (size_t)IH * HP) * sizeof(int);
    size_t smem_a = ((size_t)IH * PP * 2 + (size_t)IH * HP) * sizeof(float) + (size_t)IH * PP * 2;
    if (smem_a > 220 * 1024) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "BLO1 window too large for the tiled kernel%s%s");
    cudaFuncSetAttribute(k_blo1_norm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_n);
    cudaFuncSetAttribute(k_blo1_aggregate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    dim3 tiles(cdiv(W, BLO_TW), cdiv(H, BLO_TH));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    const bool generic = asw_dev("ASW_BLO_TILED");
    const bool templated = !generic && (win == 5 || win == 7 || win == 9 || win == 15 || win == 25 || win == 35);
    if (!templated)   // exact integer normalisers for the tiled fallback (the register-resident kernel computes its own)
        LAUNCH(ctx, "blo1_norm", (k_blo1_norm<<<dim3(tiles.x, tiles.y, g.nl), BLO_THREADS, smem_n, ctx->stream>>>(gref, gtgt, g, Nk)));
    // the dispatcher's level set (sampleRateR = 0.015: step 3, 86 levels, 255 among them) takes the streaming kernel
    if (templated && g.step == BLS_STEP && g.nl == BLS_NL && !g.last255 && !asw_dev("ASW_BLO_AGG2")) {
        switch (win) {
            case 5: ASW_TRY(launch_blo1_stream<5>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
            case 7: ASW_TRY(launch_blo1_stream<7>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
            case 9: ASW_TRY(launch_blo1_stream<9>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
            case 15: ASW_TRY(launch_blo1_stream<15>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
            case 25: ASW_TRY(launch_blo1_stream<25>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
            default: ASW_TRY(launch_blo1_stream<35>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        }
        return keys_to_disp(ctx, keys, n, disp_dev);
    }
    switch (generic ? 0 : win) {
        case 5: ASW_TRY(launch_blo1_agg2<5>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        case 7: ASW_TRY(launch_blo1_agg2<7>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        case 9: ASW_TRY(launch_blo1_agg2<9>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        case 15: ASW_TRY(launch_blo1_agg2<15>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        case 25: ASW_TRY(launch_blo1_agg2<25>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        case 35: ASW_TRY(launch_blo1_agg2<35>(ctx, gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)); break;
        default:
            LAUNCH(ctx, "blo1_aggregate_tiled", (k_blo1_aggregate<<<dim3(tiles.x, tiles.y, di_hi - di_lo), BLO_THREADS, smem_a, ctx->stream>>>(
                                                    gref, gtgt, cost, Nk, g, min_d, keys, agg_dev)));
    }
    return keys_to_disp(ctx, keys, n, disp_dev);
}
static asw_status dev_blo1(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, int disp_type, double rate_r, int win,
                           int min_d, int num_d, float* disp_dev, float* agg_dev) {
    return dev_blo1_range(ctx, dL, dR, H, W, disp_type, rate_r, win, min_d, num_d, 0, num_d, disp_dev, agg_dev);
}
