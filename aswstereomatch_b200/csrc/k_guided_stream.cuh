// k_guided_stream.cuh -- GuidedF_2 (A.cpp:2766-2854, 2976-3050) as ONE streaming kernel per view:
//     TAD C+G cost -> box -> (a, b) -> box -> q'      (k_gfs_filter, everything on chip)
//     q' -> per-slice NORM_MINMAX affine -> WTA keys   (k_gfs_wta, 4 B per disparity evaluation)
// What it replaces: the two tiled passes of k_guided_fast.cuh, which moved a 16-byte (a,b) record per disparity
// evaluation through HBM, recomputed the cost on a 2-D halo (32x64 inputs for 24x56 outputs) and held both window
// passes' operands in shared memory.
//
// Mapping.  A CTA owns a strip of 64 image columns, NS = 4 consecutive disparity slices and a band of rows, and
// walks down the band K rows per step.  Both box levels are separable with the VERTICAL pass first:
//   A/V1  thread = (slice, column): evaluates the cost of its column for the K new rows and keeps the vertical
//         window sum of (I0 c, I1 c, I2 c, c) in registers -- the last K rows live in a register ring, so a window
//         value is never re-read from anywhere -- and writes the K vertical sums to shared memory
//   H1    thread = (slice, row, run of 8 columns): sliding horizontal window over the vertical sums -> mean_p,
//         corr_Ip -> (a, b) of A.cpp:2805-2847 -> shared memory
//   V2    thread = (slice, column): vertical window sum of (a, b) through a second register ring
//   H2    thread = (slice, row, run of 8 columns): horizontal window -> q' = abar . I + bbar -> HBM (4 B)
// The vertical halo disappears (a band is entered once: K + a + 1 warm-up rows per band instead of K-1 halo rows
// per 24 output rows), the horizontal one is 2(K-1) of 64 columns, and nothing but q' is written.
//
// Borders.  Level 1 filters p with BORDER_REFLECT_101: virtual rows / columns are mapped to source pixels when the
// cost is evaluated.  Level 2 filters the (a,b) IMAGE with BORDER_REFLECT_101, i.e. out-of-image (a,b) are copies of
// in-image (a,b), not values derived from reflected costs: columns are handled by letting the V2 thread of an
// out-of-image column read the mirrored column of the strip; rows by two closed-form steps on the register ring
// (first K rows of the image -> output rows 0..a; last K rows -> output rows H-a..H-1).  Bands are aligned so that
// the top band's ring holds rows 0..K-1 after its first block and the bottom band's last block ends at row H-1.
//
// Numerics: fp32 throughout (packed FADD2/FFMA2 for the window sums, FMA in the (a,b) epilogue); every running sum
// is restarted from the ring once per K rows, so rounding does not accumulate down a column.
#pragma once
#include "k_guided_fast.cuh"

#define GFS_IW 64
#define GFS_NS 4
#define GFS_THREADS 256

// float feature record (FeatF, k_prep.cuh): gradients as exact-integer floats (|g| <= 4080), colour bytes last so that
// (g0, g1) sits in an aligned register pair.  The target plane stores NEGATED gradients: |ga - gb| = |ga + (-gb)| is one
// FADD2 + FADD.  k_features_f writes the records straight from the images.

// guidance record for the (a,b) epilogue, 32 bytes per pixel: {-mean_I0, -mean_I1, -mean_I2, rd2, rd0, rd1, -, -} with
// rd_c = (1/K^2) / (var_c + eps): the level-2 box normalisation is folded into a.
struct __align__(16) GfsMoments { float4 nm; float4 rd; };

// Fused guidance preparation of the streaming path (replaces guide_normalize + box_f32 over 6 planes + guide_finish +
// gfs_pack_guide): I = normalize(guide) (A.cpp:2774), mean_I = box(I), corr_II = box(I*I) (A.cpp:2778, 2796),
// var = corr_II - mean_I^2, den = eps + var (A.cpp:2799, 2846) -> Gi = {I} and the 32-byte moment record.
// One thread per halo column walks down a band of rows with the 6 vertical window sums in double registers
// (one row in, one row out -- box sums of these float values are exact in double, so any summation order rounds to
// the same float as cv::boxFilter's), the horizontal sums run over shared memory.  Only the guide's bytes are read.
#define GFS_GM_COLS 128
__global__ void __launch_bounds__(GFS_GM_COLS)
k_gfs_guide_moments(const uint8_t* __restrict__ guide, const int* __restrict__ mm, int H, int W, int k, float eps,
                    int band_rows, float4* __restrict__ Gi, GfsMoments* __restrict__ out) {
    __shared__ double vs[2][6][GFS_GM_COLS];
    const int a = k / 2, SW = GFS_GM_COLS - (k - 1);
    const int cx = threadIdx.x, x0 = blockIdx.x * SW;
    const int sx = border_idx(x0 - a + cx, W, 1);                 // boxFilter default border: REFLECT_101
    float sf, hf;
    minmax_scale_shift((double)mm[0], (double)mm[1], &sf, &hf);
    const int y_begin = blockIdx.y * band_rows, y_end = min(H, y_begin + band_rows);
    double V[6];
#pragma unroll
    for (int c = 0; c < 6; c++) V[c] = 0.0;
    auto add_row = [&](int y, double sign) {
        const uint8_t* p = guide + ((size_t)border_idx(y, H, 1) * W + sx) * 3;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float v = fmaf((float)p[c], sf, hf);            // cv::normalize
            V[c] += sign * (double)v;
            V[3 + c] += sign * (double)__fmul_rn(v, v);           // guidedImg.mul(guidedImg)
        }
    };
    for (int j = -a; j < a; j++) add_row(y_begin + j, 1.0);
    const double scale = 1.0 / ((double)k * k);
    const float inv = 1.0f / (float)(k * k);
    const int xo = x0 + cx;                                       // output column of this thread
    const bool writer = cx < SW && xo < W;
    int buf = 0;
    for (int y = y_begin; y < y_end; y++) {
        add_row(y + a, 1.0);
#pragma unroll
        for (int c = 0; c < 6; c++) vs[buf][c][cx] = V[c];
        add_row(y - a, -1.0);
        __syncthreads();
        if (writer) {
            float m[3], dn[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                double s1 = 0.0, s2 = 0.0;
                for (int j = 0; j < k; j++) { s1 += vs[buf][c][cx + j]; s2 += vs[buf][3 + c][cx + j]; }
                m[c] = (float)(s1 * scale);
                const float corr = (float)(s2 * scale);
                const float var = __fsub_rn(corr, __fmul_rn(m[c], m[c]));          // A.cpp:2799
                dn[c] = __fadd_rn(__fmul_rn(1.0f, eps), var);                      // A.cpp:2846
            }
            const size_t i = (size_t)y * W + xo;
            const uint8_t* p = guide + i * 3;
            Gi[i] = make_float4(fmaf((float)p[0], sf, hf), fmaf((float)p[1], sf, hf), fmaf((float)p[2], sf, hf), 0.0f);
            GfsMoments r;
            r.nm = make_float4(-m[0], -m[1], -m[2], __fdiv_rn(inv, dn[2]));
            r.rd = make_float4(__fdiv_rn(inv, dn[0]), __fdiv_rn(inv, dn[1]), 0.0f, 0.0f);
            out[i] = r;
        }
        buf ^= 1;
    }
}

struct GfsGeom {
    int H, W, Wp, Wq;          // image, padded target width, q' row pitch (floats) = strips * QW
    int x0_base, x0_step;      // target crop column of slice index di = x0_base + x0_step * di
    int D, nbands;
};
// integer form of the colour truncation (A.cpp:461-465) for integral thresholds + fp32 gradient / blend constants
struct TadStream { int thr_c, add_c; float thr_g, reg_r, reg; };

// c' = cost - c0 = regR * Cc + reg * max(G - T_G, 0)   (same value as tad_cost_prime, k_cost.cuh)
__device__ __forceinline__ float gfs_cost(const FeatF& a, const FeatF& nb, const TadStream& p) {
    uint32_t ad = __vabsdiffu4(a.bgr, nb.bgr);                                   // A.cpp:455
    int s01 = __dp4a(ad, 0x00000101u, 0u);
    int s = __dp4a(ad, 0x00010000u, (uint32_t)min(s01, 255));                    // (c0 + c1) saturates, then + c2
    int color = ((s + 1) * 43691) >> 17;                                         // round(s / 3)
    int cci = (color > p.thr_c) ? min(color + p.add_c, 255) : 0;                 // inverted truncation
    float2 d01 = __fadd2_rn(make_float2(a.g0, a.g1), make_float2(nb.g0, nb.g1));
    float d2 = a.g2 + nb.g2;
    float S = (fabsf(d01.x) + fabsf(d01.y)) + fabsf(d2);                         // exact integer
    float G = S * 0.33333334f;
    return fmaf(p.reg_r, (float)cci, p.reg * fmaxf(G - p.thr_g, 0.0f));
}

// sum of K values as a balanced tree (depth ceil(log2 K) instead of a chain of K-1 dependent adds)
template <int K>
__device__ __forceinline__ float4 gfs_pair_sum(const float4* w) {
    float4 t[(K + 1) / 2];
#pragma unroll
    for (int i = 0; i < K / 2; i++) t[i] = p4add(w[2 * i], w[2 * i + 1]);
    if (K & 1) t[K / 2] = w[K - 1];
#pragma unroll
    for (int n = (K + 1) / 2; n > 1; n = (n + 1) / 2) {
#pragma unroll
        for (int i = 0; i < n / 2; i++) t[i] = p4add(t[2 * i], t[2 * i + 1]);
        if (n & 1) t[n / 2] = t[n - 1];
    }
    return t[0];
}

// ring step: insert `nw` at slot j (a constant after unrolling) of a K-entry register ring and return the sum of the
// ring.  Slot 0 restarts the sum from the ring (K-1 adds), the others slide (s - old + new).
template <int K>
__device__ __forceinline__ float4 gfs_ring_step(float4 (&r)[K], float4& s, float4 nw, int j) {
    if (j == 0) {
        r[0] = nw;
        s = gfs_pair_sum<K>(r);
    } else {
        s = p4slide(s, r[j], nw);
        r[j] = nw;
    }
    return s;
}

// a - b on a float4 (two packed FMAs with a -1 multiplier)
__device__ __forceinline__ float4 p4sub(float4 a, float4 b) {
    const float2 m1 = make_float2(-1.0f, -1.0f);
    float2 lo = __ffma2_rn(make_float2(b.x, b.y), m1, make_float2(a.x, a.y));
    float2 hi = __ffma2_rn(make_float2(b.z, b.w), m1, make_float2(a.z, a.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Window sum over blocks of K elements without a ring of raw values (k_gfs_walk): P[j] holds the PREFIX sum of the
// previous block's elements 0..j.  Element j of the current block enters: the window = elements j+1..K-1 of the previous
// block (its total minus its prefix at j) + elements 0..j of this block (this block's prefix at j), and P[j] becomes this
// block's prefix.  Against a ring (s += new - old; ring[j] = new): the value written back is COMPUTED, so it lands in the
// slot's home register with no copy, the prefix sums restart every block (nothing accumulates, no periodic re-summation)
// and the loop-carried chain is one packed add per element.  j is a constant after unrolling.
template <int K>
__device__ __forceinline__ float4 gfs_prefix_step(float4 (&P)[K], float4 nw, int j) {
    if (j == K - 1) {                      // the previous block has fully left the window
        P[K - 1] = p4add(P[K - 2], nw);
        return P[K - 1];
    }
    const float4 tail = p4sub(P[K - 1], P[j]);
    P[j] = (j == 0) ? nw : p4add(P[j - 1], nw);
    return p4add(tail, P[j]);
}
// raw elements of the block whose prefix sums are in P
template <int K>
__device__ __forceinline__ void gfs_prefix_to_raw(const float4 (&P)[K], float4 (&r)[K]) {
    r[0] = P[0];
#pragma unroll
    for (int j = 1; j < K; j++) r[j] = p4sub(P[j], P[j - 1]);
}

// REFLECT_101 window sum centred at ring index c over a ring that holds image rows base .. base+K-1, where rows
// below `lo` (= ring index of image row 0, or -inf) and above `hi` (ring index of row H-1) are reflected.
template <int K>
__device__ __forceinline__ float4 gfs_reflect_sum(const float4 (&r)[K], int c, bool at_top) {
    constexpr int A = K / 2;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = -A; v <= A; v++) {
        int i = c + v;
        if (at_top) { if (i < 0) i = -i; }                    // ring index 0 = image row 0
        else { if (i > K - 1) i = 2 * (K - 1) - i; }          // ring index K-1 = image row H-1
        t = (v == -A) ? r[i] : p4add(t, r[i]);
    }
    return t;
}

// named barriers (id 0 is __syncthreads)
__device__ __forceinline__ void gfs_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void gfs_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// single REFLECT_101 fold (|overshoot| < len)
__device__ __forceinline__ int gfs_reflect1(int p, int len) { return p < 0 ? -p : (p >= len ? 2 * len - 2 - p : p); }
// mbarrier + TMA 1-D bulk copy (global -> shared, completion counted in bytes on the mbarrier)
__device__ __forceinline__ uint32_t gfs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gfs_mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gfs_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void gfs_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gfs_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gfs_mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(gfs_smem_u32(b)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();      // a lost copy becomes an error, not a hung GPU
    } while (!done);
}
__device__ __forceinline__ void gfs_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gfs_smem_u32(dst)), "l"(src), "r"(bytes), "r"(gfs_smem_u32(b)) : "memory");
}

#define GFS_BAR_FULL 1       // +buf : VS1[buf] written by the cost warps
#define GFS_BAR_EMPTY 3      // +buf : VS1[buf] (and the VS2 aliased onto it) consumed by the filter warps
#define GFS_BAR_FILTER 5     // among the filter warps
#define GFS_BAR_COST 7        // among the cost warps
#define GFS_BAR_FILTER2 6     // among the filter warps that have V2 / H2 items (all but the loader warp)
#define GFS_TP 68            // staged target row pitch (64 + NS - 1, padded)

// Horizontal sliding run of 8 window sums with the loads issued GFS_PF outputs ahead of their use (a filter warp
// has one other warp per scheduler to hide behind, so the shared-memory latency must be covered inside the thread).
#define GFS_PF 5
#define GFS_CPF 2           // cost rows whose feature loads are in flight
template <int K>
__device__ __forceinline__ float4 gfs_tree_sum(const float4* w) {
    return gfs_pair_sum<K>(w);
}
// level 1: window sums of (I c, c) -> (a, b) of A.cpp:2805-2847, a pre-scaled by 1/K^2
template <int K>
__device__ __forceinline__ void gfs_h1(const float4* __restrict__ src, const GfsMoments* __restrict__ gm, int ga_shift,
                                       int nga, float4* __restrict__ dst, int len, float inv) {
    constexpr int N = GFF_RUN + K - 1;
    float4 w[N], nm[GFF_RUN];
    float2 rd[GFF_RUN];
#pragma unroll
    for (int i = 0; i < K - 1 + GFS_PF; i++) w[i] = src[i];
#pragma unroll
    for (int o = 0; o < GFS_PF; o++) {
        const int ix = min(max(ga_shift + o, 0), nga - 1);
        nm[o] = gm[ix].nm; rd[o] = *(const float2*)&gm[ix].rd;
    }
    float4 s;
#pragma unroll
    for (int o = 0; o < GFF_RUN; o++) {
        if (o + GFS_PF < GFF_RUN) {
            w[K - 1 + o + GFS_PF] = src[K - 1 + o + GFS_PF];
            const int ix = min(max(ga_shift + o + GFS_PF, 0), nga - 1);
            nm[o + GFS_PF] = gm[ix].nm; rd[o + GFS_PF] = *(const float2*)&gm[ix].rd;
        }
        s = (o == 0) ? gfs_tree_sum<K>(w) : p4slide(s, w[o - 1], w[o + K - 1]);
        if (o < len) {
            const float mP = s.w;
            // cov = corr_Ip - mean_I * mean_p ; a = cov / (var + eps)                         (A.cpp:2805-2846)
            const float2 cov01 = __ffma2_rn(make_float2(nm[o].x, nm[o].y), make_float2(mP, mP), make_float2(s.x, s.y));
            const float2 a01 = __fmul2_rn(cov01, rd[o]);
            const float a2 = fmaf(nm[o].z, mP, s.z) * nm[o].w;
            // b = mean_p - a . mean_I                                                         (A.cpp:2847)
            const float b = fmaf(a01.x, nm[o].x, fmaf(a01.y, nm[o].y, fmaf(a2, nm[o].z, mP * inv)));
            dst[o] = make_float4(a01.x, a01.y, a2, b);
        }
    }
}
// level 2: window sums of (a, b) -> q' = abar . I + bbar (A.cpp:2852)
template <int K>
__device__ __forceinline__ void gfs_h2(const float4* __restrict__ src, const float4* __restrict__ iq, int iq0, int niq,
                                       float (&q)[GFF_RUN]) {
    constexpr int N = GFF_RUN + K - 1;
    float4 w[N], I[GFF_RUN];
#pragma unroll
    for (int i = 0; i < K - 1 + GFS_PF; i++) w[i] = src[i];
#pragma unroll
    for (int o = 0; o < GFS_PF; o++) I[o] = iq[min(iq0 + o, niq - 1)];
    float4 s;
#pragma unroll
    for (int o = 0; o < GFF_RUN; o++) {
        if (o + GFS_PF < GFF_RUN) {
            w[K - 1 + o + GFS_PF] = src[K - 1 + o + GFS_PF];
            I[o + GFS_PF] = iq[min(iq0 + o + GFS_PF, niq - 1)];
        }
        s = (o == 0) ? gfs_tree_sum<K>(w) : p4slide(s, w[o - 1], w[o + K - 1]);
        q[o] = fmaf(s.x, I[o].x, fmaf(s.y, I[o].y, fmaf(s.z, I[o].z, s.w)));
    }
}

template <int K>
struct GfsLayout {
    static constexpr int A = K / 2;
    static constexpr int AW = GFS_IW - (K - 1);                // (a,b) columns per strip
    static constexpr int QW = GFS_IW - 2 * (K - 1);            // q' columns per strip
    static constexpr int P1 = GFS_IW + 1, P2 = AW | 1, PQ = QW | 1;   // odd float4 pitches
    static constexpr int ROWS = GFS_NS * K;
    // float4 offsets
    static constexpr int oVS1 = 0;                             // [2][ROWS][P1]  vertical sums of (I c, c)
    static constexpr int oAB = oVS1 + 2 * ROWS * P1;           // [ROWS][P2]     (a0,a1,a2,b)
    static constexpr int oVS2 = oAB + ROWS * P2;               // [ROWS][P2]     vertical sums of (a,b)
    static constexpr int oRef = oVS2 + ROWS * P2;              // [2][K][64]     staged reference features
    static constexpr int oTgt = oRef + 2 * K * GFS_IW;         // [2][K][GFS_TP] staged target features (NS slices share them)
    static constexpr int oGM = oTgt + 2 * K * GFS_TP;          // [2][K][P2]     32-byte moment records (2 float4 each)
    static constexpr int oIQ = oGM + 2 * K * P2 * 2;           // [2][K][PQ]     guidance at the q' pixels
    static constexpr int oBar = oIQ + 2 * K * PQ;              // 4 mbarriers
    static constexpr size_t bytes = (size_t)(oBar + 2) * sizeof(float4);
};

// Warp-specialised, TMA-fed: warps 0..7 (cost warps) run A/V1 and own the level-1 rings, warps 8..15 (filter
// warps) run H1, V2, H2 and own the level-2 rings.  No thread issues a global load: the rows of the next block
// (reference / target features for the cost warps; guidance moments and guidance for the filter warps) arrive by 1-D
// bulk copies (cp.async.bulk, completion on an mbarrier) issued one block ahead by one warp of each group, so the
// copy latency is hidden behind a whole block of work.  VS1 is double-buffered between the groups: the cost
// evaluation of block u+1 overlaps the filtering of block u.
template <int K>
__global__ void __launch_bounds__(2 * GFS_THREADS, 1)
k_gfs_filter(const FeatF* __restrict__ ref, const FeatF* __restrict__ tgt, const float4* __restrict__ Gi,
             const GfsMoments* __restrict__ Gmom, const int* __restrict__ guide_mm, GfsGeom g,
             TadStream tp, float c0, float* __restrict__ qv, uint32_t* __restrict__ slice_mm, int strip0) {
    using L = GfsLayout<K>;
    constexpr int A = L::A, AW = L::AW, QW = L::QW, P1 = L::P1, P2 = L::P2, PQ = L::PQ, ROWS = L::ROWS;
    constexpr int NRUN1 = (AW + GFF_RUN - 1) / GFF_RUN, NRUN2 = (QW + GFF_RUN - 1) / GFF_RUN;
    static_assert(ROWS * NRUN1 <= GFS_THREADS && GFS_NS * AW <= GFS_THREADS, "phase does not fit one pass");
    // K = 9: the last filter warp has no V2 / H2 items and stays out of the V2 -> H2 barrier
    constexpr bool LOADER_FREE = GFS_NS * AW <= GFS_THREADS - 32 && ROWS * NRUN2 <= GFS_THREADS - 32;
    extern __shared__ float4 sm_gfs[];
    float4* VS1 = sm_gfs + L::oVS1;
    float4* AB = sm_gfs + L::oAB;
    float4* VS2 = sm_gfs + L::oVS2;
    uint64_t* mbar = (uint64_t*)(sm_gfs + L::oBar);            // [0..1] cost stages, [2..3] filter stages

    const int H = g.H, W = g.W;
    const int x0 = (blockIdx.x + strip0) * QW;                 // first q' column of the strip
    const int d0 = blockIdx.y * GFS_NS;
    // ---- band geometry ----
    const int nb = g.nbands, band = blockIdx.z;
    const int yb0 = (int)(((long long)H * band) / nb), yb1 = (int)(((long long)H * (band + 1)) / nb);
    const bool top = band == 0, bottom = band == nb - 1;
    int a0, U;                                                 // first (a,b) row of block 1, number of (a,b) blocks
    if (bottom) { U = (H - yb0 + A + K - 1) / K; a0 = H - K * U; }
    else { a0 = top ? 0 : yb0 - A; U = (yb1 + A - a0 + K - 1) / K; }
    const float inv = 1.0f / (float)(K * K);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) gfs_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // source columns of the strip: REFLECT_101 only folds the range, so [sxmin, sxmax] is contiguous
    int sxmin, sxmax;
    {   // the fold maps the interval x0-2A .. x0-2A+63 onto an interval whose ends are images of the interval's
        // ends or of the fold points k * (W-1) inside it (narrow images fold more than once)
        const int lo = x0 - 2 * A, hi = lo + GFS_IW - 1, per = W - 1;
        sxmin = min(border_idx(lo, W, 1), border_idx(hi, W, 1));
        sxmax = max(border_idx(lo, W, 1), border_idx(hi, W, 1));
        int k = lo >= 0 ? (lo + per - 1) / per : -((-lo) / per);           // ceil(lo / per)
        for (; k * per <= hi; k++) {
            if (k & 1) sxmax = W - 1; else sxmin = 0;
        }
    }
    const int nref = sxmax - sxmin + 1;
    const int xo_a = g.x0_base + g.x0_step * d0, xo_b = g.x0_base + g.x0_step * min(d0 + GFS_NS - 1, g.D - 1);
    const int xomin = min(xo_a, xo_b), ntgt = nref + abs(xo_a - xo_b);
    FeatF* sRef = (FeatF*)(sm_gfs + L::oRef);
    FeatF* sTgt = (FeatF*)(sm_gfs + L::oTgt);

    if (threadIdx.x < GFS_THREADS) {
        // =========================== cost warps: A / V1, thread = (slice, column) ===========================
        const int tid = threadIdx.x;
        const int sl1 = tid >> 6, c1 = tid & 63;
        const int di1 = min(d0 + sl1, g.D - 1);
        const int sx1 = border_idx(x0 - 2 * A + c1, W, 1);
        const int ia = sx1 - sxmin;
        const int it = ia + (g.x0_base + g.x0_step * di1) - xomin;
        // cv::normalize of the guidance (A.cpp:2774), same expression as k_guide_normalize
        float gsf, ghf;
        minmax_scale_shift((double)guide_mm[0], (double)guide_mm[1], &gsf, &ghf);
        const uint32_t tx_cost = (uint32_t)(K * (nref + ntgt) * 16);
        auto issue_cost = [&](int u) {      // lane 0 of cost warp 0: feature rows of block u -> cost stage u & 1
            const int st = u & 1;
            gfs_mbar_expect_tx(&mbar[st], tx_cost);
            const int p0 = a0 - A - 1 + K * u;
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const int sy = gfs_reflect1(p0 + j, H);
                gfs_bulk_g2s(sRef + (st * K + j) * GFS_IW, ref + (size_t)sy * W + sxmin, nref * 16, &mbar[st]);
                gfs_bulk_g2s(sTgt + (st * K + j) * GFS_TP, tgt + (size_t)sy * g.Wp + sxmin + xomin, ntgt * 16, &mbar[st]);
            }
        };
        if (tid == 0) issue_cost(0);
        float4 r1[K], s1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < K; j++) r1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        float cmin = 3.0e38f, cmax = -3.0e38f;
        for (int u = 0; u <= U; u++) {
            gfs_bar_sync(GFS_BAR_COST, GFS_THREADS);           // every cost thread is done with block u-1
            if (tid == 0 && u + 1 <= U) issue_cost(u + 1);
            const FeatF* rs = sRef + ((u & 1) * K) * GFS_IW + ia;
            const FeatF* ts = sTgt + ((u & 1) * K) * GFS_TP + it;
            float4* vs = VS1 + (u & 1) * ROWS * P1 + sl1 * P1 + c1;   // shared-memory row = j * NS + slice
            gfs_mbar_wait(&mbar[u & 1], (u >> 1) & 1);
            if (u >= 2) gfs_bar_sync(GFS_BAR_EMPTY + (u & 1), 2 * GFS_THREADS);   // the filter warps are done with block u-2
            // the staged rows and VS1 live in the same shared array: the loads are issued GFS_CPF rows ahead by hand
            // (the compiler will not move a shared load above the previous row's store)
            FeatF fa_q[GFS_CPF], fb_q[GFS_CPF];
#pragma unroll
            for (int j = 0; j < GFS_CPF; j++) { fa_q[j] = rs[j * GFS_IW]; fb_q[j] = ts[j * GFS_TP]; }
#pragma unroll
            for (int j = 0; j < K; j++) {
                const FeatF fa = fa_q[j % GFS_CPF];
                const FeatF fb = fb_q[j % GFS_CPF];
                if (j + GFS_CPF < K) { fa_q[j % GFS_CPF] = rs[(j + GFS_CPF) * GFS_IW]; fb_q[j % GFS_CPF] = ts[(j + GFS_CPF) * GFS_TP]; }
                const float cp = gfs_cost(fa, fb, tp);
                cmin = fminf(cmin, cp); cmax = fmaxf(cmax, cp);
                const float cs = cp * inv;
                const float I0 = fmaf((float)(fa.bgr & 0xFF), gsf, ghf), I1 = fmaf((float)((fa.bgr >> 8) & 0xFF), gsf, ghf);
                const float I2 = fmaf((float)((fa.bgr >> 16) & 0xFF), gsf, ghf);
                const float2 p01 = __fmul2_rn(make_float2(I0, I1), make_float2(cs, cs));
                const float4 nw = make_float4(p01.x, p01.y, I2 * cs, cs);
                vs[j * GFS_NS * P1] = gfs_ring_step<K>(r1, s1, nw, j);
            }
            __threadfence_block();
            gfs_bar_arrive(GFS_BAR_FULL + (u & 1), 2 * GFS_THREADS);   // VS1[u & 1] written, staged rows of block u consumed
        }
        // slice min / max of the raw cost: one atomic pair per warp (a warp = 32 columns of one slice)
        for (int o = 16; o > 0; o >>= 1) {
            cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        }
        if ((tid & 31) == 0 && d0 + sl1 < g.D) {
            atomicMin(&slice_mm[2 * di1], orderable_u32(__fadd_rn(c0, cmin)));
            atomicMax(&slice_mm[2 * di1 + 1], orderable_u32(__fadd_rn(c0, cmax)));
        }
    } else {
        // =========================== filter warps: H1, V2, H2 ===========================
        const int tid = threadIdx.x - GFS_THREADS;
        GfsMoments* sGM = (GfsMoments*)(sm_gfs + L::oGM);
        float4* sIQ = sm_gfs + L::oIQ;
        // staged column ranges: (a,b) columns x0-A .. x0-A+AW-1 clamped into the image, q' columns x0 .. clamped
        const int xamin = max(x0 - A, 0), xamax = min(x0 - A + AW - 1, W - 1);
        const int nga = xamax - xamin + 1, niq = min(W, x0 + QW) - x0;
        const uint32_t tx_filter = (uint32_t)(K * (2 * nga + niq) * 16);
        const int n_iter = U + (bottom ? 1 : 0);
        const int lane = tid & 31;
        // every operand of a bulk copy is CTA-uniform: one lane walks the rows (a per-lane copy would be serialised
        // through the uniform datapath anyway)
        auto issue_filter = [&](int u) {    // moment / guidance rows of block u -> filter stage u & 1
            if (lane != 0) return;
            const int st = u & 1;
            gfs_mbar_expect_tx(&mbar[2 + st], tx_filter);
            const int abase = a0 + K * (u - 1);
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const int ya = min(max(abase + j, 0), H - 1), rq = min(max(abase - A + j, 0), H - 1);
                gfs_bulk_g2s(sGM + (st * K + j) * P2, Gmom + (size_t)ya * W + xamin, nga * 32, &mbar[2 + st]);
                gfs_bulk_g2s(sIQ + (st * K + j) * PQ, Gi + (size_t)rq * W + x0, niq * 16, &mbar[2 + st]);
            }
        };
        const bool loader = tid >= GFS_THREADS - 32;
        // V2 role: (slice, (a,b) column); out-of-image columns read their mirror column
        const int sl2 = tid / AW, c2 = tid - sl2 * AW;
        const bool v2_on = tid < GFS_NS * AW;
        const int csrc = min(max(border_idx(x0 - A + c2, W, 1) - (x0 - A), 0), AW - 1);
        float4 r2[K], s2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < K; j++) r2[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        // H roles: (run, shared-memory row) with consecutive lanes on consecutive rows.  Row = j * NS + slice: a
        // quarter-warp then spans two image rows, so its loads of the per-pixel operands (identical for the NS
        // slices) touch two addresses instead of eight
        const int hrun = tid / ROWS, hrow = tid - hrun * ROWS;
        const int hj = hrow / GFS_NS, hsl = hrow - hj * GFS_NS;
        const bool h_slice_ok = d0 + hsl < g.D;
        const int ga_shift = (x0 - A) - xamin + hrun * GFF_RUN;  // staged index of the run's first (a,b) column, before clamping
        for (int u = 0; u <= n_iter; u++) {
            const int abase = a0 + K * (u - 1);                // first (a,b) row of this block
            if (u <= U) gfs_bar_sync(GFS_BAR_FULL + (u & 1), 2 * GFS_THREADS);   // also: the cost warps are done with block u
            else gfs_bar_sync(GFS_BAR_FILTER, GFS_THREADS);    // every filter thread is done with block u-1
            if (u >= 1) gfs_mbar_wait(&mbar[2 + (u & 1)], ((u - 1) >> 1) & 1);   // blocks 1, 2 are the first use of stages 1, 0
            // ---- H1: horizontal window + (a,b) epilogue ----
            if (u >= 1 && u <= U && hrun < NRUN1) {
                float4* dst = AB + hrow * P2 + hrun * GFF_RUN;
                const float4* src = VS1 + (u & 1) * ROWS * P1 + hrow * P1 + hrun * GFF_RUN;
                const GfsMoments* gm = sGM + ((u & 1) * K + hj) * P2;
                gfs_h1<K>(src, gm, ga_shift, nga, dst, min(GFF_RUN, AW - hrun * GFF_RUN), inv);
            }
            if (u + 2 <= U) { __threadfence_block(); gfs_bar_arrive(GFS_BAR_EMPTY + (u & 1), 2 * GFS_THREADS); }   // VS1[u & 1] consumed
            if (u >= 1) gfs_bar_sync(GFS_BAR_FILTER, GFS_THREADS);
            if (loader) {
                if (u + 1 <= n_iter) issue_filter(u + 1);      // filter stage (u+1) & 1: last read by H1 / H2 of block u-1
            }
            if (u == 0) continue;                              // block 0 only warms the level-1 rings up
            // ---- V2: vertical window over (a,b) ----
            if (v2_on) {
                const float4* src = AB + sl2 * P2 + csrc;
                float4* dst = VS2 + sl2 * P2 + c2;
                if (u == U + 1) {
                    // bottom of the image: ring = rows H-K .. H-1; output rows H-a .. H-1 (block rows 0 .. a-1)
#pragma unroll
                    for (int i = 0; i < A; i++) dst[i * GFS_NS * P2] = gfs_reflect_sum<K>(r2, K - A + i, false);
                } else if (top && u == 1) {
                    // top of the image: ring <- rows 0 .. K-1; output rows 0 .. a (block rows a .. K-1)
#pragma unroll
                    for (int j = 0; j < K; j++) r2[j] = src[j * GFS_NS * P2];
#pragma unroll
                    for (int i = 0; i <= A; i++) dst[(A + i) * GFS_NS * P2] = gfs_reflect_sum<K>(r2, i, true);
                } else {
                    float4 in[K];                       // AB and VS2 alias for the compiler: all loads first
#pragma unroll
                    for (int j = 0; j < K; j++) in[j] = src[j * GFS_NS * P2];
#pragma unroll
                    for (int j = 0; j < K; j++) dst[j * GFS_NS * P2] = gfs_ring_step<K>(r2, s2, in[j], j);
                }
            }
            if (!LOADER_FREE || !loader) gfs_bar_sync(GFS_BAR_FILTER2, LOADER_FREE ? GFS_THREADS - 32 : GFS_THREADS);
            // ---- H2: horizontal window + q' ----
            if (hrun < NRUN2 && h_slice_ok) {
                const int rq = abase - A + hj;
                if (rq >= yb0 && rq < yb1) {      // also drops the unused block rows of the two closed-form steps
                    const int xq0 = x0 + hrun * GFF_RUN;
                    const int len = min(GFF_RUN, QW - hrun * GFF_RUN);
                    const float4* iq = sIQ + ((u & 1) * K + hj) * PQ;
                    float q[GFF_RUN];
                    gfs_h2<K>(VS2 + hrow * P2 + hrun * GFF_RUN, iq, hrun * GFF_RUN, niq, q);
                    float4* out = (float4*)(qv + ((size_t)rq * g.D + (d0 + hsl)) * g.Wq + xq0);
                    out[0] = make_float4(q[0], q[1], q[2], q[3]);
                    if (len > 4) out[1] = make_float4(q[4], q[5], q[6], q[7]);
                }
            }
            // the next block's H1 rewrites AB (last read by V2, one filter barrier back); VS2 is rewritten after the next
            // block's first filter barrier, which every filter thread reaches only after this H2
        }
    }
}

#ifdef ASW_DEV_KERNELS
// ------------------------------------------------------------------------------------------------------------------
// k_gfs_walk -- the same filter for INTERIOR strips (every (a,b) column of the strip inside the image), K = 9.
// DEV ONLY (ASW_GFS_WALK=1 in a -DASW_DEV_KERNELS build): bit-compatible with the parity tests, measured SLOWER than the
// classic kernel (5.57 + 0.33 ms against 5.38 ms per 1080p x 256 view): the walkers carry a third of the CTA's instructions
// in 4 of its 18 warps at 18 of 32 lanes and issue at IPC 0.26 (ncu: profiles/r02_ncu_gfs.md).
// The classic kernel moves every level through shared memory twice over: V1 -> [VS1] -> H1 (runs of 8: every value read
// twice) -> [AB] -> V2 -> [VS2] -> H2 (read twice) -> q'.  ncu (profiles/r02_ncu_gfs.md): the H phases run at the speed of
// the shared-memory pipe.  Here level 2 runs horizontal-first and ONE thread walks a whole strip row:
//     cost warps (8)   A / V1 as before                          -> VS1   (16 B stored per evaluation)
//     walker warps (2) lane = (row of the block, slice): for every column of the row, in registers: level-1 horizontal
//                      window (prefix sums over blocks of K columns) -> (a,b) epilogue -> level-2 horizontal window
//                                                                 -> HS    (VS1 read ONCE; no AB, no second H read)
//     column warps (6) thread = (slice, q' column): level-2 vertical window (prefix sums over the K rows of a block)
//                      -> q' = abar . I + bbar -> HBM, coalesced along the row
// Shared-memory traffic per evaluation: 16 + 16 (VS1) + 28 (moments) + 16 + 16 (HS) + 16 (guidance) instead of
// 16 + 32 + 28 + 16 + 16 + 16 + 32 + 16.  Level 2 borders: rows as in the classic kernel (closed forms on the column
// prefix sums); columns never (interior strips).  Edge strips and K = 5, 7 run the classic kernel.
// ------------------------------------------------------------------------------------------------------------------
#define GFW_WALKERS 128       // 4 warps: (row half) x (half of the block's rows)
#define GFW_COLS 192          // 6 warps: NS * QW for K = 9
#define GFW_BAR_FULL 1        // +buf  VS1[buf] written: cost warps arrive, walkers sync
#define GFW_BAR_EMPTY 3       // +buf  VS1[buf] consumed: walkers arrive, cost warps sync
#define GFW_BAR_HSFULL 5      // +buf  HS[buf] written: walkers arrive, column warps sync
#define GFW_BAR_COST 7        // among the cost warps
#define GFW_BAR_HSEMPTY 8     // +buf  HS[buf] consumed: column warps arrive, walkers sync
#define GFW_BAR_COLS 10       // among the column warps
#define GFW_BAR_WALK 11       // among the walkers
#define GFW_PF 2              // columns the walker's loads run ahead
// 5 warpgroups: 2 cost, 1 walkers, 2 column (6 of its 8 warps have work).  640 threads start with 96 registers each; the
// roles then re-divide the register file (setmaxnreg, per warpgroup): 256 x 88 + 128 x 128 + 256 x 88 = 640 x 96
#define GFW_THREADS 640
#define GFW_REG_COST 88
#define GFW_REG_WALK 128
#define GFW_REG_COLS 88
#define GFW_STR2(x) #x
#define GFW_STR(x) GFW_STR2(x)

template <int K>
struct GfwLayout {
    static constexpr int A = K / 2;
    static constexpr int AW = GFS_IW - (K - 1), QW = GFS_IW - 2 * (K - 1);
    static constexpr int P1 = GFS_IW + 1, P2 = AW | 1, PQ = QW | 1;
    static constexpr int ROWS = GFS_NS * K;
    static constexpr int oVS1 = 0;                             // [2][ROWS][P1]
    static constexpr int oHS = oVS1 + 2 * ROWS * P1;           // [2][ROWS][PQ]  level-2 horizontal sums
    static constexpr int oRef = oHS + 2 * ROWS * PQ;           // staging as in GfsLayout
    static constexpr int oTgt = oRef + 2 * K * GFS_IW;
    static constexpr int oGM = oTgt + 2 * K * GFS_TP;
    static constexpr int oIQ = oGM + 2 * K * P2 * 2;
    static constexpr int oBar = oIQ + 2 * K * PQ;
    static constexpr size_t bytes = (size_t)(oBar + 3) * sizeof(float4);   // 6 mbarriers
};

// NSTEPS columns t0 .. t0+NSTEPS-1 of a walker's row (t0 a multiple of K, so the prefix slot of column t0+j is j).
// MODE 0: the first K columns (only the last one completes a level-1 window), 1: the next K (only the last two complete
// a level-2 window), 2: steady state.  Every queue slot is assigned unconditionally (clamped indices), so that only the
// GFW_PF slots in flight are live.
// 16-byte shared store as two 8-byte stores: the two halves of a packed result sit in two register pairs, and a
// 16-byte store would first move them into an aligned quad (4 MOVs)
__device__ __forceinline__ void gfs_sts_pairs(float4* p, float4 v) {
    const uint32_t a = gfs_smem_u32(p);
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n\tst.shared.v2.f32 [%0+8], {%3, %4};"
                 ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int K, int MODE, int NSTEPS, int NCOLS>
__device__ __forceinline__ void gfw_steps(const float4* __restrict__ src, const GfsMoments* __restrict__ gm, float4* __restrict__ dst,
                                          int t0, float inv, float4 (&Pa)[K], float4 (&Pb)[K], float4 (&vq)[K], float4 (&nmq)[K],
                                          float2 (&rdq)[K]) {
    constexpr int AW = NCOLS - (K - 1);
#pragma unroll
    for (int j = 0; j < NSTEPS; j++) {
        const int t = t0 + j;                                  // column of the walk (VS1 index relative to its first column)
        // loads of column t + PF: issued before this column's store (ptxas keeps shared loads behind shared stores)
        const int tn = min(t + GFW_PF, NCOLS - 1);
        const int cn = min(max(tn - (K - 1), 0), AW - 1);
        vq[(j + GFW_PF) % K] = src[tn];
        nmq[(j + GFW_PF) % K] = gm[cn].nm;
        rdq[(j + GFW_PF) % K] = *(const float2*)&gm[cn].rd;
        const float4 s = gfs_prefix_step<K>(Pa, vq[j], j);     // sum of VS1 columns t-K+1 .. t
        if (MODE > 0 || j >= K - 1) {
            const float4 nm = nmq[j];
            const float2 rd = rdq[j];
            const float mP = s.w;
            // cov = corr_Ip - mean_I * mean_p ; a = cov / (var + eps) ; b = mean_p - a . mean_I      (A.cpp:2805-2847)
            const float2 cov01 = __ffma2_rn(make_float2(nm.x, nm.y), make_float2(mP, mP), make_float2(s.x, s.y));
            const float2 a01 = __fmul2_rn(cov01, rd);
            const float a2 = fmaf(nm.z, mP, s.z) * nm.w;
            const float b = fmaf(a01.x, nm.x, fmaf(a01.y, nm.y, fmaf(a2, nm.z, mP * inv)));
            const float4 s2 = gfs_prefix_step<K>(Pb, make_float4(a01.x, a01.y, a2, b), (j + 1) % K);
            if (MODE == 2 || (MODE == 1 && j >= K - 2)) gfs_sts_pairs(dst + t - 2 * (K - 1), s2);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(GFW_THREADS, 1)
k_gfs_walk(const FeatF* __restrict__ ref, const FeatF* __restrict__ tgt, const float4* __restrict__ Gi,
           const GfsMoments* __restrict__ Gmom, const int* __restrict__ guide_mm, GfsGeom g,
           TadStream tp, float c0, float* __restrict__ qv, uint32_t* __restrict__ slice_mm, int strip0) {
    using L = GfwLayout<K>;
    constexpr int A = L::A, AW = L::AW, QW = L::QW, P1 = L::P1, P2 = L::P2, PQ = L::PQ, ROWS = L::ROWS;
    static_assert(GFS_NS * QW == GFW_COLS && ROWS <= 64 && ROWS % 2 == 0 && QW % 2 == 0, "role sizes");
    constexpr int NCW = GFS_THREADS + GFW_WALKERS;             // participants of the cost <-> walker barriers
    constexpr int NWC = GFW_WALKERS + GFW_COLS;                // participants of the walker <-> column barriers
    extern __shared__ float4 sm_gfs[];
    float4* VS1 = sm_gfs + L::oVS1;
    float4* HS = sm_gfs + L::oHS;
    uint64_t* mbar = (uint64_t*)(sm_gfs + L::oBar);            // [0..1] cost stages, [2..3] moment stages, [4..5] guidance stages

    const int H = g.H, W = g.W;
    const int x0 = (blockIdx.x + strip0) * QW;                 // first q' column of the strip
    const int d0 = blockIdx.y * GFS_NS;
    const int nb = g.nbands, band = blockIdx.z;
    const int yb0 = (int)(((long long)H * band) / nb), yb1 = (int)(((long long)H * (band + 1)) / nb);
    const bool top = band == 0, bottom = band == nb - 1;
    int a0, U;                                                 // first (a,b) row of block 1, number of (a,b) blocks
    if (bottom) { U = (H - yb0 + A + K - 1) / K; a0 = H - K * U; }
    else { a0 = top ? 0 : yb0 - A; U = (yb1 + A - a0 + K - 1) / K; }
    const int n_iter = U + (bottom ? 1 : 0);
    const float inv = 1.0f / (float)(K * K);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 6; i++) gfs_mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    FeatF* sRef = (FeatF*)(sm_gfs + L::oRef);
    FeatF* sTgt = (FeatF*)(sm_gfs + L::oTgt);
    GfsMoments* sGM = (GfsMoments*)(sm_gfs + L::oGM);
    float4* sIQ = sm_gfs + L::oIQ;

    if (threadIdx.x < GFS_THREADS) {
        // =========================== cost warps: A / V1, thread = (slice, column) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " GFW_STR(GFW_REG_COST) ";");   // 96 at launch
        int sxmin, sxmax;
        {   // source columns of the strip (see k_gfs_filter)
            const int lo = x0 - 2 * A, hi = lo + GFS_IW - 1, per = W - 1;
            sxmin = min(border_idx(lo, W, 1), border_idx(hi, W, 1));
            sxmax = max(border_idx(lo, W, 1), border_idx(hi, W, 1));
            int k = lo >= 0 ? (lo + per - 1) / per : -((-lo) / per);
            for (; k * per <= hi; k++) {
                if (k & 1) sxmax = W - 1; else sxmin = 0;
            }
        }
        const int nref = sxmax - sxmin + 1;
        const int xo_a = g.x0_base + g.x0_step * d0, xo_b = g.x0_base + g.x0_step * min(d0 + GFS_NS - 1, g.D - 1);
        const int xomin = min(xo_a, xo_b), ntgt = nref + abs(xo_a - xo_b);
        const int tid = threadIdx.x;
        const int sl1 = tid >> 6, c1 = tid & 63;
        const int di1 = min(d0 + sl1, g.D - 1);
        const int sx1 = border_idx(x0 - 2 * A + c1, W, 1);
        const int ia = sx1 - sxmin;
        const int it = ia + (g.x0_base + g.x0_step * di1) - xomin;
        float gsf, ghf;
        minmax_scale_shift((double)guide_mm[0], (double)guide_mm[1], &gsf, &ghf);
        const uint32_t tx_cost = (uint32_t)(K * (nref + ntgt) * 16);
        auto issue_cost = [&](int u) {
            const int st = u & 1;
            gfs_mbar_expect_tx(&mbar[st], tx_cost);
            const int p0 = a0 - A - 1 + K * u;
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const int sy = gfs_reflect1(p0 + j, H);
                gfs_bulk_g2s(sRef + (st * K + j) * GFS_IW, ref + (size_t)sy * W + sxmin, nref * 16, &mbar[st]);
                gfs_bulk_g2s(sTgt + (st * K + j) * GFS_TP, tgt + (size_t)sy * g.Wp + sxmin + xomin, ntgt * 16, &mbar[st]);
            }
        };
        if (tid == 0) issue_cost(0);
        float4 r1[K], s1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < K; j++) r1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        float cmin = 3.0e38f, cmax = -3.0e38f;
        for (int u = 0; u <= U; u++) {
            gfs_bar_sync(GFW_BAR_COST, GFS_THREADS);
            if (tid == 0 && u + 1 <= U) issue_cost(u + 1);
            const FeatF* rs = sRef + ((u & 1) * K) * GFS_IW + ia;
            const FeatF* ts = sTgt + ((u & 1) * K) * GFS_TP + it;
            float4* vs = VS1 + (u & 1) * ROWS * P1 + sl1 * P1 + c1;
            gfs_mbar_wait(&mbar[u & 1], (u >> 1) & 1);
            if (u >= 2) gfs_bar_sync(GFW_BAR_EMPTY + (u & 1), NCW);
            FeatF fa_q[GFS_CPF], fb_q[GFS_CPF];
#pragma unroll
            for (int j = 0; j < GFS_CPF; j++) { fa_q[j] = rs[j * GFS_IW]; fb_q[j] = ts[j * GFS_TP]; }
#pragma unroll
            for (int j = 0; j < K; j++) {
                const FeatF fa = fa_q[j % GFS_CPF];
                const FeatF fb = fb_q[j % GFS_CPF];
                if (j + GFS_CPF < K) { fa_q[j % GFS_CPF] = rs[(j + GFS_CPF) * GFS_IW]; fb_q[j % GFS_CPF] = ts[(j + GFS_CPF) * GFS_TP]; }
                const float cp = gfs_cost(fa, fb, tp);
                cmin = fminf(cmin, cp); cmax = fmaxf(cmax, cp);
                const float cs = cp * inv;
                const float I0 = fmaf((float)(fa.bgr & 0xFF), gsf, ghf), I1 = fmaf((float)((fa.bgr >> 8) & 0xFF), gsf, ghf);
                const float I2 = fmaf((float)((fa.bgr >> 16) & 0xFF), gsf, ghf);
                const float2 p01 = __fmul2_rn(make_float2(I0, I1), make_float2(cs, cs));
                const float4 nw = make_float4(p01.x, p01.y, I2 * cs, cs);
                vs[j * GFS_NS * P1] = gfs_ring_step<K>(r1, s1, nw, j);
            }
            __threadfence_block();
            gfs_bar_arrive(GFW_BAR_FULL + (u & 1), NCW);
        }
        for (int o = 16; o > 0; o >>= 1) {
            cmin = fminf(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
            cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        }
        if ((tid & 31) == 0 && d0 + sl1 < g.D) {
            atomicMin(&slice_mm[2 * di1], orderable_u32(__fadd_rn(c0, cmin)));
            atomicMax(&slice_mm[2 * di1 + 1], orderable_u32(__fadd_rn(c0, cmax)));
        }
    } else if (threadIdx.x < GFS_THREADS + GFW_WALKERS) {
        // =========================== walkers: H1 -> (a,b) -> H2 along one strip row ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " GFW_STR(GFW_REG_WALK) ";");
        // a walker = (row of the block, slice, half of the row): the left half turns VS1 columns 0 .. QW/2+2(K-1)-1 into
        // q' columns 0 .. QW/2-1, the right half starts QW/2 columns further (the 2(K-1) columns in between are walked twice:
        // four warps instead of two on the critical path of the CTA)
        const int wt = threadIdx.x - GFS_THREADS;
        constexpr int TPW = ROWS / 2;                          // rows per walker warp
        constexpr int NCOLS = QW / 2 + 2 * (K - 1);            // columns a walker reads
        const int ww = wt >> 5, wl = wt & 31;
        const bool w_on = wl < TPW;
        const int task = min((ww & 1) * TPW + wl, ROWS - 1);   // shared-memory row = j * NS + slice
        const int c_off = (ww >> 1) * (QW / 2);
        const int hj = task / GFS_NS;
        // the walkers stage their own operand (the moment records of the block's K rows) two blocks ahead: a stage is refilled
        // as soon as every walker is done with it
        const uint32_t tx_gm = (uint32_t)(K * AW * 32);
        auto issue_gm = [&](int u) {
            const int st = u & 1;
            gfs_mbar_expect_tx(&mbar[2 + st], tx_gm);
            const int abase = a0 + K * (u - 1);
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const int ya = min(max(abase + j, 0), H - 1);
                gfs_bulk_g2s(sGM + (st * K + j) * P2, Gmom + (size_t)ya * W + (x0 - A), AW * 32, &mbar[2 + st]);
            }
        };
        if (wt == 0) { issue_gm(1); if (2 <= U) issue_gm(2); }
        for (int u = 0; u <= U; u++) {
            gfs_bar_sync(GFW_BAR_FULL + (u & 1), NCW);         // VS1[u & 1] written
            if (u >= 1) {
                gfs_mbar_wait(&mbar[2 + (u & 1)], ((u - 1) >> 1) & 1);           // moments of block u
                if (u >= 3) gfs_bar_sync(GFW_BAR_HSEMPTY + (u & 1), NWC);        // column warps done with HS[u & 1] (block u-2)
                if (w_on) {
                    const float4* src = VS1 + (u & 1) * ROWS * P1 + task * P1 + c_off;
                    const GfsMoments* gm = sGM + ((u & 1) * K + hj) * P2 + c_off;   // interior strip: staged index = (a,b) column
                    float4* dst = HS + (u & 1) * ROWS * PQ + task * PQ + c_off;
                    float4 Pa[K], Pb[K];                       // prefix sums of the previous K columns: level 1, level 2
#pragma unroll
                    for (int j = 0; j < K; j++) { Pa[j] = make_float4(0.f, 0.f, 0.f, 0.f); Pb[j] = Pa[j]; }
                    float4 vq[K], nmq[K];
                    float2 rdq[K];
#pragma unroll
                    for (int j = 0; j < GFW_PF; j++) {
                        vq[j] = src[j];
                        nmq[j] = gm[0].nm; rdq[j] = *(const float2*)&gm[0].rd;
                    }
                    constexpr int NFULL = NCOLS / K, REM = NCOLS - NFULL * K;
                    static_assert(NFULL >= 2, "walk shorter than two blocks");
                    gfw_steps<K, 0, K, NCOLS>(src, gm, dst, 0, inv, Pa, Pb, vq, nmq, rdq);
                    gfw_steps<K, 1, K, NCOLS>(src, gm, dst, K, inv, Pa, Pb, vq, nmq, rdq);
#pragma unroll 1
                    for (int m = 2; m < NFULL; m++) gfw_steps<K, 2, K, NCOLS>(src, gm, dst, m * K, inv, Pa, Pb, vq, nmq, rdq);
                    if (REM > 0) gfw_steps<K, 2, REM, NCOLS>(src, gm, dst, NFULL * K, inv, Pa, Pb, vq, nmq, rdq);
                }
                __threadfence_block();
                gfs_bar_arrive(GFW_BAR_HSFULL + (u & 1), NWC);
                if (u + 2 <= U) {
                    gfs_bar_sync(GFW_BAR_WALK, GFW_WALKERS);   // every walker is done with the moment stage of block u
                    if (wt == 0) issue_gm(u + 2);
                }
            }
            if (u + 2 <= U) { __threadfence_block(); gfs_bar_arrive(GFW_BAR_EMPTY + (u & 1), NCW); }
        }
    } else {
        // =========================== column warps: V2 + q', thread = (slice, q' column) ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " GFW_STR(GFW_REG_COLS) ";");
        const int tid = threadIdx.x - GFS_THREADS - GFW_WALKERS;
        if (tid >= GFW_COLS) return;                           // the warpgroup's two spare warps
        const int sl2 = tid / QW, cq = tid - sl2 * QW;
        const int d = d0 + sl2;
        const bool d_ok = d < g.D;
        const int niq = min(W, x0 + QW) - x0;
        const uint32_t tx_iq = (uint32_t)(K * niq * 16);
        auto issue_iq = [&](int u) {        // guidance rows of block u -> guidance stage u & 1
            const int st = u & 1;
            gfs_mbar_expect_tx(&mbar[4 + st], tx_iq);
            const int abase = a0 + K * (u - 1);
#pragma unroll 1
            for (int j = 0; j < K; j++) {
                const int rq = min(max(abase - A + j, 0), H - 1);
                gfs_bulk_g2s(sIQ + (st * K + j) * PQ, Gi + (size_t)rq * W + x0, niq * 16, &mbar[4 + st]);
            }
        };
        if (tid == 0) { issue_iq(1); if (2 <= n_iter) issue_iq(2); }
        float4 r2[K];                                          // column prefix sums of the previous block
#pragma unroll
        for (int j = 0; j < K; j++) r2[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int u = 1; u <= n_iter; u++) {
            const int abase = a0 + K * (u - 1);
            if (u <= U) gfs_bar_sync(GFW_BAR_HSFULL + (u & 1), NWC);             // walkers done with block u
            gfs_mbar_wait(&mbar[4 + (u & 1)], ((u - 1) >> 1) & 1);               // guidance rows of block u
            const float4* src = HS + (u & 1) * ROWS * PQ + sl2 * PQ + cq;        // + j * NS * PQ
            const float4* iq = sIQ + ((u & 1) * K) * PQ + min(cq, niq - 1);
            float* qrow = qv + ((size_t)(abase - A) * g.D + d) * g.Wq + x0 + cq;  // output row of block row 0
            const size_t qpitch = (size_t)g.D * g.Wq;
            auto emit = [&](int j, float4 o) {                 // q' = abar . I + bbar (A.cpp:2852)
                const int rq = abase - A + j;
                if (rq >= yb0 && rq < yb1 && d_ok) {
                    const float4 I = iq[j * PQ];
                    qrow[(size_t)j * qpitch] = fmaf(o.x, I.x, fmaf(o.y, I.y, fmaf(o.z, I.z, o.w)));
                }
            };
            if (u == U + 1) {
                // bottom of the image: the last block = rows H-K .. H-1; output rows H-a .. H-1 (block rows 0 .. a-1)
                float4 raw[K];
                gfs_prefix_to_raw<K>(r2, raw);
#pragma unroll
                for (int i = 0; i < A; i++) emit(i, gfs_reflect_sum<K>(raw, K - A + i, false));
            } else {
                float4 in[K];
#pragma unroll
                for (int j = 0; j < K; j++) in[j] = src[j * GFS_NS * PQ];
                if (u + 2 <= U) { __threadfence_block(); gfs_bar_arrive(GFW_BAR_HSEMPTY + (u & 1), NWC); }   // HS[u & 1] consumed
                if (top && u == 1) {
                    // top of the image: rows 0 .. K-1; output rows 0 .. a (block rows a .. K-1)
#pragma unroll
                    for (int i = 0; i <= A; i++) emit(A + i, gfs_reflect_sum<K>(in, i, true));
                    r2[0] = in[0];
#pragma unroll
                    for (int j = 1; j < K; j++) r2[j] = p4add(r2[j - 1], in[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < K; j++) emit(j, gfs_prefix_step<K>(r2, in[j], j));
                }
            }
            gfs_bar_sync(GFW_BAR_COLS, GFW_COLS);              // every column thread is done with the stage of block u
            if (tid == 0 && u + 2 <= n_iter) issue_iq(u + 2);
        }
    }
}

#endif  // ASW_DEV_KERNELS

// per-slice affine of cv::normalize (A.cpp:2775) applied after the (linear) filter: q = sf * q' + (c0 * sf + hf)
__global__ void k_gfs_affine(const uint32_t* __restrict__ slice_mm, int D, float c0, float2* __restrict__ aff) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float sf, hf;
    minmax_scale_shift((double)from_orderable(slice_mm[2 * d]), (double)from_orderable(slice_mm[2 * d + 1]), &sf, &hf);
    aff[d] = make_float2(sf, (float)fma((double)c0, (double)sf, (double)hf));
}

// q' volume [H][D][Wq] -> normalised costs -> WTA keys (strict <, ascending d, NaN / inf never win; A.cpp:3032-3048).
// One thread = 4 adjacent pixels; the D loads of a thread are independent (unrolled by 8).
// AGG = false (the production path) has no store in the loop, and the loads of 8 slices are issued before the
// first use: 128 bytes in flight per thread.
// DIRECT: the launch covers the method's whole candidate range, so the winner is final: the disparity map is written
// straight away (0 where every candidate is NaN / inf, as k_keys_to_disp) and no key buffer is touched.
template <bool AGG, bool DIRECT>
__global__ void __launch_bounds__(128)
k_gfs_wta(const float* __restrict__ qv, const float2* __restrict__ aff, int D, int H, int W, int Wq, int d_label0,
          unsigned long long* __restrict__ keys, float* __restrict__ agg, float* __restrict__ disp) {
    const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= W) return;
    const size_t slice = (size_t)Wq;                         // q' is [H][D][Wq]: the D slices of a row are contiguous
    const float* p = qv + (size_t)y * D * Wq + x4;
    float best[4] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
    int bd[4] = {0, 0, 0, 0};
    const size_t n = (size_t)H * W;
    for (int d0 = 0; d0 < D; d0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (d0 + u < D) v[u] = __ldcs((const float4*)(p + (size_t)(d0 + u) * slice));
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int d = d0 + u;
            if (d < D) {
                const float2 sh = __ldg(&aff[d]);
                const float q[4] = {fmaf(v[u].x, sh.x, sh.y), fmaf(v[u].y, sh.x, sh.y), fmaf(v[u].z, sh.x, sh.y), fmaf(v[u].w, sh.x, sh.y)};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (q[k] < best[k]) { best[k] = q[k]; bd[k] = d; }
                    if (AGG && x4 + k < W) agg[(size_t)d * n + (size_t)y * W + x4 + k] = q[k];
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (DIRECT) {
            if (x4 + k < W) disp[(size_t)y * W + x4 + k] = best[k] < __int_as_float(0x7f800000) ? (float)(d_label0 + bd[k]) : 0.0f;
        } else if (x4 + k < W && best[k] < __int_as_float(0x7f800000)) {
            unsigned long long* kp = keys + (size_t)y * W + x4 + k;
            *kp = min(*kp, wta_key(best[k], d_label0 + bd[k]));
        }
    }
}

// smallest image height the streaming kernel takes: >= 2 bands of >= 2K rows each
static inline bool gfs_supported(int H, int W, int win) {
    return (win == 5 || win == 7 || win == 9) && H >= 4 * win && W >= 2;
}

template <int K>
static asw_status gfs_launch(asw_ctx* ctx, const FeatF* fref, const FeatF* ftgt, const float4* Gi, const GfsMoments* Gmom,
                             const int* guide_mm, GfsGeom g, const TadParams& tp, float* qv, uint32_t* slice_mm, float2* aff,
                             int cn, int d_label0, unsigned long long* keys, float* agg, float* disp_direct) {
    constexpr int QW = GfsLayout<K>::QW;
    const size_t smem = GfsLayout<K>::bytes;
    cudaFuncSetAttribute(k_gfs_filter<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int strips = cdiv(g.W, QW), groups = cdiv(cn, GFS_NS);
    // bands: enough CTAs for ~4 resident sets, every band at least 2K rows, at least 2 (one top, one bottom)
    int nb = cdiv(4 * ctx->sm_count, strips * groups);
    if (ctx->tune[ASW_TUNE_GFS_BANDS] > 0) nb = ctx->tune[ASW_TUNE_GFS_BANDS];
    nb = std::max(2, std::min(nb, g.H / (2 * K)));
    g.D = cn; g.nbands = nb;
    TadStream ts;
    ts.thr_c = (int)floorf(tp.thr_c); ts.add_c = (int)rintf(tp.add_c); ts.thr_g = tp.thr_g;
    ts.reg_r = (float)tp.reg_r; ts.reg = (float)tp.reg;
    // dev builds: interior strips (every (a,b) column x0-A .. x0-A+AW-1 inside the image) can take the walker kernel
    int s_lo = 0, s_hi = 0;
#ifdef ASW_DEV_KERNELS
    if (K == 9 && asw_dev("ASW_GFS_WALK")) {
        constexpr int A = K / 2, AW = GfsLayout<K>::AW;
        s_lo = 1;
        s_hi = s_lo;
        while (s_hi < strips && s_hi * QW - A + AW - 1 <= g.W - 1) s_hi++;
    }
#endif
    if (s_hi > s_lo) {
#ifdef ASW_DEV_KERNELS
        const size_t smem_w = GfwLayout<9>::bytes;
        cudaFuncSetAttribute(k_gfs_walk<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
        LAUNCH(ctx, "gfs_filter", (k_gfs_walk<9><<<dim3(s_hi - s_lo, groups, nb), GFW_THREADS, smem_w, ctx->stream>>>(
                                      fref, ftgt, Gi, Gmom, guide_mm, g, ts, tp.c0, qv, slice_mm, s_lo)));
        LAUNCH(ctx, "gfs_filter_edge", (k_gfs_filter<K><<<dim3(s_lo, groups, nb), 2 * GFS_THREADS, smem, ctx->stream>>>(
                                      fref, ftgt, Gi, Gmom, guide_mm, g, ts, tp.c0, qv, slice_mm, 0)));
        if (strips > s_hi)
            LAUNCH(ctx, "gfs_filter_edge", (k_gfs_filter<K><<<dim3(strips - s_hi, groups, nb), 2 * GFS_THREADS, smem, ctx->stream>>>(
                                          fref, ftgt, Gi, Gmom, guide_mm, g, ts, tp.c0, qv, slice_mm, s_hi)));
#endif
    } else {
        LAUNCH(ctx, "gfs_filter", (k_gfs_filter<K><<<dim3(strips, groups, nb), 2 * GFS_THREADS, smem, ctx->stream>>>(
                                      fref, ftgt, Gi, Gmom, guide_mm, g, ts, tp.c0, qv, slice_mm, 0)));
    }
    LAUNCH(ctx, "gfs_affine", (k_gfs_affine<<<cdiv(cn, 128), 128, 0, ctx->stream>>>(slice_mm, cn, tp.c0, aff)));
    const dim3 wgrid(cdiv(cdiv(g.W, 4), 128), g.H);
    if (disp_direct && agg) {
        LAUNCH(ctx, "gfs_wta", (k_gfs_wta<true, true><<<wgrid, 128, 0, ctx->stream>>>(qv, aff, cn, g.H, g.W, g.Wq, d_label0, keys, agg, disp_direct)));
    } else if (disp_direct) {
        LAUNCH(ctx, "gfs_wta", (k_gfs_wta<false, true><<<wgrid, 128, 0, ctx->stream>>>(qv, aff, cn, g.H, g.W, g.Wq, d_label0, keys, agg, disp_direct)));
    } else if (agg) {
        LAUNCH(ctx, "gfs_wta", (k_gfs_wta<true, false><<<wgrid, 128, 0, ctx->stream>>>(qv, aff, cn, g.H, g.W, g.Wq, d_label0, keys, agg, nullptr)));
    } else {
        LAUNCH(ctx, "gfs_wta", (k_gfs_wta<false, false><<<wgrid, 128, 0, ctx->stream>>>(qv, aff, cn, g.H, g.W, g.Wq, d_label0, keys, agg, nullptr)));
    }
    return ASW_OK;
}
