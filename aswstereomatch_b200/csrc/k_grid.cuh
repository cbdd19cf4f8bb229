// k_grid.cuh -- bilateral-grid ASW (createBilGrid nested-map overload A.cpp:1831-2185, quadrlinear_blGrid
// A.cpp:2227-2251, computeAdaptiveWeight_bilateralGrid A.cpp:2253-2430).  LEFT only (the RIGHT branch reads
// at(j, width) out of bounds, A.cpp:1923).
//
// Per candidate disparity (D+1 of them) a dense 4-D grid [x][y][z][w] of (double sum, int count):
//   K11 k_grid_splat   : key = (cvRound(x/sS), cvRound(y/sS), cvRound(L/sR), cvRound(R(max(0,x-d))/sR)),
//                        sum += |L-R| (integer valued -> order independent, exact), count += 1
//   K12 k_grid_pass    : the four IN-PLACE, ascending-index (recursive) 5-tap passes w, z, y, x with the
//                        reference's edge rules; counts truncated to int after every pass; fp64, same
//                        expression order, no FMA contraction (the library is built with -fmad=false)
//   K13 k_grid_slice   : corners at ceil(index) +- 1 (A.cpp:2295-2327), out-of-grid keys read 0, 4-linear
//                        interpolation in the reference's order, cost = quad(sum)/quad(count) (NaN never wins)
// Several candidates are processed per launch (blockIdx.z / grid batches) to fill the 148 SMs.
#pragma once
#include "k_cost.cuh"

struct GridDims {
    int nx, ny, nz, nw;          // last indices (inclusive): cells = (nx+1)(ny+1)(nz+1)(nw+1)
    double rate_s, rate_r;
    size_t cells;
};
__device__ __forceinline__ size_t gidx(const GridDims& g, int x, int y, int z, int w) {
    return (((size_t)x * (g.ny + 1) + y) * (g.nz + 1) + z) * (g.nw + 1) + w;
}
// cvRound = round half to even (lrint in the default rounding mode); cvCeil = ceil
__device__ __forceinline__ int cv_round(double v) { return __double2int_rn(v); }
__device__ __forceinline__ int cv_ceil(double v) { return __double2int_ru(v); }

// one candidate per blockIdx.z
__global__ void k_grid_splat(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rg, int H, int W, int d_first,
                             GridDims g, double* __restrict__ S, int* __restrict__ C) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int d = d_first + blockIdx.z;
    float lf = (float)lg[(size_t)y * W + x];
    float rf = (float)rg[(size_t)y * W + max(0, x - d)];
    int kx = cv_round((double)x / g.rate_s), ky = cv_round((double)y / g.rate_s);
    int kz = cv_round((double)lf / g.rate_r), kw = cv_round((double)rf / g.rate_r);
    size_t id = (size_t)blockIdx.z * g.cells + gidx(g, kx, ky, kz, kw);
    atomicAdd(&S[id], fabs((double)(lf - rf)));        // A.cpp:1897-1912
    atomicAdd(&C[id], 1);
}

// one recursive in-place 5-tap pass along a strided line of n+1 cells, n >= 3 (A.cpp:1936-2183).  The two edge
// rules at either end are peeled, so the interior loop is branch-free; expressions keep the reference's order.
// CT: storage type of the counts (int in shared memory; int or uint16_t in HBM: a count never exceeds the pixels of one
// spatial cell, and 10-byte instead of 12-byte cells are 17 % less traffic for the HBM-bound passes).
template <typename ST, bool BATCH, typename CT>
__device__ __forceinline__ void grid_pass_line(double* __restrict__ s, CT* __restrict__ c, ST stride, int n) {
    // rolling registers: m2, m1 = already-updated v[i-2], v[i-1]; s0 = v[i]; p1, p2 = original v[i+1], v[i+2]
    double sm2, sm1, s0 = s[0], sp1 = s[stride], sp2 = s[2 * stride], sp3 = s[3 * stride];
    double cm2, cm1, c0 = (double)c[0], cp1 = (double)c[stride], cp2 = (double)c[2 * stride], cp3 = (double)c[3 * stride];
    double ns, nc;
    int nci;
#define GRID_STORE_SHIFT(i)                                                   \
    nci = (int)nc; /* pair<double,double> -> pair<double,int>: truncation */ \
    s[(ST)(i) * stride] = ns;                                                 \
    c[(ST)(i) * stride] = (CT)nci;                                            \
    sm2 = sm1; sm1 = ns; s0 = sp1; sp1 = sp2; sp2 = sp3;                      \
    cm2 = cm1; cm1 = (double)nci; c0 = cp1; cp1 = cp2; cp2 = cp3;
    // i = 0
    ns = 0.6 * s0 + 0.3 * sp1 + 0.1 * sp2;
    nc = 0.6 * c0 + 0.3 * cp1 + 0.1 * cp2;
    sm1 = 0; cm1 = 0;
    GRID_STORE_SHIFT(0)
    if (4 <= n) { sp3 = s[(ST)4 * stride]; cp3 = (double)c[(ST)4 * stride]; }
    // i = 1
    ns = 0.2 * sm1 + 0.5 * s0 + 0.2 * sp1 + 0.1 * sp2;
    nc = 0.2 * cm1 + 0.5 * c0 + 0.2 * cp1 + 0.1 * cp2;
    GRID_STORE_SHIFT(1)
    // interior: the originals entering the window during the next 8 steps are loaded together (8 independent loads
    // in flight per thread: the global y / x passes are latency-bound otherwise), before those cells are overwritten
    if (!BATCH) {
        for (int i = 2; i <= n - 2; i++) {
            if (i + 3 <= n) { sp3 = s[(ST)(i + 3) * stride]; cp3 = (double)c[(ST)(i + 3) * stride]; }
            ns = 0.0625 * sm2 + 0.25 * sm1 + 0.375 * s0 + 0.25 * sp1 + 0.0625 * sp2;
            nc = 0.0625 * cm2 + 0.25 * cm1 + 0.375 * c0 + 0.25 * cp1 + 0.0625 * cp2;
            GRID_STORE_SHIFT(i)
        }
    }
    for (int i0 = 2; BATCH && i0 <= n - 2; i0 += 8) {
        double in_s[8];
        CT in_c[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = i0 + 3 + u;
            if (k <= n) { in_s[u] = s[(ST)k * stride]; in_c[u] = c[(ST)k * stride]; }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int i = i0 + u;
            if (i <= n - 2) {
                sp3 = in_s[u]; cp3 = (double)in_c[u];
                ns = 0.0625 * sm2 + 0.25 * sm1 + 0.375 * s0 + 0.25 * sp1 + 0.0625 * sp2;
                nc = 0.0625 * cm2 + 0.25 * cm1 + 0.375 * c0 + 0.25 * cp1 + 0.0625 * cp2;
                GRID_STORE_SHIFT(i)
            }
        }
    }
    // i = n - 1
    ns = 0.1 * sm2 + 0.2 * sm1 + 0.5 * s0 + 0.2 * sp1;
    nc = 0.1 * cm2 + 0.2 * cm1 + 0.5 * c0 + 0.2 * cp1;
    GRID_STORE_SHIFT(n - 1)
    // i = n
    ns = 0.1 * sm2 + 0.3 * sm1 + 0.6 * s0;
    nc = 0.1 * cm2 + 0.3 * cm1 + 0.6 * c0;
    nci = (int)nc;
    s[(ST)n * stride] = ns;
    c[(ST)n * stride] = (CT)nci;
#undef GRID_STORE_SHIFT
}

// axis: 0 = w, 1 = z, 2 = y, 3 = x.  One thread per line; thread index ordered so that adjacent threads
// touch adjacent memory where the axis allows it (y and x passes are fully coalesced).
template <typename CT>
__global__ void k_grid_pass(double* __restrict__ S, CT* __restrict__ C, GridDims g, int axis, int n_grids) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int X = g.nx + 1, Y = g.ny + 1, Z = g.nz + 1, Wd = g.nw + 1;
    size_t lines;
    if (axis == 0) lines = (size_t)X * Y * Z;
    else if (axis == 1) lines = (size_t)X * Y * Wd;
    else if (axis == 2) lines = (size_t)X * Z * Wd;
    else lines = (size_t)Y * Z * Wd;
    if (t >= lines * n_grids) return;
    size_t gi = t / lines, l = t - gi * lines;
    double* s = S + gi * g.cells;
    CT* c = C + gi * g.cells;
    size_t base, stride; int n;
    if (axis == 0) { base = l * Wd; stride = 1; n = g.nw; }                                   // (x,y,z) fixed
    else if (axis == 1) { size_t xy = l / Wd; int w = (int)(l - xy * Wd); base = xy * Z * Wd + w; stride = Wd; n = g.nz; }
    else if (axis == 2) { size_t x = l / ((size_t)Z * Wd); size_t zw = l - x * Z * Wd; base = x * Y * Z * Wd + zw; stride = (size_t)Z * Wd; n = g.ny; }
    else { base = l; stride = (size_t)Y * Z * Wd; n = g.nx; }
    grid_pass_line<size_t, true, CT>(s + base, c + base, stride, n);
}

// splat + w + z fused: the CTA builds the (z, w) planes it owns directly in shared memory from the ~(2 sS)^2
// pixels whose spatial key (cvRound(x/sS), cvRound(y/sS)) is the plane's (A.cpp:1897-1912; |L-R| is an integer, so
// integer shared-memory atomics are exact and order independent), runs both recursive passes there and writes the
// plane once: no zero-fill of the grid, no global atomics, no read-back before the first two passes, and no strided
// per-thread global lines (the stand-alone w pass walked 28-cell lines with a 224-byte stride between threads).
// One thread per line in shared memory, odd pitch against bank conflicts; PL planes per CTA, grid-stride.
// Tables (built on the host with the same double arithmetic, A.cpp:1897-1912 / 2290-2298):
//   klut[v]  = cvRound(v / sR) for the 256 gray values          xq[x] = x / sS, yq[y] = y / sS, gq[v] = v / sR
//   xr[kx]   = first pixel column whose cvRound(x / sS) is kx, xr[X + kx] = how many (keys are monotone, so the
//              pixels of a spatial key are one contiguous run); yr likewise for rows
// so neither kernel executes a double-precision division per pixel.
#define GRID_CONV_MAX 24      // cells of one CTA iteration / 128 threads, upper bound (host-checked)
struct GridTables {
    const int* klut;             // [256]
    const int* xr;               // [2 * X]
    const int* yr;               // [2 * Y]
    const double* xq;            // [W]
    const double* yq;            // [H]
    const double* gq;            // [256]
    int max_run;                 // longest pixel run of a spatial key (both axes)
};

template <typename CT>
__global__ void __launch_bounds__(128)
k_grid_build_wz(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rg, int H, int W, int d_first,
                double* __restrict__ S, CT* __restrict__ C, GridDims g, GridTables tb, int n_planes, int PL) {
    extern __shared__ double sm_grid[];
    const int X = g.nx + 1, Y = g.ny + 1, Z = g.nz + 1, Wd = g.nw + 1, pitch = Wd | 1, cells = Z * Wd, XY = X * Y;
    double* ss = sm_grid;                                   // [PL][Z][pitch]
    int* cc = (int*)(ss + (size_t)PL * Z * pitch);          // [PL][Z][pitch]
    int* si = (int*)ss;                                     // integer sums while splatting, aliased onto the double sums
    __shared__ int klut[256];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += 128) klut[i] = tb.klut[i];
    const int run = tb.max_run, per = run * run;
    for (int p0 = blockIdx.x * PL; p0 < n_planes; p0 += gridDim.x * PL) {
        const int np = min(PL, n_planes - p0);
        for (int i = tid; i < np * Z * pitch; i += 128) { cc[i] = 0; si[i] = 0; }
        __syncthreads();
        // splat: 4 pixels per thread and round, the global loads of a round issued together
        for (int i0 = tid; i0 < np * per; i0 += 4 * 128) {
            int cell[4], lv[4], rv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * 128;
                cell[u] = -1;
                if (i < np * per) {
                    const int pl = i / per, r = i - pl * per, oy = r / run, ox = r - oy * run;
                    const int p = p0 + pl, gi = p / XY, xy = p - gi * XY, kx = xy / Y, ky = xy - kx * Y;
                    if (ox < tb.xr[X + kx] && oy < tb.yr[Y + ky]) {
                        const int x = tb.xr[kx] + ox, y = tb.yr[ky] + oy;
                        lv[u] = lg[(size_t)y * W + x];
                        rv[u] = rg[(size_t)y * W + max(0, x - (d_first + gi))];
                        cell[u] = pl * Z * pitch;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (cell[u] >= 0) {
                    const int a = cell[u] + klut[lv[u]] * pitch + klut[rv[u]];
                    atomicAdd(&si[a], abs(lv[u] - rv[u]));   // A.cpp:1897-1912: |L-R| is an integer
                    atomicAdd(&cc[a], 1);
                }
            }
        }
        __syncthreads();
        {   // int -> double in place (different element sizes): through registers, with a barrier in between
            int tmp[GRID_CONV_MAX];
#pragma unroll
            for (int j = 0; j < GRID_CONV_MAX; j++) { const int i = tid + j * 128; if (i < np * Z * pitch) tmp[j] = si[i]; }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < GRID_CONV_MAX; j++) { const int i = tid + j * 128; if (i < np * Z * pitch) ss[i] = (double)tmp[j]; }
        }
        __syncthreads();
        for (int l = tid; l < np * Z; l += 128)             // w pass: line = (plane, z), unit stride
            grid_pass_line<int, false, int>(ss + l * pitch, cc + l * pitch, 1, g.nw);
        __syncthreads();
        for (int l = tid; l < np * Wd; l += 128) {          // z pass: line = (plane, w), stride = pitch
            const int pl = l / Wd, w = l - pl * Wd;
            grid_pass_line<int, false, int>(ss + pl * Z * pitch + w, cc + pl * Z * pitch + w, pitch, g.nz);
        }
        __syncthreads();
        const size_t base = (size_t)p0 * cells;
        for (int row = tid >> 5; row < np * Z; row += 4)     // one (plane, z) row per warp: no index divisions
            for (int w = tid & 31; w < Wd; w += 32) {
                S[base + row * Wd + w] = ss[row * pitch + w];
                C[base + row * Wd + w] = (CT)cc[row * pitch + w];
            }
        __syncthreads();
    }
}

template <typename CT>
// 64 registers (8 CTAs per SM): the 32 scattered loads per pixel are latency-bound, more resident warps hide more of it
// (84 registers / 5 CTAs: 6.5 ms; 64 / 8: 5.1 ms; 48 / 10: 5.6 ms at config 3a)
__global__ void __launch_bounds__(128, 8)
k_grid_slice(const uint8_t* __restrict__ lg, const uint8_t* __restrict__ rg, int H, int W, int d_first,
                             int cand_first, GridDims g, GridTables tb, const double* __restrict__ S, const CT* __restrict__ C,
                             unsigned long long* __restrict__ keys, float* __restrict__ agg) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    int d = d_first + blockIdx.z;
    const double* s = S + (size_t)blockIdx.z * g.cells;
    const CT* c = C + (size_t)blockIdx.z * g.cells;
    double x_ = __ldg(&tb.xq[x]), y_ = __ldg(&tb.yq[y]);                                      // x / sS, y / sS (A.cpp:2290-2293)
    double cl = __ldg(&tb.gq[lg[(size_t)y * W + x]]);                                         // L / sR
    double cr = __ldg(&tb.gq[rg[(size_t)y * W + max(0, x - d)]]);                             // R(max(0, x-d)) / sR
    int X = cv_ceil(x_), Y = cv_ceil(y_), Z = cv_ceil(cl), Q = cv_ceil(cr);
    double fx = X - x_, fy = Y - y_, fz = Z - cl, fw = Q - cr;
    // all 32 corner loads are issued before the first use (the gather is latency-bound otherwise)
    double vs[16];
    int vc[16];
    // corner k = (X -/+ 1, Y -/+ 1, Z -/+ 1, Q -/+ 1): one in-range flag and one 32-bit offset per axis end
    const int sY = g.nw + 1, sZ = sY * (g.nz + 1), sX = sZ * (g.ny + 1);    // strides of w.. are 1, sY(z), sZ(y), sX(x)
    const bool okx[2] = {X - 1 >= 0 && X - 1 <= g.nx, X + 1 >= 0 && X + 1 <= g.nx};
    const bool oky[2] = {Y - 1 >= 0 && Y - 1 <= g.ny, Y + 1 >= 0 && Y + 1 <= g.ny};
    const bool okz[2] = {Z - 1 >= 0 && Z - 1 <= g.nz, Z + 1 >= 0 && Z + 1 <= g.nz};
    const bool okq[2] = {Q - 1 >= 0 && Q - 1 <= g.nw, Q + 1 >= 0 && Q + 1 <= g.nw};
    const int base = (X - 1) * sX + (Y - 1) * sZ + (Z - 1) * sY + (Q - 1);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const bool in = okx[(k >> 3) & 1] && oky[(k >> 2) & 1] && okz[(k >> 1) & 1] && okq[k & 1];
        const int id = base + ((k & 8) ? 2 * sX : 0) + ((k & 4) ? 2 * sZ : 0) + ((k & 2) ? 2 * sY : 0) + ((k & 1) ? 2 : 0);
        vs[k] = in ? __ldg(&s[id]) : 0.0;                               // map default-insert reads (0.0, 0)
        vc[k] = in ? (int)__ldg(&c[id]) : 0;
    }
    double val[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = t == 0 ? vs[k] : (double)vc[k];
        double a[8], b[4], c2[2];
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = v[2 * k] * (1 - fw) + v[2 * k + 1] * fw;          // A.cpp:2233-2250
#pragma unroll
        for (int k = 0; k < 4; k++) b[k] = a[2 * k] * (1 - fz) + a[2 * k + 1] * fz;
#pragma unroll
        for (int k = 0; k < 2; k++) c2[k] = b[2 * k] * (1 - fy) + b[2 * k + 1] * fy;
        val[t] = c2[0] * (1 - fx) + c2[1] * fx;
    }
    double E = val[0] / val[1];                                                                // A.cpp:2348
    size_t p = (size_t)y * W + x;
    if (agg) agg[(size_t)(cand_first + blockIdx.z) * H * W + p] = (float)E;
    atomicMin(&keys[p], wta_key_d(E, d));
}

// the candidate batches of one call; CT = storage type of the counts in HBM (uint16_t needs the fused build)
template <typename CT>
static asw_status grid_batches(asw_ctx* ctx, const uint8_t* gl, const uint8_t* gr, int H, int W, int min_d, int n_cand, int batch,
                               const GridDims& g, const GridTables& tb, double* S, CT* C, unsigned long long* keys, float* agg_dev) {
    for (int c0 = 0; c0 < n_cand; c0 += batch) {
        int nb = n_cand - c0 < batch ? n_cand - c0 : batch;
        // splat + w + z fused through shared memory when a few (z, w) planes fit; otherwise zero-fill, global splat
        // and one launch per axis
        const int Zd = g.nz + 1, Wdd = g.nw + 1;
        const size_t plane_bytes = (size_t)Zd * (Wdd | 1) * 12;       // double sum (the int splat sum aliases it) + int count
        int first_axis = 0;
        // planes per CTA iteration: one line per thread in the w / z passes, and the in-place int -> double conversion
        // holds a CTA iteration's cells in GRID_CONV_MAX registers per thread
        const int PL_max = std::min(std::min(128 / std::max(Zd, Wdd), (int)((96 * 1024) / plane_bytes)),
                                    (int)((size_t)GRID_CONV_MAX * 128 / ((size_t)Zd * (Wdd | 1))));
        if (PL_max >= 1 && !asw_dev("ASW_GRID_UNFUSED")) {
            int PL = PL_max;
            size_t smem = plane_bytes * PL + 16;
            int n_planes = (g.nx + 1) * (g.ny + 1) * nb;
            cudaFuncSetAttribute(k_grid_build_wz<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            unsigned blocks = (unsigned)std::min<size_t>((size_t)(n_planes + PL - 1) / PL, (size_t)ctx->sm_count * 32);
            LAUNCH(ctx, "grid_build_wz", (k_grid_build_wz<CT><<<blocks, 128, smem, ctx->stream>>>(gl, gr, H, W, min_d + c0, S, C, g, tb, n_planes, PL)));
            first_axis = 2;
        } else {
            ASW_CUDA(ctx, cudaMemsetAsync(S, 0, g.cells * nb * sizeof(double), ctx->stream));      // A.cpp:1874-1892
            ASW_CUDA(ctx, cudaMemsetAsync(C, 0, g.cells * nb * sizeof(CT), ctx->stream));
            LAUNCH(ctx, "grid_splat", (k_grid_splat<<<dim3(cdiv(W, 128), H, nb), 128, 0, ctx->stream>>>(gl, gr, H, W, min_d + c0, g, S, (int*)C)));
        }
        for (int axis = first_axis; axis < 4; axis++) {
            size_t lines = g.cells / (size_t)((axis == 0 ? g.nw : axis == 1 ? g.nz : axis == 2 ? g.ny : g.nx) + 1) * nb;
            static const char* pass_name[4] = {"grid_pass_w", "grid_pass_z", "grid_pass_y", "grid_pass_x"};
            LAUNCH(ctx, pass_name[axis], (k_grid_pass<CT><<<(unsigned)((lines + 127) / 128), 128, 0, ctx->stream>>>(S, C, g, axis, nb)));
        }
        LAUNCH(ctx, "grid_slice", (k_grid_slice<CT><<<dim3(cdiv(W, 128), H, nb), 128, 0, ctx->stream>>>(gl, gr, H, W, min_d + c0, c0, g, tb, S, C, keys, agg_dev)));
    }
    return ASW_OK;
}

static asw_status dev_bilateral_grid(asw_ctx* ctx, const uint8_t* dL, const uint8_t* dR, int H, int W, double rate_s,
                                     double rate_r, int min_d, int num_d, float* disp_dev, float* agg_dev) {
    if (rate_s <= 0) rate_s = 16;                                  // A.cpp:1835-1843
    if (rate_r <= 0) rate_r = 0.07;
    size_t n = (size_t)H * W;
    GridDims g;
    g.rate_s = rate_s; g.rate_r = rate_r;
    g.nz = (int)lrint(255.0 / rate_r); g.nw = g.nz;                // A.cpp:1866-1871 (cvRound)
    g.nx = (int)lrint((W - 1) / rate_s); g.ny = (int)lrint((H - 1) / rate_s);
    if (g.nx < 3 || g.ny < 3 || g.nz < 3 || g.nw < 3)
        return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "bilateral grid needs at least 4 cells per axis%s%s");
    g.cells = (size_t)(g.nx + 1) * (g.ny + 1) * (g.nz + 1) * (g.nw + 1);
    uint8_t *gl, *gr;
    ASW_TRY(ws_get(ctx, WS_GRAY_L, n, &gl));
    ASW_TRY(ws_get(ctx, WS_GRAY_R, n, &gr));
    LAUNCH(ctx, "bgr2gray", (k_bgr2gray_pad<<<dim3(cdiv(W, 256), H), 256, 0, ctx->stream>>>(dL, H, W, 0, 0, gl)));   // A.cpp:2270-2277
    LAUNCH(ctx, "bgr2gray", (k_bgr2gray_pad<<<dim3(cdiv(W, 256), H), 256, 0, ctx->stream>>>(dR, H, W, 0, 0, gr)));
    int n_cand = num_d + 1;                                         // A.cpp:2258, 2279
    // batch candidates so that a launch has enough lines to fill the GPU, within a memory budget
    size_t per = g.cells * 12;
    int batch = (int)(((size_t)1536 << 20) / per);
    if (batch < 1) batch = 1;
    if (batch > n_cand) batch = n_cand;
    if (batch > 32) batch = 32;
    double* S; int* C;
    ASW_TRY(ws_get(ctx, WS_GRID_S, g.cells * batch, &S));
    ASW_TRY(ws_get(ctx, WS_GRID_C, g.cells * batch, &C));
    unsigned long long* keys;
    ASW_TRY(ws_get(ctx, WS_KEYS, n, &keys));
    ASW_TRY(init_keys(ctx, keys, n));
    // host tables (same IEEE double operations as the kernels would execute; lrint = cvRound = half-to-even)
    const int GX = g.nx + 1, GY = g.ny + 1;
    std::vector<int> ti(256 + 2 * GX + 2 * GY, 0);
    std::vector<double> td((size_t)W + H + 256);
    int* h_klut = ti.data(); int* h_xr = h_klut + 256; int* h_yr = h_xr + 2 * GX;
    double* h_xq = td.data(); double* h_yq = h_xq + W; double* h_gq = h_yq + H;
    for (int v = 0; v < 256; v++) { h_gq[v] = (double)v / rate_r; h_klut[v] = (int)lrint(h_gq[v]); }
    int max_run = 1;
    for (int x = 0; x < W; x++) {
        h_xq[x] = (double)x / rate_s;
        int k = (int)lrint(h_xq[x]);
        if (k > g.nx) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "bilateral grid: spatial key outside the grid%s%s");
        if (h_xr[GX + k]++ == 0) h_xr[k] = x;
        max_run = std::max(max_run, h_xr[GX + k]);
    }
    for (int y = 0; y < H; y++) {
        h_yq[y] = (double)y / rate_s;
        int k = (int)lrint(h_yq[y]);
        if (k > g.ny) return asw_fail(ctx, ASW_ERR_UNSUPPORTED, "bilateral grid: spatial key outside the grid%s%s");
        if (h_yr[GY + k]++ == 0) h_yr[k] = y;
        max_run = std::max(max_run, h_yr[GY + k]);
    }
    int* d_ti; double* d_td;
    ASW_TRY(ws_get(ctx, WS_GRID_TI, ti.size(), &d_ti));
    ASW_TRY(ws_get(ctx, WS_GRID_TD, td.size(), &d_td));
    {
        // the tables depend on the geometry only: uploaded (and the stream synchronised, they are host temporaries) on the
        // first call with this geometry; a batch of equal-sized pairs stays asynchronous on the ctx stream
        char key[128];
        snprintf(key, sizeof(key), "grid:%d:%d:%.17g:%.17g", H, W, rate_s, rate_r);
        const bool have_i = table_cached(ctx, 2, WS_GRID_TI, key), have_d = table_cached(ctx, 3, WS_GRID_TD, key);
        if (!have_i || !have_d) {
            ASW_CUDA(ctx, cudaMemcpyAsync(d_ti, ti.data(), ti.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
            ASW_CUDA(ctx, cudaMemcpyAsync(d_td, td.data(), td.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            ASW_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    GridTables tb;
    tb.klut = d_ti; tb.xr = d_ti + 256; tb.yr = tb.xr + 2 * GX;
    tb.xq = d_td; tb.yq = d_td + W; tb.gq = tb.yq + H; tb.max_run = max_run;
    {
        const int Zd = g.nz + 1, Wdd = g.nw + 1;
        const size_t plane_bytes = (size_t)Zd * (Wdd | 1) * 12;
        const int PL_max = std::min(std::min(128 / std::max(Zd, Wdd), (int)((96 * 1024) / plane_bytes)),
                                    (int)((size_t)GRID_CONV_MAX * 128 / ((size_t)Zd * (Wdd | 1))));
        const bool fused = PL_max >= 1 && !asw_dev("ASW_GRID_UNFUSED");
        // 16-bit counts: a count never exceeds the pixels of one spatial cell (every pass is a convex combination)
        if (fused && (size_t)max_run * max_run <= 65535 && !asw_dev("ASW_GRID_C32"))
            ASW_TRY((grid_batches<uint16_t>(ctx, gl, gr, H, W, min_d, n_cand, batch, g, tb, S, (uint16_t*)C, keys, agg_dev)));
        else
            ASW_TRY((grid_batches<int>(ctx, gl, gr, H, W, min_d, n_cand, batch, g, tb, S, C, keys, agg_dev)));
    }
    return keys_to_disp(ctx, keys, n, disp_dev);
}
