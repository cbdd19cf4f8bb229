/*
 * asw_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see asw_oracle.h).
 *
 * Restates ZhangYY12345/aswStereoMatch aswStereoMatch/methods/aswMethods.cpp
 * ("A.cpp") over OpenCV-4.13 primitive semantics.  Written from the behaviour
 * of the reference, with flat arrays instead of vector<Mat>/std::map.  Build:
 * gcc -O2 -ffp-contract=off (no FMA contraction: the reference is MSVC
 * /fp:precise, and OpenCV's fused operations are restated with explicit fma()).
 */
#include "asw_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ */
/* mini-cv                                                            */
/* ------------------------------------------------------------------ */

/* cv::borderInterpolate for BORDER_REFLECT (delta=0) / BORDER_REFLECT_101 (delta=1) */
static int border_idx(int p, int len, int delta) {
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = len - 1 - (p - len) - delta;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* cvtColor(BGR2GRAY) u8, OpenCV >= 4.2 fixed point (SURVEY B-11) */
void orc_bgr2gray(const uint8_t* bgr, int npix, uint8_t* gray) {
    for (int i = 0; i < npix; i++) {
        int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((3735 * b + 19235 * g + 9798 * r + (1 << 14)) >> 15);
    }
}

/* copyMakeBorder(..., left, right, BORDER_REFLECT) on the column axis only */
static uint8_t* pad_cols_reflect_u8(const uint8_t* src, int H, int W, int cn, int left, int right) {
    int Wp = W + left + right;
    uint8_t* dst = (uint8_t*)malloc((size_t)H * Wp * cn);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < Wp; x++) {
            int sx = border_idx(x - left, W, 0);
            memcpy(dst + ((size_t)y * Wp + x) * cn, src + ((size_t)y * W + sx) * cn, (size_t)cn);
        }
    return dst;
}

/* boxFilter(src 32F, dst 32F, Size(k,k)), normalised, anchor k/2, BORDER_REFLECT_101.
 * OpenCV accumulates 32F input in double: RowSum<float,double> (direct sums for k==3/5,
 * otherwise a running sum s += (new - old)) then ColumnSum<double,float> (running sum
 * SUM += new; out = (float)(SUM*scale); SUM -= old, scale = double 1/(k*k)).  Restated
 * with the same summation order as a single-stripe (single-thread) OpenCV run, which is
 * bit-exact against cv2 4.13 with cv2.setNumThreads(1); multi-threaded OpenCV restarts
 * the column sum per stripe and differs from itself in the last ulp on rare pixels. */
void orc_box_filter_f32(const float* src, int H, int W, int k, float* dst) {
    int a = k / 2;
    double scale = 1.0 / ((double)k * k);
    int Wp = W + k - 1;
    double* rs = (double*)malloc((size_t)H * W * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        const float* row = src + (size_t)y * W;
        double* S = (double*)malloc((size_t)Wp * sizeof(double));
        double* D = rs + (size_t)y * W;
        for (int x = 0; x < Wp; x++) S[x] = (double)row[border_idx(x - a, W, 1)];
        if (k == 3) {
            for (int x = 0; x < W; x++) D[x] = S[x] + S[x + 1] + S[x + 2];
        } else if (k == 5) {
            for (int x = 0; x < W; x++) D[x] = S[x] + S[x + 1] + S[x + 2] + S[x + 3] + S[x + 4];
        } else {
            double s = 0;
            for (int i = 0; i < k; i++) s += S[i];
            D[0] = s;
            for (int x = 0; x < W - 1; x++) { s += S[x + k] - S[x]; D[x + 1] = s; }
        }
        free(S);
    }
#pragma omp parallel for schedule(static)
    for (int x = 0; x < W; x++) {
        double sum = 0;
        for (int i = 0; i < k - 1; i++) sum += rs[(size_t)border_idx(i - a, H, 1) * W + x];
        for (int y = 0; y < H; y++) {
            double s0 = sum + rs[(size_t)border_idx(y - a + k - 1, H, 1) * W + x];
            dst[(size_t)y * W + x] = (float)(s0 * scale);
            sum = s0 - rs[(size_t)border_idx(y - a, H, 1) * W + x];
        }
    }
    free(rs);
}

/* normalize(src, dst, 0, 1, NORM_MINMAX, CV_32F): scale = (float)(1/(max-min)) (0 if
 * max-min <= DBL_EPSILON), shift = 0f - (float)(min*scale), dst = fmaf(x, scale, shift).
 * Verified bit-exact against cv2 4.13 for u8 and f32 inputs. */
static void minmax_scale_shift(double mn, double mx, float* sf, float* hf) {
    double scale = (mx - mn > DBL_EPSILON) ? 1.0 / (mx - mn) : 0.0;
    *sf = (float)scale;
    *hf = 0.0f - (float)(mn * (double)(*sf));
}
void orc_normalize_minmax_f32(const float* src, long n, float* dst) {
    double mn = src[0], mx = src[0];
    for (long i = 1; i < n; i++) {
        if (src[i] < mn) mn = src[i];
        if (src[i] > mx) mx = src[i];
    }
    float sf, hf;
    minmax_scale_shift(mn, mx, &sf, &hf);
    for (long i = 0; i < n; i++) dst[i] = fmaf(src[i], sf, hf);
}
void orc_normalize_minmax_u8(const uint8_t* src, long n, float* dst) {
    int mn = src[0], mx = src[0];
    for (long i = 1; i < n; i++) {
        if (src[i] < mn) mn = src[i];
        if (src[i] > mx) mx = src[i];
    }
    float sf, hf;
    minmax_scale_shift(mn, mx, &sf, &hf);
    for (long i = 0; i < n; i++) dst[i] = fmaf((float)src[i], sf, hf);
}

/* filter2D(src 8UC3, CV_32F, [[-3,0,3],[-10,0,10],[-3,0,3]]): correlation, anchor centre,
 * BORDER_REFLECT_101; exact integers (A.cpp:446-450). */
void orc_scharr_x_u8c3(const uint8_t* src, int H, int W, float* dst) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        int y0 = border_idx(y - 1, H, 1), y2 = border_idx(y + 1, H, 1);
        for (int x = 0; x < W; x++) {
            int x0 = border_idx(x - 1, W, 1), x2 = border_idx(x + 1, W, 1);
            for (int c = 0; c < 3; c++) {
#define PX(yy, xx) ((int)src[((size_t)(yy)*W + (xx)) * 3 + c])
                int v = 3 * (PX(y0, x2) - PX(y0, x0)) + 10 * (PX(y, x2) - PX(y, x0)) +
                        3 * (PX(y2, x2) - PX(y2, x0));
#undef PX
                dst[((size_t)y * W + x) * 3 + c] = (float)v;
            }
        }
    }
}

/* cv::addWeighted on 32F (OpenCV 4.13): double scalars, (float)fma(a, alpha, b*beta)
 * (verified bit-exact against cv2 on random inputs). */
static inline float add_weighted_f32(float a, double alpha, float b, double beta) {
    return (float)fma((double)a, alpha, (double)b * beta);
}

/* ------------------------------------------------------------------ */
/* stage 1: TAD colour + gradient cost volume (A.cpp:415-487)          */
/* ------------------------------------------------------------------ */
int orc_cost_tad_cg(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d,
                    int disp_type, double regularity, double thres_c, double thres_g, float* vol) {
    if (!L || !R || !vol || H <= 0 || W <= 0 || num_d <= 0 || min_d < 0) return ORC_BAD_ARG;
    int max_off = min_d + num_d - 1;                       /* A.cpp:423 */
    double reg_r = 1 - regularity;                         /* A.cpp:435 */
    /* LEFT: pad the right image on the left (A.cpp:442); RIGHT: pad the left image on the
     * right (A.cpp:491). ref = un-padded reference-side image, tgt = padded target. */
    const uint8_t* ref = disp_type == 0 ? L : R;
    uint8_t* tgt = disp_type == 0 ? pad_cols_reflect_u8(R, H, W, 3, max_off, 0)
                                  : pad_cols_reflect_u8(L, H, W, 3, 0, max_off);
    int Wp = W + max_off;
    float* g_ref = (float*)malloc((size_t)H * W * 3 * sizeof(float));
    float* g_tgt = (float*)malloc((size_t)H * Wp * 3 * sizeof(float));
    orc_scharr_x_u8c3(ref, H, W, g_ref);                   /* A.cpp:449 */
    orc_scharr_x_u8c3(tgt, H, Wp, g_tgt);                  /* A.cpp:450 (on the padded image) */
    const double third = 1.0 / 3;                          /* MatExpr "/ 3" -> alpha = 1./3 */
    float tg = (float)thres_g;
    /* scaleAdd(mask, thresC/255, m1) on u8 lowers to addWeighted in float (SURVEY B-2) */
    float add_c = 255.0f * (float)(thres_c / 255.0);
#pragma omp parallel for schedule(static)
    for (int di = 0; di < num_d; di++) {
        int offset = min_d + di;
        int x0 = disp_type == 0 ? max_off - offset : offset;   /* Rect(max_offset-offset,..) / Rect(offset,..) */
        float* out = vol + (size_t)di * H * W;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const uint8_t* a = ref + ((size_t)y * W + x) * 3;
                const uint8_t* b = tgt + ((size_t)y * Wp + x0 + x) * 3;
                int c0 = abs((int)a[0] - b[0]), c1 = abs((int)a[1] - b[1]), c2 = abs((int)a[2] - b[2]);
                int m1 = c0 + c1; if (m1 > 255) m1 = 255;        /* (c0+c1) materialised, saturating */
                int s = m1 + c2;
                int color = s / 3 + (s % 3 == 2);                 /* addWeighted(m1,1/3,c2,1/3) u8 (B-1) */
                int cc = 0;
                if ((double)color > thres_c) {                    /* A.cpp:461-465 */
                    long v = lrintf((float)color + add_c);
                    cc = v > 255 ? 255 : (int)v;
                }
                const float* ga = g_ref + ((size_t)y * W + x) * 3;
                const float* gb = g_tgt + ((size_t)y * Wp + x0 + x) * 3;
                float g0 = fabsf(ga[0] - gb[0]), g1 = fabsf(ga[1] - gb[1]), g2 = fabsf(ga[2] - gb[2]);
                float gm1 = g0 + g1;
                float G = add_weighted_f32(gm1, third, g2, third);   /* A.cpp:473 */
                /* A.cpp:474-482: bit = (G>T)/255 in {0,1}; not(bit) in {255,254};
                 * Gc = G*bit + T*not(bit) via scaleAdd */
                float gc = (G > tg) ? (254.0f * tg + G) : (255.0f * tg + 0.0f);
                out[(size_t)y * W + x] = add_weighted_f32((float)cc, reg_r, gc, regularity); /* A.cpp:484 */
            }
    }
    free(tgt); free(g_ref); free(g_tgt);
    return ORC_OK;
}

/* A.cpp:651-668 */
int orc_cost_tad_cg_padded(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d,
                           int disp_type, double regularity, double thres_c, double thres_g,
                           int win, float* vol) {
    if (win % 2 == 0) return ORC_BAD_ARG;
    int h = win / 2, Hp = H + 2 * h, Wp = W + 2 * h;
    float* raw = (float*)malloc((size_t)num_d * H * W * sizeof(float));
    int rc = orc_cost_tad_cg(L, R, H, W, min_d, num_d, disp_type, regularity, thres_c, thres_g, raw);
    if (rc == ORC_OK) {
        for (int d = 0; d < num_d; d++)
            for (int y = 0; y < Hp; y++) {
                int sy = border_idx(y - h, H, 0);
                for (int x = 0; x < Wp; x++)
                    vol[((size_t)d * Hp + y) * Wp + x] =
                        raw[((size_t)d * H + sy) * W + border_idx(x - h, W, 0)];
            }
    }
    free(raw);
    return rc;
}

/* getCostSAD_d (A.cpp:2442-2503) for every d as called at A.cpp:2524-2536 / 2877-2889 */
int orc_cost_sad_box(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d,
                     int disp_type, int win, float* vol) {
    if (!L || !R || !vol || H <= 0 || W <= 0 || num_d <= 0 || min_d < 0) return ORC_BAD_ARG;
    if (win % 2 == 0) return ORC_BAD_ARG;                  /* A.cpp:2458-2462 */
    int max_off = min_d + num_d - 1;
    if (max_off <= 0) return ORC_BAD_ARG;                  /* A.cpp:2472 (padded cols <= width) */
    uint8_t* lg = (uint8_t*)malloc((size_t)H * W);
    uint8_t* rg = (uint8_t*)malloc((size_t)H * W);
    orc_bgr2gray(L, H * W, lg);
    orc_bgr2gray(R, H * W, rg);
    const uint8_t* ref = disp_type == 0 ? lg : rg;
    uint8_t* tgt = disp_type == 0 ? pad_cols_reflect_u8(rg, H, W, 1, max_off, 0)
                                  : pad_cols_reflect_u8(lg, H, W, 1, 0, max_off);
    int Wp = W + max_off;
#pragma omp parallel for schedule(static)
    for (int di = 0; di < num_d; di++) {
        int d = min_d + di;
        /* LEFT: Rect(cols - W - d, ..) (A.cpp:2477); RIGHT: Rect(d, ..) (A.cpp:2492) */
        int x0 = disp_type == 0 ? Wp - W - d : d;
        float* ad = (float*)malloc((size_t)H * W * sizeof(float));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                ad[(size_t)y * W + x] =
                    (float)abs((int)ref[(size_t)y * W + x] - tgt[(size_t)y * Wp + x0 + x]);
        orc_box_filter_f32(ad, H, W, win, vol + (size_t)di * H * W);
        free(ad);
    }
    free(lg); free(rg); free(tgt);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* stage 3: WTA                                                        */
/* ------------------------------------------------------------------ */
void orc_wta(const float* vol, int D, int H, int W, int min_d, float* disp) {
    size_t n = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        double best = DBL_MAX;
        float dsp = 0.0f;   /* the reference leaves never-written pixels uninitialised; sentinel 0 */
        for (int d = 0; d < D; d++) {
            double c = vol[(size_t)d * n + i];
            if (c < best) { best = c; dsp = (float)(d + min_d); }
        }
        disp[i] = dsp;
    }
}

/* ------------------------------------------------------------------ */
/* guided filter (A.cpp:2766-2854), diagonal covariance, C = 1, 3 or 6 */
/* ------------------------------------------------------------------ */
typedef struct {
    int H, W, C, r;
    float* I;      /* [C][H][W] normalised guidance */
    float* mI;     /* [C][H][W] box(I) */
    float* den;    /* [C][H][W] (corrII - mI*mI) + eps */
} gf_guide_t;

static void gf_guide_init(gf_guide_t* g, const uint8_t* guide, int C, int H, int W, int r, double eps) {
    size_t n = (size_t)H * W;
    g->H = H; g->W = W; g->C = C; g->r = r;
    g->I = (float*)malloc(n * C * sizeof(float));
    g->mI = (float*)malloc(n * C * sizeof(float));
    g->den = (float*)malloc(n * C * sizeof(float));
    float* In = (float*)malloc(n * C * sizeof(float));
    orc_normalize_minmax_u8(guide, (long)n * C, In);         /* A.cpp:2774, one global min/max */
    for (int c = 0; c < C; c++)
        for (size_t i = 0; i < n; i++) g->I[c * n + i] = In[i * C + c];
    free(In);
    float* tmp = (float*)malloc(n * sizeof(float));
    float fe = (float)eps;                                    /* scaleAdd(ones, eps, var) in f32 */
    for (int c = 0; c < C; c++) {
        float* Ic = g->I + c * n;
        orc_box_filter_f32(Ic, H, W, r, g->mI + c * n);      /* A.cpp:2778 */
        for (size_t i = 0; i < n; i++) tmp[i] = Ic[i] * Ic[i];
        orc_box_filter_f32(tmp, H, W, r, g->den + c * n);    /* A.cpp:2796 */
        for (size_t i = 0; i < n; i++) {
            float m = g->mI[c * n + i];
            float var = g->den[c * n + i] - m * m;           /* A.cpp:2799 */
            g->den[c * n + i] = 1.0f * fe + var;             /* A.cpp:2846 denominator */
        }
    }
    free(tmp);
}
static void gf_guide_free(gf_guide_t* g) { free(g->I); free(g->mI); free(g->den); }

/* filter one cost slice with a prepared guide */
static void gf_apply(const gf_guide_t* g, const float* cost, float* q) {
    int H = g->H, W = g->W, C = g->C, r = g->r;
    size_t n = (size_t)H * W;
    float* p = (float*)malloc(n * sizeof(float));
    float* mP = (float*)malloc(n * sizeof(float));
    float* tmp = (float*)malloc(n * sizeof(float));
    float* a = (float*)malloc(n * C * sizeof(float));
    float* b = (float*)malloc(n * sizeof(float));
    orc_normalize_minmax_f32(cost, (long)n, p);               /* A.cpp:2775 */
    orc_box_filter_f32(p, H, W, r, mP);                       /* A.cpp:2780 */
    for (int c = 0; c < C; c++) {
        const float* Ic = g->I + c * n;
        for (size_t i = 0; i < n; i++) tmp[i] = Ic[i] * p[i];
        orc_box_filter_f32(tmp, H, W, r, a + c * n);          /* corrGuidP, A.cpp:2787-2792 */
        for (size_t i = 0; i < n; i++) {
            float cov = a[c * n + i] - g->mI[c * n + i] * mP[i];   /* A.cpp:2805-2815 */
            a[c * n + i] = cov / g->den[c * n + i];                /* A.cpp:2846 */
        }
    }
    for (size_t i = 0; i < n; i++) {                          /* A.cpp:2847, Vec dot left-to-right */
        float dot = a[i] * g->mI[i];
        for (int c = 1; c < C; c++) dot = dot + a[c * n + i] * g->mI[c * n + i];
        b[i] = mP[i] - dot;
    }
    for (int c = 0; c < C; c++) {                             /* A.cpp:2849 */
        orc_box_filter_f32(a + c * n, H, W, r, tmp);
        memcpy(a + c * n, tmp, n * sizeof(float));
    }
    orc_box_filter_f32(b, H, W, r, tmp);                      /* A.cpp:2850 */
    for (size_t i = 0; i < n; i++) {                          /* A.cpp:2852 */
        float dot = a[i] * g->I[i];
        for (int c = 1; c < C; c++) dot = dot + a[c * n + i] * g->I[c * n + i];
        q[i] = dot + tmp[i];
    }
    free(p); free(mP); free(tmp); free(a); free(b);
}

int orc_guided_filter(const uint8_t* guide, int C, const float* p, int H, int W, int r, double eps,
                      float* q) {
    if (!guide || !p || !q || (C != 1 && C != 3 && C != 6) || r <= 0) return ORC_BAD_ARG;
    gf_guide_t g;
    gf_guide_init(&g, guide, C, H, W, r, eps);
    gf_apply(&g, p, q);
    gf_guide_free(&g);
    return ORC_OK;
}

/* computeAdaptiveWeight_GuidedF_2 (A.cpp:2976-3050).  disp_type 1: guidance = right image
 * (A.cpp:3019) over the mirrored cost (the reference's own RIGHT cost throws). */
int orc_asw_guidedf2(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps,
                     int win, int min_d, int num_d, float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d <= 0 || win <= 0) return ORC_BAD_ARG;
    size_t n = (size_t)H * W;
    float* cost = (float*)malloc(n * num_d * sizeof(float));
    int rc = orc_cost_tad_cg(L, R, H, W, min_d, num_d, disp_type, 0.4, 10, 50, cost); /* A.cpp:2990 */
    if (rc != ORC_OK) { free(cost); return rc; }
    float* q = agg ? agg : (float*)malloc(n * num_d * sizeof(float));
    gf_guide_t g;
    gf_guide_init(&g, disp_type == 0 ? L : R, 3, H, W, win, eps);
#pragma omp parallel for schedule(dynamic)
    for (int d = 0; d < num_d; d++) gf_apply(&g, cost + d * n, q + d * n);   /* A.cpp:3004 */
    gf_guide_free(&g);
    orc_wta(q, num_d, H, W, min_d, disp);                                     /* A.cpp:3032-3048 */
    if (!agg) free(q);
    free(cost);
    return ORC_OK;
}

/* computeAdaptiveWeight_GuidedF (A.cpp:2867-2963): SAD-box cost, 6-channel guidance L (+) R_d */
int orc_asw_guidedf(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps,
                    int win, int min_d, int num_d, float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d <= 0 || win <= 0) return ORC_BAD_ARG;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d - 1;
    float* cost = (float*)malloc(n * num_d * sizeof(float));
    int rc = orc_cost_sad_box(L, R, H, W, min_d, num_d, disp_type, win, cost);  /* A.cpp:2882-2900 */
    if (rc != ORC_OK) { free(cost); return rc; }
    float* q = agg ? agg : (float*)malloc(n * num_d * sizeof(float));
    uint8_t* lb = pad_cols_reflect_u8(L, H, W, 3, 0, max_off);   /* A.cpp:2877 */
    uint8_t* rb = pad_cols_reflect_u8(R, H, W, 3, max_off, 0);   /* A.cpp:2878 */
    int Wp = W + max_off;
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < num_d; i++) {
        uint8_t* guide = (uint8_t*)malloc(n * 6);
        /* LEFT: Rect(numDisparity - i - 1, ..) of the padded right image (A.cpp:2909);
         * RIGHT: Rect(i + minDisparity, ..) of the padded left image (A.cpp:2926) */
        int x0 = disp_type == 0 ? num_d - i - 1 : i + min_d;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const uint8_t* pl = disp_type == 0 ? L + ((size_t)y * W + x) * 3
                                                   : lb + ((size_t)y * Wp + x0 + x) * 3;
                const uint8_t* pr = disp_type == 0 ? rb + ((size_t)y * Wp + x0 + x) * 3
                                                   : R + ((size_t)y * W + x) * 3;
                uint8_t* g6 = guide + ((size_t)y * W + x) * 6;
                g6[0] = pl[0]; g6[1] = pl[1]; g6[2] = pl[2];
                g6[3] = pr[0]; g6[4] = pr[1]; g6[5] = pr[2];
            }
        orc_guided_filter(guide, 6, cost + i * n, H, W, win, eps, q + i * n);   /* A.cpp:2915 */
        free(guide);
    }
    free(lb); free(rb);
    orc_wta(q, num_d, H, W, min_d, disp);                                       /* A.cpp:2945-2961 */
    if (!agg) free(q);
    free(cost);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* NCC cost (getInputImgNCC A.cpp:767-800, computeNCC A.cpp:812-1013) and GuidedF_3 (A.cpp:3063-3137).
 * Quirks reproduced: computeNCC converts with COLOR_RGB2GRAY although the data is BGR (the R and B weights are
 * swapped); the window mean comes from boxFilter (BORDER_REFLECT_101) while the window pixels come from a
 * BORDER_REFLECT-padded copy; the "correlation" is sum(l r) / (sum(l l) * sum(r r)) -- no square root; the products
 * are rounded to float (Mat::mul on CV_32F) and summed in double in window row-major order (cv::sum); the target's
 * windows and means are taken from the PADDED target image (padding of the padding at its outer columns).
 * The Mat-returning overload (the dispatcher's NCC) scans offsets min .. max-1 (strict <, A.cpp:852), keeps the
 * MINIMUM raw double cost for LEFT and never writes anything for RIGHT (its > test starts from DBL_MAX).       */
/* ------------------------------------------------------------------ */
static void rgb2gray_on_bgr(const uint8_t* bgr, size_t npix, uint8_t* gray) {    /* COLOR_RGB2GRAY applied to BGR bytes */
    for (size_t i = 0; i < npix; i++) {
        int c0 = bgr[3 * i], c1 = bgr[3 * i + 1], c2 = bgr[3 * i + 2];
        gray[i] = (uint8_t)((9798 * c0 + 19235 * c1 + 3735 * c2 + (1 << 14)) >> 15);
    }
}
/* per pixel of a gray image: mean = boxFilter(u8 -> 32F) and s2 = sum over the window of fl((p - mean)^2), double */
static void ncc_stats(const uint8_t* g, int H, int W, int win, float* mean, double* s2) {
    int h = win / 2;
    double scale = 1.0 / ((double)win * win);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            long sum = 0;
            for (int j = -h; j <= h; j++)
                for (int i = -h; i <= h; i++) sum += g[(size_t)border_idx(y + j, H, 1) * W + border_idx(x + i, W, 1)];
            float m = (float)((double)sum * scale);
            double acc = 0;
            for (int j = -h; j <= h; j++)
                for (int i = -h; i <= h; i++) {
                    float v = (float)g[(size_t)border_idx(y + j, H, 0) * W + border_idx(x + i, W, 0)] + (-m);
                    acc += (double)(v * v);
                }
            mean[(size_t)y * W + x] = m; s2[(size_t)y * W + x] = acc;
        }
}
/* raw cost of one (pixel, offset): reference-side window at (y, x) of image a (width Wa), target-side window at (y, xt) of
 * image b (width Wb) */
static double ncc_raw(const uint8_t* a, int Wa, const float* ma, const double* s2a, const uint8_t* b, int Wb,
                      const float* mb, const double* s2b, int H, int win, int y, int x, int xt) {
    int h = win / 2;
    float m0 = ma[(size_t)y * Wa + x], m1 = mb[(size_t)y * Wb + xt];
    double sxy = 0;
    for (int j = -h; j <= h; j++) {
        int sy = border_idx(y + j, H, 0);
        for (int i = -h; i <= h; i++) {
            float u = (float)a[(size_t)sy * Wa + border_idx(x + i, Wa, 0)] + (-m0);
            float v = (float)b[(size_t)sy * Wb + border_idx(xt + i, Wb, 0)] + (-m1);
            sxy += (double)(u * v);
        }
    }
    return sxy / (s2a[(size_t)y * Wa + x] * s2b[(size_t)y * Wb + xt]);
}
/* shared set-up: gray images, the padded target, the statistics of both.  LEFT: reference = left, target = right padded on
 * the left, window column x + max_off - offset; RIGHT: reference = right, target = left padded on the right, x + offset */
typedef struct { uint8_t *ref, *tgt; float *mr, *mt; double *sr, *st; int Wt; } NccSetup;
static void ncc_setup(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int max_off, NccSetup* s) {
    size_t n = (size_t)H * W;
    uint8_t* lg = (uint8_t*)malloc(n); uint8_t* rg = (uint8_t*)malloc(n);
    rgb2gray_on_bgr(L, n, lg); rgb2gray_on_bgr(R, n, rg);                  /* A.cpp:829-836 */
    s->Wt = W + max_off;
    if (disp_type == 0) { s->ref = lg; s->tgt = pad_cols_reflect_u8(rg, H, W, 1, max_off, 0); free(rg); }
    else { s->ref = rg; s->tgt = pad_cols_reflect_u8(lg, H, W, 1, 0, max_off); free(lg); }
    s->mr = (float*)malloc(n * sizeof(float)); s->sr = (double*)malloc(n * sizeof(double));
    s->mt = (float*)malloc((size_t)H * s->Wt * sizeof(float)); s->st = (double*)malloc((size_t)H * s->Wt * sizeof(double));
    ncc_stats(s->ref, H, W, win, s->mr, s->sr);
    ncc_stats(s->tgt, H, s->Wt, win, s->mt, s->st);
}
static void ncc_free(NccSetup* s) { free(s->ref); free(s->tgt); free(s->mr); free(s->mt); free(s->sr); free(s->st); }

/* computeNCC, vector overload (A.cpp:924-1013): [num_d][H][W], every slice min-max normalised to [0, 1] */
int orc_cost_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d, int disp_type, int win, float* vol) {
    if (!L || !R || !vol || H <= 0 || W <= 0 || num_d <= 0 || min_d < 0 || win <= 0 || (win & 1) == 0) return ORC_BAD_ARG;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d - 1;
    NccSetup s;
    ncc_setup(L, R, H, W, disp_type, win, max_off, &s);
#pragma omp parallel for schedule(dynamic)
    for (int di = 0; di < num_d; di++) {
        int offset = min_d + di;
        float* raw = (float*)malloc(n * sizeof(float));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int xt = disp_type == 0 ? x + max_off - offset : x + offset;
                raw[(size_t)y * W + x] = (float)ncc_raw(s.ref, W, s.mr, s.sr, s.tgt, s.Wt, s.mt, s.st, H, win, y, x, xt);
            }
        orc_normalize_minmax_f32(raw, (long)n, vol + (size_t)di * n);        /* A.cpp:974-976 */
        free(raw);
    }
    ncc_free(&s);
    return ORC_OK;
}
/* computeNCC, Mat overload (A.cpp:812-912): the dispatcher's NCC */
int orc_asw_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int min_d, int num_d, float* disp) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d <= 0 || min_d < 0 || win <= 0 || (win & 1) == 0) return ORC_BAD_ARG;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d - 1;
    for (size_t i = 0; i < n; i++) disp[i] = 0.0f;
    if (disp_type != 0) return ORC_OK;             /* RIGHT: cost > DBL_MAX never holds (A.cpp:892): nothing is written */
    NccSetup s;
    ncc_setup(L, R, H, W, 0, win, max_off, &s);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double best = DBL_MAX;
            for (int offset = min_d; offset < max_off; offset++) {        /* strict <: the last candidate is never scanned */
                double c = ncc_raw(s.ref, W, s.mr, s.sr, s.tgt, s.Wt, s.mt, s.st, H, win, y, x, x + max_off - offset);
                if (c < best) { best = c; disp[(size_t)y * W + x] = (float)offset; }
            }
        }
    ncc_free(&s);
    return ORC_OK;
}
/* computeAdaptiveWeight_GuidedF_3 (A.cpp:3063-3137): NCC cost, 6-channel guide for LEFT; the RIGHT branch builds the same
 * merge but hands the plain right image to the filter (A.cpp:3104) */
int orc_asw_guidedf3(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps, int win, int min_d,
                     int num_d, float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d <= 0 || win <= 0) return ORC_BAD_ARG;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d - 1;
    float* cost = (float*)malloc(n * num_d * sizeof(float));
    int rc = orc_cost_ncc(L, R, H, W, min_d, num_d, disp_type, win, cost);
    if (rc != ORC_OK) { free(cost); return rc; }
    float* q = agg ? agg : (float*)malloc(n * num_d * sizeof(float));
    uint8_t* rb = pad_cols_reflect_u8(R, H, W, 3, max_off, 0);
    int Wp = W + max_off;
#pragma omp parallel for schedule(dynamic)
    for (int i = 0; i < num_d; i++) {
        if (disp_type == 0) {
            uint8_t* guide = (uint8_t*)malloc(n * 6);
            int x0 = num_d - i - 1;                                         /* A.cpp:3084 */
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    const uint8_t* pl = L + ((size_t)y * W + x) * 3;
                    const uint8_t* pr = rb + ((size_t)y * Wp + x0 + x) * 3;
                    uint8_t* g6 = guide + ((size_t)y * W + x) * 6;
                    g6[0] = pl[0]; g6[1] = pl[1]; g6[2] = pl[2]; g6[3] = pr[0]; g6[4] = pr[1]; g6[5] = pr[2];
                }
            orc_guided_filter(guide, 6, cost + i * n, H, W, win, eps, q + i * n);
            free(guide);
        } else {
            orc_guided_filter(R, 3, cost + i * n, H, W, win, eps, q + i * n);   /* A.cpp:3104 */
        }
    }
    free(rb);
    orc_wta(q, num_d, H, W, min_d, disp);
    if (!agg) free(q);
    free(cost);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* traditional (Yoon-Kweon) ASW (A.cpp:1016-1156)                      */
/* ------------------------------------------------------------------ */
int orc_asw_traditional(const uint8_t* L, const uint8_t* R, int H, int W, double gamma_c,
                        double gamma_g, int disp_type, int win, int min_d, int num_d,
                        float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d < 0 || win <= 0 || (win & 1) == 0)
        return ORC_BAD_ARG;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d;                               /* A.cpp:1021: D+1 candidates */
    int h = win / 2, nw = win * win - 1, cidx = win * win / 2;
    double k = 3;
    uint8_t* lg = (uint8_t*)malloc(n);
    uint8_t* rg = (uint8_t*)malloc(n);
    orc_bgr2gray(L, (int)n, lg);                               /* A.cpp:1030-1033 */
    orc_bgr2gray(R, (int)n, rg);
    float* wl = (float*)malloc(n * nw * sizeof(float));        /* weightAllDirectLeft [nw][H][W] */
    float* wr = (float*)malloc(n * nw * sizeof(float));
    if (!wl || !wr) { free(lg); free(rg); free(wl); free(wr); return ORC_BAD_ARG; }
#pragma omp parallel for schedule(static)
    for (int pn = 0; pn < nw; pn++) {                          /* A.cpp:1044-1072 */
        int pw = pn < cidx ? pn : pn + 1;                      /* centre skipped at build */
        int j = pw / win - h, i = pw % win - h;
        double delta_g = sqrt((double)(i * i + j * j));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int nx = clampi(x + i, 0, W - 1), ny = clampi(y + j, 0, H - 1);
                double dc1 = fabs((double)(lg[(size_t)ny * W + nx] - lg[(size_t)y * W + x]));
                double dc2 = fabs((double)(rg[(size_t)ny * W + nx] - rg[(size_t)y * W + x]));
                wl[(size_t)pn * n + (size_t)y * W + x] = (float)(k * exp(-(dc1 / gamma_c + delta_g / gamma_g)));
                wr[(size_t)pn * n + (size_t)y * W + x] = (float)(k * exp(-(dc2 / gamma_c + delta_g / gamma_g)));
            }
    }
    double* best = (double*)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) { best[i] = DBL_MAX; disp[i] = 0.0f; }
    for (int offset = min_d; offset <= max_off; offset++) {    /* A.cpp:1074 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                double num = 0, den = 0;
                for (int i = 0; i < nw; i++) {                 /* A.cpp:1088-1109 */
                    int kx, ky;
                    if (i > cidx) { kx = (i + 1) / win; ky = (i + 1) % win; }
                    else { kx = i / win; ky = i % win; }
                    int nx = clampi(x - h + kx, 0, W - 1), ny = clampi(y - h + ky, 0, H - 1);
                    float w; double ad;
                    if (disp_type == 0) {
                        int xr = x - offset > 0 ? x - offset : 0;
                        int nxr = nx - offset > 0 ? nx - offset : 0;
                        w = wl[(size_t)i * n + (size_t)y * W + x] * wr[(size_t)i * n + (size_t)y * W + xr];
                        ad = fabs((double)(lg[(size_t)ny * W + nx] - rg[(size_t)ny * W + nxr]));
                    } else {                                   /* A.cpp:1113-1138 */
                        int xl = x + offset < W - 1 ? x + offset : W - 1;
                        int nxl = nx + offset < W - 1 ? nx + offset : W - 1;
                        w = wl[(size_t)i * n + (size_t)y * W + xl] * wr[(size_t)i * n + (size_t)y * W + x];
                        ad = fabs((double)(rg[(size_t)ny * W + nx] - lg[(size_t)ny * W + nxl]));
                    }
                    num += (double)w * ad;
                    den += (double)w;
                }
                double E = num / den;
                size_t p = (size_t)y * W + x;
                if (agg) agg[(size_t)(offset - min_d) * n + p] = (float)E;
                if (E < best[p]) { best[p] = E; disp[p] = (float)offset; }   /* A.cpp:1144-1150 */
            }
    }
    free(best); free(wl); free(wr); free(lg); free(rg);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* 8-direction ASW, computeAdaptiveWeight_direct8 (A.cpp:1167-1319), LEFT only: the RIGHT branch indexes the
 * weight lists with the loop variable i in [-h, h] instead of count (A.cpp:1291), i.e. out of bounds.
 * Taps: window offsets (j, i), j = row, i = column, with i == j || i == 0 || j == 0 || i + j == win - 1
 * (A.cpp:1199, 1247) -- the last test was meant as the anti-diagonal but only ever matches (h, h), which i == j
 * already covers: 3 (win - 1) taps (main diagonal, centre row, centre column), not 4 (win - 1), so the
 * count < 4 (win - 1) guard (A.cpp:1243) never fires.  No transposition here: weight and sample use the same
 * (j, i).  gamma_c = 30, gamma_g = win * 2 / 3 in INTEGER arithmetic (A.cpp:1175), k = 3, D + 1 candidates. */
/* ------------------------------------------------------------------ */
int orc_asw_direct8(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int min_d,
                    int num_d, float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d < 0 || win <= 0 || (win & 1) == 0) return ORC_BAD_ARG;
    if (disp_type != 0) return ORC_UNSUPPORTED;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d, h = win / 2;
    double k = 3, gamma_c = 30, gamma_g = (double)(win * 2 / 3);
    if (gamma_g == 0) return ORC_BAD_ARG;                      /* win = 1: division by zero in the reference */
    uint8_t* lg = (uint8_t*)malloc(n);
    uint8_t* rg = (uint8_t*)malloc(n);
    orc_bgr2gray(L, (int)n, lg);
    orc_bgr2gray(R, (int)n, rg);
    int* tj = (int*)malloc(sizeof(int) * win * win);
    int* ti = (int*)malloc(sizeof(int) * win * win);
    int nw = 0;
    for (int j = -h; j <= h; j++)                              /* A.cpp:1191-1199 */
        for (int i = -h; i <= h; i++) {
            if (i == 0 && j == 0) continue;
            if (i == j || i == 0 || j == 0 || (i + j) == win - 1) { tj[nw] = j; ti[nw] = i; nw++; }
        }
    float* wl = (float*)malloc(n * nw * sizeof(float));
    float* wr = (float*)malloc(n * nw * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nw; t++) {                             /* A.cpp:1203-1218 */
        int j = tj[t], i = ti[t];
        double delta_g = sqrt((double)(i * i + j * j));
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int nx = clampi(x + i, 0, W - 1), ny = clampi(y + j, 0, H - 1);
                double dc1 = fabs((double)(lg[(size_t)ny * W + nx] - lg[(size_t)y * W + x]));
                double dc2 = fabs((double)(rg[(size_t)ny * W + nx] - rg[(size_t)y * W + x]));
                wl[(size_t)t * n + (size_t)y * W + x] = (float)(k * exp(-(dc1 / gamma_c + delta_g / gamma_g)));
                wr[(size_t)t * n + (size_t)y * W + x] = (float)(k * exp(-(dc2 / gamma_c + delta_g / gamma_g)));
            }
    }
    double* best = (double*)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) { best[i] = DBL_MAX; disp[i] = 0.0f; }
    for (int offset = min_d; offset <= max_off; offset++) {    /* A.cpp:1225 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                double num = 0, den = 0;
                int xr = x - offset > 0 ? x - offset : 0;
                for (int t = 0; t < nw; t++) {                 /* A.cpp:1239-1262 */
                    int nx = clampi(x + ti[t], 0, W - 1), ny = clampi(y + tj[t], 0, H - 1);
                    int nxr = nx - offset > 0 ? nx - offset : 0;
                    float w = wl[(size_t)t * n + (size_t)y * W + x] * wr[(size_t)t * n + (size_t)y * W + xr];
                    num += (double)w * fabs((double)(lg[(size_t)ny * W + nx] - rg[(size_t)ny * W + nxr]));
                    den += (double)w;
                }
                double E = num / den;
                size_t p = (size_t)y * W + x;
                if (agg) agg[(size_t)(offset - min_d) * n + p] = (float)E;
                if (E < best[p]) { best[p] = E; disp[p] = (float)offset; }
            }
    }
    free(best); free(wl); free(wr); free(tj); free(ti); free(lg); free(rg);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* geodesic ASW (A.cpp:1321-1534)                                      */
/* ------------------------------------------------------------------ */
static inline float color_dist(const uint8_t* a, const uint8_t* b) {       /* A.cpp:1321-1326 */
    return (float)(fabs((double)(a[0] - b[0])) + fabs((double)(a[1] - b[1])) + fabs((double)(a[2] - b[2])));
}

/* getWinGeoDist (A.cpp:1328-1390): img = (win+2)^2 BGR window, d = (win+2)^2 floats */
static void win_geo_dist(const uint8_t* img, int stride_px, float* d, int win, int iter_time) {
    int S = win + 2;
#define IM(r, c) (img + ((size_t)(r)*stride_px + (c)) * 3)
#define DD(r, c) d[(r)*S + (c)]
    for (int it = 0; it < iter_time; it++) {
        if (it / 2 == 1) {
            for (int r = 1; r <= win; r++)
                for (int c = 1; c <= win; c++) {
                    float v;
                    v = DD(r, c - 1) + color_dist(IM(r, c - 1), IM(r, c));         DD(r, c) = fminf(DD(r, c), v);
                    v = DD(r - 1, c - 1) + color_dist(IM(r - 1, c - 1), IM(r, c)); DD(r, c) = fminf(DD(r, c), v);
                    v = DD(r - 1, c) + color_dist(IM(r - 1, c), IM(r, c));         DD(r, c) = fminf(DD(r, c), v);
                    v = DD(r - 1, c + 1) + color_dist(IM(r - 1, c + 1), IM(r, c)); DD(r, c) = fminf(DD(r, c), v);
                }
        } else if (it / 2 == 0) {
            for (int r = win; r > 0; r--)
                for (int c = win; c > 0; c--) {
                    float v;
                    v = DD(r, c + 1) + color_dist(IM(r, c + 1), IM(r, c));         DD(r, c) = fminf(DD(r, c), v);
                    v = DD(r + 1, c + 1) + color_dist(IM(r + 1, c + 1), IM(r, c)); DD(r, c) = fminf(DD(r, c), v);
                    v = DD(r + 1, c) + color_dist(IM(r + 1, c), IM(r, c));         DD(r, c) = fminf(DD(r, c), v);
                    v = DD(r + 1, c - 1) + color_dist(IM(r + 1, c - 1), IM(r, c)); DD(r, c) = fminf(DD(r, c), v);
                }
        }
    }
#undef IM
#undef DD
}

/* getGeodesicDist (A.cpp:1392-1424): dist[(y*W+x)*win*win + j*win + i] */
int orc_geodesic_dist(const uint8_t* img, int H, int W, int win, float* dist) {
    if (win % 2 == 0) return ORC_BAD_ARG;
    int h = win / 2, S = win + 2, pad = h + 1;
    int Hp = H + 2 * pad, Wp = W + 2 * pad;
    uint8_t* ext = (uint8_t*)malloc((size_t)Hp * Wp * 3);      /* A.cpp:1404 BORDER_REFLECT */
    for (int y = 0; y < Hp; y++) {
        int sy = border_idx(y - pad, H, 0);
        for (int x = 0; x < Wp; x++) {
            int sx = border_idx(x - pad, W, 0);
            memcpy(ext + ((size_t)y * Wp + x) * 3, img + ((size_t)sy * W + sx) * 3, 3);
        }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        float* d = (float*)malloc((size_t)S * S * sizeof(float));
        for (int x = 0; x < W; x++) {
            for (int i = 0; i < S * S; i++) d[i] = FLT_MAX;    /* A.cpp:1416 */
            d[(h + 1) * S + (h + 1)] = 0;                      /* A.cpp:1417 */
            win_geo_dist(ext + ((size_t)y * Wp + x) * 3, Wp, d, win, 3);
            float* o = dist + ((size_t)y * W + x) * win * win;
            for (int j = 0; j < win; j++)
                for (int i = 0; i < win; i++) o[j * win + i] = d[(j + 1) * S + (i + 1)];
        }
        free(d);
    }
    free(ext);
    return ORC_OK;
}

int orc_asw_geodesic(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win,
                     int min_d, int num_d, float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d < 0) return ORC_BAD_ARG;
    if (win % 2 == 0) return ORC_BAD_ARG;                      /* A.cpp:1440-1443 */
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d, h = win / 2, ww = win * win;  /* A.cpp:1447 */
    float* dl = (float*)malloc(n * ww * sizeof(float));
    float* dr = (float*)malloc(n * ww * sizeof(float));
    if (!dl || !dr) { free(dl); free(dr); return ORC_BAD_ARG; }
    orc_geodesic_dist(L, H, W, win, dl);                       /* A.cpp:1464-1465 */
    orc_geodesic_dist(R, H, W, win, dr);
    double* best = (double*)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) { best[i] = DBL_MAX; disp[i] = 0.0f; }
    for (int offset = min_d; offset <= max_off; offset++) {
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                double num = 0, den = 0;
                const float *pl, *pr;
                if (disp_type == 0) {
                    int xr = x - offset > 0 ? x - offset : 0;
                    pl = dl + ((size_t)y * W + x) * ww;
                    pr = dr + ((size_t)y * W + xr) * ww;
                } else {
                    int xl = x + offset < W - 1 ? x + offset : W - 1;
                    pl = dl + ((size_t)y * W + xl) * ww;
                    pr = dr + ((size_t)y * W + x) * ww;
                }
                for (int j = 0; j < win; j++)
                    for (int i = 0; i < win; i++) {            /* A.cpp:1481-1496 */
                        int nx = clampi(x - h + i, 0, W - 1), ny = clampi(y - h + j, 0, H - 1);
                        float cd;
                        if (disp_type == 0) {
                            int nxr = nx - offset > 0 ? nx - offset : 0;
                            cd = color_dist(L + ((size_t)ny * W + nx) * 3, R + ((size_t)ny * W + nxr) * 3);
                        } else {
                            int nxl = nx + offset < W - 1 ? nx + offset : W - 1;
                            cd = color_dist(R + ((size_t)ny * W + nx) * 3, L + ((size_t)ny * W + nxl) * 3);
                        }
                        float t = pl[j * win + i] * pr[j * win + i];
                        num += (double)(t * cd);
                        den += (double)t;
                    }
                double E = num / den;
                size_t p = (size_t)y * W + x;
                if (agg) agg[(size_t)(offset - min_d) * n + p] = (float)E;
                if (E < best[p]) { best[p] = E; disp[p] = (float)offset; }
            }
    }
    free(best); free(dl); free(dr);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* bilateral grid ASW (A.cpp:1831-2185, 2227-2251, 2253-2430)          */
/* ------------------------------------------------------------------ */
static int cv_round(double v) { return (int)lrint(v); }       /* round-half-even (default FE mode) */
static int cv_ceil(double v) { return (int)ceil(v); }

typedef struct { int nx, ny, nz, nw; double* s; int* c; } bgrid_t;   /* dims are last indices (inclusive) */
#define GIDX(g, x, y, z, w) ((((size_t)(x) * ((g)->ny + 1) + (y)) * ((g)->nz + 1) + (z)) * ((g)->nw + 1) + (w))

/* one recursive in-place 5-tap pass along a strided line of n+1 cells (A.cpp:1936-2183) */
static void grid_pass_line(double* s, int* c, size_t stride, int n) {
    for (int i = 0; i <= n; i++) {
#define SV(k) s[(size_t)(k)*stride]
#define CV(k) ((double)c[(size_t)(k)*stride])
        double ns, nc;
        if (i == 0) {
            ns = 0.6 * SV(i) + 0.3 * SV(i + 1) + 0.1 * SV(i + 2);
            nc = 0.6 * CV(i) + 0.3 * CV(i + 1) + 0.1 * CV(i + 2);
        } else if (i == 1) {
            ns = 0.2 * SV(i - 1) + 0.5 * SV(i) + 0.2 * SV(i + 1) + 0.1 * SV(i + 2);
            nc = 0.2 * CV(i - 1) + 0.5 * CV(i) + 0.2 * CV(i + 1) + 0.1 * CV(i + 2);
        } else if (i == n - 1) {
            ns = 0.1 * SV(i - 2) + 0.2 * SV(i - 1) + 0.5 * SV(i) + 0.2 * SV(i + 1);
            nc = 0.1 * CV(i - 2) + 0.2 * CV(i - 1) + 0.5 * CV(i) + 0.2 * CV(i + 1);
        } else if (i == n) {
            ns = 0.1 * SV(i - 2) + 0.3 * SV(i - 1) + 0.6 * SV(i);
            nc = 0.1 * CV(i - 2) + 0.3 * CV(i - 1) + 0.6 * CV(i);
        } else {
            ns = 0.0625 * SV(i - 2) + 0.25 * SV(i - 1) + 0.375 * SV(i) + 0.25 * SV(i + 1) + 0.0625 * SV(i + 2);
            nc = 0.0625 * CV(i - 2) + 0.25 * CV(i - 1) + 0.375 * CV(i) + 0.25 * CV(i + 1) + 0.0625 * CV(i + 2);
        }
        s[(size_t)i * stride] = ns;
        c[(size_t)i * stride] = (int)nc;       /* pair<double,double> -> pair<double,int>: truncation */
#undef SV
#undef CV
    }
}

/* one candidate disparity: build grid, smooth, slice -> E[H*W] (double) */
static void grid_candidate(const uint8_t* lg, const uint8_t* rg, int H, int W, int offset,
                           double rate_s, double rate_r, const bgrid_t* dims, double* E) {
    bgrid_t g = *dims;
    size_t cells = (size_t)(g.nx + 1) * (g.ny + 1) * (g.nz + 1) * (g.nw + 1);
    g.s = (double*)calloc(cells, sizeof(double));       /* A.cpp:1874-1892 */
    g.c = (int*)calloc(cells, sizeof(int));
    for (int i = 0; i < W; i++)                         /* splat, A.cpp:1897-1912 (x outer, y inner) */
        for (int j = 0; j < H; j++) {
            float lf = (float)lg[(size_t)j * W + i];
            float rf = (float)rg[(size_t)j * W + (i - offset > 0 ? i - offset : 0)];
            int kx = cv_round(i / rate_s), ky = cv_round(j / rate_s);
            int kz = cv_round(lf / rate_r), kw = cv_round(rf / rate_r);
            size_t id = GIDX(&g, kx, ky, kz, kw);
            g.s[id] = g.s[id] + fabs((double)(lf - rf));
            g.c[id] = g.c[id] + 1;
        }
    /* w pass (A.cpp:1936-1995) */
    for (int x = 0; x <= g.nx; x++) for (int y = 0; y <= g.ny; y++) for (int z = 0; z <= g.nz; z++) {
        size_t b = GIDX(&g, x, y, z, 0);
        grid_pass_line(g.s + b, g.c + b, 1, g.nw);
    }
    /* z pass (A.cpp:1998-2057) */
    for (int x = 0; x <= g.nx; x++) for (int y = 0; y <= g.ny; y++) for (int w = 0; w <= g.nw; w++) {
        size_t b = GIDX(&g, x, y, 0, w);
        grid_pass_line(g.s + b, g.c + b, (size_t)(g.nw + 1), g.nz);
    }
    /* y pass (A.cpp:2061-2120) */
    for (int w = 0; w <= g.nw; w++) for (int z = 0; z <= g.nz; z++) for (int x = 0; x <= g.nx; x++) {
        size_t b = GIDX(&g, x, 0, z, w);
        grid_pass_line(g.s + b, g.c + b, (size_t)(g.nz + 1) * (g.nw + 1), g.ny);
    }
    /* x pass (A.cpp:2123-2183) */
    for (int y = 0; y <= g.ny; y++) for (int z = 0; z <= g.nz; z++) for (int w = 0; w <= g.nw; w++) {
        size_t b = GIDX(&g, 0, y, z, w);
        grid_pass_line(g.s + b, g.c + b, (size_t)(g.ny + 1) * (g.nz + 1) * (g.nw + 1), g.nx);
    }
    /* slice (A.cpp:2290-2348) */
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            double x_ = x / rate_s, y_ = y / rate_s;
            double cl = lg[(size_t)y * W + x] / rate_r;
            double cr = rg[(size_t)y * W + (x - offset > 0 ? x - offset : 0)] / rate_r;
            int X = cv_ceil(x_), Y = cv_ceil(y_), Z = cv_ceil(cl), Q = cv_ceil(cr);
            double fx = X - x_, fy = Y - y_, fz = Z - cl, fw = Q - cr;
            double ns[16], nc[16];
            for (int k = 0; k < 16; k++) {
                int gx = X + ((k & 8) ? 1 : -1), gy = Y + ((k & 4) ? 1 : -1);
                int gz = Z + ((k & 2) ? 1 : -1), gq = Q + ((k & 1) ? 1 : -1);
                if (gx < 0 || gx > g.nx || gy < 0 || gy > g.ny || gz < 0 || gz > g.nz || gq < 0 || gq > g.nw) {
                    ns[k] = 0; nc[k] = 0;          /* std::map default-insert reads (0.0, 0) */
                } else {
                    size_t id = GIDX(&g, gx, gy, gz, gq);
                    ns[k] = g.s[id]; nc[k] = (double)g.c[id];
                }
            }
            double val[2];
            for (int t = 0; t < 2; t++) {          /* quadrlinear_blGrid, A.cpp:2227-2251 */
                const double* v = t == 0 ? ns : nc;
                double a[8], b[4], c2[2];
                for (int k = 0; k < 8; k++) a[k] = v[2 * k] * (1 - fw) + v[2 * k + 1] * fw;
                for (int k = 0; k < 4; k++) b[k] = a[2 * k] * (1 - fz) + a[2 * k + 1] * fz;
                for (int k = 0; k < 2; k++) c2[k] = b[2 * k] * (1 - fy) + b[2 * k + 1] * fy;
                val[t] = c2[0] * (1 - fx) + c2[1] * fx;
            }
            E[(size_t)y * W + x] = val[0] / val[1];    /* may be 0/0 = NaN: never wins the WTA */
        }
    free(g.s); free(g.c);
}

int orc_asw_bilateral_grid(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type,
                           double rate_s, double rate_r, int min_d, int num_d,
                           float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d < 0) return ORC_BAD_ARG;
    if (disp_type != 0) return ORC_UNSUPPORTED;   /* RIGHT reads at(j,width) out of bounds (A.cpp:1923) */
    if (rate_s <= 0) rate_s = 16;                  /* A.cpp:1835-1843 */
    if (rate_r <= 0) rate_r = 0.07;
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d;                   /* A.cpp:2258: D+1 candidates */
    bgrid_t g0; g0.s = 0; g0.c = 0;
    g0.nz = cv_round(255.0 / rate_r);              /* A.cpp:1866-1871 */
    g0.nw = cv_round(255.0 / rate_r);
    g0.nx = cv_round((W - 1) / rate_s);
    g0.ny = cv_round((H - 1) / rate_s);
    /* with fewer than 4 cells on an axis the edge rules index outside the grid */
    if (g0.nx < 3 || g0.ny < 3 || g0.nz < 3 || g0.nw < 3) return ORC_UNSUPPORTED;
    uint8_t* lg = (uint8_t*)malloc(n);
    uint8_t* rg = (uint8_t*)malloc(n);
    orc_bgr2gray(L, (int)n, lg);                   /* A.cpp:2270-2277 */
    orc_bgr2gray(R, (int)n, rg);
    double* best = (double*)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) { best[i] = DBL_MAX; disp[i] = 0.0f; }
    int ncand = max_off - min_d + 1;
    int chunk = orc_num_threads();
    double* E = (double*)malloc(n * (size_t)chunk * sizeof(double));
    for (int c0 = 0; c0 < ncand; c0 += chunk) {
        int c1 = c0 + chunk < ncand ? c0 + chunk : ncand;
#pragma omp parallel for schedule(dynamic)
        for (int ci = c0; ci < c1; ci++)
            grid_candidate(lg, rg, H, W, min_d + ci, rate_s, rate_r, &g0, E + (size_t)(ci - c0) * n);
        for (int ci = c0; ci < c1; ci++) {         /* ordered WTA, A.cpp:2349-2354 */
            const double* e = E + (size_t)(ci - c0) * n;
            for (size_t p = 0; p < n; p++) {
                if (agg) agg[(size_t)ci * n + p] = (float)e[p];
                if (e[p] < best[p]) { best[p] = e[p]; disp[p] = (float)(min_d + ci); }
            }
        }
    }
    free(E); free(best); free(lg); free(rg);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* BLO(1) ASW (A.cpp:2505-2725)                                        */
/* ------------------------------------------------------------------ */
int orc_asw_blo1(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double rate_r,
                 int win, int min_d, int num_d, float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d <= 0 || win <= 0) return ORC_BAD_ARG;
    if (disp_type != 0 && disp_type != 1) return ORC_BAD_ARG;
    if (min_d != 0) return ORC_UNSUPPORTED;        /* plane index uses 'offset' not offset-min (A.cpp:2666) */
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d - 1;
    int step = (int)(256 * rate_r);                /* A.cpp:2549 */
    if (step <= 0) return ORC_BAD_ARG;             /* the reference would loop forever */
    float* cost = (float*)malloc(n * num_d * sizeof(float));
    int rc = orc_cost_sad_box(L, R, H, W, min_d, num_d, disp_type, win, cost);   /* A.cpp:2531-2546 */
    if (rc != ORC_OK) { free(cost); return rc; }
    uint8_t* lg = (uint8_t*)malloc(n);
    uint8_t* rg = (uint8_t*)malloc(n);
    orc_bgr2gray(L, (int)n, lg);
    orc_bgr2gray(R, (int)n, rg);
    /* LEFT: reference side = left gray, target = right gray padded on the left, crop at max_off - d (A.cpp:2525, 2578);
     * RIGHT: reference side = right gray, target = left gray padded on the right, crop at d (A.cpp:2524, 2612) */
    uint8_t* rb = disp_type == 0 ? pad_cols_reflect_u8(rg, H, W, 1, max_off, 0) : pad_cols_reflect_u8(lg, H, W, 1, 0, max_off);
    if (disp_type == 1) { uint8_t* t = lg; lg = rg; rg = t; }           /* lg = the reference-side image from here on */
    int Wp = W + max_off;
#define BLO1_X0(d) (disp_type == 0 ? max_off - (d) : (d))
    int levels[257], nl = 0, is_level[256];
    memset(is_level, 0, sizeof(is_level));
    for (int i = 0; i < 256; i += step) levels[nl++] = i;                /* A.cpp:2550-2555 */
    if (levels[nl - 1] != 255) levels[nl++] = 255;                       /* A.cpp:2556-2559 */
    for (int i = 0; i < nl; i++) is_level[levels[i]] = 1;
    float* q = agg ? agg : (float*)malloc(n * num_d * sizeof(float));
    /* per pixel the two partial products of A.cpp:2666-2667 */
    float* part_lo = (float*)calloc(n * num_d, sizeof(float));
    float* part_hi = (float*)calloc(n * num_d, sizeof(float));
#pragma omp parallel for schedule(dynamic)
    for (int li = 0; li < nl; li++) {
        int k = levels[li];
        float* ml = (float*)malloc(n * sizeof(float));
        float* m = (float*)malloc(n * sizeof(float));
        float* j = (float*)malloc(n * sizeof(float));
        float* nk = (float*)malloc(n * sizeof(float));
        float* jb = (float*)malloc(n * sizeof(float));
        for (size_t i = 0; i < n; i++) ml[i] = (float)abs((int)lg[i] - k);          /* A.cpp:2571-2572 */
        /* normaliser from the LAST disparity only (A.cpp:2588) */
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float mr = (float)abs((int)rb[(size_t)y * Wp + BLO1_X0(num_d - 1) + x] - k);
                m[(size_t)y * W + x] = mr * ml[(size_t)y * W + x];
            }
        orc_box_filter_f32(m, H, W, win, nk);
        for (int d = 0; d < num_d; d++) {
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    float mr = (float)abs((int)rb[(size_t)y * Wp + BLO1_X0(d) + x] - k);     /* A.cpp:2578 / 2612 */
                    float mm = mr * ml[(size_t)y * W + x];                                   /* A.cpp:2580 */
                    m[(size_t)y * W + x] = mm * cost[(size_t)d * n + (size_t)y * W + x];     /* A.cpp:2583 */
                }
            orc_box_filter_f32(m, H, W, win, j);                                     /* A.cpp:2584 */
            for (size_t i = 0; i < n; i++) jb[i] = j[i] / nk[i];                    /* A.cpp:2594 */
            for (size_t i = 0; i < n; i++) {                                         /* A.cpp:2653-2674 */
                int I = lg[i];
                if (is_level[I]) {
                    if (I == k) q[(size_t)d * n + i] = jb[i];
                } else {
                    int lo = I / step * step, hi = lo + step;
                    if (hi > 255) hi = 255;
                    if (lo == k) part_lo[(size_t)d * n + i] = (float)(I - lo) * jb[i];
                    if (hi == k) part_hi[(size_t)d * n + i] = (float)(hi - I) * jb[i];
                }
            }
        }
        free(ml); free(m); free(j); free(nk); free(jb);
    }
    for (int d = 0; d < num_d; d++)
        for (size_t i = 0; i < n; i++)
            if (!is_level[lg[i]]) q[(size_t)d * n + i] = part_lo[(size_t)d * n + i] + part_hi[(size_t)d * n + i];
#undef BLO1_X0
    orc_wta(q, num_d, H, W, min_d, disp);                                            /* A.cpp:2675-2680, 2711-2716 */
    if (!agg) free(q);
    free(part_lo); free(part_hi); free(cost); free(lg); free(rg); free(rb);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* weighted-median machinery (A.cpp:3139-3383)                         */
/* ------------------------------------------------------------------ */
/* computeSpaceWeightGau (A.cpp:3207-3226): exp(dist2 * (-1/rateS)) with the f32 convertTo */
static void space_weight_gau(float* k, int win, double rate_s) {
    int h = win / 2;
    float alpha = (float)((1.0 / rate_s) * (-1));
    for (int y = 0; y < win; y++) {
        float yd = (float)((y - h) * (y - h));
        for (int x = 0; x < win; x++) {
            float v = (float)((x - h) * (x - h)) + yd;
            k[x * win + y] = (float)exp((double)(v * alpha));    /* dstKernel.at<float>(x, y) */
        }
    }
}
/* one window of computeColorWeightGau (A.cpp:3139-3205) around (y,x) of a REFLECT-padded image */
static inline float color_weight(const uint8_t* c, const uint8_t* q, double alpha) {
    float d0 = (float)abs((int)q[0] - c[0]), d1 = (float)abs((int)q[1] - c[1]), d2 = (float)abs((int)q[2] - c[2]);
    float m1 = d0 + d1;
    float arg = add_weighted_f32(m1, alpha, d2, alpha);          /* (d0+d1+d2)/rateR*(-1) */
    return (float)exp((double)arg);
}

/* weighted median selection rule of A.cpp:3276-3304: stable ascending sort by value,
 * accumulate weights in double, at the first element whose partial sum exceeds total/2
 * return the PREVIOUS element's value (the first one's if it is the first). */
typedef struct { float v, w; int idx; } vw_t;
static int vw_cmp(const void* a, const void* b) {
    const vw_t* x = (const vw_t*)a; const vw_t* y = (const vw_t*)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return x->idx - y->idx;
}
static float weighted_median_select(vw_t* e, int cnt, int guard_first, float keep) {
    double total = 0;
    for (int i = 0; i < cnt; i++) total += (double)e[i].w;      /* cv::sum, double accumulate */
    double half = total / 2;
    qsort(e, (size_t)cnt, sizeof(vw_t), vw_cmp);
    double partial = 0;
    for (int i = 0; i < cnt; i++) {
        partial += (double)e[i].w;
        if (partial > half) {
            if (i == 0) return guard_first ? e[0].v : keep;
            return e[i - 1].v;
        }
    }
    return keep;    /* never crossed (all-zero / NaN weights): the reference leaves the pixel unwritten */
}

int orc_asw_weighted_median(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type,
                            int win, double rate_s, double rate_r, int min_d, int num_d,
                            float* disp, float* agg) {
    if (!L || !R || !disp || H <= 0 || W <= 0 || num_d <= 0) return ORC_BAD_ARG;
    if (win % 2 == 0) return ORC_BAD_ARG;                        /* A.cpp:3238-3241 */
    if (disp_type != 0) return ORC_UNSUPPORTED;                  /* cost throws / UB (Appendix A-3, A-10) */
    size_t n = (size_t)H * W;
    int max_off = min_d + num_d - 1, h = win / 2, ww = win * win;
    int Hc = H + 2 * h, Wc = W + 2 * h;
    float* cost = (float*)malloc((size_t)num_d * Hc * Wc * sizeof(float));
    int rc = orc_cost_tad_cg_padded(L, R, H, W, min_d, num_d, 0, 0.4, 10, 50, win, cost);  /* A.cpp:3250 */
    if (rc != ORC_OK) { free(cost); return rc; }
    float* wdist = (float*)malloc((size_t)ww * sizeof(float));
    space_weight_gau(wdist, win, rate_s);                        /* A.cpp:3255 */
    uint8_t* rb = pad_cols_reflect_u8(R, H, W, 3, max_off, 0);   /* A.cpp:3246 */
    int Wr = W + max_off;
    /* src_border of computeColorWeightGau: REFLECT by h on all sides (A.cpp:3156) */
    int Hl = H + 2 * h, Wl = W + 2 * h, Wrb = Wr + 2 * h;
    uint8_t* lbb = (uint8_t*)malloc((size_t)Hl * Wl * 3);
    uint8_t* rbb = (uint8_t*)malloc((size_t)Hl * Wrb * 3);
    for (int y = 0; y < Hl; y++) {
        int sy = border_idx(y - h, H, 0);
        for (int x = 0; x < Wl; x++)
            memcpy(lbb + ((size_t)y * Wl + x) * 3, L + ((size_t)sy * W + border_idx(x - h, W, 0)) * 3, 3);
        for (int x = 0; x < Wrb; x++)
            memcpy(rbb + ((size_t)y * Wrb + x) * 3, rb + ((size_t)sy * Wr + border_idx(x - h, Wr, 0)) * 3, 3);
    }
    double alpha = (1.0 / rate_r) * (-1);
    float* q = agg ? agg : (float*)malloc(n * num_d * sizeof(float));
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < H; y++) {
        vw_t* e = (vw_t*)malloc((size_t)ww * sizeof(vw_t));
        float* wl = (float*)malloc((size_t)ww * sizeof(float));
        for (int x = 0; x < W; x++) {
            const uint8_t* cl = lbb + ((size_t)(y + h) * Wl + (x + h)) * 3;
            for (int wy = 0; wy < win; wy++)
                for (int wx = 0; wx < win; wx++)
                    wl[wy * win + wx] = color_weight(cl, lbb + ((size_t)(y + wy) * Wl + (x + wx)) * 3, alpha) *
                                        wdist[wy * win + wx];   /* weightWinsL.mul(weightDist) */
            for (int off = 0; off < num_d; off++) {             /* A.cpp:3264-3310 */
                int xr = x - off + num_d - 1;                   /* weightWinsR[y][x - offset + numDisparity - 1] */
                const uint8_t* cr = rbb + ((size_t)(y + h) * Wrb + (xr + h)) * 3;
                for (int wy = 0; wy < win; wy++)
                    for (int wx = 0; wx < win; wx++) {
                        int t = wy * win + wx;
                        e[t].v = cost[((size_t)off * Hc + (y + wy)) * Wc + (x + wx)];
                        e[t].w = wl[t] * color_weight(cr, rbb + ((size_t)(y + wy) * Wrb + (xr + wx)) * 3, alpha);
                        e[t].idx = t;
                    }
                q[(size_t)off * n + (size_t)y * W + x] = weighted_median_select(e, ww, 1, 0.0f);
            }
        }
        free(e); free(wl);
    }
    orc_wta(q, num_d, H, W, min_d, disp);                        /* A.cpp:3365-3381 */
    if (!agg) free(q);
    free(cost); free(wdist); free(rb); free(lbb); free(rbb);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* stage 4: LR check + fill + weighted-median refine (OUR SPEC, a-14)  */
/* ------------------------------------------------------------------ */
void orc_lr_check(const float* dl, const float* dr, int H, int W, float tol, uint8_t* valid) {
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float d = dl[(size_t)y * W + x];
            /* (int)d with saturation (NaN -> 0), target column clamped to the row */
            long long di = d != d ? 0 : d >= 2147483648.0f ? 2147483647LL : d <= -2147483648.0f ? -2147483648LL : (long long)(int)d;
            long long t = (long long)x - di;
            int xr = t < 0 ? 0 : t > W - 1 ? W - 1 : (int)t;
            valid[(size_t)y * W + x] = fabsf(d - dr[(size_t)y * W + xr]) <= tol ? 1 : 0;
        }
}
void orc_fill_invalid(const float* d, const uint8_t* valid, int H, int W, float* out) {
    for (int y = 0; y < H; y++) {
        const float* row = d + (size_t)y * W; const uint8_t* v = valid + (size_t)y * W;
        float* o = out + (size_t)y * W;
        for (int x = 0; x < W; x++) {
            if (v[x]) { o[x] = row[x]; continue; }
            int xl = x - 1; while (xl >= 0 && !v[xl]) xl--;
            int xr = x + 1; while (xr < W && !v[xr]) xr++;
            if (xl >= 0 && xr < W) o[x] = row[xl] < row[xr] ? row[xl] : row[xr];
            else if (xl >= 0) o[x] = row[xl];
            else if (xr < W) o[x] = row[xr];
            else o[x] = row[x];                     /* no valid pixel on the row: keep */
        }
    }
}
int orc_wmedian_refine(const uint8_t* img, const float* filled, const uint8_t* valid, int H, int W,
                       int win, double rate_s, double rate_r, float* out) {
    if (win % 2 == 0 || win <= 0) return ORC_BAD_ARG;
    int h = win / 2, ww = win * win;
    float* wdist = (float*)malloc((size_t)ww * sizeof(float));
    space_weight_gau(wdist, win, rate_s);
    double alpha = (1.0 / rate_r) * (-1);
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < H; y++) {
        vw_t* e = (vw_t*)malloc((size_t)ww * sizeof(vw_t));
        for (int x = 0; x < W; x++) {
            size_t p = (size_t)y * W + x;
            if (valid[p]) { out[p] = filled[p]; continue; }
            const uint8_t* c = img + p * 3;
            for (int wy = 0; wy < win; wy++)
                for (int wx = 0; wx < win; wx++) {
                    int sy = border_idx(y - h + wy, H, 0), sx = border_idx(x - h + wx, W, 0);  /* REFLECT, A.cpp:3156 */
                    int t = wy * win + wx;
                    e[t].v = filled[(size_t)sy * W + sx];
                    e[t].w = color_weight(c, img + ((size_t)sy * W + sx) * 3, alpha) * wdist[t];
                    e[t].idx = t;
                }
            out[p] = weighted_median_select(e, ww, 1, filled[p]);
        }
        free(e);
    }
    free(wdist);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* dispatcher (A.cpp:46-88)                                            */
/* ------------------------------------------------------------------ */
/* driver post-processing (aswStereoMatch.cpp:97-98): disparityMap.convertTo(CV_8UC1) then
 * normalize(0, 255, NORM_MINMAX) on the u8 map: scale = 255 / (max - min) (0 if max == min), shift = -min * scale in
 * double, applied in float as saturate_cast<uchar>(fma(v, (float)scale, (float)shift)) (pinned against cv2 4.13) */
/* ------------------------------------------------------------------ */
void orc_disparity_to_u8(const float* disp, int H, int W, uint8_t* out) {
    size_t n = (size_t)H * W;
    int mn = 255, mx = 0;
    for (size_t i = 0; i < n; i++) {
        long v = lrintf(disp[i]);                                  /* convertTo(CV_8UC1): cvRound + saturate */
        out[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        if (out[i] < mn) mn = out[i];
        if (out[i] > mx) mx = out[i];
    }
    double scale = 255.0 * ((double)(mx - mn) > DBL_EPSILON ? 1.0 / (double)(mx - mn) : 0.0);
    double shift = 0.0 - (double)mn * scale;
    float sf = (float)scale, hf = (float)shift;
    for (size_t i = 0; i < n; i++) {
        long v = lrintf(fmaf((float)out[i], sf, hf));
        out[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

/* ------------------------------------------------------------------ */
int orc_stereo_matching(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type,
                        int algorithm, int win, int min_d, int num_d, float* disp) {
    switch (algorithm) {
    case 2:  return orc_asw_traditional(L, R, H, W, 30, 20, disp_type, win, min_d, num_d, disp, 0);
    case 3:  return orc_asw_direct8(L, R, H, W, disp_type, win, min_d, num_d, disp, 0);           /* A.cpp:61 */
    case 4:  return orc_asw_geodesic(L, R, H, W, disp_type, win, min_d, num_d, disp, 0);
    case 5:  return orc_asw_bilateral_grid(L, R, H, W, disp_type, 10, 10, min_d, num_d, disp, 0);
    case 6:  return orc_asw_blo1(L, R, H, W, disp_type, 0.015, win, min_d, num_d, disp, 0);
    case 7:  return orc_asw_guidedf(L, R, H, W, disp_type, 1e-6, win, min_d, num_d, disp, 0);
    case 8:  return orc_asw_guidedf2(L, R, H, W, disp_type, 1e-6, win, min_d, num_d, disp, 0);
    case 9:  return orc_asw_guidedf3(L, R, H, W, disp_type, 1e-6, win, min_d, num_d, disp, 0);     /* A.cpp:79 */
    case 10: return orc_asw_weighted_median(L, R, H, W, disp_type, win, 10, 10, min_d, num_d, disp, 0);
    case 11: return orc_asw_ncc(L, R, H, W, disp_type, win, min_d, num_d, disp);                  /* A.cpp:85 */
    default: return ORC_UNSUPPORTED;   /* BM, SGBM: third-party algorithms, out of scope */
    }
}
