/*
 * asw_oracle.h -- CPU oracle for the aswStereoMatch dense-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithms (ZhangYY12345/aswStereoMatch, aswStereoMatch/methods/aswMethods.cpp,
 * cited per function below as A.cpp:<lines>) over OpenCV-4.13 primitive
 * semantics.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (aswstereomatch_b200/) never
 * links, imports or calls anything in this directory.
 *
 * Parity pin: the reference ships no tests / golden vectors.  Every function of
 * this file is pinned bit-exactly against the reference's OWN code: oracle/_ref/
 * libasw_ref.so is /root/reference's aswMethods.cpp compiled UNMODIFIED against an
 * OpenCV stand-in (oracle/refshim/, itself pinned against the real cv2 4.13), see
 * tests/test_cpu_ref.py and the fixture tests/golden/ref_methods_44x60_d6.npz made
 * by that library (tools/make_ref_golden.py).  The OpenCV-primitive stages are in
 * addition pinned against Python cv2 4.13 by oracle/cv2_restatement.py.  Only
 * stage 4 (LR check / fill / refine) has no reference counterpart: it is our own
 * specification (its median selection rule and weights are the reference's).
 *
 * Conventions: images are tightly packed, row-major; colour = BGR interleaved
 * u8 (CV_8UC3).  Volumes are [Deval][H][W] float.  disp_type: 0 = DISPARITY_LEFT,
 * 1 = DISPARITY_RIGHT (P.h:4-8).  All functions return 0 on success, <0 on the
 * argument errors for which the reference returns an empty Mat.
 */
#ifndef ASW_ORACLE_H
#define ASW_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_BAD_ARG (-1)
#define ORC_UNSUPPORTED (-4)

/* ---- mini-cv primitives (OpenCV 4.13 semantics) ---- */
void orc_bgr2gray(const uint8_t* bgr, int npix, uint8_t* gray);
void orc_box_filter_f32(const float* src, int H, int W, int ksize, float* dst);
void orc_normalize_minmax_f32(const float* src, long n, float* dst);
void orc_normalize_minmax_u8(const uint8_t* src, long n, float* dst);
void orc_scharr_x_u8c3(const uint8_t* src, int H, int W, float* dst);

/* ---- stage 1: raw cost volumes ---- */
/* A.cpp:415-487 (LEFT branch, 3 channels).  disp_type 1 = mirrored LEFT formulas
 * with the RIGHT branch's cropping (A.cpp:488-530; the reference's own RIGHT
 * branch throws). */
int orc_cost_tad_cg(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d,
                    int disp_type, double regularity, double thres_c, double thres_g, float* vol);
/* A.cpp:651-668: the above, each slice padded by win/2 with BORDER_REFLECT.
 * vol is [D][H+2h][W+2h]. */
int orc_cost_tad_cg_padded(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d,
                           int disp_type, double regularity, double thres_c, double thres_g,
                           int win, float* vol);
/* A.cpp:2442-2503 called as in A.cpp:2524-2536 (gray AD -> box). L,R are BGR. */
int orc_cost_sad_box(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d,
                     int disp_type, int win, float* vol);

/* ---- stage 3: WTA (A.cpp:3032-3048 and the 7 other inlined copies) ---- */
/* strict <, ascending d, start DBL_MAX, NaN never wins; never-written pixels get 0.0f. */
void orc_wta(const float* vol, int D, int H, int W, int min_d, float* disp);

/* ---- guided filter (A.cpp:2766-2854) ---- */
int orc_guided_filter(const uint8_t* guide, int C, const float* p, int H, int W, int r, double eps,
                      float* q);

/* ---- methods.  agg (optional, may be NULL) receives the aggregated cost volume
 *      [Deval][H][W] as float (the double E of the loop methods rounded to float). ---- */
int orc_asw_traditional(const uint8_t* L, const uint8_t* R, int H, int W, double gamma_c,
                        double gamma_g, int disp_type, int win, int min_d, int num_d,
                        float* disp, float* agg);                       /* A.cpp:1016-1156 */
int orc_asw_direct8(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win,
                    int min_d, int num_d, float* disp, float* agg);     /* A.cpp:1167-1319, LEFT only */
int orc_asw_geodesic(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win,
                     int min_d, int num_d, float* disp, float* agg);    /* A.cpp:1321-1534 */
int orc_geodesic_dist(const uint8_t* img, int H, int W, int win, float* dist /*[H][W][win*win]*/);
int orc_asw_bilateral_grid(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type,
                           double rate_s, double rate_r, int min_d, int num_d,
                           float* disp, float* agg);                    /* A.cpp:1831-2185, 2227-2430 */
int orc_asw_blo1(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double rate_r,
                 int win, int min_d, int num_d, float* disp, float* agg); /* A.cpp:2505-2725 */
int orc_asw_guidedf(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps,
                    int win, int min_d, int num_d, float* disp, float* agg); /* A.cpp:2867-2963 */
int orc_cost_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d, int disp_type, int win,
                 float* vol);                                           /* A.cpp:767-800, 924-1013 */
int orc_asw_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int min_d, int num_d,
                float* disp);                                           /* A.cpp:812-912 (dispatcher's NCC) */
int orc_asw_guidedf3(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps, int win,
                     int min_d, int num_d, float* disp, float* agg);     /* A.cpp:3063-3137 */
int orc_asw_guidedf2(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps,
                     int win, int min_d, int num_d, float* disp, float* agg); /* A.cpp:2976-3050 */
int orc_asw_weighted_median(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type,
                            int win, double rate_s, double rate_r, int min_d, int num_d,
                            float* disp, float* agg);                   /* A.cpp:3228-3383 */

/* ---- stage 4 (NOT in the reference; our specification) ---- */
void orc_lr_check(const float* dl, const float* dr, int H, int W, float tol, uint8_t* valid);
void orc_fill_invalid(const float* d, const uint8_t* valid, int H, int W, float* out);
int orc_wmedian_refine(const uint8_t* img, const float* filled, const uint8_t* valid, int H, int W,
                       int win, double rate_s, double rate_r, float* out);

/* driver post-processing (aswStereoMatch.cpp:97-98): convertTo(CV_8UC1) + normalize(0, 255, NORM_MINMAX) */
void orc_disparity_to_u8(const float* disp, int H, int W, uint8_t* out);

/* dispatcher literals (A.cpp:46-88); algorithm ids as P.h:10-24 */
int orc_stereo_matching(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type,
                        int algorithm, int win, int min_d, int num_d, float* disp);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
