/*
 * ref_api.cpp -- extern "C" face of oracle/_ref/libasw_ref.so: the reference's OWN functions
 * (aswStereoMatch/methods/aswMethods.h, compiled unmodified from /root/reference against the OpenCV
 * stand-in in oracle/refshim/) over plain pointers, so that tests can compare the C restatement
 * (oracle/asw_oracle.c) and the CUDA path with the reference's own code.  TEST INFRASTRUCTURE ONLY.
 *
 * Images: tightly packed, BGR interleaved u8.  Return value: 0 = ok, -1 = the reference returned an
 * empty Mat (its argument-error convention), -2 = the reference threw (cv::Exception), as the real
 * library would; the message is kept for ref_last_error().
 */
#include "aswMethods.h"

#include <string>

static std::string g_err;

static cv::Mat wrap_u8(const uint8_t* p, int H, int W, int cn) {
    cv::Mat m(H, W, CV_MAKETYPE(CV_8U, cn));
    for (int y = 0; y < H; y++) memcpy(m.ptr(y), p + (size_t)y * W * cn, (size_t)W * cn);
    return m;
}
static cv::Mat wrap_f32(const float* p, int H, int W) {
    cv::Mat m(H, W, CV_32FC1);
    for (int y = 0; y < H; y++) memcpy(m.ptr(y), p + (size_t)y * W, (size_t)W * 4);
    return m;
}
static int put_f32(const cv::Mat& m, int H, int W, float* out) {
    if (m.empty()) return -1;
    if (m.rows != H || m.cols != W || m.type() != CV_32FC1) { g_err = "unexpected result geometry / type"; return -3; }
    for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, m.ptr(y), (size_t)W * 4);
    return 0;
}
#define GUARD(...)                                                              \
    try { __VA_ARGS__ } catch (const cv::Exception& e) { g_err = e.what(); return -2; } \
    catch (const std::exception& e) { g_err = e.what(); return -4; }

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

/* stereoMatching (A.h:91-92): the dispatcher with its own literals */
int ref_stereo_matching(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int algorithm, int win,
                        int min_d, int num_d, float* disp) {
    GUARD(cv::Mat out;
          stereoMatching(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), out, (DisparityType)disp_type,
                         (StereoMatchingAlgorithms)algorithm, win, min_d, num_d);
          return put_f32(out, H, W, disp);)
}
int ref_adaptive_weight(const uint8_t* L, const uint8_t* R, int H, int W, double gamma_c, double gamma_g, int disp_type,
                        int win, int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), gamma_c, gamma_g,
                                               (DisparityType)disp_type, win, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_direct8(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int min_d,
                                int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_direct8(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type,
                                                       win, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_geodesic(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int min_d,
                                 int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_geodesic(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type,
                                                        win, min_d, num_d), H, W, disp);)
}
/* getGeodesicDist (A.h:140): out [H][W][win*win] */
int ref_geodesic_dist(const uint8_t* img, int H, int W, int win, float* out) {
    GUARD(std::map<cv::Point, cv::Mat, MY_COMP_Point2i> m;
          getGeodesicDist(wrap_u8(img, H, W, 3), m, win, 3);
          if (m.empty()) return -1;
          for (int y = 0; y < H; y++)
              for (int x = 0; x < W; x++) {
                  const cv::Mat& w = m[cv::Point(x, y)];
                  for (int j = 0; j < win; j++)
                      for (int i = 0; i < win; i++) out[(((size_t)y * W + x) * win + j) * win + i] = w.at<float>(j, i);
              }
          return 0;)
}
int ref_adaptive_weight_bilateral_grid(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double rate_s,
                                       double rate_r, int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_bilateralGrid(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type,
                                                             rate_s, rate_r, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_blo1(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double rate_r, int win,
                             int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_BLO1(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type, rate_r,
                                                    win, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_guidedf(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps, int win,
                                int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_GuidedF(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type, eps,
                                                       win, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_guidedf_2(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps, int win,
                                  int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_GuidedF_2(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type, eps,
                                                         win, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_weighted_median(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win,
                                        double rate_s, double rate_r, int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_WeightedMedian(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type,
                                                              win, rate_s, rate_r, min_d, num_d), H, W, disp);)
}
int ref_adaptive_weight_guidedf_3(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, double eps, int win,
                                  int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeAdaptiveWeight_GuidedF_3(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type, eps,
                                                         win, min_d, num_d), H, W, disp);)
}
/* computeNCC, Mat overload (A.h:124-125) */
int ref_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int disp_type, int win, int min_d, int num_d, float* disp) {
    GUARD(return put_f32(computeNCC(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), (DisparityType)disp_type, win, min_d, num_d), H, W, disp);)
}
/* computeNCC, vector overload (A.h:126-128): vol [num_d][H][W] */
int ref_cost_ncc(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d, int disp_type, int win, float* vol) {
    GUARD(std::vector<cv::Mat> costs;
          computeNCC(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), costs, (DisparityType)disp_type, win, min_d, num_d);
          if ((int)costs.size() != num_d) return -1;
          for (int d = 0; d < num_d; d++) { int rc = put_f32(costs[d], H, W, vol + (size_t)d * H * W); if (rc) return rc; }
          return 0;)
}
/* computeSimilarity 7-arg (A.h:112-114): vol [num_d][H][W] */
int ref_cost_tad_cg(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d, int disp_type, double regularity,
                    double thres_c, double thres_g, float* vol) {
    GUARD(std::vector<cv::Mat> costs;
          computeSimilarity(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), costs, regularity, thres_c, thres_g,
                            (DisparityType)disp_type, min_d, num_d);
          if ((int)costs.size() != num_d) return -1;
          for (int d = 0; d < num_d; d++) { int rc = put_f32(costs[d], H, W, vol + (size_t)d * H * W); if (rc) return rc; }
          return 0;)
}
/* computeSimilarity 8-arg (A.h:115-117): vol [num_d][H + 2h][W + 2h] */
int ref_cost_tad_cg_padded(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d, int disp_type,
                           double regularity, double thres_c, double thres_g, int win, float* vol) {
    GUARD(std::vector<cv::Mat> costs;
          computeSimilarity(wrap_u8(L, H, W, 3), wrap_u8(R, H, W, 3), costs, regularity, thres_c, thres_g,
                            (DisparityType)disp_type, win, min_d, num_d);
          if ((int)costs.size() != num_d) return -1;
          int h = win / 2, Hp = H + 2 * h, Wp = W + 2 * h;
          for (int d = 0; d < num_d; d++) { int rc = put_f32(costs[d], Hp, Wp, vol + (size_t)d * Hp * Wp); if (rc) return rc; }
          return 0;)
}
/* getCostSAD_d for every d exactly as computeAdaptiveWeight_BLO1 calls it (A.cpp:2511-2536): vol [num_d][H][W] */
int ref_cost_sad_box(const uint8_t* L, const uint8_t* R, int H, int W, int min_d, int num_d, int disp_type, int win, float* vol) {
    GUARD(cv::Mat l = wrap_u8(L, H, W, 3), r = wrap_u8(R, H, W, 3), lb, rb;
          cv::cvtColor(l, l, cv::COLOR_BGR2GRAY); cv::cvtColor(r, r, cv::COLOR_BGR2GRAY);
          int max_offset = min_d + num_d - 1;
          cv::copyMakeBorder(l, lb, 0, 0, 0, max_offset, cv::BORDER_REFLECT);
          cv::copyMakeBorder(r, rb, 0, 0, max_offset, 0, cv::BORDER_REFLECT);
          for (int i = min_d; i <= max_offset; i++) {
              cv::Mat c = disp_type == 0 ? getCostSAD_d(l, rb, i, DISPARITY_LEFT, win) : getCostSAD_d(lb, r, i, DISPARITY_RIGHT, win);
              int rc = put_f32(c, H, W, vol + (size_t)(i - min_d) * H * W); if (rc) return rc;
          }
          return 0;)
}
/* getGuidedFilter (A.h:163): guide u8 with cn channels, p f32 */
int ref_guided_filter(const uint8_t* guide, int cn, const float* p, int H, int W, int r, double eps, float* q) {
    GUARD(return put_f32(getGuidedFilter(wrap_u8(guide, H, W, cn), wrap_f32(p, H, W), r, eps), H, W, q);)
}

/* ---- the OpenCV stand-in's primitives, exported so that tests can pin them against the real cv2 ---- */
int shim_bgr2gray(const uint8_t* bgr, int H, int W, uint8_t* gray) {
    GUARD(cv::Mat g; cv::cvtColor(wrap_u8(bgr, H, W, 3), g, cv::COLOR_BGR2GRAY);
          for (int y = 0; y < H; y++) memcpy(gray + (size_t)y * W, g.ptr(y), W); return 0;)
}
int shim_box_filter_f32(const float* src, int H, int W, int k, float* dst) {
    GUARD(cv::Mat o; cv::boxFilter(wrap_f32(src, H, W), o, CV_32F, cv::Size(k, k)); return put_f32(o, H, W, dst);)
}
int shim_normalize_f32(const float* src, int H, int W, float* dst) {
    GUARD(cv::Mat o; cv::normalize(wrap_f32(src, H, W), o, 0, 1, cv::NORM_MINMAX, CV_32F); return put_f32(o, H, W, dst);)
}
int shim_normalize_u8c3(const uint8_t* src, int H, int W, float* dst) {
    GUARD(cv::Mat o; cv::normalize(wrap_u8(src, H, W, 3), o, 0, 1, cv::NORM_MINMAX, CV_32F);
          for (int y = 0; y < H; y++) memcpy(dst + (size_t)y * W * 3, o.ptr(y), (size_t)W * 12); return 0;)
}
int shim_scharr_x_u8c3(const uint8_t* src, int H, int W, float* dst) {
    GUARD(cv::Mat k = (cv::Mat_<char>(3, 3) << -3, 0, 3, -10, 0, 10, -3, 0, 3), o;
          cv::filter2D(wrap_u8(src, H, W, 3), o, CV_32F, k);
          for (int y = 0; y < H; y++) memcpy(dst + (size_t)y * W * 3, o.ptr(y), (size_t)W * 12); return 0;)
}
/* (a + b + c) / 3 on u8 planes and on f32 planes: the MatExpr lowering of A.cpp:459 / 473 */
int shim_mean3_u8(const uint8_t* a, const uint8_t* b, const uint8_t* c, int H, int W, uint8_t* out) {
    GUARD(cv::Mat o = (wrap_u8(a, H, W, 1) + wrap_u8(b, H, W, 1) + wrap_u8(c, H, W, 1)) / 3;
          for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, o.ptr(y), W); return 0;)
}
int shim_mean3_f32(const float* a, const float* b, const float* c, int H, int W, float* out) {
    GUARD(cv::Mat o = (wrap_f32(a, H, W) + wrap_f32(b, H, W) + wrap_f32(c, H, W)) / 3; return put_f32(o, H, W, out);)
}
/* alpha * A + beta * B on f32 (A.cpp:484) */
int shim_blend_f32(const float* a, double alpha, const float* b, double beta, int H, int W, float* out) {
    GUARD(cv::Mat o = alpha * wrap_f32(a, H, W) + beta * wrap_f32(b, H, W); return put_f32(o, H, W, out);)
}
/* color.mul(mask / 255) + T * (mask / 255) on u8 (A.cpp:464) */
int shim_trunc_u8(const uint8_t* color, double T, int H, int W, uint8_t* out) {
    GUARD(cv::Mat c = wrap_u8(color, H, W, 1), mask(c.size(), CV_8UC1);
          cv::compare(c, T, mask, cv::CMP_GT);
          cv::Mat o = c.mul(mask / 255) + T * (mask / 255);
          for (int y = 0; y < H; y++) memcpy(out + (size_t)y * W, o.ptr(y), W); return 0;)
}
int shim_exp_f32(const float* src, int H, int W, float* dst) {
    GUARD(cv::Mat o; cv::exp(wrap_f32(src, H, W), o); return put_f32(o, H, W, dst);)
}
}
