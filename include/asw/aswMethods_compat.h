/*
 * aswMethods_compat.h -- C++ host mirror of the reference's method entry points over the C ABI (asw.h).
 *
 * Same function names, argument order, defaults and error behaviour as the reference's
 * aswStereoMatch/methods/aswMethods.h (A.h:91-184): images in (cv::Mat, CV_8UC3 BGR) by value, a freshly
 * allocated CV_32FC1 disparity Mat out, an EMPTY Mat on invalid arguments (A.cpp:1440-1443, 2458-2462,
 * 3238-3241).  A reference user swaps `#include "aswMethods.h"` for this header, adds
 * `using namespace asw_b200;` and links libasw_b200.so; every method body then runs on the B200.
 *
 * With OpenCV present (<opencv2/core.hpp>) the functions take / return cv::Mat.  Without it (this build
 * container has no OpenCV C++) a minimal Mat stand-in with the same fields is used so that the shim still
 * compiles and can be tested (tests/cpp/test_compat.cpp).
 */
#ifndef ASW_ASWMETHODS_COMPAT_H
#define ASW_ASWMETHODS_COMPAT_H

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "asw.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#define ASW_HAVE_OPENCV 1
#endif
#endif

namespace asw_b200 {

// P.h:4-24 (identical integer values)
enum DisparityType { DISPARITY_LEFT = 0, DISPARITY_RIGHT = 1 };
enum StereoMatchingAlgorithms {
    BM = 0, SGBM = 1, ADAPTIVE_WEIGHT = 2, ADAPTIVE_WEIGHT_8DIRECT = 3, ADAPTIVE_WEIGHT_GEODESIC = 4,
    ADAPTIVE_WEIGHT_BILATERAL_GRID = 5, ADAPTIVE_WEIGHT_BLO1 = 6, ADAPTIVE_WEIGHT_GUIDED_FILTER = 7,
    ADAPTIVE_WEIGHT_GUIDED_FILTER_2 = 8, ADAPTIVE_WEIGHT_GUIDED_FILTER_3 = 9, ADAPTIVE_WEIGHT_MEDIAN = 10, NCC = 11
};

#ifdef ASW_HAVE_OPENCV
using Mat = cv::Mat;
inline Mat make_f32(int rows, int cols) { return Mat(rows, cols, CV_32FC1); }
inline int mat_channels(const Mat& m) { return m.channels(); }
inline bool mat_is_u8(const Mat& m) { return m.depth() == CV_8U; }
#else
// minimal stand-in: ref-counted buffer, rows / cols / channels / elemSize1 / step / data, empty()
struct Mat {
    int rows = 0, cols = 0, cn = 1, esz1 = 1;
    size_t step = 0;
    unsigned char* data = nullptr;
    std::shared_ptr<std::vector<unsigned char>> buf;
    Mat() {}
    Mat(int r, int c, int channels, int elem_size1) { create(r, c, channels, elem_size1); }
    void create(int r, int c, int channels, int elem_size1) {
        rows = r; cols = c; cn = channels; esz1 = elem_size1; step = (size_t)c * channels * elem_size1;
        buf = std::make_shared<std::vector<unsigned char>>((size_t)r * step);
        data = buf->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int channels() const { return cn; }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
};
inline Mat make_f32(int rows, int cols) { return Mat(rows, cols, 1, 4); }
inline int mat_channels(const Mat& m) { return m.cn; }
inline bool mat_is_u8(const Mat& m) { return m.esz1 == 1; }
#endif

// process-wide default context on device 0 (created on first use; nullptr when no CUDA device is usable --
// there is no CPU fallback, the methods then return an empty Mat)
inline asw_ctx* default_context() {
    static asw_ctx* ctx = [] { asw_ctx* c = nullptr; if (asw_create(0, &c) != ASW_OK) c = nullptr; return c; }();
    return ctx;
}

namespace detail {
inline bool wrap_u8(const Mat& m, asw_u8_image* out) {
    if (m.empty() || !mat_is_u8(m)) return false;
    out->data = (const uint8_t*)m.data; out->rows = m.rows; out->cols = m.cols; out->channels = mat_channels(m);
    out->step = (size_t)m.step;
    return true;
}
template <typename F>
inline Mat run(const Mat& L, const Mat& R, F&& call) {
    asw_ctx* ctx = default_context();
    asw_u8_image l, r;
    if (!ctx || !wrap_u8(L, &l) || !wrap_u8(R, &r)) return Mat();
    Mat out = make_f32(L.rows, L.cols);
    asw_f32_image d{(float*)out.data, out.rows, out.cols, (size_t)out.step};
    if (call(ctx, &l, &r, &d) != ASW_OK) return Mat();
    return out;
}
}  // namespace detail

// A.h:133-134
inline Mat computeAdaptiveWeight(Mat leftImg, Mat rightImg, double gamma_c = 30, double gamma_g = 2,
                                 DisparityType dispType = DISPARITY_LEFT, int winSize = 7, int minDisparity = 186,
                                 int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight(c, l, r, d, gamma_c, gamma_g, dispType, winSize, minDisparity, numDisparity); });
}
// A.h:135-136
inline Mat computeAdaptiveWeight_direct8(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT, int winSize = 7,
                                         int minDisparity = 186, int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_direct8(c, l, r, d, dispType, winSize, minDisparity, numDisparity); });
}
// A.h:141-142
inline Mat computeAdaptiveWeight_geodesic(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                          int winSize = 7, int minDisparity = 186, int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_geodesic(c, l, r, d, dispType, winSize, minDisparity, numDisparity); });
}
// A.h:153-155
inline Mat computeAdaptiveWeight_bilateralGrid(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                               double sampleRateS = 10, double sampleRateR = 10, int minDisparity = 186,
                                               int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_bilateral_grid(c, l, r, d, dispType, sampleRateS, sampleRateR, minDisparity, numDisparity); });
}
// A.h:158-160
inline Mat computeAdaptiveWeight_BLO1(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                      double sampleRateR = 10, int winSize = 35, int minDisparity = 186,
                                      int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_blo1(c, l, r, d, dispType, sampleRateR, winSize, minDisparity, numDisparity); });
}
// A.h:170-172
inline Mat computeAdaptiveWeight_GuidedF_3(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                           double eps = 1e-6, int winSize = 35, int minDisparity = 186, int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_guidedf_3(c, l, r, d, dispType, eps, winSize, minDisparity, numDisparity); });
}
// A.h:124-125
inline Mat computeNCC(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT, int winSize = 7,
                      int minDisparity = 0, int numDisparity = 30) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_ncc(c, l, r, d, dispType, winSize, minDisparity, numDisparity); });
}
// A.h:164-166
inline Mat computeAdaptiveWeight_GuidedF(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                         double eps = 1e-8, int winSize = 35, int minDisparity = 186, int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_guidedf(c, l, r, d, dispType, eps, winSize, minDisparity, numDisparity); });
}
// A.h:167-169
inline Mat computeAdaptiveWeight_GuidedF_2(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                           double eps = 1e-8, int winSize = 35, int minDisparity = 186, int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_guidedf_2(c, l, r, d, dispType, eps, winSize, minDisparity, numDisparity); });
}
// A.h:176-179
inline Mat computeAdaptiveWeight_WeightedMedian(Mat leftImg, Mat rightImg, DisparityType dispType = DISPARITY_LEFT,
                                                int winSize = 35, double sampleRateS = 10, double sampleRateR = 10,
                                                int minDisparity = 186, int numDisparity = 144) {
    return detail::run(leftImg, rightImg, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_adaptive_weight_weighted_median(c, l, r, d, dispType, winSize, sampleRateS, sampleRateR, minDisparity, numDisparity); });
}
// A.h:91-92: the dispatcher writes its result into disparityMap (left untouched = empty on failure / out of scope)
inline void stereoMatching(Mat srcLeft, Mat srcRight, Mat& disparityMap, DisparityType disparityType,
                           StereoMatchingAlgorithms algorithmType, int winSize = 15, int minDisparity = 0,
                           int numDisparity = 64) {
    disparityMap = detail::run(srcLeft, srcRight, [&](asw_ctx* c, const asw_u8_image* l, const asw_u8_image* r, asw_f32_image* d) {
        return asw_stereo_matching(c, l, r, d, disparityType, algorithmType, winSize, minDisparity, numDisparity); });
}

}  // namespace asw_b200
#endif
