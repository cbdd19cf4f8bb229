/*
 * opencv2/opencv.hpp -- a stand-in for the slice of OpenCV's C++ API that the reference
 * (ZhangYY12345/aswStereoMatch, aswStereoMatch/methods/aswMethods.{h,cpp}) uses.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/_ref).  The reference needs OpenCV 4.1.0 C++ headers and
 * libraries, which this image does not have (only the Python cv2 4.13 wheel).  With this header
 * on the include path the reference's OWN, UNMODIFIED aswMethods.cpp compiles with g++ into
 * oracle/_ref/libasw_ref.so (recipe: oracle/Makefile), so that the plain-C restatement in
 * oracle/asw_oracle.c can be checked against the reference's own code.
 *
 * What is restated here is OpenCV, not the reference:
 *   - cv::Mat (ref-counted buffer, ROI views, create() reusing a matching buffer),
 *   - the lazy cv::MatExpr algebra with OpenCV's fusion rules (modules/core/src/matop.cpp:
 *     MatOp::add/subtract/multiply/divide, MatOp_AddEx, MatOp_Bin, MatOp_Cmp), because the
 *     reference's arithmetic depends on how an expression such as (a + b + c) / 3 lowers
 *     (SURVEY.md Appendix B),
 *   - the primitives with OpenCV-4.13 arithmetic: cvtColor, copyMakeBorder, filter2D, boxFilter,
 *     normalize, absdiff, compare, convertTo, addWeighted, scaleAdd, multiply, divide, exp, sum ...
 * The primitives are pinned against the real cv2 4.13 by tests/test_cpu_ref.py.  Anything the
 * hot path never reaches (StereoBM / StereoSGBM) is declared and throws.
 */
#ifndef ASW_REFSHIM_OPENCV_HPP
#define ASW_REFSHIM_OPENCV_HPP

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
#define CV_8SC1 CV_MAKETYPE(CV_8S, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

typedef unsigned char uchar;
typedef signed char schar;
typedef unsigned short ushort;
typedef std::string String;

class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
namespace Error { enum { StsOk = 0, StsBadArg = -5, StsUnmatchedFormats = -205, StsNotImplemented = -213, StsAssert = -215 }; }
#define CV_Error(code, msg) throw cv::Exception((code), (msg))
#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(cv::Error::StsAssert, #expr); } while (0)

inline long long getTickCount() { return 0; }
inline double getTickFrequency() { return 1e9; }

// ---- rounding / saturation (core/fast_math.hpp, core/saturate.hpp): cvRound = round half to even ----
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
inline int cvFloor(int v) { return v; }
inline int cvCeil(int v) { return v; }

template <typename T> inline T saturate_cast(int v) { return (T)v; }
template <typename T> inline T saturate_cast(float v) { return (T)v; }
template <typename T> inline T saturate_cast(double v) { return (T)v; }
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>(cvRound(v)); }
template <> inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>(cvRound(v)); }
template <> inline schar saturate_cast<schar>(int v) { return (schar)(v < -128 ? -128 : v > 127 ? 127 : v); }
template <> inline schar saturate_cast<schar>(float v) { return saturate_cast<schar>(cvRound(v)); }
template <> inline schar saturate_cast<schar>(double v) { return saturate_cast<schar>(cvRound(v)); }
template <> inline short saturate_cast<short>(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
template <> inline short saturate_cast<short>(float v) { return saturate_cast<short>(cvRound(v)); }
template <> inline short saturate_cast<short>(double v) { return saturate_cast<short>(cvRound(v)); }
template <> inline ushort saturate_cast<ushort>(int v) { return (ushort)(v < 0 ? 0 : v > 65535 ? 65535 : v); }
template <> inline ushort saturate_cast<ushort>(float v) { return saturate_cast<ushort>(cvRound(v)); }
template <> inline ushort saturate_cast<ushort>(double v) { return saturate_cast<ushort>(cvRound(v)); }
template <> inline int saturate_cast<int>(float v) { return cvRound(v); }
template <> inline int saturate_cast<int>(double v) { return cvRound(v); }

// ---- small fixed-size types ----
template <typename T, int n> struct Vec {
    T val[n];
    typedef T value_type;
    enum { channels = n };
    Vec() { for (int i = 0; i < n; i++) val[i] = T(0); }
    Vec(T v0) : Vec() { val[0] = v0; }
    Vec(T v0, T v1) : Vec() { static_assert(n >= 2, ""); val[0] = v0; val[1] = v1; }
    Vec(T v0, T v1, T v2) : Vec() { static_assert(n >= 3, ""); val[0] = v0; val[1] = v1; val[2] = v2; }
    Vec(T v0, T v1, T v2, T v3) : Vec() { static_assert(n >= 4, ""); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    Vec(T v0, T v1, T v2, T v3, T v4, T v5) : Vec() {
        static_assert(n >= 6, ""); val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; val[4] = v4; val[5] = v5;
    }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    bool operator==(const Vec& o) const { for (int i = 0; i < n; i++) if (val[i] != o.val[i]) return false; return true; }
    bool operator!=(const Vec& o) const { return !(*this == o); }
};
typedef Vec<uchar, 2> Vec2b; typedef Vec<uchar, 3> Vec3b; typedef Vec<uchar, 4> Vec4b;
typedef Vec<short, 2> Vec2s; typedef Vec<short, 3> Vec3s;
typedef Vec<int, 2> Vec2i; typedef Vec<int, 3> Vec3i; typedef Vec<int, 4> Vec4i; typedef Vec<int, 6> Vec6i;
typedef Vec<float, 2> Vec2f; typedef Vec<float, 3> Vec3f; typedef Vec<float, 4> Vec4f; typedef Vec<float, 6> Vec6f;
typedef Vec<double, 2> Vec2d; typedef Vec<double, 3> Vec3d; typedef Vec<double, 4> Vec4d; typedef Vec<double, 6> Vec6d;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point_& o) const { return !(*this == o); }
};
typedef Point_<int> Point2i; typedef Point2i Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<int> Point3i; typedef Point3_<float> Point3f; typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size_& o) const { return !(*this == o); }
};
typedef Size_<int> Size2i; typedef Size2i Size; typedef Size_<float> Size2f;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect2i; typedef Rect2i Rect;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT_MIN, INT_MAX); }
};

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar_ all(T v) { return Scalar_(v, v, v, v); }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    bool isReal() const { return val[1] == 0 && val[2] == 0 && val[3] == 0; }
    bool operator==(const Scalar_& o) const { return val[0] == o.val[0] && val[1] == o.val[1] && val[2] == o.val[2] && val[3] == o.val[3]; }
    bool operator!=(const Scalar_& o) const { return !(*this == o); }
    Scalar_ operator-() const { return Scalar_(-val[0], -val[1], -val[2], -val[3]); }
    Scalar_& operator+=(const Scalar_& o) { for (int i = 0; i < 4; i++) val[i] += o.val[i]; return *this; }
    Scalar_& operator-=(const Scalar_& o) { for (int i = 0; i < 4; i++) val[i] -= o.val[i]; return *this; }
    Scalar_& operator*=(T s) { for (int i = 0; i < 4; i++) val[i] *= s; return *this; }
};
typedef Scalar_<double> Scalar;
inline Scalar operator-(const Scalar& a, const Scalar& b) { Scalar r = a; r -= b; return r; }
inline Scalar operator+(const Scalar& a, const Scalar& b) { Scalar r = a; r += b; return r; }
inline Scalar operator*(const Scalar& a, double s) { Scalar r = a; r *= s; return r; }

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<schar> { enum { value = CV_8S }; };
template <> struct DataDepth<char> { enum { value = CV_8S }; };
template <> struct DataDepth<ushort> { enum { value = CV_16U }; };
template <> struct DataDepth<short> { enum { value = CV_16S }; };
template <> struct DataDepth<int> { enum { value = CV_32S }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_MINMAX = 32 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_GRAY2BGR = 8, COLOR_BGR2HSV = 40, COLOR_HSV2BGR = 54 };

class Mat;
class MatExpr;

// Mat::size is both callable (size()) and comparable (a.size != b.size) in OpenCV
struct MatSize {
    const Mat* m;
    explicit MatSize(const Mat* m_) : m(m_) {}
    Size operator()() const;
    bool operator==(const MatSize& o) const;
    bool operator!=(const MatSize& o) const { return !(*this == o); }
};

class Mat {
public:
    int flags;                 // the type (depth + channels)
    int dims, rows, cols;
    uchar* data;
    size_t step;               // bytes per row
    std::shared_ptr<std::vector<uchar>> buf;
    MatSize size;

    Mat() : flags(0), dims(2), rows(0), cols(0), data(nullptr), step(0), size(this) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar& s) : Mat() { create(sz.height, sz.width, type); setTo(s); }
    Mat(const Mat& o) : flags(o.flags), dims(o.dims), rows(o.rows), cols(o.cols), data(o.data), step(o.step), buf(o.buf), size(this) {}
    Mat& operator=(const Mat& o) {
        if (this != &o) { flags = o.flags; dims = o.dims; rows = o.rows; cols = o.cols; data = o.data; step = o.step; buf = o.buf; }
        return *this;
    }
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s) { setTo(s); return *this; }

    int type() const { return flags; }
    int depth() const { return CV_MAT_DEPTH(flags); }
    int channels() const { return CV_MAT_CN(flags); }
    size_t elemSize1() const { static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2}; return (size_t)sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }

    // cv::Mat::create: a buffer of the requested size and type is kept (and stays shared with its other headers),
    // anything else is released and replaced; new memory is zero-filled here (OpenCV leaves it uninitialised: the
    // reference's never-written output pixels read as 0, the sentinel the oracle uses too)
    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && flags == type) return;
        flags = type; rows = r; cols = c; dims = 2;
        step = (size_t)c * elemSize();
        buf = std::make_shared<std::vector<uchar>>((size_t)r * step + 16, (uchar)0);
        data = (r > 0 && c > 0) ? buf->data() : nullptr;
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }

    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }

    Mat operator()(const Rect& r) const {
        if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
            throw Exception(Error::StsAssert, "Mat ROI out of range");
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat operator()(Range rr, Range cr) const {
        if (rr.start == INT_MIN) rr = Range(0, rows);
        if (cr.start == INT_MIN) cr = Range(0, cols);
        return (*this)(Rect(cr.start, rr.start, cr.end - cr.start, rr.end - rr.start));
    }
    Mat rowRange(int a, int b) const { return (*this)(Range(a, b), Range::all()); }
    Mat colRange(int a, int b) const { return (*this)(Range::all(), Range(a, b)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const {
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.flags == flags && dst.step == step) return;
        Mat src(*this);
        dst.create(rows, cols, flags);
        for (int y = 0; y < rows; y++) memmove(dst.data + (size_t)y * dst.step, src.data + (size_t)y * src.step, (size_t)cols * elemSize());
    }
    Mat reshape(int cn, int rows_ = 0) const;
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat& setTo(const Scalar& s);
    MatExpr mul(const Mat& m, double scale = 1) const;
    MatExpr mul(const MatExpr& e, double scale = 1) const;
    static MatExpr zeros(int r, int c, int type);
    static MatExpr zeros(Size s, int type);
    static MatExpr ones(int r, int c, int type);
    static MatExpr ones(Size s, int type);
};

inline Size MatSize::operator()() const { return Size(m->cols, m->rows); }
inline bool MatSize::operator==(const MatSize& o) const { return m->rows == o.m->rows && m->cols == o.m->cols; }

// Mat_<T>(rows, cols) << v0, v1, ... (the reference builds its 3x3 gradient kernel this way)
template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(DataDepth<T>::value, 1)) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
};
template <typename T> class MatCommaInitializer_ {
    Mat_<T> m; size_t i;
public:
    MatCommaInitializer_(const Mat_<T>& m_, T v) : m(m_), i(0) { *this, v; }
    MatCommaInitializer_& operator,(T v) {
        if (i >= m.total()) throw Exception(Error::StsAssert, "too many initializers");
        m.template at<T>((int)(i / m.cols), (int)(i % m.cols)) = v; i++;
        return *this;
    }
    operator Mat() const { return m; }
    operator Mat_<T>() const { return m; }
};
template <typename T, typename U> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) { return MatCommaInitializer_<T>(m, (T)v); }

// =====================================================================================================
// primitives (implemented in oracle/refshim/mini_cv.cpp)
// =====================================================================================================
void add(const Mat& a, const Mat& b, Mat& dst);
void add(const Mat& a, const Scalar& s, Mat& dst);
void subtract(const Mat& a, const Mat& b, Mat& dst);
void subtract(const Mat& a, const Scalar& s, Mat& dst);
void subtract(const Scalar& s, const Mat& a, Mat& dst);
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1, int dtype = -1);
void divide(const Mat& a, const Mat& b, Mat& dst, double scale = 1, int dtype = -1);
void divide(double scale, const Mat& b, Mat& dst, int dtype = -1);
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst);
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst, int dtype = -1);
void absdiff(const Mat& a, const Mat& b, Mat& dst);
void absdiff(const Mat& a, const Scalar& s, Mat& dst);
inline void absdiff(const Mat& a, double s, Mat& dst) { absdiff(a, Scalar(s), dst); }
void compare(const Mat& a, const Mat& b, Mat& dst, int cmpop);
void compare(const Mat& a, double s, Mat& dst, int cmpop);
void bitwise_not(const Mat& a, Mat& dst);
void split(const Mat& m, std::vector<Mat>& mv);
void merge(const std::vector<Mat>& mv, Mat& dst);
void exp(const Mat& src, Mat& dst);
Scalar sum(const Mat& m);
Scalar mean(const Mat& m);
void minMaxLoc(const Mat& m, double* minVal, double* maxVal = 0, Point* minLoc = 0, Point* maxLoc = 0);
void normalize(const Mat& src, Mat& dst, double alpha = 1, double beta = 0, int norm_type = NORM_L2, int dtype = -1);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void cvtColor(const Mat& src, Mat& dst, int code, int dstCn = 0);
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0, int borderType = BORDER_DEFAULT);
void boxFilter(const Mat& src, Mat& dst, int ddepth, Size ksize, Point anchor = Point(-1, -1), bool normalize = true, int borderType = BORDER_DEFAULT);
int borderInterpolate(int p, int len, int borderType);

// =====================================================================================================
// MatExpr: OpenCV's lazy matrix expressions (core/mat.hpp, core/src/matop.cpp).  `op` plays the role of the MatOp*.
// =====================================================================================================
class MatExpr {
public:
    enum Op { OP_IDENTITY, OP_ADDEX, OP_BIN, OP_CMP };
    Op op; int flags;
    Mat a, b, c;
    double alpha, beta;
    Scalar s;
    MatExpr() : op(OP_IDENTITY), flags(0), alpha(0), beta(0) {}
    explicit MatExpr(const Mat& m) : op(OP_IDENTITY), flags(0), a(m), alpha(1), beta(0) {}
    MatExpr(Op op_, int flags_, const Mat& a_ = Mat(), const Mat& b_ = Mat(), const Mat& c_ = Mat(), double alpha_ = 1,
            double beta_ = 1, const Scalar& s_ = Scalar())
        : op(op_), flags(flags_), a(a_), b(b_), c(c_), alpha(alpha_), beta(beta_), s(s_) {}
    operator Mat() const { Mat m; assign(m); return m; }
    void assign(Mat& m, int type = -1) const;
    Size size() const { return a.size(); }
    int type() const { return op == OP_CMP ? CV_MAKETYPE(CV_8U, a.channels()) : a.type(); }
    MatExpr mul(const MatExpr& e, double scale = 1) const;
    MatExpr mul(const Mat& m, double scale = 1) const;
};

namespace matop {
inline bool isIdentity(const MatExpr& e) { return e.op == MatExpr::OP_IDENTITY; }
inline bool isAddEx(const MatExpr& e) { return e.op == MatExpr::OP_ADDEX; }
inline bool isScaled(const MatExpr& e) { return isAddEx(e) && (!e.b.data || e.beta == 0) && e.s == Scalar(); }
inline bool isBin(const MatExpr& e, char c) { return e.op == MatExpr::OP_BIN && e.flags == c; }
inline bool isReciprocal(const MatExpr& e) { return isBin(e, '/') && (!e.b.data || e.beta == 0); }
inline void makeAddEx(MatExpr& res, const Mat& a, const Mat& b, double alpha, double beta, const Scalar& s = Scalar()) {
    res = MatExpr(MatExpr::OP_ADDEX, 0, a, b, Mat(), alpha, beta, s);
}
inline void makeBin(MatExpr& res, char op, const Mat& a, const Mat& b, double scale = 1) {
    res = MatExpr(MatExpr::OP_BIN, op, a, b, Mat(), scale, b.data ? 1 : 0);
}
inline void makeBin(MatExpr& res, char op, const Mat& a, const Scalar& s) {
    res = MatExpr(MatExpr::OP_BIN, op, a, Mat(), Mat(), 1, 0, s);
}
inline void makeCmp(MatExpr& res, int cmpop, const Mat& a, const Mat& b) { res = MatExpr(MatExpr::OP_CMP, cmpop, a, b, Mat(), 1, 1); }
inline void makeCmp(MatExpr& res, int cmpop, const Mat& a, double alpha) { res = MatExpr(MatExpr::OP_CMP, cmpop, a, Mat(), Mat(), alpha, 1); }
// virtual dispatch of matop.cpp: the first argument is the `this` MatOp of the call
void add(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res);
void add(MatExpr::Op self, const MatExpr& e, const Scalar& s, MatExpr& res);
void subtract(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res);
void subtract(MatExpr::Op self, const Scalar& s, const MatExpr& e, MatExpr& res);
void multiply(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res, double scale = 1);
void multiply(MatExpr::Op self, const MatExpr& e, double s, MatExpr& res);
void divide(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res, double scale = 1);
void divide(MatExpr::Op self, double s, const MatExpr& e, MatExpr& res);
void abs(MatExpr::Op self, const MatExpr& e, MatExpr& res);
}  // namespace matop

// ---- operators (matop.cpp, same argument orders) ----
inline MatExpr operator+(const Mat& a, const Mat& b) { MatExpr e; matop::makeAddEx(e, a, b, 1, 1); return e; }
inline MatExpr operator+(const Mat& a, const Scalar& s) { MatExpr e; matop::makeAddEx(e, a, Mat(), 1, 0, s); return e; }
inline MatExpr operator+(const Scalar& s, const Mat& a) { MatExpr e; matop::makeAddEx(e, a, Mat(), 1, 0, s); return e; }
inline MatExpr operator+(const MatExpr& e, const Mat& m) { MatExpr en; matop::add(e.op, e, MatExpr(m), en); return en; }
inline MatExpr operator+(const Mat& m, const MatExpr& e) { MatExpr en; matop::add(e.op, e, MatExpr(m), en); return en; }
inline MatExpr operator+(const MatExpr& e, const Scalar& s) { MatExpr en; matop::add(e.op, e, s, en); return en; }
inline MatExpr operator+(const Scalar& s, const MatExpr& e) { MatExpr en; matop::add(e.op, e, s, en); return en; }
inline MatExpr operator+(const MatExpr& e1, const MatExpr& e2) { MatExpr en; matop::add(e1.op, e1, e2, en); return en; }

inline MatExpr operator-(const Mat& a, const Mat& b) { MatExpr e; matop::makeAddEx(e, a, b, 1, -1); return e; }
inline MatExpr operator-(const Mat& a, const Scalar& s) { MatExpr e; matop::makeAddEx(e, a, Mat(), 1, 0, -s); return e; }
inline MatExpr operator-(const Scalar& s, const Mat& a) { MatExpr e; matop::makeAddEx(e, a, Mat(), -1, 0, s); return e; }
inline MatExpr operator-(const MatExpr& e, const Mat& m) { MatExpr en; matop::subtract(e.op, e, MatExpr(m), en); return en; }
inline MatExpr operator-(const Mat& m, const MatExpr& e) { MatExpr en; matop::subtract(e.op, MatExpr(m), e, en); return en; }
inline MatExpr operator-(const MatExpr& e, const Scalar& s) { MatExpr en; matop::add(e.op, e, -s, en); return en; }
inline MatExpr operator-(const Scalar& s, const MatExpr& e) { MatExpr en; matop::subtract(e.op, s, e, en); return en; }
inline MatExpr operator-(const MatExpr& e1, const MatExpr& e2) { MatExpr en; matop::subtract(e1.op, e1, e2, en); return en; }
inline MatExpr operator-(const Mat& m) { MatExpr e; matop::makeAddEx(e, m, Mat(), -1, 0); return e; }
inline MatExpr operator-(const MatExpr& e) { MatExpr en; matop::subtract(e.op, Scalar(0), e, en); return en; }

inline MatExpr operator*(const Mat& a, double s) { MatExpr e; matop::makeAddEx(e, a, Mat(), s, 0); return e; }
inline MatExpr operator*(double s, const Mat& a) { MatExpr e; matop::makeAddEx(e, a, Mat(), s, 0); return e; }
inline MatExpr operator*(const MatExpr& e, double s) { MatExpr en; matop::multiply(e.op, e, s, en); return en; }
inline MatExpr operator*(double s, const MatExpr& e) { MatExpr en; matop::multiply(e.op, e, s, en); return en; }

inline MatExpr operator/(const Mat& a, const Mat& b) { MatExpr e; matop::makeBin(e, '/', a, b); return e; }
inline MatExpr operator/(const Mat& a, double s) { MatExpr e; matop::makeAddEx(e, a, Mat(), 1. / s, 0); return e; }
inline MatExpr operator/(double s, const Mat& a) { MatExpr e; matop::makeBin(e, '/', a, Mat(), s); return e; }
inline MatExpr operator/(const MatExpr& e, const Mat& m) { MatExpr en; matop::divide(e.op, e, MatExpr(m), en); return en; }
inline MatExpr operator/(const Mat& m, const MatExpr& e) { MatExpr en; matop::divide(e.op, MatExpr(m), e, en); return en; }
inline MatExpr operator/(const MatExpr& e, double s) { MatExpr en; matop::multiply(e.op, e, 1. / s, en); return en; }
inline MatExpr operator/(double s, const MatExpr& e) { MatExpr en; matop::divide(e.op, s, e, en); return en; }
inline MatExpr operator/(const MatExpr& e1, const MatExpr& e2) { MatExpr en; matop::divide(e1.op, e1, e2, en); return en; }

inline MatExpr operator>(const Mat& a, double s) { MatExpr e; matop::makeCmp(e, CMP_GT, a, s); return e; }
inline MatExpr operator<(const Mat& a, double s) { MatExpr e; matop::makeCmp(e, CMP_LT, a, s); return e; }
inline MatExpr operator>(const Mat& a, const Mat& b) { MatExpr e; matop::makeCmp(e, CMP_GT, a, b); return e; }
inline MatExpr operator<(const Mat& a, const Mat& b) { MatExpr e; matop::makeCmp(e, CMP_LT, a, b); return e; }

inline MatExpr abs(const Mat& a) { MatExpr e; matop::makeBin(e, 'a', a, Scalar()); return e; }
inline MatExpr abs(const MatExpr& e) { MatExpr en; matop::abs(e.op, e, en); return en; }

inline Mat& Mat::operator=(const MatExpr& e) { e.assign(*this); return *this; }
inline Mat& operator+=(Mat& a, const Mat& b) { add(a, b, a); return a; }
inline Mat& operator-=(Mat& a, const Mat& b) { subtract(a, b, a); return a; }
inline Mat& operator*=(Mat& a, double s) { a.convertTo(a, -1, s); return a; }
inline Mat& operator/=(Mat& a, double s) { a.convertTo(a, -1, 1. / s); return a; }
inline Mat& operator+=(Mat& a, const MatExpr& e) { Mat m = e; add(a, m, a); return a; }

template <typename T> using Ptr = std::shared_ptr<T>;

// ---- calib3d: outside the dense-matching hot path; declared so the file compiles ----
class StereoMatcher {
public:
    virtual ~StereoMatcher() {}
    virtual void compute(const Mat&, const Mat&, Mat&) { throw Exception(Error::StsNotImplemented, "StereoBM / StereoSGBM are not part of the oracle shim"); }
    void setMinDisparity(int) {} void setNumDisparities(int) {} void setBlockSize(int) {} void setSpeckleWindowSize(int) {}
    void setSpeckleRange(int) {} void setDisp12MaxDiff(int) {}
};
class StereoBM : public StereoMatcher {
public:
    static Ptr<StereoBM> create(int = 0, int = 21) { return std::make_shared<StereoBM>(); }
    void setTextureThreshold(int) {} void setPreFilterCap(int) {} void setUniquenessRatio(int) {} void setPreFilterSize(int) {}
    void setPreFilterType(int) {}
};
class StereoSGBM : public StereoMatcher {
public:
    enum { MODE_SGBM = 0, MODE_HH = 1, MODE_SGBM_3WAY = 2, MODE_HH4 = 3 };
    static Ptr<StereoSGBM> create(int = 0, int = 16, int = 3, int = 0, int = 0, int = 0, int = 0, int = 0, int = 0, int = 0, int = MODE_SGBM) {
        return std::make_shared<StereoSGBM>();
    }
    void setPreFilterCap(int) {} void setUniquenessRatio(int) {} void setP1(int) {} void setP2(int) {} void setMode(int) {}
};

}  // namespace cv
#endif
