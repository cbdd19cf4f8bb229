/*
 * mini_cv.cpp -- implementation of the OpenCV stand-in declared in refshim/opencv2/opencv.hpp.
 * TEST INFRASTRUCTURE ONLY (see the header).  Arithmetic follows OpenCV 4.13 (the version of the cv2
 * wheel the primitives are pinned against in tests/test_cpu_ref.py); every non-obvious lowering cites
 * SURVEY.md Appendix B.
 */
#include "opencv2/opencv.hpp"

namespace cv {

static inline void need(bool ok, const char* what) { if (!ok) throw Exception(Error::StsAssert, what); }
static inline void unsupported(const char* what) { throw Exception(Error::StsNotImplemented, std::string("oracle shim: ") + what); }

// load / store one channel value by depth
static inline double ld(const uchar* p, int depth) {
    switch (depth) {
    case CV_8U: return *p;
    case CV_8S: return *(const schar*)p;
    case CV_16U: return *(const ushort*)p;
    case CV_16S: return *(const short*)p;
    case CV_32S: return *(const int*)p;
    case CV_32F: return *(const float*)p;
    default: return *(const double*)p;
    }
}
template <typename W> static inline void st(uchar* p, int depth, W v) {
    switch (depth) {
    case CV_8U: *p = saturate_cast<uchar>(v); break;
    case CV_8S: *(schar*)p = saturate_cast<schar>(v); break;
    case CV_16U: *(ushort*)p = saturate_cast<ushort>(v); break;
    case CV_16S: *(short*)p = saturate_cast<short>(v); break;
    case CV_32S: *(int*)p = saturate_cast<int>(v); break;
    case CV_32F: *(float*)p = (float)v; break;
    default: *(double*)p = (double)v; break;
    }
}

// write `out` into `dst` the way an OpenCV function does after _dst.create(): a dst that already has the right size and
// type keeps its buffer (other headers of that buffer see the result), otherwise dst gets the new buffer
static void deliver(Mat& dst, const Mat& out) {
    if (dst.data && dst.data != out.data && dst.rows == out.rows && dst.cols == out.cols && dst.type() == out.type()) {
        for (int y = 0; y < out.rows; y++) memcpy(dst.ptr(y), out.ptr(y), (size_t)out.cols * out.elemSize());
    } else {
        dst = out;
    }
}

// =====================================================================================================
// Mat members
// =====================================================================================================
Mat Mat::reshape(int cn, int rows_) const {
    need(isContinuous(), "reshape needs a continuous matrix");
    Mat m(*this);
    if (cn == 0) cn = channels();
    size_t total_ch = (size_t)rows * cols * channels();
    int r = rows_ == 0 ? rows : rows_;
    need(total_ch % ((size_t)r * cn) == 0, "bad reshape");
    m.flags = CV_MAKETYPE(depth(), cn); m.rows = r; m.cols = (int)(total_ch / ((size_t)r * cn)); m.step = (size_t)m.cols * m.elemSize();
    return m;
}

Mat& Mat::setTo(const Scalar& s) {
    int cn = channels(), d = depth();
    need(cn <= 4 || (s[0] == s[1] && s[1] == s[2] && s[2] == s[3]) || s.isReal(), "setTo: more than 4 channels");
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++)
            for (int c = 0; c < cn; c++) st<double>(ptr(y) + ((size_t)x * cn + c) * elemSize1(), d, s[c < 4 ? c : 0]);
    return *this;
}

// Mat::convertTo (core/src/convert_scale.simd.hpp): alpha = 1, beta = 0 is a plain saturating cast; otherwise the work
// type is float (double for 32S / 64F sources or a 64F destination) and the operation is a fused multiply-add
// (x * alpha + beta; verified against cv2 for u8 -> f32 and f32 -> f32, SURVEY B-10)
void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    Mat src(*this);
    int sd = depth(), cn = channels();
    int dd = rtype < 0 ? sd : CV_MAT_DEPTH(rtype);
    if (sd == dd && alpha == 1 && beta == 0) { src.copyTo(dst); return; }
    Mat out(rows, cols, CV_MAKETYPE(dd, cn));
    bool noscale = alpha == 1 && beta == 0;
    bool dbl = sd == CV_32S || sd == CV_64F || dd == CV_64F;
    size_t se = src.elemSize1(), de = out.elemSize1();
    float af = (float)alpha, bf = (float)beta;
    for (int y = 0; y < rows; y++) {
        const uchar* sp = src.ptr(y); uchar* dp = out.ptr(y);
        for (int i = 0; i < cols * cn; i++) {
            double v = ld(sp + i * se, sd);
            if (noscale) {
                if (sd == CV_32F) st<float>(dp + i * de, dd, (float)v); else st<double>(dp + i * de, dd, v);
            } else if (dbl) st<double>(dp + i * de, dd, fma(v, alpha, beta));
            else st<float>(dp + i * de, dd, fmaf((float)v, af, bf));
        }
    }
    deliver(dst, out);
}

MatExpr Mat::zeros(int r, int c, int type) { return MatExpr(Mat(r, c, type, Scalar::all(0))); }
MatExpr Mat::zeros(Size s, int type) { return zeros(s.height, s.width, type); }
MatExpr Mat::ones(int r, int c, int type) { return MatExpr(Mat(r, c, type, Scalar(1))); }   // like OpenCV: first channel only
MatExpr Mat::ones(Size s, int type) { return ones(s.height, s.width, type); }

// =====================================================================================================
// element-wise arithmetic (core/src/arithm.cpp).  Same-type operands only unless stated: mixed types without an
// explicit dtype raise, exactly what makes the reference's DISPARITY_RIGHT cost branch throw.
// =====================================================================================================
static void check_same(const Mat& a, const Mat& b, const char* fn) {
    if (a.rows != b.rows || a.cols != b.cols || a.channels() != b.channels())
        throw Exception(Error::StsUnmatchedFormats, std::string(fn) + ": the operands have different sizes / channel counts");
    if (a.type() != b.type())
        throw Exception(Error::StsBadArg, std::string(fn) + ": the input arrays have different types, the output array type must be explicitly specified");
}

template <typename F> static void binary(const Mat& a_, const Mat& b_, Mat& dst, const char* fn, F f) {
    Mat a(a_), b(b_);
    check_same(a, b, fn);
    Mat out(a.rows, a.cols, a.type());
    int d = a.depth(), n = a.cols * a.channels();
    size_t e = a.elemSize1();
    for (int y = 0; y < a.rows; y++) {
        const uchar *pa = a.ptr(y), *pb = b.ptr(y); uchar* po = out.ptr(y);
        for (int i = 0; i < n; i++) f(pa + i * e, pb + i * e, po + i * e, d);
    }
    deliver(dst, out);
}
template <typename F> static void unary_scalar(const Mat& a_, const Scalar& s, Mat& dst, F f) {
    Mat a(a_);
    Mat out(a.rows, a.cols, a.type());
    int d = a.depth(), cn = a.channels();
    need(cn <= 4 || s.isReal(), "scalar operation on more than 4 channels");
    size_t e = a.elemSize1();
    for (int y = 0; y < a.rows; y++) {
        const uchar* pa = a.ptr(y); uchar* po = out.ptr(y);
        for (int x = 0; x < a.cols; x++)
            for (int c = 0; c < cn; c++) {
                size_t i = (size_t)x * cn + c;
                f(pa + i * e, cn <= 4 ? s[c] : s[0], po + i * e, d);
            }
    }
    deliver(dst, out);
}

void add(const Mat& a, const Mat& b, Mat& dst) {
    binary(a, b, dst, "add", [](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = *(const float*)pa + *(const float*)pb;
        else if (d == CV_64F) *(double*)po = *(const double*)pa + *(const double*)pb;
        else st<int>(po, d, (int)ld(pa, d) + (int)ld(pb, d));
    });
}
void subtract(const Mat& a, const Mat& b, Mat& dst) {
    binary(a, b, dst, "subtract", [](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = *(const float*)pa - *(const float*)pb;
        else if (d == CV_64F) *(double*)po = *(const double*)pa - *(const double*)pb;
        else st<int>(po, d, (int)ld(pa, d) - (int)ld(pb, d));
    });
}
// array op scalar: the scalar is converted to the work type of the array (float for 32F, int-rounded for integers when
// it is integral, double otherwise)
void add(const Mat& a, const Scalar& s, Mat& dst) {
    unary_scalar(a, s, dst, [](const uchar* pa, double v, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = *(const float*)pa + (float)v;
        else st<double>(po, d, ld(pa, d) + v);
    });
}
void subtract(const Mat& a, const Scalar& s, Mat& dst) {
    unary_scalar(a, s, dst, [](const uchar* pa, double v, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = *(const float*)pa - (float)v;
        else st<double>(po, d, ld(pa, d) - v);
    });
}
void subtract(const Scalar& s, const Mat& a, Mat& dst) {
    unary_scalar(a, s, dst, [](const uchar* pa, double v, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = (float)v - *(const float*)pa;
        else st<double>(po, d, v - ld(pa, d));
    });
}
void absdiff(const Mat& a, const Mat& b, Mat& dst) {
    binary(a, b, dst, "absdiff", [](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = fabsf(*(const float*)pa - *(const float*)pb);
        else if (d == CV_64F) *(double*)po = fabs(*(const double*)pa - *(const double*)pb);
        else st<int>(po, d, std::abs((int)ld(pa, d) - (int)ld(pb, d)));
    });
}
void absdiff(const Mat& a, const Scalar& s, Mat& dst) {
    unary_scalar(a, s, dst, [](const uchar* pa, double v, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = fabsf(*(const float*)pa - (float)v);
        else st<double>(po, d, fabs(ld(pa, d) - v));
    });
}
// cv::multiply: integers go through float with a float scale (scale * a * b, left to right); 32F is a * b (* scale)
void multiply(const Mat& a, const Mat& b, Mat& dst, double scale, int dtype) {
    need(dtype < 0, "multiply: explicit dtype");
    float sf = (float)scale;
    binary(a, b, dst, "multiply", [scale, sf](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) { float x = *(const float*)pa, y = *(const float*)pb; *(float*)po = scale == 1 ? x * y : sf * x * y; }
        else if (d == CV_64F) { double x = *(const double*)pa, y = *(const double*)pb; *(double*)po = scale == 1 ? x * y : scale * x * y; }
        else if (d == CV_32S) st<double>(po, d, scale * ld(pa, d) * ld(pb, d));
        else if (scale == 1) st<int>(po, d, (int)ld(pa, d) * (int)ld(pb, d));
        else st<float>(po, d, sf * (float)ld(pa, d) * (float)ld(pb, d));
    });
}
// cv::divide on floating point is IEEE (x / 0 = inf / nan) since OpenCV 4; integers: x / 0 = 0
void divide(const Mat& a, const Mat& b, Mat& dst, double scale, int dtype) {
    need(dtype < 0, "divide: explicit dtype");
    float sf = (float)scale;
    binary(a, b, dst, "divide", [scale, sf](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) { float x = *(const float*)pa, y = *(const float*)pb; *(float*)po = scale == 1 ? x / y : sf * x / y; }
        else if (d == CV_64F) { double x = *(const double*)pa, y = *(const double*)pb; *(double*)po = scale == 1 ? x / y : scale * x / y; }
        else { double y = ld(pb, d); if (y == 0) st<int>(po, d, 0); else st<float>(po, d, sf * (float)ld(pa, d) / (float)y); }
    });
}
void divide(double scale, const Mat& b_, Mat& dst, int dtype) {
    need(dtype < 0, "divide: explicit dtype");
    Mat b(b_);
    Mat out(b.rows, b.cols, b.type());
    int d = b.depth(), n = b.cols * b.channels(); size_t e = b.elemSize1();
    for (int y = 0; y < b.rows; y++)
        for (int i = 0; i < n; i++) {
            const uchar* pb = b.ptr(y) + i * e; uchar* po = out.ptr(y) + i * e;
            if (d == CV_32F) *(float*)po = (float)scale / *(const float*)pb;
            else if (d == CV_64F) *(double*)po = scale / *(const double*)pb;
            else { double v = ld(pb, d); if (v == 0) st<int>(po, d, 0); else st<float>(po, d, (float)scale / (float)v); }
        }
    deliver(dst, out);
}
// cv::addWeighted: 8U..16S in float (float alpha / beta / gamma); 32F with double scalars, fused (SURVEY B-4; the oracle
// pins (float)fma(a, alpha, b * beta) bit-exactly against cv2 4.13)
void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst, int dtype) {
    need(dtype < 0, "addWeighted: explicit dtype");
    float af = (float)alpha, bf = (float)beta, gf = (float)gamma;
    binary(a, b, dst, "addWeighted", [=](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) { double r = fma((double)*(const float*)pa, alpha, (double)*(const float*)pb * beta); *(float*)po = (float)(gamma == 0 ? r : r + gamma); }
        else if (d == CV_64F || d == CV_32S) st<double>(po, d, ld(pa, d) * alpha + ld(pb, d) * beta + gamma);
        else st<float>(po, d, (float)ld(pa, d) * af + (float)ld(pb, d) * bf + gf);
    });
}
// cv::scaleAdd: dst = a * alpha + b; below 32F it IS addWeighted(a, alpha, b, 1, 0) (SURVEY B-2)
void scaleAdd(const Mat& a, double alpha, const Mat& b, Mat& dst) {
    if (a.depth() < CV_32F) { addWeighted(a, alpha, b, 1, 0, dst); return; }
    float af = (float)alpha;
    binary(a, b, dst, "scaleAdd", [alpha, af](const uchar* pa, const uchar* pb, uchar* po, int d) {
        if (d == CV_32F) *(float*)po = fmaf(*(const float*)pa, af, *(const float*)pb);
        else *(double*)po = fma(*(const double*)pa, alpha, *(const double*)pb);
    });
}

static inline bool cmp(double x, double y, int op) {
    switch (op) {
    case CMP_EQ: return x == y; case CMP_GT: return x > y; case CMP_GE: return x >= y;
    case CMP_LT: return x < y; case CMP_LE: return x <= y; default: return x != y;
    }
}
void compare(const Mat& a_, const Mat& b_, Mat& dst, int op) {
    Mat a(a_), b(b_);
    check_same(a, b, "compare");
    Mat out(a.rows, a.cols, CV_MAKETYPE(CV_8U, a.channels()));
    int d = a.depth(), n = a.cols * a.channels(); size_t e = a.elemSize1();
    for (int y = 0; y < a.rows; y++)
        for (int i = 0; i < n; i++) out.ptr(y)[i] = cmp(ld(a.ptr(y) + i * e, d), ld(b.ptr(y) + i * e, d), op) ? 255 : 0;
    deliver(dst, out);
}
void compare(const Mat& a_, double s, Mat& dst, int op) {
    Mat a(a_);
    Mat out(a.rows, a.cols, CV_MAKETYPE(CV_8U, a.channels()));
    int d = a.depth(), n = a.cols * a.channels(); size_t e = a.elemSize1();
    for (int y = 0; y < a.rows; y++)
        for (int i = 0; i < n; i++) out.ptr(y)[i] = cmp(ld(a.ptr(y) + i * e, d), s, op) ? 255 : 0;
    deliver(dst, out);
}
void bitwise_not(const Mat& a_, Mat& dst) {
    Mat a(a_);
    Mat out(a.rows, a.cols, a.type());
    size_t n = (size_t)a.cols * a.elemSize();
    for (int y = 0; y < a.rows; y++)
        for (size_t i = 0; i < n; i++) out.ptr(y)[i] = (uchar)~a.ptr(y)[i];
    deliver(dst, out);
}

void split(const Mat& m_, std::vector<Mat>& mv) {
    Mat m(m_);
    int cn = m.channels(); size_t e = m.elemSize1();
    mv.resize(cn);
    for (int c = 0; c < cn; c++) {
        Mat out(m.rows, m.cols, CV_MAKETYPE(m.depth(), 1));
        for (int y = 0; y < m.rows; y++)
            for (int x = 0; x < m.cols; x++) memcpy(out.ptr(y) + x * e, m.ptr(y) + ((size_t)x * cn + c) * e, e);
        deliver(mv[c], out);
    }
}
void merge(const std::vector<Mat>& mv, Mat& dst) {
    need(!mv.empty(), "merge: empty vector");
    int cn = 0;
    for (const Mat& m : mv) {
        need(m.rows == mv[0].rows && m.cols == mv[0].cols && m.depth() == mv[0].depth(), "merge: planes differ");
        cn += m.channels();
    }
    size_t e = mv[0].elemSize1();
    Mat out(mv[0].rows, mv[0].cols, CV_MAKETYPE(mv[0].depth(), cn));
    int c0 = 0;
    for (const Mat& m : mv) {
        int mc = m.channels();
        for (int y = 0; y < m.rows; y++)
            for (int x = 0; x < m.cols; x++) memcpy(out.ptr(y) + ((size_t)x * cn + c0) * e, m.ptr(y) + (size_t)x * mc * e, mc * e);
        c0 += mc;
    }
    deliver(dst, out);
}

// cv::exp on 32F: OpenCV's own polynomial, max relative error 1e-7 (SURVEY B-14); restated as the correctly rounded value
void exp(const Mat& src_, Mat& dst) {
    Mat src(src_);
    need(src.depth() == CV_32F || src.depth() == CV_64F, "exp: floating point only");
    Mat out(src.rows, src.cols, src.type());
    int n = src.cols * src.channels();
    for (int y = 0; y < src.rows; y++)
        for (int i = 0; i < n; i++) {
            if (src.depth() == CV_32F) out.ptr<float>(y)[i] = (float)::exp((double)src.ptr<float>(y)[i]);
            else out.ptr<double>(y)[i] = ::exp(src.ptr<double>(y)[i]);
        }
    deliver(dst, out);
}
// cv::sum: per channel, double accumulation in element order (SURVEY B-14)
Scalar sum(const Mat& m) {
    int cn = m.channels(), d = m.depth(); size_t e = m.elemSize1();
    need(cn <= 4, "sum: more than 4 channels");
    Scalar s;
    for (int y = 0; y < m.rows; y++)
        for (int x = 0; x < m.cols; x++)
            for (int c = 0; c < cn; c++) s[c] += ld(m.ptr(y) + ((size_t)x * cn + c) * e, d);
    return s;
}
Scalar mean(const Mat& m) {
    Scalar s = sum(m);
    double n = (double)m.total();
    return n > 0 ? s * (1. / n) : Scalar();
}
void minMaxLoc(const Mat& m, double* minVal, double* maxVal, Point* minLoc, Point* maxLoc) {
    need(!m.empty(), "minMaxLoc: empty matrix");
    int d = m.depth(), n = m.channels(); size_t e = m.elemSize1();
    need(n == 1 || (!minLoc && !maxLoc), "minMaxLoc: locations need a single channel");
    double mn = ld(m.ptr(0), d), mx = mn; Point pmn(0, 0), pmx(0, 0);
    for (int y = 0; y < m.rows; y++)
        for (int x = 0; x < m.cols * n; x++) {
            double v = ld(m.ptr(y) + x * e, d);
            if (v < mn) { mn = v; pmn = Point(x, y); }
            if (v > mx) { mx = v; pmx = Point(x, y); }
        }
    if (minVal) *minVal = mn; if (maxVal) *maxVal = mx; if (minLoc) *minLoc = pmn; if (maxLoc) *maxLoc = pmx;
}
// cv::normalize(NORM_MINMAX): one min / max over all channels; for a 32F result the scale is rounded to float first and
// the shift is dmin - (float)(smin * scale) (SURVEY B-10: bit-exact against cv2 4.13 for u8 and f32 inputs)
void normalize(const Mat& src_, Mat& dst, double a, double b, int norm_type, int dtype) {
    Mat src(src_);
    if (norm_type != NORM_MINMAX) unsupported("normalize: only NORM_MINMAX");
    double smin = 0, smax = 0;
    Mat flat = src.isContinuous() ? src.reshape(1) : src.clone().reshape(1);
    minMaxLoc(flat, &smin, &smax);
    double dmin = std::min(a, b), dmax = std::max(a, b);
    int rtype = dtype < 0 ? src.depth() : CV_MAT_DEPTH(dtype);
    double scale = (dmax - dmin) * (smax - smin > DBL_EPSILON ? 1. / (smax - smin) : 0);
    double shift;
    if (rtype == CV_32F) { scale = (float)scale; shift = (float)dmin - (float)(smin * scale); }
    else shift = dmin - smin * scale;
    src.convertTo(dst, rtype, scale, shift);
}

// =====================================================================================================
// borders, colour, filters
// =====================================================================================================
int borderInterpolate(int p, int len, int borderType) {
    if ((unsigned)p < (unsigned)len) return p;
    switch (borderType & ~BORDER_ISOLATED) {
    case BORDER_REPLICATE: return p < 0 ? 0 : len - 1;
    case BORDER_REFLECT: case BORDER_REFLECT_101: {
        int delta = (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101;
        if (len == 1) return 0;
        do { if (p < 0) p = -p - 1 + delta; else p = len - 1 - (p - len) - delta; } while ((unsigned)p >= (unsigned)len);
        return p;
    }
    case BORDER_WRAP: if (p < 0) p -= ((p - len + 1) / len) * len; if (p >= len) p %= len; return p;
    case BORDER_CONSTANT: return -1;
    }
    throw Exception(Error::StsBadArg, "unknown border type");
}
void copyMakeBorder(const Mat& src_, Mat& dst, int top, int bottom, int left, int right, int borderType, const Scalar& value) {
    Mat src(src_);
    need(top >= 0 && bottom >= 0 && left >= 0 && right >= 0, "copyMakeBorder: negative border");
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
    size_t es = src.elemSize(), e1 = src.elemSize1();
    for (int y = 0; y < out.rows; y++) {
        int sy = borderInterpolate(y - top, src.rows, borderType);
        for (int x = 0; x < out.cols; x++) {
            int sx = borderInterpolate(x - left, src.cols, borderType);
            if (sy < 0 || sx < 0) { for (int c = 0; c < src.channels(); c++) st<double>(out.ptr(y) + x * es + c * e1, src.depth(), value[c < 4 ? c : 0]); }
            else memcpy(out.ptr(y) + x * es, src.ptr(sy) + sx * es, es);
        }
    }
    deliver(dst, out);
}
// cvtColor BGR2GRAY / RGB2GRAY.  8U: 15-bit fixed point of OpenCV >= 4.2, (B 3735 + G 19235 + R 9798 + 2^14) >> 15
// (SURVEY B-11); 32F: 0.114 B + 0.587 G + 0.299 R in float
void cvtColor(const Mat& src_, Mat& dst, int code, int) {
    Mat src(src_);
    if (code != COLOR_BGR2GRAY && code != COLOR_RGB2GRAY) unsupported("cvtColor: only BGR2GRAY / RGB2GRAY");
    if (src.channels() != 3 && src.channels() != 4) throw Exception(Error::StsAssert, "cvtColor: invalid number of channels in input image (scn must be 3 or 4)");
    int cn = src.channels(), bi = code == COLOR_BGR2GRAY ? 0 : 2, ri = 2 - bi;
    Mat out(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
    for (int y = 0; y < src.rows; y++)
        for (int x = 0; x < src.cols; x++) {
            if (src.depth() == CV_8U) {
                const uchar* p = src.ptr(y) + (size_t)x * cn;
                out.ptr(y)[x] = (uchar)((3735 * p[bi] + 19235 * p[1] + 9798 * p[ri] + (1 << 14)) >> 15);
            } else if (src.depth() == CV_32F) {
                const float* p = src.ptr<float>(y) + (size_t)x * cn;
                out.ptr<float>(y)[x] = p[bi] * 0.114f + p[1] * 0.587f + p[ri] * 0.299f;
            } else unsupported("cvtColor: depth");
        }
    deliver(dst, out);
}
// filter2D: correlation (no kernel flip), anchor at the centre, border extrapolated per borderType (default REFLECT_101).
// Sums in double over the kernel in row-major order; the reference only uses small integer kernels on 8U input, for which
// every order is exact (SURVEY B-7).
void filter2D(const Mat& src_, Mat& dst, int ddepth, const Mat& kernel_, Point anchor, double delta, int borderType) {
    Mat src(src_), kernel(kernel_);
    need(kernel.channels() == 1, "filter2D: single-channel kernel");
    int dd = ddepth < 0 ? src.depth() : CV_MAT_DEPTH(ddepth);
    int ax = anchor.x < 0 ? kernel.cols / 2 : anchor.x, ay = anchor.y < 0 ? kernel.rows / 2 : anchor.y;
    int cn = src.channels(); size_t se = src.elemSize1();
    Mat out(src.rows, src.cols, CV_MAKETYPE(dd, cn));
    size_t de = out.elemSize1();
    for (int y = 0; y < src.rows; y++)
        for (int x = 0; x < src.cols; x++)
            for (int c = 0; c < cn; c++) {
                double acc = delta;
                for (int ky = 0; ky < kernel.rows; ky++) {
                    int sy = borderInterpolate(y + ky - ay, src.rows, borderType);
                    for (int kx = 0; kx < kernel.cols; kx++) {
                        double kv = ld(kernel.ptr(ky) + kx * kernel.elemSize1(), kernel.depth());
                        if (kv == 0) continue;
                        int sx = borderInterpolate(x + kx - ax, src.cols, borderType);
                        acc += kv * ld(src.ptr(sy) + ((size_t)sx * cn + c) * se, src.depth());
                    }
                }
                st<double>(out.ptr(y) + ((size_t)x * cn + c) * de, dd, acc);
            }
    deliver(dst, out);
}
// boxFilter (imgproc/src/box_filter.simd.hpp) for 32F / 64F / 8U sources with a 32F result, any channel count:
// RowSum<T, double> (direct sums for ksize 3 and 5, otherwise the running sum s += new - old per channel) followed by
// ColumnSum<double, float> (SUM += incoming row; out = (float)(SUM * scale); SUM -= outgoing row) in one stripe from the
// top, i.e. OpenCV's own summation order with one thread (SURVEY B-9)
void boxFilter(const Mat& src_, Mat& dst, int ddepth, Size ksize, Point anchor, bool normalize, int borderType) {
    Mat src(src_);
    int sd = src.depth(), dd = ddepth < 0 ? sd : CV_MAT_DEPTH(ddepth);
    if (dd != CV_32F || (sd != CV_32F && sd != CV_8U)) unsupported("boxFilter: only 8U / 32F -> 32F");
    int kw = ksize.width, kh = ksize.height, cn = src.channels();
    int ax = anchor.x < 0 ? kw / 2 : anchor.x, ay = anchor.y < 0 ? kh / 2 : anchor.y;
    int H = src.rows, W = src.cols, Wn = W * cn;
    double scale = normalize ? 1. / ((double)kw * kh) : 1;
    std::vector<double> rs((size_t)H * Wn);
    std::vector<double> S((size_t)(W + kw - 1) * cn);
    size_t se = src.elemSize1();
    for (int y = 0; y < H; y++) {
        for (int x = 0; x < W + kw - 1; x++) {
            int sx = borderInterpolate(x - ax, W, borderType);
            for (int c = 0; c < cn; c++) S[(size_t)x * cn + c] = ld(src.ptr(y) + ((size_t)sx * cn + c) * se, sd);
        }
        double* D = &rs[(size_t)y * Wn];
        int ksz_cn = kw * cn;
        if (kw == 3) { for (int i = 0; i < Wn; i++) D[i] = S[i] + S[i + cn] + S[i + 2 * cn]; }
        else if (kw == 5) { for (int i = 0; i < Wn; i++) D[i] = S[i] + S[i + cn] + S[i + 2 * cn] + S[i + 3 * cn] + S[i + 4 * cn]; }
        else {
            for (int c = 0; c < cn; c++) {
                double s = 0;
                for (int i = 0; i < ksz_cn; i += cn) s += S[c + i];
                D[c] = s;
                for (int i = 0; i < Wn - cn; i += cn) { s += S[c + i + ksz_cn] - S[c + i]; D[c + i + cn] = s; }
            }
        }
    }
    Mat out(H, W, CV_MAKETYPE(CV_32F, cn));
    std::vector<double> SUM((size_t)Wn, 0.0);
    for (int i = 0; i < kh - 1; i++) {
        const double* Sp = &rs[(size_t)borderInterpolate(i - ay, H, borderType) * Wn];
        for (int j = 0; j < Wn; j++) SUM[j] += Sp[j];
    }
    for (int y = 0; y < H; y++) {
        const double* Sp = &rs[(size_t)borderInterpolate(y - ay + kh - 1, H, borderType) * Wn];
        const double* Sm = &rs[(size_t)borderInterpolate(y - ay, H, borderType) * Wn];
        float* o = out.ptr<float>(y);
        for (int j = 0; j < Wn; j++) {
            double s0 = SUM[j] + Sp[j];
            o[j] = (float)(normalize ? s0 * scale : s0);
            SUM[j] = s0 - Sm[j];
        }
    }
    deliver(dst, out);
}

// =====================================================================================================
// MatExpr (core/src/matop.cpp)
// =====================================================================================================
void MatExpr::assign(Mat& m, int type) const {
    need(type == -1, "MatExpr::assign with an explicit type");
    switch (op) {
    case OP_IDENTITY: m = a; return;                                   // MatOp_Identity::assign (same type: header copy)
    case OP_ADDEX:                                                     // MatOp_AddEx::assign
        if (b.data) {
            if (s == Scalar() || !s.isReal()) {
                if (alpha == 1) {
                    if (beta == 1) cv::add(a, b, m);
                    else if (beta == -1) cv::subtract(a, b, m);
                    else cv::scaleAdd(b, beta, a, m);
                } else if (beta == 1) {
                    if (alpha == -1) cv::subtract(b, a, m);
                    else cv::scaleAdd(a, alpha, b, m);
                } else cv::addWeighted(a, alpha, b, beta, 0, m);
                if (!s.isReal()) cv::add(m, s, m);
            } else cv::addWeighted(a, alpha, b, beta, s[0], m);
        } else if (s.isReal() && fabs(alpha) != 1) {
            a.convertTo(m, a.type(), alpha, s[0]);
        } else if (alpha == 1) cv::add(a, s, m);
        else if (alpha == -1) cv::subtract(s, a, m);
        else { a.convertTo(m, a.type(), alpha); cv::add(m, s, m); }
        return;
    case OP_BIN:                                                       // MatOp_Bin::assign
        if (flags == '*') cv::multiply(a, b, m, alpha);
        else if (flags == '/' && b.data) cv::divide(a, b, m, alpha);
        else if (flags == '/' && !b.data) cv::divide(alpha, a, m);
        else if (flags == 'a' && b.data) cv::absdiff(a, b, m);
        else if (flags == 'a' && !b.data) cv::absdiff(a, s, m);
        else unsupported("MatOp_Bin flag");
        return;
    case OP_CMP:                                                       // MatOp_Cmp::assign
        if (b.data) cv::compare(a, b, m, flags); else cv::compare(a, alpha, m, flags);
        return;
    }
}

MatExpr Mat::mul(const Mat& m, double scale) const { MatExpr e; matop::makeBin(e, '*', *this, m, scale); return e; }
MatExpr Mat::mul(const MatExpr& me, double scale) const { MatExpr e; matop::multiply(me.op, MatExpr(*this), me, e, scale); return e; }
MatExpr MatExpr::mul(const MatExpr& e, double scale) const { MatExpr en; matop::multiply(op, *this, e, en, scale); return en; }
MatExpr MatExpr::mul(const Mat& m, double scale) const { MatExpr en; matop::multiply(op, *this, MatExpr(m), en, scale); return en; }

namespace matop {
static inline Mat materialise(const MatExpr& e) { Mat m; e.assign(m); return m; }

// MatOp::add (no subclass on this path overrides the two-expression form)
void add(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res) {
    if (self == e2.op) {
        double alpha = 1, beta = 1; Scalar s; Mat m1, m2;
        if (isAddEx(e1) && (!e1.b.data || e1.beta == 0)) { m1 = e1.a; alpha = e1.alpha; s = e1.s; } else m1 = materialise(e1);
        if (isAddEx(e2) && (!e2.b.data || e2.beta == 0)) { m2 = e2.a; beta = e2.alpha; s += e2.s; } else m2 = materialise(e2);
        makeAddEx(res, m1, m2, alpha, beta, s);
    } else add(e2.op, e1, e2, res);
}
void add(MatExpr::Op self, const MatExpr& e, const Scalar& s, MatExpr& res) {
    if (self == MatExpr::OP_ADDEX) { res = e; res.s += s; return; }              // MatOp_AddEx::add
    makeAddEx(res, materialise(e), Mat(), 1, 0, s);                              // MatOp::add
}
void subtract(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res) {
    if (self == e2.op) {
        double alpha = 1, beta = -1; Scalar s; Mat m1, m2;
        if (isAddEx(e1) && (!e1.b.data || e1.beta == 0)) { m1 = e1.a; alpha = e1.alpha; s = e1.s; } else m1 = materialise(e1);
        if (isAddEx(e2) && (!e2.b.data || e2.beta == 0)) { m2 = e2.a; beta = -e2.alpha; s -= e2.s; } else m2 = materialise(e2);
        makeAddEx(res, m1, m2, alpha, beta, s);
    } else subtract(e2.op, e1, e2, res);
}
void subtract(MatExpr::Op self, const Scalar& s, const MatExpr& e, MatExpr& res) {
    if (self == MatExpr::OP_ADDEX) { res = e; res.alpha = -res.alpha; res.beta = -res.beta; res.s = s - res.s; return; }   // MatOp_AddEx::subtract
    makeAddEx(res, materialise(e), Mat(), -1, 0, s);
}
void multiply(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res, double scale) {
    if (self == e2.op) {
        Mat m1, m2;
        if (isReciprocal(e1)) {
            if (isScaled(e2)) { scale *= e2.alpha; m2 = e2.a; } else m2 = materialise(e2);
            makeBin(res, '/', m2, e1.a, scale / e1.alpha);
        } else {
            char op = '*';
            if (isScaled(e1)) { m1 = e1.a; scale *= e1.alpha; } else m1 = materialise(e1);
            if (isScaled(e2)) { m2 = e2.a; scale *= e2.alpha; }
            else if (isReciprocal(e2)) { op = '/'; m2 = e2.a; scale /= e2.alpha; }
            else m2 = materialise(e2);
            makeBin(res, op, m1, m2, scale);
        }
    } else multiply(e2.op, e1, e2, res, scale);
}
void multiply(MatExpr::Op self, const MatExpr& e, double s, MatExpr& res) {
    if (self == MatExpr::OP_ADDEX) { res = e; res.alpha *= s; res.beta *= s; res.s *= s; return; }       // MatOp_AddEx::multiply
    if (self == MatExpr::OP_BIN && (e.flags == '*' || e.flags == '/')) { res = e; res.alpha *= s; return; }  // MatOp_Bin::multiply
    makeAddEx(res, materialise(e), Mat(), s, 0);                                                         // MatOp::multiply
}
void divide(MatExpr::Op self, const MatExpr& e1, const MatExpr& e2, MatExpr& res, double scale) {
    if (self == e2.op) {
        if (isReciprocal(e1) && isReciprocal(e2)) makeBin(res, '/', e2.a, e1.a, e1.alpha / e2.alpha);
        else {
            Mat m1, m2; char op = '/';
            if (isScaled(e1)) { m1 = e1.a; scale *= e1.alpha; } else m1 = materialise(e1);
            if (isScaled(e2)) { m2 = e2.a; scale /= e2.alpha; }
            else if (isReciprocal(e2)) { m2 = e2.a; scale /= e2.alpha; op = '*'; }
            else m2 = materialise(e2);
            makeBin(res, op, m1, m2, scale);
        }
    } else divide(e2.op, e1, e2, res, scale);
}
void divide(MatExpr::Op self, double s, const MatExpr& e, MatExpr& res) {
    if (self == MatExpr::OP_ADDEX && isScaled(e)) { makeBin(res, '/', e.a, Mat(), s / e.alpha); return; }              // MatOp_AddEx::divide
    if (self == MatExpr::OP_BIN && e.flags == '/' && (!e.b.data || e.beta == 0)) { makeAddEx(res, e.a, Mat(), s / e.alpha, 0); return; }   // MatOp_Bin::divide
    makeBin(res, '/', materialise(e), Mat(), s);
}
void abs(MatExpr::Op self, const MatExpr& e, MatExpr& res) {
    if (self == MatExpr::OP_ADDEX) {                                             // MatOp_AddEx::abs
        if ((!e.b.data || e.beta == 0) && fabs(e.alpha) == 1) { makeBin(res, 'a', e.a, -e.s * e.alpha); return; }
        if (e.b.data && e.alpha + e.beta == 0 && e.alpha * e.beta == -1) { makeBin(res, 'a', e.a, e.b); return; }
    }
    makeBin(res, 'a', materialise(e), Scalar());
}
}  // namespace matop
}  // namespace cv
