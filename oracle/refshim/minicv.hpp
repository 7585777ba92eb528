// oracle/refshim/minicv.hpp — TEST INFRASTRUCTURE ONLY.
//
// A small, FUNCTIONAL stand-in for the slice of the OpenCV 3.4 C++ API that the reference's hot-path sources touch
// (src/ORBextractor.cc, src/ORBmatcher.cc, src/LSDmatcher.cpp, src/ExtractLineSegment.cpp, src/Frame.cc,
// Thirdparty/DBoW2).  OpenCV's C++ headers and libraries are not installed in this image, so oracle/ref_build.sh
// compiles those reference files UNMODIFIED (read in place from /root/reference) against these headers into
// oracle/_ref/libref.so — the reference's own control flow running over primitives that are pinned to cv2 4.13:
//   * FAST / resize / copyMakeBorder / GaussianBlur / fastAtan2 / BFMatcher forward to the oracle's primitives
//     (oracle/orb_oracle.cpp, oracle/match_oracle.cpp), each of which tests/test_oracle_cpu.py pins bit-exactly to cv2;
//   * Mat arithmetic follows cv::gemm as probed on cv2 4.13 (tools/probe_cv_gemm.py): CV_32F products whose inner
//     length is 2..4 and equals one of the output dimensions run in float, left to right, the addend last; everything
//     else accumulates in double and rounds once.
// Nothing here is part of the product, and no reference source is copied: this file is written against the public
// OpenCV API documentation, not against OpenCV's sources.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6
#define CV_RGBA2GRAY 11
#define CV_BGRA2GRAY 10
#define CV_Assert(x) assert(x)

static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

using std::string;
typedef std::string String;

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}      // like OpenCV's saturate_cast for same-kind types
    T dot(const Point_& o) const { return x * o.x + y * o.y; }
};
template <> template <> inline Point_<int>::Point_(const Point_<float>& o) : x(cvRound(o.x)), y(cvRound(o.y)) {}
template <> template <> inline Point_<int>::Point_(const Point_<double>& o) : x(cvRound(o.x)), y(cvRound(o.y)) {}
template <class T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <class T> bool operator!=(const Point_<T>& a, const Point_<T>& b) { return !(a == b); }
// Point_<float> *= float : OpenCV narrows (float)(x*b) — a float*float product here
static inline Point_<float>& operator*=(Point_<float>& a, float b) { a.x = a.x * b; a.y = a.y * b; return a; }
static inline Point_<float>& operator*=(Point_<float>& a, double b) { a.x = (float)(a.x * b); a.y = (float)(a.y * b); return a; }
static inline Point_<float>& operator*=(Point_<float>& a, int b) { a.x = a.x * b; a.y = a.y * b; return a; }
template <class T> Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <class T> Point_<T> operator*(double b, const Point_<T>& a) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    template <class U> Point3_(const Point3_<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    T dot(const Point3_& o) const { return x * o.x + y * o.y + z * o.z; }
    Point3_ cross(const Point3_& o) const { return Point3_(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
};
template <class T> Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> Point3_<T> operator*(const Point3_<T>& a, double b) { return Point3_<T>((T)(a.x * b), (T)(a.y * b), (T)(a.z * b)); }
template <class T> Point3_<T> operator*(double b, const Point3_<T>& a) { return a * b; }
typedef Point3_<int> Point3i;
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;
typedef Size_<int> Size2i;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T a, T b, T c, T d) : x(a), y(b), width(c), height(d) {}
};
typedef Rect_<int> Rect;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double operator[](int i) const { return val[i]; }
};

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int a, int b) : start(a), end(b) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4,
       BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_BGRA2GRAY = 10, COLOR_RGBA2GRAY = 11 };

static inline int depth_size(int depth) { static const int s[7] = {1, 1, 2, 2, 4, 4, 8}; return s[depth & 7]; }

class Mat;
struct MatMul;
struct MatT;
// Mat::zeros / ones / eye are EXPRESSIONS in OpenCV: assigning one to an existing Mat of the same size and type fills it in
// place (Mat::create is a no-op then) — ORBextractor.cc:1037 relies on that to write descriptors into a rowRange() view.
struct MatInit { int rows, cols, type, kind; };

class Mat {
public:
    int flags;                 // type
    int rows, cols;
    size_t step;               // bytes per row
    uchar* data;
    std::shared_ptr<uchar> buf;  // owner (null for user data)

    Mat() : flags(0), rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& v) : Mat() { create(r, c, type); setTo(v); }
    Mat(Size s, int type, const Scalar& v) : Mat() { create(s.height, s.width, type); setTo(v); }
    Mat(int r, int c, int type, void* d, size_t st = 0) : flags(type), rows(r), cols(c), step(st ? st : (size_t)c * esz(type)), data((uchar*)d) {}
    Mat(const MatMul& e);
    Mat& operator=(const MatMul& e);
    Mat(const MatInit& e) : Mat() { *this = e; }
    Mat(const struct MatNeg& e);
    Mat& operator=(const MatInit& e) {
        create(e.rows, e.cols, e.type);
        for (int r = 0; r < rows; r++) memset(ptr(r), 0, (size_t)cols * elemSize());
        if (e.kind == 1) setTo(Scalar::all(1));
        if (e.kind == 2) for (int i = 0; i < std::min(rows, cols); i++) setd(i, i, 1);
        return *this;
    }
    template <class T> explicit Mat(const std::vector<T>& v);

    static size_t esz(int type) { return (size_t)depth_size(type & 7) * (size_t)(((type >> 3) & 63) + 1); }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags) return;
        flags = type; rows = r; cols = c; step = (size_t)c * esz(type);
        size_t n = (size_t)r * step;
        buf.reset((uchar*)calloc(n + 64, 1), free);
        data = buf.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return ((flags >> 3) & 63) + 1; }
    size_t elemSize() const { return esz(flags); }
    size_t elemSize1() const { return depth_size(flags & 7); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }

    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    // single index: element i of a row or column vector (OpenCV semantics)
    template <class T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : (cols == 1 ? *(T*)(data + (size_t)i * step) : at<T>(i / cols, i % cols)); }
    template <class T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    template <class T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <class T> const T& at(Point p) const { return at<T>(p.y, p.x); }

    Mat sub(int r0, int r1, int c0, int c1) const {
        Mat m; m.flags = flags; m.rows = r1 - r0; m.cols = c1 - c0; m.step = step; m.buf = buf;
        m.data = data + (size_t)r0 * step + (size_t)c0 * elemSize(); return m;
    }
    Mat rowRange(int a, int b) const { return sub(a, b, 0, cols); }
    Mat colRange(int a, int b) const { return sub(0, rows, a, b); }
    Mat rowRange(const Range& r) const { return r.start == INT32_MIN ? *this : rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return r.start == INT32_MIN ? *this : colRange(r.start, r.end); }
    Mat row(int i) const { return sub(i, i + 1, 0, cols); }
    Mat col(int i) const { return sub(0, rows, i, i + 1); }
    Mat operator()(const Rect& r) const { return sub(r.y, r.y + r.height, r.x, r.x + r.width); }
    Mat operator()(const Range& rr, const Range& cr) const { return rowRange(rr).colRange(cr); }

    void copyTo(Mat& dst) const {
        if (empty()) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        if (!(dst.rows == rows && dst.cols == cols && dst.flags == flags && dst.data)) dst.create(rows, cols, flags);
        size_t n = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) memmove(dst.ptr(r), ptr(r), n);
    }
    void copyTo(Mat&& dst) const { Mat d = dst; copyTo(d); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    Mat& setTo(const Scalar& s);
    Mat& operator=(const Scalar& s) { return setTo(s); }
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat reshape(int cn, int newrows = 0) const {          // header only, like cv::Mat::reshape (continuous data when the row count changes)
        Mat m = *this; const int oldcn = channels(); if (cn == 0) cn = oldcn;
        const size_t scalars = (size_t)cols * oldcn;
        m.flags = CV_MAKETYPE(depth(), cn); m.cols = (int)(scalars / cn);
        if (newrows > 0 && newrows != rows) { assert(isContinuous()); m.rows = newrows; m.cols = (int)((size_t)rows * scalars / newrows / cn); m.step = (size_t)m.cols * m.elemSize(); }
        return m;
    }

    double getd(int r, int c) const;         // element as double (8U/16S/32S/32F/64F)
    void setd(int r, int c, double v);
    MatT t() const;             // an expression, as in OpenCV: A.t()*B is ONE gemm with a transpose flag
    Mat transposed() const;
    Mat(const MatT& e);
    Mat& operator=(const MatT& e);
    Mat inv(int method = 0) const;
    double dot(const Mat& m) const;
    Mat cross(const Mat& m) const;
    Mat mul(const Mat& m, double scale = 1) const;

    static MatInit zeros(int r, int c, int type) { return MatInit{r, c, type, 0}; }
    static MatInit zeros(Size s, int type) { return MatInit{s.height, s.width, type, 0}; }
    static MatInit ones(int r, int c, int type) { return MatInit{r, c, type, 1}; }
    static MatInit eye(int r, int c, int type) { return MatInit{r, c, type, 2}; }
};

template <class T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<short> { enum { type = CV_16S }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

template <class T> Mat::Mat(const std::vector<T>& v) : Mat() {
    create((int)v.size(), 1, DataType<T>::type);
    for (size_t i = 0; i < v.size(); i++) at<T>((int)i) = v[i];
}

template <class T> class Mat_;
template <class T> struct MatCommaInitializer_ {
    Mat m; int i;
    MatCommaInitializer_(const Mat& mm) : m(mm), i(0) {}
    template <class U> MatCommaInitializer_& operator,(U v) { m.at<T>(i / m.cols, i % m.cols) = (T)v; i++; return *this; }
    operator Mat() const { return m; }
    operator Mat_<T>() const;
};
template <class T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    Mat_(const Mat& m) : Mat(m) { assert(m.empty() || m.type() == DataType<T>::type); }
    T& operator()(int r, int c) { return at<T>(r, c); }
    const T& operator()(int r, int c) const { return at<T>(r, c); }
    T& operator()(int i) { return at<T>(i); }
};
template <class T> MatCommaInitializer_<T>::operator Mat_<T>() const { return Mat_<T>(m); }
template <class T, class U> MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
    MatCommaInitializer_<T> ci(m); ci, v; return ci;
}

// ---- arithmetic (eager, except A*B which stays a product expression so that A*B+C is ONE gemm like cv::MatExpr) ----
struct MatT {                     // alpha * A^T
    Mat a; double alpha;
    MatT(const Mat& m, double s = 1);
    Mat eval() const;
    MatT t() const;               // (A^T)^T
    template <class T> T at(int r, int c) const { return eval().at<T>(r, c); }
};
struct MatMul {
    Mat a, b; double alpha; int flags;          // flags: 1 = a transposed, 2 = b transposed (cv::GEMM_1_T / GEMM_2_T)
    MatMul(const Mat& x, const Mat& y, double s = 1, int f = 0) : a(x), b(y), alpha(s), flags(f) {}
    Mat eval(const Mat* c = nullptr, double beta = 0) const;
    MatT t() const;
    Mat row(int i) const { return eval().row(i); }
    Mat col(int i) const { return eval().col(i); }
    template <class T> T at(int i) const { return eval().at<T>(i); }
    template <class T> T at(int r, int c) const { return eval().at<T>(r, c); }
    double dot(const Mat& m) const { return eval().dot(m); }
};
inline Mat::Mat(const MatMul& e) : Mat() { *this = e.eval(); }
inline Mat& Mat::operator=(const MatMul& e) { Mat r = e.eval(); return *this = r; }
struct MatNeg { Mat a; };            // -A : folds into the following product's alpha, like cv::MatExpr
inline MatNeg operator-(const Mat& a) { return MatNeg{a}; }
Mat negate(const Mat& a);
Mat operator*(const Mat& a, double s);

Mat gemm_eval(const Mat& a, const Mat& b, double alpha, const Mat* c, double beta, int flags = 0);
void gemm(const Mat& a, const Mat& b, double alpha, const Mat& c, double beta, Mat& dst, int flags = 0);
inline Mat MatMul::eval(const Mat* c, double beta) const { return gemm_eval(a, b, alpha, c, beta, flags); }
inline Mat::Mat(const MatNeg& e) : Mat() { *this = negate(e.a); }
inline MatT::MatT(const Mat& m, double s) : a(m), alpha(s) {}
inline Mat MatT::eval() const { Mat m = a.transposed(); return alpha == 1 ? m : Mat(m * alpha); }
inline MatT MatT::t() const { return MatT(a.transposed(), alpha); }
inline MatT Mat::t() const { return MatT(*this); }
inline MatT MatMul::t() const { return MatT(eval()); }
inline Mat::Mat(const MatT& e) : Mat() { *this = e.eval(); }
inline Mat& Mat::operator=(const MatT& e) { Mat r = e.eval(); return *this = r; }
inline MatT operator-(const MatT& a) { return MatT(a.a, -a.alpha); }
inline MatT operator*(double s, const MatT& a) { return MatT(a.a, a.alpha * s); }
inline MatT operator*(const MatT& a, double s) { return MatT(a.a, a.alpha * s); }
inline MatMul operator*(const MatT& a, const Mat& b) { return MatMul(a.a, b, a.alpha, 1); }
inline MatMul operator*(const Mat& a, const MatT& b) { return MatMul(a, b.a, b.alpha, 2); }
inline MatMul operator*(const MatT& a, const MatT& b) { return MatMul(a.a, b.a, a.alpha * b.alpha, 3); }

inline MatMul operator*(const Mat& a, const Mat& b) { return MatMul(a, b); }
Mat operator*(const Mat& a, double s);
inline MatMul operator*(const MatMul& a, const Mat& b) { return MatMul(a.eval(), b); }
inline MatMul operator*(const Mat& a, const MatMul& b) { return MatMul(a, b.eval()); }
inline MatMul operator*(const MatMul& a, double s) { return MatMul(a.a, a.b, a.alpha * s, a.flags); }
inline MatMul operator*(double s, const MatMul& a) { return MatMul(a.a, a.b, a.alpha * s, a.flags); }
inline MatMul operator-(const MatMul& a) { return MatMul(a.a, a.b, -a.alpha, a.flags); }
inline MatMul operator*(const MatMul& a, const MatT& b) { return MatMul(a.eval(), b.a, b.alpha, 2); }
inline MatMul operator*(const MatT& a, const MatMul& b) { return MatMul(a.a, b.eval(), a.alpha, 1); }
inline Mat operator+(const MatMul& a, const Mat& c) { return a.eval(&c, 1); }
inline Mat operator+(const Mat& c, const MatMul& a) { return a.eval(&c, 1); }
inline Mat operator-(const MatMul& a, const Mat& c) { return a.eval(&c, -1); }
Mat operator+(const Mat& a, const Mat& b);
Mat operator-(const Mat& a, const Mat& b);
inline Mat operator+(const MatMul& a, const MatMul& b) { return a.eval() + b.eval(); }
inline Mat operator-(const MatMul& a, const MatMul& b) { return a.eval() - b.eval(); }
inline Mat operator-(const Mat& a, const MatMul& b) { return a - b.eval(); }
inline MatMul operator*(const MatNeg& a, const Mat& b) { return MatMul(a.a, b, -1); }
inline Mat operator+(const MatNeg& a, const Mat& b) { return b - a.a; }
Mat operator*(const Mat& a, double s);
inline Mat operator*(double s, const Mat& a) { return a * s; }
Mat operator/(const Mat& a, double s);
inline Mat operator/(const MatMul& a, double s) { return a.eval() / s; }
Mat operator+(const Mat& a, const Scalar& s);
std::ostream& operator<<(std::ostream& os, const Mat& m);

double norm(const Mat& a, int normType = NORM_L2);
double norm(const Mat& a, const Mat& b, int normType = NORM_L2);
inline double norm(const MatMul& a, int normType = NORM_L2) { return norm(a.eval(), normType); }
template <class T> double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
template <class T> double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }

// ---- proxies for function arguments ----
class _InputArray {
public:
    const Mat* m; Mat own;
    _InputArray() : m(nullptr) {}
    _InputArray(const Mat& a) : m(&a) {}
    _InputArray(const MatMul& e) : own(e.eval()) { m = &own; }
    _InputArray(const MatT& e) : own(e.eval()) { m = &own; }
    template <class T> _InputArray(const std::vector<T>& v) : own(v) { m = &own; }
    bool empty() const { return !m || m->empty(); }
    Mat getMat(int = -1) const { return m ? *m : Mat(); }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& a) : m(&a) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void create(Size s, int t) const { m->create(s, t); }
    void release() const { m->release(); }
    Mat getMat(int = -1) const { return *m; }
    bool needed() const { return true; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline const _InputArray& noArray() { static _InputArray a; return a; }

template <class T> struct Ptr : public std::shared_ptr<T> {
    Ptr() {}
    Ptr(T* p) : std::shared_ptr<T>(p) {}
    Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
    bool empty() const { return !this->get(); }
    operator T*() const { return this->get(); }
};
template <class T, class... A> Ptr<T> makePtr(A&&... a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

float fastAtan2(float y, float x);

// ---- imgproc / features2d entry points used by the reference ----
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R = noArray(), InputArray P = noArray());
void meanStdDev(InputArray src, OutputArray mean, OutputArray stddev, InputArray mask = noArray());
void Laplacian(InputArray src, OutputArray dst, int ddepth, int ksize = 1, double scale = 1, double delta = 0, int borderType = BORDER_DEFAULT);
void convertScaleAbs(InputArray src, OutputArray dst, double alpha = 1, double beta = 0);

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

class DescriptorMatcher {
public:
    virtual ~DescriptorMatcher() {}
    virtual void knnMatch(InputArray query, InputArray train, std::vector<std::vector<DMatch> >& matches, int k,
                          InputArray mask = noArray(), bool compactResult = false) const = 0;
};
class BFMatcher : public DescriptorMatcher {
public:
    int normType; bool crossCheck;
    BFMatcher(int nt = NORM_L2, bool cc = false) : normType(nt), crossCheck(cc) {}
    void knnMatch(InputArray query, InputArray train, std::vector<std::vector<DMatch> >& matches, int k,
                  InputArray mask = noArray(), bool compactResult = false) const override;
};

class SVD {
public:
    Mat u, w, vt;
    enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    SVD() {}
    SVD(InputArray src, int flags = 0);
    static void compute(InputArray src, OutputArray w, OutputArray u, OutputArray vt, int flags = 0);
};

// FileStorage / FileNode: DBoW2's YAML save()/load() templates name them; the vocabulary is loaded from the text format
// (loadFromTextFile) in every tested path, so these are inert.
class FileNode {
public:
    operator float() const { return 0.f; }
    operator int() const { return 0; }
    operator double() const { return 0.; }
    operator std::string() const { return std::string(); }
    bool empty() const { return true; }
    size_t size() const { return 0; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    void release() {}
};
template <class T> FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
template <class T> void operator>>(const FileNode&, T&) {}

Mat imread(const std::string& name, int flags = 1);
enum { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
#define CV_LOAD_IMAGE_UNCHANGED -1

}  // namespace cv

// C-API leftovers some reference headers name
typedef cv::Point2f CvPoint2D32f;
