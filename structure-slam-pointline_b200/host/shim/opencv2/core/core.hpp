// COMPILE-CHECK SHIM ONLY — a few dozen lines of the OpenCV API surface the adapters touch, so that
// host/ORBextractor.cc and host/ExtractLineSegment_b200.cc can be syntax/type-checked in a container without
// OpenCV (tests/test_abi_cpu.py).  It is never linked into the product; real builds use the real OpenCV headers.
#pragma once
#include <cstring>
#include <vector>
#include <cstdlib>
#define CV_8U 0
#define CV_8UC1 0
typedef unsigned char uchar;
namespace cv {
template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f; typedef Point_<int> Point;
struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Size { int width, height; };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct Mat {
    int rows, cols; size_t step; uchar* data; bool owner;
    Mat() : rows(0), cols(0), step(0), data(0), owner(false) {}
    Mat(int r, int c, int) : rows(r), cols(c), step(c), data((uchar*)malloc((size_t)r * c + 1)), owner(true) {}
    int type() const { return CV_8UC1; }
    bool empty() const { return rows == 0 || cols == 0; }
    void create(int r, int c, int) { rows = r; cols = c; step = c; data = (uchar*)malloc((size_t)r * c + 1); owner = true; }
    void release() { rows = cols = 0; }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <class T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <class T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    template <class T> T& at(int i) { return ((T*)data)[i]; }
    template <class T> const T& at(int i) const { return ((const T*)data)[i]; }
    template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    Mat row(int i) const { Mat m; m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)i * step; return m; }
    bool isContinuous() const { return step == (size_t)cols; }
    Mat clone() const { return *this; }
    Mat operator()(const Rect& r) const { Mat m; m.rows = r.height; m.cols = r.width; m.step = step; m.data = data + (size_t)r.y * step + r.x; return m; }
};
Mat operator*(const Mat& a, const Mat& b); Mat operator+(const Mat& a, const Mat& b);   // declared only (type check)
struct _InputArray { const Mat* m; _InputArray(const Mat& a) : m(&a) {} _InputArray() : m(0) {} bool empty() const { return !m || m->empty(); } Mat getMat() const { return *m; } };
struct _OutputArray { Mat* m; _OutputArray(Mat& a) : m(&a) {} void create(int r, int c, int t) const { m->create(r, c, t); } void release() const { m->release(); } Mat getMat() const { return *m; } };
typedef const _InputArray& InputArray; typedef const _OutputArray& OutputArray;
}
