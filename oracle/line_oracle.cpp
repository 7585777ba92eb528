/*
 * oracle/line_oracle.cpp — CPU ORACLE for the line path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates LineSegment::ExtractLineSegment (/root/reference/src/ExtractLineSegment.cpp:18-69) and the
 * third-party code it delegates to, which is NOT in the reference tree:
 *   - cv::line_descriptor::LSDDetector::detect  (opencv_contrib 3.4 line_descriptor, LSDDetector.cpp)   [from memory]
 *   - cv::createLineSegmentDetector(LSD_REFINE_ADV) (OpenCV imgproc lsd.cpp)  — pinned against cv2 4.13 in tests/;
 *     rect_nfa follows the 4.13 binary (polygon scan with ceil/trunc limits, x86 double->int conversion)
 *   - cv::line_descriptor::BinaryDescriptor::compute (LBD, binary_descriptor.cpp)                       [from memory]
 * PARITY UNPINNED for KeyLine packaging and LBD: no runnable implementation of opencv_contrib exists in this
 * image and the reference ships no test vectors; for those parts this file IS the specification.
 * Canonical float rules as in orb_oracle.cpp: no FMA, cosf/sinf/atan2f := double evaluation narrowed to float.
 */
#include "oracle.h"
#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

extern "C" void orc_sepfilter_fixed_u8(const uint8_t*, int, int, int, uint8_t*, int, const int*, int);
extern "C" float orc_fast_atan2(float y, float x);

namespace {
typedef uint8_t uchar;
const double NOTDEF = -1024.0, M_3_2_PI = (3 * 3.14159265358979323846) / 2, M_2__PI = 2 * 3.14159265358979323846;
const double PI = 3.14159265358979323846, DEG_TO_RADS = 3.14159265358979323846 / 180;
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline int cvRoundD(double v) { return (int)lrint(v); }
inline float cosf_c(float a) { return (float)cos((double)a); }      // canonical cosf
inline float sinf_c(float a) { return (float)sin((double)a); }
inline int x86_d2i(double v) {                                      // cvttsd2si: out of range / NaN -> INT_MIN
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}

/* cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR_EXACT) on 8U — SURVEY.md A.6 (iii).  With fx, fy given, OpenCV maps
   destination to source coordinates with scale = 1 / fx (NOT src_size / dst_size; they differ whenever fx * size is not an
   integer), imgproc/resize.cpp.  For fx = 0.8 the scale is exactly 1.25, so the 8.8 weights are exact in double. */
void resize_linear_exact(const uchar* src, int sw, int sh, int sp, uchar* dst, int dw, int dh, int dp, double fx) {
    std::vector<int> xi(dw), xw(dw), yi(dh), yw(dh);
    auto tab = [fx](int dn, int sn, std::vector<int>& idx, std::vector<int>& w1) {
        const double sc = 1.0 / fx;
        for (int d = 0; d < dn; d++) {
            double s = (d + 0.5) * sc - 0.5;
            int i0 = (int)floor(s);
            double f = s - i0;
            if (i0 < 0) { i0 = 0; f = 0; }
            if (i0 >= sn - 1) { i0 = sn - 1; f = 0; }
            idx[d] = i0; w1[d] = (int)lrint(f * 256);
        }
    };
    tab(dw, sw, xi, xw); tab(dh, sh, yi, yw);
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; y++) {
        const uchar* S0 = src + (size_t)yi[y] * sp;
        const uchar* S1 = src + (size_t)std::min(yi[y] + 1, sh - 1) * sp;
        for (int x = 0; x < dw; x++) {
            int i0 = xi[x], i1 = std::min(i0 + 1, sw - 1), w1 = xw[x], w0 = 256 - w1;
            r0[x] = w0 * S0[i0] + w1 * S0[i1];
            r1[x] = w0 * S1[i0] + w1 * S1[i1];
        }
        const int v1 = yw[y], v0 = 256 - v1;
        for (int x = 0; x < dw; x++) dst[(size_t)y * dp + x] = (uchar)((v0 * r0[x] + v1 * r1[x] + 32768) >> 16);
    }
}

/* ---------------------------------------------------------------------------------------------
 * LSD (OpenCV imgproc lsd.cpp, LSD_REFINE_ADV, defaults quant 2.0, ang_th 22.5, log_eps 0, density_th 0.7,
 * n_bins 1024) on an image that is ALREADY at detection scale.  SURVEY.md A.6 "LSD internals".
 * --------------------------------------------------------------------------------------------- */
struct RegPt { int x, y; };
struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };

struct Lsd {
    int w, h;
    std::vector<double> angles, modgrad;
    std::vector<uchar> used;
    std::vector<int> order;      // pixel indices (y*w+x), bins descending, raster inside a bin
    double LOG_NT;
    std::vector<double> trace;   // per region that reached region2rect: seed, n0, n1, log_nfa, x1, y1, x2, y2, width, p

    void ll_angle(const uchar* img, int pitch, double threshold, int n_bins) {
        angles.assign((size_t)w * h, NOTDEF); modgrad.assign((size_t)w * h, 0.0);
        double max_grad = -1;
        for (int y = 0; y < h - 1; y++)
            for (int x = 0; x < w - 1; x++) {
                const uchar* p = img + (size_t)y * pitch + x;
                int DA = p[pitch + 1] - p[0], BC = p[1] - p[pitch];
                int gx = DA + BC, gy = DA - BC;
                double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
                modgrad[(size_t)y * w + x] = norm;
                if (norm > threshold) {
                    angles[(size_t)y * w + x] = orc_fast_atan2((float)gx, (float)-gy) * DEG_TO_RADS;
                    if (norm > max_grad) max_grad = norm;
                }
            }
        const double bin_coef = (max_grad > 0) ? double(n_bins - 1) / max_grad : 0;
        std::vector<int> cnt(n_bins + 1, 0);
        for (int y = 0; y < h - 1; y++)
            for (int x = 0; x < w - 1; x++) cnt[n_bins - 1 - int(modgrad[(size_t)y * w + x] * bin_coef)]++;
        int s = 0;
        for (int b = 0; b <= n_bins; b++) { int c = cnt[b]; cnt[b] = s; s += c; }
        order.resize(s);
        for (int y = 0; y < h - 1; y++)
            for (int x = 0; x < w - 1; x++) order[cnt[n_bins - 1 - int(modgrad[(size_t)y * w + x] * bin_coef)]++] = y * w + x;
    }
    bool isAligned(int x, int y, double theta, double prec) const {
        if (x < 0 || y < 0 || x >= w || y >= h) return false;
        const double a = angles[(size_t)y * w + x];
        if (a == NOTDEF) return false;
        double n_theta = theta - a;
        if (n_theta < 0) n_theta = -n_theta;
        if (n_theta > M_3_2_PI) { n_theta -= M_2__PI; if (n_theta < 0) n_theta = -n_theta; }
        return n_theta <= prec;
    }
    void region_grow(int sx, int sy, std::vector<RegPt>& reg, double& reg_angle, double prec) {
        reg.clear();
        reg.push_back(RegPt{sx, sy});
        reg_angle = angles[(size_t)sy * w + sx];
        float sumdx = float(std::cos(reg_angle)), sumdy = float(std::sin(reg_angle));
        used[(size_t)sy * w + sx] = 1;
        for (size_t i = 0; i < reg.size(); i++) {
            const RegPt rp = reg[i];
            const int xx_min = std::max(rp.x - 1, 0), xx_max = std::min(rp.x + 1, w - 1);
            const int yy_min = std::max(rp.y - 1, 0), yy_max = std::min(rp.y + 1, h - 1);
            for (int yy = yy_min; yy <= yy_max; yy++)
                for (int xx = xx_min; xx <= xx_max; xx++) {
                    uchar& u = used[(size_t)yy * w + xx];
                    if (u != 1 && isAligned(xx, yy, reg_angle, prec)) {
                        const double a = angles[(size_t)yy * w + xx];
                        u = 1;
                        reg.push_back(RegPt{xx, yy});
                        sumdx += cosf_c(float(a)); sumdy += sinf_c(float(a));
                        reg_angle = orc_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
                    }
                }
        }
    }
    static double angle_diff_signed(double a, double b) {
        double d = a - b;
        while (d <= -PI) d += M_2__PI;
        while (d > PI) d -= M_2__PI;
        return d;
    }
    static double angle_diff(double a, double b) { return std::fabs(angle_diff_signed(a, b)); }
    double get_theta(const std::vector<RegPt>& reg, double x, double y, double reg_angle, double prec) const {
        double Ixx = 0, Iyy = 0, Ixy = 0;
        for (const RegPt& r : reg) {
            const double wgt = modgrad[(size_t)r.y * w + r.x], dx = double(r.x) - x, dy = double(r.y) - y;
            Ixx += dy * dy * wgt; Iyy += dx * dx * wgt; Ixy -= dx * dy * wgt;
        }
        const double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? double(orc_fast_atan2(float(lambda - Ixx), float(Ixy)))
                                                        : double(orc_fast_atan2(float(Ixy), float(lambda - Iyy)));
        theta *= DEG_TO_RADS;
        if (angle_diff(theta, reg_angle) > prec) theta += PI;
        return theta;
    }
    void region2rect(const std::vector<RegPt>& reg, double reg_angle, double prec, double p, Rect& rec) const {
        double x = 0, y = 0, sum = 0;
        for (const RegPt& r : reg) { const double wgt = modgrad[(size_t)r.y * w + r.x]; x += double(r.x) * wgt; y += double(r.y) * wgt; sum += wgt; }
        x /= sum; y /= sum;
        const double theta = get_theta(reg, x, y, reg_angle, prec);
        const double dx = std::cos(theta), dy = std::sin(theta);
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
        for (const RegPt& r : reg) {
            const double regdx = double(r.x) - x, regdy = double(r.y) - y;
            const double l = regdx * dx + regdy * dy, ww = -regdx * dy + regdy * dx;
            if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
            if (ww > w_max) w_max = ww; else if (ww < w_min) w_min = ww;
        }
        rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy; rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min; rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }
    static double dist(double x1, double y1, double x2, double y2) { return std::sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1)); }
    static double distSq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
    bool reduce_region_radius(std::vector<RegPt>& reg, double reg_angle, double prec, double p, Rect& rec, double density, double density_th) {
        const double xc = double(reg[0].x), yc = double(reg[0].y);
        double radSq = std::max(distSq(xc, yc, rec.x1, rec.y1), distSq(xc, yc, rec.x2, rec.y2));
        while (density < density_th) {
            radSq *= 0.75 * 0.75;
            for (size_t i = 0; i < reg.size(); i++)
                if (distSq(xc, yc, double(reg[i].x), double(reg[i].y)) > radSq) {
                    used[(size_t)reg[i].y * w + reg[i].x] = 0;
                    std::swap(reg[i], reg[reg.size() - 1]);
                    reg.pop_back();
                    --i;
                }
            if (reg.size() < 2) return false;
            region2rect(reg, reg_angle, prec, p, rec);
            density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        }
        return true;
    }
    bool refine(std::vector<RegPt>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
        double density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= density_th) return true;
        const double xc = double(reg[0].x), yc = double(reg[0].y);
        const double ang_c = angles[(size_t)reg[0].y * w + reg[0].x];
        double sum = 0, s_sum = 0;
        int n = 0;
        for (const RegPt& r : reg) {
            used[(size_t)r.y * w + r.x] = 0;
            if (dist(xc, yc, double(r.x), double(r.y)) < rec.width) {
                const double ang_d = angle_diff_signed(angles[(size_t)r.y * w + r.x], ang_c);
                sum += ang_d; s_sum += ang_d * ang_d; ++n;
            }
        }
        const double mean_angle = sum / double(n);
        const double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
        const int sx = reg[0].x, sy = reg[0].y;
        region_grow(sx, sy, reg, reg_angle, tau);
        if (reg.size() < 2) return false;
        region2rect(reg, reg_angle, prec, p, rec);
        density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
        return true;
    }
    static double log_gamma_windschitl(double x) {
        return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
    }
    static double log_gamma_lanczos(double x) {
        static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
        double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5), b = 0;
        for (int n = 0; n < 7; ++n) { a -= std::log(x + double(n)); b += q[n] * std::pow(x, double(n)); }
        return a + std::log(b);
    }
    static double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }
    static bool double_equal(double a, double b) {
        if (a == b) return true;
        const double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
        double abs_max = (aa > bb) ? aa : bb;
        if (abs_max < 2.2250738585072014e-308) abs_max = 2.2250738585072014e-308;
        return (abs_diff / abs_max) <= (100.0 * 2.220446049250313e-16);
    }
    double nfa(int n, int k, double p) const {
        if (n == 0 || k == 0) return -LOG_NT;
        if (n == k) return -LOG_NT - double(n) * std::log10(p);
        const double p_term = p / (1 - p);
        const double log1term = log_gamma(double(n) + 1) - log_gamma(double(k) + 1) - log_gamma(double(n - k) + 1) +
                                double(k) * std::log(p) + double(n - k) * std::log(1.0 - p);
        double term = std::exp(log1term);
        if (double_equal(term, 0)) {
            if (k > n * p) return -log1term / 2.30258509299404568402 - LOG_NT;
            return -LOG_NT;
        }
        double bin_tail = term;
        const double tolerance = 0.1;
        for (int i = k + 1; i <= n; i++) {
            const double bin_term = double(n - i + 1) / double(i);
            const double mult_term = bin_term * p_term;
            term *= mult_term;
            bin_tail += term;
            if (bin_term < 1) {
                const double err = term * ((1 - std::pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
                if (err < tolerance * std::fabs(-std::log10(bin_tail) - LOG_NT) * bin_tail) break;
            }
        }
        return -std::log10(bin_tail) - LOG_NT;
    }
    /* rect_nfa as compiled into OpenCV 4.13 (decoded from the cv2 wheel's binary): the rectangle's four real-valued
       corners are walked from the min-y vertex; per row the span is [ceil(left limit), trunc(right limit)] with
       limits extrapolated along the current edge (so near-horizontal edges produce very wide rows), and
       double->int conversions follow x86 cvttsd2si (out of range -> INT_MIN). */
    double rect_nfa(const Rect& rec) const {
        const double half_width = 0.5 * rec.width, dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
        const double vx[4] = {rec.x1 - dyhw, rec.x2 - dyhw, rec.x2 + dyhw, rec.x1 + dyhw};
        const double vy[4] = {rec.y1 + dxhw, rec.y2 + dxhw, rec.y2 - dxhw, rec.y1 - dxhw};
        int off = 0;
        for (int i = 1; i < 4; i++) if (vy[i] < vy[off] || (vy[i] == vy[off] && vx[i] < vx[off])) off = i;
        const double Mx = vx[off], My = vy[off], Ax = vx[(off + 1) & 3], Ay = vy[(off + 1) & 3];
        const double Bx = vx[(off + 2) & 3], By = vy[(off + 2) & 3], Cx = vx[(off + 3) & 3], Cy = vy[(off + 3) & 3];
        const int cM = x86_d2i(std::ceil(My)), cA = x86_d2i(std::ceil(Ay)), cB = x86_d2i(std::ceil(By)), cC = x86_d2i(std::ceil(Cy));
        const double s1 = (cA != cM) ? (Ax - Mx) / (Ay - My) : 0.0;
        const double s2 = (cB != cA) ? (Bx - Ax) / (By - Ay) : 0.0;
        const double s3 = (cC != cM) ? (Cx - Mx) / (Cy - My) : 0.0;
        const double s4 = (cB != cC) ? (Bx - Cx) / (By - Cy) : 0.0;
        int total_pts = 0, alg_pts = 0;
        for (int y = cM; y <= cB; ++y) {
            if (y < 0 || y >= h) continue;
            const double xl = (cA < y) ? (double(y) - Ay) * s2 + Ax : (double(y) - My) * s1 + Mx;
            const double xr = (cC <= y) ? (double(y) - Cy) * s4 + Cx : (double(y) - My) * s3 + Mx;
            int xs = x86_d2i(std::ceil(xl));
            const int xe = x86_d2i(xr);
            if (xe < xs) continue;
            if (xs < 0) xs = 0;
            for (int x = xs; x <= xe && x < w; ++x) {
                ++total_pts;
                if (isAligned(x, y, rec.theta, rec.prec)) ++alg_pts;
            }
        }
        return nfa(total_pts, alg_pts, rec.p);
    }
    double rect_improve(Rect& rec) const {
        const double delta = 0.5, delta_2 = delta / 2.0, LOG_EPS = 0.0;
        double log_nfa = rect_nfa(rec);
        if (log_nfa > LOG_EPS) return log_nfa;
        Rect r = rec;
        for (int n = 0; n < 5; ++n) {
            r.p /= 2; r.prec = r.p * PI;
            const double log_nfa_new = rect_nfa(r);
            if (log_nfa_new > log_nfa) { log_nfa = log_nfa_new; rec = r; }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.width -= delta;
                const double log_nfa_new = rect_nfa(r);
                if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2; r.width -= delta;
                const double log_nfa_new = rect_nfa(r);
                if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2; r.width -= delta;
                const double log_nfa_new = rect_nfa(r);
                if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (int n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.p /= 2; r.prec = r.p * PI;
                const double log_nfa_new = rect_nfa(r);
                if (log_nfa_new > log_nfa) { rec = r; log_nfa = log_nfa_new; }
            }
        return log_nfa;
    }
    /* flsd: returns rectangles (x1,y1,x2,y2 at detection scale, before the +0.5 offset) */
    void detect(const uchar* img, int W, int H, int pitch, std::vector<Rect>& out) {
        w = W; h = H;
        const double ANG_TH = 22.5, QUANT = 2.0, DENSITY_TH = 0.7;
        const double prec = PI * ANG_TH / 180, p = ANG_TH / 180, rho = QUANT / std::sin(prec);
        ll_angle(img, pitch, rho, 1024);
        LOG_NT = 5 * (std::log10(double(w)) + std::log10(double(h))) / 2 + std::log10(11.0);
        const size_t min_reg_size = size_t(-LOG_NT / std::log10(p));
        used.assign((size_t)w * h, 0);
        std::vector<RegPt> reg;
        out.clear(); trace.clear();
        for (int idx : order) {
            if (used[idx] != 0 || angles[idx] == NOTDEF) continue;
            double reg_angle;
            region_grow(idx % w, idx / w, reg, reg_angle, prec);
            if (reg.size() < min_reg_size) continue;
            Rect rec;
            region2rect(reg, reg_angle, prec, p, rec);
            const double n0 = (double)reg.size();
            if (!refine(reg, reg_angle, prec, p, rec, DENSITY_TH)) { double t[10] = {(double)idx, n0, (double)reg.size(), -1e9, 0, 0, 0, 0, 0, 0}; trace.insert(trace.end(), t, t + 10); continue; }
            const double log_nfa = rect_improve(rec);
            { double t[10] = {(double)idx, n0, (double)reg.size(), log_nfa, rec.x1, rec.y1, rec.x2, rec.y2, rec.width, rec.p}; trace.insert(trace.end(), t, t + 10); }
            if (log_nfa <= 0.0) continue;
            out.push_back(rec);
        }
    }
};

/* Sobel 3x3 (ksize 3, no scaling) to CV_16S with BORDER_REFLECT_101 — SURVEY.md A.6 [probe] */
void sobel3(const uchar* img, int w, int h, int pitch, int16_t* dx, int16_t* dy) {
    auto R = [](int p, int len) { if (len == 1) return 0; while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p; return p; };
    for (int y = 0; y < h; y++) {
        const uchar* r0 = img + (size_t)R(y - 1, h) * pitch; const uchar* r1 = img + (size_t)y * pitch; const uchar* r2 = img + (size_t)R(y + 1, h) * pitch;
        for (int x = 0; x < w; x++) {
            const int xm = R(x - 1, w), xp = R(x + 1, w);
            dx[(size_t)y * w + x] = (int16_t)((r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]));
            dy[(size_t)y * w + x] = (int16_t)((r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]));
        }
    }
}

const int NUM_OF_BANDS = 9, WIDTH_OF_BAND = 7;
const int kCombinations[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6}, {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7},
                                  {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8}, {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

/* BinaryDescriptor::computeLBD + binaryConversion (opencv_contrib line_descriptor binary_descriptor.cpp) [memory] */
void lbd_descriptor(const orc_keyline& kl, const int16_t* dxImg, const int16_t* dyImg, int W, int H, uchar* out32) {
    static double gaussCoefL[WIDTH_OF_BAND * 3], gaussCoefG[NUM_OF_BANDS * WIDTH_OF_BAND];
    static bool init = false;
    if (!init) {
        double u = (WIDTH_OF_BAND * 3 - 1) / 2, sigma = (WIDTH_OF_BAND * 2 + 1) / 2, invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < WIDTH_OF_BAND * 3; i++) { double dis = i - u; gaussCoefL[i] = exp(dis * dis * invsigma2); }
        u = (NUM_OF_BANDS * WIDTH_OF_BAND - 1) / 2; sigma = u; invsigma2 = -1 / (2 * sigma * sigma);
        for (int i = 0; i < NUM_OF_BANDS * WIDTH_OF_BAND; i++) { double dis = i - u; gaussCoefG[i] = exp(dis * dis * invsigma2); }
        init = true;
    }
    const short heightOfLSP = WIDTH_OF_BAND * NUM_OF_BANDS, halfHeight = (heightOfLSP - 1) / 2;
    float band[8][NUM_OF_BANDS];   // pgdL, ngdL, pgdL2, ngdL2, pgdO, ngdO, pgdO2, ngdO2
    memset(band, 0, sizeof(band));
    const short realWidth = (short)W, imageWidth = (short)(W - 1), imageHeight = (short)(H - 1);
    const short lengthOfLSP = (short)kl.numOfPixels, halfWidth = (lengthOfLSP - 1) / 2;
    const float lineMiddlePointX = (float)(0.5 * (kl.sPointInOctaveX + kl.ePointInOctaveX));
    const float lineMiddlePointY = (float)(0.5 * (kl.sPointInOctaveY + kl.ePointInOctaveY));
    float dL[2], dO[2];
    dL[0] = (float)cos((double)kl.angle); dL[1] = (float)sin((double)kl.angle);
    dO[0] = -dL[1]; dO[1] = dL[0];
    float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + lineMiddlePointX;
    float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + lineMiddlePointY;
    for (short hID = 0; hID < heightOfLSP; hID++) {
        float sCorX = sCorX0, sCorY = sCorY0;
        float pgdLRowSum = 0, ngdLRowSum = 0, pgdORowSum = 0, ngdORowSum = 0;
        for (short wID = 0; wID < lengthOfLSP; wID++) {
            short tempCor = (short)roundf(sCorX);
            const short xCor = (tempCor < 0) ? 0 : (tempCor > imageWidth) ? imageWidth : tempCor;
            tempCor = (short)roundf(sCorY);
            const short yCor = (tempCor < 0) ? 0 : (tempCor > imageHeight) ? imageHeight : tempCor;
            const short dx = dxImg[yCor * realWidth + xCor], dy = dyImg[yCor * realWidth + xCor];
            const float gDL = dx * dL[0] + dy * dL[1], gDO = dx * dO[0] + dy * dO[1];
            if (gDL > 0) pgdLRowSum += gDL; else ngdLRowSum -= gDL;
            if (gDO > 0) pgdORowSum += gDO; else ngdORowSum -= gDO;
            sCorX += dL[0]; sCorY += dL[1];
        }
        sCorX0 -= dL[1]; sCorY0 += dL[0];
        float coef = (float)gaussCoefG[hID];
        pgdLRowSum = coef * pgdLRowSum; ngdLRowSum = coef * ngdLRowSum;
        const float pgdL2RowSum = pgdLRowSum * pgdLRowSum, ngdL2RowSum = ngdLRowSum * ngdLRowSum;
        pgdORowSum = coef * pgdORowSum; ngdORowSum = coef * ngdORowSum;
        const float pgdO2RowSum = pgdORowSum * pgdORowSum, ngdO2RowSum = ngdORowSum * ngdORowSum;
        auto add = [&](short b, float c) {
            band[0][b] += c * pgdLRowSum; band[1][b] += c * ngdLRowSum;
            band[2][b] += c * c * pgdL2RowSum; band[3][b] += c * c * ngdL2RowSum;
            band[4][b] += c * pgdORowSum; band[5][b] += c * ngdORowSum;
            band[6][b] += c * c * pgdO2RowSum; band[7][b] += c * c * ngdO2RowSum;
        };
        short bandID = (short)(hID / WIDTH_OF_BAND);
        add(bandID, (float)gaussCoefL[hID % WIDTH_OF_BAND + WIDTH_OF_BAND]);
        bandID--;
        if (bandID >= 0) add(bandID, (float)gaussCoefL[hID % WIDTH_OF_BAND + 2 * WIDTH_OF_BAND]);
        bandID = bandID + 2;
        if (bandID < NUM_OF_BANDS) add(bandID, (float)gaussCoefL[hID % WIDTH_OF_BAND]);
    }
    float desVec[NUM_OF_BANDS * 8];
    const float invN2 = (float)(1.0 / (WIDTH_OF_BAND * 2.0)), invN3 = (float)(1.0 / (WIDTH_OF_BAND * 3.0));
    for (short b = 0; b < NUM_OF_BANDS; b++) {
        const float invN = (b == 0 || b == NUM_OF_BANDS - 1) ? invN2 : invN3;
        const short d = b * 8;
        float temp = band[0][b] * invN; desVec[d] = temp; desVec[d + 4] = std::sqrt(band[2][b] * invN - temp * temp);
        temp = band[1][b] * invN; desVec[d + 1] = temp; desVec[d + 5] = std::sqrt(band[3][b] * invN - temp * temp);
        temp = band[4][b] * invN; desVec[d + 2] = temp; desVec[d + 6] = std::sqrt(band[6][b] * invN - temp * temp);
        temp = band[5][b] * invN; desVec[d + 3] = temp; desVec[d + 7] = std::sqrt(band[7][b] * invN - temp * temp);
    }
    float tempM = 0, tempS = 0;
    for (int b = 0; b < NUM_OF_BANDS; b++) {
        for (int i = 0; i < 4; i++) tempM += desVec[8 * b + i] * desVec[8 * b + i];
        for (int i = 4; i < 8; i++) tempS += desVec[8 * b + i] * desVec[8 * b + i];
    }
    tempM = 1 / std::sqrt(tempM); tempS = 1 / std::sqrt(tempS);
    for (int b = 0; b < NUM_OF_BANDS; b++) {
        for (int i = 0; i < 4; i++) desVec[8 * b + i] = desVec[8 * b + i] * tempM;
        for (int i = 4; i < 8; i++) desVec[8 * b + i] = desVec[8 * b + i] * tempS;
    }
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) if (desVec[i] > 0.4) desVec[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) temp += desVec[i] * desVec[i];
    temp = 1 / std::sqrt(temp);
    for (int i = 0; i < NUM_OF_BANDS * 8; i++) desVec[i] = desVec[i] * temp;
    for (int c = 0; c < 32; c++) {                                   // binaryConversion: MSB first
        const float* f1 = &desVec[8 * kCombinations[c][0]]; const float* f2 = &desVec[8 * kCombinations[c][1]];
        uchar r = 0;
        for (int i = 0; i < 8; i++) if (f1[i] > f2[i]) r += (uchar)(0x80 >> i);
        out32[c] = r;
    }
}

}  // namespace

struct orc_line {
    int nfeat;
    std::vector<double> trace;
    std::vector<float> raw;       // x1,y1,x2,y2 per raw LSD segment
    std::vector<uchar> scaled; int sw = 0, sh = 0;
    double ms[4];
};

extern "C" orc_line* orc_line_create(int lsdNFeatures) { orc_line* o = new orc_line(); o->nfeat = lsdNFeatures; return o; }
extern "C" void orc_line_destroy(orc_line* o) { delete o; }

extern "C" int orc_lsd_detect_scaled(const uchar* img, int w, int h, int pitch, float* seg4, int cap) {
    Lsd lsd; std::vector<Rect> recs;
    lsd.detect(img, w, h, pitch, recs);
    int n = 0;
    for (const Rect& r : recs) {
        if (n < cap) { seg4[4 * n] = float(r.x1 + 0.5); seg4[4 * n + 1] = float(r.y1 + 0.5); seg4[4 * n + 2] = float(r.x2 + 0.5); seg4[4 * n + 3] = float(r.y2 + 0.5); }
        n++;
    }
    return n;
}

extern "C" void orc_lbd_prep(const uchar* img, int w, int h, int pitch, int16_t* dx, int16_t* dy) {
    static const int taps5[5] = {14, 62, 104, 62, 14};        // GaussianBlur(5x5, sigma 1), 4.13 fixed point
    std::vector<uchar> blur((size_t)w * h);
    orc_sepfilter_fixed_u8(img, w, h, pitch, blur.data(), w, taps5, 5);
    sobel3(blur.data(), w, h, w, dx, dy);
}

/* cv::LineIterator(img, p1, p2, 8).count after cv::clipLine — imgproc drawing.cpp [memory], pinned to cv2.clipLine in
   tests/test_line_oracle_cpu.py.  The endpoints are cvRound()ed floats in [0, lim): 639.6 rounds to 640, one past the last
   column, and OpenCV then clips the segment to the image before counting. */
extern "C" int orc_clip_line(int w, int h, long long* px1, long long* py1, long long* px2, long long* py2) {
    long long &x1 = *px1, &y1 = *py1, &x2 = *px2, &y2 = *py2;
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return 0;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (a - y1) * (x2 - x1) / (y2 - y1); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (a - y2) * (x2 - x1) / (y2 - y1); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += (a - x1) * (y2 - y1) / (x2 - x1); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += (a - x2) * (y2 - y1) / (x2 - x1); x2 = a; c2 = 0; }
        }
    }
    return (c1 | c2) == 0;
}
extern "C" int orc_line_iterator_count(int w, int h, int ax, int ay, int bx, int by) {
    long long x1 = ax, y1 = ay, x2 = bx, y2 = by;
    if ((unsigned)ax >= (unsigned)w || (unsigned)bx >= (unsigned)w || (unsigned)ay >= (unsigned)h || (unsigned)by >= (unsigned)h)
        if (!orc_clip_line(w, h, &x1, &y1, &x2, &y2)) return 0;
    return (int)std::max(std::llabs(x2 - x1), std::llabs(y2 - y1)) + 1;
}

namespace {
/* LSDDetector::detect(image, keylines, scale=1, numOctaves=1): cv LSD (REFINE_ADV) + KeyLine packaging
   (line_descriptor LSDDetector.cpp [memory]); all segments, detection order */
void detect_keylines(const uchar* img, int w, int h, int pitch, std::vector<orc_keyline>& kls, orc_line* o, double* tms) {
    double t0 = now_ms();
    // cv::LineSegmentDetector at scale 0.8: GaussianBlur(sigma 0.6/0.8, 7x7) + INTER_LINEAR_EXACT resize (lsd.cpp)
    static const int taps7[7] = {0, 4, 56, 136, 56, 4, 0};
    std::vector<uchar> blur((size_t)w * h);
    orc_sepfilter_fixed_u8(img, w, h, pitch, blur.data(), w, taps7, 7);
    const double SCALE = 0.8;
    const int sw = cvRoundD(w * SCALE), sh = cvRoundD(h * SCALE);
    std::vector<uchar> scaled((size_t)sw * sh, 0);
    resize_linear_exact(blur.data(), w, h, w, scaled.data(), sw, sh, sw, 0.8);
    double t1 = now_ms();
    Lsd lsd; std::vector<Rect> recs;
    lsd.detect(scaled.data(), sw, sh, sw, recs);
    double t2 = now_ms();
    if (o) { o->trace = lsd.trace; o->sw = sw; o->sh = sh; o->scaled.swap(scaled); o->raw.clear(); }
    kls.clear();
    int class_counter = 0;
    for (Rect r : recs) {
        r.x1 += 0.5; r.y1 += 0.5; r.x2 += 0.5; r.y2 += 0.5;
        r.x1 /= SCALE; r.y1 /= SCALE; r.x2 /= SCALE; r.y2 /= SCALE;
        float e[4] = {float(r.x1), float(r.y1), float(r.x2), float(r.y2)};
        if (o) o->raw.insert(o->raw.end(), e, e + 4);
        // checkLineExtremes
        for (int k = 0; k < 4; k++) {
            const int lim = (k & 1) ? h : w;
            if (e[k] < 0) e[k] = 0;
            if (e[k] >= lim) e[k] = (float)lim - 1.0f;
        }
        orc_keyline kl;
        kl.startPointX = e[0]; kl.startPointY = e[1]; kl.endPointX = e[2]; kl.endPointY = e[3];   // * octaveScale (= 1)
        kl.sPointInOctaveX = e[0]; kl.sPointInOctaveY = e[1]; kl.ePointInOctaveX = e[2]; kl.ePointInOctaveY = e[3];
        kl.lineLength = (float)sqrt(pow((double)(e[0] - e[2]), 2) + pow((double)(e[1] - e[3]), 2));
        // cv::LineIterator(img, Point2f->Point (cvRound), 8-connectivity).count
        kl.numOfPixels = orc_line_iterator_count(w, h, cvRoundD(e[0]), cvRoundD(e[1]), cvRoundD(e[2]), cvRoundD(e[3]));
        kl.angle = (float)atan2((double)(kl.endPointY - kl.startPointY), (double)(kl.endPointX - kl.startPointX));
        kl.class_id = class_counter++;
        kl.octave = 0;
        kl.size = (kl.endPointX - kl.startPointX) * (kl.endPointY - kl.startPointY);
        kl.response = kl.lineLength / (float)std::max(w, h);
        kl.pt_x = (kl.endPointX + kl.startPointX) / 2; kl.pt_y = (kl.endPointY + kl.startPointY) / 2;
        kls.push_back(kl);
    }
    if (tms) { tms[0] = t1 - t0; tms[1] = t2 - t1; }
}
}  // namespace

extern "C" int orc_lsd_keylines(const uchar* img, int w, int h, int pitch, orc_keyline* out, int cap) {
    std::vector<orc_keyline> kls;
    detect_keylines(img, w, h, pitch, kls, nullptr, nullptr);
    for (int i = 0; i < (int)kls.size() && i < cap; i++) out[i] = kls[i];
    return (int)kls.size();
}

/* BinaryDescriptor::compute(image, keylines, descriptors): GaussianBlur(5x5, 1) + Sobel, then LBD per line [memory] */
extern "C" void orc_lbd_compute(const uchar* img, int w, int h, int pitch, const orc_keyline* kls, int n, uchar* ldesc) {
    std::vector<int16_t> dx((size_t)w * h), dy((size_t)w * h);
    orc_lbd_prep(img, w, h, pitch, dx.data(), dy.data());
    for (int i = 0; i < n; i++) lbd_descriptor(kls[i], dx.data(), dy.data(), w, h, ldesc + 32 * (size_t)i);
}

/* LineSegment::ExtractLineSegment — ExtractLineSegment.cpp:18-69 (scale=1, numOctaves=1) */
extern "C" int orc_line_extract(orc_line* o, const uchar* img, int w, int h, int pitch,
                                orc_keyline* klout, uchar* ldesc, double* lineeq3, int cap) {
    std::vector<orc_keyline> kls;
    double tms[2];
    detect_keylines(img, w, h, pitch, kls, o, tms);
    double t2 = now_ms();
    // ExtractLineSegment.cpp:45-51: keep the lsdNFeatures strongest (stable order on exact ties: the reference's
    // std::sort is unstable)
    if ((int)kls.size() > o->nfeat) {
        std::stable_sort(kls.begin(), kls.end(), [](const orc_keyline& a, const orc_keyline& b) { return a.response > b.response; });
        kls.resize(o->nfeat);
        for (int i = 0; i < o->nfeat; i++) kls[i].class_id = i;
    }
    double t3 = now_ms();
    const int n = std::min((int)kls.size(), cap);
    orc_lbd_compute(img, w, h, pitch, kls.data(), n, ldesc);
    for (int i = 0; i < n; i++) {
        klout[i] = kls[i];
        // line equation, ExtractLineSegment.cpp:56-68 (double cross product of f32 endpoints, normalised by |(l0,l1)|)
        const double sx = kls[i].startPointX, sy = kls[i].startPointY, ex = kls[i].endPointX, ey = kls[i].endPointY;
        double l0 = sy * 1.0 - 1.0 * ey, l1 = 1.0 * ex - sx * 1.0, l2 = sx * ey - sy * ex;
        const double nrm = sqrt(l0 * l0 + l1 * l1);
        lineeq3[3 * i] = l0 / nrm; lineeq3[3 * i + 1] = l1 / nrm; lineeq3[3 * i + 2] = l2 / nrm;
    }
    double t4 = now_ms();
    o->ms[0] = tms[0]; o->ms[1] = tms[1]; o->ms[2] = t3 - t2; o->ms[3] = t4 - t3;
    return n;
}

extern "C" int orc_line_raw_segments(const orc_line* o, float* seg4, int cap) {
    const int n = (int)o->raw.size() / 4;
    for (int i = 0; i < n && i < cap; i++) memcpy(seg4 + 4 * i, &o->raw[4 * i], 16);
    return n;
}
extern "C" void orc_line_scaled_copy(const orc_line* o, uchar* dst, int dpitch, int* w, int* h) {
    *w = o->sw; *h = o->sh;
    if (dst) for (int y = 0; y < o->sh; y++) memcpy(dst + (size_t)y * dpitch, &o->scaled[(size_t)y * o->sw], o->sw);
}
extern "C" void orc_line_stage_ms(const orc_line* o, double* ms4) { for (int i = 0; i < 4; i++) ms4[i] = o->ms[i]; }

extern "C" int orc_line_trace(const orc_line* o, double* out, int cap_rows) {
    const int n = (int)o->trace.size() / 10;
    for (int i = 0; i < n && i < cap_rows; i++) memcpy(out + 10 * i, &o->trace[10 * i], 80);
    return n;
}
