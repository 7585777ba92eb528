// ddtrig.h — correctly-rounded double sin/cos for |x| < ~1e3 via double-double arithmetic (explicit FMA only).
//
// Why: LSD's region2rect builds the rectangle from cos(theta), sin(theta); the extreme region pixels lie exactly
// ON the rectangle's end edges, so whether rect_nfa counts them depends on the last bit of cos/sin.  OpenCV (glibc)
// is almost always correctly rounded; CUDA's libm is not (<= 2 ulp).  Evaluating to ~106 bits and rounding once makes
// the device agree with the correctly-rounded value.
//
// Layout for the GPU: ONE rolled Horner routine serves both the quick (Ziv first attempt) and the full evaluation,
// and the factorial table sits in constant memory — the routine is a few KB of SASS instead of ~50 KB unrolled (the
// region walker keeps ~25 warps per SM at different program counters; code size is what its instruction cache sees).
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define DD_HD __device__ __forceinline__
#define DD_CALL __device__ __noinline__
#define DD_TAB static __constant__
#define DD_ROLLED _Pragma("unroll 1")
#else
#define DD_HD inline
#define DD_CALL inline
#define DD_TAB static const
#define DD_ROLLED
#endif

namespace ddtrig {

struct dd { double hi, lo; };

DD_HD dd two_sum(double a, double b) { const double s = a + b, bb = s - a; return dd{s, (a - (s - bb)) + (b - bb)}; }
DD_HD dd quick_two_sum(double a, double b) { const double s = a + b; return dd{s, b - (s - a)}; }
DD_HD dd two_prod(double a, double b) { const double p = a * b; return dd{p, fma(a, b, -p)}; }
DD_HD dd add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi); const dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi; s = quick_two_sum(s.hi, s.lo); s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
DD_HD dd neg(dd a) { return dd{-a.hi, -a.lo}; }
DD_HD dd mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}

// 1/n!, n = 0..29, as double-double (hi, lo)
DD_TAB double FACT_H[30] = {1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
    0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07, 2.505210838544172e-08,
    2.08767569878681e-09, 1.6059043836821613e-10, 1.1470745597729725e-11, 7.647163731819816e-13, 4.779477332387385e-14,
    2.8114572543455206e-15, 1.5619206968586225e-16, 8.22063524662433e-18, 4.110317623312165e-19, 1.9572941063391263e-20,
    8.896791392450574e-22, 3.868170170630684e-23, 1.6117375710961184e-24, 6.446950284384474e-26, 2.4795962632247976e-27,
    9.183689863795546e-29, 3.279889237069838e-30, 1.1309962886447716e-31};
DD_TAB double FACT_L[30] = {0.0, 0.0, 0.0, 9.25185853854297e-18, 2.3129646346357427e-18, 1.1564823173178714e-19, -5.300543954373577e-20,
    1.7209558293420705e-22, 2.1511947866775882e-23, -1.858393274046472e-22, 2.3767714622250297e-23, -1.448814070935912e-24,
    -1.20734505911326e-25, 1.2585294588752098e-26, 2.0655512752830745e-28, 7.03872877733453e-30, 4.399205485834081e-31,
    1.6508842730861433e-31, 1.1910679660273754e-32, 2.2141894119604265e-34, 1.4412973378659527e-36, -1.3643503830087908e-36,
    -7.911402614872376e-38, -8.843177655482344e-40, -3.6846573564509766e-41, -1.9330404233703465e-42, -1.2953730964765229e-43,
    1.4303150396787322e-45, 1.5117542744029879e-46, 1.0498015412959506e-47};

// r = x - k pi/2 as a double-double, k = rint(2x/pi)
DD_HD dd reduce(double x, double* k_out) {
    // pi/2 = C1 + C2 + C3 + C4 (C1, C2 carry 33 significant bits: k*C1, k*C2 are exact for |k| < 2^20)
    const double C1 = 1.5707963267341256, C2 = 6.077100506303966e-11, C3 = 2.0222662487959506e-21, C4 = 1.0085854035872483e-37;
    const double k = rint(x * 0.6366197723675814);
    dd r = two_sum(x, -k * C1);
    r = add(r, dd{-k * C2, 0.0});
    r = add(r, neg(two_prod(k, C3)));
    r = add(r, neg(two_prod(k, C4)));
    *k_out = k;
    return r;
}

// Horner in double-double from degree n_top (odd) down:  ps <- 1/n! - r2 ps,  pc <- 1/(n-1)! - r2 pc  for n = n_top, n_top-2, .., 1
//   sin r = r (1 - r2 (1/3! - r2 (1/5! - ...))),  cos r = 1 - r2 (1/2! - r2 (1/4! - ...))
DD_HD void horner(dd r, dd r2, int n_top, dd ps, dd pc, dd* s_out, dd* c_out) {
    DD_ROLLED
    for (int n = n_top; n >= 1; n -= 2) {
        ps = add(dd{FACT_H[n], FACT_L[n]}, neg(mul(r2, ps)));
        pc = add(dd{FACT_H[n - 1], FACT_L[n - 1]}, neg(mul(r2, pc)));
    }
    *s_out = mul(r, ps); *c_out = pc;
}

// the full ~106-bit evaluation (terms to 1/29!, 1/28!)
DD_HD void sincos_dd(dd r, dd* s_out, dd* c_out) {
    horner(r, mul(r, r), 27, dd{FACT_H[29], FACT_L[29]}, dd{FACT_H[28], FACT_L[28]}, s_out, c_out);
}

// Ziv-style first attempt: the four leading Horner steps in double-double, the tail (from 1/9! resp. 1/8!) in plain
// double.  The tail's rounding error is < 2^-50 of a quantity that contributes < 2^-18 of the result, so the
// relative error of (hi + lo) is < 2^-66; round_safe() accepts a result when rounding hi + lo cannot be changed by
// an error of 2^-62 |hi|.
DD_HD void sincos_quick(dd r, dd* s_out, dd* c_out) {
    const dd r2 = mul(r, r);
    const double z = r2.hi;
    double ts = FACT_H[21], tc = FACT_H[20];                     // tails: 1/9! - z/11! + .. + z^6/21!,  1/8! - z/10! + .. + z^6/20!
    DD_ROLLED
    for (int n = 19; n >= 9; n -= 2) { ts = FACT_H[n] - z * ts; tc = FACT_H[n - 1] - z * tc; }
    horner(r, r2, 7, dd{ts, 0.0}, dd{tc, 0.0}, s_out, c_out);
}

DD_HD bool round_safe(dd v, double* out) {
    const double e = fabs(v.hi) * 2.168404344971009e-19;          // 2^-62 |hi|
    const double a = v.hi + (v.lo - e), b = v.hi + (v.lo + e);
    *out = a;
    return a == b;
}

DD_HD void quadrant(int q, double sr, double cr, double* s_out, double* c_out) {
    double s = sr, c = cr;
    if (q == 1) { s = cr; c = -sr; }
    else if (q == 2) { s = -sr; c = -cr; }
    else if (q == 3) { s = -cr; c = sr; }
    *s_out = s; *c_out = c;
}

// sin and cos of x, each rounded once from a ~106-bit result (reference implementation: always the long evaluation)
DD_HD void sincos_cr_full(double x, double* s_out, double* c_out) {
    double k; const dd r = reduce(x, &k);
    dd s, c; sincos_dd(r, &s, &c);
    quadrant(((int)k) & 3, s.hi + s.lo, c.hi + c.lo, s_out, c_out);
}

// same result, ~4x cheaper on average: quick evaluation first, the long one only when the rounding is ambiguous (~0.7%)
DD_CALL void sincos_cr(double x, double* s_out, double* c_out) {
    double k; const dd r = reduce(x, &k);
    dd s, c;
    double sv = 0, cv = 0;
    bool done = false;
    DD_ROLLED
    for (int attempt = 0; attempt < 2 && !done; attempt++) {     // one copy of horner() in the code
        const dd r2 = mul(r, r);
        dd ps, pc; int n_top;
        if (attempt == 0) {
            const double z = r2.hi;
            double ts = FACT_H[21], tc = FACT_H[20];
            DD_ROLLED
            for (int n = 19; n >= 9; n -= 2) { ts = FACT_H[n] - z * ts; tc = FACT_H[n - 1] - z * tc; }
            ps = dd{ts, 0.0}; pc = dd{tc, 0.0}; n_top = 7;
        } else {
            ps = dd{FACT_H[29], FACT_L[29]}; pc = dd{FACT_H[28], FACT_L[28]}; n_top = 27;
        }
        horner(r, r2, n_top, ps, pc, &s, &c);
        if (attempt == 0) { const bool ok_s = round_safe(s, &sv), ok_c = round_safe(c, &cv); done = ok_s && ok_c; }
        else { sv = s.hi + s.lo; cv = c.hi + c.lo; done = true; }
    }
    quadrant(((int)k) & 3, sv, cv, s_out, c_out);
}

}  // namespace ddtrig
