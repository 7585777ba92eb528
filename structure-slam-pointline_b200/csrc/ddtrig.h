// ddtrig.h — correctly-rounded double sin/cos for |x| < ~1e3 via double-double arithmetic (explicit FMA only).
//
// Why: LSD's region2rect builds the rectangle from cos(theta), sin(theta); the extreme region pixels lie exactly
// ON the rectangle's end edges, so whether rect_nfa counts them depends on the last bit of cos/sin.  OpenCV (glibc)
// is almost always correctly rounded; CUDA's libm is not (<= 2 ulp).  Evaluating to ~106 bits and rounding once makes
// the device agree with the correctly-rounded value.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define DD_HD __host__ __device__ __forceinline__
#else
#define DD_HD inline
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

// sin and cos of x, each rounded once from a ~106-bit result
DD_HD void sincos_cr(double x, double* s_out, double* c_out) {
    // pi/2 = C1 + C2 + C3 + C4 (C1, C2 carry 33 significant bits: k*C1, k*C2 are exact for |k| < 2^20)
    const double C1 = 1.5707963267341256, C2 = 6.077100506303966e-11, C3 = 2.0222662487959506e-21, C4 = 1.0085854035872483e-37;
    const double k = rint(x * 0.6366197723675814);
    dd r = two_sum(x, -k * C1);
    r = add(r, dd{-k * C2, 0.0});
    r = add(r, neg(two_prod(k, C3)));
    r = add(r, neg(two_prod(k, C4)));
    const dd r2 = mul(r, r);
    // 1/n!, n = 2..29, as double-double
    const double fh[28] = {0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
        0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07, 2.505210838544172e-08,
        2.08767569878681e-09, 1.6059043836821613e-10, 1.1470745597729725e-11, 7.647163731819816e-13, 4.779477332387385e-14,
        2.8114572543455206e-15, 1.5619206968586225e-16, 8.22063524662433e-18, 4.110317623312165e-19, 1.9572941063391263e-20,
        8.896791392450574e-22, 3.868170170630684e-23, 1.6117375710961184e-24, 6.446950284384474e-26, 2.4795962632247976e-27,
        9.183689863795546e-29, 3.279889237069838e-30, 1.1309962886447716e-31};
    const double fl[28] = {0.0, 9.25185853854297e-18, 2.3129646346357427e-18, 1.1564823173178714e-19, -5.300543954373577e-20,
        1.7209558293420705e-22, 2.1511947866775882e-23, -1.858393274046472e-22, 2.3767714622250297e-23, -1.448814070935912e-24,
        -1.20734505911326e-25, 1.2585294588752098e-26, 2.0655512752830745e-28, 7.03872877733453e-30, 4.399205485834081e-31,
        1.6508842730861433e-31, 1.1910679660273754e-32, 2.2141894119604265e-34, 1.4412973378659527e-36, -1.3643503830087908e-36,
        -7.911402614872376e-38, -8.843177655482344e-40, -3.6846573564509766e-41, -1.9330404233703465e-42, -1.2953730964765229e-43,
        1.4303150396787322e-45, 1.5117542744029879e-46, 1.0498015412959506e-47};
    // sin r = r (1 - r2 (1/3! - r2 (1/5! - ... - r2/29!)));  cos r = 1 - r2 (1/2! - r2 (1/4! - ... - r2/28!))
    dd ps = dd{fh[27], fl[27]};                                   // 1/29!
    for (int n = 27; n >= 3; n -= 2) ps = add(dd{fh[n - 2], fl[n - 2]}, neg(mul(r2, ps)));
    ps = add(dd{1.0, 0.0}, neg(mul(r2, ps)));
    dd pc = dd{fh[26], fl[26]};                                   // 1/28!
    for (int n = 26; n >= 2; n -= 2) pc = add(dd{fh[n - 2], fl[n - 2]}, neg(mul(r2, pc)));
    pc = add(dd{1.0, 0.0}, neg(mul(r2, pc)));
    const dd sr = mul(r, ps), cr = pc;
    const int q = ((int)k) & 3;
    dd s = sr, c = cr;
    if (q == 1) { s = cr; c = neg(sr); }
    else if (q == 2) { s = neg(sr); c = neg(cr); }
    else if (q == 3) { s = neg(cr); c = sr; }
    *s_out = s.hi + s.lo; *c_out = c.hi + c.lo;
}

}  // namespace ddtrig
