// mathx.cuh — bit-exact float helpers shared by the ORB and line kernels (explicit _rn intrinsics: no FMA contraction).
#pragma once
#include <cfloat>

namespace sslpl {

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {       // cv::fastAtan2, SURVEY.md A.4
    const float sc = 57.29577951308232f;                                   // (float)(180/CV_PI)
    const float p1 = __fmul_rn(0.9997878412794807f, sc), p3 = __fmul_rn(-0.3258083974640975f, sc),
                p5 = __fmul_rn(0.1555786518463281f, sc), p7 = __fmul_rn(-0.04432655554792128f, sc);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON)); c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON)); c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__host__ __device__ __forceinline__ int reflect101(int p, int len) {      // BORDER_REFLECT_101
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

}  // namespace sslpl
