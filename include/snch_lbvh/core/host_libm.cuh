// snch_lbvh/core/host_libm.cuh — the HOST libm's float results, bit for bit, in device code.
//
// Part of the B200-native SNCH-LBVH (snch-lbvh_b200).  The reference evaluates acosf / sinf / cosf (normal-cone merge,
// core/cone.cuh:34-66, 288-302, 427-480; leaf cones, scene.cuh:909-961) and logf (scene<2>::green_weight,
// scene.cuh:606-613) through whatever libm its build target has: glibc on the CPU, the CUDA math library on the GPU.  The two
// differ in the last place on a few per cent of the arguments, and a merged cone axis is rotated about
// normalize(cross(a, b)), which amplifies one ulp by 1 / angle and hands it to every ancestor.  This library's parity bar is
// the reference's CPU build, so its device code evaluates glibc's own algorithms:
//   * acosf  — sysdeps/ieee754/flt-32/e_acosf.c (the fdlibm / msun rational approximation, float arithmetic);
//   * sinf, cosf — sysdeps/ieee754/flt-32/s_sincosf.h (ARM optimized routines: quadrant reduction and two polynomials in
//     double, one rounding), with the multiply-adds of the FMA build glibc selects on every x86-64 CPU since Haswell;
//   * logf   — sysdeps/ieee754/flt-32/e_logf.c (16-entry table and a cubic in double, one rounding).
// Each restatement was compared with glibc 2.39 on EVERY argument of its domain (acosf: all 2 130 706 434 floats in [-1, 1];
// sinf / cosf: all 2 246 049 792 floats with |x| < 120; logf: all 2 130 706 432 positive normal floats): zero differences.
// (logf gives the same floats with and without contracted multiply-adds; sinf / cosf differ on 12 / 22 arguments of the
// 2.2e9.)  Outside those domains, and in host code, the platform's function is called.  Every operation is an explicit
// round-to-nearest intrinsic, so the compiler's own contraction (-fmad) cannot change a result.
#ifndef SNCH_LBVH_B200_HOST_LIBM_CUH
#define SNCH_LBVH_B200_HOST_LIBM_CUH
#include <cmath>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define SNCH_LBVH_LIBM_CALLABLE inline __host__ __device__
#else
#define SNCH_LBVH_LIBM_CALLABLE inline
#endif

namespace lbvh
{
namespace detail
{
#ifdef __CUDA_ARCH__
namespace host_libm
{
__device__ __forceinline__ double d(unsigned long long bits) { return __longlong_as_double((long long)bits); }
// p(z) / q(z) of e_acosf.c
__device__ __forceinline__ float acos_ratio(float z)
{
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f,
                pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    float p = __fadd_rn(pS4, __fmul_rn(z, pS5));
    p = __fadd_rn(pS3, __fmul_rn(z, p));
    p = __fadd_rn(pS2, __fmul_rn(z, p));
    p = __fadd_rn(pS1, __fmul_rn(z, p));
    p = __fadd_rn(pS0, __fmul_rn(z, p));
    p = __fmul_rn(z, p);
    float q = __fadd_rn(qS3, __fmul_rn(z, qS4));
    q = __fadd_rn(qS2, __fmul_rn(z, q));
    q = __fadd_rn(qS1, __fmul_rn(z, q));
    q = __fadd_rn(1.0f, __fmul_rn(z, q));
    return __fdiv_rn(p, q);
}
// the two polynomials of s_sincosf.h (sinf_poly) with the FMA build's contractions; `neg` selects the table of negated
// coefficients (quadrants 2 and 3)
__device__ __forceinline__ float sincos_poly(double x, double x2, bool neg, int n)
{
    if ((n & 1) == 0)
    {
        const double s1c = d(0xbfc555545995a603ull), s2c = d(0x3f81107605230bc4ull), s3c = d(0xbf2994eb3774cf24ull);
        const double x3 = __dmul_rn(x, x2);
        const double s1 = __fma_rn(x2, s3c, s2c);
        const double x7 = __dmul_rn(x3, x2);
        const double s = __fma_rn(x3, s1c, x);
        return __double2float_rn(__fma_rn(x7, s1, s));
    }
    const double sg = neg ? -1.0 : 1.0;
    const double c0 = sg, c1c = __dmul_rn(sg, d(0xbfdffffffd0c621cull)), c2c = __dmul_rn(sg, d(0x3fa55553e1068f19ull)),
                 c3c = __dmul_rn(sg, d(0xbf56c087e89a359dull)), c4c = __dmul_rn(sg, d(0x3ef99343027bf8c3ull));
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __fma_rn(x2, c4c, c3c);
    const double c1 = __fma_rn(x2, c1c, c0);
    const double x6 = __dmul_rn(x4, x2);
    const double c = __fma_rn(x4, c2c, c1);
    return __double2float_rn(__fma_rn(x6, c2, c));
}
__device__ __forceinline__ unsigned int abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ffu; }
// sinf (which = 0) / cosf (which = 1) for |y| < 120
__device__ __forceinline__ float sincos_small(float y, int which)
{
    double x = (double)y;
    if (abstop12(y) < 0x3f4u) // |y| < pi / 4
    {
        if (abstop12(y) < 0x398u) return which ? 1.0f : y; // |y| < 2^-12
        return sincos_poly(x, __dmul_rn(x, x), false, which);
    }
    const double r = __dmul_rn(x, d(0x41645f306dc9c883ull));
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = __fma_rn(-(double)n, d(0x3ff921fb54442d18ull), x);
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sincos_poly(__dmul_rn(x, sgn), __dmul_rn(x, x), (n & 2) != 0, n ^ which);
}
} // namespace host_libm
#endif

SNCH_LBVH_LIBM_CALLABLE float acosf_host(float x) noexcept
{
#ifdef __CUDA_ARCH__
    const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const int hx = __float_as_int(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : __fadd_rn(pi, __fmul_rn(2.0f, pio2_lo));
    if (ix > 0x3f800000) return __fdiv_rn(__fsub_rn(x, x), __fsub_rn(x, x)); // |x| > 1 or NaN
    if (ix < 0x3f000000)
    { // |x| < 0.5
        if (ix <= 0x32800000) return __fadd_rn(pio2_hi, pio2_lo);
        const float r = host_libm::acos_ratio(__fmul_rn(x, x));
        return __fsub_rn(pio2_hi, __fsub_rn(x, __fsub_rn(pio2_lo, __fmul_rn(x, r))));
    }
    if (hx < 0)
    { // x <= -0.5
        const float z = __fmul_rn(__fadd_rn(1.0f, x), 0.5f);
        const float r = host_libm::acos_ratio(z);
        const float s = __fsqrt_rn(z);
        const float w = __fsub_rn(__fmul_rn(r, s), pio2_lo);
        return __fsub_rn(pi, __fmul_rn(2.0f, __fadd_rn(s, w)));
    }
    const float z = __fmul_rn(__fsub_rn(1.0f, x), 0.5f);
    const float s = __fsqrt_rn(z);
    const float df = __uint_as_float(__float_as_uint(s) & 0xfffff000u);
    const float c = __fdiv_rn(__fsub_rn(z, __fmul_rn(df, df)), __fadd_rn(s, df));
    const float r = host_libm::acos_ratio(z);
    const float w = __fadd_rn(__fmul_rn(r, s), c);
    return __fmul_rn(2.0f, __fadd_rn(df, w));
#else
    return ::acosf(x);
#endif
}
SNCH_LBVH_LIBM_CALLABLE float sinf_host(float x) noexcept
{
#ifdef __CUDA_ARCH__
    if (host_libm::abstop12(x) < 0x42fu) return host_libm::sincos_small(x, 0); // |x| < 120
#endif
    return ::sinf(x);
}
SNCH_LBVH_LIBM_CALLABLE float cosf_host(float x) noexcept
{
#ifdef __CUDA_ARCH__
    if (host_libm::abstop12(x) < 0x42fu) return host_libm::sincos_small(x, 1);
#endif
    return ::cosf(x);
}
SNCH_LBVH_LIBM_CALLABLE float logf_host(float x) noexcept
{
#ifdef __CUDA_ARCH__
    const unsigned int ix = __float_as_uint(x);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return ::logf(x); // zero, subnormal, negative, infinite, NaN
    if (ix == 0x3f800000u) return 0.0f;
    constexpr unsigned long long tab[32] = {
        0x3ff661ec79f8f3beull, 0xbfd57bf7808caadeull, 0x3ff571ed4aaf883dull, 0xbfd2bef0a7c06ddbull, 0x3ff49539f0f010b0ull, 0xbfd01eae7f513a67ull,
        0x3ff3c995b0b80385ull, 0xbfcb31d8a68224e9ull, 0x3ff30d190c8864a5ull, 0xbfc6574f0ac07758ull, 0x3ff25e227b0b8ea0ull, 0xbfc1aa2bc79c8100ull,
        0x3ff1bb4a4a1a343full, 0xbfba4e76ce8c0e5eull, 0x3ff12358f08ae5baull, 0xbfb1973c5a611cccull, 0x3ff0953f419900a7ull, 0xbfa252f438e10c1eull,
        0x3ff0000000000000ull, 0x0000000000000000ull, 0x3fee608cfd9a47acull, 0x3faaa5aa5df25984ull, 0x3feca4b31f026aa0ull, 0x3fbc5e53aa362eb4ull,
        0x3feb2036576afce6ull, 0x3fc526e57720db08ull, 0x3fe9c2d163a1aa2dull, 0x3fcbc2860d224770ull, 0x3fe886e6037841edull, 0x3fd1058bc8a07ee1ull,
        0x3fe767dcf5534862ull, 0x3fd4043057b6ee09ull}; // {1 / c, ln c} for the sixteen subintervals of [0.7, 1.4)
    const unsigned int tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u), k = (int)tmp >> 23;
    const double z = (double)__uint_as_float(ix - (tmp & 0xff800000u));
    const double r = __dsub_rn(__dmul_rn(z, host_libm::d(tab[2 * i])), 1.0);
    const double y0 = __dadd_rn(host_libm::d(tab[2 * i + 1]), __dmul_rn((double)k, host_libm::d(0x3fe62e42fefa39efull)));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(host_libm::d(0x3fd5575b0be00b6aull), r), host_libm::d(0xbfdffffef20a4123ull));
    y = __dadd_rn(__dmul_rn(host_libm::d(0xbfd00ea348b88334ull), r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
#else
    return ::logf(x);
#endif
}
} // namespace detail
} // namespace lbvh
#endif
