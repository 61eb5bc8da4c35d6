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
// 2.2e9.)  Outside those domains the platform's function is called.
//
// ONE source for both sides: `*_glibc` below compile for the device (explicit round-to-nearest intrinsics, so -fmad cannot
// change a result) and for the host (plain operators: build that translation unit with -ffp-contract=off), which is how
// tests/test_host_libm_cpu.py checks the very code the kernels run against the host's libm without a GPU;
// tests/test_gpu_host_libm.py checks the device build on the GPU box.  `*_host` is what the library calls: the restatement in
// device code, the platform's libm in host code.
#ifndef SNCH_LBVH_B200_HOST_LIBM_CUH
#define SNCH_LBVH_B200_HOST_LIBM_CUH
#include <cmath>
#include <cstring>
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
namespace host_libm
{
// ---- the arithmetic both sides agree on: IEEE round-to-nearest, no contraction
#ifdef __CUDA_ARCH__
SNCH_LBVH_LIBM_CALLABLE float fadd(float a, float b) { return __fadd_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE float fsub(float a, float b) { return __fsub_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE float fmul(float a, float b) { return __fmul_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE float fdiv(float a, float b) { return __fdiv_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE float fsqrt(float a) { return __fsqrt_rn(a); }
SNCH_LBVH_LIBM_CALLABLE double dadd(double a, double b) { return __dadd_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE double dsub(double a, double b) { return __dsub_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE double dmul(double a, double b) { return __dmul_rn(a, b); }
SNCH_LBVH_LIBM_CALLABLE double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
SNCH_LBVH_LIBM_CALLABLE float d2f(double a) { return __double2float_rn(a); }
SNCH_LBVH_LIBM_CALLABLE int d2i(double a) { return __double2int_rz(a); }
SNCH_LBVH_LIBM_CALLABLE unsigned int f2u(float a) { return __float_as_uint(a); }
SNCH_LBVH_LIBM_CALLABLE float u2f(unsigned int a) { return __uint_as_float(a); }
SNCH_LBVH_LIBM_CALLABLE double d(unsigned long long bits) { return __longlong_as_double((long long)bits); }
#else
SNCH_LBVH_LIBM_CALLABLE float fadd(float a, float b) { return a + b; }
SNCH_LBVH_LIBM_CALLABLE float fsub(float a, float b) { return a - b; }
SNCH_LBVH_LIBM_CALLABLE float fmul(float a, float b) { return a * b; }
SNCH_LBVH_LIBM_CALLABLE float fdiv(float a, float b) { return a / b; }
SNCH_LBVH_LIBM_CALLABLE float fsqrt(float a) { return ::sqrtf(a); }
SNCH_LBVH_LIBM_CALLABLE double dadd(double a, double b) { return a + b; }
SNCH_LBVH_LIBM_CALLABLE double dsub(double a, double b) { return a - b; }
SNCH_LBVH_LIBM_CALLABLE double dmul(double a, double b) { return a * b; }
SNCH_LBVH_LIBM_CALLABLE double dfma(double a, double b, double c) { return ::fma(a, b, c); }
SNCH_LBVH_LIBM_CALLABLE float d2f(double a) { return (float)a; }
SNCH_LBVH_LIBM_CALLABLE int d2i(double a) { return (int)a; }
SNCH_LBVH_LIBM_CALLABLE unsigned int f2u(float a)
{
    unsigned int u;
    std::memcpy(&u, &a, 4);
    return u;
}
SNCH_LBVH_LIBM_CALLABLE float u2f(unsigned int a)
{
    float f;
    std::memcpy(&f, &a, 4);
    return f;
}
SNCH_LBVH_LIBM_CALLABLE double d(unsigned long long bits)
{
    double v;
    std::memcpy(&v, &bits, 8);
    return v;
}
#endif

// p(z) / q(z) of e_acosf.c
SNCH_LBVH_LIBM_CALLABLE float acos_ratio(float z)
{
    const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f,
                pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    float p = fadd(pS4, fmul(z, pS5));
    p = fadd(pS3, fmul(z, p));
    p = fadd(pS2, fmul(z, p));
    p = fadd(pS1, fmul(z, p));
    p = fadd(pS0, fmul(z, p));
    p = fmul(z, p);
    float q = fadd(qS3, fmul(z, qS4));
    q = fadd(qS2, fmul(z, q));
    q = fadd(qS1, fmul(z, q));
    q = fadd(1.0f, fmul(z, q));
    return fdiv(p, q);
}
// the two polynomials of s_sincosf.h (sinf_poly) with the FMA build's contractions; `neg` selects the table of negated
// coefficients (quadrants 2 and 3)
SNCH_LBVH_LIBM_CALLABLE float sincos_poly(double x, double x2, bool neg, int n)
{
    if ((n & 1) == 0)
    {
        const double s1c = d(0xbfc555545995a603ull), s2c = d(0x3f81107605230bc4ull), s3c = d(0xbf2994eb3774cf24ull);
        const double x3 = dmul(x, x2);
        const double s1 = dfma(x2, s3c, s2c);
        const double x7 = dmul(x3, x2);
        const double s = dfma(x3, s1c, x);
        return d2f(dfma(x7, s1, s));
    }
    const double sg = neg ? -1.0 : 1.0;
    const double c0 = sg, c1c = dmul(sg, d(0xbfdffffffd0c621cull)), c2c = dmul(sg, d(0x3fa55553e1068f19ull)), c3c = dmul(sg, d(0xbf56c087e89a359dull)),
                 c4c = dmul(sg, d(0x3ef99343027bf8c3ull));
    const double x4 = dmul(x2, x2);
    const double c2 = dfma(x2, c4c, c3c);
    const double c1 = dfma(x2, c1c, c0);
    const double x6 = dmul(x4, x2);
    const double c = dfma(x4, c2c, c1);
    return d2f(dfma(x6, c2, c));
}
SNCH_LBVH_LIBM_CALLABLE unsigned int abstop12(float x) { return (f2u(x) >> 20) & 0x7ffu; }
// sinf (which = 0) / cosf (which = 1) for |y| < 120
SNCH_LBVH_LIBM_CALLABLE float sincos_small(float y, int which)
{
    double x = (double)y;
    if (abstop12(y) < 0x3f4u) // |y| < pi / 4
    {
        if (abstop12(y) < 0x398u) return which ? 1.0f : y; // |y| < 2^-12
        return sincos_poly(x, dmul(x, x), false, which);
    }
    const double r = dmul(x, d(0x41645f306dc9c883ull));
    const int n = (d2i(r) + 0x800000) >> 24;
    x = dfma(-(double)n, d(0x3ff921fb54442d18ull), x);
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return sincos_poly(dmul(x, sgn), dmul(x, x), (n & 2) != 0, n ^ which);
}
} // namespace host_libm

// ---- glibc's algorithms (both sides) -------------------------------------------------------------------------------------
SNCH_LBVH_LIBM_CALLABLE float acosf_glibc(float x) noexcept
{
    using namespace host_libm;
    const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    const int hx = (int)f2u(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : fadd(pi, fmul(2.0f, pio2_lo));
    if (ix > 0x3f800000) return fdiv(fsub(x, x), fsub(x, x)); // |x| > 1 or NaN
    if (ix < 0x3f000000)
    { // |x| < 0.5
        if (ix <= 0x32800000) return fadd(pio2_hi, pio2_lo);
        const float r = acos_ratio(fmul(x, x));
        return fsub(pio2_hi, fsub(x, fsub(pio2_lo, fmul(x, r))));
    }
    if (hx < 0)
    { // x <= -0.5
        const float z = fmul(fadd(1.0f, x), 0.5f);
        const float r = acos_ratio(z);
        const float s = fsqrt(z);
        const float w = fsub(fmul(r, s), pio2_lo);
        return fsub(pi, fmul(2.0f, fadd(s, w)));
    }
    const float z = fmul(fsub(1.0f, x), 0.5f);
    const float s = fsqrt(z);
    const float df = u2f(f2u(s) & 0xfffff000u);
    const float c = fdiv(fsub(z, fmul(df, df)), fadd(s, df));
    const float r = acos_ratio(z);
    const float w = fadd(fmul(r, s), c);
    return fmul(2.0f, fadd(df, w));
}
SNCH_LBVH_LIBM_CALLABLE float sinf_glibc(float x) noexcept
{
    if (host_libm::abstop12(x) < 0x42fu) return host_libm::sincos_small(x, 0); // |x| < 120
    return ::sinf(x);
}
SNCH_LBVH_LIBM_CALLABLE float cosf_glibc(float x) noexcept
{
    if (host_libm::abstop12(x) < 0x42fu) return host_libm::sincos_small(x, 1);
    return ::cosf(x);
}
SNCH_LBVH_LIBM_CALLABLE float logf_glibc(float x) noexcept
{
    using namespace host_libm;
    const unsigned int ix = f2u(x);
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
    const double z = (double)u2f(ix - (tmp & 0xff800000u));
    const double r = dsub(dmul(z, d(tab[2 * i])), 1.0);
    const double y0 = dadd(d(tab[2 * i + 1]), dmul((double)k, d(0x3fe62e42fefa39efull)));
    const double r2 = dmul(r, r);
    double y = dadd(dmul(d(0x3fd5575b0be00b6aull), r), d(0xbfdffffef20a4123ull));
    y = dadd(dmul(d(0xbfd00ea348b88334ull), r2), y);
    y = dadd(dmul(y, r2), dadd(y0, r));
    return d2f(y);
}

// ---- what the library calls: the restatement in device code, the platform's libm in host code -----------------------------
#ifdef __CUDA_ARCH__
SNCH_LBVH_LIBM_CALLABLE float acosf_host(float x) noexcept { return acosf_glibc(x); }
SNCH_LBVH_LIBM_CALLABLE float sinf_host(float x) noexcept { return sinf_glibc(x); }
SNCH_LBVH_LIBM_CALLABLE float cosf_host(float x) noexcept { return cosf_glibc(x); }
SNCH_LBVH_LIBM_CALLABLE float logf_host(float x) noexcept { return logf_glibc(x); }
#else
SNCH_LBVH_LIBM_CALLABLE float acosf_host(float x) noexcept { return ::acosf(x); }
SNCH_LBVH_LIBM_CALLABLE float sinf_host(float x) noexcept { return ::sinf(x); }
SNCH_LBVH_LIBM_CALLABLE float cosf_host(float x) noexcept { return ::cosf(x); }
SNCH_LBVH_LIBM_CALLABLE float logf_host(float x) noexcept { return ::logf(x); }
#endif
} // namespace detail
} // namespace lbvh
#endif
