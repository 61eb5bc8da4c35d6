// snch_lbvh/core/utility.cuh — decorators, vector helpers and numeric constants of the drop-in C++ API.
//
// Part of the B200-native SNCH-LBVH (snch-lbvh_b200).  This header re-creates the NAMES and MEANING of the reference's
// core/utility.cuh (tyanyuy3125/snch-lbvh) so that user code written against the reference compiles unchanged; the
// implementation is dimension/precision-generic (one template per operation instead of one overload per CUDA vector
// type).  Float operation order follows the reference where results feed parity-critical paths (dot: utility.cuh:236-239,
// length: :373-398, normalize: :427-447, cross: :449-463).
#ifndef SNCH_LBVH_B200_UTILITY_CUH
#define SNCH_LBVH_B200_UTILITY_CUH
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <type_traits>
#include <cuda_runtime.h>
#include <vector_types.h>

#ifdef __CUDACC__
#define SNCH_LBVH_DEVICE __device__
#define SNCH_LBVH_HOST __host__
#else
#define SNCH_LBVH_DEVICE
#define SNCH_LBVH_HOST
#endif
#define SNCH_LBVH_CALLABLE inline SNCH_LBVH_DEVICE SNCH_LBVH_HOST
#define SNCH_LBVH_HOST_DEVICE SNCH_LBVH_DEVICE SNCH_LBVH_HOST
#define SNCH_LBVH_DEVICE_INLINE inline SNCH_LBVH_DEVICE

namespace lbvh
{
// literal constants instead of std::numeric_limits: usable in device code without --expt-relaxed-constexpr
template <typename T> SNCH_LBVH_CALLABLE T epsilon() noexcept { return sizeof(T) == sizeof(float) ? T(FLT_EPSILON) : T(DBL_EPSILON); }
template <typename T> SNCH_LBVH_CALLABLE T one_minus_epsilon() noexcept { return T(1) - epsilon<T>(); }
template <typename T> SNCH_LBVH_CALLABLE T infinity() noexcept { return T(INFINITY); }

// vector_of<T, dim>::type — the CUDA vector type with `dim` components of T
template <typename T, unsigned int dim> struct vector_of;
template <> struct vector_of<float, 2> { using type = float2; };
template <> struct vector_of<float, 3> { using type = float3; };
template <> struct vector_of<float, 4> { using type = float4; };
template <> struct vector_of<double, 2> { using type = double2; };
template <> struct vector_of<double, 3> { using type = double3; };
template <> struct vector_of<double, 4> { using type = double4; };
template <typename T, unsigned int dim> using vector_of_t = typename vector_of<T, dim>::type;

namespace detail
{
// compile-time description of a CUDA vector type: scalar, arity and component access by index
template <typename V> struct vec_traits;
#define SNCH_LBVH_VEC_TRAITS(V, S, N) \
    template <> struct vec_traits<V> { using scalar = S; static constexpr int size = N; };
SNCH_LBVH_VEC_TRAITS(float2, float, 2)
SNCH_LBVH_VEC_TRAITS(float3, float, 3)
SNCH_LBVH_VEC_TRAITS(float4, float, 4)
SNCH_LBVH_VEC_TRAITS(double2, double, 2)
SNCH_LBVH_VEC_TRAITS(double3, double, 3)
SNCH_LBVH_VEC_TRAITS(double4, double, 4)
SNCH_LBVH_VEC_TRAITS(int2, int, 2)
SNCH_LBVH_VEC_TRAITS(int3, int, 3)
SNCH_LBVH_VEC_TRAITS(int4, int, 4)
SNCH_LBVH_VEC_TRAITS(uint2, unsigned int, 2)
SNCH_LBVH_VEC_TRAITS(uint3, unsigned int, 3)
SNCH_LBVH_VEC_TRAITS(uint4, unsigned int, 4)
#undef SNCH_LBVH_VEC_TRAITS
template <typename V> using scalar_of = typename vec_traits<V>::scalar;
template <typename V, typename = void> struct is_vec : std::false_type {};
template <typename V> struct is_vec<V, std::void_t<typename vec_traits<V>::scalar>> : std::true_type {};
template <typename V> using enable_vec = std::enable_if_t<is_vec<V>::value, int>;
template <typename V> using enable_real_vec = std::enable_if_t<is_vec<V>::value && std::is_floating_point<scalar_of<V>>::value, int>;

// components are laid out x,y,z,w contiguously in every CUDA vector type
template <typename V> SNCH_LBVH_CALLABLE scalar_of<V> &at(V &v, int i) noexcept { return reinterpret_cast<scalar_of<V> *>(&v)[i]; }
template <typename V> SNCH_LBVH_CALLABLE scalar_of<V> at(const V &v, int i) noexcept { return reinterpret_cast<const scalar_of<V> *>(&v)[i]; }
template <typename V, typename F> SNCH_LBVH_CALLABLE V map2(const V &a, const V &b, F f) noexcept
{
    V r;
    for (int i = 0; i < vec_traits<V>::size; ++i) at(r, i) = f(at(a, i), at(b, i));
    return r;
}
SNCH_LBVH_CALLABLE float sqrt_of(float x) noexcept { return ::sqrtf(x); }
SNCH_LBVH_CALLABLE double sqrt_of(double x) noexcept { return ::sqrt(x); }
SNCH_LBVH_CALLABLE float abs_of(float x) noexcept { return ::fabsf(x); }
// Natural logarithm with the HOST libm's bits on the device.  scene<2>::green_weight is |ln r| / 2 pi and feeds the sampling
// pdf; CUDA's logf and glibc's differ in the last place on a fifth of the arguments.  glibc's logf (2.28 and later,
// sysdeps/ieee754/flt-32/e_logf.c — the algorithm of ARM's optimized routines) is a 16-entry table and a cubic evaluated in
// double and rounded once; the restatement below gave glibc 2.39's result on every positive normal float (2 130 706 432
// arguments, with and without contracted multiply-adds: the contraction never reaches the rounded float).  Zero, subnormal,
// negative, infinite and NaN arguments keep the device libm (green_weight clamps its argument at 1e-2).
SNCH_LBVH_CALLABLE float log_of(float x) noexcept
{
#ifdef __CUDA_ARCH__
    const unsigned int ix = __float_as_uint(x);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return ::logf(x);
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
    const double r = __dsub_rn(__dmul_rn(z, __longlong_as_double((long long)tab[2 * i])), 1.0);
    const double y0 = __dadd_rn(__longlong_as_double((long long)tab[2 * i + 1]), __dmul_rn((double)k, __longlong_as_double(0x3fe62e42fefa39efll)));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(__longlong_as_double(0x3fd5575b0be00b6all), r), __longlong_as_double((long long)0xbfdffffef20a4123ull));
    y = __dadd_rn(__dmul_rn(__longlong_as_double((long long)0xbfd00ea348b88334ull), r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return (float)y;
#else
    return ::logf(x);
#endif
}
SNCH_LBVH_CALLABLE double abs_of(double x) noexcept { return ::fabs(x); }
// std::min / std::max semantics (return the first argument on ties and when a NaN is involved the way `b < a ? b : a` does)
template <typename T> SNCH_LBVH_CALLABLE T min_of(T a, T b) noexcept { return (b < a) ? b : a; }
template <typename T> SNCH_LBVH_CALLABLE T max_of(T a, T b) noexcept { return (a < b) ? b : a; }
} // namespace detail

SNCH_LBVH_CALLABLE float3 vec4_to_vec3(float4 p) { return make_float3(p.x, p.y, p.z); }
SNCH_LBVH_CALLABLE double3 vec4_to_vec3(double4 p) { return make_double3(p.x, p.y, p.z); }
SNCH_LBVH_CALLABLE float4 vec3_to_vec4(float3 p, float w = 0.0f) { return make_float4(p.x, p.y, p.z, w); }
SNCH_LBVH_CALLABLE double4 vec3_to_vec4(double3 p, double w = 0.0) { return make_double4(p.x, p.y, p.z, w); }

// x*x' + y*y' (+ z*z'), accumulated left to right
template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE detail::scalar_of<V> dot(const V &a, const V &b) noexcept
{
    detail::scalar_of<V> s = detail::at(a, 0) * detail::at(b, 0);
    for (int i = 1; i < detail::vec_traits<V>::size; ++i) s = s + detail::at(a, i) * detail::at(b, i);
    return s;
}
SNCH_LBVH_CALLABLE bool inrange(float val, float low, float high) { return val >= low && val <= high; }
SNCH_LBVH_CALLABLE bool inrange(double val, double low, double high) { return val >= low && val <= high; }

template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V cwiseabs(const V &a) noexcept
{
    V r;
    for (int i = 0; i < detail::vec_traits<V>::size; ++i) detail::at(r, i) = detail::abs_of(detail::at(a, i));
    return r;
}
// component-wise fmin / fmax (NaN-suppressing, like the CUDA ::fminf the reference calls)
template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V cwisemin(const V &a, const V &b) noexcept
{
    return detail::map2(a, b, [](detail::scalar_of<V> x, detail::scalar_of<V> y) { return ::fmin(x, y); });
}
template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V cwisemax(const V &a, const V &b) noexcept
{
    return detail::map2(a, b, [](detail::scalar_of<V> x, detail::scalar_of<V> y) { return ::fmax(x, y); });
}
template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE detail::scalar_of<V> squared_length(const V &a) noexcept
{
    return dot(a, a);
}
template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE detail::scalar_of<V> length(const V &a) noexcept
{
    return detail::sqrt_of(dot(a, a));
}
template <typename V, detail::enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V normalize(const V &v)
{
    const detail::scalar_of<V> n = length(v);
    V r;
    for (int i = 0; i < detail::vec_traits<V>::size; ++i) detail::at(r, i) = detail::at(v, i) / n;
    return r;
}
SNCH_LBVH_CALLABLE float3 cross(const float3 &u, const float3 &v) noexcept
{
    return make_float3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
SNCH_LBVH_CALLABLE double3 cross(const double3 &u, const double3 &v) noexcept
{
    return make_double3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
template <typename T> SNCH_LBVH_CALLABLE void lbvh_swap(T &a, T &b)
{
    T t = a;
    a = b;
    b = t;
}
// indexed component access for every CUDA vector type (int/uint/float/double x 2/3/4)
template <typename V, detail::enable_vec<V> = 0> SNCH_LBVH_CALLABLE detail::scalar_of<V> &get(V &vec, int index) { return detail::at(vec, index); }
template <typename V, detail::enable_vec<V> = 0> SNCH_LBVH_CALLABLE detail::scalar_of<V> get(const V &vec, int index) { return detail::at(vec, index); }

namespace detail
{
// small arithmetic helpers used by the geometry headers (not part of the reference's public names)
template <typename V, enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V sub(const V &a, const V &b) noexcept
{
    return map2(a, b, [](scalar_of<V> x, scalar_of<V> y) { return x - y; });
}
template <typename V, enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V add(const V &a, const V &b) noexcept
{
    return map2(a, b, [](scalar_of<V> x, scalar_of<V> y) { return x + y; });
}
template <typename V, enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V scale(const V &a, scalar_of<V> s) noexcept
{
    V r;
    for (int i = 0; i < vec_traits<V>::size; ++i) at(r, i) = at(a, i) * s;
    return r;
}
template <typename V, enable_real_vec<V> = 0> SNCH_LBVH_CALLABLE V splat(scalar_of<V> s) noexcept
{
    V r;
    for (int i = 0; i < vec_traits<V>::size; ++i) at(r, i) = s;
    return r;
}
} // namespace detail

// Which side of the segment p0->p1 (2-D, left normal) / of the plane through p0,p1,p2 (3-D, (p1-p0)x(p2-p0)) a point lies
// on: the sign of the projection onto that normal, 0 only when it is exactly zero (utility.cuh:569-636).
SNCH_LBVH_CALLABLE int checkPointSide(float2 p0, float2 p1, float2 point)
{
    const float2 d = detail::sub(p1, p0), w = detail::sub(point, p0);
    const float c = w.x * -d.y + w.y * d.x;
    return c > 0.0f ? 1 : (c < 0.0f ? -1 : 0);
}
SNCH_LBVH_CALLABLE int checkPointSide(float3 p0, float3 p1, float3 p2, float3 point)
{
    const float3 n = cross(detail::sub(p1, p0), detail::sub(p2, p0));
    const float d = dot(detail::sub(point, p0), n);
    return d > 0.0f ? 1 : (d < 0.0f ? -1 : 0);
}
// Unclamped parameter of the orthogonal projection of `point` onto the line p0->p1 (0 at p0, 1 at p1; 0 when the segment
// is shorter than 1e-4) / unclamped (u, v) of its projection onto the triangle's plane w.r.t. edges p0->p1 and p0->p2
// (utility.cuh:638-707; like there, a degenerate triangle divides by zero).
SNCH_LBVH_CALLABLE float computeProjectionRatio(float2 p0, float2 p1, float2 point)
{
    const float2 d = detail::sub(p1, p0);
    const float l2 = dot(d, d);
    return l2 < 1e-8f ? 0.0f : dot(detail::sub(point, p0), d) / l2;
}
SNCH_LBVH_CALLABLE float2 computeProjectionRatio(float3 p0, float3 p1, float3 p2, float3 point)
{
    const float3 e0 = detail::sub(p1, p0), e1 = detail::sub(p2, p0), w = detail::sub(point, p0);
    const float d00 = dot(e0, e0), d01 = dot(e0, e1), d11 = dot(e1, e1), d20 = dot(w, e0), d21 = dot(w, e1);
    const float den = d00 * d11 - d01 * d01;
    return make_float2((d20 * d11 - d21 * d01) / den, (d21 * d00 - d20 * d01) / den);
}
} // namespace lbvh
#endif // SNCH_LBVH_B200_UTILITY_CUH
