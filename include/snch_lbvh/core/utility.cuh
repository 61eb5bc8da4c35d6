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

#include "host_libm.cuh"

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
SNCH_LBVH_CALLABLE float log_of(float x) noexcept { return logf_host(x); } // glibc's bits on the device too (host_libm.cuh)
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
