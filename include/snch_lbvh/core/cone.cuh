// snch_lbvh/core/cone.cuh — normal cones of the spatialized normal cone hierarchy (drop-in C++ API).
//
// A cone bounds the normals of every silhouette element below a BVH node: axis (unit), half_angle (radians; < 0 means
// "no silhouette element below", >= pi/2 means "cannot prune") and radius (bound on |element - box centre|).
// `overlap` is the view-cone test of SNCH traversal, `merge` the union used by the bottom-up refit.  Names, member
// layout (20 B in 3-D / 16 B in 2-D: the device arrays of the built tree use it) and float operation order follow the
// reference's core/cone.cuh; one template per operation covers 2-D/3-D and float/double.
#ifndef SNCH_LBVH_B200_CONE_CUH
#define SNCH_LBVH_B200_CONE_CUH
#include "aabb.cuh"

namespace lbvh
{
template <typename T, unsigned int dim> struct cone
{
    vector_of_t<T, dim> axis;
    T half_angle;
    T radius;
};
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE bool is_valid(const cone<T, dim> &c) noexcept { return c.half_angle >= T(0); }

namespace detail
{
template <typename T> SNCH_LBVH_CALLABLE T pi() noexcept { return T(3.14159265358979323846); }
template <typename T> SNCH_LBVH_CALLABLE T half_pi() noexcept { return T(1.57079632679489661923); }
SNCH_LBVH_CALLABLE float acos_of(float x) noexcept { return acosf_host(x); } // glibc's bits on the device too (host_libm.cuh)
SNCH_LBVH_CALLABLE double acos_of(double x) noexcept { return ::acos(x); }
SNCH_LBVH_CALLABLE float asin_of(float x) noexcept { return ::asinf(x); }
SNCH_LBVH_CALLABLE double asin_of(double x) noexcept { return ::asin(x); }
SNCH_LBVH_CALLABLE float atan2_of(float y, float x) noexcept { return ::atan2f(y, x); }
SNCH_LBVH_CALLABLE double atan2_of(double y, double x) noexcept { return ::atan2(y, x); }
SNCH_LBVH_CALLABLE float cos_of(float x) noexcept { return cosf_host(x); }
SNCH_LBVH_CALLABLE double cos_of(double x) noexcept { return ::cos(x); }
SNCH_LBVH_CALLABLE float sin_of(float x) noexcept { return sinf_host(x); }
SNCH_LBVH_CALLABLE double sin_of(double x) noexcept { return ::sin(x); }
// angle between two unit vectors, dot clamped into acos' domain
template <typename V> SNCH_LBVH_CALLABLE scalar_of<V> angle_between(const V &a, const V &b) noexcept
{
    using T = scalar_of<V>;
    return acos_of(max_of(T(-1), min_of(T(1), dot(a, b))));
}
} // namespace detail

// branch-free orthonormal basis of the plane perpendicular to the unit vector n (Duff et al. 2017)       cone.cuh:24-42
template <typename V3, std::enable_if_t<detail::vec_traits<V3>::size == 3, int> = 0>
SNCH_LBVH_CALLABLE void compute_orthonormal_basis(const V3 &n, V3 *b1, V3 *b2) noexcept
{
    using T = detail::scalar_of<V3>;
    const T sign = T(::copysignf(1.0f, float(n.z)));
    const T a = T(-1) / (sign + n.z);
    const T b = n.x * n.y * a;
    b1->x = T(1) + sign * n.x * n.x * a;
    b1->y = sign * b;
    b1->z = -sign * n.x;
    b2->x = b;
    b2->y = sign + n.y * n.y * a;
    b2->z = -n.y;
}
// extent of the box half-diagonal e seen along the unit direction n: |e| projected onto the line (2-D) / plane (3-D)
// perpendicular to n, taken over the box's corners via component-wise absolute basis vectors            cone.cuh:44-76
template <typename V, std::enable_if_t<detail::vec_traits<V>::size == 2, int> = 0>
SNCH_LBVH_CALLABLE detail::scalar_of<V> project_to_plane(const V &n, const V &e) noexcept
{
    V b;
    b.x = -n.y;
    b.y = n.x;
    return detail::abs_of(dot(e, cwiseabs(b)));
}
template <typename V, std::enable_if_t<detail::vec_traits<V>::size == 3, int> = 0>
SNCH_LBVH_CALLABLE detail::scalar_of<V> project_to_plane(const V &n, const V &e) noexcept
{
    V b1, b2;
    compute_orthonormal_basis(n, &b1, &b2);
    const detail::scalar_of<V> r1 = dot(e, cwiseabs(b1)), r2 = dot(e, cwiseabs(b2));
    return detail::sqrt_of(r1 * r1 + r2 * r2);
}

// Can the subtree bounded by (cone bc, box b) contain an element that is a silhouette as seen from o?
// True when the cone cannot prune (half_angle >= pi/2), when o is inside the box (dist_to_box = squared distance < eps),
// or when the plane perpendicular to some direction of the view cone (axis o -> box centre, half-angle = what the box
// subtends) meets the normal cone: pi/2 within [theta - sum, theta + sum].  The angle interval tested last is reported
// through min/max_angle_range.                                                                       cone.cuh:78-264
template <typename T, unsigned int dim>
SNCH_LBVH_CALLABLE bool overlap(const cone<T, dim> &bc, const vector_of_t<T, dim> &o, const aabb<T, dim> &b, const T dist_to_box,
                                T *min_angle_range, T *max_angle_range) noexcept
{
    using V = vector_of_t<T, dim>;
    const T hp = detail::half_pi<T>();
    *min_angle_range = T(0);
    *max_angle_range = hp;
    if (bc.half_angle >= hp || dist_to_box < epsilon<T>()) return true;
    const V c = centroid(b);
    V view = detail::sub(c, o);
    const T l = length(view);
    for (unsigned int i = 0; i < dim; ++i) detail::at(view, i) /= l;
    const T theta = detail::angle_between(bc.axis, view);
    if (inrange(hp, theta - bc.half_angle, theta + bc.half_angle)) return true;
    T view_half;
    if (l > bc.radius) view_half = detail::asin_of(bc.radius / l); // o outside the cone's bounding sphere
    else
    {
        const V e = detail::sub(b.upper, c);
        const T s = l - dot(e, cwiseabs(view));
        if (s <= T(0)) return true;
        view_half = detail::atan2_of(project_to_plane(view, e), s);
    }
    const T sum = bc.half_angle + view_half;
    *min_angle_range = theta - sum;
    *max_angle_range = theta + sum;
    return sum >= hp ? true : inrange(hp, *min_angle_range, *max_angle_range);
}

// rotate u towards v by theta: 2-D about the sign of their cross product, 3-D Rodrigues about normalize(u x v)  cone.cuh:266-318
template <typename V, std::enable_if_t<detail::vec_traits<V>::size == 2, int> = 0>
SNCH_LBVH_CALLABLE V rotate(const V &u, const V &v, detail::scalar_of<V> theta)
{
    using T = detail::scalar_of<V>;
    theta *= T(::copysign(1.0, double(u.x * v.y - u.y * v.x)));
    const T ct = detail::cos_of(theta), st = detail::sin_of(theta);
    V r;
    r.x = ct * u.x - st * u.y;
    r.y = st * u.x + ct * u.y;
    return r;
}
template <typename V, std::enable_if_t<detail::vec_traits<V>::size == 3, int> = 0>
SNCH_LBVH_CALLABLE V rotate(const V &u, const V &v, detail::scalar_of<V> theta)
{
    using T = detail::scalar_of<V>;
    const T ct = detail::cos_of(theta), st = detail::sin_of(theta);
    const V w = normalize(cross(u, v));
    const V k = detail::scale(w, T(1) - ct); // (1 - cos) * w
    V r;
    r.x = (ct + k.x * w.x) * u.x + (k.y * w.x - st * w.z) * u.y + (k.z * w.x + st * w.y) * u.z;
    r.y = (k.x * w.y + st * w.z) * u.x + (ct + k.y * w.y) * u.y + (k.z * w.y - st * w.x) * u.z;
    r.z = (k.x * w.z - st * w.y) * u.x + (k.y * w.z + st * w.x) * u.y + (ct + k.z * w.z) * u.z;
    return r;
}

// Union of two child cones expressed around the parent's box centre.  Invalid children are ignored (both invalid ->
// invalid).  The wider cone absorbs the other when it already covers it; otherwise the axis is rotated to the middle of the
// joint angular range.  A joint range of 2*pi or more yields half_angle = pi ("cannot prune"): the reference leaves the
// field unset there (cone.cuh:454-459) — pi is what the algorithm it was derived from uses.  cone.cuh:320-538
template <typename T, unsigned int dim>
SNCH_LBVH_CALLABLE cone<T, dim> merge(const cone<T, dim> &cone_a, const cone<T, dim> &cone_b, const vector_of_t<T, dim> &origin_a,
                                      const vector_of_t<T, dim> &origin_b, const vector_of_t<T, dim> &new_origin) noexcept
{
    cone<T, dim> ret;
    ret.axis = detail::splat<vector_of_t<T, dim>>(T(0));
    ret.half_angle = -detail::pi<T>();
    ret.radius = T(0);
    const bool va = is_valid(cone_a), vb = is_valid(cone_b);
    if (!(va && vb)) return va ? cone_a : (vb ? cone_b : ret);
    const bool b_wider = cone_b.half_angle > cone_a.half_angle;
    const cone<T, dim> &wide = b_wider ? cone_b : cone_a, &thin = b_wider ? cone_a : cone_b;
    const T ra = cone_a.radius * cone_a.radius + squared_length(detail::sub(new_origin, origin_a));
    const T rb = cone_b.radius * cone_b.radius + squared_length(detail::sub(new_origin, origin_b));
    ret.radius = detail::sqrt_of(detail::max_of(ra, rb));
    ret.axis = wide.axis;
    const T theta = detail::angle_between(wide.axis, thin.axis);
    if (detail::min_of(theta + thin.half_angle, detail::pi<T>()) <= wide.half_angle)
    {
        ret.half_angle = wide.half_angle;
        return ret;
    }
    const T mid = (wide.half_angle + theta + thin.half_angle) / T(2);
    if (mid >= detail::pi<T>())
    {
        ret.half_angle = detail::pi<T>();
        return ret;
    }
    ret.axis = rotate(wide.axis, thin.axis, mid - wide.half_angle);
    ret.half_angle = mid;
    return ret;
}
} // namespace lbvh
#endif // SNCH_LBVH_B200_CONE_CUH
