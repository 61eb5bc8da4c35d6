// snch_lbvh/core/aabb.cuh — axis-aligned boxes, lines/rays/spheres and their box tests (drop-in C++ API).
//
// Names, member layout (`upper` before `lower`: 24 B in 3-D, 16 B in 2-D — the device arrays of the built tree use exactly
// this layout) and results follow the reference's core/aabb.cuh; the code is written once per operation over the
// dimension instead of once per vector type.  Cited lines are the reference semantics each routine reproduces.
#ifndef SNCH_LBVH_B200_AABB_CUH
#define SNCH_LBVH_B200_AABB_CUH
#include "utility.cuh"

namespace lbvh
{
template <typename T, unsigned int dim> struct aabb
{
    using vector_type = vector_of_t<T, dim>;
    vector_type upper;
    vector_type lower;

    aabb() = default;
    SNCH_LBVH_HOST_DEVICE aabb(vector_type upper_, vector_type lower_) : upper(upper_), lower(lower_) {}
    // degenerate box around one point, padded by machine epsilon on the first three axes (aabb.cuh:21-50; w stays 0 in 4-D)
    SNCH_LBVH_HOST_DEVICE aabb(const vector_type &p)
    {
        for (unsigned int i = 0; i < dim; ++i)
        {
            const bool pad = i < 3;
            detail::at(upper, i) = pad ? detail::at(p, i) + epsilon<T>() : T(0);
            detail::at(lower, i) = pad ? detail::at(p, i) - epsilon<T>() : T(0);
        }
    }
};

// closed-interval overlap on every axis (aabb.cuh:53-82)
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE bool intersects(const aabb<T, dim> &lhs, const aabb<T, dim> &rhs) noexcept
{
    for (unsigned int i = 0; i < dim; ++i)
        if (detail::at(lhs.upper, i) < detail::at(rhs.lower, i) || detail::at(rhs.upper, i) < detail::at(lhs.lower, i)) return false;
    return true;
}
// grow by the epsilon-padded point (aabb.cuh:84-101)
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE void expand_to_include(aabb<T, dim> *box, const vector_of_t<T, dim> &p)
{
    for (unsigned int i = 0; i < dim; ++i)
    {
        detail::at(box->lower, i) = ::fmin(detail::at(box->lower, i), detail::at(p, i) - epsilon<T>());
        detail::at(box->upper, i) = ::fmax(detail::at(box->upper, i), detail::at(p, i) + epsilon<T>());
    }
}
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE aabb<T, dim> merge(const aabb<T, dim> &lhs, const aabb<T, dim> &rhs) noexcept
{
    aabb<T, dim> m;
    for (unsigned int i = 0; i < dim; ++i)
    {
        detail::at(m.upper, i) = ::fmax(detail::at(lhs.upper, i), detail::at(rhs.upper, i));
        detail::at(m.lower, i) = ::fmin(detail::at(lhs.lower, i), detail::at(rhs.lower, i));
    }
    return m;
}
// SQUARED distance from a point to the box (0 inside), clamp-then-subtract per axis (aabb.cuh:130-158)
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE T mindist(const aabb<T, dim> &lhs, const vector_of_t<T, dim> &rhs) noexcept
{
    T acc = T(0);
    for (unsigned int i = 0; i < dim; ++i)
    {
        const T d = ::fmin(detail::at(lhs.upper, i), ::fmax(detail::at(lhs.lower, i), detail::at(rhs, i))) - detail::at(rhs, i);
        acc = (i == 0) ? d * d : acc + d * d;
    }
    return acc;
}
// Roussopoulos MINMAXDIST (squared): on one axis the nearer face, on the others the farther face (aabb.cuh:160-260)
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE T minmaxdist(const aabb<T, dim> &lhs, const vector_of_t<T, dim> &rhs) noexcept
{
    T nearer[dim], farther[dim];
    for (unsigned int i = 0; i < dim; ++i)
    {
        const T lo = detail::at(lhs.lower, i) - detail::at(rhs, i), hi = detail::at(lhs.upper, i) - detail::at(rhs, i);
        const bool past_centre = (detail::at(lhs.upper, i) + detail::at(lhs.lower, i)) * T(0.5) < detail::at(rhs, i);
        nearer[i] = past_centre ? hi * hi : lo * lo;
        farther[i] = past_centre ? lo * lo : hi * hi;
    }
    T best = T(0);
    for (unsigned int k = 0; k < dim; ++k)
    {
        T s = T(0);
        for (unsigned int i = 0; i < dim; ++i) s = (i == 0) ? (i == k ? nearer[i] : farther[i]) : s + (i == k ? nearer[i] : farther[i]);
        best = (k == 0) ? s : ::fmin(best, s);
    }
    return best;
}
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE vector_of_t<T, dim> centroid(const aabb<T, dim> &box) noexcept
{
    vector_of_t<T, dim> c;
    for (unsigned int i = 0; i < dim; ++i) detail::at(c, i) = (detail::at(box.upper, i) + detail::at(box.lower, i)) * T(0.5);
    return c;
}

namespace detail
{
template <typename T, unsigned int dim> struct directed
{
    vector_of_t<T, dim> origin;
    vector_of_t<T, dim> dir;
    vector_of_t<T, dim> dir_inv; // 1 / dir per axis (inf for axis-parallel directions: the slab test is NaN-robust)
    SNCH_LBVH_HOST_DEVICE directed(const vector_of_t<T, dim> &origin_, const vector_of_t<T, dim> &dir_) : origin(origin_), dir(dir_)
    {
        for (unsigned int i = 0; i < dim; ++i) at(dir_inv, i) = 1 / at(dir_, i);
    }
};
// slab test: parametric entry/exit of origin + t*dir through the box; fmin/fmax drop NaN lanes (0 * inf)
template <typename T, unsigned int dim>
SNCH_LBVH_CALLABLE void slab_interval(const directed<T, dim> &l, const aabb<T, dim> &box, T *tmin, T *tmax) noexcept
{
    for (unsigned int i = 0; i < dim; ++i)
    {
        const T t1 = (at(box.lower, i) - at(l.origin, i)) * at(l.dir_inv, i);
        const T t2 = (at(box.upper, i) - at(l.origin, i)) * at(l.dir_inv, i);
        *tmin = (i == 0) ? ::fmin(t1, t2) : ::fmax(*tmin, ::fmin(t1, t2));
        *tmax = (i == 0) ? ::fmax(t1, t2) : ::fmin(*tmax, ::fmax(t1, t2));
    }
}
} // namespace detail

template <typename T, unsigned int dim> struct line : detail::directed<T, dim>
{
    SNCH_LBVH_HOST_DEVICE line(const vector_of_t<T, dim> &origin, const vector_of_t<T, dim> &dir) : detail::directed<T, dim>(origin, dir) {}
};
template <typename T, unsigned int dim> struct ray : detail::directed<T, dim>
{
    SNCH_LBVH_HOST_DEVICE ray(const vector_of_t<T, dim> &origin, const vector_of_t<T, dim> &dir) : detail::directed<T, dim>(origin, dir) {}
};
template <typename T, unsigned int dim> struct sphere
{
    vector_of_t<T, dim> origin;
    float radius;
    SNCH_LBVH_HOST_DEVICE sphere(const vector_of_t<T, dim> &origin_, const float radius_) : origin(origin_), radius(radius_) {}
};

// infinite line vs box: any parameter, either sign (aabb.cuh:329-364)
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE bool intersects(const line<T, dim> &l, const aabb<T, dim> &box) noexcept
{
    T tmin, tmax;
    detail::slab_interval(l, box, &tmin, &tmax);
    return tmax >= tmin;
}
// ray vs box within max_dist; *distance = entry parameter clamped to >= 0 (aabb.cuh:366-431)
template <typename T, unsigned int dim>
SNCH_LBVH_CALLABLE bool intersects_d(const ray<T, dim> &r, const aabb<T, dim> &box, const T max_dist, T *distance) noexcept
{
    T tmin, tmax;
    detail::slab_interval(r, box, &tmin, &tmax);
    if (!(tmax >= tmin && tmax >= T(0) && tmin <= max_dist)) return false;
    *distance = tmin >= T(0) ? tmin : T(0);
    return true;
}
// closest box point to the sphere centre within the radius (aabb.cuh:433-449)
template <typename T, unsigned int dim> SNCH_LBVH_CALLABLE bool intersect_sphere(const sphere<T, dim> &sph, const aabb<T, dim> &box) noexcept
{
    T d2 = T(0);
    for (unsigned int i = 0; i < dim; ++i)
    {
        const T c = detail::max_of(detail::at(box.lower, i), detail::min_of(detail::at(sph.origin, i), detail::at(box.upper, i)));
        const T d = c - detail::at(sph.origin, i);
        d2 = (i == 0) ? d * d : d2 + d * d;
    }
    return d2 <= T(sph.radius * sph.radius);
}
} // namespace lbvh
#endif // SNCH_LBVH_B200_AABB_CUH
