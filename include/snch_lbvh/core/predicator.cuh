// snch_lbvh/core/predicator.cuh — query descriptors and their factory functions (drop-in C++ API).
//
// Plain aggregates that name WHAT is asked of the tree; query_device() overloads (query.cuh) dispatch on their type.
// Same type names, members and factory spellings as the reference's core/predicator.cuh.
#ifndef SNCH_LBVH_B200_PREDICATOR_CUH
#define SNCH_LBVH_B200_PREDICATOR_CUH
#include "aabb.cuh"

namespace lbvh
{
// every primitive the infinite line l passes through                                                    predicator.cuh:8-41
template <typename Real, unsigned int dim> struct query_line_intersect
{
    using vector_type = vector_of_t<Real, dim>;
    SNCH_LBVH_HOST_DEVICE query_line_intersect(const line<Real, dim> &l_) : l(l_) {}
    line<Real, dim> l;
};
// first (or, with TestOnly, any) primitive hit by ray r within max_dist                                  predicator.cuh:43-79
template <typename Real, unsigned int dim, bool TestOnly> struct query_ray_intersect
{
    using vector_type = vector_of_t<Real, dim>;
    SNCH_LBVH_HOST_DEVICE query_ray_intersect(const ray<Real, dim> &r_, const float max_dist_) : r(r_), max_dist(max_dist_) {}
    ray<Real, dim> r;
    float max_dist;
};
// the sphere a primitive is importance-sampled in (sample_object_in_sphere)                               predicator.cuh:81-113
template <typename Real, unsigned int dim> struct query_sphere_intersect
{
    using vector_type = vector_of_t<Real, dim>;
    SNCH_LBVH_HOST_DEVICE query_sphere_intersect(const sphere<Real, dim> &sph_) : sph(sph_) {}
    sphere<Real, dim> sph;
};
// every leaf whose box overlaps target                                                                  predicator.cuh:115-136
template <typename Real, unsigned int dim> struct query_overlap
{
    SNCH_LBVH_HOST_DEVICE query_overlap(const aabb<Real, dim> &tgt) : target(tgt) {}
    query_overlap() = default;
    SNCH_LBVH_CALLABLE bool operator()(const aabb<Real, dim> &box) noexcept { return intersects(box, target); }
    aabb<Real, dim> target;
};
// nearest primitive to target                                                                           predicator.cuh:138-171
template <typename Real, unsigned int dim> struct query_nearest
{
    using vector_type = vector_of_t<Real, dim>;
    SNCH_LBVH_HOST_DEVICE query_nearest(const vector_type &tgt) : target(tgt) {}
    query_nearest() = default;
    vector_type target;
};
// nearest silhouette element as seen from target; flip_normal_orientation swaps which side counts as front-facing  :173-207
template <typename Real, unsigned int dim> struct query_nearest_silhouette
{
    using vector_type = vector_of_t<Real, dim>;
    SNCH_LBVH_HOST_DEVICE query_nearest_silhouette(const vector_type &tgt, const bool flip) : target(tgt), flip_normal_orientation(flip) {}
    query_nearest_silhouette() = default;
    vector_type target;
    bool flip_normal_orientation;
};

template <typename Real, unsigned int dim> SNCH_LBVH_CALLABLE query_line_intersect<Real, dim> line_intersect(const line<Real, dim> &l) noexcept
{
    return query_line_intersect<Real, dim>(l);
}
template <bool TestOnly = false, typename Real, unsigned int dim>
SNCH_LBVH_CALLABLE query_ray_intersect<Real, dim, TestOnly> ray_intersect(const ray<Real, dim> &r, const float max_dist) noexcept
{
    return query_ray_intersect<Real, dim, TestOnly>(r, max_dist);
}
template <typename Real, unsigned int dim> SNCH_LBVH_CALLABLE query_sphere_intersect<Real, dim> sphere_intersect(const sphere<Real, dim> &s) noexcept
{
    return query_sphere_intersect<Real, dim>(s);
}
template <typename Real, unsigned int dim> SNCH_LBVH_CALLABLE query_overlap<Real, dim> overlaps(const aabb<Real, dim> &region) noexcept
{
    return query_overlap<Real, dim>(region);
}
SNCH_LBVH_CALLABLE query_nearest<float, 2> nearest(const float2 &point) noexcept { return query_nearest<float, 2>(point); }
SNCH_LBVH_CALLABLE query_nearest<double, 2> nearest(const double2 &point) noexcept { return query_nearest<double, 2>(point); }
SNCH_LBVH_CALLABLE query_nearest<float, 3> nearest(const float3 &point) noexcept { return query_nearest<float, 3>(point); }
SNCH_LBVH_CALLABLE query_nearest<double, 3> nearest(const double3 &point) noexcept { return query_nearest<double, 3>(point); }
SNCH_LBVH_CALLABLE query_nearest_silhouette<float, 2> nearest_silhouette(const float2 &p, const bool flip) noexcept { return {p, flip}; }
SNCH_LBVH_CALLABLE query_nearest_silhouette<double, 2> nearest_silhouette(const double2 &p, const bool flip) noexcept { return {p, flip}; }
SNCH_LBVH_CALLABLE query_nearest_silhouette<float, 3> nearest_silhouette(const float3 &p, const bool flip) noexcept { return {p, flip}; }
SNCH_LBVH_CALLABLE query_nearest_silhouette<double, 3> nearest_silhouette(const double3 &p, const bool flip) noexcept { return {p, flip}; }
} // namespace lbvh
#endif // SNCH_LBVH_B200_PREDICATOR_CUH
