// snch_lbvh/core/sample.cuh — importance sampling of a primitive inside a sphere (SampleTriangleInSphere) and of a point
// on a primitive (drop-in C++ API; same signatures as the reference's core/sample.cuh).
#ifndef SNCH_LBVH_B200_SAMPLE_CUH
#define SNCH_LBVH_B200_SAMPLE_CUH
#include "bvh.cuh"
#include "predicator.cuh"

namespace lbvh
{
// point on object `object_idx` from two uniforms                                                          sample.cuh:7-21
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename SampleFunc>
SNCH_LBVH_DEVICE vector_of_t<Real, dim> sample_on_object(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh,
                                                         const int object_idx, SampleFunc sample_func, float u, float v)
{
    return sample_func(bvh.objects[object_idx], u, v);
}

// One stochastic root-to-leaf descent driven by the single uniform u: at each node a child is chosen with probability
// proportional to weight(sphere centre, child box centre) among the children whose boxes meet the sphere; u is re-stretched
// to [0,1) after each choice and the path probability accumulates.  If the leaf reached really meets the sphere the
// result is (object index, path probability / measure(object)); otherwise (-1, 0) — the reference leaves the pdf
// unset on a miss (Q19).                                                                                  sample.cuh:23-92
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename SphereIntersectionTestFunc, typename MeasurementFunc,
          typename WeightFunc>
SNCH_LBVH_DEVICE thrust::pair<int, float> sample_object_in_sphere(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh,
                                                                  const query_sphere_intersect<Real, dim> q,
                                                                  SphereIntersectionTestFunc sphere_intersects, MeasurementFunc measure,
                                                                  WeightFunc weight, float u) noexcept
{
    if (bvh.num_objects == 0) return thrust::make_pair(-1, 0.0f);
    std::uint32_t n = 0;
    Real path = Real(1);
    for (;;)
    {
        const auto &nd = bvh.nodes[n];
        if (nd.object_idx != 0xFFFFFFFFu)
        {
            const auto &obj = bvh.objects[nd.object_idx];
            if (!sphere_intersects(q.sph, obj)) return thrust::make_pair(-1, 0.0f);
            float pdf = path;
            pdf /= measure(obj);
            return thrust::make_pair(static_cast<int>(nd.object_idx), pdf);
        }
        const aabb<Real, dim> &lb = bvh.aabbs[nd.left_idx], &rb = bvh.aabbs[nd.right_idx];
        const Real wl = intersect_sphere(q.sph, lb) ? weight(q.sph.origin, centroid(lb)) : 0;
        const Real wr = intersect_sphere(q.sph, rb) ? weight(q.sph.origin, centroid(rb)) : 0;
        const Real total = wl + wr;
        if (!(total > 0)) return thrust::make_pair(-1, 0.0f);
        const Real pl = wl / total;
        if (u < pl)
        {
            u /= pl;
            path = pl * path;
            n = nd.left_idx;
        }
        else
        {
            const Real pr = 1.0f - pl;
            u = (u - pl) / pr;
            path = pr * path;
            n = nd.right_idx;
        }
    }
}
} // namespace lbvh
#endif // SNCH_LBVH_B200_SAMPLE_CUH
