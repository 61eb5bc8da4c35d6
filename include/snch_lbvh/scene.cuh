// snch_lbvh/scene.cuh — lbvh::scene<2> (line segments / silhouette vertices) and lbvh::scene<3> (triangles / silhouette
// edges) of the drop-in C++ API: the geometry layer with the per-primitive functors that query_device() and
// sample_object_in_sphere() are parameterised with.
//
// Interface kept from the reference's scene.cuh: constructors from vertex / index ranges, compute_silhouettes(),
// build_bvh(), get_bvh_device_ptr() (throws std::runtime_error("BVH is not built yet.")), the public data members, the
// nested primitive types with their member layout (they live in device arrays that user kernels index), and the functor
// types aabb_getter, cone_getter, distance_calculator, silhouette_distance_calculator, intersect_test, intersect_sphere,
// green_weight, measurement_getter, sample_on_object.
//
// What is different underneath:
//   * scene<3> is built by libsnch_b200.so (snch_scene3_* in include/snch_b200.h): O(N) hashed edge adjacency instead of
//     std::map, and the fused sm_100a build pipeline.  All device data lives in ONE arena owned by the library handle;
//     vertices_d / silhouettes_d are non-owning views into it and bvh_dev points at its reference-layout arrays.  The same
//     handle serves the batched query entry points (closest_points(), closest_silhouettes(), intersect(), sample_in_spheres()).
//   * scene<2> evaluates its getters in this header and builds through lbvh::bvh -> snch_lbvh_build (dim = 2); its batched
//     entry points run on the library's 2-D scene (snch_scene2_* in include/snch_b200.h), created on first use.
// Float operation order inside the functors follows the reference so results agree to rounding (DESIGN.md "Parity rules").
#ifndef SNCH_LBVH_B200_SCENE_CUH
#define SNCH_LBVH_B200_SCENE_CUH
#include "lbvh.cuh"

#include <memory>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace lbvh
{
constexpr float bvh_offset = 1e-3f; // absolute padding of 2-D leaf boxes (scene.cuh:12)

// ---- primitive-level geometry ------------------------------------------------------------------------------------------
// uniform point on a triangle from two uniforms, folded across the diagonal                            scene.cuh:14-27
SNCH_LBVH_CALLABLE float3 sample_triangle(const float3 &pa, const float3 &pb, const float3 &pc, float u, float v)
{
    if (u + v > 1.0f)
    {
        u = 1.0f - u;
        v = 1.0f - v;
    }
    const float w = 1.0f - u - v;
    return make_float3(w * pa.x + u * pb.x + v * pc.x, w * pa.y + u * pb.y + v * pc.y, w * pa.z + u * pb.z + v * pc.z);
}
SNCH_LBVH_CALLABLE float2 sample_line(const float2 &pa, const float2 &pb, const float u) // scene.cuh:29-32
{
    return make_float2(pa.x + u * (pb.x - pa.x), pa.y + u * (pb.y - pa.y));
}

namespace detail
{
// shared tail of both silhouette classifiers: signs of the view direction against the two adjacent normals, with a
// dead band of `precision` around zero
SNCH_LBVH_CALLABLE bool opposite_facing(float dot0, float dot1, float sign, float precision)
{
    if (abs_of(dot0) <= precision) return sign * dot1 > precision;
    if (abs_of(dot1) <= precision) return sign * dot0 > precision;
    return dot0 * dot1 < 0.0f;
}
} // namespace detail

// Is the vertex shared by two segments with unit normals n0, n1 a silhouette as seen along view_dir (|view_dir| = d)?
// At (almost) zero distance the sign of the turn decides.                                               scene.cuh:112-141
SNCH_LBVH_CALLABLE bool is_silhouette_vertex(const float2 &n0, const float2 &n1, const float2 &view_dir, float d, bool flip_normal_orientation)
{
    const float precision = 1e-3f;
    const float sign = flip_normal_orientation ? 1.0f : -1.0f;
    if (d <= precision) return sign * (n0.x * n1.y - n0.y * n1.x) > precision;
    const float2 unit = make_float2(view_dir.x / d, view_dir.y / d);
    return detail::opposite_facing(dot(unit, n0), dot(unit, n1), sign, precision);
}
// Same for the edge pa-pb shared by two faces.  The view direction is used UN-normalised, as in the reference
// (scene.cuh:157): the dead band therefore scales with distance.                        scene.cuh:143-174
SNCH_LBVH_CALLABLE bool is_silhouette_edge(const float3 &pa, const float3 &pb, const float3 &n0, const float3 &n1, const float3 &view_dir, float d,
                                           bool flip_normal_orientation)
{
    const float precision = 1e-3f;
    const float sign = flip_normal_orientation ? 1.0f : -1.0f;
    if (d <= precision)
    {
        const float3 edge_dir = normalize(detail::sub(pb, pa));
        return sign * ::atan2f(dot(edge_dir, cross(n0, n1)), dot(n0, n1)) > precision;
    }
    return detail::opposite_facing(dot(view_dir, n0), dot(view_dir, n1), sign, precision);
}

// closest point on segment pa-pb to x: returns the distance, the point and its parameter                 scene.cuh:176-285
template <typename V, detail::enable_real_vec<V> = 0>
SNCH_LBVH_CALLABLE detail::scalar_of<V> find_closest_point_line_segment(const V &pa, const V &pb, const V &x, V *pt, detail::scalar_of<V> *t)
{
    using T = detail::scalar_of<V>;
    const V u = detail::sub(pb, pa);
    const T c1 = dot(u, detail::sub(x, pa));
    if (c1 <= T(0))
    {
        *pt = pa;
        *t = T(0);
        return length(detail::sub(x, pa));
    }
    const T c2 = dot(u, u);
    if (c2 <= c1)
    {
        *pt = pb;
        *t = T(1);
        return length(detail::sub(x, pb));
    }
    *t = c1 / c2;
    for (int i = 0; i < detail::vec_traits<V>::size; ++i) detail::at(*pt, i) = detail::at(pa, i) + detail::at(u, i) * (*t);
    return length(detail::sub(x, *pt));
}

// closest point on triangle pa,pb,pc to x (Ericson, Real-Time Collision Detection 5.1.5): Voronoi region tests in the
// order vertex A, B, C, edge AB, AC, BC, interior.  Returns the distance; *pt the point; *t = (weight of pa, weight of pb).
//                                                                                                         scene.cuh:34-110
SNCH_LBVH_CALLABLE float find_closest_point_triangle(const float3 &pa, const float3 &pb, const float3 &pc, const float3 &x, float3 *pt, float2 *t)
{
    const float3 ab = detail::sub(pb, pa), ac = detail::sub(pc, pa), ax = detail::sub(x, pa);
    auto finish = [&](const float3 &p, float wa, float wb)
    {
        *pt = p;
        t->x = wa;
        t->y = wb;
        return length(detail::sub(x, p));
    };
    const float d1 = dot(ab, ax), d2 = dot(ac, ax);
    if (d1 <= 0.0f && d2 <= 0.0f) return finish(pa, 1.0f, 0.0f);
    const float3 bx = detail::sub(x, pb);
    const float d3 = dot(ab, bx), d4 = dot(ac, bx);
    if (d3 >= 0.0f && d4 <= d3) return finish(pb, 0.0f, 1.0f);
    const float3 cx = detail::sub(x, pc);
    const float d5 = dot(ab, cx), d6 = dot(ac, cx);
    if (d6 >= 0.0f && d5 <= d6) return finish(pc, 0.0f, 0.0f);
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f)
    {
        const float v = d1 / (d1 - d3);
        return finish(make_float3(pa.x + ab.x * v, pa.y + ab.y * v, pa.z + ab.z * v), 1.0f - v, v);
    }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f)
    {
        const float w = d2 / (d2 - d6);
        return finish(make_float3(pa.x + ac.x * w, pa.y + ac.y * w, pa.z + ac.z * w), 1.0f - w, 0.0f);
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f)
    {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        return finish(make_float3(pb.x + (pc.x - pb.x) * w, pb.y + (pc.y - pb.y) * w, pb.z + (pc.z - pb.z) * w), 0.0f, 1.0f - w);
    }
    const float denom = 1.0f / (va + vb + vc);
    const float v = vb * denom, w = vc * denom;
    return finish(make_float3(pa.x + ab.x * v + ac.x * w, pa.y + ab.y * v + ac.y * w, pa.z + ab.z * v + ac.z * w), 1.0f - v - w, v);
}

namespace detail
{
// non-owning view of a device array inside the scene's arena, spelled like the thrust::device_vector members it replaces
// (`v.data().get()`, `v.size()`)
template <typename T> class device_view
{
public:
    struct pointer
    {
        T *p;
        T *get() const noexcept { return p; }
        operator T *() const noexcept { return p; }
    };
    device_view() = default;
    device_view(T *p, std::size_t n) : p_(p), n_(n) {}
    pointer data() const noexcept { return pointer{p_}; }
    std::size_t size() const noexcept { return n_; }
    bool empty() const noexcept { return n_ == 0; }

private:
    T *p_ = nullptr;
    std::size_t n_ = 0;
};

// leaf cone shared by both dimensions (scene.cuh:447-500, 887-961): axis = normalised sum of the owned silhouette
// elements' normals, radius = farthest element centre from the box centre, half-angle = widest deviation of an adjacent
// face normal from the axis; -pi without owned elements, pi when one of them lies on the boundary or the axis degenerates.
template <unsigned int dim, typename Element, int kSlots>
SNCH_LBVH_CALLABLE cone<float, dim> leaf_cone(const aabb<float, dim> &box, const int *owned, const Element *elements)
{
    using V = vector_of_t<float, dim>;
    const V box_centre = centroid(box);
    cone<float, dim> ret;
    ret.axis = splat<V>(0.0f);
    ret.half_angle = pi<float>();
    ret.radius = 0.0f;
    bool any = false, all_two_sided = true;
    for (int i = 0; i < kSlots; ++i)
    {
        if (owned[i] == -1) continue;
        const Element &el = elements[owned[i]];
        ret.axis = add(ret.axis, el.normal());
        ret.radius = max_of(ret.radius, length(sub(el.centroid(), box_centre)));
        all_two_sided = all_two_sided && el.has_face(0) && el.has_face(1);
        any = true;
    }
    if (!any) ret.half_angle = -pi<float>();
    else if (all_two_sided)
    {
        const float norm = length(ret.axis);
        if (norm > epsilon<float>())
        {
            for (unsigned int a = 0; a < dim; ++a) at(ret.axis, a) /= norm;
            ret.half_angle = 0.0f;
            for (int i = 0; i < kSlots; ++i)
            {
                if (owned[i] == -1) continue;
                const Element &el = elements[owned[i]];
                for (int f = 0; f < 2; ++f)
                {
                    const V n = el.has_face(f) ? el.normal(f, true) : splat<V>(0.0f);
                    ret.half_angle = max_of(ret.half_angle, angle_between(ret.axis, n));
                }
            }
        }
    }
    return ret;
}
// nearest silhouette element among the ones a primitive owns, inside a shrinking radius (scene.cuh:518-541, 978-1003)
template <typename Point, typename Element, int kSlots>
SNCH_LBVH_CALLABLE bool nearest_owned_silhouette(const Point &origin, const int *owned, const Element *elements, float max_radius_squared,
                                                 float &distance, bool flip, float min_radius_squared)
{
    bool any = false;
    for (int i = 0; i < kSlots; ++i)
    {
        if (owned[i] == -1) continue;
        if (elements[owned[i]].find_closest_silhouette_point(origin, max_radius_squared, distance, flip, min_radius_squared))
        {
            any = true;
            max_radius_squared = distance * distance;
        }
    }
    return any;
}
SNCH_LBVH_CALLABLE float reciprocal(float x)
{
#ifdef __CUDA_ARCH__
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
} // namespace detail

template <unsigned int dim> class scene;

// =========================================================================================================================
// 2-D: polylines.  Primitive = line segment, silhouette element = vertex between two segments.             scene.cuh:287-703
// =========================================================================================================================
template <> class scene<2>
{
public:
    // indices = (previous vertex, this vertex, next vertex, unused): face 0 is the segment leaving the vertex, face 1 the
    // one arriving; -1 where there is none
    struct silhouette_vertex
    {
        int4 indices;
        const float2 *vertices;
        SNCH_LBVH_HOST_DEVICE silhouette_vertex(const int4 indices_, const float2 *vertices_) : indices(indices_), vertices(vertices_) {}
        SNCH_LBVH_HOST_DEVICE silhouette_vertex() : indices(make_int4(-1, -1, -1, -1)), vertices(nullptr) {}

        SNCH_LBVH_HOST_DEVICE aabb<float, 2> bounding_box() const
        {
            const float2 p = vertices[indices.y];
            return aabb<float, 2>(make_float2(p.x + bvh_offset, p.y + bvh_offset), make_float2(p.x - bvh_offset, p.y - bvh_offset));
        }
        SNCH_LBVH_HOST_DEVICE float2 centroid() const { return vertices[indices.y]; }
        SNCH_LBVH_HOST_DEVICE bool has_face(int f_index) const { return f_index == 0 ? indices.z != -1 : indices.x != -1; }
        // right-hand normal of the adjacent segment (face 0: this->next, face 1: previous->this)
        SNCH_LBVH_HOST_DEVICE float2 normal(int f_index, bool do_normalize = true) const
        {
            const int i = f_index == 0 ? 1 : 0;
            const float2 pa = vertices[get(indices, i)], pb = vertices[get(indices, i + 1)];
            const float2 n = make_float2(pb.y - pa.y, -(pb.x - pa.x));
            return do_normalize ? normalize(n) : n;
        }
        SNCH_LBVH_HOST_DEVICE float2 normal() const
        {
            float2 n = make_float2(0.0f, 0.0f);
            for (int f = 0; f < 2; ++f)
                if (has_face(f)) n = detail::add(n, normal(f, false));
            return normalize(n);
        }
        SNCH_LBVH_HOST_DEVICE bool find_closest_silhouette_point(const float2 origin, const float max_radius_squared, float &distance,
                                                                 const bool flip_normal_orientation, const float min_radius_squared) const
        {
            if (min_radius_squared >= max_radius_squared) return false;
            const float2 view_dir = detail::sub(origin, vertices[indices.y]);
            const float d = length(view_dir);
            if (d * d > max_radius_squared) return false;
            bool is_silhouette = !has_face(0) || !has_face(1);
            if (!is_silhouette) is_silhouette = is_silhouette_vertex(normal(0), normal(1), view_dir, d, flip_normal_orientation);
            if (!is_silhouette) return false;
            distance = d;
            return true;
        }
    };

    struct line_segment
    {
        int2 vertex_indices;
        int2 silhouette_indices; // the end vertices this segment "owns" (first segment in input order that touches them)
        const float2 *vertices;
        const silhouette_vertex *silhouettes;
        SNCH_LBVH_HOST_DEVICE line_segment(const int2 vertex_indices_, const int2 silhouette_indices_, const float2 *vertices_,
                                           const silhouette_vertex *silhouettes_)
            : vertex_indices(vertex_indices_), silhouette_indices(silhouette_indices_), vertices(vertices_), silhouettes(silhouettes_)
        {
        }
        line_segment() = default;
        SNCH_LBVH_HOST_DEVICE float2 normal() const
        {
            const float2 s = detail::sub(vertices[vertex_indices.y], vertices[vertex_indices.x]);
            return normalize(make_float2(-s.y, s.x));
        }
    };

    struct measurement_getter
    {
        SNCH_LBVH_HOST_DEVICE float operator()(const line_segment &o) const noexcept
        {
            return length(detail::sub(o.vertices[o.vertex_indices.x], o.vertices[o.vertex_indices.y]));
        }
    };
    struct aabb_getter
    {
        SNCH_LBVH_HOST_DEVICE lbvh::aabb<float, 2> operator()(const line_segment &ls) const noexcept
        {
            const float2 p0 = ls.vertices[ls.vertex_indices.x], p1 = ls.vertices[ls.vertex_indices.y];
            return lbvh::aabb<float, 2>(make_float2(detail::max_of(p0.x, p1.x) + bvh_offset, detail::max_of(p0.y, p1.y) + bvh_offset),
                                        make_float2(detail::min_of(p0.x, p1.x) - bvh_offset, detail::min_of(p0.y, p1.y) - bvh_offset));
        }
    };
    struct cone_getter
    {
        SNCH_LBVH_HOST_DEVICE lbvh::cone<float, 2> operator()(const line_segment &ls) const noexcept
        {
            const int owned[2] = {ls.silhouette_indices.x, ls.silhouette_indices.y};
            return detail::leaf_cone<2, silhouette_vertex, 2>(aabb_getter()(ls), owned, ls.silhouettes);
        }
    };
    struct distance_calculator
    {
        SNCH_LBVH_HOST_DEVICE float operator()(const float2 point, const line_segment &o) const noexcept
        {
            float2 pt;
            float t;
            return find_closest_point_line_segment(o.vertices[o.vertex_indices.x], o.vertices[o.vertex_indices.y], point, &pt, &t);
        }
    };
    struct silhouette_distance_calculator
    {
        SNCH_LBVH_HOST_DEVICE bool operator()(const float2 origin, const line_segment &o, const float max_radius_squared, float &distance,
                                              const bool flip_normal_orientation, const float min_radius_squared) const noexcept
        {
            const int owned[2] = {o.silhouette_indices.x, o.silhouette_indices.y};
            return detail::nearest_owned_silhouette<float2, silhouette_vertex, 2>(origin, owned, o.silhouettes, max_radius_squared, distance,
                                                                                 flip_normal_orientation, min_radius_squared);
        }
    };
    // ray vs segment by Cramer's rule; returns (hit, ray parameter t, segment parameter s), s accepted in [-1e-3, 1+1e-3]
    struct intersect_test
    {
        SNCH_LBVH_HOST_DEVICE thrust::tuple<bool, float, float> operator()(const ray<float, 2> &r, const line_segment &o) const noexcept
        {
            const float2 p0 = o.vertices[o.vertex_indices.x], p1 = o.vertices[o.vertex_indices.y];
            const float2 seg = detail::sub(p1, p0);
            const float D = r.dir.x * (-seg.y) + r.dir.y * seg.x;
            if (detail::abs_of(D) < epsilon<float>()) return thrust::make_tuple(false, 0.0f, 0.0f);
            const float inv = detail::reciprocal(D);
            const float2 w = detail::sub(p0, r.origin);
            const float t = (w.x * (-seg.y) - w.y * (-seg.x)) * inv;
            const float s = (r.dir.x * w.y - r.dir.y * w.x) * inv;
            if (s >= -1e-3f && s <= 1.0f + 1e-3f && t >= 0.0f) return thrust::make_tuple(true, t, s);
            return thrust::make_tuple(false, 0.0f, 0.0f);
        }
    };
    struct intersect_sphere
    {
        SNCH_LBVH_HOST_DEVICE bool operator()(const sphere<float, 2> &sph, const line_segment &o) const noexcept
        {
            const float2 p1 = o.vertices[o.vertex_indices.x], d = detail::sub(o.vertices[o.vertex_indices.y], p1);
            float t = ((sph.origin.x - p1.x) * d.x + (sph.origin.y - p1.y) * d.y) / (d.x * d.x + d.y * d.y);
            t = detail::max_of(0.0f, detail::min_of(1.0f, t));
            const float dx = (p1.x + t * d.x) - sph.origin.x, dy = (p1.y + t * d.y) - sph.origin.y;
            return dx * dx + dy * dy <= sph.radius * sph.radius;
        }
    };
    struct green_weight // |ln r| / 2 pi with r clamped at 1e-2
    {
        SNCH_LBVH_HOST_DEVICE float operator()(const float2 &x, const float2 &y) const noexcept
        {
            const float r = detail::max_of(length(detail::sub(x, y)), 1e-2f);
            return detail::abs_of(detail::log_of(r) / (detail::pi<float>() * 2.0f));
        }
    };
    struct sample_on_object
    {
        SNCH_LBVH_HOST_DEVICE float2 operator()(const line_segment &o, float u, float) { return sample_line(o.vertices[o.vertex_indices.x], o.vertices[o.vertex_indices.y], u); }
    };

    using bvh_type = lbvh::bvh<float, 2, line_segment, aabb_getter, cone_getter>;

    scene() = default;
    template <typename VerticesInputIterator, typename IndicesInputIterator>
    scene(VerticesInputIterator vertices_first, VerticesInputIterator vertices_last, IndicesInputIterator indices_first, IndicesInputIterator indices_last)
        : vertices_h(vertices_first, vertices_last), indices_h(indices_first, indices_last)
    {
        vertices_store_.upload(vertices_h.size() ? &vertices_h[0] : nullptr, vertices_h.size());
        vertices_d = detail::device_view<float2>(vertices_store_.data(), vertices_store_.size());
    }

    // one silhouette record per VERTEX: (previous, self, next) from the segments that touch it            scene.cuh:634-656
    void compute_silhouettes()
    {
        silhouettes_h.clear();
        silhouettes_h.resize(vertices_h.size(), silhouette_vertex(make_int4(-1, -1, -1, -1), vertices_store_.data()));
        for (const int2 &seg : indices_h)
        {
            silhouette_vertex &from = silhouettes_h[seg.x];
            from.indices.y = seg.x;
            from.indices.z = seg.y;
            silhouette_vertex &to = silhouettes_h[seg.y];
            to.indices.x = seg.x;
            to.indices.y = seg.y;
        }
        silhouettes_store_.upload(silhouettes_h.size() ? &silhouettes_h[0] : nullptr, silhouettes_h.size());
        silhouettes_d = detail::device_view<silhouette_vertex>(silhouettes_store_.data(), silhouettes_store_.size());
    }
    // a vertex is owned by the first segment (input order) that touches it                                 scene.cuh:658-681
    void build_bvh()
    {
        std::vector<char> seen(vertices_h.size(), 0);
        lines.clear();
        for (const int2 &seg : indices_h)
        {
            int2 owned = make_int2(-1, -1);
            int k = 0;
            for (int i = 0; i < 2; ++i)
            {
                const int v = get(seg, i);
                if (!seen[v])
                {
                    seen[v] = 1;
                    get(owned, k++) = v;
                }
            }
            lines.push_back(line_segment(seg, owned, vertices_store_.data(), silhouettes_store_.data()));
        }
        p_bvh = std::make_unique<bvh_type>(lines.begin(), lines.end(), true);
        bvh_dev = p_bvh->get_device_repr();
        batched_.reset(); // the batched entry points rebuild their records from the new geometry on next use
    }
    const lbvh::bvh_device<float, 2, line_segment> &get_bvh_device_ptr() const
    {
        if (!p_bvh) throw std::runtime_error("BVH is not built yet.");
        return bvh_dev;
    }

    // ---- batched queries (one launch per call; device OR host pointers, see include/snch_b200.h "2-D scenes") ----------------
    // Each is the batched form of the per-thread call named next to it and returns the same values per query.  They run on
    // the library's own 2-D scene (fused two-child records, persistent kernels), created from this scene's vertices and
    // segments the first time one of them is called after build_bvh().
    // query_device(bvh_dev, nearest(p), distance_calculator())
    void closest_points(const float2 *points, std::size_t n, unsigned int *out_index, float *out_distance, cudaStream_t stream = nullptr) const
    {
        detail::check_status(snch_closest_point_batch2(batched(stream), &points->x, n, out_index, out_distance, stream));
    }
    // query_device(bvh_dev, nearest_silhouette(p, flip), silhouette_distance_calculator()); r_max optional search radii
    // out_vertex / out_point (optional): index in silhouettes_d of a silhouette vertex attaining the distance, and its position
    void closest_silhouettes(const float2 *points, const unsigned char *flip, const float *r_max, std::size_t n, float *out_distance,
                             cudaStream_t stream = nullptr, unsigned int *out_vertex = nullptr, float2 *out_point = nullptr) const
    {
        detail::check_status(snch_closest_silhouette_batch2(batched(stream), &points->x, flip, r_max, n, out_distance, out_vertex,
                                                            out_point ? &out_point->x : nullptr, stream));
    }
    // query_device(bvh_dev, ray_intersect<any_hit>(ray(o, d), t_max), intersect_test()); snch_hit = {t, s, 0, segment}
    void intersect(const float2 *origins, const float2 *directions, const float *t_max, std::size_t n, snch_hit *out_hits, unsigned char *out_found,
                   bool any_hit = false, cudaStream_t stream = nullptr) const
    {
        detail::check_status(snch_intersect_batch2(batched(stream), &origins->x, &directions->x, t_max, n, out_hits, out_found, any_hit ? 1 : 0, stream));
    }
    // sample_object_in_sphere(...) followed by sample_on_object(...); circles = (x, y, radius), rnd = (u, u1) per query
    void sample_in_spheres(const float3 *circles, const float2 *rnd, std::size_t n, int *out_index, float *out_pdf, float2 *out_point,
                           cudaStream_t stream = nullptr) const
    {
        detail::check_status(snch_sample_in_sphere_batch2(batched(stream), &circles->x, &rnd->x, n, out_index, out_pdf, out_point ? &out_point->x : nullptr, stream));
    }

public:
    thrust::host_vector<float2> vertices_h;
    detail::device_view<float2> vertices_d;
    thrust::host_vector<int2> indices_h;
    std::vector<line_segment> lines;
    thrust::host_vector<silhouette_vertex> silhouettes_h;
    detail::device_view<silhouette_vertex> silhouettes_d;
    std::unique_ptr<bvh_type> p_bvh;
    lbvh::bvh_device<float, 2, line_segment> bvh_dev;

private:
    struct handle2_deleter
    {
        void operator()(snch_scene2 *s) const noexcept { snch_scene2_destroy(s); }
    };
    const snch_scene2 *batched(cudaStream_t stream) const
    {
        if (!p_bvh) throw std::runtime_error("BVH is not built yet.");
        if (!batched_)
        {
            snch_scene2 *h = nullptr;
            detail::check_status(snch_scene2_create(vertices_h.size() ? &vertices_h[0].x : nullptr, static_cast<std::uint32_t>(vertices_h.size()),
                                                    indices_h.size() ? &indices_h[0].x : nullptr, static_cast<std::uint32_t>(indices_h.size()), 0, &h));
            batched_.reset(h);
            detail::check_status(snch_scene2_compute_silhouettes(h));
            detail::check_status(snch_scene2_build(h, stream));
        }
        return batched_.get();
    }
    detail::device_buffer<float2> vertices_store_;
    detail::device_buffer<silhouette_vertex> silhouettes_store_;
    mutable std::unique_ptr<snch_scene2, handle2_deleter> batched_;
};

// =========================================================================================================================
// 3-D: triangle meshes.  Primitive = triangle, silhouette element = edge between two faces.               scene.cuh:705-1268
// =========================================================================================================================
template <> class scene<3>
{
public:
    // indices = (vertex opposite the edge in the face that runs it high->low, lower end, higher end, vertex opposite in the
    // face that runs it low->high); -1 where that face does not exist
    struct silhouette_edge
    {
        int4 indices;
        const float3 *vertices;
        SNCH_LBVH_HOST_DEVICE silhouette_edge(const int4 indices_, const float3 *vertices_) : indices(indices_), vertices(vertices_) {}
        SNCH_LBVH_HOST_DEVICE silhouette_edge() : indices(make_int4(-1, -1, -1, -1)), vertices(nullptr) {}

        SNCH_LBVH_HOST_DEVICE aabb<float, 3> bounding_box() const
        {
            aabb<float, 3> box(vertices[indices.y]);
            expand_to_include(&box, vertices[indices.z]);
            return box;
        }
        SNCH_LBVH_HOST_DEVICE float3 centroid() const
        {
            const float3 pa = vertices[indices.y], pb = vertices[indices.z];
            return make_float3((pa.x + pb.x) / 2, (pa.y + pb.y) / 2, (pa.z + pb.z) / 2);
        }
        SNCH_LBVH_HOST_DEVICE bool has_face(int f_index) const { return f_index == 0 ? indices.w != -1 : indices.x != -1; }
        // face 0 = (lower, higher, opposite w), face 1 = (higher, lower, opposite x): both counter-clockwise
        SNCH_LBVH_HOST_DEVICE float3 normal(int f_index, bool do_normalize = true) const
        {
            const float3 pa = vertices[f_index == 0 ? indices.y : indices.z];
            const float3 pb = vertices[f_index == 0 ? indices.z : indices.y];
            const float3 pc = vertices[f_index == 0 ? indices.w : indices.x];
            const float3 n = cross(detail::sub(pb, pa), detail::sub(pc, pa));
            return do_normalize ? normalize(n) : n;
        }
        SNCH_LBVH_HOST_DEVICE float3 normal() const // area-weighted mean of the adjacent face normals
        {
            float3 n = make_float3(0.0f, 0.0f, 0.0f);
            for (int f = 0; f < 2; ++f)
                if (has_face(f)) n = detail::add(n, normal(f, false));
            return normalize(n);
        }
        SNCH_LBVH_HOST_DEVICE bool find_closest_silhouette_point(const float3 origin, const float max_radius_squared, float &distance,
                                                                 const bool flip_normal_orientation, const float min_radius_squared) const
        {
            if (min_radius_squared >= max_radius_squared) return false;
            const float3 pa = vertices[indices.y], pb = vertices[indices.z];
            float3 closest;
            float t;
            const float d = find_closest_point_line_segment(pa, pb, origin, &closest, &t);
            if (d * d > max_radius_squared) return false;
            bool is_silhouette = !has_face(0) || !has_face(1); // boundary edges always are
            if (!is_silhouette)
                is_silhouette = is_silhouette_edge(pa, pb, normal(0), normal(1), detail::sub(origin, closest), d, flip_normal_orientation);
            if (!is_silhouette) return false;
            distance = d;
            return true;
        }
    };

    struct triangle
    {
        int3 vertex_indices;
        int3 silhouette_indices; // the edges this triangle "owns" (first triangle in input order that references them)
        const float3 *vertices;
        const silhouette_edge *silhouettes;
        SNCH_LBVH_HOST_DEVICE triangle(const int3 vertex_indices_, const int3 silhouette_indices_, const float3 *vertices_,
                                       const silhouette_edge *silhouettes_)
            : vertex_indices(vertex_indices_), silhouette_indices(silhouette_indices_), vertices(vertices_), silhouettes(silhouettes_)
        {
        }
        triangle() = default;
        SNCH_LBVH_HOST_DEVICE float3 normal() const
        {
            const float3 pa = vertices[vertex_indices.x];
            return normalize(cross(detail::sub(vertices[vertex_indices.z], pa), detail::sub(vertices[vertex_indices.y], pa)));
        }
    };
    static_assert(sizeof(silhouette_edge) == 32 && sizeof(triangle) == 40, "the library's arena uses these record layouts");

    struct measurement_getter // area
    {
        SNCH_LBVH_HOST_DEVICE float operator()(const triangle &o) const noexcept
        {
            const float3 pa = o.vertices[o.vertex_indices.x];
            return length(cross(detail::sub(o.vertices[o.vertex_indices.z], pa), detail::sub(o.vertices[o.vertex_indices.y], pa))) / 2;
        }
    };
    struct aabb_getter
    {
        SNCH_LBVH_HOST_DEVICE lbvh::aabb<float, 3> operator()(const triangle &tri) const noexcept
        {
            lbvh::aabb<float, 3> box(tri.vertices[tri.vertex_indices.x]);
            expand_to_include(&box, tri.vertices[tri.vertex_indices.y]);
            expand_to_include(&box, tri.vertices[tri.vertex_indices.z]);
            return box;
        }
    };
    struct cone_getter
    {
        SNCH_LBVH_HOST_DEVICE lbvh::cone<float, 3> operator()(const triangle &tri) const noexcept
        {
            const int owned[3] = {tri.silhouette_indices.x, tri.silhouette_indices.y, tri.silhouette_indices.z};
            return detail::leaf_cone<3, silhouette_edge, 3>(aabb_getter()(tri), owned, tri.silhouettes);
        }
    };
    struct distance_calculator
    {
        SNCH_LBVH_HOST_DEVICE float operator()(const float3 point, const triangle &o) const noexcept
        {
            float3 pt;
            float2 t;
            return find_closest_point_triangle(o.vertices[o.vertex_indices.x], o.vertices[o.vertex_indices.y], o.vertices[o.vertex_indices.z], point,
                                               &pt, &t);
        }
    };
    struct silhouette_distance_calculator
    {
        SNCH_LBVH_HOST_DEVICE bool operator()(const float3 origin, const triangle &o, const float max_radius_squared, float &distance,
                                              const bool flip_normal_orientation, const float min_radius_squared) const noexcept
        {
            const int owned[3] = {o.silhouette_indices.x, o.silhouette_indices.y, o.silhouette_indices.z};
            return detail::nearest_owned_silhouette<float3, silhouette_edge, 3>(origin, owned, o.silhouettes, max_radius_squared, distance,
                                                                               flip_normal_orientation, min_radius_squared);
        }
    };
    // Moeller-Trumbore; returns (hit, t, (u, v)); degenerate when |det| < eps; u, v in [0,1], u+v <= 1, t >= 0   scene.cuh:1005-1052
    struct intersect_test
    {
        SNCH_LBVH_HOST_DEVICE thrust::tuple<bool, float, float2> operator()(const ray<float, 3> &r, const triangle &o) const noexcept
        {
            const auto miss = thrust::make_tuple(false, 0.0f, make_float2(0.0f, 0.0f));
            const float3 v0 = o.vertices[o.vertex_indices.x];
            const float3 e1 = detail::sub(o.vertices[o.vertex_indices.y], v0), e2 = detail::sub(o.vertices[o.vertex_indices.z], v0);
            const float3 h = cross(r.dir, e2);
            const float det = dot(e1, h);
            if (detail::abs_of(det) < epsilon<float>()) return miss;
            const float inv_det = detail::reciprocal(det);
            const float3 s = detail::sub(r.origin, v0);
            const float u = dot(s, h) * inv_det;
            if (u < 0.0f || u > 1.0f) return miss;
            const float3 q = cross(s, e1);
            const float v = dot(r.dir, q) * inv_det;
            if (v < 0.0f || u + v > 1.0f) return miss;
            const float t = dot(e2, q) * inv_det;
            if (t >= 0.0f) return thrust::make_tuple(true, t, make_float2(u, v));
            return miss;
        }
    };
    // Sphere vs triangle: project the centre on the plane; inside the triangle -> plane distance decides, otherwise the
    // distance to ONE vertex chosen by the sign of the barycentrics (never an edge point: the reference's approximation,
    // SURVEY quirk Q10, kept because sample_object_in_sphere's pdf depends on it).                        scene.cuh:1054-1117
    struct intersect_sphere
    {
        SNCH_LBVH_HOST_DEVICE bool operator()(const sphere<float, 3> &sph, const triangle &o) const noexcept
        {
            const float3 p1 = o.vertices[o.vertex_indices.x], p2 = o.vertices[o.vertex_indices.y], p3 = o.vertices[o.vertex_indices.z];
            const float3 c = sph.origin;
            const float3 n = normalize(cross(detail::sub(p2, p1), detail::sub(p3, p1)));
            const float plane_d = dot(n, p1);
            const float dist_to_plane = dot(n, c) - plane_d;
            const float3 proj = make_float3(c.x - dist_to_plane * n.x, c.y - dist_to_plane * n.y, c.z - dist_to_plane * n.z);
            const float3 v0 = detail::sub(p3, p1), v1 = detail::sub(p2, p1), v2 = detail::sub(proj, p1);
            const float d00 = dot(v0, v0), d01 = dot(v0, v1), d02 = dot(v0, v2), d11 = dot(v1, v1), d12 = dot(v1, v2);
            const float inv = 1.0f / (d00 * d11 - d01 * d01);
            const float u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
            if (u >= 0 && v >= 0 && u + v <= 1) return detail::abs_of(dist_to_plane) <= sph.radius;
            const float3 nearest = u < 0 ? p1 : (v < 0 ? p3 : p2);
            const float3 d = detail::sub(nearest, c);
            return d.x * d.x + d.y * d.y + d.z * d.z <= sph.radius * sph.radius;
        }
    };
    struct green_weight // 1 / (4 pi r) with r clamped at 1e-4
    {
        SNCH_LBVH_HOST_DEVICE float operator()(const float3 &x, const float3 &y) const noexcept
        {
            const float r = detail::max_of(length(detail::sub(x, y)), 1e-4f);
            return 1.0f / (detail::pi<float>() * 4.0f * r);
        }
    };
    struct sample_on_object
    {
        SNCH_LBVH_HOST_DEVICE float3 operator()(const triangle &o, float u, float v)
        {
            return sample_triangle(o.vertices[o.vertex_indices.x], o.vertices[o.vertex_indices.y], o.vertices[o.vertex_indices.z], u, v);
        }
    };

    // What p_bvh points at once the tree is built: the reference exposes a lbvh::bvh object here; this one forwards to the
    // library handle and downloads host mirrors on demand instead of on every build (the reference forces three
    // device-to-host copies of the whole tree per build).
    class built_tree
    {
    public:
        using node_type = detail::node;
        using aabb_type = aabb<float, 3>;
        using cone_type = cone<float, 3>;
        explicit built_tree(scene<3> *owner) : owner_(owner) {}
        bool query_host_enabled() const noexcept { return true; }
        lbvh::bvh_device<float, 3, triangle> get_device_repr() const noexcept { return owner_->bvh_dev; }
        const thrust::host_vector<node_type> &nodes_host() { return fetch(nodes_h_, SNCH_EXPORT_NODES); }
        const thrust::host_vector<aabb_type> &aabbs_host() { return fetch(aabbs_h_, SNCH_EXPORT_AABBS); }
        const thrust::host_vector<cone_type> &cones_host() { return fetch(cones_h_, SNCH_EXPORT_CONES); }
        const std::vector<triangle> &objects_host() const noexcept { return owner_->triangles; }

    private:
        template <typename T> const thrust::host_vector<T> &fetch(thrust::host_vector<T> &dst, int kind)
        {
            const std::size_t n = owner_->bvh_dev.num_nodes;
            if (dst.size() != n)
            {
                dst.resize(n);
                if (n) detail::check_status(snch_scene_export(owner_->handle_.get(), kind, &dst[0], n * sizeof(T)));
            }
            return dst;
        }
        scene<3> *owner_;
        thrust::host_vector<node_type> nodes_h_;
        thrust::host_vector<aabb_type> aabbs_h_;
        thrust::host_vector<cone_type> cones_h_;
    };

    scene() = default;
    template <typename VerticesInputIterator, typename IndicesInputIterator>
    scene(VerticesInputIterator vertices_first, VerticesInputIterator vertices_last, IndicesInputIterator indices_first, IndicesInputIterator indices_last,
          int device = 0)
        : vertices_h(vertices_first, vertices_last), indices_h(indices_first, indices_last)
    {
        snch_scene *h = nullptr;
        detail::check_status(snch_scene3_create(vertices_h.size() ? &vertices_h[0].x : nullptr, static_cast<std::uint32_t>(vertices_h.size()),
                                                indices_h.size() ? &indices_h[0].x : nullptr, static_cast<std::uint32_t>(indices_h.size()), device, &h));
        handle_.reset(h);
    }
    scene(const scene &) = delete; // owns a device arena; movable only
    scene &operator=(const scene &) = delete;
    scene(scene &&o) noexcept { *this = std::move(o); }
    scene &operator=(scene &&o) noexcept
    {
        vertices_h = std::move(o.vertices_h), indices_h = std::move(o.indices_h), edge_indices_h = std::move(o.edge_indices_h);
        triangles = std::move(o.triangles), silhouettes_h = std::move(o.silhouettes_h);
        vertices_d = o.vertices_d, silhouettes_d = o.silhouettes_d, bvh_dev = o.bvh_dev;
        handle_ = std::move(o.handle_);
        p_bvh.reset(o.p_bvh ? new built_tree(this) : nullptr);
        o.p_bvh.reset();
        return *this;
    }

    // Edge ids in first-seen order over the triangles' (a,b),(b,c),(c,a) edges; returns the number of edges.  scene.cuh:1135-1166
    int assign_edge_indices()
    {
        require_handle();
        detail::check_status(snch_scene_compute_silhouettes(handle_.get()));
        snch_build_stats st;
        detail::check_status(snch_scene_stats(handle_.get(), &st));
        edge_indices_h.resize(indices_h.size());
        if (!indices_h.empty())
            detail::check_status(snch_scene_export(handle_.get(), SNCH_EXPORT_TRI_EDGES, &edge_indices_h[0], indices_h.size() * sizeof(int3)));
        return static_cast<int>(st.num_edges);
    }
    // Per edge the int4 (opposite, low, high, opposite) with the reference's orientation rule (a later face overwrites an
    // earlier one on non-manifold input, quirk Q18).  Host mirror silhouettes_h; the device copy is made by build_bvh().  :1167-1204
    void compute_silhouettes()
    {
        const int n_edges = assign_edge_indices();
        std::vector<int4> packed(static_cast<std::size_t>(n_edges));
        if (n_edges) detail::check_status(snch_scene_export(handle_.get(), SNCH_EXPORT_EDGES, packed.data(), packed.size() * sizeof(int4)));
        silhouettes_h.resize(packed.size());
        for (std::size_t e = 0; e < packed.size(); ++e) silhouettes_h[e] = silhouette_edge(packed[e], nullptr); // pointers: see build_bvh()
    }
    // First-owner assignment of edges to triangles, upload, and the device build.                             scene.cuh:1205-1229
    void build_bvh(cudaStream_t stream = nullptr)
    {
        require_handle();
        detail::check_status(snch_scene_build(handle_.get(), nullptr, stream));
        snch_bvh_device_pod pod;
        detail::check_status(snch_scene_device_repr(handle_.get(), &pod));
        const float3 *dv = static_cast<const float3 *>(pod.vertices);
        const silhouette_edge *de = static_cast<const silhouette_edge *>(pod.silhouettes);
        vertices_d = detail::device_view<const float3>(dv, pod.num_vertices);
        silhouettes_d = detail::device_view<const silhouette_edge>(de, pod.num_silhouettes);
        bvh_dev = lbvh::bvh_device<float, 3, triangle>(pod.num_nodes, pod.num_objects, static_cast<detail::node *>(pod.nodes),
                                                       static_cast<aabb<float, 3> *>(pod.aabbs), static_cast<cone<float, 3> *>(pod.cones),
                                                       static_cast<triangle *>(pod.objects));
        for (auto &se : silhouettes_h) se.vertices = dv;
        std::vector<int3> owned(indices_h.size());
        if (!owned.empty()) detail::check_status(snch_scene_export(handle_.get(), SNCH_EXPORT_TRI_OWNED, owned.data(), owned.size() * sizeof(int3)));
        triangles.clear();
        triangles.reserve(owned.size());
        for (std::size_t i = 0; i < owned.size(); ++i) triangles.push_back(triangle(indices_h[i], owned[i], dv, de));
        p_bvh.reset(new built_tree(this));
    }
    const lbvh::bvh_device<float, 3, triangle> &get_bvh_device_ptr() const
    {
        if (!p_bvh) throw std::runtime_error("BVH is not built yet.");
        return bvh_dev;
    }

    // ---- batched queries (one launch per call; device OR host pointers, see include/snch_b200.h) ----------------------------
    // Each is the batched form of the per-thread call named next to it and returns the same values per query.
    // query_device(bvh_dev, nearest(p), distance_calculator())
    void closest_points(const float3 *points, std::size_t n, unsigned int *out_index, float *out_distance, cudaStream_t stream = nullptr) const
    {
        detail::check_status(snch_closest_point_batch(built(), &points->x, n, out_index, out_distance, stream));
    }
    // query_device(bvh_dev, nearest_silhouette(p, flip), silhouette_distance_calculator()); r_max optional search radii
    // out_edge / out_point (optional): index in silhouettes_d of an edge attaining the distance and the closest point on it —
    // the values silhouette_edge::find_closest_silhouette_point computes and the reference drops
    void closest_silhouettes(const float3 *points, const unsigned char *flip, const float *r_max, std::size_t n, float *out_distance,
                             cudaStream_t stream = nullptr, unsigned int *out_edge = nullptr, float3 *out_point = nullptr) const
    {
        detail::check_status(snch_closest_silhouette_batch(built(), &points->x, flip, r_max, n, out_distance, out_edge, out_point ? &out_point->x : nullptr, stream));
    }
    // query_device(bvh_dev, ray_intersect<any_hit>(ray(o, d), t_max), intersect_test())
    void intersect(const float3 *origins, const float3 *directions, const float *t_max, std::size_t n, snch_hit *out_hits, unsigned char *out_found,
                   bool any_hit = false, cudaStream_t stream = nullptr) const
    {
        detail::check_status(snch_intersect_batch(built(), &origins->x, &directions->x, t_max, n, out_hits, out_found, any_hit ? 1 : 0, stream));
    }
    // sample_object_in_sphere(...) followed by sample_on_object(...); rnd = (u, u1, u2) per query
    void sample_in_spheres(const float4 *spheres, const float3 *rnd, std::size_t n, int *out_index, float *out_pdf, float3 *out_point,
                           cudaStream_t stream = nullptr) const
    {
        detail::check_status(snch_sample_in_sphere_batch(built(), &spheres->x, &rnd->x, n, out_index, out_pdf, out_point ? &out_point->x : nullptr, stream));
    }
    snch_scene *native_handle() const noexcept { return handle_.get(); }

public:
    thrust::host_vector<float3> vertices_h;
    detail::device_view<const float3> vertices_d;
    thrust::host_vector<int3> indices_h;
    thrust::host_vector<int3> edge_indices_h;
    std::vector<triangle> triangles;
    thrust::host_vector<silhouette_edge> silhouettes_h;
    detail::device_view<const silhouette_edge> silhouettes_d;
    std::unique_ptr<built_tree> p_bvh;
    lbvh::bvh_device<float, 3, triangle> bvh_dev;

private:
    struct handle_deleter
    {
        void operator()(snch_scene *s) const noexcept { snch_scene_destroy(s); }
    };
    void require_handle() const
    {
        if (!handle_) throw std::runtime_error("snch_lbvh: scene<3> was default-constructed (no geometry)");
    }
    const snch_scene *built() const
    {
        if (!p_bvh) throw std::runtime_error("BVH is not built yet.");
        return handle_.get();
    }
    std::unique_ptr<snch_scene, handle_deleter> handle_;
};
} // namespace lbvh
#endif // SNCH_LBVH_B200_SCENE_CUH
