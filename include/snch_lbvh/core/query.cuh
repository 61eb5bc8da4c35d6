// snch_lbvh/core/query.cuh — per-thread traversals callable from user kernels (drop-in C++ API).
//
// Same five query_device() overloads as the reference's core/query.cuh (line intersection, ray intersection with
// closest-hit / any-hit, box overlap, nearest primitive, nearest silhouette), same functor call conventions, same return
// types, operating on the reference-layout arrays of lbvh::bvh_device.  They exist so that code which calls the reference
// from its own kernels keeps compiling and returns the same answers; throughput-critical callers should use the batched
// entry points of include/snch_b200.h, which run the warp-scheduled kernels of libsnch_b200.so.
//
// Differences in HOW the tree is walked (never in what is returned, up to the documented tie rules):
//   * nearest / nearest_silhouette descend into the nearer child first and discard a child as soon as its box is farther
//     than the best answer so far — at push time and again at pop time.  The reference visits left-then-right with a
//     MINMAXDIST filter and opens 4-6x more nodes.  Equal-distance primitives: the first one reached is
//     reported (the reference reports the last one it reaches; either is a member of the argmin set).
//   * a tree with a single primitive (root is a leaf) is handled everywhere (the reference reads out of bounds, Q6).
// ray / line / overlap walks keep the reference's visiting order, so hit records on ties and truncated output buffers
// come out the same.
#ifndef SNCH_LBVH_B200_QUERY_CUH
#define SNCH_LBVH_B200_QUERY_CUH
#include "bvh.cuh"
#include "predicator.cuh"

namespace lbvh
{
namespace detail
{
constexpr int kTraversalStack = 64; // depth bound of a tree over 62-bit augmented keys
constexpr std::uint32_t kNoObject = 0xFFFFFFFFu;

template <typename Key> struct keyed_stack
{
    std::uint32_t node[kTraversalStack];
    Key key[kTraversalStack];
    int size = 0;
    SNCH_LBVH_DEVICE_INLINE void push(std::uint32_t n, Key k)
    {
        node[size] = n;
        key[size] = k;
        ++size;
    }
    SNCH_LBVH_DEVICE_INLINE bool empty() const { return size == 0; }
    SNCH_LBVH_DEVICE_INLINE void pop(std::uint32_t *n, Key *k)
    {
        --size;
        *n = node[size];
        *k = key[size];
    }
};

// Depth-first walk that tests both children of every opened node with `accept_box` and hands accepted leaves to
// `on_leaf(object index)`; children are examined left then right and opened last-in-first-out.
template <typename Bvh, typename BoxTest, typename LeafFn> SNCH_LBVH_DEVICE_INLINE void walk_all(const Bvh &bvh, BoxTest accept_box, LeafFn on_leaf)
{
    if (bvh.num_objects == 0) return;
    if (bvh.nodes[0].object_idx != kNoObject)
    { // single-primitive tree
        if (accept_box(bvh.aabbs[0])) on_leaf(bvh.nodes[0].object_idx);
        return;
    }
    std::uint32_t open[kTraversalStack];
    int top = 0;
    open[top++] = 0;
    while (top > 0)
    {
        const auto &nd = bvh.nodes[open[--top]];
        const std::uint32_t child[2] = {nd.left_idx, nd.right_idx};
        for (int c = 0; c < 2; ++c)
        {
            if (!accept_box(bvh.aabbs[child[c]])) continue;
            const std::uint32_t obj = bvh.nodes[child[c]].object_idx;
            if (obj != kNoObject) on_leaf(obj);
            else open[top++] = child[c];
        }
    }
}
} // namespace detail

// Objects intersected by the infinite line q.l.  element_intersects(line, object) -> pair<bool, data>; every hit is
// written as pair<object index, data> through outiter while fewer than max_buffer_size were found.  Returns the number
// of hits (which may exceed max_buffer_size).                                                            query.cuh:12-77
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename OutputIterator, typename IntersectionTestFunc>
SNCH_LBVH_DEVICE unsigned int query_device(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh,
                                           const query_line_intersect<Real, dim> q, IntersectionTestFunc element_intersects,
                                           OutputIterator outiter, const unsigned int max_buffer_size) noexcept
{
    unsigned int num_found = 0;
    detail::walk_all(
        bvh, [&](const aabb<Real, dim> &box) { return intersects(q.l, box); },
        [&](std::uint32_t obj)
        {
            auto flag_data = element_intersects(q.l, bvh.objects[obj]);
            if (!flag_data.first) return;
            if (num_found < max_buffer_size) *outiter++ = thrust::pair<unsigned int, decltype(flag_data.second)>(obj, flag_data.second);
            ++num_found;
        });
    return num_found;
}

// Objects whose leaf box overlaps q.target: their indices go through outiter (at most max_buffer_size of them); returns
// how many overlap.                                                                                      query.cuh:171-236
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename OutputIterator>
SNCH_LBVH_DEVICE unsigned int query_device(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh, const query_overlap<Real, dim> q,
                                           OutputIterator outiter, const unsigned int max_buffer_size = 0xFFFFFFFF) noexcept
{
    unsigned int num_found = 0;
    detail::walk_all(
        bvh, [&](const aabb<Real, dim> &box) { return intersects(q.target, box); },
        [&](std::uint32_t obj)
        {
            if (num_found < max_buffer_size) *outiter++ = obj;
            ++num_found;
        });
    return num_found;
}

// Nearest hit of ray q.r with parameter t < q.max_dist.  element_intersects(ray, object) -> tuple<bool hit, Real t, uv>.
// Returns tuple<found, t (+inf if none), uv, object index (0xFFFFFFFF if none)>, or just `found` when TestOnly (any hit).
// Front-to-back: of two intersected children the one entered first is opened first; subtrees entered beyond the best
// hit are skipped; among equal t the first triangle reached wins (strict <).                             query.cuh:79-169
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename IntersectionTestFunc, bool TestOnly>
SNCH_LBVH_DEVICE auto query_device(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh,
                                   const query_ray_intersect<Real, dim, TestOnly> q, IntersectionTestFunc element_intersects) noexcept
{
    using uv_type = std::conditional_t<dim == 3, float2, float>;
    Real best_t = infinity<Real>();
    bool found = false;
    uv_type uv{};
    unsigned int best_obj = detail::kNoObject;
    detail::keyed_stack<Real> todo;
    if (bvh.num_objects > 0) todo.push(0, infinity<Real>());
    while (!todo.empty())
    {
        std::uint32_t n;
        Real entry;
        todo.pop(&n, &entry);
        if (entry > best_t) continue;
        const auto &nd = bvh.nodes[n];
        if (nd.object_idx != detail::kNoObject)
        {
            auto hit = element_intersects(q.r, bvh.objects[nd.object_idx]);
            if (thrust::get<0>(hit) && thrust::get<1>(hit) < q.max_dist && thrust::get<1>(hit) < best_t)
            {
                best_t = thrust::get<1>(hit);
                found = true;
                if constexpr (TestOnly) return true;
                uv = thrust::get<2>(hit);
                best_obj = nd.object_idx;
            }
            continue;
        }
        Real tl, tr;
        const bool hl = intersects_d(q.r, bvh.aabbs[nd.left_idx], Real(q.max_dist), &tl);
        const bool hr = intersects_d(q.r, bvh.aabbs[nd.right_idx], Real(q.max_dist), &tr);
        if (hl && hr)
        {
            const bool right_first = tr < tl;
            todo.push(right_first ? nd.left_idx : nd.right_idx, right_first ? tl : tr); // farther child: opened later
            todo.push(right_first ? nd.right_idx : nd.left_idx, right_first ? tr : tl);
        }
        else if (hl) todo.push(nd.left_idx, tl);
        else if (hr) todo.push(nd.right_idx, tr);
    }
    if constexpr (TestOnly) return false;
    else return thrust::make_tuple(found, best_t, uv, best_obj);
}

// Nearest object to q.target.  calc_dist(point, object) -> distance.  Returns pair<object index, distance>
// (0xFFFFFFFF, +inf for an empty tree).  Distances are compared squared, like the reference (query.cuh:284-291), so the
// returned value is sqrt(d*d) of the functor's result.                                                   query.cuh:238-318
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename DistanceCalculator>
SNCH_LBVH_DEVICE thrust::pair<unsigned int, Real> query_device(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh,
                                                               const query_nearest<Real, dim> &q, DistanceCalculator calc_dist) noexcept
{
    Real best2 = infinity<Real>();
    unsigned int best_obj = detail::kNoObject;
    auto test_leaf = [&](std::uint32_t obj)
    {
        Real d = calc_dist(q.target, bvh.objects[obj]);
        d *= d;
        if (d < best2 || (best_obj == detail::kNoObject && d <= best2)) // the first object at any distance up to +inf, never a NaN
        {
            best2 = d;
            best_obj = obj;
        }
    };
    if (bvh.num_objects == 0) return thrust::make_pair(best_obj, best2);
    if (bvh.nodes[0].object_idx != detail::kNoObject)
    {
        test_leaf(bvh.nodes[0].object_idx);
        return thrust::make_pair(best_obj, detail::sqrt_of(best2));
    }
    detail::keyed_stack<Real> todo;
    std::uint32_t n = 0;
    for (;;)
    {
        const auto &nd = bvh.nodes[n];
        std::uint32_t child[2] = {nd.left_idx, nd.right_idx};
        Real m[2] = {mindist(bvh.aabbs[child[0]], q.target), mindist(bvh.aabbs[child[1]], q.target)};
        if (m[1] < m[0])
        {
            lbvh_swap(m[0], m[1]);
            lbvh_swap(child[0], child[1]);
        }
        std::uint32_t next = detail::kNoObject;
        for (int c = 0; c < 2; ++c)
        {
            if (!(m[c] < best2 || best_obj == detail::kNoObject)) continue;
            const std::uint32_t obj = bvh.nodes[child[c]].object_idx;
            if (obj != detail::kNoObject) test_leaf(obj);
            else if (next == detail::kNoObject) next = child[c];
            else todo.push(child[c], m[c]);
        }
        while (next == detail::kNoObject && !todo.empty())
        {
            Real key;
            todo.pop(&next, &key);
            if (!(key < best2)) next = detail::kNoObject;
        }
        if (next == detail::kNoObject) break;
        n = next;
    }
    return thrust::make_pair(best_obj, detail::sqrt_of(best2));
}

// Distance to the nearest silhouette element as seen from q.target (+inf if none is reachable).  A subtree is entered
// only if its normal cone is valid and overlap() says it may hold a silhouette element for this view point — the
// reference's own pruning predicate, so the set of reachable leaves is the reference's.
// calc_dist(point, object, max_radius_squared, Real &dist, flip, min_radius_squared) -> found.           query.cuh:320-423
template <typename Real, unsigned int dim, typename Objects, bool IsConst, typename DistanceCalculator>
SNCH_LBVH_DEVICE Real query_device(const detail::basic_device_bvh<Real, dim, Objects, IsConst> &bvh,
                                   const query_nearest_silhouette<Real, dim> &q, DistanceCalculator calc_dist) noexcept
{
    Real best = infinity<Real>();
    auto test_leaf = [&](std::uint32_t obj)
    {
        Real d = infinity<Real>();
        const bool ok = calc_dist(q.target, bvh.objects[obj], best * best, d, q.flip_normal_orientation, 0.0f);
        if (ok && d <= best) best = d;
    };
    auto may_hold_silhouette = [&](std::uint32_t n, Real m2)
    {
        Real lo, hi;
        return is_valid(bvh.cones[n]) && overlap(bvh.cones[n], q.target, bvh.aabbs[n], m2, &lo, &hi);
    };
    if (bvh.num_objects == 0) return best;
    if (bvh.nodes[0].object_idx != detail::kNoObject)
    {
        if (may_hold_silhouette(0, mindist(bvh.aabbs[0], q.target))) test_leaf(bvh.nodes[0].object_idx);
        return best;
    }
    detail::keyed_stack<Real> todo;
    std::uint32_t n = 0;
    for (;;)
    {
        const auto &nd = bvh.nodes[n];
        std::uint32_t child[2] = {nd.left_idx, nd.right_idx};
        Real m[2] = {mindist(bvh.aabbs[child[0]], q.target), mindist(bvh.aabbs[child[1]], q.target)};
        if (m[1] < m[0])
        {
            lbvh_swap(m[0], m[1]);
            lbvh_swap(child[0], child[1]);
        }
        std::uint32_t next = detail::kNoObject;
        for (int c = 0; c < 2; ++c)
        {
            if (!(m[c] <= best * best) || !may_hold_silhouette(child[c], m[c])) continue;
            const std::uint32_t obj = bvh.nodes[child[c]].object_idx;
            if (obj != detail::kNoObject) test_leaf(obj);
            else if (next == detail::kNoObject) next = child[c];
            else todo.push(child[c], m[c]);
        }
        while (next == detail::kNoObject && !todo.empty())
        {
            Real key;
            todo.pop(&next, &key);
            if (!(key <= best * best)) next = detail::kNoObject;
        }
        if (next == detail::kNoObject) break;
        n = next;
    }
    return best;
}
} // namespace lbvh
#endif // SNCH_LBVH_B200_QUERY_CUH
