// query.cu — batched traversal kernels (one launch per query batch).
//
// Each kernel returns what the reference's per-thread query_device()/sample_object_in_sphere() returns for the same
// query (query.cuh:79-169, 238-318, 325-423; sample.cuh:23-92), but walks the two-child traversal records
// (layout.h) near-child-first with the running best distance applied when a child is *pushed*, not only when it is
// popped — the reference pops 4-6x more nodes than any exact traversal must open.
//
// Result equivalence (see DESIGN.md "Parity rules"):
//   closest   : min over all triangles of the reference's own point-triangle distance; index = any argmin (ties, Q3)
//   silhouette: min over the leaves that pass the reference's cone test chain (same predicate, same libm calls)
//   ray       : smallest t with t < max_dist; prim = any triangle attaining it (Q4)
//   sample    : identical single-path descent (deterministic given u)
#include "scene.h"
#include "snch_math.cuh"

namespace snch
{

constexpr int kQueryThreads = 128;
constexpr int kStackDepth = 64; // >= 62 levels possible with the 62-bit augmented key

struct NodeBoxes
{
    V3 lo0, hi0, lo1, hi1;
};
SNCH_DI NodeBoxes unpack_boxes(float4 a, float4 b, float4 c)
{
    NodeBoxes n;
    n.lo0 = V3{a.x, a.y, a.z};
    n.hi0 = V3{a.w, b.x, b.y};
    n.lo1 = V3{b.z, b.w, c.x};
    n.hi1 = V3{c.y, c.z, c.w};
    return n;
}
SNCH_DI V3 load_point(const float *__restrict__ q, uint64_t i) { return V3{q[3 * i], q[3 * i + 1], q[3 * i + 2]}; }

// ---------------------------------------------------------------------------------------------------------------
// nearest primitive                                                                      query.cuh:238-318
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQueryThreads)
    k_closest(SceneView sv, const float *__restrict__ q, uint64_t n, uint32_t *__restrict__ out_idx, float *__restrict__ out_dist)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = load_point(q, i);
    uint32_t stk_n[kStackDepth];
    float stk_k[kStackDepth];
    int sp = 0;
    float best2 = INFINITY;
    uint32_t best = kNone;
    uint32_t node = 0;
    for (;;)
    {
        const float4 *np = reinterpret_cast<const float4 *>(sv.bnode + node);
        const float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2), d = __ldg(np + 3);
        const NodeBoxes nb = unpack_boxes(a, b, c);
        float m0 = box_mindist2(nb.lo0, nb.hi0, p), m1 = box_mindist2(nb.lo1, nb.hi1, p);
        uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y);
        if (m1 < m0)
        {
            const float tm = m0;
            m0 = m1;
            m1 = tm;
            const uint32_t tr = r0;
            r0 = r1;
            r1 = tr;
        }
        uint32_t next = kNone;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
        {
            const float m = ch ? m1 : m0;
            const uint32_t r = ch ? r1 : r0;
            if (!(m < best2)) continue;
            if (r & kLeafFlag)
            {
                const float4 *tp = reinterpret_cast<const float4 *>(sv.ltri + (r & ~kLeafFlag));
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float dist = point_triangle_distance(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, p);
                dist *= dist; // the reference squares the distance it got back (query.cuh:284-285)
                if (dist < best2)
                {
                    best2 = dist;
                    best = __float_as_uint(t0.w);
                }
            }
            else if (next == kNone) next = r;
            else
            {
                stk_n[sp] = r;
                stk_k[sp] = m;
                ++sp;
            }
        }
        if (next == kNone)
        {
            while (sp > 0)
            {
                --sp;
                if (stk_k[sp] < best2)
                {
                    next = stk_n[sp];
                    break;
                }
            }
            if (next == kNone) break;
        }
        node = next;
    }
    out_idx[i] = best;
    out_dist[i] = sqrtf(best2);
}

// ---------------------------------------------------------------------------------------------------------------
// nearest silhouette                                                                     query.cuh:325-423
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQueryThreads)
    k_silhouette(SceneView sv, const float *__restrict__ q, const uint8_t *__restrict__ flipv, const float *__restrict__ rmax, uint64_t n,
                 float *__restrict__ out_dist)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p = load_point(q, i);
    const bool flip = flipv ? (flipv[i] != 0) : false;
    float best = rmax ? rmax[i] : INFINITY;
    float best2 = best * best;
    bool found = false;
    uint32_t stk_n[kStackDepth];
    float stk_k[kStackDepth];
    int sp = 0;
    uint32_t node = 0;
    for (;;)
    {
        const float4 *np = reinterpret_cast<const float4 *>(sv.snode + node);
        const float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2), d = __ldg(np + 3), e = __ldg(np + 4), f = __ldg(np + 5);
        const NodeBoxes nb = unpack_boxes(a, b, c);
        const float m0 = box_mindist2(nb.lo0, nb.hi0, p), m1 = box_mindist2(nb.lo1, nb.hi1, p);
        const uint32_t r0 = __float_as_uint(f.z), r1 = __float_as_uint(f.w);
        // the reference's per-child test: is_valid(cone) && overlap(cone, p, box, mindist^2)   (query.cuh:366-367),
        // evaluated only for children that can still beat the current best
        bool h0 = (m0 <= best2) && (d.w >= 0.0f) && cone_overlap(V3{d.x, d.y, d.z}, d.w, e.x, p, nb.lo0, nb.hi0, m0);
        bool h1 = (m1 <= best2) && (f.x >= 0.0f) && cone_overlap(V3{e.y, e.z, e.w}, f.x, f.y, p, nb.lo1, nb.hi1, m1);
        const bool swap = m1 < m0;
        uint32_t next = kNone;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
        {
            const bool second = (ch == 1) != swap; // visit the nearer child first
            const bool h = second ? h1 : h0;
            const float m = second ? m1 : m0;
            const uint32_t r = second ? r1 : r0;
            if (!h || !(m <= best2)) continue;
            if (r & kLeafFlag)
            {
                const uint32_t payload = r & ~kLeafFlag;
                const uint32_t first = payload >> 2, cnt = payload & 3u;
                for (uint32_t k = 0; k < cnt; ++k)
                { // silhouette_distance_calculator over the owned edges          scene.cuh:978-1003, 788-824
                    const float4 *ep = reinterpret_cast<const float4 *>(sv.ledge + first + k);
                    const float4 e0 = __ldg(ep), e1 = __ldg(ep + 1), e2 = __ldg(ep + 2);
                    const V3 pa = V3{e0.x, e0.y, e0.z}, pb = V3{e0.w, e1.x, e1.y};
                    V3 cp;
                    const float dist = point_segment_distance(pa, pb, p, &cp);
                    if (dist * dist > best2) continue;
                    bool is_sil = isnan(e1.z); // boundary edge
                    if (!is_sil) is_sil = is_silhouette_edge(pa, pb, V3{e1.z, e1.w, e2.x}, V3{e2.y, e2.z, e2.w}, p - cp, dist, flip);
                    if (is_sil && dist <= best)
                    {
                        best = dist;
                        best2 = dist * dist;
                        found = true;
                    }
                }
            }
            else if (next == kNone) next = r;
            else
            {
                stk_n[sp] = r;
                stk_k[sp] = m;
                ++sp;
            }
        }
        if (next == kNone)
        {
            while (sp > 0)
            {
                --sp;
                if (stk_k[sp] <= best2)
                {
                    next = stk_n[sp];
                    break;
                }
            }
            if (next == kNone) break;
        }
        node = next;
    }
    out_dist[i] = found ? best : INFINITY;
}

// ---------------------------------------------------------------------------------------------------------------
// ray intersection (closest hit / any hit)                                               query.cuh:79-169
// ---------------------------------------------------------------------------------------------------------------
template <bool kAnyHit>
__global__ void __launch_bounds__(kQueryThreads)
    k_intersect(SceneView sv, const float *__restrict__ org, const float *__restrict__ dir, const float *__restrict__ tmaxv, uint64_t n,
                snch_hit *__restrict__ hits, uint8_t *__restrict__ found_out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 o = load_point(org, i), dv = load_point(dir, i);
    const V3 dinv = V3{1.0f / dv.x, 1.0f / dv.y, 1.0f / dv.z}; // aabb.cuh:305-312
    const float max_dist = tmaxv ? tmaxv[i] : INFINITY;
    float best_t = INFINITY, best_u = 0.f, best_v = 0.f;
    uint32_t best_prim = kNone;
    bool found = false;
    uint32_t stk_n[kStackDepth];
    float stk_k[kStackDepth];
    int sp = 0;
    uint32_t node = 0;
    for (;;)
    {
        const float4 *np = reinterpret_cast<const float4 *>(sv.bnode + node);
        const float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2), d = __ldg(np + 3);
        const NodeBoxes nb = unpack_boxes(a, b, c);
        float e0, e1;
        bool h0 = box_ray(nb.lo0, nb.hi0, o, dinv, max_dist, &e0);
        bool h1 = box_ray(nb.lo1, nb.hi1, o, dinv, max_dist, &e1);
        const uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y);
        const bool swap = h0 && h1 && (e1 < e0); // the reference visits L first on ties (query.cuh:141)
        uint32_t next = kNone;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
        {
            const bool second = (ch == 1) != swap;
            const bool h = second ? h1 : h0;
            const float en = second ? e1 : e0;
            const uint32_t r = second ? r1 : r0;
            if (!h || en > best_t) continue; // same rejection the reference applies at pop time (query.cuh:106)
            if (r & kLeafFlag)
            {
                const float4 *tp = reinterpret_cast<const float4 *>(sv.ltri + (r & ~kLeafFlag));
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                float t, u, v;
                if (ray_triangle(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, o, dv, &t, &u, &v) && t < max_dist &&
                    t < best_t)
                {
                    best_t = t;
                    best_u = u;
                    best_v = v;
                    best_prim = __float_as_uint(t0.w);
                    found = true;
                    if (kAnyHit)
                    {
                        found_out[i] = 1;
                        return;
                    }
                }
            }
            else if (next == kNone) next = r;
            else
            {
                stk_n[sp] = r;
                stk_k[sp] = en;
                ++sp;
            }
        }
        if (next == kNone)
        {
            while (sp > 0)
            {
                --sp;
                if (!(stk_k[sp] > best_t))
                {
                    next = stk_n[sp];
                    break;
                }
            }
            if (next == kNone) break;
        }
        node = next;
    }
    if (found_out) found_out[i] = found ? 1 : 0;
    if (!kAnyHit && hits)
    {
        snch_hit h;
        h.t = best_t;
        h.u = best_u;
        h.v = best_v;
        h.prim = best_prim;
        hits[i] = h;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// SampleTriangleInSphere                                              sample.cuh:23-92 + 7-21, scene.cuh:14-27
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQueryThreads)
    k_sample(SceneView sv, const float *__restrict__ sph, const float *__restrict__ rnd, uint64_t n, int32_t *__restrict__ out_idx,
             float *__restrict__ out_pdf, float *__restrict__ out_pt)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 ctr = V3{sph[4 * i], sph[4 * i + 1], sph[4 * i + 2]};
    const float radius = sph[4 * i + 3];
    float u = rnd[3 * i];
    float path = 1.0f;
    int32_t idx = -1;
    float pdf = 0.0f;
    V3 pt = V3{0.f, 0.f, 0.f};
    uint32_t node = 0;
    for (;;)
    {
        const float4 *np = reinterpret_cast<const float4 *>(sv.bnode + node);
        const float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2), d = __ldg(np + 3);
        const NodeBoxes nb = unpack_boxes(a, b, c);
        const V3 c0 = V3{(nb.hi0.x + nb.lo0.x) * 0.5f, (nb.hi0.y + nb.lo0.y) * 0.5f, (nb.hi0.z + nb.lo0.z) * 0.5f};
        const V3 c1 = V3{(nb.hi1.x + nb.lo1.x) * 0.5f, (nb.hi1.y + nb.lo1.y) * 0.5f, (nb.hi1.z + nb.lo1.z) * 0.5f};
        const float w0 = box_sphere(nb.lo0, nb.hi0, ctr, radius) ? green_weight3(ctr, c0) : 0.0f;
        const float w1 = box_sphere(nb.lo1, nb.hi1, ctr, radius) ? green_weight3(ctr, c1) : 0.0f;
        const float total = w0 + w1;
        if (!(total > 0.0f)) break;
        const float p0 = w0 / total;
        uint32_t r;
        if (u < p0)
        {
            u /= p0;
            r = __float_as_uint(d.x);
            path = p0 * path;
        }
        else
        {
            const float p1 = 1.0f - p0;
            u = (u - p0) / p1;
            r = __float_as_uint(d.y);
            path = p1 * path;
        }
        if (r & kLeafFlag)
        {
            const float4 *tp = reinterpret_cast<const float4 *>(sv.ltri + (r & ~kLeafFlag));
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            const V3 pa = V3{t0.x, t0.y, t0.z}, pb = V3{t1.x, t1.y, t1.z}, pc = V3{t2.x, t2.y, t2.z};
            if (sphere_triangle(pa, pb, pc, ctr, radius))
            {
                idx = (int32_t)__float_as_uint(t0.w);
                pdf = path / triangle_area(pa, pb, pc);
                float su = rnd[3 * i + 1], sv2 = rnd[3 * i + 2];
                if (su + sv2 > 1.0f)
                {
                    su = 1.0f - su;
                    sv2 = 1.0f - sv2;
                }
                const float w = 1.0f - su - sv2;
                pt = V3{w * pa.x + su * pb.x + sv2 * pc.x, w * pa.y + su * pb.y + sv2 * pc.y, w * pa.z + su * pb.z + sv2 * pc.z};
            }
            break;
        }
        node = r;
    }
    out_idx[i] = idx;
    out_pdf[i] = pdf;
    if (out_pt)
    {
        out_pt[3 * i] = pt.x;
        out_pt[3 * i + 1] = pt.y;
        out_pt[3 * i + 2] = pt.z;
    }
}

// empty scene: the reference's construct() returns early and every pointer is null (bvh.cuh:383-386); batched calls
// on an empty scene return the sentinels
__global__ void k_fill_empty(uint64_t n, uint32_t *idx, float *dist, snch_hit *hits, uint8_t *found, int32_t *sidx, float *pdf, float *pt)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (idx) idx[i] = kNone;
    if (dist) dist[i] = INFINITY;
    if (hits)
    {
        snch_hit h;
        h.t = INFINITY;
        h.u = h.v = 0.f;
        h.prim = kNone;
        hits[i] = h;
    }
    if (found) found[i] = 0;
    if (sidx) sidx[i] = -1;
    if (pdf) pdf[i] = 0.f;
    if (pt) pt[3 * i] = pt[3 * i + 1] = pt[3 * i + 2] = 0.f;
}

static inline unsigned grid_for(uint64_t n) { return (unsigned)((n + kQueryThreads - 1) / kQueryThreads); }

int launch_closest(const SceneView &v, const float *q, uint64_t n, uint32_t *idx, float *dist, cudaStream_t st)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0) k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, idx, dist, nullptr, nullptr, nullptr, nullptr, nullptr);
    else k_closest<<<grid_for(n), kQueryThreads, 0, st>>>(v, q, n, idx, dist);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}
int launch_silhouette(const SceneView &v, const float *q, const uint8_t *flip, const float *rmax, uint64_t n, float *dist, cudaStream_t st)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0) k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, nullptr, dist, nullptr, nullptr, nullptr, nullptr, nullptr);
    else k_silhouette<<<grid_for(n), kQueryThreads, 0, st>>>(v, q, flip, rmax, n, dist);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}
int launch_intersect(const SceneView &v, const float *o, const float *d, const float *tmax, uint64_t n, snch_hit *hits, uint8_t *found,
                     int any_hit, cudaStream_t st)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0) k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, nullptr, nullptr, hits, found, nullptr, nullptr, nullptr);
    else if (any_hit) k_intersect<true><<<grid_for(n), kQueryThreads, 0, st>>>(v, o, d, tmax, n, hits, found);
    else k_intersect<false><<<grid_for(n), kQueryThreads, 0, st>>>(v, o, d, tmax, n, hits, found);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}
int launch_sample(const SceneView &v, const float *sph, const float *rnd, uint64_t n, int32_t *idx, float *pdf, float *pt, cudaStream_t st)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0) k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, nullptr, nullptr, nullptr, nullptr, idx, pdf, pt);
    else k_sample<<<grid_for(n), kQueryThreads, 0, st>>>(v, sph, rnd, n, idx, pdf, pt);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}

} // namespace snch
