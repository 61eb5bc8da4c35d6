// scene2.cu — 2-D scenes (polylines: line segments / silhouette vertices) behind the C-ABI: snch_scene2_* and *_batch2.
//
// Replaces lbvh::scene<2> of the reference (scene.cuh:287-703) and the per-thread query_device() / sample_object_in_sphere()
// calls a user kernel makes on its bvh_device (query.cuh:79-169, 238-318, 325-423; sample.cuh:23-92) with batched launches,
// the 2-D twin of the scene<3> path:
//   * compute_silhouettes / ownership: the reference's host loops (scene.cuh:634-681), O(N)
//   * leaf boxes and leaf normal cones: scene<2>::aabb_getter / cone_getter evaluated per segment on the GPU
//   * tree: snch_lbvh_build(dim = 2) — Morton, radix sort of (key, index), Karras hierarchy, one box + cone refit
//   * traversal records: N2 = both child boxes | both child refs | both child cones (96 B, three 32 B sectors: the
//     closest-point, ray and sampling kernels read two of them), L2 = segment end points + object index | the (at most
//     two) silhouette vertices the segment owns with their precomputed unit normals (96 B)
//   * kernels: persistent lanes with warp-level refill over Morton-ordered batches (query_common.cuh), nearer child first,
//     bounds applied at push and at pop.
// The per-primitive arithmetic and the box / cone predicates are the drop-in headers' own functions (include/snch_lbvh/core:
// mindist, overlap, intersects_d, intersect_sphere, merge; scene.cuh: the scene<2> functors), i.e. the code the per-thread
// path of the same library runs, so both paths return the same values; parity with the UNMODIFIED reference headers (run on
// the CPU and on the GPU) and with committed golden vectors is pinned by tests/test_gpu_scene2.py.
#include "query_common.cuh"
#include "sort_scan.cuh"

#include "../../include/snch_lbvh/lbvh.cuh"
#include "../../include/snch_lbvh/scene.cuh"

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace snch
{
namespace
{
using Scene2T = lbvh::scene<2>;
using SegT = Scene2T::line_segment;
using SilT = Scene2T::silhouette_vertex;
using Box2 = lbvh::aabb<float, 2>;
using Cone2 = lbvh::cone<float, 2>;
static_assert(sizeof(SegT) == 32 && sizeof(SilT) == 32 && sizeof(Box2) == 16 && sizeof(Cone2) == 16, "reference layouts (SURVEY 8)");

constexpr uint32_t kLeaf2 = 0x80000000u;     // child ref: leaf flag | owned-vertex count << 29 | leaf position (Morton order)
constexpr uint32_t kRefIndex2 = 0x1FFFFFFFu; // internal node id or leaf position
constexpr uint32_t kMaxSegments2 = kRefIndex2;

struct __align__(32) N2
{
    float box[8];  // child 0 {lower.x, lower.y, upper.x, upper.y}, child 1
    uint32_t ref[2];
    uint32_t pad[6];
    float cone[8]; // child 0 {axis.x, axis.y, half_angle, radius}, child 1
};
struct __align__(32) L2
{
    float p0x, p0y, p1x, p1y;
    uint32_t object, owned, pad0, pad1;
    float vert[2][8]; // owned silhouette vertex: {x, y, n0.x, n0.y, n1.x, n1.y, face bits (bit 0: has_face(0), bit 1: has_face(1)), vertex id bits}
};
static_assert(sizeof(N2) == 96 && sizeof(L2) == 96, "record sizes");

struct View2
{
    uint32_t n; // segments
    const N2 *n2;
    const L2 *l2;
    const Box2 *aabbs;   // reference layout (root box / cone of a one-segment scene)
    const Cone2 *cones;
};

SNCH_DI Box2 make_box(float lx, float ly, float ux, float uy) { return Box2(make_float2(ux, uy), make_float2(lx, ly)); }
SNCH_DI float dist_point_segment(float2 p0, float2 p1, float2 x)
{
    float2 pt;
    float t;
    return lbvh::find_closest_point_line_segment(p0, p1, x, &pt, &t);
}

// ---------------------------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k2_leaf_inputs(const SegT *__restrict__ lines, uint32_t n, Box2 *boxes, Cone2 *cones)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const SegT s = lines[i];
    boxes[i] = Scene2T::aabb_getter()(s);  // scene.cuh:419-432
    cones[i] = Scene2T::cone_getter()(s);  // scene.cuh:434-500
}
// leaf record k (Morton order) from object sorted_idx[k]
__global__ void __launch_bounds__(128) k2_leaf_records(const SegT *__restrict__ lines, const uint32_t *__restrict__ sorted_idx, uint32_t n, L2 *l2)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t obj = sorted_idx[k];
    const SegT s = lines[obj];
    L2 r;
    const float2 p0 = s.vertices[s.vertex_indices.x], p1 = s.vertices[s.vertex_indices.y];
    r.p0x = p0.x, r.p0y = p0.y, r.p1x = p1.x, r.p1y = p1.y;
    r.object = obj;
    r.pad0 = r.pad1 = 0;
    uint32_t cnt = 0;
    for (int j = 0; j < 2; ++j)
    {
        for (int c = 0; c < 8; ++c) r.vert[j][c] = 0.0f;
    }
    const int owned[2] = {s.silhouette_indices.x, s.silhouette_indices.y};
    for (int j = 0; j < 2; ++j)
    {
        if (owned[j] == -1) continue;
        const SilT sv = s.silhouettes[owned[j]];
        const float2 p = sv.centroid();
        const bool f0 = sv.has_face(0), f1 = sv.has_face(1);
        const float2 n0 = f0 ? sv.normal(0) : make_float2(0.f, 0.f), n1 = f1 ? sv.normal(1) : make_float2(0.f, 0.f);
        float *o = r.vert[cnt];
        o[0] = p.x, o[1] = p.y, o[2] = n0.x, o[3] = n0.y, o[4] = n1.x, o[5] = n1.y;
        o[6] = __uint_as_float((f0 ? 1u : 0u) | (f1 ? 2u : 0u));
        o[7] = __int_as_float(owned[j]); // the silhouette vertex's id: what out_vertex of a silhouette query reports
        ++cnt;
    }
    r.owned = cnt;
    l2[k] = r;
}
__global__ void __launch_bounds__(128)
    k2_node_records(const RefNode *__restrict__ nodes, const Box2 *__restrict__ aabbs, const Cone2 *__restrict__ cones, const L2 *__restrict__ l2,
                    uint32_t n, N2 *n2)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= n) return;
    const RefNode nd = nodes[i];
    N2 r;
    const uint32_t child[2] = {nd.left, nd.right};
    for (int c = 0; c < 2; ++c)
    {
        const Box2 b = aabbs[child[c]];
        const Cone2 cn = cones[child[c]];
        r.box[4 * c + 0] = b.lower.x, r.box[4 * c + 1] = b.lower.y, r.box[4 * c + 2] = b.upper.x, r.box[4 * c + 3] = b.upper.y;
        r.cone[4 * c + 0] = cn.axis.x, r.cone[4 * c + 1] = cn.axis.y, r.cone[4 * c + 2] = cn.half_angle, r.cone[4 * c + 3] = cn.radius;
        if (child[c] >= n - 1)
        {
            const uint32_t k = child[c] - (n - 1);
            r.ref[c] = kLeaf2 | (l2[k].owned << 29) | k;
        }
        else r.ref[c] = child[c];
    }
    for (int c = 0; c < 6; ++c) r.pad[c] = 0;
    n2[i] = r;
}

// ---------------------------------------------------------------------------------------------------------------
// traversal helpers
// ---------------------------------------------------------------------------------------------------------------
struct Pair2
{
    Box2 b0, b1;
    uint32_t r0, r1;
};
SNCH_DI Pair2 load_pair(const N2 *n2, uint32_t node)
{
    float4 a, b, c, d;
    ld256(n2 + node, a, b);
    ld256(reinterpret_cast<const char *>(n2 + node) + 32, c, d);
    Pair2 p{make_box(a.x, a.y, a.z, a.w), make_box(b.x, b.y, b.z, b.w), __float_as_uint(c.x), __float_as_uint(c.y)};
    return p;
}
SNCH_DI float2 load_point2(const float *__restrict__ q, uint64_t i) { return make_float2(__ldg(q + 2 * i), __ldg(q + 2 * i + 1)); }

// nearest segment                                                                         query.cuh:238-318
__global__ void __launch_bounds__(kQueryThreads)
    k2_closest(View2 v, const float *__restrict__ q, const uint32_t *__restrict__ perm, uint32_t n, uint32_t *__restrict__ out_idx,
               float *__restrict__ out_dist, unsigned long long *counter)
{
    const int lane = threadIdx.x & 31;
    Feeder fd{0u, 0u, false};
    StackEntry stk[kStackDepth];
    int sp = 0;
    float2 p = make_float2(0.f, 0.f);
    bool busy = false;
    float best2 = INFINITY;
    uint32_t best_obj = kNone, slot = kNone, node = kNone;
    auto test_leaf = [&](uint32_t k)
    {
        float4 a, b;
        ld256(v.l2 + k, a, b);
        float d = dist_point_segment(make_float2(a.x, a.y), make_float2(a.z, a.w), p);
        d *= d;
        if (d < best2 || (best_obj == kNone && d <= best2)) // the first object at any distance up to +inf (the reference's `<=`), never a NaN
        {
            best2 = d;
            best_obj = __float_as_uint(b.x);
        }
    };
    for (;;)
    {
        if (busy && node == kNone)
        {
            if (out_idx) out_idx[slot] = best_obj;
            out_dist[slot] = sqrtf(best2);
            busy = false;
        }
        const unsigned idle = __ballot_sync(kFull, !busy);
        if (idle)
        {
            const uint32_t s = feeder_take(fd, idle, !busy, lane, n, counter);
            if (s != kNone)
            {
                slot = perm ? __ldg(perm + s) : s;
                p = load_point2(q, slot);
                best2 = INFINITY;
                best_obj = kNone;
                busy = true;
                sp = 0;
                node = 0;
                if (v.n == 1)
                {
                    test_leaf(0);
                    node = kNone;
                }
            }
            if (fd.exhausted && __all_sync(kFull, !busy)) break;
        }
        if (node != kNone)
        {
            const Pair2 pr = load_pair(v.n2, node);
            const float m0 = lbvh::mindist(pr.b0, p), m1 = lbvh::mindist(pr.b1, p);
            const bool swap = m1 < m0;
            uint32_t next = kNone;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            {
                const bool second = (ch == 1) != swap;
                const float m = second ? m1 : m0;
                const uint32_t r = second ? pr.r1 : pr.r0;
                if (!(m < best2 || best_obj == kNone)) continue;
                if (r & kLeaf2) test_leaf(r & kRefIndex2);
                else if (next == kNone) next = r;
                else stk[sp++] = StackEntry{r, m};
            }
            while (next == kNone && sp > 0)
            {
                const StackEntry se = stk[--sp];
                if (se.key < best2) next = se.node;
            }
            node = next;
        }
    }
}

// nearest silhouette vertex                                                               query.cuh:325-423
SNCH_DI bool may_hold_silhouette2(const float *cone4, float2 o, const Box2 &box, float m2)
{
    Cone2 c;
    c.axis = make_float2(cone4[0], cone4[1]);
    c.half_angle = cone4[2];
    c.radius = cone4[3];
    float lo, hi;
    return lbvh::is_valid(c) && lbvh::overlap(c, o, box, m2, &lo, &hi);
}
__global__ void __launch_bounds__(kQueryThreads)
    k2_silhouette(View2 v, const float *__restrict__ q, const uint8_t *__restrict__ flipv, const float *__restrict__ rmax,
                  const uint32_t *__restrict__ perm, uint32_t n, float *__restrict__ out_dist, uint32_t *__restrict__ out_vertex,
                  float *__restrict__ out_point, unsigned long long *counter)
{
    const int lane = threadIdx.x & 31;
    Feeder fd{0u, 0u, false};
    StackEntry stk[kStackDepth];
    int sp = 0;
    float2 p = make_float2(0.f, 0.f);
    bool busy = false, flip = false, found = false;
    float best = INFINITY;
    uint32_t slot = kNone, node = kNone;
    uint32_t best_vid = kNone; // the optional outputs: id of the silhouette vertex that attains `best`, and the vertex itself —
    float2 best_pt = make_float2(0.f, 0.f); // `p` of silhouette_vertex::find_closest_silhouette_point (scene.cuh:354-356)
    // silhouette_distance_calculator over the owned vertices (scene.cuh:518-541, 345-378)
    auto test_leaf = [&](uint32_t k, uint32_t cnt)
    {
        float max_r2 = best * best;
        float d_found = INFINITY;
        bool ok = false;
        uint32_t vid = kNone;
        float2 vpt = make_float2(0.f, 0.f);
        for (uint32_t j = 0; j < cnt; ++j)
        {
            float4 a, b;
            ld256(reinterpret_cast<const char *>(v.l2 + k) + 32 + 32 * j, a, b);
            const float2 view_dir = make_float2(p.x - a.x, p.y - a.y);
            const float d = lbvh::length(view_dir);
            if (0.0f >= max_r2 || d * d > max_r2) continue;
            const uint32_t faces = __float_as_uint(b.z);
            bool is_sil = faces != 3u;
            if (!is_sil) is_sil = lbvh::is_silhouette_vertex(make_float2(a.z, a.w), make_float2(b.x, b.y), view_dir, d, flip);
            if (is_sil && d * d <= max_r2)
            {
                d_found = d;
                ok = true;
                max_r2 = d * d;
                vid = __float_as_uint(b.w);
                vpt = make_float2(a.x, a.y);
            }
        }
        if (ok && d_found <= best)
        {
            best = d_found;
            found = true;
            best_vid = vid;
            best_pt = vpt;
        }
    };
    for (;;)
    {
        if (busy && node == kNone)
        {
            out_dist[slot] = found ? best : INFINITY;
            if (out_vertex) out_vertex[slot] = found ? best_vid : kNone;
            if (out_point)
            {
                out_point[2 * (uint64_t)slot] = found ? best_pt.x : 0.0f;
                out_point[2 * (uint64_t)slot + 1] = found ? best_pt.y : 0.0f;
            }
            busy = false;
        }
        const unsigned idle = __ballot_sync(kFull, !busy);
        if (idle)
        {
            const uint32_t s = feeder_take(fd, idle, !busy, lane, n, counter);
            if (s != kNone)
            {
                slot = perm ? __ldg(perm + s) : s;
                p = load_point2(q, slot);
                flip = flipv ? (__ldg(flipv + slot) != 0) : false;
                best = rmax ? __ldg(rmax + slot) : INFINITY;
                found = false;
                busy = true;
                sp = 0;
                node = 0;
                if (v.n == 1)
                { // the root is the only leaf: its own box and cone gate the test (query.cuh:352-358 would read out of bounds, Q6)
                    const Box2 rb = v.aabbs[0];
                    const Cone2 rc = v.cones[0];
                    const float m = lbvh::mindist(rb, p);
                    const float c4[4] = {rc.axis.x, rc.axis.y, rc.half_angle, rc.radius};
                    if (m <= best * best && may_hold_silhouette2(c4, p, rb, m)) test_leaf(0, v.l2[0].owned);
                    node = kNone;
                }
            }
            if (fd.exhausted && __all_sync(kFull, !busy)) break;
        }
        if (node != kNone)
        {
            const Pair2 pr = load_pair(v.n2, node);
            float4 e, f;
            ld256(reinterpret_cast<const char *>(v.n2 + node) + 64, e, f);
            const float c0[4] = {e.x, e.y, e.z, e.w}, c1[4] = {f.x, f.y, f.z, f.w};
            const float m0 = lbvh::mindist(pr.b0, p), m1 = lbvh::mindist(pr.b1, p);
            const float best2 = best * best;
            const bool h0 = (m0 <= best2) && may_hold_silhouette2(c0, p, pr.b0, m0);
            const bool h1 = (m1 <= best2) && may_hold_silhouette2(c1, p, pr.b1, m1);
            const bool swap = m1 < m0;
            uint32_t next = kNone;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            {
                const bool second = (ch == 1) != swap;
                const bool h = second ? h1 : h0;
                const float m = second ? m1 : m0;
                const uint32_t r = second ? pr.r1 : pr.r0;
                if (!h || !(m <= best * best)) continue;
                if (r & kLeaf2) test_leaf(r & kRefIndex2, (r >> 29) & 3u);
                else if (next == kNone) next = r;
                else stk[sp++] = StackEntry{r, m};
            }
            while (next == kNone && sp > 0)
            {
                const StackEntry se = stk[--sp];
                if (se.key <= best * best) next = se.node;
            }
            node = next;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One query per WARP (small batches): the cooperative walk of query.cu (solo_closest / solo_silhouette) on the 2-D records.
// A batch that does not fill the machine runs as long as its most expensive query — the centre of a closed curve sees every
// segment at the same distance — and 32 lanes on one query shorten exactly that path.  Same predicates, same distance
// functions, same results.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWide2Stack = 512; // entries per warp; above kWide2Stack - 128 one entry is popped per step (growth <= tree depth)
__global__ void __launch_bounds__(kQueryThreads)
    k2_closest_wide(View2 v, const float *__restrict__ q, const uint32_t *__restrict__ perm, uint32_t n, uint32_t *__restrict__ out_idx,
                    float *__restrict__ out_dist, unsigned long long *counter)
{
    __shared__ StackEntry s_st[kQueryThreads / 32][kWide2Stack];
    const int lane = threadIdx.x & 31;
    StackEntry *st = s_st[threadIdx.x >> 5];
    const unsigned lt = (1u << lane) - 1u;
    for (;;)
    {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, 4ull);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        for (unsigned long long s = base; s < base + 4ull && s < n; ++s)
        {
            const uint32_t slot = perm ? __ldg(perm + s) : (uint32_t)s;
            const float2 p = load_point2(q, slot);
            float b2 = INFINITY;
            uint32_t bi = kNone;
            int sp = 0;
            if (v.n == 1)
            {
                float4 a, b;
                ld256(v.l2, a, b);
                const float d = dist_point_segment(make_float2(a.x, a.y), make_float2(a.z, a.w), p);
                b2 = d * d;
                bi = __float_as_uint(b.x);
            }
            else
            {
                if (lane == 0) st[0] = StackEntry{0u, 0.0f};
                sp = 1;
            }
            __syncwarp();
            while (sp > 0)
            {
                const int take = sp > kWide2Stack - 128 ? 1 : (sp < 32 ? sp : 32);
                sp -= take;
                StackEntry e = StackEntry{kNone, INFINITY};
                if (lane < take) e = st[sp + lane];
                __syncwarp();
                uint32_t cand = 0xFFFFFFFFu, cand_i = kNone, pr0 = kNone, pr1 = kNone;
                float pk0 = 0.0f, pk1 = 0.0f;
                if (e.node != kNone && (e.key < b2 || bi == kNone))
                {
                    const Pair2 pr = load_pair(v.n2, e.node);
                    const float m0 = lbvh::mindist(pr.b0, p), m1 = lbvh::mindist(pr.b1, p);
                    const bool far1 = m0 < m1; // the farther child is pushed first, so the nearer one is popped first
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
                    {
                        const bool one = (ch == 0) == far1;
                        const float m = one ? m1 : m0;
                        const uint32_t r = one ? pr.r1 : pr.r0;
                        if (!(m < b2 || bi == kNone)) continue;
                        if (r & kLeaf2)
                        {
                            float4 a, b;
                            ld256(v.l2 + (r & kRefIndex2), a, b);
                            float d = dist_point_segment(make_float2(a.x, a.y), make_float2(a.z, a.w), p);
                            d *= d;
                            if ((d < b2 || (bi == kNone && d <= b2)) && __float_as_uint(d) < cand)
                            {
                                cand = __float_as_uint(d);
                                cand_i = __float_as_uint(b.x);
                            }
                        }
                        else if (pr0 == kNone)
                        {
                            pr0 = r;
                            pk0 = m;
                        }
                        else
                        {
                            pr1 = r;
                            pk1 = m;
                        }
                    }
                }
                const unsigned c1 = __ballot_sync(kFull, pr0 != kNone), c2 = __ballot_sync(kFull, pr1 != kNone);
                const int off = sp + __popc(c1 & lt) + __popc(c2 & lt);
                if (pr0 != kNone) st[off] = StackEntry{pr0, pk0};
                if (pr1 != kNone) st[off + 1] = StackEntry{pr1, pk1};
                sp += __popc(c1) + __popc(c2);
                const uint32_t mn = __reduce_min_sync(kFull, cand);
                if (mn != 0xFFFFFFFFu)
                {
                    const int wl = __ffs(__ballot_sync(kFull, cand == mn)) - 1;
                    b2 = __uint_as_float(mn);
                    bi = __shfl_sync(kFull, cand_i, wl);
                }
                __syncwarp();
            }
            if (lane == 0)
            {
                if (out_idx) out_idx[slot] = bi;
                out_dist[slot] = sqrtf(b2);
            }
        }
    }
}
__global__ void __launch_bounds__(kQueryThreads)
    k2_silhouette_wide(View2 v, const float *__restrict__ q, const uint8_t *__restrict__ flipv, const float *__restrict__ rmax,
                       const uint32_t *__restrict__ perm, uint32_t n, float *__restrict__ out_dist, uint32_t *__restrict__ out_vertex,
                       float *__restrict__ out_point, unsigned long long *counter)
{
    __shared__ StackEntry s_st[kQueryThreads / 32][kWide2Stack];
    const int lane = threadIdx.x & 31;
    StackEntry *st = s_st[threadIdx.x >> 5];
    const unsigned lt = (1u << lane) - 1u;
    for (;;)
    {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, 4ull);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        for (unsigned long long s = base; s < base + 4ull && s < n; ++s)
        {
            const uint32_t slot = perm ? __ldg(perm + s) : (uint32_t)s;
            const float2 p = load_point2(q, slot);
            const bool flip = flipv ? (__ldg(flipv + slot) != 0) : false;
            float best = rmax ? __ldg(rmax + slot) : INFINITY;
            bool found = false;
            uint32_t best_vid = kNone, cvid = kNone; // c*: this lane's candidate of the current step
            float2 best_pt = make_float2(0.f, 0.f), cpt = make_float2(0.f, 0.f);
            // silhouette_distance_calculator over the owned vertices of leaf k (scene.cuh:518-541, 345-378), bound = `bound`
            auto test_leaf = [&](uint32_t k, uint32_t cnt, float bound, uint32_t &cand)
            {
                float max_r2 = bound * bound;
                for (uint32_t j = 0; j < cnt; ++j)
                {
                    float4 a, b;
                    ld256(reinterpret_cast<const char *>(v.l2 + k) + 32 + 32 * j, a, b);
                    const float2 view_dir = make_float2(p.x - a.x, p.y - a.y);
                    const float d = lbvh::length(view_dir);
                    if (0.0f >= max_r2 || d * d > max_r2) continue;
                    bool is_sil = __float_as_uint(b.z) != 3u;
                    if (!is_sil) is_sil = lbvh::is_silhouette_vertex(make_float2(a.z, a.w), make_float2(b.x, b.y), view_dir, d, flip);
                    if (is_sil && d * d <= max_r2 && d <= bound)
                    {
                        max_r2 = d * d;
                        bound = d;
                        cand = __float_as_uint(d);
                        cvid = __float_as_uint(b.w);
                        cpt = make_float2(a.x, a.y);
                    }
                }
            };
            int sp = 0;
            if (v.n == 1)
            {
                const Box2 rb = v.aabbs[0];
                const Cone2 rc = v.cones[0];
                const float m = lbvh::mindist(rb, p);
                const float c4[4] = {rc.axis.x, rc.axis.y, rc.half_angle, rc.radius};
                uint32_t cand = 0xFFFFFFFFu;
                if (m <= best * best && may_hold_silhouette2(c4, p, rb, m)) test_leaf(0, v.l2[0].owned, best, cand);
                if (cand != 0xFFFFFFFFu)
                {
                    best = __uint_as_float(cand);
                    found = true;
                    best_vid = cvid;
                    best_pt = cpt;
                }
            }
            else
            {
                if (lane == 0) st[0] = StackEntry{0u, 0.0f};
                sp = 1;
            }
            __syncwarp();
            while (sp > 0)
            {
                const int take = sp > kWide2Stack - 128 ? 1 : (sp < 32 ? sp : 32);
                sp -= take;
                StackEntry e = StackEntry{kNone, INFINITY};
                if (lane < take) e = st[sp + lane];
                __syncwarp();
                uint32_t cand = 0xFFFFFFFFu, pr0 = kNone, pr1 = kNone;
                float pk0 = 0.0f, pk1 = 0.0f;
                const float best2 = best * best;
                if (e.node != kNone && e.key <= best2)
                {
                    const Pair2 pr = load_pair(v.n2, e.node);
                    float4 ce, cf;
                    ld256(reinterpret_cast<const char *>(v.n2 + e.node) + 64, ce, cf);
                    const float c0[4] = {ce.x, ce.y, ce.z, ce.w}, c1[4] = {cf.x, cf.y, cf.z, cf.w};
                    const float m0 = lbvh::mindist(pr.b0, p), m1 = lbvh::mindist(pr.b1, p);
                    const bool h0 = (m0 <= best2) && may_hold_silhouette2(c0, p, pr.b0, m0);
                    const bool h1 = (m1 <= best2) && may_hold_silhouette2(c1, p, pr.b1, m1);
                    const bool far1 = m0 < m1;
                    float bound = best;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch)
                    {
                        const bool one = (ch == 0) == far1;
                        if (!(one ? h1 : h0)) continue;
                        const float m = one ? m1 : m0;
                        const uint32_t r = one ? pr.r1 : pr.r0;
                        if (r & kLeaf2)
                        {
                            test_leaf(r & kRefIndex2, (r >> 29) & 3u, bound, cand);
                            if (cand != 0xFFFFFFFFu) bound = __uint_as_float(cand);
                        }
                        else if (pr0 == kNone)
                        {
                            pr0 = r;
                            pk0 = m;
                        }
                        else
                        {
                            pr1 = r;
                            pk1 = m;
                        }
                    }
                }
                const unsigned c1m = __ballot_sync(kFull, pr0 != kNone), c2m = __ballot_sync(kFull, pr1 != kNone);
                const int off = sp + __popc(c1m & lt) + __popc(c2m & lt);
                if (pr0 != kNone) st[off] = StackEntry{pr0, pk0};
                if (pr1 != kNone) st[off + 1] = StackEntry{pr1, pk1};
                sp += __popc(c1m) + __popc(c2m);
                const uint32_t mn = __reduce_min_sync(kFull, cand);
                if (mn != 0xFFFFFFFFu)
                {
                    best = __uint_as_float(mn);
                    found = true;
                    const int wl = __ffs(__ballot_sync(kFull, cand == mn)) - 1;
                    best_vid = __shfl_sync(kFull, cvid, wl);
                    best_pt = make_float2(__shfl_sync(kFull, cpt.x, wl), __shfl_sync(kFull, cpt.y, wl));
                }
                __syncwarp();
            }
            if (lane == 0)
            {
                out_dist[slot] = found ? best : INFINITY;
                if (out_vertex) out_vertex[slot] = found ? best_vid : kNone;
                if (out_point)
                {
                    out_point[2 * (uint64_t)slot] = found ? best_pt.x : 0.0f;
                    out_point[2 * (uint64_t)slot + 1] = found ? best_pt.y : 0.0f;
                }
            }
        }
    }
}

// ray vs segments                                                                         query.cuh:79-169
SNCH_DI bool ray_segment(float2 p0, float2 p1, float2 org, float2 dir, float *t, float *s) // scene.cuh:543-577
{
    const float2 seg = make_float2(p1.x - p0.x, p1.y - p0.y);
    const float D = dir.x * (-seg.y) + dir.y * seg.x;
    if (fabsf(D) < FLT_EPSILON) return false;
    const float inv = __frcp_rn(D);
    const float2 w = make_float2(p0.x - org.x, p0.y - org.y);
    const float tt = (w.x * (-seg.y) - w.y * (-seg.x)) * inv;
    const float ss = (dir.x * w.y - dir.y * w.x) * inv;
    if (ss >= -1e-3f && ss <= 1.0f + 1e-3f && tt >= 0.0f)
    {
        *t = tt;
        *s = ss;
        return true;
    }
    return false;
}
template <bool kAnyHit>
__global__ void __launch_bounds__(kQueryThreads)
    k2_intersect(View2 v, const float *__restrict__ org, const float *__restrict__ dirs, const float *__restrict__ tmax,
                 const uint32_t *__restrict__ perm, uint32_t n, snch_hit *__restrict__ hits, uint8_t *__restrict__ out_found,
                 unsigned long long *counter)
{
    const int lane = threadIdx.x & 31;
    Feeder fd{0u, 0u, false};
    StackEntry stk[kStackDepth];
    int sp = 0;
    float2 o = make_float2(0.f, 0.f), d = make_float2(1.f, 0.f);
    lbvh::ray<float, 2> ry(o, d); // carries 1 / dir per axis: rebuilt once per query
    bool busy = false, found = false;
    float best_t = INFINITY, max_dist = INFINITY, hit_s = 0.0f;
    uint32_t best_obj = kNone, slot = kNone, node = kNone;
    auto test_leaf = [&](uint32_t k)
    {
        float4 a, b;
        ld256(v.l2 + k, a, b);
        float t, s;
        if (ray_segment(make_float2(a.x, a.y), make_float2(a.z, a.w), o, d, &t, &s) && t < max_dist && t < best_t)
        {
            best_t = t;
            hit_s = s;
            best_obj = __float_as_uint(b.x);
            found = true;
        }
    };
    for (;;)
    {
        if (busy && node == kNone)
        {
            if (out_found) out_found[slot] = found ? 1 : 0;
            if (!kAnyHit && hits) hits[slot] = snch_hit{found ? best_t : INFINITY, found ? hit_s : 0.0f, 0.0f, found ? best_obj : kNone};
            busy = false;
        }
        const unsigned idle = __ballot_sync(kFull, !busy);
        if (idle)
        {
            const uint32_t s = feeder_take(fd, idle, !busy, lane, n, counter);
            if (s != kNone)
            {
                slot = perm ? __ldg(perm + s) : s;
                o = load_point2(org, slot);
                d = load_point2(dirs, slot);
                ry = lbvh::ray<float, 2>(o, d);
                max_dist = tmax ? __ldg(tmax + slot) : INFINITY;
                best_t = INFINITY;
                best_obj = kNone;
                found = false;
                hit_s = 0.0f;
                busy = true;
                sp = 0;
                node = 0;
                if (v.n == 1)
                {
                    test_leaf(0);
                    node = kNone;
                }
            }
            if (fd.exhausted && __all_sync(kFull, !busy)) break;
        }
        if (node != kNone)
        {
            const Pair2 pr = load_pair(v.n2, node);
            float t0, t1;
            const bool h0 = lbvh::intersects_d(ry, pr.b0, max_dist, &t0) && !(t0 > best_t);
            const bool h1 = lbvh::intersects_d(ry, pr.b1, max_dist, &t1) && !(t1 > best_t);
            // The reference's order (query.cuh:128-160): the nearer hit child next (L on ties), the other — leaf or internal — on
            // the stack with its entry distance; a leaf is tested when the walk REACHES it, not where it is met.  The first hit
            // at a given t wins (t < best_t), so with segments meeting at a vertex, duplicated segments or a ray starting on the
            // polyline (t = +0 / -0) the order decides which segment — and which sign of zero — is reported.
            const bool swap = h0 && h1 && t1 < t0;
            uint32_t next = kNone;
            if (h0 && h1) stk[sp++] = swap ? StackEntry{pr.r0, t0} : StackEntry{pr.r1, t1};
            if (h0 || h1) next = (h0 && !swap) ? pr.r0 : pr.r1;
            for (;;)
            {
                if (next != kNone && (next & kLeaf2))
                {
                    test_leaf(next & kRefIndex2);
                    next = kNone;
                    if (kAnyHit && found) sp = 0;
                }
                if (next != kNone || sp == 0) break;
                const StackEntry se = stk[--sp];
                if (!(se.key > best_t)) next = se.node;
            }
            node = next;
        }
    }
}

// SampleSegmentInCircle: the 2-D instantiation of sample_object_in_sphere + sample_on_object   sample.cuh:7-92
__global__ void __launch_bounds__(kQueryThreads)
    k2_sample(View2 v, const float *__restrict__ sph, const float *__restrict__ rnd, uint32_t n, int32_t *__restrict__ out_idx,
              float *__restrict__ out_pdf, float *__restrict__ out_pt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const lbvh::sphere<float, 2> s(make_float2(__ldg(sph + 3 * (uint64_t)i), __ldg(sph + 3 * (uint64_t)i + 1)), __ldg(sph + 3 * (uint64_t)i + 2));
    float u = __ldg(rnd + 2 * (uint64_t)i);
    const float u1 = __ldg(rnd + 2 * (uint64_t)i + 1);
    float path = 1.0f;
    uint32_t ref = v.n == 1 ? (kLeaf2 | 0u) : 0u;
    int32_t idx = -1;
    float pdf = 0.0f;
    float2 pt = make_float2(0.f, 0.f);
    for (;;)
    {
        if (ref & kLeaf2)
        {
            float4 a, b;
            ld256(v.l2 + (ref & kRefIndex2), a, b);
            const float2 p1 = make_float2(a.x, a.y), p2 = make_float2(a.z, a.w);
            // scene<2>::intersect_sphere (scene.cuh:579-604)
            const float2 dd = make_float2(p2.x - p1.x, p2.y - p1.y);
            const float len_sq = dd.x * dd.x + dd.y * dd.y;
            float t = ((s.origin.x - p1.x) * dd.x + (s.origin.y - p1.y) * dd.y) / len_sq;
            t = fmaxf(0.0f, fminf(1.0f, t));
            const float dx = (p1.x + t * dd.x) - s.origin.x, dy = (p1.y + t * dd.y) - s.origin.y;
            if (dx * dx + dy * dy <= s.radius * s.radius)
            {
                idx = (int32_t)__float_as_uint(b.x);
                pdf = path / lbvh::length(make_float2(p1.x - p2.x, p1.y - p2.y)); // measurement_getter (scene.cuh:408-417)
                pt = lbvh::sample_line(p1, p2, u1);
            }
            break;
        }
        const Pair2 pr = load_pair(v.n2, ref);
        const float wl = lbvh::intersect_sphere(s, pr.b0) ? Scene2T::green_weight()(s.origin, lbvh::centroid(pr.b0)) : 0.0f;
        const float wr = lbvh::intersect_sphere(s, pr.b1) ? Scene2T::green_weight()(s.origin, lbvh::centroid(pr.b1)) : 0.0f;
        const float total = wl + wr;
        if (!(total > 0.0f)) break;
        const float pl = wl / total;
        if (u < pl)
        {
            u /= pl;
            path = pl * path;
            ref = pr.r0;
        }
        else
        {
            const float prr = 1.0f - pl;
            u = (u - pl) / prr;
            path = prr * path;
            ref = pr.r1;
        }
    }
    out_idx[i] = idx;
    out_pdf[i] = pdf;
    if (out_pt)
    {
        out_pt[2 * (uint64_t)i] = pt.x;
        out_pt[2 * (uint64_t)i + 1] = pt.y;
    }
}
__global__ void k2_fill_empty(uint64_t n, uint32_t *idx, float *dist, snch_hit *hits, uint8_t *found, int32_t *sidx, float *pdf, float *pt)
{
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (idx) idx[i] = kNone;
    if (dist) dist[i] = INFINITY;
    if (hits) hits[i] = snch_hit{INFINITY, 0.0f, 0.0f, kNone};
    if (found) found[i] = 0;
    if (sidx) sidx[i] = -1;
    if (pdf) pdf[i] = 0.0f;
    if (pt) pt[2 * i] = pt[2 * i + 1] = 0.0f;
}

enum Kind2
{
    K2_NULL,
    K2_HOST,
    K2_DEVICE
};
Kind2 kind_of(const void *p)
{
    if (!p) return K2_NULL;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return K2_HOST;
    }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? K2_DEVICE : K2_HOST;
}
} // namespace
} // namespace snch

// ---------------------------------------------------------------------------------------------------------------
// the handle
// ---------------------------------------------------------------------------------------------------------------
struct snch_scene2
{
    int device = 0;
    uint32_t n_verts = 0, n_segs = 0;
    bool have_silhouettes = false, built = false;
    std::vector<float2> verts_h;
    std::vector<int2> segs_h;
    std::vector<int4> sil_h;   // silhouette_vertex::indices per vertex
    std::vector<int2> owned_h; // line_segment::silhouette_indices per segment
    // device
    float2 *verts = nullptr;
    snch::SilT *sil = nullptr;
    snch::SegT *lines = nullptr;
    snch::RefNode *nodes = nullptr;
    snch::Box2 *aabbs = nullptr;
    snch::Cone2 *cones = nullptr;
    uint32_t *sorted_idx = nullptr, *morton = nullptr;
    snch::N2 *n2 = nullptr;
    snch::L2 *l2 = nullptr;
    int collision = 0;
    float build_ms = 0.0f;
    snch::QueryTuning tuning;
    cudaMemPool_t pool = nullptr; // per-call scratch and staging: stream-ordered, freed blocks stay cached (no driver allocation per call)
    std::mutex mu;
    void free_device()
    {
        for (void *p : {(void *)verts, (void *)sil, (void *)lines, (void *)nodes, (void *)aabbs, (void *)cones, (void *)sorted_idx, (void *)morton,
                        (void *)n2, (void *)l2})
            if (p) cudaFree(p);
        verts = nullptr, sil = nullptr, lines = nullptr, nodes = nullptr, aabbs = nullptr, cones = nullptr, sorted_idx = nullptr, morton = nullptr;
        n2 = nullptr, l2 = nullptr;
    }
    snch::View2 view() const { return snch::View2{n_segs, n2, l2, aabbs, cones}; }
};

namespace snch
{
namespace
{
struct DeviceGuard
{
    int prev = 0;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
template <typename T> int dev_alloc(T **p, uint64_t count)
{
    if (cudaMalloc((void **)p, (count ? count : 1) * sizeof(T)) != cudaSuccess)
    {
        cudaGetLastError();
        set_error("snch_scene2: out of device memory");
        return SNCH_ERR_OOM;
    }
    return SNCH_OK;
}

// Host or device buffers of one batch: host arrays are staged through stream-ordered device allocations and copied back
// before the call returns (the call then synchronises `st`); device arrays are used in place.
int ensure_pool2(snch_scene2 *s)
{
    std::lock_guard<std::mutex> lock(s->mu);
    if (s->pool) return SNCH_OK;
    cudaMemPoolProps props;
    std::memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = s->device;
    SNCH_CUDA(cudaMemPoolCreate(&s->pool, &props));
    uint64_t keep = ~0ull;
    SNCH_CUDA(cudaMemPoolSetAttribute(s->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    return SNCH_OK;
}
struct Stage2
{
    cudaStream_t st;
    cudaMemPool_t pool;
    bool host = false, device = false;
    int status = SNCH_OK;
    struct Out
    {
        void *host, *dev;
        size_t bytes;
    };
    std::vector<void *> owned;
    std::vector<Out> outs;
    Stage2(cudaStream_t s, cudaMemPool_t p) : st(s), pool(p) {}
    void note(Kind2 k)
    {
        if (k == K2_HOST) host = true;
        if (k == K2_DEVICE) device = true;
    }
    template <typename T> const T *in(const T *p, size_t bytes)
    {
        const Kind2 k = kind_of(p);
        note(k);
        if (k != K2_HOST || status != SNCH_OK) return p;
        void *d = nullptr;
        if (cudaMallocFromPoolAsync(&d, bytes ? bytes : 16, pool, st) != cudaSuccess)
        {
            cudaGetLastError();
            set_error("out of device memory for the staged query batch");
            status = SNCH_ERR_OOM;
            return nullptr;
        }
        owned.push_back(d);
        cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, st);
        return (const T *)d;
    }
    template <typename T> T *out(T *p, size_t bytes)
    {
        const Kind2 k = kind_of(p);
        note(k);
        if (k != K2_HOST || status != SNCH_OK) return p;
        void *d = nullptr;
        if (cudaMallocFromPoolAsync(&d, bytes ? bytes : 16, pool, st) != cudaSuccess)
        {
            cudaGetLastError();
            set_error("out of device memory for the staged query batch");
            status = SNCH_ERR_OOM;
            return nullptr;
        }
        owned.push_back(d);
        outs.push_back(Out{p, d, bytes});
        return (T *)d;
    }
    bool mixed() const { return host && device; }
    int finish()
    {
        for (const Out &o : outs) cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, st);
        for (void *d : owned) cudaFreeAsync(d, st);
        owned.clear();
        if (host) SNCH_CUDA(cudaStreamSynchronize(st));
        SNCH_CUDA(cudaGetLastError());
        return SNCH_OK;
    }
    ~Stage2()
    {
        for (void *d : owned) cudaFreeAsync(d, st);
    }
};
struct Scratch2
{
    cudaStream_t st;
    unsigned char *p = nullptr;
    Scratch2(cudaStream_t s, cudaMemPool_t pool, uint64_t bytes) : st(s)
    {
        if (cudaMallocFromPoolAsync((void **)&p, bytes ? bytes : 256, pool, st) != cudaSuccess)
        {
            cudaGetLastError();
            p = nullptr;
        }
    }
    ~Scratch2()
    {
        if (p) cudaFreeAsync(p, st);
    }
};
int check_ready(const snch_scene2 *s, uint64_t n, const char *what)
{
    if (!s)
    {
        set_error(std::string(what) + ": null scene");
        return SNCH_ERR_INVALID;
    }
    if (!s->built)
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    if (n > 0xFFF00000ull)
    {
        set_error(std::string(what) + ": at most 2^32 - 2^20 queries per call");
        return SNCH_ERR_INVALID;
    }
    DeviceGuard g(s->device);
    return ensure_pool2(const_cast<snch_scene2 *>(s)); // the pool is a cache, not scene state
}
} // namespace
} // namespace snch

using namespace snch;

extern "C" int snch_scene2_create(const float *xy, uint32_t n_verts, const int32_t *seg, uint32_t n_segs, int device, snch_scene2 **out)
{
    if (!out || (n_verts && !xy) || (n_segs && !seg))
    {
        set_error("snch_scene2_create: null argument");
        return SNCH_ERR_INVALID;
    }
    if (n_segs > kMaxSegments2)
    {
        set_error("snch_scene2_create: too many segments for 29-bit leaf references");
        return SNCH_ERR_INVALID;
    }
    for (uint64_t i = 0; i < 2ull * n_segs; ++i)
        if (seg[i] < 0 || (uint32_t)seg[i] >= n_verts)
        {
            set_error("snch_scene2_create: segment vertex index out of range");
            return SNCH_ERR_INVALID;
        }
    snch_scene2 *s = new (std::nothrow) snch_scene2();
    if (!s) return SNCH_ERR_OOM;
    s->device = device;
    s->tuning.wide_max_n_sil = 131072; // "query.wide_max_n" of a 2-D scene
    s->tuning.sort_rays = 0;           // 2-D ray batches keep the caller's order unless asked ("query.sort_rays")
    s->n_verts = n_verts;
    s->n_segs = n_segs;
    s->verts_h.resize(n_verts);
    s->segs_h.resize(n_segs);
    for (uint32_t i = 0; i < n_verts; ++i) s->verts_h[i] = make_float2(xy[2 * i], xy[2 * i + 1]);
    for (uint32_t i = 0; i < n_segs; ++i) s->segs_h[i] = make_int2(seg[2 * i], seg[2 * i + 1]);
    *out = s;
    return SNCH_OK;
}
extern "C" int snch_scene2_destroy(snch_scene2 *s)
{
    if (!s) return SNCH_OK;
    {
        DeviceGuard g(s->device);
        s->free_device();
        if (s->pool)
        {
            cudaDeviceSynchronize();
            cudaMemPoolDestroy(s->pool);
        }
    }
    delete s;
    return SNCH_OK;
}
// one record per VERTEX: (previous, self, next) from the segments that touch it, later segments overwrite   scene.cuh:634-656
extern "C" int snch_scene2_compute_silhouettes(snch_scene2 *s)
{
    if (!s)
    {
        set_error("snch_scene2_compute_silhouettes: null scene");
        return SNCH_ERR_INVALID;
    }
    s->sil_h.assign(s->n_verts, make_int4(-1, -1, -1, -1));
    for (const int2 &sg : s->segs_h)
    {
        int4 &from = s->sil_h[sg.x];
        from.y = sg.x;
        from.z = sg.y;
        int4 &to = s->sil_h[sg.y];
        to.x = sg.x;
        to.y = sg.y;
    }
    s->have_silhouettes = true;
    return SNCH_OK;
}
extern "C" int snch_scene2_build(snch_scene2 *s, snch_stream stream)
{
    if (!s)
    {
        set_error("snch_scene2_build: null scene");
        return SNCH_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || s->device >= count)
    {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return SNCH_ERR_CUDA;
    }
    if (!s->have_silhouettes)
    { // the reference's build_bvh() reads silhouettes_d; without compute_silhouettes() every vertex record is the default one
        s->sil_h.assign(s->n_verts, make_int4(-1, -1, -1, -1));
    }
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    s->free_device();
    s->built = false;
    const uint32_t n = s->n_segs, nv = s->n_verts;
    // a vertex is owned by the first segment (input order) that touches it                              scene.cuh:658-681
    s->owned_h.assign(n, make_int2(-1, -1));
    {
        std::vector<char> seen(nv, 0);
        for (uint32_t i = 0; i < n; ++i)
        {
            const int vs[2] = {s->segs_h[i].x, s->segs_h[i].y};
            int k = 0;
            for (int j = 0; j < 2; ++j)
                if (!seen[vs[j]])
                {
                    seen[vs[j]] = 1;
                    (k == 0 ? s->owned_h[i].x : s->owned_h[i].y) = vs[j];
                    ++k;
                }
        }
    }
    int rc;
    if ((rc = dev_alloc(&s->verts, nv)) || (rc = dev_alloc(&s->sil, nv)) || (rc = dev_alloc(&s->lines, n))) return rc;
    const uint32_t nn = n ? 2 * n - 1 : 0;
    if ((rc = dev_alloc(&s->nodes, nn)) || (rc = dev_alloc(&s->aabbs, nn)) || (rc = dev_alloc(&s->cones, nn)) || (rc = dev_alloc(&s->sorted_idx, n)) ||
        (rc = dev_alloc(&s->morton, n)) || (rc = dev_alloc(&s->n2, n ? n - 1 : 0)) || (rc = dev_alloc(&s->l2, n)))
        return rc;
    std::vector<SilT> sil(nv);
    for (uint32_t i = 0; i < nv; ++i) sil[i] = SilT(s->sil_h[i], s->verts);
    std::vector<SegT> lines(n);
    for (uint32_t i = 0; i < n; ++i) lines[i] = SegT(s->segs_h[i], s->owned_h[i], s->verts, s->sil);
    cudaEvent_t e0, e1;
    SNCH_CUDA(cudaEventCreate(&e0));
    SNCH_CUDA(cudaEventCreate(&e1));
    SNCH_CUDA(cudaMemcpyAsync(s->verts, s->verts_h.data(), (size_t)nv * sizeof(float2), cudaMemcpyHostToDevice, st));
    SNCH_CUDA(cudaMemcpyAsync(s->sil, sil.data(), (size_t)nv * sizeof(SilT), cudaMemcpyHostToDevice, st));
    SNCH_CUDA(cudaMemcpyAsync(s->lines, lines.data(), (size_t)n * sizeof(SegT), cudaMemcpyHostToDevice, st));
    SNCH_CUDA(cudaStreamSynchronize(st)); // the staging vectors die with this scope
    SNCH_CUDA(cudaEventRecord(e0, st));
    if (n > 0)
    {
        Box2 *leaf_boxes = nullptr;
        Cone2 *leaf_cones = nullptr;
        SNCH_CUDA(cudaMallocAsync((void **)&leaf_boxes, (size_t)n * sizeof(Box2), st));
        SNCH_CUDA(cudaMallocAsync((void **)&leaf_cones, (size_t)n * sizeof(Cone2), st));
        const unsigned g128 = (n + 127) / 128;
        k2_leaf_inputs<<<g128, 128, 0, st>>>(s->lines, n, leaf_boxes, leaf_cones);
        rc = snch_lbvh_build(2, n, leaf_boxes, leaf_cones, nullptr, s->nodes, s->aabbs, s->cones, s->sorted_idx, s->morton, &s->collision, st);
        cudaFreeAsync(leaf_boxes, st);
        cudaFreeAsync(leaf_cones, st);
        if (rc != SNCH_OK) return rc;
        k2_leaf_records<<<g128, 128, 0, st>>>(s->lines, s->sorted_idx, n, s->l2);
        if (n > 1) k2_node_records<<<(n - 1 + 127) / 128, 128, 0, st>>>(s->nodes, s->aabbs, s->cones, s->l2, n, s->n2);
    }
    SNCH_CUDA(cudaEventRecord(e1, st));
    SNCH_CUDA(cudaStreamSynchronize(st));
    SNCH_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&s->build_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    s->built = true;
    return SNCH_OK;
}
extern "C" int snch_scene2_stats(const snch_scene2 *s, snch_build_stats *out)
{
    if (!s || !out)
    {
        set_error("snch_scene2_stats: null argument");
        return SNCH_ERR_INVALID;
    }
    std::memset(out, 0, sizeof *out);
    out->num_objects = s->n_segs;
    out->num_nodes = s->built && s->n_segs ? 2 * s->n_segs - 1 : 0;
    out->num_edges = s->n_verts; // silhouette elements of a 2-D scene are its vertices
    out->num_vertices = s->n_verts;
    out->morton_collision = (uint32_t)s->collision;
    out->build_ms = s->build_ms;
    return SNCH_OK;
}
extern "C" int snch_scene2_device_repr(const snch_scene2 *s, snch_bvh_device_pod *out)
{
    if (!s || !out)
    {
        set_error("snch_scene2_device_repr: null argument");
        return SNCH_ERR_INVALID;
    }
    if (!s->built)
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    std::memset(out, 0, sizeof *out);
    out->num_objects = s->n_segs;
    out->num_nodes = s->n_segs ? 2 * s->n_segs - 1 : 0;
    if (s->n_segs)
    { // bvh.cuh:383-386: an empty tree has no arrays at all
        out->nodes = s->nodes, out->aabbs = s->aabbs, out->cones = s->cones, out->objects = s->lines;
    }
    out->vertices = s->verts;
    out->silhouettes = s->sil;
    out->num_vertices = s->n_verts;
    out->num_silhouettes = s->n_verts;
    return SNCH_OK;
}
extern "C" int snch_scene2_export(const snch_scene2 *s, int kind, void *host_dst, uint64_t bytes)
{
    if (!s || !host_dst)
    {
        set_error("snch_scene2_export: null argument");
        return SNCH_ERR_INVALID;
    }
    const uint64_t n = s->n_segs, nn = n ? 2 * n - 1 : 0;
    const void *src = nullptr;
    uint64_t need = 0;
    bool on_host = false;
    switch (kind)
    {
    case SNCH_EXPORT_NODES: src = s->nodes, need = nn * 16; break;
    case SNCH_EXPORT_AABBS: src = s->aabbs, need = nn * 16; break;
    case SNCH_EXPORT_CONES: src = s->cones, need = nn * 16; break;
    case SNCH_EXPORT_MORTON_SORTED: src = s->morton, need = n * 4; break;
    case SNCH_EXPORT_SORTED_INDEX: src = s->sorted_idx, need = n * 4; break;
    case SNCH_EXPORT_EDGES: src = s->sil_h.data(), need = (uint64_t)s->n_verts * 16, on_host = true; break;
    case SNCH_EXPORT_TRI_OWNED: src = s->owned_h.data(), need = n * 8, on_host = true; break;
    default: set_error("snch_scene2_export: kind not available for 2-D scenes"); return SNCH_ERR_INVALID;
    }
    if (on_host ? (kind == SNCH_EXPORT_EDGES ? !s->have_silhouettes && !s->built : !s->built) : !s->built)
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    if (bytes != need)
    {
        set_error("snch_scene2_export: destination size does not match the array");
        return SNCH_ERR_INVALID;
    }
    if (need == 0) return SNCH_OK;
    if (on_host) std::memcpy(host_dst, src, need);
    else
    {
        DeviceGuard g(s->device);
        SNCH_CUDA(cudaMemcpy(host_dst, src, need, cudaMemcpyDeviceToHost));
    }
    return SNCH_OK;
}
extern "C" int snch_scene2_set_option(snch_scene2 *s, const char *name, int64_t value)
{
    if (!s || !name)
    {
        set_error("snch_scene2_set_option: null argument");
        return SNCH_ERR_INVALID;
    }
    const std::string k(name);
    if (k == "query.sort_min_n") s->tuning.sort_min_n = (int)value;
    else if (k == "query.sort_bits") s->tuning.sort_bits = (int)value;
    else if (k == "query.sort_rays") s->tuning.sort_rays = (int)value;
    else if (k == "query.blocks_per_sm") s->tuning.blocks_per_sm = (int)value;
    else if (k == "query.wide_max_n") s->tuning.wide_max_n_sil = (int)value;
    else
    {
        set_error("snch_scene2_set_option: unknown option '" + k + "'");
        return SNCH_ERR_INVALID;
    }
    return SNCH_OK;
}

extern "C" int snch_closest_point_batch2(const snch_scene2 *s, const float *points_xy, uint64_t n, uint32_t *out_index, float *out_distance,
                                         snch_stream stream)
{
    int rc = check_ready(s, n, "snch_closest_point_batch2");
    if (rc != SNCH_OK) return rc;
    if (n == 0) return SNCH_OK;
    if (!points_xy || !out_distance)
    {
        set_error("snch_closest_point_batch2: null argument");
        return SNCH_ERR_INVALID;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Stage2 sg(st, s->pool);
    const float *q = sg.in(points_xy, n * 8);
    uint32_t *idx = sg.out(out_index, n * 4);
    float *dist = sg.out(out_distance, n * 4);
    if (sg.status != SNCH_OK) return sg.status;
    if (sg.mixed())
    {
        set_error("snch_closest_point_batch2: host and device pointers mixed in one call");
        return SNCH_ERR_POINTER_KIND;
    }
    if (s->n_segs == 0) k2_fill_empty<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, idx, dist, nullptr, nullptr, nullptr, nullptr, nullptr);
    else
    {
        Scratch2 scr(st, s->pool, query_scratch_bytes(n, s->tuning));
        if (!scr.p)
        {
            set_error("out of device memory for the per-call query scratch");
            return SNCH_ERR_OOM;
        }
        unsigned long long *counter;
        const uint32_t *perm;
        rc = prepare_batch(s->tuning, true, q, 2, nullptr, (uint32_t)n, scr.p, st, &counter, &perm, nullptr, 2);
        if (rc != SNCH_OK) return rc;
        if (n < (uint64_t)s->tuning.wide_max_n_sil) // 2-D scenes share one threshold (measured crossover 0.1-0.3M queries)
            k2_closest_wide<<<persistent_grid(k2_closest_wide, s->tuning, (uint32_t)(n * 16)), kQueryThreads, 0, st>>>(s->view(), q, perm, (uint32_t)n, idx,
                                                                                                                 dist, counter);
        else
            k2_closest<<<persistent_grid(k2_closest, s->tuning, (uint32_t)n), kQueryThreads, 0, st>>>(s->view(), q, perm, (uint32_t)n, idx, dist, counter);
    }
    return sg.finish();
}
extern "C" int snch_closest_silhouette_batch2(const snch_scene2 *s, const float *points_xy, const uint8_t *flip, const float *r_max, uint64_t n,
                                              float *out_distance, uint32_t *out_vertex, float *out_point_xy, snch_stream stream)
{
    int rc = check_ready(s, n, "snch_closest_silhouette_batch2");
    if (rc != SNCH_OK) return rc;
    if (n == 0) return SNCH_OK;
    if (!points_xy || !out_distance)
    {
        set_error("snch_closest_silhouette_batch2: null argument");
        return SNCH_ERR_INVALID;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Stage2 sg(st, s->pool);
    const float *q = sg.in(points_xy, n * 8);
    const uint8_t *fl = sg.in(flip, n);
    const float *rm = sg.in(r_max, n * 4);
    float *dist = sg.out(out_distance, n * 4);
    uint32_t *vid = sg.out(out_vertex, n * 4);
    float *pt = sg.out(out_point_xy, n * 8);
    if (sg.status != SNCH_OK) return sg.status;
    if (sg.mixed())
    {
        set_error("snch_closest_silhouette_batch2: host and device pointers mixed in one call");
        return SNCH_ERR_POINTER_KIND;
    }
    if (s->n_segs == 0) k2_fill_empty<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, vid, dist, nullptr, nullptr, nullptr, nullptr, pt);
    else
    {
        Scratch2 scr(st, s->pool, query_scratch_bytes(n, s->tuning));
        if (!scr.p)
        {
            set_error("out of device memory for the per-call query scratch");
            return SNCH_ERR_OOM;
        }
        unsigned long long *counter;
        const uint32_t *perm;
        rc = prepare_batch(s->tuning, true, q, 2, nullptr, (uint32_t)n, scr.p, st, &counter, &perm, nullptr, 2);
        if (rc != SNCH_OK) return rc;
        if (n < (uint64_t)s->tuning.wide_max_n_sil)
            k2_silhouette_wide<<<persistent_grid(k2_silhouette_wide, s->tuning, (uint32_t)(n * 16)), kQueryThreads, 0, st>>>(s->view(), q, fl, rm, perm,
                                                                                                                       (uint32_t)n, dist, vid, pt, counter);
        else
            k2_silhouette<<<persistent_grid(k2_silhouette, s->tuning, (uint32_t)n), kQueryThreads, 0, st>>>(s->view(), q, fl, rm, perm, (uint32_t)n, dist,
                                                                                                           vid, pt, counter);
    }
    return sg.finish();
}
extern "C" int snch_intersect_batch2(const snch_scene2 *s, const float *origins_xy, const float *dirs_xy, const float *t_max, uint64_t n,
                                     snch_hit *out_hits, uint8_t *out_found, int any_hit, snch_stream stream)
{
    int rc = check_ready(s, n, "snch_intersect_batch2");
    if (rc != SNCH_OK) return rc;
    if (n == 0) return SNCH_OK;
    if (!origins_xy || !dirs_xy || (!out_hits && !out_found) || (any_hit && !out_found))
    {
        set_error("snch_intersect_batch2: null argument");
        return SNCH_ERR_INVALID;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Stage2 sg(st, s->pool);
    const float *o = sg.in(origins_xy, n * 8);
    const float *d = sg.in(dirs_xy, n * 8);
    const float *tm = sg.in(t_max, n * 4);
    snch_hit *hits = any_hit ? nullptr : sg.out(out_hits, n * sizeof(snch_hit));
    uint8_t *found = sg.out(out_found, n);
    if (sg.status != SNCH_OK) return sg.status;
    if (sg.mixed())
    {
        set_error("snch_intersect_batch2: host and device pointers mixed in one call");
        return SNCH_ERR_POINTER_KIND;
    }
    if (s->n_segs == 0) k2_fill_empty<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, nullptr, nullptr, hits, found, nullptr, nullptr, nullptr);
    else
    {
        Scratch2 scr(st, s->pool, s->tuning.sort_rays != 0 ? query_scratch_bytes(n, s->tuning) : query_scratch_bytes(0, s->tuning)); // unordered: header only
        if (!scr.p)
        {
            set_error("out of device memory for the per-call query scratch");
            return SNCH_ERR_OOM;
        }
        unsigned long long *counter;
        const uint32_t *perm;
        rc = prepare_batch(s->tuning, s->tuning.sort_rays != 0, o, 2, nullptr, (uint32_t)n, scr.p, st, &counter, &perm, nullptr, 2);
        if (rc != SNCH_OK) return rc;
        if (any_hit)
            k2_intersect<true><<<persistent_grid(k2_intersect<true>, s->tuning, (uint32_t)n), kQueryThreads, 0, st>>>(s->view(), o, d, tm, perm, (uint32_t)n,
                                                                                                                 hits, found, counter);
        else
            k2_intersect<false><<<persistent_grid(k2_intersect<false>, s->tuning, (uint32_t)n), kQueryThreads, 0, st>>>(s->view(), o, d, tm, perm,
                                                                                                                   (uint32_t)n, hits, found, counter);
    }
    return sg.finish();
}
extern "C" int snch_sample_in_sphere_batch2(const snch_scene2 *s, const float *circles_xyr, const float *rnd2, uint64_t n, int32_t *out_index,
                                            float *out_pdf, float *out_point_xy, snch_stream stream)
{
    int rc = check_ready(s, n, "snch_sample_in_sphere_batch2");
    if (rc != SNCH_OK) return rc;
    if (n == 0) return SNCH_OK;
    if (!circles_xyr || !rnd2 || !out_index || !out_pdf)
    {
        set_error("snch_sample_in_sphere_batch2: null argument");
        return SNCH_ERR_INVALID;
    }
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    Stage2 sg(st, s->pool);
    const float *sp = sg.in(circles_xyr, n * 12);
    const float *rn = sg.in(rnd2, n * 8);
    int32_t *idx = sg.out(out_index, n * 4);
    float *pdf = sg.out(out_pdf, n * 4);
    float *pt = sg.out(out_point_xy, n * 8);
    if (sg.status != SNCH_OK) return sg.status;
    if (sg.mixed())
    {
        set_error("snch_sample_in_sphere_batch2: host and device pointers mixed in one call");
        return SNCH_ERR_POINTER_KIND;
    }
    if (s->n_segs == 0) k2_fill_empty<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, nullptr, nullptr, nullptr, nullptr, idx, pdf, pt);
    else k2_sample<<<(unsigned)((n + kQueryThreads - 1) / kQueryThreads), kQueryThreads, 0, st>>>(s->view(), sp, rn, (uint32_t)n, idx, pdf, pt);
    return sg.finish();
}
