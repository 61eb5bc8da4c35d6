// build.cu — LBVH + SNCH construction on the GPU.
//
// Computes what lbvh::bvh<...>::construct() computes (bvh.cuh:380-613) with hand-written kernels:
//   k_scene_box   leaf boxes -> scene box (block reduction + ordered-int atomics)          [bvh.cuh:423-432]
//   k_morton      30-bit Morton code of each leaf-box centroid, (key, index) pairs         [bvh.cuh:436-449]
//   radix sort    stable LSD sort of 8-byte pairs (payload is NOT dragged along)           [bvh.cuh:452-454]
//   k_hierarchy   Karras 2012 ranges/splits on the augmented key (morton<<32 | index)      [bvh.cuh:109-229,459-515]
//   k_refit       leaf boxes/cones in sorted order, then ONE bottom-up climb that merges boxes AND normal cones and
//                 emits the 64 B / 96 B two-child traversal records                       [bvh.cuh:520-604]
// The augmented key reproduces the reference topology on both of its paths (unique 32-bit codes, or its 64-bit fallback
// on collisions) — SURVEY 7 step 5.
#include "scene.h"
#include "snch_math.cuh"
#include "sort_scan.cuh"
#include "build_ctx.h"

#include <chrono>
#include <cstdio>
#include <cstring>

namespace snch
{

// ---------------------------------------------------------------------------------------------------------------
// host: edge adjacency (scene.cuh:1135-1229).  Same numbering and ownership as the reference (first-seen edge ids,
// last-writer-wins face slots, first triangle in input order owns an edge), O(N) with an open-addressing table.
// ---------------------------------------------------------------------------------------------------------------
void compute_adjacency_host(snch_scene *s)
{
    const auto t0 = std::chrono::steady_clock::now();
    const uint32_t n = s->n_tris;
    const int32_t *tri = s->h_tri.data();
    s->h_tri_edges.assign((size_t)3 * n, -1);
    s->h_tri_owned.assign((size_t)3 * n, -1);
    uint64_t cap = 16;
    while (cap < (uint64_t)6 * n + 16) cap <<= 1;
    std::vector<uint64_t> keys(cap, ~0ull);
    std::vector<int32_t> vals(cap, -1);
    int32_t E = 0;
    for (uint32_t i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j)
        {
            int32_t I = tri[3 * i + j], J = tri[3 * i + (j + 1) % 3];
            if (I > J) std::swap(I, J);
            const uint64_t key = ((uint64_t)(uint32_t)I << 32) | (uint32_t)J;
            uint64_t h = (key * 0x9E3779B97F4A7C15ull) >> 17;
            for (;;)
            {
                h &= cap - 1;
                if (keys[h] == key) break;
                if (keys[h] == ~0ull)
                {
                    keys[h] = key;
                    vals[h] = E++;
                    break;
                }
                ++h;
            }
            s->h_tri_edges[3 * (size_t)i + j] = vals[h];
        }
    s->n_edges = (uint32_t)E;
    s->h_edges4.assign((size_t)4 * E, -1);
    std::vector<uint8_t> seen((size_t)E, 0);
    for (uint32_t i = 0; i < n; ++i)
    {
        const int32_t *vi = tri + 3 * (size_t)i;
        int p = 0;
        for (int j = 0; j < 3; ++j)
        {
            const int I = j - 1 < 0 ? 2 : j - 1;
            int J = j, K = j + 1 > 2 ? 0 : j + 1;
            int orientation = 1;
            if (vi[J] > vi[K])
            {
                std::swap(J, K);
                orientation = -1;
            }
            const int32_t e = s->h_tri_edges[3 * (size_t)i + j];
            int32_t *se = &s->h_edges4[4 * (size_t)e];
            se[orientation == 1 ? 0 : 3] = vi[I];
            se[1] = vi[J];
            se[2] = vi[K];
            if (!seen[e])
            {
                seen[e] = 1;
                s->h_tri_owned[3 * (size_t)i + p++] = e;
            }
        }
    }
    s->silhouettes_done = true;
    s->adjacency_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------------------------------
// arena layout
// ---------------------------------------------------------------------------------------------------------------
void layout_arena(ArenaHeader &h, uint32_t nV, uint32_t nT, uint32_t nE)
{
    std::memset(&h, 0, sizeof h);
    h.magic = kArenaMagic;
    h.version = kArenaVersion;
    h.n_tris = nT;
    h.n_verts = nV;
    h.n_edges = nE;
    h.n_nodes = nT ? 2 * nT - 1 : 0;
    h.n_internal = nT ? nT - 1 : 0;
    const uint64_t n_rec = nT > 1 ? nT - 1 : (nT ? 1 : 0); // traversal records (a 1-leaf tree gets one dummy root)
    uint64_t off = align_up(sizeof(ArenaHeader), 256);
    auto take = [&](uint64_t bytes)
    {
        const uint64_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    h.off_vertices = take((uint64_t)nV * sizeof(float3));
    h.off_edges = take((uint64_t)nE * sizeof(RefEdge));
    h.off_objects = take((uint64_t)nT * sizeof(RefTriangle));
    h.off_tri_edges = take((uint64_t)nT * sizeof(int3));
    h.off_nodes = take((uint64_t)h.n_nodes * sizeof(RefNode));
    h.off_aabbs = take((uint64_t)h.n_nodes * sizeof(RefAabb));
    h.off_cones = take((uint64_t)h.n_nodes * sizeof(RefCone));
    h.off_morton = take((uint64_t)nT * 4);
    h.off_sorted_idx = take((uint64_t)nT * 4);
    h.off_ranges = take((uint64_t)h.n_internal * 8);
    h.off_q1 = take((uint64_t)h.n_nodes);
    h.off_bnode = take(n_rec * sizeof(BNode));
    h.off_snode = take(n_rec * sizeof(SNode));
    h.off_ltri = take((uint64_t)nT * sizeof(LTri));
    h.off_ledge = take((uint64_t)nE * sizeof(LEdge));
    h.off_edge_off = take((uint64_t)nT * 4);
    h.total_bytes = off;
}

void resolve_view(snch_scene *s)
{
    const ArenaHeader &h = s->hdr;
    unsigned char *b = s->arena;
    SceneView &v = s->view;
    v.n_tris = h.n_tris;
    v.n_verts = h.n_verts;
    v.n_edges = h.n_edges;
    v.n_internal = h.n_internal;
    v.vertices = (const float3 *)(b + h.off_vertices);
    v.edges = (const RefEdge *)(b + h.off_edges);
    v.objects = (const RefTriangle *)(b + h.off_objects);
    v.bnode = (const BNode *)(b + h.off_bnode);
    v.snode = (const SNode *)(b + h.off_snode);
    v.ltri = (const LTri *)(b + h.off_ltri);
    v.ledge = (const LEdge *)(b + h.off_ledge);
    v.edge_off = (const uint32_t *)(b + h.off_edge_off);
}

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
SNCH_DI int f2ord(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
SNCH_DI float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }
SNCH_DI V3 ldv(const float3 *v, int i)
{
    const float3 p = v[i];
    return V3{p.x, p.y, p.z};
}

__global__ void k_init_box(int *box)
{
    if (threadIdx.x < 3) box[threadIdx.x] = f2ord(INFINITY);
    else if (threadIdx.x < 6) box[threadIdx.x] = f2ord(-INFINITY);
}

__global__ void __launch_bounds__(256) k_scene_box(BuildCtx c)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n; i += gridDim.x * blockDim.x)
    {
        const int3 vi = c.objects[i].v;
        const Box b = tri_box(ldv(c.verts, vi.x), ldv(c.verts, vi.y), ldv(c.verts, vi.z));
        lo[0] = fminf(lo[0], b.lo.x);
        lo[1] = fminf(lo[1], b.lo.y);
        lo[2] = fminf(lo[2], b.lo.z);
        hi[0] = fmaxf(hi[0], b.hi.x);
        hi[1] = fmaxf(hi[1], b.hi.y);
        hi[2] = fmaxf(hi[2], b.hi.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    __shared__ float sm[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
        for (int a = 0; a < 3; ++a)
        {
            sm[warp][a] = lo[a];
            sm[warp][3 + a] = hi[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6)
    {
        float v = sm[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fminf(v, sm[w][threadIdx.x]) : fmaxf(v, sm[w][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&c.scene_box[threadIdx.x], f2ord(v));
        else atomicMax(&c.scene_box[threadIdx.x], f2ord(v));
    }
}

__global__ void __launch_bounds__(256) k_morton(BuildCtx c)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    const V3 wlo = V3{ord2f(c.scene_box[0]), ord2f(c.scene_box[1]), ord2f(c.scene_box[2])};
    const V3 whi = V3{ord2f(c.scene_box[3]), ord2f(c.scene_box[4]), ord2f(c.scene_box[5])};
    const int3 vi = c.objects[i].v;
    const Box b = tri_box(ldv(c.verts, vi.x), ldv(c.verts, vi.y), ldv(c.verts, vi.z));
    c.morton[i] = morton30(b, wlo, whi);
    c.sorted_idx[i] = i;
}

// common_upper_bits on the augmented 64-bit key (morton<<32 | object index)        morton_code.cuh:144-167
SNCH_DI int key_delta(const uint32_t *__restrict__ m, const uint32_t *__restrict__ id, int n, uint32_t mi, uint32_t ii, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t mj = m[j];
    return mi != mj ? __clz(mi ^ mj) : 32 + __clz(ii ^ id[j]);
}

// Karras 2012: bvh.cuh:109-167 (determine_range), :169-199 (find_split), :200-229 (children + parent links)
__global__ void __launch_bounds__(256) k_hierarchy(BuildCtx c)
{
    const int n = (int)c.n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const uint32_t *__restrict__ m = c.morton;
    const uint32_t *__restrict__ id = c.sorted_idx;
    const uint32_t mi = m[i], ii = id[i];
    if (mi == m[i + 1]) c.counters[0] = 1u; // duplicate Morton codes: the reference's 64-bit path (bvh.cuh:459-476)
    int first, last;
    if (i == 0)
    {
        first = 0;
        last = n - 1;
        c.nodes[0].parent = kNone;
    }
    else
    {
        const int dl = key_delta(m, id, n, mi, ii, i - 1), dr = key_delta(m, id, n, mi, ii, i + 1);
        const int d = dr > dl ? 1 : -1;
        const int dmin = min(dl, dr);
        int lmax = 2;
        while (key_delta(m, id, n, mi, ii, i + d * lmax) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t > 0; t >>= 1)
            if (key_delta(m, id, n, mi, ii, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d;
        first = min(i, j);
        last = max(i, j);
    }
    // split: highest differing bit between first and last key
    const uint32_t mf = m[first], idf = id[first];
    const int dnode = key_delta(m, id, n, mf, idf, last);
    int split = first, stride = last - first;
    do
    {
        stride = (stride + 1) >> 1;
        const int mid = split + stride;
        if (mid < last && key_delta(m, id, n, mf, idf, mid) > dnode) split = mid;
    } while (stride > 1);
    uint32_t left = (uint32_t)split, right = (uint32_t)split + 1;
    if (first == split) left += (uint32_t)(n - 1);
    if (last == split + 1) right += (uint32_t)(n - 1);
    c.nodes[i].left = left;
    c.nodes[i].right = right;
    c.nodes[i].object = kNone;
    c.nodes[left].parent = (uint32_t)i;
    c.nodes[right].parent = (uint32_t)i;
    c.ranges[i] = make_uint2((uint32_t)first, (uint32_t)last);
}

SNCH_DI void store_aabb(RefAabb *dst, Box b)
{
    float2 *p = reinterpret_cast<float2 *>(dst);
    p[0] = make_float2(b.hi.x, b.hi.y);
    p[1] = make_float2(b.hi.z, b.lo.x);
    p[2] = make_float2(b.lo.y, b.lo.z);
}
SNCH_DI Box load_aabb_cg(const RefAabb *src)
{
    const float2 *p = reinterpret_cast<const float2 *>(src);
    const float2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
    Box r;
    r.hi = V3{a.x, a.y, b.x};
    r.lo = V3{b.y, c.x, c.y};
    return r;
}
SNCH_DI void store_cone(RefCone *dst, Cone c)
{
    float *p = reinterpret_cast<float *>(dst);
    p[0] = c.axis.x;
    p[1] = c.axis.y;
    p[2] = c.axis.z;
    p[3] = c.half_angle;
    p[4] = c.radius;
}
SNCH_DI Cone load_cone_cg(const RefCone *src)
{
    const float *p = reinterpret_cast<const float *>(src);
    Cone c;
    c.axis = V3{__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)};
    c.half_angle = __ldcg(p + 3);
    c.radius = __ldcg(p + 4);
    return c;
}
// `left` / `right` are the children's node ids (internal < n_internal <= leaf), as in RefNode
SNCH_DI void store_records(BNode *bn, SNode *sn, Box lb, Box rb, Cone lc, Cone rc, uint32_t lbref, uint32_t rbref, uint32_t lsref, uint32_t rsref,
                           uint32_t parent)
{
    const float4 a = make_float4(lb.lo.x, lb.lo.y, lb.lo.z, lb.hi.x);
    const float4 b = make_float4(lb.hi.y, lb.hi.z, rb.lo.x, rb.lo.y);
    const float4 c = make_float4(rb.lo.z, rb.hi.x, rb.hi.y, rb.hi.z);
    float4 *bp = reinterpret_cast<float4 *>(bn);
    bp[0] = a;
    bp[1] = b;
    bp[2] = c;
    bp[3] = make_float4(__uint_as_float(lbref), __uint_as_float(rbref), __uint_as_float(parent), 0.0f);
    float4 *sp = reinterpret_cast<float4 *>(sn);
    sp[0] = a;
    sp[1] = b;
    sp[2] = c;
    sp[3] = make_float4(lc.axis.x, lc.axis.y, lc.axis.z, lc.half_angle);
    sp[4] = make_float4(lc.radius, rc.axis.x, rc.axis.y, rc.axis.z);
    sp[5] = make_float4(rc.half_angle, rc.radius, __uint_as_float(lsref), __uint_as_float(rsref));
}

// Leaf part of the refit, one thread per leaf k in Morton order: leaf box (scene.cuh:870-885), leaf normal cone from the
// owned silhouette edges (scene.cuh:887-961), the LTri / LEdge traversal records, and the reference-layout leaf entries.
// Returns the packed edge reference (first_edge << 2 | count).
SNCH_DI uint32_t refit_leaf(const BuildCtx &c, uint32_t k, Box &box, Cone &cone)
{
    const uint32_t ni = c.n - 1;
    const uint32_t obj = c.sorted_idx[k];
    const RefTriangle t = c.objects[obj];
    const V3 pa = ldv(c.verts, t.v.x), pb = ldv(c.verts, t.v.y), pc = ldv(c.verts, t.v.z);
    box = tri_box(pa, pb, pc);
    {
        float4 *lt = reinterpret_cast<float4 *>(c.ltri + k);
        lt[0] = make_float4(pa.x, pa.y, pa.z, __uint_as_float(obj));
        lt[1] = make_float4(pb.x, pb.y, pb.z, 0.0f);
        lt[2] = make_float4(pc.x, pc.y, pc.z, 0.0f);
        // unit normal and the reach of the triangle from its first vertex: the cheap lower bound of tri_lower_bound2() (query.cu)
        const V3 ab = pb - pa, ac = pc - pa;
        const V3 nn = normalize(cross(ab, ac)); // NaN for a degenerate triangle: the bound then never rejects
        lt[3] = make_float4(nn.x, nn.y, nn.z, fmaxf(len(ab), len(ac)) * 1.000001f);
    }
    // ---- leaf normal cone from the owned silhouette edges, and their traversal records
    const V3 bc = box_centroid(box);
    cone.axis = V3{0.f, 0.f, 0.f};
    cone.half_angle = kPi;
    cone.radius = 0.f;
    // LEdge is indexed by EDGE ID: the reference numbers edges in first-seen order and gives each to the first triangle (in
    // input order) that touches it (scene.cuh:1135-1229), so the ids a triangle owns are consecutive — first owned id + s for
    // its s-th owned edge (checked by tests/test_host_adjacency.py) — and a leaf's edges need no offset table or scan
    const int owned[3] = {t.owned.x, t.owned.y, t.owned.z};
    const uint32_t eoff = owned[0] >= 0 ? (uint32_t)owned[0] : 0u;
    uint32_t cnt = 0;
    bool all_two = true;
    V3 fn0[3], fn1[3];
#pragma unroll
    for (int s = 0; s < 3; ++s)
    {
        if (owned[s] == -1) continue;
        const int4 id = c.edges[owned[s]].indices;
        const V3 ea = ldv(c.verts, id.y), eb = ldv(c.verts, id.z);
        const bool has0 = id.w != -1, has1 = id.x != -1;
        V3 n0 = V3{0.f, 0.f, 0.f}, n1 = V3{0.f, 0.f, 0.f}, nsum = V3{0.f, 0.f, 0.f};
        if (has0)
        { // silhouette_edge::normal(0): (pb-pa) x (pc-pa) with pa=v[1], pb=v[2], pc=v[3]      scene.cuh:747-772
            n0 = cross(eb - ea, ldv(c.verts, id.w) - ea);
            nsum = V3{nsum.x + n0.x, nsum.y + n0.y, nsum.z + n0.z};
        }
        if (has1)
        { // normal(1): pa=v[2], pb=v[1], pc=v[0]
            n1 = cross(ea - eb, ldv(c.verts, id.x) - eb);
            nsum = V3{nsum.x + n1.x, nsum.y + n1.y, nsum.z + n1.z};
        }
        const V3 en = normalize(nsum);
        cone.axis = V3{cone.axis.x + en.x, cone.axis.y + en.y, cone.axis.z + en.z};
        const V3 ec = V3{(ea.x + eb.x) / 2, (ea.y + eb.y) / 2, (ea.z + eb.z) / 2};
        cone.radius = std_max(cone.radius, len(ec - bc));
        all_two = all_two && has0 && has1;
        const V3 u0 = has0 ? normalize(n0) : V3{0.f, 0.f, 0.f};
        const V3 u1 = has1 ? normalize(n1) : V3{0.f, 0.f, 0.f};
        fn0[s] = u0;
        fn1[s] = u1;
        float4 *le = reinterpret_cast<float4 *>(c.ledge + owned[s]);
        const bool boundary = !(has0 && has1);
        le[0] = make_float4(ea.x, ea.y, ea.z, eb.x);
        le[1] = make_float4(eb.y, eb.z, u0.x, u0.y);
        le[2] = make_float4(u0.z, u1.x, u1.y, u1.z);
        // the edge's id (what out_edge of a silhouette query reports) and the boundary flag — a flag of its own: a zero-area face has
        // a NaN normal, which the reference feeds to its silhouette test (-> "not a silhouette"), not a missing face
        le[3] = make_float4(__int_as_float(owned[s]), __int_as_float(boundary ? 1 : 0), 0.0f, 0.0f);
        ++cnt;
    }
    if (cnt == 0) cone.half_angle = -kPi;
    else if (!all_two) cone.half_angle = kPi;
    else
    {
        const float an = len(cone.axis);
        if (an > FLT_EPSILON)
        {
            cone.axis = V3{cone.axis.x / an, cone.axis.y / an, cone.axis.z / an};
            cone.half_angle = 0.0f;
#pragma unroll
            for (int s = 0; s < 3; ++s)
            {
                if (owned[s] == -1) continue;
                cone.half_angle = std_max(cone.half_angle, lbvh::detail::acosf_host(std_max(-1.0f, std_min(1.0f, dot(cone.axis, fn0[s])))));
                cone.half_angle = std_max(cone.half_angle, lbvh::detail::acosf_host(std_max(-1.0f, std_min(1.0f, dot(cone.axis, fn1[s])))));
            }
        }
    }
    const uint32_t packed = (eoff << 2) | cnt;
    c.edge_off[k] = packed;

    const uint32_t cur = ni + k;
    store_aabb(c.aabbs + cur, box);
    store_cone(c.cones + cur, cone);
    c.nodes[cur].left = kNone;
    c.nodes[cur].right = kNone;
    c.nodes[cur].object = obj;
    c.q1[cur] = 0;
    return packed;
}

// single-leaf tree: dummy root record whose second child can never be entered (NaN box, invalid cone)
SNCH_DI void refit_single_leaf(const BuildCtx &c, Box box, Cone cone, uint32_t packed)
{
    c.nodes[0].parent = kNone;
    const float qnan = __int_as_float(0x7FC00000);
    Box nb;
    nb.lo = nb.hi = V3{qnan, qnan, qnan};
    Cone ncone;
    ncone.axis = V3{0.f, 0.f, 0.f};
    ncone.half_angle = -kPi;
    ncone.radius = 0.f;
    store_records(c.bnode, c.snode, box, nb, cone, ncone, kLeafFlag | 0u, kLeafFlag | 0u, kLeafFlag | packed, kLeafFlag | 0u,
                  kNone); // left = leaf 0, right = a leaf that does not exist behind an invalid cone
}

// The reference's bottom-up climb (bvh.cuh:520-554 boxes, :556-604 cones, fused; with the fences its cone pass lacks, Q7)
// from node `cur` whose box / cone / record references the thread holds: the second arriver at a parent merges.
SNCH_DI void refit_climb(const BuildCtx &c, uint32_t cur, Box box, Cone cone, uint32_t cur_bref, uint32_t cur_sref, bool taint,
                         uint32_t parent)
{
    const uint32_t ni = c.n - 1;
    while (parent != kNone)
    {
        // topology is immutable here: fetched while the arrival atomic is in flight
        const uint4 nd = __ldg(reinterpret_cast<const uint4 *>(c.nodes + parent)); // parent, left, right, object
        // one acq_rel atomic instead of fence / atomic / fence: releases this thread's stores to the first arriver's
        // sibling, acquires the sibling's stores for the second arriver (MEMBAR.ALL instead of two MEMBAR.SC)
        uint32_t old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(c.flags + parent) : "memory");
        if (old == 0) return; // first arriver: the sibling subtree is not finished yet
        const uint32_t l = nd.y, r = nd.z;
        const bool cur_is_left = (l == cur);
        const uint32_t sib = cur_is_left ? r : l;
        const Box sbox = load_aabb_cg(c.aabbs + sib);
        const Cone scone = load_cone_cg(c.cones + sib);
        const bool staint = __ldcg(c.q1 + sib) != 0;
        uint32_t sib_bref = sib, sib_sref = sib;
        if (sib >= ni)
        {
            sib_bref = kLeafFlag | (sib - ni);
            sib_sref = kLeafFlag | __ldcg(c.edge_off + (sib - ni));
        }
        const Box lb = cur_is_left ? box : sbox, rb = cur_is_left ? sbox : box;
        const Cone lc = cur_is_left ? cone : scone, rc = cur_is_left ? scone : cone;
        const Box pbx = box_merge(lb, rb);
        bool q1;
        const Cone pcn = cone_merge(lc, rc, box_centroid(lb), box_centroid(rb), box_centroid(pbx), &q1);
        if (q1) atomicAdd(c.counters + 1, 1u);
        taint = taint || staint || q1;
        store_aabb(c.aabbs + parent, pbx);
        store_cone(c.cones + parent, pcn);
        c.q1[parent] = taint ? 1 : 0;
        const uint32_t gp = nd.x;
        store_records(c.bnode + parent, c.snode + parent, lb, rb, lc, rc, cur_is_left ? cur_bref : sib_bref, cur_is_left ? sib_bref : cur_bref,
                      cur_is_left ? cur_sref : sib_sref, cur_is_left ? sib_sref : cur_sref, gp);
        cur = parent;
        box = pbx;
        cone = pcn;
        cur_bref = cur_sref = parent;
        parent = gp;
    }
}

// v1 refit ("build.refit_kernel" = 0): one thread per leaf, every thread climbs on its own.  profiles/r01k: the climb ran
// at 4.1 of 32 lanes and was 75% of the kernel's warp instructions; the two fences per level were 44% of its stall samples.
__global__ void __launch_bounds__(128) k_refit(BuildCtx c)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= c.n) return;
    Box box;
    Cone cone;
    const uint32_t packed = refit_leaf(c, k, box, cone);
    if (c.n == 1)
    {
        refit_single_leaf(c, box, cone, packed);
        return;
    }
    const uint32_t cur = c.n - 1 + k;
    refit_climb(c, cur, box, cone, kLeafFlag | k, kLeafFlag | packed, false, c.nodes[cur].parent);
}

// v2 refit (default): one CTA per kRefitLeaves consecutive leaves (Morton order).  An internal node whose Karras range lies
// inside the CTA's leaf interval has its index, both subtrees and its arrival flag inside the CTA, so those nodes — all but
// ~2% — are merged in ROUNDS out of shared memory: every thread that holds a finished node bumps its parent's flag (a
// shared-memory atomic, no fence), second arrivers merge, then the surviving holders are compacted to the low threads, so
// a round's merges run on full warps.  The CTA's slice of the topology (parent / children / range of internal nodes
// [b0, b1)) is staged in shared memory once, so a round touches global memory only to store its products.  Nodes whose
// parent's range leaves the interval ("escapes": the roots of the maximal in-CTA subtrees) are appended to a list that
// k_refit_top finishes with the global flags and fences of the v1 climb.  Every merge is the same pure function of the same
// two children as in v1, so all products are bit-identical (tests/test_gpu_build.py).
constexpr int kRefitLeaves = 128;
struct RefitShared
{
    float f[11][2 * kRefitLeaves]; // lo.xyz hi.xyz axis.xyz half_angle radius, by local id: leaves [0,B), internal B + (i - b0)
    uint32_t bref[2 * kRefitLeaves], sref[2 * kRefitLeaves];
    uint32_t n_parent[kRefitLeaves], n_left[kRefitLeaves], n_right[kRefitLeaves]; // topology of internal nodes b0 + j
    uint2 range[kRefitLeaves];
    uint32_t flag[kRefitLeaves];
    uint2 queue[kRefitLeaves]; // (local id, parent)
    uint8_t taint[2 * kRefitLeaves]; // Q1 taint of the subtree (exported as q1)
    uint32_t count[2];
};
SNCH_DI void sh_put(RefitShared &sh, uint32_t id, Box b, Cone cn, uint32_t bref, uint32_t sref)
{
    sh.f[0][id] = b.lo.x;
    sh.f[1][id] = b.lo.y;
    sh.f[2][id] = b.lo.z;
    sh.f[3][id] = b.hi.x;
    sh.f[4][id] = b.hi.y;
    sh.f[5][id] = b.hi.z;
    sh.f[6][id] = cn.axis.x;
    sh.f[7][id] = cn.axis.y;
    sh.f[8][id] = cn.axis.z;
    sh.f[9][id] = cn.half_angle;
    sh.f[10][id] = cn.radius;
    sh.bref[id] = bref;
    sh.sref[id] = sref;
}
SNCH_DI void sh_get(const RefitShared &sh, uint32_t id, Box &b, Cone &cn)
{
    b.lo = V3{sh.f[0][id], sh.f[1][id], sh.f[2][id]};
    b.hi = V3{sh.f[3][id], sh.f[4][id], sh.f[5][id]};
    cn.axis = V3{sh.f[6][id], sh.f[7][id], sh.f[8][id]};
    cn.half_angle = sh.f[9][id];
    cn.radius = sh.f[10][id];
}
// Register budget, measured (build ms @1M triangles): 6 CTAs/SM at 76 registers, no spills 0.497; **8 CTAs at 64 registers, 56 B of
// spills 0.488**; 10 CTAs at 48 registers, 308 B 0.566; 12 CTAs at 40 registers, 476 B 0.768.  The kernel's time is (waves of CTAs) x
// (rounds) x (latency of one cone merge on one thread), so it wants residency — until the spills of the merge eat it.
__global__ void __launch_bounds__(kRefitLeaves, 8) k_refit_coop(BuildCtx c)
{
    __shared__ RefitShared sh;
    constexpr uint32_t B = kRefitLeaves;
    const uint32_t tid = threadIdx.x;
    const uint32_t b0 = blockIdx.x * B, b1 = min(b0 + B, c.n);
    const uint32_t ni = c.n - 1;
    const uint32_t k = b0 + tid;
    sh.flag[tid] = 0;
    sh.taint[tid] = 0;
    if (tid < 2) sh.count[tid] = 0;
    if (k < ni)
    { // internal node k: its range ends or starts at leaf k, so only nodes [b0, b1) can lie inside the interval
        const uint4 nd = *reinterpret_cast<const uint4 *>(c.nodes + k); // parent, left, right, object
        sh.n_parent[tid] = nd.x;
        sh.n_left[tid] = nd.y;
        sh.n_right[tid] = nd.z;
        sh.range[tid] = c.ranges[k];
    }
    else sh.range[tid] = make_uint2(0u, 0xFFFFFFFFu); // never inside
    // item held by this thread: a finished node (local id) and its parent
    bool active = k < c.n;
    uint32_t cl = tid, parent = kNone;
    if (active)
    {
        Box box;
        Cone cone;
        const uint32_t packed = refit_leaf(c, k, box, cone);
        sh_put(sh, tid, box, cone, kLeafFlag | k, kLeafFlag | packed);
        if (c.n == 1)
        {
            refit_single_leaf(c, box, cone, packed);
            active = false;
        }
        else parent = c.nodes[ni + k].parent;
    }
    __syncthreads();
    for (uint32_t round = 0;; ++round)
    {
        if (active)
        {
            active = false;
            if (parent != kNone)
            {
                const uint32_t pj = parent - b0; // (wraps for parent < b0: then >= B)
                bool inside = false;
                if (pj < B)
                {
                    const uint2 rg = sh.range[pj];
                    inside = rg.x >= b0 && rg.y < b1;
                }
                const uint32_t cur = cl < B ? ni + b0 + cl : b0 + (cl - B);
                if (!inside) c.escapes[atomicAdd(c.counters + 2, 1u)] = cur;
                else if (atomicAdd(&sh.flag[pj], 1u) != 0)
                { // second arriver: the sibling's entry was written in an earlier round (or by the leaf pass)
                    const uint32_t l = sh.n_left[pj], r = sh.n_right[pj], gp = sh.n_parent[pj];
                    const bool cur_is_left = (l == cur);
                    const uint32_t sib = cur_is_left ? r : l;
                    const uint32_t sl = sib >= ni ? sib - ni - b0 : B + (sib - b0);
                    Box box, sbox;
                    Cone cone, scone;
                    sh_get(sh, cl, box, cone);
                    sh_get(sh, sl, sbox, scone);
                    const uint32_t l_id = cur_is_left ? cl : sl, r_id = cur_is_left ? sl : cl;
                    const Box lb = cur_is_left ? box : sbox, rb = cur_is_left ? sbox : box;
                    const Cone lc = cur_is_left ? cone : scone, rc = cur_is_left ? scone : cone;
                    const Box pbx = box_merge(lb, rb);
                    bool q1;
                    const Cone pcn = cone_merge(lc, rc, box_centroid(lb), box_centroid(rb), box_centroid(pbx), &q1);
                    if (q1) atomicAdd(c.counters + 1, 1u);
                    const bool taint = sh.taint[cl] || sh.taint[sl] || q1;
                    store_aabb(c.aabbs + parent, pbx);
                    store_cone(c.cones + parent, pcn);
                    c.q1[parent] = taint ? 1 : 0;
                    store_records(c.bnode + parent, c.snode + parent, lb, rb, lc, rc, sh.bref[l_id], sh.bref[r_id], sh.sref[l_id], sh.sref[r_id], gp);
                    const uint32_t pl = B + pj;
                    sh_put(sh, pl, pbx, pcn, parent, parent);
                    sh.taint[pl] = taint ? 1 : 0;
                    cl = pl;
                    parent = gp;
                    active = true;
                }
            }
        }
        // compact the surviving holders onto the low threads
        if (tid == 0) sh.count[(round + 1) & 1] = 0;
        if (active) sh.queue[atomicAdd(&sh.count[round & 1], 1u)] = make_uint2(cl, parent);
        __syncthreads();
        const uint32_t total = sh.count[round & 1];
        if (total == 0) break;
        active = tid < total;
        uint2 it = make_uint2(0u, kNone);
        if (active) it = sh.queue[tid];
        __syncthreads();
        cl = it.x;
        parent = it.y;
    }
}

// Finishes the refit above the CTA subtrees: one thread per escape (the root of a maximal in-CTA subtree, ~12 per CTA),
// each continuing the reference's climb with the global arrival flags.  Everything it reads was stored by k_refit_coop.
__global__ void __launch_bounds__(128) k_refit_top(BuildCtx c)
{
    const uint32_t n_esc = c.counters[2];
    const uint32_t ni = c.n - 1;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n_esc; e += gridDim.x * blockDim.x)
    {
        const uint32_t cur = c.escapes[e];
        const Box box = load_aabb_cg(c.aabbs + cur);
        const Cone cone = load_cone_cg(c.cones + cur);
        uint32_t bref = cur, sref = cur;
        if (cur >= ni)
        {
            bref = kLeafFlag | (cur - ni);
            sref = kLeafFlag | c.edge_off[cur - ni];
        }
        refit_climb(c, cur, box, cone, bref, sref, c.q1[cur] != 0, c.nodes[cur].parent);
    }
}

__global__ void k_patch_pointers(RefEdge *edges, uint32_t n_edges, RefTriangle *objects, uint32_t n_tris, const float3 *verts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_edges) edges[i].vertices = verts;
    if (i < n_tris)
    {
        objects[i].vertices = verts;
        objects[i].silhouettes = edges;
    }
}

// reference-layout edge / triangle structs from the device adjacency arrays (adjacency.cu)
__global__ void k_assemble_topology(const int32_t *__restrict__ tri, const int32_t *__restrict__ tri_edges, const int32_t *__restrict__ owned,
                                    const int4 *__restrict__ e4, uint32_t n_tris, uint32_t n_edges, RefEdge *edges, RefTriangle *objects,
                                    int32_t *tri_edges_out, const float3 *verts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_edges)
    {
        RefEdge r;
        r.indices = e4[i];
        r.vertices = verts;
        r.pad_ = 0;
        edges[i] = r;
    }
    if (i < n_tris)
    {
        RefTriangle t;
        t.v = make_int3(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
        t.owned = make_int3(owned[3 * i], owned[3 * i + 1], owned[3 * i + 2]);
        t.vertices = verts;
        t.silhouettes = edges;
        objects[i] = t;
        tri_edges_out[3 * i] = tri_edges[3 * i];
        tri_edges_out[3 * i + 1] = tri_edges[3 * i + 1];
        tri_edges_out[3 * i + 2] = tri_edges[3 * i + 2];
    }
}

int patch_pointers(snch_scene *s, cudaStream_t stream)
{
    const ArenaHeader &h = s->hdr;
    const uint32_t m = h.n_edges > h.n_tris ? h.n_edges : h.n_tris;
    if (m == 0) return SNCH_OK;
    k_patch_pointers<<<(m + 255) / 256, 256, 0, stream>>>((RefEdge *)(s->arena + h.off_edges), h.n_edges,
                                                           (RefTriangle *)(s->arena + h.off_objects), h.n_tris,
                                                           (const float3 *)(s->arena + h.off_vertices));
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
int build_device(snch_scene *s, cudaStream_t stream)
{
    SNCH_CUDA(cudaSetDevice(s->device));
    const uint32_t nV = s->n_verts, nT = s->n_tris, nE = s->n_edges;
    ArenaHeader h;
    layout_arena(h, nV, nT, nE);
    const uint32_t prev_collision = s->hdr.collision;
    // refit-only: Morton order, hierarchy and edge slots of the previous build stay; only geometry-dependent products change
    const bool refit = s->opt_refit_only && s->built && s->arena && s->arena_bytes == h.total_bytes && s->arena_has_topology;
    s->built = false; // set again only when every step below has succeeded: a failed rebuild never leaves a half-written tree queryable
    if (!s->arena || s->arena_bytes != h.total_bytes)
    {
        if (s->arena) cudaFree(s->arena);
        s->arena = nullptr;
        s->arena_bytes = 0;
        s->arena_has_topology = false;
        if (cudaMalloc(&s->arena, h.total_bytes) != cudaSuccess)
        {
            cudaGetLastError();
            s->arena = nullptr;
            set_error("cudaMalloc of the scene arena failed");
            return SNCH_ERR_OOM;
        }
        s->arena_bytes = h.total_bytes;
    }
    s->hdr = h;
    resolve_view(s);
    unsigned char *b = s->arena;

    // ---- uploads (host arrays -> arena); the reference does these in its constructors (scene.cuh:1131, bvh.cuh:330).
    // Topology (edges, triangles, per-triangle edge ids) is written once per compute_silhouettes(); later builds of the
    // same scene (snch_scene_update_vertices) only refresh the vertices.
    if (nV) SNCH_CUDA(cudaMemcpyAsync(b + h.off_vertices, s->h_xyz.data(), (size_t)nV * 12, cudaMemcpyHostToDevice, stream));
    std::vector<RefEdge> he;
    std::vector<RefTriangle> ho;
    if (!s->arena_has_topology && s->adjacency_on_device && nT)
    {
        const uint32_t m = nE > nT ? nE : nT;
        k_assemble_topology<<<(m + 255) / 256, 256, 0, stream>>>(s->adj_tri, s->adj_tri_edges, s->adj_tri_owned, s->adj_edges4, nT, nE,
                                                                  (RefEdge *)(b + h.off_edges), (RefTriangle *)(b + h.off_objects),
                                                                  (int32_t *)(b + h.off_tri_edges), (const float3 *)(b + h.off_vertices));
        SNCH_CUDA(cudaGetLastError());
    }
    else if (!s->arena_has_topology)
    {
        he.resize(nE);
        for (uint32_t e = 0; e < nE; ++e)
        {
            he[e].indices = make_int4(s->h_edges4[4 * e], s->h_edges4[4 * e + 1], s->h_edges4[4 * e + 2], s->h_edges4[4 * e + 3]);
            he[e].vertices = (const float3 *)(b + h.off_vertices);
            he[e].pad_ = 0;
        }
        ho.resize(nT);
        for (uint32_t i = 0; i < nT; ++i)
        {
            ho[i].v = make_int3(s->h_tri[3 * i], s->h_tri[3 * i + 1], s->h_tri[3 * i + 2]);
            ho[i].owned = make_int3(s->h_tri_owned[3 * i], s->h_tri_owned[3 * i + 1], s->h_tri_owned[3 * i + 2]);
            ho[i].vertices = (const float3 *)(b + h.off_vertices);
            ho[i].silhouettes = (const RefEdge *)(b + h.off_edges);
        }
        if (nE) SNCH_CUDA(cudaMemcpyAsync(b + h.off_edges, he.data(), (size_t)nE * sizeof(RefEdge), cudaMemcpyHostToDevice, stream));
        if (nT)
        {
            SNCH_CUDA(cudaMemcpyAsync(b + h.off_objects, ho.data(), (size_t)nT * sizeof(RefTriangle), cudaMemcpyHostToDevice, stream));
            SNCH_CUDA(cudaMemcpyAsync(b + h.off_tri_edges, s->h_tri_edges.data(), (size_t)nT * 12, cudaMemcpyHostToDevice, stream));
        }
    }
    SNCH_CUDA(cudaStreamSynchronize(stream)); // staging vectors go out of scope below; also isolates build_ms
    s->arena_has_topology = true;
    if (s->adjacency_on_device && s->adj)
    { // the arena now holds the only copy the library needs; keep the host vectors for exports if they were fetched
        free_adjacency(s);
    }
    if (nT == 0)
    {
        SNCH_CUDA(cudaMemcpyAsync(b, &s->hdr, sizeof(ArenaHeader), cudaMemcpyHostToDevice, stream));
        SNCH_CUDA(cudaStreamSynchronize(stream));
        s->built = true;
        s->build_ms = 0.f;
        return SNCH_OK;
    }

    // ---- scratch: sort ping-pong + sort/scan counters + flags + box + counters
    const uint64_t sort_elems = sort_scratch_elems(nT);
    uint64_t so = 0;
    auto stake = [&](uint64_t bytes)
    {
        const uint64_t o = so;
        so = align_up(so + bytes, 256);
        return o;
    };
    const uint64_t o_ktmp = stake((uint64_t)nT * 4), o_vtmp = stake((uint64_t)nT * 4);
    const uint64_t o_sort = stake(sort_elems * 4);
    const uint64_t o_flags = stake((uint64_t)nT * 4), o_esc = stake((uint64_t)nT * 4), o_box = stake(64), o_cnt = stake(64);
    if (s->scratch_bytes < so)
    {
        if (s->scratch) cudaFree(s->scratch);
        s->scratch = nullptr;
        if (cudaMalloc(&s->scratch, so) != cudaSuccess)
        {
            cudaGetLastError();
            set_error("cudaMalloc of the build scratch failed");
            return SNCH_ERR_OOM;
        }
        s->scratch_bytes = so;
    }
    unsigned char *sc = s->scratch;

    BuildCtx c;
    c.n = nT;
    c.n_edges = nE;
    c.verts = (const float3 *)(b + h.off_vertices);
    c.edges = (const RefEdge *)(b + h.off_edges);
    c.objects = (const RefTriangle *)(b + h.off_objects);
    c.nodes = (RefNode *)(b + h.off_nodes);
    c.aabbs = (RefAabb *)(b + h.off_aabbs);
    c.cones = (RefCone *)(b + h.off_cones);
    c.morton = (uint32_t *)(b + h.off_morton);
    c.sorted_idx = (uint32_t *)(b + h.off_sorted_idx);
    c.ranges = (uint2 *)(b + h.off_ranges);
    c.q1 = (uint8_t *)(b + h.off_q1);
    c.bnode = (BNode *)(b + h.off_bnode);
    c.snode = (SNode *)(b + h.off_snode);
    c.ltri = (LTri *)(b + h.off_ltri);
    c.ledge = (LEdge *)(b + h.off_ledge);
    c.edge_off = (uint32_t *)(b + h.off_edge_off);
    c.scene_box = (int *)(sc + o_box);
    c.flags = (uint32_t *)(sc + o_flags);
    c.escapes = (uint32_t *)(sc + o_esc);
    c.counters = (uint32_t *)(sc + o_cnt);

    struct EventPair
    { // destroyed on every exit path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair()
        {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } ev;
    SNCH_CUDA(cudaEventCreate(&ev.a));
    SNCH_CUDA(cudaEventCreate(&ev.b));
    SNCH_CUDA(cudaEventRecord(ev.a, stream));

    const unsigned g256 = (nT + 255) / 256;
    SNCH_CUDA(cudaMemsetAsync(c.flags, 0, (size_t)nT * 4, stream));
    SNCH_CUDA(cudaMemsetAsync(c.counters, 0, 64, stream));
    k_init_box<<<1, 32, 0, stream>>>(c.scene_box);
    k_scene_box<<<g256 < 1184 ? g256 : 1184, 256, 0, stream>>>(c);
    int launches = 2; // k_init_box, k_scene_box
    if (!refit)
    {
        k_morton<<<g256, 256, 0, stream>>>(c);
        launches += 1 + radix_sort_pairs(c.morton, c.sorted_idx, (uint32_t *)(sc + o_ktmp), (uint32_t *)(sc + o_vtmp), nT, 30,
                                         (uint32_t *)(sc + o_sort), stream);
        if (nT > 1) k_hierarchy<<<(nT - 1 + 255) / 256, 256, 0, stream>>>(c);
        launches += (nT > 1 ? 1 : 0);
    }
    if (s->opt_refit_kernel == 0) k_refit<<<(nT + 127) / 128, 128, 0, stream>>>(c);
    else
    {
        const unsigned ctas = (nT + kRefitLeaves - 1) / kRefitLeaves;
        k_refit_coop<<<ctas, kRefitLeaves, 0, stream>>>(c);
        if (nT > 1)
        {
            k_refit_top<<<ctas < 8 ? 1 : ctas / 8, 128, 0, stream>>>(c); // one thread per escape of the pass above (~12 per CTA); grid-stride beyond
            launches += 1;
        }
    }
    launches += 1;
    s->build_launches = (uint64_t)launches;
    SNCH_CUDA(cudaGetLastError());
    SNCH_CUDA(cudaEventRecord(ev.b, stream));

    // ---- stats read-back + header
    uint32_t counters[2] = {0, 0};
    int boxi[6];
    SNCH_CUDA(cudaMemcpyAsync(counters, c.counters, 8, cudaMemcpyDeviceToHost, stream));
    SNCH_CUDA(cudaMemcpyAsync(boxi, c.scene_box, 24, cudaMemcpyDeviceToHost, stream));
    SNCH_CUDA(cudaStreamSynchronize(stream));
    SNCH_CUDA(cudaEventElapsedTime(&s->build_ms, ev.a, ev.b));
    s->hdr.collision = refit ? prev_collision : counters[0];
    s->hdr.q1_nodes = counters[1];
    for (int a = 0; a < 3; ++a)
    {
        int lo = boxi[a], hi = boxi[3 + a];
        lo = lo >= 0 ? lo : lo ^ 0x7FFFFFFF;
        hi = hi >= 0 ? hi : hi ^ 0x7FFFFFFF;
        std::memcpy(&s->hdr.scene_lo[a], &lo, 4);
        std::memcpy(&s->hdr.scene_hi[a], &hi, 4);
    }
    SNCH_CUDA(cudaMemcpyAsync(b, &s->hdr, sizeof(ArenaHeader), cudaMemcpyHostToDevice, stream));
    SNCH_CUDA(cudaStreamSynchronize(stream));
    if (s->hdr.collision && s->opt_print_collision) std::printf("Morton code collision detected.\n");
    s->built = true;
    return SNCH_OK;
}

} // namespace snch
