// query.cu — batched traversal kernels (one launch per query batch).
//
// Each kernel returns what the reference's per-thread query_device()/sample_object_in_sphere() returns for the same
// query (query.cuh:79-169, 238-318, 325-423; sample.cuh:23-92).  What differs is how the batch is scheduled:
//
//   1. ORDER.   Large batches are visited in Morton order of the query points (k_query_bounds -> k_query_keys -> the
//      build's own radix sort): the 32 lanes of a warp then walk nearly the same root-to-leaf paths, so a node record
//      fetched by one lane is an L1 hit for its neighbours.  Results are written back to the caller's slots.
//   2. PERSISTENT LANES.  The grid is sized to the machine (SMs x resident CTAs), not to the batch.  A warp draws chunks
//      of consecutive queries from one global counter and every lane that finishes its query immediately takes the next
//      one of the chunk (ballot + popc rank, no per-lane atomics).  Query cost varies by >100x (a star radius below the
//      closest distance prunes at the root; a point in the hole of the torus opens thousands of nodes): with one query
//      per thread a warp ran at 3.2 of 32 lanes (profiles/r01a_*); refilling keeps the lanes occupied.
//   3. TRAVERSAL.  Two-child records (layout.h), nearer child first, the running best applied when a child is pushed
//      and again when it is popped.  The closest-point kernel seeds its bound with the triangle that answered the lane's
//      previous (neighbouring) query.  The silhouette kernel decides the reference's normal-cone test (cone.cuh:168-212)
//      in sine space with a guard band and only evaluates the acosf/asinf/atan2f chain inside the band, so prune
//      decisions are the reference's while ~99.9% of tests cost a fifth of it.
//
// Result equivalence (DESIGN.md "Parity rules"):
//   closest   : min over all triangles of the reference's own point-triangle distance; index = any argmin (ties, Q3)
//   silhouette: min over the leaves that pass the reference's cone test chain (same predicate)
//   ray       : smallest t with t < max_dist; prim = any triangle attaining it (Q4)
//   sample    : identical single-path descent (deterministic given u)
#include "query_common.cuh"
#include "scene.h"
#include "snch_math.cuh"
#include "sort_scan.cuh"

namespace snch
{

// ---------------------------------------------------------------------------------------------------------------
// query ordering
// ---------------------------------------------------------------------------------------------------------------
SNCH_DI int f2ord_q(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
SNCH_DI float ord2f_q(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__global__ void k_query_box_init(int *box)
{
    if (threadIdx.x < 3) box[threadIdx.x] = f2ord_q(INFINITY);
    else if (threadIdx.x < 6) box[threadIdx.x] = f2ord_q(-INFINITY);
}
// bounding box of the finite query points (stride = floats per query: 3 for points/origins, 4 for spheres)
__global__ void __launch_bounds__(256) k_query_bounds(const float *__restrict__ q, int stride, int dims, uint32_t n, int *box)
{
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            if (a >= dims) continue; // 2-D batches: the third axis keeps an empty range and contributes no key bits
            const float x = __ldg(q + (uint64_t)stride * i + a);
            if (fabsf(x) <= FLT_MAX)
            {
                lo[a] = fminf(lo[a], x);
                hi[a] = fmaxf(hi[a], x);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFull, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFull, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0)
    {
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            atomicMin(box + a, f2ord_q(lo[a]));
            atomicMax(box + 3 + a, f2ord_q(hi[a]));
        }
    }
}
// 30-bit Morton key of each query point inside the batch's own bounding box (a scheduling hint only: any key is correct)
// With `radius` (silhouette star radii) the top 4 key bits are the radius octave relative to the batch extent, so the 32
// queries a warp walks together also have similar search radii (a scheduling hint only: any key is correct).
__global__ void __launch_bounds__(256) k_query_keys(const float *__restrict__ q, int stride, int dims, uint32_t n, const int *__restrict__ box,
                                                    const float *__restrict__ radius, int radius_desc, const float *__restrict__ dirs,
                                                    uint32_t *__restrict__ keys, uint32_t *__restrict__ perm)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        if (a >= dims) continue;
        const float lo = ord2f_q(box[a]), hi = ord2f_q(box[3 + a]);
        const float x = __ldg(q + (uint64_t)stride * i + a);
        float t = (x - lo) / fmaxf(hi - lo, FLT_MIN) * 1024.0f;
        t = fminf(fmaxf(t, 0.0f), 1023.0f); // NaN -> 0
        code |= expand_bits10((uint32_t)t) << (2 - a);
    }
    if (radius)
    {
        const float ext = fmaxf(fmaxf(ord2f_q(box[3]) - ord2f_q(box[0]), ord2f_q(box[4]) - ord2f_q(box[1])),
                                fmaxf(dims > 2 ? ord2f_q(box[5]) - ord2f_q(box[2]) : 0.0f, FLT_MIN));
        const float r = __ldg(radius + i);
        // radius class = floor(2^e * log2(r / extent)) read off the float bit patterns (exponent + top e mantissa bits)
        const int e = radius_desc > 1 ? radius_desc - 1 : 0; // extra class bits beyond the octave
        const int oct = (int)((__float_as_uint(fmaxf(r, 0.0f)) >> (23 - e)) & (0xFFu << e | ((1u << e) - 1u))) -
                        (int)((__float_as_uint(ext) >> (23 - e)) & (0xFFu << e | ((1u << e) - 1u)));
        const int top = (16 << e) - 1;
        uint32_t cls = (uint32_t)min(max(oct + (13 << e), 0), top); // >= 4 x extent (incl. +inf) -> top; NaN -> 0
        if (radius_desc) cls = (uint32_t)top - cls; // largest search radii (the most expensive walks) first: the kernel's tail is then made of cheap queries
        code = (cls << (26 - e)) | (code >> (4 + e));
    }
    if (dirs)
    { // rays: direction octant first (a warp's rays then descend the same side of every split), origin Morton code below it
        const uint32_t oct = (__ldg(dirs + 3 * (uint64_t)i) < 0.0f ? 4u : 0u) | (__ldg(dirs + 3 * (uint64_t)i + 1) < 0.0f ? 2u : 0u) |
                             (__ldg(dirs + 3 * (uint64_t)i + 2) < 0.0f ? 1u : 0u);
        code = (oct << 27) | (code >> 3);
    }
    keys[i] = code;
    perm[i] = i;
}

// ---------------------------------------------------------------------------------------------------------------
// nearest primitive                                                                      query.cuh:238-318
// warp-cooperative ("packet") traversal for Morton-ordered batches
// ---------------------------------------------------------------------------------------------------------------
// The 32 lanes of a warp hold 32 consecutive queries of the ordered batch and walk the tree TOGETHER: one node record is
// fetched per warp (a single broadcast sector per load instead of 32 divergent ones), every lane tests its own query
// against it, and __ballot_sync decides which children any lane still needs.  The traversal stack is per warp (node +
// lane mask, in shared memory); per-lane state is just the running best.  Lanes only ever skip work they could not use
// (mindist >= their best / their cone test failed), so each lane's result is exactly what its own traversal returns.
SNCH_DI float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
SNCH_DI float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Can the triangle with first vertex pa, unit normal n and reach ra (no point of it is farther than ra from pa) be nearer to x
// than `bestd` = sqrt(best2)?  Lower bound of its distance: the distance to its plane, combined with how far the projection
// of x lies beyond that reach.  For the triangles an exact search has to LOOK at — those whose box comes within the current
// best, a band whose population grows with sqrt(N) for a query far from a fine surface — this is tight to second order where
// the box is tight to first order, so most of them are rejected for ~25 instructions instead of the ~120 of the reference's
// point-triangle distance.
// Sound in float, so results stay bit-identical: (i) the bound's own roundings (|n| = 1 +- 2e-7, the dot products, the
// approximate square roots) are covered by the 4e-6 s / 2e-6 s terms, s = |x - pa|^2; (ii) the reference's distance function
// rounds its closest point in ABSOLUTE terms (~4e-7 of the coordinate magnitude M <= |x|_inf + |x - pa|), so the distance it
// would return may undercut the true one by that much: the triangle is rejected only if the bound exceeds
// (bestd + 2e-6 M)^2 (1 + 4e-6), i.e. only if that returned distance could not have passed `dist^2 < best2`.  NaN (degenerate
// triangle, infinite best) never rejects.
SNCH_DI bool tri_cannot_improve(V3 pa, V3 n, float ra, V3 x, float xmax, float bestd)
{
    const V3 w = x - pa;
    const float s = dot(w, w), pd = dot(n, w), pd2 = pd * pd;
    const float lat = fmaxf(sqrt_approx(fmaxf(s - pd2 - 4e-6f * s, 0.0f)) - ra, 0.0f);
    const float lb2 = pd2 + lat * lat - 2e-6f * s;
    const float t = bestd + 2e-6f * (xmax + sqrt_approx(s));
    return lb2 > t * t * 1.000004f;
}

// COOPERATIVE WALK of one query by a whole warp.  A query that is (nearly) equidistant to a large part of the mesh (the centre
// of a sphere, the axis of a torus) needs tens of thousands of nodes; walked by one lane, or by a packet it shares with 31
// neighbours that need the same nodes, that is a 5 ms critical path — the whole run time of a batch too small to fill the
// machine (config C1: 64K queries), whatever the size of the GPU.  Here the warp keeps a shared stack of (node, key) entries, pops
// up to 32 per step, every lane opens one node — both child boxes, leaf children tested at once — the survivors are pushed with
// warp-aggregated offsets and the running best is the warp minimum.  The result is the minimum over the same triangles with
// the same distance function, so the distance is the one the sequential walk returns (the index is an argmin, as everywhere:
// ties, Q3).  Tried inside the packet kernel too (for subtrees only one lane still needs, and as a second pass for packets
// over a step budget): 51 -> 59 ms on the 16.7M-query batch (registers 48 -> 56, one resident CTA fewer), so large batches keep
// the plain packets.
constexpr int kSoloStack = 512;  // entries per warp; above kSoloStack - 128 the walk pops one entry per step (growth <= tree depth)
SNCH_DI void solo_closest(const SceneView &sv, StackEntry *st, uint32_t root, int owner, int lane, V3 p_lane, float &best2_lane, uint32_t &best_lane,
                          uint32_t &best_leaf_lane, bool use_lb)
{
    const V3 p = V3{__shfl_sync(kFull, p_lane.x, owner), __shfl_sync(kFull, p_lane.y, owner), __shfl_sync(kFull, p_lane.z, owner)};
    float b2 = __shfl_sync(kFull, best2_lane, owner);
    uint32_t bi = __shfl_sync(kFull, best_lane, owner), bl = __shfl_sync(kFull, best_leaf_lane, owner);
    const unsigned lt = (1u << lane) - 1u;
    const float pmax = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z));
    if (lane == 0) st[0] = StackEntry{root, 0.0f};
    int sp = 1;
    __syncwarp();
    while (sp > 0)
    {
        const int take = sp > kSoloStack - 128 ? 1 : (sp < 32 ? sp : 32);
        sp -= take;
        StackEntry e = StackEntry{kNone, INFINITY};
        if (lane < take) e = st[sp + lane];
        __syncwarp(); // every popped entry is read before this step's pushes reuse the slots
        uint32_t cand = 0xFFFFFFFFu, cand_i = kNone, cand_leaf = kNone; // squared distance as ordered bits (>= +0)
        const float bd = sqrt_approx(b2) * 1.000001f;
        uint32_t pr0 = kNone, pr1 = kNone;
        float pk0 = 0.0f, pk1 = 0.0f;
        if (e.node != kNone && e.key < b2)
        {
            float4 a, b, c, d;
            ld256(sv.bnode + e.node, a, b);
            ld256(reinterpret_cast<const char *>(sv.bnode + e.node) + 32, c, d);
            const NodeBoxes nb = unpack_boxes(a, b, c);
            const float m0 = box_mindist2(nb.lo0, nb.hi0, p), m1 = box_mindist2(nb.lo1, nb.hi1, p);
            const uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y);
            const bool far1 = m0 < m1; // the farther child is pushed first, so the nearer one is popped first
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            {
                const bool one = (ch == 0) == far1; // iteration 0: the farther child, iteration 1: the nearer child
                const float m = one ? m1 : m0;
                const uint32_t r = one ? r1 : r0;
                if (!(m < b2)) continue;
                if (r & kLeafFlag)
                {
                    const uint32_t k = r & ~kLeafFlag;
                    const LTri *tp = sv.ltri + k;
                    float4 t0, t1, t2, t3;
                    ld256(tp, t0, t1);
                    ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
                    if (use_lb && tri_cannot_improve(V3{t0.x, t0.y, t0.z}, V3{t3.x, t3.y, t3.z}, t3.w, p, pmax, bd)) continue;
                    float dist = point_triangle_distance(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, p);
                    dist *= dist;
                    if (dist < b2 && __float_as_uint(dist) < cand)
                    {
                        cand = __float_as_uint(dist);
                        cand_i = __float_as_uint(t0.w);
                        cand_leaf = k;
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
        { // some lane improved the bound (cand < b2 by construction)
            const int wl = __ffs(__ballot_sync(kFull, cand == mn)) - 1;
            b2 = __uint_as_float(mn);
            bi = __shfl_sync(kFull, cand_i, wl);
            bl = __shfl_sync(kFull, cand_leaf, wl);
        }
        __syncwarp();
    }
    if (lane == owner)
    {
        best2_lane = b2;
        best_lane = bi;
        best_leaf_lane = bl;
    }
}

struct PacketStack
{
    uint2 e[kStackDepth];
};

__global__ void __launch_bounds__(kQueryThreads)
    k_closest_packet(SceneView sv, const float *__restrict__ q, const uint32_t *__restrict__ perm, uint32_t n, uint32_t *__restrict__ out_idx,
                     float *__restrict__ out_dist, unsigned long long *counter, int use_seed)
{
    __shared__ PacketStack s_stack[kQueryThreads / 32];
    const int lane = threadIdx.x & 31;
    uint2 *stk = s_stack[threadIdx.x >> 5].e;
    uint32_t best_leaf = kNone;
    const bool use_lb = !(use_seed & 2); // "query.seed" bit 1 switches the cheap lower bound off (A/B, knob tests)
    use_seed &= 1;
    for (;;)
    {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, 32ull);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        const uint32_t s = (uint32_t)base + lane;
        const bool valid = s < n;
        const uint32_t slot = valid ? (perm ? __ldg(perm + s) : s) : 0u;
        const V3 p = valid ? load_point(q, slot) : V3{0.f, 0.f, 0.f};
        const float pmax = fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z));
        float best2 = INFINITY;
        uint32_t best = kNone;
        if (use_seed && valid && best_leaf != kNone)
        { // the triangle that answered this lane's previous (neighbouring) query bounds this one
            const LTri *tp = sv.ltri + best_leaf;
            float4 t0, t1, t2, t3;
            ld256(tp, t0, t1);
            ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
            float dist = point_triangle_distance(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, p);
            dist *= dist;
            if (dist < INFINITY)
            {
                best2 = dist;
                best = __float_as_uint(t0.w);
            }
            else best_leaf = kNone;
        }
        unsigned mask = __ballot_sync(kFull, valid);
        uint32_t node = 0;
        int sp = 0;
        for (;;)
        {
            float4 a, b, c, d;
            ld256(sv.bnode + node, a, b);
            ld256(reinterpret_cast<const char *>(sv.bnode + node) + 32, c, d);
            const NodeBoxes nb = unpack_boxes(a, b, c);
            const float m0 = box_mindist2(nb.lo0, nb.hi0, p), m1 = box_mindist2(nb.lo1, nb.hi1, p);
            const uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y); // warp-uniform
            const bool in = (mask >> lane) & 1u;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            { // leaf children are tested at once: they tighten the bounds before the sibling subtree is considered
                const uint32_t r = ch ? r1 : r0;
                if (!(r & kLeafFlag)) continue;
                const bool w = in && ((ch ? m1 : m0) < best2);
                if (!__any_sync(kFull, w)) continue;
                const uint32_t k = r & ~kLeafFlag;
                const LTri *tp = sv.ltri + k;
                float4 t0, t1, t2, t3;
                ld256(tp, t0, t1);
                ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
                if (w && !(use_lb && tri_cannot_improve(V3{t0.x, t0.y, t0.z}, V3{t3.x, t3.y, t3.z}, t3.w, p, pmax, sqrt_approx(best2) * 1.000001f)))
                { // only the lanes that need this triangle: the Voronoi-region branches of the test then split the warp over the
                  // regions of THOSE lanes, not over the regions of all 32 query points
                    float dist = point_triangle_distance(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, p);
                    dist *= dist; // the reference squares the distance it got back (query.cuh:284-285)
                    if (dist < best2)
                    {
                        best2 = dist;
                        best = __float_as_uint(t0.w);
                        best_leaf = k;
                    }
                }
            }
            const bool w0 = in && !(r0 & kLeafFlag) && m0 < best2;
            const bool w1 = in && !(r1 & kLeafFlag) && m1 < best2;
            const unsigned b0 = __ballot_sync(kFull, w0), b1 = __ballot_sync(kFull, w1);
            if (b0 && b1)
            { // both needed: enter the child most lanes are nearer to, keep the other with the lanes that want it
                const unsigned near1 = __ballot_sync(kFull, (w0 && w1) ? (m1 < m0) : w1);
                const bool first1 = 2 * __popc(near1) > __popc(b0 | b1);
                if (lane == 0) stk[sp] = first1 ? make_uint2(r0, b0) : make_uint2(r1, b1); // (warp-uniform value: one writer)
                __syncwarp();
                ++sp;
                node = first1 ? r1 : r0;
                mask = first1 ? b1 : b0;
            }
            else if (b0 | b1)
            {
                node = b0 ? r0 : r1;
                mask = b0 | b1;
            }
            else
            {
                if (sp == 0) break;
                --sp;
                const uint2 e = stk[sp];
                __syncwarp(); // every lane has read the entry before a later push may overwrite the slot
                node = e.x;
                mask = e.y;
            }
        }
        if (valid)
        {
            out_idx[slot] = best;
            out_dist[slot] = sqrtf(best2);
        }
    }
}

// One query per WARP, walked with solo_closest from the root.  For batches too small to fill the machine with packets the
// run time is the critical path of the most expensive query (config C1, 64K queries on a sphere: 5.2 ms for the points near
// the centre, to which every triangle is equally near); 32 lanes on one query shorten exactly that path.
__global__ void __launch_bounds__(kQueryThreads)
    k_closest_wide(SceneView sv, const float *__restrict__ q, const uint32_t *__restrict__ perm, uint32_t n, uint32_t *__restrict__ out_idx,
                   float *__restrict__ out_dist, unsigned long long *counter, int use_seed)
{
    __shared__ StackEntry s_solo[kQueryThreads / 32][kSoloStack];
    const int lane = threadIdx.x & 31;
    StackEntry *solo = s_solo[threadIdx.x >> 5];
    uint32_t best_leaf = kNone;
    const bool use_lb = !(use_seed & 2); // "query.seed" bit 1 switches the cheap lower bound off (A/B, knob tests)
    use_seed &= 1;
    constexpr unsigned long long kRun = 4; // consecutive (neighbouring) queries per draw
    for (;;)
    {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, kRun);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        for (unsigned long long s = base; s < base + kRun && s < n; ++s)
        {
            const uint32_t slot = perm ? __ldg(perm + s) : (uint32_t)s;
            const V3 p = load_point(q, slot); // every lane holds the query
            float best2 = INFINITY;
            uint32_t best = kNone;
            if (use_seed && best_leaf != kNone)
            {
                const LTri *tp = sv.ltri + best_leaf;
                float4 t0, t1, t2, t3;
                ld256(tp, t0, t1);
                ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
                float dist = point_triangle_distance(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, p);
                dist *= dist;
                if (dist < INFINITY)
                {
                    best2 = dist;
                    best = __float_as_uint(t0.w);
                }
                else best_leaf = kNone;
            }
            solo_closest(sv, solo, 0u, 0, lane, p, best2, best, best_leaf, use_lb);
            best_leaf = __shfl_sync(kFull, best_leaf, 0);
            if (lane == 0)
            {
                out_idx[slot] = best;
                out_dist[slot] = sqrtf(best2);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// nearest silhouette                                                                     query.cuh:325-423
// ---------------------------------------------------------------------------------------------------------------
// The reference's view-cone test (cone.cuh:168-212) is, in real arithmetic,  |acos(t) - pi/2| <= alpha + beta  (or
// alpha + beta >= pi/2), with t = axis . dir(o -> box centre), alpha the cone half-angle and beta the half-angle the box
// subtends (asin(radius/l) outside the cone's sphere, atan2(projected extent, l - d) inside).  Equivalently
// |t| <= sin(alpha + beta).  The filter evaluates that in sine space with fast intrinsics and answers only when the
// inequality holds or fails by more than kConeBand (>= 10x the rounding of either evaluation); inside the band — and for
// every branch decision the reference takes on exact float values — it defers to cone_overlap(), the reference's own
// operation sequence.  Decisions are therefore the reference's; only their cost changes.
constexpr float kConeBand = 2e-5f;
// kMode 0: the reference's libm chain, verbatim ("query.cone_filter" = 0: the A/B the parity tests sweep).
// kMode 1: the sine-space filter on the MUFU approximations (rel. error <= 2^-22, i.e. <= 5e-7 on every quantity compared
//    against the 2e-5 band); the two exact-value branch decisions of the reference (l > radius, s <= 0) and the
//    ill-conditioned corner (view cone within ~6 degrees of a half space, where cos(beta) amplifies the error of sin(beta))
//    are handed to the exact chain, which is kept OUT OF LINE (one copy per kernel instead of one per hand-off site: the
//    walk's hot loop shrinks by ~2000 instructions, profiles/r01j showed 0.74 "no instruction" stall cycles per issue).
__device__ __noinline__ bool cone_overlap_ool(V3 axis, float half_angle, float radius, V3 o, V3 lo, V3 hi, float md2)
{
    return cone_overlap(axis, half_angle, radius, o, lo, hi, md2);
}
template <int kMode> SNCH_DI bool cone_test(V3 axis, float half_angle, float radius, V3 o, V3 lo, V3 hi, float md2)
{
    if (kMode == 0) return cone_overlap(axis, half_angle, radius, o, lo, hi, md2);
    if (half_angle >= kHalfPi || md2 < FLT_EPSILON) return true;
    const V3 c = V3{(hi.x + lo.x) * 0.5f, (hi.y + lo.y) * 0.5f, (hi.z + lo.z) * 0.5f};
    const V3 w = c - o;
    const float l2 = w.x * w.x + w.y * w.y + w.z * w.z;
    const float rl = rsqrt_approx(l2);
    const float l = l2 * rl;
    if (!(fabsf(l - radius) > 4e-6f * radius)) return cone_overlap_ool(axis, half_angle, radius, o, lo, hi, md2); // also NaN / l2 == 0
    const float t = fabsf(__fmaf_rn(axis.x, w.x, __fmaf_rn(axis.y, w.y, axis.z * w.z))) * rl;
    float sa, ca;
    __sincosf(half_angle, &sa, &ca);
    float sb, cb;
    if (l > radius)
    {
        sb = radius * rl;
        const float cb2 = fmaxf(__fmaf_rn(-sb, sb, 1.0f), 0.0f);
        if (cb2 < 0.01f) return cone_overlap_ool(axis, half_angle, radius, o, lo, hi, md2);
        cb = sqrt_approx(cb2);
    }
    else
    {
        const V3 v = V3{w.x * rl, w.y * rl, w.z * rl};
        const V3 e = hi - c;
        const float d = __fmaf_rn(e.x, fabsf(v.x), __fmaf_rn(e.y, fabsf(v.y), e.z * fabsf(v.z)));
        const float s = l - d;
        const float sband = kConeBand * l;
        if (s < -sband) return true; // the reference returns true for s <= 0
        if (!(s > sband)) return cone_overlap_ool(axis, half_angle, radius, o, lo, hi, md2);
        // project_to_plane(v, e)                                                        cone.cuh:34-42, 58-66
        const float sign = copysignf(1.0f, v.z);
        const float ia = -__frcp_rn(sign + v.z);
        const float bb = v.x * v.y * ia;
        const float b1x = __fmaf_rn(sign * v.x * v.x, ia, 1.0f), b1y = sign * bb, b1z = -sign * v.x;
        const float b2x = bb, b2y = __fmaf_rn(v.y * v.y, ia, sign), b2z = -v.y;
        const float r1 = __fmaf_rn(e.x, fabsf(b1x), __fmaf_rn(e.y, fabsf(b1y), e.z * fabsf(b1z)));
        const float r2 = __fmaf_rn(e.x, fabsf(b2x), __fmaf_rn(e.y, fabsf(b2y), e.z * fabsf(b2z)));
        const float pr2 = __fmaf_rn(r1, r1, r2 * r2);
        const float rh = rsqrtf(__fmaf_rn(s, s, pr2));
        sb = sqrtf(pr2) * rh;
        cb = s * rh;
    }
    const float cg = __fmaf_rn(ca, cb, -sa * sb); // cos(alpha + beta)
    const float sg = __fmaf_rn(sa, cb, ca * sb);  // sin(alpha + beta)
    if (cg <= -kConeBand) return true;            // alpha + beta > pi/2
    if (cg >= kConeBand)
    {
        if (t <= sg - kConeBand) return true;
        if (t >= sg + kConeBand) return false;
    }
    return cone_overlap_ool(axis, half_angle, radius, o, lo, hi, md2); // inside the band (or NaN): the reference's own sequence
}

// silhouette_distance_calculator over the owned edges of one leaf (scene.cuh:978-1003 -> silhouette_edge::
// find_closest_silhouette_point, :788-824), `bound` = the running best (inclusive, as there).  Returns true when an edge
// improved it; `bound` / `slot` (LEdge index of that edge) are updated.
SNCH_DI bool leaf_silhouette(const SceneView &sv, uint32_t first, uint32_t cnt, V3 p, bool flip, float &bound, uint32_t &slot)
{
    float b2 = bound * bound;
    bool hit = false;
    for (uint32_t k = 0; k < cnt; ++k)
    {
        float4 e0, e1, e2, e3;
        ld256(sv.ledge + first + k, e0, e1);
        ld256(reinterpret_cast<const char *>(sv.ledge + first + k) + 32, e2, e3);
        const V3 pa = V3{e0.x, e0.y, e0.z}, pb = V3{e0.w, e1.x, e1.y};
        V3 cp;
        const float dist = point_segment_distance(pa, pb, p, &cp);
        // scene.cuh:791 `if (min_radius_squared >= max_radius_squared) return false` with min = 0: a bound whose square is zero (a
        // star radius of 0: the query point lies on the surface) finds nothing, not even an edge at distance 0
        if (0.0f >= b2 || dist * dist > b2) continue;
        bool is_sil = __float_as_uint(e3.y) != 0u; // boundary edge
        if (!is_sil) is_sil = is_silhouette_edge(pa, pb, V3{e1.z, e1.w, e2.x}, V3{e2.y, e2.z, e2.w}, p - cp, dist, flip);
        if (is_sil && dist <= bound)
        {
            bound = dist;
            b2 = dist * dist;
            slot = first + k;
            hit = true;
        }
    }
    return hit;
}
// The optional outputs of a silhouette query (the reference computes the point at scene.cuh:796-799 and
// drops it, and leaves the index as a TODO at query.cuh:386,411): the id of the silhouette edge that attains the distance
// (its index in scene<3>::silhouettes) and the closest point on it — recomputed from the edge's record with the very
// call that produced the distance, so |p - point| is that distance bit for bit.
SNCH_DI void write_silhouette_point(const SceneView &sv, uint32_t slot, uint32_t ledge_slot, V3 p, bool found, uint32_t *__restrict__ out_edge,
                                    float *__restrict__ out_point)
{
    uint32_t id = kNone;
    V3 cp = V3{0.f, 0.f, 0.f};
    if (found)
    {
        float4 e0, e1, e2, e3;
        ld256(sv.ledge + ledge_slot, e0, e1);
        ld256(reinterpret_cast<const char *>(sv.ledge + ledge_slot) + 32, e2, e3);
        point_segment_distance(V3{e0.x, e0.y, e0.z}, V3{e0.w, e1.x, e1.y}, p, &cp);
        id = __float_as_uint(e3.x);
    }
    if (out_edge) out_edge[slot] = id;
    if (out_point)
    {
        out_point[3 * (uint64_t)slot] = cp.x;
        out_point[3 * (uint64_t)slot + 1] = cp.y;
        out_point[3 * (uint64_t)slot + 2] = cp.z;
    }
}

// COOPERATIVE WALK of one silhouette query by a whole warp: every lane opens one node (both child boxes, the reference's
// per-child cone test, the edges of leaf children), survivors are pushed with warp-aggregated offsets, the bound is the
// warp minimum.  The answer is the minimum over the silhouette edges of the leaves the reference's predicate chain
// reaches, which does not depend on the order of the walk.  `Stk` is the warp's (node, key) stack: a flat array in the
// one-query-per-warp kernel, the warp's slice of the per-lane stack levels when the per-lane kernel finishes its last
// walks this way.  kCap = its capacity; above kCap - 128 the walk pops one entry per step (growth <= tree depth).
struct FlatStack
{
    StackEntry *e;
    SNCH_DI StackEntry get(int i) const { return e[i]; }
    SNCH_DI void put(int i, StackEntry v) const { e[i] = v; }
};
template <int kFilter, bool kEdge, int kCap, typename Stk>
SNCH_DI float solo_silhouette(const SceneView &sv, const Stk st, int lane, V3 p, bool flip, float best, uint32_t &best_slot)
{
    float best2 = best * best;
    bool found = false;
    const unsigned lt = (1u << lane) - 1u;
    if (lane == 0) st.put(0, StackEntry{0u, 0.0f});
    int sp = 1;
    __syncwarp();
    while (sp > 0)
    {
        const int take = sp > kCap - 128 ? 1 : (sp < 32 ? sp : 32);
        sp -= take;
        StackEntry en = StackEntry{kNone, INFINITY};
        if (lane < take) en = st.get(sp + lane);
        __syncwarp();
        uint32_t cand = 0xFFFFFFFFu, cand_slot = kNone; // distance as ordered bits (>= +0)
        uint32_t pr0 = kNone, pr1 = kNone;
        float pk0 = 0.0f, pk1 = 0.0f;
        if (en.node != kNone && en.key <= best2)
        {
            float4 a, b, c, d, e, f;
            ld256(sv.snode + en.node, a, b);
            ld256(reinterpret_cast<const char *>(sv.snode + en.node) + 32, c, d);
            ld256(reinterpret_cast<const char *>(sv.snode + en.node) + 64, e, f);
            const NodeBoxes nb = unpack_boxes(a, b, c);
            const float m0 = box_mindist2(nb.lo0, nb.hi0, p), m1 = box_mindist2(nb.lo1, nb.hi1, p);
            const uint32_t r0 = __float_as_uint(f.z), r1 = __float_as_uint(f.w);
            const bool h0 = (m0 <= best2) && (d.w >= 0.0f) && cone_test<kFilter>(V3{d.x, d.y, d.z}, d.w, e.x, p, nb.lo0, nb.hi0, m0);
            const bool h1 = (m1 <= best2) && (f.x >= 0.0f) && cone_test<kFilter>(V3{e.y, e.z, e.w}, f.x, f.y, p, nb.lo1, nb.hi1, m1);
            const bool far1 = m0 < m1; // the farther child is pushed first, so the nearer one is popped first
            float ob = best;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            {
                const bool one = (ch == 0) == far1;
                if (!(one ? h1 : h0)) continue;
                const float m = one ? m1 : m0;
                const uint32_t r = one ? r1 : r0;
                if (r & kLeafFlag)
                {
                    const uint32_t payload = r & ~kLeafFlag;
                    if (leaf_silhouette(sv, payload >> 2, payload & 3u, p, flip, ob, cand_slot)) cand = __float_as_uint(ob);
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
        if (pr0 != kNone) st.put(off, StackEntry{pr0, pk0});
        if (pr1 != kNone) st.put(off + 1, StackEntry{pr1, pk1});
        sp += __popc(c1) + __popc(c2);
        const uint32_t mn = __reduce_min_sync(kFull, cand);
        if (mn != 0xFFFFFFFFu)
        { // cand <= best by construction
            best = __uint_as_float(mn);
            best2 = best * best;
            found = true;
            if (kEdge) best_slot = __shfl_sync(kFull, cand_slot, __ffs(__ballot_sync(kFull, cand == mn)) - 1);
        }
        __syncwarp();
    }
    return found ? best : INFINITY;
}

// Per-lane traversal with a WARP-SHARED leaf queue.  A leaf costs up to three edge tests (~100 instructions each) and only
// one lane in ~15 reaches one in a given step: tested inline, that code ran at 2 of 32 lanes and was 42% of all issued
// instructions (profiles/r01b_*); parked per lane and drained by each lane for itself, at 3-4 of 32 (profiles/r01k).  Here
//   * a lane that reaches a leaf appends (first edge, edge count | owner lane) to a queue in shared memory; when the
//     queue holds a warp's worth — or a lane has finished walking and needs its answer — the warp tests the queued
//     leaves one per lane, reading the owner's query through shuffles, and hands results back through a shared
//     per-lane minimum.  The answer is min over the silhouette edges within the bound, which does not depend on the
//     order or grouping of the tests, so results are identical to the sequential loop's;
//   * the lowest kSStack levels of each lane's traversal stack live in shared memory (conflict-free: the bank depends
//     on the lane only), deeper levels spill to local memory;
//   * kEdge: the per-lane minimum is the 64-bit key (distance bits, LEdge slot), so the edge that attains the answer comes
//     back with it (snch_closest_silhouette_batch out_edge / out_point); instantiated only when those outputs are asked for;
//   * TAIL: once the batch has no more queries to hand out, a warp left with at most `tail_lanes` walking lanes stops
//     walking them one lane each: it drains its queue and puts those queries, with the bound each has found so far, on a
//     list that the next launch (k_silhouette_wide) finishes ONE QUERY PER WARP from the root.  The few queries that open
//     thousands of nodes (a point in the hole of the torus) were a fixed ~8 ms single-lane tail of every unbounded batch;
//     same predicate chain, same edge tests, same minimum.  The edge found before the restart travels with the entry
//     (tail_prev): its distance is the restart's inclusive bound, but far from the origin the ROUNDED distance to an edge
//     can be smaller than the rounded distance to the box that holds it, so the restart is not certain to reach it again
//     (tests/test_gpu_fuzz.py seed 7: coordinates near 1000); when the restart finds nothing within the bound, that edge
//     is the answer.
constexpr int kSStack = 12;
constexpr int kLeafQueue = 96;   // >= kLeafFlushAt - 1 + 64 (every lane can add two leaves per step)
constexpr int kLeafFlushAt = 32; // upper limit of "query.sil_flush"
template <bool kEdge> struct SilResult
{
    using T = uint32_t;
    static constexpr T kEmpty = 0xFFFFFFFFu;
    SNCH_DI static T make(float d, uint32_t) { return __float_as_uint(d); } // distances are >= +0: uint order = float order
    SNCH_DI static float dist(T r) { return __uint_as_float(r); }
    SNCH_DI static uint32_t slot(T) { return kNone; }
};
template <> struct SilResult<true>
{
    using T = unsigned long long;
    static constexpr T kEmpty = ~0ull;
    SNCH_DI static T make(float d, uint32_t s) { return ((T)__float_as_uint(d) << 32) | s; }
    SNCH_DI static float dist(T r) { return __uint_as_float((uint32_t)(r >> 32)); }
    SNCH_DI static uint32_t slot(T r) { return (uint32_t)r; }
};
template <int kFilter, bool kEdge>
__global__ void __launch_bounds__(kQueryThreads, 8)
    k_silhouette_coop(SceneView sv, const float *__restrict__ q, const uint8_t *__restrict__ flipv, const float *__restrict__ rmax,
                      const uint32_t *__restrict__ perm, uint32_t n, float *__restrict__ out_dist, uint32_t *__restrict__ out_edge,
                      float *__restrict__ out_point, unsigned long long *counter, int tail_lanes, uint32_t *__restrict__ tail_slot,
                      float *__restrict__ tail_bound, uint32_t *__restrict__ tail_prev, uint32_t flush_at, uint32_t chunk)
{
    using Res = SilResult<kEdge>;
    __shared__ StackEntry s_stk[kSStack][kQueryThreads];
    struct WarpQueue // one base address per warp: payloads, owner bytes and the fill count are immediate offsets from it
    {
        uint32_t payload[kLeafQueue];
        uint8_t owner[kLeafQueue];
        uint32_t count;
    };
    __shared__ WarpQueue s_wq[kQueryThreads / 32];
    __shared__ typename Res::T s_result[kQueryThreads];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WarpQueue &wq = s_wq[wid];
    Feeder fd{0u, 0u, false, chunk};
    StackEntry lstk[kStackDepth - kSStack];
    int sp = 0;
    V3 p = V3{0.f, 0.f, 0.f};
    bool flip = false, found = false, busy = false, pend = false, tail = false;
    float best = INFINITY, best2 = INFINITY;
    uint32_t best_slot = kNone; // kEdge: LEdge slot of the edge that attains `best`
    uint32_t slot = kNone, node = kNone;
    s_result[threadIdx.x] = Res::kEmpty;
    if (lane == 0) wq.count = 0;
    __syncwarp();
    for (;;)
    {
        // ---- 1. queued leaves
        uint32_t qc = wq.count;
        if (qc >= flush_at || tail || __any_sync(kFull, pend && node == kNone))
        {
            __syncwarp(); // every lane has read the count before lane 0 resets it below
            while (qc > 0)
            {
                const uint32_t take = qc < 32u ? qc : 32u;
                qc -= take;
                const bool mine = (uint32_t)lane < take;
                const uint32_t payload = mine ? wq.payload[qc + lane] : 0u;
                const int owner = mine ? (int)wq.owner[qc + lane] : lane;
                const V3 op = V3{__shfl_sync(kFull, p.x, owner), __shfl_sync(kFull, p.y, owner), __shfl_sync(kFull, p.z, owner)};
                float ob = __shfl_sync(kFull, best, owner);
                const bool oflip = __shfl_sync(kFull, (int)flip, owner) != 0;
                uint32_t eslot = kNone;
                if (leaf_silhouette(sv, payload >> 2, payload & 3u, op, oflip, ob, eslot)) atomicMin(&s_result[(wid << 5) + owner], Res::make(ob, eslot));
                __syncwarp();
            }
            const typename Res::T rb = s_result[threadIdx.x];
            if (rb != Res::kEmpty)
            {
                const float v = Res::dist(rb);
                if (v <= best)
                {
                    best = v;
                    best2 = v * v;
                    found = true;
                    if (kEdge) best_slot = Res::slot(rb);
                }
                s_result[threadIdx.x] = Res::kEmpty;
            }
            pend = false;
            if (lane == 0) wq.count = 0;
            __syncwarp();
            if (tail) break;
        }
        // ---- 2. finished walks hand in their answer; idle lanes take the next query
        if (busy && node == kNone)
        {
            out_dist[slot] = found ? best : INFINITY;
            if (kEdge) write_silhouette_point(sv, slot, best_slot, p, found, out_edge, out_point);
            busy = false;
        }
        const unsigned idle = __ballot_sync(kFull, !busy);
        if (idle)
        {
            const uint32_t s = feeder_take(fd, idle, !busy, lane, n, counter);
            if (s != kNone)
            {
                slot = perm ? __ldg(perm + s) : s;
                p = load_point(q, slot);
                flip = flipv ? (__ldg(flipv + slot) != 0) : false;
                best = rmax ? __ldg(rmax + slot) : INFINITY;
                best2 = best * best;
                found = false;
                busy = true;
                sp = 0;
                node = 0;
            }
            if (fd.exhausted)
            {
                const int walking = __popc(__ballot_sync(kFull, busy));
                if (walking == 0) break;
                if (walking <= tail_lanes)
                { // nothing left to hand out and few lanes still walk: drain the queue (top of the loop), then hand them to the tail launch
                    tail = true;
                    __syncwarp();
                    continue;
                }
            }
        }
        // ---- 3. one traversal step
        if (node != kNone)
        {
            float4 a, b, c, d, e, f;
            ld256(sv.snode + node, a, b);
            ld256(reinterpret_cast<const char *>(sv.snode + node) + 32, c, d);
            ld256(reinterpret_cast<const char *>(sv.snode + node) + 64, e, f);
            const NodeBoxes nb = unpack_boxes(a, b, c);
            const float m0 = box_mindist2(nb.lo0, nb.hi0, p), m1 = box_mindist2(nb.lo1, nb.hi1, p);
            const uint32_t r0 = __float_as_uint(f.z), r1 = __float_as_uint(f.w);
            // the reference's per-child test: is_valid(cone) && overlap(cone, p, box, mindist^2)   (query.cuh:366-367),
            // evaluated only for children that can still beat the current best
            const bool h0 = (m0 <= best2) && (d.w >= 0.0f) && cone_test<kFilter>(V3{d.x, d.y, d.z}, d.w, e.x, p, nb.lo0, nb.hi0, m0);
            const bool h1 = (m1 <= best2) && (f.x >= 0.0f) && cone_test<kFilter>(V3{e.y, e.z, e.w}, f.x, f.y, p, nb.lo1, nb.hi1, m1);
            const bool swap = m1 < m0;
            uint32_t next = kNone;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            {
                const bool second = (ch == 1) != swap; // visit the nearer child first
                const bool h = second ? h1 : h0;
                const float m = second ? m1 : m0;
                const uint32_t r = second ? r1 : r0;
                if (!h) continue;
                if (r & kLeafFlag)
                {
                    const uint32_t pos = atomicAdd(&wq.count, 1u);
                    wq.payload[pos] = r & ~kLeafFlag;
                    wq.owner[pos] = (uint8_t)lane;
                    pend = true;
                }
                else if (next == kNone) next = r;
                else
                {
                    const StackEntry se = StackEntry{r, m};
                    if (sp < kSStack) s_stk[sp][threadIdx.x] = se;
                    else lstk[sp - kSStack] = se;
                    ++sp;
                }
            }
            if (next == kNone)
            {
                while (sp > 0)
                {
                    --sp;
                    const StackEntry se = sp < kSStack ? s_stk[sp][threadIdx.x] : lstk[sp - kSStack];
                    if (se.key <= best2)
                    {
                        next = se.node;
                        break;
                    }
                }
            }
            node = next;
        }
        __syncwarp(); // queue appends of this step are visible to the whole warp before the next count is read
    }
    if (!tail) return;
    // ---- TAIL: the queue is drained and every result is folded into found / best.  Lanes whose walk had already ended hand in;
    // the queries still being walked go on the tail list with the bound found so far, and the launch that follows
    // (k_silhouette_wide over that list) finishes each of them on 32 lanes.
    if (busy && node == kNone)
    {
        out_dist[slot] = found ? best : INFINITY;
        if (kEdge) write_silhouette_point(sv, slot, best_slot, p, found, out_edge, out_point);
    }
    else if (busy)
    {
        const uint32_t at = (uint32_t)atomicAdd(counter + 1, 1ull); // < gridDim.x * blockDim.x entries by construction
        tail_slot[at] = slot;
        tail_bound[at] = best;
        tail_prev[at] = found ? (kEdge ? best_slot : 0u) : kNone; // the LEdge slot that attains `best`, if an edge has been found
    }
}

// One query per WARP (solo_silhouette from the root): for batches too small to fill the machine with per-lane walks the
// run time is the critical path of the most expensive query.
constexpr int kSoloStackSil = 512;
template <int kFilter, bool kEdge>
__global__ void __launch_bounds__(kQueryThreads)
    k_silhouette_wide(SceneView sv, const float *__restrict__ q, const uint8_t *__restrict__ flipv, const float *__restrict__ rmax,
                      const uint32_t *__restrict__ perm, uint32_t n, float *__restrict__ out_dist, uint32_t *__restrict__ out_edge,
                      float *__restrict__ out_point, unsigned long long *counter, const unsigned long long *__restrict__ list_n,
                      const float *__restrict__ list_bound, const uint32_t *__restrict__ list_prev)
{
    __shared__ StackEntry s_solo[kQueryThreads / 32][kSoloStackSil];
    const int lane = threadIdx.x & 31;
    const FlatStack solo{s_solo[threadIdx.x >> 5]};
    // list mode (the tail of k_silhouette_coop): `perm` is the list, *list_n its length, list_bound[i] the bound of entry i and
    // list_prev[i] the edge that set it (kNone: the bound is still the query's own radius)
    if (list_n) n = (uint32_t)*list_n;
    const unsigned long long kRun = list_n ? 1 : 4;
    for (;;)
    {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, kRun);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) break;
        for (unsigned long long s = base; s < base + kRun && s < n; ++s)
        {
            const uint32_t slot = perm ? __ldg(perm + s) : (uint32_t)s;
            const V3 p = load_point(q, slot);
            const bool flip = flipv ? (__ldg(flipv + slot) != 0) : false;
            const float r = list_n ? __ldg(list_bound + s) : (rmax ? __ldg(rmax + slot) : INFINITY);
            uint32_t eslot = kNone;
            float ans = solo_silhouette<kFilter, kEdge, kSoloStackSil>(sv, solo, lane, p, flip, r, eslot);
            if (list_n && !(ans < INFINITY))
            {
                const uint32_t prev = __ldg(list_prev + s);
                if (prev != kNone)
                {
                    ans = r;
                    eslot = prev;
                }
            }
            if (lane == 0)
            {
                out_dist[slot] = ans;
                if (kEdge) write_silhouette_point(sv, slot, eslot, p, ans < INFINITY, out_edge, out_point);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ray intersection (closest hit / any hit)                                               query.cuh:79-169
// ---------------------------------------------------------------------------------------------------------------
template <bool kAnyHit>
__global__ void __launch_bounds__(kQueryThreads)
    k_intersect(SceneView sv, const float *__restrict__ org, const float *__restrict__ dir, const float *__restrict__ tmaxv,
                const uint32_t *__restrict__ perm, uint32_t n, snch_hit *__restrict__ hits, uint8_t *__restrict__ found_out,
                unsigned long long *counter)
{
    const int lane = threadIdx.x & 31;
    Feeder fd{0u, 0u, false};
    StackEntry stk[kStackDepth];
    int sp = 0;
    V3 o = V3{0.f, 0.f, 0.f}, dv = o, dinv = o;
    float max_dist = INFINITY, best_t = INFINITY, best_u = 0.f, best_v = 0.f;
    uint32_t best_prim = kNone, slot = kNone, node = kNone;
    bool found = false;
    for (;;)
    {
        const unsigned idle = __ballot_sync(kFull, node == kNone);
        if (idle)
        {
            const uint32_t s = feeder_take(fd, idle, node == kNone, lane, n, counter);
            if (s != kNone)
            {
                slot = perm ? __ldg(perm + s) : s;
                o = load_point(org, slot);
                dv = load_point(dir, slot);
                dinv = V3{1.0f / dv.x, 1.0f / dv.y, 1.0f / dv.z}; // aabb.cuh:305-312
                max_dist = tmaxv ? __ldg(tmaxv + slot) : INFINITY;
                best_t = INFINITY;
                best_u = best_v = 0.f;
                best_prim = kNone;
                found = false;
                sp = 0;
                node = 0;
            }
            if (fd.exhausted && __all_sync(kFull, node == kNone)) break;
        }
        if (node != kNone)
        {

            float4 a, b, c, d;
            ld256(sv.bnode + node, a, b);
            ld256(reinterpret_cast<const char *>(sv.bnode + node) + 32, c, d);
            const NodeBoxes nb = unpack_boxes(a, b, c);
            float e0, e1;
            const bool h0 = box_ray(nb.lo0, nb.hi0, o, dinv, max_dist, &e0);
            const bool h1 = box_ray(nb.lo1, nb.hi1, o, dinv, max_dist, &e1);
            const uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y);
            const bool swap = h0 && h1 && (e1 < e0); // the reference visits L first on ties (query.cuh:141)
            uint32_t next = kNone;
            bool done = false;
    #pragma unroll
            for (int ch = 0; ch < 2; ++ch)
            {
                const bool second = (ch == 1) != swap;
                const bool h = second ? h1 : h0;
                const float en = second ? e1 : e0;
                const uint32_t r = second ? r1 : r0;
                if (done || !h || en > best_t) continue; // same rejection the reference applies at pop time (query.cuh:106)
                if (r & kLeafFlag)
                {
                    const LTri *tp = sv.ltri + (r & ~kLeafFlag);
                    float4 t0, t1, t2, t3;
                    ld256(tp, t0, t1);
                    ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
                    float t, u, v;
                    if (ray_triangle(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, o, dv, &t, &u, &v) && t < max_dist &&
                        t < best_t)
                    {
                        best_t = t;
                        best_u = u;
                        best_v = v;
                        best_prim = __float_as_uint(t0.w);
                        found = true;
                        if (kAnyHit) done = true;
                    }
                }
                else if (next == kNone) next = r;
                else
                {
                    stk[sp] = StackEntry{r, en};
                    ++sp;
                }
            }
            if (done) next = kNone;
            else if (next == kNone)
            {
                while (sp > 0)
                {
                    --sp;
                    const StackEntry se = stk[sp];
                    if (!(se.key > best_t))
                    {
                        next = se.node;
                        break;
                    }
                }
            }
            if (next == kNone)
            {
                if (found_out) found_out[slot] = found ? 1 : 0;
                if (!kAnyHit && hits)
                {
                    snch_hit h;
                    h.t = best_t;
                    h.u = best_u;
                    h.v = best_v;
                    h.prim = best_prim;
                    hits[slot] = h;
                }
            }
            node = next;
        }
    }
}

// Per-lane walk in the REFERENCE'S ORDER with the leaf tests taken out of the traversal step (default, "query.ray_kernel" = 1).
// profiles/r2a: in k_intersect above a warp runs the Moeller-Trumbore code on the 2-3 lanes that happen to stand at a leaf in
// almost every step, and tests a far leaf child before the near subtree has been searched.  Here a hit leaf child is not
// tested where it is met:
//   * the children of a node are taken nearer first, as in the reference (query.cuh:128-160, L first on ties).  If the first
//     hit child is an internal node the lane descends into it and the other hit child — leaf or internal — goes on the stack
//     with its entry distance; if the first hit child is a LEAF the lane parks on it (`pleaf`), the other child goes on the
//     stack and the lane waits;
//   * when `flush_lanes` lanes of the warp are parked (or no lane can walk) the warp tests all parked leaves at once — the
//     triangle code runs on many lanes instead of two — and each of those lanes then pops its next entry under the
//     reference's pop-time rejection (query.cuh:106) with its NEW best t; a popped leaf parks the lane again.
//   So every lane tests exactly the leaves, in exactly the order, of the reference's stack walk: same t, same triangle.
//   * the lowest kRStack stack levels live in shared memory (bank = lane), deeper ones in local memory;
//   * finished lanes take their next ray when `refill_lanes` of them are idle (the refill code — six loads, three divisions —
//     is then not run for one lane at a time in nearly every step).
constexpr int kRStack = 12;
template <bool kAnyHit>
__global__ void __launch_bounds__(kQueryThreads, 10)
    k_intersect_parked(SceneView sv, const float *__restrict__ org, const float *__restrict__ dir, const float *__restrict__ tmaxv,
                       const uint32_t *__restrict__ perm, uint32_t n, snch_hit *__restrict__ hits, uint8_t *__restrict__ found_out,
                       unsigned long long *counter, int flush_lanes, int refill_lanes)
{
    __shared__ StackEntry s_stk[kRStack][kQueryThreads];
    const int lane = threadIdx.x & 31;
    Feeder fd{0u, 0u, false};
    StackEntry lstk[kStackDepth - kRStack];
    int sp = 0;
    V3 o = V3{0.f, 0.f, 0.f}, dv = o, dinv = o;
    float max_dist = INFINITY, best_t = INFINITY, best_u = 0.f, best_v = 0.f;
    uint32_t best_prim = kNone, slot = kNone, node = kNone, pleaf = kNone;
    bool busy = false;
    for (;;)
    {
        // ---- 1. parked leaves
        bool pop = false; // this lane needs the next entry of its stack (set by the leaf test above / the step below, served in 3b)
        const unsigned parked = __ballot_sync(kFull, pleaf != kNone);
        if (parked && (__popc(parked) >= flush_lanes || !__any_sync(kFull, node != kNone)))
        {
            if (pleaf != kNone)
            {
                const LTri *tp = sv.ltri + pleaf;
                float4 t0, t1, t2, t3;
                ld256(tp, t0, t1);
                ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
                float t, u, v;
                if (ray_triangle(V3{t0.x, t0.y, t0.z}, V3{t1.x, t1.y, t1.z}, V3{t2.x, t2.y, t2.z}, o, dv, &t, &u, &v) && t < max_dist && t < best_t)
                {
                    best_t = t;
                    best_u = u;
                    best_v = v;
                    best_prim = __float_as_uint(t0.w);
                    if (kAnyHit) sp = 0;
                }
                pleaf = kNone;
                pop = true;
            }
        }
        // ---- 2. idle lanes take the next ray
        const unsigned idle = __ballot_sync(kFull, !busy);
        if (idle)
        {
            if (!fd.exhausted && (__popc(idle) >= refill_lanes || !__any_sync(kFull, node != kNone || pop)))
            {
                const uint32_t s = feeder_take(fd, idle, !busy, lane, n, counter);
                if (s != kNone)
                {
                    slot = perm ? __ldg(perm + s) : s;
                    o = load_point(org, slot);
                    dv = load_point(dir, slot);
                    dinv = V3{1.0f / dv.x, 1.0f / dv.y, 1.0f / dv.z}; // aabb.cuh:305-312
                    max_dist = tmaxv ? __ldg(tmaxv + slot) : INFINITY;
                    best_t = INFINITY;
                    best_u = best_v = 0.f;
                    best_prim = kNone;
                    busy = true;
                    sp = 0;
                    node = 0;
                    pleaf = kNone;
                }
            }
            if (fd.exhausted && idle == kFull) break;
        }
        // ---- 3. one traversal step
        if (node != kNone)
        {
            float4 a, b, c, d;
            ld256(sv.bnode + node, a, b);
            ld256(reinterpret_cast<const char *>(sv.bnode + node) + 32, c, d);
            const NodeBoxes nb = unpack_boxes(a, b, c);
            float e0, e1;
            bool h0 = box_ray(nb.lo0, nb.hi0, o, dinv, max_dist, &e0);
            bool h1 = box_ray(nb.lo1, nb.hi1, o, dinv, max_dist, &e1);
            h0 = h0 && !(e0 > best_t); // the rejection the reference applies when it pops the child (query.cuh:106)
            h1 = h1 && !(e1 > best_t);
            uint32_t r0 = __float_as_uint(d.x), r1 = __float_as_uint(d.y);
            if (h0 && h1 && e1 < e0)
            { // nearer child first; L first on ties (query.cuh:141)
                const uint32_t tr = r0;
                r0 = r1;
                r1 = tr;
                const float te = e0;
                e0 = e1;
                e1 = te;
            }
            if (h0 && h1)
            {
                const StackEntry se = StackEntry{r1, e1};
                if (sp < kRStack) s_stk[sp][threadIdx.x] = se;
                else lstk[sp - kRStack] = se;
                ++sp;
            }
            node = kNone;
            if (h0 || h1)
            {
                const uint32_t r = h0 ? r0 : r1;
                if (r & kLeafFlag) pleaf = r & ~kLeafFlag;
                else node = r;
            }
            else pop = true;
        }
        // ---- 3b. next entry of the stack that survives the pop-time rejection: an internal node to walk, a leaf to park on, or
        // the end of the ray — ONE site for the lanes coming from the leaf test and from a dead end, so they run it together
        if (pop)
        {
            while (sp > 0)
            {
                --sp;
                const StackEntry se = sp < kRStack ? s_stk[sp][threadIdx.x] : lstk[sp - kRStack];
                if (se.key > best_t) continue;
                if (se.node & kLeafFlag) pleaf = se.node & ~kLeafFlag;
                else node = se.node;
                pop = false;
                break;
            }
            if (pop)
            { // stack empty: the ray is done
                if (found_out) found_out[slot] = best_prim != kNone ? 1 : 0;
                if (!kAnyHit && hits)
                {
                    snch_hit h;
                    h.t = best_t;
                    h.u = best_u;
                    h.v = best_v;
                    h.prim = best_prim;
                    hits[slot] = h;
                }
                busy = false;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// SampleTriangleInSphere                                              sample.cuh:23-92 + 7-21, scene.cuh:14-27
// One root-to-leaf path per query: no stack, no variance in length beyond the tree depth -> one query per thread.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kQueryThreads)
    k_sample(SceneView sv, const float *__restrict__ sph, const float *__restrict__ rnd, const uint32_t *__restrict__ perm, uint32_t n,
             int32_t *__restrict__ out_idx, float *__restrict__ out_pdf, float *__restrict__ out_pt)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint64_t i = perm ? __ldg(perm + s) : s;
    const V3 ctr = V3{__ldg(sph + 4 * i), __ldg(sph + 4 * i + 1), __ldg(sph + 4 * i + 2)};
    const float radius = __ldg(sph + 4 * i + 3);
    float u = __ldg(rnd + 3 * i);
    float path = 1.0f;
    int32_t idx = -1;
    float pdf = 0.0f;
    V3 pt = V3{0.f, 0.f, 0.f};
    uint32_t node = 0;
    for (;;)
    {
        float4 a, b, c, d;
        ld256(sv.bnode + node, a, b);
        ld256(reinterpret_cast<const char *>(sv.bnode + node) + 32, c, d);
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
            const LTri *tp = sv.ltri + (r & ~kLeafFlag);
            float4 t0, t1, t2, t3;
            ld256(tp, t0, t1);
            ld256(reinterpret_cast<const char *>(tp) + 32, t2, t3);
            const V3 pa = V3{t0.x, t0.y, t0.z}, pb = V3{t1.x, t1.y, t1.z}, pc = V3{t2.x, t2.y, t2.z};
            if (sphere_triangle(pa, pb, pc, ctr, radius))
            {
                idx = (int32_t)__float_as_uint(t0.w);
                pdf = path / triangle_area(pa, pb, pc);
                float su = __ldg(rnd + 3 * i + 1), sv2 = __ldg(rnd + 3 * i + 2);
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

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// Counts the traversal launch and, when kernel timing is on ("query.time_kernels"), brackets it with its own CUDA event pair
// on the launching stream; elapsed times are folded into the counters when they are read (QueryCounters::fold).
struct TraversalTimer
{
    QueryCounters *qc;
    cudaStream_t st;
    QueryCounters::Pair pair{nullptr, nullptr};
    bool timed = false;
    TraversalTimer(QueryCounters *qc_, cudaStream_t st_) : qc(qc_), st(st_)
    {
        if (!qc) return;
        qc->launches += 1;
        qc->traversal_launches += 1;
        if (qc->time_kernels.load()) timed = qc->begin(pair, st);
    }
    ~TraversalTimer()
    {
        if (timed) qc->end(pair, st);
    }
};
bool QueryCounters::begin(Pair &p, cudaStream_t st)
{
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!spare.empty())
        {
            p = spare.back();
            spare.pop_back();
        }
    }
    if (!p.a && (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess))
    {
        cudaGetLastError();
        return false;
    }
    return cudaEventRecord(p.a, st) == cudaSuccess;
}
void QueryCounters::end(const Pair &p, cudaStream_t st)
{
    cudaEventRecord(p.b, st);
    std::lock_guard<std::mutex> lock(mu);
    pending.push_back(p);
}
void QueryCounters::fold()
{
    std::lock_guard<std::mutex> lock(mu);
    for (const Pair &p : pending)
    {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) traversal_ms += ms;
        else cudaGetLastError();
        spare.push_back(p);
    }
    pending.clear();
}
void QueryCounters::reset()
{
    fold();
    launches = 0;
    traversal_launches = 0;
    std::lock_guard<std::mutex> lock(mu);
    traversal_ms = 0.0;
}
void QueryCounters::release()
{
    fold();
    std::lock_guard<std::mutex> lock(mu);
    for (const Pair &p : spare)
    {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    spare.clear();
}

static inline unsigned grid_for(uint64_t n) { return (unsigned)((n + kQueryThreads - 1) / kQueryThreads); }

// bytes of device scratch one batch of n queries needs (ordering buffers + sort counters + the work counter)
uint64_t query_scratch_bytes(uint64_t n, const QueryTuning &t)
{
    uint64_t b = kScratchHeader + kTailBytes; // work counters + query box | tail list of the silhouette kernel
    if (t.sort_min_n > 0 && n >= (uint64_t)t.sort_min_n) b += 4 * align_up(n * 4, 256) + align_up(sort_scratch_elems(n) * 4, 256);
    return b;
}

// Lays the scratch out and, for batches worth ordering, produces the Morton permutation.  *perm_out = nullptr otherwise.
int prepare_batch(const QueryTuning &t, bool order, const float *pts, int stride, const float *radius, uint32_t n, unsigned char *scratch,
                  cudaStream_t st, unsigned long long **counter_out, const uint32_t **perm_out, QueryCounters *qc, int dims, int radius_desc,
                  const float *dirs)
{
    SNCH_CUDA(cudaMemsetAsync(scratch, 0, 64, st));
    *counter_out = reinterpret_cast<unsigned long long *>(scratch);
    *perm_out = nullptr;
    if (!(order && t.sort_min_n > 0 && n >= (uint32_t)t.sort_min_n)) return SNCH_OK;
    int *box = reinterpret_cast<int *>(scratch + 64);
    const uint64_t a = align_up((uint64_t)n * 4, 256);
    unsigned char *ord = scratch + kScratchHeader + kTailBytes;
    uint32_t *keys = reinterpret_cast<uint32_t *>(ord);
    uint32_t *perm = reinterpret_cast<uint32_t *>(ord + a);
    uint32_t *ktmp = reinterpret_cast<uint32_t *>(ord + 2 * a);
    uint32_t *vtmp = reinterpret_cast<uint32_t *>(ord + 3 * a);
    uint32_t *sscr = reinterpret_cast<uint32_t *>(ord + 4 * a);
    const unsigned g = (n + 255) / 256;
    k_query_box_init<<<1, 32, 0, st>>>(box);
    k_query_bounds<<<g < 1184 ? g : 1184, 256, 0, st>>>(pts, stride, dims, n, box);
    k_query_keys<<<g, 256, 0, st>>>(pts, stride, dims, n, box, radius, radius_desc, dirs, keys, perm);
    int bits = t.sort_bits < 8 ? 8 : (t.sort_bits > 30 ? 30 : t.sort_bits);
    const int sort_launches = radix_sort_pairs(keys, perm, ktmp, vtmp, n, bits, sscr, st, 30 - bits);
    if (qc) qc->launches += 3 + sort_launches;
    SNCH_CUDA(cudaGetLastError());
    *perm_out = perm;
    return SNCH_OK;
}

// silhouette traversal: one query per warp for batches too small to fill the machine ("query.wide_max_n_sil"), per-lane walks
// with the warp-shared leaf queue otherwise; cone test per "query.cone_filter"; the edge-carrying instantiation only when
// the caller asked for the edge index or the point
template <int kFilter, bool kEdge>
static void launch_silhouette_kernel(const SceneView &v, const QueryTuning &t, const float *q, const uint8_t *flip, const float *rmax,
                                     const uint32_t *perm, uint32_t n, float *dist, uint32_t *edge, float *point, unsigned long long *counter,
                                     unsigned char *tail, cudaStream_t st, QueryCounters *qc)
{
    if (n < (uint32_t)t.wide_max_n_sil)
    {
        if (qc) qc->last_kernel = kEdge ? "k_silhouette_wide<edge>" : "k_silhouette_wide";
        k_silhouette_wide<kFilter, kEdge><<<persistent_grid(k_silhouette_wide<kFilter, kEdge>, t, n < (1u << 26) ? n * 16 : n), kQueryThreads, 0, st>>>(
            v, q, flip, rmax, perm, n, dist, edge, point, counter, nullptr, nullptr, nullptr);
        return;
    }
    if (qc) qc->last_kernel = kEdge ? "k_silhouette_coop<edge>" : "k_silhouette_coop";
    const unsigned grid = persistent_grid(k_silhouette_coop<kFilter, kEdge>, t, n);
    // the tail list holds at most one entry per resident lane
    uint32_t *tail_slot = reinterpret_cast<uint32_t *>(tail);
    float *tail_bound = reinterpret_cast<float *>(tail + kTailEntries * 4);
    uint32_t *tail_prev = reinterpret_cast<uint32_t *>(tail + kTailEntries * 8);
    const int tl = (t.sil_tail > 0 && (uint64_t)grid * kQueryThreads <= kTailEntries) ? (t.sil_tail > 31 ? 31 : t.sil_tail) : 0;
    const uint32_t flush_at = (uint32_t)(t.sil_flush < 1 ? 1 : (t.sil_flush > kLeafFlushAt ? kLeafFlushAt : t.sil_flush));
    k_silhouette_coop<kFilter, kEdge><<<grid, kQueryThreads, 0, st>>>(v, q, flip, rmax, perm, n, dist, edge, point, counter, tl, tail_slot, tail_bound, tail_prev, flush_at,
                                                                      (uint32_t)(t.sil_chunk > 0 ? t.sil_chunk : sil_chunk_for_host(n)));
    if (tl)
    { // finish the listed queries one per warp; its own work counter is counter[2], the list length counter[1]
        if (qc) qc->launches += 1;
        k_silhouette_wide<kFilter, kEdge><<<grid < 592 ? grid : 592, kQueryThreads, 0, st>>>(v, q, flip, rmax, tail_slot, 0u, dist, edge, point, counter + 2, counter + 1,
                                                                                          tail_bound, tail_prev);
    }
}
static void launch_silhouette_lanes(const SceneView &v, const QueryTuning &t, const float *q, const uint8_t *flip, const float *rmax,
                                    const uint32_t *perm, uint32_t n, float *dist, uint32_t *edge, float *point, unsigned long long *counter,
                                    unsigned char *tail, cudaStream_t st, QueryCounters *qc)
{
    const bool e = edge || point;
    if (t.cone_filter && e) launch_silhouette_kernel<1, true>(v, t, q, flip, rmax, perm, n, dist, edge, point, counter, tail, st, qc);
    else if (t.cone_filter) launch_silhouette_kernel<1, false>(v, t, q, flip, rmax, perm, n, dist, edge, point, counter, tail, st, qc);
    else if (e) launch_silhouette_kernel<0, true>(v, t, q, flip, rmax, perm, n, dist, edge, point, counter, tail, st, qc);
    else launch_silhouette_kernel<0, false>(v, t, q, flip, rmax, perm, n, dist, edge, point, counter, tail, st, qc);
}

// One query per warp (k_closest_wide) instead of packets: for batches too small to fill the machine, and for batches SPARSE
// relative to the mesh — with fewer than ~2 queries per triangle the 32 Morton neighbours of a packet share the top of the
// tree but hardly any leaf (their candidate sets, a cap of radius ~sqrt(2 d h) around each closest point, no longer overlap),
// so the packet walks the union of 32 disjoint leaf sets one node at a time while the cooperative walk opens 32 nodes of one
// query per step.  Measured: 8.4M queries on 10M triangles 115.9 (packets) vs 95.7 ms; 2.1M on 4M triangles 30.2 vs 18.3 ms;
// 4.2M on 1M triangles 18.5 vs 25.4 ms (packets stay).
static bool use_wide_closest(const QueryTuning &t, uint64_t n, uint64_t n_tris)
{
    return t.wide_max_n > 0 && (n < (uint64_t)t.wide_max_n || n < 2 * n_tris);
}
static void launch_closest_kernel(const SceneView &v, const QueryTuning &t, const float *q, const uint32_t *perm, uint32_t n, uint32_t *idx,
                                  float *dist, unsigned long long *counter, cudaStream_t st, QueryCounters *qc)
{
    if (use_wide_closest(t, n, v.n_tris))
    {
        if (qc) qc->last_kernel = "k_closest_wide";
        k_closest_wide<<<persistent_grid(k_closest_wide, t, n < (1u << 26) ? n * 16 : n), kQueryThreads, 0, st>>>(v, q, perm, n, idx, dist, counter, t.seed);
    }
    else
    {
        if (qc) qc->last_kernel = "k_closest_packet";
        k_closest_packet<<<persistent_grid(k_closest_packet, t, n), kQueryThreads, 0, st>>>(v, q, perm, n, idx, dist, counter, t.seed);
    }
}

int launch_closest(const SceneView &v, const QueryTuning &t, const float *q, uint64_t n, uint32_t *idx, float *dist, unsigned char *scratch,
                   cudaStream_t st, QueryCounters *qc)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0)
    {
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, idx, dist, nullptr, nullptr, nullptr, nullptr, nullptr);
        SNCH_CUDA(cudaGetLastError());
        return SNCH_OK;
    }
    unsigned long long *counter;
    const uint32_t *perm;
    const int rc = prepare_batch(t, true, q, 3, nullptr, (uint32_t)n, scratch, st, &counter, &perm, qc);
    if (rc != SNCH_OK) return rc;
    TraversalTimer tt(qc, st);
    launch_closest_kernel(v, t, q, perm, (uint32_t)n, idx, dist, counter, st, qc);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}
int launch_silhouette(const SceneView &v, const QueryTuning &t, const float *q, const uint8_t *flip, const float *rmax, uint64_t n,
                      float *dist, uint32_t *edge, float *point, unsigned char *scratch, cudaStream_t st, QueryCounters *qc)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0)
    {
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, edge, dist, nullptr, nullptr, nullptr, nullptr, point);
        SNCH_CUDA(cudaGetLastError());
        return SNCH_OK;
    }
    unsigned long long *counter;
    const uint32_t *perm;
    const int rc = prepare_batch(t, true, q, 3, t.sort_radius ? rmax : nullptr, (uint32_t)n, scratch, st, &counter, &perm, qc, 3,
                                 t.sort_radius >= 2 ? t.sort_radius - 1 : 0);
    if (rc != SNCH_OK) return rc;
    TraversalTimer tt(qc, st);
    launch_silhouette_lanes(v, t, q, flip, rmax, perm, (uint32_t)n, dist, edge, point, counter, scratch + kScratchHeader, st, qc);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}
static void launch_intersect_kernel(const SceneView &v, const QueryTuning &t, const float *o, const float *d, const float *tmax, const uint32_t *perm,
                                    uint32_t n, snch_hit *hits, uint8_t *found, bool any_hit, unsigned long long *counter, cudaStream_t st,
                                    QueryCounters *qc)
{
    // "query.ray_kernel" = 0: leaves tested where they are met (k_intersect), for every batch — the A/B of the knob tests.  Its
    // hit flag and t equal the reference's, but among triangles hit at the SAME t (shared edges, duplicates, t = +0 / -0 for a
    // ray that starts on the surface) it reports the first in ITS order, so the triangle — and the sign of a zero t — can differ.
    if (t.ray_kernel == 0)
    {
        if (qc) qc->last_kernel = "k_intersect";
        if (any_hit) k_intersect<true><<<persistent_grid(k_intersect<true>, t, n), kQueryThreads, 0, st>>>(v, o, d, tmax, perm, n, hits, found, counter);
        else k_intersect<false><<<persistent_grid(k_intersect<false>, t, n), kQueryThreads, 0, st>>>(v, o, d, tmax, perm, n, hits, found, counter);
        return;
    }
    if (qc) qc->last_kernel = "k_intersect_parked";
    // The reference-order walk for every batch.  A batch that does not fill the machine (one ray per lane: config C1, 64K rays)
    // runs as long as its longest ray, so there a lane tests its leaf at once (flush at 1 parked lane): 0.106 ms on C1 against
    // 0.124 with the large-batch setting and 0.100 for k_intersect (tools/ray_small_exp.py); "query.ray_kernel" = 2 keeps the
    // large-batch setting for every batch.
    const bool small = t.ray_kernel == 1 && n < (1u << 20);
    const int fl = small ? 1 : (t.ray_flush < 1 ? 1 : t.ray_flush), rl = t.ray_refill < 1 ? 1 : t.ray_refill;
    if (any_hit)
        k_intersect_parked<true><<<persistent_grid(k_intersect_parked<true>, t, n), kQueryThreads, 0, st>>>(v, o, d, tmax, perm, n, hits, found, counter, fl, rl);
    else
        k_intersect_parked<false><<<persistent_grid(k_intersect_parked<false>, t, n), kQueryThreads, 0, st>>>(v, o, d, tmax, perm, n, hits, found, counter, fl, rl);
}
int launch_intersect(const SceneView &v, const QueryTuning &t, const float *o, const float *d, const float *tmax, uint64_t n, snch_hit *hits,
                     uint8_t *found, int any_hit, unsigned char *scratch, cudaStream_t st, QueryCounters *qc)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0)
    {
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, nullptr, nullptr, any_hit ? nullptr : hits, found, nullptr, nullptr, nullptr);
        SNCH_CUDA(cudaGetLastError());
        return SNCH_OK;
    }
    unsigned long long *counter;
    const uint32_t *perm;
    // "query.sort_rays" = -1 (default): rays are visited in Morton order of their origins when the traversal records (128 B per
    // triangle) do not fit the 126 MB L2 — 16.7M rays: 7.38 -> 6.80 ms at 4M triangles, 13.5 -> 8.5 ms at 10M, but 5.12 -> 5.90 ms
    // at 1M, where the ordering costs more than it returns (tools/c4_ray_exp.py, profiles/r2x_c4_ray_order.json)
    // ... and the batch has 4M rays or more: a 2M-ray shard is latency-bound and the ordering's own launches cost what the better
    // locality returns (1.47 ms in the caller's order, 1.53 ordered, per rank of the 8-GPU C4 run; a wash at 4.2M rays)
    const int sr = t.sort_rays >= 0 ? t.sort_rays : ((v.n_tris >= (1u << 21) && n >= (1ull << 22)) ? 1 : 0);
    const int rc = prepare_batch(t, sr != 0, o, 3, nullptr, (uint32_t)n, scratch, st, &counter, &perm, qc, 3, 0, sr >= 2 ? d : nullptr);
    if (rc != SNCH_OK) return rc;
    TraversalTimer tt(qc, st);
    launch_intersect_kernel(v, t, o, d, tmax, perm, (uint32_t)n, hits, found, any_hit != 0, counter, st, qc);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}
int launch_sample(const SceneView &v, const QueryTuning &t, const float *sph, const float *rnd, uint64_t n, int32_t *idx, float *pdf,
                  float *pt, unsigned char *scratch, cudaStream_t st, QueryCounters *qc)
{
    if (n == 0) return SNCH_OK;
    if (v.n_tris == 0)
    {
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, nullptr, nullptr, nullptr, nullptr, idx, pdf, pt);
        SNCH_CUDA(cudaGetLastError());
        return SNCH_OK;
    }
    (void)t;
    (void)scratch; // one short root-to-leaf path per query: neither ordering nor work stealing pays for itself here
    TraversalTimer tt(qc, st);
    if (qc) qc->last_kernel = "k_sample";
    k_sample<<<grid_for(n), kQueryThreads, 0, st>>>(v, sph, rnd, nullptr, (uint32_t)n, idx, pdf, pt);
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// One wavefront walk-on-stars step: per walker the four reference calls an Elaina-style stage makes,
//   (i, d) = nearest(p);  s = nearest_silhouette(p, flip) within d;  R = min(d, s)   [star radius]
//   hit    = ray_intersect(p, dir, t_max = R);  (tri, pdf, y) = sample_object_in_sphere(sphere(p, R), u)
// as ONE call: the Morton ordering of the walkers is computed once and shared by all four traversals, and the star
// radius never leaves the device.  Every output equals what the four *_batch calls return for the same inputs.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_star_radius(const float *__restrict__ pts, const float *__restrict__ d_closest, const float *__restrict__ d_sil, uint32_t n,
                              float *__restrict__ radius, float *__restrict__ spheres)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = fminf(d_closest[i], d_sil[i]);
    if (radius) radius[i] = r;
    reinterpret_cast<float4 *>(spheres)[i] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], r);
}

uint64_t wost_scratch_bytes(uint64_t n, const QueryTuning &t) { return query_scratch_bytes(n, t) + 4 * align_up(n * 4, 256) + align_up(n * 16, 256); }

int launch_wost_step(const SceneView &v, const QueryTuning &t, const WostBuffers &io, uint64_t n, unsigned char *scratch, cudaStream_t st,
                     QueryCounters *qc)
{
    if (n == 0) return SNCH_OK;
    const uint64_t qs = query_scratch_bytes(n, t), a4 = align_up(n * 4, 256);
    float *d_closest = io.closest_distance ? io.closest_distance : reinterpret_cast<float *>(scratch + qs);
    float *d_sil = io.silhouette_distance ? io.silhouette_distance : reinterpret_cast<float *>(scratch + qs + a4);
    float *radius = io.star_radius ? io.star_radius : reinterpret_cast<float *>(scratch + qs + 2 * a4);
    uint32_t *c_index = io.closest_index ? io.closest_index : reinterpret_cast<uint32_t *>(scratch + qs + 3 * a4);
    float *spheres = reinterpret_cast<float *>(scratch + qs + 4 * a4);
    if (v.n_tris == 0)
    {
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, c_index, d_closest, io.hits, io.found, io.sample_index, io.sample_pdf,
                                                            io.sample_point);
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, io.silhouette_edge, d_sil, nullptr, nullptr, nullptr, nullptr, io.silhouette_point);
        k_fill_empty<<<grid_for(n), kQueryThreads, 0, st>>>(n, nullptr, radius, nullptr, nullptr, nullptr, nullptr, nullptr);
        SNCH_CUDA(cudaGetLastError());
        return SNCH_OK;
    }
    const uint32_t m = (uint32_t)n;
    unsigned long long *counter;
    const uint32_t *perm;
    const int rc = prepare_batch(t, true, io.points, 3, nullptr, m, scratch, st, &counter, &perm, qc);
    if (rc != SNCH_OK) return rc;
    {
        TraversalTimer tt(qc, st);
        launch_closest_kernel(v, t, io.points, perm, m, c_index, d_closest, counter, st, qc);
    }
    SNCH_CUDA(cudaMemsetAsync(counter, 0, 24, st));
    {
        TraversalTimer tt(qc, st);
        launch_silhouette_lanes(v, t, io.points, io.flip, d_closest, perm, m, d_sil, io.silhouette_edge, io.silhouette_point, counter, scratch + kScratchHeader, st,
                                qc);
    }
    k_star_radius<<<(m + 255) / 256, 256, 0, st>>>(io.points, d_closest, d_sil, m, radius, spheres);
    if (qc) qc->launches += 1;
    if (io.dirs && (io.hits || io.found))
    {
        SNCH_CUDA(cudaMemsetAsync(counter, 0, 8, st));
        TraversalTimer tt(qc, st);
        // the walkers' Morton order is already there: the rays use it unless ordering is switched off ("query.sort_rays" = 0)
        launch_intersect_kernel(v, t, io.points, io.dirs, radius, t.sort_rays ? perm : nullptr, m, io.hits, io.found, false, counter, st, qc);
    }
    if (io.rnd && io.sample_index && io.sample_pdf)
    {
        TraversalTimer tt(qc, st);
        if (qc) qc->last_kernel = "k_sample";
        k_sample<<<grid_for(n), kQueryThreads, 0, st>>>(v, spheres, io.rnd, perm, m, io.sample_index, io.sample_pdf, io.sample_point);
    }
    SNCH_CUDA(cudaGetLastError());
    return SNCH_OK;
}

} // namespace snch
