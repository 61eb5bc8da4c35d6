// layout.h — device memory layout of a built scene (one pointer-free arena per GPU).
//
// The arena starts with an ArenaHeader; every array is addressed by a byte offset from the arena base, so the whole
// allocation can be broadcast verbatim to other GPUs.  Two families of arrays live in it:
//   * the reference-layout arrays that lbvh::bvh_device exposes (bvh.cuh:48-54) — kept for drop-in compatibility and
//     for bit-exact parity dumps;
//   * the traversal records the batched kernels read: one record per INTERNAL node holding both children, so a
//     traversal step is one contiguous, sector-aligned read instead of the reference's 3-4 dependent gathers.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace snch
{

constexpr uint64_t kArenaMagic = 0x534e43484c425648ull; // "SNCHLBVH"
constexpr uint32_t kArenaVersion = 5; // 5: LEdge carries its boundary flag in id.y (was: NaN in n0.x)
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr uint32_t kNone = 0xFFFFFFFFu;

// reference record layouts -------------------------------------------------------------------
struct RefNode // bvh.cuh:27-33
{
    uint32_t parent, left, right, object;
};
struct RefAabb // aabb.cuh:12-15 (upper first)
{
    float ux, uy, uz, lx, ly, lz;
};
struct RefCone // cone.cuh:10-16
{
    float ax, ay, az, half_angle, radius;
};
struct RefEdge // scene<3>::silhouette_edge, scene.cuh:709-712 (int4 + pointer, 32 B)
{
    int4 indices;
    const float3 *vertices;
    uint64_t pad_;
};
struct RefTriangle // scene<3>::triangle, scene.cuh:826-831 (40 B)
{
    int3 v;
    int3 owned;
    const float3 *vertices;
    const RefEdge *silhouettes;
};
static_assert(sizeof(RefNode) == 16 && sizeof(RefAabb) == 24 && sizeof(RefCone) == 20, "reference layout");
static_assert(sizeof(RefEdge) == 32 && sizeof(RefTriangle) == 40, "reference layout");

// traversal records ---------------------------------------------------------------------------------------------
// Both children of one internal node.  Box floats are packed lo.xyz, hi.xyz per child over three float4.
// ref: internal child -> node index; leaf child -> kLeafFlag | payload
//   BNode payload  = sorted leaf position k   (LTri[k])
//   SNode payload  = (first_edge << 2) | edge_count   (LEdge[first_edge .. first_edge+count))
struct __align__(32) BNode // 64 B: closest-point, ray, sphere sampling (two 256-bit loads)
{
    float4 a; // lo0.x lo0.y lo0.z hi0.x
    float4 b; // hi0.y hi0.z lo1.x lo1.y
    float4 c; // lo1.z hi1.x hi1.y hi1.z
    uint32_t ref0, ref1, parent, pad;
};
struct __align__(32) SNode // 96 B: silhouette traversal = boxes + both normal cones + refs (3 sectors)
{
    float4 a, b, c; // boxes as in BNode
    float4 d;       // axis0.xyz half0
    float4 e;       // radius0 axis1.xyz
    float4 f;       // half1 radius1 ref0(bits) ref1(bits)
};
struct __align__(32) LTri // 64 B (two sectors, two 256-bit loads), Morton (leaf) order
{
    float4 v0; // xyz, w = object index bits
    float4 v1;
    float4 v2;
    float4 plane; // unit normal xyz (NaN if degenerate), w = max(|v1 - v0|, |v2 - v0|) rounded up: the cheap distance lower bound
};
struct __align__(32) LEdge // 64 B, grouped by owning leaf in Morton order
{
    float4 a; // pa.xyz pb.x
    float4 b; // pb.y pb.z n0.x n0.y       (unit normals; NaN for a zero-area face, 0 for a missing one)
    float4 c; // n0.z n1.xyz
    float4 id; // x = edge id bits (index into scene<3>::silhouettes), y = 1 (bits) for a boundary edge (fewer than two faces: always a silhouette), zw = 0
};
static_assert(sizeof(BNode) == 64 && sizeof(SNode) == 96 && sizeof(LTri) == 64 && sizeof(LEdge) == 64, "record sizes");

struct ArenaHeader
{
    uint64_t magic;
    uint32_t version, n_tris, n_verts, n_edges, n_nodes, n_internal;
    uint32_t collision, q1_nodes;
    float scene_lo[3], scene_hi[3];
    uint64_t total_bytes;
    // input geometry + adjacency
    uint64_t off_vertices;  // float3[n_verts]
    uint64_t off_edges;     // RefEdge[n_edges]   (pointer field patched per device)
    uint64_t off_objects;   // RefTriangle[n_tris] (pointer fields patched per device), ORIGINAL order
    uint64_t off_tri_edges; // int3[n_tris]
    // reference-layout tree
    uint64_t off_nodes, off_aabbs, off_cones;
    uint64_t off_morton, off_sorted_idx, off_ranges, off_q1;
    // traversal records
    uint64_t off_bnode, off_snode, off_ltri, off_ledge, off_edge_off;
    uint64_t reserved[8];
};

// Device-side view with resolved pointers (built on the host from header + base; passed by value to kernels).
struct SceneView
{
    uint32_t n_tris, n_verts, n_edges, n_internal;
    const float3 *vertices;
    const RefEdge *edges;
    const RefTriangle *objects;
    const BNode *bnode;
    const SNode *snode;
    const LTri *ltri;
    const LEdge *ledge;
    const uint32_t *edge_off; // per sorted leaf: (first_edge << 2) | edge_count
};

inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

} // namespace snch
