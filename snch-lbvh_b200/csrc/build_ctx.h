// build_ctx.h — kernel argument block shared by the scene builder (build.cu) and the generic leaf-array builder
// (generic_build.cu), and the one kernel both launch.
#pragma once
#include "layout.h"

namespace snch
{
struct BuildCtx
{
    uint32_t n, n_edges;
    const float3 *verts;
    const RefEdge *edges;
    const RefTriangle *objects;
    RefNode *nodes;
    RefAabb *aabbs;
    RefCone *cones;
    uint32_t *morton, *sorted_idx;
    uint2 *ranges;
    uint8_t *q1;
    BNode *bnode;
    SNode *snode;
    LTri *ltri;
    LEdge *ledge;
    uint32_t *edge_off;
    // scratch
    int *scene_box; // 6 ordered ints: lo xyz, hi xyz
    uint32_t *flags;
    uint32_t *escapes;  // k_refit_coop -> k_refit_top: roots of the per-CTA subtrees (node ids), counters[2] of them
    uint32_t *counters; // [0] collision, [1] q1 events, [2] escapes
};

// Karras 2012 on the augmented key (morton << 32 | index): uses n, morton, sorted_idx, nodes, ranges, counters[0]
__global__ void k_hierarchy(BuildCtx c);

} // namespace snch
