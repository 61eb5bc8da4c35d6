/* snch_b200.h — C-ABI of the B200-native SNCH-LBVH hot path (libsnch_b200.so).
 *
 * The reference (tyanyuy3125/snch-lbvh) is a header-only C++/Thrust library with no FFI layer: its contract is the
 * source-level API in namespace lbvh.  This C-ABI is the boundary underneath our drop-in headers
 * (the .cuh files under include/snch_lbvh): every entry point names the reference interface it replaces.  Plain pointers and sizes
 * only; no C++ or torch types.  All functions return 0 on success and a negative snch_status otherwise; the message
 * of the last failure on the calling thread is returned by snch_last_error().  Nothing here ever falls back to a CPU
 * implementation: if CUDA is unavailable the call fails with SNCH_ERR_CUDA.
 *
 * Pointer convention for the *_batch calls: query and result pointers may be DEVICE pointers (zero-copy: the kernels
 * read/write them directly on `stream`) or HOST pointers (the library stages them through its own pinned buffers and
 * copies H2D / D2H on `stream`, then synchronises the stream before returning).  The kind is detected per pointer
 * with cudaPointerGetAttributes; all pointers of one call must be of the same kind.
 */
#ifndef SNCH_B200_H
#define SNCH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNCH_B200_ABI_VERSION 2

typedef struct snch_scene snch_scene; /* opaque; owns one device arena on one GPU */
typedef void *snch_stream;            /* a cudaStream_t (NULL = legacy default stream) */

enum snch_status
{
    SNCH_OK = 0,
    SNCH_ERR_INVALID = -1,   /* bad argument */
    SNCH_ERR_NOT_BUILT = -2, /* "BVH is not built yet."  (scene.cuh:686,1250) */
    SNCH_ERR_CUDA = -3,      /* CUDA runtime failure / no device */
    SNCH_ERR_OOM = -4,
    SNCH_ERR_POINTER_KIND = -5
};

const char *snch_last_error(void);
int snch_abi_version(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Scene lifetime.   Replaces lbvh::scene<3>::scene(vFirst,vLast,iFirst,iLast)            scene.cuh:1128-1133
 *                            lbvh::scene<3>::compute_silhouettes()                        scene.cuh:1167-1204
 *                            lbvh::scene<3>::build_bvh() -> lbvh::bvh<...>::construct()   scene.cuh:1205-1229, bvh.cuh:380-613
 * xyz: n_verts x 3 floats, tri: n_tris x 3 vertex indices (HOST pointers, copied).
 * ------------------------------------------------------------------------------------------------------------- */
int snch_scene3_create(const float *xyz, uint32_t n_verts, const int32_t *tri, uint32_t n_tris, int device, snch_scene **out);
int snch_scene_destroy(snch_scene *s);
/* host-side edge adjacency: first-seen edge ids, silhouette int4, first-owner rule (quirks Q17/Q18 reproduced) */
int snch_scene_compute_silhouettes(snch_scene *s);

typedef struct snch_build_options
{
    uint32_t struct_size;       /* = sizeof(snch_build_options) */
    uint32_t keep_reference_layout; /* 1 (default): keep nodes/aabbs/cones/objects arrays of bvh_device (bvh.cuh:48-54) */
    uint32_t print_collision;   /* 1: print "Morton code collision detected." like bvh.cuh:466 (default 0) */
    uint32_t refit_only;        /* 1: keep Morton order and topology of the previous build and only recompute leaf boxes/cones and
                                 * the bottom-up box + cone merge (after snch_scene_update_vertices; the scene must be built).  The
                                 * reference has no counterpart (it rebuilds); fcpw exposes refit() (ext/fcpw/include/fcpw/fcpw.h:97-106).
                                 * With unchanged vertices the result is bit-identical to a full build. */
} snch_build_options;

/* Morton -> radix sort -> Karras hierarchy -> fused AABB + normal-cone refit -> traversal records; asynchronous on `stream`
 * except for the final stats read-back.  May be called again after snch_scene_update_vertices(). */
int snch_scene_build(snch_scene *s, const snch_build_options *opts, snch_stream stream);

/* New vertex positions for the same topology (n_verts x 3 floats, HOST or DEVICE pointer); takes effect at the next
 * snch_scene_build().  No reference counterpart: lbvh::scene<3> must be re-created (scene.cuh:1128-1133). */
int snch_scene_update_vertices(snch_scene *s, const float *xyz, snch_stream stream);

typedef struct snch_build_stats
{
    uint32_t num_objects, num_nodes, num_edges, num_vertices;
    uint32_t morton_collision;  /* the reference's 64-bit key path would have been taken (bvh.cuh:464) */
    uint32_t q1_nodes;          /* internal nodes whose cone union exceeded pi (half_angle defined as pi) */
    float build_ms;             /* device time of the last snch_scene_build (CUDA events) */
    float adjacency_ms;         /* host time of snch_scene_compute_silhouettes */
    uint64_t arena_bytes;
    float scene_lower[3], scene_upper[3];
} snch_build_stats;
int snch_scene_stats(const snch_scene *s, snch_build_stats *out);

/* ---------------------------------------------------------------------------------------------------------------
 * Reference-layout view.  Replaces lbvh::bvh<...>::get_device_repr() / scene<3>::get_bvh_device_ptr()
 *                         (bvh.cuh:367-378, scene.cuh:1242-1252).  Non-owning DEVICE pointers, valid until the scene is
 *                         rebuilt or destroyed.  Record layouts are the reference's:
 *   nodes   16 B {parent,left,right,object_idx}      aabbs 24 B {upper xyz, lower xyz}     cones 20 B {axis xyz, half_angle, radius}
 *   objects 40 B scene<3>::triangle {int3 v; int3 owned_edges; const float3* vertices; const silhouette_edge* silhouettes}
 *   silhouettes 32 B {int4 indices; const float3* vertices}       vertices 12 B float3
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct snch_bvh_device_pod
{
    uint32_t num_nodes, num_objects;
    void *nodes, *aabbs, *cones, *objects;
    void *vertices, *silhouettes;
    uint32_t num_vertices, num_silhouettes;
} snch_bvh_device_pod;
int snch_scene_device_repr(const snch_scene *s, snch_bvh_device_pod *out);

/* Copy one build product to HOST memory (parity dumps).  `bytes` must equal the product's size. */
enum snch_export_kind
{
    SNCH_EXPORT_NODES = 0,       /* (2N-1) x 4 u32 */
    SNCH_EXPORT_AABBS = 1,       /* (2N-1) x 6 f32 */
    SNCH_EXPORT_CONES = 2,       /* (2N-1) x 5 f32 */
    SNCH_EXPORT_MORTON_SORTED = 3, /* N u32 */
    SNCH_EXPORT_SORTED_INDEX = 4,  /* N u32 */
    SNCH_EXPORT_RANGES = 5,      /* (N-1) x 2 u32 : Karras [first,last] */
    SNCH_EXPORT_EDGES = 6,       /* E x 4 i32 : silhouette_edge::indices */
    SNCH_EXPORT_TRI_EDGES = 7,   /* N x 3 i32 : edge_indices_h */
    SNCH_EXPORT_TRI_OWNED = 8,   /* N x 3 i32 : triangle::silhouette_indices */
    SNCH_EXPORT_Q1_TAINT = 9     /* (2N-1) u8 */
};
int snch_scene_export(const snch_scene *s, int kind, void *host_dst, size_t bytes);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched queries (one launch for n queries).  Each replaces a per-thread reference call made from a user kernel.
 * ------------------------------------------------------------------------------------------------------------- */
/* query_device(bvh, nearest(p), scene<3>::distance_calculator())  -> pair<object idx, distance>     query.cuh:238-318
 * Index rule on exact ties: any member of the argmin set (documented deviation). */
int snch_closest_point_batch(const snch_scene *s, const float *points_xyz, uint64_t n, uint32_t *out_index, float *out_distance,
                             snch_stream stream);

/* query_device(bvh, nearest_silhouette(p, flip), scene<3>::silhouette_distance_calculator()) -> distance   query.cuh:325-423
 * r_max (NULL = unbounded, the reference behaviour): search radius per query; result is +inf when no silhouette point
 * lies within it — identical to filtering the unbounded answer, except that a radius whose square is zero finds
 * nothing (r_max is the reference's max_radius: `min_radius_squared >= max_radius_squared -> false`, scene.cuh:791).  flip: one
 * byte per query, or NULL = false.
 * Optional outputs (NULL = not wanted; asking for neither costs nothing) — what the reference computes and drops:
 *   out_edge      index in scene<3>::silhouettes of an edge attaining the distance ("TODO: identify nearest index",
 *                 query.cuh:386,411); 0xFFFFFFFF when the distance is +inf.  On exact ties any attaining edge.
 *   out_point_xyz the closest point ON that edge — `closest_pos` of silhouette_edge::find_closest_silhouette_point
 *                 (scene.cuh:796-799), which the reference discards at :817-821; (0,0,0) when the distance is +inf. */
int snch_closest_silhouette_batch(const snch_scene *s, const float *points_xyz, const uint8_t *flip, const float *r_max, uint64_t n,
                                  float *out_distance, uint32_t *out_edge, float *out_point_xyz, snch_stream stream);

/* query_device(bvh, ray_intersect<TestOnly>(ray, max_dist), scene<3>::intersect_test())                     query.cuh:79-169
 * -> tuple<found, t, uv, prim idx>.  any_hit != 0 is TestOnly: only `found` is meaningful. */
typedef struct snch_hit
{
    float t;          /* +inf when not found */
    float u, v;       /* barycentrics of the hit (0 when not found) */
    uint32_t prim;    /* 0xFFFFFFFF when not found */
} snch_hit;
int snch_intersect_batch(const snch_scene *s, const float *origins_xyz, const float *dirs_xyz, const float *t_max, uint64_t n,
                         snch_hit *out_hits, uint8_t *out_found, int any_hit, snch_stream stream);

/* sample_object_in_sphere(bvh, sphere_intersect(sph), intersect_sphere(), measurement_getter(), green_weight(), u)   sample.cuh:23-92
 * followed by sample_on_object(bvh, idx, scene<3>::sample_on_object(), u1, u2)                               sample.cuh:7-21
 * spheres: x,y,z,radius; rnd: 3 uniforms per query {u (branch), u1, u2 (point)}.  out_index = -1 and pdf = 0 on a miss
 * (pdf is uninitialised in the reference).  out_point may be NULL. */
int snch_sample_in_sphere_batch(const snch_scene *s, const float *spheres_xyzr, const float *rnd_uvw, uint64_t n, int32_t *out_index,
                                float *out_pdf, float *out_point_xyz, snch_stream stream);

/* One wavefront walk-on-stars step per walker — the call sequence of an Elaina-style WoSt stage (README.md:5,9), as one
 * launch sequence that shares the Morton ordering of the walkers and keeps the star radius on the device:
 *   (closest_index, closest_distance) = query_device(bvh, nearest(p), distance_calculator())                 query.cuh:238-318
 *   silhouette_distance = query_device(bvh, nearest_silhouette(p, flip), ...) restricted to r_max = closest_distance   query.cuh:325-423
 *   star_radius = min(closest_distance, silhouette_distance)
 *   hits/found  = query_device(bvh, ray_intersect(ray(p, dir), star_radius), intersect_test())              query.cuh:79-169
 *   sample_*    = sample_object_in_sphere(bvh, sphere_intersect(sphere(p, star_radius)), ...) + sample_on_object   sample.cuh:7-92
 * Every output equals what the four *_batch calls return for the same inputs.  Required: points_xyz; every other pointer may be
 * NULL (dirs NULL skips the ray stage, rnd NULL the sampling stage; NULL outputs are not written).  All HOST or all DEVICE pointers. */
typedef struct snch_wost_io
{
    uint32_t struct_size; /* = sizeof(snch_wost_io) */
    uint32_t reserved;
    const float *points_xyz;  /* n x 3 */
    const uint8_t *flip;      /* n, or NULL = false */
    const float *dirs_xyz;    /* n x 3 */
    const float *rnd_uvw;     /* n x 3 */
    uint32_t *closest_index;
    float *closest_distance;
    float *silhouette_distance;
    float *star_radius;
    snch_hit *hits;
    uint8_t *found;
    int32_t *sample_index;
    float *sample_pdf;
    float *sample_point_xyz;
    uint32_t *silhouette_edge;    /* as out_edge / out_point_xyz of snch_closest_silhouette_batch */
    float *silhouette_point_xyz;
} snch_wost_io;
int snch_wost_step_batch(const snch_scene *s, const snch_wost_io *io, uint64_t n, snch_stream stream);

/* Scheduling knobs of the batched kernels; results never depend on them (tests/test_gpu_queries.py sweeps them).
 *   "query.sort_min_n"  batches at least this large are visited in Morton order of the query points (default 16384; 0 = never)
 *   "query.sort_bits"   key bits of that ordering (8..30, default 24)
 *   "query.sort_rays"   ray batches: -1 = by size (default: Morton order of the origins for 4M rays or more on 2M triangles or more, caller's order otherwise),
 *                       0 = caller's order, 1 = Morton order of the origins, 2 = direction octant, then origin
 *   "query.cone_filter" silhouette normal-cone test: 0 = cone.cuh:168-212 verbatim; 1 = guard-banded sine-space evaluation on MUFU
 *                       approximations, the verbatim chain out of line for everything inside the band (default)
 *   "query.seed"        closest point: bit 0 = bound each query by the triangle that answered the lane's previous query (default 1);
 *                       bit 1 = switch the per-triangle lower bound off
 *   "query.sil_flush"   silhouette: queued leaves of a warp that trigger their (one per lane) tests (1..32, default 24)
 *   "query.sil_chunk"   silhouette: queries a warp draws per atomic (0 = 64, or 16 below 12M queries)
 *   "query.sil_tail"    silhouette: when a batch has been handed out, a warp with at most this many walking lanes passes them to a
 *                       one-query-per-warp finishing launch (default 8; 0 = never)
 *   "query.sort_radius" bounded silhouette batches: 0 = Morton order only, 1 = search-radius octave then Morton, 2 = the same with
 *                       the largest radii first (default 2: the longest walks start first)
 *   "query.wide_max_n"  closest point: batches smaller than this — or with fewer than 2 queries per triangle — are walked one query
 *                       per warp (default 2097152; 0 = never)
 *   "query.wide_max_n_sil"  the same for silhouette batches (default 262144)
 *   "query.ray_kernel"  1 = reference-order ray walk with parked leaves (default; batches under 1M rays test a parked leaf at once),
 *                       2 = the same with the large-batch setting for every batch, 0 = leaves tested where they are met (same hit
 *                       flag and t; among triangles hit at the same t it keeps the first in its own order, not the reference's)
 *   "query.ray_flush" / "query.ray_refill"  parked / idle lanes of a warp that trigger the triangle tests / the next draw (8 / 8)
 *   "query.host_chunk"  host-pointer batches: queries per pipeline chunk (default 8388608; 0 = one chunk)
 *   "query.host_first"  host-pointer batches that are split: queries of the first chunk, whose H2D copy nothing overlaps
 *                       (default 0 = an eighth of the batch within [512K, 2M]; -1 = like the other chunks)
 *   "query.host_split_min"  host-pointer batches of at least this many queries are split even when they fit one chunk
 *                       (default 3145728; 0 = only batches larger than "query.host_chunk")
 *   "query.blocks_per_sm" cap on resident CTAs per SM of the persistent kernels (default 0 = occupancy limit)
 *   "build.refit_kernel" 1 = CTA-cooperative refit (default), 0 = per-thread climb; "sort.onesweep" 1 = onesweep radix sort
 *                       (default), 0 = three-kernel passes; "sort.lookback" predecessor tiles a tile reads per round trip of its
 *                       look-back (8, default) or 1; "adjacency.device" 1 = GPU silhouette adjacency (default when a device
 *                       is present), 0 = host passes
 * The reference has no counterpart (its queries are per-thread device functions scheduled by the caller's kernel).
 *   "query.time_kernels" bracket every traversal kernel with CUDA events on the launching stream (default 0; see snch_scene_counter) */
int snch_scene_set_option(snch_scene *s, const char *name, int64_t value);

/* Scene preparation:
 *   "adjacency.device"  1 = snch_scene_compute_silhouettes runs on the GPU (half-edge sort, adjacency.cu), 0 = host passes
 *                       (same arrays bit for bit), -1 (default) = GPU when a CUDA device is present */
/* Launch accounting since creation / the last reset (what bench.py reports as gpu_launches and roofline.achieved):
 *   "query.launches"            kernels launched by the *_batch calls (ordering + traversal)
 *   "query.traversal_launches"  traversal kernels among them
 *   "query.traversal_ms"        device time of the traversal kernels alone; needs "query.time_kernels" = 1 (synchronises the last one)
 *   "build.launches"            kernels of the last snch_scene_build
 *   "adjacency.device_ms"       device time (upload + kernels, CUDA events) of the last GPU snch_scene_compute_silhouettes;
 *                               snch_build_stats.adjacency_ms is the host's wall clock around it, allocations included */
int snch_scene_counter(snch_scene *s, const char *name, double *value, int reset);
/* name of the traversal kernel the last *_batch call on this scene launched (static string; what bench.py puts in roofline.kernel) */
const char *snch_scene_last_kernel(const snch_scene *s);

/* ---------------------------------------------------------------------------------------------------------------
 * Generic builder.  Replaces lbvh::bvh<Real,dim,Object,AABBGetter,ConeGetter,MortonCalc>::construct()   bvh.cuh:380-613
 * for ANY object type: the caller (include/snch_lbvh/core/bvh.cuh) evaluates its getters per object and passes the leaf
 * arrays; this runs scene box -> Morton (default_morton_code_calculator, bvh.cuh:232-304, unless morton_codes is given)
 * -> stable sort -> Karras hierarchy on (code << 32 | index) -> one bottom-up box + cone refit.
 * All pointers are DEVICE pointers.  dim = 2 or 3, float.  Record layouts are the reference's:
 *   leaf_aabbs n x {upper[dim], lower[dim]}   leaf_cones n x {axis[dim], half_angle, radius}     (object order)
 *   nodes (2n-1) x {parent,left,right,object}   aabbs / cones (2n-1) records, leaves at n-1.. in Morton order
 * Optional outputs: sorted_index_out[n], morton_sorted_out[n] (device), *collision_out (host; forces a stream sync). */
int snch_lbvh_build(int dim, uint32_t n, const void *leaf_aabbs, const void *leaf_cones, const uint32_t *morton_codes, void *nodes,
                    void *aabbs, void *cones, uint32_t *sorted_index_out, uint32_t *morton_sorted_out, int *collision_out,
                    snch_stream stream);

/* ---------------------------------------------------------------------------------------------------------------
 * 2-D scenes (polylines).  Replaces lbvh::scene<2> (scene.cuh:287-703): scene<2>::scene(vFirst,vLast,iFirst,iLast) :621-627,
 * compute_silhouettes() :629-656 (one silhouette_vertex {previous, self, next} per vertex), build_bvh() :658-681 (a vertex is
 * owned by the first segment in input order that touches it) -> bvh::construct(), and the per-thread calls a user kernel makes
 * on its bvh_device<float, 2, line_segment>:
 *   query_device(bvh, nearest(p), scene<2>::distance_calculator())                                       query.cuh:238-318
 *   query_device(bvh, nearest_silhouette(p, flip), scene<2>::silhouette_distance_calculator())           query.cuh:325-423
 *   query_device(bvh, ray_intersect<TestOnly>(ray, max_dist), scene<2>::intersect_test())                query.cuh:79-169
 *   sample_object_in_sphere(bvh, sphere_intersect(circle), scene<2>::intersect_sphere(), measurement_getter(), green_weight(), u)
 *   + sample_on_object(bvh, idx, scene<2>::sample_on_object(), u1, .)                                    sample.cuh:7-92
 * xy: n_verts x 2 floats, seg: n_segs x 2 vertex indices (HOST pointers, copied).  Query / result pointers are all DEVICE or
 * all HOST pointers (staged inside the call, which then synchronises `stream`).  Sentinels as in 3-D.
 * snch_scene2_device_repr fills the reference-layout view: nodes 16 B, aabbs 16 B {upper xy, lower xy}, cones 16 B {axis xy,
 * half_angle, radius}, objects 32 B scene<2>::line_segment, silhouettes 32 B scene<2>::silhouette_vertex, vertices float2.
 * snch_scene2_export kinds: NODES, AABBS (4 f32), CONES (4 f32), MORTON_SORTED, SORTED_INDEX, EDGES (= silhouette_vertex::indices,
 * n_verts x 4 i32), TRI_OWNED (= line_segment::silhouette_indices, n_segs x 2 i32).
 * snch_hit for a 2-D ray: t, u = segment parameter s, v = 0, prim.   Circles: x, y, radius.  rnd2: {u (branch), u1 (point)}.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct snch_scene2 snch_scene2;
int snch_scene2_create(const float *xy, uint32_t n_verts, const int32_t *seg, uint32_t n_segs, int device, snch_scene2 **out);
int snch_scene2_destroy(snch_scene2 *s);
int snch_scene2_compute_silhouettes(snch_scene2 *s);
int snch_scene2_build(snch_scene2 *s, snch_stream stream);
int snch_scene2_stats(const snch_scene2 *s, snch_build_stats *out);
int snch_scene2_device_repr(const snch_scene2 *s, snch_bvh_device_pod *out);
int snch_scene2_export(const snch_scene2 *s, int kind, void *host_dst, uint64_t bytes);
int snch_scene2_set_option(snch_scene2 *s, const char *name, int64_t value); /* "query.sort_min_n", "query.sort_bits", "query.sort_rays", "query.blocks_per_sm", "query.wide_max_n" (batches below it walk one query per warp; default 131072) */
int snch_closest_point_batch2(const snch_scene2 *s, const float *points_xy, uint64_t n, uint32_t *out_index, float *out_distance,
                              snch_stream stream);
/* out_vertex / out_point_xy (optional, NULL = not wanted): index in scene<2>::silhouettes (= the vertex index) of a silhouette
 * vertex attaining the distance and that vertex's position — `p` of silhouette_vertex::find_closest_silhouette_point
 * (scene.cuh:354-356); 0xFFFFFFFF / (0,0) when the distance is +inf */
int snch_closest_silhouette_batch2(const snch_scene2 *s, const float *points_xy, const uint8_t *flip, const float *r_max, uint64_t n,
                                   float *out_distance, uint32_t *out_vertex, float *out_point_xy, snch_stream stream);
int snch_intersect_batch2(const snch_scene2 *s, const float *origins_xy, const float *dirs_xy, const float *t_max, uint64_t n,
                          snch_hit *out_hits, uint8_t *out_found, int any_hit, snch_stream stream);
int snch_sample_in_sphere_batch2(const snch_scene2 *s, const float *circles_xyr, const float *rnd2, uint64_t n, int32_t *out_index,
                                 float *out_pdf, float *out_point_xy, snch_stream stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Replication (multi-GPU): the built scene lives in ONE pointer-free arena; a replica is a byte copy of it plus a
 * pointer patch.  The reference is single-GPU (no counterpart); this is the exchange step of its north-star deployment —
 * build on one GPU, broadcast over NVLink, every GPU traverses its shard of the query batch against its own copy.
 *
 *   one process per GPU  : snch_comm_* + snch_scene_broadcast — ncclBroadcast received directly into the replica's arena.
 *                          NCCL is libnccl.so.2 itself, resolved on first use (SNCH_NCCL_LIB overrides the path); nothing else
 *                          in this library needs it.  The 128-byte id from snch_comm_unique_id (rank 0) reaches the other ranks by
 *                          whatever bootstrap the job has (MPI, a file, torch.distributed, ...); a job that already owns an
 *                          ncclComm_t wraps it with snch_comm_adopt instead.
 *   one process, n GPUs  : snch_scene_replicate_local — cudaMemcpyPeerAsync fan-out, peer access enabled where possible.
 *   bring your own bytes : snch_scene_arena + snch_scene_adopt_arena.
 * A replica answers every query and export like the original; it cannot be rebuilt (rebuild the original, then
 * snch_scene_rebroadcast into the replicas' existing arenas: the warm form of the exchange).
 * ------------------------------------------------------------------------------------------------------------- */
#define SNCH_COMM_ID_BYTES 128
typedef struct snch_comm snch_comm;
int snch_comm_unique_id(void *id_out, uint64_t bytes);                                   /* ncclGetUniqueId */
int snch_comm_create(const void *id, uint64_t bytes, int rank, int world, int device, snch_comm **out); /* ncclCommInitRank (collective) */
int snch_comm_adopt(void *nccl_comm, int rank, int world, int device, snch_comm **out);  /* wrap the caller's ncclComm_t (not destroyed here) */
int snch_comm_destroy(snch_comm *c);
/* collective over `comm`, asynchronous on `stream` until the final synchronisation.  Root: `scene` = the built scene, *out = NULL.
 * Other ranks: `scene` = NULL, *out = the replica on comm's device. */
int snch_scene_broadcast(const snch_scene *scene, int root, snch_comm *comm, snch_stream stream, snch_scene **out);
/* collective: the root's (rebuilt) arena into the existing arenas of its replicas (same vertex / triangle / edge counts) */
int snch_scene_rebroadcast(snch_scene *scene, int root, snch_comm *comm, snch_stream stream);
/* replicas of `scene` on devices[0..n) of this process (out[i] on devices[i]; a device may be the scene's own) */
int snch_scene_replicate_local(const snch_scene *scene, const int *devices, int n, snch_scene **out);
int snch_scene_arena(const snch_scene *s, void **device_ptr, uint64_t *bytes);
/* arena_copy: DEVICE pointer on `device` holding a byte-exact copy of another scene's arena (copied; caller keeps ownership) */
int snch_scene_adopt_arena(const void *arena_copy, uint64_t bytes, int device, snch_stream stream, snch_scene **out);

/* Serialisation: the arena written to / read from a file verbatim (header with magic, version and size
 * first).  A loaded scene answers queries and exports like the one that was saved; it cannot be rebuilt (like an adopted one).
 * No reference counterpart: its scene embeds raw device pointers (scene.cuh:831-839). */
int snch_scene_save(const snch_scene *s, const char *path);
int snch_scene_load(const char *path, int device, snch_stream stream, snch_scene **out);

/* Self-test of the device math: out[i] = f(x[i]) evaluated ON THE DEVICE by the routine the build uses where the reference's
 * CPU build calls its libm (include/snch_lbvh/core/host_libm.cuh): which = 0 acosf, 1 sinf, 2 cosf, 3 logf.  Host pointers.
 * The cone refit (core/cone.cuh:34-66, 288-302, 427-480) and scene<2>::green_weight (scene.cuh:606-613) are bit-identical to
 * the reference's CPU build because these are; tests/test_gpu_host_libm.py compares them with the host's libm. */
int snch_selftest_host_libm(int which, const float *x, uint64_t n, float *out, int device);

#ifdef __cplusplus
}
#endif
#endif /* SNCH_B200_H */
