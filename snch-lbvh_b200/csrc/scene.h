// scene.h — internal definition of the opaque snch_scene handle and the entry points shared by the translation units.
#pragma once
#include "../../include/snch_b200.h"
#include "layout.h"

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

namespace snch
{
// Scheduling knobs of the batched query kernels (snch_scene_set_option).  Results never depend on them.
struct QueryTuning
{
    int sort_min_n = 16384; // batches at least this large are visited in Morton order of the query points (0 = never)
    int sort_bits = 24;     // Morton key bits the ordering sorts on (top bits of the 30-bit code)
    int sort_radius = 2;    // bounded silhouette batches: 1 = order by search-radius octave first, then Morton code;
                            // 2 = the same with the largest radii first (longest walks start first, the tail is made of cheap queries);
                            // 3, 4 = largest first with 2 / 4 classes per octave
    int sort_rays = -1;     // ray batches: -1 = Morton order of the origins when the tree has 2M triangles or more (its records exceed L2) and the
                            // batch 4M rays or more, caller's order otherwise (at 1M triangles ordering costs more than it returns); 0 = caller's order,
                            // 1 = Morton order of the origins, 2 = direction octant, then origin
    int cone_filter = 1;    // silhouette normal-cone test: 0 = the reference's libm chain verbatim, 1 = guard-banded sine-space filter on
                            // MUFU approximations with the exact chain out of line (decisions identical; 52.3 vs 69 ms on C3)
    int sil_flush = 24;     // silhouette: queued leaves of a warp that trigger their tests (1..32; fewer = bounds tighten sooner, tests run on fewer lanes)
    int sil_chunk = 0;      // silhouette: queries a warp draws per atomic (0 = 64 for batches of 12M and more, 16 below)
    int sil_tail = 8;       // silhouette: once the batch is handed out, a warp with at most this many walking lanes finishes them
                            // cooperatively, one query at a time on 32 lanes (0 = never)
    int wide_max_n = 2097152; // closest point: batches smaller than this walk ONE query per warp (32 lanes on one query: shortens the critical
                            // path of pathological queries in batches too small to fill the machine; 0 = never)
    int wide_max_n_sil = 262144; // silhouette: the same for k_silhouette_wide (measured crossover 0.25-0.5M queries)
    int seed = 1;           // closest point: bit 0 = bound each query by the triangle that answered the lane's previous query;
                            // bit 1 = switch the per-triangle lower bound OFF (A/B)
    int ray_kernel = 1;     // ray traversal: 1 = reference-order walk with parked leaves (k_intersect_parked; batches under 1M rays test a parked
                            // leaf at once), 2 = the same with the large-batch flush for every batch, 0 = leaves tested where they are met
                            // (k_intersect: same hit flag and t; among triangles hit at the same t its own order picks, not the reference's)
    int ray_flush = 8;      // k_intersect_parked: parked lanes of a warp that trigger the triangle tests
    int ray_refill = 8;
    int ray_prefetch = 1;   // k_intersect_parked: ask L2 for the record of a child when it is pushed     // k_intersect_parked: idle lanes of a warp that trigger the next draw of rays
    int blocks_per_sm = 0;  // cap on resident CTAs per SM of the persistent kernels (0 = occupancy limit)
    int host_first = 0;       // host-pointer batches that are split: queries of the FIRST chunk (nothing overlaps its H2D copy);
                              // 0 = an eighth of the batch within [512K, 2M], -1 = like the other chunks
    int host_split_min = 3 << 20; // host-pointer batches of at least this many queries are split (short first chunk + the rest) even
                              // when they fit one chunk; 0 = only batches larger than "query.host_chunk" are split
    int host_chunk = 1 << 23; // host-pointer batches: queries per pipeline chunk (H2D / kernels / D2H overlap); 0 = one chunk.
                              // Measured on C3 (16.7M queries): 0 -> 60.5 ms, 8M -> 59.5, 4M -> 60.0, 2M -> 63.5, 1M -> 73.1: every extra
                              // launch pays its own tail and orders a sparser batch, so chunks stay large
};
// Launch accounting of the batched queries (snch_scene_counter): kernels launched, traversal kernels among them, the name of
// the last traversal kernel, and — when "query.time_kernels" is set — device time of the traversal kernels alone (one CUDA
// event pair per launch on the launching stream, folded on demand: nothing in the launch path waits for an event).  Safe
// for concurrent batches from several host threads / streams.
struct QueryCounters
{
    std::atomic<uint64_t> launches{0}, traversal_launches{0};
    std::atomic<int> time_kernels{0};
    std::atomic<const char *> last_kernel{""};
    std::mutex mu; // guards everything below
    double traversal_ms = 0.0;
    struct Pair
    {
        cudaEvent_t a, b;
    };
    std::vector<Pair> pending, spare;
    bool begin(Pair &p, cudaStream_t st);
    void end(const Pair &p, cudaStream_t st);
    void fold();
    void reset();
    void release();
};
} // namespace snch

struct snch_scene
{
    int device = 0;
    uint32_t n_verts = 0, n_tris = 0, n_edges = 0;
    // host copies of the input and of the adjacency (edge ids, silhouette int4, ownership)
    std::vector<float> h_xyz;
    std::vector<int32_t> h_tri;
    std::vector<int32_t> h_edges4, h_tri_edges, h_tri_owned;
    bool silhouettes_done = false, built = false, adopted = false;
    // device-side adjacency (adjacency.cu): one allocation holding tri | tri_edges | tri_owned | edges4 until the first
    // build assembles the reference-layout structs from it; then the arena is the only copy
    int adjacency_mode = -1; // "adjacency.device": 1 = GPU, 0 = host passes, -1 = GPU when a CUDA device is present
    bool adjacency_on_device = false, arena_has_topology = false;
    unsigned char *adj = nullptr;
    int32_t *adj_tri = nullptr, *adj_tri_edges = nullptr, *adj_tri_owned = nullptr;
    int4 *adj_edges4 = nullptr;
    // one device arena
    unsigned char *arena = nullptr;
    uint64_t arena_bytes = 0;
    snch::ArenaHeader hdr{};
    snch::SceneView view{};
    // build scratch (kept between builds)
    unsigned char *scratch = nullptr;
    uint64_t scratch_bytes = 0;
    // stream-ordered pool for per-call scratch (query ordering, work counters, staging of host-pointer batches)
    cudaMemPool_t pool = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr, compute_b = nullptr; // streams of the host-pointer pipeline (created on first use)
    std::mutex mu;
    snch::QueryTuning tuning;
    snch::QueryCounters counters;
    uint64_t build_launches = 0;
    // stats
    float build_ms = 0.f, adjacency_ms = 0.f, adjacency_device_ms = 0.f; // adjacency: host wall clock / CUDA events around the device work
    uint32_t opt_print_collision = 0, opt_refit_only = 0;
    int opt_refit_kernel = 1; // "build.refit_kernel": 1 = block-cooperative rounds (default), 0 = one climbing thread per leaf
};

namespace snch
{
void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
#define SNCH_CUDA(call)                                              \
    do                                                               \
    {                                                                \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return snch::cuda_fail(e__, #call);  \
    } while (0)

// capi.cu — replica creation: validate the header + allocate on `device`; (caller fills the arena); patch pointers
int adopt_begin(const ArenaHeader &h, uint64_t bytes, int device, const char *who, snch_scene **out);
int adopt_end(snch_scene *s, cudaStream_t stream);

// adjacency.cu
int compute_adjacency_device(snch_scene *s);
int fetch_adjacency_host(snch_scene *s);
void free_adjacency(snch_scene *s);

// build.cu
void compute_adjacency_host(snch_scene *s);
int build_device(snch_scene *s, cudaStream_t stream);
void layout_arena(ArenaHeader &h, uint32_t nV, uint32_t nT, uint32_t nE);
void resolve_view(snch_scene *s);
int patch_pointers(snch_scene *s, cudaStream_t stream);

// query.cu — n <= 2^32 - 2^20 per launch (the C-ABI splits larger batches); `scratch` has query_scratch_bytes(n) bytes
uint64_t query_scratch_bytes(uint64_t n, const QueryTuning &t);
int launch_closest(const SceneView &v, const QueryTuning &t, const float *q, uint64_t n, uint32_t *idx, float *dist, unsigned char *scratch,
                   cudaStream_t st, QueryCounters *qc);
int launch_silhouette(const SceneView &v, const QueryTuning &t, const float *q, const uint8_t *flip, const float *rmax, uint64_t n,
                      float *dist, uint32_t *edge, float *point, unsigned char *scratch, cudaStream_t st, QueryCounters *qc);
int launch_intersect(const SceneView &v, const QueryTuning &t, const float *o, const float *d, const float *tmax, uint64_t n, snch_hit *hits,
                     uint8_t *found, int any_hit, unsigned char *scratch, cudaStream_t st, QueryCounters *qc);
int launch_sample(const SceneView &v, const QueryTuning &t, const float *sph, const float *rnd, uint64_t n, int32_t *idx, float *pdf,
                  float *pt, unsigned char *scratch, cudaStream_t st, QueryCounters *qc);
// one wavefront walk-on-stars step over device buffers (snch_wost_step_batch); scratch has wost_scratch_bytes(n) bytes
struct WostBuffers
{
    const float *points;
    const uint8_t *flip;
    const float *dirs, *rnd;
    uint32_t *closest_index;
    float *closest_distance, *silhouette_distance, *star_radius;
    uint32_t *silhouette_edge;
    float *silhouette_point;
    snch_hit *hits;
    uint8_t *found;
    int32_t *sample_index;
    float *sample_pdf, *sample_point;
};
uint64_t wost_scratch_bytes(uint64_t n, const QueryTuning &t);
int launch_wost_step(const SceneView &v, const QueryTuning &t, const WostBuffers &io, uint64_t n, unsigned char *scratch, cudaStream_t st,
                     QueryCounters *qc);
} // namespace snch
