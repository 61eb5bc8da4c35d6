// scene.h — internal definition of the opaque snch_scene handle and the entry points shared by the translation units.
#pragma once
#include "../../include/snch_b200.h"
#include "layout.h"

#include <string>
#include <vector>

struct snch_scene
{
    int device = 0;
    uint32_t n_verts = 0, n_tris = 0, n_edges = 0;
    // host copies of the input and of the adjacency (edge ids, silhouette int4, ownership)
    std::vector<float> h_xyz;
    std::vector<int32_t> h_tri;
    std::vector<int32_t> h_edges4, h_tri_edges, h_tri_owned;
    bool silhouettes_done = false, built = false, adopted = false;
    // one device arena
    unsigned char *arena = nullptr;
    uint64_t arena_bytes = 0;
    snch::ArenaHeader hdr{};
    snch::SceneView view{};
    // build scratch (kept between builds)
    unsigned char *scratch = nullptr;
    uint64_t scratch_bytes = 0;
    // staging for host-pointer batches
    unsigned char *pinned = nullptr;
    uint64_t pinned_bytes = 0;
    unsigned char *dstage = nullptr;
    uint64_t dstage_bytes = 0;
    // stats
    float build_ms = 0.f, adjacency_ms = 0.f;
    uint32_t opt_print_collision = 0;
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

// build.cu
void compute_adjacency_host(snch_scene *s);
int build_device(snch_scene *s, cudaStream_t stream);
void resolve_view(snch_scene *s);
int patch_pointers(snch_scene *s, cudaStream_t stream);

// query.cu
int launch_closest(const SceneView &v, const float *q, uint64_t n, uint32_t *idx, float *dist, cudaStream_t st);
int launch_silhouette(const SceneView &v, const float *q, const uint8_t *flip, const float *rmax, uint64_t n, float *dist, cudaStream_t st);
int launch_intersect(const SceneView &v, const float *o, const float *d, const float *tmax, uint64_t n, snch_hit *hits, uint8_t *found,
                     int any_hit, cudaStream_t st);
int launch_sample(const SceneView &v, const float *sph, const float *rnd, uint64_t n, int32_t *idx, float *pdf, float *pt, cudaStream_t st);
} // namespace snch
