// adjacency.cu — silhouette-edge adjacency of a triangle mesh on the GPU.
//
// Replaces the host passes of the reference, scene<3>::assign_edge_indices / compute_silhouettes / the ownership loop of
// build_bvh (scene.cuh:1135-1229: std::map + unordered_map, serial), and produces bit-identical products:
//   * edge ids in FIRST-SEEN order of the half-edges h = 3*triangle + slot                         scene.cuh:1135-1165
//   * silhouette int4 {opposite vertex of the last +oriented face, v_lo, v_hi, opposite of the last -oriented face};
//     non-manifold input: the last writer in half-edge order wins, as on the host (quirk Q18)      scene.cuh:1167-1204
//   * ownership: the triangle holding the first-seen half-edge of an edge owns it, packed in slot order (quirk Q17)
//                                                                                                  scene.cuh:1207-1225
// Method: the 3N half-edges are stably sorted by (v_lo, v_hi) with two LSD radix sorts of (key, h) — the build's sort —
// so every edge becomes one run whose members are in half-edge order.  The run head is the first-seen half-edge; a 0/1
// flag per half-edge scanned in half-edge order numbers the edges in first-seen order without a second sort.
#include "scene.h"
#include "sort_scan.cuh"

#include <chrono>

namespace snch
{
namespace
{
__device__ __forceinline__ void half_edge(const int32_t *__restrict__ tri, uint32_t h, int32_t &lo, int32_t &hi)
{
    const uint32_t i = h / 3u, j = h - 3u * i;
    const int32_t a = tri[3u * i + j], b = tri[3u * i + (j == 2u ? 0u : j + 1u)];
    lo = a < b ? a : b;
    hi = a < b ? b : a;
}

__global__ void k_halfedge_hi(const int32_t *__restrict__ tri, uint32_t m, uint32_t *__restrict__ key, uint32_t *__restrict__ val)
{
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= m) return;
    int32_t lo, hi;
    half_edge(tri, h, lo, hi);
    key[h] = (uint32_t)hi;
    val[h] = h;
}

__global__ void k_halfedge_lo(const int32_t *__restrict__ tri, uint32_t m, const uint32_t *__restrict__ val, uint32_t *__restrict__ key)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int32_t lo, hi;
    half_edge(tri, val[k], lo, hi);
    key[k] = (uint32_t)lo;
}

// head[h] = 1 when half-edge h opens a run of equal (lo, hi), i.e. it is the first triangle slot that saw this edge
__global__ void k_mark_heads(const int32_t *__restrict__ tri, uint32_t m, const uint32_t *__restrict__ val, uint32_t *__restrict__ head)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint32_t h = val[k];
    int32_t lo, hi;
    half_edge(tri, h, lo, hi);
    bool first = k == 0;
    if (!first)
    {
        int32_t plo, phi;
        half_edge(tri, val[k - 1], plo, phi);
        first = plo != lo || phi != hi;
    }
    head[h] = first ? 1u : 0u;
}

// one thread per run: writes the edge record and the edge id of every member half-edge
__global__ void k_fill_edges(const int32_t *__restrict__ tri, uint32_t m, const uint32_t *__restrict__ val, const uint32_t *__restrict__ head,
                             const uint32_t *__restrict__ edge_id, int4 *__restrict__ edges4, int32_t *__restrict__ tri_edges)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint32_t h0 = val[k];
    if (!head[h0]) return;
    const int32_t e = (int32_t)edge_id[h0];
    int32_t lo, hi;
    half_edge(tri, h0, lo, hi);
    int32_t opp_pos = -1, opp_neg = -1;
    for (uint32_t kk = k; kk < m; ++kk)
    {
        const uint32_t h = val[kk];
        if (kk != k && head[h]) break;
        const uint32_t i = h / 3u, j = h - 3u * i;
        const int32_t a = tri[3u * i + j], b = tri[3u * i + (j == 2u ? 0u : j + 1u)];
        const int32_t opp = tri[3u * i + (j == 0u ? 2u : j - 1u)];
        if (a > b) opp_neg = opp; // the face runs the edge from its larger to its smaller vertex: orientation -1
        else opp_pos = opp;
        tri_edges[h] = e;
    }
    edges4[e] = make_int4(opp_pos, lo, hi, opp_neg);
}

__global__ void k_owned(uint32_t n, const uint32_t *__restrict__ head, const uint32_t *__restrict__ edge_id, int32_t *__restrict__ owned)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t o[3] = {-1, -1, -1};
    int p = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (head[3u * i + j]) o[p++] = (int32_t)edge_id[3u * i + j];
    owned[3u * i] = o[0];
    owned[3u * i + 1] = o[1];
    owned[3u * i + 2] = o[2];
}
} // namespace

void free_adjacency(snch_scene *s)
{
    if (s->adj)
    {
        cudaSetDevice(s->device);
        cudaFree(s->adj);
    }
    s->adj = nullptr;
    s->adj_tri = s->adj_tri_edges = s->adj_tri_owned = nullptr;
    s->adj_edges4 = nullptr;
}

// Leaves tri / tri_edges / tri_owned / edges4 in one device allocation owned by the scene (s->adj); build_device()
// assembles the reference-layout structs from it and releases it.  One host synchronisation, at the end.
int compute_adjacency_device(snch_scene *s)
{
    const auto t0 = std::chrono::steady_clock::now();
    SNCH_CUDA(cudaSetDevice(s->device));
    free_adjacency(s);
    s->h_edges4.clear();
    s->h_tri_edges.clear();
    s->h_tri_owned.clear();
    const uint32_t n = s->n_tris;
    const uint64_t m64 = (uint64_t)3 * n;
    if (n == 0)
    {
        s->n_edges = 0;
        s->silhouettes_done = true;
        s->adjacency_on_device = true;
        s->adjacency_ms = 0.f;
        return SNCH_OK;
    }
    const uint32_t m = (uint32_t)m64;
    cudaStream_t st = nullptr;

    // persistent part: tri | tri_edges | tri_owned   (edges4 follows once the edge count is known)
    // transient part: key, val, key_tmp, val_tmp, head, edge_id, sort/scan scratch
    const uint64_t sort_elems = sort_scratch_elems(m), scan_elems = scan_scratch_elems(m);
    uint64_t off = 0;
    auto take = [&](uint64_t bytes)
    {
        const uint64_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    const uint64_t o_key = take(m64 * 4), o_val = take(m64 * 4), o_ktmp = take(m64 * 4), o_vtmp = take(m64 * 4);
    const uint64_t o_head = take(m64 * 4), o_eid = take(m64 * 4), o_sort = take(sort_elems * 4), o_scan = take(scan_elems * 4);
    const uint64_t tmp_bytes = off;
    // the persistent arrays are sized for the worst case (every half-edge its own edge) so that no allocation falls between
    // the kernels; build_device() releases them as soon as the arena holds the topology
    uint64_t poff = 0;
    auto ptake = [&](uint64_t bytes)
    {
        const uint64_t o = poff;
        poff = align_up(poff + bytes, 256);
        return o;
    };
    const uint64_t p_tri = ptake(m64 * 4), p_te = ptake(m64 * 4), p_to = ptake(m64 * 4), p_e4 = ptake(m64 * 16);
    unsigned char *tmp = nullptr, *tri_buf = nullptr, *adj = nullptr;
    if (cudaMalloc(&tmp, tmp_bytes) != cudaSuccess || cudaMalloc(&tri_buf, m64 * 4) != cudaSuccess || cudaMalloc(&adj, poff) != cudaSuccess)
    {
        cudaGetLastError();
        if (tmp) cudaFree(tmp);
        if (tri_buf) cudaFree(tri_buf);
        set_error("cudaMalloc of the adjacency arrays failed");
        return SNCH_ERR_OOM;
    }
    auto fail = [&](cudaError_t e, const char *what)
    {
        cudaFree(tmp);
        cudaFree(tri_buf);
        cudaFree(adj);
        return cuda_fail(e, what);
    };
#define ADJ_CUDA(call)                                   \
    do                                                   \
    {                                                    \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return fail(e__, #call); \
    } while (0)
    int32_t *tri = (int32_t *)tri_buf;
    uint32_t *key = (uint32_t *)(tmp + o_key), *val = (uint32_t *)(tmp + o_val);
    uint32_t *ktmp = (uint32_t *)(tmp + o_ktmp), *vtmp = (uint32_t *)(tmp + o_vtmp);
    uint32_t *head = (uint32_t *)(tmp + o_head), *eid = (uint32_t *)(tmp + o_eid);
    uint32_t *sort_scr = (uint32_t *)(tmp + o_sort), *scan_scr = (uint32_t *)(tmp + o_scan);

    // device-side time of the pass (upload, two sorts, scan, fills) between two events: adjacency_ms is the host's wall clock and
    // also pays three cudaMalloc and two cudaFree, which the driver answers in anything from 0.1 to 100+ ms
    struct EventPair
    {
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair()
        {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } ev;
    ADJ_CUDA(cudaEventCreate(&ev.a));
    ADJ_CUDA(cudaEventCreate(&ev.b));
    ADJ_CUDA(cudaEventRecord(ev.a, st));
    ADJ_CUDA(cudaMemcpyAsync(tri, s->h_tri.data(), m64 * 4, cudaMemcpyHostToDevice, st));
    int bits = 1;
    while (bits < 32 && (1ull << bits) < (uint64_t)s->n_verts) ++bits;
    const unsigned g = (m + 255) / 256;
    k_halfedge_hi<<<g, 256, 0, st>>>(tri, m, key, val);
    radix_sort_pairs(key, val, ktmp, vtmp, m, bits, sort_scr, st);
    k_halfedge_lo<<<g, 256, 0, st>>>(tri, m, val, key);
    radix_sort_pairs(key, val, ktmp, vtmp, m, bits, sort_scr, st);
    k_mark_heads<<<g, 256, 0, st>>>(tri, m, val, head);
    exclusive_scan_u32(head, eid, m, scan_scr, st);
    ADJ_CUDA(cudaGetLastError());
    uint32_t last[2] = {0, 0};
    ADJ_CUDA(cudaMemcpyAsync(&last[0], eid + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    ADJ_CUDA(cudaMemcpyAsync(&last[1], head + (m - 1), 4, cudaMemcpyDeviceToHost, st));
    int32_t *adj_tri = (int32_t *)(adj + p_tri), *adj_tri_edges = (int32_t *)(adj + p_te), *adj_tri_owned = (int32_t *)(adj + p_to);
    int4 *adj_edges4 = (int4 *)(adj + p_e4);
    ADJ_CUDA(cudaMemcpyAsync(adj_tri, tri, m64 * 4, cudaMemcpyDeviceToDevice, st));
    k_fill_edges<<<g, 256, 0, st>>>(tri, m, val, head, eid, adj_edges4, adj_tri_edges);
    k_owned<<<(n + 255) / 256, 256, 0, st>>>(n, head, eid, adj_tri_owned);
    ADJ_CUDA(cudaGetLastError());
    ADJ_CUDA(cudaEventRecord(ev.b, st));
    ADJ_CUDA(cudaStreamSynchronize(st));
    ADJ_CUDA(cudaEventElapsedTime(&s->adjacency_device_ms, ev.a, ev.b));
#undef ADJ_CUDA
    s->adj = adj;
    s->adj_tri = adj_tri;
    s->adj_tri_edges = adj_tri_edges;
    s->adj_tri_owned = adj_tri_owned;
    s->adj_edges4 = adj_edges4;
    cudaFree(tmp);
    cudaFree(tri_buf);
    s->n_edges = last[0] + last[1];
    s->silhouettes_done = true;
    s->adjacency_on_device = true;
    s->adjacency_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return SNCH_OK;
}

// Copies the device adjacency into the host vectors (parity exports before the scene is built).
int fetch_adjacency_host(snch_scene *s)
{
    if (!s->adjacency_on_device || !s->adj || !s->h_tri_edges.empty() || s->n_tris == 0) return SNCH_OK;
    SNCH_CUDA(cudaSetDevice(s->device));
    const size_t m = (size_t)3 * s->n_tris;
    s->h_tri_edges.resize(m);
    s->h_tri_owned.resize(m);
    s->h_edges4.resize((size_t)4 * s->n_edges);
    SNCH_CUDA(cudaMemcpy(s->h_tri_edges.data(), s->adj_tri_edges, m * 4, cudaMemcpyDeviceToHost));
    SNCH_CUDA(cudaMemcpy(s->h_tri_owned.data(), s->adj_tri_owned, m * 4, cudaMemcpyDeviceToHost));
    if (s->n_edges) SNCH_CUDA(cudaMemcpy(s->h_edges4.data(), s->adj_edges4, (size_t)16 * s->n_edges, cudaMemcpyDeviceToHost));
    return SNCH_OK;
}

} // namespace snch
