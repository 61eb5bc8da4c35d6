// capi.cu — the C-ABI (include/snch_b200.h): scene lifetime, build, exports, batched query entry points, replication.
#include "scene.h"
#include "sort_scan.cuh"

#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <new>

namespace snch
{
static thread_local std::string g_last_error;
void set_error(const std::string &msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char *what)
{
    g_last_error = std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + what;
    cudaGetLastError();
    return SNCH_ERR_CUDA;
}

enum PtrKind
{
    PK_NULL,
    PK_HOST,
    PK_DEVICE
};
static PtrKind ptr_kind(const void *p)
{
    if (!p) return PK_NULL;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return PK_HOST;
    }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? PK_DEVICE : PK_HOST;
}

// Per-call device scratch from the scene's stream-ordered pool: allocation and release are ordered on the caller's
// stream, so concurrent batches on different streams (or host threads) never share a buffer.
static int ensure_pool(snch_scene *s)
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
    uint64_t keep = ~0ull; // keep freed blocks cached in the pool: steady-state calls never reach the driver allocator
    SNCH_CUDA(cudaMemPoolSetAttribute(s->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    return SNCH_OK;
}
struct PoolBuffer
{
    snch_scene *s;
    cudaStream_t st;
    unsigned char *p = nullptr;
    int status = SNCH_OK;
    PoolBuffer(snch_scene *s_, cudaStream_t st_, uint64_t bytes) : s(s_), st(st_)
    {
        status = ensure_pool(s);
        if (status != SNCH_OK) return;
        if (cudaMallocFromPoolAsync((void **)&p, bytes ? bytes : 256, s->pool, st) != cudaSuccess)
        {
            cudaGetLastError();
            p = nullptr;
            set_error("out of device memory for the per-call query scratch");
            status = SNCH_ERR_OOM;
        }
    }
    ~PoolBuffer()
    {
        if (p) cudaFreeAsync(p, st);
    }
    PoolBuffer(const PoolBuffer &) = delete;
    PoolBuffer &operator=(const PoolBuffer &) = delete;
};

// Host-pointer batches: inputs are copied into a device staging area, results copied back.  The batch is cut into
// chunks ("query.host_chunk" queries each) that flow through three streams — H2D copies on a scene-owned copy stream,
// kernels on the caller's stream, D2H copies on a second copy stream — so chunk k+1's inputs and chunk k-1's results
// cross PCIe while chunk k is traversed.  Consecutive chunks alternate between the caller's stream and a second compute
// stream (each with its own ordering scratch): a persistent traversal kernel ends with a tail of a few expensive queries
// that occupy single lanes for 2-3 ms (measured: 8 back-to-back chunks on one stream cost +20 ms), and the next chunk's
// CTAs fill the SMs as the previous chunk's drain.  Results do not depend on the chunking (every query is answered on
// its own).
static int ensure_copy_streams(snch_scene *s)
{
    std::lock_guard<std::mutex> lock(s->mu);
    if (s->copy_in) return SNCH_OK;
    SNCH_CUDA(cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking));
    SNCH_CUDA(cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking));
    SNCH_CUDA(cudaStreamCreateWithFlags(&s->compute_b, cudaStreamNonBlocking));
    return SNCH_OK;
}
struct Stager
{
    static constexpr int kMaxArrays = 16, kMaxChunks = 32, kLanes = 2;
    snch_scene *s;
    cudaStream_t st; // the caller's stream: kernels, allocation order, final synchronisation
    unsigned char *base;
    uint64_t m; // queries staged
    uint64_t used = 0;
    int status = SNCH_OK;
    struct Arr
    {
        const unsigned char *src; // host input  (null for outputs)
        unsigned char *dst;       // host output (null for inputs)
        unsigned char *dev;
        uint32_t stride; // bytes per query
    };
    Arr arrs[kMaxArrays];
    int n_arrs = 0;
    cudaEvent_t ev[2 * kMaxChunks + 2];
    int n_ev = 0;
    uint64_t chunk = 0;
    int n_chunks = 0;
    Stager(snch_scene *s_, unsigned char *base_, cudaStream_t st_, uint64_t m_) : s(s_), st(st_), base(base_), m(m_) {}
    ~Stager()
    {
        for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
    }
    Stager(const Stager &) = delete;
    Stager &operator=(const Stager &) = delete;
    static uint64_t pad(uint64_t b) { return align_up(b, 256); }
    template <typename T> const T *in(const T *host, uint32_t stride)
    {
        if (!host || status != SNCH_OK) return nullptr;
        unsigned char *d = base + used;
        used += pad(m * stride);
        arrs[n_arrs++] = Arr{reinterpret_cast<const unsigned char *>(host), nullptr, d, stride};
        return reinterpret_cast<const T *>(d);
    }
    template <typename T> T *out(T *host, uint32_t stride)
    {
        if (!host || status != SNCH_OK) return nullptr;
        unsigned char *d = base + used;
        used += pad(m * stride);
        arrs[n_arrs++] = Arr{nullptr, reinterpret_cast<unsigned char *>(host), d, stride};
        return reinterpret_cast<T *>(d);
    }
    int fail(cudaError_t e, const char *what)
    {
        status = cuda_fail(e, what);
        if (s->copy_in)
        { // nothing may still be touching the staging area when the caller releases it
            cudaStreamSynchronize(s->copy_in);
            cudaStreamSynchronize(s->copy_out);
            cudaStreamSynchronize(s->compute_b);
            cudaGetLastError();
        }
        return status;
    }
    cudaEvent_t new_event()
    {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
        {
            fail(cudaGetLastError(), "cudaEventCreate");
            return nullptr;
        }
        ev[n_ev++] = e;
        return e;
    }
    // launch(o, c, stream, lane) enqueues the kernels for queries [o, o + c) of the staged arrays on `stream`, using the
    // lane's (0 or 1) scratch area
    template <typename Launch> int run(uint64_t chunk_queries, Launch &&launch)
    {
        if (status != SNCH_OK) return status;
        chunk = chunk_queries ? chunk_queries : m;
        if ((m + chunk - 1) / chunk > (uint64_t)(kMaxChunks - 1)) chunk = (m + kMaxChunks - 2) / (kMaxChunks - 1);
        chunk = (chunk + 31) & ~31ull; // keep every chunk's staged arrays 32 B aligned (float3 x 32, byte x 32)
        // Chunk schedule: nothing overlaps the first chunk's H2D copy, so a batch of at least "query.host_split_min" queries
        // starts with a SHORT chunk ("query.host_first": 0 = an eighth of the batch within [512K, 2M], -1 = like the others)
        // and the rest is cut into equal parts of at most `chunk` queries.  Measured on C3 (profiles/r2x_e2e_chunks.json):
        // 16.7M queries 53.1 -> 51.8 ms, 8.4M 28.2 -> 26.9, 4.2M 14.8 -> 14.3; a 2M batch gains nothing from a split.
        uint64_t starts[kMaxChunks + 1];
        n_chunks = 0;
        starts[0] = 0;
        uint64_t first = chunk;
        if (s->tuning.host_first > 0) first = (uint64_t)s->tuning.host_first;
        else if (s->tuning.host_first == 0)
        {
            first = m / 8;
            first = first < (1ull << 19) ? (1ull << 19) : (first > (1ull << 21) ? (1ull << 21) : first);
        }
        first = ((first < chunk ? first : chunk) + 31) & ~31ull;
        const bool split_small = s->tuning.host_first >= 0 && s->tuning.host_split_min > 0 && m >= (uint64_t)s->tuning.host_split_min && m > 2 * first;
        if (chunk_queries && (m > chunk || split_small))
        {
            const uint64_t rest = m - first, parts = (rest + chunk - 1) / chunk;
            const uint64_t each = ((rest + parts - 1) / parts + 31) & ~31ull;
            starts[++n_chunks] = first;
            while (starts[n_chunks] < m)
            {
                starts[n_chunks + 1] = starts[n_chunks] + each < m ? starts[n_chunks] + each : m;
                ++n_chunks;
            }
        }
        else starts[n_chunks = 1] = m;
        if (n_chunks <= 1)
        { // small batch: everything on the caller's stream
            for (int i = 0; i < n_arrs; ++i)
                if (arrs[i].src)
                {
                    const cudaError_t e = cudaMemcpyAsync(arrs[i].dev, arrs[i].src, m * arrs[i].stride, cudaMemcpyHostToDevice, st);
                    if (e != cudaSuccess) return fail(e, "H2D staging copy");
                }
            const int rc = launch((uint64_t)0, m, st, 0);
            if (rc != SNCH_OK) return status = rc;
            for (int i = 0; i < n_arrs; ++i)
                if (arrs[i].dst)
                {
                    const cudaError_t e = cudaMemcpyAsync(arrs[i].dst, arrs[i].dev, m * arrs[i].stride, cudaMemcpyDeviceToHost, st);
                    if (e != cudaSuccess) return fail(e, "D2H staging copy");
                }
            const cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return fail(e, "cudaStreamSynchronize");
            return SNCH_OK;
        }
        if (ensure_copy_streams(s) != SNCH_OK) return status = SNCH_ERR_CUDA;
        cudaError_t e;
        // the staging area was allocated in `st` order: the copy stream starts after it
        cudaEvent_t ready = new_event();
        if (!ready) return status;
        if ((e = cudaEventRecord(ready, st)) != cudaSuccess) return fail(e, "cudaEventRecord");
        if ((e = cudaStreamWaitEvent(s->copy_in, ready, 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent");
        if ((e = cudaStreamWaitEvent(s->compute_b, ready, 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent");
        cudaStream_t lanes[2] = {st, s->compute_b};
        cudaEvent_t h2d_done[kMaxChunks];
        for (int c = 0; c < n_chunks; ++c)
        {
            const uint64_t o = starts[c], cnt = starts[c + 1] - o;
            for (int i = 0; i < n_arrs; ++i)
                if (arrs[i].src)
                {
                    e = cudaMemcpyAsync(arrs[i].dev + o * arrs[i].stride, arrs[i].src + o * arrs[i].stride, cnt * arrs[i].stride,
                                        cudaMemcpyHostToDevice, s->copy_in);
                    if (e != cudaSuccess) return fail(e, "H2D staging copy");
                }
            if (!(h2d_done[c] = new_event())) return status;
            if ((e = cudaEventRecord(h2d_done[c], s->copy_in)) != cudaSuccess) return fail(e, "cudaEventRecord");
        }
        for (int c = 0; c < n_chunks; ++c)
        {
            const uint64_t o = starts[c], cnt = starts[c + 1] - o;
            cudaStream_t cs = lanes[c & 1];
            if ((e = cudaStreamWaitEvent(cs, h2d_done[c], 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent");
            const int rc = launch(o, cnt, cs, c & 1);
            if (rc != SNCH_OK)
            {
                cudaStreamSynchronize(s->copy_in); // nothing may still be touching the staging area when it is released
                cudaStreamSynchronize(s->copy_out);
                cudaStreamSynchronize(s->compute_b);
                return status = rc;
            }
            cudaEvent_t kernels_done = new_event();
            if (!kernels_done) return status;
            if ((e = cudaEventRecord(kernels_done, cs)) != cudaSuccess) return fail(e, "cudaEventRecord");
            if ((e = cudaStreamWaitEvent(s->copy_out, kernels_done, 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent");
            for (int i = 0; i < n_arrs; ++i)
                if (arrs[i].dst)
                {
                    e = cudaMemcpyAsync(arrs[i].dst + o * arrs[i].stride, arrs[i].dev + o * arrs[i].stride, cnt * arrs[i].stride,
                                        cudaMemcpyDeviceToHost, s->copy_out);
                    if (e != cudaSuccess) return fail(e, "D2H staging copy");
                }
        }
        cudaEvent_t all_out = new_event();
        if (!all_out) return status;
        if ((e = cudaEventRecord(all_out, s->copy_out)) != cudaSuccess) return fail(e, "cudaEventRecord");
        if ((e = cudaStreamWaitEvent(st, all_out, 0)) != cudaSuccess) return fail(e, "cudaStreamWaitEvent"); // release after the last D2H
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e, "cudaStreamSynchronize");
        return SNCH_OK;
    }
};

// one launch handles at most this many queries (32-bit slots in the kernels); larger batches are split
constexpr uint64_t kMaxLaunch = 1ull << 30;

// queries per pipeline chunk of a host-pointer batch of m queries ("query.host_chunk"; 0 = one chunk)
static uint64_t host_chunk(const snch_scene *s, uint64_t m)
{
    uint64_t c = s->tuning.host_chunk > 0 ? (uint64_t)s->tuning.host_chunk : m;
    if (c > m) c = m;
    if ((m + c - 1) / c > (uint64_t)(Stager::kMaxChunks - 1)) c = (m + Stager::kMaxChunks - 2) / (Stager::kMaxChunks - 1);
    return (c + 31) & ~31ull;
}

static int check_built(const snch_scene *s)
{
    if (!s)
    {
        set_error("null scene");
        return SNCH_ERR_INVALID;
    }
    if (!s->built)
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    return SNCH_OK;
}
// all non-null pointers must be of one kind; returns PK_DEVICE / PK_HOST, or PK_NULL on a mix
static PtrKind common_kind(std::initializer_list<const void *> ptrs)
{
    PtrKind k = PK_NULL;
    for (const void *p : ptrs)
    {
        const PtrKind pk = ptr_kind(p);
        if (pk == PK_NULL) continue;
        if (k == PK_NULL) k = pk;
        else if (k != pk) return PK_NULL;
    }
    return k;
}
} // namespace snch

using namespace snch;

extern "C"
{

const char *snch_last_error(void) { return g_last_error.c_str(); }
int snch_abi_version(void) { return SNCH_B200_ABI_VERSION; }

int snch_scene3_create(const float *xyz, uint32_t n_verts, const int32_t *tri, uint32_t n_tris, int device, snch_scene **out)
{
    if (!out || (n_verts && !xyz) || (n_tris && !tri))
    {
        set_error("snch_scene3_create: null argument");
        return SNCH_ERR_INVALID;
    }
    *out = nullptr;
    if (n_tris > (1u << 27))
    { // child references keep bit 31 as the leaf flag and a leaf's edge payload is (first edge id << 2 | count): with at most
      // 3 edges per triangle that bounds a scene at 2^27 = 134M triangles (a 76 GB arena)
        set_error("snch_scene3_create: more than 2^27 triangles (32-bit child references / edge payloads)");
        return SNCH_ERR_INVALID;
    }
    for (uint64_t i = 0; i < (uint64_t)3 * n_tris; ++i)
        if (tri[i] < 0 || (uint32_t)tri[i] >= n_verts)
        {
            set_error("snch_scene3_create: vertex index out of range");
            return SNCH_ERR_INVALID;
        }
    if (device < 0)
    {
        set_error("snch_scene3_create: bad device ordinal");
        return SNCH_ERR_INVALID;
    }
    // no CUDA call here: creation and compute_silhouettes() are host-only (as in the reference, whose scene constructor
    // only fills host vectors before the first device_vector copy); the device is first touched by snch_scene_build().
    snch_scene *s = new (std::nothrow) snch_scene();
    if (!s)
    {
        set_error("out of host memory");
        return SNCH_ERR_OOM;
    }
    s->device = device;
    s->n_verts = n_verts;
    s->n_tris = n_tris;
    s->h_xyz.assign(xyz, xyz + (size_t)3 * n_verts);
    s->h_tri.assign(tri, tri + (size_t)3 * n_tris);
    *out = s;
    return SNCH_OK;
}

int snch_scene_destroy(snch_scene *s)
{
    if (!s) return SNCH_OK;
    cudaSetDevice(s->device);
    if (s->arena) cudaFree(s->arena);
    if (s->adj) cudaFree(s->adj);
    if (s->scratch) cudaFree(s->scratch);
    if (s->pool) cudaMemPoolDestroy(s->pool);
    if (s->copy_in) cudaStreamDestroy(s->copy_in);
    if (s->copy_out) cudaStreamDestroy(s->copy_out);
    if (s->compute_b) cudaStreamDestroy(s->compute_b);
    s->counters.release();
    delete s;
    return SNCH_OK;
}

int snch_scene_compute_silhouettes(snch_scene *s)
{
    if (!s || s->adopted)
    {
        set_error("snch_scene_compute_silhouettes: invalid scene");
        return SNCH_ERR_INVALID;
    }
    // The reference does this on the host (scene.cuh:1135-1229).  Here it runs on the GPU (adjacency.cu) whenever a CUDA
    // device is present; "adjacency.device" = 0 selects the host passes, which produce the same arrays bit for bit.
    s->arena_has_topology = false;
    bool on_device = s->adjacency_mode == 1;
    if (s->adjacency_mode < 0)
    {
        int count = 0;
        on_device = cudaGetDeviceCount(&count) == cudaSuccess && s->device < count;
        if (!on_device) cudaGetLastError();
    }
    if (on_device) return compute_adjacency_device(s);
    free_adjacency(s);
    s->adjacency_on_device = false;
    compute_adjacency_host(s);
    return SNCH_OK;
}

int snch_scene_build(snch_scene *s, const snch_build_options *opts, snch_stream stream)
{
    if (!s || s->adopted)
    {
        set_error("snch_scene_build: invalid scene");
        return SNCH_ERR_INVALID;
    }
    if (opts)
    {
        if (opts->struct_size != sizeof(snch_build_options))
        {
            set_error("snch_scene_build: snch_build_options.struct_size mismatch");
            return SNCH_ERR_INVALID;
        }
        s->opt_print_collision = opts->print_collision;
        s->opt_refit_only = opts->refit_only;
    }
    else s->opt_refit_only = 0;
    if (s->opt_refit_only && !s->built)
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    // the reference's build_bvh() silently uses whatever compute_silhouettes() left behind; an un-prepared scene has
    // no edges at all, which makes every leaf cone invalid.  Do the same (no implicit call).
    if (!s->silhouettes_done)
    {
        s->n_edges = 0;
        s->adjacency_on_device = false;
        s->arena_has_topology = false;
        s->h_edges4.clear();
        s->h_tri_edges.assign((size_t)3 * s->n_tris, -1);
        s->h_tri_owned.assign((size_t)3 * s->n_tris, -1);
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return SNCH_ERR_CUDA;
    }
    if (s->device >= count)
    {
        set_error("snch_scene_build: bad device ordinal");
        return SNCH_ERR_INVALID;
    }
    return build_device(s, (cudaStream_t)stream);
}

int snch_scene_stats(const snch_scene *s, snch_build_stats *out)
{
    if (!s || !out)
    {
        set_error("snch_scene_stats: null argument");
        return SNCH_ERR_INVALID;
    }
    std::memset(out, 0, sizeof *out);
    out->num_objects = s->n_tris;
    out->num_vertices = s->n_verts;
    out->num_edges = s->n_edges;
    out->adjacency_ms = s->adjacency_ms;
    if (s->built)
    {
        out->num_nodes = s->hdr.n_nodes;
        out->morton_collision = s->hdr.collision;
        out->q1_nodes = s->hdr.q1_nodes;
        out->build_ms = s->build_ms;
        out->arena_bytes = s->arena_bytes;
        for (int a = 0; a < 3; ++a)
        {
            out->scene_lower[a] = s->hdr.scene_lo[a];
            out->scene_upper[a] = s->hdr.scene_hi[a];
        }
    }
    return SNCH_OK;
}

int snch_scene_device_repr(const snch_scene *s, snch_bvh_device_pod *out)
{
    if (!out)
    {
        set_error("snch_scene_device_repr: null argument");
        return SNCH_ERR_INVALID;
    }
    const int st = check_built(s);
    if (st != SNCH_OK) return st;
    const ArenaHeader &h = s->hdr;
    out->num_nodes = h.n_nodes;
    out->num_objects = h.n_tris;
    out->num_vertices = h.n_verts;
    out->num_silhouettes = h.n_edges;
    const bool empty = h.n_tris == 0;
    out->nodes = empty ? nullptr : s->arena + h.off_nodes;
    out->aabbs = empty ? nullptr : s->arena + h.off_aabbs;
    out->cones = empty ? nullptr : s->arena + h.off_cones;
    out->objects = empty ? nullptr : s->arena + h.off_objects;
    out->vertices = s->arena + h.off_vertices;
    out->silhouettes = s->arena + h.off_edges;
    return SNCH_OK;
}

int snch_scene_export(const snch_scene *s, int kind, void *host_dst, size_t bytes)
{
    if (!host_dst && bytes)
    {
        set_error("snch_scene_export: null destination");
        return SNCH_ERR_INVALID;
    }
    if (s && !s->adopted && s->silhouettes_done && s->adjacency_on_device && s->adj)
    {
        const int fs = fetch_adjacency_host(const_cast<snch_scene *>(s));
        if (fs != SNCH_OK) return fs;
    }
    if (s && !s->adopted && s->silhouettes_done && (!s->adjacency_on_device || !s->h_tri_edges.empty() || s->n_tris == 0) &&
        (kind == SNCH_EXPORT_EDGES || kind == SNCH_EXPORT_TRI_EDGES || kind == SNCH_EXPORT_TRI_OWNED))
    { // adjacency products are available as soon as compute_silhouettes() ran
        const std::vector<int32_t> &v = kind == SNCH_EXPORT_EDGES ? s->h_edges4 : (kind == SNCH_EXPORT_TRI_EDGES ? s->h_tri_edges : s->h_tri_owned);
        if (bytes != v.size() * 4)
        {
            set_error("snch_scene_export: size mismatch");
            return SNCH_ERR_INVALID;
        }
        if (bytes) std::memcpy(host_dst, v.data(), bytes);
        return SNCH_OK;
    }
    const int st = check_built(s);
    if (st != SNCH_OK) return st;
    const ArenaHeader &h = s->hdr;
    uint64_t off = 0, want = 0;
    const void *host_src = nullptr;
    std::vector<int32_t> tmp;
    switch (kind)
    {
    case SNCH_EXPORT_NODES: off = h.off_nodes; want = (uint64_t)h.n_nodes * 16; break;
    case SNCH_EXPORT_AABBS: off = h.off_aabbs; want = (uint64_t)h.n_nodes * 24; break;
    case SNCH_EXPORT_CONES: off = h.off_cones; want = (uint64_t)h.n_nodes * 20; break;
    case SNCH_EXPORT_MORTON_SORTED: off = h.off_morton; want = (uint64_t)h.n_tris * 4; break;
    case SNCH_EXPORT_SORTED_INDEX: off = h.off_sorted_idx; want = (uint64_t)h.n_tris * 4; break;
    case SNCH_EXPORT_RANGES: off = h.off_ranges; want = (uint64_t)h.n_internal * 8; break;
    case SNCH_EXPORT_Q1_TAINT: off = h.off_q1; want = (uint64_t)h.n_nodes; break;
    case SNCH_EXPORT_TRI_EDGES: off = h.off_tri_edges; want = (uint64_t)h.n_tris * 12; break;
    case SNCH_EXPORT_EDGES:
    case SNCH_EXPORT_TRI_OWNED:
    {
        // unpack from the reference-layout structs living in the arena (works for adopted replicas too)
        const bool edges = kind == SNCH_EXPORT_EDGES;
        const uint64_t cnt = edges ? h.n_edges : h.n_tris;
        want = cnt * (edges ? 16 : 12);
        if (bytes != want) break;
        const uint64_t rec = edges ? sizeof(RefEdge) : sizeof(RefTriangle);
        std::vector<unsigned char> raw(cnt * rec);
        SNCH_CUDA(cudaSetDevice(s->device));
        if (cnt) SNCH_CUDA(cudaMemcpy(raw.data(), s->arena + (edges ? h.off_edges : h.off_objects), cnt * rec, cudaMemcpyDeviceToHost));
        int32_t *dst = (int32_t *)host_dst;
        for (uint64_t i = 0; i < cnt; ++i)
        {
            if (edges) std::memcpy(dst + 4 * i, raw.data() + i * rec, 16);
            else std::memcpy(dst + 3 * i, raw.data() + i * rec + 12, 12);
        }
        return SNCH_OK;
    }
    default:
        set_error("snch_scene_export: unknown kind");
        return SNCH_ERR_INVALID;
    }
    if (bytes != want)
    {
        char buf[128];
        std::snprintf(buf, sizeof buf, "snch_scene_export: kind %d needs %llu bytes, got %llu", kind, (unsigned long long)want,
                      (unsigned long long)bytes);
        set_error(buf);
        return SNCH_ERR_INVALID;
    }
    if (host_src)
    {
        std::memcpy(host_dst, host_src, bytes);
        return SNCH_OK;
    }
    SNCH_CUDA(cudaSetDevice(s->device));
    if (bytes) SNCH_CUDA(cudaMemcpy(host_dst, s->arena + off, bytes, cudaMemcpyDeviceToHost));
    return SNCH_OK;
}

// ---- batched queries ----------------------------------------------------------------------------------------------
// Every entry point: validate, classify the pointers, then per slice of <= kMaxLaunch queries either launch directly on
// the caller's device buffers or stage host buffers through pool memory (H2D, launch, D2H, synchronise).
int snch_closest_point_batch(const snch_scene *cs, const float *pts, uint64_t n, uint32_t *out_index, float *out_distance, snch_stream stream)
{
    int st = check_built(cs);
    if (st != SNCH_OK) return st;
    if (n == 0) return SNCH_OK;
    if (!pts || !out_index || !out_distance)
    {
        set_error("snch_closest_point_batch: null argument");
        return SNCH_ERR_INVALID;
    }
    snch_scene *s = const_cast<snch_scene *>(cs);
    SNCH_CUDA(cudaSetDevice(s->device));
    const PtrKind k = common_kind({pts, out_index, out_distance});
    if (k == PK_NULL)
    {
        set_error("snch_closest_point_batch: mixed host/device pointers");
        return SNCH_ERR_POINTER_KIND;
    }
    cudaStream_t cst = (cudaStream_t)stream;
    for (uint64_t off = 0; off < n; off += kMaxLaunch)
    {
        const uint64_t m = n - off < kMaxLaunch ? n - off : kMaxLaunch;
        const uint64_t cm = k == PK_DEVICE ? m : host_chunk(s, m);
        const uint64_t qs = query_scratch_bytes(cm, s->tuning);
        const uint64_t stage = k == PK_DEVICE ? 0 : Stager::pad(m * 12) + 2 * Stager::pad(m * 4);
        PoolBuffer buf(s, cst, k == PK_DEVICE ? qs : Stager::kLanes * qs + stage);
        if (buf.status != SNCH_OK) return buf.status;
        if (k == PK_DEVICE)
        {
            st = launch_closest(s->view, s->tuning, pts + 3 * off, m, out_index + off, out_distance + off, buf.p, cst, &s->counters);
            if (st != SNCH_OK) return st;
            continue;
        }
        Stager sg(s, buf.p + Stager::kLanes * qs, cst, m);
        const float *dq = sg.in(pts + 3 * off, 12);
        uint32_t *di = sg.out(out_index + off, 4);
        float *dd = sg.out(out_distance + off, 4);
        st = sg.run(cm, [&](uint64_t o, uint64_t c, cudaStream_t ls, int lane) {
            return launch_closest(s->view, s->tuning, dq + 3 * o, c, di + o, dd + o, buf.p + lane * qs, ls, &s->counters);
        });
        if (st != SNCH_OK) return st;
    }
    return SNCH_OK;
}

int snch_closest_silhouette_batch(const snch_scene *cs, const float *pts, const uint8_t *flip, const float *r_max, uint64_t n,
                                  float *out_distance, uint32_t *out_edge, float *out_point, snch_stream stream)
{
    int st = check_built(cs);
    if (st != SNCH_OK) return st;
    if (n == 0) return SNCH_OK;
    if (!pts || !out_distance)
    {
        set_error("snch_closest_silhouette_batch: null argument");
        return SNCH_ERR_INVALID;
    }
    snch_scene *s = const_cast<snch_scene *>(cs);
    SNCH_CUDA(cudaSetDevice(s->device));
    const PtrKind k = common_kind({pts, flip, r_max, out_distance, out_edge, out_point});
    if (k == PK_NULL)
    {
        set_error("snch_closest_silhouette_batch: mixed host/device pointers");
        return SNCH_ERR_POINTER_KIND;
    }
    cudaStream_t cst = (cudaStream_t)stream;
    for (uint64_t off = 0; off < n; off += kMaxLaunch)
    {
        const uint64_t m = n - off < kMaxLaunch ? n - off : kMaxLaunch;
        const uint64_t cm = k == PK_DEVICE ? m : host_chunk(s, m);
        const uint64_t qs = query_scratch_bytes(cm, s->tuning);
        const uint64_t stage = k == PK_DEVICE ? 0 : 2 * Stager::pad(m * 12) + Stager::pad(m) + 3 * Stager::pad(m * 4);
        PoolBuffer buf(s, cst, k == PK_DEVICE ? qs : Stager::kLanes * qs + stage);
        if (buf.status != SNCH_OK) return buf.status;
        const uint8_t *fo = flip ? flip + off : nullptr;
        const float *ro = r_max ? r_max + off : nullptr;
        uint32_t *eo = out_edge ? out_edge + off : nullptr;
        float *po = out_point ? out_point + 3 * off : nullptr;
        if (k == PK_DEVICE)
        {
            st = launch_silhouette(s->view, s->tuning, pts + 3 * off, fo, ro, m, out_distance + off, eo, po, buf.p, cst, &s->counters);
            if (st != SNCH_OK) return st;
            continue;
        }
        Stager sg(s, buf.p + Stager::kLanes * qs, cst, m);
        const float *dq = sg.in(pts + 3 * off, 12);
        const uint8_t *df = sg.in(fo, 1);
        const float *dr = sg.in(ro, 4);
        float *dd = sg.out(out_distance + off, 4);
        uint32_t *de = sg.out(eo, 4);
        float *dp = sg.out(po, 12);
        st = sg.run(cm, [&](uint64_t o, uint64_t c, cudaStream_t ls, int lane) {
            return launch_silhouette(s->view, s->tuning, dq + 3 * o, df ? df + o : nullptr, dr ? dr + o : nullptr, c, dd + o, de ? de + o : nullptr,
                                     dp ? dp + 3 * o : nullptr, buf.p + lane * qs, ls, &s->counters);
        });
        if (st != SNCH_OK) return st;
    }
    return SNCH_OK;
}

int snch_intersect_batch(const snch_scene *cs, const float *org, const float *dir, const float *t_max, uint64_t n, snch_hit *out_hits,
                         uint8_t *out_found, int any_hit, snch_stream stream)
{
    int st = check_built(cs);
    if (st != SNCH_OK) return st;
    if (n == 0) return SNCH_OK;
    if (!org || !dir || (any_hit && !out_found) || (!any_hit && !out_hits && !out_found))
    {
        set_error("snch_intersect_batch: null argument");
        return SNCH_ERR_INVALID;
    }
    snch_scene *s = const_cast<snch_scene *>(cs);
    SNCH_CUDA(cudaSetDevice(s->device));
    const PtrKind k = common_kind({org, dir, t_max, out_hits, out_found});
    if (k == PK_NULL)
    {
        set_error("snch_intersect_batch: mixed host/device pointers");
        return SNCH_ERR_POINTER_KIND;
    }
    cudaStream_t cst = (cudaStream_t)stream;
    for (uint64_t off = 0; off < n; off += kMaxLaunch)
    {
        const uint64_t m = n - off < kMaxLaunch ? n - off : kMaxLaunch;
        const uint64_t cm = k == PK_DEVICE ? m : host_chunk(s, m);
        const uint64_t qs = query_scratch_bytes(cm, s->tuning);
        const uint64_t stage = k == PK_DEVICE ? 0 : 2 * Stager::pad(m * 12) + Stager::pad(m * 4) + Stager::pad(m * 16) + Stager::pad(m);
        PoolBuffer buf(s, cst, k == PK_DEVICE ? qs : Stager::kLanes * qs + stage);
        if (buf.status != SNCH_OK) return buf.status;
        const float *to = t_max ? t_max + off : nullptr;
        snch_hit *ho = out_hits ? out_hits + off : nullptr;
        uint8_t *fo = out_found ? out_found + off : nullptr;
        if (k == PK_DEVICE)
        {
            st = launch_intersect(s->view, s->tuning, org + 3 * off, dir + 3 * off, to, m, ho, fo, any_hit, buf.p, cst, &s->counters);
            if (st != SNCH_OK) return st;
            continue;
        }
        Stager sg(s, buf.p + Stager::kLanes * qs, cst, m);
        const float *dor = sg.in(org + 3 * off, 12);
        const float *ddi = sg.in(dir + 3 * off, 12);
        const float *dtm = sg.in(to, 4);
        snch_hit *dh = sg.out(ho, (uint32_t)sizeof(snch_hit));
        uint8_t *df = sg.out(fo, 1);
        st = sg.run(cm, [&](uint64_t o, uint64_t c, cudaStream_t ls, int lane) {
            return launch_intersect(s->view, s->tuning, dor + 3 * o, ddi + 3 * o, dtm ? dtm + o : nullptr, c, dh ? dh + o : nullptr,
                                    df ? df + o : nullptr, any_hit, buf.p + lane * qs, ls, &s->counters);
        });
        if (st != SNCH_OK) return st;
    }
    return SNCH_OK;
}

int snch_sample_in_sphere_batch(const snch_scene *cs, const float *spheres, const float *rnd, uint64_t n, int32_t *out_index, float *out_pdf,
                                float *out_point, snch_stream stream)
{
    int st = check_built(cs);
    if (st != SNCH_OK) return st;
    if (n == 0) return SNCH_OK;
    if (!spheres || !rnd || !out_index || !out_pdf)
    {
        set_error("snch_sample_in_sphere_batch: null argument");
        return SNCH_ERR_INVALID;
    }
    snch_scene *s = const_cast<snch_scene *>(cs);
    SNCH_CUDA(cudaSetDevice(s->device));
    const PtrKind k = common_kind({spheres, rnd, out_index, out_pdf, out_point});
    if (k == PK_NULL)
    {
        set_error("snch_sample_in_sphere_batch: mixed host/device pointers");
        return SNCH_ERR_POINTER_KIND;
    }
    cudaStream_t cst = (cudaStream_t)stream;
    for (uint64_t off = 0; off < n; off += kMaxLaunch)
    {
        const uint64_t m = n - off < kMaxLaunch ? n - off : kMaxLaunch;
        const uint64_t cm = k == PK_DEVICE ? m : host_chunk(s, m);
        const uint64_t qs = query_scratch_bytes(cm, s->tuning);
        const uint64_t stage = k == PK_DEVICE ? 0 : Stager::pad(m * 16) + 2 * Stager::pad(m * 12) + 2 * Stager::pad(m * 4);
        PoolBuffer buf(s, cst, k == PK_DEVICE ? qs : Stager::kLanes * qs + stage);
        if (buf.status != SNCH_OK) return buf.status;
        float *po = out_point ? out_point + 3 * off : nullptr;
        if (k == PK_DEVICE)
        {
            st = launch_sample(s->view, s->tuning, spheres + 4 * off, rnd + 3 * off, m, out_index + off, out_pdf + off, po, buf.p, cst, &s->counters);
            if (st != SNCH_OK) return st;
            continue;
        }
        Stager sg(s, buf.p + Stager::kLanes * qs, cst, m);
        const float *ds = sg.in(spheres + 4 * off, 16);
        const float *dr = sg.in(rnd + 3 * off, 12);
        int32_t *di = sg.out(out_index + off, 4);
        float *dp = sg.out(out_pdf + off, 4);
        float *dpt = sg.out(po, 12);
        st = sg.run(cm, [&](uint64_t o, uint64_t c, cudaStream_t ls, int lane) {
            return launch_sample(s->view, s->tuning, ds + 4 * o, dr + 3 * o, c, di + o, dp + o, dpt ? dpt + 3 * o : nullptr, buf.p + lane * qs, ls,
                                 &s->counters);
        });
        if (st != SNCH_OK) return st;
    }
    return SNCH_OK;
}

int snch_scene_counter(snch_scene *s, const char *name, double *value, int reset)
{
    if (!s || !name || !value)
    {
        set_error("snch_scene_counter: null argument");
        return SNCH_ERR_INVALID;
    }
    const std::string k(name);
    QueryCounters &c = s->counters;
    if (k == "query.launches") *value = (double)c.launches.load();
    else if (k == "query.traversal_launches") *value = (double)c.traversal_launches.load();
    else if (k == "query.traversal_ms")
    {
        cudaSetDevice(s->device);
        c.fold();
        std::lock_guard<std::mutex> lock(c.mu);
        *value = c.traversal_ms;
    }
    else if (k == "build.launches") *value = (double)s->build_launches;
    else if (k == "adjacency.device_ms") *value = (double)s->adjacency_device_ms;
    else
    {
        set_error("snch_scene_counter: unknown counter '" + k + "'");
        return SNCH_ERR_INVALID;
    }
    if (reset)
    {
        cudaSetDevice(s->device);
        c.reset();
    }
    return SNCH_OK;
}

const char *snch_scene_last_kernel(const snch_scene *s) { return s ? s->counters.last_kernel.load() : ""; }

int snch_scene_set_option(snch_scene *s, const char *name, int64_t value)
{
    if (!s || !name)
    {
        set_error("snch_scene_set_option: null argument");
        return SNCH_ERR_INVALID;
    }
    const std::string k(name);
    QueryTuning &t = s->tuning;
    if (k == "query.sort_min_n") t.sort_min_n = (int)value;
    else if (k == "query.sort_bits") t.sort_bits = (int)value;
    else if (k == "query.sort_rays") t.sort_rays = (int)value;
    else if (k == "query.cone_filter") t.cone_filter = (int)value;
    else if (k == "query.seed") t.seed = (int)value;
    else if (k == "query.sort_radius") t.sort_radius = (int)value;
    else if (k == "query.sil_tail") t.sil_tail = (int)value;
    else if (k == "query.sil_flush") t.sil_flush = (int)value;
    else if (k == "query.sil_chunk") t.sil_chunk = (int)value;
    else if (k == "query.wide_max_n") t.wide_max_n = (int)value;
    else if (k == "query.wide_max_n_sil") t.wide_max_n_sil = (int)value;
    else if (k == "query.ray_kernel") t.ray_kernel = (int)value;
    else if (k == "query.ray_flush") t.ray_flush = (int)value;
    else if (k == "query.ray_refill") t.ray_refill = (int)value;
    else if (k == "query.blocks_per_sm") t.blocks_per_sm = (int)value;
    else if (k == "query.host_chunk") t.host_chunk = (int)(value < 0 ? 0 : value);
    else if (k == "query.host_first") t.host_first = (int)(value < 0 ? -1 : value);
    else if (k == "query.host_split_min") t.host_split_min = (int)(value < 0 ? 0 : value);
    else if (k == "query.time_kernels") s->counters.time_kernels = (int)value;
    else if (k == "adjacency.device") s->adjacency_mode = (int)value;
    else if (k == "build.refit_kernel") s->opt_refit_kernel = (int)value;
    else if (k == "sort.lookback") set_sort_lookback((int)value); // process-wide: 1 = one predecessor tile per L2 round trip, else 8
    else if (k == "sort.onesweep") set_sort_onesweep((int)value); // process-wide (A/B of the two radix sorts)
    else
    {
        set_error("snch_scene_set_option: unknown option '" + k + "'");
        return SNCH_ERR_INVALID;
    }
    return SNCH_OK;
}

// ---- replication ------------------------------------------------------------------------------------------------------
int snch_scene_arena(const snch_scene *s, void **device_ptr, uint64_t *bytes)
{
    const int st = check_built(s);
    if (st != SNCH_OK) return st;
    if (!device_ptr || !bytes)
    {
        set_error("snch_scene_arena: null argument");
        return SNCH_ERR_INVALID;
    }
    *device_ptr = s->arena;
    *bytes = s->arena_bytes;
    return SNCH_OK;
}

// Replica creation in two steps shared by adopt / load / broadcast / peer fan-out: adopt_begin validates the (untrusted)
// header and allocates the arena on `device`; the caller fills it; adopt_end re-patches the embedded pointers.
extern "C++"
{
namespace snch
{
int adopt_begin(const ArenaHeader &h, uint64_t bytes, int device, const char *who, snch_scene **out)
{
    *out = nullptr;
    SNCH_CUDA(cudaSetDevice(device));
    if (h.magic != kArenaMagic || h.version != kArenaVersion || h.total_bytes != bytes)
    {
        set_error(std::string(who) + ": not a scene arena (magic/version/size mismatch)");
        return SNCH_ERR_INVALID;
    }
    { // the header is untrusted input (a file, a peer's bytes): every count and offset must be exactly what this library lays
      // out for (n_verts, n_tris, n_edges) — the traversal kernels and the pointer patch index the arena through them
        ArenaHeader want;
        if (h.n_tris <= 0x3FFFFFFFu) layout_arena(want, h.n_verts, h.n_tris, h.n_edges);
        const bool same = h.n_tris <= 0x3FFFFFFFu && want.total_bytes == h.total_bytes && want.n_nodes == h.n_nodes && want.n_internal == h.n_internal &&
                          want.off_vertices == h.off_vertices && want.off_edges == h.off_edges && want.off_objects == h.off_objects &&
                          want.off_tri_edges == h.off_tri_edges && want.off_nodes == h.off_nodes && want.off_aabbs == h.off_aabbs &&
                          want.off_cones == h.off_cones && want.off_morton == h.off_morton && want.off_sorted_idx == h.off_sorted_idx &&
                          want.off_ranges == h.off_ranges && want.off_q1 == h.off_q1 && want.off_bnode == h.off_bnode && want.off_snode == h.off_snode &&
                          want.off_ltri == h.off_ltri && want.off_ledge == h.off_ledge && want.off_edge_off == h.off_edge_off;
        if (!same)
        {
            set_error(std::string(who) + ": arena header is inconsistent with its own counts (corrupt or foreign arena)");
            return SNCH_ERR_INVALID;
        }
    }
    snch_scene *s = new (std::nothrow) snch_scene();
    if (!s)
    {
        set_error("out of host memory");
        return SNCH_ERR_OOM;
    }
    s->device = device;
    s->adopted = true;
    s->n_verts = h.n_verts;
    s->n_tris = h.n_tris;
    s->n_edges = h.n_edges;
    if (cudaMalloc(&s->arena, bytes) != cudaSuccess)
    {
        cudaGetLastError();
        delete s;
        set_error("cudaMalloc of the adopted arena failed");
        return SNCH_ERR_OOM;
    }
    s->arena_bytes = bytes;
    s->hdr = h;
    *out = s;
    return SNCH_OK;
}
int adopt_end(snch_scene *s, cudaStream_t cst)
{
    resolve_view(s);
    int st = patch_pointers(s, cst); // reference-layout structs embed raw pointers (scene.cuh:831-839)
    if (st == SNCH_OK)
    {
        const cudaError_t e = cudaStreamSynchronize(cst);
        if (e != cudaSuccess) st = cuda_fail(e, "cudaStreamSynchronize");
    }
    if (st == SNCH_OK) s->built = true;
    return st;
}
} // namespace snch
} // extern "C++"

static int adopt_from(const void *src, bool src_is_host, uint64_t bytes, int device, cudaStream_t cst, snch_scene **out, const char *who)
{
    *out = nullptr;
    SNCH_CUDA(cudaSetDevice(device));
    ArenaHeader h;
    if (src_is_host) std::memcpy(&h, src, sizeof h);
    else
    {
        SNCH_CUDA(cudaMemcpyAsync(&h, src, sizeof h, cudaMemcpyDeviceToHost, cst));
        SNCH_CUDA(cudaStreamSynchronize(cst));
    }
    snch_scene *s = nullptr;
    int st = adopt_begin(h, bytes, device, who, &s);
    if (st != SNCH_OK) return st;
    const cudaError_t e = cudaMemcpyAsync(s->arena, src, bytes, src_is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, cst);
    if (e != cudaSuccess)
    {
        snch_scene_destroy(s);
        return cuda_fail(e, "arena copy");
    }
    st = adopt_end(s, cst);
    if (st != SNCH_OK)
    {
        snch_scene_destroy(s);
        return st;
    }
    *out = s;
    return SNCH_OK;
}

int snch_scene_adopt_arena(const void *arena_copy, uint64_t bytes, int device, snch_stream stream, snch_scene **out)
{
    if (!arena_copy || !out || bytes < sizeof(ArenaHeader))
    {
        set_error("snch_scene_adopt_arena: bad argument");
        return SNCH_ERR_INVALID;
    }
    return adopt_from(arena_copy, false, bytes, device, (cudaStream_t)stream, out, "snch_scene_adopt_arena");
}

// ---- serialisation ----------------------------------------------------------------------------------------------------
int snch_scene_save(const snch_scene *s, const char *path)
{
    const int st = check_built(s);
    if (st != SNCH_OK) return st;
    if (!path)
    {
        set_error("snch_scene_save: null path");
        return SNCH_ERR_INVALID;
    }
    SNCH_CUDA(cudaSetDevice(s->device));
    std::FILE *f = std::fopen(path, "wb");
    if (!f)
    {
        set_error(std::string("snch_scene_save: cannot open '") + path + "' for writing");
        return SNCH_ERR_INVALID;
    }
    // chunked D2H through one pinned bounce buffer: the arena can be GBs, the host copy need not be
    const uint64_t chunk = 64ull << 20;
    void *bounce = nullptr;
    if (cudaMallocHost(&bounce, chunk) != cudaSuccess)
    {
        cudaGetLastError();
        std::fclose(f);
        set_error("snch_scene_save: out of pinned host memory");
        return SNCH_ERR_OOM;
    }
    int rc = SNCH_OK;
    for (uint64_t off = 0; off < s->arena_bytes && rc == SNCH_OK; off += chunk)
    {
        const uint64_t m = s->arena_bytes - off < chunk ? s->arena_bytes - off : chunk;
        const cudaError_t e = cudaMemcpy(bounce, s->arena + off, m, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = cuda_fail(e, "arena D2H");
        else if (std::fwrite(bounce, 1, m, f) != m)
        {
            set_error(std::string("snch_scene_save: short write to '") + path + "'");
            rc = SNCH_ERR_INVALID;
        }
    }
    cudaFreeHost(bounce);
    if (std::fclose(f) != 0 && rc == SNCH_OK)
    {
        set_error(std::string("snch_scene_save: close of '") + path + "' failed");
        rc = SNCH_ERR_INVALID;
    }
    return rc;
}

int snch_scene_load(const char *path, int device, snch_stream stream, snch_scene **out)
{
    if (!path || !out)
    {
        set_error("snch_scene_load: null argument");
        return SNCH_ERR_INVALID;
    }
    *out = nullptr;
    std::FILE *f = std::fopen(path, "rb");
    if (!f)
    {
        set_error(std::string("snch_scene_load: cannot open '") + path + "'");
        return SNCH_ERR_INVALID;
    }
    std::vector<unsigned char> buf;
    ArenaHeader h;
    int rc = SNCH_OK;
    if (std::fread(&h, 1, sizeof h, f) != sizeof h || h.magic != kArenaMagic || h.version != kArenaVersion || h.total_bytes < sizeof h ||
        h.total_bytes > (1ull << 40))
    {
        set_error(std::string("snch_scene_load: '") + path + "' is not a scene arena (magic/version mismatch)");
        rc = SNCH_ERR_INVALID;
    }
    else
    {
        buf.resize(h.total_bytes);
        std::memcpy(buf.data(), &h, sizeof h);
        const size_t rest = (size_t)h.total_bytes - sizeof h;
        if (std::fread(buf.data() + sizeof h, 1, rest, f) != rest || std::fgetc(f) != EOF)
        {
            set_error(std::string("snch_scene_load: '") + path + "' is truncated or has trailing bytes");
            rc = SNCH_ERR_INVALID;
        }
    }
    std::fclose(f);
    if (rc != SNCH_OK) return rc;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
    {
        cudaGetLastError();
        set_error("snch_scene_load: no such CUDA device (this library has no CPU fallback)");
        return SNCH_ERR_CUDA;
    }
    return adopt_from(buf.data(), true, buf.size(), device, (cudaStream_t)stream, out, "snch_scene_load");
}

int snch_scene_update_vertices(snch_scene *s, const float *xyz, snch_stream stream)
{
    if (!s || s->adopted || (!xyz && s->n_verts))
    {
        set_error("snch_scene_update_vertices: invalid scene or null vertices");
        return SNCH_ERR_INVALID;
    }
    const size_t bytes = (size_t)s->n_verts * 12;
    if (bytes == 0) return SNCH_OK;
    if (ptr_kind(xyz) == PK_DEVICE)
    {
        SNCH_CUDA(cudaSetDevice(s->device));
        SNCH_CUDA(cudaMemcpyAsync(s->h_xyz.data(), xyz, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        SNCH_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
    else std::memcpy(s->h_xyz.data(), xyz, bytes);
    return SNCH_OK;
}

int snch_wost_step_batch(const snch_scene *cs, const snch_wost_io *io, uint64_t n, snch_stream stream)
{
    int st = check_built(cs);
    if (st != SNCH_OK) return st;
    if (!io || io->struct_size != sizeof(snch_wost_io))
    {
        set_error("snch_wost_step_batch: null io or snch_wost_io.struct_size mismatch");
        return SNCH_ERR_INVALID;
    }
    if (n == 0) return SNCH_OK;
    if (!io->points_xyz)
    {
        set_error("snch_wost_step_batch: points_xyz is required");
        return SNCH_ERR_INVALID;
    }
    snch_scene *s = const_cast<snch_scene *>(cs);
    SNCH_CUDA(cudaSetDevice(s->device));
    const PtrKind k = common_kind({io->points_xyz, io->flip, io->dirs_xyz, io->rnd_uvw, io->closest_index, io->closest_distance,
                                   io->silhouette_distance, io->star_radius, io->hits, io->found, io->sample_index, io->sample_pdf,
                                   io->sample_point_xyz, io->silhouette_edge, io->silhouette_point_xyz});
    if (k == PK_NULL)
    {
        set_error("snch_wost_step_batch: mixed host/device pointers");
        return SNCH_ERR_POINTER_KIND;
    }
    cudaStream_t cst = (cudaStream_t)stream;
    for (uint64_t off = 0; off < n; off += kMaxLaunch)
    {
        const uint64_t m = n - off < kMaxLaunch ? n - off : kMaxLaunch;
        const uint64_t cm = k == PK_DEVICE ? m : host_chunk(s, m);
        const uint64_t qs = wost_scratch_bytes(cm, s->tuning);
        const uint64_t stage = k == PK_DEVICE ? 0 : 5 * Stager::pad(m * 12) + 2 * Stager::pad(m) + 7 * Stager::pad(m * 4) + Stager::pad(m * 16);
        PoolBuffer buf(s, cst, k == PK_DEVICE ? qs : Stager::kLanes * qs + stage);
        if (buf.status != SNCH_OK) return buf.status;
        WostBuffers w;
        auto at = [&](auto *p, uint64_t stride) { return p ? p + stride * off : p; };
        if (k == PK_DEVICE)
        {
            w.points = at(io->points_xyz, 3);
            w.flip = at(io->flip, 1);
            w.dirs = at(io->dirs_xyz, 3);
            w.rnd = at(io->rnd_uvw, 3);
            w.closest_index = at(io->closest_index, 1);
            w.closest_distance = at(io->closest_distance, 1);
            w.silhouette_distance = at(io->silhouette_distance, 1);
            w.star_radius = at(io->star_radius, 1);
            w.hits = at(io->hits, 1);
            w.found = at(io->found, 1);
            w.sample_index = at(io->sample_index, 1);
            w.sample_pdf = at(io->sample_pdf, 1);
            w.sample_point = at(io->sample_point_xyz, 3);
            w.silhouette_edge = at(io->silhouette_edge, 1);
            w.silhouette_point = at(io->silhouette_point_xyz, 3);
            st = launch_wost_step(s->view, s->tuning, w, m, buf.p, cst, &s->counters);
            if (st != SNCH_OK) return st;
            continue;
        }
        Stager sg(s, buf.p + Stager::kLanes * qs, cst, m);
        w.points = sg.in(at(io->points_xyz, 3), 12);
        w.flip = sg.in(at(io->flip, 1), 1);
        w.dirs = sg.in(at(io->dirs_xyz, 3), 12);
        w.rnd = sg.in(at(io->rnd_uvw, 3), 12);
        w.closest_index = sg.out(at(io->closest_index, 1), 4);
        w.closest_distance = sg.out(at(io->closest_distance, 1), 4);
        w.silhouette_distance = sg.out(at(io->silhouette_distance, 1), 4);
        w.star_radius = sg.out(at(io->star_radius, 1), 4);
        w.hits = sg.out(at(io->hits, 1), (uint32_t)sizeof(snch_hit));
        w.found = sg.out(at(io->found, 1), 1);
        w.sample_index = sg.out(at(io->sample_index, 1), 4);
        w.sample_pdf = sg.out(at(io->sample_pdf, 1), 4);
        w.sample_point = sg.out(at(io->sample_point_xyz, 3), 12);
        w.silhouette_edge = sg.out(at(io->silhouette_edge, 1), 4);
        w.silhouette_point = sg.out(at(io->silhouette_point_xyz, 3), 12);
        st = sg.run(cm, [&](uint64_t o, uint64_t c, cudaStream_t ls, int lane) {
            WostBuffers wc;
            auto sl = [&](auto *p, uint64_t stride) { return p ? p + stride * o : p; };
            wc.points = sl(w.points, 3);
            wc.flip = sl(w.flip, 1);
            wc.dirs = sl(w.dirs, 3);
            wc.rnd = sl(w.rnd, 3);
            wc.closest_index = sl(w.closest_index, 1);
            wc.closest_distance = sl(w.closest_distance, 1);
            wc.silhouette_distance = sl(w.silhouette_distance, 1);
            wc.star_radius = sl(w.star_radius, 1);
            wc.hits = sl(w.hits, 1);
            wc.found = sl(w.found, 1);
            wc.sample_index = sl(w.sample_index, 1);
            wc.sample_pdf = sl(w.sample_pdf, 1);
            wc.sample_point = sl(w.sample_point, 3);
            wc.silhouette_edge = sl(w.silhouette_edge, 1);
            wc.silhouette_point = sl(w.silhouette_point, 3);
            return launch_wost_step(s->view, s->tuning, wc, c, buf.p + lane * qs, ls, &s->counters);
        });
        if (st != SNCH_OK) return st;
    }
    return SNCH_OK;
}

} // extern "C"
