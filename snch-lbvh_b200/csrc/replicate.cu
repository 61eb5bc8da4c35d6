// replicate.cu — multi-GPU replication of a built scene behind the C-ABI (include/snch_b200.h "Replication").
//
// The built scene is ONE pointer-free arena, so replicating it is a byte copy plus the pointer patch of adopt_end():
//   * one process, several GPUs: snch_scene_replicate_local — cudaMemcpyPeerAsync fan-out over NVLink (peer access enabled
//     where the topology allows it), one stream per destination so the copies overlap;
//   * one process per GPU: snch_scene_broadcast — ncclBroadcast on the caller's (or a library-made) communicator, received
//     DIRECTLY into the replica's own arena allocation (no staging buffer, no second copy).
// NCCL is reached through libnccl.so.2 itself — resolved with dlopen on first use so that single-GPU users (and boxes without
// NCCL) load libsnch_b200.so without it; no PyTorch on this path.  In a process that already runs torch.distributed the
// soname resolves to the copy torch loaded, so both share one NCCL.
#include "scene.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <new>
#include <vector>

struct snch_comm
{
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    bool owned = false; // created here (destroyed by snch_comm_destroy) vs adopted from the caller
};

namespace snch
{
namespace
{
struct NcclApi
{
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};
NcclApi *nccl_api()
{
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (api.lib) return &api;
    const char *env = std::getenv("SNCH_NCCL_LIB");
    void *h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h)
    {
        set_error(std::string("libnccl.so.2 not found (set SNCH_NCCL_LIB): ") + (dlerror() ? dlerror() : ""));
        return nullptr;
    }
    NcclApi a;
    a.lib = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(h, "ncclBroadcast"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast || !a.AllReduce || !a.GetErrorString)
    {
        set_error("libnccl.so.2 lacks a required symbol");
        return nullptr;
    }
    api = a;
    return &api;
}
int nccl_fail(NcclApi *api, ncclResult_t r, const char *what)
{
    set_error(std::string("NCCL error '") + api->GetErrorString(r) + "' in " + what);
    return SNCH_ERR_CUDA;
}
#define SNCH_NCCL(api, call)                                        \
    do                                                              \
    {                                                               \
        const ncclResult_t r__ = (call);                            \
        if (r__ != ncclSuccess) return nccl_fail(api, r__, #call);  \
    } while (0)
} // namespace
} // namespace snch

using namespace snch;

extern "C"
{

int snch_comm_unique_id(void *id_out, uint64_t bytes)
{
    if (!id_out || bytes < sizeof(ncclUniqueId))
    {
        set_error("snch_comm_unique_id: need a buffer of at least SNCH_COMM_ID_BYTES bytes");
        return SNCH_ERR_INVALID;
    }
    NcclApi *api = nccl_api();
    if (!api) return SNCH_ERR_CUDA;
    ncclUniqueId id;
    SNCH_NCCL(api, api->GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof id);
    return SNCH_OK;
}

int snch_comm_create(const void *id, uint64_t bytes, int rank, int world, int device, snch_comm **out)
{
    if (!id || !out || bytes < sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world || device < 0)
    {
        set_error("snch_comm_create: bad argument");
        return SNCH_ERR_INVALID;
    }
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api) return SNCH_ERR_CUDA;
    SNCH_CUDA(cudaSetDevice(device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof uid);
    snch_comm *c = new (std::nothrow) snch_comm();
    if (!c)
    {
        set_error("out of host memory");
        return SNCH_ERR_OOM;
    }
    const ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess)
    {
        delete c;
        return nccl_fail(api, r, "ncclCommInitRank");
    }
    c->rank = rank;
    c->world = world;
    c->device = device;
    c->owned = true;
    *out = c;
    return SNCH_OK;
}

int snch_comm_adopt(void *nccl_comm, int rank, int world, int device, snch_comm **out)
{
    if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world || device < 0)
    {
        set_error("snch_comm_adopt: bad argument");
        return SNCH_ERR_INVALID;
    }
    if (!nccl_api()) return SNCH_ERR_CUDA;
    snch_comm *c = new (std::nothrow) snch_comm();
    if (!c)
    {
        set_error("out of host memory");
        return SNCH_ERR_OOM;
    }
    c->comm = (ncclComm_t)nccl_comm;
    c->rank = rank;
    c->world = world;
    c->device = device;
    *out = c;
    return SNCH_OK;
}

int snch_comm_destroy(snch_comm *c)
{
    if (!c) return SNCH_OK;
    NcclApi *api = nccl_api();
    if (c->owned && c->comm && api)
    {
        cudaSetDevice(c->device);
        api->CommDestroy(c->comm);
    }
    delete c;
    return SNCH_OK;
}

// Root: `scene` is the built scene (stays the root's), *out = NULL.  Other ranks: `scene` = NULL, *out = the replica.
int snch_scene_broadcast(const snch_scene *scene, int root, snch_comm *comm, snch_stream stream, snch_scene **out)
{
    if (!comm || !out || root < 0 || root >= comm->world)
    {
        set_error("snch_scene_broadcast: bad argument");
        return SNCH_ERR_INVALID;
    }
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api) return SNCH_ERR_CUDA;
    const bool is_root = comm->rank == root;
    if (is_root && (!scene || !scene->built))
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    if (is_root && scene->device != comm->device)
    {
        set_error("snch_scene_broadcast: the scene lives on another device than the communicator");
        return SNCH_ERR_INVALID;
    }
    SNCH_CUDA(cudaSetDevice(comm->device));
    cudaStream_t st = (cudaStream_t)stream;
    // 1. the header (the replicas size their arena from it); the root's copy sits at the start of its arena
    ArenaHeader *dh = nullptr;
    if (!is_root) SNCH_CUDA(cudaMalloc(&dh, sizeof(ArenaHeader)));
    const void *hsrc = is_root ? (const void *)scene->arena : (const void *)dh;
    ncclResult_t r = api->Broadcast(hsrc, is_root ? (void *)scene->arena : (void *)dh, sizeof(ArenaHeader), ncclUint8, root, comm->comm, st);
    if (r != ncclSuccess)
    {
        if (dh) cudaFree(dh);
        return nccl_fail(api, r, "ncclBroadcast(header)");
    }
    snch_scene *rep = nullptr;
    int local_rc = SNCH_OK;
    if (!is_root)
    {
        ArenaHeader h;
        cudaError_t e = cudaMemcpyAsync(&h, dh, sizeof h, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) local_rc = cuda_fail(e, "header read-back");
        else local_rc = adopt_begin(h, h.total_bytes, comm->device, "snch_scene_broadcast", &rep);
    }
    // every rank learns whether ALL of them accepted the header and allocated their arena before the big transfer starts: a
    // rank that bailed out alone would leave the others blocked inside the collective
    {
        int *dflag = is_root ? nullptr : reinterpret_cast<int *>(dh);
        if (is_root) SNCH_CUDA(cudaMalloc(&dflag, sizeof(int)));
        int agreed = local_rc;
        cudaError_t e = cudaMemcpyAsync(dflag, &local_rc, sizeof(int), cudaMemcpyHostToDevice, st);
        ncclResult_t rr = ncclSuccess;
        if (e == cudaSuccess) rr = api->AllReduce(dflag, dflag, 1, ncclInt32, ncclMin, comm->comm, st);
        if (e == cudaSuccess && rr == ncclSuccess) e = cudaMemcpyAsync(&agreed, dflag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && rr == ncclSuccess) e = cudaStreamSynchronize(st);
        cudaFree(dflag);
        dh = nullptr;
        if (rr != ncclSuccess || e != cudaSuccess || agreed != SNCH_OK)
        {
            if (rep) snch_scene_destroy(rep);
            if (rr != ncclSuccess) return nccl_fail(api, rr, "ncclAllReduce(status)");
            if (e != cudaSuccess) return cuda_fail(e, "status exchange");
            if (local_rc == SNCH_OK) set_error("snch_scene_broadcast: another rank rejected the arena header or could not allocate its replica");
            return local_rc != SNCH_OK ? local_rc : agreed;
        }
    }
    // 2. the arena, received in place
    const uint64_t bytes = is_root ? scene->arena_bytes : rep->arena_bytes;
    r = api->Broadcast(is_root ? scene->arena : rep->arena, is_root ? scene->arena : rep->arena, bytes, ncclUint8, root, comm->comm, st);
    if (r != ncclSuccess)
    {
        if (rep) snch_scene_destroy(rep);
        return nccl_fail(api, r, "ncclBroadcast(arena)");
    }
    if (is_root)
    {
        SNCH_CUDA(cudaStreamSynchronize(st));
        return SNCH_OK;
    }
    const int rc = adopt_end(rep, st);
    if (rc != SNCH_OK)
    {
        snch_scene_destroy(rep);
        return rc;
    }
    *out = rep;
    return SNCH_OK;
}

// Re-broadcast of an already replicated scene into the replicas' existing arenas (same sizes): what a per-frame geometry
// update costs, and the WARM timing of the exchange step (communicator and allocations exist).
int snch_scene_rebroadcast(snch_scene *scene, int root, snch_comm *comm, snch_stream stream)
{
    if (!comm || !scene || !scene->built || root < 0 || root >= comm->world)
    {
        set_error("snch_scene_rebroadcast: bad argument");
        return SNCH_ERR_INVALID;
    }
    NcclApi *api = nccl_api();
    if (!api) return SNCH_ERR_CUDA;
    SNCH_CUDA(cudaSetDevice(comm->device));
    cudaStream_t st = (cudaStream_t)stream;
    SNCH_NCCL(api, api->Broadcast(scene->arena, scene->arena, scene->arena_bytes, ncclUint8, root, comm->comm, st));
    if (comm->rank != root)
    {
        scene->built = false;
        ArenaHeader h;
        SNCH_CUDA(cudaMemcpyAsync(&h, scene->arena, sizeof h, cudaMemcpyDeviceToHost, st));
        SNCH_CUDA(cudaStreamSynchronize(st));
        if (h.magic != kArenaMagic || h.total_bytes != scene->arena_bytes || h.n_tris != scene->hdr.n_tris || h.n_edges != scene->hdr.n_edges ||
            h.n_verts != scene->hdr.n_verts)
        {
            set_error("snch_scene_rebroadcast: the root's scene has another size than this replica");
            return SNCH_ERR_INVALID;
        }
        scene->hdr = h;
        return adopt_end(scene, st);
    }
    SNCH_CUDA(cudaStreamSynchronize(st));
    return SNCH_OK;
}

int snch_scene_replicate_local(const snch_scene *scene, const int *devices, int n, snch_scene **out)
{
    if (!scene || !scene->built)
    {
        set_error("BVH is not built yet.");
        return SNCH_ERR_NOT_BUILT;
    }
    if (n < 0 || (n && (!devices || !out)))
    {
        set_error("snch_scene_replicate_local: bad argument");
        return SNCH_ERR_INVALID;
    }
    int count = 0;
    SNCH_CUDA(cudaGetDeviceCount(&count));
    for (int i = 0; i < n; ++i) out[i] = nullptr;
    std::vector<cudaStream_t> streams((size_t)n, nullptr);
    int rc = SNCH_OK;
    for (int i = 0; i < n && rc == SNCH_OK; ++i)
    {
        const int d = devices[i];
        if (d < 0 || d >= count)
        {
            set_error("snch_scene_replicate_local: bad device ordinal");
            rc = SNCH_ERR_INVALID;
            break;
        }
        rc = adopt_begin(scene->hdr, scene->arena_bytes, d, "snch_scene_replicate_local", &out[i]); // (sets device d)
        if (rc != SNCH_OK) break;
        if (d != scene->device)
        { // direct NVLink path where the topology allows it; the copy below is staged by the driver otherwise
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d, scene->device) == cudaSuccess && can)
            {
                const cudaError_t e = cudaDeviceEnablePeerAccess(scene->device, 0);
                if (e != cudaSuccess) cudaGetLastError(); // already enabled
            }
        }
        cudaError_t e = cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(out[i]->arena, d, scene->arena, scene->device, scene->arena_bytes, streams[i]);
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyPeerAsync");
    }
    for (int i = 0; i < n; ++i)
    {
        if (!out[i]) continue;
        cudaSetDevice(devices[i]);
        if (rc == SNCH_OK && streams[i]) rc = adopt_end(out[i], streams[i]);
        if (streams[i])
        {
            cudaStreamSynchronize(streams[i]);
            cudaStreamDestroy(streams[i]);
        }
    }
    if (rc != SNCH_OK)
        for (int i = 0; i < n; ++i)
            if (out[i])
            {
                snch_scene_destroy(out[i]);
                out[i] = nullptr;
            }
    cudaSetDevice(scene->device);
    return rc;
}

} // extern "C"
