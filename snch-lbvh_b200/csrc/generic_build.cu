// generic_build.cu — LBVH + SNCH construction from caller-supplied LEAF boxes and cones (2-D or 3-D, float).
//
// This is what stands behind the generic class template lbvh::bvh<Real, dim, Object, AABBGetter, ConeGetter, MortonCalc>
// of the drop-in headers (include/snch_lbvh/core/bvh.cuh): the header evaluates the user's getters per object in a small
// templated kernel (user types cannot cross a C-ABI) and hands the resulting leaf arrays to snch_lbvh_build(), which runs
// the same pipeline as the scene builder — scene box, Morton codes, radix sort of (key, index), Karras hierarchy on the
// augmented key, one bottom-up climb that merges boxes AND cones — and fills the reference-layout arrays
// nodes / aabbs / cones of lbvh::bvh_device (bvh.cuh:27-54).  Replaces bvh::construct(), bvh.cuh:380-613, for any Object.
#include "build_ctx.h"
#include "scene.h"
#include "sort_scan.cuh"

#include "../../include/snch_lbvh/core/cone.cuh"
#include "../../include/snch_lbvh/core/morton_code.cuh"

namespace snch
{
namespace
{
template <int D> using BoxT = lbvh::aabb<float, D>;
template <int D> using ConeT = lbvh::cone<float, D>;

__device__ __forceinline__ int ford(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordf(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

template <int D> __global__ void k_gen_box_init(int *box)
{
    if (threadIdx.x < D) box[threadIdx.x] = ford(INFINITY);            // lower
    else if (threadIdx.x < 2 * D) box[threadIdx.x] = ford(-INFINITY); // upper
}
// scene box = merge of all leaf boxes (bvh.cuh:429-432); exact (min/max), so the reduction order is irrelevant
template <int D> __global__ void __launch_bounds__(256) k_gen_box(const BoxT<D> *__restrict__ leaf, uint32_t n, int *box)
{
    float lo[D], hi[D];
    for (int a = 0; a < D; ++a)
    {
        lo[a] = INFINITY;
        hi[a] = -INFINITY;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const BoxT<D> b = leaf[i];
        for (int a = 0; a < D; ++a)
        {
            lo[a] = fminf(lo[a], lbvh::detail::at(b.lower, a));
            hi[a] = fmaxf(hi[a], lbvh::detail::at(b.upper, a));
        }
    }
    for (int a = 0; a < D; ++a)
        for (int o = 16; o > 0; o >>= 1)
        {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0)
        for (int a = 0; a < D; ++a)
        {
            atomicMin(box + a, ford(lo[a]));
            atomicMax(box + D + a, ford(hi[a]));
        }
}
// default_morton_code_calculator (bvh.cuh:232-304): code of the box centroid normalised by the scene box, IEEE division
template <int D>
__global__ void __launch_bounds__(256)
    k_gen_morton(const BoxT<D> *__restrict__ leaf, uint32_t n, const int *__restrict__ box, const uint32_t *__restrict__ user_codes,
                 uint32_t *__restrict__ keys, uint32_t *__restrict__ idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    idx[i] = i;
    if (user_codes)
    {
        keys[i] = user_codes[i];
        return;
    }
    const BoxT<D> b = leaf[i];
    lbvh::vector_of_t<float, D> p;
    for (int a = 0; a < D; ++a)
    {
        const float c = __fmul_rn(__fadd_rn(lbvh::detail::at(b.upper, a), lbvh::detail::at(b.lower, a)), 0.5f);
        const float wl = ordf(box[a]), wu = ordf(box[D + a]);
        lbvh::detail::at(p, a) = __fdiv_rn(__fsub_rn(c, wl), __fsub_rn(wu, wl));
    }
    keys[i] = lbvh::morton_code(p);
}
// leaves in Morton order: record N-1+k holds object sorted_idx[k] (bvh.cuh:489-499); parent links come from k_hierarchy
template <int D>
__global__ void __launch_bounds__(256)
    k_gen_leaves(const BoxT<D> *__restrict__ leaf_box, const ConeT<D> *__restrict__ leaf_cone, const uint32_t *__restrict__ sorted_idx,
                 uint32_t n, RefNode *nodes, BoxT<D> *aabbs, ConeT<D> *cones)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t obj = sorted_idx[k];
    const uint32_t at = n - 1 + k;
    aabbs[at] = leaf_box[obj];
    cones[at] = leaf_cone[obj];
    nodes[at].left = kNone;
    nodes[at].right = kNone;
    nodes[at].object = obj;
    if (n == 1) nodes[at].parent = kNone;
}
// bottom-up refit of boxes and cones in one climb (bvh.cuh:520-604); the second arriver at a node merges its children
template <int D>
__global__ void __launch_bounds__(128) k_gen_refit(uint32_t n, const RefNode *__restrict__ nodes, BoxT<D> *aabbs, ConeT<D> *cones, uint32_t *flags)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t cur = n - 1 + k;
    uint32_t parent = nodes[cur].parent;
    while (parent != kNone)
    {
        __threadfence();
        if (atomicAdd(flags + parent, 1u) == 0) return;
        __threadfence();
        const uint32_t l = nodes[parent].left, r = nodes[parent].right;
        // children were written by other threads: read them past L1
        BoxT<D> lb, rb;
        ConeT<D> lc, rc;
        {
            const float *pl = reinterpret_cast<const float *>(aabbs + l), *pr = reinterpret_cast<const float *>(aabbs + r);
            float *dl = reinterpret_cast<float *>(&lb), *dr = reinterpret_cast<float *>(&rb);
            for (int i = 0; i < 2 * D; ++i)
            {
                dl[i] = __ldcg(pl + i);
                dr[i] = __ldcg(pr + i);
            }
            const float *ql = reinterpret_cast<const float *>(cones + l), *qr = reinterpret_cast<const float *>(cones + r);
            float *el = reinterpret_cast<float *>(&lc), *er = reinterpret_cast<float *>(&rc);
            for (int i = 0; i < D + 2; ++i)
            {
                el[i] = __ldcg(ql + i);
                er[i] = __ldcg(qr + i);
            }
        }
        const BoxT<D> pb = lbvh::merge(lb, rb);
        aabbs[parent] = pb;
        cones[parent] = lbvh::merge(lc, rc, lbvh::centroid(lb), lbvh::centroid(rb), lbvh::centroid(pb));
        cur = parent;
        parent = nodes[cur].parent;
    }
}

struct AsyncBuf
{
    void *p = nullptr;
    cudaStream_t st;
    explicit AsyncBuf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 16, st); }
    ~AsyncBuf()
    {
        if (p) cudaFreeAsync(p, st);
    }
};

template <int D>
int build_generic(uint32_t n, const void *leaf_aabbs, const void *leaf_cones, const uint32_t *user_codes, void *nodes, void *aabbs,
                  void *cones, uint32_t *sorted_index_out, uint32_t *morton_out, int *collision_out, cudaStream_t st)
{
    if (collision_out) *collision_out = 0;
    if (n == 0) return SNCH_OK; // bvh.cuh:383-386: nothing to build
    const uint64_t a4 = align_up((uint64_t)n * 4, 256);
    const uint64_t sort_bytes = align_up(sort_scratch_elems(n) * 4, 256);
    AsyncBuf buf(st);
    // keys | idx | keys tmp | idx tmp | flags | ranges | sort scratch | box + counters
    const uint64_t total = 5 * a4 + align_up((uint64_t)n * 8, 256) + sort_bytes + 512;
    if (buf.alloc(total) != cudaSuccess)
    {
        cudaGetLastError();
        set_error("snch_lbvh_build: out of device memory for the build scratch");
        return SNCH_ERR_OOM;
    }
    unsigned char *b = (unsigned char *)buf.p;
    uint32_t *keys = (uint32_t *)b, *idx = (uint32_t *)(b + a4), *ktmp = (uint32_t *)(b + 2 * a4), *vtmp = (uint32_t *)(b + 3 * a4);
    uint32_t *flags = (uint32_t *)(b + 4 * a4);
    uint2 *ranges = (uint2 *)(b + 5 * a4);
    uint32_t *sscr = (uint32_t *)(b + 5 * a4 + align_up((uint64_t)n * 8, 256));
    int *box = (int *)(b + total - 512);
    uint32_t *counters = (uint32_t *)(b + total - 256);
    SNCH_CUDA(cudaMemsetAsync(flags, 0, (size_t)n * 4, st));
    SNCH_CUDA(cudaMemsetAsync(counters, 0, 64, st));
    const unsigned g = (n + 255) / 256;
    const BoxT<D> *lb = (const BoxT<D> *)leaf_aabbs;
    const ConeT<D> *lc = (const ConeT<D> *)leaf_cones;
    k_gen_box_init<D><<<1, 32, 0, st>>>(box);
    k_gen_box<D><<<g < 1184 ? g : 1184, 256, 0, st>>>(lb, n, box);
    k_gen_morton<D><<<g, 256, 0, st>>>(lb, n, box, user_codes, keys, idx);
    radix_sort_pairs(keys, idx, ktmp, vtmp, n, user_codes ? 32 : 30, sscr, st); // 2-D codes interleave by 3 too (bits 0..28)
    k_gen_leaves<D><<<g, 256, 0, st>>>(lb, lc, idx, n, (RefNode *)nodes, (BoxT<D> *)aabbs, (ConeT<D> *)cones);
    if (n > 1)
    {
        BuildCtx c{};
        c.n = n;
        c.morton = keys;
        c.sorted_idx = idx;
        c.nodes = (RefNode *)nodes;
        c.ranges = ranges;
        c.counters = counters;
        k_hierarchy<<<(n - 1 + 255) / 256, 256, 0, st>>>(c);
        k_gen_refit<D><<<(n + 127) / 128, 128, 0, st>>>(n, (const RefNode *)nodes, (BoxT<D> *)aabbs, (ConeT<D> *)cones, flags);
    }
    SNCH_CUDA(cudaGetLastError());
    if (sorted_index_out) SNCH_CUDA(cudaMemcpyAsync(sorted_index_out, idx, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    if (morton_out) SNCH_CUDA(cudaMemcpyAsync(morton_out, keys, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    if (collision_out)
    {
        uint32_t h = 0;
        SNCH_CUDA(cudaMemcpyAsync(&h, counters, 4, cudaMemcpyDeviceToHost, st));
        SNCH_CUDA(cudaStreamSynchronize(st));
        *collision_out = (int)h;
    }
    return SNCH_OK;
}
} // namespace
} // namespace snch

extern "C" int snch_lbvh_build(int dim, uint32_t n, const void *leaf_aabbs, const void *leaf_cones, const uint32_t *morton_codes, void *nodes,
                               void *aabbs, void *cones, uint32_t *sorted_index_out, uint32_t *morton_sorted_out, int *collision_out,
                               snch_stream stream)
{
    using namespace snch;
    if ((dim != 2 && dim != 3) || (n && (!leaf_aabbs || !leaf_cones || !nodes || !aabbs || !cones)))
    {
        set_error("snch_lbvh_build: dim must be 2 or 3 and the leaf / output arrays non-null");
        return SNCH_ERR_INVALID;
    }
    if (n > 0x7FFFFFFFu / 2)
    {
        set_error("snch_lbvh_build: too many objects for 32-bit node ids");
        return SNCH_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return SNCH_ERR_CUDA;
    }
    return dim == 2 ? build_generic<2>(n, leaf_aabbs, leaf_cones, morton_codes, nodes, aabbs, cones, sorted_index_out, morton_sorted_out,
                                       collision_out, (cudaStream_t)stream)
                    : build_generic<3>(n, leaf_aabbs, leaf_cones, morton_codes, nodes, aabbs, cones, sorted_index_out, morton_sorted_out,
                                       collision_out, (cudaStream_t)stream);
}

// ---- self-test of the device restatements of the host libm (include/snch_lbvh/core/host_libm.cuh) ------------------------
namespace snch
{
namespace
{
__global__ void k_selftest_libm(int which, const float *__restrict__ x, uint64_t n, float *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    {
        const float v = x[i];
        out[i] = which == 0 ? lbvh::detail::acosf_host(v) : which == 1 ? lbvh::detail::sinf_host(v) : which == 2 ? lbvh::detail::cosf_host(v) : lbvh::detail::logf_host(v);
    }
}
} // namespace
} // namespace snch

extern "C" int snch_selftest_host_libm(int which, const float *x, uint64_t n, float *out, int device)
{
    using namespace snch;
    if (which < 0 || which > 3 || (n && (!x || !out)))
    {
        set_error("snch_selftest_host_libm: which must be 0..3 and the arrays non-null");
        return SNCH_ERR_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
    {
        cudaGetLastError();
        set_error("no CUDA device available (this library has no CPU fallback)");
        return SNCH_ERR_CUDA;
    }
    if (n == 0) return SNCH_OK;
    int prev = 0;
    SNCH_CUDA(cudaGetDevice(&prev));
    SNCH_CUDA(cudaSetDevice(device));
    float *dx = nullptr, *dy = nullptr;
    cudaError_t e = cudaMalloc(&dx, n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dy, n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
    {
        k_selftest_libm<<<1184, 256>>>(which, dx, n, dy);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, dy, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(dx);
    cudaFree(dy);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(e, "snch_selftest_host_libm");
    return SNCH_OK;
}
