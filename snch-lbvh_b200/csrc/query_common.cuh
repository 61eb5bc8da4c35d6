// query_common.cuh — scheduling pieces shared by the batched traversal kernels of query.cu (3-D scenes) and scene2.cu
// (2-D scenes): persistent-lane work distribution, 256-bit record loads, stack entries, batch ordering.
#ifndef SNCH_QUERY_COMMON_CUH
#define SNCH_QUERY_COMMON_CUH
#include "scene.h"
#include "snch_math.cuh"

namespace snch
{

constexpr int kQueryThreads = 128;
constexpr int kStackDepth = 64; // >= 62 levels possible with the 62-bit augmented key
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kChunk = 64; // consecutive queries a warp draws per atomic
// batches under 1M queries draw 32 at a time: twice the warps, one query per lane — such a batch is latency-bound (C1: 64K rays)
SNCH_DI uint32_t chunk_for(uint32_t n) { return n < (1u << 20) ? 32u : kChunk; }
// The per-lane silhouette kernel: batches are ordered LARGEST search radius first, so the first draws hold the longest walks; a
// draw of 64 gives every lane of the first warps two of them back to back — half of the run time of a 2M-query shard (the per-GPU
// share of the C3 batch on 8 GPUs).  Below 12M queries a warp draws 16 at a time: 8.35 -> 7.35 ms at 2M, 13.8 -> 13.2 at 4M,
// unchanged at 8M; at 16.7M draws of 64 stay (48.9 vs 49.3 ms) (profiles/r2h_shard_exp.json).  (Draws that SHRINK towards the end of the
// batch — guided self-scheduling — were measured too and gave nothing: the long walks are at the start of the batch.)
static inline uint32_t sil_chunk_for_host(uint64_t n) { return n < (12u << 20) ? 16u : kChunk; }
static inline uint32_t chunk_for_host(uint64_t n) { return n < (1u << 20) ? 32u : kChunk; }

struct NodeBoxes
{
    V3 lo0, hi0, lo1, hi1;
};
SNCH_DI NodeBoxes unpack_boxes(float4 a, float4 b, float4 c)
{
    NodeBoxes n;
    n.lo0 = V3{a.x, a.y, a.z};
    n.hi0 = V3{a.w, b.x, b.y};
    n.lo1 = V3{b.z, b.w, c.x};
    n.hi1 = V3{c.y, c.z, c.w};
    return n;
}
// One 32-byte sector per instruction (LDG.E.256, sm_100): a divergent warp pays one L1 tag lookup per lane per
// instruction, so a 64 B / 96 B record costs 2 / 3 lookups instead of 4 / 6 with 128-bit loads.  p must be 32 B aligned.
SNCH_DI void ld256(const void *p, float4 &lo, float4 &hi)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}
// traversal stack entry: node reference + the key it was pushed with, moved with one 64-bit local access
struct __align__(8) StackEntry
{
    uint32_t node;
    float key;
};
SNCH_DI V3 load_point(const float *__restrict__ q, uint64_t i) { return V3{__ldg(q + 3 * i), __ldg(q + 3 * i + 1), __ldg(q + 3 * i + 2)}; }

// ---------------------------------------------------------------------------------------------------------------
// warp-level work distribution
// ---------------------------------------------------------------------------------------------------------------
struct Feeder
{
    uint32_t next, end; // warp-uniform: the unclaimed part of the warp's current chunk
    bool exhausted;     // the global counter ran past n
    uint32_t chunk = 0; // queries per draw (0 = chunk_for(n))
};
constexpr uint64_t kScratchHeader = 256; // work counters (u64 x 8: [0] batch, [1] tail list length, [2] tail work) + query box (6 ordered ints at +64)
constexpr uint64_t kTailEntries = 1u << 18; // tail list of the silhouette kernel: one (slot, bound, edge found so far) triple per resident lane at most
constexpr uint64_t kTailBytes = kTailEntries * 12;
// Gives every idle lane (bit set in `idle`) the next query slot of the warp's chunk; draws a new chunk when needed.
// Returns the slot for this lane or kNone.  Warp-convergent call.
SNCH_DI uint32_t feeder_take(Feeder &f, unsigned idle, bool lane_idle, int lane, uint32_t n, unsigned long long *counter)
{
    if (f.next == f.end && !f.exhausted)
    {
        const uint32_t ch = f.chunk ? f.chunk : chunk_for(n);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)ch);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n) f.exhausted = true;
        else
        {
            f.next = (uint32_t)base;
            f.end = (uint32_t)min((unsigned long long)n, base + ch);
        }
    }
    const uint32_t avail = f.end - f.next;
    const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
    const uint32_t want = __popc(idle);
    const uint32_t take = min(avail, want);
    uint32_t slot = kNone;
    if (lane_idle && rank < take) slot = f.next + rank;
    f.next += take;
    return slot;
}

// Lays the per-batch scratch out (query_scratch_bytes) and, for batches worth ordering, produces the Morton permutation of the
// query points (`dims` = 2 or 3 coordinates at pts[stride * i]).  *perm_out = nullptr otherwise.            query.cu
int prepare_batch(const QueryTuning &t, bool order, const float *pts, int stride, const float *radius, uint32_t n, unsigned char *scratch,
                  cudaStream_t st, unsigned long long **counter_out, const uint32_t **perm_out, QueryCounters *qc, int dims = 3,
                  int radius_desc = 0, const float *dirs = nullptr);

// grid of a persistent kernel: SMs x resident CTAs, or fewer when the batch has fewer chunks than that
template <typename K> static inline unsigned persistent_grid(K kernel, const QueryTuning &t, uint32_t n)
{
    static thread_local int cached_dev = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev)
    {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kQueryThreads, 0);
    if (per_sm < 1) per_sm = 1;
    if (t.blocks_per_sm > 0 && t.blocks_per_sm < per_sm) per_sm = t.blocks_per_sm;
    const uint64_t full = (uint64_t)sms * per_sm;
    const uint64_t per_cta = (uint64_t)chunk_for_host(n) * (kQueryThreads / 32);
    const uint64_t need = ((uint64_t)n + per_cta - 1) / per_cta;
    return (unsigned)(need < full ? (need ? need : 1) : full);
}

} // namespace snch
#endif // SNCH_QUERY_COMMON_CUH
