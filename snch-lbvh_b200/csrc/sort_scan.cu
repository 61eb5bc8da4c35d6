// sort_scan.cu — see sort_scan.cuh
#include "sort_scan.cuh"

namespace snch
{

// ---------------------------------------------------------------------------------------------------------------
// exclusive scan
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        const uint32_t w = warp_sums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - w;
        if (lane == 31) block_total = wi;
    }
    __syncthreads();
    const uint32_t res = incl - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const uint32_t *__restrict__ in, uint64_t n, uint32_t *__restrict__ tile_sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
        if (base + j < n) s += in[base + j];
    uint32_t total;
    block_exclusive_scan_1024(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
    k_scan_apply(const uint32_t *in, uint32_t *out, uint64_t n, const uint32_t *__restrict__ tile_offsets)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t a[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
        a[j] = (base + j < n) ? in[base + j] : 0u;
        s += a[j];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan_1024(s, &total) + (tile_offsets ? tile_offsets[blockIdx.x] : 0u);
#pragma unroll
    for (int j = 0; j < kScanItems; ++j)
    {
        if (base + j < n) out[base + j] = run;
        run += a[j];
    }
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *scratch, cudaStream_t stream)
{
    if (n == 0) return 0;
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles == 1)
    {
        k_scan_apply<<<1, kScanThreads, 0, stream>>>(in, out, n, nullptr);
        return 1;
    }
    uint32_t *tile_sums = scratch;
    k_scan_reduce<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, n, tile_sums);
    const int inner = exclusive_scan_u32(tile_sums, tile_sums, tiles, scratch + tiles + 1, stream);
    k_scan_apply<<<(unsigned)tiles, kScanThreads, 0, stream>>>(in, out, n, tile_sums);
    return inner + 2;
}

// ---------------------------------------------------------------------------------------------------------------
// LSD radix sort, 8 bits per pass: histogram -> scan -> stable scatter
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
    k_radix_hist(const uint32_t *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ counts, uint32_t tiles)
{
    __shared__ uint32_t h[kSortBins];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int it = 0; it < kSortItems; ++it)
    {
        const uint64_t idx = base + (uint64_t)it * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & (kSortBins - 1)], 1u);
    }
    __syncthreads();
    counts[(uint64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x]; // digit-major
}

__global__ void __launch_bounds__(kSortThreads)
    k_radix_scatter(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
                    uint32_t *__restrict__ vout, uint64_t n, int shift, const uint32_t *__restrict__ offsets, uint32_t tiles)
{
    constexpr int kWarps = kSortThreads / 32;
    __shared__ uint32_t wc[kWarps][kSortBins];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kWarps * kSortBins; i += kSortThreads) (&wc[0][0])[i] = 0;
    __syncthreads();

    // each warp owns a contiguous 512-key slice of the tile; item `it` of lane l is key it*32+l of that slice
    const uint64_t wbase = (uint64_t)blockIdx.x * kSortTile + (uint64_t)warp * (32 * kSortItems);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t key[kSortItems];
    uint32_t rank[kSortItems];
#pragma unroll
    for (int it = 0; it < kSortItems; ++it)
    {
        const uint64_t idx = wbase + (uint64_t)it * 32 + lane;
        const bool valid = idx < n;
        key[it] = valid ? kin[idx] : 0xFFFFFFFFu;
        const uint32_t d = valid ? ((key[it] >> shift) & (kSortBins - 1)) : (uint32_t)(kSortBins + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid)
        {
            old = wc[warp][d];
            wc[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[it] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d turns the per-warp counts of digit d into per-warp global bases
        const int d = threadIdx.x;
        uint32_t run = offsets[(uint64_t)d * tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
        {
            const uint32_t c = wc[w][d];
            wc[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kSortItems; ++it)
    {
        const uint64_t idx = wbase + (uint64_t)it * 32 + lane;
        if (idx < n)
        {
            const uint32_t d = (key[it] >> shift) & (kSortBins - 1);
            const uint32_t dst = wc[warp][d] + rank[it];
            kout[dst] = key[it];
            vout[dst] = vin[idx];
        }
    }
}

static int radix_sort_pairs_multipass(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int bits,
                                      uint32_t *scratch, cudaStream_t stream, int first_bit)
{
    if (n == 0) return 0;
    int launches = 0;
    const uint32_t tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    uint32_t *counts = scratch;
    uint32_t *scan_scratch = scratch + (uint64_t)tiles * kSortBins;
    const int passes = (bits + 7) / 8;
    uint32_t *sk = keys, *sv = vals, *dk = keys_tmp, *dv = vals_tmp;
    for (int p = 0; p < passes; ++p)
    {
        const int shift = first_bit + 8 * p;
        k_radix_hist<<<tiles, kSortThreads, 0, stream>>>(sk, n, shift, counts, tiles);
        launches += 2 + exclusive_scan_u32(counts, counts, (uint64_t)tiles * kSortBins, scan_scratch, stream);
        k_radix_scatter<<<tiles, kSortThreads, 0, stream>>>(sk, sv, dk, dv, n, shift, counts, tiles);
        uint32_t *t = sk;
        sk = dk;
        dk = t;
        t = sv;
        sv = dv;
        dv = t;
    }
    if (sk != keys)
    {
        cudaMemcpyAsync(keys, sk, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
        cudaMemcpyAsync(vals, sv, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    }
    return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// Onesweep LSD radix sort (Adinets & Merrill 2022): ONE histogram kernel for all passes, then one kernel per pass in which
// every tile ranks its keys, publishes its digit counts and obtains its global offsets by decoupled look-back over the
// preceding tiles — no per-pass histogram / scan launches and no second read of the keys.  5 launches for 4 passes instead
// of 20; stable (tile ids are taken in launch order, ranks preserve input order).
// ---------------------------------------------------------------------------------------------------------------
constexpr uint32_t kOsAgg = 1u << 30, kOsPrefix = 2u << 30, kOsValue = (1u << 30) - 1u;
constexpr int kOsMaxPasses = 4;
constexpr int kOsWindow = 8;

__global__ void __launch_bounds__(kSortThreads)
    k_os_hist(const uint32_t *__restrict__ keys, uint64_t n, int first_bit, int passes, uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t h[kOsMaxPasses][kSortBins];
    for (int p = 0; p < kOsMaxPasses; ++p) h[p][threadIdx.x] = 0;
    __syncthreads();
    // a thread walks 16 CONSECUTIVE keys and adds each run of equal digits once: mesh-ordered Morton keys share their high
    // digits with their neighbours (one shared-memory atomic per run instead of 32 colliding ones per warp)
    for (uint64_t base = ((uint64_t)blockIdx.x * kSortThreads + threadIdx.x) * kSortItems; base < n;
         base += (uint64_t)gridDim.x * kSortTile)
    {
        uint32_t k[kSortItems];
        const bool full = base + kSortItems <= n;
        if (full)
        {
#pragma unroll
            for (int j = 0; j < kSortItems / 4; ++j)
            {
                const uint4 v = reinterpret_cast<const uint4 *>(keys + base)[j];
                k[4 * j] = v.x;
                k[4 * j + 1] = v.y;
                k[4 * j + 2] = v.z;
                k[4 * j + 3] = v.w;
            }
        }
        else
        {
#pragma unroll
            for (int j = 0; j < kSortItems; ++j) k[j] = base + j < n ? keys[base + j] : 0u;
        }
        const int cnt = full ? kSortItems : (int)(n - base);
        for (int p = 0; p < passes; ++p)
        {
            const int shift = first_bit + 8 * p;
            uint32_t prev = (k[0] >> shift) & (kSortBins - 1), run = 0;
#pragma unroll
            for (int j = 0; j < kSortItems; ++j)
            {
                if (j >= cnt) break;
                const uint32_t d = (k[j] >> shift) & (kSortBins - 1);
                if (d != prev)
                {
                    atomicAdd(&h[p][prev], run);
                    prev = d;
                    run = 0;
                }
                ++run;
            }
            if (run) atomicAdd(&h[p][prev], run);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p)
    {
        const uint32_t c = h[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * kSortBins + threadIdx.x], c);
    }
}

// exclusive scan of one value per thread over the 256-thread CTA
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *warp_sums /* [8] shared */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w)
        if (w < warp) before += warp_sums[w];
    __syncthreads();
    return incl - v + before;
}

static int g_sort_onesweep = 1, g_sort_lookback = 8;
void set_sort_onesweep(int on) { g_sort_onesweep = on; }
void set_sort_lookback(int window) { g_sort_lookback = window; }

// kItems keys per thread: 16 (4096-key tiles) for large inputs, 8 when that would leave SMs without a tile
template <int kItems>
__global__ void __launch_bounds__(kSortThreads)
    k_os_pass(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout, uint32_t *__restrict__ vout,
              uint64_t n, int shift, const uint32_t *__restrict__ ghist, volatile uint32_t *status, uint32_t *tile_counter, int lookback)
{
    constexpr int kWarps = kSortThreads / 32;
    constexpr int kTile = kSortThreads * kItems;
    __shared__ uint32_t wc[kWarps][kSortBins];
    __shared__ uint32_t s_keys[kTile], s_vals[kTile];
    __shared__ uint32_t s_dstart[kSortBins], s_gbase[kSortBins];
    __shared__ uint32_t s_warp_sums[kWarps];
    __shared__ uint32_t s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u); // tiles are taken in start order: every predecessor is running
    for (int i = threadIdx.x; i < kWarps * kSortBins; i += kSortThreads) (&wc[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tbase = (uint64_t)tile * kTile;
    const uint32_t tcount = (uint32_t)(n - tbase < (uint64_t)kTile ? n - tbase : (uint64_t)kTile);

    // each warp owns a contiguous (32 * kItems)-key slice of the tile; item `it` of lane l is key it*32+l of that slice
    const uint64_t wbase = tbase + (uint64_t)warp * (32 * kItems);
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t key[kItems];
    uint32_t rank[kItems];
#pragma unroll
    for (int it = 0; it < kItems; ++it)
    {
        const uint64_t idx = wbase + (uint64_t)it * 32 + lane;
        const bool valid = idx < n;
        key[it] = valid ? kin[idx] : 0xFFFFFFFFu;
        const uint32_t d = valid ? ((key[it] >> shift) & (kSortBins - 1)) : (uint32_t)(kSortBins + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (lane == leader && valid)
        {
            old = wc[warp][d];
            wc[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[it] = old + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    {
        // thread d: per-warp counts of digit d -> per-warp offsets inside the digit; tile count -> published for look-back
        const int d = threadIdx.x;
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
        {
            const uint32_t c = wc[w][d];
            wc[w][d] = cnt;
            cnt += c;
        }
        volatile uint32_t *mine = status + (uint64_t)tile * kSortBins + d;
        if (tile > 0) *mine = cnt | kOsAgg;
        const uint32_t dstart = block_exclusive_scan_256(cnt, s_warp_sums);          // where digit d starts inside the tile
        const uint32_t gdigit = block_exclusive_scan_256(ghist[d], s_warp_sums);     // where digit d starts in the output
        uint32_t excl = 0;
        if (lookback <= 1)
        {
            for (int64_t t = (int64_t)tile - 1; t >= 0; --t)
            {
                uint32_t v;
                do
                {
                    v = status[(uint64_t)t * kSortBins + d];
                } while ((v & ~kOsValue) == 0);
                excl += v & kOsValue;
                if (v & kOsPrefix) break;
            }
        }
        else
        { // kOsWindow predecessors are read at once (independent loads in flight), then consumed in order: the walk back to
          // the nearest published prefix costs one L2 round trip per window instead of one per tile
            bool done = tile == 0;
            for (int64_t t = (int64_t)tile - 1; !done; t -= kOsWindow)
            {
                uint32_t v[kOsWindow];
#pragma unroll
                for (int j = 0; j < kOsWindow; ++j) v[j] = t - j >= 0 ? status[(uint64_t)(t - j) * kSortBins + d] : (2u << 30) /* prefix of nothing */;
#pragma unroll
                for (int j = 0; j < kOsWindow; ++j)
                {
                    if (done) continue;
                    uint32_t x = v[j];
                    while ((x & ~kOsValue) == 0) x = status[(uint64_t)(t - j) * kSortBins + d];
                    excl += x & kOsValue;
                    done = (x & kOsPrefix) != 0;
                }
            }
        }
        *mine = (excl + cnt) | kOsPrefix;
        s_dstart[d] = dstart;
        s_gbase[d] = gdigit + excl - dstart;
    }
    __syncthreads();
    // reorder through shared memory so that each digit's run leaves the CTA as one contiguous, coalesced store
#pragma unroll
    for (int it = 0; it < kItems; ++it)
    {
        const uint64_t idx = wbase + (uint64_t)it * 32 + lane;
        if (idx < n)
        {
            const uint32_t d = (key[it] >> shift) & (kSortBins - 1);
            const uint32_t pos = s_dstart[d] + wc[warp][d] + rank[it];
            s_keys[pos] = key[it];
            s_vals[pos] = vin[idx];
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < tcount; i += kSortThreads)
    {
        const uint32_t k = s_keys[i];
        const uint32_t dst = s_gbase[(k >> shift) & (kSortBins - 1)] + i;
        kout[dst] = k;
        vout[dst] = s_vals[i];
    }
}


int radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int bits,
                     uint32_t *scratch, cudaStream_t stream, int first_bit)
{
    if (n == 0) return 0;
    const int passes = (bits + 7) / 8;
    if (!g_sort_onesweep || n >= (1ull << 30) || passes > kOsMaxPasses)
        return radix_sort_pairs_multipass(keys, vals, keys_tmp, vals_tmp, n, bits, scratch, stream, first_bit);
    const bool small_tiles = (n + kSortTile - 1) / kSortTile < 1024; // fewer 4096-key tiles than ~2 waves of CTAs: halve them
    const uint32_t tile_keys = small_tiles ? kSortTile / 2 : kSortTile;
    const uint32_t tiles = (uint32_t)((n + tile_keys - 1) / tile_keys);
    // scratch: [passes][256] global histograms | [passes] tile counters (padded to 8) | [passes][tiles][256] tile status
    uint32_t *ghist = scratch;
    uint32_t *counters = scratch + kOsMaxPasses * kSortBins;
    uint32_t *status = counters + 8;
    cudaMemsetAsync(scratch, 0, ((uint64_t)kOsMaxPasses * kSortBins + 8 + (uint64_t)passes * tiles * kSortBins) * sizeof(uint32_t), stream);
    const uint32_t hist_tiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
    k_os_hist<<<hist_tiles < 592u ? hist_tiles : 592u, kSortThreads, 0, stream>>>(keys, n, first_bit, passes, ghist);
    uint32_t *sk = keys, *sv = vals, *dk = keys_tmp, *dv = vals_tmp;
    for (int p = 0; p < passes; ++p)
    {
        if (small_tiles)
            k_os_pass<kSortItems / 2><<<tiles, kSortThreads, 0, stream>>>(sk, sv, dk, dv, n, first_bit + 8 * p, ghist + p * kSortBins,
                                                                        status + (uint64_t)p * tiles * kSortBins, counters + p, g_sort_lookback);
        else
            k_os_pass<kSortItems><<<tiles, kSortThreads, 0, stream>>>(sk, sv, dk, dv, n, first_bit + 8 * p, ghist + p * kSortBins,
                                                                    status + (uint64_t)p * tiles * kSortBins, counters + p, g_sort_lookback);
        uint32_t *t = sk;
        sk = dk;
        dk = t;
        t = sv;
        sv = dv;
        dv = t;
    }
    int launches = 1 + passes;
    if (sk != keys)
    {
        cudaMemcpyAsync(keys, sk, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
        cudaMemcpyAsync(vals, sv, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    }
    return launches;
}

} // namespace snch
