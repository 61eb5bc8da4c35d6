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

int radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int bits,
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

} // namespace snch
