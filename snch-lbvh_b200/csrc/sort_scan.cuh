// sort_scan.cuh — device-wide primitives used by the build (and by query re-ordering):
//   * exclusive_scan_u32 : reduce-then-scan over up to 16 Mi elements per level (recursive above that)
//   * radix_sort_pairs   : stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass — onesweep (one histogram
//                          kernel + one decoupled-look-back kernel per pass); the histogram/scan/scatter multi-kernel form
//                          remains for n >= 2^30 and for A/B ("sort.onesweep" = 0)
//
// Replaces thrust::stable_sort_by_key in the reference (bvh.cuh:452-454), which drags a 48-byte payload through every
// pass; here only (key, index) = 8 bytes move and the payload is gathered once afterwards.
// Hand-written for sm_100a: warp-level __match_any_sync ranking, shared-memory digit counters, no library calls.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace snch
{

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems; // 4096 keys per CTA
constexpr int kSortBins = 256;

// scratch sizing helpers (in u32 elements)
inline uint64_t scan_scratch_elems(uint64_t n)
{
    uint64_t total = 0;
    while (n > 1)
    {
        n = (n + kScanTile - 1) / kScanTile;
        total += n + 1;
        if (n <= 1) break;
    }
    return total + 8;
}
inline uint64_t sort_scratch_elems(uint64_t n)
{
    const uint64_t tiles = (n + kSortTile - 1) / kSortTile;
    const uint64_t counts = tiles * kSortBins;
    const uint64_t multipass = counts + scan_scratch_elems(counts) + 8;
    const uint64_t onesweep = 4 * kSortBins + 8 + 4 * (2 * counts + kSortBins); // global histograms, tile counters, per-pass status of (half) tiles
    return multipass > onesweep ? multipass : onesweep;
}

// both return the number of kernels they launched (the library reports launch counts, snch_scene_counter)
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, uint32_t *scratch, cudaStream_t stream);

// Sorts (keys, vals) by key bits [first_bit, first_bit + bits), ascending, stable.  keys/vals are overwritten with the
// result; keys_tmp/vals_tmp are ping-pong buffers of the same size; scratch has sort_scratch_elems(n) u32.
int radix_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint64_t n, int bits,
                      uint32_t *scratch, cudaStream_t stream, int first_bit = 0);

// process-wide switch between the onesweep and the multi-kernel sort (snch_scene_set_option "sort.onesweep")
void set_sort_onesweep(int on);
void set_sort_lookback(int window); // 1 = serial look-back, else windows of 8 tiles

} // namespace snch
