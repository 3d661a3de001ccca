// Stable LSD radix sort of (u32 key, u32 value) pairs and an exclusive scan, with the element count optionally
// read from device memory so that the whole frame can be enqueued without a host round trip.
//
// One pass over an 8-bit digit is three launches:
//   radix_hist    : per-chunk digit histogram  -> counts[digit][chunk]           (chunk = 4096 consecutive elements)
//   radix_offsets : per digit, exclusive scan of counts over the chunks, in place, + digit totals (one warp per digit)
//   radix_scatter : stable rank inside the chunk (warp match + per-warp counters) and scatter
// Everything the sorts of one frame touch is L2-resident on B200 (126 MB), so the passes are L2-bound, not HBM-bound.
#pragma once
#include "common.cuh"

namespace mb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                           // per thread
constexpr int kSortChunk = kSortThreads * kSortItems;    // 4096
constexpr int kSortWarps = kSortThreads / 32;

inline int64_t sort_chunks(int64_t n) { return (n + kSortChunk - 1) / kSortChunk; }

struct SortWorkspace {
    uint32_t *counts;   // [256][max_chunks]
    uint32_t *totals;   // [256] keys per digit of the current pass
    uint32_t *keys_tmp, *vals_tmp;
    int64_t max_chunks;
};

inline size_t sort_workspace_bytes(int64_t max_n) {
    int64_t ch = sort_chunks(max_n) + 1;
    return align_up(256 * ch * sizeof(uint32_t)) + align_up(256 * sizeof(uint32_t)) + 2 * align_up((size_t)(max_n + 1) * sizeof(uint32_t));
}

inline SortWorkspace carve_sort_workspace(void *p, int64_t max_n) {
    Carver c(p);
    SortWorkspace w;
    w.max_chunks = sort_chunks(max_n) + 1;
    w.counts = c.take<uint32_t>(256 * w.max_chunks);
    w.totals = c.take<uint32_t>(256);
    w.keys_tmp = c.take<uint32_t>(max_n + 1);
    w.vals_tmp = c.take<uint32_t>(max_n + 1);
    return w;
}

// n = n_host if >= 0 else min(*n_dev, max_n)
int radix_sort_pairs(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n_host,
                     const uint32_t *n_dev, int64_t max_n, int begin_bit, int end_bit, const SortWorkspace &ws,
                     cudaStream_t stream, bool debug);

// out[i] = sum_{j<i} f(j) for i < n, where f(j) = src[j] or src[index[j]] (gather); *total = sum of all.
// partials: scratch of scan_blocks(n) uint32.  n is a host value.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanChunk = kScanThreads * kScanItems;
inline int64_t scan_blocks(int64_t n) { return (n + kScanChunk - 1) / kScanChunk; }
int exclusive_scan_gather(const uint32_t *src, const uint32_t *index, uint32_t *out, uint32_t *total, int64_t n,
                          uint32_t *partials, cudaStream_t stream, bool debug);

}  // namespace mb
