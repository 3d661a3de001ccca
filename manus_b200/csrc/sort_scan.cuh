// Stable LSD radix sort of (u32 key, u32 value) pairs and an exclusive scan, with the element count optionally
// read from device memory so that the whole frame can be enqueued without a host round trip.
//
// Onesweep scheme: ONE histogram kernel reads the keys once and counts the 8-bit digits of every pass (digit counts do
// not depend on the order of the keys), then each pass is a single kernel: a CTA takes the next chunk of 4096 keys
// (dynamic chunk id, so that every predecessor is already running), ranks its keys stably (warp match + per-warp
// counters), publishes its per-digit counts, obtains the counts of all earlier chunks by decoupled look-back over
// [chunk][digit] status words (2 flag bits + 30-bit count) and scatters.  A sort of k passes is k + 2 launches
// (memset, histogram, k passes).  Everything the sorts of one frame touch is L2-resident on B200 (126 MB).
#pragma once
#include "common.cuh"

namespace mb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                           // per thread
constexpr int kSortChunk = kSortThreads * kSortItems;    // 4096
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortMaxPasses = 4;

inline int64_t sort_chunks(int64_t n) { return (n + kSortChunk - 1) / kSortChunk; }

struct SortWorkspace {
    uint32_t *zeroed;        // start of the region cleared by one memset per sort: hist | cursors | status
    size_t zeroed_bytes;
    uint32_t *hist;          // [kSortMaxPasses][256] digit counts of each pass
    uint32_t *cursor;        // [kSortMaxPasses] next chunk id of each pass
    uint32_t *status;        // [kSortMaxPasses][max_chunks][256] look-back words
    uint32_t *keys_tmp, *vals_tmp;
    int64_t max_chunks;
};

inline size_t sort_workspace_bytes(int64_t max_n) {
    const int64_t ch = sort_chunks(max_n) + 1;
    return align_up(kSortMaxPasses * 256 * sizeof(uint32_t)) + align_up(64 * sizeof(uint32_t)) +
           align_up((size_t)kSortMaxPasses * ch * 256 * sizeof(uint32_t)) + 2 * align_up((size_t)(max_n + 1) * sizeof(uint32_t));
}

inline SortWorkspace carve_sort_workspace(void *p, int64_t max_n) {
    Carver c(p);
    SortWorkspace w;
    w.max_chunks = sort_chunks(max_n) + 1;
    w.hist = c.take<uint32_t>(kSortMaxPasses * 256);
    w.cursor = c.take<uint32_t>(64);
    w.status = c.take<uint32_t>((size_t)kSortMaxPasses * w.max_chunks * 256);
    w.zeroed = w.hist;
    w.zeroed_bytes = (size_t)((char *)(w.status + (size_t)kSortMaxPasses * w.max_chunks * 256) - (char *)w.hist);
    w.keys_tmp = c.take<uint32_t>(max_n + 1);
    w.vals_tmp = c.take<uint32_t>(max_n + 1);
    return w;
}

// n = n_host if >= 0 else min(*n_dev, max_n).
// hist_ready: the caller cleared ws.zeroed (ws.zeroed_bytes) BEFORE the kernel that produced the keys ran, and that kernel
// counted the digits of every key it wrote into ws.hist (hist_smem_* below): the sort is then the passes only -- no memset and
// no histogram launch between the producer and the first pass.
// runs (zeroed by the caller, indexed by key; needs at least one pass): the LAST pass also records where every key's run of equal keys
// lies in the output -- runs[k] = (~first position, last position + 1) for the keys that occur, (0, 0) for the others (atomicMax
// on zero-initialised words; the caller decodes x = ~runs[k].x) -- so that nobody has to read the sorted keys again to find the
// boundaries.
int radix_sort_pairs(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n_host,
                     const uint32_t *n_dev, int64_t max_n, int begin_bit, int end_bit, const SortWorkspace &ws,
                     cudaStream_t stream, bool debug, bool hist_ready = false, uint2 *runs = nullptr);

inline int sort_passes(int begin_bit, int end_bit) { return (end_bit - begin_bit + 7) / 8; }

#ifdef __CUDACC__
// Digit counting inside the kernel that produces the keys.  h = kSortMaxPasses * 256 words of shared memory, zeroed by
// hist_smem_zero (+ a barrier) at the start of the kernel; hist_smem_count for every key the kernel writes (any subset of a
// warp's lanes may call it: lanes that arrive together are grouped with __activemask); hist_smem_flush once per CTA after a
// barrier that follows the last count.  Digits from `agg_from` on are counted once per group of lanes holding the same digit
// (the high bytes of depth keys and of tile ids are nearly constant inside a warp: 32 same-address shared atomics otherwise).
__device__ __forceinline__ void hist_smem_zero(uint32_t *h) {
    for (int e = threadIdx.x; e < kSortMaxPasses * 256; e += blockDim.x) h[e] = 0;
}

__device__ __forceinline__ void hist_smem_count(uint32_t *h, uint32_t key, int passes, int agg_from) {
    const unsigned act = __activemask();
    const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
#pragma unroll
    for (int p = 0; p < kSortMaxPasses; ++p) {
        if (p < passes) {
            const uint32_t d = (key >> (8 * p)) & 255u;
            if (p < agg_from) {
                atomicAdd(&h[p * 256 + d], 1u);
            } else {
                const unsigned peers = __match_any_sync(act, d);
                if ((peers & lt) == 0) atomicAdd(&h[p * 256 + d], (uint32_t)__popc(peers));
            }
        }
    }
}

__device__ __forceinline__ void hist_smem_flush(const uint32_t *h, int passes, uint32_t *hist) {
    for (int e = threadIdx.x; e < passes * 256; e += blockDim.x)
        if (h[e]) atomicAdd(&hist[e], h[e]);
}
#endif

// out[i] = sum_{j<i} f(j) for i < n, where f(j) = src[j] or src[index[j]] (gather); *total = sum of all.
// partials: scratch of scan_blocks(n) uint32.  n is a host value.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanChunk = kScanThreads * kScanItems;
inline int64_t scan_blocks(int64_t n) { return (n + kScanChunk - 1) / kScanChunk; }
int exclusive_scan_gather(const uint32_t *src, const uint32_t *index, uint32_t *out, uint32_t *total, int64_t n,
                          uint32_t *partials, cudaStream_t stream, bool debug);

}  // namespace mb
