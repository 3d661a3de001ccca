// Stable LSD radix sort (u32 keys, u32 values) and exclusive scan.  See sort_scan.cuh for the scheme.
#include "sort_scan.cuh"

namespace mb {

__device__ __forceinline__ int64_t resolve_n(int64_t n_host, const uint32_t *n_dev, int64_t max_n) {
    if (n_host >= 0) return n_host;
    int64_t n = (int64_t)(*n_dev);
    return n < max_n ? n : max_n;
}

// hist[pass][digit] += number of keys whose digit of that pass is `digit`; one read of the keys for all passes
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n_host,
                                                                  const uint32_t *__restrict__ n_dev, int64_t max_n,
                                                                  int begin_bit, int passes, uint32_t *__restrict__ hist) {
    const int64_t n = resolve_n(n_host, n_dev, max_n);
    __shared__ uint32_t h[kSortMaxPasses][256];
    const int tid = threadIdx.x;
#pragma unroll
    for (int p = 0; p < kSortMaxPasses; ++p) h[p][tid] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kSortThreads * 4;
    const bool vec = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0;
    for (int64_t i = ((int64_t)blockIdx.x * kSortThreads + tid) * 4; i < n; i += stride) {
        uint32_t k[4];
        int cnt = 4;
        if (vec && i + 4 <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(keys + i);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
        } else {
            cnt = (int)(n - i < 4 ? n - i : 4);
            for (int e = 0; e < cnt; ++e) k[e] = keys[i + e];
        }
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + 8 * p;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (e < cnt) atomicAdd(&h[p][(k[e] >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; ++p)
        if (h[p][tid]) atomicAdd(&hist[p * 256 + tid], h[p][tid]);
}

enum : uint32_t { kFlagAggregate = 1u << 30, kFlagPrefix = 2u << 30, kCountMask = (1u << 30) - 1u };
#ifndef MB_LOOKBACK
#define MB_LOOKBACK 8
#endif
#ifndef MB_RANK_MATCH
#define MB_RANK_MATCH 0
#endif
#ifndef MB_SORT_MINBLOCKS
#define MB_SORT_MINBLOCKS 2
#endif
constexpr int kLookBack = MB_LOOKBACK;   // predecessors polled per look-back step (independent loads: one L2 round trip per step)

// One radix pass over one chunk per CTA.  Ranks are stable; the chunk is first reordered in shared memory (digit-major),
// so that the scatter to global memory writes runs of consecutive addresses, and the (key, value) pairs are parked there
// while the look-back for the counts of the earlier chunks is in flight.
__global__ void __launch_bounds__(kSortThreads, MB_SORT_MINBLOCKS) radix_pass_kernel(const uint32_t *__restrict__ keys_in,
                                                                  const uint32_t *__restrict__ vals_in,
                                                                  uint32_t *__restrict__ keys_out,
                                                                  uint32_t *__restrict__ vals_out, int64_t n_host,
                                                                  const uint32_t *__restrict__ n_dev, int64_t max_n,
                                                                  int shift, const uint32_t *__restrict__ hist,
                                                                  volatile uint32_t *status, uint32_t *cursor, uint2 *runs) {
    const int64_t n = resolve_n(n_host, n_dev, max_n);
    const int64_t nchunks = (n + kSortChunk - 1) / kSortChunk;
    __shared__ uint32_t warp_cnt[kSortWarps][256];
    __shared__ uint32_t base_s[256];
    __shared__ uint32_t scan_s[2][kSortWarps];
    __shared__ uint32_t chunk_s;
    __shared__ uint32_t keys_s[kSortChunk], vals_s[kSortChunk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    if (tid == 0) chunk_s = atomicAdd(cursor, 1u);
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) warp_cnt[w][tid] = 0;
    __syncthreads();
    const int64_t chunk = chunk_s;
    if (chunk >= nchunks) return;
    const int64_t chunk_base = chunk * kSortChunk;
    const int nvalid = (int)(n - chunk_base < kSortChunk ? n - chunk_base : kSortChunk);
    // every key has the same digit (e.g. the exponent byte of the depths of one hand-sized scene): the stable pass is the
    // identity -- copy the chunk, no ranking and no look-back
    if (__syncthreads_or(hist[tid] == (uint32_t)n)) {
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            const int e = r * kSortThreads + tid;
            if (e < nvalid) {
                const uint32_t k = keys_in[chunk_base + e];
                if (keys_out) keys_out[chunk_base + e] = k;
                vals_out[chunk_base + e] = vals_in[chunk_base + e];
                if (runs) {
                    const uint32_t pos = (uint32_t)(chunk_base + e);
                    if (e == 0 || keys_in[chunk_base + e - 1] != k) atomicMax(&runs[k].x, ~pos);
                    if (e == nvalid - 1 || keys_in[chunk_base + e + 1] != k) atomicMax(&runs[k].y, pos + 1u);
                }
            }
        }
        return;
    }
    // stable rank of every key among the keys of its warp's 512-element slice with the same digit
    const int wbase = warp * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int e = wbase + r * 32 + lane;
        key[r] = e < nvalid ? keys_in[chunk_base + e] : 0xffffffffu;
    }
    // peers = lanes of the warp holding the same digit: eight ballots per round (fixed cost, unlike match.any whose cost
    // grows with the number of distinct digits in the warp), all rounds independent
    uint32_t peers[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const bool valid = wbase + r * 32 + lane < nvalid;
        const uint32_t digit = (key[r] >> shift) & 255u;
#if MB_RANK_MATCH
        const uint32_t m = __match_any_sync(0xffffffffu, valid ? digit : 256u);
#else
        uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (digit >> b) & 1u;
            const uint32_t v = __ballot_sync(0xffffffffu, bit);
            m &= bit ? v : ~v;
        }
#endif
        peers[r] = valid ? m : 0u;
    }
    // the first lane of every peer group adds the group size to the warp's digit counter; the shared-memory atomics of
    // consecutive rounds are issued back to back (no register dependency between rounds)
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const uint32_t digit = (key[r] >> shift) & 255u;
        uint32_t old = 0;
        if (peers[r] && (peers[r] & lt) == 0) old = atomicAdd(&warp_cnt[warp][digit], (uint32_t)__popc(peers[r]));
        __syncwarp();
        rank[r] = old;
    }
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int leader = peers[r] ? __ffs(peers[r]) - 1 : lane;
        rank[r] = __shfl_sync(0xffffffffu, rank[r], leader) + __popc(peers[r] & lt);
    }
    // the values travel with the keys from here on; their loads overlap the scans below
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int e = wbase + r * 32 + lane;
        val[r] = e < nvalid ? vals_in[chunk_base + e] : 0u;
    }
    __syncthreads();
    // thread `tid` owns digit `tid`
    uint32_t mine = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        const uint32_t t = warp_cnt[w][tid];
        warp_cnt[w][tid] = mine;     // exclusive prefix over the warps of this chunk
        mine += t;
    }
    status[chunk * 256 + tid] = (chunk == 0 ? kFlagPrefix : kFlagAggregate) | mine;
    // two block scans of 256 values at once: first global position of the digit (histogram of all keys) and first
    // position of the digit inside this chunk
    uint32_t digit_base, local_start;
    {
        const uint32_t v = hist[tid];
        uint32_t incl = v, incl_l = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o), tl = __shfl_up_sync(0xffffffffu, incl_l, o);
            if (lane >= o) { incl += t; incl_l += tl; }
        }
        if (lane == 31) { scan_s[0][warp] = incl; scan_s[1][warp] = incl_l; }
        __syncthreads();     // also: every thread has read its warp_cnt column before the prefixes are used below
        uint32_t wp = 0, wpl = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            wp += (w < warp) ? scan_s[0][w] : 0u;
            wpl += (w < warp) ? scan_s[1][w] : 0u;
        }
        digit_base = wp + incl - v;
        local_start = wpl + incl_l - mine;
    }
    base_s[tid] = local_start;
    __syncthreads();
    // park the pairs digit-major in shared memory
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        if (wbase + r * 32 + lane < nvalid) {
            const uint32_t digit = (key[r] >> shift) & 255u;
            const uint32_t lp = base_s[digit] + warp_cnt[warp][digit] + rank[r];
            keys_s[lp] = key[r];
            vals_s[lp] = val[r];
        }
    }
    // decoupled look-back: keys with this digit in all earlier chunks (kLookBack status words per step)
    uint32_t before = 0;
    {
        int64_t p = chunk - 1;
        bool found = chunk == 0;
        while (!found) {
            uint32_t v[kLookBack];
#pragma unroll
            for (int k = 0; k < kLookBack; ++k) v[k] = p - k >= 0 ? status[(p - k) * 256 + tid] : kFlagPrefix;
            int used = 0;
#pragma unroll
            for (int k = 0; k < kLookBack; ++k) {
                const uint32_t f = v[k] & ~kCountMask;
                if (used == k && f != 0 && !found) {     // contiguous run of published words, up to the first prefix
                    before += v[k] & kCountMask;
                    ++used;
                    found = f == kFlagPrefix;
                }
            }
            p -= used;
        }
        if (chunk > 0) status[chunk * 256 + tid] = kFlagPrefix | (before + mine);
    }
    __syncthreads();     // base_s (local starts) has been read by everyone, the parked pairs are complete
    base_s[tid] = digit_base + before - local_start;     // global position = base_s[digit] + position in the chunk
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
        const int lp = r * kSortThreads + tid;
        if (lp < nvalid) {
            const uint32_t k = keys_s[lp];
            const uint32_t pos = base_s[(k >> shift) & 255u] + (uint32_t)lp;
            if (keys_out) keys_out[pos] = k;
            vals_out[pos] = vals_s[lp];
            // last pass of a sort whose caller wants the runs of equal keys: inside a chunk equal keys are neighbours (digit-major
            // and stable) and land on consecutive positions, so the first / last element of each run of the chunk bounds the
            // key's global run from below / above (a run that spans chunks: the extremes over its chunks)
            if (runs) {
                if (lp == 0 || keys_s[lp - 1] != k) atomicMax(&runs[k].x, ~pos);
                if (lp == nvalid - 1 || keys_s[lp + 1] != k) atomicMax(&runs[k].y, pos + 1u);
            }
        }
    }
}

__global__ void copy_pairs_kernel(const uint32_t *__restrict__ ki, const uint32_t *__restrict__ vi, uint32_t *__restrict__ ko,
                                  uint32_t *__restrict__ vo, int64_t n_host, const uint32_t *__restrict__ n_dev,
                                  int64_t max_n) {
    const int64_t n = resolve_n(n_host, n_dev, max_n);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        ko[i] = ki[i];
        vo[i] = vi[i];
    }
}

int radix_sort_pairs(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n_host,
                     const uint32_t *n_dev, int64_t max_n, int begin_bit, int end_bit, const SortWorkspace &ws,
                     cudaStream_t stream, bool debug, bool hist_ready, uint2 *runs) {
    const int64_t bound = n_host >= 0 ? n_host : max_n;
    if (bound <= 0) return MB_OK;
    const int passes = sort_passes(begin_bit, end_bit);
    const int64_t chunks = sort_chunks(bound);
    if (chunks > ws.max_chunks || passes > kSortMaxPasses) {
        set_error("radix_sort_pairs: workspace too small (%lld chunks > %lld) or too many passes (%d)", (long long)chunks,
                  (long long)ws.max_chunks, passes);
        return MB_ERR_WORKSPACE;
    }
    if (passes == 0) {
        const int grid = (int)(chunks < (int64_t)sm_count() * 8 ? chunks : (int64_t)sm_count() * 8);
        copy_pairs_kernel<<<grid, 256, 0, stream>>>(keys_in, vals_in, keys_out, vals_out, n_host, n_dev, max_n);
        return check_launch("copy_pairs", debug, stream);
    }
    int rc = MB_OK;
    if (!hist_ready) {
        MB_CUDA(cudaMemsetAsync(ws.zeroed, 0, ws.zeroed_bytes, stream));
        {
            KernelTimer kt("radix_hist", stream);
            const int64_t want = (bound + kSortThreads * 16 - 1) / (kSortThreads * 16);
            const int grid = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
            radix_hist_kernel<<<grid, kSortThreads, 0, stream>>>(keys_in, n_host, n_dev, max_n, begin_bit, passes, ws.hist);
        }
        rc = check_launch("radix_hist", debug, stream);
        if (rc) return rc;
    }
    // ping-pong so that the last pass lands in keys_out / vals_out
    uint32_t *src_k = keys_in, *src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        uint32_t *dst_k = to_out ? keys_out : ws.keys_tmp;
        uint32_t *dst_v = to_out ? vals_out : ws.vals_tmp;
        {
            KernelTimer kt("radix_pass", stream);
            radix_pass_kernel<<<(int)chunks, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n_host, n_dev, max_n,
                                                                      begin_bit + 8 * p, ws.hist + p * 256,
                                                                      ws.status + (size_t)p * ws.max_chunks * 256, ws.cursor + p,
                                                                      p == passes - 1 ? runs : nullptr);
        }
        rc = check_launch("radix pass", debug, stream);
        if (rc) return rc;
        src_k = dst_k;
        src_v = dst_v;
    }
    return MB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// exclusive scan with optional gather
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t scan_fetch(const uint32_t *src, const uint32_t *index, int64_t i) {
    return index ? src[index[i]] : src[i];
}

__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(const uint32_t *__restrict__ src,
                                                                     const uint32_t *__restrict__ index, int64_t n,
                                                                     uint32_t *__restrict__ partials) {
    __shared__ uint32_t ws[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanChunk;
    uint32_t s = 0;
#pragma unroll 4
    for (int r = 0; r < kScanItems; ++r) {
        const int64_t i = base + r * kScanThreads + threadIdx.x;
        if (i < n) s += scan_fetch(src, index, i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += ws[w];
        partials[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t *__restrict__ src,
                                                                  const uint32_t *__restrict__ index, int64_t n,
                                                                  const uint32_t *__restrict__ partials,
                                                                  uint32_t *__restrict__ out, uint32_t *__restrict__ total) {
    __shared__ uint32_t ws[kScanThreads / 32];
    __shared__ uint32_t block_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // prefix of all earlier blocks
    uint32_t p = 0;
    for (int b = tid; b < (int)blockIdx.x; b += kScanThreads) p += partials[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if (lane == 0) ws[warp] = p;
    __syncthreads();
    if (tid == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += ws[w];
        block_prefix = t;
    }
    __syncthreads();
    const uint32_t bp = block_prefix;
    __syncthreads();
    // each warp scans 32*kScanItems consecutive elements, round by round
    const int64_t base = (int64_t)blockIdx.x * kScanChunk + (int64_t)warp * (32 * kScanItems);
    uint32_t excl[kScanItems];
    uint32_t carry = 0;
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        const int64_t i = base + r * 32 + lane;
        const uint32_t v = (i < n) ? scan_fetch(src, index, i) : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        excl[r] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) ws[warp] = carry;
    __syncthreads();
    uint32_t wp = 0;
    for (int w = 0; w < warp; ++w) wp += ws[w];
#pragma unroll
    for (int r = 0; r < kScanItems; ++r) {
        const int64_t i = base + r * 32 + lane;
        if (i < n) out[i] = bp + wp + excl[r];
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        uint32_t t = bp;
        for (int w = 0; w < kScanThreads / 32; ++w) t += ws[w];
        *total = t;
    }
}

int exclusive_scan_gather(const uint32_t *src, const uint32_t *index, uint32_t *out, uint32_t *total, int64_t n,
                          uint32_t *partials, cudaStream_t stream, bool debug) {
    const int blocks = (int)(n > 0 ? scan_blocks(n) : 1);
    {
        KernelTimer kt("scan_partials", stream);
        scan_partials_kernel<<<blocks, kScanThreads, 0, stream>>>(src, index, n, partials);
    }
    {
        KernelTimer kt("scan_apply", stream);
        scan_apply_kernel<<<blocks, kScanThreads, 0, stream>>>(src, index, n, partials, out, total);
    }
    return check_launch("exclusive_scan", debug, stream);
}

}  // namespace mb

extern "C" size_t mb_sort_workspace_bytes(int64_t max_n) { return mb::sort_workspace_bytes(max_n < 1 ? 1 : max_n); }

extern "C" int mb_radix_sort_pairs(uint32_t *keys_in, uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out,
                                   int64_t n_host, const uint32_t *n_dev, int64_t max_n, int32_t end_bit, void *workspace,
                                   size_t workspace_bytes, mb_stream_t stream) {
    MB_REQUIRE(keys_in && vals_in && keys_out && vals_out && workspace, "mb_radix_sort_pairs: null pointer");
    MB_REQUIRE(n_host >= 0 || n_dev != nullptr, "mb_radix_sort_pairs: n_host < 0 needs n_dev");
    MB_REQUIRE(end_bit >= 0 && end_bit <= 32, "mb_radix_sort_pairs: end_bit out of range");
    if (n_host >= 0) max_n = n_host;
    if (workspace_bytes < mb::sort_workspace_bytes(max_n < 1 ? 1 : max_n)) {
        mb::set_error("mb_radix_sort_pairs: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    mb::SortWorkspace ws = mb::carve_sort_workspace(workspace, max_n < 1 ? 1 : max_n);
    return mb::radix_sort_pairs(keys_in, vals_in, keys_out, vals_out, n_host, n_dev, max_n, 0, end_bit, ws,
                                (cudaStream_t)stream, false);
}
