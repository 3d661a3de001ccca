// Stable LSD radix sort (u32 keys, u32 values) and exclusive scan.  See sort_scan.cuh for the scheme.
#include "sort_scan.cuh"

namespace mb {

__device__ __forceinline__ int64_t resolve_n(int64_t n_host, const uint32_t *n_dev, int64_t max_n) {
    if (n_host >= 0) return n_host;
    int64_t n = (int64_t)(*n_dev);
    return n < max_n ? n : max_n;
}

// counts[digit * nchunks + chunk] = number of keys of `chunk` whose digit is `digit`
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n_host,
                                                                  const uint32_t *__restrict__ n_dev, int64_t max_n,
                                                                  int shift, uint32_t *__restrict__ counts) {
    const int64_t n = resolve_n(n_host, n_dev, max_n);
    const int64_t nchunks = (n + kSortChunk - 1) / kSortChunk;
    __shared__ uint32_t hist[256];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        hist[tid] = 0;
        __syncthreads();
        const int64_t base = chunk * kSortChunk;
#pragma unroll 4
        for (int r = 0; r < kSortItems; ++r) {
            const int64_t idx = base + r * kSortThreads + tid;
            const bool valid = idx < n;
            const uint32_t digit = valid ? ((keys[idx] >> shift) & 255u) : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&hist[digit], __popc(peers));
        }
        __syncthreads();
        counts[(int64_t)tid * nchunks + chunk] = hist[tid];
        __syncthreads();
    }
}

// One warp per digit: in-place exclusive scan of counts[digit][0..nchunks) and the digit's total.  The prefix over the
// 256 digit totals is taken by every scatter CTA itself (256 values, one block scan), so no single-CTA pass remains.
__global__ void __launch_bounds__(256) radix_offsets_kernel(uint32_t *__restrict__ counts, uint32_t *__restrict__ totals,
                                                            int64_t n_host, const uint32_t *__restrict__ n_dev, int64_t max_n) {
    const int64_t n = resolve_n(n_host, n_dev, max_n);
    const int64_t nchunks = (n + kSortChunk - 1) / kSortChunk;
    const int lane = threadIdx.x & 31;
    const int digit = blockIdx.x * 8 + (threadIdx.x >> 5);
    uint32_t *row = counts + (int64_t)digit * nchunks;
    uint32_t carry = 0;
    for (int64_t base = 0; base < nchunks; base += 32) {
        const int64_t c = base + lane;
        const uint32_t v = c < nchunks ? row[c] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (c < nchunks) row[c] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) totals[digit] = carry;
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                     const uint32_t *__restrict__ vals_in,
                                                                     uint32_t *__restrict__ keys_out,
                                                                     uint32_t *__restrict__ vals_out, int64_t n_host,
                                                                     const uint32_t *__restrict__ n_dev, int64_t max_n,
                                                                     int shift, const uint32_t *__restrict__ offsets,
                                                                     const uint32_t *__restrict__ totals) {
    const int64_t n = resolve_n(n_host, n_dev, max_n);
    const int64_t nchunks = (n + kSortChunk - 1) / kSortChunk;
    __shared__ uint32_t warp_cnt[kSortWarps][256];
    __shared__ uint32_t digit_warp_sum[kSortWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    // first global position of digit `tid` = exclusive prefix of the digit totals
    uint32_t digit_base;
    {
        const uint32_t v = totals[tid];
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) digit_warp_sum[warp] = incl;
        __syncthreads();
        uint32_t wp = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) wp += (w < warp) ? digit_warp_sum[w] : 0u;
        digit_base = wp + incl - v;
    }
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) warp_cnt[w][tid] = 0;
        __syncthreads();
        const int64_t base = chunk * kSortChunk + (int64_t)warp * (32 * kSortItems);
        uint32_t key[kSortItems];
        uint32_t rank[kSortItems];
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            const int64_t idx = base + r * 32 + lane;
            const bool valid = idx < n;
            key[r] = valid ? keys_in[idx] : 0xffffffffu;
            const uint32_t digit = valid ? ((key[r] >> shift) & 255u) : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = warp_cnt[warp][digit];
                warp_cnt[warp][digit] = old + __popc(peers);
            }
            __syncwarp();
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = old + __popc(peers & lt);
        }
        __syncthreads();
        {   // thread `tid` owns digit `tid`: exclusive prefix over warps + global base of (digit, chunk)
            uint32_t run = digit_base + offsets[(int64_t)tid * nchunks + chunk];
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const uint32_t t = warp_cnt[w][tid];
                warp_cnt[w][tid] = run;
                run += t;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kSortItems; ++r) {
            const int64_t idx = base + r * 32 + lane;
            if (idx < n) {
                const uint32_t digit = (key[r] >> shift) & 255u;
                const uint32_t pos = warp_cnt[warp][digit] + rank[r];
                keys_out[pos] = key[r];
                vals_out[pos] = vals_in[idx];
            }
        }
        __syncthreads();
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
                     cudaStream_t stream, bool debug) {
    const int64_t bound = n_host >= 0 ? n_host : max_n;
    if (bound <= 0) return MB_OK;
    const int passes = (end_bit - begin_bit + 7) / 8;
    int64_t chunks = sort_chunks(bound);
    if (chunks > ws.max_chunks) {
        set_error("radix_sort_pairs: workspace too small (%lld chunks > %lld)", (long long)chunks, (long long)ws.max_chunks);
        return MB_ERR_WORKSPACE;
    }
    const int grid = (int)(chunks < (int64_t)sm_count() * 8 ? chunks : (int64_t)sm_count() * 8);
    if (passes == 0) {
        copy_pairs_kernel<<<grid, 256, 0, stream>>>(keys_in, vals_in, keys_out, vals_out, n_host, n_dev, max_n);
        return check_launch("copy_pairs", debug, stream);
    }
    // ping-pong so that the last pass lands in keys_out / vals_out
    uint32_t *src_k = keys_in, *src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        uint32_t *dst_k = to_out ? keys_out : ws.keys_tmp;
        uint32_t *dst_v = to_out ? vals_out : ws.vals_tmp;
        const int shift = begin_bit + 8 * p;
        {
            KernelTimer kt("radix_hist", stream);
            radix_hist_kernel<<<grid, kSortThreads, 0, stream>>>(src_k, n_host, n_dev, max_n, shift, ws.counts);
        }
        {
            KernelTimer kt("radix_offsets", stream);
            radix_offsets_kernel<<<32, 256, 0, stream>>>(ws.counts, ws.totals, n_host, n_dev, max_n);
        }
        {
            KernelTimer kt("radix_scatter", stream);
            radix_scatter_kernel<<<grid, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n_host, n_dev, max_n, shift,
                                                                   ws.counts, ws.totals);
        }
        int rc = check_launch("radix pass", debug, stream);
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
