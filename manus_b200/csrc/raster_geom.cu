// Rasterizer stage 1 and binning:  per-Gaussian projection, depth ordering, instance
// emission, per-tile ordering, tile ranges and the 48-B blend records (Appendix A.2, re-designed: Gaussians are
// depth-sorted ONCE (P keys), instances are emitted in depth order and then only need a stable partition by tile id
// (ceil(log2(tiles)/8) = 2 radix passes over D) instead of upstream's 6-pass 64-bit sort over D).
//
// This translation unit is compiled with --fmad=false: every quantity that feeds an integer decision (near cull,
// radius ceil, tile rectangle truncation, depth order) is evaluated with individually rounded IEEE operations in a
// fixed order, so radii / tiles touched / instance lists are reproducible bit-for-bit on any IEEE machine.
#include "raster_state.cuh"
#include "project_fwd.cuh"

namespace mb {

struct PreArgs {
    int P, W, H, gx, gy, deg, M;
    float tanx, tany, focx, focy, scale_mod;
    const float *means3D, *opac, *colors, *cov3D_precomp, *scales, *rots, *shs, *view, *proj, *campos, *tanfov_dev;
    GeomState g;
    int32_t *radii;
    uint32_t *depth_hist;     // [4][256] digit counts of the depth keys (the depth sort's histogram, zeroed before the launch)
};

template <bool kPrecompCov, bool kSH>
__global__ void __launch_bounds__(256) preprocess_kernel(PreArgs a) {   // `a` is a by-value copy: the camera override below is local
    __shared__ float cam[36];
    const int tid = threadIdx.x;
    if (tid < 16) cam[tid] = a.view[tid];
    else if (tid < 32) cam[tid] = a.proj[tid - 16];
    else if (tid < 35) cam[tid] = a.campos[tid - 32];
    __syncthreads();
    if (a.tanfov_dev) {   // camera intrinsics from device memory (same fp32 expressions as raster_dims on the host)
        a.tanx = a.tanfov_dev[0]; a.tany = a.tanfov_dev[1];
        a.focx = a.W / (2.0f * a.tanx); a.focy = a.H / (2.0f * a.tany);
    }
    const float *v = cam, *p = cam + 16;
    const int i = blockIdx.x * 256 + tid;
    bool visible = false;
    uint32_t my_tiles = 0, my_key = 0;
    __shared__ uint32_t hist_s[kSortMaxPasses * 256];
    hist_smem_zero(hist_s);
    if (i < a.P) {
        const float mx = a.means3D[3 * i], my = a.means3D[3 * i + 1], mz = a.means3D[3 * i + 2];
        // A.1 step 4: 3-D covariance
        float c6[6];
        if (kPrecompCov) {
#pragma unroll
            for (int k = 0; k < 6; ++k) c6[k] = a.cov3D_precomp[6 * (size_t)i + k];
        } else {
            float R[9], L[9];
            quat_to_rot(a.rots[4 * i], a.rots[4 * i + 1], a.rots[4 * i + 2], a.rots[4 * i + 3], R);
            const float s[3] = {a.scale_mod * a.scales[3 * i], a.scale_mod * a.scales[3 * i + 1], a.scale_mod * a.scales[3 * i + 2]};
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) L[3 * r + k] = R[3 * r + k] * s[k];
            float S[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc += L[3 * r + k] * L[3 * c + k];
                    S[3 * r + c] = acc;
                }
            c6[0] = S[0]; c6[1] = S[1]; c6[2] = S[2]; c6[3] = S[4]; c6[4] = S[5]; c6[5] = S[8];
#pragma unroll
            for (int k = 0; k < 6; ++k) a.g.cov3D[6 * (size_t)i + k] = c6[k];
        }
        float rgb[3];
        if (kSH) {   // step 10: colour from SH in the world-space view direction
            float dx = mx - cam[32], dy = my - cam[33], dz = mz - cam[34];
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            dx *= inv; dy *= inv; dz *= inv;
            float basis[16];
            sh_basis(a.deg, dx, dy, dz, basis);
            const int nb = (a.deg + 1) * (a.deg + 1);
            const float *sh = a.shs + (size_t)i * a.M * 3;
            uint32_t mask = 0;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float r = 0.f;
                for (int k = 0; k < nb; ++k) r += basis[k] * sh[3 * k + ch];
                r += 0.5f;
                if (r < 0.f) { mask |= 1u << ch; r = 0.f; }
                rgb[ch] = r;
            }
            a.g.clamped[i] = mask;
        } else {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) rgb[ch] = a.colors[3 * (size_t)i + ch];
        }
        Projected pr;
        project_forward(v, p, a.tanx, a.tany, a.focx, a.focy, a.W, a.H, a.gx, a.gy, mx, my, mz, c6, a.opac[i], rgb, pr);
        visible = pr.visible;
        if (visible) {
            a.g.rect[i] = pr.rect;
            a.g.rec[i] = pr.rec;
        }
        a.radii[i] = pr.radius;
        my_tiles = pr.tiles;
        a.g.tiles_touched[i] = pr.tiles;
        a.g.depth_key[i] = pr.key;
        a.g.ident[i] = (uint32_t)i;
        my_key = pr.key;
    }
    // per-CTA totals: visible Gaussians and instances (num_rendered is known as soon as this kernel has run)
    __shared__ uint32_t s_vis, s_tiles;
    if (tid == 0) { s_vis = 0; s_tiles = 0; }
    __syncthreads();      // also: hist_s is zeroed
    if (i < a.P) hist_smem_count(hist_s, my_key, 4, 2);       // digit counts of the depth sort that follows (bytes 2, 3 aggregated)
    const unsigned vis = __ballot_sync(0xffffffffu, visible);
    const uint32_t wt = __reduce_add_sync(0xffffffffu, my_tiles);
    if ((tid & 31) == 0 && vis) {
        atomicAdd(&s_vis, (uint32_t)__popc(vis));
        atomicAdd(&s_tiles, wt);
    }
    __syncthreads();
    if (tid == 0 && s_vis) {
        atomicAdd(&a.g.counters[kCntVisible], s_vis);
        atomicAdd(&a.g.counters[kCntRendered], s_tiles);
    }
    hist_smem_flush(hist_s, 4, a.depth_hist);
}

// Instance emission, warp-cooperative: a warp owns 32 consecutive depth-sorted Gaussians; their (tile id, gaussian id)
// runs are contiguous in the instance list, so the lanes walk the warp's slots in order (coalesced 128-B stores) and
// find the owning Gaussian of each slot by a shuffle binary search over the warp's inclusive tile counts.  The few
// Gaussians that cover more than kBigTiles tiles (up to the whole screen) are queued and written by whole CTAs in
// emit_big_kernel, so that no warp is left with thousands of slots.
constexpr uint32_t kBigTiles = 64;
enum : uint32_t { kScanAggregate = 1u << 30, kScanPrefix = 2u << 30, kScanMask = (1u << 30) - 1u };

// The instance offsets (exclusive scan of the tile counts in depth order) are computed here as well, single pass:
// a CTA takes the next chunk of 1024 depth-sorted Gaussians (dynamic chunk id), scans its counts, publishes the chunk
// total and gets the total of all earlier chunks by decoupled look-back (warp 0 reads 32 predecessors per step).
#ifndef MB_EMIT_ITEMS
#define MB_EMIT_ITEMS 2
#endif
constexpr int kEmitItems = MB_EMIT_ITEMS;          // Gaussians per thread
constexpr int kEmitChunk = 256 * kEmitItems;       // Gaussians per chunk

__global__ void __launch_bounds__(256) emit_instances_kernel(int P, int gx, const uint32_t *__restrict__ sorted_idx,
                                                             const uint32_t *__restrict__ tiles_touched,
                                                             const ushort4 *__restrict__ rect,
                                                             uint32_t *__restrict__ counters, volatile uint32_t *status,
                                                             int64_t capacity, uint32_t *__restrict__ tile_out,
                                                             uint32_t *__restrict__ gid_out, uint2 *__restrict__ big_list,
                                                             int key_passes, uint32_t *__restrict__ key_hist) {
    __shared__ uint32_t s_chunk, s_base, s_warp[8];
    // digit counts of the tile ids this CTA writes: the histogram of the stable partition by tile that follows
    __shared__ uint32_t hist_s[kSortMaxPasses * 256];
    hist_smem_zero(hist_s);
    if (blockIdx.x == 0 && threadIdx.x == 0 && (int64_t)counters[kCntRendered] > capacity) counters[kCntOverflow] = 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nchunks = (P + kEmitChunk - 1) / kEmitChunk;
    while (true) {
        __syncthreads();   // s_chunk / s_base / s_warp of the previous chunk are no longer read
        if (threadIdx.x == 0) s_chunk = atomicAdd(&counters[kCntEmitCursor], 1u);
        __syncthreads();
        const int chunk = (int)s_chunk;
        if (chunk >= nchunks) {      // two barriers after the last count of this CTA
            hist_smem_flush(hist_s, key_passes, key_hist);
            return;
        }
        // a warp owns 32 * kEmitItems consecutive depth-sorted Gaussians, item k of lane l is Gaussian first + 32 k + l;
        // the three dependent loads of all items are in flight together
        const int first = chunk * kEmitChunk + warp * (32 * kEmitItems);
        uint32_t g[kEmitItems], cnt[kEmitItems];
        ushort4 r[kEmitItems];
#pragma unroll
        for (int k = 0; k < kEmitItems; ++k) {
            const int i = first + 32 * k + lane;
            g[k] = i < P ? sorted_idx[i] : 0xffffffffu;
        }
#pragma unroll
        for (int k = 0; k < kEmitItems; ++k) cnt[k] = g[k] != 0xffffffffu ? tiles_touched[g[k]] : 0u;
#pragma unroll
        for (int k = 0; k < kEmitItems; ++k) r[k] = cnt[k] ? rect[g[k]] : make_ushort4(0, 0, 1, 1);
        // scan of the full counts (big Gaussians keep their slots; they are written by emit_big_kernel)
        uint32_t excl[kEmitItems], run = 0;
#pragma unroll
        for (int k = 0; k < kEmitItems; ++k) {
            uint32_t incl = cnt[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            excl[k] = run + incl - cnt[k];
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_warp[warp] = run;
        __syncthreads();
        uint32_t warp_excl = 0, chunk_total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t t = s_warp[w];
            if (w < warp) warp_excl += t;
            chunk_total += t;
        }
        if (warp == 0) {
            if (lane == 0) status[chunk] = (chunk == 0 ? kScanPrefix : kScanAggregate) | chunk_total;
            uint32_t before = 0;
            if (chunk > 0) {
                int p = chunk - 1;   // nearest predecessor not yet accounted for
                while (true) {
                    const int idx = p - lane;
                    const uint32_t v = idx >= 0 ? status[idx] : (uint32_t)kScanPrefix;   // virtual zero prefix in front of chunk 0
                    const unsigned not_ready = __ballot_sync(0xffffffffu, (v & ~kScanMask) == 0);
                    const unsigned is_prefix = __ballot_sync(0xffffffffu, (v & ~kScanMask) == kScanPrefix);
                    const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
                    const int first_pf = is_prefix ? __ffs(is_prefix) - 1 : 32;
                    const int take = min(first_nr, first_pf < 32 ? first_pf + 1 : 32);   // lanes [0, take) are usable
                    const uint32_t part = __reduce_add_sync(0xffffffffu, lane < take ? (v & kScanMask) : 0u);
                    before += part;
                    if (first_pf < first_nr) break;
                    p -= take;
                }
                if (lane == 0) status[chunk] = kScanPrefix | (before + chunk_total);
            }
            if (lane == 0) s_base = before;
        }
        __syncthreads();
        const uint32_t wbase = s_base + warp_excl;
#pragma unroll
        for (int k = 0; k < kEmitItems; ++k) {
            const uint32_t off = wbase + excl[k];
            uint32_t c = cnt[k];
            if (c > kBigTiles) {
                big_list[atomicAdd(&counters[kCntBig], 1u)] = make_uint2((uint32_t)(first + 32 * k + lane), off);
                c = 0;
            }
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            const uint32_t width = (uint32_t)(r[k].z - r[k].x);
            for (uint32_t s0 = 0; s0 < total; s0 += 32) {
                const uint32_t sl = s0 + lane;
                int lo = 0, hi = 31;   // first lane whose inclusive count exceeds sl
#pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const int mid = (lo + hi) >> 1;
                    const uint32_t v = __shfl_sync(0xffffffffu, incl, mid);
                    if (v > sl) hi = mid; else lo = mid + 1;
                }
                const uint32_t o_incl = __shfl_sync(0xffffffffu, incl, lo), o_cnt = __shfl_sync(0xffffffffu, c, lo);
                const uint32_t o_g = __shfl_sync(0xffffffffu, g[k], lo), o_w = __shfl_sync(0xffffffffu, width, lo);
                const uint32_t o_x0 = __shfl_sync(0xffffffffu, (uint32_t)r[k].x, lo), o_y0 = __shfl_sync(0xffffffffu, (uint32_t)r[k].y, lo);
                const uint32_t o_off = __shfl_sync(0xffffffffu, off, lo);
                const uint32_t local = sl - (o_incl - o_cnt);
                const int64_t pos = (int64_t)o_off + local;
                if (sl < total && pos < capacity) {
                    const uint32_t ty = local / o_w, tx = local - ty * o_w;
                    const uint32_t tile = (o_y0 + ty) * (uint32_t)gx + o_x0 + tx;
                    tile_out[pos] = tile;
                    gid_out[pos] = o_g;
                    hist_smem_count(hist_s, tile, key_passes, 1);
                }
            }
        }
    }
}

// one CTA per queued big Gaussian (grid-stride)
__global__ void __launch_bounds__(256) emit_big_kernel(int gx, const uint32_t *__restrict__ sorted_idx,
                                                       const uint32_t *__restrict__ tiles_touched,
                                                       const ushort4 *__restrict__ rect, const uint32_t *__restrict__ counters,
                                                       int64_t capacity, const uint2 *__restrict__ big_list,
                                                       uint32_t *__restrict__ tile_out, uint32_t *__restrict__ gid_out,
                                                       int key_passes, uint32_t *__restrict__ key_hist) {
    const uint32_t nbig = counters[kCntBig];
    if (blockIdx.x >= nbig) return;
    __shared__ uint32_t hist_s[kSortMaxPasses * 256];
    hist_smem_zero(hist_s);
    __syncthreads();
    for (uint32_t e = blockIdx.x; e < nbig; e += gridDim.x) {
        const uint2 entry = big_list[e];
        const uint32_t g = sorted_idx[entry.x], cnt = tiles_touched[g];
        const ushort4 r = rect[g];
        const uint32_t w = (uint32_t)(r.z - r.x);
        const int64_t off = entry.y;
        for (uint32_t local = threadIdx.x; local < cnt; local += blockDim.x) {
            const int64_t pos = off + local;
            if (pos < capacity) {
                const uint32_t ty = local / w, tx = local - ty * w;
                const uint32_t tile = ((uint32_t)r.y + ty) * (uint32_t)gx + (uint32_t)r.x + tx;
                tile_out[pos] = tile;
                gid_out[pos] = g;
                hist_smem_count(hist_s, tile, key_passes, 1);
            }
        }
    }
    __syncthreads();
    hist_smem_flush(hist_s, key_passes, key_hist);
}

// Work order of the tile kernels: tiles by descending weight (list length for the forward, deepest last-contributor for
// the backward) so that the longest lists start first.  Counting sort over 129 quarter-octave buckets in ONE launch of a
// few co-resident CTAs: bucket counts go to global memory, the CTAs meet at an arrival counter, then every tile takes the
// next slot of its bucket (the order inside a bucket does not matter).  The last CTA to leave clears the workspace.
__device__ __forceinline__ int weight_bucket(uint32_t w) {
    if (w == 0) return 0;
    const int e = 31 - __clz(w);
    const int m = e >= 2 ? (int)((w >> (e - 2)) & 3u) : (int)((w << (2 - e)) & 3u);
    return 1 + 4 * e + m;
}

constexpr int kOrderThreads = 256;

//
// decode != 0 (the forward's launch): on entry ranges[t] is what the last pass of the partition by tile left there --
// (~first position, last position + 1) of tile t's run in the sorted list, (0, 0) for an empty tile; the kernel rewrites it as
// the list range (first, last + 1) while it is at it (every tile is visited by exactly one thread), so no kernel reads the sorted
// instance list just to find the tile boundaries.
__device__ __forceinline__ uint2 decode_range(uint2 r, int decode) {
    return (decode && r.y) ? make_uint2(~r.x, r.y) : r;
}

__global__ void __launch_bounds__(kOrderThreads) tile_order_kernel(const uint32_t *__restrict__ weight, uint2 *ranges,
                                                                   int tiles, uint32_t *__restrict__ order, uint32_t *ws, int decode) {
    __shared__ uint32_t hist[kOrderBuckets], base[kOrderBuckets];
    __shared__ bool s_last;
    uint32_t *g_hist = ws, *g_cursor = ws + kOrderBuckets;
    volatile uint32_t *arrive = ws + 2 * kOrderBuckets;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int b = tid; b < kOrderBuckets; b += kOrderThreads) hist[b] = 0;
    __syncthreads();
    const int stride = gridDim.x * kOrderThreads;
    for (int t0 = blockIdx.x * kOrderThreads; t0 < tiles; t0 += stride) {
        const int t = t0 + tid;
        int bkt = -1;
        if (t < tiles) {
            if (weight) bkt = weight_bucket(weight[t]);
            else {
                const uint2 r = decode_range(ranges[t], decode);
                bkt = weight_bucket(r.y - r.x);
            }
        }
        // empty tiles are the majority: one atomic per warp for bucket 0
        const unsigned zeros = __ballot_sync(0xffffffffu, bkt == 0);
        if (bkt > 0) atomicAdd(&hist[bkt], 1u);
        else if (bkt == 0 && lane == __ffs(zeros) - 1) atomicAdd(&hist[0], (uint32_t)__popc(zeros));
    }
    __syncthreads();
    for (int b = tid; b < kOrderBuckets; b += kOrderThreads)
        if (hist[b]) atomicAdd(&g_hist[b], hist[b]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd((uint32_t *)arrive, 1u);
        while (arrive[0] < gridDim.x) {
        }
        __threadfence();
    }
    __syncthreads();
    // start of every bucket in the descending order
    if (tid < 32) {
        uint32_t run = 0;
        for (int b0 = kOrderBuckets - 1; b0 >= 0; b0 -= 32) {
            const int b = b0 - lane;
            const uint32_t v = b >= 0 ? __ldcg(&g_hist[b]) : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (b >= 0) base[b] = run + incl - v;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    for (int t0 = blockIdx.x * kOrderThreads; t0 < tiles; t0 += stride) {
        const int t = t0 + tid;
        int bkt = -1;
        if (t < tiles) {
            if (weight) bkt = weight_bucket(weight[t]);
            else {
                const uint2 r = decode_range(ranges[t], decode);
                bkt = weight_bucket(r.y - r.x);
                if (decode && r.y) ranges[t] = r;
            }
        }
        const unsigned zeros = __ballot_sync(0xffffffffu, bkt == 0);
        const int leader = __ffs(zeros) - 1;
        uint32_t zstart = 0;
        if (bkt == 0 && lane == leader) zstart = atomicAdd(&g_cursor[0], (uint32_t)__popc(zeros));
        if (zeros) zstart = __shfl_sync(0xffffffffu, zstart, leader);
        if (bkt > 0) order[base[bkt] + atomicAdd(&g_cursor[bkt], 1u)] = (uint32_t)t;
        else if (bkt == 0) order[base[0] + zstart + __popc(zeros & lanemask_lt())] = (uint32_t)t;
    }
    // leave the workspace zeroed for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd((uint32_t *)arrive + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last)
        for (int b = tid; b < kOrderWs; b += kOrderThreads) ws[b] = 0;
}

// The CTAs of the two ordering kernels meet at an arrival counter, so all of them must be resident at the same time: a
// COOPERATIVE launch (at most 32 CTAs) makes the driver start the grid only when every CTA fits on the device at once, whatever
// else is running (the other views of a step share the GPU); the guarantee also holds for the kernel node of a captured graph.
template <typename... Params, typename... Args>
static cudaError_t launch_cooperative(void (*kernel)(Params...), int grid, int block, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeCooperative;
    attr.val.cooperative = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Params>(args)...);
}

int tile_order(const uint32_t *weight_or_null, uint2 *ranges_or_null, int tiles, uint32_t *order, uint32_t *ws,
               cudaStream_t s, bool debug, bool decode_runs) {
    KernelTimer kt("tile_order", s);
    const int grid = max(1, min((tiles + kOrderThreads - 1) / kOrderThreads, min(32, sm_count())));
    MB_CUDA(launch_cooperative(tile_order_kernel, grid, kOrderThreads, s, weight_or_null, ranges_or_null, tiles, order, ws, decode_runs ? 1 : 0));
    return check_launch("tile_order", debug, s);
}

__global__ void __launch_bounds__(kOrderThreads) segment_items_kernel(const uint32_t *__restrict__ maxlast, int tiles,
                                                                      uint2 *__restrict__ items, uint32_t *__restrict__ n_items,
                                                                      uint32_t *ws) {
    __shared__ uint32_t hist[kOrderBuckets], base[kOrderBuckets];
    __shared__ bool s_last;
    uint32_t *g_hist = ws, *g_cursor = ws + kOrderBuckets;
    volatile uint32_t *arrive = ws + 2 * kOrderBuckets;
    const int tid = threadIdx.x, lane = tid & 31;
    const int bfull = weight_bucket((uint32_t)kSeg);
    for (int b = tid; b < kOrderBuckets; b += kOrderThreads) hist[b] = 0;
    __syncthreads();
    const int stride = gridDim.x * kOrderThreads;
    for (int t = blockIdx.x * kOrderThreads + tid; t < tiles; t += stride) {
        const uint32_t ml = maxlast[t];
        const uint32_t nfull = ml / kSeg, rem = ml % kSeg;
        if (nfull) atomicAdd(&hist[bfull], nfull);
        if (rem) atomicAdd(&hist[weight_bucket(rem)], 1u);
    }
    __syncthreads();
    for (int b = tid; b < kOrderBuckets; b += kOrderThreads)
        if (hist[b]) atomicAdd(&g_hist[b], hist[b]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        atomicAdd((uint32_t *)arrive, 1u);
        while (arrive[0] < gridDim.x) {
        }
        __threadfence();
    }
    __syncthreads();
    if (tid < 32) {
        uint32_t run = 0;
        for (int b0 = kOrderBuckets - 1; b0 >= 0; b0 -= 32) {
            const int b = b0 - lane;
            const uint32_t v = b >= 0 ? __ldcg(&g_hist[b]) : 0u;
            uint32_t incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (b >= 0) base[b] = run + incl - v;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (blockIdx.x == 0 && lane == 0) *n_items = run;
    }
    __syncthreads();
    for (int t = blockIdx.x * kOrderThreads + tid; t < tiles; t += stride) {
        const uint32_t ml = maxlast[t];
        const uint32_t nfull = ml / kSeg, rem = ml % kSeg;
        if (nfull) {
            const uint32_t p = base[bfull] + atomicAdd(&g_cursor[bfull], nfull);
            for (uint32_t sgm = 0; sgm < nfull; ++sgm) items[p + sgm] = make_uint2((uint32_t)t, sgm);
        }
        if (rem) {
            const int b = weight_bucket(rem);
            items[base[b] + atomicAdd(&g_cursor[b], 1u)] = make_uint2((uint32_t)t, nfull);
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd((uint32_t *)arrive + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last)
        for (int b = tid; b < kOrderWs; b += kOrderThreads) ws[b] = 0;
}

int segment_items(const uint32_t *maxlast, int tiles, uint2 *items, uint32_t *n_items, uint32_t *ws, cudaStream_t s, bool debug) {
    KernelTimer kt("tile_order", s);
    const int grid = max(1, min((tiles + kOrderThreads - 1) / kOrderThreads, min(32, sm_count())));
    MB_CUDA(launch_cooperative(segment_items_kernel, grid, kOrderThreads, s, maxlast, tiles, items, n_items, ws));
    return check_launch("segment_items", debug, s);
}

int validate_raster_inputs(const mb_raster_inputs *in, const char *who, bool need_opacities, bool need_arrays) {
    MB_REQUIRE(in != nullptr, "%s: null inputs", who);
    MB_REQUIRE(in->num_points >= 0 && in->image_width > 0 && in->image_height > 0, "%s: bad sizes P=%d W=%d H=%d", who,
               in->num_points, in->image_width, in->image_height);
    if (!need_arrays) {      // stages that only read the opaque state (binning + tile kernels): camera, sizes, background
        MB_REQUIRE(in->background && in->viewmatrix && in->projmatrix && in->campos, "%s: camera tensors missing", who);
        MB_REQUIRE(in->tanfov_dev || (in->tanfovx > 0.f && in->tanfovy > 0.f), "%s: tanfov must be positive", who);
        return MB_OK;
    }
    MB_REQUIRE((in->colors_precomp != nullptr) != (in->shs != nullptr),
               "%s: Please provide excatly one of either SHs or precomputed colors!", who);
    const bool sr = in->scales != nullptr && in->rotations != nullptr;
    MB_REQUIRE((in->cov3D_precomp != nullptr) != sr && ((in->scales != nullptr) == (in->rotations != nullptr)),
               "%s: Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!", who);
    MB_REQUIRE(in->num_points == 0 || (in->means3D && (in->opacities || !need_opacities)), "%s: means3D / opacities missing", who);
    MB_REQUIRE(in->background && in->viewmatrix && in->projmatrix && in->campos, "%s: camera tensors missing", who);
    if (in->shs) {
        MB_REQUIRE(in->sh_degree >= 0 && in->sh_degree <= 3, "%s: sh_degree %d not in 0..3", who, in->sh_degree);
        MB_REQUIRE(in->sh_coeffs >= (in->sh_degree + 1) * (in->sh_degree + 1) && in->sh_coeffs <= 16,
                   "%s: shs has %d coefficients, degree %d needs %d (max 16)", who, in->sh_coeffs, in->sh_degree,
                   (in->sh_degree + 1) * (in->sh_degree + 1));
    }
    MB_REQUIRE(in->tanfov_dev || (in->tanfovx > 0.f && in->tanfovy > 0.f), "%s: tanfov must be positive", who);
    return MB_OK;
}

static int tile_bits(int tiles) {
    int b = 0;
    while ((1 << b) < tiles) ++b;
    return b < 1 ? 1 : b;
}

// emission + per-tile ordering + ranges; shared with the render entry point (raster_blend.cu)
int build_instances(const mb_raster_inputs *in, const RasterDims &d, const GeomState &g, const BinningState &b,
                    const ImageState &im, int64_t capacity, cudaStream_t s) {
    const bool dbg = in->debug != 0;
    const int grid_p = max(1, min((d.P + kEmitChunk - 1) / kEmitChunk, sm_count() * 8));
    // (im.ranges was cleared by the caller)
    // the emission kernels count the digits of the tile ids they write (histogram of the partition by tile below): its
    // workspace is cleared here, in front of them
    SortWorkspace ws = carve_sort_workspace(b.sort_ws, capacity > 0 ? capacity : 1);
    const int key_bits = tile_bits(d.tiles), key_passes = sort_passes(0, key_bits);
    MB_CUDA(cudaMemsetAsync(ws.zeroed, 0, ws.zeroed_bytes, s));
    {
    KernelTimer kt("emit_instances", s);
    emit_instances_kernel<<<grid_p, 256, 0, s>>>(d.P, d.gx, g.sorted_idx, g.tiles_touched, g.rect, g.counters, g.scan_status,
                                                 capacity, b.tile_a, b.gid_a, g.big_list, key_passes, ws.hist);
    }
    int rc = check_launch("emit_instances", dbg, s);
    if (rc) return rc;
    {
    KernelTimer kt("emit_big", s);
    emit_big_kernel<<<sm_count() * 2, 256, 0, s>>>(d.gx, g.sorted_idx, g.tiles_touched, g.rect, g.counters, capacity, g.big_list,
                                                  b.tile_a, b.gid_a, key_passes, ws.hist);
    }
    rc = check_launch("emit_big", dbg, s);
    if (rc) return rc;
    // stable partition by tile id; its last pass leaves every tile's run of the sorted list in im.ranges (encoded: decoded by
    // tile_order, which the caller launches next)
    return radix_sort_pairs(b.tile_a, b.gid_a, b.tile_b, b.gid_b, -1, g.counters + kCntRendered, capacity, 0, key_bits, ws, s, dbg, true,
                            im.ranges);
}

}  // namespace mb

using namespace mb;

extern "C" size_t mb_raster_geom_bytes(int32_t num_points) { return GeomState::carve(nullptr, num_points).bytes; }

extern "C" size_t mb_raster_binning_bytes(int64_t capacity, int32_t w, int32_t h) {
    const int tiles = ((w + kTile - 1) / kTile) * ((h + kTile - 1) / kTile);
    return BinningState::carve(nullptr, capacity, tiles).bytes;
}

extern "C" size_t mb_raster_image_bytes(int32_t w, int32_t h) { return ImageState::carve(nullptr, w, h).bytes; }

extern "C" int mb_raster_forward_geom(const mb_raster_inputs *in, void *geom, size_t geom_bytes, int32_t *radii,
                                      int64_t *num_rendered_host, mb_stream_t stream) {
    int rc = validate_raster_inputs(in, "mb_raster_forward_geom");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterDims d = raster_dims(in);
    MB_REQUIRE(geom != nullptr && (d.P == 0 || radii != nullptr), "mb_raster_forward_geom: null geom / radii");
    GeomState g = GeomState::carve(geom, d.P);
    if (geom_bytes < g.bytes) {
        set_error("mb_raster_forward_geom: geom buffer has %zu bytes, needs %zu", geom_bytes, g.bytes);
        return MB_ERR_WORKSPACE;
    }
    const bool dbg = in->debug != 0;
    // counters + look-back words of the instance-offset scan + histograms / cursors / look-back words of the depth sort
    // (contiguous): the projection kernel counts the digits of the depth keys it writes, so the sort below is its passes only
    MB_CUDA(cudaMemsetAsync(g.counters, 0, g.zeroed_bytes(d.P), s));
    if (d.P > 0) {
        SortWorkspace ws = carve_sort_workspace(g.sort_ws, d.P);
        PreArgs a;
        a.depth_hist = ws.hist;
        a.P = d.P; a.W = d.W; a.H = d.H; a.gx = d.gx; a.gy = d.gy; a.deg = in->sh_degree; a.M = in->sh_coeffs;
        a.tanx = in->tanfovx; a.tany = in->tanfovy; a.focx = d.focx; a.focy = d.focy; a.scale_mod = in->scale_modifier;
        a.means3D = in->means3D; a.opac = in->opacities; a.cov3D_precomp = in->cov3D_precomp; a.scales = in->scales;
        a.rots = in->rotations; a.shs = in->shs; a.colors = in->colors_precomp; a.view = in->viewmatrix; a.proj = in->projmatrix; a.campos = in->campos;
        a.tanfov_dev = in->tanfov_dev;
        a.g = g; a.radii = radii;
        const int grid = (d.P + 255) / 256;
        {
        KernelTimer kt("preprocess", s);
        if (in->cov3D_precomp) {
            if (in->shs) preprocess_kernel<true, true><<<grid, 256, 0, s>>>(a);
            else preprocess_kernel<true, false><<<grid, 256, 0, s>>>(a);
        } else {
            if (in->shs) preprocess_kernel<false, true><<<grid, 256, 0, s>>>(a);
            else preprocess_kernel<false, false><<<grid, 256, 0, s>>>(a);
        }
        }
        rc = check_launch("preprocess", dbg, s);
        if (rc) return rc;
        // depth order (stable; culled Gaussians carry key 0xffffffff and sink to the end)
        rc = radix_sort_pairs(g.depth_key, g.ident, g.sorted_key, g.sorted_idx, d.P, nullptr, d.P, 0, 32, ws, s, dbg, true);
        if (rc) return rc;
    }
    if (num_rendered_host) {
        // 64-bit host slot, 32-bit device counter: clear the high word first
        *num_rendered_host = 0;
        MB_CUDA(cudaMemcpyAsync(num_rendered_host, g.counters + kCntRendered, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    return MB_OK;
}

extern "C" int mb_raster_query(const void *geom, int64_t *num_rendered, int64_t *num_visible, int32_t *overflow,
                               mb_stream_t stream) {
    MB_REQUIRE(geom != nullptr, "mb_raster_query: null geom");
    uint32_t host[kNumCounters];
    GeomState g = GeomState::carve(const_cast<void *>(geom), 1);
    MB_CUDA(cudaMemcpyAsync(host, g.counters, sizeof(host), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (num_rendered) *num_rendered = host[kCntRendered];
    if (num_visible) *num_visible = host[kCntVisible];
    if (overflow) *overflow = (int32_t)host[kCntOverflow];
    return MB_OK;
}

__global__ void mark_visible_kernel(const float *__restrict__ means3D, int P, const float *__restrict__ view,
                                    uint8_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float z = view[2] * means3D[3 * i] + view[6] * means3D[3 * i + 1] + view[10] * means3D[3 * i + 2] + view[14];
    out[i] = z > kNearZ ? 1 : 0;
}

extern "C" int mb_mark_visible(const float *means3D, int32_t num_points, const float *viewmatrix, const float *projmatrix,
                               uint8_t *out, mb_stream_t stream) {
    MB_REQUIRE(num_points >= 0, "mb_mark_visible: negative count");
    if (num_points == 0) return MB_OK;
    MB_REQUIRE(means3D && viewmatrix && projmatrix && out, "mb_mark_visible: null pointer");
    mark_visible_kernel<<<(num_points + 255) / 256, 256, 0, (cudaStream_t)stream>>>(means3D, num_points, viewmatrix, out);
    return check_launch("mark_visible", false, (cudaStream_t)stream);
}
