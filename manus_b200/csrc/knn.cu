// distCUDA2: mean squared distance from every point to its 3 nearest other points (exact).
// Replaces simple_knn._C.distCUDA2 (call site /root/reference/src/models/gaussian.py:110).
//
// Scheme: 30-bit Morton codes on the bounding box -> radix sort (sort_scan.cu) -> points gathered into Morton order
// (one coalesced float4 stream) -> axis-aligned bounds of every run of 1024 sorted points -> one thread per point:
// the +-3 Morton neighbours give a rejection radius, then every box closer than the current 3rd-best distance is
// scanned.  A warp holds 32 Morton-adjacent points, so its lanes visit almost the same boxes and the box reads are
// broadcasts.  Pruning is conservative (a box is skipped only if strictly farther than the 3rd best), so the result
// is the exact 3-NN answer up to fp32 rounding of the squared distances.
#include <float.h>

#include "sort_scan.cuh"

namespace mb {

constexpr int kBox = 1024;
constexpr int kFine = 64;     // second, finer level of boxes inside every kBox run

struct KnnWorkspace {
    uint32_t *minmax;  // 6 ordered-uint encoded floats: min xyz, max xyz
    uint32_t *codes, *ident, *codes_sorted, *order;
    float4 *pts;       // Morton order: x, y, z, original index (bits)
    float4 *box_lo, *box_hi;
    float4 *fine_lo, *fine_hi;   // bounds of every run of kFine sorted points (nearest-point search only)
    void *sort_ws;
    size_t bytes;
    static KnnWorkspace carve(void *p, int64_t n) {
        Carver c(p);
        KnnWorkspace w;
        const size_t m = (size_t)(n > 0 ? n : 1), boxes = (m + kBox - 1) / kBox;
        w.minmax = c.take<uint32_t>(8);
        w.codes = c.take<uint32_t>(m);
        w.ident = c.take<uint32_t>(m);
        w.codes_sorted = c.take<uint32_t>(m);
        w.order = c.take<uint32_t>(m);
        w.pts = c.take<float4>(m);
        w.box_lo = c.take<float4>(boxes);
        w.box_hi = c.take<float4>(boxes);
        w.fine_lo = c.take<float4>(boxes * (kBox / kFine));
        w.fine_hi = c.take<float4>(boxes * (kBox / kFine));
        w.sort_ws = c.take<char>(sort_workspace_bytes((int64_t)m));
        w.bytes = c.off;
        return w;
    }
};

// order-preserving float <-> uint mapping for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f2o(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void knn_init_kernel(uint32_t *minmax) {
    if (threadIdx.x < 3) minmax[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) minmax[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) knn_bounds_kernel(const float *__restrict__ pts, int n, uint32_t *minmax) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = pts[3 * (size_t)i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&minmax[k], f2o(lo[k]));
            atomicMax(&minmax[3 + k], f2o(hi[k]));
        }
}

__device__ __forceinline__ uint32_t spread10(uint32_t v) {   // 10 bits -> every third bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256) knn_morton_kernel(const float *__restrict__ pts, int n, const uint32_t *__restrict__ minmax,
                                                         uint32_t *__restrict__ codes, uint32_t *__restrict__ ident) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = o2f(minmax[k]), hi = o2f(minmax[3 + k]);
        const float ext = hi - lo;
        float r = ext > 0.f ? (pts[3 * (size_t)i + k] - lo) / ext : 0.f;
        r = fminf(fmaxf(r * 1023.f, 0.f), 1023.f);
        code |= spread10((uint32_t)r) << (2 - k);
    }
    codes[i] = code;
    ident[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) knn_gather_kernel(const float *__restrict__ pts, int n, const uint32_t *__restrict__ order,
                                                         float4 *__restrict__ sorted) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = order[j];
    sorted[j] = make_float4(pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2], __uint_as_float(i));
}

__global__ void __launch_bounds__(256) knn_box_kernel(const float4 *__restrict__ sorted, int n, float4 *__restrict__ box_lo,
                                                      float4 *__restrict__ box_hi, int bsz = kBox) {
    __shared__ float slo[8][3], shi[8][3];
    const int box = blockIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int j = box * bsz + threadIdx.x; j < min(n, (box + 1) * bsz); j += 256) {
        const float4 p = sorted[j];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) { slo[threadIdx.x >> 5][k] = lo[k]; shi[threadIdx.x >> 5][k] = hi[k]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
#pragma unroll
            for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], slo[w][k]); hi[k] = fmaxf(hi[k], shi[w][k]); }
        box_lo[box] = make_float4(lo[0], lo[1], lo[2], 0.f);
        box_hi[box] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
}

__device__ __forceinline__ void best3_insert(float d, float &b0, float &b1, float &b2) {
    if (d < b2) {
        if (d < b1) {
            b2 = b1;
            if (d < b0) { b1 = b0; b0 = d; }
            else b1 = d;
        } else b2 = d;
    }
}

__device__ __forceinline__ float dist2(const float4 &a, const float4 &b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(256) knn_search_kernel(const float4 *__restrict__ sorted, int n, const float4 *__restrict__ box_lo,
                                                         const float4 *__restrict__ box_hi, int boxes, float *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float4 p = sorted[j];
    // rejection radius from the Morton neighbourhood
    float r0 = FLT_MAX, r1 = FLT_MAX, r2 = FLT_MAX;
    for (int k = max(0, j - 3); k <= min(n - 1, j + 3); ++k)
        if (k != j) best3_insert(dist2(p, sorted[k]), r0, r1, r2);
    const float reject = r2;
    float b0 = FLT_MAX, b1 = FLT_MAX, b2 = FLT_MAX;
    for (int b = 0; b < boxes; ++b) {
        const float4 lo = box_lo[b], hi = box_hi[b];
        const float ex = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.f), ey = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.f),
                    ez = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.f);
        const float bd = ex * ex + ey * ey + ez * ez;
        if (bd > reject || bd > b2) continue;
        const int end = min(n, (b + 1) * kBox);
        for (int k = b * kBox; k < end; ++k)
            if (k != j) best3_insert(dist2(p, sorted[k]), b0, b1, b2);
    }
    // fewer than 4 points in total: missing neighbours count as distance 0
    if (b0 == FLT_MAX) b0 = 0.f;
    if (b1 == FLT_MAX) b1 = 0.f;
    if (b2 == FLT_MAX) b2 = 0.f;
    out[__float_as_uint(p.w)] = (b0 + b1 + b2) / 3.0f;
}

// ---------------------------------------------------------------------------------------------------------
// Nearest reference point of every query point (exact): the "contact distance" of the composite scene.
// Replaces get_contact_dist (src/utils/gaussian_utils.py:521-554: an O(N*M) taichi loop, first index of the minimum) and
// get_contact_map (:514-518: chunked torch.cdist().min()).  Both sets are Morton-ordered on their common bounding box;
// a query starts from the reference points around its own Morton position (a tight upper bound), then descends only into
// the 1024-point boxes, and inside them the 64-point boxes, that can still hold something closer.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nearest_consider(float d2, uint32_t idx, float &best_d2, float &best_s, uint32_t &best_i) {
    // the reference compares sqrt(d2) with '<' in ascending reference index: emulate that order-independently
    if (d2 > best_d2 * 1.000001f) return;
    const float sq = sqrtf(d2);
    if (sq < best_s || (sq == best_s && idx < best_i)) {
        best_s = sq;
        best_i = idx;
        best_d2 = fminf(best_d2, d2);
    }
}

__global__ void __launch_bounds__(256) nearest_search_kernel(const float4 *__restrict__ queries, int nq, const uint32_t *__restrict__ qcodes,
                                                             const float4 *__restrict__ refs, int nr, const uint32_t *__restrict__ rcodes,
                                                             const float4 *__restrict__ box_lo, const float4 *__restrict__ box_hi, int boxes,
                                                             const float4 *__restrict__ fine_lo, const float4 *__restrict__ fine_hi,
                                                             float *__restrict__ out_dist, int32_t *__restrict__ out_idx) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nq) return;
    const float4 p = queries[j];
    // position of the query's Morton code among the reference codes
    const uint32_t code = qcodes[j];
    int lo = 0, hi = nr;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rcodes[mid] < code) lo = mid + 1; else hi = mid;
    }
    float best_d2 = FLT_MAX, best_s = FLT_MAX;
    uint32_t best_i = 0xffffffffu;
    for (int k = max(0, lo - 4); k < min(nr, lo + 4); ++k) {
        const float4 r = refs[k];
        nearest_consider(dist2(p, r), __float_as_uint(r.w), best_d2, best_s, best_i);
    }
    for (int b = 0; b < boxes; ++b) {
        const float4 blo = box_lo[b], bhi = box_hi[b];
        const float ex = fmaxf(fmaxf(blo.x - p.x, p.x - bhi.x), 0.f), ey = fmaxf(fmaxf(blo.y - p.y, p.y - bhi.y), 0.f),
                    ez = fmaxf(fmaxf(blo.z - p.z, p.z - bhi.z), 0.f);
        if (ex * ex + ey * ey + ez * ez > best_d2 * 1.000001f) continue;
        const int fend = min((nr + kFine - 1) / kFine, (b + 1) * (kBox / kFine));
        for (int f = b * (kBox / kFine); f < fend; ++f) {
            const float4 flo = fine_lo[f], fhi = fine_hi[f];
            const float fx = fmaxf(fmaxf(flo.x - p.x, p.x - fhi.x), 0.f), fy = fmaxf(fmaxf(flo.y - p.y, p.y - fhi.y), 0.f),
                        fz = fmaxf(fmaxf(flo.z - p.z, p.z - fhi.z), 0.f);
            if (fx * fx + fy * fy + fz * fz > best_d2 * 1.000001f) continue;
            const int end = min(nr, (f + 1) * kFine);
            for (int k = f * kFine; k < end; ++k) {
                const float4 r = refs[k];
                nearest_consider(dist2(p, r), __float_as_uint(r.w), best_d2, best_s, best_i);
            }
        }
    }
    const uint32_t qi = __float_as_uint(p.w);
    out_dist[qi] = best_s;
    out_idx[qi] = (int32_t)best_i;
}

}  // namespace mb

using namespace mb;

extern "C" size_t mb_nearest_workspace_bytes(int32_t num_queries, int32_t num_refs) {
    return KnnWorkspace::carve(nullptr, num_queries).bytes + KnnWorkspace::carve(nullptr, num_refs).bytes;
}

extern "C" int mb_nearest_point(const float *queries, int32_t nq, const float *refs, int32_t nr, float *out_dist, int32_t *out_index,
                                void *workspace, size_t workspace_bytes, mb_stream_t stream) {
    MB_REQUIRE(nq >= 0 && nr > 0, "mb_nearest_point: need at least one reference point (nq=%d nr=%d)", nq, nr);
    if (nq == 0) return MB_OK;
    MB_REQUIRE(queries && refs && out_dist && out_index && workspace, "mb_nearest_point: null pointer");
    if (workspace_bytes < mb_nearest_workspace_bytes(nq, nr)) {
        set_error("mb_nearest_point: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    KnnWorkspace wq = KnnWorkspace::carve(workspace, nq);
    KnnWorkspace wr = KnnWorkspace::carve(reinterpret_cast<char *>(workspace) + wq.bytes, nr);
    cudaStream_t s = (cudaStream_t)stream;
    const int gq = (nq + 255) / 256, gr = (nr + 255) / 256, boxes = (nr + kBox - 1) / kBox;
    KernelTimer kt("nearest_point", s);
    // common bounding box -> comparable Morton codes
    knn_init_kernel<<<1, 32, 0, s>>>(wr.minmax);
    knn_bounds_kernel<<<min(gq, sm_count() * 8), 256, 0, s>>>(queries, nq, wr.minmax);
    knn_bounds_kernel<<<min(gr, sm_count() * 8), 256, 0, s>>>(refs, nr, wr.minmax);
    knn_morton_kernel<<<gr, 256, 0, s>>>(refs, nr, wr.minmax, wr.codes, wr.ident);
    knn_morton_kernel<<<gq, 256, 0, s>>>(queries, nq, wr.minmax, wq.codes, wq.ident);
    int rc = check_launch("nearest_morton", false, s);
    if (rc) return rc;
    rc = radix_sort_pairs(wr.codes, wr.ident, wr.codes_sorted, wr.order, nr, nullptr, nr, 0, 30, carve_sort_workspace(wr.sort_ws, nr), s, false);
    if (rc) return rc;
    rc = radix_sort_pairs(wq.codes, wq.ident, wq.codes_sorted, wq.order, nq, nullptr, nq, 0, 30, carve_sort_workspace(wq.sort_ws, nq), s, false);
    if (rc) return rc;
    knn_gather_kernel<<<gr, 256, 0, s>>>(refs, nr, wr.order, wr.pts);
    knn_gather_kernel<<<gq, 256, 0, s>>>(queries, nq, wq.order, wq.pts);
    knn_box_kernel<<<boxes, 256, 0, s>>>(wr.pts, nr, wr.box_lo, wr.box_hi, kBox);
    knn_box_kernel<<<(nr + kFine - 1) / kFine, 256, 0, s>>>(wr.pts, nr, wr.fine_lo, wr.fine_hi, kFine);
    nearest_search_kernel<<<gq, 256, 0, s>>>(wq.pts, nq, wq.codes_sorted, wr.pts, nr, wr.codes_sorted, wr.box_lo, wr.box_hi, boxes, wr.fine_lo,
                                             wr.fine_hi, out_dist, out_index);
    return check_launch("nearest_search", false, s);
}

extern "C" size_t mb_knn_workspace_bytes(int32_t num_points) { return KnnWorkspace::carve(nullptr, num_points).bytes; }

extern "C" int mb_dist2_knn3(const float *points, int32_t n, float *out, void *workspace, size_t workspace_bytes,
                             mb_stream_t stream) {
    MB_REQUIRE(n >= 0, "mb_dist2_knn3: negative count");
    if (n == 0) return MB_OK;
    MB_REQUIRE(points && out && workspace, "mb_dist2_knn3: null pointer");
    KnnWorkspace w = KnnWorkspace::carve(workspace, n);
    if (workspace_bytes < w.bytes) {
        set_error("mb_dist2_knn3: workspace has %zu bytes, needs %zu", workspace_bytes, w.bytes);
        return MB_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = (n + 255) / 256, boxes = (n + kBox - 1) / kBox;
    knn_init_kernel<<<1, 32, 0, s>>>(w.minmax);
    knn_bounds_kernel<<<min(grid, sm_count() * 8), 256, 0, s>>>(points, n, w.minmax);
    knn_morton_kernel<<<grid, 256, 0, s>>>(points, n, w.minmax, w.codes, w.ident);
    int rc = check_launch("knn_morton", false, s);
    if (rc) return rc;
    SortWorkspace sw = carve_sort_workspace(w.sort_ws, n);
    rc = radix_sort_pairs(w.codes, w.ident, w.codes_sorted, w.order, n, nullptr, n, 0, 30, sw, s, false);
    if (rc) return rc;
    knn_gather_kernel<<<grid, 256, 0, s>>>(points, n, w.order, w.pts);
    knn_box_kernel<<<boxes, 256, 0, s>>>(w.pts, n, w.box_lo, w.box_hi);
    knn_search_kernel<<<grid, 256, 0, s>>>(w.pts, n, w.box_lo, w.box_hi, boxes, out);
    return check_launch("knn_search", false, s);
}
