// Per-tile alpha blending, forward and backward (A.4), plus the per-Gaussian backward of the
// projection.
//
// Work item = one 16x16 tile (kWarps = 8) or one 16x8 half tile (kWarps = 4); a warp owns an 8x4 pixel block.  The tile's
// depth-sorted id list is staged into shared memory by the TMA engine (cp.async.bulk + mbarrier, 3-slot ring, two batches
// ahead); every thread then gathers the 48-B record of one listed Gaussian with three 128-bit cp.async loads (records
// are written once per Gaussian by the projection kernel and stay L2 resident), double buffered against the blending of
// the previous batch.  Only the part of a list that is actually consumed is ever gathered: the forward stops a tile
// when all its pixels are saturated, the backward starts at the tile's deepest last contributor.
//
// Per pixel the arithmetic is the reference's sequential front-to-back product, so results do not depend on the
// schedule.  The inner loops are warp-convergent: per round, the 32 lanes box-test 32 list entries against the warp's
// pixel block (one ballot), and only the survivors are evaluated per pixel; a survivor whose power is below its cut-off
// on all 32 pixels costs no exponential (warp vote).  The backward sums each of its nine per-Gaussian partials over the
// warp with a 12-shuffle reduce-scatter and issues them as one vector of RED.ADD.F32 from nine lanes.
#include <stdlib.h>

#include "raster_state.cuh"
#include "project_bwd.cuh"

namespace mb {

constexpr int kAccStride = 12;              // floats per Gaussian in the gradient accumulator
// accumulator slots: moments of q = G * dL/dalpha over the pixels, d = mean2D - pixel:
//   0 sum q dx | 1 sum q dy | 2 sum q dx^2 | 3 sum q dx dy | 4 sum q dy^2 | 5 sum q (= dL/dopacity) | 6,7,8 dL/dcolour
// The consumer (project_backward) turns them into dL/dmean2D and dL/dconic with the Gaussian's conic and opacity, once per
// Gaussian instead of once per (Gaussian, pixel).

int build_instances(const mb_raster_inputs *in, const RasterDims &d, const GeomState &g, const BinningState &b,
                    const ImageState &im, int64_t capacity, cudaStream_t s);

// Optional per-CTA timeline of the tile kernels (tools/cta_trace.py): compiled in with -DMB_TRACE_CTA only.
#ifdef MB_TRACE_CTA
constexpr int kTraceMax = 1 << 17;
__device__ ulonglong2 g_trace[2][kTraceMax];   // [kernel][cta] = (start ns, end ns)
__device__ uint32_t g_trace_work[2][kTraceMax];
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MB_TRACE_BEGIN() const unsigned long long trace_t0 = trace_now()
#define MB_TRACE_END(k, work)                                                                       \
    do {                                                                                            \
        __syncthreads();                                                                            \
        if (threadIdx.x == 0 && blockIdx.x < kTraceMax) {                                           \
            g_trace[k][blockIdx.x] = make_ulonglong2(trace_t0, trace_now());                        \
            g_trace_work[k][blockIdx.x] = (uint32_t)(work);                                         \
        }                                                                                           \
    } while (0)
#else
#define MB_TRACE_BEGIN()
#define MB_TRACE_END(k, work)
#endif

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

#ifndef MB_GATHER_AHEAD
#define MB_GATHER_AHEAD 1
#endif
constexpr int kPre = MB_GATHER_AHEAD;       // batches of records in flight ahead of the one being blended
constexpr int kRecSlots = kPre + 1, kIdSlots = kPre + 2;

template <int kWarps>
struct StageSmem {
    static constexpr int B = kWarps * 32;       // list entries per batch = threads per CTA
    Record rec[kRecSlots][B];                   // gathered records: the batch being blended + kPre batches in flight
    uint32_t ids[kIdSlots][B + 4];              // TMA-staged slices of the id list (16-B aligned source => up to 3 ids of slack)
    uint64_t bar[kIdSlots];
    uint32_t red;                               // per-CTA scratch (max last contributor)
};

// Streams a tile's id list through shared memory: sequence step i handles batch i (forward) or nb-1-i (backward).
// Pipeline: the id slice of batch i + kPre + 1 is requested from the TMA engine and the records of batch i + kPre are
// gathered (three 128-bit cp.async per thread, L2 hits) while batch i is blended.
template <int kWarps>
struct ListStager {
    static constexpr int B = kWarps * 32;
    StageSmem<kWarps> &sm;
    const uint32_t *list;      // sorted id list (whole frame)
    const Record *rec;         // per-Gaussian records
    uint32_t first;            // index of the tile's first instance in `list`
    int len, nb;
    bool reverse;

    __device__ __forceinline__ int batch_of(int i) const { return reverse ? nb - 1 - i : i; }

    __device__ __forceinline__ void issue_ids(int i) {   // one elected thread
        const int b = batch_of(i), cnt = min(B, len - b * B);
        const uint32_t start = first + (uint32_t)(b * B), o = start & 3u;
        const uint32_t bytes = ((o + (uint32_t)cnt) * 4u + 15u) & ~15u;
        const int slot = i % kIdSlots;
        mbar_expect_tx(&sm.bar[slot], bytes);
        bulk_g2s(&sm.ids[slot][0], list + (start - o), bytes, &sm.bar[slot]);
    }
    __device__ __forceinline__ void gather(int i) {      // all threads; one commit group per call
        const int b = batch_of(i), cnt = min(B, len - b * B);
        const uint32_t o = (first + (uint32_t)(b * B)) & 3u;
        const int slot = i % kIdSlots;
        mbar_wait(&sm.bar[slot], (uint32_t)((i / kIdSlots) & 1));
        if ((int)threadIdx.x < cnt) {
            const char *src = reinterpret_cast<const char *>(rec + sm.ids[slot][o + threadIdx.x]);
            char *dst = reinterpret_cast<char *>(&sm.rec[i % kRecSlots][threadIdx.x]);
            cp_async16(dst, src);
            cp_async16(dst + 16, src + 16);
            cp_async16(dst + 32, src + 32);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void prologue() {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < kIdSlots; ++k) mbar_init(&sm.bar[k], 1);
            mbar_fence_init();
            sm.red = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k <= kPre && k < nb; ++k) issue_ids(k);
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            if (k < nb) gather(k);
            else cp_async_commit();
        }
    }
    // top of sequence step i: keep the pipeline full, then make batch i visible to every thread
    __device__ __forceinline__ void advance(int i) {
        if (threadIdx.x == 0 && i + kPre + 1 < nb) issue_ids(i + kPre + 1);
        if (i + kPre < nb) gather(i + kPre);
        else cp_async_commit();
        cp_async_wait<kPre>();
        __syncthreads();
    }
    // leaving after step i with copies possibly in flight: nothing may land in shared memory after the CTA is gone
    __device__ __forceinline__ void drain(int i) {
        cp_async_wait<0>();
        if (threadIdx.x == 0)
            for (int k = i + kPre + 1; k < nb && k <= i + kPre + 1; ++k) mbar_wait(&sm.bar[k % kIdSlots], (uint32_t)((k / kIdSlots) & 1));
    }
    __device__ __forceinline__ int count(int i) const { return min(B, len - batch_of(i) * B); }
    __device__ __forceinline__ const float4 *records(int i) const { return reinterpret_cast<const float4 *>(&sm.rec[i % kRecSlots][0]); }
    __device__ __forceinline__ const uint32_t *ids(int i) const { return &sm.ids[i % kIdSlots][(first + (uint32_t)(batch_of(i) * B)) & 3u]; }
};

#ifndef MB_FWD_FAST_EXP
#define MB_FWD_FAST_EXP 1
#endif
#ifndef MB_FWD_GROUP
#define MB_FWD_GROUP 4
#endif
constexpr int kGroup = MB_FWD_GROUP;   // survivors evaluated together in the forward (independent alpha evaluations)

// exp(x) for x <= 0 (|x| < 100) with one MUFU.EX2, like expf(), in 6 instructions instead of expf()'s range split: exp(x) = 2^t 2^r with
// t = fl(x log2e) and r = the rounding residual of that product plus x (log2e - fl(log2e)); |r| < 1e-6, so 2^r = 1 + r ln2.
__device__ __forceinline__ float exp_neg(float x) {
    const float L = 1.4426950216293334961f, Llo = 1.925963033500011079e-08f, ln2 = 0.693147182464599609375f;
    const float t = x * L;
    float r = fmaf(x, L, -t);
    r = fmaf(x, Llo, r);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return fmaf(e, r * ln2, e);
}

// pixel of a thread: the work item covers 8/kWarps... see kernel comments; vw = virtual warp index inside the 16x16 tile
__device__ __forceinline__ void pixel_of_thread(int tile, int gx, int vw, int lane, int &px, int &py) {
    px = (tile % gx) * kTile + (vw & 1) * 8 + (lane & 7);
    py = (tile / gx) * kTile + (vw >> 1) * 4 + (lane >> 3);
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 24 / kWarps) blend_forward_kernel(
    const Record *__restrict__ recs, const uint32_t *__restrict__ list, const uint2 *__restrict__ ranges,
    const uint32_t *__restrict__ order, int W, int H, int gx, const float *__restrict__ bg, float *__restrict__ out_color,
    float *__restrict__ final_T, uint32_t *__restrict__ n_contrib, uint32_t *__restrict__ tile_maxlast,
    float4 *__restrict__ ckpt) {
    constexpr int kSplit = 8 / kWarps;
    static_assert(kSeg % (kWarps * 32) == 0, "segment length must be a multiple of the batch size");
    MB_TRACE_BEGIN();
    __shared__ __align__(128) StageSmem<kWarps> sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int item = blockIdx.x / kSplit, sub = blockIdx.x % kSplit;
    const int tile = order ? (int)order[item] : item;
    const uint2 range = ranges[tile];
    int px, py;
    pixel_of_thread(tile, gx, sub * kWarps + warp, lane, px, py);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const float bcx = (float)(px - (lane & 7)) + 3.5f, bcy = (float)(py - (lane >> 3)) + 1.5f;   // centre of the warp's pixel block

    ListStager<kWarps> st{sm, list, recs, range.x, (int)(range.y - range.x), 0, false};
    st.nb = (st.len + st.B - 1) / st.B;
    st.prologue();

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t last = 0;
    bool done = !inside;
    bool wdone = __all_sync(0xffffffffu, done);
    const int ck_idx = (sub * kWarps + warp) * 32 + lane;   // pixel index inside the 16x16 tile
    for (int i = 0; i < st.nb; ++i) {
        st.advance(i);
        const int cnt = st.count(i);
        const float4 *r = st.records(i);
        const uint32_t base = (uint32_t)(i * st.B);
        // state in front of list position `base`, once per kSeg entries: where a backward segment starts
        if (i > 0 && base % kSeg == 0)
            ckpt[BinningState::ckpt_slot((uint32_t)tile, range.x, base / kSeg - 1) * kTilePixels + ck_idx] = make_float4(T, C0, C1, C2);
        // 32 list entries per round: lane k box-tests entry j0+k against the warp's 8x4 pixel block, the survivors are
        // blended in list order with the exact reference arithmetic
        for (int j0 = 0; j0 < cnt && !wdone; j0 += 32) {
            bool hit = false;
            if (j0 + lane < cnt) {
                const float4 ra = r[3 * (j0 + lane)], rc = r[3 * (j0 + lane) + 2];
                hit = fabsf(ra.x - bcx) <= rc.z + 3.5f && fabsf(ra.y - bcy) <= rc.w + 1.5f;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                // up to kGroup survivors per step: everything that does not depend on the running transmittance (record
                // fetch, power, exponential, alpha) is evaluated for all four first (independent instruction streams),
                // then the four are applied in list order
                float alpha[kGroup], cr[kGroup], cg[kGroup], cb[kGroup];
                uint32_t pos1[kGroup];
                bool valid[kGroup];
#pragma unroll
                for (int q = 0; q < kGroup; ++q) {
                    const bool has = m != 0;
                    const int j = j0 + (has ? __ffs(m) - 1 : 0);
                    m &= m - 1;
                    const float4 ra = r[3 * j], rb = r[3 * j + 1];
                    const float blue = r[3 * j + 2].x;
                    const float dx = ra.x - fx, dy = ra.y - fy;
                    const float power = -0.5f * (ra.z * dx * dx + rb.x * dy * dy) - ra.w * dx * dy;
#if MB_FWD_FAST_EXP
                    alpha[q] = fminf(kAlphaMax, rb.y * exp_neg(power));      // power > 0 gives garbage here and is rejected below
#else
                    alpha[q] = fminf(kAlphaMax, rb.y * expf(power));
#endif
                    valid[q] = has && power <= 0.0f && alpha[q] >= kAlphaMin;
                    cr[q] = rb.z; cg[q] = rb.w; cb[q] = blue;
                    pos1[q] = base + (uint32_t)j + 1u;
                }
#pragma unroll
                for (int q = 0; q < kGroup; ++q) {
                    const bool ok = valid[q] && !done;
                    const float test_T = T * (1.0f - alpha[q]);
                    const bool stop = ok && test_T < kTMin;
                    if (ok && !stop) {
                        const float w = alpha[q] * T;
                        C0 += cr[q] * w;
                        C1 += cg[q] * w;
                        C2 += cb[q] * w;
                        T = test_T;
                        last = pos1[q];
                    }
                    done = done || stop;
                }
                if (__all_sync(0xffffffffu, done)) {
                    wdone = true;
                    break;
                }
            }
        }
        // all reads of this batch are complete after the barrier; it also counts finished pixels (early exit)
        const int ndone = __syncthreads_count(done);
        if (ndone == st.B) {
            st.drain(i);
            break;
        }
    }
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && wmax) atomicMax(&sm.red, wmax);
    // tiles whose list spans more than one segment: final (T, C) in the tile's spare checkpoint slot, for the backward
    if (st.len > kSeg)
        ckpt[BinningState::ckpt_slot((uint32_t)tile, range.x, (uint32_t)((st.len + kSeg - 1) / kSeg - 1)) * kTilePixels + ck_idx] =
            make_float4(T, C0, C1, C2);
    if (inside) {
        const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = C0 + T * bg[0];
        out_color[plane + pix] = C1 + T * bg[1];
        out_color[2 * plane + pix] = C2 + T * bg[2];
    }
    __syncthreads();
    if (tid == 0 && sm.red) atomicMax(&tile_maxlast[tile], sm.red);
    MB_TRACE_END(0, sm.red);
}

// 1 / x by the special-function unit (MUFU.RCP, <= 1 ulp) instead of the IEEE division sequence: used for 1 / (1 - alpha) in
// the backward, where upstream itself reconstructs T by repeated division and the result feeds sums of thousands of terms
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// lane -> which of the 9 reduced values it owns after the reduce-scatter (or -1)
__device__ __forceinline__ int reduce_slot(int lane) {
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    const int k2 = 2 * b2 + b1;        // index into the 3-array
    const int k3 = 3 * b3 + k2;        // index into the 5-array
    const int k4 = 5 * b4 + k3;        // index into the 9-array
    const bool valid = (lane & 1) == 0 && k2 < 3 && k3 < 5 && k4 < 9 && !(b2 && b1) && !(b3 && k2 >= 2) && !(b4 && k3 >= 4);
    return valid ? k4 : -1;
}

// Sum each of v[0..8] over the 32 lanes: 5+3+2+1+1 = 12 shuffles.  Afterwards the lane with reduce_slot(lane) == k
// holds the total of v[k] in the return value.
__device__ __forceinline__ float reduce_scatter9(const float (&v)[9], int lane) {
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
    float a[5], b[3], c[2];
#pragma unroll
    for (int k = 0; k < 5; ++k) {   // keep v[0..4] (low half) or v[5..8] (high half)
        const float lo = v[k], hi = k < 4 ? v[5 + k] : 0.f;
        const float recv = __shfl_xor_sync(0xffffffffu, h4 ? lo : hi, 16);
        a[k] = (h4 ? hi : lo) + recv;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {   // keep a[0..2] or a[3..4]
        const float lo = a[k], hi = k < 2 ? a[3 + k] : 0.f;
        const float recv = __shfl_xor_sync(0xffffffffu, h3 ? lo : hi, 8);
        b[k] = (h3 ? hi : lo) + recv;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {   // keep b[0..1] or b[2]
        const float lo = b[k], hi = k < 1 ? b[2 + k] : 0.f;
        const float recv = __shfl_xor_sync(0xffffffffu, h2 ? lo : hi, 4);
        c[k] = (h2 ? hi : lo) + recv;
    }
    const float recv = __shfl_xor_sync(0xffffffffu, h1 ? c[0] : c[1], 2);
    float d = (h1 ? c[1] : c[0]) + recv;
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// ---- packed fp32x2 arithmetic (sm_100a FFMA2 / FMUL2 / FADD2): one issue slot for the two pixels of a lane ----
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// exp(x) of both halves for x <= 0 (|x| < 100): exp(x) = 2^t 2^r with t = fl(x log2e) and r = the rounding residual of that
// product plus x (log2e - fl(log2e)); |r| < 1e-6, so 2^r = 1 + r ln2.  MUFU.EX2 at the core like expf(), no range split needed.
__device__ __forceinline__ float2 exp_pair(float2 x) {
    const float L = 1.4426950216293334961f, Llo = 1.925963033500011079e-08f, ln2 = 0.693147182464599609375f;
    const float2 t = __fmul2_rn(x, splat(L));
    float2 r = __ffma2_rn(x, splat(L), neg2(t));
    r = __ffma2_rn(x, splat(Llo), r);
    const float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
    return __ffma2_rn(e, __fmul2_rn(r, splat(ln2)), e);
}

#ifndef MB_BWD_SMEM_REDUCE
#define MB_BWD_SMEM_REDUCE 1
#endif
constexpr int kRedStride = 36;     // floats per row of a warp's reduction scratch: 16-B aligned rows, conflict-free 128-bit reads

struct SegmentSmem {
    Record rec[kSeg];          // the segment's records, list order
    uint32_t ids[kSeg + 4];    // the segment's slice of the id list (16-B aligned source => up to 3 ids of slack in front)
    uint64_t bar;
#if MB_BWD_SMEM_REDUCE
    alignas(16) float red[4][8 * kRedStride];   // per warp: 8 of the 9 per-Gaussian partials x 32 lanes (transposed reduction)
#endif
};

#if MB_BWD_SMEM_REDUCE
// Sum each of v[0..8] over the 32 lanes through shared memory.  v[0..7]: every lane stores its eight values as one row element each
// (8 STS, conflict free), then lane (k = lane / 4, part = lane % 4) loads 8 consecutive lanes' values of row k with two 128-bit
// loads (conflict free per quarter warp with the 36-float row stride), adds them (7 FADD) and the four parts meet with two
// shuffles: lane 4k holds the total of v[k].  v[8] takes the 5-step butterfly (all lanes hold its total).  ~36 issue slots against
// ~52 for the register-only reduce-scatter, and the selects of that version (ALU pipe) become LSU work.
// Returns the value the lane issues its RED with: slot k = lane / 4 on lanes 4k, slot 8 on lane 1 (see reduce_slot_smem).
__device__ __forceinline__ float reduce9_smem(const float (&v)[9], float *red, int lane) {
    __syncwarp();      // the previous survivor's loads are done
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k * kRedStride + lane] = v[k];
    float w = v[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    __syncwarp();
    const float4 *row = reinterpret_cast<const float4 *>(red + (lane >> 2) * kRedStride + (lane & 3) * 8);
    const float4 a = row[0], b = row[1];
    float t = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    return lane == 1 ? w : t;
}
__device__ __forceinline__ int reduce_slot_smem(int lane) { return lane == 1 ? 8 : ((lane & 3) == 0 ? lane >> 2 : -1); }
#endif

#ifndef MB_BWD2_CTAS_PER_SM
#define MB_BWD2_CTAS_PER_SM 8
#endif
// Backward, two pixels per lane.  Work item = one kSeg-entry segment of one 16x16 tile's list (as above); the CTA's four
// warps own the tile's four 8x8 pixel blocks, lane l the pixels (l & 7, l >> 3) and (l & 7, (l >> 3) + 4) of its block.  The
// pair shares dx, and everything per pixel that is plain fp32 arithmetic is issued once for both as a packed FFMA2 / FMUL2 /
// FADD2; per-survivor overhead (list walk, record fetch, the nine-value warp reduction, the RED) is paid once per 64 pixels.
// The recurrence is upstream's, restated without selects: an inactive (pixel, Gaussian) pair takes alpha = G = 0, for which
// every update below is the identity and every partial an exact zero.  State per pixel: T and B = the "colour behind" term
// the NEXT contributor sees, B' = alpha c + (1 - alpha) B (upstream's accum_rec, evaluated one contributor earlier).
__global__ void __launch_bounds__(128, MB_BWD2_CTAS_PER_SM) blend_backward2_kernel(
    const Record *__restrict__ recs, const uint32_t *__restrict__ list, const uint2 *__restrict__ ranges,
    const uint2 *__restrict__ items, const uint32_t *__restrict__ n_items, const uint32_t *__restrict__ tile_maxlast, int W, int H,
    int gx, const float *__restrict__ bg, const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
    const float4 *__restrict__ ckpt, const float *__restrict__ dL_dout, int64_t sc, int64_t sy, int64_t sx,
    float *__restrict__ acc) {
    __shared__ __align__(128) SegmentSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int item = blockIdx.x;
    if ((uint32_t)item >= *n_items) return;
    MB_TRACE_BEGIN();
    const uint2 it = items[item];
    const int tile = (int)it.x;
    const uint32_t s0 = it.y * (uint32_t)kSeg;
    const uint32_t s1 = min(s0 + (uint32_t)kSeg, tile_maxlast[tile]);   // list positions [s0, s1) of this tile
    const int len = (int)(s1 - s0);
    const uint2 range = ranges[tile];
    const int lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 8 + (lane >> 3);   // pixel 0 inside the tile; pixel 1 is 4 rows below
    const int px = (tile % gx) * kTile + lx, py = (tile / gx) * kTile + ly;
    const float fx = (float)px;
    const float2 nfy = make_float2(-(float)py, -(float)(py + 4));
    const bool in0 = px < W && py < H, in1 = px < W && py + 4 < H;
    const size_t pix0 = (size_t)py * W + px, pix1 = pix0 + 4 * (size_t)W;
    const float2 T_final = make_float2(in0 ? final_T[pix0] : 0.f, in1 ? final_T[pix1] : 0.f);
    const uint32_t last0 = in0 ? n_contrib[pix0] : 0u, last1 = in1 ? n_contrib[pix1] : 0u;
    float2 dp0 = splat(0.f), dp1 = dp0, dp2 = dp0;
    if (in0) {
        const float *gp = dL_dout + (int64_t)py * sy + (int64_t)px * sx;
        dp0.x = gp[0]; dp1.x = gp[sc]; dp2.x = gp[2 * sc];
    }
    if (in1) {
        const float *gp = dL_dout + (int64_t)(py + 4) * sy + (int64_t)px * sx;
        dp0.y = gp[0]; dp1.y = gp[sc]; dp2.y = gp[2 * sc];
    }
    const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
    const float2 bg_dot = make_float2(b0 * dp0.x + b1 * dp1.x + b2 * dp2.x, b0 * dp0.y + b1 * dp1.y + b2 * dp2.y);
    const float2 nTfbg = make_float2(-T_final.x * bg_dot.x, -T_final.y * bg_dot.y);   // (-T_final / (1 - alpha)) bg_dot = rinv * nTfbg
#if MB_BWD_SMEM_REDUCE
    const int slot = reduce_slot_smem(lane);
    float *red = sm.red[warp];
#else
    const int slot = reduce_slot(lane);
#endif
    const float bcx = (float)(px - (lane & 7)) + 3.5f, bcy = (float)(py - (lane >> 3)) + 3.5f;   // centre of the warp's 8x8 block

    float2 T = T_final, B0 = splat(0.f), B1 = B0, B2 = B0;
    {   // contributions behind this segment: resume from the forward's state in front of position s1 (per pixel)
        const uint32_t nseg_list = (range.y - range.x + (uint32_t)kSeg - 1u) / (uint32_t)kSeg;
        const size_t ck_base = BinningState::ckpt_slot((uint32_t)tile, range.x, it.y) * kTilePixels;
        const size_t fin_base = BinningState::ckpt_slot((uint32_t)tile, range.x, nseg_list - 1u) * kTilePixels;
        // the forward indexes a tile's pixels by (8x4 block, lane)
        const int idx0 = ((lx >> 3) + 2 * (ly >> 2)) * 32 + (lx & 7) + 8 * (ly & 3), idx1 = idx0 + 64;
        if (last0 > s1) {
            const float4 ck = ckpt[ck_base + idx0], fin = ckpt[fin_base + idx0];
            const float inv = 1.0f / ck.x;
            T.x = ck.x; B0.x = (fin.y - ck.y) * inv; B1.x = (fin.z - ck.z) * inv; B2.x = (fin.w - ck.w) * inv;
        }
        if (last1 > s1) {
            const float4 ck = ckpt[ck_base + idx1], fin = ckpt[fin_base + idx1];
            const float inv = 1.0f / ck.x;
            T.y = ck.x; B0.y = (fin.y - ck.y) * inv; B1.y = (fin.z - ck.z) * inv; B2.y = (fin.w - ck.w) * inv;
        }
    }

    // The whole segment is staged at once: ids by one bulk (TMA) copy, then every thread gathers the records of up to
    // kSeg / 128 entries (three 128-bit cp.async each, L2 hits).  One barrier; after it the four warps walk the segment
    // independently (their 8x8 blocks keep different survivors, a per-batch barrier made them wait for the slowest).
    const uint32_t first = range.x + s0, o = first & 3u;
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
        const uint32_t bytes = ((o + (uint32_t)len) * 4u + 15u) & ~15u;
        mbar_expect_tx(&sm.bar, bytes);
        bulk_g2s(&sm.ids[0], list + (first - o), bytes, &sm.bar);
    }
    __syncthreads();
    mbar_wait(&sm.bar, 0u);
    const uint32_t *ids = &sm.ids[o];
    for (int e = tid; e < len; e += 128) {
        const char *src = reinterpret_cast<const char *>(recs + ids[e]);
        char *dst = reinterpret_cast<char *>(&sm.rec[e]);
        cp_async16(dst, src);
        cp_async16(dst + 16, src + 16);
        cp_async16(dst + 32, src + 32);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const float4 *r = reinterpret_cast<const float4 *>(&sm.rec[0]);

    // nothing to do for this warp's pixels behind their deepest last contributor (positions relative to s0)
    const uint32_t wlast_abs = __reduce_max_sync(0xffffffffu, max(last0, last1));
    const int wlast = wlast_abs > s0 ? (int)min(wlast_abs - s0, (uint32_t)len) : 0;
    const uint32_t lrel0 = last0 > s0 ? last0 - s0 : 0u, lrel1 = last1 > s0 ? last1 - s0 : 0u;
    {
        const int jend = wlast;   // entries at or behind the warp's deepest last contributor cannot matter
        for (int j0 = ((jend - 1) >> 5) << 5; j0 >= 0; j0 -= 32) {
            bool hit = false;
            if (j0 + lane < jend) {
                const float4 ra = r[3 * (j0 + lane)], rc = r[3 * (j0 + lane) + 2];
                hit = fabsf(ra.x - bcx) <= rc.z + 3.5f && fabsf(ra.y - bcy) <= rc.w + 3.5f;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {   // survivors back to front
                const int k = 31 - __clz(m);
                m &= ~(1u << k);
                const int j = j0 + k;
                const uint32_t pos = (uint32_t)j;
                const float4 ra = r[3 * j], rb = r[3 * j + 1];
                const float2 rc = *reinterpret_cast<const float2 *>(&r[3 * j + 2]);
                // power = -0.5 (conic.x dx^2 + conic.z dy^2) - conic.y dx dy ; dx is common to the pair
                const float dx = ra.x - fx;
                const float2 dy = __fadd2_rn(splat(ra.y), nfy);
                const float hxx = (ra.z * dx) * dx, nbdx = -ra.w * dx;
                const float2 sq = __ffma2_rn(__fmul2_rn(splat(rb.x), dy), dy, splat(hxx));
                const float2 power = __ffma2_rn(splat(nbdx), dy, __fmul2_rn(sq, splat(-0.5f)));
                const bool cand0 = (pos < lrel0) && (power.x <= 0.0f) && !(power.x < rc.y);
                const bool cand1 = (pos < lrel1) && (power.y <= 0.0f) && !(power.y < rc.y);
                if (!__any_sync(0xffffffffu, cand0 || cand1)) continue;
                float2 G = exp_pair(power);
                float2 alpha = __fmul2_rn(splat(rb.y), G);
                alpha.x = fminf(kAlphaMax, alpha.x);
                alpha.y = fminf(kAlphaMax, alpha.y);
                const bool act0 = cand0 && (alpha.x >= kAlphaMin), act1 = cand1 && (alpha.y >= kAlphaMin);
                // inactive pairs: alpha = G = 0 (G may be inf / NaN there: power > 0 is not excluded before the exponential)
                G.x = act0 ? G.x : 0.f; alpha.x = act0 ? alpha.x : 0.f;
                G.y = act1 ? G.y : 0.f; alpha.y = act1 ? alpha.y : 0.f;
                const float2 om = __ffma2_rn(alpha, splat(-1.0f), splat(1.0f));   // 1 - alpha
                const float2 rinv = make_float2(fast_rcp(om.x), fast_rcp(om.y));  // exactly 1 for an inactive pair
                T = __fmul2_rn(T, rinv);
                const float c0 = rb.z, c1 = rb.w, c2 = rc.x;
                // dL/dalpha = ((c - B) . dL/dpixel) T + (-T_final / (1 - alpha)) (bg . dL/dpixel)
                float2 dot = __fmul2_rn(__ffma2_rn(B0, splat(-1.0f), splat(c0)), dp0);
                dot = __ffma2_rn(__ffma2_rn(B1, splat(-1.0f), splat(c1)), dp1, dot);
                dot = __ffma2_rn(__ffma2_rn(B2, splat(-1.0f), splat(c2)), dp2, dot);
                const float2 dL_dalpha = __ffma2_rn(dot, T, __fmul2_rn(rinv, nTfbg));
                const float2 dch = __fmul2_rn(alpha, T);   // dL/dcolour weight: alpha * (T in front of this Gaussian)
                B0 = __ffma2_rn(alpha, splat(c0), __fmul2_rn(om, B0));
                B1 = __ffma2_rn(alpha, splat(c1), __fmul2_rn(om, B1));
                B2 = __ffma2_rn(alpha, splat(c2), __fmul2_rn(om, B2));
                // moments of q = G dL/dalpha (the 0.99 clamp is not masked: upstream behaviour)
                const float2 q = __fmul2_rn(G, dL_dalpha);
                const float2 qx = __fmul2_rn(q, splat(dx)), qy = __fmul2_rn(q, dy);
                const float2 qxx = __fmul2_rn(qx, splat(dx)), qxy = __fmul2_rn(qx, dy), qyy = __fmul2_rn(qy, dy);
                const float2 g0 = __fmul2_rn(dch, dp0), g1 = __fmul2_rn(dch, dp1), g2 = __fmul2_rn(dch, dp2);
                float v[9];
                v[0] = qx.x + qx.y; v[1] = qy.x + qy.y;
                v[2] = qxx.x + qxx.y; v[3] = qxy.x + qxy.y; v[4] = qyy.x + qyy.y;
                v[5] = q.x + q.y;
                v[6] = g0.x + g0.y; v[7] = g1.x + g1.y; v[8] = g2.x + g2.y;
#if MB_BWD_SMEM_REDUCE
                const float total = reduce9_smem(v, red, lane);
#else
                const float total = reduce_scatter9(v, lane);
#endif
                if (slot >= 0) red_add(acc + (size_t)ids[j] * kAccStride + slot, total);
            }
        }
    }
    MB_TRACE_END(1, len);
}

struct PreBwdArgs {
    int P, W, H, deg, M;
    float tanx, tany, focx, focy, scale_mod;
    const float *means3D, *cov3D, *scales, *rots, *shs, *view, *proj, *campos, *tanfov_dev;
    const int32_t *radii;
    const uint32_t *clamped;
    const Record *rec;     // the forward's blend records (opacity as the forward used it)
    const float *acc;
    float *dL_dmeans2D, *dL_dcolors, *dL_dopacity, *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscales, *dL_drots;
};

template <bool kSH, bool kScaleRot>
__global__ void __launch_bounds__(256) preprocess_backward_kernel(PreBwdArgs a) {
    __shared__ float cam[36];
    const int tid = threadIdx.x;
    if (tid < 16) cam[tid] = a.view[tid];
    else if (tid < 32) cam[tid] = a.proj[tid - 16];
    else if (tid < 35) cam[tid] = a.campos[tid - 32];
    __syncthreads();
    if (a.tanfov_dev) {
        a.tanx = a.tanfov_dev[0]; a.tany = a.tanfov_dev[1];
        a.focx = a.W / (2.0f * a.tanx); a.focy = a.H / (2.0f * a.tany);
    }
    const float *v = cam, *p = cam + 16;
    const int i = blockIdx.x * 256 + tid;
    if (i >= a.P) return;
    float gmean[3] = {0.f, 0.f, 0.f}, gcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g2x = 0.f, g2y = 0.f, gop = 0.f, gc[3] = {0.f, 0.f, 0.f};
    const bool vis = a.radii[i] > 0;
    const float mx = a.means3D[3 * i], my = a.means3D[3 * i + 1], mz = a.means3D[3 * i + 2];
    if (vis) {
        const float *ac = a.acc + (size_t)i * kAccStride;
        gop = ac[5]; gc[0] = ac[6]; gc[1] = ac[7]; gc[2] = ac[8];
        const float *c6 = a.cov3D + 6 * (size_t)i;
        float g2[2];
        project_backward(v, p, a.tanx, a.tany, a.focx, a.focy, a.W, a.H, mx, my, mz, c6, a.rec[i].b.y, ac, g2, gmean, gcov);
        g2x = g2[0]; g2y = g2[1];

        if (kSH) {
            float dx = mx - cam[32], dy = my - cam[33], dz = mz - cam[34];
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            dx *= inv; dy *= inv; dz *= inv;
            float basis[16], bxg[16], byg[16], bzg[16];
            sh_basis(a.deg, dx, dy, dz, basis);
            sh_basis_grad(a.deg, dx, dy, dz, bxg, byg, bzg);
            const int nbas = (a.deg + 1) * (a.deg + 1);
            const uint32_t mask = a.clamped[i];
            const float *sh = a.shs + (size_t)i * a.M * 3;
            float *dsh = a.dL_dsh + (size_t)i * a.M * 3;
            float gd[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < a.M; ++k)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float go = ((mask >> ch) & 1u) ? 0.f : gc[ch];
                    if (k < nbas) {
                        dsh[3 * k + ch] = basis[k] * go;
                        const float s = sh[3 * k + ch] * go;
                        gd[0] += bxg[k] * s; gd[1] += byg[k] * s; gd[2] += bzg[k] * s;
                    } else dsh[3 * k + ch] = 0.f;
                }
            const float dot = dx * gd[0] + dy * gd[1] + dz * gd[2];
            gmean[0] += (gd[0] - dx * dot) * inv;
            gmean[1] += (gd[1] - dy * dot) * inv;
            gmean[2] += (gd[2] - dz * dot) * inv;
        }
        if (kScaleRot) {
            const float s[3] = {a.scale_mod * a.scales[3 * i], a.scale_mod * a.scales[3 * i + 1], a.scale_mod * a.scales[3 * i + 2]};
            const float q0 = a.rots[4 * i], q1 = a.rots[4 * i + 1], q2 = a.rots[4 * i + 2], q3 = a.rots[4 * i + 3];
            float R[9], L[9], dLm[9], dR[9], dq[4];
            quat_to_rot(q0, q1, q2, q3, R);
            const float Gm[9] = {gcov[0], 0.5f * gcov[1], 0.5f * gcov[2], 0.5f * gcov[1], gcov[3], 0.5f * gcov[4],
                                 0.5f * gcov[2], 0.5f * gcov[4], gcov[5]};
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) L[3 * r + k] = R[3 * r + k] * s[k];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    dLm[3 * r + k] = 2.f * (Gm[3 * r] * L[k] + Gm[3 * r + 1] * L[3 + k] + Gm[3 * r + 2] * L[6 + k]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                a.dL_dscales[3 * i + k] = (dLm[k] * R[k] + dLm[3 + k] * R[3 + k] + dLm[6 + k] * R[6 + k]) * a.scale_mod;
#pragma unroll
                for (int r = 0; r < 3; ++r) dR[3 * r + k] = dLm[3 * r + k] * s[k];
            }
            quat_to_rot_bwd(q0, q1, q2, q3, dR, dq);
#pragma unroll
            for (int k = 0; k < 4; ++k) a.dL_drots[4 * i + k] = dq[k];
        }
    } else {
        if (kSH) {
            float *dsh = a.dL_dsh + (size_t)i * a.M * 3;
            for (int k = 0; k < a.M * 3; ++k) dsh[k] = 0.f;
        }
        if (kScaleRot) {
#pragma unroll
            for (int k = 0; k < 3; ++k) a.dL_dscales[3 * i + k] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) a.dL_drots[4 * i + k] = 0.f;
        }
    }
    a.dL_dmeans2D[3 * i] = g2x; a.dL_dmeans2D[3 * i + 1] = g2y; a.dL_dmeans2D[3 * i + 2] = 0.f;
    a.dL_dopacity[i] = gop;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a.dL_dcolors[3 * i + k] = kSH ? 0.f : gc[k];
        a.dL_dmeans3D[3 * i + k] = gmean[k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * (size_t)i + k] = gcov[k];
}

// work-item shape of the tile kernels: 8 warps = one 16x16 tile per CTA, 4 warps = one 16x8 half tile per CTA
static int blend_warps() {
    static int cached = 0;
    if (!cached) {
        const char *e = getenv("MB_BLEND_WARPS");
        const int v = e ? atoi(e) : 4;
        cached = (v == 8 || v == 2) ? v : 4;
    }
    return cached;
}

}  // namespace mb

using namespace mb;

extern "C" int mb_raster_state_layout(int32_t num_points, int64_t capacity, int32_t w, int32_t h, int64_t *out, int32_t n_out) {
    MB_REQUIRE(out != nullptr && n_out >= 8, "mb_raster_state_layout: need room for 8 offsets");
    GeomState g = GeomState::carve(nullptr, num_points);
    BinningState b = BinningState::carve(nullptr, capacity, ((w + kTile - 1) / kTile) * ((h + kTile - 1) / kTile));
    ImageState im = ImageState::carve(nullptr, w, h);
    out[0] = (int64_t)((char *)im.final_T - (char *)nullptr);
    out[1] = (int64_t)((char *)im.n_contrib - (char *)nullptr);
    out[2] = (int64_t)((char *)im.ranges - (char *)nullptr);
    out[3] = (int64_t)((char *)im.tile_maxlast - (char *)nullptr);
    out[4] = (int64_t)((char *)b.gid_b - (char *)nullptr);
    out[5] = (int64_t)((char *)b.tile_b - (char *)nullptr);
    out[6] = (int64_t)((char *)g.rec - (char *)nullptr);
    out[7] = (int64_t)((char *)g.counters - (char *)nullptr);
    return MB_OK;
}

extern "C" int mb_raster_forward_render(const mb_raster_inputs *in, void *geom, void *binning, size_t binning_bytes,
                                        int64_t capacity, void *image_buf, size_t image_bytes, float *out_color,
                                        mb_stream_t stream) {
    int rc = validate_raster_inputs(in, "mb_raster_forward_render", false, false);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterDims d = raster_dims(in);
    MB_REQUIRE(geom && binning && image_buf && out_color, "mb_raster_forward_render: null buffer");
    MB_REQUIRE(capacity >= 0 && capacity < (int64_t)0xffffffff, "mb_raster_forward_render: capacity out of range");
    GeomState g = GeomState::carve(geom, d.P);
    BinningState b = BinningState::carve(binning, capacity, d.tiles);
    ImageState im = ImageState::carve(image_buf, d.W, d.H);
    if (binning_bytes < b.bytes || image_bytes < im.bytes) {
        set_error("mb_raster_forward_render: binning %zu/%zu or image %zu/%zu bytes too small", binning_bytes, b.bytes,
                  image_bytes, im.bytes);
        return MB_ERR_WORKSPACE;
    }
    MB_CUDA(cudaMemsetAsync(im.ranges, 0, (size_t)((char *)im.order_fwd - (char *)im.ranges), s));   // ranges + tile_maxlast + order workspace
    const bool dbg = in->debug != 0;
    const uint32_t *order = nullptr;
    if (d.P > 0 && capacity > 0) {
        rc = build_instances(in, d, g, b, im, capacity, s);
        if (rc) return rc;
        rc = tile_order(nullptr, im.ranges, d.tiles, im.order_fwd, im.order_ws, s, dbg, true);   // also decodes the tile ranges
        if (rc) return rc;
        order = im.order_fwd;
    }
    {
        KernelTimer kt("blend_forward", s);
        if (blend_warps() == 4)
            blend_forward_kernel<4><<<d.tiles * 2, 128, 0, s>>>(g.rec, b.gid_b, im.ranges, order, d.W, d.H, d.gx, in->background,
                                                                out_color, im.final_T, im.n_contrib, im.tile_maxlast, b.ckpt);
        else if (blend_warps() == 2)
            blend_forward_kernel<2><<<d.tiles * 4, 64, 0, s>>>(g.rec, b.gid_b, im.ranges, order, d.W, d.H, d.gx, in->background,
                                                               out_color, im.final_T, im.n_contrib, im.tile_maxlast, b.ckpt);
        else
            blend_forward_kernel<8><<<d.tiles, 256, 0, s>>>(g.rec, b.gid_b, im.ranges, order, d.W, d.H, d.gx, in->background,
                                                            out_color, im.final_T, im.n_contrib, im.tile_maxlast, b.ckpt);
    }
    return check_launch("blend_forward", dbg, s);
}

extern "C" size_t mb_raster_backward_scratch_bytes(int32_t num_points) {
    return align_up((size_t)(num_points > 0 ? num_points : 1) * kAccStride * sizeof(float));
}

static int raster_backward_impl(const mb_raster_inputs *in, const int32_t *radii, const void *geom, const void *binning,
                                int64_t capacity, const void *image_buf, const float *dL_dout, int64_t stride_c,
                                int64_t stride_y, int64_t stride_x, void *grad_scratch, size_t scratch_bytes,
                                float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D,
                                float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations, bool blend_only,
                                mb_stream_t stream) {
    // the backward never reads the opacities (they are part of the saved blend records), like upstream's, whose
    // rasterize_gaussians_backward does not take them
    int rc = validate_raster_inputs(in, "mb_raster_backward", false, !blend_only);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterDims d = raster_dims(in);
    if (d.P == 0) return MB_OK;
    MB_REQUIRE(radii && geom && binning && image_buf && dL_dout && grad_scratch, "mb_raster_backward: null buffer");
    if (!blend_only) {
        MB_REQUIRE(dL_dmeans2D && dL_dcolors && dL_dopacity && dL_dmeans3D && dL_dcov3D, "mb_raster_backward: null output");
        MB_REQUIRE(!in->shs || dL_dsh, "mb_raster_backward: dL_dsh required when shs is given");
        MB_REQUIRE(in->cov3D_precomp || (dL_dscales && dL_drotations), "mb_raster_backward: dL_dscales / dL_drotations required");
    }
    if (scratch_bytes < mb_raster_backward_scratch_bytes(d.P)) {
        set_error("mb_raster_backward: scratch too small");
        return MB_ERR_WORKSPACE;
    }
    const bool dbg = in->debug != 0;
    GeomState g = GeomState::carve(const_cast<void *>(geom), d.P);
    BinningState b = BinningState::carve(const_cast<void *>(binning), capacity, d.tiles);
    ImageState im = ImageState::carve(const_cast<void *>(image_buf), d.W, d.H);
    float *acc = reinterpret_cast<float *>(grad_scratch);
    MB_CUDA(cudaMemsetAsync(acc, 0, (size_t)d.P * kAccStride * sizeof(float), s));
    if (capacity > 0) {
        uint32_t *n_items = g.counters + kCntBwdItems;
        rc = segment_items(im.tile_maxlast, d.tiles, b.bwd_items, n_items, im.order_ws + kOrderWs, s, dbg);
        if (rc) return rc;
        {
            KernelTimer kt("blend_backward", s);
            // upper bound of the item count (the real one is on the device; surplus CTAs exit at once)
            const int64_t bound = (int64_t)d.tiles + capacity / kSeg + 1;
            const int64_t max_items = bound < b.max_items ? bound : b.max_items;
            blend_backward2_kernel<<<(unsigned)max_items, 128, 0, s>>>(g.rec, b.gid_b, im.ranges, b.bwd_items, n_items, im.tile_maxlast,
                                                                      d.W, d.H, d.gx, in->background, im.final_T, im.n_contrib, b.ckpt,
                                                                      dL_dout, stride_c, stride_y, stride_x, acc);
        }
        rc = check_launch("blend_backward", dbg, s);
        if (rc) return rc;
    }
    if (blend_only) return MB_OK;      // the accumulator rows are consumed by mb_pose_backward_from_raster
    PreBwdArgs a;
    a.P = d.P; a.W = d.W; a.H = d.H; a.deg = in->sh_degree; a.M = in->sh_coeffs;
    a.tanx = in->tanfovx; a.tany = in->tanfovy; a.focx = d.focx; a.focy = d.focy; a.scale_mod = in->scale_modifier;
    a.means3D = in->means3D; a.cov3D = in->cov3D_precomp ? in->cov3D_precomp : g.cov3D; a.scales = in->scales;
    a.rots = in->rotations; a.shs = in->shs; a.view = in->viewmatrix; a.proj = in->projmatrix; a.campos = in->campos;
    a.tanfov_dev = in->tanfov_dev;
    a.radii = radii; a.clamped = g.clamped; a.rec = g.rec; a.acc = acc;
    a.dL_dmeans2D = dL_dmeans2D; a.dL_dcolors = dL_dcolors; a.dL_dopacity = dL_dopacity; a.dL_dmeans3D = dL_dmeans3D;
    a.dL_dcov3D = dL_dcov3D; a.dL_dsh = dL_dsh; a.dL_dscales = dL_dscales; a.dL_drots = dL_drotations;
    const int grid = (d.P + 255) / 256;
    const bool sh = in->shs != nullptr, sr = in->cov3D_precomp == nullptr;
    KernelTimer kt("preprocess_backward", s);
    if (sh && sr) preprocess_backward_kernel<true, true><<<grid, 256, 0, s>>>(a);
    else if (sh) preprocess_backward_kernel<true, false><<<grid, 256, 0, s>>>(a);
    else if (sr) preprocess_backward_kernel<false, true><<<grid, 256, 0, s>>>(a);
    else preprocess_backward_kernel<false, false><<<grid, 256, 0, s>>>(a);
    return check_launch("preprocess_backward", dbg, s);
}

extern "C" int mb_raster_backward(const mb_raster_inputs *in, const int32_t *radii, const void *geom, const void *binning,
                                  int64_t capacity, const void *image_buf, const float *dL_dout, int64_t stride_c,
                                  int64_t stride_y, int64_t stride_x, void *grad_scratch, size_t scratch_bytes,
                                  float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D,
                                  float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                                  mb_stream_t stream) {
    return raster_backward_impl(in, radii, geom, binning, capacity, image_buf, dL_dout, stride_c, stride_y, stride_x, grad_scratch,
                                scratch_bytes, dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales,
                                dL_drotations, false, stream);
}

extern "C" int mb_raster_backward_blend(const mb_raster_inputs *in, const int32_t *radii, const void *geom, const void *binning,
                                        int64_t capacity, const void *image_buf, const float *dL_dout, int64_t stride_c,
                                        int64_t stride_y, int64_t stride_x, void *grad_scratch, size_t scratch_bytes,
                                        mb_stream_t stream) {
    return raster_backward_impl(in, radii, geom, binning, capacity, image_buf, dL_dout, stride_c, stride_y, stride_x, grad_scratch,
                                scratch_bytes, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, true, stream);
}

#ifdef MB_TRACE_CTA
// (variant builds only) copies the per-CTA timeline of the last launch of tile kernel `kernel` (0 forward) to the host
extern "C" int mb_debug_cta_trace(int kernel, unsigned long long *times_host /*[n][2]*/, uint32_t *work_host /*[n]*/, int n) {
    if (cudaDeviceSynchronize() != cudaSuccess) return MB_ERR_CUDA;
    if (n > kTraceMax) n = kTraceMax;
    MB_CUDA(cudaMemcpyFromSymbol(times_host, g_trace, sizeof(ulonglong2) * (size_t)n, sizeof(ulonglong2) * (size_t)kTraceMax * kernel));
    MB_CUDA(cudaMemcpyFromSymbol(work_host, g_trace_work, sizeof(uint32_t) * (size_t)n, sizeof(uint32_t) * (size_t)kTraceMax * kernel));
    return MB_OK;
}
#endif
