// Fused pre-raster step: linear-blend skinning of means and covariances, covariance build from (log-scale, quaternion),
// SH -> RGB in the canonical-space view direction, sigmoid / exp activations -- forward and backward in one kernel each.
//
// Replaces ~100 small PyTorch kernels per frame of the reference (src/modules/hand_dynamic.py:86-137,
// src/models/gaussian.py:48-93, src/utils/gaussian_utils.py:248-314,431-449, src/utils/sh_utils.py:57-120): the per-Gaussian
// 4x4 transform is never written to HBM and its inverse (torch.linalg.inv on N 4x4 matrices in the reference) is the
// closed-form affine inverse.
//
// Persistent CTAs walk tiles of 128 Gaussians (one thread per Gaussian).  Every per-Gaussian array of a tile is one
// dense run in HBM (f_rest 23 KB, skin weights 10.5 KB, xyz / scales / quaternions / ... 0.5-2 KB), so a tile is fetched by
// a handful of 1-D bulk TMA copies (cp.async.bulk + mbarrier) into the CTA's shared-memory stage, results are written over
// the thread's own input rows and leave the same way (cp.async.bulk shared -> global).
// One stage per CTA and four CTAs per SM: the loads of one CTA overlap the arithmetic of the others (a double-buffered
// stage per CTA with half as many CTAs was measured slower).
//
// HBM bytes per Gaussian (fp32, K=16, B bones): forward 236 + 4B read, 52 written; backward re-reads the parameters
// and the 52 B of upstream gradients and writes 236 B of parameter gradients.
#include <stdlib.h>

#include "common.cuh"
#include "project_bwd.cuh"
#include "project_fwd.cuh"
#include "sort_scan.cuh"

namespace mb {

#ifndef MB_POSE_THREADS
#define MB_POSE_THREADS 128
#endif
constexpr int kPoseThreads = MB_POSE_THREADS;   // Gaussians per tile = threads per CTA
constexpr int kMaxBones = 64;
constexpr int kMaxArrays = 12;
constexpr int kMaxViews = 8;      // views of one step whose pose backward runs as ONE pass over the parameters

struct ViewArgs {      // what differs between the views of a step (multi-view pose backward)
    const float *acc;            // [N,12] accumulator rows of the view's tile backward
    const int32_t *radii;        // [N]
    const float *bone_tf, *bones_posed;   // the view's pose (one of the two; rest_inv is shared)
    const float *campos, *view, *proj, *tanfov_dev;
    float tanx, tany;
    float *g_means2D;            // [N,3] out
};

struct PoseArgs {
    int N, n_skinned, B, deg, K, iso;
    const float *xyz, *log_scale, *quat, *opacity_logit, *f_dc, *f_rest, *skin, *bone_tf, *campos;
    // forward outputs
    float *posed_xyz, *cov6, *colors, *opacity, *tf_out;
    // backward inputs / outputs
    const float *g_posed_xyz, *g_cov6, *g_colors, *g_opacity;
    float *g_xyz, *g_log_scale, *g_quat, *g_opacity_logit, *g_f_dc, *g_f_rest, *g_skin;
    int accumulate;   // backward: add the gradients to the output buffers (bulk TMA reduce-add) instead of overwriting them
    // backward fused with the rasterizer's projection backward (mb_pose_backward_from_raster): the upstream gradients are
    // derived in the kernel from the blend backward's accumulator rows instead of being read from the four g_* arrays
    const float *acc;          // [N,12]: five moments of the blend backward (project_bwd.cuh), dL/dopacity, dL/dcolour rgb, 3 unused; nullptr = not fused
    const int32_t *radii;      // [N]
    const float *view, *proj;  // [16] each
    const float *tanfov_dev;   // optional [2] in device memory (replaces tanx / tany)
    float tanx, tany;
    int W, H;
    float *g_means2D;          // [N,3] out: (dL/dmean2D x, y, 0)
    // optional densification statistics of the visible Gaussians (gaussian.py:335-338, gaussian_utils.py:461-473), updated
    // with atomics because the views of a step run concurrently: accum += |dL/dmean2D.xy|, denom += 1, max_radii = max(., radius)
    float *stat_accum, *stat_denom, *stat_maxrad;
    // bone transforms built in the kernel prologue: T_b = bones_posed[b] * rest_inv[b] for b < n_posed, identity for the
    // other B - n_posed rows (hand_dynamic.py:93-102); nullptr = take bone_tf as given
    const float *bones_posed, *rest_inv;
    int n_posed;
    // forward fused with the rasterizer's projection (mb_pose_project_forward): record, radius, tile rectangle, depth key of
    // every Gaussian are written from registers, the posed arrays only if their pointers are given
    int n_views;               // > 0: multi-view backward (views[]); the single-view fields above are unused
    ViewArgs views[kMaxViews];
    Record *rec;               // nullptr = plain pose forward
    ushort4 *rect;
    uint32_t *tiles_touched, *depth_key, *ident, *counters;
    uint32_t *depth_hist;      // [4][256] digit counts of the depth keys (the depth sort's histogram, zeroed before the launch)
    int32_t *radii_out;
    int gx, gy;
};

constexpr int kAccRow = 12;          // floats per accumulator row (kAccStride of raster_blend.cu)
constexpr int kCamFloats = 40;       // view 16 | proj 16 | tanx, tany, focx, focy | pad (fused backward only)

// Shared-memory image of one tile of kPoseThreads Gaussians (offsets in floats, every array 16-B aligned).  Rows are
// dense: the wide rows have odd word counts in the shipped configurations (45 = 15 SH coefficients x 3, 21 bones), so
// the per-thread row reads are bank-conflict free without padding.
struct TileLayout {
    int fr, sk, xyz, ls, quat, opac, fdc, cov, gpx, gcov, gcol, gop, floats;
};

__host__ __device__ inline TileLayout tile_layout(int K, int B, int iso, bool backward) {
    TileLayout L;
    const int T = kPoseThreads;
    int o = 0;
    L.fr = o; o += T * (K - 1) * 3;
    L.sk = o; o += T * B;
    L.xyz = o; o += T * 3;
    L.ls = o; o += T * (iso ? 1 : 3);
    L.quat = o; o += T * 4;
    L.opac = o; o += T;
    L.fdc = o; o += T * 3;
    L.cov = o; o += backward ? 0 : T * 6;
    L.gpx = o; o += backward ? T * 3 : 0;
    L.gcov = o; o += backward ? T * 6 : 0;
    L.gcol = o; o += backward ? T * 3 : 0;
    L.gop = o; o += backward ? T : 0;
    L.floats = o;
    return L;
}

inline size_t pose_smem_bytes(int B, int K, int iso, bool backward) {
    // bones (13 floats each) + camera centre + 2 mbarriers + camera matrices, then the tile stage
    return sizeof(float) * ((size_t)kMaxBones * 13 + 8 + kCamFloats) + sizeof(float) * (size_t)tile_layout(K, B, iso, backward).floats;
}

struct Transfer {   // one dense array of a tile: global <-> shared
    const float *g;
    int s;          // offset in the stage (floats)
    uint32_t bytes;
};

__device__ __forceinline__ bool bulk_ok(const Transfer &t) {
    return t.bytes && ((reinterpret_cast<uintptr_t>(t.g) | t.bytes) & 15u) == 0;
}

// 1-D bulk asynchronous copy shared -> global (TMA engine); completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// the same with an fp32 add at the destination (TMA reduction: the adds happen in L2, nothing is read back)
__device__ __forceinline__ void bulk_s2g_add(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct PoseLocal {
    float A[9], t[3], s;     // blended transform (identity for static Gaussians)
    float qn[4], qnorm;      // normalised quaternion
    float R[9], S[3], L[9];  // L = R diag(S)
    float x[3];
    bool skinned;
};

__device__ __forceinline__ void mat3_mul(const float *a, const float *b, float *c) {   // c = a b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}

__device__ __forceinline__ float mat3_inverse(const float *a, float *inv) {
    const float c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const float det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const float id = 1.0f / det;
    inv[0] = c00 * id; inv[1] = (a[2] * a[7] - a[1] * a[8]) * id; inv[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    inv[3] = c01 * id; inv[4] = (a[0] * a[8] - a[2] * a[6]) * id; inv[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    inv[6] = c02 * id; inv[7] = (a[1] * a[6] - a[0] * a[7]) * id; inv[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return det;
}

// Bit b of the mask: this Gaussian's weight of bone b is not zero.  The weights do not depend on the view, so kernels that
// blend the same row against several views' bones build the mask once.
__device__ __forceinline__ uint64_t bone_mask(const float *w_row, int B) {
    uint64_t m = 0;
    for (int b = 0; b < B; ++b) m |= (uint64_t)(w_row[b] != 0.f) << b;
    return m;
}

// p.A | p.t | p.s = sum_b w_b T_b over the bones of the mask, in ascending b (the same terms in the same order as a loop over
// all bones that skips zero weights).  Every lane walks its OWN non-zero bones: a warp of unrelated Gaussians takes as many
// trips as its busiest lane has bones (~11 of 21 for MANO-style weights) instead of one per bone any lane uses (~20); lanes
// on different bones read different rows of the 13-float table, 13 being odd the rows fall in different banks.
__device__ __forceinline__ void blend_bones(const float *w_row, uint64_t mask, const float *bones_s, PoseLocal &p) {
#pragma unroll
    for (int k = 0; k < 9; ++k) p.A[k] = 0.f;
    p.t[0] = p.t[1] = p.t[2] = 0.f; p.s = 0.f;
    while (mask) {
        const int b = __ffsll((long long)mask) - 1;
        mask &= mask - 1;
        const float w = w_row[b];
        const float *T = bones_s + 13 * b;
        p.A[0] += w * T[0]; p.A[1] += w * T[1]; p.A[2] += w * T[2]; p.t[0] += w * T[3];
        p.A[3] += w * T[4]; p.A[4] += w * T[5]; p.A[5] += w * T[6]; p.t[1] += w * T[7];
        p.A[6] += w * T[8]; p.A[7] += w * T[9]; p.A[8] += w * T[10]; p.t[2] += w * T[11];
        p.s += w * T[12];
    }
}

// everything both directions need: blended transform, rotation, scales.  `st` = the tile's stage, `row` = thread's row.
__device__ __forceinline__ void pose_common(const PoseArgs &a, const TileLayout &L, const float *st, int i, int row,
                                            const float *bones_s, PoseLocal &p) {
    p.x[0] = st[L.xyz + 3 * row]; p.x[1] = st[L.xyz + 3 * row + 1]; p.x[2] = st[L.xyz + 3 * row + 2];
    p.skinned = i < a.n_skinned;
    if (p.skinned) {
        const float *w_row = st + L.sk + row * a.B;
        blend_bones(w_row, bone_mask(w_row, a.B), bones_s, p);
    } else {
        p.A[0] = p.A[4] = p.A[8] = 1.f;
        p.A[1] = p.A[2] = p.A[3] = p.A[5] = p.A[6] = p.A[7] = 0.f;
        p.t[0] = p.t[1] = p.t[2] = 0.f; p.s = 1.f;
    }
    const float4 q = *reinterpret_cast<const float4 *>(st + L.quat + 4 * row);
    p.qnorm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    p.qn[0] = q.x / p.qnorm; p.qn[1] = q.y / p.qnorm; p.qn[2] = q.z / p.qnorm; p.qn[3] = q.w / p.qnorm;
    quat_to_rot(p.qn[0], p.qn[1], p.qn[2], p.qn[3], p.R);
    if (a.iso) {
        p.S[0] = p.S[1] = p.S[2] = expf(st[L.ls + row]);
    } else {
        p.S[0] = expf(st[L.ls + 3 * row]); p.S[1] = expf(st[L.ls + 3 * row + 1]); p.S[2] = expf(st[L.ls + 3 * row + 2]);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) p.L[3 * r + k] = p.R[3 * r + k] * p.S[k];
}

// view direction in canonical space: d = x - inv(tf)[c;1], returns |d| and the unit vector; ci = inv(tf)[c;1]
__device__ __forceinline__ float view_dir(const PoseLocal &p, const float *cam, const float *Ainv, float *ci, float *dir) {
    if (p.skinned) {
        const float is = 1.0f / p.s;
        const float u0 = cam[0] - p.t[0] * is, u1 = cam[1] - p.t[1] * is, u2 = cam[2] - p.t[2] * is;
        ci[0] = Ainv[0] * u0 + Ainv[1] * u1 + Ainv[2] * u2;
        ci[1] = Ainv[3] * u0 + Ainv[4] * u1 + Ainv[5] * u2;
        ci[2] = Ainv[6] * u0 + Ainv[7] * u1 + Ainv[8] * u2;
    } else {
        ci[0] = cam[0]; ci[1] = cam[1]; ci[2] = cam[2];
    }
    const float d0 = p.x[0] - ci[0], d1 = p.x[1] - ci[1], d2 = p.x[2] - ci[2];
    const float n = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
    dir[0] = d0 / n; dir[1] = d1 / n; dir[2] = d2 / n;
    return n;
}

// Persistent tile pipeline shared by both directions.  Every array of a tile is one dense run in global memory, so a
// tile is brought in by a handful of 1-D bulk (TMA) copies into the stage that is not being computed on, and written
// back the same way; runs that are not 16-B sized / aligned (last tile, skinned/static boundary, odd tensor offsets) take
// a cooperative load / store path through the same shared-memory image.
template <bool kBackward, bool kFused = false>
struct TilePipe {
    const PoseArgs &a;
    TileLayout L;
    float *stage0;
    uint64_t *bar;

    __device__ __forceinline__ float *stage(int k) const { return stage0 + (k & 1) * L.floats; }

    __device__ __forceinline__ int inputs(int tile, Transfer *t) const {
        const int base = tile * kPoseThreads, cnt = min(kPoseThreads, a.N - base);
        const int nsk = max(0, min(cnt, a.n_skinned - base));
        const int rs = (a.K - 1) * 3, lsw = a.iso ? 1 : 3;
        int n = 0;
        t[n++] = {a.f_rest ? a.f_rest + (size_t)base * rs : nullptr, L.fr, (uint32_t)(rs > 0 ? cnt * rs * 4 : 0)};
        t[n++] = {a.skin ? a.skin + (size_t)base * a.B : nullptr, L.sk, (uint32_t)(nsk * a.B * 4)};
        t[n++] = {a.xyz + (size_t)base * 3, L.xyz, (uint32_t)(cnt * 12)};
        t[n++] = {a.log_scale + (size_t)base * lsw, L.ls, (uint32_t)(cnt * lsw * 4)};
        t[n++] = {a.quat + (size_t)base * 4, L.quat, (uint32_t)(cnt * 16)};
        t[n++] = {a.opacity_logit + base, L.opac, (uint32_t)(cnt * 4)};
        t[n++] = {a.f_dc + (size_t)base * 3, L.fdc, (uint32_t)(cnt * 12)};
        if (kBackward && kFused) {
            // fused with the rasterizer: the tile's accumulator rows (12 floats each) and radii land in the region of the four
            // upstream-gradient arrays (13 floats per Gaussian) and are transposed in place (run_tiles)
            t[n++] = {a.acc + (size_t)base * kAccRow, L.gpx, (uint32_t)(cnt * kAccRow * 4)};
            t[n++] = {reinterpret_cast<const float *>(a.radii) + base, L.gpx + kPoseThreads * kAccRow, (uint32_t)(cnt * 4)};
        } else if (kBackward && a.n_views == 0) {
            t[n++] = {a.g_posed_xyz + (size_t)base * 3, L.gpx, (uint32_t)(cnt * 12)};
            t[n++] = {a.g_cov6 + (size_t)base * 6, L.gcov, (uint32_t)(cnt * 24)};
            t[n++] = {a.g_colors + (size_t)base * 3, L.gcol, (uint32_t)(cnt * 12)};
            t[n++] = {a.g_opacity + base, L.gop, (uint32_t)(cnt * 4)};
        }
        return n;
    }
    __device__ __forceinline__ int outputs(int tile, Transfer *t) const {
        const int base = tile * kPoseThreads, cnt = min(kPoseThreads, a.N - base);
        const int nsk = max(0, min(cnt, a.n_skinned - base));
        const int rs = (a.K - 1) * 3, lsw = a.iso ? 1 : 3;
        int n = 0;
        if (!kBackward) {
            // (the projecting forward may leave the posed arrays out: they then never touch HBM)
            t[n++] = {a.posed_xyz ? a.posed_xyz + (size_t)base * 3 : nullptr, L.xyz, (uint32_t)(a.posed_xyz ? cnt * 12 : 0)};
            t[n++] = {a.cov6 ? a.cov6 + (size_t)base * 6 : nullptr, L.cov, (uint32_t)(a.cov6 ? cnt * 24 : 0)};
            t[n++] = {a.colors ? a.colors + (size_t)base * 3 : nullptr, L.fdc, (uint32_t)(a.colors ? cnt * 12 : 0)};
            t[n++] = {a.opacity ? a.opacity + base : nullptr, L.opac, (uint32_t)(a.opacity ? cnt * 4 : 0)};
        } else {
            t[n++] = {a.g_f_rest ? a.g_f_rest + (size_t)base * rs : nullptr, L.fr, (uint32_t)(a.g_f_rest && rs > 0 ? cnt * rs * 4 : 0)};
            t[n++] = {a.g_skin ? a.g_skin + (size_t)base * a.B : nullptr, L.sk, (uint32_t)(a.g_skin ? nsk * a.B * 4 : 0)};
            t[n++] = {a.g_xyz + (size_t)base * 3, L.gpx, (uint32_t)(cnt * 12)};
            t[n++] = {a.g_log_scale + (size_t)base * lsw, L.ls, (uint32_t)(cnt * lsw * 4)};
            t[n++] = {a.g_quat + (size_t)base * 4, L.quat, (uint32_t)(cnt * 16)};
            t[n++] = {a.g_opacity_logit + base, L.gop, (uint32_t)(cnt * 4)};
            t[n++] = {a.g_f_dc + (size_t)base * 3, L.gcol, (uint32_t)(cnt * 12)};
        }
        return n;
    }
    // elected thread: start the bulk copies of `tile` into stage k
    __device__ __forceinline__ void prefetch(int tile, int k) const {
        Transfer t[kMaxArrays];
        const int n = inputs(tile, t);
        uint32_t total = 0;
        for (int j = 0; j < n; ++j)
            if (bulk_ok(t[j])) total += t[j].bytes;
        if (total == 0) return;
        mbar_expect_tx(&bar[k & 1], total);
        for (int j = 0; j < n; ++j)
            if (bulk_ok(t[j])) bulk_g2s(stage(k) + t[j].s, t[j].g, t[j].bytes, &bar[k & 1]);
    }
    // all threads: the tile's inputs are complete in stage k after this returns (phase = per-stage mbarrier parity bits)
    __device__ __forceinline__ void acquire(int tile, int k, uint32_t &phase) const {
        Transfer t[kMaxArrays];
        const int n = inputs(tile, t);
        bool any_bulk = false, any_plain = false;
        for (int j = 0; j < n; ++j) {
            if (bulk_ok(t[j])) any_bulk = true;
            else if (t[j].bytes) {
                any_plain = true;
                float *dst = stage(k) + t[j].s;
                const int words = (int)(t[j].bytes >> 2);
                for (int e = threadIdx.x; e < words; e += kPoseThreads) dst[e] = t[j].g[e];
            }
        }
        if (any_bulk) {
            mbar_wait(&bar[k & 1], (phase >> (k & 1)) & 1u);
            phase ^= 1u << (k & 1);
        }
        if (any_plain) __syncthreads();
    }
    // all threads, after the results have been written into stage k
    __device__ __forceinline__ void release(int tile, int k) const {
        fence_proxy_async();
        __syncthreads();
        Transfer t[kMaxArrays];
#ifdef MB_POSE_NOSTORE      // experiment switch: loads only
        const int n = 0;
#else
        const int n = outputs(tile, t);
#endif
        for (int j = 0; j < n; ++j) {
            if (!t[j].bytes) continue;
            float *dst = const_cast<float *>(t[j].g);
            const float *src = stage(k) + t[j].s;
            const bool add = kBackward && a.accumulate;
            if (bulk_ok(t[j])) {
                if (threadIdx.x == 0) {
                    if (add) bulk_s2g_add(dst, src, t[j].bytes);
                    else bulk_s2g(dst, src, t[j].bytes);
                }
            } else {
                const int words = (int)(t[j].bytes >> 2);
                if (add)
                    for (int e = threadIdx.x; e < words; e += kPoseThreads) atomicAdd(dst + e, src[e]);
                else
                    for (int e = threadIdx.x; e < words; e += kPoseThreads) dst[e] = src[e];
            }
        }
        if (threadIdx.x == 0) bulk_commit();
    }
};

struct NoPost {
    __device__ __forceinline__ void operator()(int) const {}
};

// body(L, stage, i, row, bones, cam) runs for every Gaussian of the tile; post(tile) runs once per tile on ALL threads after it
// (warp-collective work: the body is skipped by the threads past the end of the last tile)
template <bool kBackward, bool kFused = false, typename Body, typename Post = NoPost>
__device__ __forceinline__ void run_tiles(const PoseArgs &a, float *smem, Body body, Post post = Post()) {
    float *bones_s = smem, *cam_s = bones_s + kMaxBones * 13;
    uint64_t *bar = reinterpret_cast<uint64_t *>(cam_s + 4);
    float *rcam_s = cam_s + 8;      // fused backward: view | proj | tanx, tany, focx, focy
    TilePipe<kBackward, kFused> pipe{a, tile_layout(a.K, a.B, a.iso, kBackward), rcam_s + kCamFloats, bar};
    if (kFused) {
        if (threadIdx.x < 16) rcam_s[threadIdx.x] = a.view[threadIdx.x];
        else if (threadIdx.x < 32) rcam_s[threadIdx.x] = a.proj[threadIdx.x - 16];
        else if (threadIdx.x == 32) {
            const float tx = a.tanfov_dev ? a.tanfov_dev[0] : a.tanx, ty = a.tanfov_dev ? a.tanfov_dev[1] : a.tany;
            rcam_s[32] = tx; rcam_s[33] = ty;
            rcam_s[34] = a.W / (2.0f * tx); rcam_s[35] = a.H / (2.0f * ty);
        }
    }
    for (int j = threadIdx.x; j < a.B * 13; j += kPoseThreads) {
        const int b = j / 13, e = j - 13 * b, idx = e < 12 ? e : 15;
        if (a.bones_posed == nullptr) bones_s[j] = a.bone_tf[16 * b + idx];
        else if (b >= a.n_posed) bones_s[j] = (idx % 5 == 0) ? 1.f : 0.f;      // appended identity rows
        else {      // element (r, c) of bones_posed[b] * rest_inv[b]
            const int r = idx >> 2, c = idx & 3;
            const float *P = a.bones_posed + 16 * b + 4 * r, *R = a.rest_inv + 16 * b + c;
            bones_s[j] = fmaf(P[3], R[12], fmaf(P[2], R[8], fmaf(P[1], R[4], P[0] * R[0])));
        }
    }
    if (threadIdx.x < 3) cam_s[threadIdx.x] = a.campos[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int ntiles = (a.N + kPoseThreads - 1) / kPoseThreads;
    uint32_t phase = 0;
    // one stage per CTA, four CTAs per SM: the loads of one CTA overlap the arithmetic of the others
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();            // plain-path stores of the previous tile have read the stage
        if (threadIdx.x == 0) {
            bulk_wait_read_all();   // the bulk stores of the previous tile have finished reading the stage
            pipe.prefetch(tile, 0);
        }
        __syncthreads();
        pipe.acquire(tile, 0, phase);
        const int row = threadIdx.x, i = tile * kPoseThreads + row;
        if (kBackward && kFused) {
            // rows of the accumulator (array of structures) -> the per-array slots the body reads: every thread takes its
            // row into registers, then writes (moments m0 m1, radius | moments m2 m3 m4 | colour grad | opacity grad)
            float *st = pipe.stage(0);
            float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
            int rad = 0;
            if (i < a.N) {
                const float4 *rp = reinterpret_cast<const float4 *>(st + pipe.L.gpx + kAccRow * row);
                r0 = rp[0]; r1 = rp[1]; r2 = rp[2];
                rad = reinterpret_cast<const int *>(st + pipe.L.gpx + kPoseThreads * kAccRow)[row];
            }
            __syncthreads();
            if (i < a.N) {
                st[pipe.L.gpx + 3 * row] = r0.x; st[pipe.L.gpx + 3 * row + 1] = r0.y; st[pipe.L.gpx + 3 * row + 2] = __int_as_float(rad);
                st[pipe.L.gcov + 6 * row] = r0.z; st[pipe.L.gcov + 6 * row + 1] = r0.w; st[pipe.L.gcov + 6 * row + 2] = r1.x;
                st[pipe.L.gop + row] = r1.y;
                st[pipe.L.gcol + 3 * row] = r1.z; st[pipe.L.gcol + 3 * row + 1] = r1.w; st[pipe.L.gcol + 3 * row + 2] = r2.x;
            }
        }
#ifndef MB_POSE_NOCOMPUTE   // experiment switch (tools/pose_bench.py): data movement only
        if (i < a.N) body(pipe.L, pipe.stage(0), i, row, bones_s, cam_s);
#endif
        post(tile);
        pipe.release(tile, 0);
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

// kProject: the rasterizer's projection (A.1) runs on the posed mean / covariance / colour / opacity while they are in
// registers (mb_pose_project_forward): the 52 B per Gaussian between the two stages need not cross HBM, one launch less
template <int DEG, bool kProject>
__global__ void __launch_bounds__(kPoseThreads) pose_forward_kernel(PoseArgs a) {
    extern __shared__ __align__(128) float smem[];
    constexpr int nb = (DEG + 1) * (DEG + 1);
    bool visible = false;        // of this thread's Gaussian in the current tile (kProject)
    bool has_key = false;
    uint32_t my_tiles = 0, my_key = 0;
    // kProject: digit counts of the depth keys this CTA writes = the histogram of the depth sort that follows the kernel
    __shared__ uint32_t hist_s[kProject ? kSortMaxPasses * 256 : 1];
    if (kProject) hist_smem_zero(hist_s);       // run_tiles starts with a barrier
    run_tiles<false, kProject>(a, smem, [&](const TileLayout &L, float *st, int i, int row, const float *bones_s, const float *cam_s) {
        PoseLocal p;
        pose_common(a, L, st, i, row, bones_s, p);
        // mean
        float px[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) px[r] = p.A[3 * r] * p.x[0] + p.A[3 * r + 1] * p.x[1] + p.A[3 * r + 2] * p.x[2] + p.t[r];
        // covariance: Sigma' = (A L)(A L)^T
        float Bm[9];
        if (p.skinned) mat3_mul(p.A, p.L, Bm);
        else {
#pragma unroll
            for (int k = 0; k < 9; ++k) Bm[k] = p.L[k];
        }
        float c6[6];
        c6[0] = Bm[0] * Bm[0] + Bm[1] * Bm[1] + Bm[2] * Bm[2];
        c6[1] = Bm[0] * Bm[3] + Bm[1] * Bm[4] + Bm[2] * Bm[5];
        c6[2] = Bm[0] * Bm[6] + Bm[1] * Bm[7] + Bm[2] * Bm[8];
        c6[3] = Bm[3] * Bm[3] + Bm[4] * Bm[4] + Bm[5] * Bm[5];
        c6[4] = Bm[3] * Bm[6] + Bm[4] * Bm[7] + Bm[5] * Bm[8];
        c6[5] = Bm[6] * Bm[6] + Bm[7] * Bm[7] + Bm[8] * Bm[8];
        // colour
        float Ainv[9], ci[3], dir[3], basis[16];
        if (p.skinned) mat3_inverse(p.A, Ainv);
        view_dir(p, cam_s, Ainv, ci, dir);
        sh_basis(DEG, dir[0], dir[1], dir[2], basis);
        const float *fr = st + L.fr + row * (a.K - 1) * 3;
        float rgb[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = basis[0] * st[L.fdc + 3 * row + c];
#pragma unroll
            for (int k = 1; k < nb; ++k) v += basis[k] * fr[3 * (k - 1) + c];
            rgb[c] = fmaxf(v + 0.5f, 0.f);
        }
        const float ol = st[L.opac + row];
        if (a.tf_out && p.skinned) {
            float *o = a.tf_out + 16 * (size_t)i;
            o[0] = p.A[0]; o[1] = p.A[1]; o[2] = p.A[2]; o[3] = p.t[0];
            o[4] = p.A[3]; o[5] = p.A[4]; o[6] = p.A[5]; o[7] = p.t[1];
            o[8] = p.A[6]; o[9] = p.A[7]; o[10] = p.A[8]; o[11] = p.t[2];
            // bottom row = sum_b w_b T_b[3,:]
            float b0 = 0.f, b1 = 0.f, b2 = 0.f;
            for (int b = 0; b < a.B; ++b) {
                const float w = st[L.sk + row * a.B + b];
                b0 += w * a.bone_tf[16 * b + 12]; b1 += w * a.bone_tf[16 * b + 13]; b2 += w * a.bone_tf[16 * b + 14];
            }
            o[12] = b0; o[13] = b1; o[14] = b2; o[15] = p.s;
        }
        // results over the thread's own input rows (xyz -> posed xyz, f_dc -> colour, logit -> opacity)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            st[L.xyz + 3 * row + r] = px[r];
            st[L.fdc + 3 * row + r] = rgb[r];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) st[L.cov + 6 * row + k] = c6[k];
        const float op = 1.0f / (1.0f + expf(-ol));
        st[L.opac + row] = op;
        if (kProject) {
            const float *rc = cam_s + 8;      // view | proj | tanx, tany, focx, focy
            Projected pr;
            project_forward(rc, rc + 16, rc[32], rc[33], rc[34], rc[35], a.W, a.H, a.gx, a.gy, px[0], px[1], px[2], c6, op, rgb, pr);
            visible = pr.visible;
            my_tiles = pr.tiles;
            if (pr.visible) {
                a.rect[i] = pr.rect;
                a.rec[i] = pr.rec;
            }
            a.radii_out[i] = pr.radius;
            a.tiles_touched[i] = pr.tiles;
            a.depth_key[i] = pr.key;
            a.ident[i] = (uint32_t)i;
            my_key = pr.key;
            has_key = true;
        }
    }, [&](int) {
        if (kProject) {
            if (has_key) hist_smem_count(hist_s, my_key, 4, 2);
            has_key = false;
        }
        if (kProject) {      // per-warp totals of the tile: visible Gaussians and instances (num_rendered)
            const unsigned vis = __ballot_sync(0xffffffffu, visible);
            const uint32_t wt = __reduce_add_sync(0xffffffffu, my_tiles);
            if ((threadIdx.x & 31) == 0 && vis) {
                atomicAdd(&a.counters[kCntVisible], (uint32_t)__popc(vis));
                atomicAdd(&a.counters[kCntRendered], wt);
            }
            visible = false;
            my_tiles = 0;
        }
    });
    if (kProject) {
        __syncthreads();
        hist_smem_flush(hist_s, 4, a.depth_hist);
    }
}

// kFused: upstream gradients from the rasterizer's accumulator rows (mb_pose_backward_from_raster)
template <int DEG, bool kFused>
__global__ void __launch_bounds__(kPoseThreads, kFused ? 4 : 1) pose_backward_kernel(PoseArgs a) {
    extern __shared__ __align__(128) float smem[];
    constexpr int nb = (DEG + 1) * (DEG + 1);
    run_tiles<true, kFused>(a, smem, [&](const TileLayout &L, float *st, int i, int row, const float *bones_s, const float *cam_s) {
        PoseLocal p;
        pose_common(a, L, st, i, row, bones_s, p);
        float gx[3] = {0.f, 0.f, 0.f}, dA[9], dt[3] = {0.f, 0.f, 0.f}, ds = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) dA[k] = 0.f;
        float Bm[9], dB[9], dL[9];
        if (p.skinned) mat3_mul(p.A, p.L, Bm);
        else {
#pragma unroll
            for (int k = 0; k < 9; ++k) Bm[k] = p.L[k];
        }
        // upstream gradients of the posed mean and covariance: from the rasterizer's arrays, or (fused) from its accumulator row
        float gp[3], g6[6];
        if (kFused) {
            const float m[5] = {st[L.gpx + 3 * row], st[L.gpx + 3 * row + 1], st[L.gcov + 6 * row], st[L.gcov + 6 * row + 1], st[L.gcov + 6 * row + 2]};
            const int rad = __float_as_int(st[L.gpx + 3 * row + 2]);
            float g2[2] = {0.f, 0.f};
            gp[0] = gp[1] = gp[2] = 0.f;
#pragma unroll
            for (int k = 0; k < 6; ++k) g6[k] = 0.f;
            if (rad > 0) {
                // the posed mean, covariance and opacity exactly as pose_forward_kernel wrote them for the rasterizer's forward
                float pm[3], c6[6];
#pragma unroll
                for (int r = 0; r < 3; ++r) pm[r] = p.A[3 * r] * p.x[0] + p.A[3 * r + 1] * p.x[1] + p.A[3 * r + 2] * p.x[2] + p.t[r];
                c6[0] = Bm[0] * Bm[0] + Bm[1] * Bm[1] + Bm[2] * Bm[2];
                c6[1] = Bm[0] * Bm[3] + Bm[1] * Bm[4] + Bm[2] * Bm[5];
                c6[2] = Bm[0] * Bm[6] + Bm[1] * Bm[7] + Bm[2] * Bm[8];
                c6[3] = Bm[3] * Bm[3] + Bm[4] * Bm[4] + Bm[5] * Bm[5];
                c6[4] = Bm[3] * Bm[6] + Bm[4] * Bm[7] + Bm[5] * Bm[8];
                c6[5] = Bm[6] * Bm[6] + Bm[7] * Bm[7] + Bm[8] * Bm[8];
                const float op = 1.0f / (1.0f + expf(-st[L.opac + row]));
                const float *rc = cam_s + 8;
                project_backward(rc, rc + 16, rc[32], rc[33], rc[34], rc[35], a.W, a.H, pm[0], pm[1], pm[2], c6, op, m, g2, gp, g6);
            }
            a.g_means2D[3 * (size_t)i] = g2[0]; a.g_means2D[3 * (size_t)i + 1] = g2[1]; a.g_means2D[3 * (size_t)i + 2] = 0.f;
            if (a.stat_accum && rad > 0) {
                atomicAdd(a.stat_accum + i, sqrtf(g2[0] * g2[0] + g2[1] * g2[1]));
                atomicAdd(a.stat_denom + i, 1.0f);
                atomicMax(reinterpret_cast<int *>(a.stat_maxrad) + i, __float_as_int((float)rad));   // non-negative floats order like ints
            }
        } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) gp[r] = st[L.gpx + 3 * row + r];
#pragma unroll
            for (int k = 0; k < 6; ++k) g6[k] = st[L.gcov + 6 * row + k];
        }
        // ---- mean: x' = A x + t
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            gx[r] = p.A[r] * gp[0] + p.A[3 + r] * gp[1] + p.A[6 + r] * gp[2];
            dt[r] = gp[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) dA[3 * r + c] = gp[r] * p.x[c];
        }
        // ---- covariance: Sigma' = Bm Bm^T, Bm = A L ; Gs = symmetrised dL/dSigma'
        const float Gs[9] = {g6[0], 0.5f * g6[1], 0.5f * g6[2], 0.5f * g6[1], g6[3], 0.5f * g6[4], 0.5f * g6[2], 0.5f * g6[4], g6[5]};
        mat3_mul(Gs, Bm, dB);
#pragma unroll
        for (int k = 0; k < 9; ++k) dB[k] *= 2.f;
        if (p.skinned) {
            // dA += dB L^T ; dL = A^T dB
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    dA[3 * r + c] += dB[3 * r] * p.L[3 * c] + dB[3 * r + 1] * p.L[3 * c + 1] + dB[3 * r + 2] * p.L[3 * c + 2];
                    dL[3 * r + c] = p.A[r] * dB[c] + p.A[3 + r] * dB[3 + c] + p.A[6 + r] * dB[6 + c];
                }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) dL[k] = dB[k];
        }
        float dS[3], dR[9], dqn[4];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dS[k] = dL[k] * p.R[k] + dL[3 + k] * p.R[3 + k] + dL[6 + k] * p.R[6 + k];
#pragma unroll
            for (int r = 0; r < 3; ++r) dR[3 * r + k] = dL[3 * r + k] * p.S[k];
        }
        quat_to_rot_bwd(p.qn[0], p.qn[1], p.qn[2], p.qn[3], dR, dqn);
        const float qd = p.qn[0] * dqn[0] + p.qn[1] * dqn[1] + p.qn[2] * dqn[2] + p.qn[3] * dqn[3];
        float4 gq;
        gq.x = (dqn[0] - p.qn[0] * qd) / p.qnorm; gq.y = (dqn[1] - p.qn[1] * qd) / p.qnorm;
        gq.z = (dqn[2] - p.qn[2] * qd) / p.qnorm; gq.w = (dqn[3] - p.qn[3] * qd) / p.qnorm;
        *reinterpret_cast<float4 *>(st + L.quat + 4 * row) = gq;
        if (a.iso) st[L.ls + row] = dS[0] * p.S[0] + dS[1] * p.S[1] + dS[2] * p.S[2];
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) st[L.ls + 3 * row + k] = dS[k] * p.S[k];
        }
        // ---- colour
        float Ainv[9], ci[3], dir[3], basis[16], bxg[16], byg[16], bzg[16];
        if (p.skinned) mat3_inverse(p.A, Ainv);
        const float dn = view_dir(p, cam_s, Ainv, ci, dir);
        sh_basis(DEG, dir[0], dir[1], dir[2], basis);
        sh_basis_grad(DEG, dir[0], dir[1], dir[2], bxg, byg, bzg);
        float *fr = st + L.fr + row * (a.K - 1) * 3;
        float go[3], gd[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = basis[0] * st[L.fdc + 3 * row + c];
#pragma unroll
            for (int k = 1; k < nb; ++k) v += basis[k] * fr[3 * (k - 1) + c];
            go[c] = (v + 0.5f >= 0.f) ? st[L.gcol + 3 * row + c] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) st[L.gcol + 3 * row + c] = basis[0] * go[c];   // g_f_dc
#pragma unroll
        for (int k = 1; k < nb; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float sv = fr[3 * (k - 1) + c] * go[c];
                gd[0] += bxg[k] * sv; gd[1] += byg[k] * sv; gd[2] += bzg[k] * sv;
                if (a.g_f_rest) fr[3 * (k - 1) + c] = basis[k] * go[c];
            }
        if (a.g_f_rest)
            for (int k = nb; k < a.K; ++k)   // coefficients above the active degree get no gradient
#pragma unroll
                for (int c = 0; c < 3; ++c) fr[3 * (k - 1) + c] = 0.f;
        const float dot = dir[0] * gd[0] + dir[1] * gd[1] + dir[2] * gd[2];
        const float gdd[3] = {(gd[0] - dir[0] * dot) / dn, (gd[1] - dir[1] * dot) / dn, (gd[2] - dir[2] * dot) / dn};
#pragma unroll
        for (int r = 0; r < 3; ++r) gx[r] += gdd[r];
        if (p.skinned) {
            // ci = Ainv u, u = c - t/s ; gci = -gdd
            float du[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) du[r] = -(Ainv[r] * gdd[0] + Ainv[3 + r] * gdd[1] + Ainv[6 + r] * gdd[2]);
            const float is = 1.0f / p.s;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                dt[r] -= du[r] * is;
#pragma unroll
                for (int c = 0; c < 3; ++c) dA[3 * r + c] -= du[r] * ci[c];
            }
            ds += (p.t[0] * du[0] + p.t[1] * du[1] + p.t[2] * du[2]) * is * is;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) st[L.gpx + 3 * row + r] = gx[r];   // g_xyz
        const float sg = 1.0f / (1.0f + expf(-st[L.opac + row]));
        st[L.gop + row] = st[L.gop + row] * sg * (1.0f - sg);          // g_opacity_logit
        if (a.g_skin && p.skinned) {
            float *w_row = st + L.sk + row * a.B;
            for (int b = 0; b < a.B; ++b) {
                const float *T = bones_s + 13 * b;
                w_row[b] = dA[0] * T[0] + dA[1] * T[1] + dA[2] * T[2] + dt[0] * T[3] + dA[3] * T[4] + dA[4] * T[5] +
                           dA[5] * T[6] + dt[1] * T[7] + dA[6] * T[8] + dA[7] * T[9] + dA[8] * T[10] + dt[2] * T[11] +
                           ds * T[12];
            }
        }
    });
}

// ---- multi-view pose backward --------------------------------------------------------------------------------------
// The V views of one optimisation step (gradient accumulation, hand_dynamic.py:248,259-277) differ only in pose, camera and
// in what their tile backward left in the accumulator rows; the parameters are the same.  One pass: a tile's parameters are
// staged ONCE, every thread walks the V views of its Gaussian (its 48-byte accumulator row and radius come straight from global
// memory, one view ahead), sums the parameter gradients in registers and writes them once -- instead of V launches that each
// re-read 236 + 4B bytes of parameters per Gaussian and read-modify-write 236 bytes of gradients.  The f_rest gradient is
// rank one per view, sum_v basis_k(dir_v) go_v[c]: the loop keeps (dir_v, go_v) and the rows are rebuilt after the last view
// has read the coefficients.
constexpr int kViewCam = 44;      // floats per view in shared memory: campos 3 (+1) | view 16 | proj 16 | tanx, tany, focx, focy | pad

__host__ __device__ inline int multi_bones_floats(int B, int V) { return (V * B * 13 + 3) & ~3; }   // keeps what follows 16-byte aligned

inline size_t pose_multi_smem_bytes(int B, int K, int iso, int V) {
    return sizeof(float) * ((size_t)multi_bones_floats(B, V) + (size_t)V * kViewCam + 8) + sizeof(float) * (size_t)tile_layout(K, B, iso, true).floats;
}

template <int DEG>
__global__ void __launch_bounds__(kPoseThreads, 4) pose_backward_multi_kernel(PoseArgs a) {
    extern __shared__ __align__(128) float smem[];
    constexpr int nb = (DEG + 1) * (DEG + 1);
    const int V = a.n_views;
    float *bones_all = smem, *cams = bones_all + multi_bones_floats(a.B, V);
    uint64_t *bar = reinterpret_cast<uint64_t *>(cams + V * kViewCam);
    TilePipe<true, false> pipe{a, tile_layout(a.K, a.B, a.iso, true), cams + V * kViewCam + 8, bar};
    for (int j = threadIdx.x; j < V * a.B * 13; j += kPoseThreads) {
        const int v = j / (a.B * 13), jj = j - v * (a.B * 13);
        const int b = jj / 13, e = jj - 13 * b, idx = e < 12 ? e : 15;
        const ViewArgs &w = a.views[v];
        if (w.bones_posed == nullptr) bones_all[j] = w.bone_tf[16 * b + idx];
        else if (b >= a.n_posed) bones_all[j] = (idx % 5 == 0) ? 1.f : 0.f;
        else {
            const int r = idx >> 2, c = idx & 3;
            const float *P = w.bones_posed + 16 * b + 4 * r, *R = a.rest_inv + 16 * b + c;
            bones_all[j] = fmaf(P[3], R[12], fmaf(P[2], R[8], fmaf(P[1], R[4], P[0] * R[0])));
        }
    }
    for (int j = threadIdx.x; j < V * kViewCam; j += kPoseThreads) {
        const int v = j / kViewCam, e = j - v * kViewCam;
        const ViewArgs &w = a.views[v];
        float val = 0.f;
        if (e < 3) val = w.campos[e];
        else if (e >= 4 && e < 20) val = w.view[e - 4];
        else if (e >= 20 && e < 36) val = w.proj[e - 20];
        else if (e >= 36 && e < 40) {
            const float tx = w.tanfov_dev ? w.tanfov_dev[0] : w.tanx, ty = w.tanfov_dev ? w.tanfov_dev[1] : w.tany;
            val = e == 36 ? tx : e == 37 ? ty : e == 38 ? a.W / (2.0f * tx) : a.H / (2.0f * ty);
        }
        cams[j] = val;
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const TileLayout &L = pipe.L;
    const int ntiles = (a.N + kPoseThreads - 1) / kPoseThreads;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_wait_read_all();
            pipe.prefetch(tile, 0);
        }
        __syncthreads();
        const int row = threadIdx.x, i = tile * kPoseThreads + row;
        // the first view's accumulator row is requested before the tile's parameters have arrived
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0;
        int rad = 0;
        if (i < a.N) {
            const float4 *rp = reinterpret_cast<const float4 *>(a.views[0].acc + (size_t)i * kAccRow);
            r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
            rad = __ldg(a.views[0].radii + i);
        }
        pipe.acquire(tile, 0, phase);
        if (i < a.N) {
            float *st = pipe.stage(0);
            PoseLocal p;
            // ---- view-independent part: rotation, scales, L = R diag(S)
            p.x[0] = st[L.xyz + 3 * row]; p.x[1] = st[L.xyz + 3 * row + 1]; p.x[2] = st[L.xyz + 3 * row + 2];
            p.skinned = i < a.n_skinned;
            const uint64_t nzmask = p.skinned ? bone_mask(st + L.sk + row * a.B, a.B) : 0ull;
            {
                const float4 q = *reinterpret_cast<const float4 *>(st + L.quat + 4 * row);
                p.qnorm = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
                p.qn[0] = q.x / p.qnorm; p.qn[1] = q.y / p.qnorm; p.qn[2] = q.z / p.qnorm; p.qn[3] = q.w / p.qnorm;
                quat_to_rot(p.qn[0], p.qn[1], p.qn[2], p.qn[3], p.R);
                if (a.iso) p.S[0] = p.S[1] = p.S[2] = expf(st[L.ls + row]);
                else { p.S[0] = expf(st[L.ls + 3 * row]); p.S[1] = expf(st[L.ls + 3 * row + 1]); p.S[2] = expf(st[L.ls + 3 * row + 2]); }
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int k = 0; k < 3; ++k) p.L[3 * r + k] = p.R[3 * r + k] * p.S[k];
            }
            const float op = 1.0f / (1.0f + expf(-st[L.opac + row]));
            const float *fr = st + L.fr + row * (a.K - 1) * 3;
            const float fdc[3] = {st[L.fdc + 3 * row], st[L.fdc + 3 * row + 1], st[L.fdc + 3 * row + 2]};
            float gx[3] = {0.f, 0.f, 0.f}, dLs[9], gop = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) dLs[k] = 0.f;
            float go_v[kMaxViews][3], dir_v[kMaxViews][3];
            for (int v = 0; v < V; ++v) {
                const float4 c0 = r0, c1 = r1, c2 = r2;
                const int crad = rad;
                if (v + 1 < V) {      // next view's row in flight while this one is processed
                    const float4 *rp = reinterpret_cast<const float4 *>(a.views[v + 1].acc + (size_t)i * kAccRow);
                    r0 = __ldg(rp); r1 = __ldg(rp + 1); r2 = __ldg(rp + 2);
                    rad = __ldg(a.views[v + 1].radii + i);
                }
                const float *bones_s = bones_all + v * a.B * 13, *cam = cams + v * kViewCam;
                // ---- the view's blended transform
                if (p.skinned) {
                    blend_bones(st + L.sk + row * a.B, nzmask, bones_s, p);
                } else {
                    p.A[0] = p.A[4] = p.A[8] = 1.f;
                    p.A[1] = p.A[2] = p.A[3] = p.A[5] = p.A[6] = p.A[7] = 0.f;
                    p.t[0] = p.t[1] = p.t[2] = 0.f; p.s = 1.f;
                }
                float Bm[9];
                if (p.skinned) mat3_mul(p.A, p.L, Bm);
                else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) Bm[k] = p.L[k];
                }
                // ---- projection backward from the view's accumulator row
                float gp[3] = {0.f, 0.f, 0.f}, g6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, g2[2] = {0.f, 0.f};
                if (crad > 0) {
                    float pm[3], c6[6];
#pragma unroll
                    for (int r = 0; r < 3; ++r) pm[r] = p.A[3 * r] * p.x[0] + p.A[3 * r + 1] * p.x[1] + p.A[3 * r + 2] * p.x[2] + p.t[r];
                    c6[0] = Bm[0] * Bm[0] + Bm[1] * Bm[1] + Bm[2] * Bm[2];
                    c6[1] = Bm[0] * Bm[3] + Bm[1] * Bm[4] + Bm[2] * Bm[5];
                    c6[2] = Bm[0] * Bm[6] + Bm[1] * Bm[7] + Bm[2] * Bm[8];
                    c6[3] = Bm[3] * Bm[3] + Bm[4] * Bm[4] + Bm[5] * Bm[5];
                    c6[4] = Bm[3] * Bm[6] + Bm[4] * Bm[7] + Bm[5] * Bm[8];
                    c6[5] = Bm[6] * Bm[6] + Bm[7] * Bm[7] + Bm[8] * Bm[8];
                    const float m[5] = {c0.x, c0.y, c0.z, c0.w, c1.x};
                    project_backward(cam + 4, cam + 20, cam[36], cam[37], cam[38], cam[39], a.W, a.H, pm[0], pm[1], pm[2], c6, op, m, g2, gp, g6);
                }
                float *gm = a.views[v].g_means2D + 3 * (size_t)i;
                gm[0] = g2[0]; gm[1] = g2[1]; gm[2] = 0.f;
                if (a.stat_accum && crad > 0) {
                    atomicAdd(a.stat_accum + i, sqrtf(g2[0] * g2[0] + g2[1] * g2[1]));
                    atomicAdd(a.stat_denom + i, 1.0f);
                    atomicMax(reinterpret_cast<int *>(a.stat_maxrad) + i, __float_as_int((float)crad));
                }
                gop += c1.y;
                // ---- mean and covariance
#pragma unroll
                for (int r = 0; r < 3; ++r) gx[r] += p.A[r] * gp[0] + p.A[3 + r] * gp[1] + p.A[6 + r] * gp[2];
                const float Gs[9] = {g6[0], 0.5f * g6[1], 0.5f * g6[2], 0.5f * g6[1], g6[3], 0.5f * g6[4], 0.5f * g6[2], 0.5f * g6[4], g6[5]};
                float dB[9];
                mat3_mul(Gs, Bm, dB);
                if (p.skinned) {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 3; ++c) dLs[3 * r + c] += 2.f * (p.A[r] * dB[c] + p.A[3 + r] * dB[3 + c] + p.A[6 + r] * dB[6 + c]);
                } else {
#pragma unroll
                    for (int k = 0; k < 9; ++k) dLs[k] += 2.f * dB[k];
                }
                // ---- colour: the view direction in canonical space
                float Ainv[9], ci[3], dir[3], basis[16], bxg[16], byg[16], bzg[16];
                if (p.skinned) mat3_inverse(p.A, Ainv);
                const float dn = view_dir(p, cam, Ainv, ci, dir);
                sh_basis(DEG, dir[0], dir[1], dir[2], basis);
                sh_basis_grad(DEG, dir[0], dir[1], dir[2], bxg, byg, bzg);
                const float gcol[3] = {c1.z, c1.w, c2.x};
                float go[3], gd[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float val = basis[0] * fdc[c];
#pragma unroll
                    for (int k = 1; k < nb; ++k) val += basis[k] * fr[3 * (k - 1) + c];
                    go[c] = (val + 0.5f >= 0.f) ? gcol[c] : 0.f;
                }
#pragma unroll
                for (int k = 1; k < nb; ++k)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float sv = fr[3 * (k - 1) + c] * go[c];
                        gd[0] += bxg[k] * sv; gd[1] += byg[k] * sv; gd[2] += bzg[k] * sv;
                    }
                const float dot = dir[0] * gd[0] + dir[1] * gd[1] + dir[2] * gd[2];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    gx[r] += (gd[r] - dir[r] * dot) / dn;
                    go_v[v][r] = go[r];
                    dir_v[v][r] = dir[r];
                }
            }
            // ---- view-independent tail: scale / rotation gradients from the summed dL/dL
            float dS[3], dR[9], dqn[4];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dS[k] = dLs[k] * p.R[k] + dLs[3 + k] * p.R[3 + k] + dLs[6 + k] * p.R[6 + k];
#pragma unroll
                for (int r = 0; r < 3; ++r) dR[3 * r + k] = dLs[3 * r + k] * p.S[k];
            }
            quat_to_rot_bwd(p.qn[0], p.qn[1], p.qn[2], p.qn[3], dR, dqn);
            const float qd = p.qn[0] * dqn[0] + p.qn[1] * dqn[1] + p.qn[2] * dqn[2] + p.qn[3] * dqn[3];
            float4 gq;
            gq.x = (dqn[0] - p.qn[0] * qd) / p.qnorm; gq.y = (dqn[1] - p.qn[1] * qd) / p.qnorm;
            gq.z = (dqn[2] - p.qn[2] * qd) / p.qnorm; gq.w = (dqn[3] - p.qn[3] * qd) / p.qnorm;
            *reinterpret_cast<float4 *>(st + L.quat + 4 * row) = gq;
            if (a.iso) st[L.ls + row] = dS[0] * p.S[0] + dS[1] * p.S[1] + dS[2] * p.S[2];
            else {
#pragma unroll
                for (int k = 0; k < 3; ++k) st[L.ls + 3 * row + k] = dS[k] * p.S[k];
            }
            // ---- SH gradients: sum over the views of basis_k(dir_v) go_v (the coefficient rows are no longer needed)
            // (the sums over the views stay in registers -- most of the view loop's state is dead here -- and each coefficient is
            // stored once)
            float gdc[3] = {0.f, 0.f, 0.f};
            float *frw = st + L.fr + row * (a.K - 1) * 3;
            if (a.g_f_rest) {
                constexpr int nr = nb > 1 ? (nb - 1) * 3 : 1;
                float gfr[nr];
#pragma unroll
                for (int k = 0; k < nr; ++k) gfr[k] = 0.f;
                for (int v = 0; v < V; ++v) {
                    float basis[16];
                    sh_basis(DEG, dir_v[v][0], dir_v[v][1], dir_v[v][2], basis);
#pragma unroll
                    for (int c = 0; c < 3; ++c) gdc[c] += basis[0] * go_v[v][c];
#pragma unroll
                    for (int k = 1; k < nb; ++k)
#pragma unroll
                        for (int c = 0; c < 3; ++c) gfr[3 * (k - 1) + c] += basis[k] * go_v[v][c];
                }
#pragma unroll
                for (int k = 1; k < nb; ++k)
#pragma unroll
                    for (int c = 0; c < 3; ++c) frw[3 * (k - 1) + c] = gfr[3 * (k - 1) + c];
                for (int k = nb; k < a.K; ++k)
#pragma unroll
                    for (int c = 0; c < 3; ++c) frw[3 * (k - 1) + c] = 0.f;
            } else {
                for (int v = 0; v < V; ++v) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) gdc[c] += MB_SH_C0 * go_v[v][c];
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) st[L.gcol + 3 * row + c] = gdc[c];
#pragma unroll
            for (int r = 0; r < 3; ++r) st[L.gpx + 3 * row + r] = gx[r];
            st[L.gop + row] = gop * op * (1.0f - op);
        }
        pipe.release(tile, 0);
    }
    if (threadIdx.x == 0) bulk_wait_all();
}

// Rebuilds the SH-coefficient gradients of a sum over R views from what every view contributes through its DC term.
// For one view, g_f_dc[c] = basis_0 * go[c] and g_f_rest[k][c] = basis_k(dir) * go[c], where go is the colour gradient after
// the clamp mask and dir the view direction in canonical space (function of the view's bone transforms and camera centre).
// So ranks exchange g_f_dc (3 floats per Gaussian and view, all-gather) plus the tiny per-view (bone_tf, campos) and every
// rank recomputes dir per view: 12 R bytes per Gaussian on the wire instead of an all-reduce of 192.
struct ShViewsArgs {
    int N, n_skinned, B, K, R;
    int64_t s_bone, s_cam, s_g;     // floats between consecutive views in bone_tf_all / campos_all / g_fdc_all
    const float *xyz, *skin, *bone_tf_all, *campos_all, *g_fdc_all;
    float *g_f_dc, *g_f_rest;
};

constexpr int kShvTile = 128;   // Gaussians per CTA
constexpr int kShvNz = 16;      // non-zero skin weights kept per Gaussian in compacted form
constexpr int kShvNzStride = 2 * kShvNz + 1;   // odd: the per-thread rows do not collide on shared-memory banks

// shared memory (floats): bones [R][B][13] | cameras [R][3] | skin tile [128][B] | compacted weights [128][2 kShvNz + 1];
// the output tile [128][3 (K-1)] (+1 pad per row) re-uses the skin / compacted-weight area once every thread is done with it
template <int DEG>
__global__ void __launch_bounds__(kShvTile, 6) sh_grad_from_views_kernel(ShViewsArgs a) {
    constexpr int nb = (DEG + 1) * (DEG + 1);
    extern __shared__ float sh_smem[];
    float *bones_s = sh_smem, *cam_s = bones_s + (size_t)a.R * a.B * 13;
    float *skin_s = cam_s + 3 * a.R + ((3 * a.R) & 1);
    const int rs = (a.K - 1) * 3, rs_pad = rs | 1;            // odd row stride: conflict-free row-wise access
    float *nz_s = skin_s + (size_t)kShvTile * a.B;
    float *out_s = skin_s;
    const int tid = threadIdx.x, base = blockIdx.x * kShvTile, cnt = min(kShvTile, a.N - base);
    for (int j = tid; j < a.R * a.B * 13; j += kShvTile) {
        const int rb = j / 13, e = j - 13 * rb;
        const int r = rb / a.B, b = rb - r * a.B;
        bones_s[j] = a.bone_tf_all[(size_t)r * a.s_bone + 16 * b + (e < 12 ? e : 15)];
    }
    for (int j = tid; j < a.R * 3; j += kShvTile) cam_s[j] = a.campos_all[(size_t)(j / 3) * a.s_cam + j % 3];
    // the tile's skin weights are one dense run: coalesced load, row-wise use
    const int nsk = max(0, min(cnt, a.n_skinned - base));
    for (int j = tid; j < nsk * a.B; j += kShvTile) skin_s[j] = a.skin[(size_t)base * a.B + j];
    __syncthreads();
    const int i = base + tid;
    float acc[nb][3];
#pragma unroll
    for (int k = 0; k < nb; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
    if (tid < cnt) {
        const float x[3] = {a.xyz[3 * (size_t)i], a.xyz[3 * (size_t)i + 1], a.xyz[3 * (size_t)i + 2]};
        const bool skinned = i < a.n_skinned;
        // the weight row is sparse (a few bones per Gaussian): compact it once into (weight, bone) pairs so that every view
        // only walks the non-zero entries; rows with more than kShvNz non-zeros keep the full scan
        const float *w_row = skin_s + tid * a.B;
        float *nz = nz_s + tid * kShvNzStride;
        int nnz = 0;
        if (skinned) {
            for (int b = 0; b < a.B; ++b) {
                const float w = w_row[b];
                if (w == 0.f) continue;
                if (nnz == kShvNz) { nnz = -1; break; }
                nz[2 * nnz] = w;
                nz[2 * nnz + 1] = __int_as_float(b);
                ++nnz;
            }
        }
        for (int r = 0; r < a.R; ++r) {
            // this view's DC gradient: requested first, used last (its latency hides behind the blend)
            const float *g = a.g_fdc_all + (size_t)r * a.s_g + (size_t)i * 3;
            const float g0 = g[0], g1 = g[1], g2 = g[2];
            PoseLocal p;
            p.x[0] = x[0]; p.x[1] = x[1]; p.x[2] = x[2];
            p.skinned = skinned;
            float Ainv[9], ci[3], dir[3], basis[16];
            if (skinned) {
#pragma unroll
                for (int k = 0; k < 9; ++k) p.A[k] = 0.f;
                p.t[0] = p.t[1] = p.t[2] = 0.f; p.s = 0.f;
                const float *bones = bones_s + (size_t)r * a.B * 13;
                const int n_e = nnz >= 0 ? nnz : a.B;
                for (int e = 0; e < n_e; ++e) {
                    const float w = nnz >= 0 ? nz[2 * e] : w_row[e];
                    const int b = nnz >= 0 ? __float_as_int(nz[2 * e + 1]) : e;
                    const float *T = bones + 13 * b;
                    p.A[0] += w * T[0]; p.A[1] += w * T[1]; p.A[2] += w * T[2]; p.t[0] += w * T[3];
                    p.A[3] += w * T[4]; p.A[4] += w * T[5]; p.A[5] += w * T[6]; p.t[1] += w * T[7];
                    p.A[6] += w * T[8]; p.A[7] += w * T[9]; p.A[8] += w * T[10]; p.t[2] += w * T[11];
                    p.s += w * T[12];
                }
                mat3_inverse(p.A, Ainv);
            }
            view_dir(p, cam_s + 3 * r, Ainv, ci, dir);
            sh_basis(DEG, dir[0], dir[1], dir[2], basis);
            const float inv_c0 = 1.0f / MB_SH_C0;
            const float go[3] = {g0 * inv_c0, g1 * inv_c0, g2 * inv_c0};
#pragma unroll
            for (int k = 0; k < nb; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[k][c] += basis[k] * go[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) a.g_f_dc[3 * (size_t)i + c] = acc[0][c];
    }
    __syncthreads();              // every thread of the tile is done with the skin / compacted-weight area: it becomes the output tile
    if (tid < cnt) {
        float *row = out_s + tid * rs_pad;
#pragma unroll
        for (int k = 1; k < nb; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) row[3 * (k - 1) + c] = acc[k][c];
        for (int k = nb; k < a.K; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) row[3 * (k - 1) + c] = 0.f;
    }
    __syncthreads();
    // the tile's f_rest gradients are one dense run in global memory: coalesced store
    float *out = a.g_f_rest + (size_t)base * rs;
    for (int j = tid; j < cnt * rs; j += kShvTile) {
        const int rowi = j / rs, e = j - rowi * rs;
        out[j] = out_s[rowi * rs_pad + e];
    }
}

static int validate_pose(const mb_pose_inputs *in, const char *who) {
    MB_REQUIRE(in != nullptr, "%s: null inputs", who);
    MB_REQUIRE(in->num_points >= 0 && in->num_skinned >= 0 && in->num_skinned <= in->num_points, "%s: bad counts N=%d skinned=%d",
               who, in->num_points, in->num_skinned);
    MB_REQUIRE(in->sh_degree >= 0 && in->sh_degree <= 3, "%s: sh_degree %d not in 0..3", who, in->sh_degree);
    MB_REQUIRE(in->sh_coeffs >= (in->sh_degree + 1) * (in->sh_degree + 1) && in->sh_coeffs <= 16,
               "%s: %d SH coefficients, degree %d needs %d (max 16)", who, in->sh_coeffs, in->sh_degree,
               (in->sh_degree + 1) * (in->sh_degree + 1));
    if (in->num_points == 0) return MB_OK;
    MB_REQUIRE(in->xyz && in->log_scale && in->quat && in->opacity_logit && in->f_dc && in->campos, "%s: null parameter tensor", who);
    MB_REQUIRE(in->sh_coeffs == 1 || in->f_rest, "%s: f_rest missing", who);
    if (in->num_skinned > 0) {
        MB_REQUIRE(in->skin_wts && (in->bone_tf || (in->bones_posed && in->bones_rest_inv)), "%s: skin_wts / bone_tf missing", who);
        MB_REQUIRE(!in->bones_posed || (in->bones_rest_inv && in->num_posed_bones >= 0 && in->num_posed_bones <= in->num_bones),
                   "%s: bones_posed needs bones_rest_inv and 0 <= num_posed_bones <= num_bones", who);
        MB_REQUIRE(in->num_bones > 0 && in->num_bones <= kMaxBones, "%s: num_bones %d not in 1..%d", who, in->num_bones, kMaxBones);
    }
    return MB_OK;
}

static PoseArgs pose_args(const mb_pose_inputs *in) {
    PoseArgs a = {};
    a.N = in->num_points; a.n_skinned = in->num_skinned; a.B = in->num_skinned > 0 ? in->num_bones : 0;
    a.deg = in->sh_degree; a.K = in->sh_coeffs; a.iso = in->isotropic;
    a.xyz = in->xyz; a.log_scale = in->log_scale; a.quat = in->quat; a.opacity_logit = in->opacity_logit;
    a.f_dc = in->f_dc; a.f_rest = in->f_rest; a.skin = in->skin_wts; a.bone_tf = in->bone_tf; a.campos = in->campos;
    a.bones_posed = in->bones_posed; a.rest_inv = in->bones_rest_inv; a.n_posed = in->num_posed_bones;
    return a;
}

template <int DEG>
static int launch_pose(PoseArgs a, bool backward, cudaStream_t s) {
    const size_t smem = pose_smem_bytes(a.B, a.K, a.iso, backward);
    const int ntiles = (a.N + kPoseThreads - 1) / kPoseThreads;
    // (the fused projection adds the 4 KB static digit histogram of the depth sort)
    int per_sm = (int)((size_t)(220 * 1024) / (smem + (!backward && a.rec ? sizeof(uint32_t) * kSortMaxPasses * 256 : 0)));
    if (per_sm < 1) {
        set_error("pose kernel: a tile of %d Gaussians with %d SH coefficients and %d bones needs %zu bytes of shared memory", kPoseThreads,
                  a.K, a.B, smem);
        return MB_ERR_INVALID;
    }
    if (per_sm > 16) per_sm = 16;
    const int grid = min(ntiles, sm_count() * per_sm);
    // raise the dynamic shared-memory limit once per (kernel, device): not a stream operation, kept out of the per-frame path
    static thread_local int limit[2][16] = {};
    int dev = 0;
    MB_CUDA(cudaGetDevice(&dev));
    int &cur = limit[backward ? 1 : 0][dev & 15];
    if (cur < (int)smem) {
        if (backward) {
            MB_CUDA(cudaFuncSetAttribute(pose_backward_kernel<DEG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MB_CUDA(cudaFuncSetAttribute(pose_backward_kernel<DEG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        } else {
            MB_CUDA(cudaFuncSetAttribute(pose_forward_kernel<DEG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            MB_CUDA(cudaFuncSetAttribute(pose_forward_kernel<DEG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        cur = (int)smem;
    }
    if (backward) {
        KernelTimer kt("pose_backward", s);
        if (a.acc) pose_backward_kernel<DEG, true><<<grid, kPoseThreads, smem, s>>>(a);
        else pose_backward_kernel<DEG, false><<<grid, kPoseThreads, smem, s>>>(a);
    } else {
        KernelTimer kt("pose_forward", s);
        if (a.rec) pose_forward_kernel<DEG, true><<<grid, kPoseThreads, smem, s>>>(a);
        else pose_forward_kernel<DEG, false><<<grid, kPoseThreads, smem, s>>>(a);
    }
    return check_launch(backward ? "pose_backward" : "pose_forward", false, s);
}

static int launch_pose_deg(const PoseArgs &a, bool backward, cudaStream_t s) {
    switch (a.deg) {
        case 0: return launch_pose<0>(a, backward, s);
        case 1: return launch_pose<1>(a, backward, s);
        case 2: return launch_pose<2>(a, backward, s);
        default: return launch_pose<3>(a, backward, s);
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_pose_forward(const mb_pose_inputs *in, float *posed_xyz, float *posed_cov6, float *colors, float *opacity,
                               float *tf_out, mb_stream_t stream) {
    int rc = validate_pose(in, "mb_pose_forward");
    if (rc) return rc;
    if (in->num_points == 0) return MB_OK;
    MB_REQUIRE(posed_xyz && posed_cov6 && colors && opacity, "mb_pose_forward: null output");
    PoseArgs a = pose_args(in);
    a.posed_xyz = posed_xyz; a.cov6 = posed_cov6; a.colors = colors; a.opacity = opacity; a.tf_out = tf_out;
    return launch_pose_deg(a, false, (cudaStream_t)stream);
}

extern "C" int mb_pose_project_forward(const mb_pose_inputs *in, const mb_raster_inputs *raster, void *geom, size_t geom_bytes,
                                       int32_t *radii, int64_t *num_rendered_host, float *posed_xyz, float *posed_cov6, float *colors,
                                       float *opacity, mb_stream_t stream) {
    int rc = validate_pose(in, "mb_pose_project_forward");
    if (rc) return rc;
    MB_REQUIRE(raster != nullptr && raster->num_points == in->num_points, "mb_pose_project_forward: raster inputs missing or of another size");
    MB_REQUIRE(raster->viewmatrix && raster->projmatrix && raster->image_width > 0 && raster->image_height > 0 &&
                   (raster->tanfov_dev || (raster->tanfovx > 0.f && raster->tanfovy > 0.f)),
               "mb_pose_project_forward: camera missing");
    MB_REQUIRE(raster->scale_modifier == 1.0f && raster->shs == nullptr, "mb_pose_project_forward: scale_modifier must be 1, colours come from the pose step");
    cudaStream_t s = (cudaStream_t)stream;
    const RasterDims d = raster_dims(raster);
    MB_REQUIRE(geom != nullptr && (d.P == 0 || radii != nullptr), "mb_pose_project_forward: null geom / radii");
    GeomState g = GeomState::carve(geom, d.P);
    if (geom_bytes < g.bytes) {
        set_error("mb_pose_project_forward: geom buffer has %zu bytes, needs %zu", geom_bytes, g.bytes);
        return MB_ERR_WORKSPACE;
    }
    // counters + look-back words of the instance-offset scan + the depth sort's histograms / cursors / look-back words
    // (contiguous), as mb_raster_forward_geom: the kernel counts the digits of the depth keys it writes
    MB_CUDA(cudaMemsetAsync(g.counters, 0, g.zeroed_bytes(d.P), s));
    if (d.P > 0) {
        SortWorkspace ws = carve_sort_workspace(g.sort_ws, d.P);
        PoseArgs a = pose_args(in);
        a.depth_hist = ws.hist;
        a.posed_xyz = posed_xyz; a.cov6 = posed_cov6; a.colors = colors; a.opacity = opacity;
        a.view = raster->viewmatrix; a.proj = raster->projmatrix; a.tanfov_dev = raster->tanfov_dev;
        a.tanx = raster->tanfovx; a.tany = raster->tanfovy; a.W = d.W; a.H = d.H; a.gx = d.gx; a.gy = d.gy;
        a.rec = g.rec; a.rect = g.rect; a.tiles_touched = g.tiles_touched; a.depth_key = g.depth_key; a.ident = g.ident;
        a.counters = g.counters; a.radii_out = radii;
        rc = launch_pose_deg(a, false, s);
        if (rc) return rc;
        rc = radix_sort_pairs(g.depth_key, g.ident, g.sorted_key, g.sorted_idx, d.P, nullptr, d.P, 0, 32, ws, s, raster->debug != 0, true);
        if (rc) return rc;
    }
    if (num_rendered_host) {
        *num_rendered_host = 0;
        MB_CUDA(cudaMemcpyAsync(num_rendered_host, g.counters + kCntRendered, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    return MB_OK;
}

static int pose_backward_impl(const mb_pose_inputs *in, const float *g_posed_xyz, const float *g_posed_cov6,
                              const float *g_colors, const float *g_opacity, float *g_xyz, float *g_log_scale, float *g_quat,
                              float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts, int accumulate,
                              mb_stream_t stream) {
    int rc = validate_pose(in, "mb_pose_backward");
    if (rc) return rc;
    if (in->num_points == 0) return MB_OK;
    MB_REQUIRE(g_posed_xyz && g_posed_cov6 && g_colors && g_opacity, "mb_pose_backward: null upstream gradient");
    // g_f_rest may be NULL: the SH gradient of one view is rank one, g_f_rest[k][c] = basis_k(dir) * g_f_dc[c] / basis_0, and
    // can be rebuilt from g_f_dc by mb_sh_grad_from_views (what the data-parallel step exchanges instead of 45 floats)
    MB_REQUIRE(g_xyz && g_log_scale && g_quat && g_opacity_logit && g_f_dc, "mb_pose_backward: null output");
    PoseArgs a = pose_args(in);
    a.g_posed_xyz = g_posed_xyz; a.g_cov6 = g_posed_cov6; a.g_colors = g_colors; a.g_opacity = g_opacity;
    a.g_xyz = g_xyz; a.g_log_scale = g_log_scale; a.g_quat = g_quat; a.g_opacity_logit = g_opacity_logit;
    a.g_f_dc = g_f_dc; a.g_f_rest = g_f_rest; a.g_skin = g_skin_wts;
    a.accumulate = accumulate;
    return launch_pose_deg(a, true, (cudaStream_t)stream);
}

extern "C" int mb_pose_backward(const mb_pose_inputs *in, const float *g_posed_xyz, const float *g_posed_cov6,
                                const float *g_colors, const float *g_opacity, float *g_xyz, float *g_log_scale, float *g_quat,
                                float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts, mb_stream_t stream) {
    return pose_backward_impl(in, g_posed_xyz, g_posed_cov6, g_colors, g_opacity, g_xyz, g_log_scale, g_quat, g_opacity_logit, g_f_dc,
                              g_f_rest, g_skin_wts, 0, stream);
}

extern "C" int mb_pose_backward_from_raster(const mb_pose_inputs *in, const mb_raster_inputs *raster, const int32_t *radii,
                                           const void *grad_scratch, float *dL_dmeans2D, float *g_xyz, float *g_log_scale,
                                           float *g_quat, float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts,
                                           int32_t accumulate, float *xyz_gradient_accum, float *denom, float *max_radii2D,
                                           mb_stream_t stream) {
    int rc = validate_pose(in, "mb_pose_backward_from_raster");
    if (rc) return rc;
    MB_REQUIRE(raster != nullptr && raster->num_points == in->num_points, "mb_pose_backward_from_raster: raster inputs missing or of another size");
    if (in->num_points == 0) return MB_OK;
    MB_REQUIRE(raster->viewmatrix && raster->projmatrix && raster->image_width > 0 && raster->image_height > 0 &&
                   (raster->tanfov_dev || (raster->tanfovx > 0.f && raster->tanfovy > 0.f)),
               "mb_pose_backward_from_raster: camera missing");
    MB_REQUIRE(raster->shs == nullptr && raster->scales == nullptr && raster->scale_modifier == 1.0f,
               "mb_pose_backward_from_raster: the rasterizer must have run on this pose's colours and covariances (scale_modifier 1)");
    MB_REQUIRE(radii && grad_scratch && dL_dmeans2D, "mb_pose_backward_from_raster: null radii / accumulator / dL_dmeans2D");
    MB_REQUIRE(g_xyz && g_log_scale && g_quat && g_opacity_logit && g_f_dc, "mb_pose_backward_from_raster: null output");
    PoseArgs a = pose_args(in);
    a.g_xyz = g_xyz; a.g_log_scale = g_log_scale; a.g_quat = g_quat; a.g_opacity_logit = g_opacity_logit;
    a.g_f_dc = g_f_dc; a.g_f_rest = g_f_rest; a.g_skin = g_skin_wts;
    a.accumulate = accumulate;
    a.acc = reinterpret_cast<const float *>(grad_scratch); a.radii = radii; a.g_means2D = dL_dmeans2D;
    MB_REQUIRE((xyz_gradient_accum != nullptr) == (denom != nullptr) && (denom != nullptr) == (max_radii2D != nullptr),
               "mb_pose_backward_from_raster: give all three densification statistics or none");
    a.stat_accum = xyz_gradient_accum; a.stat_denom = denom; a.stat_maxrad = max_radii2D;
    a.view = raster->viewmatrix; a.proj = raster->projmatrix; a.tanfov_dev = raster->tanfov_dev;
    a.tanx = raster->tanfovx; a.tany = raster->tanfovy; a.W = raster->image_width; a.H = raster->image_height;
    return launch_pose_deg(a, true, (cudaStream_t)stream);
}

template <int DEG>
static int launch_pose_multi(const PoseArgs &a, cudaStream_t s) {
    const size_t smem = pose_multi_smem_bytes(a.B, a.K, a.iso, a.n_views);
    int per_sm = (int)((size_t)(220 * 1024) / smem);
    if (per_sm < 1) {
        set_error("multi-view pose backward: %d views x %d bones and a tile of %d Gaussians need %zu bytes of shared memory", a.n_views, a.B,
                  kPoseThreads, smem);
        return MB_ERR_INVALID;
    }
    if (per_sm > 4) per_sm = 4;
    const int ntiles = (a.N + kPoseThreads - 1) / kPoseThreads;
    const int grid = min(ntiles, sm_count() * per_sm);
    static thread_local int limit[16] = {};
    int dev = 0;
    MB_CUDA(cudaGetDevice(&dev));
    if (limit[dev & 15] < (int)smem) {
        MB_CUDA(cudaFuncSetAttribute(pose_backward_multi_kernel<DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        limit[dev & 15] = (int)smem;
    }
    KernelTimer kt("pose_backward", s);
    pose_backward_multi_kernel<DEG><<<grid, kPoseThreads, smem, s>>>(a);
    return check_launch("pose_backward_multi", false, s);
}

extern "C" int mb_pose_backward_from_raster_views(const mb_pose_inputs *in, int32_t num_views, const mb_view_inputs *views, float *g_xyz,
                                                 float *g_log_scale, float *g_quat, float *g_opacity_logit, float *g_f_dc, float *g_f_rest,
                                                 int32_t accumulate, float *xyz_gradient_accum, float *denom, float *max_radii2D,
                                                 mb_stream_t stream) {
    MB_REQUIRE(in != nullptr && views != nullptr && num_views >= 1 && num_views <= kMaxViews, "mb_pose_backward_from_raster_views: 1..%d views", kMaxViews);
    // the per-view bone transforms replace those of `in`; everything else of `in` is validated as usual
    mb_pose_inputs chk = *in;
    chk.bone_tf = views[0].bone_tf; chk.bones_posed = views[0].bones_posed;
    int rc = validate_pose(&chk, "mb_pose_backward_from_raster_views");
    if (rc) return rc;
    if (in->num_points == 0) return MB_OK;
    MB_REQUIRE(g_xyz && g_log_scale && g_quat && g_opacity_logit && g_f_dc, "mb_pose_backward_from_raster_views: null output");
    MB_REQUIRE((xyz_gradient_accum != nullptr) == (denom != nullptr) && (denom != nullptr) == (max_radii2D != nullptr),
               "mb_pose_backward_from_raster_views: give all three densification statistics or none");
    PoseArgs a = pose_args(in);
    a.g_xyz = g_xyz; a.g_log_scale = g_log_scale; a.g_quat = g_quat; a.g_opacity_logit = g_opacity_logit;
    a.g_f_dc = g_f_dc; a.g_f_rest = g_f_rest; a.g_skin = nullptr;
    a.accumulate = accumulate;
    a.stat_accum = xyz_gradient_accum; a.stat_denom = denom; a.stat_maxrad = max_radii2D;
    a.n_views = num_views;
    for (int v = 0; v < num_views; ++v) {
        const mb_view_inputs &w = views[v];
        const mb_raster_inputs *r = w.raster;
        MB_REQUIRE(r != nullptr && r->num_points == in->num_points, "mb_pose_backward_from_raster_views: view %d: raster inputs missing or of another size", v);
        MB_REQUIRE(r->viewmatrix && r->projmatrix && r->image_width > 0 && r->image_height > 0 && (r->tanfov_dev || (r->tanfovx > 0.f && r->tanfovy > 0.f)),
                   "mb_pose_backward_from_raster_views: view %d: camera missing", v);
        MB_REQUIRE(r->shs == nullptr && r->scales == nullptr && r->scale_modifier == 1.0f,
                   "mb_pose_backward_from_raster_views: view %d: the rasterizer must have run on this pose's colours and covariances", v);
        MB_REQUIRE(v == 0 || (r->image_width == views[0].raster->image_width && r->image_height == views[0].raster->image_height),
                   "mb_pose_backward_from_raster_views: the views of a step share the image size");
        MB_REQUIRE(w.radii && w.grad_scratch && w.dL_dmeans2D && w.campos, "mb_pose_backward_from_raster_views: view %d: null radii / accumulator / dL_dmeans2D / campos", v);
        MB_REQUIRE(in->num_skinned == 0 || w.bone_tf || (w.bones_posed && in->bones_rest_inv), "mb_pose_backward_from_raster_views: view %d: bone transforms missing", v);
        MB_REQUIRE((w.bones_posed != nullptr) == (views[0].bones_posed != nullptr), "mb_pose_backward_from_raster_views: give every view's bones the same way");
        ViewArgs &o = a.views[v];
        o.acc = reinterpret_cast<const float *>(w.grad_scratch); o.radii = w.radii; o.bone_tf = w.bone_tf; o.bones_posed = w.bones_posed;
        o.campos = w.campos; o.view = r->viewmatrix; o.proj = r->projmatrix; o.tanfov_dev = r->tanfov_dev; o.tanx = r->tanfovx; o.tany = r->tanfovy;
        o.g_means2D = w.dL_dmeans2D;
    }
    a.bones_posed = views[0].bones_posed;      // selects the in-kernel product (rest_inv / n_posed come from `in`)
    a.W = views[0].raster->image_width; a.H = views[0].raster->image_height;
    cudaStream_t s = (cudaStream_t)stream;
    switch (a.deg) {
        case 0: return launch_pose_multi<0>(a, s);
        case 1: return launch_pose_multi<1>(a, s);
        case 2: return launch_pose_multi<2>(a, s);
        default: return launch_pose_multi<3>(a, s);
    }
}

extern "C" int mb_pose_backward_accumulate(const mb_pose_inputs *in, const float *g_posed_xyz, const float *g_posed_cov6,
                                           const float *g_colors, const float *g_opacity, float *g_xyz, float *g_log_scale,
                                           float *g_quat, float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts,
                                           mb_stream_t stream) {
    return pose_backward_impl(in, g_posed_xyz, g_posed_cov6, g_colors, g_opacity, g_xyz, g_log_scale, g_quat, g_opacity_logit, g_f_dc,
                              g_f_rest, g_skin_wts, 1, stream);
}


extern "C" int mb_sh_grad_from_views(const float *xyz, int32_t num_points, const float *skin_wts, int32_t num_skinned, int32_t num_bones,
                                     int32_t sh_degree, int32_t sh_coeffs, int32_t num_views, const float *bone_tf_all,
                                     const float *campos_all, const float *g_f_dc_all, int64_t view_stride, float *g_f_dc,
                                     float *g_f_rest, mb_stream_t stream) {
    MB_REQUIRE(num_points >= 0 && num_views >= 1 && num_skinned >= 0 && num_skinned <= num_points, "mb_sh_grad_from_views: bad counts");
    MB_REQUIRE(sh_degree >= 0 && sh_degree <= 3 && sh_coeffs >= (sh_degree + 1) * (sh_degree + 1) && sh_coeffs <= 16,
               "mb_sh_grad_from_views: bad SH degree / coefficient count");
    if (num_points == 0) return MB_OK;
    MB_REQUIRE(xyz && campos_all && g_f_dc_all && g_f_dc && (sh_coeffs == 1 || g_f_rest), "mb_sh_grad_from_views: null pointer");
    MB_REQUIRE(num_skinned == 0 || (skin_wts && bone_tf_all && num_bones > 0 && num_bones <= kMaxBones), "mb_sh_grad_from_views: skinning inputs missing");
    const int nbones = num_skinned > 0 ? num_bones : 0;
    ShViewsArgs a = {num_points, num_skinned, nbones, sh_coeffs, num_views,
                     view_stride ? view_stride : (int64_t)nbones * 16, view_stride ? view_stride : 3,
                     view_stride ? view_stride : (int64_t)num_points * 3,
                     xyz, skin_wts, bone_tf_all, campos_all, g_f_dc_all, g_f_dc, g_f_rest};
    const size_t work = (size_t)kShvTile * a.B + (size_t)kShvTile * kShvNzStride, outt = (size_t)kShvTile * (((a.K - 1) * 3) | 1);
    const size_t smem = sizeof(float) * ((size_t)a.R * a.B * 13 + (size_t)a.R * 3 + 1 + (work > outt ? work : outt));
    MB_REQUIRE(smem <= 200 * 1024, "mb_sh_grad_from_views: %d views x %d bones do not fit in shared memory", num_views, num_bones);
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = (num_points + kShvTile - 1) / kShvTile;
    KernelTimer kt("sh_grad_from_views", s);
#define MB_LAUNCH_SHV(D)                                                                                             \
    do {                                                                                                             \
        if (smem > 48 * 1024) MB_CUDA(cudaFuncSetAttribute(sh_grad_from_views_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        sh_grad_from_views_kernel<D><<<grid, kShvTile, smem, s>>>(a);                                                    \
    } while (0)
    switch (sh_degree) {
        case 0: MB_LAUNCH_SHV(0); break;
        case 1: MB_LAUNCH_SHV(1); break;
        case 2: MB_LAUNCH_SHV(2); break;
        default: MB_LAUNCH_SHV(3); break;
    }
#undef MB_LAUNCH_SHV
    return check_launch("sh_grad_from_views", false, s);
}
