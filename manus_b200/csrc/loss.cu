// Fused photometric loss of the training step, forward AND gradient in one pass over the images:
//     loss = w_l1 * mean|pred - gt| + w_ssim * (1 - mean(ssim_map(pred, gt)))
// exactly as the reference evaluates it (src/utils/loss_utils.py:22-97 called from src/modules/base.py:323-365 on HWC
// tensors): `channel = img.size(-3)` is the image HEIGHT there, so the 11x11 Gaussian window of `ssim` slides over the
// (W, 3) plane of every image row -- an 11-tap filter along x and a 3x3 mixing of the colour channels
// (M[c][c'] = g[5 + c' - c]), zero padded, with no vertical extent.  Rows are therefore independent.
//
// One CTA owns 512 consecutive pixels of one row: it stages pred / gt with a 10-pixel halo in shared memory, filters the
// five moment images (mu1, mu2, E11, E22, E12) for 522 pixels, evaluates the SSIM map and its derivatives with respect to
// (mu1, E11, E12), filters those derivative maps again (the window is self-adjoint) and writes d loss / d pred for its
// 512 pixels.  HBM traffic: pred and gt read once (+4 % halo), the gradient written once: 36 B per pixel.  The reference
// runs 5 grouped convolutions forward and their adjoints backward (~30 full-image passes).
// The kernel is bound by fp32 issue and shared-memory reads, not by HBM: every thread filters TWO adjacent pixels (each
// staged value is loaded once for both), on (pred, gt) pairs with packed FFMA2.
// Partial sums leave per CTA and are added in a fixed order by a second one-CTA kernel, so the loss is reproducible.
#include "common.cuh"

namespace mb {

constexpr int kLossChunk = 512;              // output pixels per CTA
constexpr int kLossHalo = 10;                // two 11-tap filters back to back
constexpr int kLossLoad = kLossChunk + 2 * kLossHalo;    // 532 staged pixels
constexpr int kLossMid = kLossChunk + kLossHalo;         // 522 pixels with an SSIM value
constexpr int kLossThreads = 288;            // two adjacent pixels per thread: 261 threads in stage A, 256 in stage B
constexpr int kLoadH = kLossLoad / 2, kMidH = kLossMid / 2;

// the 11 taps as torch computes them in fp32 (loss_utils.py:38-45: exp(-(x-5)^2 / (2 * 1.5^2)) / sum)
__device__ __constant__ float kTaps[11] = {0x1.0d956cp-10f, 0x1.f1fe02p-8f, 0x1.26eb18p-5f, 0x1.bff0fep-4f, 0x1.b43c3ep-3f, 0x1.106560p-2f,
                                           0x1.b43c3ep-3f, 0x1.bff0fep-4f,  0x1.26eb18p-5f, 0x1.f1fe02p-8f, 0x1.0d956cp-10f};

// two fp32 FMAs per instruction (FFMA2 on sm_100): the filters run on (pred, gt) pairs
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// out[c] = sum_c' g[5 + c' - c] v[c']  (the colour axis of the window; symmetric), on pairs
__device__ __forceinline__ void mix3(const float2 *v, float2 *out) {
    const float2 g0 = make_float2(kTaps[5], kTaps[5]), g1 = make_float2(kTaps[6], kTaps[6]), g2 = make_float2(kTaps[7], kTaps[7]);
    out[0] = fma2(g2, v[2], fma2(g1, v[1], mul2(g0, v[0])));
    out[1] = fma2(g1, v[2], fma2(g0, v[1], mul2(g1, v[0])));
    out[2] = fma2(g0, v[2], fma2(g1, v[1], mul2(g2, v[0])));
}
__device__ __forceinline__ void mix3(const float *v, float *out) {
    const float g0 = kTaps[5], g1 = kTaps[6], g2 = kTaps[7];
    out[0] = g0 * v[0] + g1 * v[1] + g2 * v[2];
    out[1] = g1 * v[0] + g0 * v[1] + g1 * v[2];
    out[2] = g2 * v[0] + g1 * v[1] + g0 * v[2];
}

__device__ __forceinline__ float fast_rcp_loss(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Shared-memory images are split into even and odd pixels so that a thread working on pixels (2t, 2t + 1) reads, at every
// tap, an address that is consecutive across the warp.
__global__ void __launch_bounds__(kLossThreads) photometric_loss_kernel(const float *__restrict__ pred, int64_t ps_y, int64_t ps_x,
                                                                        int64_t ps_c, const float *__restrict__ gt,
                                                                        int H, int W, float w_l1, float w_ssim,
                                                                        float *__restrict__ d_pred, double2 *__restrict__ partials) {
    __shared__ float2 spq[2][3][kLoadH];        // [parity][channel][pixel / 2] = (pred, gt), zero outside the row
    __shared__ float2 sdA[2][3][kMidH];         // derivative maps (d/dmu1, d/dE11) per channel
    __shared__ float sdB[2][3][kMidH];          // d/dE12 per channel
    __shared__ float red[2][kLossThreads / 32];
    const int tid = threadIdx.x, y = blockIdx.y;
    const int x0 = blockIdx.x * kLossChunk;                     // first output pixel of this CTA
    const float *prow = pred + (int64_t)y * ps_y, *grow = gt + (size_t)y * W * 3;
    // stage pixels [x0 - 10, x0 + 522) of the row; gt is dense HWC, pred is read with its own strides (the rasterizer's
    // [3,H,W] output viewed as HWC: three coalesced plane reads)
    // all global loads of a thread are issued before the first shared-memory store (the staging is latency bound otherwise)
    constexpr int kIter = (kLossLoad * 3 + kLossThreads - 1) / kLossThreads;
    float vg[kIter], vp[kIter];
#pragma unroll
    for (int n = 0; n < kIter; ++n) {
        const int e = tid + n * kLossThreads;
        const int j = e / 3, xe = x0 - kLossHalo + j;
        vg[n] = (e < kLossLoad * 3 && xe >= 0 && xe < W) ? grow[(ptrdiff_t)(x0 - kLossHalo) * 3 + e] : 0.f;
        const int c2 = e / kLossLoad, j2 = e - c2 * kLossLoad, xe2 = x0 - kLossHalo + j2;   // channel-major order: coalesced for planar pred
        vp[n] = (e < kLossLoad * 3 && xe2 >= 0 && xe2 < W) ? prow[(int64_t)xe2 * ps_x + (int64_t)c2 * ps_c] : 0.f;
    }
#pragma unroll
    for (int n = 0; n < kIter; ++n) {
        const int e = tid + n * kLossThreads;
        if (e < kLossLoad * 3) {
            const int j = e / 3, c = e - 3 * j;
            spq[j & 1][c][j >> 1].y = vg[n];
            const int c2 = e / kLossLoad, j2 = e - c2 * kLossLoad;
            spq[j2 & 1][c2][j2 >> 1].x = vp[n];
        }
    }
    __syncthreads();
    const float inv_n = 1.0f / ((float)H * (float)W * 3.0f);
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float up = -w_ssim * inv_n;                           // d loss / d ssim_map
    float my_l1 = 0.f, my_ss = 0.f;
    // stage A: thread t owns mid pixels m = 2t, 2t + 1 (image pixel x0 - 5 + m); window pixel k of m = staged pixel m + k
    if (tid < kMidH) {
        float2 amu[2][3], aee[2][3];                            // (mu1, mu2) and (E11, E22) sums of the two pixels
        float a12[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            amu[0][c] = amu[1][c] = aee[0][c] = aee[1][c] = make_float2(0.f, 0.f);
            a12[0][c] = a12[1][c] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) {                          // staged pixel 2t + k serves tap k of pixel 2t and tap k - 1 of pixel 2t + 1
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 pq = spq[k & 1][c][tid + (k >> 1)];
                const float2 sq = mul2(pq, pq);
                const float x12 = pq.x * pq.y;
                if (k < 11) {
                    const float2 w = make_float2(kTaps[k], kTaps[k]);
                    amu[0][c] = fma2(w, pq, amu[0][c]); aee[0][c] = fma2(w, sq, aee[0][c]); a12[0][c] += w.x * x12;
                }
                if (k > 0) {
                    const float2 w = make_float2(kTaps[k - 1], kTaps[k - 1]);
                    amu[1][c] = fma2(w, pq, amu[1][c]); aee[1][c] = fma2(w, sq, aee[1][c]); a12[1][c] += w.x * x12;
                }
            }
        }
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            const int m = 2 * tid + px, x = x0 - 5 + m;
            float2 mu[3], ee[3];
            float e12[3];
            mix3(amu[px], mu); mix3(aee[px], ee); mix3(a12[px], e12);
            const bool in = x >= 0 && x < W;
            const bool own = in && m >= 5 && m < 5 + kLossChunk;   // pixels whose loss terms this CTA counts
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float m1 = mu[c].x, m2 = mu[c].y;
                const float s1 = ee[c].x - m1 * m1, s2 = ee[c].y - m2 * m2, s12 = e12[c] - m1 * m2;
                const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2;
                const float B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
                const float iden = fast_rcp_loss(B1 * B2);
                const float t = A1 * A2 * iden;                 // ssim
                if (own) {
                    my_ss += t;
                    const float2 pq = spq[(m + 5) & 1][c][(m + 5) >> 1];
                    my_l1 += fabsf(pq.x - pq.y);
                }
                // d ssim / d mu1, d E11, d E12 (sigma1^2 = E11 - mu1^2, sigma12 = E12 - mu1 mu2), times d loss / d ssim
                const float dmu = (2.f * m2 * (A2 - A1)) * iden - t * (2.f * m1 * (B2 - B1)) * iden;
                const float u = in ? up : 0.f;                  // no output pixel outside the row: its derivative maps are zero
                sdA[m & 1][c][m >> 1] = make_float2(u * dmu, u * (-t * fast_rcp_loss(B2)));
                sdB[m & 1][c][m >> 1] = u * (2.f * A1 * iden);
            }
        }
    }
    __syncthreads();
    // stage B: thread t owns output pixels 2t, 2t + 1 (image pixel x0 + o); the adjoint window covers derivative pixels o .. o + 10
    if (tid < kLossChunk / 2) {
        float2 bA[2][3];
        float bB[2][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { bA[0][c] = bA[1][c] = make_float2(0.f, 0.f); bB[0][c] = bB[1][c] = 0.f; }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 dA = sdA[k & 1][c][tid + (k >> 1)];
                const float dB = sdB[k & 1][c][tid + (k >> 1)];
                if (k < 11) { const float2 w = make_float2(kTaps[k], kTaps[k]); bA[0][c] = fma2(w, dA, bA[0][c]); bB[0][c] += w.x * dB; }
                if (k > 0) { const float2 w = make_float2(kTaps[k - 1], kTaps[k - 1]); bA[1][c] = fma2(w, dA, bA[1][c]); bB[1][c] += w.x * dB; }
            }
        }
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            const int o = 2 * tid + px, x = x0 + o;
            if (x >= W) continue;
            float2 gA[3];
            float gB[3];
            mix3(bA[px], gA); mix3(bB[px], gB);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 pq = spq[(o + kLossHalo) & 1][c][(o + kLossHalo) >> 1];
                const float diff = pq.x - pq.y;
                const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                d_pred[((size_t)y * W + x) * 3 + c] = gA[c].x + 2.f * pq.x * gA[c].y + pq.y * gB[c] + w_l1 * inv_n * sgn;
            }
        }
    }
    // per-CTA sums of |pred - gt| and ssim
    const int lane = tid & 31, warp = tid >> 5;
    my_l1 = warp_sum(my_l1);
    my_ss = warp_sum(my_ss);
    if (lane == 0) { red[0][warp] = my_l1; red[1][warp] = my_ss; }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0.0, s1 = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) { s0 += red[0][w]; s1 += red[1][w]; }
        partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = make_double2(s0, s1);
    }
}

// out[0] = loss, out[1] = mean |pred - gt|, out[2] = mean ssim; fixed summation order
__global__ void __launch_bounds__(1024) photometric_loss_finalize_kernel(const double2 *__restrict__ partials, int n, double inv_n,
                                                                         float w_l1, float w_ssim, float *__restrict__ out) {
    __shared__ double r0[32], r1[32];
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) { s0 += partials[i].x; s1 += partials[i].y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s0; r1[threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < 32; ++w) { t0 += r0[w]; t1 += r1[w]; }
        const double l1 = t0 * inv_n, ss = t1 * inv_n;
        out[0] = (float)(w_l1 * l1 + w_ssim * (1.0 - ss));
        out[1] = (float)l1;
        out[2] = (float)ss;
    }
}

}  // namespace mb

using namespace mb;

extern "C" size_t mb_photometric_loss_workspace_bytes(int32_t height, int32_t width) {
    const size_t chunks = (size_t)((width + kLossChunk - 1) / kLossChunk) * (size_t)(height > 0 ? height : 1);
    return align_up(chunks * sizeof(double2));
}

extern "C" int mb_photometric_loss(const float *pred, int64_t pred_stride_y, int64_t pred_stride_x, int64_t pred_stride_c,
                                   const float *gt, int32_t height, int32_t width, float w_l1, float w_ssim,
                                   float *loss_out, float *d_pred, void *workspace, size_t workspace_bytes, mb_stream_t stream) {
    MB_REQUIRE(height > 0 && width > 0, "mb_photometric_loss: bad image size %d x %d", width, height);
    MB_REQUIRE(pred && gt && loss_out && d_pred && workspace, "mb_photometric_loss: null pointer");
    if (workspace_bytes < mb_photometric_loss_workspace_bytes(height, width)) {
        set_error("mb_photometric_loss: workspace too small");
        return MB_ERR_WORKSPACE;
    }
    MB_REQUIRE(height <= 65535, "mb_photometric_loss: image height %d exceeds the grid limit", height);
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 grid((width + kLossChunk - 1) / kLossChunk, height);
    double2 *partials = reinterpret_cast<double2 *>(workspace);
    {
        KernelTimer kt("photometric_loss", s);
        photometric_loss_kernel<<<grid, kLossThreads, 0, s>>>(pred, pred_stride_y, pred_stride_x, pred_stride_c, gt, height, width, w_l1, w_ssim,
                                                              d_pred, partials);
    }
    int rc = check_launch("photometric_loss", false, s);
    if (rc) return rc;
    photometric_loss_finalize_kernel<<<1, 1024, 0, s>>>(partials, (int)(grid.x * grid.y), 1.0 / ((double)height * width * 3.0), w_l1, w_ssim,
                                                        loss_out);
    return check_launch("photometric_loss_finalize", false, s);
}
