// Fused photometric loss of the training step, forward AND gradient in one pass over the images:
//     loss = w_l1 * mean|pred - gt| + w_ssim * (1 - mean(ssim_map(pred, gt)))
// exactly as the reference evaluates it (src/utils/loss_utils.py:22-97 called from src/modules/base.py:323-365 on HWC
// tensors): `channel = img.size(-3)` is the image HEIGHT there, so the 11x11 Gaussian window of `ssim` slides over the
// (W, 3) plane of every image row -- an 11-tap filter along x and a 3x3 mixing of the colour channels
// (M[c][c'] = g[5 + c' - c]), zero padded, with no vertical extent.  Rows are therefore independent.
//
// One CTA owns 256 consecutive pixels of one row: it stages pred / gt with a 10-pixel halo in shared memory, filters the
// five moment images (mu1, mu2, E11, E22, E12) for 266 pixels, evaluates the SSIM map and its derivatives with respect to
// (mu1, E11, E12), filters those derivative maps again (the window is self-adjoint) and writes d loss / d pred for its
// 256 pixels.  HBM traffic: pred and gt read once (+8 % halo), the gradient written once: 36 B per pixel.  The reference
// runs 5 grouped convolutions forward and their adjoints backward (~30 full-image passes).
// Partial sums leave per CTA and are added in a fixed order by a second one-CTA kernel, so the loss is reproducible.
#include "common.cuh"

namespace mb {

constexpr int kLossChunk = 256;              // output pixels per CTA
constexpr int kLossHalo = 10;                // two 11-tap filters back to back
constexpr int kLossLoad = kLossChunk + 2 * kLossHalo;    // 276 staged pixels
constexpr int kLossMid = kLossChunk + kLossHalo;         // 266 pixels with an SSIM value
constexpr int kLossThreads = 288;

// the 11 taps as torch computes them in fp32 (loss_utils.py:38-45: exp(-(x-5)^2 / (2 * 1.5^2)) / sum)
__device__ __constant__ float kTaps[11] = {0x1.0d956cp-10f, 0x1.f1fe02p-8f, 0x1.26eb18p-5f, 0x1.bff0fep-4f, 0x1.b43c3ep-3f, 0x1.106560p-2f,
                                           0x1.b43c3ep-3f, 0x1.bff0fep-4f,  0x1.26eb18p-5f, 0x1.f1fe02p-8f, 0x1.0d956cp-10f};

// out[c] = sum_c' g[5 + c' - c] v[c']  (the colour axis of the window; symmetric)
__device__ __forceinline__ void mix3(const float *v, float *out) {
    const float g0 = kTaps[5], g1 = kTaps[6], g2 = kTaps[7];
    out[0] = g0 * v[0] + g1 * v[1] + g2 * v[2];
    out[1] = g1 * v[0] + g0 * v[1] + g1 * v[2];
    out[2] = g2 * v[0] + g1 * v[1] + g0 * v[2];
}

__global__ void __launch_bounds__(kLossThreads) photometric_loss_kernel(const float *__restrict__ pred, int64_t ps_y, int64_t ps_x,
                                                                        int64_t ps_c, const float *__restrict__ gt,
                                                                        int H, int W, float w_l1, float w_ssim,
                                                                        float *__restrict__ d_pred, double2 *__restrict__ partials) {
    __shared__ float sp[kLossLoad * 3], sg[kLossLoad * 3];      // staged rows, pixel-major (HWC)
    __shared__ float sd[kLossMid * 9];                          // derivative maps: d/dmu1, d/dE11, d/dE12 per channel
    __shared__ float red[2][kLossThreads / 32];
    const int tid = threadIdx.x, y = blockIdx.y;
    const int x0 = blockIdx.x * kLossChunk;                     // first output pixel of this CTA
    const float *prow = pred + (int64_t)y * ps_y, *grow = gt + (size_t)y * W * 3;
    // stage pixels [x0 - 10, x0 + 266) of the row; zero outside the image (conv2d zero padding).  gt is dense HWC; pred is
    // read with its own strides (the rasterizer's [3,H,W] output viewed as HWC: three coalesced plane reads)
    for (int e = tid; e < kLossLoad * 3; e += kLossThreads) {
        const int xe = x0 - kLossHalo + e / 3;
        const bool in = xe >= 0 && xe < W;
        const ptrdiff_t off = (ptrdiff_t)(x0 - kLossHalo) * 3 + e;
        sg[e] = in ? grow[off] : 0.f;
    }
    for (int e = tid; e < kLossLoad * 3; e += kLossThreads) {
        const int c = e / kLossLoad, j = e - c * kLossLoad, xe = x0 - kLossHalo + j;   // channel-major order: coalesced for planar pred
        const bool in = xe >= 0 && xe < W;
        sp[j * 3 + c] = in ? prow[(int64_t)xe * ps_x + (int64_t)c * ps_c] : 0.f;
    }
    __syncthreads();
    const float inv_n = 1.0f / ((float)H * (float)W * 3.0f);
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float my_l1 = 0.f, my_ss = 0.f;
    // stage A: pixel m of [0, 266) is image pixel x0 - 5 + m; its window covers staged pixels m .. m + 10
    for (int m = tid; m < kLossMid; m += kLossThreads) {
        const int x = x0 - 5 + m;
        float d[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (x >= 0 && x < W) {
            float a1[3] = {0.f, 0.f, 0.f}, a2[3] = {0.f, 0.f, 0.f}, a11[3] = {0.f, 0.f, 0.f}, a22[3] = {0.f, 0.f, 0.f}, a12[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const float w = kTaps[k];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float p = sp[(m + k) * 3 + c], q = sg[(m + k) * 3 + c];
                    a1[c] += w * p; a2[c] += w * q;
                    a11[c] += w * (p * p); a22[c] += w * (q * q); a12[c] += w * (p * q);
                }
            }
            float mu1[3], mu2[3], e11[3], e22[3], e12[3];
            mix3(a1, mu1); mix3(a2, mu2); mix3(a11, e11); mix3(a22, e22); mix3(a12, e12);
            const bool own = m >= 5 && m < 5 + kLossChunk;       // pixels whose loss terms this CTA counts
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float m1 = mu1[c], m2 = mu2[c];
                const float s1 = e11[c] - m1 * m1, s2 = e22[c] - m2 * m2, s12 = e12[c] - m1 * m2;
                const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2;
                const float B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
                const float den = B1 * B2, iden = 1.0f / den;
                const float ssim = A1 * A2 * iden;
                if (own) {
                    my_ss += ssim;
                    my_l1 += fabsf(sp[(m + 5) * 3 + c] - sg[(m + 5) * 3 + c]);
                }
                // d ssim / d mu1, d E11, d E12 (sigma1^2 = E11 - mu1^2, sigma12 = E12 - mu1 mu2), times d loss / d ssim
                const float up = -w_ssim * inv_n;
                const float dmu = ((2.f * m2 * A2 - 2.f * m2 * A1) * den - A1 * A2 * (2.f * m1 * B2 - 2.f * m1 * B1)) * (iden * iden);
                d[c] = up * dmu;
                d[3 + c] = up * (-(A1 * A2) * iden / B2);
                d[6 + c] = up * (2.f * A1 * iden);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) sd[m * 9 + k] = d[k];
    }
    __syncthreads();
    // stage B: output pixel t of [0, 256) is image pixel x0 + t; the adjoint window covers derivative pixels t .. t + 10
    for (int t = tid; t < kLossChunk; t += kLossThreads) {
        const int x = x0 + t;
        if (x >= W) continue;
        float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = kTaps[k];
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[j] += w * sd[(t + k) * 9 + j];
        }
        float g1[3], g11[3], g12[3];
        mix3(acc, g1); mix3(acc + 3, g11); mix3(acc + 6, g12);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float p = sp[(t + kLossHalo) * 3 + c], q = sg[(t + kLossHalo) * 3 + c];
            const float diff = p - q;
            const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            d_pred[((size_t)y * W + x) * 3 + c] = g1[c] + 2.f * p * g11[c] + q * g12[c] + w_l1 * inv_n * sgn;
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
