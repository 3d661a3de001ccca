// Per-frame skin-weight lookup: trilinear interpolation of a channel-last voxel grid of bone weights at every Gaussian,
// followed by row normalisation -- forward and backward.
//
// Replaces skinning_weights_from_voxel_grid (src/utils/gaussian_utils.py:167-196: torch grid_sample on a [1,C,D,H,W] view of
// the [D,H,W,C] grid, align_corners=True, zero padding, then w / w.sum(-1)), which HandGaussianModel.get_skin_weights calls
// every training step (src/models/hand_gaussian.py:65-76) directly in front of the LBS step.
//
// The grid is read as it is stored ([D,H,W,C], C = 21 bones contiguous): every one of the 8 corners of a cell is one
// contiguous 4C-byte read (84 B at C = 21).  672 B gathered + 4C B written per Gaussian; the backward re-gathers the
// corners and reduces the coordinate gradient over the channels with warp shuffles.
// The gradient w.r.t. the grid (only needed when the weights themselves are optimised) is scattered with RED.ADD.
#include "common.cuh"

namespace mb {

constexpr int kSkinWarps = 8;   // Gaussians per CTA

struct SkinArgs {
    int N, D, H, W, C;
    const float *xyz, *grid, *center, *scale;
    float *out;                 // forward: skin_wts [N,C]
    const float *g_out;         // backward: dL/dskin_wts [N,C]
    float *g_xyz, *g_grid;      // backward outputs ([N,3]; [D,H,W,C] or null, accumulated)
};

// grid_sample's un-normalisation with align_corners=True: ((coord + 1) / 2) * (size - 1), coord = (xyz - centre) / scale.
// Returns the element offset of the cell's base corner, an 8-bit mask of the corners that lie inside the grid (zero
// padding) and the fractional position.
struct Cell {
    int off;            // ((z * H + y) * W + x) * C of the base corner (may be negative for cells that straddle the border)
    unsigned mask;      // bit k set: corner k = (dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2) is inside the grid
    float fx, fy, fz;
};

__device__ __forceinline__ Cell locate(const SkinArgs &a, float x, float y, float z) {
    const float nx = (x - a.center[0]) / a.scale[0], ny = (y - a.center[1]) / a.scale[1], nz = (z - a.center[2]) / a.scale[2];
    const float px = ((nx + 1.f) / 2.f) * (float)(a.W - 1), py = ((ny + 1.f) / 2.f) * (float)(a.H - 1), pz = ((nz + 1.f) / 2.f) * (float)(a.D - 1);
    const float bx = floorf(px), by = floorf(py), bz = floorf(pz);
    // far outside the grid every corner is out of range anyway: clamp before the int conversion
    const int ix = (int)fminf(fmaxf(bx, -2.f), (float)a.W), iy = (int)fminf(fmaxf(by, -2.f), (float)a.H), iz = (int)fminf(fmaxf(bz, -2.f), (float)a.D);
    Cell c;
    c.fx = px - bx; c.fy = py - by; c.fz = pz - bz;
    c.off = ((iz * a.H + iy) * a.W + ix) * a.C;
    const unsigned mx = (ix >= 0 && ix < a.W ? 0x55u : 0u) | (ix + 1 >= 0 && ix + 1 < a.W ? 0xaau : 0u);
    const unsigned my = (iy >= 0 && iy < a.H ? 0x33u : 0u) | (iy + 1 >= 0 && iy + 1 < a.H ? 0xccu : 0u);
    const unsigned mz = (iz >= 0 && iz < a.D ? 0x0fu : 0u) | (iz + 1 >= 0 && iz + 1 < a.D ? 0xf0u : 0u);
    c.mask = mx & my & mz;
    return c;
}

// A warp owns 32 consecutive Gaussians: lane l locates Gaussian base + l (coordinates -> cell, once per Gaussian instead of
// once per lane), then the warp walks its Gaussians with lane c owning channels c, c + 32, ... (kR per lane); the cell is
// broadcast with 5 shuffles and the 8 corner rows are contiguous 4C-byte reads.  The walk is unrolled by 4 so that 32
// corner reads are in flight per warp (the lookup is latency bound otherwise).
template <bool kBackward, int kR>
__global__ void __launch_bounds__(kSkinWarps * 32) skin_weights_kernel(SkinArgs a) {
    const int lane = threadIdx.x & 31;
    const int base = (blockIdx.x * kSkinWarps + (threadIdx.x >> 5)) * 32;
    if (base >= a.N) return;
    Cell mine = {0, 0u, 0.f, 0.f, 0.f};
    if (base + lane < a.N) mine = locate(a, a.xyz[3 * (size_t)(base + lane)], a.xyz[3 * (size_t)(base + lane) + 1], a.xyz[3 * (size_t)(base + lane) + 2]);
    const int cnt = min(32, a.N - base);
    const int sy = a.W * a.C, sz = a.H * a.W * a.C;
    const float kx = ((float)(a.W - 1) / 2.f) / a.scale[0], ky = ((float)(a.H - 1) / 2.f) / a.scale[1], kz = ((float)(a.D - 1) / 2.f) / a.scale[2];
    float my_gx = 0.f, my_gy = 0.f, my_gz = 0.f;     // backward: coordinate gradient of Gaussian base + lane
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
        const int off = __shfl_sync(0xffffffffu, mine.off, j);
        const unsigned mask = __shfl_sync(0xffffffffu, mine.mask, j);
        const float fx = __shfl_sync(0xffffffffu, mine.fx, j), fy = __shfl_sync(0xffffffffu, mine.fy, j), fz = __shfl_sync(0xffffffffu, mine.fz, j);
        const size_t i = (size_t)(base + j);
        float v[kR][8], go[kR];
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const int c = lane + 32 * r;
            const float *p = a.grid + off + c;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                v[r][k] = (c < a.C && ((mask >> k) & 1u)) ? p[(k & 1) * a.C + ((k >> 1) & 1) * sy + (k >> 2) * sz] : 0.f;
            if (kBackward) go[r] = c < a.C ? a.g_out[i * a.C + c] : 0.f;
        }
        const float ax[2] = {1.f - fx, fx}, ay[2] = {1.f - fy, fy}, az[2] = {1.f - fz, fz};
        float raw[kR], tot = 0.f;
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            raw[r] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[r] += (ax[k & 1] * ay[(k >> 1) & 1] * az[k >> 2]) * v[r][k];
            tot += raw[r];
        }
        const float S = warp_sum(tot);
        if (!kBackward) {
#pragma unroll
            for (int r = 0; r < kR; ++r) {
                const int c = lane + 32 * r;
                if (c < a.C) a.out[i * a.C + c] = raw[r] / S;
            }
            continue;
        }
        // backward: through the normalisation, then through the trilinear weights
        float dotp = 0.f;
#pragma unroll
        for (int r = 0; r < kR; ++r) dotp += go[r] * (raw[r] / S);
        dotp = warp_sum(dotp);
        float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            const int c = lane + 32 * r;
            const float g_raw = c < a.C ? (go[r] - dotp) / S : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t = v[r][k] * g_raw;
                gx += ((k & 1) ? t : -t) * (ay[(k >> 1) & 1] * az[k >> 2]);
                gy += (((k >> 1) & 1) ? t : -t) * (ax[k & 1] * az[k >> 2]);
                gz += ((k >> 2) ? t : -t) * (ax[k & 1] * ay[(k >> 1) & 1]);
                if (a.g_grid && c < a.C && ((mask >> k) & 1u))
                    red_add(a.g_grid + off + c + (k & 1) * a.C + ((k >> 1) & 1) * sy + (k >> 2) * sz, (ax[k & 1] * ay[(k >> 1) & 1] * az[k >> 2]) * g_raw);
            }
        }
        gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
        if (lane == j) { my_gx = gx; my_gy = gy; my_gz = gz; }
    }
    if (kBackward && lane < cnt) {
        a.g_xyz[3 * (size_t)(base + lane)] = my_gx * kx;
        a.g_xyz[3 * (size_t)(base + lane) + 1] = my_gy * ky;
        a.g_xyz[3 * (size_t)(base + lane) + 2] = my_gz * kz;
    }
}

template <bool kBackward>
static void launch_skin(const SkinArgs &a, cudaStream_t s) {
    const int warps = (a.N + 31) / 32;
    const int grid = (warps + kSkinWarps - 1) / kSkinWarps;
    if (a.C <= 32) skin_weights_kernel<kBackward, 1><<<grid, kSkinWarps * 32, 0, s>>>(a);
    else skin_weights_kernel<kBackward, 2><<<grid, kSkinWarps * 32, 0, s>>>(a);
}

static int validate_skin(const SkinArgs &a, const char *who) {
    MB_REQUIRE(a.N >= 0 && a.D > 0 && a.H > 0 && a.W > 0 && a.C > 0 && a.C <= 64, "%s: bad sizes N=%d grid=%dx%dx%dx%d (C <= 64)", who, a.N,
               a.D, a.H, a.W, a.C);
    MB_REQUIRE(a.N == 0 || (a.xyz && a.grid && a.center && a.scale), "%s: null input", who);
    MB_REQUIRE((int64_t)a.D * a.H * a.W * a.C < ((int64_t)1 << 31), "%s: grid has more than 2^31 elements", who);
    return MB_OK;
}

}  // namespace mb

using namespace mb;

extern "C" int mb_skin_weights_forward(const float *xyz, int32_t num_points, const float *grid_weights, int32_t depth, int32_t height,
                                       int32_t width, int32_t channels, const float *grid_center, const float *grid_scale,
                                       float *skin_wts, mb_stream_t stream) {
    SkinArgs a = {num_points, depth, height, width, channels, xyz, grid_weights, grid_center, grid_scale, skin_wts, nullptr, nullptr, nullptr};
    int rc = validate_skin(a, "mb_skin_weights_forward");
    if (rc) return rc;
    if (num_points == 0) return MB_OK;
    MB_REQUIRE(skin_wts != nullptr, "mb_skin_weights_forward: null output");
    cudaStream_t s = (cudaStream_t)stream;
    KernelTimer kt("skin_weights_forward", s);
    launch_skin<false>(a, s);
    return check_launch("skin_weights_forward", false, s);
}

extern "C" int mb_skin_weights_backward(const float *xyz, int32_t num_points, const float *grid_weights, int32_t depth, int32_t height,
                                        int32_t width, int32_t channels, const float *grid_center, const float *grid_scale,
                                        const float *g_skin_wts, float *g_xyz, float *g_grid_weights, mb_stream_t stream) {
    SkinArgs a = {num_points, depth, height, width, channels, xyz, grid_weights, grid_center, grid_scale, nullptr, g_skin_wts, g_xyz, g_grid_weights};
    int rc = validate_skin(a, "mb_skin_weights_backward");
    if (rc) return rc;
    if (num_points == 0) return MB_OK;
    MB_REQUIRE(g_skin_wts && g_xyz, "mb_skin_weights_backward: null gradient pointer");
    cudaStream_t s = (cudaStream_t)stream;
    KernelTimer kt("skin_weights_backward", s);
    launch_skin<true>(a, s);
    return check_launch("skin_weights_backward", false, s);
}
