// Per-tile alpha blending, forward and backward (A.4), plus the per-Gaussian backward of the
// projection.  One CTA of 256 threads owns one 16x16 tile; its depth-sorted instance list is a contiguous run of 48-B
// records that is streamed through shared memory by the TMA engine (cp.async.bulk + mbarrier, double buffered) while
// the threads blend the previous batch.  A warp owns an 8x4 pixel block.  Arithmetic order per pixel is the reference
// order (sequential front-to-back product), so results do not depend on the schedule.
#include "raster_state.cuh"

namespace mb {

constexpr int kBatch = 256;                 // records per pipeline stage (12 KB)
constexpr int kAccStride = 12;              // floats per Gaussian in the gradient accumulator
// accumulator slots: 0,1 mean2D.xy | 2,3,4 conic (x,y,w) | 5 opacity | 6,7,8 colour

int build_instances(const mb_raster_inputs *in, const RasterDims &d, const GeomState &g, const BinningState &b,
                    const ImageState &im, int64_t capacity, cudaStream_t s);

__device__ __forceinline__ void pixel_of_thread(int tile, int gx, int tid, int &px, int &py) {
    const int warp = tid >> 5, lane = tid & 31;
    px = (tile % gx) * kTile + (warp & 1) * 8 + (lane & 7);
    py = (tile / gx) * kTile + (warp >> 1) * 4 + (lane >> 3);
}

__global__ void __launch_bounds__(256) blend_forward_kernel(const Record *__restrict__ records,
                                                            const uint2 *__restrict__ ranges, int W, int H, int gx,
                                                            const float *__restrict__ bg, float *__restrict__ out_color,
                                                            float *__restrict__ final_T, uint32_t *__restrict__ n_contrib) {
    __shared__ __align__(128) Record stage[2][kBatch];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const uint2 range = ranges[tile];
    const int len = (int)(range.y - range.x);
    const int nb = (len + kBatch - 1) / kBatch;
    int px, py;
    pixel_of_thread(tile, gx, tid, px, py);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const Record *src = records + range.x;
    if (tid == 0 && nb > 0) {
        const uint32_t bytes = (uint32_t)min(kBatch, len) * (uint32_t)sizeof(Record);
        mbar_expect_tx(&bar[0], bytes);
        bulk_g2s(&stage[0][0], src, bytes, &bar[0]);
    }

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last = 0;
    bool done = !inside;
    for (int b = 0; b < nb; ++b) {
        const int cur = b & 1;
        if (tid == 0 && b + 1 < nb) {   // prefetch the next batch into the buffer released at the end of iteration b-1
            const uint32_t bytes = (uint32_t)min(kBatch, len - (b + 1) * kBatch) * (uint32_t)sizeof(Record);
            mbar_expect_tx(&bar[cur ^ 1], bytes);
            bulk_g2s(&stage[cur ^ 1][0], src + (size_t)(b + 1) * kBatch, bytes, &bar[cur ^ 1]);
        }
        mbar_wait(&bar[cur], (uint32_t)((b >> 1) & 1));
        const int cnt = min(kBatch, len - b * kBatch);
        if (!done) {
            const float4 *st = reinterpret_cast<const float4 *>(&stage[cur][0]);
            for (int j = 0; j < cnt; ++j) {
                const float4 ra = st[3 * j], rb = st[3 * j + 1];
                ++contributor;
                const float dx = ra.x - fx, dy = ra.y - fy;
                const float power = -0.5f * (ra.z * dx * dx + rb.x * dy * dy) - ra.w * dx * dy;
                if (power > 0.0f) continue;
                const float alpha = fminf(kAlphaMax, rb.y * expf(power));
                if (alpha < kAlphaMin) continue;
                const float test_T = T * (1.0f - alpha);
                if (test_T < kTMin) {
                    done = true;
                    break;
                }
                const float w = alpha * T;
                C0 += rb.z * w;
                C1 += rb.w * w;
                C2 += st[3 * j + 2].x * w;
                T = test_T;
                last = contributor;
            }
        }
        // all reads of stage[cur] are complete after this barrier; it also counts finished pixels (early exit)
        const int ndone = __syncthreads_count(done);
        if (ndone == 256) {
            if (tid == 0 && b + 1 < nb) mbar_wait(&bar[cur ^ 1], (uint32_t)(((b + 1) >> 1) & 1));   // drain in-flight copy
            break;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = C0 + T * bg[0];
        out_color[plane + pix] = C1 + T * bg[1];
        out_color[2 * plane + pix] = C2 + T * bg[2];
    }
}

__global__ void __launch_bounds__(256) blend_backward_kernel(const Record *__restrict__ records,
                                                             const uint2 *__restrict__ ranges, int W, int H, int gx,
                                                             const float *__restrict__ bg, const float *__restrict__ final_T,
                                                             const uint32_t *__restrict__ n_contrib,
                                                             const float *__restrict__ dL_dout, int64_t sc, int64_t sy,
                                                             int64_t sx, float *__restrict__ acc) {
    __shared__ __align__(128) Record stage[2][kBatch];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t warp_max[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const uint2 range = ranges[tile];
    if (range.y <= range.x) return;
    int px, py;
    pixel_of_thread(tile, gx, tid, px, py);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const size_t pix = (size_t)py * W + px;
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t last = inside ? n_contrib[pix] : 0u;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) {
        const float *gp = dL_dout + (int64_t)py * sy + (int64_t)px * sx;
        dp0 = gp[0];
        dp1 = gp[sc];
        dp2 = gp[2 * sc];
    }
    const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;

    // only instances in front of the deepest last-contributor of the tile can matter
    uint32_t m = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) warp_max[warp] = m;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = max(m, warp_max[w]);
    const int len = (int)m;                       // process list positions [0, len) back to front
    const int nb = (len + kBatch - 1) / kBatch;
    if (nb == 0) return;
    const Record *src = records + range.x;
    auto issue = [&](int b, int buf) {
        const uint32_t bytes = (uint32_t)min(kBatch, len - b * kBatch) * (uint32_t)sizeof(Record);
        mbar_expect_tx(&bar[buf], bytes);
        bulk_g2s(&stage[buf][0], src + (size_t)b * kBatch, bytes, &bar[buf]);
    };
    if (tid == 0) issue(nb - 1, 0);

    float T = T_final, a0 = 0.f, a1 = 0.f, a2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    int it = 0;
    for (int b = nb - 1; b >= 0; --b, ++it) {
        const int cur = it & 1;
        if (tid == 0 && b > 0) issue(b - 1, cur ^ 1);
        mbar_wait(&bar[cur], (uint32_t)((it >> 1) & 1));
        const int cnt = min(kBatch, len - b * kBatch);
        const float4 *st = reinterpret_cast<const float4 *>(&stage[cur][0]);
        for (int j = cnt - 1; j >= 0; --j) {
            const uint32_t pos = (uint32_t)(b * kBatch + j);
            const float4 ra = st[3 * j], rb = st[3 * j + 1];
            const float dx = ra.x - fx, dy = ra.y - fy;
            const float power = -0.5f * (ra.z * dx * dx + rb.x * dy * dy) - ra.w * dx * dy;
            const float G = expf(power);
            const float alpha = fminf(kAlphaMax, rb.y * G);
            const bool active = (pos < last) && (power <= 0.0f) && (alpha >= kAlphaMin);
            if (!__any_sync(0xffffffffu, active)) continue;
            const float4 rc = st[3 * j + 2];
            float g_m2x = 0.f, g_m2y = 0.f, g_cx = 0.f, g_cy = 0.f, g_cw = 0.f, g_op = 0.f, g_c0 = 0.f, g_c1 = 0.f, g_c2 = 0.f;
            if (active) {
                T = T / (1.0f - alpha);
                const float dch = alpha * T;
                float dL_dalpha;
                a0 = last_alpha * lc0 + (1.f - last_alpha) * a0;
                a1 = last_alpha * lc1 + (1.f - last_alpha) * a1;
                a2 = last_alpha * lc2 + (1.f - last_alpha) * a2;
                lc0 = rb.z; lc1 = rb.w; lc2 = rc.x;
                dL_dalpha = (lc0 - a0) * dp0 + (lc1 - a1) * dp1 + (lc2 - a2) * dp2;
                g_c0 = dch * dp0; g_c1 = dch * dp1; g_c2 = dch * dp2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = rb.y * dL_dalpha;   // the 0.99 clamp is not masked (upstream behaviour)
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * ra.z - gdy * ra.w;
                const float dG_ddely = -gdy * rb.x - gdx * ra.w;
                g_m2x = dL_dG * dG_ddelx * ddelx_dx;
                g_m2y = dL_dG * dG_ddely * ddely_dy;
                g_cx = -0.5f * gdx * dx * dL_dG;
                g_cy = -0.5f * gdx * dy * dL_dG;
                g_cw = -0.5f * gdy * dy * dL_dG;
                g_op = G * dL_dalpha;
            }
            g_m2x = warp_sum(g_m2x); g_m2y = warp_sum(g_m2y);
            g_cx = warp_sum(g_cx); g_cy = warp_sum(g_cy); g_cw = warp_sum(g_cw);
            g_op = warp_sum(g_op);
            g_c0 = warp_sum(g_c0); g_c1 = warp_sum(g_c1); g_c2 = warp_sum(g_c2);
            if (lane == 0) {
                float *dst = acc + (size_t)__float_as_uint(rc.y) * kAccStride;
                red_add(dst + 0, g_m2x); red_add(dst + 1, g_m2y);
                red_add(dst + 2, g_cx); red_add(dst + 3, g_cy); red_add(dst + 4, g_cw);
                red_add(dst + 5, g_op);
                red_add(dst + 6, g_c0); red_add(dst + 7, g_c1); red_add(dst + 8, g_c2);
            }
        }
        __syncthreads();   // stage[cur] may be overwritten by the copy issued in the next iteration
    }
}

struct PreBwdArgs {
    int P, W, H, deg, M;
    float tanx, tany, focx, focy, scale_mod;
    const float *means3D, *cov3D, *scales, *rots, *shs, *view, *proj, *campos;
    const int32_t *radii;
    const uint32_t *clamped;
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
    const float *v = cam, *p = cam + 16;
    const int i = blockIdx.x * 256 + tid;
    if (i >= a.P) return;
    float gmean[3] = {0.f, 0.f, 0.f}, gcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float g2x = 0.f, g2y = 0.f, gop = 0.f, gc[3] = {0.f, 0.f, 0.f};
    const bool vis = a.radii[i] > 0;
    const float mx = a.means3D[3 * i], my = a.means3D[3 * i + 1], mz = a.means3D[3 * i + 2];
    if (vis) {
        const float *ac = a.acc + (size_t)i * kAccStride;
        g2x = ac[0]; g2y = ac[1];
        const float gx = ac[2], gy = ac[3], gz = ac[4];
        gop = ac[5]; gc[0] = ac[6]; gc[1] = ac[7]; gc[2] = ac[8];
        const float *c6 = a.cov3D + 6 * (size_t)i;
        // cov2D backward (A.4)
        const float t0 = v[0] * mx + v[4] * my + v[8] * mz + v[12];
        const float t1 = v[1] * mx + v[5] * my + v[9] * mz + v[13];
        const float tz = v[2] * mx + v[6] * my + v[10] * mz + v[14];
        const float limx = 1.3f * a.tanx, limy = 1.3f * a.tany;
        const float rx = t0 / tz, ry = t1 / tz;
        const float xm = (rx < -limx || rx > limx) ? 0.f : 1.f, ym = (ry < -limy || ry > limy) ? 0.f : 1.f;
        const float tx = fminf(limx, fmaxf(-limx, rx)) * tz, ty = fminf(limy, fmaxf(-limy, ry)) * tz;
        const float J00 = a.focx / tz, J02 = -(a.focx * tx) / (tz * tz);
        const float J11 = a.focy / tz, J12 = -(a.focy * ty) / (tz * tz);
        float M0[3], M1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            M0[k] = J00 * v[4 * k + 0] + J02 * v[4 * k + 2];
            M1[k] = J11 * v[4 * k + 1] + J12 * v[4 * k + 2];
        }
        const float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
        float SM0[3], SM1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            SM0[k] = S[3 * k] * M0[0] + S[3 * k + 1] * M0[1] + S[3 * k + 2] * M0[2];
            SM1[k] = S[3 * k] * M1[0] + S[3 * k + 1] * M1[1] + S[3 * k + 2] * M1[2];
        }
        const float ca = M0[0] * SM0[0] + M0[1] * SM0[1] + M0[2] * SM0[2] + kLowPass;
        const float cb = M0[0] * SM1[0] + M0[1] * SM1[1] + M0[2] * SM1[2];
        const float cc = M1[0] * SM1[0] + M1[1] * SM1[1] + M1[2] * SM1[2] + kLowPass;
        const float denom = ca * cc - cb * cb;
        const float d2 = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (d2 != 0.f) {
            dL_da = d2 * (-cc * cc * gx + 2.f * cb * cc * gy + (denom - ca * cc) * gz);
            dL_dc = d2 * (-ca * ca * gz + 2.f * ca * cb * gy + (denom - ca * cc) * gx);
            dL_db = d2 * 2.f * (cb * cc * gx - (denom + 2.f * cb * cb) * gy + ca * cb * gz);
            gcov[0] = M0[0] * M0[0] * dL_da + M0[0] * M1[0] * dL_db + M1[0] * M1[0] * dL_dc;
            gcov[3] = M0[1] * M0[1] * dL_da + M0[1] * M1[1] * dL_db + M1[1] * M1[1] * dL_dc;
            gcov[5] = M0[2] * M0[2] * dL_da + M0[2] * M1[2] * dL_db + M1[2] * M1[2] * dL_dc;
            gcov[1] = 2.f * M0[0] * M0[1] * dL_da + (M0[0] * M1[1] + M0[1] * M1[0]) * dL_db + 2.f * M1[0] * M1[1] * dL_dc;
            gcov[2] = 2.f * M0[0] * M0[2] * dL_da + (M0[0] * M1[2] + M0[2] * M1[0]) * dL_db + 2.f * M1[0] * M1[2] * dL_dc;
            gcov[4] = 2.f * M0[2] * M0[1] * dL_da + (M0[1] * M1[2] + M0[2] * M1[1]) * dL_db + 2.f * M1[1] * M1[2] * dL_dc;
        }
        float dM0[3], dM1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dM0[k] = 2.f * SM0[k] * dL_da + SM1[k] * dL_db;
            dM1[k] = 2.f * SM1[k] * dL_dc + SM0[k] * dL_db;
        }
        const float dJ00 = v[0] * dM0[0] + v[4] * dM0[1] + v[8] * dM0[2];
        const float dJ02 = v[2] * dM0[0] + v[6] * dM0[1] + v[10] * dM0[2];
        const float dJ11 = v[1] * dM1[0] + v[5] * dM1[1] + v[9] * dM1[2];
        const float dJ12 = v[2] * dM1[0] + v[6] * dM1[1] + v[10] * dM1[2];
        const float itz = 1.f / tz, tz2 = itz * itz, tz3 = tz2 * itz;
        const float dtx = xm * -a.focx * tz2 * dJ02, dty = ym * -a.focy * tz2 * dJ12;
        const float dtz = -a.focx * tz2 * dJ00 - a.focy * tz2 * dJ11 + (2.f * a.focx * tx) * tz3 * dJ02 + (2.f * a.focy * ty) * tz3 * dJ12;
        gmean[0] = v[0] * dtx + v[1] * dty + v[2] * dtz;
        gmean[1] = v[4] * dtx + v[5] * dty + v[6] * dtz;
        gmean[2] = v[8] * dtx + v[9] * dty + v[10] * dtz;
        // projection backward
        const float hx = p[0] * mx + p[4] * my + p[8] * mz + p[12];
        const float hy = p[1] * mx + p[5] * my + p[9] * mz + p[13];
        const float hw = p[3] * mx + p[7] * my + p[11] * mz + p[15];
        const float mw = 1.0f / (hw + 0.0000001f);
        const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        gmean[0] += (p[0] * mw - p[3] * mul1) * g2x + (p[1] * mw - p[3] * mul2) * g2y;
        gmean[1] += (p[4] * mw - p[7] * mul1) * g2x + (p[5] * mw - p[7] * mul2) * g2y;
        gmean[2] += (p[8] * mw - p[11] * mul1) * g2x + (p[9] * mw - p[11] * mul2) * g2y;

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

}  // namespace mb

using namespace mb;

extern "C" int mb_raster_forward_render(const mb_raster_inputs *in, void *geom, void *binning, size_t binning_bytes,
                                        int64_t capacity, void *image_buf, size_t image_bytes, float *out_color,
                                        mb_stream_t stream) {
    int rc = validate_raster_inputs(in, "mb_raster_forward_render");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterDims d = raster_dims(in);
    MB_REQUIRE(geom && binning && image_buf && out_color, "mb_raster_forward_render: null buffer");
    MB_REQUIRE(capacity >= 0 && capacity < (int64_t)0xffffffff, "mb_raster_forward_render: capacity out of range");
    GeomState g = GeomState::carve(geom, d.P);
    BinningState b = BinningState::carve(binning, capacity);
    ImageState im = ImageState::carve(image_buf, d.W, d.H);
    if (binning_bytes < b.bytes || image_bytes < im.bytes) {
        set_error("mb_raster_forward_render: binning %zu/%zu or image %zu/%zu bytes too small", binning_bytes, b.bytes,
                  image_bytes, im.bytes);
        return MB_ERR_WORKSPACE;
    }
    if (d.P > 0 && capacity > 0) {
        rc = build_instances(in, d, g, b, im, capacity, s);
        if (rc) return rc;
    } else {
        MB_CUDA(cudaMemsetAsync(im.ranges, 0, sizeof(uint2) * (size_t)d.tiles, s));
    }
    {
        KernelTimer kt("blend_forward", s);
        blend_forward_kernel<<<d.tiles, 256, 0, s>>>(b.records, im.ranges, d.W, d.H, d.gx, in->background, out_color, im.final_T,
                                                     im.n_contrib);
    }
    return check_launch("blend_forward", in->debug != 0, s);
}

extern "C" size_t mb_raster_backward_scratch_bytes(int32_t num_points) {
    return align_up((size_t)(num_points > 0 ? num_points : 1) * kAccStride * sizeof(float));
}

extern "C" int mb_raster_backward(const mb_raster_inputs *in, const int32_t *radii, const void *geom, const void *binning,
                                  int64_t capacity, const void *image_buf, const float *dL_dout, int64_t stride_c,
                                  int64_t stride_y, int64_t stride_x, void *grad_scratch, size_t scratch_bytes,
                                  float *dL_dmeans2D, float *dL_dcolors, float *dL_dopacity, float *dL_dmeans3D,
                                  float *dL_dcov3D, float *dL_dsh, float *dL_dscales, float *dL_drotations,
                                  mb_stream_t stream) {
    int rc = validate_raster_inputs(in, "mb_raster_backward");
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterDims d = raster_dims(in);
    if (d.P == 0) return MB_OK;
    MB_REQUIRE(radii && geom && binning && image_buf && dL_dout && grad_scratch, "mb_raster_backward: null buffer");
    MB_REQUIRE(dL_dmeans2D && dL_dcolors && dL_dopacity && dL_dmeans3D && dL_dcov3D, "mb_raster_backward: null output");
    MB_REQUIRE(!in->shs || dL_dsh, "mb_raster_backward: dL_dsh required when shs is given");
    MB_REQUIRE(in->cov3D_precomp || (dL_dscales && dL_drotations), "mb_raster_backward: dL_dscales / dL_drotations required");
    if (scratch_bytes < mb_raster_backward_scratch_bytes(d.P)) {
        set_error("mb_raster_backward: scratch too small");
        return MB_ERR_WORKSPACE;
    }
    const bool dbg = in->debug != 0;
    GeomState g = GeomState::carve(const_cast<void *>(geom), d.P);
    BinningState b = BinningState::carve(const_cast<void *>(binning), capacity);
    ImageState im = ImageState::carve(const_cast<void *>(image_buf), d.W, d.H);
    float *acc = reinterpret_cast<float *>(grad_scratch);
    MB_CUDA(cudaMemsetAsync(acc, 0, (size_t)d.P * kAccStride * sizeof(float), s));
    if (capacity > 0) {
        {
            KernelTimer kt("blend_backward", s);
            blend_backward_kernel<<<d.tiles, 256, 0, s>>>(b.records, im.ranges, d.W, d.H, d.gx, in->background, im.final_T,
                                                          im.n_contrib, dL_dout, stride_c, stride_y, stride_x, acc);
        }
        rc = check_launch("blend_backward", dbg, s);
        if (rc) return rc;
    }
    PreBwdArgs a;
    a.P = d.P; a.W = d.W; a.H = d.H; a.deg = in->sh_degree; a.M = in->sh_coeffs;
    a.tanx = in->tanfovx; a.tany = in->tanfovy; a.focx = d.focx; a.focy = d.focy; a.scale_mod = in->scale_modifier;
    a.means3D = in->means3D; a.cov3D = in->cov3D_precomp ? in->cov3D_precomp : g.cov3D; a.scales = in->scales;
    a.rots = in->rotations; a.shs = in->shs; a.view = in->viewmatrix; a.proj = in->projmatrix; a.campos = in->campos;
    a.radii = radii; a.clamped = g.clamped; a.acc = acc;
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
