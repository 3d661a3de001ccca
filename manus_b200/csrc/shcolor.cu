// SH -> RGB for callers that hold the materialised per-Gaussian transform tf [N,4,4] (the reference's calculate_colors_from_sh,
// src/utils/gaussian_utils.py:431-449, with eval_sh of src/utils/sh_utils.py:57-120): the view direction is taken in canonical
// space, d = cano_mean - (inv(tf) [campos; 1])[:3].  The reference runs torch.linalg.inv on N 4x4 matrices (batched LU) and ~60
// elementwise kernels; here one thread per Gaussian solves tf y = [c; 1] by the adjugate (closed form, any invertible 4x4),
// evaluates the basis and the three dot products, forward and backward.  tf = None (static object): d = posed_mean - campos.
//
// Backward: colour = max(sum_k basis_k(dir) f_k + 0.5, 0); go = clamp mask * dL/dcolour; dL/df_k = basis_k go;
// dL/ddir = sum_k grad basis_k (f_k . go); dL/dd = (I - dir dir^T) dL/ddir / |d|; dL/dcano_mean = dL/dd (tf given) or
// dL/dposed_mean = dL/dd (tf = None); dL/dy = -dL/dd; y = tf^-1 h  =>  dL/dtf = -(tf^-T dL/dy) y^T.
#include "common.cuh"

namespace mb {

// adj(M) for a row-major 4x4 and det(M); inv(M) = adj / det
__device__ __forceinline__ float adjugate4(const float *m, float *adj) {
    const float s0 = m[0] * m[5] - m[4] * m[1], s1 = m[0] * m[6] - m[4] * m[2], s2 = m[0] * m[7] - m[4] * m[3];
    const float s3 = m[1] * m[6] - m[5] * m[2], s4 = m[1] * m[7] - m[5] * m[3], s5 = m[2] * m[7] - m[6] * m[3];
    const float c5 = m[10] * m[15] - m[14] * m[11], c4 = m[9] * m[15] - m[13] * m[11], c3 = m[9] * m[14] - m[13] * m[10];
    const float c2 = m[8] * m[15] - m[12] * m[11], c1 = m[8] * m[14] - m[12] * m[10], c0 = m[8] * m[13] - m[12] * m[9];
    adj[0] = m[5] * c5 - m[6] * c4 + m[7] * c3;   adj[1] = -m[1] * c5 + m[2] * c4 - m[3] * c3;
    adj[2] = m[13] * s5 - m[14] * s4 + m[15] * s3; adj[3] = -m[9] * s5 + m[10] * s4 - m[11] * s3;
    adj[4] = -m[4] * c5 + m[6] * c2 - m[7] * c1;  adj[5] = m[0] * c5 - m[2] * c2 + m[3] * c1;
    adj[6] = -m[12] * s5 + m[14] * s2 - m[15] * s1; adj[7] = m[8] * s5 - m[10] * s2 + m[11] * s1;
    adj[8] = m[4] * c4 - m[5] * c2 + m[7] * c0;   adj[9] = -m[0] * c4 + m[1] * c2 - m[3] * c0;
    adj[10] = m[12] * s4 - m[13] * s2 + m[15] * s0; adj[11] = -m[8] * s4 + m[9] * s2 - m[11] * s0;
    adj[12] = -m[4] * c3 + m[5] * c1 - m[6] * c0; adj[13] = m[0] * c3 - m[1] * c1 + m[2] * c0;
    adj[14] = -m[12] * s3 + m[13] * s1 - m[14] * s0; adj[15] = m[8] * s3 - m[9] * s1 + m[10] * s0;
    return s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
}

struct ShColorArgs {
    int N, deg, K;
    const float *means;      // cano_means (tf given) or posed_means (tf = NULL), [N,3]
    const float *features;   // [N,K,3]
    const float *tf;         // [N,4,4] or NULL
    const float *campos;     // [3]
    float *colors;           // forward out [N,3]
    const float *g_colors;   // backward in [N,3]
    float *g_means, *g_features, *g_tf;   // backward out ([N,3], [N,K,3], [N,4,4] or NULL)
};

template <bool kBackward>
__global__ void __launch_bounds__(128) sh_colors_kernel(ShColorArgs a) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= a.N) return;
    const float c[3] = {a.campos[0], a.campos[1], a.campos[2]};
    float y[4] = {c[0], c[1], c[2], 1.f}, adj[16], idet = 0.f;
    if (a.tf) {
        float m[16];
        const float4 *t4 = reinterpret_cast<const float4 *>(a.tf + 16 * (size_t)i);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 v = t4[r];
            m[4 * r] = v.x; m[4 * r + 1] = v.y; m[4 * r + 2] = v.z; m[4 * r + 3] = v.w;
        }
        idet = 1.0f / adjugate4(m, adj);
#pragma unroll
        for (int r = 0; r < 4; ++r) y[r] = (adj[4 * r] * c[0] + adj[4 * r + 1] * c[1] + adj[4 * r + 2] * c[2] + adj[4 * r + 3]) * idet;
    }
    const float d0 = a.means[3 * (size_t)i] - y[0], d1 = a.means[3 * (size_t)i + 1] - y[1], d2 = a.means[3 * (size_t)i + 2] - y[2];
    const float n = sqrtf(d0 * d0 + d1 * d1 + d2 * d2), in = 1.0f / n;
    const float dir[3] = {d0 * in, d1 * in, d2 * in};
    float basis[16];
    sh_basis(a.deg, dir[0], dir[1], dir[2], basis);
    const int nb = (a.deg + 1) * (a.deg + 1);
    const float *f = a.features + (size_t)i * a.K * 3;
    float col[3] = {0.5f, 0.5f, 0.5f};
    for (int k = 0; k < nb; ++k)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) col[ch] += basis[k] * f[3 * k + ch];
    if (!kBackward) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) a.colors[3 * (size_t)i + ch] = fmaxf(col[ch], 0.f);
        return;
    }
    float go[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) go[ch] = col[ch] >= 0.f ? a.g_colors[3 * (size_t)i + ch] : 0.f;
    float bx[16], by[16], bz[16], gd[3] = {0.f, 0.f, 0.f};
    sh_basis_grad(a.deg, dir[0], dir[1], dir[2], bx, by, bz);
    float *gf = a.g_features + (size_t)i * a.K * 3;
    for (int k = 0; k < a.K; ++k)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            if (k < nb) {
                gf[3 * k + ch] = basis[k] * go[ch];
                const float sv = f[3 * k + ch] * go[ch];
                gd[0] += bx[k] * sv; gd[1] += by[k] * sv; gd[2] += bz[k] * sv;
            } else gf[3 * k + ch] = 0.f;
        }
    const float dot = dir[0] * gd[0] + dir[1] * gd[1] + dir[2] * gd[2];
    const float g[3] = {(gd[0] - dir[0] * dot) * in, (gd[1] - dir[1] * dot) * in, (gd[2] - dir[2] * dot) * in};   // dL/dd
#pragma unroll
    for (int r = 0; r < 3; ++r) a.g_means[3 * (size_t)i + r] = g[r];
    if (a.tf && a.g_tf) {
        // z = tf^-T (-g, 0): z_j = sum_r inv[r][j] * (-g_r) ; dL/dtf[j][k] = -z_j y_k ... with the sign folded: dL/dtf = -(tf^-T gy) y^T, gy = -g
        float z[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = (adj[j] * g[0] + adj[4 + j] * g[1] + adj[8 + j] * g[2]) * idet;   // = tf^-T g (rows 0..2 of inv)
        float4 *o = reinterpret_cast<float4 *>(a.g_tf + 16 * (size_t)i);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = make_float4(z[j] * y[0], z[j] * y[1], z[j] * y[2], z[j] * y[3]);
    }
}

}  // namespace mb

using namespace mb;

static int sh_colors_common(ShColorArgs &a, const char *who) {
    MB_REQUIRE(a.N >= 0 && a.deg >= 0 && a.deg <= 3 && a.K >= (a.deg + 1) * (a.deg + 1) && a.K <= 16, "%s: bad sizes N=%d degree=%d K=%d", who, a.N, a.deg,
               a.K);
    MB_REQUIRE(a.N == 0 || (a.means && a.features && a.campos), "%s: null input", who);
    return MB_OK;
}

extern "C" int mb_sh_colors_forward(const float *means, const float *features, const float *tf, const float *campos, int32_t num_points,
                                    int32_t sh_degree, int32_t sh_coeffs, float *colors, mb_stream_t stream) {
    ShColorArgs a = {};
    a.N = num_points; a.deg = sh_degree; a.K = sh_coeffs; a.means = means; a.features = features; a.tf = tf; a.campos = campos; a.colors = colors;
    int rc = sh_colors_common(a, "mb_sh_colors_forward");
    if (rc || num_points == 0) return rc;
    MB_REQUIRE(colors != nullptr, "mb_sh_colors_forward: null output");
    cudaStream_t s = (cudaStream_t)stream;
    KernelTimer kt("sh_colors", s);
    sh_colors_kernel<false><<<(num_points + 127) / 128, 128, 0, s>>>(a);
    return check_launch("sh_colors_forward", false, s);
}

extern "C" int mb_sh_colors_backward(const float *means, const float *features, const float *tf, const float *campos, int32_t num_points,
                                     int32_t sh_degree, int32_t sh_coeffs, const float *g_colors, float *g_means, float *g_features, float *g_tf,
                                     mb_stream_t stream) {
    ShColorArgs a = {};
    a.N = num_points; a.deg = sh_degree; a.K = sh_coeffs; a.means = means; a.features = features; a.tf = tf; a.campos = campos;
    a.g_colors = g_colors; a.g_means = g_means; a.g_features = g_features; a.g_tf = g_tf;
    int rc = sh_colors_common(a, "mb_sh_colors_backward");
    if (rc || num_points == 0) return rc;
    MB_REQUIRE(g_colors && g_means && g_features, "mb_sh_colors_backward: null gradient buffer");
    cudaStream_t s = (cudaStream_t)stream;
    KernelTimer kt("sh_colors_backward", s);
    sh_colors_kernel<true><<<(num_points + 127) / 128, 128, 0, s>>>(a);
    return check_launch("sh_colors_backward", false, s);
}
