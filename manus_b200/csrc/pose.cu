// Fused pre-raster step: linear-blend skinning of means and covariances, covariance build from (log-scale, quaternion),
// SH -> RGB in the canonical-space view direction, sigmoid / exp activations -- forward and backward in one kernel each.
//
// Replaces ~100 small PyTorch kernels per frame of the reference (src/modules/hand_dynamic.py:86-137,
// src/models/gaussian.py:48-93, src/utils/gaussian_utils.py:248-314,431-449, src/utils/sh_utils.py:57-120): the per-Gaussian
// 4x4 transform is never written to HBM and its inverse (torch.linalg.inv on N 4x4 matrices in the reference) is the
// closed-form affine inverse.  One thread per Gaussian, 128 Gaussians per CTA; the wide rows (SH coefficients, skin
// weights) move through shared memory so that every HBM access is a coalesced run; strides are padded to odd word
// counts so the per-thread row reads are bank-conflict free.
//
// HBM bytes per Gaussian (fp32, K=16, B bones): forward 236 + 4B read, 52 written; backward re-reads the parameters
// and the 52 B of upstream gradients and writes 236 B of parameter gradients.
#include "common.cuh"

namespace mb {

constexpr int kPoseThreads = 128;
constexpr int kMaxBones = 64;

struct PoseArgs {
    int N, n_skinned, B, deg, K, iso;
    const float *xyz, *log_scale, *quat, *opacity_logit, *f_dc, *f_rest, *skin, *bone_tf, *campos;
    // forward outputs
    float *posed_xyz, *cov6, *colors, *opacity, *tf_out;
    // backward inputs / outputs
    const float *g_posed_xyz, *g_cov6, *g_colors, *g_opacity;
    float *g_xyz, *g_log_scale, *g_quat, *g_opacity_logit, *g_f_dc, *g_f_rest, *g_skin;
};

__host__ __device__ inline int odd_stride(int n) { return n | 1; }

inline size_t pose_smem_bytes(int B, int K) {
    const int rs = (K - 1) * 3;
    return sizeof(float) * ((size_t)B * 13 + 4 + (size_t)kPoseThreads * (rs > 0 ? odd_stride(rs) : 0) +
                            (size_t)kPoseThreads * (B > 0 ? odd_stride(B) : 0));
}

// coalesced copy of `rows` rows of `width` floats (dense in global memory) into padded shared rows
__device__ __forceinline__ void stage_rows_in(float *dst, int dst_stride, const float *__restrict__ src, int rows, int width) {
    const int total = rows * width;
    for (int j = threadIdx.x; j < total; j += kPoseThreads) {
        const int r = j / width, c = j - r * width;
        dst[r * dst_stride + c] = src[j];
    }
}
__device__ __forceinline__ void stage_rows_out(float *__restrict__ dst, const float *src, int src_stride, int rows, int width) {
    const int total = rows * width;
    for (int j = threadIdx.x; j < total; j += kPoseThreads) {
        const int r = j / width, c = j - r * width;
        dst[j] = src[r * src_stride + c];
    }
}

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

// everything both directions need: blended transform, rotation, scales
__device__ __forceinline__ void pose_common(const PoseArgs &a, int i, int row, const float *bones_s, const float *w_s, int ws,
                                            PoseLocal &p) {
    p.x[0] = a.xyz[3 * (size_t)i]; p.x[1] = a.xyz[3 * (size_t)i + 1]; p.x[2] = a.xyz[3 * (size_t)i + 2];
    p.skinned = i < a.n_skinned;
    if (p.skinned) {
#pragma unroll
        for (int k = 0; k < 9; ++k) p.A[k] = 0.f;
        p.t[0] = p.t[1] = p.t[2] = 0.f; p.s = 0.f;
        for (int b = 0; b < a.B; ++b) {
            const float w = w_s[row * ws + b];
            if (w == 0.f) continue;
            const float *T = bones_s + 13 * b;
            p.A[0] += w * T[0]; p.A[1] += w * T[1]; p.A[2] += w * T[2]; p.t[0] += w * T[3];
            p.A[3] += w * T[4]; p.A[4] += w * T[5]; p.A[5] += w * T[6]; p.t[1] += w * T[7];
            p.A[6] += w * T[8]; p.A[7] += w * T[9]; p.A[8] += w * T[10]; p.t[2] += w * T[11];
            p.s += w * T[12];
        }
    } else {
        p.A[0] = p.A[4] = p.A[8] = 1.f;
        p.A[1] = p.A[2] = p.A[3] = p.A[5] = p.A[6] = p.A[7] = 0.f;
        p.t[0] = p.t[1] = p.t[2] = 0.f; p.s = 1.f;
    }
    const float q0 = a.quat[4 * (size_t)i], q1 = a.quat[4 * (size_t)i + 1], q2 = a.quat[4 * (size_t)i + 2], q3 = a.quat[4 * (size_t)i + 3];
    p.qnorm = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    p.qn[0] = q0 / p.qnorm; p.qn[1] = q1 / p.qnorm; p.qn[2] = q2 / p.qnorm; p.qn[3] = q3 / p.qnorm;
    quat_to_rot(p.qn[0], p.qn[1], p.qn[2], p.qn[3], p.R);
    if (a.iso) {
        p.S[0] = p.S[1] = p.S[2] = expf(a.log_scale[i]);
    } else {
        p.S[0] = expf(a.log_scale[3 * (size_t)i]); p.S[1] = expf(a.log_scale[3 * (size_t)i + 1]); p.S[2] = expf(a.log_scale[3 * (size_t)i + 2]);
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

__device__ __forceinline__ void load_block_inputs(const PoseArgs &a, int base, int cnt, float *bones_s, float *cam_s, float *fr_s,
                                                  int frs, float *w_s, int ws) {
    for (int j = threadIdx.x; j < a.B * 13; j += kPoseThreads) {
        const int b = j / 13, e = j - 13 * b;
        bones_s[j] = a.bone_tf[16 * b + (e < 12 ? e : 15)];
    }
    if (threadIdx.x < 3) cam_s[threadIdx.x] = a.campos[threadIdx.x];
    const int rs = (a.K - 1) * 3;
    if (rs > 0) stage_rows_in(fr_s, frs, a.f_rest + (size_t)base * rs, cnt, rs);
    const int nsk = min(cnt, a.n_skinned - base);
    if (nsk > 0) stage_rows_in(w_s, ws, a.skin + (size_t)base * a.B, nsk, a.B);
    __syncthreads();
}

__global__ void __launch_bounds__(kPoseThreads) pose_forward_kernel(PoseArgs a) {
    extern __shared__ float smem[];
    const int rs = (a.K - 1) * 3, frs = rs > 0 ? odd_stride(rs) : 0, ws = a.B > 0 ? odd_stride(a.B) : 0;
    float *bones_s = smem, *cam_s = bones_s + a.B * 13, *fr_s = cam_s + 4, *w_s = fr_s + kPoseThreads * frs;
    const int base = blockIdx.x * kPoseThreads;
    const int cnt = min(kPoseThreads, a.N - base);
    load_block_inputs(a, base, cnt, bones_s, cam_s, fr_s, frs, w_s, ws);
    const int row = threadIdx.x, i = base + row;
    if (row >= cnt) return;
    PoseLocal p;
    pose_common(a, i, row, bones_s, w_s, ws, p);
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
    sh_basis(a.deg, dir[0], dir[1], dir[2], basis);
    const int nb = (a.deg + 1) * (a.deg + 1);
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = basis[0] * a.f_dc[3 * (size_t)i + c];
        for (int k = 1; k < nb; ++k) v += basis[k] * fr_s[row * frs + 3 * (k - 1) + c];
        rgb[c] = fmaxf(v + 0.5f, 0.f);
    }
    const float ol = a.opacity_logit[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        a.posed_xyz[3 * (size_t)i + r] = px[r];
        a.colors[3 * (size_t)i + r] = rgb[r];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) a.cov6[6 * (size_t)i + k] = c6[k];
    a.opacity[i] = 1.0f / (1.0f + expf(-ol));
    if (a.tf_out && p.skinned) {
        float *o = a.tf_out + 16 * (size_t)i;
        o[0] = p.A[0]; o[1] = p.A[1]; o[2] = p.A[2]; o[3] = p.t[0];
        o[4] = p.A[3]; o[5] = p.A[4]; o[6] = p.A[5]; o[7] = p.t[1];
        o[8] = p.A[6]; o[9] = p.A[7]; o[10] = p.A[8]; o[11] = p.t[2];
        // bottom row = sum_b w_b T_b[3,:]
        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
        for (int b = 0; b < a.B; ++b) {
            const float w = w_s[row * ws + b];
            b0 += w * a.bone_tf[16 * b + 12]; b1 += w * a.bone_tf[16 * b + 13]; b2 += w * a.bone_tf[16 * b + 14];
        }
        o[12] = b0; o[13] = b1; o[14] = b2; o[15] = p.s;
    }
}

__global__ void __launch_bounds__(kPoseThreads) pose_backward_kernel(PoseArgs a) {
    extern __shared__ float smem[];
    const int rs = (a.K - 1) * 3, frs = rs > 0 ? odd_stride(rs) : 0, ws = a.B > 0 ? odd_stride(a.B) : 0;
    float *bones_s = smem, *cam_s = bones_s + a.B * 13, *fr_s = cam_s + 4, *w_s = fr_s + kPoseThreads * frs;
    const int base = blockIdx.x * kPoseThreads;
    const int cnt = min(kPoseThreads, a.N - base);
    load_block_inputs(a, base, cnt, bones_s, cam_s, fr_s, frs, w_s, ws);
    const int row = threadIdx.x, i = base + row;
    if (row < cnt) {
        PoseLocal p;
        pose_common(a, i, row, bones_s, w_s, ws, p);
        float gx[3] = {0.f, 0.f, 0.f}, dA[9], dt[3] = {0.f, 0.f, 0.f}, ds = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) dA[k] = 0.f;
        // ---- mean: x' = A x + t
        const float gp[3] = {a.g_posed_xyz[3 * (size_t)i], a.g_posed_xyz[3 * (size_t)i + 1], a.g_posed_xyz[3 * (size_t)i + 2]};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            gx[r] = p.A[r] * gp[0] + p.A[3 + r] * gp[1] + p.A[6 + r] * gp[2];
            dt[r] = gp[r];
#pragma unroll
            for (int c = 0; c < 3; ++c) dA[3 * r + c] = gp[r] * p.x[c];
        }
        // ---- covariance: Sigma' = Bm Bm^T, Bm = A L ; Gs = symmetrised dL/dSigma'
        const float *g6 = a.g_cov6 + 6 * (size_t)i;
        const float Gs[9] = {g6[0], 0.5f * g6[1], 0.5f * g6[2], 0.5f * g6[1], g6[3], 0.5f * g6[4], 0.5f * g6[2], 0.5f * g6[4], g6[5]};
        float Bm[9], dB[9], dL[9];
        if (p.skinned) mat3_mul(p.A, p.L, Bm);
        else {
#pragma unroll
            for (int k = 0; k < 9; ++k) Bm[k] = p.L[k];
        }
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
#pragma unroll
        for (int k = 0; k < 4; ++k) a.g_quat[4 * (size_t)i + k] = (dqn[k] - p.qn[k] * qd) / p.qnorm;
        if (a.iso) a.g_log_scale[i] = dS[0] * p.S[0] + dS[1] * p.S[1] + dS[2] * p.S[2];
        else {
#pragma unroll
            for (int k = 0; k < 3; ++k) a.g_log_scale[3 * (size_t)i + k] = dS[k] * p.S[k];
        }
        // ---- colour
        float Ainv[9], ci[3], dir[3], basis[16], bxg[16], byg[16], bzg[16];
        if (p.skinned) mat3_inverse(p.A, Ainv);
        const float dn = view_dir(p, cam_s, Ainv, ci, dir);
        sh_basis(a.deg, dir[0], dir[1], dir[2], basis);
        sh_basis_grad(a.deg, dir[0], dir[1], dir[2], bxg, byg, bzg);
        const int nb = (a.deg + 1) * (a.deg + 1);
        float go[3], gd[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f0 = a.f_dc[3 * (size_t)i + c];
            float v = basis[0] * f0;
            for (int k = 1; k < nb; ++k) v += basis[k] * fr_s[row * frs + 3 * (k - 1) + c];
            go[c] = (v + 0.5f >= 0.f) ? a.g_colors[3 * (size_t)i + c] : 0.f;
            a.g_f_dc[3 * (size_t)i + c] = basis[0] * go[c];
        }
        for (int k = 1; k < a.K; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float *slot = fr_s + row * frs + 3 * (k - 1) + c;
                if (k < nb) {
                    const float s = *slot * go[c];
                    gd[0] += bxg[k] * s; gd[1] += byg[k] * s; gd[2] += bzg[k] * s;
                    *slot = basis[k] * go[c];
                } else *slot = 0.f;
            }
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
        for (int r = 0; r < 3; ++r) a.g_xyz[3 * (size_t)i + r] = gx[r];
        const float sg = 1.0f / (1.0f + expf(-a.opacity_logit[i]));
        a.g_opacity_logit[i] = a.g_opacity[i] * sg * (1.0f - sg);
        if (a.g_skin && p.skinned) {
            for (int b = 0; b < a.B; ++b) {
                const float *T = bones_s + 13 * b;
                w_s[row * ws + b] = dA[0] * T[0] + dA[1] * T[1] + dA[2] * T[2] + dt[0] * T[3] + dA[3] * T[4] + dA[4] * T[5] +
                                    dA[5] * T[6] + dt[1] * T[7] + dA[6] * T[8] + dA[7] * T[9] + dA[8] * T[10] + dt[2] * T[11] +
                                    ds * T[12];
            }
        }
    }
    __syncthreads();
    if (rs > 0) stage_rows_out(a.g_f_rest + (size_t)base * rs, fr_s, frs, cnt, rs);
    const int nsk = min(cnt, a.n_skinned - base);
    if (a.g_skin && nsk > 0) stage_rows_out(a.g_skin + (size_t)base * a.B, w_s, ws, nsk, a.B);
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
        MB_REQUIRE(in->skin_wts && in->bone_tf, "%s: skin_wts / bone_tf missing", who);
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
    return a;
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
    const size_t smem = pose_smem_bytes(a.B, a.K);
    MB_CUDA(cudaFuncSetAttribute(pose_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelTimer kt("pose_forward", (cudaStream_t)stream);
        pose_forward_kernel<<<(a.N + kPoseThreads - 1) / kPoseThreads, kPoseThreads, smem, (cudaStream_t)stream>>>(a);
    }
    return check_launch("pose_forward", false, (cudaStream_t)stream);
}

extern "C" int mb_pose_backward(const mb_pose_inputs *in, const float *g_posed_xyz, const float *g_posed_cov6,
                                const float *g_colors, const float *g_opacity, float *g_xyz, float *g_log_scale, float *g_quat,
                                float *g_opacity_logit, float *g_f_dc, float *g_f_rest, float *g_skin_wts, mb_stream_t stream) {
    int rc = validate_pose(in, "mb_pose_backward");
    if (rc) return rc;
    if (in->num_points == 0) return MB_OK;
    MB_REQUIRE(g_posed_xyz && g_posed_cov6 && g_colors && g_opacity, "mb_pose_backward: null upstream gradient");
    MB_REQUIRE(g_xyz && g_log_scale && g_quat && g_opacity_logit && g_f_dc && (in->sh_coeffs == 1 || g_f_rest),
               "mb_pose_backward: null output");
    PoseArgs a = pose_args(in);
    a.g_posed_xyz = g_posed_xyz; a.g_cov6 = g_posed_cov6; a.g_colors = g_colors; a.g_opacity = g_opacity;
    a.g_xyz = g_xyz; a.g_log_scale = g_log_scale; a.g_quat = g_quat; a.g_opacity_logit = g_opacity_logit;
    a.g_f_dc = g_f_dc; a.g_f_rest = g_f_rest; a.g_skin = g_skin_wts;
    const size_t smem = pose_smem_bytes(a.B, a.K);
    MB_CUDA(cudaFuncSetAttribute(pose_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelTimer kt("pose_backward", (cudaStream_t)stream);
        pose_backward_kernel<<<(a.N + kPoseThreads - 1) / kPoseThreads, kPoseThreads, smem, (cudaStream_t)stream>>>(a);
    }
    return check_launch("pose_backward", false, (cudaStream_t)stream);
}
