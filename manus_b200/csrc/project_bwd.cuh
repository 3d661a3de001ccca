// Backward of the projection step of the rasterizer for ONE Gaussian (computeCov2D backward + the
// gradient of the projected mean), shared by preprocess_backward_kernel (raster_blend.cu) and by the fused pose backward
// (pose.cu), which applies it to the rasterizer's accumulator rows without a round trip through HBM.
#pragma once
#include "common.cuh"

namespace mb {

// v = viewmatrix[16], p = projmatrix[16] (column-major as upstream reads them); (mx, my, mz) world-space mean; c6 = 3-D
// covariance (xx,xy,xz,yy,yz,zz); opacity as the forward used it; W, H image size.
// m[5] = the blend backward's moments of q = G dL/dalpha over the Gaussian's pixels, d = mean2D - pixel:
//   sum q dx, sum q dy, sum q dx^2, sum q dx dy, sum q dy^2.   With dL/dG = opacity dL/dalpha and
//   dG/dd = -G (conic.x dx + conic.y dy, conic.z dy + conic.y dx) they give (A.4)
//   dL/dmean2D = -opacity (conic.x m0 + conic.y m1, conic.z m1 + conic.y m0) * (W/2, H/2)   [NDC-scaled, upstream's ddelx_dx]
//   dL/dconic  = -0.5 opacity (m2, m3, m4).
// Writes g2[2] = dL/dmean2D, gmean[3] = dL/dmean3D and gcov[6] = dL/dcov3D.
__device__ __forceinline__ void project_backward(const float *v, const float *p, float tanx, float tany, float focx, float focy,
                                                 int W, int H, float mx, float my, float mz, const float *c6, float opacity,
                                                 const float *m, float *g2, float *gmean, float *gcov) {
    // cov2D backward (A.4)
    const float t0 = v[0] * mx + v[4] * my + v[8] * mz + v[12];
    const float t1 = v[1] * mx + v[5] * my + v[9] * mz + v[13];
    const float tz = v[2] * mx + v[6] * my + v[10] * mz + v[14];
    const float limx = 1.3f * tanx, limy = 1.3f * tany;
    const float rx = t0 / tz, ry = t1 / tz;
    const float xm = (rx < -limx || rx > limx) ? 0.f : 1.f, ym = (ry < -limy || ry > limy) ? 0.f : 1.f;
    const float tx = fminf(limx, fmaxf(-limx, rx)) * tz, ty = fminf(limy, fmaxf(-limy, ry)) * tz;
    const float J00 = focx / tz, J02 = -(focx * tx) / (tz * tz);
    const float J11 = focy / tz, J12 = -(focy * ty) / (tz * tz);
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
    // the conic of the forward (A.1 step 6) and the per-Gaussian part of the blend backward
    const float di = 1.0f / denom;
    const float conx = cc * di, cony = -cb * di, conz = ca * di;
    const float g2x = -opacity * (conx * m[0] + cony * m[1]) * (0.5f * W);
    const float g2y = -opacity * (conz * m[1] + cony * m[0]) * (0.5f * H);
    const float gx = -0.5f * opacity * m[2], gy = -0.5f * opacity * m[3], gz = -0.5f * opacity * m[4];
    g2[0] = g2x; g2[1] = g2y;
    const float d2 = 1.0f / (denom * denom + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) gcov[k] = 0.f;
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
    const float dtx = xm * -focx * tz2 * dJ02, dty = ym * -focy * tz2 * dJ12;
    const float dtz = -focx * tz2 * dJ00 - focy * tz2 * dJ11 + (2.f * focx * tx) * tz3 * dJ02 + (2.f * focy * ty) * tz3 * dJ12;
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
}

}  // namespace mb
