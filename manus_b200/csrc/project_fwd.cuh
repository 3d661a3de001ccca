// Forward projection of ONE Gaussian: cull, EWA 2-D covariance, conic, radius, tile rectangle,
// depth key and the 48-B blend record.  Shared by preprocess_kernel (raster_geom.cu) and by the fused pose + projection
// forward (pose.cu), which applies it to the posed mean / covariance / colour / opacity while they are still in registers.
//
// Every quantity that feeds an integer decision (near cull, radius ceil, tile rectangle truncation, depth order) is built
// from individually rounded IEEE operations in a fixed order -- explicit __fmul_rn / __fadd_rn, which the compiler never
// contracts into FMAs -- so radii / tiles touched / instance lists are reproducible bit for bit whatever the flags of the
// including translation unit (they equal the C oracle's, which is compiled without contraction).
#pragma once
#include "raster_state.cuh"

namespace mb {

__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
// ((a x + b y) + c z) + d, every operation rounded
__device__ __forceinline__ float affine3(float a, float b, float c, float d, float x, float y, float z) {
    return add_(add_(add_(mul_(a, x), mul_(b, y)), mul_(c, z)), d);
}
__device__ __forceinline__ float dot3(float a, float b, float c, float x, float y, float z) {
    return add_(add_(mul_(a, x), mul_(b, y)), mul_(c, z));
}

__device__ __forceinline__ void tile_rect(float px, float py, int rad, int gx, int gy, int &x0, int &y0, int &x1, int &y1) {
    // division by BLOCK = 16: the multiplication by 1/16 is the same IEEE result (power of two)
    const float r = (float)rad, t = (float)kTile, it = 1.0f / (float)kTile;
    x0 = min(gx, max(0, (int)mul_(sub_(px, r), it)));
    y0 = min(gy, max(0, (int)mul_(sub_(py, r), it)));
    x1 = min(gx, max(0, (int)mul_(sub_(add_(add_(px, r), t), 1.0f), it)));   // (px + rad + BLOCK - 1) / BLOCK, left to right
    y1 = min(gy, max(0, (int)mul_(sub_(add_(add_(py, r), t), 1.0f), it)));
}

struct Projected {
    int radius;          // upstream's radii (0 = culled)
    uint32_t tiles;      // instances this Gaussian emits
    uint32_t key;        // depth key (0xffffffff = culled: sinks to the end of the depth order)
    ushort4 rect;        // tile rectangle (valid when tiles > 0)
    Record rec;          // valid when radius > 0
    bool visible;
};

// v = viewmatrix[16], p = projmatrix[16] (column-major as upstream reads them); c6 = 3-D covariance (xx,xy,xz,yy,yz,zz)
__device__ __forceinline__ void project_forward(const float *v, const float *p, float tanx, float tany, float focx, float focy, int W,
                                                int H, int gx, int gy, float mx, float my, float mz, const float *c6, float op,
                                                const float *rgb, Projected &out) {
    out.radius = 0; out.tiles = 0; out.key = 0xffffffffu; out.visible = false;
    out.rect = make_ushort4(0, 0, 0, 0);
    // A.1 step 2: view space
    const float tx0 = affine3(v[0], v[4], v[8], v[12], mx, my, mz);
    const float ty0 = affine3(v[1], v[5], v[9], v[13], mx, my, mz);
    const float tz = affine3(v[2], v[6], v[10], v[14], mx, my, mz);
    if (!(tz > kNearZ)) return;
    // step 3
    const float hx = affine3(p[0], p[4], p[8], p[12], mx, my, mz);
    const float hy = affine3(p[1], p[5], p[9], p[13], mx, my, mz);
    const float hw = affine3(p[3], p[7], p[11], p[15], mx, my, mz);
    const float pw = div_(1.0f, add_(hw, 0.0000001f));
    const float ndcx = mul_(hx, pw), ndcy = mul_(hy, pw);
    // step 5: EWA projection with the clamped view-space point
    const float limx = mul_(1.3f, tanx), limy = mul_(1.3f, tany);
    float rx = div_(tx0, tz), ry = div_(ty0, tz);
    rx = rx < -limx ? -limx : (rx > limx ? limx : rx);
    ry = ry < -limy ? -limy : (ry > limy ? limy : ry);
    const float tx = mul_(rx, tz), ty = mul_(ry, tz);
    const float tz2 = mul_(tz, tz);
    const float J00 = div_(focx, tz), J02 = -div_(mul_(focx, tx), tz2);
    const float J11 = div_(focy, tz), J12 = -div_(mul_(focy, ty), tz2);
    float M0[3], M1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        M0[k] = add_(mul_(J00, v[4 * k + 0]), mul_(J02, v[4 * k + 2]));
        M1[k] = add_(mul_(J11, v[4 * k + 1]), mul_(J12, v[4 * k + 2]));
    }
    const float S[9] = {c6[0], c6[1], c6[2], c6[1], c6[3], c6[4], c6[2], c6[4], c6[5]};
    float SM0[3], SM1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        SM0[k] = dot3(S[3 * k], S[3 * k + 1], S[3 * k + 2], M0[0], M0[1], M0[2]);
        SM1[k] = dot3(S[3 * k], S[3 * k + 1], S[3 * k + 2], M1[0], M1[1], M1[2]);
    }
    const float ca = add_(dot3(M0[0], M0[1], M0[2], SM0[0], SM0[1], SM0[2]), kLowPass);
    const float cb = dot3(M0[0], M0[1], M0[2], SM1[0], SM1[1], SM1[2]);
    const float cc = add_(dot3(M1[0], M1[1], M1[2], SM1[0], SM1[1], SM1[2]), kLowPass);
    // steps 6-7
    const float det = sub_(mul_(ca, cc), mul_(cb, cb));
    if (det == 0.0f) return;
    const float di = div_(1.0f, det);
    const float mid = mul_(0.5f, add_(ca, cc));
    float disc = sub_(mul_(mid, mid), det);
    if (disc < 0.1f) disc = 0.1f;
    const float sq = __fsqrt_rn(disc);
    const float l1 = add_(mid, sq), l2 = sub_(mid, sq);
    const int rad = (int)ceilf(mul_(3.0f, __fsqrt_rn(l1 > l2 ? l1 : l2)));
    // steps 8-9
    const float px = mul_(sub_(mul_(add_(ndcx, 1.0f), (float)W), 1.0f), 0.5f), py = mul_(sub_(mul_(add_(ndcy, 1.0f), (float)H), 1.0f), 0.5f);
    int x0, y0, x1, y1;
    tile_rect(px, py, rad, gx, gy, x0, y0, x1, y1);
    if ((x1 - x0) * (y1 - y0) <= 0) return;
    out.visible = true;
    out.radius = rad;
    out.key = __float_as_uint(tz);
    // below this power, op * exp(power) < 1/255 with a wide margin (NaN for op < 0: never skips)
    const float cut = sub_(-logf(mul_(255.0f, op)), 1e-4f);
    // half extents of the bounding box of { power >= cut }: dx^2 <= 2|cut| cov2D.xx, dy^2 <= 2|cut| cov2D.yy
    // (NaN when no pixel can pass the alpha gate: such a record never survives the tile kernels' box test)
    float ex = add_(mul_(__fsqrt_rn(mul_(mul_(-2.0f, cut), ca)), 1.0001f), 0.01f);
    float ey = add_(mul_(__fsqrt_rn(mul_(mul_(-2.0f, cut), cc)), 1.0001f), 0.01f);
    // an indefinite 2-D covariance (det < 0: only possible with a non-PSD cov3D_precomp) has no bounded
    // { power >= cut } set: keep upstream's whole rectangle and let the per-pixel gates decide
    if (det < 0.0f && cut <= 0.0f) ex = ey = 1.0e9f;
    // Instances are only emitted for the tiles of upstream's rectangle that this box reaches: in the others every pixel
    // fails the alpha >= 1/255 gate, so dropping them changes neither image nor gradients (it only shortens the lists;
    // radii and visibility stay upstream's).
    uint32_t tiles = 0;
    if (ex >= 0.f && ey >= 0.f) {
        const float it = 1.0f / (float)kTile;
        x0 = max(x0, (int)floorf(mul_(sub_(px, ex), it)));
        y0 = max(y0, (int)floorf(mul_(sub_(py, ey), it)));
        x1 = min(x1, (int)floorf(mul_(add_(px, ex), it)) + 1);
        y1 = min(y1, (int)floorf(mul_(add_(py, ey), it)) + 1);
        tiles = (uint32_t)(max(x1 - x0, 0) * max(y1 - y0, 0));
    }
    if (tiles == 0) x0 = y0 = x1 = y1 = 0;
    out.tiles = tiles;
    out.rect = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
    out.rec.a = make_float4(px, py, mul_(cc, di), mul_(-cb, di));
    out.rec.b = make_float4(mul_(ca, di), op, rgb[0], rgb[1]);
    out.rec.c = make_float4(rgb[2], cut, ex, ey);
}

}  // namespace mb
