// Shared helpers for the sm_100a kernels: error plumbing, mbarrier / bulk-copy (TMA) PTX wrappers, small math.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/manus_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "manus_b200 kernels are written for sm_100a only"
#endif

namespace mb {

constexpr int kTile = 16;            // 16x16 pixel tiles (upstream BLOCK_X/BLOCK_Y)
constexpr int kTilePixels = 256;
constexpr float kNearZ = 0.2f;       // near cull, A.1 step 2
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kAlphaMax = 0.99f;
constexpr float kTMin = 0.0001f;
constexpr float kLowPass = 0.3f;     // px^2 added to the 2-D covariance diagonal

void set_error(const char *fmt, ...);
int check_launch(const char *what, bool debug_sync, cudaStream_t s);

#define MB_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            mb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return MB_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

#define MB_REQUIRE(cond, ...)                                                                   \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            mb::set_error(__VA_ARGS__);                                                         \
            return MB_ERR_INVALID;                                                              \
        }                                                                                       \
    } while (0)

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Carves consecutive 256-B aligned arrays out of one opaque byte buffer (same offsets in forward and backward).
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *p) : base(reinterpret_cast<char *>(p)) {}
    template <typename T>
    T *take(size_t count) {
        T *r = reinterpret_cast<T *>(base + off);
        off = align_up(off + count * sizeof(T));
        return r;
    }
};

int sm_count();

// Optional per-kernel timing with CUDA events on the launching stream (mb_profile_enable / mb_profile_report).
// Disabled by default: then the constructor / destructor do nothing.
struct KernelTimer {
    KernelTimer(const char *name, cudaStream_t s);
    ~KernelTimer();
    int slot;
    cudaStream_t stream;
};

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barriers visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk asynchronous copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fire-and-forget float add (RED.E.ADD.F32 in SASS)
__device__ __forceinline__ void red_add(float *addr, float v) { atomicAdd(addr, v); }

struct Mat3 {
    float m[9];  // row-major
};

// row-major standard rotation from a quaternion (r,x,y,z) used as given
__device__ __host__ __forceinline__ void quat_to_rot(float r, float x, float y, float z, float *R) {
    R[0] = 1.f - 2.f * (y * y + z * z);
    R[1] = 2.f * (x * y - r * z);
    R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);
    R[4] = 1.f - 2.f * (x * x + z * z);
    R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);
    R[7] = 2.f * (y * z + r * x);
    R[8] = 1.f - 2.f * (x * x + y * y);
}

// dL/dq for R = quat_to_rot(q) given dL/dR (row-major)
__device__ __host__ __forceinline__ void quat_to_rot_bwd(float r, float x, float y, float z, const float *dR, float *dq) {
    dq[0] = 2.f * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
    dq[1] = 2.f * (y * dR[1] + z * dR[2] + y * dR[3] - 2.f * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - 2.f * x * dR[8]);
    dq[2] = 2.f * (-2.f * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - 2.f * y * dR[8]);
    dq[3] = 2.f * (-2.f * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - 2.f * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
}

// real SH basis constants (src/utils/sh_utils.py:26-44)
#define MB_SH_C0 0.28209479177387814f
#define MB_SH_C1 0.4886025119029199f
#define MB_SH_C2_0 1.0925484305920792f
#define MB_SH_C2_1 -1.0925484305920792f
#define MB_SH_C2_2 0.31539156525252005f
#define MB_SH_C2_3 -1.0925484305920792f
#define MB_SH_C2_4 0.5462742152960396f
#define MB_SH_C3_0 -0.5900435899266435f
#define MB_SH_C3_1 2.890611442640554f
#define MB_SH_C3_2 -0.4570457994644658f
#define MB_SH_C3_3 0.3731763325901154f
#define MB_SH_C3_4 -0.4570457994644658f
#define MB_SH_C3_5 1.445305721320277f
#define MB_SH_C3_6 -0.5900435899266435f

// The 16 degree<=3 basis values at unit direction (x,y,z); entries above `deg` are left untouched.
__device__ __host__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float *b) {
    b[0] = MB_SH_C0;
    if (deg > 0) {
        b[1] = -MB_SH_C1 * y;
        b[2] = MB_SH_C1 * z;
        b[3] = -MB_SH_C1 * x;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = MB_SH_C2_0 * xy;
            b[5] = MB_SH_C2_1 * yz;
            b[6] = MB_SH_C2_2 * (2.0f * zz - xx - yy);
            b[7] = MB_SH_C2_3 * xz;
            b[8] = MB_SH_C2_4 * (xx - yy);
            if (deg > 2) {
                b[9] = MB_SH_C3_0 * y * (3.f * xx - yy);
                b[10] = MB_SH_C3_1 * xy * z;
                b[11] = MB_SH_C3_2 * y * (4.f * zz - xx - yy);
                b[12] = MB_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy);
                b[13] = MB_SH_C3_4 * x * (4.f * zz - xx - yy);
                b[14] = MB_SH_C3_5 * z * (xx - yy);
                b[15] = MB_SH_C3_6 * x * (xx - 3.f * yy);
            }
        }
    }
}

// d(basis_k)/d(x,y,z) for k < (deg+1)^2 : three arrays of 16
__device__ __host__ __forceinline__ void sh_basis_grad(int deg, float x, float y, float z, float *bx, float *by, float *bz) {
    bx[0] = by[0] = bz[0] = 0.f;
    if (deg > 0) {
        bx[1] = 0.f; by[1] = -MB_SH_C1; bz[1] = 0.f;
        bx[2] = 0.f; by[2] = 0.f; bz[2] = MB_SH_C1;
        bx[3] = -MB_SH_C1; by[3] = 0.f; bz[3] = 0.f;
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            bx[4] = MB_SH_C2_0 * y; by[4] = MB_SH_C2_0 * x; bz[4] = 0.f;
            bx[5] = 0.f; by[5] = MB_SH_C2_1 * z; bz[5] = MB_SH_C2_1 * y;
            bx[6] = MB_SH_C2_2 * -2.f * x; by[6] = MB_SH_C2_2 * -2.f * y; bz[6] = MB_SH_C2_2 * 4.f * z;
            bx[7] = MB_SH_C2_3 * z; by[7] = 0.f; bz[7] = MB_SH_C2_3 * x;
            bx[8] = MB_SH_C2_4 * 2.f * x; by[8] = MB_SH_C2_4 * -2.f * y; bz[8] = 0.f;
            if (deg > 2) {
                bx[9] = MB_SH_C3_0 * 6.f * xy; by[9] = MB_SH_C3_0 * 3.f * (xx - yy); bz[9] = 0.f;
                bx[10] = MB_SH_C3_1 * yz; by[10] = MB_SH_C3_1 * xz; bz[10] = MB_SH_C3_1 * xy;
                bx[11] = MB_SH_C3_2 * -2.f * xy; by[11] = MB_SH_C3_2 * (4.f * zz - xx - 3.f * yy); bz[11] = MB_SH_C3_2 * 8.f * yz;
                bx[12] = MB_SH_C3_3 * -6.f * xz; by[12] = MB_SH_C3_3 * -6.f * yz; bz[12] = MB_SH_C3_3 * 3.f * (2.f * zz - xx - yy);
                bx[13] = MB_SH_C3_4 * (4.f * zz - 3.f * xx - yy); by[13] = MB_SH_C3_4 * -2.f * xy; bz[13] = MB_SH_C3_4 * 8.f * xz;
                bx[14] = MB_SH_C3_5 * 2.f * xz; by[14] = MB_SH_C3_5 * -2.f * yz; bz[14] = MB_SH_C3_5 * (xx - yy);
                bx[15] = MB_SH_C3_6 * 3.f * (xx - yy); by[15] = MB_SH_C3_6 * -6.f * xy; bz[15] = 0.f;
            }
        }
    }
}

}  // namespace mb
