// Fused Adam step over the flat parameter buffer.
//
// The reference optimises six tensors with one torch.optim.Adam(lr=0, eps=1e-15) whose param groups differ only in their
// learning rate (src/models/gaussian.py:133-141: xyz, f_dc, f_rest, opacity, scaling, rotation) -- per step that is ~50
// elementwise / foreach kernels.  Here parameters, gradients and both moments live in flat fp32 buffers with one contiguous
// segment per group (manus_b200.dist.FlatGaussians), so the whole step is ONE kernel: 28 B of HBM traffic per scalar
// (read p, g, m, v; write p, m, v), each touched once.  The arithmetic follows torch's Adam operation by operation
// (lerp for the first moment, mul + addcmul for the second, sqrt / bias-correction, addcdiv), so results agree with
// torch.optim.Adam to fp32 rounding.  A [begin, end) element range lets a rank update only its shard (ZeRO-1 style
// data parallelism: reduce-scatter -> sharded Adam -> all-gather).
#include "common.cuh"

namespace mb {

constexpr int kAdamMaxSegments = 8;

struct AdamArgs {
    float *param, *exp_avg, *exp_avg_sq;
    const float *grad;
    int64_t begin, end;                      // element range of this call inside the flat buffer
    int num_segments;
    int64_t seg_end[kAdamMaxSegments];       // exclusive end offset of every segment (ascending)
    float step_size[kAdamMaxSegments];       // lr / (1 - beta1^t) per segment
    float w1, beta2, w2, eps, bc2_sqrt, grad_scale;   // w1 = 1 - beta1, w2 = 1 - beta2 (rounded from double like torch's scalars)
};

__device__ __forceinline__ void adam_one(float &p, float &m, float &v, float g, float step_size, const AdamArgs &a) {
    g *= a.grad_scale;
    m = m + a.w1 * (g - m);                                  // exp_avg.lerp_(grad, 1 - beta1)  (weight < 0.5 branch of lerp)
    v = v * a.beta2 + (a.w2 * g) * g;                       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;       // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p + (-step_size) * (m / denom);                      // param.addcdiv_(exp_avg, denom, value = -step_size)
}

__global__ void __launch_bounds__(256) fused_adam_kernel(AdamArgs a) {
    // 4 consecutive scalars per thread (128-bit accesses when the range start is 16-B aligned, scalar otherwise)
    const int64_t first = a.begin + ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (first >= a.end) return;
    int seg = 0;
    while (seg < a.num_segments - 1 && first >= a.seg_end[seg]) ++seg;
    const bool vec = first + 4 <= a.end && first + 4 <= a.seg_end[seg] && (first & 3) == 0;
    if (vec) {
        float4 p = *reinterpret_cast<float4 *>(a.param + first), m = *reinterpret_cast<float4 *>(a.exp_avg + first);
        float4 v = *reinterpret_cast<float4 *>(a.exp_avg_sq + first);
        const float4 g = *reinterpret_cast<const float4 *>(a.grad + first);
        const float ss = a.step_size[seg];
        adam_one(p.x, m.x, v.x, g.x, ss, a);
        adam_one(p.y, m.y, v.y, g.y, ss, a);
        adam_one(p.z, m.z, v.z, g.z, ss, a);
        adam_one(p.w, m.w, v.w, g.w, ss, a);
        *reinterpret_cast<float4 *>(a.param + first) = p;
        *reinterpret_cast<float4 *>(a.exp_avg + first) = m;
        *reinterpret_cast<float4 *>(a.exp_avg_sq + first) = v;
    } else {
        for (int64_t i = first; i < first + 4 && i < a.end; ++i) {
            while (seg < a.num_segments - 1 && i >= a.seg_end[seg]) ++seg;
            float p = a.param[i], m = a.exp_avg[i], v = a.exp_avg_sq[i];
            adam_one(p, m, v, a.grad[i], a.step_size[seg], a);
            a.param[i] = p; a.exp_avg[i] = m; a.exp_avg_sq[i] = v;
        }
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_fused_adam(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t begin, int64_t end,
                             int32_t num_segments, const int64_t *segment_end_host, const double *lr_host, int64_t step, double beta1,
                             double beta2, double eps, float grad_scale, mb_stream_t stream) {
    MB_REQUIRE(param && grad && exp_avg && exp_avg_sq, "mb_fused_adam: null buffer");
    MB_REQUIRE(num_segments >= 1 && num_segments <= kAdamMaxSegments && segment_end_host && lr_host, "mb_fused_adam: 1..%d segments",
               kAdamMaxSegments);
    MB_REQUIRE(step >= 1 && begin >= 0 && end >= begin, "mb_fused_adam: bad step / range");
    if (end == begin) return MB_OK;
    AdamArgs a;
    a.param = param; a.grad = grad; a.exp_avg = exp_avg; a.exp_avg_sq = exp_avg_sq;
    a.begin = begin; a.end = end; a.num_segments = num_segments;
    // bias corrections exactly as torch computes them on the host (Python floats = doubles)
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    for (int s = 0; s < num_segments; ++s) {
        MB_REQUIRE(s == 0 || segment_end_host[s] >= segment_end_host[s - 1], "mb_fused_adam: segment ends must ascend");
        a.seg_end[s] = segment_end_host[s];
        a.step_size[s] = (float)(lr_host[s] / bc1);
    }
    a.w1 = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.w2 = (float)(1.0 - beta2); a.eps = (float)eps;
    a.bc2_sqrt = (float)sqrt(bc2); a.grad_scale = grad_scale;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t threads = (end - begin + 3) / 4;
    KernelTimer kt("fused_adam", s);
    fused_adam_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(a);
    return check_launch("fused_adam", false, s);
}
