// Bandwidth probes: what can a plain kernel get out of HBM on this GPU for a read-only stream, a write-only stream and a copy?
// Built by tools/bw_probe.py into tools/probe/libbw_probe.so (experiment tooling, not part of the product library).
#include <cuda_runtime.h>
#include <stdint.h>

__global__ void __launch_bounds__(256) read_kernel(const float4 *__restrict__ src, size_t n, float *sink) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}

__global__ void __launch_bounds__(256) write_kernel(float4 *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

__global__ void __launch_bounds__(256) copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// each CTA streams its own contiguous chunk (like a tile pipeline) instead of the interleaved grid-stride pattern
__global__ void __launch_bounds__(256) read_chunked_kernel(const float4 *__restrict__ src, size_t n, size_t chunk, float *sink) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t c = blockIdx.x; c * chunk < n; c += gridDim.x) {
        const size_t lo = c * chunk, hi = lo + chunk < n ? lo + chunk : n;
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const float4 v = src[i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) *sink = acc.x;
}

extern "C" int bw_probe(int mode, void *a, void *b, size_t bytes, int grid, size_t chunk_bytes, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = bytes / 16;
    if (mode == 0) read_kernel<<<grid, 256, 0, s>>>((const float4 *)a, n, (float *)b);
    else if (mode == 1) write_kernel<<<grid, 256, 0, s>>>((float4 *)a, n);
    else if (mode == 2) copy_kernel<<<grid, 256, 0, s>>>((const float4 *)a, (float4 *)b, n);
    else read_chunked_kernel<<<grid, 256, 0, s>>>((const float4 *)a, n, chunk_bytes / 16, (float *)b);
    return (int)cudaGetLastError();
}
