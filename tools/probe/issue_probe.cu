// Issue-rate probes for the blend kernels' instruction mix on sm_100a: how many warp instructions per clock and SM do
// FFMA, packed FFMA2, FSEL, SHFL and MUFU sustain, alone and interleaved?  (experiment tooling, not part of the product)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, float seed) {
    float a[8], s = seed, t = seed * 0.5f;
    u64 p[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = seed + k + threadIdx.x; p[k] = pack(a[k], a[k] + 1.f); }
    const u64 ps = pack(s, s), pt = pack(t, t);
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) a[k] = fmaf(a[k], s, t);                                      // 8 FFMA
            if (MODE == 1) p[k] = ffma2(p[k], ps, pt);                                   // 8 FFMA2
            if (MODE == 2) { a[k] = fmaf(a[k], s, t); a[k] = a[k] > 1.f ? a[k] : t; }    // FFMA + FSETP+FSEL (or FMNMX)
            if (MODE == 3) { p[k] = ffma2(p[k], ps, pt); a[k] = (i & (1 << k)) ? a[k] : t; }   // FFMA2 + select
            if (MODE == 4) a[k] = __shfl_xor_sync(0xffffffffu, a[k], 1 + k);             // 8 SHFL
            if (MODE == 5) { a[k] = fmaf(a[k], s, t); a[k] = __shfl_xor_sync(0xffffffffu, a[k], 16); }   // FFMA + SHFL
            if (MODE == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));      // 8 MUFU
            if (MODE == 7) { a[k] = a[k] * s; }                                          // 8 FMUL
            if (MODE == 8) { a[k] = a[k] + s; }                                          // 8 FADD
            if (MODE == 9) { a[k] = fmaf(a[k], s, t); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(k + 4) & 7])); }
            if (MODE == 10) { a[k] = fmaf(a[k], s, t); a[k] = __int_as_float(__float_as_int(a[k]) + lane); }   // FFMA + IADD
        }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += a[k] + __uint_as_float((uint32_t)(p[k] >> 32)) + __uint_as_float((uint32_t)p[k]);
    if (r == 123.456f) out[0] = r;
}

static const char *names[] = {"FFMA", "FFMA2", "FFMA+FSETP/FSEL", "FFMA2+LOP/SEL", "SHFL", "FFMA+SHFL", "MUFU.EX2", "FMUL", "FADD", "FFMA+MUFU", "FFMA+IADD"};

template <int MODE>
static void run(float *out, int sms, float mhz, int per_iter) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int ctas = sms * 8;
    probe<MODE><<<ctas, 256>>>(out, 1.0001f);
    cudaEventRecord(e0);
    probe<MODE><<<ctas, 256>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)ctas * 8 * ITERS * per_iter;
    const double clocks = ms * 1e-3 * mhz * 1e6;
    printf("%-18s %7.3f ms  %6.2f listed warp-instr / clk / SM\n", names[MODE], ms, warp_instr / clocks / sms);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *out;
    cudaMalloc(&out, 4);
    const float mhz = p.clockRate / 1000.0f;
    printf("%s: %d SMs, %.0f MHz (max)\n", p.name, p.multiProcessorCount, mhz);
    const int sms = p.multiProcessorCount;
    run<0>(out, sms, mhz, 8); run<1>(out, sms, mhz, 8); run<2>(out, sms, mhz, 16); run<3>(out, sms, mhz, 16);
    run<4>(out, sms, mhz, 8); run<5>(out, sms, mhz, 16); run<6>(out, sms, mhz, 8); run<7>(out, sms, mhz, 8);
    run<8>(out, sms, mhz, 8); run<9>(out, sms, mhz, 16); run<10>(out, sms, mhz, 16);
    return 0;
}
