// The one exchange of the view-sharded step: SUM of the per-Gaussian gradients over the ranks, written
// for NVLink 5 / NVSwitch multicast memory instead of calling a collectives library.
//
// Every rank holds its gradient buffer at the same offset of a symmetric allocation that is also mapped through an NVSwitch
// MULTICAST address.  Rank r owns 1/R of every piece: for each 16 bytes of its share it issues ONE multimem.ld_reduce -- the
// switch reads the 16 bytes from all R GPUs, adds them in fp32 and returns the sum -- followed by ONE multimem.st, which the
// switch writes to all R GPUs.  Per GPU and direction ~(1 + 1/R) x the buffer crosses NVLink once (a ring all-reduce moves
// 2 (R-1)/R x over R-1 sequential steps), no staging buffers, no protocol flags: two instructions per 16 bytes.  Every element
// is reduced once, by the switch, and the same sum lands on every rank (bitwise identical replicas, which the replicated Adam
// step relies on).  The caller brackets the launch with cross-GPU barriers (all ranks have finished producing their gradients
// / all ranks' stores have landed); between them nobody else touches the pieces.
#include "common.cuh"

namespace mb {

constexpr int kMaxPieces = 8;

struct AllReduceArgs {
    float *mc;                        // multicast address of the symmetric buffer
    long long off[kMaxPieces];        // piece = [off, off + cnt) floats of the buffer
    long long cnt[kMaxPieces];
    int pieces, rank, world;
};

__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float *mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc_addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f32x4(float *mc_addr, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float multimem_ld_reduce_f32(const float *mc_addr) {
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc_addr) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st_f32(float *mc_addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc_addr), "f"(v) : "memory");
}

#ifndef MB_AR_UNROLL
#define MB_AR_UNROLL 4
#endif

__global__ void __launch_bounds__(512) multimem_allreduce_kernel(AllReduceArgs a) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < a.pieces; ++p) {
        float *base = a.mc + a.off[p];
        // the 16-byte-aligned body of the piece in units of float4, split evenly over the ranks; the (< 4 float) head and tail of a
        // piece that is not 16-byte aligned / sized go to rank 0 as scalars
        const long long head = (4 - ((a.off[p]) & 3)) & 3;
        const long long h = head < a.cnt[p] ? head : a.cnt[p];
        const long long n4 = (a.cnt[p] - h) >> 2, tail = a.cnt[p] - h - 4 * n4;
        const long long lo = n4 * a.rank / a.world, hi = n4 * (a.rank + 1) / a.world;
        float *body = base + h;
        long long i = lo + tid;
        // MB_AR_UNROLL independent reductions in flight per thread (each is a round trip through the switch)
        for (; i + (MB_AR_UNROLL - 1) * nthr < hi; i += MB_AR_UNROLL * nthr) {
            float4 v[MB_AR_UNROLL];
#pragma unroll
            for (int u = 0; u < MB_AR_UNROLL; ++u) v[u] = multimem_ld_reduce_f32x4(body + 4 * (i + u * nthr));
#pragma unroll
            for (int u = 0; u < MB_AR_UNROLL; ++u) multimem_st_f32x4(body + 4 * (i + u * nthr), v[u]);
        }
        for (; i < hi; i += nthr) multimem_st_f32x4(body + 4 * i, multimem_ld_reduce_f32x4(body + 4 * i));
        if (a.rank == 0 && tid < h + tail) {
            float *e = tid < h ? base + tid : body + 4 * n4 + (tid - h);
            multimem_st_f32(e, multimem_ld_reduce_f32(e));
        }
    }
}

// The same sum through plain peer-to-peer accesses (no multicast): rank r reads its 1/R share of every piece from all R buffers,
// adds them in rank order and writes the sum into all R buffers.  1.75 x the buffer per GPU and direction at R = 8 -- more than
// the in-switch reduction -- but it runs on the NVLink bandwidth the switch's reduction rate leaves idle, so the exchange can
// split every piece between the two kernels (mb_multimem_allreduce / mb_p2p_allreduce on two streams inside the same barriers).
struct P2PArgs {
    float *buf[8];                    // the symmetric buffer on every rank (peer pointers)
    long long off[kMaxPieces], cnt[kMaxPieces];
    int pieces, rank, world;
};

__global__ void __launch_bounds__(512) p2p_allreduce_kernel(P2PArgs a) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    for (int p = 0; p < a.pieces; ++p) {
        const long long n4 = a.cnt[p] >> 2;       // pieces handed to this kernel are 16-byte aligned and sized
        const long long lo = n4 * a.rank / a.world, hi = n4 * (a.rank + 1) / a.world;
        for (long long i = lo + tid; i < hi; i += nthr) {
            float4 v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < a.world) v[r] = __ldcg(reinterpret_cast<const float4 *>(a.buf[r] + a.off[p]) + i);
            float4 s = v[0];
#pragma unroll
            for (int r = 1; r < 8; ++r)
                if (r < a.world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < a.world) __stcg(reinterpret_cast<float4 *>(a.buf[r] + a.off[p]) + i, s);
        }
    }
}

}  // namespace mb

using namespace mb;

extern "C" int mb_p2p_allreduce(float *const *peer_buffers, const int64_t *piece_offsets, const int64_t *piece_counts, int32_t num_pieces,
                                int32_t rank, int32_t world, int32_t max_ctas, mb_stream_t stream) {
    MB_REQUIRE(peer_buffers != nullptr && world >= 1 && world <= 8 && rank >= 0 && rank < world, "mb_p2p_allreduce: 1..8 ranks");
    MB_REQUIRE(num_pieces >= 1 && num_pieces <= kMaxPieces && piece_offsets && piece_counts, "mb_p2p_allreduce: 1..%d pieces", kMaxPieces);
    P2PArgs a;
    a.pieces = num_pieces; a.rank = rank; a.world = world;
    for (int r = 0; r < 8; ++r) a.buf[r] = r < world ? peer_buffers[r] : nullptr;
    long long most = 0;
    for (int p = 0; p < num_pieces; ++p) {
        MB_REQUIRE(piece_offsets[p] >= 0 && piece_counts[p] >= 0 && (piece_offsets[p] & 3) == 0 && (piece_counts[p] & 3) == 0,
                   "mb_p2p_allreduce: pieces must start and end on 16-byte boundaries");
        a.off[p] = piece_offsets[p]; a.cnt[p] = piece_counts[p];
        most = piece_counts[p] > most ? piece_counts[p] : most;
    }
    if (most == 0) return MB_OK;
    long long grid = (most / 4 / world + 511) / 512;
    const int cap = max_ctas > 0 ? max_ctas : sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    cudaStream_t s = (cudaStream_t)stream;
    KernelTimer kt("p2p_allreduce", s);
    p2p_allreduce_kernel<<<(unsigned)grid, 512, 0, s>>>(a);
    return check_launch("p2p_allreduce", false, s);
}

extern "C" int mb_multimem_allreduce(float *multicast_base, const int64_t *piece_offsets, const int64_t *piece_counts, int32_t num_pieces,
                                     int32_t rank, int32_t world, int32_t max_ctas, mb_stream_t stream) {
    MB_REQUIRE(multicast_base != nullptr, "mb_multimem_allreduce: no multicast address (NVSwitch multicast is not available for this allocation)");
    MB_REQUIRE(num_pieces >= 1 && num_pieces <= kMaxPieces && piece_offsets && piece_counts, "mb_multimem_allreduce: 1..%d pieces", kMaxPieces);
    MB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "mb_multimem_allreduce: bad rank %d of %d", rank, world);
    AllReduceArgs a;
    a.mc = multicast_base; a.pieces = num_pieces; a.rank = rank; a.world = world;
    long long most = 0;
    for (int p = 0; p < num_pieces; ++p) {
        MB_REQUIRE(piece_offsets[p] >= 0 && piece_counts[p] >= 0, "mb_multimem_allreduce: negative piece");
        a.off[p] = piece_offsets[p]; a.cnt[p] = piece_counts[p];
        most = piece_counts[p] > most ? piece_counts[p] : most;
    }
    if (most == 0) return MB_OK;
    // enough CTAs to keep the links busy, few enough to leave the SMs to the kernels this exchange runs beside
    const long long share4 = most / 4 / world + 1;
    long long grid = (share4 + 512 * MB_AR_UNROLL - 1) / (512 * MB_AR_UNROLL);
    const int cap = max_ctas > 0 ? max_ctas : sm_count();
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    cudaStream_t s = (cudaStream_t)stream;
    KernelTimer kt("multimem_allreduce", s);
    multimem_allreduce_kernel<<<(unsigned)grid, 512, 0, s>>>(a);
    return check_launch("multimem_allreduce", false, s);
}
