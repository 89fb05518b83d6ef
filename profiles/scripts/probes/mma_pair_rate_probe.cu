// Dispatch-rate probe, 2-CTA form: tcgen05.mma.cta_group::2 (M = 256: 128 rows in each CTA of a cluster of two), issued by the leader
// CTA, A operand in tensor memory, kind::i8 and kind::mxf4 (block-scaled), N = 16 .. 128.  Question: is the ~42-cycle dispatch floor of
// kind::mxf4 at N <= 64 (profiles/r2_mma_dispatch_rate_call20.txt) per instruction (then a pair does twice the rows per dispatch) or
// per 128 rows?  Operand contents do not matter: only the timing does.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I qgtc_ppopp22_b200/csrc -I include -o mma_pair_rate_probe mma_pair_rate_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "tc05_common.cuh"
using namespace qgtc::tc05;

__device__ __forceinline__ void mma_f4_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_f4_m(int n, int m) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | (((uint32_t)m >> 4) << 24);
}

// mode 0: i8 pair, 1: mxf4 pair.  out[0] = issue cycles, out[1] = cycles until the commit arrived (leader CTA)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) probe(int mode, int N, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0) tmem_alloc_pair(&tmem_slot, 512);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    {
        uint32_t r[32];
        for (int e = 0; e < 32; ++e) r[e] = 0x7F7F7F7Fu;
        for (int c = 256; c < 512; c += 32) tmem_st32(tb + ((uint32_t)(32 * warp) << 16) + (uint32_t)c, r);
        tmem_wait_st();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (rank == 0 && warp == 0) {
        const uint32_t idesc = mode == 1 ? make_idesc_f4_m(N, 256) : make_idesc_i8(N, 256);
        // each CTA holds N / 2 columns of B: K-direction stride (N / 16) * 128 bytes
        const uint32_t lbo = (uint32_t)(N / 16) * 128u, sbo = 128u;
        const uint64_t bdesc = make_smem_desc(smem_u32(smem), lbo, sbo);
        const uint32_t d_tmem = tb, a_tmem = tb + 256u, sf = tb + 480u;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < reps; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t a = a_tmem + (uint32_t)(8 * u);
                    if (mode == 0) mma_i8_ts_pair(d_tmem, a, bdesc, idesc, 1u);
                    else mma_f4_ts_pair(d_tmem, a, bdesc, idesc, sf, sf + 4u);
                }
            }
            t1 = clock64();
            tc_commit_pair(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, nullptr, 0, false);
        const long long t2 = clock64();
        const long long tt0 = __shfl_sync(0xffffffffu, t0, __ffs(__ballot_sync(0xffffffffu, t0 != 0)) - 1);
        const long long tt1 = __shfl_sync(0xffffffffu, t1, __ffs(__ballot_sync(0xffffffffu, t1 != 0)) - 1);
        if (lane == 0) {
            out[0] = tt1 - tt0;
            out[1] = t2 - tt0;
        }
    } else if (rank == 1 && warp == 0) {
        mbar_wait(&bar, 0, nullptr, 0, false);      // the multicast commit arrives here too
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_pair(tb, 512);
    }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[2] = {"i8   cta_group::2 (M = 256)", "mxf4 cta_group::2 (M = 256)"};
    const int reps = 2048;
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {16, 32, 64, 128, 256}) {
            for (int rep = 0; rep < 2; ++rep) {
                out[0] = out[1] = 0;
                probe<<<2, 128, 64 * 1024>>>(mode, N, reps, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s N=%d error %s\n", names[mode], N, cudaGetErrorString(e)); return 1; }
            }
            printf("%s  N=%3d  cycles/MMA: issue %.1f  complete %.1f   (per 128 rows: %.1f)\n", names[mode], N, (double)out[0] / reps, (double)out[1] / reps, (double)out[1] / reps / 2.0);
        }
    return 0;
}
