// Dispatch-rate probe: how many cycles does one tcgen05.mma (M = 128) take when many are issued back to back by one thread
// (and by two warps at once) -- kind::i8 (K = 32) and kind::mxf4 (K = 64), A operand from tensor memory (TS) or from shared
// memory (SS), N = 16 .. 256.  The operands are whatever the memories hold: only the timing matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I qgtc_ppopp22_b200/csrc -I include -o mma_rate_probe mma_rate_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "tc05_common.cuh"
using namespace qgtc::tc05;

__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}

// mode 0: i8 TS, 1: i8 SS, 2: mxf4 TS.  issuers = 1 or 2 warps, each issuing `reps` MMAs into its own accumulator.
__global__ void __launch_bounds__(128) probe(int mode, int N, int reps, int issuers, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    {   // fill the A / scale columns (256..511) with something harmless: 0x7F bytes are unit UE8M0 scales and small E2M1 / u8 values
        uint32_t r[32];
        for (int e = 0; e < 32; ++e) r[e] = 0x7F7F7F7Fu;
        for (int c = 256; c < 512; c += 32) tmem_st32(tb + ((uint32_t)(32 * warp) << 16) + (uint32_t)c, r);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp < issuers) {
        const uint32_t idesc = mode == 2 ? make_idesc_f4(N) : make_idesc_i8(N, 128);
        const uint32_t lbo = (uint32_t)(N / 8) * 128u, sbo = 128u;
        const uint64_t bdesc = make_smem_desc(smem_u32(smem), lbo, sbo);
        const uint64_t adesc = make_smem_desc(smem_u32(smem + 32768), 16u * 128u, 128u);
        const uint32_t d_tmem = tb + (uint32_t)(warp * 128 < 256 - N ? warp * 128 : 0);
        const uint32_t a_tmem = tb + 256u, sf = tb + 480u;
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < reps; i += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t a = a_tmem + (uint32_t)(8 * u);
                    if (mode == 0) mma_i8_ts(d_tmem, a, bdesc, idesc, 1u);
                    else if (mode == 1) mma_i8_ss(d_tmem, adesc, bdesc, idesc);
                    else mma_f4_ts(d_tmem, a, bdesc, idesc, sf, sf + 4u, 1u);
                }
            }
            t1 = clock64();           // issue time
            tc_commit(&bar[warp]);
        }
        __syncwarp();
        mbar_wait(&bar[warp], 0, nullptr, 0, false);
        const long long t2 = clock64();   // completion time
        if (lane == 0) {
            // only the elected lane has t0 / t1: broadcast below
        }
        const long long tt0 = __shfl_sync(0xffffffffu, t0, __ffs(__ballot_sync(0xffffffffu, t0 != 0)) - 1);
        const long long tt1 = __shfl_sync(0xffffffffu, t1, __ffs(__ballot_sync(0xffffffffu, t1 != 0)) - 1);
        if (lane == 0) {
            out[warp * 2 + 0] = tt1 - tt0;
            out[warp * 2 + 1] = t2 - tt0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tb, 512);
    }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[3] = {"i8   TS (A in TMEM)", "i8   SS (A in smem)", "mxf4 TS (A in TMEM)"};
    const int reps = 2048;
    for (int issuers = 1; issuers <= 2; ++issuers)
        for (int mode = 0; mode < 3; ++mode)
            for (int N : {16, 32, 64, 128, 256}) {
                if (issuers == 2 && N > 128) continue;   // two accumulators of N columns in the first 256 columns
                for (int rep = 0; rep < 2; ++rep) {
                    out[0] = out[1] = out[2] = out[3] = 0;
                    probe<<<1, 128, 64 * 1024>>>(mode, N, reps, issuers, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                printf("%s  N=%3d  issuers=%d  cycles/MMA: issue %.1f  complete %.1f", names[mode], N, issuers, (double)out[0] / reps, (double)out[1] / reps);
                if (issuers == 2) printf("   (warp 1: issue %.1f complete %.1f; per MMA of both: %.1f)", (double)out[2] / reps, (double)out[3] / reps, (double)(out[1] > out[3] ? out[1] : out[3]) / (2.0 * reps));
                printf("\n");
            }
    return 0;
}
