// Launch-floor probe: what does one launch of a persistent 148-CTA kernel cost before any work, in a CUDA graph of 50 dependent launches?
//   small / big     : 64-byte vs 28 KB of kernel parameters (the grouped engine passes its problem table by value: GrpArgs ~ 29 KB)
//   empty / setup   : no body vs the grouped engine's set-up (576 threads, 200 KB dynamic shared memory, mbarrier init, 512 TMEM columns)
//   plain / pdl     : stream order vs programmatic dependent launch (launch_dependents at the top, griddepcontrol.wait before the exit)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I qgtc_ppopp22_b200/csrc -I include -o launch_floor_probe launch_floor_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "tc05_common.cuh"
using namespace qgtc::tc05;

struct Small { int* out; int n; int pad[13]; };
struct Big { int* out; int n; int tab[7200]; };

template <typename P, bool SETUP, bool PDL>
__global__ void __launch_bounds__(576, 1) probe(const __grid_constant__ P p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bars[16];
    if (PDL) asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    int v = 0;
    if constexpr (sizeof(P) > 1024) v = p.tab[(blockIdx.x * 37 + threadIdx.x) % 7200];
    if (SETUP) {
        const int warp = threadIdx.x >> 5;
        if (threadIdx.x == 0) {
            for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
            fence_mbar_init();
        }
        if (warp == 16) tmem_alloc(&tmem_slot, 512);
        smem[threadIdx.x] = (unsigned char)v;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        v += smem[(threadIdx.x + 1) % 576] + (int)tmem_slot;
        __syncthreads();
        if (warp == 16) tmem_dealloc(tmem_slot, 512);
    }
    if (PDL) asm volatile("griddepcontrol.wait;\n" ::: "memory");
    if (v == 0x7fffffff) p.out[0] = v;
}

template <typename P, bool SETUP, bool PDL>
static float run(const char* name, int* out) {
    P p{};
    p.out = out;
    p.n = 1;
    const int smem_bytes = SETUP ? 200 * 1024 : 0;
    auto k = probe<P, SETUP, PDL>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(576);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = PDL ? 1 : 0;
    const int L = 50;
    for (int i = 0; i < 3; ++i) cudaLaunchKernelEx(&cfg, k, p);
    cudaStreamSynchronize(st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < L; ++i) cudaLaunchKernelEx(&cfg, k, p);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const float us = ms * 1e3f / (10 * L);
    printf("%-28s %6.2f us per launch  (%s)\n", name, us, cudaGetErrorString(cudaGetLastError()));
    return us;
}

int main() {
    int* out;
    cudaMalloc(&out, 64);
    run<Small, false, false>("small params, empty, plain", out);
    run<Big, false, false>("28 KB params, empty, plain", out);
    run<Small, true, false>("small params, setup, plain", out);
    run<Big, true, false>("28 KB params, setup, plain", out);
    run<Small, false, true>("small params, empty, pdl", out);
    run<Big, false, true>("28 KB params, empty, pdl", out);
    run<Small, true, true>("small params, setup, pdl", out);
    run<Big, true, true>("28 KB params, setup, pdl", out);
    return 0;
}
