// Probe for the next-round 4-bit operand path: one CTA, tcgen05.mma kind::mxf4 (E2M1 x E2M1 -> f32, unit UE8M0 block scales),
// A operand from tensor memory (TS form), B from shared memory (K-major, no swizzle).  Finds out (1) whether the
// instruction descriptor / scale-factor set-up below computes plain sums of products, and (2) how the 64 K positions
// of the A operand in TMEM (8 columns x 8 nibbles) pair with the 64 K positions of the B operand in shared memory
// (2 core matrices x 16 bytes x 2 nibbles).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o mxf4_probe mxf4_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void st1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};\n" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(v) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    return v;
}

constexpr int N = 64;
// out[test][m][n] floats.  test 0: all ones (expect 64).  tests 1 + 6 * a + t (a = A position 0..63, t = bit 0..5): A has 1.0 only at
// position a, B has 1.0 at the positions p with bit t of p set -> out = bit t of the B position paired with A position a.
__global__ void __launch_bounds__(128) probe(float* out, int ntests) {
    __shared__ __align__(1024) uint8_t bsm[2 * (N / 8) * 128];   // [kgroup 2][n/8][8 rows][16 B]
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_slot;
    const uint32_t lane_base = tb + ((uint32_t)(32 * warp) << 16);
    const uint32_t d_col = 0, a_col = 64, sf_col = 96;
    // unit scale factors (UE8M0 127 = 2^0) everywhere in 8 columns
    for (int c = 0; c < 8; ++c) st1(lane_base + sf_col + c, 0x7F7F7F7Fu);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    // idesc: a_format = b_format = 1 (MXF4 E2M1), K-major both, n_dim = N >> 3, scale_format = 1 (UE8M0), m_dim = 128 >> 4, k_size = 0 (K = 64)
    const uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const uint64_t bdesc = make_smem_desc(smem_u32(bsm), (N / 8) * 128, 128);
    uint32_t phase = 0;
    for (int test = 0; test < ntests; ++test) {
        const int a_pos = test == 0 ? -1 : (test - 1) / 6, tbit = test == 0 ? -1 : (test - 1) % 6;
        // A: 8 columns per lane
        for (int c = 0; c < 8; ++c) {
            uint32_t v = 0;
            if (test == 0) v = 0x22222222u;                                   // 1.0 in every nibble
            else if (a_pos / 8 == c) v = 0x2u << (4 * (a_pos % 8));          // 1.0 at nibble (a_pos % 8) of column c
            st1(lane_base + a_col + c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        // B: position p = kg * 32 + byte * 2 + nib
        for (int i = threadIdx.x; i < (int)sizeof(bsm); i += blockDim.x) {
            const int kg = i / ((N / 8) * 128), byte = i % 16;
            uint8_t v = 0;
            for (int nib = 0; nib < 2; ++nib) {
                const int p = kg * 32 + byte * 2 + nib;
                const bool on = test == 0 ? true : ((p >> tbit) & 1);
                if (on) v |= (uint8_t)(0x2u << (4 * nib));
            }
            bsm[i] = v;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (threadIdx.x == 0) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %6, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n\t}\n"
                         ::"r"(tb + d_col), "r"(tb + a_col), "l"(bdesc), "r"(idesc), "r"(tb + sf_col), "r"(tb + sf_col + 4), "r"(0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        }
        // wait for the MMA
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        tc_fence_after();
        for (int n = 0; n < N; ++n) {
            const uint32_t v = ld1(lane_base + d_col + n);
            out[((size_t)test * 128 + (32 * warp + lane)) * N + n] = __uint_as_float(v);
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(128u) : "memory");
}

int main() {
    const int ntests = 1 + 64 * 6;
    float* d;
    cudaMalloc(&d, sizeof(float) * ntests * 128 * N);
    cudaMemset(d, 0xFF, sizeof(float) * ntests * 128 * N);
    probe<<<1, 128>>>(d, ntests);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    float* h = new float[(size_t)ntests * 128 * N];
    cudaMemcpy(h, d, sizeof(float) * ntests * 128 * N, cudaMemcpyDeviceToHost);
    printf("test 0 (all ones, expect 64): D[0][0]=%g D[5][7]=%g D[127][63]=%g\n", h[0], h[5 * N + 7], h[127 * N + 63]);
    bool uniform = true;
    for (int i = 0; i < 128 * N; ++i) uniform &= h[i] == h[0];
    printf("test 0 uniform over the tile: %d\n", (int)uniform);
    printf("A position (column c, nibble i) -> paired B position (kgroup*32 + byte*2 + nib); value seen\n");
    for (int a = 0; a < 64; ++a) {
        int p = 0; float vmax = 0;
        bool consistent = true;
        for (int t = 0; t < 6; ++t) {
            const float* o = h + (size_t)(1 + a * 6 + t) * 128 * N;
            const float v = o[3 * N + 9];
            for (int i = 0; i < 128 * N; ++i) consistent &= o[i] == v;
            if (v != 0.f) p |= 1 << t;
            if (v > vmax) vmax = v;
        }
        printf("  A(c=%d,i=%d) -> B %2d  (value %g, uniform %d)%s", a / 8, a % 8, p, vmax, (int)consistent, (a % 4 == 3) ? "\n" : "");
    }
    return 0;
}
