// Bit-GEMM engine 2: tcgen05.mma kind::i8 (u8 x u8 -> s32) with TMEM accumulators, sm_100a.
//
// Data flow of one CTA (persistent, one per SM), for a 128-row adjacency tile:
//
//   packed 1-bit adjacency rows --cp.async--> smem ring --LDS.128--> registers
//        --(w >> c) & 0x01010101, 2 ALU ops per 4 bytes--> 32 registers = 128 u8 of one row
//        --tcgen05.st.32x32b.x32--> TMEM  (the MMA's A operand, one row per TMEM lane)
//   feature "image" (u8, K-major core matrices, produced once per call by prep_b_kernel)
//        --cp.async.bulk + mbarrier--> smem stage (the MMA's B operand, no-swizzle descriptor)
//   tcgen05.mma.cta_group::1.kind::i8  D[tmem] (+)= A[tmem] x B[smem], 4 per 128-bit K chunk
//   tcgen05.ld --> requantize + pack planes (or float) in registers --> coalesced global stores
//
// The adjacency never exists as bytes in shared or global memory: 1 bit per element is read from
// HBM, the 8x inflation happens in registers and goes straight to tensor memory.  The K order
// inside a 32-element step is permuted (byte i of column c is bit c+8i of the packed word) so the
// expansion is one shift and one mask per register; the feature image uses the same permutation.
//
// Warp roles: warps 0-7 adjacency producers, 8-11 epilogue, 12 feature loader, 13 MMA issuer.
// Zero-tile jumping: a producer warp whose 32 rows x 128 bits are all zero skips the expansion and
// reports it; if the whole tile chunk is zero the MMA issuer skips its 4 MMAs.
#include "gemm_common.cuh"

namespace qgtc {

namespace {

constexpr int kProducerWarps = 8;
constexpr int kEpilogueWarp0 = 8;
constexpr int kLoaderWarp = 12;
constexpr int kMmaWarp = 13;
constexpr int kThreads = 14 * 32;
constexpr int kADepth = 4;        // cp.async ring depth per producer thread (in its own items)
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;

struct Tc05Args {
    GemmArgs g;
    const uint8_t* bimg;            // feature image: [count][gdk][NP x 128 B]
    unsigned long long bimg_gstride;  // bytes per problem
    int NP;                         // padded N (multiple of 16, <= 256)
    int stages;                     // A (TMEM) / B (smem) ring depth
    int dbufs;                      // accumulator buffers (1 or 2)
    int mtiles;                     // 128-row tiles per problem
    int total_tiles;                // count * mtiles
    int epi;
    unsigned int* error_flag;       // device word set by the watchdog
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  ~2^26 polls is far beyond any legitimate wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* error_flag, int where) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            if (error_flag) atomicExch(error_flag, 0x80000000u | (unsigned)where);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], kind::i8
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16_ca_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start address,
// leading (K-direction) and stride (N-direction) byte offsets between 8x16-byte core matrices, in
// 16-byte units; version 1 (Blackwell) in bits 46-47.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = u8, both K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc_i8(int n) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

struct SmemLayout {
    // offsets in bytes from the 128-byte aligned base
    int b_stage0, b_stage_bytes, a_ring, bars, flags, misc, total;
};
__host__ __device__ inline SmemLayout smem_layout(int NP, int stages) {
    SmemLayout L;
    L.b_stage0 = 0;
    L.b_stage_bytes = NP * 128;
    L.a_ring = L.b_stage0 + stages * L.b_stage_bytes;
    L.bars = L.a_ring + kProducerWarps * kADepth * 32 * 16;
    L.flags = L.bars + (3 * kMaxStages + 4) * 8;
    L.misc = L.flags + kMaxStages * 8;
    L.total = L.misc + 32;
    return L;
}

// One (tile, chunk) item of the persistent schedule
struct Item {
    int tl;  // CTA-local tile counter
    int c;   // K chunk
};

// ---------------------------------------------------------------------------------------------
// Feature image: X2 (packed planes, raw words, zero-extension) -> u8 values in the MMA's K order,
// laid out as the exact shared-memory image of each 128-element K chunk (core matrices of 8 rows
// x 16 bytes; K-direction stride (NP/8)*128 B, N-direction stride 128 B).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_b_kernel(GemmArgs P, uint8_t* __restrict__ img, unsigned long long gstride, int NP) {
    const int gdk = P.gdk, n_lim = p8(P.N);
    const long long per_problem = (long long)gdk * NP * 4;  // (chunk, n, s)
    const long long total = per_problem * P.count;
    const unsigned long long gdk4 = (unsigned long long)gdk * 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(i / per_problem);
        const long long r = i % per_problem;
        const int s = (int)(r & 3), n = (int)((r >> 2) % NP), c = (int)((r >> 2) / NP);
        const Words x2{P.x2 + (unsigned long long)z * P.x2_bstride, P.x2_words};
        uint32_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (n < n_lim) {
            for (int q = 0; q < P.bit2; ++q) {
                const uint32_t w = ldw(x2, (unsigned long long)q * P.w_offset + (unsigned long long)n * gdk4 + 4ull * c + s);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] |= ((w >> j) & 0x01010101u) << q;
            }
        }
        uint8_t* dst = img + (unsigned long long)z * gstride + (unsigned long long)c * NP * 128 +
                       (unsigned long long)(2 * s) * (NP / 8) * 128 + (unsigned long long)(n >> 3) * 128 + (n & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(dst + (NP / 8) * 128) = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// The GEMM kernel
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc05_kernel(const Tc05Args T) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GemmArgs& P = T.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NP = T.NP, S = T.stages, gdk = P.gdk;
    const SmemLayout L = smem_layout(NP, S);

    unsigned char* b_stages = smem_raw + L.b_stage0;
    uint4* a_ring = reinterpret_cast<uint4*>(smem_raw + L.a_ring);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
    uint64_t* a_full = bars;                    // [kMaxStages] producers -> MMA      (count 4 warps)
    uint64_t* b_full = bars + kMaxStages;       // [kMaxStages] bulk copy -> MMA      (count 1 + tx)
    uint64_t* empty = bars + 2 * kMaxStages;    // [kMaxStages] MMA commit -> producers, loader
    uint64_t* d_full = bars + 3 * kMaxStages;   // [2] MMA -> epilogue                (count 2: commit + plain arrive)
    uint64_t* d_empty = d_full + 2;             // [2] epilogue -> MMA                (count 4 warps)
    volatile unsigned char* flags = smem_raw + L.flags;  // [stage][8]: per producer warp "chunk is non-zero"
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + L.misc);
    volatile uint32_t* d_valid = reinterpret_cast<volatile uint32_t*>(smem_raw + L.misc + 8);  // [2]

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&a_full[s], 4);
            mbar_init(&b_full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 2);
            mbar_init(&d_empty[b], 4);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int dcols = T.dbufs * NP;  // accumulators first, then the A stages (32 columns each)

    const int my_tiles = (T.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long n_items = (long long)my_tiles * gdk;
    const unsigned long long gdk4 = (unsigned long long)gdk * 4;
    const int m_lim = p8(P.M);
    const bool nojump = (P.flags & QGTC_FLAG_NO_ZERO_JUMP) != 0;

    if (warp < kProducerWarps) {
        // ===================== adjacency producers =====================
        const int q = warp & 3, par = warp >> 2;   // TMEM lane quarter; this warp takes chunks with it % 2 == par
        const int row_in_tile = 32 * q + lane;
        uint4* my_ring = a_ring + (warp * kADepth) * 32 + lane;   // slot k at my_ring[k * 32]
        const long long my_items = (n_items - par + 1) / 2;       // items par, par+2, ...

        auto fetch = [&](long long k) {  // k-th item of this warp
            if (k < my_items) {
                const long long it = 2 * k + par;
                const int tl = (int)(it / gdk), c = (int)(it % gdk);
                const int tile = (int)blockIdx.x + tl * (int)gridDim.x;
                const int z = tile / T.mtiles, m = (tile % T.mtiles) * 128 + row_in_tile;
                const uint32_t* base = P.x1 + (unsigned long long)z * P.x1_bstride;
                const unsigned long long idx = (unsigned long long)m * gdk4 + 4ull * c;
                int nb = 0;
                if (m < m_lim && idx < P.x1_words) {
                    const unsigned long long left = P.x1_words - idx;
                    nb = left >= 4 ? 16 : (int)left * 4;
                }
                cp_async16_ca_zfill(my_ring + (k % kADepth) * 32, nb ? (const void*)(base + idx) : (const void*)base, nb);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int k = 0; k < kADepth; ++k) fetch(k);

        for (long long k = 0; k < my_items; ++k) {
            const long long it = 2 * k + par;
            const int s = (int)(it % S);
            const uint32_t use = (uint32_t)(it / S);
            cp_async_wait<kADepth - 1>();
            const uint4 v = my_ring[(k % kADepth) * 32];
            fetch(k + kADepth);  // refill the slot just read (same thread reads and writes it)
            const bool nz = nojump || __any_sync(0xffffffffu, (v.x | v.y | v.z | v.w) != 0u);
            mbar_wait(&empty[s], (use & 1) ^ 1, T.error_flag, 1);   // MMA of the previous use of this stage is done
            tc_fence_after();
            uint32_t r[32];
            if (nz) {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                    for (int cidx = 0; cidx < 8; ++cidx) r[ks * 8 + cidx] = (w[ks] >> cidx) & 0x01010101u;
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            tmem_st32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(dcols + s * 32), r);
            tmem_wait_st();
            tc_fence_before();
            if (lane == 0) flags[s * 8 + warp] = nz ? 1 : 0;
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
        }
        cp_async_wait<0>();
    } else if (warp == kLoaderWarp) {
        // ===================== feature-image loader =====================
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)NP * 128u;
            for (long long it = 0; it < n_items; ++it) {
                const int s = (int)(it % S);
                const uint32_t use = (uint32_t)(it / S);
                const int tl = (int)(it / gdk), c = (int)(it % gdk);
                const int tile = (int)blockIdx.x + tl * (int)gridDim.x;
                const int z = tile / T.mtiles;
                mbar_wait(&empty[s], (use & 1) ^ 1, T.error_flag, 2);
                mbar_expect_tx(&b_full[s], bytes);
                bulk_g2s(b_stages + (size_t)s * L.b_stage_bytes, T.bimg + (unsigned long long)z * T.bimg_gstride + (unsigned long long)c * bytes,
                         bytes, &b_full[s]);
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_i8(NP);
            const uint32_t lbo = (uint32_t)(NP / 8) * 128u, sbo = 128u;
            long long it = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int buf = tl % T.dbufs;
                const uint32_t duse = (uint32_t)(tl / T.dbufs);
                mbar_wait(&d_empty[buf], (duse & 1) ^ 1, T.error_flag, 3);   // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NP);
                uint32_t issued = 0;
                for (int c = 0; c < gdk; ++c, ++it) {
                    const int s = (int)(it % S);
                    const uint32_t use = (uint32_t)(it / S);
                    mbar_wait(&a_full[s], use & 1, T.error_flag, 4);
                    mbar_wait(&b_full[s], use & 1, T.error_flag, 5);
                    tc_fence_after();
                    const int par = (int)(it & 1);
                    const uint32_t nz = *reinterpret_cast<volatile uint32_t*>(smem_raw + L.flags + s * 8 + par * 4);
                    if (nz) {
                        const uint32_t a_tmem = tmem_base + (uint32_t)(dcols + s * 32);
                        const uint32_t b_addr = smem_u32(b_stages + (size_t)s * L.b_stage_bytes);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            mma_i8_ts(d_tmem, a_tmem + 8 * ks, make_smem_desc(b_addr + (uint32_t)ks * 2u * lbo, lbo, sbo), idesc, issued | (uint32_t)(ks > 0));
                        }
                        issued = 1;
                    }
                    tc_commit(&empty[s]);   // arrives when the MMAs above have finished reading A (TMEM) and B (smem)
                }
                d_valid[buf] = issued;
                mbar_arrive(&d_full[buf]);  // publishes d_valid
                tc_commit(&d_full[buf]);    // arrives when the accumulator is complete
            }
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
        const int q = warp & 3;
        const int M = P.M, N = P.N, ob = P.out_bit;
        const unsigned keep = (1u << ob) - 1u;
        const int wpr = s128(N) * 4;
        const unsigned long long opt_offset = (unsigned long long)p8(M) * wpr;
        for (int tl = 0; tl < my_tiles; ++tl) {
            const int buf = tl % T.dbufs;
            const uint32_t duse = (uint32_t)(tl / T.dbufs);
            const int tile = (int)blockIdx.x + tl * (int)gridDim.x;
            const int z = tile / T.mtiles, m = (tile % T.mtiles) * 128 + 32 * q + lane;
            mbar_wait(&d_full[buf], duse & 1, T.error_flag, 6);
            tc_fence_after();
            const bool valid = d_valid[buf] != 0;
            const uint32_t d_tmem = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * NP);
            if (EPI == kEpiFloat) {
                float* out = reinterpret_cast<float*>(P.out) + (unsigned long long)z * P.out_bstride;
                for (int cb = 0; cb < NP / 16; ++cb) {
                    int v[16];
                    tmem_ld16(d_tmem + 16 * cb, v);
                    if (m < M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = 16 * cb + j;
                            if (n < N) out[(size_t)m * N + n] = valid ? (float)v[j] : 0.0f;
                        }
                    }
                }
            } else {
                uint32_t* out = reinterpret_cast<uint32_t*>(P.out) + (unsigned long long)z * P.out_bstride;
                uint32_t pw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const bool row_ok = m < p8(M);
                for (int cb = 0; cb < NP / 16; ++cb) {
                    int v[16];
                    tmem_ld16(d_tmem + 16 * cb, v);
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        uint32_t four = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int n = 16 * cb + 4 * g4 + e;
                            uint32_t x = (valid && m < M && n < N) ? ((uint32_t)requant(v[4 * g4 + e], ob) & keep) : 0u;
                            four |= x << (8 * e);
                        }
                        const int nib_pos = 28 - 4 * (((16 * cb + 4 * g4) & 31) >> 2);
#pragma unroll
                        for (int p = 0; p < 8; ++p)
                            if (p < ob) pw[p] |= nibble_of_plane(four, p) << nib_pos;
                    }
                    if ((cb & 1) || cb == NP / 16 - 1) {   // a 32-column word is complete (or the tile ends on a half word)
                        const int w = cb >> 1;
                        if (row_ok) {
#pragma unroll
                            for (int p = 0; p < 8; ++p)
                                if (p < ob) out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w] = pw[p];
                        }
#pragma unroll
                        for (int p = 0; p < 8; ++p) pw[p] = 0;
                    }
                }
                if (row_ok) {   // zero pad words of the row (output buffers are not pre-zeroed)
                    for (int w = (NP + 31) / 32; w < wpr; ++w)
                        for (int p = 0; p < ob; ++p) out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w] = 0u;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

unsigned int* g_error_flag = nullptr;  // one device word per process (first device used)

}  // namespace

bool tc05_supports(const GemmArgs& a, int epi) {
    if (epi != kEpiBitsRow && epi != kEpiFloat) return false;
    if (a.bit1 != 1) return false;                 // the producers expand a single adjacency plane
    if (p8(a.N) > 256) return false;               // one MMA N block
    // worth it only when there is enough work to amortise the feature-image prepass
    const double ops = 2.0 * a.M * (double)a.K * a.N * a.count;
    return ops >= 2.0e9;
}

int launch_gemm_tc05(const GemmArgs& a, int epi, cudaStream_t st) {
    if (epi != kEpiBitsRow && epi != kEpiFloat) return QGTC_ERR_UNSUPPORTED;
    if (a.bit1 != 1 || p8(a.N) > 256) return QGTC_ERR_UNSUPPORTED;
    Tc05Args T;
    T.g = a;
    T.NP = ((p8(a.N) + 15) / 16) * 16;
    T.dbufs = (2 * T.NP + 4 * 32 <= kTmemCols) ? 2 : 1;
    int stages = (kTmemCols - T.dbufs * T.NP) / 32;
    if (stages > kMaxStages) stages = kMaxStages;
    while (stages > 2 && smem_layout(T.NP, stages).total > 200 * 1024) --stages;
    stages &= ~1;   // producers alternate by chunk parity: keep the ring even
    if (stages < 2) return QGTC_ERR_UNSUPPORTED;
    T.stages = stages;
    T.mtiles = (p8(a.M) + 127) / 128;
    T.total_tiles = T.mtiles * a.count;
    T.epi = epi;
    T.bimg_gstride = (unsigned long long)a.gdk * T.NP * 128;

    cudaError_t e;
    if (!g_error_flag) {
        e = cudaMalloc(&g_error_flag, sizeof(unsigned int));
        if (e != cudaSuccess) return record_cuda_error(e, "cudaMalloc(tc05 error flag)");
        cudaMemset(g_error_flag, 0, sizeof(unsigned int));
    }
    T.error_flag = g_error_flag;
    uint8_t* img = nullptr;
    const size_t img_bytes = (size_t)T.bimg_gstride * a.count;
    e = cudaMallocAsync(reinterpret_cast<void**>(&img), img_bytes, st);
    if (e != cudaSuccess) return record_cuda_error(e, "cudaMallocAsync(feature image)");
    T.bimg = img;

    const long long prep_items = (long long)a.gdk * T.NP * 4 * a.count;
    long long pgrid = (prep_items + 255) / 256;
    if (pgrid > kNumSMs * 8) pgrid = kNumSMs * 8;
    prep_b_kernel<<<(int)pgrid, 256, 0, st>>>(a, img, T.bimg_gstride, T.NP);
    note_launch();
    int rc = check_launch("prep_b_kernel");

    if (rc == QGTC_OK) {
        const int smem = smem_layout(T.NP, T.stages).total + 128;
        const int grid = T.total_tiles < kNumSMs ? T.total_tiles : kNumSMs;
        auto launch = [&](auto kfn) -> int {
            cudaError_t e2 = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e2 != cudaSuccess) return record_cuda_error(e2, "cudaFuncSetAttribute(gemm_tc05)");
            kfn<<<grid, kThreads, smem, st>>>(T);
            note_launch();
            return check_launch("gemm_tc05_kernel");
        };
        rc = (epi == kEpiFloat) ? launch(gemm_tc05_kernel<kEpiFloat>) : launch(gemm_tc05_kernel<kEpiBitsRow>);
    }
    e = cudaFreeAsync(img, st);
    if (rc == QGTC_OK && e != cudaSuccess) rc = record_cuda_error(e, "cudaFreeAsync(feature image)");
    return rc;
}

}  // namespace qgtc
