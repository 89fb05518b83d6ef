// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers and the fused requantize + plane-pack epilogue of one
// accumulator tile, shared by the bit-GEMM engines that run on the 5th-generation tensor cores
// (gemm_tc05.cu: uniform batches, TMA-fed; gemm_grp.cu: grouped variable-shape problems).  sm_100a only.
#pragma once

#include <cuda.h>

#include "gemm_common.cuh"

namespace qgtc {
namespace tc05 {
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
// try_wait with a suspend-time hint: the thread may stay suspended (no issue slots used) until the phase
// completes or the hint expires
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU.  Two forms, chosen by `hinted`:
//  - hinted: try_wait with a long suspend-time hint and no sleep;
//  - else `kBackoffNs` > 0 puts the warp to sleep between polls: roles with long expected waits (epilogue,
//    loaders) must not steal issue slots from the expanders -- the scheduler favours high warp ids, and a
//    failed plain try_wait comes back after only ~50-100 cycles.
template <int kBackoffNs = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* error_flag, int where, bool hinted = false) {
    uint32_t spins = 0;
    if (hinted) {   // a failed hinted try_wait sleeps on the barrier (NANOSLEEP.SYNCS), so the bound is generous in time
        while (!mbar_try_wait_hint(bar, parity, 1000000u)) {
            if (++spins > (1u << 16)) {   // ~65 s of 1 ms suspensions (a sanitizer run is orders of magnitude slower than normal): only a protocol deadlock gets here
                if (error_flag) { atomicExch_system(error_flag, 0x80000000u | (unsigned)where); __threadfence_system(); }
                __trap();
            }
        }
        return;
    }
    while (!mbar_try_wait(bar, parity)) {
        if (kBackoffNs > 0) __nanosleep(kBackoffNs);
        if (++spins > (1u << 24)) {
            if (error_flag) { atomicExch_system(error_flag, 0x80000000u | (unsigned)where); __threadfence_system(); }
            __trap();
        }
    }
}
// The same operations on a shared-window address formed once by the caller (role loops: no generic-to-shared conversion per use)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint_a(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
template <int kBackoffNs = 0>
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity, unsigned int* error_flag, int where, bool hinted = false) {
    uint32_t spins = 0;
    if (hinted) {
        while (!mbar_try_wait_hint_a(bar, parity, 1000000u)) {
            if (++spins > (1u << 16)) {
                if (error_flag) { atomicExch_system(error_flag, 0x80000000u | (unsigned)where); __threadfence_system(); }
                __trap();
            }
        }
        return;
    }
    while (!mbar_try_wait_a(bar, parity)) {
        if (kBackoffNs > 0) __nanosleep(kBackoffNs);
        if (++spins > (1u << 24)) {
            if (error_flag) { atomicExch_system(error_flag, 0x80000000u | (unsigned)where); __threadfence_system(); }
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
// D[tmem] (+)= A[tmem] * B[smem desc], kind::mxf4 (E2M1 x E2M1 -> f32, K = 64) with UE8M0 block scales read from TMEM
__device__ __forceinline__ void mma_f4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%5], p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sfa), "r"(sfb), "r"(accumulate)
        : "memory");
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
// one lane of a converged warp (cute::elect_one_sync)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
// ---- 2-CTA ("pair") forms: one tcgen05.mma.cta_group::2 covers 256 rows (128 per CTA) with each CTA holding half
// of the B columns; issued by the leader CTA (cluster rank 0) only.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
// commit that arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mma_i8_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in the leader CTA's shared memory
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 3-D tiled TMA load (cute::SM90_TMA_LOAD_3D): box of the tensor map at element coordinates (c0, c1, c2)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st16r(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
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
__host__ __device__ constexpr uint32_t make_idesc_i8(int n, int m = 128) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | (((uint32_t)m >> 4) << 24);
}

// Block-scaled instruction descriptor (cute::UMMA::InstrDescriptorBlockScaled), kind::mxf4: A = B = E2M1 (format 1), both
// K-major, UE8M0 scales, M = 128, K = 64.
__host__ __device__ constexpr uint32_t make_idesc_f4(int n) {
    return (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
}


// transpose8x8 (8 x 8 bit-matrix transpose) lives in common.cuh: the pack kernels use it too
// Requantize of QGTC_layer_hidden (reference kernel.h:31-37) followed by the "keep bits 0..ob-1" of the packing.
// NN: the accumulator is known to be non-negative (scaled expansion modes: no int32 wrap possible).
template <bool NN>
__device__ __forceinline__ uint32_t requant_keep(int acc, int mx, unsigned keep) {
    const int rq = NN ? (acc > mx ? mx - 1 : acc) : (acc > mx ? mx - 1 : (acc < 0 ? 1 : acc));
    return (uint32_t)rq & keep;
}

// Requantized, masked-to-ob-bits bytes of four accumulators, first one in the low byte.
// FAST (accumulators = 2^dshift x sum with dshift <= 3, non-negative, ob <= 4): clamp to 2^dshift (2^ob + 1) so that every
// value fits a byte, gather the four low bytes with three PRMTs and finish byte-parallel: >> dshift; the byte 2^ob + 1 ("above the range")
// becomes 2^ob - 1, the byte 2^ob becomes 0 (reference kernel.h:31-37,350), everything else passes.
// F4RAW (with FAST): the inputs are the bit patterns of f32 sums (exact non-negative integers): clamp as floats and add the magic
// number 2^23 -- the low byte of the result is the clamped value and bytes 1-2 are zero, which is all the byte gather reads
template <bool NN, bool FAST, bool F4RAW = false>
__device__ __forceinline__ uint32_t requant4(int a0, int a1, int a2, int a3, int dshift, int ob, int mx, unsigned keep) {
    if (FAST && F4RAW) {
        const float lim = (float)(mx + 1);
        const uint32_t t0 = __float_as_uint(fminf(__int_as_float(a0), lim) + 8388608.0f), t1 = __float_as_uint(fminf(__int_as_float(a1), lim) + 8388608.0f);
        const uint32_t t2 = __float_as_uint(fminf(__int_as_float(a2), lim) + 8388608.0f), t3 = __float_as_uint(fminf(__int_as_float(a3), lim) + 8388608.0f);
        const uint32_t lo = __byte_perm(t0, t1, 0x1140u), hi = __byte_perm(t2, t3, 0x1140u);
        const uint32_t four = __byte_perm(lo, hi, 0x5410u);
        const uint32_t over = (four >> ob) & four & 0x01010101u;
        return (four & (keep * 0x01010101u)) | (over * keep);
    }
    if (FAST) {
        const int lim = (mx + 1) << dshift;
        const uint32_t t0 = (uint32_t)min(a0, lim), t1 = (uint32_t)min(a1, lim), t2 = (uint32_t)min(a2, lim), t3 = (uint32_t)min(a3, lim);
        const uint32_t lo = __byte_perm(t0, t1, 0x1140u), hi = __byte_perm(t2, t3, 0x1140u);   // bytes 1..3 of every t are zero
        uint32_t four = (__byte_perm(lo, hi, 0x5410u) >> dshift) & 0x1F1F1F1Fu;
        const uint32_t over = (four >> ob) & four & 0x01010101u;                                 // bytes equal to 2^ob + 1
        return (four & (keep * 0x01010101u)) | (over * keep);
    }
    return requant_keep<NN>(a0 >> dshift, mx, keep) | (requant_keep<NN>(a1 >> dshift, mx, keep) << 8) |
           (requant_keep<NN>(a2 >> dshift, mx, keep) << 16) | (requant_keep<NN>(a3 >> dshift, mx, keep) << 24);
}

// Requantize + plane-pack epilogue of one accumulator tile for one warp (bitMM2Bit): the warp owns TMEM lane
// quarter q (one row per lane) and every other 32-column word.  OB > 0 fixes the number of output planes at
// compile time; OB == 0 takes it from `ob_rt`.  Four or more planes are packed through 8 x 8 bit transposes
// (22 operations per eight values for all planes) instead of one multiply-gather per plane and four values.
// F4: the accumulators are f32 sums (exact non-negative integers < 2^23): the magic-number add turns them into ints
template <int OB, bool NN, bool FAST, bool F4, int STEP>
__device__ __forceinline__ void epi_bits_tile(uint32_t d_tmem, int NP, int half, int N, int ob_rt, int dshift, uint32_t rmask, bool row_ok,
                                              uint32_t* out_row, unsigned long long opt_offset) {
    constexpr int kPlanes = OB > 0 ? OB : 8;
    const int ob = OB > 0 ? OB : ob_rt;
    const unsigned keep = (1u << ob) - 1u;
    const int mx = 1 << ob;
    for (int w = half; w * 32 < NP; w += STEP) {
        const bool two = w * 32 + 32 <= NP;          // the tile may end on a 16-column half word
        if (OB >= 4 && two && w * 32 + 32 <= N) {
            // ---- whole 32-column word: transpose path
            int v0[16], v1[16];
            tmem_ld16(d_tmem + 32 * w, v0);
            tmem_ld16(d_tmem + 32 * w + 16, v1);
            if (F4 && !FAST) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    v0[e] = __float_as_int(__int_as_float(v0[e]) + 8388608.0f) - 0x4B000000;
                    v1[e] = __float_as_int(__int_as_float(v1[e]) + 8388608.0f) - 0x4B000000;
                }
            }
            uint32_t xs[4], ys[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int* vv = g < 2 ? &v0[8 * g] : &v1[8 * (g - 2)];
                // the transpose wants the first column in the top byte: reverse the four bytes
                uint32_t x = __byte_perm(requant4<NN, FAST, F4>(vv[0], vv[1], vv[2], vv[3], dshift, ob, mx, keep), 0u, 0x0123u);
                uint32_t y = __byte_perm(requant4<NN, FAST, F4>(vv[4], vv[5], vv[6], vv[7], dshift, ob, mx, keep), 0u, 0x0123u);
                transpose8x8(x, y);
                xs[g] = x;
                ys[g] = y;
            }
            if (row_ok) {
#pragma unroll
                for (int p = 0; p < kPlanes; ++p) {
                    const uint32_t sel = (uint32_t)((p & 3) | ((4 + (p & 3)) << 4));
                    const uint32_t lo = p >= 4 ? __byte_perm(xs[3], xs[2], sel) : __byte_perm(ys[3], ys[2], sel);
                    const uint32_t hi = p >= 4 ? __byte_perm(xs[1], xs[0], sel) : __byte_perm(ys[1], ys[0], sel);
                    out_row[(unsigned long long)p * opt_offset + w] = __byte_perm(lo, hi, 0x5410u) & rmask;
                }
            }
            continue;
        }
        // ---- general path (few planes, or a word cut by N or by the tile): multiply-gather per plane
        uint32_t pw[kPlanes];
#pragma unroll
        for (int p = 0; p < kPlanes; ++p) pw[p] = 0u;
        for (int hcb = 0; hcb < (two ? 2 : 1); ++hcb) {
            const int cb = 2 * w + hcb;
            int v[16];
            tmem_ld16(d_tmem + 16 * cb, v);
            if (F4 && !FAST) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __float_as_int(__int_as_float(v[e]) + 8388608.0f) - 0x4B000000;
            }
            const bool whole = 16 * cb + 16 <= N;   // else: one byte mask per four values
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
                uint32_t four = requant4<NN, FAST, F4>(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3], dshift, ob, mx, keep);
                uint32_t mask = rmask;
                if (!whole) {
                    const int nv = N - (16 * cb + 4 * g4);   // valid columns of this group of four
                    mask = nv >= 4 ? rmask : (nv <= 0 ? 0u : (rmask & ((1u << (8 * nv)) - 1u)));
                }
                four &= mask;
                const int nib_pos = 28 - 4 * (4 * hcb + g4);
#pragma unroll
                for (int p = 0; p < kPlanes; ++p)
                    if (p < ob) pw[p] |= nibble_of_plane(four, p) << nib_pos;
            }
        }
        if (row_ok) {
#pragma unroll
            for (int p = 0; p < kPlanes; ++p)
                if (p < ob) out_row[(unsigned long long)p * opt_offset + w] = pw[p];
        }
    }
}

}  // namespace tc05
}  // namespace qgtc
