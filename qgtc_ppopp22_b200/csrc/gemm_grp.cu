// Bit-GEMM engine 3: GROUPED tcgen05 engine -- many independent problems of different M and K in ONE launch
// ("fused per-batch forward / grouped launch"), sm_100a.
//
// One launch executes one op of the Cluster-GCN / batched-GIN forward (reference main_qgtc.py:128-154) for ALL batches of
// a rank: problem z has its own M, K, operand and output pointers; N, the bit widths and the epilogue are common.  It is
// also the engine for everything the TMA-fed batch engine (gemm_tc05.cu) refuses: multi-plane A operands (bit1 > 1: the
// X.W and o.W products, kernel.h:245-391 handles any act_bit), operands that end before the index formulas do
// (zero-extension), the column-major epilogue (kernel.h:651-810, intended semantics) and single small problems (one launch,
// no feature-image scratch, no allocation).
//
// Data flow of one CTA (persistent; work unit = one 128-row tile of one problem, units are taken round-robin):
//
//   A operand: packed planes R(bit1, M, K) in global memory
//        --cp.async 16 B per (plane, chunk) per lane, zero-fill beyond the tensor--> lane-private staging ring in smem
//        --LDS.128--> registers: byte k of row m = sum_p 2^p bit_p  (K permuted inside each 32-group, as in gemm_tc05.cu)
//        --tcgen05.st.32x32b.x32--> TMEM stage (the MMA's A operand; row = TMEM lane)
//   B operand: packed planes C(bit2, K, N) in global memory (raw words, whatever produced them)
//        --cp.async--> staging ring --converter warps: re-assemble u8 values in the MMA's K order--> smem stage in the
//        canonical K-major core-matrix layout (no prep kernel, no feature image in HBM) --fence.proxy.async--> MMA
//   tcgen05.mma.cta_group::1.kind::i8  D[tmem] (+)= A[tmem] x B[smem desc]; the first MMA of a unit overwrites D
//   tcgen05.ld -> requantize + pack planes (row-major words / column-major words) or float -> global stores
//
// Warp roles: 0-7 A expanders (lane quarter x 2 groups; a group takes whole pipeline items round-robin, each warp loads
// the rows it expands itself: no loader warp, no cross-thread hand-over of the packed adjacency), 8-11 epilogue,
// 12-15 B converters, 16 MMA issuer (also owns the TMEM allocation).
// Zero-tile jumping as in gemm_tc05.cu: on the fly (an all-zero 128 x 128-bit chunk skips its 4 MMAs) and, with a per-problem
// tile bitmap (qgtc_tile_bitmap), all-zero pipeline stages are skipped by every role (stage 0 of a unit is always visited:
// it initialises the accumulator).
#include <cstdlib>
#include <cstring>

#include "tc05_common.cuh"

namespace qgtc {

using namespace tc05;

unsigned int* tc05_error_flag_device();   // gemm_tc05.cu: the watchdog word (mapped host memory)

namespace {

constexpr int kExpWarps = 8;
constexpr int kEpiWarp0 = kExpWarps;          // 4 epilogue warps (warp % 4 == TMEM lane quarter)
constexpr int kConvWarp0 = kEpiWarp0 + 4;
constexpr int kConvWarps = 4;
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kIssuerWarp = kConvWarp0 + kConvWarps;
constexpr int kIssuer2Warp = kIssuerWarp + 1;      // fused update GEMM (second tcgen05.mma chain); idle otherwise
constexpr int kThreads = (kIssuer2Warp + 1) * 32;  // 576
constexpr int kARing = 3;                     // staged pipeline items per expander warp (cp.async groups in flight)
constexpr int kMaxBRing = 4;                  // ... per converter thread: 2..4 by shared-memory budget (GrpArgs::b_ring)
constexpr int kMaxTasks = 2;                  // (chunk, column) conversion tasks per converter thread and item: cps * NP <= 256
constexpr int kMaxStagesG = 4;
constexpr int kMaxCpsG = 4;
constexpr int kGrpMaxProblems = 224;          // kernel-parameter table (80 B each): one launch takes up to this many problems
constexpr int kUnitTable = 4096;              // launches of up to this many units carry a unit -> problem table (no search on the device)
// The tables travel BY VALUE as kernel parameters, and a launch costs ~45 ns per KB of parameters (0.6 us with 64 B, 1.8 us with
// 28 KB: profiles/r2_launch_floor_call64.txt), so the width instances exist for three table sizes and the launcher picks the
// smallest that holds the launch: 2.5 KB (a rank's share of an epoch at 4-8 GPUs), 8 KB (75 batches of <= 20 tiles), 22 KB.
constexpr int kTierSProblems = 24, kTierSUnits = 512;
constexpr int kTierMProblems = 80, kTierMUnits = 1536;
constexpr int kMaxUnitsPerCta = 384;          // per-CTA unit records in shared memory (64 B each); larger launches are split by the host
constexpr int kTraceRoles = 8, kTraceEvents = 256;   // QGTC_GRP_TRACE=1: [role][event][tag, clock64] of CTA 0

struct GrpProblem {
    const uint32_t* x1;
    const uint32_t* x2;
    void* out;
    const uint32_t* bitmap;                   // optional tile bitmap of this problem's 1-bit A operand, or null
    unsigned long long x1_words, x2_words;    // words really present (zero-extension beyond)
    int M, K;
    int unit0;                                // index of this problem's first work unit
    int mt;                                   // (per-CTA unit records only) the unit's 128-row tile
    void* out2;                               // fused update: the second op's output of this problem
    unsigned long long pad_;
};
static_assert(sizeof(GrpProblem) == 80, "problem descriptors are 80 bytes");
constexpr int kDescWords = (int)(sizeof(GrpProblem) / 4);

struct GrpHead {
    int count, units;
    int N, NP, bit1, bit2, out_bit, epi, pad128;
    int cps, stages, dbufs, groups;           // chunks per item, A/B ring depth, accumulator buffers, expander groups in use
    int a_slot_bytes, b_slot_bytes, b_stage_bytes, b_ring, max_nst, mode2, any_bitmap;
    int off_bstage, off_bring, off_bars, off_flags, off_misc, off_units;   // shared-memory offsets from the 1024-byte aligned base
    unsigned int cps_inv, np_inv;             // ceil(2^32 / cps), ceil(2^32 / NP): divisions by multiply-high (cps_inv 0: cps == 1)
    int units_cap;                            // ceil(units / grid): the per-CTA work lists are sized for this many units
    int conv_alt;                             // <= 64 conversion tasks per item: the two converter half-groups take alternate items (else 0)
    int may_reuse;                            // some item may find its B stage already converted (shared x2): run the re-use scan
    unsigned flags;
    unsigned int* error_flag;
    unsigned long long* trace;                // clock64 trace buffer (debug instance only), else null
    // fused update GEMM (0 = off): every unit's requantized tile (A2, K2 = N <= 64) times one shared weight matrix W
    int fused, N2, NP2, bit_w, out_bit2, epi2, pad128_2, write_out1;
    int a2_col, d2_col, off_b2;               // TMEM columns of the A2 operand / the second accumulator, smem offset of W's image
    const uint32_t* w2;
    unsigned long long w2_words;
};
template <int NPROB, int NUNIT>
struct GrpArgsT : GrpHead {
    static constexpr int kUnits = NUNIT;
    GrpProblem p[NPROB];
    unsigned char unit_z[NUNIT];              // problem of unit u (valid when units <= NUNIT)
};
using GrpArgs = GrpArgsT<kGrpMaxProblems, kUnitTable>;
using GrpArgsM = GrpArgsT<kTierMProblems, kTierMUnits>;
using GrpArgsS = GrpArgsT<kTierSProblems, kTierSUnits>;
static_assert(kGrpMaxProblems <= 256, "unit_z holds problem indices in a byte");

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
// whole piece known to exist: no source-size operand, shared-window destination address formed by the caller
__device__ __forceinline__ void cp_async16_a(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// bytes of the 16-byte piece starting at word idx that really exist in the operand
__device__ __forceinline__ int piece_bytes(unsigned long long idx, unsigned long long words) {
    if (idx >= words) return 0;
    const unsigned long long left = words - idx;
    return left >= 4 ? 16 : (int)left * 4;
}

// Per-CTA work lists, built in shared memory by the kernel prologue (all threads, before the dependency wait) so that the
// role loops -- single warps running latency-bound instruction streams -- do no searching, dividing or bitmap walking:
//   UnitRec  one per work unit of this CTA (units u = blockIdx.x, blockIdx.x + gridDim.x, ...): operand pointers, sizes, row tile
//   ItemRec  one per pipeline item (a stage of cps K chunks of a unit), in processing order: all stages of a unit without a
//            tile bitmap, else stage 0 (it initialises the accumulator) plus every stage with a non-zero chunk bit
struct UnitRec {
    const uint32_t* x1;
    const uint32_t* x2;
    void* out;
    const uint32_t* bm;                       // bitmap words of this tile, or null
    unsigned long long x1_words, x2_words;
    int M, gdk, mt, item0;                    // rows, K chunks, 128-row tile, first item of the unit
    void* out2;                               // fused update: second output
    unsigned long long pad_;
};
static_assert(sizeof(UnitRec) == sizeof(GrpProblem), "a unit record is built in place of the copied descriptor");
struct ItemRec {
    int c0;                                   // first K chunk
    unsigned short unit;                      // index into the unit records
    unsigned char nch, flags;                 // chunks (1..cps); bit 0: first item of its unit, bit 1: last
};
static_assert(sizeof(ItemRec) == 8, "item records are 8 bytes");

// stages of a tile with a bitmap: visit(stage) for stage 0 and every stage holding a non-zero chunk; returns their number
template <typename F>
__device__ __forceinline__ int walk_bitmap_stages(const uint32_t* bm, int gdk, int cps, unsigned cps_inv, F visit) {
    int n = 1, covered_end = cps;
    visit(0);
    const int bmw = (gdk + 31) >> 5;
    for (int w = 0; w < bmw; ++w) {
        const int left = gdk - 32 * w;
        uint32_t m = __ldg(bm + w) & (left >= 32 ? 0xffffffffu : ((1u << left) - 1u));
        while (m) {
            const int k = covered_end - 32 * w;          // chunks of this word already inside an emitted stage
            if (k > 0) m &= k >= 32 ? 0u : ~((1u << k) - 1u);
            if (!m) break;
            const int c = 32 * w + __ffs(m) - 1;
            const int st = cps_inv ? (int)__umulhi((unsigned)c, cps_inv) : c;
            covered_end = (st + 1) * cps;
            visit(st);
            ++n;
        }
    }
    return n;
}

#define QGTC_GTRACE(role, tag)                                                                              \
    do {                                                                                                    \
        if (TRACE && T.trace && blockIdx.x == 0 && tn < kTraceEvents) {                                     \
            T.trace[((role) * kTraceEvents + tn) * 2] = (unsigned long long)(tag);                          \
            T.trace[((role) * kTraceEvents + tn) * 2 + 1] = (unsigned long long)clock64();                  \
            ++tn;                                                                                           \
        }                                                                                                   \
    } while (0)

// Re-assemble NB bit planes of one 128-bit K chunk into 32 registers of four u8 values: byte i of r[8 s + c] = value of K
// element 32 s + c + 8 i = sum over planes of bit (c + 8 i) of the plane's word s, moved to bit p (the K permutation of
// gemm_tc05.cu: a register is a mask of the packed word).  Plane p's 16 bytes are at src + p * stride.  With the plane
// index known at compile time every (plane, register) step is one shift and one AND-OR.  Returns the OR of all words.
// SCALED (the B operand in expansion mode 2, see the expanders): element of column c additionally carries 2^(3 - c % 4).
template <int NB, bool SCALED>
__device__ __forceinline__ uint32_t assemble_planes(const unsigned char* src, int stride, uint32_t (&r)[32]) {
    uint32_t any = 0u;
#pragma unroll
    for (int p = 0; p < NB; ++p) {
        const uint4 pv = *reinterpret_cast<const uint4*>(src + p * stride);
        const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
        any |= pv.x | pv.y | pv.z | pv.w;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int to = p + (SCALED ? 3 - (c & 3) : 0);     // compile-time after unrolling; <= 7 (SCALED needs NB <= 5)
                const uint32_t sh = c >= to ? (w[ks] >> (c - to)) : (w[ks] << (to - c));
                const uint32_t v = sh & (0x01010101u << to);
                r[ks * 8 + c] = p == 0 ? v : (r[ks * 8 + c] | v);
            }
        }
    }
    return any;
}
// any plane count (1..8): the common widths unrolled, the rest through a runtime plane loop
template <bool SCALED>
__device__ __forceinline__ uint32_t assemble_planes_rt(int nb, const unsigned char* src, int stride, uint32_t (&r)[32]) {
    switch (nb) {
        case 1: return assemble_planes<1, SCALED>(src, stride, r);
        case 2: return assemble_planes<2, SCALED>(src, stride, r);
        case 4: return assemble_planes<4, SCALED>(src, stride, r);
        case 8: if (!SCALED) return assemble_planes<8, false>(src, stride, r); break;
        default: break;
    }
    uint32_t any = 0u;
#pragma unroll
    for (int e = 0; e < 32; ++e) r[e] = 0u;
#pragma unroll 1
    for (int p = 0; p < nb; ++p) {
        const uint4 pv = *reinterpret_cast<const uint4*>(src + p * stride);
        const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
        any |= pv.x | pv.y | pv.z | pv.w;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int c = 0; c < 8; ++c) r[ks * 8 + c] |= ((w[ks] >> c) & 0x01010101u) << (p + (SCALED ? 3 - (c & 3) : 0));
        }
    }
    return any;
}

// BW > 0: the plane count is the kernel's compile-time bit width (one unrolled variant in the code instead of a switch over five:
// the role loops of this kernel run out of the instruction cache -- collapsing the switches cut an epoch by 5-10 %)
template <int BW, bool SCALED>
__device__ __forceinline__ uint32_t assemble_planes_w(int nb, const unsigned char* src, int stride, uint32_t (&r)[32]) {
    if (BW > 0 && !(SCALED && BW > 5)) return assemble_planes<(BW > 0 ? BW : 1), SCALED>(src, stride, r);
    return assemble_planes_rt<SCALED>(nb, src, stride, r);
}

// Half of a chunk (words 2 h and 2 h + 1 of every plane -> 16 registers = k groups 4 h .. 4 h + 3): when an item has at most 64
// (chunk, column) tasks, two converter threads share each task so that all four converter warps work on every item.
template <int NB, bool SCALED>
__device__ __forceinline__ void assemble_half(const unsigned char* src, int stride, int h, uint32_t (&r)[16]) {
#pragma unroll
    for (int p = 0; p < NB; ++p) {
        const uint4 pv = *reinterpret_cast<const uint4*>(src + p * stride);
        const uint32_t w[2] = {h ? pv.z : pv.x, h ? pv.w : pv.y};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int to = p + (SCALED ? 3 - (c & 3) : 0);
                const uint32_t sh = c >= to ? (w[ks] >> (c - to)) : (w[ks] << (to - c));
                const uint32_t v = sh & (0x01010101u << to);
                r[ks * 8 + c] = p == 0 ? v : (r[ks * 8 + c] | v);
            }
        }
    }
}
template <bool SCALED>
__device__ __forceinline__ void assemble_half_rt(int nb, const unsigned char* src, int stride, int h, uint32_t (&r)[16]) {
    switch (nb) {
        case 1: assemble_half<1, SCALED>(src, stride, h, r); return;
        case 2: assemble_half<2, SCALED>(src, stride, h, r); return;
        case 4: assemble_half<4, SCALED>(src, stride, h, r); return;
        case 8: if (!SCALED) { assemble_half<8, false>(src, stride, h, r); return; } break;
        default: break;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) r[e] = 0u;
#pragma unroll 1
    for (int p = 0; p < nb; ++p) {
        const uint4 pv = *reinterpret_cast<const uint4*>(src + p * stride);
        const uint32_t w[2] = {h ? pv.z : pv.x, h ? pv.w : pv.y};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
            for (int c = 0; c < 8; ++c) r[ks * 8 + c] |= ((w[ks] >> c) & 0x01010101u) << (p + (SCALED ? 3 - (c & 3) : 0));
        }
    }
}

template <int BW, bool SCALED>
__device__ __forceinline__ void assemble_half_w(int nb, const unsigned char* src, int stride, int h, uint32_t (&r)[16]) {
    if (BW > 0 && !(SCALED && BW > 5)) { assemble_half<(BW > 0 ? BW : 1), SCALED>(src, stride, h, r); return; }
    assemble_half_rt<SCALED>(nb, src, stride, h, r);
}

// fast path first: a plain try_wait (already-complete phases return at once), then the bounded suspending wait
__device__ __forceinline__ void gwait(uint64_t* bar, uint32_t parity, unsigned int* ef, int where) {
    if (!mbar_try_wait(bar, parity)) mbar_wait(bar, parity, ef, where, true);
}

// A1: single-plane A operand (the adjacency products): the lean expansion.  TRACE: debug instance with a clock64 timeline of CTA 0.
// BW: 1 / 2 / 4 / 8 = the launcher guarantees bit2 == out_bit == BW (and bit1 == BW unless A1, and bit_w == out_bit2 == BW when fused):
// the conversions and epilogues are compiled for that width only; 0 = any widths (run-time switches).
// ARGS: the parameter block, i.e. the size of the by-value problem / unit tables (GrpArgs, GrpArgsM, GrpArgsS).
template <bool A1, bool TRACE, int BW = 0, typename ARGS = GrpArgs>
__global__ void __launch_bounds__(kThreads, 1) gemm_grp_kernel(const __grid_constant__ ARGS T) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NP = T.NP, S = T.stages, CPS = T.cps;

    unsigned char* a_rings = smem;                                   // [kExpWarps][kARing][a_slot_bytes]
    unsigned char* b_stages = smem + T.off_bstage;                   // [S][cps * NP * 128]
    unsigned char* b_ring = smem + T.off_bring;                      // [T.b_ring][b_slot_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T.off_bars);
    uint64_t* full = bars;                         // [kMaxStagesG] 4 expander warps + kConvWarps converter warps -> MMA
    uint64_t* empty = bars + kMaxStagesG;          // [kMaxStagesG] MMA commit -> expanders, converters
    uint64_t* d_full = bars + 2 * kMaxStagesG;     // [2] MMA commit -> epilogue
    uint64_t* d_empty = d_full + 2;                // [2] 4 epilogue warps -> MMA
    uint64_t* a2_full = d_empty + 2;               // fused update: 4 epilogue warps wrote the A2 operand -> second issuer
    uint64_t* d2_full = a2_full + 1;               //               second MMA chain done -> epilogue
    uint64_t* b2_ready = d2_full + 1;              //               converters built W's image (once)
    volatile unsigned char* flags = smem + T.off_flags;   // [stage][chunk][lane quarter]: that 32 x 128-bit block has a one
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T.off_misc);
    int* n_items_slot = reinterpret_cast<int*>(smem + T.off_misc + 8);
    UnitRec* units = reinterpret_cast<UnitRec*>(smem + T.off_units);          // [units of this CTA]
    int tn = 0;                                    // trace cursor (per thread; lane 0 of the traced warps writes)
    (void)tn;
    const bool tlane = lane == 0;
    if (threadIdx.x == 32) QGTC_GTRACE(7, 1);

    // let the next kernel of the stream be scheduled as SMs free up (it waits for this grid in its own griddepcontrol.wait)
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStagesG; ++s) {
            mbar_init(&full[s], 4 + (T.conv_alt ? kConvWarps / 2 : kConvWarps));
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], 4);
        }
        mbar_init(a2_full, 4);
        mbar_init(d2_full, 1);
        mbar_init(b2_ready, kConvWarps);
        fence_mbar_init();
    }
    if (warp == kIssuerWarp) tmem_alloc(tmem_slot, 512);
    if (threadIdx.x == 32) QGTC_GTRACE(7, 5);

    if (T.any_bitmap && !(T.flags & QGTC_FLAG_X1_STATIC)) asm volatile("griddepcontrol.wait;\n" ::: "memory");   // the lists read the tile bitmaps
    // ---- work lists (kernel parameters, and tile bitmaps: nothing here depends on the previous kernel)
    // (no run-time divisions in the prologue or the role preambles: this code runs once, cold, on every launch's critical path)
    const int units_cap = T.units_cap;           // ceil(units / grid): the launcher sized the lists for this many units per CTA
    const int units_here = units_cap - (((int)blockIdx.x + (units_cap - 1) * (int)gridDim.x >= T.units) ? 1 : 0);
    ItemRec* items = reinterpret_cast<ItemRec*>(smem + T.off_units + units_cap * (int)sizeof(UnitRec));
    unsigned char* reuse_flag = reinterpret_cast<unsigned char*>(items + units_cap * T.max_nst);
    // One warp per unit: the problem index is a warp-uniform constant-bank read, the 64-byte descriptor one read per lane
    // from a single line (per-thread searches through the parameter table serialise in the constant cache).
    for (int k = warp; k < units_here; k += kThreads / 32) {
        const int u = (int)blockIdx.x + k * (int)gridDim.x;
        int z;
        if (T.units <= ARGS::kUnits) {
            z = T.unit_z[u];
        } else {
            int lo = 0, hi = T.count - 1;      // largest z with unit0[z] <= u
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (T.p[mid].unit0 <= u) lo = mid; else hi = mid - 1;
            }
            z = lo;
        }
        uint32_t* raw = reinterpret_cast<uint32_t*>(&units[k]);
        if (lane < kDescWords) raw[lane] = reinterpret_cast<const uint32_t*>(&T.p[z])[lane];
        __syncwarp();
        if (lane == 0) {
            const GrpProblem P = *reinterpret_cast<const GrpProblem*>(raw);
            UnitRec r;
            r.x1 = P.x1; r.x2 = P.x2; r.out = P.out; r.out2 = P.out2; r.pad_ = 0;
            r.x1_words = P.x1_words; r.x2_words = P.x2_words;
            r.M = P.M;
            r.gdk = s128(P.K);
            r.mt = u - P.unit0;
            r.bm = P.bitmap ? P.bitmap + (size_t)r.mt * ((r.gdk + 31) >> 5) : nullptr;
            r.item0 = r.bm ? walk_bitmap_stages(r.bm, r.gdk, CPS, T.cps_inv, [](int) {}) : (T.cps_inv ? (int)__umulhi((unsigned)(r.gdk + CPS - 1), T.cps_inv) : r.gdk);   // count, for now
            units[k] = r;
        }
    }
    if (threadIdx.x == 32) QGTC_GTRACE(7, 6);
    __syncthreads();
    if (threadIdx.x == 32) QGTC_GTRACE(7, 7);
    if (warp == 0) {       // item lists: an exclusive scan of the per-unit item counts, then lane k writes the items of unit base + k
        int carry = 0;
        for (int base = 0; base < units_here; base += 32) {
            const int k = base + lane;
            const int c = k < units_here ? units[k].item0 : 0;
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (k < units_here) {
                const UnitRec& r = units[k];
                const int first = carry + incl - c, gdk = r.gdk;
                int i = first;
                auto put = [&](int st) {
                    ItemRec it;
                    it.c0 = st * CPS;
                    it.unit = (unsigned short)k;
                    it.nch = (unsigned char)(gdk - it.c0 < CPS ? gdk - it.c0 : CPS);
                    it.flags = (unsigned char)(i == first ? 1 : 0);
                    items[i++] = it;
                };
                if (r.bm) walk_bitmap_stages(r.bm, gdk, CPS, T.cps_inv, put);
                else for (int st = 0; st * CPS < gdk; ++st) put(st);
                items[i - 1].flags |= 2;
                units[k].item0 = first;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            *n_items_slot = carry;
            reinterpret_cast<volatile unsigned char*>(n_items_slot)[4] = 0;     // the "never a re-use" flag byte (launches without the re-use scan)
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_items = *n_items_slot;
    if (threadIdx.x == 32) QGTC_GTRACE(7, 2);
    // Programmatic dependent launch: this grid may start while the previous kernel of the stream is still running (its CTAs
    // take over SMs as that kernel's CTAs retire).  Memory the previous kernel writes -- or still reads -- may only be touched
    // after griddepcontrol.wait (no-op for a normal launch).  Each role waits where it first needs it: the epilogue before
    // its first store, expanders / converters before their first load unless the caller declared that operand static
    // (QGTC_FLAG_X1_STATIC / QGTC_FLAG_X2_STATIC: adjacency, weights), so those loads overlap the previous kernel's tail.
    if (threadIdx.x == 32) QGTC_GTRACE(7, 3);

    const int dcols = T.dbufs * NP;                // accumulators first, then the A stages
    const int stage_cols = CPS * 32;
    const bool nojump = (T.flags & QGTC_FLAG_NO_ZERO_JUMP) != 0;
    const bool mode2 = A1 && T.mode2 != 0;         // launcher: bit2 <= 5 and 8 x the largest possible sum fits an int32
    const int dshift = mode2 ? 3 : 0;
    unsigned int* const ef = T.error_flag;

    if (warp < kExpWarps) {
        // ===================== A expanders =====================
        // warp = (TMEM lane quarter q, group g); the groups take the items round-robin (G <= S: see gemm_tc05.cu).  A warp
        // loads the rows it expands itself: cp.async into a lane-private ring, kARing - 1 of its own items ahead.
        const int q = warp & 3, g = warp >> 2;
        const int G = T.groups;
        if (g < G) {
            unsigned char* ring = a_rings + (size_t)warp * kARing * T.a_slot_bytes + lane * 16;   // piece i of this lane at + i * 512
            const int trole = warp == 0 ? 0 : (warp == 4 ? 1 : -1);
            const bool tr = TRACE && trole >= 0 && tlane;
            if (!(T.flags & QGTC_FLAG_X1_STATIC)) asm volatile("griddepcontrol.wait;\n" ::: "memory");
            if (tr) QGTC_GTRACE(trole, 10);
            int s = g;                             // stage and use count of the item being consumed (g < G <= S)
            uint32_t use = 0;
            // software pipeline: iteration k issues the loads of own item k and consumes own item k - (kARing - 1)
            int it_l = g, slot_l = 0, slot_c = 0;
            int lu_unit = -1;                      // unit whose row pointer / bounds are cached below
            unsigned long long lu_act = 0, lu_row0 = 0;
            const uint32_t* lu_rowp = nullptr;
            bool lu_row_ok = false, lu_fast = false;
#pragma unroll 1
            for (int it_c = g - (kARing - 1) * G; it_c < n_items; it_c += G, it_l += G) {
                if (it_l < n_items) {
                    const ItemRec I = items[it_l];
                    const UnitRec& U = units[I.unit];
                    const int nb1 = A1 ? 1 : T.bit1;
                    if ((int)I.unit != lu_unit) {
                        // per-unit part of the address arithmetic (64-bit products, bounds): once per unit, not per item -- this
                        // instruction stream sits between two of the warp's items, i.e. on the group's critical path
                        lu_unit = I.unit;
                        const int m = U.mt * 128 + 32 * q + lane;
                        const unsigned long long gdk4 = (unsigned long long)U.gdk * 4;
                        lu_act = (unsigned long long)p8(U.M) * gdk4;                                     // reference kernel.h:265
                        lu_row0 = (unsigned long long)m * gdk4;
                        lu_row_ok = m < p8(U.M);
                        // every piece of this row, in every plane, lies inside the tensor: plain 16-byte copies from a row pointer
                        lu_fast = lu_row_ok && lu_row0 + (unsigned long long)(nb1 - 1) * lu_act + gdk4 <= U.x1_words;
                        lu_rowp = U.x1 + lu_row0;
                    }
                    const unsigned long long act_offset = lu_act;
                    const unsigned long long idx0 = lu_row0 + 4ull * I.c0;
                    const bool row_ok = lu_row_ok;
                    unsigned char* dst = ring + slot_l * T.a_slot_bytes;
                    if (lu_fast) {
                        const uint32_t d0 = smem_u32(dst);
                        const uint32_t* src0 = lu_rowp + 4 * I.c0;
                        for (int p = 0; p < nb1; ++p, src0 += act_offset) {
#pragma unroll
                            for (int j = 0; j < kMaxCpsG; ++j)
                                if (j < I.nch) cp_async16_a(d0 + (uint32_t)((p * CPS + j) * 512), src0 + 4 * j);
                        }
                    } else if (row_ok && idx0 + (unsigned long long)(nb1 - 1) * act_offset + 4ull * I.nch <= U.x1_words) {   // all pieces inside the tensor
                        for (int p = 0; p < nb1; ++p)
                            for (int j = 0; j < I.nch; ++j)
                                cp_async16_zfill(dst + (p * CPS + j) * 512, U.x1 + idx0 + (unsigned long long)p * act_offset + 4 * j, 16);
                    } else {
                        for (int p = 0; p < nb1; ++p)
                            for (int j = 0; j < I.nch; ++j) {
                                const unsigned long long idx = idx0 + (unsigned long long)p * act_offset + 4ull * j;
                                const int nb = row_ok ? piece_bytes(idx, U.x1_words) : 0;
                                cp_async16_zfill(dst + (p * CPS + j) * 512, nb ? (const void*)(U.x1 + idx) : (const void*)U.x1, nb);
                            }
                    }
                    if (++slot_l == kARing) slot_l = 0;
                }
                cp_async_commit();
                if (tr) QGTC_GTRACE(trole, 11);
                if (it_c < 0) continue;
                cp_async_wait<kARing - 1>();       // this lane's pieces of item it_c have landed (only this lane reads them)
                if (tr) QGTC_GTRACE(trole, 12);
                const int nch = items[it_c].nch;
                const unsigned char* src = ring + slot_c * T.a_slot_bytes;
                if (++slot_c == kARing) slot_c = 0;
                const uint32_t st_s = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(dcols + s * stage_cols);
                unsigned nzmask = 0;
#pragma unroll 1
                for (int j = 0; j < nch; ++j) {
                    uint32_t r[32];
                    uint32_t any;
                    if (A1 && mode2) {
                        // expansion mode 2 (gemm_tc05.cu): bit c of every byte stays where it is inside its nibble -- one AND per
                        // register, one shift per word -- so A holds 2^(c % 4); the converters scale the B rows by 2^(3 - c % 4)
                        // and the accumulators hold 8 x the sums (the epilogue shifts them back, exactly)
                        const uint4 pv = *reinterpret_cast<const uint4*>(src + j * 512);
                        const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
                        any = pv.x | pv.y | pv.z | pv.w;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t hi = w[ks] >> 4;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                r[ks * 8 + c] = w[ks] & (0x01010101u << c);
                                r[ks * 8 + 4 + c] = hi & (0x01010101u << c);
                            }
                        }
                    } else {
                        any = A1 ? assemble_planes<1, false>(src + j * 512, 0, r) : assemble_planes_w<BW, false>(T.bit1, src + j * 512, CPS * 512, r);
                    }
                    if (nojump || __any_sync(0xffffffffu, any != 0u)) nzmask |= 1u << j;
                    if (j == 0) {   // the first chunk is expanded before the wait: only the store sits behind it
                        if (tr) QGTC_GTRACE(trole, 13);
                        gwait(&empty[s], (use & 1) ^ 1, ef, 1);     // the MMAs of this stage's previous use are done
                        tc_fence_after();
                        if (tr) QGTC_GTRACE(trole, 14);
                    }
                    tmem_st32(st_s + (uint32_t)(32 * j), r);
                }
                if (lane < nch) flags[(s * kMaxCpsG + lane) * 4 + q] = (unsigned char)((nzmask >> lane) & 1u);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                if (tr) QGTC_GTRACE(trole, 15);
                s += G;
                if (s >= S) { s -= S; ++use; }
            }
            cp_async_wait<0>();
        }
    } else if (warp >= kConvWarp0 && warp < kConvWarp0 + kConvWarps) {
        // ===================== B converters =====================
        // thread t takes the (chunk, column) tasks t and t + 128 of every item: bit2 pieces of 16 B (128 K bits of one column,
        // one plane each) in, 128 u8 values in the MMA's K order out, as 8 core-matrix rows of 16 B.
        const int t = threadIdx.x - kConvWarp0 * 32;
        const int p8n = p8(T.N);
        const int pn = T.pad128 ? p128(T.N) : p8n;                     // columns per plane of X2: kernel.h:266 / :836
        // At most 64 tasks per item (cps * NP <= 64, e.g. N = 16): one task per thread and 64 threads per item.  conv_alt (default): the two
        // half-groups (threads 0-63 / 64-127) take alternate items -- the converters' per-item chain (issue the loads, wait, convert, arrive)
        // is the engine's critical path, and this runs two of them side by side.  Otherwise two threads share a task, each converting
        // one half of the chunk (threads 0-63: words 0-1, threads 64-127: words 2-3).
        const bool few = CPS * NP <= 64;
        const bool split = few && !T.conv_alt;
        const int half = split ? (t >> 6) : 0;
        const int cg = T.conv_alt ? (t >> 6) : 0;                      // this thread's item parity; items advance by 1 + conv_alt
        int tj[kMaxTasks], tn_[kMaxTasks];                             // this thread's tasks: chunk within the item, column
#pragma unroll
        for (int i = 0; i < kMaxTasks; ++i) {
            const int task = few ? ((t & 63) + i * 64 * 1024) : (t + i * kConvThreads);   // few: one task per thread (the second is out of range)
            tj[i] = (int)__umulhi((unsigned)task, T.np_inv);
            tn_[i] = task - tj[i] * NP;
        }
        // Stage re-use (one byte per item, written and read by the converters only): an item whose B chunk is identical to those of the S
        // items before it -- the weight products, where every unit multiplies with the same W -- finds its stage already
        // holding the converted image: the converters neither load nor convert anything for it.
        const int n_scan = T.may_reuse ? n_items : 0;   // launcher: no two problems share x2 and no problem spans the grid -> nothing to find
        for (int it = t; it < n_scan; it += kConvThreads) {
            bool same = it >= S;
            const ItemRec I = items[it];
            const UnitRec& U = units[I.unit];
            for (int d = 1; same && d <= S; ++d) {
                const ItemRec J = items[it - d];
                const UnitRec& V = units[J.unit];
                same = J.c0 == I.c0 && J.nch == I.nch && V.x2 == U.x2 && V.x2_words == U.x2_words && V.gdk == U.gdk;
            }
            reuse_flag[it] = same ? 1 : 0;
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(kConvThreads) : "memory");     // the four converter warps only
        if (!(T.flags & QGTC_FLAG_X2_STATIC)) asm volatile("griddepcontrol.wait;\n" ::: "memory");
        if (T.fused) {
            // Fused update: the shared weight matrix W (packed C(bit_w, K2 = N, N2), one K chunk) -> the same u8 core-matrix
            // image as a B stage, once per CTA.  One column per thread, staged through ring slot 0 (free until the main loop).
            const int p8n2 = p8(T.N2), pn2 = (T.epi2 == kEpiFloat && !T.pad128_2) ? p8n2 : p128(T.N2);
            const unsigned long long w_off2 = 4ull * (unsigned long long)pn2;          // S128(K2) == 1
            unsigned char* st0 = b_ring + t * 16;
            if (t < T.NP2) {
                for (int qq = 0; qq < T.bit_w; ++qq) {
                    const unsigned long long idx = (unsigned long long)qq * w_off2 + 4ull * t;
                    const int nb = t < p8n2 ? piece_bytes(idx, T.w2_words) : 0;
                    cp_async16_zfill(st0 + qq * (kConvThreads * 16), nb ? (const void*)(T.w2 + idx) : (const void*)T.w2, nb);
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            if (t < T.NP2) {
                uint32_t o[32];
                assemble_planes_w<BW, false>(T.bit_w, st0, kConvThreads * 16, o);
                unsigned char* dstc = smem + T.off_b2 + (size_t)(t >> 3) * 128 + (t & 7) * 16;
#pragma unroll
                for (int kg = 0; kg < 8; ++kg)
                    *reinterpret_cast<uint4*>(dstc + (size_t)kg * (T.NP2 / 8) * 128) = make_uint4(o[4 * kg], o[4 * kg + 1], o[4 * kg + 2], o[4 * kg + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(b2_ready);
            asm volatile("bar.sync 1, %0;\n" ::"n"(kConvThreads) : "memory");     // slot 0 is re-used by the main loop's first item
        }
        // without the scan every item reads the one zero byte next to n_items (index * 0): no flag register lives across the loop
        const unsigned char* rf = T.may_reuse ? reuse_flag : reinterpret_cast<const unsigned char*>(n_items_slot) + 4;
        unsigned char* ring = b_ring + t * 16;                         // piece (i, plane) of this thread at + (i * bit2 + plane) * 2048
        const bool tr = TRACE && warp == kConvWarp0 && tlane;
        if (tr) QGTC_GTRACE(2, 20);
        const int RB = T.b_ring;                                       // staged items in flight (2..4, by shared-memory budget)
        int s = cg, slot_l = 0, slot_c = 0;        // (cg < 2 <= S)
        uint32_t use = 0;
        int lu_unit = -1;                          // unit whose column pointers / bounds are cached below
        unsigned long long lu_woff = 0;
        const uint32_t* lu_colp[kMaxTasks] = {nullptr, nullptr};
        bool lu_fast[kMaxTasks] = {false, false};
#pragma unroll 1
        for (int it_c = cg - (1 + T.conv_alt) * (RB - 1), it_l = cg; it_c < n_items; it_c += 1 + T.conv_alt, it_l += 1 + T.conv_alt) {
            if (it_l < n_items) {
                const ItemRec I = items[it_l];
                if (!rf[it_l * T.may_reuse]) {     // not a stage re-use: stage this item's packed words
                    const UnitRec& U = units[I.unit];
                    const unsigned long long gdk4 = (unsigned long long)U.gdk * 4;
                    if ((int)I.unit != lu_unit) {
                        // per-unit part of the address arithmetic, once per unit (this stream is the converters' whole critical path
                        // between two items: ~35 instructions of 64-bit index and bounds arithmetic per copy before, clock64 trace)
                        lu_unit = I.unit;
                        lu_woff = gdk4 * (unsigned long long)pn;
#pragma unroll
                        for (int i = 0; i < kMaxTasks; ++i) {
                            const unsigned long long col0 = (unsigned long long)tn_[i] * gdk4;
                            lu_colp[i] = U.x2 + col0;
                            // every piece of this column, in every plane, lies inside the tensor
                            lu_fast[i] = tn_[i] < p8n && col0 + (unsigned long long)(T.bit2 - 1) * lu_woff + gdk4 <= U.x2_words;
                        }
                    }
                    const unsigned long long w_offset = lu_woff;
                    unsigned char* dst = ring + slot_l * T.b_slot_bytes;
#pragma unroll
                    for (int i = 0; i < kMaxTasks; ++i) {
                        if (tj[i] < I.nch && lu_fast[i]) {
                            const uint32_t* src0 = lu_colp[i] + 4 * (I.c0 + tj[i]);
                            uint32_t d0 = smem_u32(dst) + (uint32_t)(i * T.bit2) * (uint32_t)(kConvThreads * 16);
                            for (int qq = 0; qq < T.bit2; ++qq, src0 += w_offset, d0 += (uint32_t)(kConvThreads * 16)) cp_async16_a(d0, src0);
                        } else if (tj[i] < I.nch) {
                            const unsigned long long idx0 = (unsigned long long)tn_[i] * gdk4 + 4ull * (I.c0 + tj[i]);
                            const bool col_ok = tn_[i] < p8n;
                            for (int qq = 0; qq < T.bit2; ++qq) {
                                const unsigned long long idx = idx0 + (unsigned long long)qq * w_offset;
                                const int nb = col_ok ? piece_bytes(idx, U.x2_words) : 0;
                                cp_async16_zfill(dst + (i * T.bit2 + qq) * (kConvThreads * 16), nb ? (const void*)(U.x2 + idx) : (const void*)U.x2, nb);
                            }
                        }
                    }
                    if (++slot_l == RB) slot_l = 0;
                }
            }
            cp_async_commit();
            if (tr) QGTC_GTRACE(2, 21);
            if (it_c < 0) continue;
            const ItemRec I = items[it_c];
            const bool reuse = rf[it_c * T.may_reuse] != 0;
            if (!reuse) {
                if (RB == 4) cp_async_wait<3>();
                else if (RB == 3) cp_async_wait<2>();
                else cp_async_wait<1>();
            }
            if (tr) QGTC_GTRACE(2, 22);
            const int nch = I.nch;
            // the MMAs that read this stage's previous content are done (also orders this arrival after the stage's previous phase)
            gwait(&empty[s], (use & 1) ^ 1, ef, 2);
            if (tr) QGTC_GTRACE(2, 23);
            if (!reuse) {
                const unsigned char* src = ring + slot_c * T.b_slot_bytes;
                if (++slot_c == RB) slot_c = 0;
                unsigned char* stage = b_stages + (size_t)s * T.b_stage_bytes;
#pragma unroll
                for (int i = 0; i < kMaxTasks; ++i) {
                    if (tj[i] >= nch) continue;
                    if (split) {
                        if (i == 0) {
                            uint32_t o[16];
                            if (A1 && mode2) assemble_half_w<BW, true>(T.bit2, src, kConvThreads * 16, half, o);
                            else assemble_half_w<BW, false>(T.bit2, src, kConvThreads * 16, half, o);
                            unsigned char* dsth = stage + (size_t)tj[0] * NP * 128 + (size_t)(tn_[0] >> 3) * 128 + (tn_[0] & 7) * 16 + (size_t)(4 * half) * (NP / 8) * 128;
#pragma unroll
                            for (int kg = 0; kg < 4; ++kg)
                                *reinterpret_cast<uint4*>(dsth + (size_t)kg * (NP / 8) * 128) = make_uint4(o[4 * kg], o[4 * kg + 1], o[4 * kg + 2], o[4 * kg + 3]);
                        }
                        continue;
                    }
                    uint32_t o[32];
                    if (A1 && mode2) assemble_planes_w<BW, true>(T.bit2, src + i * T.bit2 * (kConvThreads * 16), kConvThreads * 16, o);
                    else assemble_planes_w<BW, false>(T.bit2, src + i * T.bit2 * (kConvThreads * 16), kConvThreads * 16, o);
                    // image of one K chunk: [k group (8)][column block][8 columns][16 B]; k group kg = registers 4 kg .. 4 kg + 3
                    unsigned char* dstc = stage + (size_t)tj[i] * NP * 128 + (size_t)(tn_[i] >> 3) * 128 + (tn_[i] & 7) * 16;
#pragma unroll
                    for (int kg = 0; kg < 8; ++kg)
                        *reinterpret_cast<uint4*>(dstc + (size_t)kg * (NP / 8) * 128) = make_uint4(o[4 * kg], o[4 * kg + 1], o[4 * kg + 2], o[4 * kg + 3]);
                }
                fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's (async proxy) reads
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            if (tr) QGTC_GTRACE(2, 24);
            s += 1 + T.conv_alt;
            if (s >= S) { s -= S; ++use; }
        }
        cp_async_wait<0>();
    } else if (warp == kIssuerWarp) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_i8(NP, 128);
        const uint32_t lbo = (uint32_t)(NP / 8) * 128u, sbo = 128u;
        const uint64_t desc0 = make_smem_desc(smem_u32(b_stages), lbo, sbo);
        const uint32_t desc_hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
        const uint32_t stage_lo = (uint32_t)T.b_stage_bytes >> 4, chunk_lo = ((uint32_t)NP * 128u) >> 4, kstep_lo = (2u * lbo) >> 4;
        int s = 0, buf = 0;
        uint32_t use = 0, duse = 0, started = 0;
#pragma unroll 1
        for (int it = 0; it < n_items; ++it) {
            const ItemRec I = items[it];
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NP);
            if (I.flags & 1) {
                if (tlane) QGTC_GTRACE(3, 30);
                gwait(&d_empty[buf], (duse & 1) ^ 1, ef, 3);             // the epilogue has drained this accumulator's previous tile
                started = 0;                                             // the unit's first MMA overwrites the accumulator
            }
            gwait(&full[s], use & 1, ef, 4);
            tc_fence_after();
            if (tlane) QGTC_GTRACE(3, 32);
            const int nch = I.nch;
            const uint32_t fw = lane < kMaxCpsG ? *reinterpret_cast<volatile uint32_t*>(smem + T.off_flags + (s * kMaxCpsG + lane) * 4) : 0u;
            // an all-zero chunk adds nothing and is skipped -- except the unit's very first MMAs, which initialise D
            uint32_t todo = __ballot_sync(0xffffffffu, fw != 0u) & ((1u << nch) - 1u);
            if (!started) todo |= 1u;
            if (elect_one()) {
                const uint32_t a_tmem = tmem_base + (uint32_t)(dcols + s * stage_cols);
                const uint32_t lo_s = lo0 + (uint32_t)s * stage_lo;
                uint32_t acc = started;
#pragma unroll 1
                for (int j = 0; j < nch; ++j) {
                    if ((todo >> j) & 1u) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t desc = ((uint64_t)desc_hi << 32) | (uint64_t)(lo_s + (uint32_t)j * chunk_lo + (uint32_t)ks * kstep_lo);
                            mma_i8_ts(d_tmem, a_tmem + (uint32_t)(32 * j + 8 * ks), desc, idesc, acc);
                            acc = 1u;
                        }
                    }
                }
                tc_commit(&empty[s]);          // arrives when these MMAs have finished reading A (TMEM) and B (smem)
                if (I.flags & 2) tc_commit(&d_full[buf]);
            }
            started = 1u;                      // todo != 0 whenever started was 0
            __syncwarp();
            if (tlane) QGTC_GTRACE(3, 33);
            if (++s == S) { s = 0; ++use; }
            if (I.flags & 2) {
                if (++buf == T.dbufs) { buf = 0; ++duse; }
            }
        }
    } else if (warp == kIssuer2Warp) {
        // ===================== second issuer: the fused update GEMM D2 = A2 x W, one chain per unit =====================
        if (T.fused) {
            const int NP2 = T.NP2;
            const uint32_t idesc2 = make_idesc_i8(NP2, 128);
            const uint32_t lbo2 = (uint32_t)(NP2 / 8) * 128u;
            const uint64_t desc2 = make_smem_desc(smem_u32(smem + T.off_b2), lbo2, 128u);
            const uint32_t desc2_hi = (uint32_t)(desc2 >> 32), desc2_lo = (uint32_t)desc2, kstep2 = (2u * lbo2) >> 4;
            const int nk2 = (T.N + 31) >> 5;                 // K2 = N <= 64: one or two MMAs of K = 32
            gwait(b2_ready, 0, ef, 20);
#pragma unroll 1
            for (int tl = 0; tl < units_here; ++tl) {
                gwait(a2_full, (uint32_t)tl & 1, ef, 21);
                tc_fence_after();
                if (elect_one()) {
                    for (int ks = 0; ks < nk2; ++ks) {
                        const uint64_t desc = ((uint64_t)desc2_hi << 32) | (uint64_t)(desc2_lo + (uint32_t)ks * kstep2);
                        mma_i8_ts(tmem_base + (uint32_t)T.d2_col, tmem_base + (uint32_t)(T.a2_col + 8 * ks), desc, idesc2, ks > 0 ? 1u : 0u);
                    }
                    tc_commit(d2_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (4 warps: one per TMEM lane quarter) =====================
        const int q = warp & 3;
        const int N = T.N, ob = T.out_bit;
        int buf = 0;
        uint32_t duse = 0;
        asm volatile("griddepcontrol.wait;\n" ::: "memory");       // the output may alias memory the previous kernel still reads
#pragma unroll 1
        for (int tl = 0; tl < units_here; ++tl) {
            const UnitRec& P = units[tl];
            const int M = P.M;
            const int m = P.mt * 128 + 32 * q + lane;
            const bool tr = TRACE && warp == kEpiWarp0 && tlane;
            if (tr) QGTC_GTRACE(4, 40);
            mbar_wait<64>(&d_full[buf], duse & 1, ef, 6, false);
            tc_fence_after();
            if (tr) QGTC_GTRACE(4, 41);
            const uint32_t d_tmem = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * NP);
            if (T.fused) {
                // ---- stage 1: requantized bytes of this row (natural order: register j = columns 4j .. 4j+3), NP <= 64
                const bool row_ok = m < p8(M);
                const uint32_t rmask = m < M ? 0xffffffffu : 0u;
                const unsigned keep = (1u << ob) - 1u;
                const int mx = 1 << ob;
                uint32_t nat[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) nat[j] = 0u;
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    if (16 * cb < NP) {
                        int v[16];
                        tmem_ld16(d_tmem + 16 * cb, v);
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            uint32_t four = (mode2 && ob <= 4)
                                                ? requant4<true, true>(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3], dshift, ob, mx, keep)
                                                : requant4<false, false>(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3], dshift, ob, mx, keep);
                            const int nv = N - (16 * cb + 4 * g4);       // valid columns of this group of four
                            const uint32_t mask = nv >= 4 ? rmask : (nv <= 0 ? 0u : (rmask & ((1u << (8 * nv)) - 1u)));
                            nat[4 * cb + g4] = four & mask;
                        }
                    }
                }
                if (T.write_out1 && row_ok) {        // the first op's own result R(ob, M, N), when the caller wants it
                    uint32_t* out = reinterpret_cast<uint32_t*>(P.out);
                    const int wpr = s128(N) * 4;
                    const unsigned long long opt_offset = (unsigned long long)p8(M) * wpr;
                    for (int p = 0; p < ob; ++p) {
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            if (32 * w < NP) {
                                uint32_t word = 0u;
#pragma unroll
                                for (int g = 0; g < 8; ++g) word |= nibble_of_plane(nat[8 * w + g], p) << (28 - 4 * g);
                                out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w] = word;
                            }
                        }
                        for (int w = (NP + 31) / 32; w < wpr; ++w) out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w] = 0u;
                    }
                }
                // ---- A2 operand in the K permutation W's image uses: byte i of register 8 s + c <-> BIT c + 8 i of packed word s, and
                // the packed words are MSB-first (column e of a 32-group is bit 31 - e): byte i of register 8 s + c = value of
                // column 32 s + 31 - c - 8 i, i.e. the natural gather below with the register order and the bytes reversed.
                uint32_t a2[16];
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll
                    for (int cq = 0; cq < 2; ++cq) {
                        const uint32_t r0 = nat[8 * s2 + cq], r1 = nat[8 * s2 + cq + 2], r2 = nat[8 * s2 + cq + 4], r3 = nat[8 * s2 + cq + 6];
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt) {
                            const uint32_t sel = (uint32_t)(tt | ((4 + tt) << 4));
                            a2[8 * s2 + 7 - (4 * cq + tt)] = __byte_perm(__byte_perm(r0, r1, sel), __byte_perm(r2, r3, sel), 0x0145u);
                        }
                    }
                }
                const uint32_t a2_tmem = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)T.a2_col;
                if (N > 32) tmem_st16r(a2_tmem, a2);
                else tmem_st8(a2_tmem, a2);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(a2_full);            // -> second issuer
                    mbar_arrive(&d_empty[buf]);      // the main accumulator is drained
                }
                // ---- stage 2: the update GEMM's accumulator -> second output
                gwait(d2_full, (uint32_t)tl & 1, ef, 22);
                tc_fence_after();
                const int N2 = T.N2, NP2 = T.NP2, ob2 = T.out_bit2;
                const uint32_t d2_tmem = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)T.d2_col;
                if (T.epi2 == kEpiFloat) {
                    float* out2 = reinterpret_cast<float*>(P.out2);
                    for (int cb = 0; cb < NP2 / 16; ++cb) {
                        int v[16];
                        tmem_ld16(d2_tmem + 16 * cb, v);
                        if (m < M) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const int n = 16 * cb + j;
                                if (n < N2) out2[(size_t)m * N2 + n] = (float)v[j];
                            }
                        }
                    }
                } else {
                    uint32_t* out2 = reinterpret_cast<uint32_t*>(P.out2);
                    const int wpr2 = s128(N2) * 4;
                    const unsigned long long opt2 = (unsigned long long)p8(M) * wpr2;
                    uint32_t* out2_row = out2 + (unsigned long long)m * wpr2;
                    if (BW > 0) {
                        epi_bits_tile<BW, false, false, false, 1>(d2_tmem, NP2, 0, N2, ob2, 0, rmask, row_ok, out2_row, opt2);
                    } else
                    switch (ob2) {
                        case 1: epi_bits_tile<1, false, false, false, 1>(d2_tmem, NP2, 0, N2, ob2, 0, rmask, row_ok, out2_row, opt2); break;
                        case 2: epi_bits_tile<2, false, false, false, 1>(d2_tmem, NP2, 0, N2, ob2, 0, rmask, row_ok, out2_row, opt2); break;
                        case 4: epi_bits_tile<4, false, false, false, 1>(d2_tmem, NP2, 0, N2, ob2, 0, rmask, row_ok, out2_row, opt2); break;
                        case 8: epi_bits_tile<8, false, false, false, 1>(d2_tmem, NP2, 0, N2, ob2, 0, rmask, row_ok, out2_row, opt2); break;
                        default: epi_bits_tile<0, false, false, false, 1>(d2_tmem, NP2, 0, N2, ob2, 0, rmask, row_ok, out2_row, opt2); break;
                    }
                    if (row_ok) {
                        for (int w = (NP2 + 31) / 32; w < wpr2; ++w)
                            for (int p = 0; p < ob2; ++p) out2[(unsigned long long)p * opt2 + (unsigned long long)m * wpr2 + w] = 0u;
                    }
                }
                tc_fence_before();               // D2 / A2 are free again once all four warps have passed (next unit's a2_full)
                __syncwarp();
                if (tr) QGTC_GTRACE(4, 42);
                if (++buf == T.dbufs) { buf = 0; ++duse; }
                continue;
            }
            if (T.epi == kEpiFloat) {
                float* out = reinterpret_cast<float*>(P.out);
                for (int cb = 0; cb < NP / 16; ++cb) {
                    int v[16];
                    tmem_ld16(d_tmem + 16 * cb, v);
                    if (m < M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = 16 * cb + j;
                            if (n < N) out[(size_t)m * N + n] = (float)(v[j] >> dshift);
                        }
                    }
                }
            } else if (T.epi == kEpiBitsRow) {
                uint32_t* out = reinterpret_cast<uint32_t*>(P.out);
                const int wpr = s128(N) * 4;
                const unsigned long long opt_offset = (unsigned long long)p8(M) * wpr;
                const bool row_ok = m < p8(M);
                const uint32_t rmask = m < M ? 0xffffffffu : 0u;
                uint32_t* out_row = out + (unsigned long long)m * wpr;
#define QGTC_GRP_EPI(OBV, NNV, FASTV) epi_bits_tile<OBV, NNV, FASTV, false, 1>(d_tmem, NP, 0, N, ob, dshift, rmask, row_ok, out_row, opt_offset)
                if (BW > 0) {
                    if (mode2) QGTC_GRP_EPI(BW, true, (BW <= 4));
                    else QGTC_GRP_EPI(BW, false, false);
                } else if (mode2) {   // 8 x sums, non-negative: byte-parallel requantize up to 4 output bits
                    switch (ob) {
                        case 1: QGTC_GRP_EPI(1, true, true); break;
                        case 2: QGTC_GRP_EPI(2, true, true); break;
                        case 4: QGTC_GRP_EPI(4, true, true); break;
                        case 8: QGTC_GRP_EPI(8, true, false); break;
                        default: if (ob <= 4) QGTC_GRP_EPI(0, true, true); else QGTC_GRP_EPI(0, true, false); break;
                    }
                } else {
                    switch (ob) {
                        case 1: QGTC_GRP_EPI(1, false, false); break;
                        case 2: QGTC_GRP_EPI(2, false, false); break;
                        case 4: QGTC_GRP_EPI(4, false, false); break;
                        case 8: QGTC_GRP_EPI(8, false, false); break;
                        default: QGTC_GRP_EPI(0, false, false); break;
                    }
                }
#undef QGTC_GRP_EPI
                if (row_ok) {   // zero pad words of the row (output buffers are not pre-zeroed)
                    for (int w = (NP + 31) / 32; w < wpr; ++w)
                        for (int p = 0; p < ob; ++p) out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w] = 0u;
                }
            } else {
                // column-major planes C128(ob, M, N) (bitMM2Bit_col, intended semantics): word (p, n, mw) holds rows
                // 32 mw .. 32 mw + 31 of column n, MSB-first.  This warp's 32 rows are exactly word mw = 4 mt + q of every
                // column: one ballot per (column, plane); every word of the image is written by exactly one warp
                // (pad columns and pad rows as zeros), so the output needs no pre-zeroing.
                uint32_t* out = reinterpret_cast<uint32_t*>(P.out);
                const unsigned long long kw = (unsigned long long)s128(M) * 4;
                const unsigned long long plane = kw * (unsigned long long)p128(N);
                const int mw = P.mt * 4 + q;
                const unsigned keep = (1u << ob) - 1u;
                if ((unsigned long long)mw < kw) {
                    for (int cb = 0; cb < NP / 16; ++cb) {
                        int v[16];
                        tmem_ld16(d_tmem + 16 * cb, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = 16 * cb + j;
                            const uint32_t val = (m < M && n < N) ? ((uint32_t)requant(v[j] >> dshift, ob) & keep) : 0u;
                            for (int p = 0; p < ob; ++p) {
                                const uint32_t word = __brev(__ballot_sync(0xffffffffu, (val >> p) & 1u));
                                if (lane == ((j + p) & 31)) out[(unsigned long long)p * plane + (unsigned long long)n * kw + mw] = word;
                            }
                        }
                    }
                    for (int n = NP + lane; n < p128(N); n += 32)
                        for (int p = 0; p < ob; ++p) out[(unsigned long long)p * plane + (unsigned long long)n * kw + mw] = 0u;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
            if (tr) QGTC_GTRACE(4, 42);
            if (++buf == T.dbufs) { buf = 0; ++duse; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 32) QGTC_GTRACE(7, 4);
    if (warp == kIssuerWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

struct GrpPlan {
    int NP, cps, stages, dbufs, groups, a_slot, b_slot, b_stage, b_ring, off_bstage, off_bring, off_bars, off_flags, off_misc, off_units, smem_total;
    int NP2, a2_col, d2_col, off_b2;      // fused update (NP2 == 0: none)
};

// pipeline geometry for (N, bit1, bit2) and, with N2 > 0, a fused update GEMM of width N2 with bit_w weight planes; false when
// the shape does not fit this engine
bool plan_for(int N, int bit1, int bit2, GrpPlan& pl, int N2 = 0, int bit_w = 0) {
    pl.NP = ((p8(N) + 15) / 16) * 16;
    if (pl.NP > 256 || bit1 < 1 || bit1 > 8 || bit2 < 1 || bit2 > 8) return false;
    pl.NP2 = N2 > 0 ? ((p8(N2) + 15) / 16) * 16 : 0;
    if (N2 > 0 && (pl.NP > 64 || pl.NP2 > 128 || bit_w < 1 || bit_w > 8)) return false;     // A2 = one requantized tile of K2 = N <= 64
    const int extra_cols = N2 > 0 ? 16 + pl.NP2 : 0;      // A2 operand (16 columns) + second accumulator
    int cps = 256 / pl.NP;                       // the converters take at most kMaxTasks x 128 (chunk, column) tasks per item
    if (cps > kMaxCpsG) cps = kMaxCpsG;
    while (cps > 1 && cps * bit1 > 8) cps >>= 1; // staging: at most 8 pieces of 16 B per lane and item
    if (cps == 3) cps = 2;
    pl.cps = cps;
    pl.dbufs = (2 * pl.NP + 2 * cps * 32 + extra_cols <= 512) ? 2 : 1;
    for (int stages = kMaxStagesG; stages >= 2; --stages) {
        if (pl.dbufs * pl.NP + stages * cps * 32 + extra_cols > 512) continue;
        pl.a2_col = pl.dbufs * pl.NP + stages * cps * 32;
        pl.d2_col = pl.a2_col + 16;
        pl.stages = stages;
        pl.groups = stages < 2 ? 1 : 2;          // G <= S keeps the parity-only mbarrier waits unambiguous (see gemm_tc05.cu)
        pl.a_slot = bit1 * cps * 512;
        const int tasks_per_thread = (cps * pl.NP + kConvThreads - 1) / kConvThreads;      // 1 or 2 (cps * NP <= 256)
        pl.b_slot = tasks_per_thread * bit2 * kConvThreads * 16;
        if (bit_w * kConvThreads * 16 > pl.b_slot) pl.b_slot = bit_w * kConvThreads * 16;   // slot 0 also stages W's planes once
        pl.b_stage = cps * pl.NP * 128;
        int off = kExpWarps * kARing * pl.a_slot;
        off = (off + 1023) & ~1023;
        pl.off_bstage = off;
        off += stages * pl.b_stage;
        pl.off_b2 = off;                         // W's image of the fused update (one K chunk)
        off += pl.NP2 * 128;
        pl.off_bring = off;
        // converter staging ring: as deep as shared memory allows (the converters' loads are issued b_ring - 1 items ahead;
        // two is not enough to cover the L2 / HBM latency of a short item), leaving room for the per-CTA work lists
        pl.b_ring = kMaxBRing;
        while (pl.b_ring > 2 && off + pl.b_ring * pl.b_slot + 512 + 24 * 1024 > 227 * 1024) --pl.b_ring;
        off += pl.b_ring * pl.b_slot;
        pl.off_bars = off;
        off += (2 * kMaxStagesG + 4 + 3) * 8;
        pl.off_flags = off;
        off += kMaxStagesG * kMaxCpsG * 4;
        pl.off_misc = off;
        off += 16;
        pl.off_units = off;                      // per-CTA work lists (unit + item records): sized at launch
        pl.smem_total = off + 1024;              // slack for the manual 1024-byte alignment
        if (pl.smem_total + 4 * 1024 <= 227 * 1024) return true;
    }
    return false;
}

}  // namespace

struct GroupedProblemHost {     // what the ABI layer hands over (host memory)
    const uint32_t* x1;
    const uint32_t* x2;
    void* out;
    const uint32_t* bitmap;
    unsigned long long x1_words, x2_words;
    int M, K;
    void* out2;                 // fused update only
};
struct GroupedFusedHost {       // the second op of a fused launch: out2 = op2(requant(op1), W)
    const uint32_t* w;
    unsigned long long w_words;
    int N2, bit_w, out_bit2, epi2, pad128_2, write_out1;
};

bool grp_supports(int N, int bit1, int bit2, int epi) {
    if (epi != kEpiBitsRow && epi != kEpiFloat && epi != kEpiBitsCol) return false;
    GrpPlan pl;
    return plan_for(N, bit1, bit2, pl);
}
bool grp_supports_fused(int N, int bit1, int bit2, int N2, int bit_w, int epi2) {
    if (epi2 != kEpiBitsRow && epi2 != kEpiFloat) return false;
    GrpPlan pl;
    return plan_for(N, bit1, bit2, pl, N2, bit_w);
}

unsigned long long* g_grp_trace = nullptr;     // clock64 trace buffer of the debug instance (QGTC_GRP_TRACE=1)
constexpr size_t kGrpTraceWords = (size_t)kTraceRoles * kTraceEvents * 2;

// One op over `count` problems.  Returns QGTC_ERR_UNSUPPORTED (before any work is enqueued) when the shape does not fit.
int launch_gemm_grouped(const GroupedProblemHost* probs, int count, int N, int bit1, int bit2, int out_bit, int epi, int pad128,
                        unsigned flags, cudaStream_t st, const GroupedFusedHost* fu) {
    GrpPlan pl;
    if (count <= 0) return QGTC_ERR_INVALID_ARG;
    if (!grp_supports(N, bit1, bit2, epi)) return QGTC_ERR_UNSUPPORTED;
    if (fu) {
        if (epi != kEpiBitsRow || !grp_supports_fused(N, bit1, bit2, fu->N2, fu->bit_w, fu->epi2)) return QGTC_ERR_UNSUPPORTED;
        if (!plan_for(N, bit1, bit2, pl, fu->N2, fu->bit_w)) return QGTC_ERR_UNSUPPORTED;
    } else if (!plan_for(N, bit1, bit2, pl)) {
        return QGTC_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < count; ++i)    // a single problem larger than one launch can hold: not this engine
        if ((long long)((p8(probs[i].M) + 127) / 128) > (long long)kNumSMs * kMaxUnitsPerCta) return QGTC_ERR_UNSUPPORTED;
    unsigned int* ef = tc05_error_flag_device();
    static const bool want_trace = [] { const char* t = getenv("QGTC_GRP_TRACE"); return t && t[0] == '1'; }();
    if (want_trace && !g_grp_trace && cudaMalloc(&g_grp_trace, kGrpTraceWords * 8) != cudaSuccess) { cudaGetLastError(); g_grp_trace = nullptr; }
    static thread_local GrpArgs T;     // ~22 KB: filled per launch; passed by value as the kernel parameter, or its used part in a smaller block
    int first = 0;
    int limit = kGrpMaxProblems;       // problems per launch; halved when the per-CTA work lists would not fit shared memory
    while (first < count) {
        // as many problems as the parameter table and the per-CTA work lists hold
        long long units = 0;
        int n = 0, max_nst = 1, max_gdk = 1, any_bitmap = 0;
        while (first + n < count && n < limit) {
            const GroupedProblemHost& h = probs[first + n];
            const long long u = (p8(h.M) + 127) / 128;
            if (n > 0 && units + u > (long long)kNumSMs * kMaxUnitsPerCta) break;
            GrpProblem& p = T.p[n];
            p.x1 = h.x1; p.x2 = h.x2; p.out = h.out; p.out2 = fu ? h.out2 : nullptr; p.pad_ = 0;
            p.bitmap = bit1 == 1 ? h.bitmap : nullptr;
            if (p.bitmap) any_bitmap = 1;
            p.x1_words = h.x1_words; p.x2_words = h.x2_words;
            p.M = h.M; p.K = h.K;
            p.unit0 = (int)units;
            p.mt = 0;
            if (units + u <= kUnitTable)
                for (long long k = 0; k < u; ++k) T.unit_z[units + k] = (unsigned char)n;
            const int nst = (s128(h.K) + pl.cps - 1) / pl.cps;
            if (nst > max_nst) max_nst = nst;
            if (s128(h.K) > max_gdk) max_gdk = s128(h.K);
            units += u;
            ++n;
        }
        {
            const long long grid_ = units < kNumSMs ? units : kNumSMs;
            const long long upc = (units + grid_ - 1) / grid_;
            const long long lists = upc * (long long)sizeof(UnitRec) + upc * max_nst * (long long)(sizeof(ItemRec) + 1);
            if (pl.smem_total + lists > 227 * 1024) {
                if (n <= 1) return QGTC_ERR_UNSUPPORTED;
                limit = n / 2;
                continue;              // retry this launch with fewer problems
            }
        }
        T.count = n;
        T.units = (int)units;
        T.N = N; T.NP = pl.NP; T.bit1 = bit1; T.bit2 = bit2; T.out_bit = out_bit; T.epi = epi; T.pad128 = pad128;
        T.cps = pl.cps; T.stages = pl.stages; T.dbufs = pl.dbufs; T.groups = pl.groups;
        T.a_slot_bytes = pl.a_slot; T.b_slot_bytes = pl.b_slot; T.b_stage_bytes = pl.b_stage; T.b_ring = pl.b_ring;
        T.off_bstage = pl.off_bstage; T.off_bring = pl.off_bring; T.off_bars = pl.off_bars; T.off_flags = pl.off_flags; T.off_misc = pl.off_misc;
        T.off_units = pl.off_units;
        T.fused = fu ? 1 : 0;
        T.N2 = fu ? fu->N2 : 0; T.NP2 = pl.NP2; T.bit_w = fu ? fu->bit_w : 0; T.out_bit2 = fu ? fu->out_bit2 : 0;
        T.epi2 = fu ? fu->epi2 : 0; T.pad128_2 = fu ? fu->pad128_2 : 0; T.write_out1 = fu ? fu->write_out1 : 0;
        T.a2_col = pl.a2_col; T.d2_col = pl.d2_col; T.off_b2 = pl.off_b2;
        T.w2 = fu ? fu->w : nullptr; T.w2_words = fu ? fu->w_words : 0;
        T.np_inv = (unsigned int)((0x100000000ull + (unsigned long long)pl.NP - 1ull) / (unsigned long long)pl.NP);
        T.cps_inv = pl.cps == 1 ? 0u : (unsigned int)((0x100000000ull + (unsigned long long)pl.cps - 1ull) / (unsigned long long)pl.cps);
        T.flags = flags;
        T.error_flag = ef;
        T.trace = g_grp_trace;
        if (g_grp_trace) cudaMemsetAsync(g_grp_trace, 0, kGrpTraceWords * 8, st);
        const int grid = T.units < kNumSMs ? T.units : kNumSMs;
        const int units_per_cta = (T.units + grid - 1) / grid;
        T.units_cap = units_per_cta;
        {   // stage re-use needs two items of one CTA with the same B chunk: problems sharing x2 (the weight products) or one problem
            // with more tiles than CTAs.  A cheap sufficient test -- a miss only costs the optimisation, never a result.
            static const bool force_scan = [] { const char* e = getenv("QGTC_GRP_REUSE_SCAN"); return e && e[0] == '1'; }();
            int mr = force_scan ? 1 : 0;
            for (int i = 0; i < n && !mr; ++i)
                mr = (i > 0 && T.p[i].x2 == T.p[i - 1].x2) || (p8(T.p[i].M) + 127) / 128 > grid;
            T.may_reuse = mr;
        }
        {
            static const bool no_alt = [] { const char* e = getenv("QGTC_GRP_CONV_ALT"); return e && e[0] == '0'; }();
            T.conv_alt = (!no_alt && pl.cps * pl.NP <= 64) ? 1 : 0;
        }
        const int smem_total = pl.smem_total + units_per_cta * (int)sizeof(UnitRec) + units_per_cta * max_nst * ((int)sizeof(ItemRec) + 1);
        T.max_nst = max_nst;
        T.any_bitmap = any_bitmap;
        {   // expansion mode 2 (1-bit A only): the 8 x scaled sums must fit an int32 and the scaled features a byte
            static const bool no_mode2 = [] { const char* e = getenv("QGTC_GRP_MODE2"); return e && e[0] == '0'; }();
            const double max_sum = (double)max_gdk * 128.0 * (double)((1u << bit2) - 1u);
            T.mode2 = (!no_mode2 && bit1 == 1 && bit2 <= 5 && max_sum * 8.0 < 2147483648.0) ? 1 : 0;
        }
        auto launch_with = [&](auto kfn, const auto& TA) -> int {
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return record_cuda_error(e, "cudaFuncSetAttribute(gemm_grp)");
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = smem_total;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous kernel's tail
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            static const bool no_pdl = [] { const char* e = getenv("QGTC_GRP_NO_PDL"); return e && e[0] == '1'; }();
            cfg.numAttrs = no_pdl ? 0 : 1;
            e = cudaLaunchKernelEx(&cfg, kfn, TA);
            note_launch();
            if (e != cudaSuccess) return record_cuda_error(e, "cudaLaunchKernelEx(gemm_grp)");
            return check_launch("gemm_grp_kernel");
        };
        auto launch = [&](auto kfn) -> int { return launch_with(kfn, T); };
        // the same launch with a smaller parameter block: head + the n problems + the unit table
        auto shrink = [&](auto& TS) {
            static_cast<GrpHead&>(TS) = static_cast<const GrpHead&>(T);
            memcpy(TS.p, T.p, (size_t)n * sizeof(GrpProblem));
            memcpy(TS.unit_z, T.unit_z, (size_t)T.units);
        };
        int rc;
        // compile-time width instances (QGTC_GRP_WIDTH=0: the generic instances, for A/B runs)
        static const bool no_width = [] { const char* e = getenv("QGTC_GRP_WIDTH"); return e && e[0] == '0'; }();
        int bw = 0;
        if (!no_width && (bit2 == 1 || bit2 == 2 || bit2 == 4 || bit2 == 8) && (epi != kEpiBitsRow || T.fused || out_bit == bit2) &&
            (epi != kEpiBitsCol || out_bit == bit2) && (bit1 == 1 || bit1 == bit2) &&
            (!T.fused || (out_bit == bit2 && T.bit_w == bit2 && (T.epi2 == kEpiFloat || T.out_bit2 == bit2))))
            bw = bit2;
        static const bool no_tiers = [] { const char* e = getenv("QGTC_GRP_PARAM_TIERS"); return e && e[0] == '0'; }();
        const int tier = (no_tiers || g_grp_trace || bw == 0 || (bit1 != 1 && bw == 1)) ? 0
                         : (n <= kTierSProblems && T.units <= kTierSUnits) ? 2
                         : (n <= kTierMProblems && T.units <= kTierMUnits) ? 1 : 0;
#define QGTC_GRP_TIER_CASE(A1_, BW_)                                                         \
    if (tier == 2) { shrink(TS); rc = launch_with(gemm_grp_kernel<A1_, false, BW_, GrpArgsS>, TS); } \
    else { shrink(TM); rc = launch_with(gemm_grp_kernel<A1_, false, BW_, GrpArgsM>, TM); }
        if (tier != 0) {
            static thread_local GrpArgsS TS;
            static thread_local GrpArgsM TM;
            if (bit1 == 1) {
                switch (bw) {
                    case 1: QGTC_GRP_TIER_CASE(true, 1) break;
                    case 2: QGTC_GRP_TIER_CASE(true, 2) break;
                    case 4: QGTC_GRP_TIER_CASE(true, 4) break;
                    default: QGTC_GRP_TIER_CASE(true, 8) break;
                }
            } else {
                switch (bw) {
                    case 2: QGTC_GRP_TIER_CASE(false, 2) break;
                    case 4: QGTC_GRP_TIER_CASE(false, 4) break;
                    default: QGTC_GRP_TIER_CASE(false, 8) break;
                }
            }
        } else if (g_grp_trace) rc = bit1 == 1 ? launch(gemm_grp_kernel<true, true>) : launch(gemm_grp_kernel<false, true>);
#undef QGTC_GRP_TIER_CASE
        else if (bit1 == 1) {
            switch (bw) {
                case 1: rc = launch(gemm_grp_kernel<true, false, 1>); break;
                case 2: rc = launch(gemm_grp_kernel<true, false, 2>); break;
                case 4: rc = launch(gemm_grp_kernel<true, false, 4>); break;
                case 8: rc = launch(gemm_grp_kernel<true, false, 8>); break;
                default: rc = launch(gemm_grp_kernel<true, false>); break;
            }
        } else {
            switch (bw) {
                case 2: rc = launch(gemm_grp_kernel<false, false, 2>); break;
                case 4: rc = launch(gemm_grp_kernel<false, false, 4>); break;
                case 8: rc = launch(gemm_grp_kernel<false, false, 8>); break;
                default: rc = launch(gemm_grp_kernel<false, false>); break;
            }
        }
        if (rc != QGTC_OK) return rc;
        first += n;
    }
    return QGTC_OK;
}

// debugging aid (not part of the public ABI): copy the last clock64 trace of the grouped engine to the host
extern "C" int qgtc_debug_grp_trace(unsigned long long* dst, size_t max_words) {
    if (!g_grp_trace) return 0;
    const size_t n = max_words < kGrpTraceWords ? max_words : kGrpTraceWords;
    cudaDeviceSynchronize();
    cudaMemcpy(dst, g_grp_trace, n * 8, cudaMemcpyDeviceToHost);
    return (int)n;
}

}  // namespace qgtc
