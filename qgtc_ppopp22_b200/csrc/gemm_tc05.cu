// Bit-GEMM engine 2: tcgen05.mma kind::i8 (u8 x u8 -> s32) with TMEM accumulators, sm_100a.
//
// Data flow of one CTA (persistent, one per SM), for a 128-row adjacency tile:
//
//   packed 1-bit adjacency tile  --TMA (cp.async.bulk.tensor, 64B/32B swizzle)--> smem ring (8 slots)
//        --LDS.128--> registers --one AND per register (expansion mode 2)--> 32 registers
//        = 128 u8 of one row  --tcgen05.st.32x32b.x32--> TMEM (the MMA's A operand, row = lane)
//   feature "image" (u8, K-major core matrices, produced once per call by prep_b_kernel)
//        --cp.async.bulk + mbarrier--> smem stage (the MMA's B operand, no-swizzle descriptor)
//   tcgen05.mma.cta_group::1.kind::i8  D[tmem] (+)= A[tmem] x B[smem], 4 per 128-bit K chunk
//   tcgen05.ld --> requantize + pack planes (or float) in registers --> global stores
//
// The adjacency never exists as bytes in shared or global memory: 1 bit per element is read from
// HBM, the 8x inflation happens in registers and goes straight to tensor memory.  The K order
// inside a 32-element step is permuted (byte i of column c is bit c+8i of the packed word) so that a
// register of four A bytes is a mask of the packed word; the feature image uses the same permutation
// (and, in expansion mode 2, carries the power of two that turns the masked bit back into a 1).
//
// A pipeline stage ("item") covers `cps` K chunks of 128 bits (cps * NP = 256, so a stage is always 16
// column-64-equivalent MMAs) to amortise the per-stage barrier round trips.
// Warp roles: warps 0-15 adjacency expanders (warp = TMEM lane quarter of one of four groups; a group
// expands whole items, the groups take the items round-robin), 16-23 epilogue (lane quarter x column half),
// 24 feature loader, 25 adjacency TMA loader, 26-27 MMA issuers (items round-robin; an item is signalled on its owner's `full`
// barrier and only its owner commits to its `empty` barrier -- see the issuer loop; pair mode keeps the in-order protocol).
// The hot configurations (seven chunks per item on the 4-bit path, four on the u8 path) are compile-time instances (CPST).
// Zero-tile jumping: an expander warp whose 32 rows x 128 bits are all zero reports it; if the whole
// 128 x 128-bit tile chunk is zero the MMA issuer skips its 4 MMAs.  With a precomputed tile bitmap
// (qgtc_tile_bitmap) all-zero stages are skipped by every role.
#include <cuda.h>
#include <mutex>
#include <utility>
#include <vector>

#include <cstdlib>

#include "tc05_common.cuh"

namespace qgtc {

using namespace tc05;

namespace {

// Warp layout.  QGTC_TC05_WIDE = 0: 16 expander warps (a warp takes all chunks of its item), 8 epilogue warps, 28 warps,
// 72 registers.  QGTC_TC05_WIDE = 1: 24 expander warps (two warps per lane quarter share an item's chunks: the stage is
// filled sooner after it is freed), 4 epilogue warps, 32 warps, 64 registers.
#ifndef QGTC_TC05_WIDE
#define QGTC_TC05_WIDE 0
#endif
constexpr int kProducerWarps = QGTC_TC05_WIDE ? 24 : 16;
constexpr int kEpilogueWarp0 = kProducerWarps;
constexpr int kEpilogueWarps = QGTC_TC05_WIDE ? 4 : 8;
constexpr int kEpiStep = kEpilogueWarps / 4;   // epilogue warps per TMEM lane quarter: they take the 32-column words round-robin
constexpr int kLoaderWarp = kEpilogueWarp0 + kEpilogueWarps;
constexpr int kALoaderWarp = kLoaderWarp + 1;
constexpr int kMmaWarp = kLoaderWarp + 2;      // first MMA issuer (also allocates TMEM); issuer i is warp kMmaWarp + i
constexpr int kMaxIssuers = 2;    // (the per-owner `full` barriers below assume at most two)    // Tc05Args::ni of them are launched; they take the pipeline items round-robin
constexpr int kMmaWarp2 = kMmaWarp + 1;
constexpr int kMaxThreads = (kMmaWarp + kMaxIssuers) * 32;
constexpr int kRawSlots = 8;      // TMA ring of packed adjacency tiles (128 rows x 16*cps bytes each)
constexpr int kMaxStages = 4;
constexpr int kMaxCps = 4;        // K chunks (128 bits each) per pipeline stage, u8 operand path
constexpr int kMaxCps4 = 8;       // ... 4-bit operand path (a chunk takes 16 TMEM columns instead of 32)
constexpr int kSfCols = 16;       // 4-bit path: TMEM columns holding the (all-ones) block scale factors: 4 for A, 12 for B (N up to 256)
constexpr int kTmemCols = 512;
// split release of a stage (first / second half of its chunks): built and measured on B200 -- parity-green, 4 % slower on the
// u8 path (the second wait sits in the middle of the fill and the extra state spills), no change on the 4-bit path; kept as a
// compile-time experiment (-DQGTC_TC05_SPLIT=1), profiles/r2_tc05_ab_call23_split_release.txt, r2_tc05_ab_call29_split_u8.txt
#ifndef QGTC_TC05_SPLIT
#define QGTC_TC05_SPLIT 0
#endif
#ifndef QGTC_TC05_EPI_POLL1
#define QGTC_TC05_EPI_POLL1 1
#endif
#ifndef QGTC_TC05_DENSE_ISSUE
#define QGTC_TC05_DENSE_ISSUE 1
#endif
#ifndef QGTC_TC05_OWN_FULL
#define QGTC_TC05_OWN_FULL 1
#endif
#ifndef QGTC_TC05_OWNER_COMMIT
#define QGTC_TC05_OWNER_COMMIT 1
#endif

struct Tc05Args {
    GemmArgs g;
    const uint8_t* bimg;            // feature image: [count][gdk][NP x 128 B]
    unsigned long long bimg_gstride;  // bytes per problem (0 when every problem aliases one feature matrix)
    int NP;                         // padded N (multiple of 16, <= 256)
    int stages;                     // A (TMEM) / B (smem) ring depth
    int cps;                        // K chunks per stage (1, 2 or 4)
    int nst;                        // stages per tile = ceil(gdk / cps)
    int wch;                        // chunks each expander warp handles per item (= cps)
    int wpi;                        // expander warps per item (4: one per TMEM lane quarter); the 16 warps form 16 / wpi groups
    int dbufs;                      // accumulator buffers (1 or 2)
    int mtiles;                     // 128-row tiles per problem
    int total_tiles;                // count * mtiles
    int units;                      // work units: tiles, or 256-row super-tiles in pair mode (count * ceil(mtiles / 2))
    int spg;                        // pair mode: super-tiles per problem
    int epi;
    int a_batched;                  // 1: the tensor map's third coordinate is the problem index
    const uint32_t* bitmap;         // optional tile bitmap (qgtc_tile_bitmap): [count][mtiles][bmw] words of chunk bits
    int bmw;                        // bitmap words per tile = ceil(gdk / 32)
    int rslots;                     // raw adjacency ring slots in use (<= kRawSlots)
    unsigned int cps_inv;           // ceil(2^32 / cps): stage of a chunk by one multiply
    int amode;                      // 3 = 4-bit operand path (kind::mxf4, adjacency bit -> E2M1 nibble, features as E2M1); else adjacency expansion: 0 = bit -> 1 (shift + AND), 1 = bit -> 128 (IMAD + AND), 2 = bit c -> 2^(c%4) (AND only; feature rows pre-scaled)
    int ni;                         // number of MMA issuer warps (2..kMaxIssuers)
    unsigned int one;               // == 1, opaque to the compiler (keeps mul.hi from becoming a shift)
    unsigned int epi_sleep;         // ns the epilogue warps sleep between polls of the accumulator barrier
    unsigned int* error_flag;       // device word set by the watchdog
    unsigned long long* trace;      // optional clock64 trace of CTA 0 (QGTC_TC05_TRACE=1), else null
};
constexpr int kTraceItems = 512;    // per role: [item][8] timestamps
#define QGTC_TRACE(role, item, slot)                                                             \
    do {                                                                                         \
        if (DBG && T.trace && blockIdx.x == 0 && (item) < kTraceItems)                           \
            T.trace[((role) * kTraceItems + (item)) * 8 + (slot)] = (unsigned long long)clock64(); \
    } while (0)

// PTX wrappers and the epilogue helpers live in tc05_common.cuh (shared with gemm_grp.cu)
constexpr int kFlagsOff = (5 * kMaxStages + 2 * kRawSlots + 6) * 8;   // flag bytes after the mbarrier arrays
struct SmemLayout {
    // offsets in bytes from the 1024-byte aligned base
    int raw0, raw_slot_bytes, b_stage0, b_stage_bytes, bars, flags, misc, total;
};
__host__ __device__ inline SmemLayout smem_layout(int NP, int stages, int cps, int pair = 0, int rslots = kRawSlots, int f4 = 0) {
    SmemLayout L;
    L.raw0 = 0;
    L.raw_slot_bytes = 128 * 16 * cps;                 // 2/4/8 KB, keeps every slot 1024-byte aligned for the swizzle
    if (L.raw_slot_bytes < 1024) L.raw_slot_bytes = 1024;
    L.raw_slot_bytes = (L.raw_slot_bytes + 1023) & ~1023;
    L.b_stage0 = L.raw0 + rslots * L.raw_slot_bytes;
    L.b_stage_bytes = cps * NP * ((pair || f4) ? 64 : 128);    // pair mode: each CTA holds half of the feature columns; 4-bit path: half the bytes
    L.bars = L.b_stage0 + stages * L.b_stage_bytes;
    L.flags = L.bars + kFlagsOff;
    L.misc = L.flags + kMaxStages * 32;                // flags: [stage][chunk <= 8][lane quarter]
    L.total = L.misc + 32;
    return L;
}

// word w of a tile's chunk bits; without a bitmap every existing chunk counts
__device__ __forceinline__ uint32_t tile_word(const uint32_t* bm_tile, int w, int gdk) {
    const int left = gdk - 32 * w;
    const uint32_t valid = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
    return bm_tile ? (__ldg(bm_tile + w) & valid) : valid;
}

// Work-unit geometry.  Without pairing a unit is one 128-row tile; in pair mode it is a 256-row super-tile whose
// halves go to the two CTAs of the cluster (the second half may lie beyond the matrix: TMA then reads zeros and the
// epilogue writes nothing).
struct UnitGeo {
    int z, mt;          // problem, this CTA's 128-row tile
};
template <bool PAIR>
__device__ __forceinline__ UnitGeo unit_geo(const Tc05Args& T, int u, int rank) {
    UnitGeo g;
    if (PAIR) { g.z = u / T.spg; g.mt = 2 * (u - g.z * T.spg) + rank; }
    else      { g.z = u / T.mtiles; g.mt = u - g.z * T.mtiles; }
    return g;
}
// chunk bits of word w of a unit (pair: OR of both tiles, so that both CTAs walk the same stages)
template <bool PAIR>
__device__ __forceinline__ uint32_t unit_word(const Tc05Args& T, int u, int w, int gdk) {
    if (!T.bitmap) return tile_word(nullptr, w, gdk);
    if (!PAIR) return tile_word(T.bitmap + (size_t)u * T.bmw, w, gdk);
    const int z = u / T.spg, mt0 = 2 * (u - z * T.spg);
    uint32_t r = tile_word(T.bitmap + ((size_t)z * T.mtiles + mt0) * T.bmw, w, gdk);
    if (mt0 + 1 < T.mtiles) r |= tile_word(T.bitmap + ((size_t)z * T.mtiles + mt0 + 1) * T.bmw, w, gdk);
    return r;
}

// Zero-tile jumping from the precomputed bitmap: a stage (chunks [st * cps, st * cps + cps), any cps) of a unit is
// visited iff one of its chunk bits is set; without a bitmap every stage is.  Every role walks the stages of a unit with
// this iterator, so they all see the same item sequence.  With a bitmap the walk costs one find-first-set per NON-ZERO
// stage (all-zero words are skipped 32 chunks at a time), which is what keeps very sparse inputs fast.
struct StageIter {
    int w, st_next, covered_end;   // current bitmap word, next stage (no bitmap), first chunk not yet covered by an emitted stage
    uint32_t m;                    // chunk bits of word w still to visit
    __device__ __forceinline__ void begin() { w = -1; st_next = 0; covered_end = 0; m = 0u; }
    // next non-zero stage of unit u, or -1
    template <bool PAIR>
    __device__ __forceinline__ int next(const Tc05Args& T, int u, int cps, int gdk) {
        if (!T.bitmap) return st_next < T.nst ? st_next++ : -1;
        for (;;) {
            if (m == 0u) {
                if (++w >= T.bmw) return -1;
                m = unit_word<PAIR>(T, u, w, gdk);
                const int n = covered_end - 32 * w;          // chunks of this word already inside the last emitted stage
                if (n > 0) m &= n >= 32 ? 0u : ~((1u << n) - 1u);
                continue;
            }
            const int c = 32 * w + __ffs(m) - 1;
            const int st = T.cps_inv ? (int)__umulhi((unsigned)c, T.cps_inv) : c;   // c / cps (cps_inv = ceil(2^32 / cps), exact for c < 2^28; 0: cps == 1)
            covered_end = (st + 1) * cps;
            const int n = covered_end - 32 * w;
            m &= n >= 32 ? 0u : ~((1u << n) - 1u);
            return st;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Feature image: X2 (packed planes, raw words, zero-extension) -> u8 values in the MMA's K order,
// laid out as the exact shared-memory image of each 128-element K chunk (core matrices of 8 rows
// x 16 bytes; K-direction stride (NP/8)*128 B, N-direction stride 128 B).
// ---------------------------------------------------------------------------------------------
constexpr int kPrepChunks = 32;   // K chunks per prep job: 512 contiguous bytes of every (plane, column) run

// BIT2 = number of feature planes, MODE2 = expansion mode 2 (rows scaled by 2^(3 - c % 4)): both fixed at compile time so
// that every (plane, j) step is one constant shift and one AND-OR
template <int BIT2, bool MODE2>
__global__ void __launch_bounds__(256) prep_b_kernel(GemmArgs P, uint8_t* __restrict__ img, unsigned long long gstride, int NP, int halves) {
    // One block per (problem, tile of 32 K chunks, block of 8 columns).  Phase 1 stages the job's packed words in
    // shared memory, [plane][column][chunk x 4 words (+4 pad)]: consecutive threads read consecutive 16-byte pieces of
    // a column's K run (coalesced 512-byte runs).  Phase 2: thread = (chunk, k group, column), column fastest: the
    // eight columns of one (chunk, k group) store one contiguous 128-byte row of core matrices.
    extern __shared__ uint4 pk4[];   // [bit2][8][ld / 4]
    // This grid is itself launched with programmatic stream serialization: it may become resident while the previous kernel of the
    // stream still runs (which hides its launch latency), so it first waits for that kernel to complete -- x2 may be its output and
    // the image buffer may still be read by it -- and only then lets the GEMM grid start (whose adjacency pipeline runs beside
    // this grid; its feature-image loader waits for this whole grid before the first copy).
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    constexpr int bit2 = BIT2;
    const int gdk = P.gdk, n_lim = p8(P.N);
    const unsigned long long gdk4 = (unsigned long long)gdk * 4;
    const int ktiles = (gdk + kPrepChunks - 1) / kPrepChunks, nblocks = NP / 8, nh = NP / halves;
    const int ld4 = kPrepChunks + 1;          // uint4 units per (plane, column) row; the pad keeps phase 2 conflict-free
    const uint32_t* pk = reinterpret_cast<const uint32_t*>(pk4);
    const long long jobs = (long long)P.count * ktiles * nblocks;
    for (long long job = blockIdx.x; job < jobs; job += gridDim.x) {
        const int nb = (int)(job % nblocks);
        const long long t = job / nblocks;
        const int kt = (int)(t % ktiles), z = (int)(t / ktiles);
        const int c_base = kt * kPrepChunks, nch = min(kPrepChunks, gdk - c_base);
        const Words x2{P.x2 + (unsigned long long)z * P.x2_bstride, P.x2_words};
        for (int i = threadIdx.x; i < bit2 * 8 * nch; i += blockDim.x) {
            const int cc = i % nch, r = i / nch, col = r & 7, q = r >> 3, n = nb * 8 + col;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (n < n_lim) w = ldw4(x2, (unsigned long long)q * P.w_offset + (unsigned long long)n * gdk4 + 4ull * (c_base + cc));
            pk4[(size_t)(q * 8 + col) * ld4 + cc] = w;
        }
        __syncthreads();
        uint8_t* base = img + (unsigned long long)z * gstride + (unsigned long long)c_base * NP * 128;
        const int h = (nb * 8) / nh, nbl = nb - h * (nh / 8);        // pair mode keeps the two column halves of a chunk apart
        for (int i = threadIdx.x; i < nch * 64; i += blockDim.x) {
            const int col = i & 7, kg = (i >> 3) & 7, cc = i >> 6, s = kg >> 1, j0 = (kg & 1) * 4;
            uint32_t o[4] = {0u, 0u, 0u, 0u};
            // byte i of o[j] = value of K element (32 s + 4 (j0 + j) + i) = sum over planes of bit (j0 + j + 8 i) of the
            // plane's word s, moved to bit q (+ 3 - j in expansion mode 2: the expanders leave adjacency bit c at
            // 2^(c % 4), the rows carry the complementary factor).  After w >>= j0 the source bit is j: one constant
            // shift and one AND-OR per (plane, j).
#pragma unroll
            for (int q = 0; q < BIT2; ++q) {
                const uint32_t w = pk[((size_t)(q * 8 + col) * ld4 + cc) * 4 + s] >> j0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int to = q + (MODE2 ? 3 - j : 0);   // compile-time after unrolling
                    const uint32_t sh = to >= j ? (w << (to - j)) : (w >> (j - to));
                    o[j] |= sh & (0x01010101u << to);
                }
            }
            // image of one K chunk: [half][k group (8)][column block][8 columns][16 B]; halves == 2 in pair mode
            *reinterpret_cast<uint4*>(base + (size_t)cc * NP * 128 + (size_t)h * nh * 128 + (size_t)kg * (nh / 8) * 128 + (size_t)nbl * 128 + col * 16) =
                make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
    }
}

// 4-bit operand path: packed feature planes (1 or 2 bits) -> E2M1 nibbles in the MMA's K order.  One K chunk of 128
// elements is four groups (one per 32-bit word wd of the packed run) of [column block][8 columns][16 B]; the 16 bytes of a
// column hold, for j = 0..3, the eight nibbles i = 0..7 of the feature at bit j + 4 i of the word, scaled to pair with
// the adjacency nibble of the expanders (bit j of the nibble = 0.5, 1, 2, 2): feature f in {0..3} is stored as 2 f, f,
// f / 2, f / 2 -- all exact in E2M1 (codes 0,4,6,7 / 0,2,4,5 / 0,1,2,3) -- so that every product is f itself.
template <int BIT2>
__global__ void __launch_bounds__(256) prep_b4_kernel(GemmArgs P, uint8_t* __restrict__ img, unsigned long long gstride, int NP) {
    extern __shared__ uint4 pk4[];   // [bit2][8][ld4]
    asm volatile("griddepcontrol.wait;\n" ::: "memory");              // see prep_b_kernel
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    const int gdk = P.gdk, n_lim = p8(P.N);
    const unsigned long long gdk4 = (unsigned long long)gdk * 4;
    const int ktiles = (gdk + kPrepChunks - 1) / kPrepChunks, nblocks = NP / 8;
    const int ld4 = kPrepChunks + 1;
    const uint32_t* pk = reinterpret_cast<const uint32_t*>(pk4);
    const long long jobs = (long long)P.count * ktiles * nblocks;
    for (long long job = blockIdx.x; job < jobs; job += gridDim.x) {
        const int nb = (int)(job % nblocks);
        const long long t = job / nblocks;
        const int kt = (int)(t % ktiles), z = (int)(t / ktiles);
        const int c_base = kt * kPrepChunks, nch = min(kPrepChunks, gdk - c_base);
        const Words x2{P.x2 + (unsigned long long)z * P.x2_bstride, P.x2_words};
        for (int i = threadIdx.x; i < BIT2 * 8 * nch; i += blockDim.x) {
            const int cc = i % nch, r = i / nch, col = r & 7, q = r >> 3, n = nb * 8 + col;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (n < n_lim) w = ldw4(x2, (unsigned long long)q * P.w_offset + (unsigned long long)n * gdk4 + 4ull * (c_base + cc));
            pk4[(size_t)(q * 8 + col) * ld4 + cc] = w;
        }
        __syncthreads();
        uint8_t* base = img + (unsigned long long)z * gstride + (unsigned long long)c_base * NP * 64;
        for (int i = threadIdx.x; i < nch * 32; i += blockDim.x) {
            const int col = i & 7, wd = (i >> 3) & 3, cc = i >> 5;
            const uint32_t w0 = pk[((size_t)(0 * 8 + col) * ld4 + cc) * 4 + wd];
            const uint32_t w1 = BIT2 > 1 ? pk[((size_t)(1 * 8 + col) * ld4 + cc) * 4 + wd] : 0u;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t b0 = (w0 >> j) & 0x11111111u, b1 = (w1 >> j) & 0x11111111u;
                if (j == 0) o[j] = ((b0 | b1) << 2) | (b1 << 1) | (b0 & b1);          // 2 f : codes 0, 4, 6, 7
                else if (j == 1) o[j] = (b1 << 2) | ((b0 & ~b1) << 1) | (b0 & b1);    //   f : codes 0, 2, 4, 5
                else o[j] = b0 | (b1 << 1);                                           // f / 2 : codes 0, 1, 2, 3
            }
            *reinterpret_cast<uint4*>(base + (size_t)cc * NP * 64 + (size_t)wd * (NP / 8) * 128 + (size_t)nb * 128 + col * 16) =
                make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// The GEMM kernel
// ---------------------------------------------------------------------------------------------
// DBG = 2 adds the clock64 trace and the timing-experiment knobs (flag bits 0x10000..0x400000), DBG = 1 only the
// trace (the production code path with seven clock reads per item, for timelines that can be trusted); the production
// instance carries none of that code in its (latency-bound) role loops.
// CPST > 0 fixes the chunks per item at compile time (the hot configurations: 7 on the 4-bit path, 4 on the u8 path), so the
// expander loop carries no per-chunk predicates or address arithmetic; 0 reads T.cps.
template <int EPI, bool PAIR, int DBG, int AMODE, int CPST = 0>   // AMODE >= 0 fixes the expansion mode at compile time, -1 reads T.amode
__global__ void __launch_bounds__(kMaxThreads, 1) gemm_tc05_kernel(const Tc05Args T, const __grid_constant__ CUtensorMap tmapA) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    // the swizzled TMA destinations need 1024-byte alignment: align by hand, the launch adds the slack
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const GemmArgs& P = T.g;
    // the next kernel of the stream (the next call's prep grid, launched programmatically) may become resident as SMs free up; it
    // waits for this whole grid before it touches memory
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NP = T.NP, S = T.stages, gdk = P.gdk, CPS = CPST > 0 ? CPST : T.cps;
    constexpr bool F4 = AMODE == 3;     // 4-bit operand path (only as a compile-time instance)
    // Only the owner of an item commits to its `empty` barrier (not in pair mode).  Every issuer still WAITS for every item's
    // `full` phase (parity waits must not skip a phase), but a non-owner's tcgen05.commit would queue behind the owner's MMAs
    // in the tensor-core instruction queue and stall that warp until they are dispatched (clock64 trace: its own next item was
    // picked up 300-450 cycles late, the tensor pipe idling meanwhile).  Every other waiter of `empty[s]` sees every phase of
    // it (an expander group stays on one stage because G == S; the feature loader walks all items), so completion order
    // between the issuers does not matter.
    constexpr bool kOwnerCommit = !PAIR && !QGTC_TC05_WIDE && QGTC_TC05_OWNER_COMMIT;   // (the wide layout has G < S)
    // Per-owner `full` barriers (not in pair mode): item i is signalled on full[s] when issuer 0 owns it and on full_b[s] when
    // issuer 1 does, so an issuer meets consecutive phases of every barrier it waits on and never has to observe the other
    // issuer's items (with one barrier per stage it had to: a parity wait must not skip a phase, and every such wait costs
    // ~160 cycles even on a completed barrier -- clock64 trace, profiles/r2_tc05_trace_f4_n64_call25.txt).
    constexpr bool kOwnFull = kOwnerCommit && QGTC_TC05_OWN_FULL;   // (a non-owner that skips an item cannot commit to its `empty` barrier either)
    // Split release (compile-time item size, no pairing): the issuer commits to `empty` after the MMAs of the first kHalf chunks
    // of an item and to `empty2` after the rest, so the expanders refill the first chunks of a stage while the tensor pipe
    // still works on its last ones -- the ring (three stages: fill + pick-up + drain + release = three drains) gets slack.
    constexpr bool SPLIT = CPST > 1 && !PAIR && !F4 && QGTC_TC05_SPLIT;   // (the 4-bit path is paced by the MMA dispatch rate, not by the ring)
    constexpr int kHalf = SPLIT ? (CPST + 1) / 2 : 0;
    constexpr int kCh = F4 ? kMaxCps4 : kMaxCps;          // chunks per item, compile-time bound
    constexpr int kChunkCols = F4 ? 16 : 32;              // TMEM columns of one expanded chunk
    const SmemLayout L = smem_layout(NP, S, CPS, PAIR ? 1 : 0, T.rslots, F4 ? 1 : 0);
    const int rank = PAIR ? (int)cluster_ctarank() : 0;            // 0 = leader (issues the MMAs)
    const int group_id = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int ngroups = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    unsigned char* raw_slots = smem_raw + L.raw0;
    unsigned char* b_stages = smem_raw + L.b_stage0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
    uint64_t* full = bars;                          // [kMaxStages] 4 expander warps + the bulk copy's expect_tx -> MMA
    uint64_t* empty = bars + kMaxStages;            // [kMaxStages] MMA commit -> expanders, feature loader
    uint64_t* raw_full = bars + 2 * kMaxStages;     // [kRawSlots]  adjacency TMA -> expanders
    uint64_t* raw_empty = raw_full + kRawSlots;     // [kRawSlots]  4 expander warps -> adjacency loader
    uint64_t* d_full = raw_empty + kRawSlots;       // [2] MMA -> epilogue   (count 2: one commit per issuer)
    uint64_t* d_empty = d_full + 2;                 // [2] epilogue -> MMA   (count 8 warps)
    uint64_t* peer_full = d_empty + 2;              // [kMaxStages] pair mode, leader only: the peer CTA's stage is ready (relayed)
    uint64_t* peer_dempty = peer_full + kMaxStages; // [2]          pair mode, leader only: the peer's epilogue drained the accumulator
    uint64_t* empty2 = peer_dempty + 2;             // [kMaxStages] split release (SPLIT): the MMAs of the second half of the stage's chunks are done
    uint64_t* full_b = empty2 + kMaxStages;         // [kMaxStages] per-owner `full` barriers (kOwnFull): the items of issuer 1 (issuer 0's use `full`)
    constexpr uint32_t kFullBOff = (uint32_t)(4 * kMaxStages + 2 * kRawSlots + 6) * 8u;   // byte offset of full_b from full
    volatile unsigned char* flags = smem_raw + L.flags;  // [stage][chunk][lane quarter]: that expander warp's 32 x 128 bits are non-zero
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + L.misc);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(&full[s], T.wpi + 1);         // the expander warps of one item + the feature copy's expect_tx
            mbar_init(&full_b[s], T.wpi + 1);
            mbar_init(&empty[s], kOwnerCommit ? 1 : T.ni);   // the item's owner commits (pair mode: every issuer, see the issuer loop)
            mbar_init(&empty2[s], kOwnerCommit ? 1 : T.ni);
        }
        for (int s = 0; s < kRawSlots; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], T.wpi);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], T.ni);     // every issuer commits its share of the tile
            mbar_init(&d_empty[b], kEpilogueWarps);
            mbar_init(&peer_dempty[b], 1);
        }
        for (int s = 0; s < kMaxStages; ++s) mbar_init(&peer_full[s], 1);
        fence_mbar_init();
    }
    if (warp == kALoaderWarp && lane == 0) prefetch_tmap(&tmapA);
    if (warp == kMmaWarp) {
        if (PAIR) tmem_alloc_pair(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    // number of pipeline items of this CTA = non-zero stages of its tiles (all stages without a bitmap)
    int* item_count = reinterpret_cast<int*>(smem_raw + L.misc + 16);
    if (threadIdx.x == 0) *item_count = 0;
    __syncthreads();
    {
        const int units_here = (T.units - group_id + ngroups - 1) / ngroups;
        int cnt = 0;
        if (!T.bitmap) {
            if (threadIdx.x == 0) cnt = units_here * T.nst;
        } else {
            for (int ul = threadIdx.x; ul < units_here; ul += blockDim.x) {
                StageIter it;
                it.begin();
                while (it.next<PAIR>(T, group_id + ul * ngroups, T.cps, P.gdk) >= 0) ++cnt;
            }
        }
        if (cnt) atomicAdd(item_count, cnt);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();      // the peer's barriers must be initialised before anything is signalled across
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int sf_col = T.dbufs * NP;      // 4-bit path: kSfCols columns of unit block scales after the accumulators
    const int dcols = T.dbufs * NP + (F4 ? kSfCols : 0);   // accumulators (and scales) first, then the A stages
    const int stage_cols = CPS * kChunkCols;

    const int my_tiles = (T.units - group_id + ngroups - 1) / ngroups;   // work units of this CTA (pair)
    const int n_items = *item_count;      // launch_gemm_tc05 guarantees this fits 31 bits
    const bool nojump = (P.flags & QGTC_FLAG_NO_ZERO_JUMP) != 0;
    // timing-experiment knobs (results are wrong when set; used only by profiles/scripts/tc05_knobs.py)
    const bool dbg_no_st = DBG == 2 && (P.flags & 0x10000u) != 0, dbg_no_mma = DBG == 2 && (P.flags & 0x20000u) != 0;
    const bool dbg_no_bulk = DBG == 2 && (P.flags & 0x40000u) != 0, dbg_no_ldg = DBG == 2 && (P.flags & 0x80000u) != 0;
    const bool dbg_alt_d = DBG == 2 && (P.flags & 0x400000u) != 0;   // alternate the MMAs of a tile between both accumulators (is the D dependency the pace?)
    // Adjacency expansion modes (T.amode, chosen by the launcher so that no scaled sum can reach 2^31):
    //  2: bit c of every byte stays where it is within its nibble -- one AND per register (plus one shift per
    //     word for the upper nibble) -- so the A operand holds 2^(c % 4); prep_b_kernel scales the feature rows by
    //     2^(3 - c % 4) (needs bit2 <= 5) and the accumulators hold 8 x the sums;
    //  1: bit c is moved to the byte's MSB by an integer multiply (FMA pipe) and one AND: A holds 0x80, 128 x sums;
    //  0: bit c is moved to bit 0 by a shift and an AND (both ALU pipe): exact sums.
    // The epilogue shifts the accumulators back (exact: every term carries the factor).
    const int amode = AMODE >= 0 ? AMODE : T.amode;
    const bool hinted = DBG != 2 || (P.flags & 0x200000u) == 0;      // knob set: mbarrier waits poll with nanosleep back-off instead of a suspend hint

    if (warp < kProducerWarps) {
        // ===================== adjacency expanders =====================
        // warp = (TMEM lane quarter q, slot t).  An item (stage) is expanded by wpi warps, each taking
        // wch chunks; the 16 warps form G = 16 / wpi groups that take the items round-robin, so a warp
        // has G item periods to get through its (latency-bound) wait / expand / store / signal chain.
        const int q = warp & 3, t = warp >> 2;
        const int spi = T.wpi >> 2;                // expander warps per lane quarter and item
        const int g = t / spi, c0 = (t - g * spi) * T.wch;   // group, first chunk of this warp's share of an item
        // Each warp expands all chunks of its item for its 32 rows (wpi = 4 warps per item: the fixed cost of an
        // item -- two waits, two arrivals, addresses -- is paid once per 32 rows x stage), so the groups of four
        // warps take the items round-robin and a warp has G item periods for its wait / expand / store / signal chain.
        // G <= S is required: a warp that reaches item X has seen the MMAs of item X-G-S complete (it waited for
        // them at item X-G), which only implies that item X-2S is done -- no mbarrier phase aliasing -- if G <= S
        // (MMA completion is in item order, see the issuers).  With three stages the fourth group idles.
        const int WCH = (CPST > 0 && !QGTC_TC05_WIDE) ? CPST : T.wch;
        const int G = min(min(kProducerWarps / T.wpi, S), T.rslots);
        const int row = 32 * q + lane;             // row of the tile == TMEM lane
        const int my_items = g < G ? (n_items - g + G - 1) / G : 0;   // warps beyond the last group idle
        // byte offset of chunk j of this row inside a raw slot = row_off + ((j ^ swz) << 4): the TMA swizzle
        // (Swizzle<log2(cps),4,3>) XORs address bits [4,4+log2 cps) with bits [7,7+log2 cps)
        const uint32_t row_off = (uint32_t)(row * 16 * CPS);
        const uint32_t swz = (row_off >> 7) & (CPS == 4 ? 3u : (CPS == 2 ? 1u : 0u));   // 3 chunks per stage: no swizzle (48-byte rows are conflict-free)
        // multipliers 2^(7-c) of the scaled expansion; derived from a kernel argument so that ptxas keeps them
        // as IMAD operands (FMA pipe) instead of folding them back into ALU-pipe shifts
        uint32_t mult[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mult[c] = T.one << (7 - c);

        const int rslots = T.rslots;
        int slot = g % rslots, s = g % S;          // raw ring slot and TMEM/B stage of the current item
        int own = kOwnFull ? g % T.ni : 0;         // issuer that owns the current item (item % ni)
        const int own_step = kOwnFull ? G % T.ni : 0;
        uint32_t raw_use = (uint32_t)(g / rslots), use = (uint32_t)(g / S);
        const uint32_t st_addr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(dcols + kChunkCols * c0);
        // traced expander warps: 0 (role 0), 1 (role 3), 3 (role 4), 8 (role 6)
        const int trole = warp == 0 ? 0 : (warp == 1 ? 3 : (warp == 3 ? 4 : (warp == 8 ? 6 : -1)));
        const bool tr = DBG && (trole >= 0 && lane == 0);
        // shared-window addresses, formed once: this lane's row in slot 0 (swizzled: bits [4, 4 + log2 cps) of the row offset
        // are zero, so chunk j sits at (row address | swz << 4) ^ (j << 4)), this warp's flag bytes, the barrier arrays
        const bool swizzled = CPS == 4 || CPS == 2;
        const uint32_t raw_row0 = smem_u32(raw_slots) + row_off + (swizzled ? (swz << 4) : 0u);
        const uint32_t slot_bytes = (uint32_t)L.raw_slot_bytes;
        // (one base register: the barrier arrays and the flag bytes sit at constant offsets from the first barrier)
        const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * kMaxStages, raw_full0 = full0 + 16u * kMaxStages, raw_empty0 = raw_full0 + 8u * kRawSlots;
        const uint32_t flag_addr0 = full0 + (uint32_t)kFlagsOff + (uint32_t)(c0 * 4 + q + lane * 4);
        const uint32_t empty2_0 = smem_u32(empty2);
        auto lds128 = [](uint32_t addr) -> uint4 {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
            return v;
        };
        // expansion of one 128-bit chunk of this lane's row into the 32 TMEM columns' registers
        auto expand = [&](const uint4& pv, uint32_t (&r)[32]) {
            const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
            if (amode == 2) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t hi = w[ks] >> 4;
#pragma unroll
                    for (int cidx = 0; cidx < 4; ++cidx) {
                        r[ks * 8 + cidx] = w[ks] & (0x01010101u << cidx);
                        r[ks * 8 + 4 + cidx] = hi & (0x01010101u << cidx);
                    }
                }
            } else if (amode == 1) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                    for (int cidx = 0; cidx < 7; ++cidx) r[ks * 8 + cidx] = (w[ks] * mult[cidx]) & 0x80808080u;
                    r[ks * 8 + 7] = w[ks] & 0x80808080u;
                }
            } else {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    r[ks * 8] = w[ks] & 0x01010101u;
#pragma unroll
                    for (int cidx = 1; cidx < 8; ++cidx) r[ks * 8 + cidx] = (w[ks] >> cidx) & 0x01010101u;
                }
            }
        };
        // 4-bit path: one 128-bit chunk -> 16 registers of eight E2M1 nibbles.  Register 4 wd + j, nibble i = bit j + 4 i of
        // word wd, left at nibble bit j: 0.5, 1.0 or 2.0 (bit 3 would be the sign: it is moved to bit 2); prep_b4_kernel
        // scales the feature rows by 2, 1, 0.5, 0.5 so that every product is the feature value itself.
        auto expand4 = [&](const uint4& pv, uint32_t (&r)[16]) {
            const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
            for (int wd = 0; wd < 4; ++wd) {
                r[4 * wd + 0] = w[wd] & 0x11111111u;
                r[4 * wd + 1] = w[wd] & 0x22222222u;
                r[4 * wd + 2] = w[wd] & 0x44444444u;
                r[4 * wd + 3] = (w[wd] >> 1) & 0x44444444u;   // (as IMAD.HI on the FMA pipe instead: measured 3 % slower)
            }
        };
        // zero-tile jumping on the fly: every lane notes which of its chunks are non-zero (two ORs, a predicate and a
        // predicated OR per chunk); ONE warp reduction per item turns that into the per-chunk flags of this lane quarter
        const bool check_zero = !(nojump || PAIR);
        for (int k = 0; k < my_items; ++k) {
            if (tr) QGTC_TRACE(trole, k, 0);
            if (!dbg_no_ldg) mbar_wait_a(raw_full0 + 8u * (uint32_t)slot, raw_use & 1, T.error_flag, 7, hinted);
            uint4 v[kCh];
            const uint32_t src = raw_row0 + (uint32_t)slot * slot_bytes;
            auto load_chunk = [&](int j) -> uint4 {
                return j < WCH ? lds128(swizzled ? (src ^ ((uint32_t)(c0 + j) << 4)) : (src + ((uint32_t)(c0 + j) << 4))) : make_uint4(0u, 0u, 0u, 0u);
            };
            // 4-bit path: all chunks up front (16 result registers per chunk leave room); u8 path: one chunk ahead of the
            // expansion (32 result registers per chunk: loading everything first spills the loop state)
            constexpr int kAhead = F4 ? kCh : 2;
#pragma unroll
            for (int j = 0; j < kAhead; ++j) v[j] = load_chunk(j);
            if (tr) QGTC_TRACE(trole, k, 1);
            // expand the first chunk BEFORE waiting for the stage: the registers hold it while the MMAs of the
            // stage's previous use drain, so only the tcgen05.st sits behind the wait.  An all-zero 32 x 128 block is
            // expanded like any other (to zeros); it only clears its flag so that the MMA can be skipped when the
            // other three lane quarters are zero too.
            uint32_t lane_nz = 0;
            const uint32_t st_s = st_addr + (uint32_t)(s * stage_cols);
            if (F4) {
                {
                    uint32_t r[16];
                    if ((v[0].x | v[0].y | v[0].z | v[0].w) != 0u) lane_nz |= 1u;
                    expand4(v[0], r);
                    if (tr) QGTC_TRACE(trole, k, 2);
                    mbar_wait_a<32>(empty0 + 8u * (uint32_t)s, (use & 1) ^ 1, T.error_flag, 1, hinted);   // the MMAs of the previous use of this stage are done
                    tc_fence_after();
                    if (tr) QGTC_TRACE(trole, k, 3);
                    tmem_st16r(st_s, r);
                }
#pragma unroll
                for (int j = 1; j < kCh; ++j) {
                    if (SPLIT && j == kHalf) {   // the second half of the stage's previous use has been multiplied too
                        mbar_wait_a<32>(empty2_0 + 8u * (uint32_t)s, (use & 1) ^ 1, T.error_flag, 14, hinted);
                        tc_fence_after();
                    }
                    if (j < WCH) {   // warp-uniform
                        uint32_t r[16];
                        if ((v[j].x | v[j].y | v[j].z | v[j].w) != 0u) lane_nz |= 1u << j;
                        expand4(v[j], r);
                        tmem_st16r(st_s + (uint32_t)(16 * j), r);
                    }
                }
            } else {
                {
                    uint32_t r[32];
                    if ((v[0].x | v[0].y | v[0].z | v[0].w) != 0u) lane_nz |= 1u;
                    expand(v[0], r);
                    if (tr) QGTC_TRACE(trole, k, 2);
                    mbar_wait_a<32>(empty0 + 8u * (uint32_t)s, (use & 1) ^ 1, T.error_flag, 1, hinted);   // the MMAs of the previous use of this stage are done
                    tc_fence_after();
                    if (tr) QGTC_TRACE(trole, k, 3);
                    if (!dbg_no_st) tmem_st32(st_s, r);
                    else if (r[5] == 0xdeadbeefu) flags[255] = 1;  // keep the expansion alive
                }
#pragma unroll
                for (int j = 1; j < kCh; ++j) {
                    if (j + 1 < kCh) v[j + 1] = load_chunk(j + 1);
                    if (SPLIT && j == kHalf) {   // the second half of the stage's previous use has been multiplied too
                        mbar_wait_a<32>(empty2_0 + 8u * (uint32_t)s, (use & 1) ^ 1, T.error_flag, 14, hinted);
                        tc_fence_after();
                    }
                    if (j < WCH) {   // warp-uniform
                        uint32_t r[32];
                        if ((v[j].x | v[j].y | v[j].z | v[j].w) != 0u) lane_nz |= 1u << j;
                        expand(v[j], r);
                        if (!dbg_no_st) tmem_st32(st_s + (uint32_t)(32 * j), r);
                        else if (r[5] == 0xdeadbeefu) flags[255] = 1;
                    }
                }
            }
            // every lane holds its raw words by now: the slot can be refilled
            const uint32_t nzmask = check_zero ? __reduce_or_sync(0xffffffffu, lane_nz) : 0xffffffffu;   // also converges the warp
            if (!check_zero) __syncwarp();
            if (lane == 0) mbar_arrive_a(raw_empty0 + 8u * (uint32_t)slot);
            if (lane < WCH) {
                const uint32_t bit = (nzmask >> lane) & 1u;
                asm volatile("st.shared.u8 [%0], %1;\n" ::"r"(flag_addr0 + 32u * (uint32_t)s), "r"(bit) : "memory");
            }
            if (tr) QGTC_TRACE(trole, k, 4);
            if (!dbg_no_st) tmem_wait_st();
            tc_fence_before();
            if (tr) QGTC_TRACE(trole, k, 5);
            __syncwarp();   // every lane's tcgen05.st has completed, the flags are written
            if (lane == 0) mbar_arrive_a(full0 + (own ? kFullBOff : 0u) + 8u * (uint32_t)s);
            if (tr) QGTC_TRACE(trole, k, 6);
            slot += G;
            if (slot >= rslots) { slot -= rslots; ++raw_use; }
            s += G;
            if (s >= S) { s -= S; ++use; }
            own += own_step;
            if (own >= T.ni) own -= T.ni;
        }
    } else if (warp == kALoaderWarp) {
        // ===================== adjacency TMA loader =====================
        if (lane == 0 && !dbg_no_ldg) {
            const uint32_t bytes = (uint32_t)(128 * 16 * CPS);
            int slot = 0;
            uint32_t use = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int u = group_id + tl * ngroups;
                const UnitGeo ug = unit_geo<PAIR>(T, u, rank);
                const int z = T.a_batched ? ug.z : 0;
                const int row0 = ug.mt * 128;                // may lie beyond the matrix in pair mode: TMA fills zeros
                StageIter it;
                it.begin();
                for (int st; (st = it.next<PAIR>(T, u, CPS, gdk)) >= 0;) {
                    const int c0 = st * CPS;                    // first chunk of a non-zero stage
                    mbar_wait<128>(&raw_empty[slot], (use & 1) ^ 1, T.error_flag, 8, hinted);
                    mbar_expect_tx(&raw_full[slot], bytes);
                    tma_load_3d(raw_slots + (size_t)slot * L.raw_slot_bytes, &tmapA, c0 * 4, row0, z, &raw_full[slot]);
                    if (++slot == T.rslots) { slot = 0; ++use; }
                }
            }
        }
    } else if (warp == kLoaderWarp) {
        // ===================== feature-image loader =====================
        if (lane == 0) {
            // programmatic dependent launch: everything up to here (and the adjacency pipeline) overlaps the tail
            // of prep_b_kernel; the image itself is complete and visible only after this wait
            asm volatile("griddepcontrol.wait;\n" ::: "memory");
            const uint32_t chunk_bytes = (uint32_t)NP * (F4 ? 64u : 128u);
            int s = 0, own = 0;       // own: issuer that owns the item (item % ni), selects the `full` barrier
            uint32_t use = 0;
            for (int tl = 0; tl < my_tiles; ++tl) {
                const int u = group_id + tl * ngroups;
                const int z = unit_geo<PAIR>(T, u, rank).z;
                const uint8_t* img = T.bimg + (unsigned long long)z * T.bimg_gstride;
                StageIter it;
                it.begin();
                for (int st; (st = it.next<PAIR>(T, u, CPS, gdk)) >= 0;) {
                    const int c0 = st * CPS;
                    const int nch = (gdk - c0 < CPS) ? gdk - c0 : CPS;
                    mbar_wait<128>(SPLIT ? &empty2[s] : &empty[s], (use & 1) ^ 1, T.error_flag, 2, hinted);   // every MMA that reads the stage's feature chunks is done
                    uint64_t* fbar = (kOwnFull && own) ? &full_b[s] : &full[s];
                    if (kOwnFull && ++own == T.ni) own = 0;
                    if (dbg_no_bulk) {
                        mbar_arrive(fbar);
                    } else if (PAIR) {
                        // this CTA's half of every chunk (the image keeps the two halves of a chunk apart)
                        const uint32_t half_bytes = chunk_bytes >> 1;
                        mbar_expect_tx(&full[s], half_bytes * (uint32_t)nch);
                        for (int j = 0; j < nch; ++j)
                            bulk_g2s(b_stages + (size_t)s * L.b_stage_bytes + (size_t)j * half_bytes,
                                     img + (unsigned long long)(c0 + j) * chunk_bytes + (unsigned long long)rank * half_bytes, half_bytes, &full[s]);
                    } else {
                        mbar_expect_tx(fbar, chunk_bytes * (uint32_t)nch);
                        bulk_g2s(b_stages + (size_t)s * L.b_stage_bytes, img + (unsigned long long)c0 * chunk_bytes, chunk_bytes * (uint32_t)nch, fbar);
                    }
                    if (++s == S) { s = 0; ++use; }
                }
            }
        }
    } else if (warp >= kMmaWarp) {
        // ===================== MMA issuers =====================
        // T.ni warps issue: a tcgen05.mma of 128 x 64 x 32 keeps the tensor pipe busy ~32 cycles, but one
        // warp needs ~1300 cycles of (latency-bound) instruction stream per item: wait, flags, 16 MMAs with their
        // descriptors, commit.  The issuers take the items round-robin.  Every MMA accumulates (the epilogue
        // leaves the accumulator zeroed), so the order between the issuers does not matter.
        // Protocol, pair mode (and builds without kOwnFull / kOwnerCommit): EVERY issuer waits for EVERY item's `full` phase and
        // commits to its `empty` barrier (count T.ni); only the owner issues MMAs in between.  A non-owner's commit tracks the
        // MMAs of its own earlier items, so "item i is done" implies "all earlier items are done", and no issuer ever skips a
        // phase of a barrier it waits on (mbarrier waits only see parity).  Default (kOwnFull): per-owner `full` barriers and
        // owner-only commits, see the constants at the top of the kernel -- an issuer only ever touches its own items.
        // Each warp runs with warp-uniform state (addresses and descriptors live in uniform registers);
        // one elected lane issues the tcgen05 instructions.
        if (PAIR && rank != 0) {
            // ---- peer CTA of a pair: it issues no MMAs.  Its two otherwise idle issuer warps relay "this CTA's
            // stage is ready" (warp 25) and "this CTA's epilogue has drained the accumulator" (warp 27) to the leader.
            if (lane == 0) {
                if (warp == kMmaWarp) {
                    int s = 0;
                    uint32_t use = 0;
                    for (int it = 0; it < n_items; ++it) {
                        mbar_wait<64>(&full[s], use & 1, T.error_flag, 9, hinted);
                        tc_fence_after();
                        tc_fence_before();
                        mbar_arrive_leader(&peer_full[s]);
                        if (++s == S) { s = 0; ++use; }
                    }
                } else if (warp == kMmaWarp2) {
                    int buf = 0;
                    uint32_t duse = 0;
                    for (int tl = 0; tl < my_tiles; ++tl) {
                        mbar_wait<256>(&d_empty[buf], duse & 1, T.error_flag, 10, hinted);
                        tc_fence_after();
                        tc_fence_before();
                        mbar_arrive_leader(&peer_dempty[buf]);
                        if (++buf == T.dbufs) { buf = 0; ++duse; }
                    }
                }
            }
        } else {
            const int who = warp - kMmaWarp;
            const int trole = who == 0 ? 1 : (who == 1 ? 5 : -1);   // traced issuers
            // The CTA owns all 512 TMEM columns, so the allocation starts at column 0 / lane 0; using the
            // literal keeps every MMA operand provably warp-uniform (uniform registers, no per-MMA R2UR).
            if (tmem_base != 0u) {
                if (T.error_flag) { atomicExch_system(T.error_flag, 0x40000000u | tmem_base); __threadfence_system(); }
                __trap();
            }
            const uint32_t tb = 0u;
            const uint32_t idesc = F4 ? make_idesc_f4(NP) : make_idesc_i8(NP, PAIR ? 256 : 128);
            const uint32_t sfa = tb + (uint32_t)sf_col, sfb = tb + (uint32_t)sf_col + 4u;   // 4-bit path: unit scales (every byte 0x7F)
            const uint32_t lbo = (uint32_t)(NP / (PAIR ? 16 : 8)) * 128u, sbo = 128u;   // pair: this CTA's smem holds NP/2 columns
            // descriptor = constant high word | (low word that only differs in the 14-bit start address)
            const uint64_t desc0 = make_smem_desc(smem_u32(b_stages), lbo, sbo);
            const uint32_t desc_hi = (uint32_t)(desc0 >> 32);
            const uint32_t lo0 = (uint32_t)desc0;
            const uint32_t stage_lo = (uint32_t)L.b_stage_bytes >> 4;
            const uint32_t chunk_lo = ((uint32_t)NP * ((PAIR || F4) ? 64u : 128u)) >> 4;
            const uint32_t kstep_lo = (2u * lbo) >> 4;   // one MMA's K: two core matrices (u8: 2 x 16 elements, 4-bit: 2 x 32)
            int s = 0, buf = 0, item = 0, owner = 0;   // owner = item % T.ni
            uint32_t use = 0, duse = 0;
            uint32_t own_par = 0;   // kOwnFull: bit s = parity of this issuer's next wait on its `full` barrier of stage s
            auto acquire = [&](int s_, uint32_t use_) -> uint32_t {
                if (kOwnFull) {
                    mbar_wait(who ? &full_b[s_] : &full[s_], (own_par >> s_) & 1u, T.error_flag, 4, hinted);
                    own_par ^= 1u << s_;
                } else {
                    mbar_wait(&full[s_], use_ & 1, T.error_flag, 4, hinted);
                }
                if (lane == 0 && trole >= 0) QGTC_TRACE(trole, item, 4);
                if (PAIR) {   // ... and the peer CTA's half of the stage; its zero flags are not visible here: issue everything
                    mbar_wait(&peer_full[s_], use_ & 1, T.error_flag, 11, hinted);
                    tc_fence_after();
                    return dbg_no_mma ? 0u : 0xFu;
                }
                tc_fence_after();
                // one word per chunk: the four lane quarters' "non-zero" bytes
                const uint32_t w = lane < kCh ? *reinterpret_cast<volatile uint32_t*>(smem_raw + L.flags + s_ * 32 + lane * 4) : 0u;
                const uint32_t m4 = __ballot_sync(0xffffffffu, w != 0u) & ((1u << kCh) - 1u);
                return dbg_no_mma ? 0u : m4;
            };
            for (int tl = 0; tl < my_tiles; ++tl) {
                mbar_wait(&d_empty[buf], duse & 1, T.error_flag, 3, hinted);   // the epilogue has drained and zeroed this accumulator
                if (PAIR) mbar_wait(&peer_dempty[buf], duse & 1, T.error_flag, 12, hinted);
                tc_fence_after();
                const uint32_t d_tmem = tb + (uint32_t)(buf * NP);
                const int u = group_id + tl * ngroups;
                StageIter it;
                it.begin();
                for (int st; (st = it.next<PAIR>(T, u, CPS, gdk)) >= 0;) {
                  {
                    if (owner == who) {
                        const int c_first = st * CPS;                      // first chunk of this non-zero stage
                        if (lane == 0 && trole >= 0) QGTC_TRACE(trole, item, 0);
                        const uint32_t m_cur = acquire(s, use);
                        if (lane == 0 && trole >= 0) QGTC_TRACE(trole, item, 1);
                        const int nch = (gdk - c_first < CPS) ? gdk - c_first : CPS;
                        const uint32_t a_tmem = tb + (uint32_t)(dcols + s * stage_cols);
                        const uint32_t lo_s = lo0 + (uint32_t)s * stage_lo;
                        if (elect_one()) {
                            if (CPST > 0 && QGTC_TC05_DENSE_ISSUE && nch == CPST && m_cur == ((1u << CPST) - 1u)) {
                                // dense item of the compile-time size (the common case): one branch-free MMA sequence -- the issuer
                                // warp gets an issue slot every 5-10 cycles next to the expanders, so its instruction count per
                                // item, not the tensor pipe, sets the pick-up latency
#pragma unroll
                                for (int j = 0; j < (CPST > 0 ? CPST : 1); ++j) {
                                    if (SPLIT && j == kHalf) tc_commit(&empty[s]);   // the first chunks of the stage may be refilled
                                    if (F4) {
#pragma unroll
                                        for (int mm = 0; mm < 2; ++mm) {
                                            const uint64_t desc = ((uint64_t)desc_hi << 32) | (uint64_t)(lo_s + (uint32_t)j * chunk_lo + (uint32_t)mm * kstep_lo);
                                            mma_f4_ts(d_tmem, a_tmem + (uint32_t)(16 * j + 8 * mm), desc, idesc, sfa, sfb, 1u);
                                        }
                                    } else {
#pragma unroll
                                        for (int ks = 0; ks < 4; ++ks) {
                                            const uint64_t desc = ((uint64_t)desc_hi << 32) | (uint64_t)(lo_s + (uint32_t)j * chunk_lo + (uint32_t)ks * kstep_lo);
                                            mma_i8_ts(d_tmem, a_tmem + (uint32_t)(32 * j + 8 * ks), desc, idesc, 1u);
                                        }
                                    }
                                }
                            } else
#pragma unroll
                            for (int j = 0; j < kCh; ++j) {
                                if (SPLIT && j == kHalf) tc_commit(&empty[s]);   // the first chunks of the stage may be refilled
                                if (j < nch && ((m_cur >> j) & 1u)) {
                                    if (F4) {   // two MMAs of K = 64 per 128-bit chunk (8 TMEM columns, two 32-element core matrices each)
#pragma unroll
                                        for (int mm = 0; mm < 2; ++mm) {
                                            const uint64_t desc = ((uint64_t)desc_hi << 32) | (uint64_t)(lo_s + (uint32_t)j * chunk_lo + (uint32_t)mm * kstep_lo);
                                            mma_f4_ts(d_tmem, a_tmem + (uint32_t)(16 * j + 8 * mm), desc, idesc, sfa, sfb, 1u);
                                        }
                                    } else {
#pragma unroll
                                        for (int ks = 0; ks < 4; ++ks) {
                                            const uint64_t desc = ((uint64_t)desc_hi << 32) | (uint64_t)(lo_s + (uint32_t)j * chunk_lo + (uint32_t)ks * kstep_lo);
                                            const uint32_t d_use = (dbg_alt_d && (ks & 1)) ? (d_tmem ^ (uint32_t)NP) : d_tmem;
                                            if (PAIR) mma_i8_ts_pair(d_use, a_tmem + (uint32_t)(32 * j + 8 * ks), desc, idesc, 1u);
                                            else mma_i8_ts(d_use, a_tmem + (uint32_t)(32 * j + 8 * ks), desc, idesc, 1u);
                                        }
                                    }
                                }
                            }
                            // arrives when these MMAs have finished reading A (TMEM) and B (smem); pair: in both CTAs
                            if (PAIR) tc_commit_pair(&empty[s]);
                            else tc_commit(SPLIT ? &empty2[s] : &empty[s]);
                        }
                        __syncwarp();
                        if (lane == 0 && trole >= 0) QGTC_TRACE(trole, item, 3);
                    } else if (!kOwnFull) {
                        // not this issuer's item: observe its phase and add this issuer's (in-order) commit
                        if (lane == 0 && trole >= 0) QGTC_TRACE(trole, item, 5);
                        mbar_wait(&full[s], use & 1, T.error_flag, 13, hinted);
                        if (lane == 0 && trole >= 0) QGTC_TRACE(trole, item, 6);
                        if (!kOwnerCommit) {
                            if (elect_one()) {
                                if (PAIR) tc_commit_pair(&empty[s]);
                                else tc_commit(&empty[s]);
                                if (SPLIT) tc_commit(&empty2[s]);
                            }
                            __syncwarp();
                        }
                    }
                    if (++owner == T.ni) owner = 0;
                    if (++s == S) { s = 0; ++use; }
                    ++item;
                  }
                }
                if (elect_one()) {   // arrives when this issuer's share of the accumulator is complete
                    if (PAIR) tc_commit_pair(&d_full[buf]);
                    else tc_commit(&d_full[buf]);
                }
                if (++buf == T.dbufs) { buf = 0; ++duse; }
            }
        }
    } else {
        // ===================== epilogue (8 warps: TMEM lane quarter x column half) =====================
        // a warp owns the 16-column blocks cb with (cb >> 1) % 2 == half, i.e. every other 32-column word
        const int q = warp & 3, half = (warp - kEpilogueWarp0) >> 2;   // half = index among the kEpiStep warps of this lane quarter
        const int M = P.M, N = P.N, ob = P.out_bit;
        const int dshift = amode == 1 ? 7 : (amode == 2 ? 3 : 0);   // the accumulators hold 2^dshift x the sums (4-bit path: f32 sums)
        const int wpr = s128(N) * 4;
        const unsigned long long opt_offset = (unsigned long long)p8(M) * wpr;
        int buf = 0;
        uint32_t duse = 0;
        // Accumulators are kept zeroed between tiles (every MMA accumulates): clear this warp's share of
        // both buffers once, then after each read.  Column blocks follow the same ownership as the reads.
        auto zero_my_columns = [&](int b) {
            uint32_t z[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) z[e] = 0u;
            for (int w32 = half; w32 * 32 < NP; w32 += kEpiStep) {
                const uint32_t col = (uint32_t)(b * NP + w32 * 32);
                if (w32 * 32 + 32 <= NP) tmem_st32(tmem_base + ((uint32_t)(32 * q) << 16) + col, z);
                else tmem_st16(tmem_base + ((uint32_t)(32 * q) << 16) + col, z);   // NP % 32 == 16: a trailing half word
            }
            tmem_wait_st();
        };
        if (F4 && half == 0) {   // unit UE8M0 block scales (2^0 = 0x7F) in every byte of the scale columns of this lane quarter
            uint32_t sf[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) sf[e] = 0x7F7F7F7Fu;
            static_assert(kSfCols == 16, "scale columns are written by one x16 store");
            tmem_st16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)sf_col, sf);
            tmem_wait_st();
        }
        for (int b = 0; b < T.dbufs; ++b) {
            zero_my_columns(b);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[b]);
        }
        for (int tl = 0; tl < my_tiles; ++tl) {
            const UnitGeo ug = unit_geo<PAIR>(T, group_id + tl * ngroups, rank);
            const int z = ug.z, m = ug.mt * 128 + 32 * q + lane;   // pair mode: rows past the matrix are masked by m < P8(M)
            if (warp == kEpilogueWarp0 && lane == 0) QGTC_TRACE(2, tl, 0);
            // Long wait (a whole tile).  ONE epilogue warp polls the barrier; the others park on a named barrier, which costs them a
            // single instruction: a polling warp comes back every ~40 ns whatever sleep it asks for (ncu: the instruction count does
            // not depend on it), and eight of them were a fifth of all instructions this kernel executed -- in a kernel paced by
            // the SM's instruction throughput.
            if (!QGTC_TC05_EPI_POLL1 || warp == kEpilogueWarp0) {
                uint32_t spins = 0;
                while (!mbar_try_wait(&d_full[buf], duse & 1)) {
                    __nanosleep(T.epi_sleep);
                    if (++spins > (1u << 24)) {
                        if (T.error_flag) { atomicExch_system(T.error_flag, 0x80000000u | 6u); __threadfence_system(); }
                        __trap();
                    }
                }
            }
            if (QGTC_TC05_EPI_POLL1) asm volatile("bar.sync 2, %0;\n" ::"n"(kEpilogueWarps * 32) : "memory");
            tc_fence_after();
            if (warp == kEpilogueWarp0 && lane == 0) QGTC_TRACE(2, tl, 1);
            const uint32_t d_tmem = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * NP);
            if (EPI == kEpiFloat) {
                float* out = reinterpret_cast<float*>(P.out) + (unsigned long long)z * P.out_bstride;
                for (int cb = 0; cb < NP / 16; ++cb) {
                    if (((cb >> 1) % kEpiStep) != half) continue;
                    int v[16];
                    tmem_ld16(d_tmem + 16 * cb, v);
                    if (m < M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = 16 * cb + j;
                            if (n < N) out[(size_t)m * N + n] = F4 ? __int_as_float(v[j]) : (float)(v[j] >> dshift);
                        }
                    }
                }
            } else {
                uint32_t* out = reinterpret_cast<uint32_t*>(P.out) + (unsigned long long)z * P.out_bstride;
                const bool row_ok = m < p8(M);
                const uint32_t rmask = m < M ? 0xffffffffu : 0u;
                uint32_t* out_row = out + (unsigned long long)m * wpr;
                // the plane loop is unrolled for the common widths (a runtime bound issues all 8 planes predicated)
#define QGTC_EPI_CASE(OBV, NNV, FASTV) epi_bits_tile<OBV, NNV, FASTV, F4, kEpiStep>(d_tmem, NP, half, N, ob, dshift, rmask, row_ok, out_row, opt_offset)
                if (F4 || dshift == 3) {   // 4-bit path (exact sums) / expansion mode 2 (8 x sums): non-negative; byte-parallel requantize up to 4 bits
                    switch (ob) {
                        case 1: QGTC_EPI_CASE(1, true, true); break;
                        case 2: QGTC_EPI_CASE(2, true, true); break;
                        case 4: QGTC_EPI_CASE(4, true, true); break;
                        case 8: QGTC_EPI_CASE(8, true, false); break;
                        default: if (ob <= 4) QGTC_EPI_CASE(0, true, true); else QGTC_EPI_CASE(0, true, false); break;
                    }
                } else if (dshift > 0) {   // scaled expansion: sums cannot wrap, accumulators are non-negative
                    switch (ob) {
                        case 1: QGTC_EPI_CASE(1, true, false); break;
                        case 2: QGTC_EPI_CASE(2, true, false); break;
                        case 4: QGTC_EPI_CASE(4, true, false); break;
                        case 8: QGTC_EPI_CASE(8, true, false); break;
                        default: QGTC_EPI_CASE(0, true, false); break;
                    }
                } else {
                    QGTC_EPI_CASE(0, false, false);
                }
#undef QGTC_EPI_CASE
                if (row_ok && half == 0) {   // zero pad words of the row (output buffers are not pre-zeroed)
                    for (int w = (NP + 31) / 32; w < wpr; ++w)
                        for (int p = 0; p < ob; ++p) out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w] = 0u;
                }
            }
            zero_my_columns(buf);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[buf]);
            if (warp == kEpilogueWarp0 && lane == 0) QGTC_TRACE(2, tl, 2);
            if (++buf == T.dbufs) { buf = 0; ++duse; }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();      // neither CTA may retire (or free its TMEM) while the other can still touch it
    else __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

unsigned int* g_error_flag = nullptr;  // device alias of g_error_flag_host
unsigned int* g_error_flag_host = nullptr;  // one mapped host word per process: where a watchdog fired (0x80000000 | site)
// kernel timing (qgtc_debug_tc05_time): cudaEvent pairs around the GEMM kernel launches, read by qgtc_debug_tc05_kernel_times
bool g_time_on = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_time_events;   // guarded by g_time_mu
std::mutex g_time_mu;
unsigned long long* g_trace = nullptr;  // clock64 trace buffer, only with QGTC_TC05_TRACE=1
constexpr size_t kTraceWords = 8 * (size_t)kTraceItems * 8;

}  // namespace

// the watchdog word of the tcgen05 engines (mapped host memory, so the code survives the trap that follows it); null on failure
unsigned int* tc05_error_flag_device() {
    if (!g_error_flag) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&g_error_flag_host), sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        *g_error_flag_host = 0u;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_error_flag), g_error_flag_host, 0) != cudaSuccess) { cudaGetLastError(); g_error_flag = nullptr; }
    }
    return g_error_flag;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// x1 must hold the whole R(1, M, K) plane of every problem for the TMA path (no zero-extension cases)
static bool a_in_bounds(const GemmArgs& a) { return a.x1_words >= (unsigned long long)p8(a.M) * a.gdk * 4; }

bool tc05_supports(const GemmArgs& a, int epi) {
    if (epi != kEpiBitsRow && epi != kEpiFloat) return false;
    if (a.bit1 != 1) return false;                 // the expanders handle a single adjacency plane
    if (p8(a.N) > 256) return false;               // one MMA N block
    if (!a_in_bounds(a) || !encode_tiled_fn()) return false;
    // measured single-call latency (profiles/r1_single_call_sweep.txt): tcgen05 wins from the smallest measured
    // problem up (759 x 759 x 16, 0.018 GOP: 6.2 us vs 8.0 us; 2258 x 2258 x 64: 8.2 us vs 37 us); below that the
    // two launches and the pipeline prologue are assumed to cost more than the single legacy kernel
    const double ops = 2.0 * a.M * (double)a.K * a.N * a.count;
    return ops >= 1.5e7;
}

int launch_gemm_tc05(const GemmArgs& a, int epi, cudaStream_t st) {
    if (epi != kEpiBitsRow && epi != kEpiFloat) return QGTC_ERR_UNSUPPORTED;
    if (a.bit1 != 1 || p8(a.N) > 256 || !a_in_bounds(a)) return QGTC_ERR_UNSUPPORTED;
    EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return QGTC_ERR_UNSUPPORTED;
    Tc05Args T;
    T.g = a;
    T.NP = ((p8(a.N) + 15) / 16) * 16;
    static const bool pair_env = [] { const char* e = getenv("QGTC_TC05_PAIR"); return e && e[0] == '1'; }();
    // pair mode (opt-in: QGTC_FLAG_TC05_PAIR or QGTC_TC05_PAIR=1): two CTAs of a cluster share every MMA
    // (cta_group::2, M = 256) and each streams half of the feature image.  Measured on B200 it does not beat the
    // single-CTA form (the adjacency expanders, not the tensor front end or L2, bound the kernel): DESIGN.md 5.4.
    const bool pair = (pair_env || (a.flags & QGTC_FLAG_TC05_PAIR)) && (T.NP % 16 == 0) && ((long long)((p8(a.M) + 127) / 128) * a.count >= 2);
    // 4-bit operand path (kind::mxf4, K = 64 per MMA; QGTC_TC05_FP4=0 disables): features of at most 2 bits are exact in
    // E2M1 with the row scaling of prep_b4_kernel, the f32 accumulators hold exact sums below 2^23.  Half the TMEM per
    // chunk: items of up to 7 chunks in three stages next to the double-buffered accumulator and the scale columns.
    static const bool f4_env = [] { const char* e = getenv("QGTC_TC05_FP4"); return !(e && e[0] == '0'); }();
    static const bool trace_env = [] { const char* t = getenv("QGTC_TC05_TRACE"); return t && t[0] == '1'; }();
    const double max_sum = (double)a.gdk * 128.0 * (double)((1u << a.bit2) - 1u);
    static const bool trace_f4_env = [] { const char* t = getenv("QGTC_TC05_TRACE_F4"); return t && t[0] == '1'; }();   // trace the 4-bit path instead of the u8 path
    const bool f4 = f4_env && (!trace_env || trace_f4_env) && !pair && (a.flags & 0x6F0000u) == 0 && a.bit2 <= 2 && max_sum < 8388608.0;
    if (f4) {
        // chunks per item: the largest of {7, 5, 4, 3, 2, 1} that fits three stages and K (6 and 8 chunks would put the
        // 128 rows of a raw slot on a stride with shared-memory bank conflicts; 4 and 2 use the TMA swizzle)
        // accumulators: double-buffered up to N = 128, single above (as on the u8 path)
        T.dbufs = T.NP <= 128 ? 2 : 1;
        int cap = (kTmemCols - T.dbufs * T.NP - kSfCols) / 48;
        const int smem_cap = (180 * 1024) / (3 * (T.NP * 64 + 2 * 128 * 16));   // three feature stages + six raw slots per chunk of an item
        if (cap > smem_cap) cap = smem_cap;
        const int lim = cap < a.gdk ? cap : a.gdk;
        T.cps = lim >= 7 ? 7 : (lim >= 5 ? 5 : lim);
        static const int cps4_env = [] { const char* e = getenv("QGTC_TC05_CPS4"); const int v = e ? atoi(e) : 0; return (v == 1 || v == 2 || v == 3 || v == 4 || v == 5 || v == 7) ? v : 0; }();
        if (cps4_env && cps4_env <= T.cps) T.cps = cps4_env;   // experiment knob
        T.amode = 3;
    } else {
        // chunks per pipeline item: 256 / N_pad, at most 4 -- three A stages of 128 columns beside the double-buffered
        // accumulator for N <= 64.  Measured alternative (QGTC_TC05_CPS=3: four stages of three chunks, 2 x 64 + 4 x 96
        // = 512 columns): 1907 vs 2135 TOPS -- a fourth chunk per item is worth more than a fourth stage.
        T.cps = T.NP <= 64 ? 4 : (T.NP <= 128 ? 2 : 1);
        static const int cps_env = [] { const char* e = getenv("QGTC_TC05_CPS"); const int v = e ? atoi(e) : 0; return (v >= 1 && v <= kMaxCps) ? v : 0; }();
        if (cps_env && cps_env <= T.cps) T.cps = cps_env;       // experiment knob: smaller items (3: four stages for N <= 64)
        if (T.cps > a.gdk) T.cps = a.gdk;                       // short K: do not pad stages with empty chunks
    }
    T.nst = (a.gdk + T.cps - 1) / T.cps;
    T.cps_inv = T.cps == 1 ? 0u : (unsigned int)((0x100000000ull + (unsigned long long)T.cps - 1ull) / (unsigned long long)T.cps);
    // expander warps per item: one per TMEM lane quarter takes all chunks (wide layout: two per quarter split them)
    T.wch = (QGTC_TC05_WIDE && !f4 && T.cps >= 2) ? T.cps / 2 : T.cps;
    T.wpi = 4 * T.cps / T.wch;
    if (!f4) T.dbufs = (2 * T.NP + 2 * T.cps * 32 <= kTmemCols) ? 2 : 1;
    int stages = f4 ? (kTmemCols - T.dbufs * T.NP - kSfCols) / (T.cps * 16) : (kTmemCols - T.dbufs * T.NP) / (T.cps * 32);
    if (stages > kMaxStages) stages = kMaxStages;
    T.rslots = f4 ? 6 : kRawSlots;      // 4-bit path: raw slots are up to 14 KB
    // (Converting the features inside this kernel, by an expander group without items, was built and measured on
    // B200 -- 4096 x 4096 x 64, 96 graphs: +17 us at 1 bit, +83 us at 8 bits per launch against 8-17 us for the
    // separate prep kernel: the GEMM warps have no issue slots to spare.  Removed)
    while (stages > 2 && smem_layout(T.NP, stages, T.cps, pair, T.rslots, f4).total > 200 * 1024) --stages;
    if (stages < 2) return QGTC_ERR_UNSUPPORTED;
    T.stages = stages;
    T.mtiles = (p8(a.M) + 127) / 128;
    T.total_tiles = T.mtiles * a.count;
    T.spg = (T.mtiles + 1) / 2;
    T.units = pair ? T.spg * a.count : T.total_tiles;
    T.epi = epi;
    T.one = 1u;
    {
        static const int sleep_env = [] { const char* e = getenv("QGTC_TC05_EPI_SLEEP"); return e ? atoi(e) : 0; }();
        T.epi_sleep = sleep_env > 0 ? (unsigned)sleep_env : 512u;
    }
    {
        static const int ni_env = [] { const char* e = getenv("QGTC_TC05_ISSUERS"); return e ? atoi(e) : 0; }();
        T.ni = (ni_env >= 1 && ni_env <= kMaxIssuers) ? ni_env : 2;
        if (pair && T.ni < 2) T.ni = 2;   // the peer CTA's first two issuer warps are its relays
    }
    if (!f4) {   // expansion mode: the scaled sums must fit an int32 (mode 1: K < 65793 for 8-bit features)
        static const int amode_env = [] { const char* e = getenv("QGTC_TC05_AMODE"); return e ? atoi(e) : -1; }();
        const bool ok2 = a.bit2 <= 5 && max_sum * 8.0 < 2147483648.0, ok1 = max_sum * 128.0 < 2147483648.0;
        T.amode = ok2 ? 2 : (ok1 ? 1 : 0);
        if (amode_env == 0 || (amode_env == 1 && ok1) || (amode_env == 2 && ok2)) T.amode = amode_env;   // experiment knob
    }
    T.bitmap = a.tile_bitmap;
    T.bmw = (a.gdk + 31) / 32;
    if ((long long)((T.total_tiles + kNumSMs - 1) / kNumSMs + 1) * T.nst > (1ll << 30)) return QGTC_ERR_UNSUPPORTED;
    const bool b_shared = a.count > 1 && a.x2_bstride == 0;   // every problem aliases one feature matrix: one image
    const unsigned long long chunk_img_bytes = (unsigned long long)T.NP * (f4 ? 64 : 128);   // image bytes of one 128-element K chunk
    T.bimg_gstride = b_shared ? 0ull : (unsigned long long)a.gdk * chunk_img_bytes;

    // tensor map of the packed adjacency: uint32 [count][P8(M)][gdk*4], box = [1][128][4*cps]
    alignas(64) CUtensorMap tmap;
    {
        T.a_batched = (a.count > 1 && a.x1_bstride != 0) ? 1 : 0;
        const cuuint64_t row_bytes = (cuuint64_t)a.gdk * 16;
        cuuint64_t dims[3] = {(cuuint64_t)a.gdk * 4, (cuuint64_t)p8(a.M), (cuuint64_t)(T.a_batched ? a.count : 1)};
        cuuint64_t strides[2] = {row_bytes, T.a_batched ? (cuuint64_t)a.x1_bstride * 4 : row_bytes * (cuuint64_t)p8(a.M)};
        cuuint32_t box[3] = {(cuuint32_t)(4 * T.cps), 128u, 1u};
        cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUtensorMapSwizzle sw = T.cps == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : (T.cps == 2 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
        const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t*>(a.x1), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return QGTC_ERR_UNSUPPORTED;
    }

    cudaError_t e;
    if (!g_error_flag) {   // mapped host memory: the watchdog's code survives the trap that follows it
        e = cudaHostAlloc(reinterpret_cast<void**>(&g_error_flag_host), sizeof(unsigned int), cudaHostAllocMapped);
        if (e != cudaSuccess) return record_cuda_error(e, "cudaHostAlloc(tc05 error flag)");
        *g_error_flag_host = 0u;
        e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_error_flag), g_error_flag_host, 0);
        if (e != cudaSuccess) return record_cuda_error(e, "cudaHostGetDevicePointer(tc05 error flag)");
    }
    T.error_flag = g_error_flag;
    static const bool want_trace = [] { const char* t = getenv("QGTC_TC05_TRACE"); return t && t[0] == '1'; }();
    if (want_trace && !g_trace) {
        if (cudaMalloc(&g_trace, kTraceWords * 8) != cudaSuccess) g_trace = nullptr;
    }
    if (g_trace) cudaMemsetAsync(g_trace, 0, kTraceWords * 8, st);
    T.trace = g_trace;
    {   // keep the stream-ordered pool from handing memory back to the OS at every synchronisation
        static thread_local int pool_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != pool_dev) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            cudaGetLastError();
            pool_dev = dev;
        }
    }
    uint8_t* img = nullptr;
    const int img_problems = b_shared ? 1 : a.count;
    const size_t img_bytes = (((size_t)a.gdk * chunk_img_bytes * img_problems) + 255) & ~(size_t)255;
    e = cudaMallocAsync(reinterpret_cast<void**>(&img), img_bytes, st);
    if (e != cudaSuccess) return record_cuda_error(e, "cudaMallocAsync(feature image)");
    T.bimg = img;

    int rc = QGTC_OK;
    {
        GemmArgs pa = a;
        pa.count = img_problems;
        long long pgrid = (long long)img_problems * ((a.gdk + kPrepChunks - 1) / kPrepChunks) * (T.NP / 8);
        if (pgrid > kNumSMs * 8) pgrid = kNumSMs * 8;
        const size_t psmem = (size_t)a.bit2 * 8 * (kPrepChunks + 1) * sizeof(uint4);    // <= 33 KB
        const unsigned long long gs = (unsigned long long)a.gdk * chunk_img_bytes;
        const int halves = pair ? 2 : 1;
        cudaLaunchConfig_t pcfg = {};
        pcfg.gridDim = dim3((unsigned)pgrid);
        pcfg.blockDim = dim3(256);
        pcfg.dynamicSmemBytes = psmem;
        pcfg.stream = st;
        cudaLaunchAttribute pattr[1];
        pattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        pattr[0].val.programmaticStreamSerializationAllowed = 1;
        pcfg.attrs = pattr;
        static const bool prep_pdl = [] { const char* e = getenv("QGTC_TC05_PREP_PDL"); return !(e && e[0] == '0'); }();
        pcfg.numAttrs = prep_pdl ? 1 : 0;
        const int np_ = T.NP;
        if (f4) {
            if (a.bit2 == 1) cudaLaunchKernelEx(&pcfg, prep_b4_kernel<1>, pa, img, gs, np_);
            else cudaLaunchKernelEx(&pcfg, prep_b4_kernel<2>, pa, img, gs, np_);
        } else
#define QGTC_PREP(B)                                                                              \
    case B:                                                                                       \
        if (T.amode == 2) cudaLaunchKernelEx(&pcfg, prep_b_kernel<B, true>, pa, img, gs, np_, halves); \
        else cudaLaunchKernelEx(&pcfg, prep_b_kernel<B, false>, pa, img, gs, np_, halves);          \
        break;
        switch (a.bit2) {
            QGTC_PREP(1) QGTC_PREP(2) QGTC_PREP(3) QGTC_PREP(4) QGTC_PREP(5) QGTC_PREP(6) QGTC_PREP(7) QGTC_PREP(8)
            default: cudaFreeAsync(img, st); return QGTC_ERR_UNSUPPORTED;
        }
#undef QGTC_PREP
        note_launch();
        rc = check_launch("prep_b_kernel");
    }

    if (rc == QGTC_OK) {
        const int smem = smem_layout(T.NP, T.stages, T.cps, pair, T.rslots, f4).total + 1024;
        const int groups = pair ? (T.units < kNumSMs / 2 ? T.units : kNumSMs / 2) : (T.units < kNumSMs ? T.units : kNumSMs);
        auto launch = [&](auto kfn) -> int {
            cudaError_t e2 = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e2 != cudaSuccess) return record_cuda_error(e2, "cudaFuncSetAttribute(gemm_tc05)");
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(pair ? 2 * groups : groups);
            cfg.blockDim = dim3((kMmaWarp + T.ni) * 32);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = pair ? 2 : 1;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            // start while prep_b_kernel drains (the kernel's image loader does the griddepcontrol.wait)
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pair ? 1 : 2;
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            if (g_time_on && cudaEventCreate(&ev0) == cudaSuccess && cudaEventCreate(&ev1) == cudaSuccess) cudaEventRecord(ev0, st);
            e2 = cudaLaunchKernelEx(&cfg, kfn, T, tmap);
            if (ev0 && ev1) {
                cudaEventRecord(ev1, st);
                std::lock_guard<std::mutex> lk(g_time_mu);
                g_time_events.emplace_back(ev0, ev1);
            }
            note_launch();
            if (e2 != cudaSuccess) return record_cuda_error(e2, "cudaLaunchKernelEx(gemm_tc05)");
            return check_launch("gemm_tc05_kernel");
        };
        const bool knobs = (T.g.flags & 0x6F0000u) != 0, trace = T.trace != nullptr;
        // compile-time chunks per item for the hot configurations (QGTC_TC05_FIXED=0: the runtime-cps instances, for A/B runs)
        static const bool fixed_env = [] { const char* e = getenv("QGTC_TC05_FIXED"); return !(e && e[0] == '0'); }();
        const bool fixed_ok = fixed_env && !QGTC_TC05_WIDE && T.wch == T.cps;
        const bool fl = epi == kEpiFloat;
        if (knobs) {
            if (pair) rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, true, 2, -1>) : launch(gemm_tc05_kernel<kEpiBitsRow, true, 2, -1>);
            else rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 2, -1>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 2, -1>);
        } else if (pair) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, true, 0, -1>) : launch(gemm_tc05_kernel<kEpiBitsRow, true, 0, -1>);
        } else if (T.amode == 3 && T.cps == 7 && fixed_ok && trace && !fl) {
            rc = launch(gemm_tc05_kernel<kEpiBitsRow, false, 1, 3, 7>);    // the production code path + clock reads
        } else if (T.amode == 3 && T.cps == 7 && fixed_ok) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 3, 7>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 3, 7>);
        } else if (T.amode == 3) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 3>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 3>);
        } else if (trace && !fl && T.amode == 2 && T.cps == 4 && fixed_ok) {
            rc = launch(gemm_tc05_kernel<kEpiBitsRow, false, 1, 2, 4>);    // the production code path + clock reads
        } else if (trace && !fl && T.amode == 2) {
            rc = launch(gemm_tc05_kernel<kEpiBitsRow, false, 1, 2>);
        } else if (trace) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 2, -1>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 2, -1>);
        } else if (T.amode == 2 && T.cps == 4 && fixed_ok) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 2, 4>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 2, 4>);
        } else if (T.amode == 1 && T.cps == 4 && fixed_ok) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 1, 4>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 1, 4>);
        } else if (T.amode == 2) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 2>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 2>);
        } else if (T.amode == 1) {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 1>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 1>);
        } else {
            rc = fl ? launch(gemm_tc05_kernel<kEpiFloat, false, 0, 0>) : launch(gemm_tc05_kernel<kEpiBitsRow, false, 0, 0>);
        }
    }
    e = cudaFreeAsync(img, st);
    if (rc == QGTC_OK && e != cudaSuccess) rc = record_cuda_error(e, "cudaFreeAsync(feature image)");
    return rc;
}

// debugging aid (not part of the public ABI): the watchdog word (0 = no watchdog fired); readable after a trap
extern "C" unsigned int qgtc_debug_tc05_error_flag() { return g_error_flag_host ? *reinterpret_cast<volatile unsigned int*>(g_error_flag_host) : 0u; }

// measurement aid (not part of the public ABI): with timing on, every gemm_tc05_kernel launch is bracketed by a
// pair of CUDA events on its stream (which also keeps the kernel from overlapping the tail of prep_b_kernel);
// qgtc_debug_tc05_kernel_times waits for them and returns the kernels' durations in ms, oldest first.
extern "C" void qgtc_debug_tc05_time(int on) { g_time_on = on != 0; }
extern "C" int qgtc_debug_tc05_kernel_times(float* dst, int max_n) {
    int n = 0;
    std::lock_guard<std::mutex> lk(g_time_mu);
    for (auto& pr : g_time_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess && n < max_n) dst[n++] = ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_time_events.clear();
    return n;
}

// debugging aid (not part of the public ABI): copy the last clock64 trace to the host
extern "C" int qgtc_debug_tc05_trace(unsigned long long* dst, size_t max_words) {
    if (!g_trace) return 0;
    const size_t n = max_words < kTraceWords ? max_words : kTraceWords;
    cudaDeviceSynchronize();
    cudaMemcpy(dst, g_trace, n * 8, cudaMemcpyDeviceToHost);
    return (int)n;
}

}  // namespace qgtc
