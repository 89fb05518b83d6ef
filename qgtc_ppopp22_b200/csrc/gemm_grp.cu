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
constexpr int kThreads = (kIssuerWarp + 1) * 32;   // 544
constexpr int kARing = 3;                     // staged pipeline items per expander warp (cp.async groups in flight)
constexpr int kBRing = 2;                     // ... per converter thread
constexpr int kMaxTasks = 2;                  // (chunk, column) conversion tasks per converter thread and item: cps * NP <= 256
constexpr int kMaxStagesG = 4;
constexpr int kMaxCpsG = 4;
constexpr int kGrpMaxProblems = 224;          // kernel-parameter table (64 B each): one launch takes up to this many problems

struct GrpProblem {
    const uint32_t* x1;
    const uint32_t* x2;
    void* out;
    const uint32_t* bitmap;                   // optional tile bitmap of this problem's 1-bit A operand, or null
    unsigned long long x1_words, x2_words;    // words really present (zero-extension beyond)
    int M, K;
    int unit0;                                // index of this problem's first work unit
    int pad_;
};
static_assert(sizeof(GrpProblem) == 64, "problem descriptors are 64 bytes");

struct GrpArgs {
    int count, units;
    int N, NP, bit1, bit2, out_bit, epi, pad128;
    int cps, stages, dbufs, groups;           // chunks per item, A/B ring depth, accumulator buffers, expander groups in use
    int a_slot_bytes, b_slot_bytes, b_stage_bytes;
    int off_bstage, off_bring, off_bars, off_flags, off_misc;   // shared-memory offsets from the 1024-byte aligned base
    unsigned int cps_inv;
    unsigned flags;
    unsigned int* error_flag;
    GrpProblem p[kGrpMaxProblems];
};

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
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

// The sequence of pipeline items of one CTA.  Every role walks it with its own Walk and must see the same sequence:
// units u = blockIdx.x, blockIdx.x + gridDim.x, ...; inside a unit the stages (cps chunks each) in increasing order, all of
// them without a tile bitmap, else stage 0 plus every stage with a non-zero chunk bit.
struct Walk {
    int tl, u, z, mt, gdk, nst;        // unit counter of this CTA, unit, problem, row tile, K chunks, stages of the problem
    const uint32_t* bm;                // bitmap words of tile (z, mt), or null
    int bmw, w, st_next, covered_end;  // stage iterator (see StageIter in gemm_tc05.cu)
    uint32_t m;
    bool first, in_unit, done;
    int item;                          // number of items next() has returned so far (the last one's CTA-wide index + 1)

    __device__ __forceinline__ void init() {
        tl = -1; item = 0; nst = 0; st_next = 0; bm = nullptr; first = false; in_unit = false; done = false;
        m = 0u; w = 0; bmw = 0; covered_end = 0; u = 0; z = 0; mt = 0; gdk = 0;
    }

    // position on this CTA's next unit; false when the work is exhausted
    __device__ __forceinline__ bool next_unit(const GrpArgs& T) {
        ++tl;
        u = (int)blockIdx.x + tl * (int)gridDim.x;
        if (u >= T.units) return false;
        int lo = 0, hi = T.count - 1;      // largest z with unit0[z] <= u
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (T.p[mid].unit0 <= u) lo = mid; else hi = mid - 1;
        }
        z = lo;
        mt = u - T.p[z].unit0;
        gdk = s128(T.p[z].K);
        nst = (gdk + T.cps - 1) / T.cps;
        bmw = (gdk + 31) >> 5;
        bm = T.p[z].bitmap ? T.p[z].bitmap + (size_t)mt * bmw : nullptr;
        st_next = 0; w = -1; m = 0u; covered_end = 0; first = true;
        return true;
    }
    // next stage of the current unit, or -1
    __device__ __forceinline__ int next_stage(const GrpArgs& T) {
        if (!bm) return st_next < nst ? st_next++ : -1;
        if (first) { first = false; covered_end = T.cps; return 0; }
        for (;;) {
            if (m == 0u) {
                if (++w >= bmw) return -1;
                const int left = gdk - 32 * w;
                const uint32_t valid = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                m = __ldg(bm + w) & valid;
                const int n = covered_end - 32 * w;
                if (n > 0) m &= n >= 32 ? 0u : ~((1u << n) - 1u);
                continue;
            }
            const int c = 32 * w + __ffs(m) - 1;
            const int st = T.cps_inv ? (int)__umulhi((unsigned)c, T.cps_inv) : c;
            covered_end = (st + 1) * T.cps;
            const int n = covered_end - 32 * w;
            m &= n >= 32 ? 0u : ~((1u << n) - 1u);
            return st;
        }
    }
    // next item of the CTA: (z, mt) are the fields above, c0 / nch its chunks.  false (for good) when the work is exhausted.
    __device__ __forceinline__ bool next(const GrpArgs& T, int& c0, int& nch) {
        for (;;) {
            if (in_unit) {
                const int st = next_stage(T);
                if (st >= 0) {
                    c0 = st * T.cps;
                    nch = gdk - c0 < T.cps ? gdk - c0 : T.cps;
                    ++item;
                    return true;
                }
                in_unit = false;
            }
            if (done || !next_unit(T)) { done = true; return false; }
            in_unit = true;
        }
    }
};

template <bool A1>   // A1: single-plane A operand (the adjacency products): the lean expansion
__global__ void __launch_bounds__(kThreads, 1) gemm_grp_kernel(const __grid_constant__ GrpArgs T) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NP = T.NP, S = T.stages, CPS = T.cps;

    unsigned char* a_rings = smem;                                   // [kExpWarps][kARing][a_slot_bytes]
    unsigned char* b_stages = smem + T.off_bstage;                   // [S][cps * NP * 128]
    unsigned char* b_ring = smem + T.off_bring;                      // [kBRing][b_slot_bytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T.off_bars);
    uint64_t* full = bars;                         // [kMaxStagesG] 4 expander warps + kConvWarps converter warps -> MMA
    uint64_t* empty = bars + kMaxStagesG;          // [kMaxStagesG] MMA commit -> expanders, converters
    uint64_t* d_full = bars + 2 * kMaxStagesG;     // [2] MMA commit -> epilogue
    uint64_t* d_empty = d_full + 2;                // [2] 4 epilogue warps -> MMA
    volatile unsigned char* flags = smem + T.off_flags;   // [stage][chunk][lane quarter]: that 32 x 128-bit block has a one
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + T.off_misc);

    // let the next kernel of the stream be scheduled as SMs free up (it waits for this grid in its own griddepcontrol.wait)
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStagesG; ++s) {
            mbar_init(&full[s], 4 + kConvWarps);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&d_full[b], 1);
            mbar_init(&d_empty[b], 4);
        }
        fence_mbar_init();
    }
    if (warp == kIssuerWarp) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results (this kernel's
    // operands) are complete and visible only after this wait.  No-op for a normal launch.
    asm volatile("griddepcontrol.wait;\n" ::: "memory");

    const int dcols = T.dbufs * NP;                // accumulators first, then the A stages
    const int stage_cols = CPS * 32;
    const bool nojump = (T.flags & QGTC_FLAG_NO_ZERO_JUMP) != 0;
    unsigned int* const ef = T.error_flag;

    if (warp < kExpWarps) {
        // ===================== A expanders =====================
        const int q = warp & 3, g = warp >> 2;
        const int G = T.groups;
        if (g < G) {
            unsigned char* ring = a_rings + (size_t)warp * kARing * T.a_slot_bytes;
            Walk ld, cs;
            ld.init();
            cs.init();
            // issue the cp.async loads of an item into a ring slot: piece (p, j) of this lane's row at [(p * CPS + j) * 32 + lane] x 16 B
            auto issue = [&](const Walk& wk, int c0, int nch, int slot) {
                const GrpProblem& P = T.p[wk.z];
                const int m = wk.mt * 128 + 32 * q + lane;
                const unsigned long long gdk4 = (unsigned long long)wk.gdk * 4;
                const unsigned long long act_offset = (unsigned long long)p8(P.M) * gdk4;     // reference kernel.h:265
                const bool row_ok = m < p8(P.M);
                unsigned char* dst = ring + (size_t)slot * T.a_slot_bytes + lane * 16;
                for (int p = 0; p < T.bit1; ++p) {
                    for (int j = 0; j < nch; ++j) {
                        const unsigned long long idx = (unsigned long long)p * act_offset + (unsigned long long)m * gdk4 + 4ull * (c0 + j);
                        const int nb = row_ok ? piece_bytes(idx, P.x1_words) : 0;
                        cp_async16_zfill(dst + (size_t)(p * CPS + j) * 512, nb ? (const void*)(P.x1 + idx) : (const void*)P.x1, nb);
                    }
                }
            };
            // advance a walk to this group's next item (items are dealt to the groups round-robin)
            auto next_own = [&](Walk& wk, int& c0, int& nch) -> bool {
                while (wk.next(T, c0, nch))
                    if ((wk.item - 1) % G == g) return true;
                return false;
            };
            int lc0 = 0, lnch = 0;
            bool more = true;
#pragma unroll 1
            for (int i = 0; i < kARing - 1; ++i) {
                if (more && (more = next_own(ld, lc0, lnch))) issue(ld, lc0, lnch, i);
                cp_async_commit();
            }
            int k = 0, c0 = 0, nch = 0;
#pragma unroll 1
            while (next_own(cs, c0, nch)) {
                if (more && (more = next_own(ld, lc0, lnch))) issue(ld, lc0, lnch, (k + kARing - 1) % kARing);
                cp_async_commit();
                cp_async_wait<kARing - 1>();       // this lane's pieces of item k have landed (only this lane reads them)
                const int it = cs.item - 1;        // CTA-wide item index: stage and phase
                const int s = it % S;
                const uint32_t use = (uint32_t)(it / S);
                const unsigned char* src = ring + (size_t)(k % kARing) * T.a_slot_bytes + lane * 16;
                const uint32_t st_s = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(dcols + s * stage_cols);
                unsigned nzmask = 0;
#pragma unroll 1
                for (int j = 0; j < nch; ++j) {
                    uint32_t r[32];
                    uint32_t any;
                    if (A1) {
                        const uint4 pv = *reinterpret_cast<const uint4*>(src + (size_t)j * 512);
                        const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
                        any = pv.x | pv.y | pv.z | pv.w;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            r[ks * 8] = w[ks] & 0x01010101u;
#pragma unroll
                            for (int c = 1; c < 8; ++c) r[ks * 8 + c] = (w[ks] >> c) & 0x01010101u;
                        }
                    } else {
                        any = 0u;
#pragma unroll
                        for (int e = 0; e < 32; ++e) r[e] = 0u;
                        for (int p = 0; p < T.bit1; ++p) {
                            const uint4 pv = *reinterpret_cast<const uint4*>(src + (size_t)(p * CPS + j) * 512);
                            const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
                            any |= pv.x | pv.y | pv.z | pv.w;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                                for (int c = 0; c < 8; ++c) r[ks * 8 + c] |= ((w[ks] >> c) & 0x01010101u) << p;
                            }
                        }
                    }
                    if (nojump || __any_sync(0xffffffffu, any != 0u)) nzmask |= 1u << j;
                    if (j == 0) {   // the first chunk is expanded before the wait: only the store sits behind it
                        mbar_wait<32>(&empty[s], (use & 1) ^ 1, ef, 1, true);     // the MMAs of this stage's previous use are done
                        tc_fence_after();
                    }
                    tmem_st32(st_s + (uint32_t)(32 * j), r);
                }
                if (lane < nch) flags[(s * kMaxCpsG + lane) * 4 + q] = (unsigned char)((nzmask >> lane) & 1u);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                ++k;
            }
            cp_async_wait<0>();
        }
    } else if (warp >= kConvWarp0 && warp < kConvWarp0 + kConvWarps) {
        // ===================== B converters =====================
        // thread t takes the (chunk, column) tasks t, t + 128 of every item: bit2 pieces of 16 B (128 K bits of one column,
        // one plane each) in, 128 u8 values in the MMA's K order out, as 8 core-matrix rows of 16 B.
        const int t = threadIdx.x - kConvWarp0 * 32;
        Walk ld, cs;
        ld.init();
        cs.init();
        const int p8n = p8(T.N);
        auto issue = [&](const Walk& wk, int c0, int nch, int slot) {
            const GrpProblem& P = T.p[wk.z];
            const unsigned long long gdk4 = (unsigned long long)wk.gdk * 4;
            const unsigned long long w_offset = gdk4 * (unsigned long long)(T.pad128 ? p128(T.N) : p8n);   // kernel.h:266 / :836
            unsigned char* dst = b_ring + (size_t)slot * T.b_slot_bytes + t * 16;
            for (int i = 0; i < kMaxTasks; ++i) {
                const int task = t + i * kConvThreads;
                if (task >= nch * NP) break;
                const int j = task / NP, n = task - j * NP;
                for (int qq = 0; qq < T.bit2; ++qq) {
                    const unsigned long long idx = (unsigned long long)qq * w_offset + (unsigned long long)n * gdk4 + 4ull * (c0 + j);
                    const int nb = n < p8n ? piece_bytes(idx, P.x2_words) : 0;
                    cp_async16_zfill(dst + (size_t)(i * T.bit2 + qq) * (kConvThreads * 16), nb ? (const void*)(P.x2 + idx) : (const void*)P.x2, nb);
                }
            }
        };
        int lc0 = 0, lnch = 0;
        bool more = true;
#pragma unroll 1
        for (int i = 0; i < kBRing - 1; ++i) {
            if (more && (more = ld.next(T, lc0, lnch))) issue(ld, lc0, lnch, i);
            cp_async_commit();
        }
        int k = 0, c0 = 0, nch = 0;
#pragma unroll 1
        while (cs.next(T, c0, nch)) {
            if (more && (more = ld.next(T, lc0, lnch))) issue(ld, lc0, lnch, (k + kBRing - 1) % kBRing);
            cp_async_commit();
            cp_async_wait<kBRing - 1>();
            const int s = k % S;
            const uint32_t use = (uint32_t)(k / S);
            mbar_wait<32>(&empty[s], (use & 1) ^ 1, ef, 2, true);        // the MMAs that read this stage's previous content are done
            const unsigned char* src = b_ring + (size_t)(k % kBRing) * T.b_slot_bytes + t * 16;
            unsigned char* stage = b_stages + (size_t)s * T.b_stage_bytes;
#pragma unroll 1
            for (int i = 0; i < kMaxTasks; ++i) {
                const int task = t + i * kConvThreads;
                if (task >= nch * NP) break;
                const int j = task / NP, n = task - j * NP;
                uint32_t o[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) o[e] = 0u;
                for (int qq = 0; qq < T.bit2; ++qq) {
                    const uint4 pv = *reinterpret_cast<const uint4*>(src + (size_t)(i * T.bit2 + qq) * (kConvThreads * 16));
                    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
                    // byte i of o[8 s + c] = value of K element 32 s + c + 8 i (the expanders' K permutation)
#pragma unroll
                    for (int sdx = 0; sdx < 4; ++sdx) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) o[sdx * 8 + c] |= ((w[sdx] >> c) & 0x01010101u) << qq;
                    }
                }
                // image of one K chunk: [k group (8)][column block][8 columns][16 B]; k group kg = registers 4 kg .. 4 kg + 3
                unsigned char* dstc = stage + (size_t)j * NP * 128 + (size_t)(n >> 3) * 128 + (n & 7) * 16;
#pragma unroll
                for (int kg = 0; kg < 8; ++kg)
                    *reinterpret_cast<uint4*>(dstc + (size_t)kg * (NP / 8) * 128) = make_uint4(o[4 * kg], o[4 * kg + 1], o[4 * kg + 2], o[4 * kg + 3]);
            }
            fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's (async proxy) reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            ++k;
        }
        cp_async_wait<0>();
    } else if (warp == kIssuerWarp) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_i8(NP, 128);
        const uint32_t lbo = (uint32_t)(NP / 8) * 128u, sbo = 128u;
        const uint64_t desc0 = make_smem_desc(smem_u32(b_stages), lbo, sbo);
        const uint32_t desc_hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
        const uint32_t stage_lo = (uint32_t)T.b_stage_bytes >> 4, chunk_lo = ((uint32_t)NP * 128u) >> 4, kstep_lo = (2u * lbo) >> 4;
        Walk wk;
        wk.init();
        int s = 0, buf = 0;
        uint32_t use = 0, duse = 0;
#pragma unroll 1
        while (wk.next_unit(T)) {
            mbar_wait(&d_empty[buf], (duse & 1) ^ 1, ef, 3, true);       // the epilogue has drained this accumulator's previous tile
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * NP);
            uint32_t started = 0;                                         // 0: the next MMA overwrites the accumulator
#pragma unroll 1
            for (int st; (st = wk.next_stage(T)) >= 0;) {
                const int c_first = st * CPS;
                const int nch = wk.gdk - c_first < CPS ? wk.gdk - c_first : CPS;
                mbar_wait(&full[s], use & 1, ef, 4, true);
                tc_fence_after();
                const uint32_t fw = lane < kMaxCpsG ? *reinterpret_cast<volatile uint32_t*>(smem + T.off_flags + (s * kMaxCpsG + lane) * 4) : 0u;
                const uint32_t m4 = __ballot_sync(0xffffffffu, fw != 0u);
                const uint32_t a_tmem = tmem_base + (uint32_t)(dcols + s * stage_cols);
                const uint32_t lo_s = lo0 + (uint32_t)s * stage_lo;
                // an all-zero chunk adds nothing and is skipped -- except the unit's very first MMAs, which initialise D
                uint32_t todo = m4 & ((1u << nch) - 1u);
                if (!started) todo |= 1u;
                if (elect_one()) {
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
                }
                started = 1u;                      // todo != 0 whenever started was 0
                __syncwarp();
                if (++s == S) { s = 0; ++use; }
            }
            if (elect_one()) tc_commit(&d_full[buf]);
            __syncwarp();
            if (++buf == T.dbufs) { buf = 0; ++duse; }
        }
    } else {
        // ===================== epilogue (4 warps: one per TMEM lane quarter) =====================
        const int q = warp & 3;
        const int N = T.N, ob = T.out_bit;
        Walk wk;
        wk.init();
        int buf = 0;
        uint32_t duse = 0;
#pragma unroll 1
        while (wk.next_unit(T)) {
            const GrpProblem& P = T.p[wk.z];
            const int M = P.M;
            const int m = wk.mt * 128 + 32 * q + lane;
            mbar_wait<256>(&d_full[buf], duse & 1, ef, 6, false);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * NP);
            if (T.epi == kEpiFloat) {
                float* out = reinterpret_cast<float*>(P.out);
                for (int cb = 0; cb < NP / 16; ++cb) {
                    int v[16];
                    tmem_ld16(d_tmem + 16 * cb, v);
                    if (m < M) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = 16 * cb + j;
                            if (n < N) out[(size_t)m * N + n] = (float)v[j];
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
#define QGTC_GRP_EPI(OBV) epi_bits_tile<OBV, false, false, false, 1>(d_tmem, NP, 0, N, ob, 0, rmask, row_ok, out_row, opt_offset)
                switch (ob) {
                    case 1: QGTC_GRP_EPI(1); break;
                    case 2: QGTC_GRP_EPI(2); break;
                    case 4: QGTC_GRP_EPI(4); break;
                    case 8: QGTC_GRP_EPI(8); break;
                    default: QGTC_GRP_EPI(0); break;
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
                const int mw = wk.mt * 4 + q;
                const unsigned keep = (1u << ob) - 1u;
                if ((unsigned long long)mw < kw) {
                    for (int cb = 0; cb < NP / 16; ++cb) {
                        int v[16];
                        tmem_ld16(d_tmem + 16 * cb, v);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = 16 * cb + j;
                            const uint32_t val = (m < M && n < N) ? ((uint32_t)requant(v[j], ob) & keep) : 0u;
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
            if (++buf == T.dbufs) { buf = 0; ++duse; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kIssuerWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

struct GrpPlan {
    int NP, cps, stages, dbufs, groups, a_slot, b_slot, b_stage, off_bstage, off_bring, off_bars, off_flags, off_misc, smem_total;
};

// pipeline geometry for (N, bit1, bit2); false when the shape does not fit this engine
bool plan_for(int N, int bit1, int bit2, GrpPlan& pl) {
    pl.NP = ((p8(N) + 15) / 16) * 16;
    if (pl.NP > 256 || bit1 < 1 || bit1 > 8 || bit2 < 1 || bit2 > 8) return false;
    int cps = 256 / pl.NP;                       // the converters take at most kMaxTasks x 128 (chunk, column) tasks per item
    if (cps > kMaxCpsG) cps = kMaxCpsG;
    while (cps > 1 && cps * bit1 > 8) cps >>= 1; // staging: at most 8 pieces of 16 B per lane and item
    if (cps == 3) cps = 2;
    pl.cps = cps;
    pl.dbufs = (2 * pl.NP + 2 * cps * 32 <= 512) ? 2 : 1;
    for (int stages = kMaxStagesG; stages >= 2; --stages) {
        if (pl.dbufs * pl.NP + stages * cps * 32 > 512) continue;
        pl.stages = stages;
        pl.groups = stages < 2 ? 1 : 2;          // G <= S keeps the parity-only mbarrier waits unambiguous (see gemm_tc05.cu)
        pl.a_slot = bit1 * cps * 512;
        pl.b_slot = kMaxTasks * bit2 * kConvThreads * 16;
        pl.b_stage = cps * pl.NP * 128;
        int off = kExpWarps * kARing * pl.a_slot;
        off = (off + 1023) & ~1023;
        pl.off_bstage = off;
        off += stages * pl.b_stage;
        pl.off_bring = off;
        off += kBRing * pl.b_slot;
        pl.off_bars = off;
        off += (2 * kMaxStagesG + 4) * 8;
        pl.off_flags = off;
        off += kMaxStagesG * kMaxCpsG * 4;
        pl.off_misc = off;
        off += 16;
        pl.smem_total = off + 1024;              // slack for the manual 1024-byte alignment
        if (pl.smem_total <= 224 * 1024) return true;
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
};

bool grp_supports(int N, int bit1, int bit2, int epi) {
    if (epi != kEpiBitsRow && epi != kEpiFloat && epi != kEpiBitsCol) return false;
    GrpPlan pl;
    return plan_for(N, bit1, bit2, pl);
}

// One op over `count` problems.  Returns QGTC_ERR_UNSUPPORTED (before any work is enqueued) when the shape does not fit.
int launch_gemm_grouped(const GroupedProblemHost* probs, int count, int N, int bit1, int bit2, int out_bit, int epi, int pad128,
                        unsigned flags, cudaStream_t st) {
    GrpPlan pl;
    if (count <= 0) return QGTC_ERR_INVALID_ARG;
    if (!grp_supports(N, bit1, bit2, epi) || !plan_for(N, bit1, bit2, pl)) return QGTC_ERR_UNSUPPORTED;
    unsigned int* ef = tc05_error_flag_device();
    static thread_local GrpArgs T;     // 14 KB: filled per launch, passed by value as the kernel parameter
    for (int first = 0; first < count; first += kGrpMaxProblems) {
        const int n = count - first < kGrpMaxProblems ? count - first : kGrpMaxProblems;
        long long units = 0;
        for (int i = 0; i < n; ++i) {
            const GroupedProblemHost& h = probs[first + i];
            GrpProblem& p = T.p[i];
            p.x1 = h.x1; p.x2 = h.x2; p.out = h.out;
            p.bitmap = bit1 == 1 ? h.bitmap : nullptr;
            p.x1_words = h.x1_words; p.x2_words = h.x2_words;
            p.M = h.M; p.K = h.K;
            p.unit0 = (int)units;
            p.pad_ = 0;
            units += (p8(h.M) + 127) / 128;
            if (units > (1ll << 30)) return QGTC_ERR_UNSUPPORTED;
        }
        T.count = n;
        T.units = (int)units;
        T.N = N; T.NP = pl.NP; T.bit1 = bit1; T.bit2 = bit2; T.out_bit = out_bit; T.epi = epi; T.pad128 = pad128;
        T.cps = pl.cps; T.stages = pl.stages; T.dbufs = pl.dbufs; T.groups = pl.groups;
        T.a_slot_bytes = pl.a_slot; T.b_slot_bytes = pl.b_slot; T.b_stage_bytes = pl.b_stage;
        T.off_bstage = pl.off_bstage; T.off_bring = pl.off_bring; T.off_bars = pl.off_bars; T.off_flags = pl.off_flags; T.off_misc = pl.off_misc;
        T.cps_inv = pl.cps == 1 ? 0u : (unsigned int)((0x100000000ull + (unsigned long long)pl.cps - 1ull) / (unsigned long long)pl.cps);
        T.flags = flags;
        T.error_flag = ef;
        auto launch = [&](auto kfn) -> int {
            cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_total);
            if (e != cudaSuccess) return record_cuda_error(e, "cudaFuncSetAttribute(gemm_grp)");
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(T.units < kNumSMs ? T.units : kNumSMs);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = pl.smem_total;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous kernel's tail
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            static const bool no_pdl = [] { const char* e = getenv("QGTC_GRP_NO_PDL"); return e && e[0] == '1'; }();
            cfg.numAttrs = no_pdl ? 0 : 1;
            e = cudaLaunchKernelEx(&cfg, kfn, T);
            note_launch();
            if (e != cudaSuccess) return record_cuda_error(e, "cudaLaunchKernelEx(gemm_grp)");
            return check_launch("gemm_grp_kernel");
        };
        const int rc = bit1 == 1 ? launch(gemm_grp_kernel<true>) : launch(gemm_grp_kernel<false>);
        if (rc != QGTC_OK) return rc;
    }
    return QGTC_OK;
}

}  // namespace qgtc
