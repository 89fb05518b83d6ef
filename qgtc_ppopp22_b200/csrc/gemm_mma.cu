// Bit-GEMM engine 1: warp-level mma.sync.m16n8k32 (u8 x u8 -> s32) on re-assembled planes.
//
// Computes acc[m][n] = sum_{p<bit1} sum_{q<bit2} 2^(p+q) * popc(X1[p][m] & X2[q][n]) (reference
// kernel.h:292-342) as the value-level product sum_k a[m][k]*b[k][n] with a = sum_p 2^p X1[p],
// b = sum_q 2^q X2[q]: identical in wrapping int32 arithmetic, and its cost does not grow with
// bit1*bit2.  The packed planes are the only thing read from global memory; the 1-bit -> 8-bit
// expansion happens in registers with a K permutation chosen so that one shift and one mask
// yield a whole operand register (both operands use the same permutation, so the dot product is
// unchanged): byte i of register (slot s, quad lane t) is bit (t + 4s + 8i) of the 32-bit K word.
//
// This engine is the general one: any M, K, N, bit widths 1..8, raw-memory operands with
// zero-extension, all four epilogues.  It is also the "legacy tensor path" candidate of the
// engine measurement; the tcgen05 engine lives in gemm_tc05.cu.
#include "gemm_common.cuh"

namespace qgtc {

namespace {

constexpr int kStages = 3;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// bytes of a 16-byte piece starting at word idx that really exist in the operand
__device__ __forceinline__ int valid_bytes(unsigned long long idx, unsigned long long words) {
    if (idx >= words) return 0;
    const unsigned long long left = words - idx;
    return left >= 4 ? 16 : (int)left * 4;
}

template <int WARPS, int NT>
struct Smem {
    static constexpr int BM = 16 * WARPS, BN = 8 * NT;
    // per stage: packed A [bit1][BM] x 16 B, packed B [bit2][BN] x 16 B  (sized at launch)
    // then: Bs fragments [4 ksteps][NT][32 lanes][2] u32 ; the epilogue tile aliases the stages.
    static constexpr int kBsWords = 4 * NT * 32 * 2;
    static constexpr int kCsPitch = BN + 4;
};

template <int WARPS, int NT, int EPI>
__global__ void __launch_bounds__(WARPS * 32) gemm_mma_kernel(const GemmArgs P) {
    using S = Smem<WARPS, NT>;
    constexpr int BM = S::BM, BN = S::BN, NTHREADS = WARPS * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int bit1 = P.bit1, bit2 = P.bit2, gdk = P.gdk;
    const unsigned long long gdk4 = (unsigned long long)gdk * 4;

    const uint32_t* x1 = P.x1 + (unsigned long long)blockIdx.z * P.x1_bstride;
    const uint32_t* x2 = P.x2 + (unsigned long long)blockIdx.z * P.x2_bstride;

    const int stage_words = (bit1 * BM + bit2 * BN) * 4;
    uint32_t* stage_base = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* Bs = stage_base + kStages * stage_words;
    // the reference computes every row/column of its 8x8 fragments, i.e. up to P8(M)/P8(N)
    const int m_lim = p8(P.M), n_lim = p8(P.N);

    auto issue_chunk = [&](int c) {
        if (c < gdk) {
            uint32_t* st = stage_base + (c % kStages) * stage_words;
            const int a_items = bit1 * BM, items = a_items + bit2 * BN;
            for (int i = tid; i < items; i += NTHREADS) {
                const uint32_t* base;
                unsigned long long idx, words;
                bool ok;
                if (i < a_items) {
                    const int p = i / BM, r = i % BM, m = m0 + r;
                    ok = m < m_lim;
                    idx = (unsigned long long)p * P.act_offset + (unsigned long long)m * gdk4 + 4ull * c;
                    base = x1; words = P.x1_words;
                } else {
                    const int k = i - a_items, q = k / BN, j = k % BN, n = n0 + j;
                    ok = n < n_lim;
                    idx = (unsigned long long)q * P.w_offset + (unsigned long long)n * gdk4 + 4ull * c;
                    base = x2; words = P.x2_words;
                }
                const int nb = ok ? valid_bytes(idx, words) : 0;
                cp_async16_zfill(st + i * 4, nb ? (const void*)(base + idx) : (const void*)base, nb);
            }
        }
        cp_async_commit();
    };

    int acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0;

    issue_chunk(0);
    issue_chunk(1);
    const bool zero_jump = !(P.flags & QGTC_FLAG_NO_ZERO_JUMP);

    for (int c = 0; c < gdk; ++c) {
        cp_async_wait<kStages - 2>();
        __syncthreads();  // chunk c visible to all; everyone is done with chunk c-1 (its stage and Bs are free)
        issue_chunk(c + kStages - 1);

        const uint32_t* Apk = stage_base + (c % kStages) * stage_words;  // [bit1][BM][4]
        const uint32_t* Bpk = Apk + bit1 * BM * 4;                       // [bit2][BN][4]

        // ---- expand the B chunk once per CTA into mma fragment order -------------------------
        for (int e = tid; e < NT * 32; e += NTHREADS) {
            const int j = e >> 5, ln = e & 31, col = 8 * j + (ln >> 2), sh = ln & 3;
            uint32_t b0[4] = {0, 0, 0, 0}, b1[4] = {0, 0, 0, 0};
            for (int q = 0; q < bit2; ++q) {
                const uint4 w = *reinterpret_cast<const uint4*>(Bpk + (q * BN + col) * 4);
                const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    b0[s] |= ((ws[s] >> sh) & 0x01010101u) << q;
                    b1[s] |= ((ws[s] >> (sh + 4)) & 0x01010101u) << q;
                }
            }
#pragma unroll
            for (int s = 0; s < 4; ++s)
                *reinterpret_cast<uint2*>(Bs + ((s * NT + j) * 32 + ln) * 2) = make_uint2(b0[s], b1[s]);
        }
        __syncthreads();

        // ---- warp tile: 16 rows x BN columns, 4 k-steps of 32 ---------------------------------
        const uint32_t* Arow0 = Apk + (warp * 16 + g) * 4;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            uint32_t a[4] = {0, 0, 0, 0};
            for (int p = 0; p < bit1; ++p) {
                const uint32_t w0 = Arow0[p * BM * 4 + s];
                const uint32_t w8 = Arow0[p * BM * 4 + 32 + s];
                a[0] |= ((w0 >> t) & 0x01010101u) << p;
                a[1] |= ((w8 >> t) & 0x01010101u) << p;
                a[2] |= ((w0 >> (t + 4)) & 0x01010101u) << p;
                a[3] |= ((w8 >> (t + 4)) & 0x01010101u) << p;
            }
            // zero-tile jumping at 16-row x 32-bit granularity: an all-zero A fragment adds nothing
            if (zero_jump && !__any_sync(0xffffffffu, (a[0] | a[1] | a[2] | a[3]) != 0u)) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const uint2 b = *reinterpret_cast<const uint2*>(Bs + ((s * NT + j) * 32 + lane) * 2);
                mma_u8(acc[j], a, b.x, b.y);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // all stages idle: the epilogue tile may alias them

    const int M = P.M, N = P.N;
    if (EPI == kEpiFloat) {
        float* out = reinterpret_cast<float*>(P.out) + (unsigned long long)blockIdx.z * P.out_bstride;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + warp * 16 + g + 8 * h;
                const int n = n0 + 8 * j + 2 * t;
                if (m < M) {
                    if (n < N) out[(size_t)m * N + n] = (float)acc[j][2 * h];
                    if (n + 1 < N) out[(size_t)m * N + n + 1] = (float)acc[j][2 * h + 1];
                }
            }
        }
        return;
    }

    // ---- bit epilogues: requantize into a byte tile, then pack planes from shared memory -------
    unsigned char* Cs = smem_raw;  // [BM][kCsPitch]
    constexpr int PITCH = S::kCsPitch;
    const int ob = P.out_bit;
    const unsigned keep = (1u << ob) - 1u;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = warp * 16 + g + 8 * h, cc = 8 * j + 2 * t;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                unsigned v = (unsigned)requant(acc[j][2 * h + e], ob) & keep;
                if (EPI != kEpiBitsColQuirk && !(m0 + r < M && n0 + cc + e < N)) v = 0;  // kernel.h:367-368
                Cs[r * PITCH + cc + e] = (unsigned char)v;
            }
        }
    }
    __syncthreads();
    uint32_t* out = reinterpret_cast<uint32_t*>(P.out) + (unsigned long long)blockIdx.z * P.out_bstride;

    if (EPI == kEpiBitsRow) {
        const int wpr = s128(N) * 4;                                       // words per output row
        const unsigned long long opt_offset = (unsigned long long)p8(M) * wpr;  // kernel.h:267
        constexpr int WPB = (BN + 31) / 32;                                // words this block owns per row
        const int w_first = n0 >> 5;
        const bool last_block = blockIdx.y == gridDim.y - 1;
        const int w_cover = last_block ? wpr - w_first : WPB;              // the last block also writes the zero pad
        const int items = ob * BM * w_cover;
        for (int i = tid; i < items; i += NTHREADS) {
            const int wi = i % w_cover, r = (i / w_cover) % BM, p = i / (w_cover * BM);
            const int m = m0 + r;
            if (m >= p8(M)) continue;
            uint32_t word = 0;
            if (wi < WPB) {
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const int cc = wi * 32 + q4 * 4;
                    if (cc < BN) {
                        const uint32_t four = *reinterpret_cast<const uint32_t*>(Cs + r * PITCH + cc);
                        word |= nibble_of_plane(four, p) << (28 - 4 * q4);
                    }
                }
            }
            out[(unsigned long long)p * opt_offset + (unsigned long long)m * wpr + w_first + wi] = word;
        }
    } else if (EPI == kEpiBitsCol) {
        // output pre-zeroed by the launcher; word (p, n, mw) holds rows 32mw..32mw+31 of column n
        const unsigned long long kw = (unsigned long long)s128(M) * 4;
        const unsigned long long plane = (unsigned long long)s128(M) * p128(N) * 4;
        constexpr int RG = BM / 32;
        const int items = ob * BN * RG;
        for (int i = tid; i < items; i += NTHREADS) {
            const int v = i % RG, j = (i / RG) % BN, p = i / (RG * BN);
            const int n = n0 + j;
            if (n >= N || m0 + 32 * v >= M) continue;
            uint32_t word = 0;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) word |= (uint32_t)((Cs[(32 * v + r) * PITCH + j] >> p) & 1u) << (31 - r);
            out[(unsigned long long)p * plane + (unsigned long long)n * kw + (m0 >> 5) + v] = word;
        }
    } else {  // kEpiBitsColQuirk: byte-exact image of what QGTC_layer_hidden_col stores (kernel.h:764-806)
        unsigned char* outb = reinterpret_cast<unsigned char*>(out);
        const unsigned long long gdh = (unsigned long long)s128(M);
        const unsigned long long plane_bytes = gdh * p128(N) * 16;
        constexpr int TR = BM / 8, TC = BN / 8;
        const int items = ob * TR * TC * 8;  // (plane, tile row, tile col, half*4 + j)
        for (int i = tid; i < items; i += NTHREADS) {
            const int hj = i & 7, half = hj >> 2, j = hj & 3;
            const int tc = (i >> 3) % TC, tr = ((i >> 3) / TC) % TR, p = (i >> 3) / (TC * TR);
            const int bx = (m0 >> 3) + tr, by = (n0 >> 3) + tc;
            if (bx >= s8(M) || by >= s8(N)) continue;
            unsigned byte = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const int ln = 8 * j + b, gy = ln & 7, gx = ln >> 3;
                const int idx = (half ? 4 : 0) + ln;         // the defect: +4, not +32 (kernel.h:782)
                const int col = idx >> 3, row = idx & 7;     // Cs was stored col-major (kernel.h:764)
                const bool inb = (by * 8 + gy < N) && (bx * 8 + gx + 4 * half < M);  // transposed test, :777-778
                const unsigned bit = inb ? ((Cs[(tr * 8 + row) * PITCH + tc * 8 + col] >> p) & 1u) : 0u;
                byte |= bit << (7 - b);
            }
            outb[(unsigned long long)p * plane_bytes + (unsigned long long)(by * 8 + 4 * half + j) * gdh * 16 + (bx ^ 3)] = (unsigned char)byte;
        }
    }
}

template <int WARPS, int NT, int EPI>
int launch_one(const GemmArgs& a, cudaStream_t st) {
    using S = Smem<WARPS, NT>;
    const int stage_words = (a.bit1 * S::BM + a.bit2 * S::BN) * 4;
    size_t smem = (size_t)(kStages * stage_words + S::kBsWords) * 4;
    const size_t cs = (size_t)S::BM * S::kCsPitch;
    if (smem < cs) smem = cs;
    auto kfn = gemm_mma_kernel<WARPS, NT, EPI>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return record_cuda_error(e, "cudaFuncSetAttribute(gemm_mma)");
    }
    dim3 grid((p8(a.M) + S::BM - 1) / S::BM, (a.N + S::BN - 1) / S::BN, a.count);
    kfn<<<grid, WARPS * 32, smem, st>>>(a);
    note_launch();
    return check_launch("gemm_mma_kernel");
}

template <int EPI>
int launch_epi(const GemmArgs& a, cudaStream_t st) {
    // column tile: 16/32/64 wide; row tile: shrink while the grid cannot fill the SMs
    const int nt = a.N <= 16 ? 2 : (a.N <= 32 ? 4 : 8);
    const long long nblk = (long long)((a.N + 8 * nt - 1) / (8 * nt)) * a.count;
    const long long rows = p8(a.M);
    int warps = 8;
    if (((rows + 127) / 128) * nblk < 2 * kNumSMs) warps = 4;
    if (((rows + 63) / 64) * nblk < 2 * kNumSMs) warps = 2;
#define QGTC_CASE(W, T) if (warps == W && nt == T) return launch_one<W, T, EPI>(a, st);
    QGTC_CASE(8, 8) QGTC_CASE(8, 4) QGTC_CASE(8, 2)
    QGTC_CASE(4, 8) QGTC_CASE(4, 4) QGTC_CASE(4, 2)
    QGTC_CASE(2, 8) QGTC_CASE(2, 4) QGTC_CASE(2, 2)
#undef QGTC_CASE
    return QGTC_ERR_INVALID_ARG;
}

}  // namespace

int launch_gemm_mma(const GemmArgs& a, int epi, cudaStream_t st) {
    switch (epi) {
        case kEpiBitsRow: return launch_epi<kEpiBitsRow>(a, st);
        case kEpiFloat: return launch_epi<kEpiFloat>(a, st);
        case kEpiBitsCol: return launch_epi<kEpiBitsCol>(a, st);
        case kEpiBitsColQuirk: return launch_epi<kEpiBitsColQuirk>(a, st);
    }
    return QGTC_ERR_INVALID_ARG;
}

}  // namespace qgtc
