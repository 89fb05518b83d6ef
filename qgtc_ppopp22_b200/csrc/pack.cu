// Quantize-and-pack (val2bit), unpack (bit2val) and tile counters.
//
// Replaces Quantize_val + QGTC_layer_input / PackFcWeight128 (reference kernel.h:49-106,204-242)
// and UnPackFc* (kernel.h:109-201).  One fused pass: fp32 is read once with 128-bit loads, the
// int32 temporary of the reference (QGTC_device.cu:63) does not exist, every output word is
// written (no pre-zeroed output needed).  HBM-bound: 4 B/element read, nbits/8 B/element written.
#include <cstdlib>

#include "common.cuh"

namespace qgtc {

// ------------------------------------------------------------------------------------------
// ROW layout, shuffle variant (small or oddly aligned inputs, one or two planes): a warp handles (row, 128-column group)
// units: lane l owns columns 4l..4l+3 (one float4, or four scalars when the row is not 16-byte aligned), turns them into
// one nibble per plane, and an 8-lane OR-butterfly assembles the four words.  Finer work units than the word-per-lane
// kernel below (more parallelism for small matrices) and coalesced for any alignment.
// ------------------------------------------------------------------------------------------
constexpr int kPackUnroll = 4;

// Work decomposition without divisions in the loop: `splits` warps share a row (each takes a contiguous
// range of 128-column groups), rows are strided over the remaining warps.
template <bool kVec>
__global__ void __launch_bounds__(256) pack_row_shfl_kernel(const float* __restrict__ x, int H, int W, int nbits, int splits,
                                                             uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int gdy = s128(W);
    const int rows = p8(H);
    const size_t wpr = (size_t)gdy * 4, plane = (size_t)rows * wpr;
    const float ub = (float)(1 << nbits);
    const uint32_t keep4 = ((1u << nbits) - 1u) * 0x01010101u;
    const unsigned warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned part = warp_id % (unsigned)splits, row0 = warp_id / (unsigned)splits, row_step = nwarps / (unsigned)splits;
    if (row_step == 0 || row0 >= row_step) return;   // leftover warps when nwarps is not a multiple of splits
    const int per = (gdy + splits - 1) / splits;
    const int cg_begin = (int)part * per, cg_end = min(gdy, cg_begin + per);
    const int p_of_lane = lane & 7, w_of_lane = lane >> 3;

    for (unsigned r = row0; r < (unsigned)rows; r += row_step) {
        const bool row_ok = r < (unsigned)H;
        const float* xrow = x + (size_t)r * W;
        uint32_t* orow = out + (size_t)r * wpr;
        for (int cg0 = cg_begin; cg0 < cg_end; cg0 += kPackUnroll) {
            float4 v[kPackUnroll];
#pragma unroll
            for (int j = 0; j < kPackUnroll; ++j) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = (cg0 + j) * 128 + 4 * lane;
                if (row_ok && cg0 + j < cg_end && c < W) {
                    const float* src = xrow + c;
                    if (kVec) {
                        v[j] = __ldcs(reinterpret_cast<const float4*>(src));  // streamed once
                    } else {
                        v[j].x = __ldcs(src);
                        if (c + 1 < W) v[j].y = __ldcs(src + 1);
                        if (c + 2 < W) v[j].z = __ldcs(src + 2);
                        if (c + 3 < W) v[j].w = __ldcs(src + 3);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kPackUnroll; ++j) {
                if (cg0 + j >= cg_end) break;  // warp-uniform
                const uint32_t a0 = quantize_bits(v[j].x, ub), a1 = quantize_bits(v[j].y, ub);
                const uint32_t a2 = quantize_bits(v[j].z, ub), a3 = quantize_bits(v[j].w, ub);
                const uint32_t t = __byte_perm(__byte_perm(a3, a2, 0x0040u), __byte_perm(a1, a0, 0x0040u), 0x5410u) & keep4;
                unsigned mine = 0;
                for (int p = 0; p < nbits; ++p) {
                    // bit p of the four values -> nibble, first column in the nibble's MSB
                    const unsigned nib = (((t >> p) & 0x01010101u) * 0x01020408u) >> 24;
                    unsigned w = nib << (28 - 4 * p_of_lane);
                    w |= __shfl_xor_sync(0xffffffffu, w, 1);
                    w |= __shfl_xor_sync(0xffffffffu, w, 2);
                    w |= __shfl_xor_sync(0xffffffffu, w, 4);
                    if (p_of_lane == p) mine = w;
                }
                if (p_of_lane < nbits) orow[(size_t)p_of_lane * plane + (size_t)(cg0 + j) * 4 + w_of_lane] = mine;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Batched ROW pack: one launch packs a whole list of matrices (the 75 dense batch adjacencies of a Cluster-GCN epoch are
// 759 x 759 each: one launch per matrix is 4 us of launch latency for 2.3 MB).  Work unit = (matrix, row, 128-column
// group), units of all matrices in one index space; the matrix of a unit is found by a binary search over the unit
// prefix sums in the kernel parameters (once per unit = 512 bytes of input).  Rows of odd width are only 4-byte aligned:
// a lane then reads the two aligned 16-byte pieces that cover its four columns and picks them apart with the row's
// misalignment (0..3 floats, warp-uniform per row), so every request is still a full, coalesced 512-byte warp access.
// ------------------------------------------------------------------------------------------
constexpr int kPackBatchMax = 448;      // problems per launch (28 bytes each: a 12.5 KB kernel parameter); longer lists are split
struct PackProblem {
    const float* in;
    uint32_t* out;
    int H, W;
    unsigned int unit0;                 // first unit of this matrix
};
struct PackBatchArgs {
    int count, nbits;
    unsigned int units;
    PackProblem p[kPackBatchMax];
};

__global__ void __launch_bounds__(256) pack_row_batched_kernel(const __grid_constant__ PackBatchArgs A) {
    const int lane = threadIdx.x & 31;
    const int nbits = A.nbits;
    const float ub = (float)(1 << nbits);
    const uint32_t keep4 = ((1u << nbits) - 1u) * 0x01010101u;
    const unsigned warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    const int p_of_lane = lane & 7, w_of_lane = lane >> 3;
    int z = 0;
    for (unsigned u = warp_id; u < A.units; u += nwarps) {
        if (!(u >= A.p[z].unit0 && (z + 1 == A.count || u < A.p[z + 1].unit0))) {   // usually the same matrix as the last unit
            int lo = 0, hi = A.count - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (A.p[mid].unit0 <= u) lo = mid; else hi = mid - 1;
            }
            z = lo;
        }
        const PackProblem& P = A.p[z];
        const int H = P.H, W = P.W;
        const int gdy = s128(W);
        const size_t wpr = (size_t)gdy * 4, plane = (size_t)p8(H) * wpr;
        const unsigned lu = u - P.unit0;
        const int r = (int)(lu / (unsigned)gdy), cg = (int)(lu - (unsigned)r * (unsigned)gdy);
        const int c = cg * 128 + 4 * lane;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < H && c < W) {
            const float* src = P.in + (size_t)r * W + c;
            const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(src) >> 2) & 3u);   // floats past a 16-byte boundary (warp-uniform)
            if (mis == 0 && c + 4 <= W) {
                v = __ldcs(reinterpret_cast<const float4*>(src));
            } else if (c + 4 <= W && c + 8 <= W + (int)mis) {
                // two aligned pieces cover the four columns; both stay inside the row's own storage except possibly the very
                // first / last piece of the whole matrix, which is excluded by the conditions around this branch
                const float4* a = reinterpret_cast<const float4*>(src - mis);
                const bool first_ok = (r > 0 || c >= (int)mis);          // piece a[0] starts before the matrix only for row 0, column 0
                if (first_ok) {
                    const float4 lo4 = __ldcs(a), hi4 = __ldcs(a + 1);
                    const float f[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                    v.x = mis == 1 ? f[1] : (mis == 2 ? f[2] : f[3]);
                    v.y = mis == 1 ? f[2] : (mis == 2 ? f[3] : f[4]);
                    v.z = mis == 1 ? f[3] : (mis == 2 ? f[4] : f[5]);
                    v.w = mis == 1 ? f[4] : (mis == 2 ? f[5] : f[6]);
                } else {
                    v.x = __ldcs(src); v.y = __ldcs(src + 1); v.z = __ldcs(src + 2); v.w = __ldcs(src + 3);
                }
            } else {
                v.x = __ldcs(src);
                if (c + 1 < W) v.y = __ldcs(src + 1);
                if (c + 2 < W) v.z = __ldcs(src + 2);
                if (c + 3 < W) v.w = __ldcs(src + 3);
            }
        }
        const uint32_t a0 = quantize_bits(v.x, ub), a1 = quantize_bits(v.y, ub);
        const uint32_t a2 = quantize_bits(v.z, ub), a3 = quantize_bits(v.w, ub);
        const uint32_t t = __byte_perm(__byte_perm(a3, a2, 0x0040u), __byte_perm(a1, a0, 0x0040u), 0x5410u) & keep4;
        unsigned mine = 0;
        for (int p = 0; p < nbits; ++p) {
            const unsigned nib = (((t >> p) & 0x01010101u) * 0x01020408u) >> 24;
            unsigned w = nib << (28 - 4 * p_of_lane);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if (p_of_lane == p) mine = w;
        }
        if (p_of_lane < nbits) P.out[(size_t)p_of_lane * plane + (size_t)r * wpr + (size_t)cg * 4 + w_of_lane] = mine;
    }
}

// Batched ROW pack, word-per-lane form (rows of at least 16 output words): a lane owns one output word = 32 consecutive
// columns = 128 bytes of the fp32 row, as in pack_row_kernel below; a unit is (matrix, row, group of 32 words).  Rows of odd
// width are only 4-byte aligned: the lane reads the nine aligned 16-byte pieces that cover its 32 floats and the register
// selection is fixed at compile time per misalignment (0..3 floats, the same for every lane of a row).  ~150 instructions per
// 4 KB of input instead of ~100 per 512 bytes in the shuffle form above.
template <int MIS>
__device__ __forceinline__ void quantize_from_pieces(const float4 (&a)[9], float ub, uint32_t keep4, uint32_t (&t)[8]) {
    const float f[36] = {a[0].x, a[0].y, a[0].z, a[0].w, a[1].x, a[1].y, a[1].z, a[1].w, a[2].x, a[2].y, a[2].z, a[2].w,
                         a[3].x, a[3].y, a[3].z, a[3].w, a[4].x, a[4].y, a[4].z, a[4].w, a[5].x, a[5].y, a[5].z, a[5].w,
                         a[6].x, a[6].y, a[6].z, a[6].w, a[7].x, a[7].y, a[7].z, a[7].w, a[8].x, a[8].y, a[8].z, a[8].w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t a0 = quantize_bits(f[MIS + 4 * k], ub), a1 = quantize_bits(f[MIS + 4 * k + 1], ub);
        const uint32_t a2 = quantize_bits(f[MIS + 4 * k + 2], ub), a3 = quantize_bits(f[MIS + 4 * k + 3], ub);
        t[k] = __byte_perm(__byte_perm(a3, a2, 0x0040u), __byte_perm(a1, a0, 0x0040u), 0x5410u) & keep4;
    }
}

template <int NB>
__global__ void __launch_bounds__(256) pack_row_batched_word_kernel(const __grid_constant__ PackBatchArgs A) {
    const int lane = threadIdx.x & 31;
    const float ub = (float)(1 << NB);
    const uint32_t keep4 = ((1u << NB) - 1u) * 0x01010101u;
    const unsigned warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int kU = 1;                 // units per iteration (2: measured slower -- 125 registers halve the occupancy)
    int z = 0;
    for (unsigned u0 = warp_id * kU; u0 < A.units; u0 += nwarps * kU) {
        float4 a[kU][9];
        const float* src[kU];
        uint32_t* dst[kU];
        size_t plane[kU];
        int col0[kU], W_[kU];
        unsigned mis[kU];
        int mode[kU];                     // 0: nothing to do, 1: zeros (padding row / columns), 2: aligned pieces, 3: element-wise tail
#pragma unroll
        for (int i = 0; i < kU; ++i) {
            const unsigned u = u0 + i;
            mode[i] = 0;
            if (u >= A.units) continue;
            if (!(u >= A.p[z].unit0 && (z + 1 == A.count || u < A.p[z + 1].unit0))) {
                int lo = 0, hi = A.count - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (A.p[mid].unit0 <= u) lo = mid; else hi = mid - 1;
                }
                z = lo;
            }
            const PackProblem& P = A.p[z];
            const int H = P.H, W = P.W;
            const int wpr = s128(W) * 4, groups = (wpr + 31) >> 5;
            const unsigned lu = u - P.unit0;
            const int r = (int)(lu / (unsigned)groups), g = (int)(lu - (unsigned)r * (unsigned)groups);
            const int wi = g * 32 + lane;
            if (wi >= wpr) continue;
            plane[i] = (size_t)p8(H) * wpr;
            dst[i] = P.out + (size_t)r * wpr + wi;
            col0[i] = wi * 32;
            W_[i] = W;
            mode[i] = 1;
            if (r < H && col0[i] < W) {
                src[i] = P.in + (size_t)r * W + col0[i];
                mis[i] = (unsigned)((reinterpret_cast<uintptr_t>(src[i]) >> 2) & 3u);
                // all nine aligned pieces lie inside the matrix: not before its first element, not past its last
                const bool whole = col0[i] + 32 <= W && (mis[i] == 0 || ((r > 0 || col0[i] > 0) && (col0[i] + 36 <= W || r + 1 < H)));
                mode[i] = whole ? 2 : 3;
                if (whole) {
                    const float4* ap = reinterpret_cast<const float4*>(src[i] - mis[i]);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(a[i][k].x), "=f"(a[i][k].y), "=f"(a[i][k].z), "=f"(a[i][k].w) : "l"(ap + k));
                    a[i][8] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (mis[i]) asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(a[i][8].x), "=f"(a[i][8].y), "=f"(a[i][8].z), "=f"(a[i][8].w) : "l"(ap + 8));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kU; ++i) {
            if (mode[i] == 0) continue;
            uint32_t t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = 0u;
            if (mode[i] == 2) {
                switch (mis[i]) {      // the same for every lane of a row
                    case 0: quantize_from_pieces<0>(a[i], ub, keep4, t); break;
                    case 1: quantize_from_pieces<1>(a[i], ub, keep4, t); break;
                    case 2: quantize_from_pieces<2>(a[i], ub, keep4, t); break;
                    default: quantize_from_pieces<3>(a[i], ub, keep4, t); break;
                }
            } else if (mode[i] == 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t q[4] = {0x4B400000u, 0x4B400000u, 0x4B400000u, 0x4B400000u};   // quantize_bits(0)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (col0[i] + 4 * k + e < W_[i]) q[e] = quantize_bits(__ldg(src[i] + 4 * k + e), ub);
                    t[k] = __byte_perm(__byte_perm(q[3], q[2], 0x0040u), __byte_perm(q[1], q[0], 0x0040u), 0x5410u) & keep4;
                }
            }
            uint32_t w[NB];
            planes_from_bytes<NB>(t, w);
#pragma unroll
            for (int p = 0; p < NB; ++p) dst[i][(size_t)p * plane[i]] = w[p];
        }
    }
}

// ------------------------------------------------------------------------------------------
// ROW layout: word [p][r][w] holds columns 32w..32w+31 of row r, MSB-first.
// A lane owns one output word = 32 consecutive columns = one 128-byte line of the fp32 row (eight 16-byte loads, all in
// flight together); a warp covers 1024 columns of a row per unit and stores 32 consecutive words per plane (one coalesced
// 128-byte store per plane).  No shuffles, no float->int conversions: the quantizer rounds with a magic-number add
// (quantize_bits), the plane words come from multiply-gathers (up to 3 planes) or 8 x 8 bit transposes (4 planes and up).
// ------------------------------------------------------------------------------------------
template <int NB, bool kVec>
__global__ void __launch_bounds__(256) pack_row_kernel(const float* __restrict__ x, int H, int W, uint32_t* __restrict__ out, int vec8) {
    const int lane = threadIdx.x & 31;
    const int rows = p8(H);
    const int wpr = s128(W) * 4;                       // words per output row
    const int groups = (wpr + 31) >> 5;                // 32-word groups per row
    const size_t plane = (size_t)rows * wpr;
    const float ub = (float)(1 << NB);
    const uint32_t keep4 = ((1u << NB) - 1u) * 0x01010101u;
    const long long units = (long long)rows * groups;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long u = warp_id; u < units; u += nwarps) {
        const int r = (int)(u / groups), g = (int)(u - (long long)r * groups);
        const int wi = g * 32 + lane;
        if (wi >= wpr) continue;
        const int col0 = wi * 32;
        uint32_t t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = 0u;
        if (r < H && col0 < W) {
            const float* src = x + (size_t)r * W + col0;
            if (kVec && col0 + 32 <= W) {
                float4 v[8];
                if (vec8) {     // 32-byte aligned rows: four 256-bit loads, one whole sector each (LDG.E.256)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                                     : "=f"(v[2 * k].x), "=f"(v[2 * k].y), "=f"(v[2 * k].z), "=f"(v[2 * k].w), "=f"(v[2 * k + 1].x),
                                       "=f"(v[2 * k + 1].y), "=f"(v[2 * k + 1].z), "=f"(v[2 * k + 1].w)
                                     : "l"(src + 8 * k));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                                     : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w)
                                     : "l"(reinterpret_cast<const float4*>(src) + k));
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t a0 = quantize_bits(v[k].x, ub), a1 = quantize_bits(v[k].y, ub);
                    const uint32_t a2 = quantize_bits(v[k].z, ub), a3 = quantize_bits(v[k].w, ub);
                    // low bytes of a0..a3, first element in the top byte
                    t[k] = __byte_perm(__byte_perm(a3, a2, 0x0040u), __byte_perm(a1, a0, 0x0040u), 0x5410u) & keep4;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t a[4] = {0x4B400000u, 0x4B400000u, 0x4B400000u, 0x4B400000u};   // quantize_bits(0)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (col0 + 4 * k + i < W) a[i] = quantize_bits(__ldg(src + 4 * k + i), ub);
                    t[k] = __byte_perm(__byte_perm(a[3], a[2], 0x0040u), __byte_perm(a[1], a[0], 0x0040u), 0x5410u) & keep4;
                }
            }
        }
        uint32_t w[NB];
        planes_from_bytes<NB>(t, w);
        uint32_t* dst = out + (size_t)r * wpr + wi;
#pragma unroll
        for (int p = 0; p < NB; ++p) dst[(size_t)p * plane] = w[p];
    }
}

// ------------------------------------------------------------------------------------------
// COL layouts: word [p][c][kw] holds rows 32kw..32kw+31 of column c, MSB-first (a transposing
// pack).  One warp handles a 32-row x 32-column tile: every row read is a coalesced 128-byte
// request (all 32 in flight), each lane collects the 32 quantized values of its own column (four per
// register) and extracts the plane words like the row kernel.
// ------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) pack_col_kernel(const float* __restrict__ x, int H, int W, int col_limit,
                                                        uint32_t* __restrict__ out, unsigned long long out_words) {
    const int lane = threadIdx.x & 31;
    const int gdk = s128(H);
    const int kwords = gdk * 4;                       // words per column
    const int cgroups = (col_limit + 31) >> 5;
    const unsigned units = (unsigned)kwords * (unsigned)cgroups;
    const size_t plane = (size_t)gdk * p128(W) * 4;   // reference kernel.h:88: always the P128 stride
    const float ub = (float)(1 << NB);
    const uint32_t keep4 = ((1u << NB) - 1u) * 0x01010101u;
    const unsigned warp_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;

    for (unsigned u = warp_id; u < units; u += nwarps) {
        const int kw = (int)(u % (unsigned)kwords), c = (int)(u / (unsigned)kwords) * 32 + lane;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int r = kw * 32 + i;
            v[i] = (r < H && c < W) ? __ldg(x + (size_t)r * W + c) : 0.f;
        }
        uint32_t t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t a0 = quantize_bits(v[4 * k], ub), a1 = quantize_bits(v[4 * k + 1], ub);
            const uint32_t a2 = quantize_bits(v[4 * k + 2], ub), a3 = quantize_bits(v[4 * k + 3], ub);
            t[k] = __byte_perm(__byte_perm(a3, a2, 0x0040u), __byte_perm(a1, a0, 0x0040u), 0x5410u) & keep4;
        }
        uint32_t w[NB];
        planes_from_bytes<NB>(t, w);
        if (c < col_limit) {
#pragma unroll
            for (int p = 0; p < NB; ++p) {
                const unsigned long long idx = (unsigned long long)p * plane + (unsigned long long)c * kwords + (unsigned long long)kw;
                if (idx < out_words) out[idx] = w[p];  // C8 with P8(W) != P128(W): the reference writes out of bounds here; dropped
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// bit2val: one thread per output element (debug / test op in the reference).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_kernel(Words in, int nbits, int H, int W, int col_major, int* __restrict__ out) {
    const long long n = (long long)H * W;
    const size_t wpr = (size_t)s128(W) * 4, plane_r = (size_t)p8(H) * wpr;
    const size_t kw = (size_t)s128(H) * 4, plane_c = (size_t)s128(H) * p128(W) * 4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / W), c = (int)(e % W);
        int v = 0;
        for (int p = 0; p < nbits; ++p) {
            uint32_t wd;
            int bit;
            if (!col_major) { wd = ldw(in, p * plane_r + r * wpr + (c >> 5)); bit = 31 - (c & 31); }
            else            { wd = ldw(in, p * plane_c + c * kw + (r >> 5));  bit = 31 - (r & 31); }
            v += (int)((wd >> bit) & 1u) << p;
        }
        out[e] = v;
    }
}

// ------------------------------------------------------------------------------------------
// Non-zero 8-row x 128-bit tiles of the A operand (zero-tile statistics, reference kernel.h:574-591).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_nonzero_tiles_kernel(Words x1, int M, int K, int bit1, unsigned long long* __restrict__ nz) {
    const int gdx = s8(M), gdk = s128(K);
    const long long tiles = (long long)bit1 * gdx * gdk;
    const size_t act_offset = (size_t)p8(M) * gdk * 4;
    unsigned long long local = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < tiles; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % gdk);
        const int bx = (int)((t / gdk) % gdx);
        const int p = (int)(t / ((long long)gdk * gdx));
        uint32_t any = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint4 v = ldw4(x1, p * act_offset + (size_t)(bx * 8 + r) * gdk * 4 + (size_t)i * 4);
            any |= v.x | v.y | v.z | v.w;
        }
        local += any != 0u;
    }
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(nz, local);
}

// ------------------------------------------------------------------------------------------
// Tile bitmap for zero-tile jumping in the tcgen05 engine: bit c of word w of tile (z, mt) is set iff the
// 128-row x 128-bit block (rows 128mt.., K chunk 32w + c) of plane 0 of problem z has any bit set.
// One warp per (z, mt, w, row group of 8): lane = chunk, 8 coalesced row reads (independent loads), one ballot, one atomicOr
// into the (pre-zeroed) bitmap word.  (The first version walked all 128 rows of a tile in one warp: 57 us for a 759-node batch,
// pure load latency.)
// ------------------------------------------------------------------------------------------
constexpr int kBmRows = 8;       // rows per warp
__global__ void __launch_bounds__(256) tile_bitmap_kernel(const uint32_t* __restrict__ x1, unsigned long long x1_words,
                                                           unsigned long long bstride, int M, int K, int count,
                                                           uint32_t* __restrict__ bitmap) {
    const int lane = threadIdx.x & 31;
    const int gdk = s128(K), mtiles = (p8(M) + 127) / 128, bmw = (gdk + 31) / 32;
    constexpr int kGroups = 128 / kBmRows;
    const long long units = (long long)count * mtiles * bmw * kGroups;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const unsigned long long gdk4 = (unsigned long long)gdk * 4;
    for (long long uu = warp_id; uu < units; uu += nwarps) {
        const int rg = (int)(uu % kGroups);
        const long long u = uu / kGroups;                 // bitmap word index: (z, mt, w)
        const int w = (int)(u % bmw), mt = (int)((u / bmw) % mtiles), z = (int)(u / ((long long)bmw * mtiles));
        const Words a{x1 + (unsigned long long)z * bstride, x1_words};
        const int c = 32 * w + lane;
        uint32_t any = 0;
        if (c < gdk) {
            const int r0 = mt * 128 + rg * kBmRows;
#pragma unroll
            for (int r = 0; r < kBmRows; ++r) {
                if (r0 + r < p8(M)) {
                    const uint4 v = ldw4(a, (unsigned long long)(r0 + r) * gdk4 + 4ull * c);
                    any |= v.x | v.y | v.z | v.w;
                }
            }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, any != 0u);
        if (lane == 0 && word) atomicOr(bitmap + u, word);
    }
}

// ------------------------------------------------------------------------------------------
// Batch construction straight from an edge list: packed 1-bit adjacency R(1, n, n) without the dense fp32
// n x n intermediate the reference materialises (sampler.py:87-89).  One thread per edge, atomicOr of a bit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edges_to_bits_kernel(const int* __restrict__ rows, const int* __restrict__ cols, long long n_edges,
                                                             int H, int W, uint32_t* __restrict__ out) {
    const size_t wpr = (size_t)s128(W) * 4;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x) {
        const int r = __ldg(rows + e), c = __ldg(cols + e);
        if (r >= 0 && r < H && c >= 0 && c < W) atomicOr(out + (size_t)r * wpr + (c >> 5), 1u << (31 - (c & 31)));
    }
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
static int grid_for(long long work_items_per_block_units, int blocks_per_sm) {
    long long g = work_items_per_block_units;
    const long long cap = (long long)kNumSMs * blocks_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_val2bit_row_batched(const float* const* ins, const int* Hs, const int* Ws, uint32_t* const* outs, int count, int nbits, cudaStream_t st);

int launch_val2bit(const float* in, int H, int W, int nbits, int layout, uint32_t* out, size_t out_words, cudaStream_t st) {
    if (layout == QGTC_LAYOUT_ROW) {
        const int rows = p8(H), wpr = s128(W) * 4, gdy = s128(W);
        const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
        // Two kernels, chosen by measurement (profiles/r2_pack_bench_*.txt): the word-per-lane kernel (256-bit loads, no shuffles,
        // bit transposes) wins on large 16-byte-aligned inputs at every width (4096^2: 0.64-0.67 of HBM vs 0.38-0.59; 8192^2: 0.79-0.80);
        // the shuffle kernel has finer work units (small matrices) and is coalesced for any alignment.  QGTC_PACK_ROW=word|shfl forces one.
        static const int force = [] { const char* e = getenv("QGTC_PACK_ROW"); return !e ? 0 : (e[0] == 'w' ? 1 : (e[0] == 's' ? 2 : 0)); }();
        const long long word_units = (long long)rows * ((wpr + 31) / 32);     // warp-level units of the word kernel: (row, 32 output words)
        bool use_word = vec && W >= 1024 && word_units >= 2ll * kNumSMs * 8;
        if (force == 1) use_word = vec;
        if (force == 2) use_word = false;
        // (a single large matrix with unaligned rows stays on the shuffle kernel: routed through the batched word kernel, 2258^2 measured
        // 0.27 of HBM at 1-2 bits against 0.34-0.36 -- too few warp-level units for one matrix; profiles/r2_pack_bench_call48.txt)
        if (use_word) {
            const int grid = grid_for((word_units + 7) / 8, 8);
            const int vec8 = (W % 8 == 0) && ((reinterpret_cast<uintptr_t>(in) & 31u) == 0) ? 1 : 0;
#define QGTC_PACK_ROW(NB) \
    case NB: pack_row_kernel<NB, true><<<grid, 256, 0, st>>>(in, H, W, out, vec8); break;
            switch (nbits) {
                QGTC_PACK_ROW(1) QGTC_PACK_ROW(2) QGTC_PACK_ROW(3) QGTC_PACK_ROW(4)
                QGTC_PACK_ROW(5) QGTC_PACK_ROW(6) QGTC_PACK_ROW(7) QGTC_PACK_ROW(8)
                default: return QGTC_ERR_UNSUPPORTED;
            }
#undef QGTC_PACK_ROW
            note_launch();
            return check_launch("pack_row_kernel");
        }
        const long long units = (long long)rows * gdy;
        const long long blocks = (units + 8 * kPackUnroll - 1) / (8 * kPackUnroll);
        const int grid = grid_for(blocks, 8);
        // warps sharing a row: only when there are fewer rows than warps
        const long long nwarps = (long long)grid * 8;
        int splits = 1;
        if (rows < nwarps) {
            splits = (int)(nwarps / rows);
            const int max_splits = (gdy + kPackUnroll - 1) / kPackUnroll;
            if (splits > max_splits) splits = max_splits;
            if (splits < 1) splits = 1;
        }
        if (vec) pack_row_shfl_kernel<true><<<grid, 256, 0, st>>>(in, H, W, nbits, splits, out);
        else pack_row_shfl_kernel<false><<<grid, 256, 0, st>>>(in, H, W, nbits, splits, out);
        note_launch();
        return check_launch("pack_row_shfl_kernel");
    }
    const int col_limit = layout == QGTC_LAYOUT_COL128 ? p128(W) : p8(W);
    if (layout == QGTC_LAYOUT_COL8 && nbits > 1 && p8(W) != p128(W)) {
        // words between the C8 allocation's planes are never addressed by the C128 index formula
        cudaError_t e = cudaMemsetAsync(out, 0, out_words * sizeof(uint32_t), st);
        if (e != cudaSuccess) return record_cuda_error(e, "cudaMemsetAsync(val2bit C8)");
    }
    const long long units = (long long)s128(H) * 4 * ((col_limit + 31) >> 5);
    if (units >= (1ll << 31)) return QGTC_ERR_UNSUPPORTED;
    const int grid = grid_for((units + 7) / 8, 8);
#define QGTC_PACK_COL(NB) \
    case NB: pack_col_kernel<NB><<<grid, 256, 0, st>>>(in, H, W, col_limit, out, (unsigned long long)out_words); break;
    switch (nbits) {
        QGTC_PACK_COL(1) QGTC_PACK_COL(2) QGTC_PACK_COL(3) QGTC_PACK_COL(4)
        QGTC_PACK_COL(5) QGTC_PACK_COL(6) QGTC_PACK_COL(7) QGTC_PACK_COL(8)
        default: return QGTC_ERR_UNSUPPORTED;
    }
#undef QGTC_PACK_COL
    note_launch();
    return check_launch("pack_col_kernel");
}

// Batched ROW pack of `count` matrices (ins[i]: H[i] x W[i] fp32, outs[i]: its R(nbits, H, W) words) in ceil(count / kPackBatchMax) launches.
int launch_val2bit_row_batched(const float* const* ins, const int* Hs, const int* Ws, uint32_t* const* outs, int count, int nbits, cudaStream_t st) {
    static thread_local PackBatchArgs A;
    for (int base = 0; base < count; base += kPackBatchMax) {
        const int n = count - base < kPackBatchMax ? count - base : kPackBatchMax;
        unsigned long long units = 0;
        for (int i = 0; i < n; ++i) {
            A.p[i].in = ins[base + i];
            A.p[i].out = outs[base + i];
            A.p[i].H = Hs[base + i];
            A.p[i].W = Ws[base + i];
            A.p[i].unit0 = (unsigned int)units;
            units += (unsigned long long)p8(Hs[base + i]) * (unsigned long long)s128(Ws[base + i]);
            if (units >= (1ull << 31)) return QGTC_ERR_UNSUPPORTED;
        }
        A.count = n;
        A.nbits = nbits;
        // rows of at least 16 output words everywhere: the word-per-lane form (units = (matrix, row, group of 32 words)); else the shuffle form
        static const int form_env = [] { const char* e = getenv("QGTC_PACK_BATCHED"); return e ? (e[0] == 'w' ? 1 : (e[0] == 's' ? 2 : 0)) : 0; }();
        bool wide = true;
        for (int i = 0; i < n; ++i) wide = wide && s128(A.p[i].W) * 4 >= 16;
        if (form_env == 1) wide = true;
        if (form_env == 2) wide = false;
        if (wide) {
            units = 0;
            for (int i = 0; i < n; ++i) {
                A.p[i].unit0 = (unsigned int)units;
                units += (unsigned long long)p8(A.p[i].H) * (unsigned long long)((s128(A.p[i].W) * 4 + 31) >> 5);
            }
        }
        A.units = (unsigned int)units;
        const int grid = grid_for((long long)((units + 7) / 8), 8);
        if (wide) {
#define QGTC_PACK_BW(NB) case NB: pack_row_batched_word_kernel<NB><<<grid, 256, 0, st>>>(A); break;
            switch (nbits) {
                QGTC_PACK_BW(1) QGTC_PACK_BW(2) QGTC_PACK_BW(3) QGTC_PACK_BW(4) QGTC_PACK_BW(5) QGTC_PACK_BW(6) QGTC_PACK_BW(7) QGTC_PACK_BW(8)
                default: return QGTC_ERR_UNSUPPORTED;
            }
#undef QGTC_PACK_BW
        } else {
            pack_row_batched_kernel<<<grid, 256, 0, st>>>(A);
        }
        note_launch();
        if (int rc = check_launch("pack_row_batched_kernel")) return rc;
    }
    return QGTC_OK;
}

int launch_bit2val(const uint32_t* in, size_t in_words, int nbits, int H, int W, int col_major, int* out, cudaStream_t st) {
    const long long n = (long long)H * W;
    const int grid = grid_for((n + 255) / 256, 8);
    unpack_kernel<<<grid, 256, 0, st>>>(Words{in, (unsigned long long)in_words}, nbits, H, W, col_major, out);
    note_launch();
    return check_launch("unpack_kernel");
}

int launch_count_nonzero_tiles(const uint32_t* x1, size_t x1_words, int M, int K, int bit1, unsigned long long* nz_dev, cudaStream_t st) {
    const long long tiles = (long long)bit1 * s8(M) * s128(K);
    const int grid = grid_for((tiles + 255) / 256, 8);
    count_nonzero_tiles_kernel<<<grid, 256, 0, st>>>(Words{x1, (unsigned long long)x1_words}, M, K, bit1, nz_dev);
    note_launch();
    return check_launch("count_nonzero_tiles_kernel");
}

int launch_edges_to_bits(const int* rows, const int* cols, long long n_edges, int H, int W, uint32_t* out, size_t out_words, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, out_words * sizeof(uint32_t), st);
    if (e != cudaSuccess) return record_cuda_error(e, "cudaMemsetAsync(edges2bit)");
    if (n_edges > 0) {
        const int grid = grid_for((n_edges + 255) / 256, 8);
        edges_to_bits_kernel<<<grid, 256, 0, st>>>(rows, cols, n_edges, H, W, out);
        note_launch();
        return check_launch("edges_to_bits_kernel");
    }
    return QGTC_OK;
}

int launch_tile_bitmap(const uint32_t* x1, size_t x1_words, size_t bstride, int M, int K, int count, uint32_t* bitmap, cudaStream_t st) {
    const long long words = (long long)count * ((p8(M) + 127) / 128) * ((s128(K) + 31) / 32);
    cudaError_t e = cudaMemsetAsync(bitmap, 0, (size_t)words * sizeof(uint32_t), st);
    if (e != cudaSuccess) return record_cuda_error(e, "cudaMemsetAsync(tile bitmap)");
    const long long units = words * (128 / kBmRows);
    const int grid = grid_for((units + 7) / 8, 8);
    tile_bitmap_kernel<<<grid, 256, 0, st>>>(x1, (unsigned long long)x1_words, (unsigned long long)bstride, M, K, count, bitmap);
    note_launch();
    return check_launch("tile_bitmap_kernel");
}

}  // namespace qgtc
