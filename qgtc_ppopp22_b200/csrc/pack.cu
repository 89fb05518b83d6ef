// Quantize-and-pack (val2bit), unpack (bit2val) and tile counters.
//
// Replaces Quantize_val + QGTC_layer_input / PackFcWeight128 (reference kernel.h:49-106,204-242)
// and UnPackFc* (kernel.h:109-201).  One fused pass: fp32 is read once with 128-bit loads, the
// int32 temporary of the reference (QGTC_device.cu:63) does not exist, every output word is
// written (no pre-zeroed output needed).  HBM-bound: 4 B/element read, nbits/8 B/element written.
#include "common.cuh"

namespace qgtc {

// ------------------------------------------------------------------------------------------
// ROW layout: word [p][r][w] holds columns 32w..32w+31 of row r, MSB-first.
// One warp handles (row, 128-column group) units: lane l owns columns 4l..4l+3 (one float4),
// turns them into one nibble per plane, and an 8-lane OR-butterfly assembles the four words.
// ------------------------------------------------------------------------------------------
constexpr int kPackUnroll = 4;

template <bool kVec>
__global__ void __launch_bounds__(256) pack_row_kernel(const float* __restrict__ x, int H, int W, int nbits,
                                                        uint32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int gdy = s128(W);
    const int rows = p8(H);
    const long long units = (long long)rows * gdy;
    const size_t wpr = (size_t)gdy * 4, plane = (size_t)rows * wpr;
    const float ub = (float)(1 << nbits);
    const unsigned keep = (1u << nbits) - 1u;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;

    for (long long u0 = warp_id * kPackUnroll; u0 < units; u0 += nwarps * kPackUnroll) {
        float4 v[kPackUnroll];
#pragma unroll
        for (int j = 0; j < kPackUnroll; ++j) {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            const long long u = u0 + j;
            if (u < units) {
                const int r = (int)(u / gdy), c = (int)(u % gdy) * 128 + 4 * lane;
                if (r < H && c < W) {
                    const float* src = x + (size_t)r * W + c;
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
        }
#pragma unroll
        for (int j = 0; j < kPackUnroll; ++j) {
            const long long u = u0 + j;
            if (u >= units) break;  // warp-uniform
            const int r = (int)(u / gdy), cg = (int)(u % gdy);
            const unsigned q0 = (unsigned)quantize_val(v[j].x, ub) & keep;
            const unsigned q1 = (unsigned)quantize_val(v[j].y, ub) & keep;
            const unsigned q2 = (unsigned)quantize_val(v[j].z, ub) & keep;
            const unsigned q3 = (unsigned)quantize_val(v[j].w, ub) & keep;
            const unsigned t = (q0 << 24) | (q1 << 16) | (q2 << 8) | q3;
            unsigned mine = 0;
            for (int p = 0; p < nbits; ++p) {
                // bit p of the four values -> nibble, first column in the nibble's MSB
                const unsigned nib = (((t >> p) & 0x01010101u) * 0x01020408u) >> 24;
                unsigned w = nib << (28 - 4 * (lane & 7));
                w |= __shfl_xor_sync(0xffffffffu, w, 1);
                w |= __shfl_xor_sync(0xffffffffu, w, 2);
                w |= __shfl_xor_sync(0xffffffffu, w, 4);
                if ((lane & 7) == p) mine = w;
            }
            if ((lane & 7) < nbits)
                out[(size_t)(lane & 7) * plane + (size_t)r * wpr + (size_t)cg * 4 + (lane >> 3)] = mine;
        }
    }
}

// ------------------------------------------------------------------------------------------
// COL layouts: word [p][c][kw] holds rows 32kw..32kw+31 of column c, MSB-first (a transposing
// pack).  One warp handles a 128-row x 32-column tile: every row read is a coalesced 128-byte
// request, each lane accumulates the four words of its own column and stores them as one uint4.
// ------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) pack_col_kernel(const float* __restrict__ x, int H, int W, int col_limit,
                                                        uint32_t* __restrict__ out, unsigned long long out_words) {
    const int lane = threadIdx.x & 31;
    const int gdk = s128(H);
    const int cgroups = (col_limit + 31) >> 5;
    const long long units = (long long)gdk * cgroups;
    const size_t plane = (size_t)gdk * p128(W) * 4;  // reference kernel.h:88: always the P128 stride
    const float ub = (float)(1 << NB);
    const unsigned keep = (1u << NB) - 1u;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;

    for (long long u = warp_id; u < units; u += nwarps) {
        const int kb = (int)(u % gdk), c = (int)(u / gdk) * 32 + lane;
        uint32_t w[NB][4];
#pragma unroll
        for (int p = 0; p < NB; ++p) w[p][0] = w[p][1] = w[p][2] = w[p][3] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
                const int r = kb * 128 + j * 32 + i;
                float v = 0.f;
                if (r < H && c < W) v = __ldg(x + (size_t)r * W + c);
                const unsigned q = (unsigned)quantize_val(v, ub) & keep;
#pragma unroll
                for (int p = 0; p < NB; ++p) w[p][j] |= ((q >> p) & 1u) << (31 - i);
            }
        }
        if (c < col_limit) {
#pragma unroll
            for (int p = 0; p < NB; ++p) {
                const unsigned long long idx = (unsigned long long)p * plane + (unsigned long long)c * gdk * 4 + (unsigned long long)kb * 4;
                if (idx + 4 <= out_words)  // C8 with P8(W) != P128(W): the reference writes out of bounds here; dropped
                    *reinterpret_cast<uint4*>(out + idx) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
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
// host-side launchers
// ------------------------------------------------------------------------------------------
static int grid_for(long long work_items_per_block_units, int blocks_per_sm) {
    long long g = work_items_per_block_units;
    const long long cap = (long long)kNumSMs * blocks_per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

int launch_val2bit(const float* in, int H, int W, int nbits, int layout, uint32_t* out, size_t out_words, cudaStream_t st) {
    if (layout == QGTC_LAYOUT_ROW) {
        const long long units = (long long)p8(H) * s128(W);
        const long long blocks = (units + 8 * kPackUnroll - 1) / (8 * kPackUnroll);
        const int grid = grid_for(blocks, 8);
        const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15u) == 0);
        if (vec) pack_row_kernel<true><<<grid, 256, 0, st>>>(in, H, W, nbits, out);
        else pack_row_kernel<false><<<grid, 256, 0, st>>>(in, H, W, nbits, out);
        note_launch();
        return check_launch("pack_row_kernel");
    }
    const int col_limit = layout == QGTC_LAYOUT_COL128 ? p128(W) : p8(W);
    if (layout == QGTC_LAYOUT_COL8 && nbits > 1 && p8(W) != p128(W)) {
        // words between the C8 allocation's planes are never addressed by the C128 index formula
        cudaError_t e = cudaMemsetAsync(out, 0, out_words * sizeof(uint32_t), st);
        if (e != cudaSuccess) return record_cuda_error(e, "cudaMemsetAsync(val2bit C8)");
    }
    const long long units = (long long)s128(H) * ((col_limit + 31) >> 5);
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

}  // namespace qgtc
