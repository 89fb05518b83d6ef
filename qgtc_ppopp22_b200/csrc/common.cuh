// Shared helpers of the qgtc_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "qgtc_b200.h"

namespace qgtc {

// padding rules of the packed formats (reference utility.h:33-45)
__host__ __device__ constexpr int s8(int x) { return (x + 7) >> 3; }
__host__ __device__ constexpr int s128(int x) { return (x + 127) >> 7; }
__host__ __device__ constexpr int p8(int x) { return s8(x) << 3; }
__host__ __device__ constexpr int p128(int x) { return s128(x) << 7; }

constexpr int kNumSMs = 148;  // B200

// bookkeeping shared by the ABI layer
void note_launch(int n = 1);
int record_cuda_error(cudaError_t e, const char* where);  // returns QGTC_ERR_CUDA
int check_launch(const char* where);                      // cudaGetLastError -> status

// A packed operand as the kernels see it: raw words + how many really exist.
struct Words {
    const uint32_t* p;
    unsigned long long n;  // words available (zero-extension beyond)
};

// Zero-extended loads on raw word memory (DESIGN.md "UB corners", SURVEY 8c rule 10).
__device__ __forceinline__ uint32_t ldw(const Words& w, unsigned long long idx) {
    return idx < w.n ? __ldg(w.p + idx) : 0u;
}
// 4 consecutive words starting at idx (idx % 4 == 0, base 16-byte aligned).
__device__ __forceinline__ uint4 ldw4(const Words& w, unsigned long long idx) {
    if (idx + 4 <= w.n) return __ldg(reinterpret_cast<const uint4*>(w.p + idx));
    uint4 r;
    r.x = ldw(w, idx);
    r.y = ldw(w, idx + 1);
    r.z = ldw(w, idx + 2);
    r.w = ldw(w, idx + 3);
    return r;
}

// Requantize of QGTC_layer_hidden (reference kernel.h:31-37 with max=2^ob, min=0); the caller
// keeps only bits 0..ob-1, so acc == 2^ob contributes no bits.
__device__ __forceinline__ int requant(int acc, int ob) {
    const int mx = 1 << ob;
    return acc > mx ? mx - 1 : (acc < 0 ? 1 : acc);
}

// Quantize of val2bit (reference kernel.h:39-44,66-68): clip with strict compares, then
// round-half-even; NaN -> 0 (cvt.rni.s32.f32 semantics).
__device__ __forceinline__ int quantize_val(float x, float ub) {
    float y = x;
    if (x < 0.0f) y = 1.0f;
    else if (x > ub) y = ub - 1.0f;
    return __float2int_rn(y);
}

// 8 x 8 bit-matrix transpose (Hacker's Delight 7-3) of eight bytes: x = rows 0..3 (row 0 in the top byte),
// y = rows 4..7.  Afterwards byte (p - 4) of x (p >= 4) / byte p of y (p < 4), counted from the low end, holds
// bit p of the eight rows with row 0 in its MSB -- the MSB-first order of the packed plane words.
__device__ __forceinline__ void transpose8x8(uint32_t& x, uint32_t& y) {
    uint32_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AAu;  x = x ^ t ^ (t << 7);
    t = (y ^ (y >> 7)) & 0x00AA00AAu;  y = y ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCCu; x = x ^ t ^ (t << 14);
    t = (y ^ (y >> 14)) & 0x0000CCCCu; y = y ^ t ^ (t << 14);
    t = (x & 0xF0F0F0F0u) | ((y >> 4) & 0x0F0F0F0Fu);
    y = ((x << 4) & 0xF0F0F0F0u) | (y & 0x0F0F0F0Fu);
    x = t;
}


// val2bit quantizer without the quarter-rate float->int conversion: after the reference's clip (strict compares) the value
// lies in [0, 2^b] (or is NaN: fmaxf / fminf return the non-NaN operand, so NaN -> 0 as cvt.rni does), and adding 1.5 * 2^23
// rounds it to nearest-even into the low mantissa bits.  The caller keeps the low byte (or fewer bits): q == 2^b -> 0 bits.
__device__ __forceinline__ uint32_t quantize_bits(float x, float ub) {
    float y = x < 0.0f ? 1.0f : x;
    y = x > ub ? ub - 1.0f : y;
    y = fminf(fmaxf(y, 0.0f), ub);
    return __float_as_uint(y + 12582912.0f);
}

// Plane words of 32 quantized values: t[k] holds elements 4k .. 4k+3, the first one in the top byte (values masked to NB
// bits).  w[p] = bit p of the 32 elements, element 0 in the MSB (the packed formats' word order).  Up to three planes: one
// multiply-gather per plane and register; from four planes up: 8 x 8 bit transposes (22 operations per eight values for all
// planes) and byte permutes.
template <int NB>
__device__ __forceinline__ void planes_from_bytes(const uint32_t (&t)[8], uint32_t (&w)[NB]) {
    if (NB <= 3) {
#pragma unroll
        for (int p = 0; p < NB; ++p) {
            uint32_t word = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) word |= ((((t[k] >> p) & 0x01010101u) * 0x01020408u) >> 24) << (28 - 4 * k);
            w[p] = word;
        }
    } else {
        uint32_t xs[4], ys[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            xs[g] = t[2 * g];
            ys[g] = t[2 * g + 1];
            transpose8x8(xs[g], ys[g]);
        }
#pragma unroll
        for (int p = 0; p < NB; ++p) {
            const uint32_t sel = (uint32_t)((p & 3) | ((4 + (p & 3)) << 4));
            const uint32_t lo = p >= 4 ? __byte_perm(xs[3], xs[2], sel) : __byte_perm(ys[3], ys[2], sel);
            const uint32_t hi = p >= 4 ? __byte_perm(xs[1], xs[0], sel) : __byte_perm(ys[1], ys[0], sel);
            w[p] = __byte_perm(lo, hi, 0x5410u);
        }
    }
}

}  // namespace qgtc
