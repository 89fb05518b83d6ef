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

}  // namespace qgtc
