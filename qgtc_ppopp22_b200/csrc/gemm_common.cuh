// Arguments and epilogue helpers shared by the bit-GEMM engines.
#pragma once

#include "common.cuh"

namespace qgtc {

enum Epilogue : int {
    kEpiBitsRow = 0,      // bitMM2Bit      -> R(out_bit, M, N)          reference kernel.h:347-389
    kEpiFloat = 1,        // bitMM2Int      -> float [M, N]              reference kernel.h:905-930
    kEpiBitsCol = 2,      // bitMM2Bit_col  -> C128(out_bit, M, N), intended semantics
    kEpiBitsColQuirk = 3  // bitMM2Bit_col as shipped (kernel.h:764-806)
};

struct GemmArgs {
    const uint32_t* x1;  // A operand, R(bit1, M, K) indexing on raw words
    const uint32_t* x2;  // B operand, C(bit2, K, N) indexing on raw words
    void* out;
    unsigned long long x1_words, x2_words;              // words really present (zero-extension beyond)
    unsigned long long x1_bstride, x2_bstride, out_bstride;  // per-problem strides of a batch (words / elements)
    unsigned long long act_offset;                      // A plane stride  = P8(M)*S128(K)*4   (kernel.h:265)
    unsigned long long w_offset;                        // B plane stride  = S128(K)*P128(N)*4 or *P8(N)*4 (kernel.h:266,836)
    int M, K, N;
    int bit1, bit2, out_bit;
    int gdk;   // S128(K): number of 128-bit K chunks
    int count; // batch size
    unsigned flags;
    const uint32_t* tile_bitmap;  // optional (qgtc_tile_bitmap): [count][ceil(P8(M)/128)][ceil(S128(K)/32)] words, or null
};

// Gather bit p of four consecutive byte values (lowest address = first column) into a nibble
// whose MSB is the first column: the MSB-first convention of the packed words.
__device__ __forceinline__ uint32_t nibble_of_plane(uint32_t four_bytes, int p) {
    return (((four_bytes >> p) & 0x01010101u) * 0x08040201u) >> 24;
}

}  // namespace qgtc
