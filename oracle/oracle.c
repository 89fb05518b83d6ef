/*
 * CPU bit-plane oracle for the QGTC hot path, C restatement.  TEST INFRASTRUCTURE ONLY.
 *
 * Same contract as oracle/bitplane_oracle.py (which it is cross-checked against in
 * tests/test_oracle.py): a plain restatement of the reference kernels' arithmetic in
 * /root/reference/QGTC_module/kernel.h, used (a) as the checker at sizes numpy is too slow
 * for and (b) as bench.py's cpu_baseline ("port": AND+popcount over bit planes on the host
 * cores, OpenMP).  Only tests/, __graft_entry__.smoke() and bench.py may load this library;
 * the product (qgtc_ppopp22_b200/) never does.
 *
 * Packed tensors are flat uint32 word arrays; reads past the end return 0 and writes past the
 * end are dropped (zero-extension rule, SURVEY.md 8c rule 10).
 *
 * Parity pin: hand-derived KATs + golden vectors produced by the unmodified reference kernels
 * on a B200 (tests/golden/), see the header of bitplane_oracle.py.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define S8(x) (((x) + 7) >> 3)
#define S128(x) (((x) + 127) >> 7)
#define P8(x) (S8(x) << 3)
#define P128(x) (S128(x) << 7)

enum { LAYOUT_ROW = 0, LAYOUT_COL128 = 1, LAYOUT_COL8 = 2 };

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

size_t oracle_packed_words(int layout, int bits, int H, int W) {
    if (layout == LAYOUT_ROW) return (size_t)bits * P8(H) * S128(W) * 4;       /* QGTC_device.cu:115 */
    if (layout == LAYOUT_COL128) return (size_t)bits * S128(H) * 4 * P128(W);  /* QGTC_device.cu:97  */
    return (size_t)bits * S128(H) * 4 * P8(W);                                 /* QGTC_device.cu:83  */
}

/* kernel.h:39-44 + :66-68 */
static inline int32_t quantize_one(float x, int nbits) {
    const float ub = (float)(1 << nbits);
    float y = x;
    if (x < 0.0f) y = 1.0f;
    else if (x > ub) y = ub - 1.0f;
    if (isnan(y)) return 0;                 /* __float2int_rn(NaN) == 0 */
    return (int32_t)nearbyintf(y);          /* default FE_TONEAREST == round-half-even */
}

static inline uint32_t ld_ext(const uint32_t* p, size_t words, size_t idx) { return idx < words ? p[idx] : 0u; }

/* kernel.h:49-71, :75-106, :204-242 ; QGTC_device.cu:44-130 */
int oracle_val2bit(const float* x, int H, int W, int nbits, int layout, uint32_t* out, size_t out_words) {
    if (out_words != oracle_packed_words(layout, nbits, H, W)) return -1;
    memset(out, 0, out_words * sizeof(uint32_t));
    if (layout == LAYOUT_ROW) {
        const size_t wpr = (size_t)S128(W) * 4, plane = (size_t)P8(H) * wpr;
#pragma omp parallel for schedule(static)
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c) {
                const int32_t q = quantize_one(x[(size_t)r * W + c], nbits);
                for (int p = 0; p < nbits; ++p)
                    if ((q >> p) & 1) out[p * plane + r * wpr + (c >> 5)] |= 1u << (31 - (c & 31));
            }
    } else {
        const size_t kw = (size_t)S128(H) * 4, plane = (size_t)S128(H) * P128(W) * 4; /* kernel.h:88 */
#pragma omp parallel for schedule(static)
        for (int c = 0; c < W; ++c)
            for (int r = 0; r < H; ++r) {
                const int32_t q = quantize_one(x[(size_t)r * W + c], nbits);
                for (int p = 0; p < nbits; ++p) {
                    const size_t idx = p * plane + c * kw + (r >> 5);
                    if (((q >> p) & 1) && idx < out_words) out[idx] |= 1u << (31 - (r & 31));
                }
            }
    }
    return 0;
}

/* kernel.h:109-139, :173-201 ; QGTC_device.cu:135-206 */
int oracle_bit2val(const uint32_t* in, size_t in_words, int nbits, int H, int W, int col_major, int32_t* out) {
    memset(out, 0, (size_t)H * W * sizeof(int32_t));
    if (!col_major) {
        const size_t wpr = (size_t)S128(W) * 4, plane = (size_t)P8(H) * wpr;
        for (int p = 0; p < nbits; ++p)
            for (int r = 0; r < H; ++r)
                for (int c = 0; c < W; ++c)
                    out[(size_t)r * W + c] += (int32_t)((ld_ext(in, in_words, p * plane + r * wpr + (c >> 5)) >> (31 - (c & 31))) & 1u) << p;
    } else {
        const size_t kw = (size_t)S128(H) * 4, plane = (size_t)S128(H) * P128(W) * 4;
        for (int p = 0; p < nbits; ++p)
            for (int r = 0; r < H; ++r)
                for (int c = 0; c < W; ++c)
                    out[(size_t)r * W + c] += (int32_t)((ld_ext(in, in_words, p * plane + c * kw + (r >> 5)) >> (31 - (r & 31))) & 1u) << p;
    }
    return 0;
}

/* Gather one operand's planes into a dense zero-extended scratch [bits][rows][kw] (uint32). */
static uint32_t* gather_planes(const uint32_t* x, size_t words, int bits, int rows, size_t kw, size_t plane_stride) {
    uint32_t* g = (uint32_t*)calloc((size_t)bits * rows * kw + 2, sizeof(uint32_t));
    if (!g) return NULL;
    for (int p = 0; p < bits; ++p)
        for (int r = 0; r < rows; ++r)
            for (size_t w = 0; w < kw; ++w)
                g[((size_t)p * rows + r) * kw + w] = ld_ext(x, words, p * plane_stride + r * kw + w);
    return g;
}

/* kernel.h:265-342: acc[m][n] = sum_{p,q} popc(X1[p][m] & X2[q][n]) << (p+q), wrapping int32.
 * acc has rows x cols entries where rows/cols may be M/N or P8(M)/P8(N) (offsets do not change). */
int oracle_bitgemm_acc(const uint32_t* x1, size_t x1_words, const uint32_t* x2, size_t x2_words,
                       int M, int K, int N, int bit1, int bit2, int pad128_b, int rows, int cols, int32_t* acc) {
    const size_t kw = (size_t)S128(K) * 4;
    const size_t act_offset = (size_t)P8(M) * kw;
    const size_t w_offset = kw * (size_t)(pad128_b ? P128(N) : P8(N));
    uint32_t* a = gather_planes(x1, x1_words, bit1, rows, kw, act_offset);
    uint32_t* b = gather_planes(x2, x2_words, bit2, cols, kw, w_offset);
    if (!a || !b) { free(a); free(b); return -2; }
    const size_t kq = kw / 2; /* kw is a multiple of 4 */
#pragma omp parallel for schedule(static)
    for (int m = 0; m < rows; ++m) {
        for (int n = 0; n < cols; ++n) {
            uint32_t c = 0;
            for (int q = 0; q < bit2; ++q) {
                const uint64_t* bq = (const uint64_t*)(b + ((size_t)q * cols + n) * kw);
                for (int p = 0; p < bit1; ++p) {
                    const uint64_t* ap = (const uint64_t*)(a + ((size_t)p * rows + m) * kw);
                    uint32_t cnt = 0;
                    for (size_t i = 0; i < kq; ++i) cnt += (uint32_t)__builtin_popcountll(ap[i] & bq[i]);
                    c += cnt << (p + q);
                }
            }
            acc[(size_t)m * cols + n] = (int32_t)c;
        }
    }
    free(a); free(b);
    return 0;
}

/* kernel.h:31-37 with (bitwidth=ob, max=2^ob, min=0) */
static inline int32_t requant_one(int32_t acc, int ob) {
    const int32_t mx = 1 << ob;
    if (acc > mx) return mx - 1;
    if (acc < 0) return 1;
    return acc;
}

/* QGTC_layer_hidden, kernel.h:245-391 -> R(out_bit, M, N) */
int oracle_bitmm2bit(const uint32_t* x1, size_t x1_words, const uint32_t* x2, size_t x2_words,
                     int M, int K, int N, int bit1, int bit2, int out_bit, uint32_t* out, size_t out_words) {
    if (out_words != oracle_packed_words(LAYOUT_ROW, out_bit, M, N)) return -1;
    int32_t* acc = (int32_t*)malloc((size_t)M * N * sizeof(int32_t));
    if (!acc) return -2;
    int rc = oracle_bitgemm_acc(x1, x1_words, x2, x2_words, M, K, N, bit1, bit2, 1, M, N, acc);
    if (rc) { free(acc); return rc; }
    memset(out, 0, out_words * sizeof(uint32_t));
    const size_t wpr = (size_t)S128(N) * 4, plane = (size_t)P8(M) * wpr;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            const int32_t v = requant_one(acc[(size_t)m * N + n], out_bit);
            for (int p = 0; p < out_bit; ++p)
                if ((v >> p) & 1) out[p * plane + m * wpr + (n >> 5)] |= 1u << (31 - (n & 31));
        }
    free(acc);
    return 0;
}

/* QGTC_layer_output_PAD8/_PAD128, kernel.h:816-1054 -> float [M,N] */
int oracle_bitmm2int(const uint32_t* x1, size_t x1_words, const uint32_t* x2, size_t x2_words,
                     int M, int K, int N, int bit1, int bit2, int pad128, float* out) {
    int32_t* acc = (int32_t*)malloc((size_t)M * N * sizeof(int32_t));
    if (!acc) return -2;
    int rc = oracle_bitgemm_acc(x1, x1_words, x2, x2_words, M, K, N, bit1, bit2, pad128, M, N, acc);
    if (!rc) for (size_t i = 0; i < (size_t)M * N; ++i) out[i] = (float)acc[i];
    free(acc);
    return rc;
}

/* QGTC_layer_hidden_col, kernel.h:651-810 -> C128(out_bit, M, N).
 * quirk=0: intended column-major pack.  quirk=1: exactly what the shipped kernel stores
 * (second half read at +4 instead of +32, kernel.h:782; transposed bounds test, :777-778). */
int oracle_bitmm2bit_col(const uint32_t* x1, size_t x1_words, const uint32_t* x2, size_t x2_words,
                         int M, int K, int N, int bit1, int bit2, int out_bit, int quirk,
                         uint32_t* out, size_t out_words) {
    if (out_words != oracle_packed_words(LAYOUT_COL128, out_bit, M, N)) return -1;
    const int PM = P8(M), PN = P8(N);
    int32_t* acc = (int32_t*)malloc((size_t)PM * PN * sizeof(int32_t));
    if (!acc) return -2;
    int rc = oracle_bitgemm_acc(x1, x1_words, x2, x2_words, M, K, N, bit1, bit2, 1, PM, PN, acc);
    if (rc) { free(acc); return rc; }
    memset(out, 0, out_words * sizeof(uint32_t));
    const size_t gdh = S128(M);
    if (!quirk) {
        const size_t kw = gdh * 4, plane = gdh * P128(N) * 4;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                const int32_t v = requant_one(acc[(size_t)m * PN + n], out_bit);
                for (int p = 0; p < out_bit; ++p)
                    if ((v >> p) & 1) out[p * plane + n * kw + (m >> 5)] |= 1u << (31 - (m & 31));
            }
    } else {
        uint8_t* ob = (uint8_t*)out; /* little-endian host, as the GPU */
        const size_t plane_bytes = gdh * P128(N) * 16;
        for (int bx = 0; bx < S8(M); ++bx)
            for (int by = 0; by < S8(N); ++by) {
                int32_t cs[64];
                for (int col = 0; col < 8; ++col)
                    for (int row = 0; row < 8; ++row)
                        cs[col * 8 + row] = requant_one(acc[(size_t)(bx * 8 + row) * PN + by * 8 + col], out_bit);
                for (int p = 0; p < out_bit; ++p)
                    for (int half = 0; half < 2; ++half)
                        for (int j = 0; j < 4; ++j) {
                            uint8_t byte = 0;
                            for (int i = 0; i < 8; ++i) {
                                const int lane = 8 * j + i, gy = lane % 8, gx = lane / 8;
                                const int inb = (by * 8 + gy < N) && (bx * 8 + gx + 4 * half < M);
                                const int bit = inb ? ((cs[(half ? 4 : 0) + lane] >> p) & 1) : 0;
                                byte |= (uint8_t)(bit << (7 - i));
                            }
                            ob[p * plane_bytes + (size_t)(by * 8 + 4 * half + j) * gdh * 16 + (bx ^ 3)] = byte;
                        }
            }
    }
    free(acc);
    return 0;
}

/* kernel.h:394-648 counters (intended 8x128-bit tile test; see bitplane_oracle.tile_counts) */
void oracle_tile_counts(const uint32_t* x1, size_t x1_words, int M, int K, int N, int bit1, int bit2,
                        uint64_t* total, uint64_t* nonzero) {
    const size_t gdk = S128(K), act_offset = (size_t)P8(M) * gdk * 4;
    uint64_t nz = 0;
    for (int p = 0; p < bit1; ++p)
        for (int bx = 0; bx < S8(M); ++bx)
            for (size_t i = 0; i < gdk; ++i) {
                uint32_t any = 0;
                for (int r = 0; r < 8; ++r)
                    for (int w = 0; w < 4; ++w)
                        any |= ld_ext(x1, x1_words, p * act_offset + (size_t)(bx * 8 + r) * gdk * 4 + i * 4 + w);
                nz += any != 0;
            }
    *total = 32ull * S8(M) * S8(N) * bit1 * bit2 * gdk;
    *nonzero = 32ull * S8(N) * bit2 * nz;
}
