/*
 * qgtc_b200.h -- C ABI of the B200-native QGTC hot path (libqgtc_b200.so).
 *
 * Every entry point replaces one launcher of the reference extension
 * (QGTC_PPoPP22/QGTC_module); the citation after each prototype is the reference interface it
 * stands in for.  The PyTorch extension `QGTC` (qgtc_ppopp22_b200/csrc/qgtc_torch.cpp) and the
 * ctypes mirror (qgtc_ppopp22_b200/capi.py) are thin callers of exactly these functions.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Pointers named *_dev are device pointers
 *     valid on the CUDA device that is current when the call is made; *_host are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All device
 *     entry points are asynchronous with respect to the host unless stated otherwise.
 *   - packed tensors are arrays of 32-bit words ("raw memory").  Every *_words argument is the
 *     number of words the buffer really holds: reads that a reference index formula would place
 *     beyond it return 0 and writes beyond it are dropped (zero-extension; DESIGN.md "UB corners").
 *   - packed device buffers must be 16-byte aligned (QGTC_ERR_ALIGNMENT otherwise).
 *   - output buffers need NOT be zero-initialised: every word/element of the output is written.
 *   - return value: QGTC_OK (0) or a qgtc_status error code; nothing is printed, nothing exits.
 *
 * Layouts (bits = number of planes, MSB-first words: element e of a 32-group is bit 31-(e%32)):
 *   QGTC_LAYOUT_ROW     R(bits,H,W)    words [bits][P8(H)][S128(W)*4]           kernel.h:204-242
 *   QGTC_LAYOUT_COL128  C128(bits,H,W) words [bits][P128(W)][S128(H)*4]         kernel.h:75-106
 *   QGTC_LAYOUT_COL8    C8(bits,H,W)   allocation bits*S128(H)*4*P8(W) words, written with the
 *                                      C128 plane stride (QGTC_device.cu:83 vs kernel.h:88)
 *   with S8(x)=ceil(x/8), S128(x)=ceil(x/128), P8=8*S8, P128=128*S128 (utility.h:33-45).
 */
#ifndef QGTC_B200_H
#define QGTC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QGTC_B200_ABI_VERSION 1

typedef enum qgtc_status {
    QGTC_OK = 0,
    QGTC_ERR_INVALID_ARG = 1, /* null pointer, non-positive dimension, bad enum            */
    QGTC_ERR_UNSUPPORTED = 2, /* bit width outside [1,8], dimension beyond 32-bit indexing  */
    QGTC_ERR_ALIGNMENT = 3,   /* packed buffer not 16-byte aligned                          */
    QGTC_ERR_SIZE = 4,        /* output buffer smaller than the op's result                 */
    QGTC_ERR_CUDA = 5,        /* a CUDA runtime call or launch failed (see qgtc_last_cuda_error) */
    QGTC_ERR_NO_DEVICE = 6    /* no sm_100 device is current                                */
} qgtc_status;

typedef enum qgtc_layout { QGTC_LAYOUT_ROW = 0, QGTC_LAYOUT_COL128 = 1, QGTC_LAYOUT_COL8 = 2 } qgtc_layout;

/* GEMM engine selection (A/B measurement, DESIGN.md "tensor path chosen by measurement"). */
typedef enum qgtc_gemm_engine {
    QGTC_ENGINE_AUTO = 0,    /* library picks (tcgen05 when the shape qualifies, else mma.sync) */
    QGTC_ENGINE_MMA_SYNC = 1,/* legacy warp-level mma.sync.m16n8k32.u8 on re-assembled planes   */
    QGTC_ENGINE_TCGEN05 = 2, /* tcgen05.mma, TMEM accumulators, TMA-fed uniform batches (1-bit X1 only) */
    QGTC_ENGINE_GROUPED = 3  /* tcgen05.mma, grouped variable-shape engine: any bit1, in-kernel operand conversion */
} qgtc_gemm_engine;

/* flags for the GEMM entry points */
#define QGTC_FLAG_NO_ZERO_JUMP 0x1u /* do not skip all-zero adjacency tiles (for A/B timing)        */
#define QGTC_FLAG_COL_REF_QUIRK 0x2u /* bitmm2bit_col: reproduce the shipped kernel's stores exactly */
#define QGTC_FLAG_TC05_PAIR 0x4u    /* tcgen05 engine: 2-CTA clusters, tcgen05.mma.cta_group::2 (M = 256)  */
/* Grouped engine, programmatic dependent launch: the caller declares that an operand was NOT written by the work queued
 * just before this call on the stream (adjacency and its tile bitmap, weights).  The kernel may then load it while the
 * preceding kernel is still draining; operands without the flag are read only after that kernel has completed. */
#define QGTC_FLAG_X1_STATIC 0x8u
#define QGTC_FLAG_X2_STATIC 0x10u

int qgtc_abi_version(void);
const char* qgtc_strerror(int status);
/* text of the last CUDA error seen by this thread inside the library ("" if none) */
const char* qgtc_last_cuda_error(void);

/* number of 32-bit words of a packed tensor; allocation rule of QGTC_device.cu:83,97,115 */
size_t qgtc_packed_words(int layout, int bits, int H, int W);
/* reported 2-D tensor shape of the same (rows, cols) */
void qgtc_packed_shape(int layout, int bits, int H, int W, int64_t* rows, int64_t* cols);

/* val2bit: quantize fp32 [H,W] (row-major, contiguous) and pack nbits planes.
 * Replaces val2bit_cuda, QGTC_device.cu:44-130 (Quantize_val + QGTC_layer_input / PackFcWeight128),
 * fused into one pass with no int32 temporary. */
int qgtc_val2bit(const float* in_dev, int H, int W, int nbits, int col_major, int output_layer,
                 int32_t* out_dev, size_t out_words, void* stream);

/* val2bit of a whole list of matrices (same nbits and layout) in one call.  The ROW layout (col_major == 0: the batch
 * adjacencies of sampler.py:97-102, one QGTC.val2bit per batch there) is packed by ONE kernel launch per 448 matrices --
 * a 759 x 759 adjacency alone is 4 us of launch latency for 2.3 MB; the column-major layouts run one launch per matrix.
 * Every matrix gets exactly the words qgtc_val2bit would write. */
typedef struct qgtc_pack_problem {
    const float* in_dev;   /* fp32 [H, W], row-major, contiguous */
    int32_t* out_dev;      /* packed result, 16-byte aligned */
    size_t out_words;      /* capacity of out_dev in 32-bit words (>= qgtc_packed_words) */
    int H, W;
} qgtc_pack_problem;
int qgtc_val2bit_batched(const qgtc_pack_problem* problems_host, int count, int nbits, int col_major, int output_layer,
                         void* stream);

/* bit2val: unpack to int32 [H,W].  Replaces bit2val_cuda, QGTC_device.cu:135-206 (silent: the
 * reference's debug printf of the whole tensor is not reproduced). */
int qgtc_bit2val(const int32_t* in_dev, size_t in_words, int nbits, int H, int W, int col_major,
                 int output_layer, int32_t* out_dev, void* stream);

/* bitMM2Bit: R(out_bit,M,N) <- requantize( X1[R(bit1,M,K)] x X2[C128(bit2,K,N)] ).
 * Replaces bitMM2Bit_cuda, QGTC_device.cu:211-266 (kernel QGTC_layer_hidden, kernel.h:245-391). */
int qgtc_bitmm2bit(const int32_t* x1_dev, size_t x1_words, const int32_t* x2_dev, size_t x2_words,
                   int M, int K, int N, int bit1, int bit2, int out_bit,
                   int32_t* out_dev, size_t out_words, int engine, unsigned flags, void* stream);

/* bitMM2Bit_col: same product, output packed C128(out_bit,M,N) so it can be the next B operand.
 * Replaces bitMM2Bit_col_cuda, QGTC_device.cu:441-489 (kernel.h:651-810).  Default = intended
 * semantics; QGTC_FLAG_COL_REF_QUIRK = byte-exact reproduction of the shipped (defective) kernel. */
int qgtc_bitmm2bit_col(const int32_t* x1_dev, size_t x1_words, const int32_t* x2_dev, size_t x2_words,
                       int M, int K, int N, int bit1, int bit2, int out_bit,
                       int32_t* out_dev, size_t out_words, int engine, unsigned flags, void* stream);

/* bitMM2Int: float [M,N] <- (float) acc.  pad128 selects the X2 plane stride (C128 vs C8).
 * Replaces bitMM2Int_cuda, QGTC_device.cu:495-542 (kernel.h:816-1054). */
int qgtc_bitmm2int(const int32_t* x1_dev, size_t x1_words, const int32_t* x2_dev, size_t x2_words,
                   int M, int K, int N, int bit1, int bit2, int pad128,
                   float* out_dev, size_t out_elems, int engine, unsigned flags, void* stream);

/* Tile counters of bitMM2Bit_base_cnt / bitMM2Bit_zerojump_cnt (QGTC_device.cu:272-373,
 * kernel.h:394-648): *total_host = counter_global increment, *nonzero_host = counter increment
 * of one call (32 lanes x tiles x plane pairs x k steps).  SYNCHRONOUS on `stream`. */
int qgtc_tile_counts(const int32_t* x1_dev, size_t x1_words, int M, int K, int N, int bit1, int bit2,
                     uint64_t* total_host, uint64_t* nonzero_host, void* stream);

/* Precomputed tile bitmap of a 1-bit adjacency (zero-tile jumping, reference kernel.h:574-591 lifted to the
 * tile the tensor core consumes): bit c of word w of tile (problem z, row tile mt) is set iff rows
 * 128mt..128mt+127 x K bits 128(32w+c)..+127 contain a one.  Words: count * ceil(P8(M)/128) * ceil(S128(K)/32)
 * (qgtc_tile_bitmap_words).  Compute it once per adjacency (e.g. when the batch is packed) and pass it to
 * qgtc_bitmm2bit_tilemap / qgtc_bitmm2int_tilemap: all-zero tiles are then neither loaded, expanded nor
 * multiplied.  Results are identical to the calls without a bitmap. */
size_t qgtc_tile_bitmap_words(int M, int K, int count);
int qgtc_tile_bitmap(const int32_t* x1_dev, size_t x1_words, size_t x1_stride_words, int M, int K, int count,
                     uint32_t* bitmap_dev, size_t bitmap_words, void* stream);
int qgtc_bitmm2bit_tilemap(const int32_t* x1_dev, size_t x1_stride_words, const int32_t* x2_dev, size_t x2_stride_words,
                           const uint32_t* bitmap_dev, int count, int M, int K, int N, int bit2, int out_bit,
                           int32_t* out_dev, size_t out_stride_words, int engine, unsigned flags, void* stream);
int qgtc_bitmm2int_tilemap(const int32_t* x1_dev, size_t x1_words, const int32_t* x2_dev, size_t x2_words,
                           const uint32_t* bitmap_dev, int M, int K, int N, int bit2, int pad128,
                           float* out_dev, size_t out_elems, int engine, unsigned flags, void* stream);

/* Packed 1-bit adjacency R(1,H,W) directly from an edge list (int32 row / column indices on the device): the
 * result equals qgtc_val2bit of the dense 0/1 matrix with ones at the listed positions (duplicates collapse,
 * out-of-range edges are ignored), without materialising the dense fp32 H x W matrix the reference builds
 * (sampler.py:87-89, torch.sparse...to_dense()). */
int qgtc_edges2bit(const int32_t* rows_dev, const int32_t* cols_dev, size_t n_edges, int H, int W,
                   int32_t* out_dev, size_t out_words, void* stream);

/* Batched bit-GEMM: `count` independent problems of identical shape, one launch (persistent CTAs).
 * Problem i uses x1 + i*x1_stride_words, x2 + i*x2_stride_words, out + i*out_stride_words.
 * This is the throughput form of bitMM2Bit_profile (QGTC_device.cu:379-434: 200 back-to-back
 * launches) and of the per-batch loop of main_qgtc.py:113-154. */
int qgtc_bitmm2bit_batched(const int32_t* x1_dev, size_t x1_stride_words, const int32_t* x2_dev,
                           size_t x2_stride_words, int count, int M, int K, int N, int bit1, int bit2,
                           int out_bit, int32_t* out_dev, size_t out_stride_words, int engine,
                           unsigned flags, void* stream);

/* Host-buffer form used for end-to-end timing: copies the packed operands host->device, runs
 * bitMM2Bit and copies the packed result back, all on `stream`, then synchronises.  Scratch
 * device memory is owned by the library and grows on demand. */
int qgtc_bitmm2bit_host(const int32_t* x1_host, size_t x1_words, const int32_t* x2_host, size_t x2_words,
                        int M, int K, int N, int bit1, int bit2, int out_bit,
                        int32_t* out_host, size_t out_words, int engine, unsigned flags, void* stream);

/* Pipelined host-buffer form for a batch: the problems are processed in groups; the upload of group g+1 and the
 * download of group g-1 overlap the GEMM of group g on library-owned streams (host buffers should be pinned for
 * the copies to be asynchronous).  Synchronises before returning.  Same results as `count` calls of
 * qgtc_bitmm2bit_host. */
int qgtc_bitmm2bit_host_batched(const int32_t* x1_host, size_t x1_stride_words, const int32_t* x2_host, size_t x2_stride_words,
                                int count, int M, int K, int N, int bit1, int bit2, int out_bit,
                                int32_t* out_host, size_t out_stride_words, int engine, unsigned flags, void* stream);

/* The same for several feature matrices per adjacency (e.g. one step of the bit-width sweep: 1/2/4/8-bit features of the
 * same graphs): every group's adjacency is uploaded ONCE and multiplied with each variant's features.  Same results as
 * `nvar` calls of qgtc_bitmm2bit_host_batched. */
typedef struct qgtc_host_variant {
    const int32_t* x2_host;
    size_t x2_stride_words;
    int bit2, out_bit;
    int32_t* out_host;
    size_t out_stride_words;
} qgtc_host_variant;
int qgtc_bitmm2bit_host_batched_multi(const int32_t* x1_host, size_t x1_stride_words, int count, int M, int K, int N, int bit1,
                                      const qgtc_host_variant* variants, int nvar, int engine, unsigned flags, void* stream);

/* Grouped bit-GEMM: ONE launch for `count` independent problems that share N, the bit widths and the kind of result but
 * have their own M, K, operands and output -- e.g. one op of the Cluster-GCN / batched-GIN forward for all batches of a
 * rank (the per-batch loop of main_qgtc.py:113-154 turned inside out: 6 launches per epoch instead of 6 x batches).
 * Results are identical to `count` calls of qgtc_bitmm2bit (QGTC_GROUP_BITS_ROW; QGTC_device.cu:211-266),
 * qgtc_bitmm2int (QGTC_GROUP_FLOAT; QGTC_device.cu:495-542) or qgtc_bitmm2bit_col (QGTC_GROUP_BITS_COL;
 * QGTC_device.cu:441-489, intended semantics): same raw-memory indexing, same zero-extension rule.
 * `problems_host` is an array in HOST memory (read before the call returns; its pointers are device pointers).
 * out_size: words of the packed result buffer, or float elements for QGTC_GROUP_FLOAT.  tile_bitmap_dev (optional, 1-bit X1
 * only): the problem's qgtc_tile_bitmap; all-zero tiles are then skipped.  pad128 is used by QGTC_GROUP_FLOAT only. */
typedef struct qgtc_group_problem {
    const int32_t* x1_dev;
    size_t x1_words;
    const int32_t* x2_dev;
    size_t x2_words;
    void* out_dev;
    size_t out_size;
    const uint32_t* tile_bitmap_dev;
    int32_t M, K;
} qgtc_group_problem;
typedef enum qgtc_group_kind { QGTC_GROUP_BITS_ROW = 0, QGTC_GROUP_FLOAT = 1, QGTC_GROUP_BITS_COL = 2 } qgtc_group_kind;
int qgtc_bitmm_grouped(const qgtc_group_problem* problems_host, int count, int N, int bit1, int bit2, int out_bit,
                       int kind, int pad128, int engine, unsigned flags, void* stream);

/* Fused aggregation + update: for every problem  out2 = op2( bitMM2Bit(X1, X2) , W )  in ONE launch -- the pairs
 * main_qgtc.py:148+150, :152+153 (GCN) and :131+132, :133+134, :135+136 (GIN).  The first op is bitMM2Bit (result R(out_bit, M, N),
 * N <= 64); its requantized tile goes straight back into tensor memory as the A operand of the second op, which is
 * bitMM2Bit(o, W, M, N, N2, out_bit, bit_w, out_bit2) (kind2 = QGTC_GROUP_BITS_ROW) or bitMM2Int(o, W, M, N, N2, out_bit, bit_w, pad128)
 * (kind2 = QGTC_GROUP_FLOAT) with one weight matrix W = C(bit_w, N, N2) shared by all problems (N2 <= 128).  out2 is bit-identical
 * to the two separate calls.  problems[i].out_dev / out_size are only used when write_out1 != 0 (the first op's own result is then
 * written as well).  Returns QGTC_ERR_UNSUPPORTED when the shapes do not fit (the caller then issues the two ops separately). */
typedef struct qgtc_fused_update {
    const int32_t* w_dev;
    size_t w_words;
    int N2, bit_w, out_bit2, kind2, pad128;
} qgtc_fused_update;
int qgtc_bitmm_grouped_fused(const qgtc_group_problem* problems_host, int count, int N, int bit1, int bit2, int out_bit,
                             const qgtc_fused_update* upd, void* const* out2_dev, const size_t* out2_size, int write_out1,
                             int engine, unsigned flags, void* stream);

/* Number of kernels this library has launched from the calling process (bench.py gpu_launches). */
uint64_t qgtc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QGTC_B200_H */
