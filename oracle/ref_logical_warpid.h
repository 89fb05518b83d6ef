/*
 * Oracle-only build shim (TEST INFRASTRUCTURE): pre-empts the reference's utility.h so that the
 * kernels' warp index is the LOGICAL one (threadIdx.x >> 5) instead of the %warpid special
 * register the reference reads (utility.h:49).  %warpid is a physical warp-slot id; on a B200 two
 * 1024-thread CTAs can share an SM (2048 threads/SM; the reference's pack and hidden kernels need
 * exactly 32 registers), where the second CTA sees ids 32..63 and maps to the wrong tiles.
 * Everything else keeps the reference's meaning; only the eight names its kernels use are
 * provided (ceil-div steps, padded sizes, lane id, byte flip, the byte type).
 * Used only for the "QGTC_ref_lw" variant built by build_ref.py; the unmodified variant
 * "QGTC_ref" is built without this header.
 */
#ifndef UTILITY_H
#define UTILITY_H
typedef unsigned char uin8;
#define STEP8(n)   (((n) + 7) / 8)
#define STEP128(n) (((n) + 127) / 128)
#define PAD8(n)    (8 * STEP8(n))
#define PAD128(n)  (128 * STEP128(n))
#define FLIPBITS(v, nb) ((v) ^ ((1u << (nb)) - 1u))
#define GET_LANEID const unsigned laneid = threadIdx.x & 31u;
#define GET_WARPID const unsigned warpid = threadIdx.x >> 5;
#endif
