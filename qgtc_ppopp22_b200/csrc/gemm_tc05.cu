// Bit-GEMM engine 2: tcgen05.mma kind::i8 with TMEM accumulators (placeholder until the engine lands).
#include "gemm_common.cuh"

namespace qgtc {

bool tc05_supports(const GemmArgs&, int) { return false; }
int launch_gemm_tc05(const GemmArgs&, int, cudaStream_t) { return QGTC_ERR_UNSUPPORTED; }

}  // namespace qgtc
