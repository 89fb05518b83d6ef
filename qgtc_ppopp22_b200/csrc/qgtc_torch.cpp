// PyTorch extension `QGTC`: the reference's Python op surface (QGTC_module/QGTC_host.cpp:259-271)
// as a thin pybind11 layer over the C ABI in include/qgtc_b200.h.
//
// Same eight functions, same positional-only signatures, same returned shapes and dtypes
// (QGTC_device.cu:83,97,115,223,456,507).  Differences, all deliberate:
//   * outputs are allocated on the tensor's device (torch::empty) and the work is enqueued on the
//     current PyTorch stream; no host synchronisation, no cudaGetDeviceProperties per call;
//   * dtype / size mistakes raise RuntimeError instead of being undefined behaviour, CUDA errors
//     raise instead of exit(-1);
//   * bit2val does not print the tensor;  bitMM2Bit_col implements the intended column-major
//     output (set QGTC_COL_REF_QUIRK=1 in the environment for the shipped kernel's exact bytes).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "qgtc_b200.h"

namespace {

#define CHECK_CUDA(x) TORCH_CHECK((x).is_cuda(), #x " must be a CUDA tensor")
#define CHECK_CONTIGUOUS(x) TORCH_CHECK((x).is_contiguous(), #x " must be contiguous")
#define CHECK_INPUT(x) \
    CHECK_CUDA(x);     \
    CHECK_CONTIGUOUS(x)

void check_status(int rc, const char* op) {
    if (rc == QGTC_OK) return;
    if (rc == QGTC_ERR_CUDA) TORCH_CHECK(false, "QGTC.", op, ": ", qgtc_strerror(rc), ": ", qgtc_last_cuda_error());
    TORCH_CHECK(false, "QGTC.", op, ": ", qgtc_strerror(rc));
}

void* current_stream() { return (void*)at::cuda::getCurrentCUDAStream().stream(); }

torch::Tensor packed_input(const torch::Tensor& t, const char* name) {
    TORCH_CHECK(t.scalar_type() == torch::kInt32, name, " must be an int32 packed tensor");
    if ((reinterpret_cast<uintptr_t>(t.data_ptr()) & 15u) != 0) return t.clone();  // re-align odd views
    return t;
}

int engine_from_env() {
    static const int e = [] {
        const char* s = std::getenv("QGTC_ENGINE");
        if (!s) return (int)QGTC_ENGINE_AUTO;
        if (!std::strcmp(s, "mma")) return (int)QGTC_ENGINE_MMA_SYNC;
        if (!std::strcmp(s, "tcgen05")) return (int)QGTC_ENGINE_TCGEN05;
        if (!std::strcmp(s, "grouped")) return (int)QGTC_ENGINE_GROUPED;
        return (int)QGTC_ENGINE_AUTO;
    }();
    return e;
}

unsigned flags_from_env() {
    static const unsigned f = [] {
        unsigned v = 0;
        const char* z = std::getenv("QGTC_NO_ZERO_JUMP");
        if (z && z[0] == '1') v |= QGTC_FLAG_NO_ZERO_JUMP;
        return v;
    }();
    return f;
}

// cumulative counters, like the reference's __device__ globals (kernel.h:13-14), printed as %d
unsigned long long g_counter = 0, g_counter_global = 0;

torch::Tensor val2bit(torch::Tensor input, const int nbits, const bool col_major, const bool output_layer) {
    CHECK_INPUT(input);
    TORCH_CHECK(input.dim() == 2, "input must be 2-D");
    TORCH_CHECK(input.scalar_type() == torch::kFloat32, "input must be float32");
    const c10::cuda::CUDAGuard guard(input.device());
    const int H = (int)input.size(0), W = (int)input.size(1);
    const int layout = !col_major ? QGTC_LAYOUT_ROW : (output_layer ? QGTC_LAYOUT_COL8 : QGTC_LAYOUT_COL128);
    TORCH_CHECK(nbits >= 1 && nbits <= 8, "nbits must be in [1, 8]");
    TORCH_CHECK(H > 0 && W > 0, "input must be non-empty");
    int64_t r = 0, c = 0;
    qgtc_packed_shape(layout, nbits, H, W, &r, &c);
    auto out = torch::empty({r, c}, input.options().dtype(torch::kInt32));
    check_status(qgtc_val2bit(input.data_ptr<float>(), H, W, nbits, col_major, output_layer, out.data_ptr<int32_t>(),
                              (size_t)out.numel(), current_stream()), "val2bit");
    return out;
}

// extension beyond the reference surface: val2bit of a list of matrices in one call (ROW layout: one kernel launch for the list)
std::vector<torch::Tensor> val2bit_batched(std::vector<torch::Tensor> inputs, const int nbits, const bool col_major, const bool output_layer) {
    TORCH_CHECK(!inputs.empty(), "inputs must be non-empty");
    TORCH_CHECK(nbits >= 1 && nbits <= 8, "nbits must be in [1, 8]");
    const c10::cuda::CUDAGuard guard(inputs[0].device());
    const int layout = !col_major ? QGTC_LAYOUT_ROW : (output_layer ? QGTC_LAYOUT_COL8 : QGTC_LAYOUT_COL128);
    std::vector<qgtc_pack_problem> probs(inputs.size());
    std::vector<torch::Tensor> outs;
    outs.reserve(inputs.size());
    for (size_t i = 0; i < inputs.size(); ++i) {
        const torch::Tensor& x = inputs[i];
        CHECK_INPUT(x);
        TORCH_CHECK(x.dim() == 2 && x.scalar_type() == torch::kFloat32, "inputs must be 2-D float32");
        TORCH_CHECK(x.device() == inputs[0].device(), "inputs must live on one device");
        const int H = (int)x.size(0), W = (int)x.size(1);
        TORCH_CHECK(H > 0 && W > 0, "inputs must be non-empty");
        int64_t r = 0, c = 0;
        qgtc_packed_shape(layout, nbits, H, W, &r, &c);
        outs.push_back(torch::empty({r, c}, x.options().dtype(torch::kInt32)));
        probs[i] = qgtc_pack_problem{x.data_ptr<float>(), outs.back().data_ptr<int32_t>(), (size_t)outs.back().numel(), H, W};
    }
    check_status(qgtc_val2bit_batched(probs.data(), (int)probs.size(), nbits, col_major, output_layer, current_stream()), "val2bit_batched");
    return outs;
}

torch::Tensor bit2val(torch::Tensor input, const int nbits, const int height, const int width, const bool col_major,
                      const bool output_layer) {
    CHECK_INPUT(input);
    const c10::cuda::CUDAGuard guard(input.device());
    auto in = packed_input(input, "input");
    TORCH_CHECK(height > 0 && width > 0, "height and width must be positive");
    auto out = torch::empty({height, width}, input.options().dtype(torch::kInt32));
    check_status(qgtc_bit2val(in.data_ptr<int32_t>(), (size_t)in.numel(), nbits, height, width, col_major, output_layer,
                              out.data_ptr<int32_t>(), current_stream()), "bit2val");
    return out;
}

torch::Tensor mm2bit_impl(const torch::Tensor& X1, const torch::Tensor& X2, int M, int K, int N, int bit1, int bit2, int out_bit,
                          const char* op) {
    CHECK_INPUT(X1);
    CHECK_INPUT(X2);
    const c10::cuda::CUDAGuard guard(X1.device());
    auto a = packed_input(X1, "bit_X1"), b = packed_input(X2, "bit_X2");
    TORCH_CHECK(M > 0 && K > 0 && N > 0, "dimensions must be positive");
    int64_t r = 0, c = 0;
    qgtc_packed_shape(QGTC_LAYOUT_ROW, out_bit, M, N, &r, &c);
    auto out = torch::empty({r, c}, X1.options().dtype(torch::kInt32));
    check_status(qgtc_bitmm2bit(a.data_ptr<int32_t>(), (size_t)a.numel(), b.data_ptr<int32_t>(), (size_t)b.numel(), M, K, N, bit1,
                                bit2, out_bit, out.data_ptr<int32_t>(), (size_t)out.numel(), engine_from_env(), flags_from_env(),
                                current_stream()), op);
    return out;
}

torch::Tensor bitMM2Bit(torch::Tensor bit_X1, torch::Tensor bit_X2, const int X1_height, const int X1_width, const int X2_width,
                        const int bit1, const int bit2, const int output_bit) {
    return mm2bit_impl(bit_X1, bit_X2, X1_height, X1_width, X2_width, bit1, bit2, output_bit, "bitMM2Bit");
}

// 200 repetitions timed with CUDA events and the reference's stdout line (QGTC_device.cu:379-434).
// The repetitions are one batched launch whose 200 problems alias the same operands and output:
// the same work as the reference's 200 back-to-back launches without 200 launch latencies.
torch::Tensor bitMM2Bit_profile(torch::Tensor bit_X1, torch::Tensor bit_X2, const int X1_height, const int X1_width,
                                const int X2_width, const int bit1, const int bit2, const int output_bit) {
    CHECK_INPUT(bit_X1);
    CHECK_INPUT(bit_X2);
    const c10::cuda::CUDAGuard guard(bit_X1.device());
    auto a = packed_input(bit_X1, "bit_X1"), b = packed_input(bit_X2, "bit_X2");
    const int M = X1_height, K = X1_width, N = X2_width, PROF = 200;
    TORCH_CHECK(M > 0 && K > 0 && N > 0, "dimensions must be positive");
    TORCH_CHECK((size_t)a.numel() >= qgtc_packed_words(QGTC_LAYOUT_ROW, bit1, M, K) &&
                    (size_t)b.numel() >= qgtc_packed_words(QGTC_LAYOUT_COL128, bit2, K, N),
                "bitMM2Bit_profile needs well-formed operands (R(bit1,M,K) and C128(bit2,K,N))");
    int64_t r = 0, c = 0;
    qgtc_packed_shape(QGTC_LAYOUT_ROW, output_bit, M, N, &r, &c);
    auto out = torch::empty({r, c}, bit_X1.options().dtype(torch::kInt32));
    cudaStream_t st = (cudaStream_t)current_stream();
    auto run = [&](int count) {
        check_status(qgtc_bitmm2bit_batched(a.data_ptr<int32_t>(), 0, b.data_ptr<int32_t>(), 0, count, M, K, N, bit1, bit2,
                                            output_bit, out.data_ptr<int32_t>(), 0, engine_from_env(), flags_from_env(), st),
                     "bitMM2Bit_profile");
    };
    run(2);  // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    run(PROF);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::printf("X1_height %d, X1_width: %d, X2_width: %d, TFLOPs: %.3f\n", M, K, N,
                2.0 * M * K * N * PROF / (ms / 1e3) / 1e12);
    std::fflush(stdout);
    return out;
}

// base_cnt only advances counter_global (kernel.h:452), zerojump_cnt only counter (kernel.h:587)
void add_counts(const torch::Tensor& X1, int M, int K, int N, int bit1, int bit2, bool global_counter) {
    const c10::cuda::CUDAGuard guard(X1.device());     // the count kernel runs on the tensor's device and its current stream
    auto a = packed_input(X1, "bit_X1");
    uint64_t total = 0, nz = 0;
    check_status(qgtc_tile_counts(a.data_ptr<int32_t>(), (size_t)a.numel(), M, K, N, bit1, bit2, &total, &nz, current_stream()),
                 "tile_counts");
    if (global_counter) g_counter_global += total;
    else g_counter += nz;
}

torch::Tensor bitMM2Bit_base_cnt(torch::Tensor bit_X1, torch::Tensor bit_X2, const int X1_height, const int X1_width,
                                 const int X2_width, const int bit1, const int bit2, const int output_bit) {
    auto out = mm2bit_impl(bit_X1, bit_X2, X1_height, X1_width, X2_width, bit1, bit2, output_bit, "bitMM2Bit_base_cnt");
    add_counts(bit_X1, X1_height, X1_width, X2_width, bit1, bit2, true);
    std::printf("counter_global: %d\n", (int)g_counter_global);  // low 32 bits, as the reference's %d (kernel.h:27)
    std::fflush(stdout);
    return out;
}

torch::Tensor bitMM2Bit_zerojump_cnt(torch::Tensor bit_X1, torch::Tensor bit_X2, const int X1_height, const int X1_width,
                                     const int X2_width, const int bit1, const int bit2, const int output_bit) {
    auto out = mm2bit_impl(bit_X1, bit_X2, X1_height, X1_width, X2_width, bit1, bit2, output_bit, "bitMM2Bit_zerojump_cnt");
    add_counts(bit_X1, X1_height, X1_width, X2_width, bit1, bit2, false);
    std::printf("counter: %d\n", (int)g_counter);
    std::fflush(stdout);
    return out;
}

torch::Tensor bitMM2Bit_col(torch::Tensor bit_X1, torch::Tensor bit_X2, const int X1_height, const int X1_width, const int X2_width,
                            const int bit1, const int bit2, const int output_bit) {
    CHECK_INPUT(bit_X1);
    CHECK_INPUT(bit_X2);
    const c10::cuda::CUDAGuard guard(bit_X1.device());
    auto a = packed_input(bit_X1, "bit_X1"), b = packed_input(bit_X2, "bit_X2");
    TORCH_CHECK(X1_height > 0 && X1_width > 0 && X2_width > 0, "dimensions must be positive");
    int64_t r = 0, c = 0;
    qgtc_packed_shape(QGTC_LAYOUT_COL128, output_bit, X1_height, X2_width, &r, &c);
    auto out = torch::empty({r, c}, bit_X1.options().dtype(torch::kInt32));
    const char* q = std::getenv("QGTC_COL_REF_QUIRK");
    const unsigned flags = flags_from_env() | ((q && q[0] == '1') ? QGTC_FLAG_COL_REF_QUIRK : 0u);
    check_status(qgtc_bitmm2bit_col(a.data_ptr<int32_t>(), (size_t)a.numel(), b.data_ptr<int32_t>(), (size_t)b.numel(), X1_height,
                                    X1_width, X2_width, bit1, bit2, output_bit, out.data_ptr<int32_t>(), (size_t)out.numel(),
                                    engine_from_env(), flags, current_stream()), "bitMM2Bit_col");
    return out;
}

torch::Tensor bitMM2Int(torch::Tensor bit_X1, torch::Tensor bit_X2, const int X1_height, const int X1_width, const int X2_width,
                        const int bit1, const int bit2, const bool pad_128) {
    CHECK_INPUT(bit_X1);
    CHECK_INPUT(bit_X2);
    const c10::cuda::CUDAGuard guard(bit_X1.device());
    auto a = packed_input(bit_X1, "bit_X1"), b = packed_input(bit_X2, "bit_X2");
    TORCH_CHECK(X1_height > 0 && X1_width > 0 && X2_width > 0, "dimensions must be positive");
    auto out = torch::empty({X1_height, X2_width}, bit_X1.options().dtype(torch::kFloat32));
    check_status(qgtc_bitmm2int(a.data_ptr<int32_t>(), (size_t)a.numel(), b.data_ptr<int32_t>(), (size_t)b.numel(), X1_height,
                                X1_width, X2_width, bit1, bit2, pad_128 ? 1 : 0, out.data_ptr<float>(), (size_t)out.numel(),
                                engine_from_env(), flags_from_env(), current_stream()), "bitMM2Int");
    return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// static_operands (grouped ops): bit 0 = every X1 (and its tile bitmap), bit 1 = every X2 was NOT produced by the work queued
// just before this call (adjacency, weights): QGTC_FLAG_X1_STATIC / QGTC_FLAG_X2_STATIC, see include/qgtc_b200.h.
// Extensions beyond the reference's eight functions (keyword arguments allowed here): the grouped launch -- one op of the
// forward for ALL batches in one kernel (include/qgtc_b200.h: qgtc_bitmm_grouped) -- the tile bitmap for zero-tile jumping
// and batch construction from an edge list.
// ---------------------------------------------------------------------------------------------------------------------
std::vector<torch::Tensor> grouped_impl(const std::vector<torch::Tensor>& X1s, const std::vector<torch::Tensor>& X2s,
                                        const std::vector<int64_t>& Ms, const std::vector<int64_t>& Ks, int N, int bit1, int bit2,
                                        int out_bit, int kind, bool pad_128, const std::vector<torch::Tensor>& tile_maps, unsigned static_flags,
                                        const char* op) {
    const size_t count = X1s.size();
    TORCH_CHECK(count > 0 && X2s.size() == count && Ms.size() == count && Ks.size() == count, op, ": operand lists must be non-empty and of equal length");
    TORCH_CHECK(tile_maps.empty() || tile_maps.size() == count, op, ": tile_maps must be empty or hold one bitmap per problem");
    TORCH_CHECK(N > 0, "dimensions must be positive");
    const c10::cuda::CUDAGuard guard(X1s[0].device());
    std::vector<torch::Tensor> keep;          // re-aligned copies (rare) must outlive the launch call
    std::vector<qgtc_group_problem> probs(count);
    std::vector<int64_t> offs(count), sizes(count), rows(count), cols(count);
    int64_t total = 0;
    for (size_t i = 0; i < count; ++i) {
        TORCH_CHECK(Ms[i] > 0 && Ks[i] > 0, "dimensions must be positive");
        if (kind == QGTC_GROUP_FLOAT) { rows[i] = Ms[i]; cols[i] = N; }
        else qgtc_packed_shape(kind == QGTC_GROUP_BITS_ROW ? QGTC_LAYOUT_ROW : QGTC_LAYOUT_COL128, out_bit, (int)Ms[i], N, &rows[i], &cols[i]);
        sizes[i] = rows[i] * cols[i];
        offs[i] = total;
        total += (sizes[i] + 3) & ~(int64_t)3;          // every problem's output stays 16-byte aligned
    }
    auto flat = torch::empty({total}, X1s[0].options().dtype(kind == QGTC_GROUP_FLOAT ? torch::kFloat32 : torch::kInt32));
    std::vector<torch::Tensor> outs(count);
    for (size_t i = 0; i < count; ++i) {
        CHECK_INPUT(X1s[i]);
        CHECK_INPUT(X2s[i]);
        TORCH_CHECK(X1s[i].device() == X1s[0].device() && X2s[i].device() == X1s[0].device(), op, ": all operands must live on one device");
        keep.push_back(packed_input(X1s[i], "bit_X1"));
        keep.push_back(packed_input(X2s[i], "bit_X2"));
        const torch::Tensor &a = keep[keep.size() - 2], &b = keep[keep.size() - 1];
        outs[i] = flat.narrow(0, offs[i], sizes[i]).view({rows[i], cols[i]});
        qgtc_group_problem& p = probs[i];
        p.x1_dev = a.data_ptr<int32_t>();
        p.x1_words = (size_t)a.numel();
        p.x2_dev = b.data_ptr<int32_t>();
        p.x2_words = (size_t)b.numel();
        p.out_dev = outs[i].data_ptr();
        p.out_size = (size_t)sizes[i];
        p.tile_bitmap_dev = nullptr;
        if (!tile_maps.empty() && tile_maps[i].defined() && tile_maps[i].numel() > 0) {
            TORCH_CHECK(tile_maps[i].is_cuda() && tile_maps[i].scalar_type() == torch::kInt32 && tile_maps[i].is_contiguous(), "tile_map must be a contiguous int32 CUDA tensor");
            TORCH_CHECK((size_t)tile_maps[i].numel() >= qgtc_tile_bitmap_words((int)Ms[i], (int)Ks[i], 1), "tile_map too small for its problem");
            p.tile_bitmap_dev = reinterpret_cast<const uint32_t*>(tile_maps[i].data_ptr<int32_t>());
        }
        p.M = (int32_t)Ms[i];
        p.K = (int32_t)Ks[i];
    }
    check_status(qgtc_bitmm_grouped(probs.data(), (int)count, N, bit1, bit2, out_bit, kind, pad_128 ? 1 : 0, engine_from_env(),
                                    flags_from_env() | (static_flags & (QGTC_FLAG_X1_STATIC | QGTC_FLAG_X2_STATIC)), current_stream()), op);
    return outs;
}

std::vector<torch::Tensor> bitMM2Bit_grouped(std::vector<torch::Tensor> X1s, std::vector<torch::Tensor> X2s, std::vector<int64_t> Ms,
                                             std::vector<int64_t> Ks, int N, int bit1, int bit2, int out_bit, std::vector<torch::Tensor> tile_maps,
                                             int static_operands) {
    return grouped_impl(X1s, X2s, Ms, Ks, N, bit1, bit2, out_bit, QGTC_GROUP_BITS_ROW, true, tile_maps, (unsigned)static_operands << 3, "bitMM2Bit_grouped");
}
std::vector<torch::Tensor> bitMM2Bit_col_grouped(std::vector<torch::Tensor> X1s, std::vector<torch::Tensor> X2s, std::vector<int64_t> Ms,
                                                 std::vector<int64_t> Ks, int N, int bit1, int bit2, int out_bit) {
    return grouped_impl(X1s, X2s, Ms, Ks, N, bit1, bit2, out_bit, QGTC_GROUP_BITS_COL, true, {}, 0u, "bitMM2Bit_col_grouped");
}
std::vector<torch::Tensor> bitMM2Int_grouped(std::vector<torch::Tensor> X1s, std::vector<torch::Tensor> X2s, std::vector<int64_t> Ms,
                                             std::vector<int64_t> Ks, int N, int bit1, int bit2, bool pad_128, std::vector<torch::Tensor> tile_maps,
                                             int static_operands) {
    return grouped_impl(X1s, X2s, Ms, Ks, N, bit1, bit2, 1, QGTC_GROUP_FLOAT, pad_128, tile_maps, (unsigned)static_operands << 3, "bitMM2Int_grouped");
}

// Fused aggregation + update for all problems of a list in one launch (qgtc_bitmm_grouped_fused): returns the second op's
// outputs.  Falls back to two grouped launches when the shapes do not fit the fused kernel (identical results).
std::vector<torch::Tensor> fused_impl(const std::vector<torch::Tensor>& X1s, const std::vector<torch::Tensor>& X2s, const std::vector<int64_t>& Ms,
                                      const std::vector<int64_t>& Ks, int N, int bit1, int bit2, int out_bit, const torch::Tensor& W, int N2, int bit_w,
                                      int out_bit2, bool float_out, bool pad_128, const std::vector<torch::Tensor>& tile_maps, unsigned static_flags) {
    const char* op = "bitMM2Bit_grouped_fused";
    const size_t count = X1s.size();
    TORCH_CHECK(count > 0 && X2s.size() == count && Ms.size() == count && Ks.size() == count, op, ": operand lists must be non-empty and of equal length");
    TORCH_CHECK(tile_maps.empty() || tile_maps.size() == count, op, ": tile_maps must be empty or hold one bitmap per problem");
    CHECK_INPUT(W);
    const c10::cuda::CUDAGuard guard(X1s[0].device());
    auto w = packed_input(W, "bit_W");
    std::vector<torch::Tensor> keep;
    std::vector<qgtc_group_problem> probs(count);
    std::vector<void*> out2_ptr(count);
    std::vector<size_t> out2_size(count);
    std::vector<int64_t> offs(count), sizes(count), rows(count), cols(count);
    int64_t total = 0;
    for (size_t i = 0; i < count; ++i) {
        TORCH_CHECK(Ms[i] > 0 && Ks[i] > 0, "dimensions must be positive");
        if (float_out) { rows[i] = Ms[i]; cols[i] = N2; }
        else qgtc_packed_shape(QGTC_LAYOUT_ROW, out_bit2, (int)Ms[i], N2, &rows[i], &cols[i]);
        sizes[i] = rows[i] * cols[i];
        offs[i] = total;
        total += (sizes[i] + 3) & ~(int64_t)3;
    }
    auto flat = torch::empty({total}, X1s[0].options().dtype(float_out ? torch::kFloat32 : torch::kInt32));
    std::vector<torch::Tensor> outs(count);
    for (size_t i = 0; i < count; ++i) {
        CHECK_INPUT(X1s[i]);
        CHECK_INPUT(X2s[i]);
        keep.push_back(packed_input(X1s[i], "bit_X1"));
        keep.push_back(packed_input(X2s[i], "bit_X2"));
        const torch::Tensor &a = keep[keep.size() - 2], &b = keep[keep.size() - 1];
        outs[i] = flat.narrow(0, offs[i], sizes[i]).view({rows[i], cols[i]});
        qgtc_group_problem& p = probs[i];
        p.x1_dev = a.data_ptr<int32_t>();
        p.x1_words = (size_t)a.numel();
        p.x2_dev = b.data_ptr<int32_t>();
        p.x2_words = (size_t)b.numel();
        p.out_dev = nullptr;
        p.out_size = 0;
        p.tile_bitmap_dev = nullptr;
        if (!tile_maps.empty() && tile_maps[i].defined() && tile_maps[i].numel() > 0) {
            TORCH_CHECK(tile_maps[i].is_cuda() && tile_maps[i].scalar_type() == torch::kInt32 && tile_maps[i].is_contiguous(), "tile_map must be a contiguous int32 CUDA tensor");
            TORCH_CHECK((size_t)tile_maps[i].numel() >= qgtc_tile_bitmap_words((int)Ms[i], (int)Ks[i], 1), "tile_map too small for its problem");
            p.tile_bitmap_dev = reinterpret_cast<const uint32_t*>(tile_maps[i].data_ptr<int32_t>());
        }
        p.M = (int32_t)Ms[i];
        p.K = (int32_t)Ks[i];
        out2_ptr[i] = outs[i].data_ptr();
        out2_size[i] = (size_t)sizes[i];
    }
    const qgtc_fused_update upd{w.data_ptr<int32_t>(), (size_t)w.numel(), N2, bit_w, out_bit2, float_out ? QGTC_GROUP_FLOAT : QGTC_GROUP_BITS_ROW, pad_128 ? 1 : 0};
    const unsigned fl = flags_from_env() | (static_flags & (QGTC_FLAG_X1_STATIC | QGTC_FLAG_X2_STATIC));
    const int rc = qgtc_bitmm_grouped_fused(probs.data(), (int)count, N, bit1, bit2, out_bit, &upd, out2_ptr.data(), out2_size.data(), 0, engine_from_env(), fl,
                                            current_stream());
    if (rc == QGTC_ERR_UNSUPPORTED) {      // shapes outside the fused kernel: the same two ops, one grouped launch each
        auto mid = grouped_impl(X1s, X2s, Ms, Ks, N, bit1, bit2, out_bit, QGTC_GROUP_BITS_ROW, true, tile_maps, static_flags, op);
        std::vector<torch::Tensor> Ws(count, W);
        std::vector<int64_t> K2(count, (int64_t)N);
        return grouped_impl(mid, Ws, Ms, K2, N2, out_bit, bit_w, float_out ? 1 : out_bit2, float_out ? QGTC_GROUP_FLOAT : QGTC_GROUP_BITS_ROW, pad_128, {},
                            QGTC_FLAG_X2_STATIC, op);
    }
    check_status(rc, op);
    return outs;
}

std::vector<torch::Tensor> bitMM2Bit_grouped_fused(std::vector<torch::Tensor> X1s, std::vector<torch::Tensor> X2s, std::vector<int64_t> Ms, std::vector<int64_t> Ks,
                                                   int N, int bit1, int bit2, int out_bit, torch::Tensor W, int N2, int bit_w, int out_bit2, bool float_out,
                                                   bool pad_128, std::vector<torch::Tensor> tile_maps, int static_operands) {
    return fused_impl(X1s, X2s, Ms, Ks, N, bit1, bit2, out_bit, W, N2, bit_w, out_bit2, float_out, pad_128, tile_maps, (unsigned)static_operands << 3);
}

torch::Tensor tile_bitmap(torch::Tensor bit_X1, const int M, const int K) {
    CHECK_INPUT(bit_X1);
    const c10::cuda::CUDAGuard guard(bit_X1.device());
    auto a = packed_input(bit_X1, "bit_X1");
    TORCH_CHECK(M > 0 && K > 0, "dimensions must be positive");
    const size_t words = qgtc_tile_bitmap_words(M, K, 1);
    auto bm = torch::empty({(int64_t)words}, bit_X1.options().dtype(torch::kInt32));
    check_status(qgtc_tile_bitmap(a.data_ptr<int32_t>(), (size_t)a.numel(), 0, M, K, 1, reinterpret_cast<uint32_t*>(bm.data_ptr<int32_t>()), words,
                                  current_stream()), "tile_bitmap");
    return bm;
}

torch::Tensor edges2bit(torch::Tensor rows, torch::Tensor cols, const int H, const int W) {
    CHECK_INPUT(rows);
    CHECK_INPUT(cols);
    TORCH_CHECK(rows.scalar_type() == torch::kInt32 && cols.scalar_type() == torch::kInt32 && rows.numel() == cols.numel(), "rows / cols must be int32 tensors of equal length");
    const c10::cuda::CUDAGuard guard(rows.device());
    int64_t r = 0, c = 0;
    qgtc_packed_shape(QGTC_LAYOUT_ROW, 1, H, W, &r, &c);
    auto out = torch::empty({r, c}, rows.options().dtype(torch::kInt32));
    check_status(qgtc_edges2bit(rows.data_ptr<int32_t>(), cols.data_ptr<int32_t>(), (size_t)rows.numel(), H, W, out.data_ptr<int32_t>(), (size_t)out.numel(),
                                current_stream()), "edges2bit");
    return out;
}

}  // namespace

// Same names, order and docstrings as the reference binding (QGTC_host.cpp:259-271); no py::arg,
// so every argument is positional and required, as there.
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("val2bit", &val2bit, "quantize a [ int32 --> bit ] tensor (CUDA)");
    m.def("bit2val", &bit2val, "quantize a [ bit --> int32 ] tensor (CUDA)");
    m.def("bitMM2Bit", &bitMM2Bit, "QGTC [ bit_A x bit_B --> bit_C ] forward (CUDA)");
    m.def("bitMM2Bit_profile", &bitMM2Bit_profile, "QGTC [ bit_A x bit_B --> bit_C ] forward (CUDA) for profiling (200)");
    m.def("bitMM2Bit_base_cnt", &bitMM2Bit_base_cnt, "QGTC [ bit_A x bit_B --> bit_C ] forward (CUDA) base count");
    m.def("bitMM2Bit_zerojump_cnt", &bitMM2Bit_zerojump_cnt, "QGTC [ bit_A x bit_B --> bit_C ] forward (CUDA) zero jump count");
    m.def("bitMM2Bit_col", &bitMM2Bit_col, "QGTC [ bit_A x bit_B --> bit_C (column major) ] forward (CUDA)");
    m.def("bitMM2Int", &bitMM2Int, "QGTC [ bit_A x bit_B --> float32 ] forward (CUDA)");
    // extensions (not in the reference): grouped launches, tile bitmap, edge-list packing
    namespace py = pybind11;
    m.def("bitMM2Bit_grouped", &bitMM2Bit_grouped, "one launch: bitMM2Bit of every problem of a list (same N and bit widths)", py::arg("X1s"),
          py::arg("X2s"), py::arg("Ms"), py::arg("Ks"), py::arg("N"), py::arg("bit1"), py::arg("bit2"), py::arg("out_bit"),
          py::arg("tile_maps") = std::vector<torch::Tensor>(), py::arg("static_operands") = 0);
    m.def("bitMM2Bit_col_grouped", &bitMM2Bit_col_grouped, "one launch: bitMM2Bit_col (column-major output) of every problem of a list", py::arg("X1s"),
          py::arg("X2s"), py::arg("Ms"), py::arg("Ks"), py::arg("N"), py::arg("bit1"), py::arg("bit2"), py::arg("out_bit"));
    m.def("bitMM2Int_grouped", &bitMM2Int_grouped, "one launch: bitMM2Int of every problem of a list (same N and bit widths)", py::arg("X1s"),
          py::arg("X2s"), py::arg("Ms"), py::arg("Ks"), py::arg("N"), py::arg("bit1"), py::arg("bit2"), py::arg("pad_128"),
          py::arg("tile_maps") = std::vector<torch::Tensor>(), py::arg("static_operands") = 0);
    m.def("bitMM2Bit_grouped_fused", &bitMM2Bit_grouped_fused,
          "one launch: out2[i] = op2(bitMM2Bit(X1s[i], X2s[i]), W) with op2 = bitMM2Bit (float_out = False) or bitMM2Int (True) and one shared W",
          py::arg("X1s"), py::arg("X2s"), py::arg("Ms"), py::arg("Ks"), py::arg("N"), py::arg("bit1"), py::arg("bit2"), py::arg("out_bit"), py::arg("W"),
          py::arg("N2"), py::arg("bit_w"), py::arg("out_bit2"), py::arg("float_out") = false, py::arg("pad_128") = false,
          py::arg("tile_maps") = std::vector<torch::Tensor>(), py::arg("static_operands") = 0);
    m.def("val2bit_batched", &val2bit_batched, "val2bit of a list of float32 matrices in one call (ROW layout: one launch for the whole list)");
    m.def("tile_bitmap", &tile_bitmap, "non-zero 128 x 128-bit tile bitmap of a packed 1-bit adjacency (zero-tile jumping)");
    m.def("edges2bit", &edges2bit, "packed 1-bit adjacency R(1,H,W) from int32 edge endpoints (no dense fp32 intermediate)");
    m.attr("__backend__") = "qgtc_b200";
}
