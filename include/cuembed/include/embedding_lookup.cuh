// Drop-in replacement for cuembed/include/embedding_lookup.cuh
// (NVIDIA/cuEmbed @ 90dd8436): the same host templates in namespace cuembed,
// forwarding to the B200-native kernels in libcuembed_b200.so through the C
// ABI of include/cuembed_b200.h.  Callers keep
//     #include "cuembed/include/embedding_lookup.cuh"
// put <repo>/include on the include path and link -lcuembed_b200.
//
//   EmbeddingForward   replaces cuembed/include/embedding_lookup.cuh:245-308
//   EmbeddingBackward  replaces cuembed/include/embedding_lookup.cuh:423-483
// Misuse aborts with a "Check failed" message like CUEMBED_ASSERT (:151-158).
#ifndef CUEMBED_INCLUDE_EMBEDDING_LOOKUP_CUH_
#define CUEMBED_INCLUDE_EMBEDDING_LOOKUP_CUH_

#include <cuda_runtime.h>

#include <cstdlib>
#include <iostream>

#include "cuembed/include/embedding_lookup_types.cuh"
#include "cuembed_b200.h"

namespace cuembed {

#define CUEMBED_ASSERT(condition)                                           \
  do {                                                                      \
    if (!(condition)) {                                                     \
      std::cerr << "Check failed: " #condition << " at " << __FILE__ << ":" \
                << __LINE__ << std::endl;                                   \
      std::abort();                                                         \
    }                                                                       \
  } while (0)

namespace b200_detail {
// A non-zero C-ABI code is a violated precondition: abort like the reference.
inline void CheckCode(int code, const char* what) {
  if (code != CUEMBED_OK) {
    std::cerr << what << ": " << cuembed_error_string(code) << std::endl;
    std::abort();
  }
}
inline int ModeCode(CombineMode mode) {
  return mode == CombineMode::kSum
             ? CUEMBED_SUM
             : (mode == CombineMode::kMean ? CUEMBED_MEAN : CUEMBED_CONCAT);
}
}  // namespace b200_detail

// Pooled embedding lookup, fixed hotness or CSR.  Template parameters and
// arguments as in the reference; bf16 tables are additionally supported.
template <typename InputT, typename OutputT, typename IndexT, typename OffsetT,
          bool fp16_math = false>
void EmbeddingForward(const InputT* params, const int embed_width,
                      const IndexT* indices, const OffsetT* offsets,
                      const GetElemT<InputT>* weights, const int batch_size,
                      const int num_hots, const CombineMode mode, OutputT* ret,
                      const cudaStream_t stream = 0) {
  using ElemT = GetElemT<InputT>;
  if constexpr (b200_detail::IsMappedTable<InputT>::value) {
    // structured InputT: `params` is a host pointer to one MappedTable
    CUEMBED_ASSERT(params != nullptr && params->row_map != nullptr);
    CUEMBED_ASSERT(!fp16_math && mode != CombineMode::kConcat);
    b200_detail::CheckCode(
        cuembed_forward_mapped(
            params->rows, b200_detail::DTypeCode<ElemT>::value, embed_width,
            indices, b200_detail::ITypeCode<IndexT>::value, offsets,
            b200_detail::ITypeCode<OffsetT>::value, weights, batch_size,
            num_hots, b200_detail::ModeCode(mode), ret,
            b200_detail::DTypeCode<GetElemT<OutputT>>::value, params->row_map,
            params->cache, reinterpret_cast<cuembed_stream_t>(stream)),
        "EmbeddingForward<MappedTable>");
    return;
  } else {
  b200_detail::CheckCode(
      cuembed_forward(params, b200_detail::DTypeCode<ElemT>::value, embed_width,
                      indices, b200_detail::ITypeCode<IndexT>::value, offsets,
                      b200_detail::ITypeCode<OffsetT>::value, weights,
                      batch_size, num_hots, b200_detail::ModeCode(mode),
                      fp16_math ? 1 : 0, ret,
                      b200_detail::DTypeCode<GetElemT<OutputT>>::value,
                      reinterpret_cast<cuembed_stream_t>(stream)),
      "EmbeddingForward");
  }
}

// Debug aid (new): validates indices / offsets on the device and aborts with
// the first offending position; synchronises the stream.  The kernels carry
// no bounds checks, like the reference's (embedding_lookup_ops.cuh:59).
template <typename IndexT, typename OffsetT>
void DebugCheckLookup(const IndexT* indices, const long long nnz,
                      const long long num_rows, const OffsetT* offsets,
                      const int batch_size, const cudaStream_t stream = 0) {
  long long first = -1;
  const int code = cuembed_debug_check_lookup(
      indices, b200_detail::ITypeCode<IndexT>::value, nnz, num_rows, offsets,
      b200_detail::ITypeCode<OffsetT>::value, batch_size, &first,
      reinterpret_cast<cuembed_stream_t>(stream));
  if (code != CUEMBED_OK) {
    std::cerr << "DebugCheckLookup: " << cuembed_error_string(code)
              << " (first offending position " << first << ")" << std::endl;
    std::abort();
  }
}

// Gradient w.r.t. the table from the transposed COO indices; full or
// compressed.  Deterministic (fixed summation order, no atomics).
template <typename GradT, typename IndexT>
void EmbeddingBackward(const GradT* grad_y, const int embed_width,
                       const int num_grad_embedding_rows, const int nnz,
                       const IndexT* transpose_indices,
                       const IndexT* transpose_sample_ids,
                       const IndexT* transpose_remapped_indices,
                       const GradT* transpose_weights,
                       const bool skip_grad_init, GradT* grad_embedding,
                       IndexT* inverse_mapping,
                       const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_backward(grad_y, b200_detail::DTypeCode<GradT>::value,
                       embed_width, num_grad_embedding_rows, nnz,
                       b200_detail::ITypeCode<IndexT>::value,
                       transpose_indices, transpose_sample_ids,
                       transpose_remapped_indices, transpose_weights,
                       skip_grad_init ? 1 : 0, grad_embedding, inverse_mapping,
                       reinterpret_cast<cuembed_stream_t>(stream)),
      "EmbeddingBackward");
}

// ---- additions of the B200 build (no counterpart in the reference; its README
// lists "optimizer" kernels and "multiple tables" as future work, :110,119) ----

enum class SparseOptimizer { kSgd = CUEMBED_OPT_SGD, kAdagrad = CUEMBED_OPT_ADAGRAD };

// Backward fused with a sparse optimizer step on the touched rows of `params`
// (include/cuembed_b200.h: cuembed_backward_update).  transpose_indices are
// table rows; `state` is the fp32 Adagrad accumulator (nullptr for SGD).
// Scratch through the reference's work / lwork two-call protocol.
template <typename GradT, typename IndexT>
void EmbeddingBackwardUpdate(const GradT* grad_y, const int embed_width,
                             const int nnz, const IndexT* transpose_indices,
                             const IndexT* transpose_sample_ids,
                             const GradT* transpose_weights,
                             const SparseOptimizer optimizer, const float lr,
                             const float eps, GradT* params, float* state,
                             char* work, size_t* lwork,
                             const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_backward_update(grad_y, b200_detail::DTypeCode<GradT>::value,
                              embed_width, nnz,
                              b200_detail::ITypeCode<IndexT>::value,
                              transpose_indices, transpose_sample_ids,
                              transpose_weights, static_cast<int>(optimizer),
                              lr, eps, params, state, work, lwork,
                              reinterpret_cast<cuembed_stream_t>(stream)),
      "EmbeddingBackwardUpdate");
}

// Pooled lookups into `num_tables` tables of one row shape in one launch
// (cuembed_forward_multi).  HOST arrays of per-table device pointers / sizes;
// offsets / weights / modes may be nullptr (fixed hotness / unweighted / sum).
template <typename InputT, typename OutputT, typename IndexT, typename OffsetT>
void EmbeddingForwardMulti(const int num_tables, const InputT* const* params,
                           const int embed_width, const IndexT* const* indices,
                           const OffsetT* const* offsets,
                           const GetElemT<InputT>* const* weights,
                           const int* batch_sizes, const int* num_hots,
                           const CombineMode* modes, OutputT* const* rets,
                           const long long out_row_stride = 0,
                           const cudaStream_t stream = 0) {
  int mode_codes[64];
  int* mc = nullptr;
  if (modes != nullptr) {
    CUEMBED_ASSERT(num_tables <= 64);
    for (int t = 0; t < num_tables; ++t)
      mode_codes[t] = b200_detail::ModeCode(modes[t]);
    mc = mode_codes;
  }
  b200_detail::CheckCode(
      cuembed_forward_multi(
          num_tables, reinterpret_cast<const void* const*>(params),
          b200_detail::DTypeCode<GetElemT<InputT>>::value, embed_width,
          reinterpret_cast<const void* const*>(indices),
          b200_detail::ITypeCode<IndexT>::value,
          reinterpret_cast<const void* const*>(offsets),
          b200_detail::ITypeCode<OffsetT>::value,
          reinterpret_cast<const void* const*>(weights), batch_sizes, num_hots,
          mc, reinterpret_cast<void* const*>(rets),
          b200_detail::DTypeCode<GetElemT<OutputT>>::value, out_row_stride,
          reinterpret_cast<cuembed_stream_t>(stream)),
      "EmbeddingForwardMulti");
}

}  // namespace cuembed

#endif  // CUEMBED_INCLUDE_EMBEDDING_LOOKUP_CUH_
