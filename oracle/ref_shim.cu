// ref_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Type-erased extern "C" entry points over the reference's OWN CPU templates,
// compiled unchanged from where they lie under /root/reference:
//   utils/include/embedding_lookup_cpu.hpp  (EmbeddingForwardCpu :35-94,
//                                            EmbeddingBackwardCpu :96-144)
//   utils/include/index_transforms_cpu.hpp  (ExtractRowIds*Cpu :35-64,
//       ComputeCompressedGradIndicesCpu :66-77, TransposeCpu :86-125)
// Built by oracle/Makefile into oracle/_ref/libcuembed_ref.so (git-ignored,
// travels to the GPU box).  No reference source is copied into this repo: the
// headers are reached through -I/root/reference at build time; absl's CHECK is
// replaced by oracle/shim/absl/log/{check,log}.h.
//
// The signatures are identical to oracle/cuembed_oracle.c so tests and the
// bench can swap one for the other.  bfloat16 is NOT supported by the
// reference (README.md:112,118); the bf16 instantiations below exist only
// because the templates happen to compile with one operator overload, and are
// reported as "unpinned by reference tests".
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstring>

#include "cuembed/include/embedding_lookup_types.cuh"

namespace cuembed {
// The single ambiguity that keeps EmbeddingForwardCpu<__nv_bfloat16,...> from
// compiling (float * bf16 at utils/include/embedding_lookup_cpu.hpp:75).
inline float operator*(const float& lhs, const __nv_bfloat16 rhs) {
  return lhs * __bfloat162float(rhs);
}
}  // namespace cuembed

#include "utils/include/embedding_lookup_cpu.hpp"
#include "utils/include/index_transforms_cpu.hpp"

namespace {

using cuembed::CombineMode;

CombineMode ToMode(int mode) {
  return mode == 0 ? CombineMode::kSum
                   : (mode == 1 ? CombineMode::kMean : CombineMode::kConcat);
}

template <typename InT, typename OutT, typename IndexT, typename OffsetT>
int ForwardTyped(const void* params, int embed_width, int batch_size,
                 int num_hots, const void* indices, const void* offsets,
                 const void* weights, void* ret, int mode, int fp16_math,
                 int sample_begin, int sample_end) {
  if (sample_end > batch_size) sample_end = batch_size;
  const int n = sample_end - sample_begin;
  if (n <= 0) return 0;
  const InT* p = static_cast<const InT*>(params);
  const IndexT* idx = static_cast<const IndexT*>(indices);
  const OffsetT* off = static_cast<const OffsetT*>(offsets);
  const InT* w = static_cast<const InT*>(weights);
  OutT* out = static_cast<OutT*>(ret);
  // Batch slicing by pointer offsets only (no reference change).
  if (off != nullptr) {
    off += sample_begin;  // offsets stay absolute into indices / weights
    out += static_cast<int64_t>(sample_begin) * embed_width;
  } else {
    idx += static_cast<int64_t>(sample_begin) * num_hots;
    if (w != nullptr) w += static_cast<int64_t>(sample_begin) * num_hots;
    const int64_t per_sample =
        (mode == 2) ? static_cast<int64_t>(num_hots) * embed_width
                    : embed_width;
    out += static_cast<int64_t>(sample_begin) * per_sample;
  }
  if (off != nullptr && mode == 2) return -3;
  if (fp16_math) {
    cuembed::EmbeddingForwardCpu<InT, OutT, IndexT, OffsetT, true>(
        p, embed_width, n, num_hots, idx, off, w, out, ToMode(mode));
  } else {
    cuembed::EmbeddingForwardCpu<InT, OutT, IndexT, OffsetT, false>(
        p, embed_width, n, num_hots, idx, off, w, out, ToMode(mode));
  }
  return 0;
}

template <typename InT, typename OutT>
int ForwardIdx(int it, int ot, const void* params, int embed_width,
               int batch_size, int num_hots, const void* indices,
               const void* offsets, const void* weights, void* ret, int mode,
               int fp16_math, int sb, int se) {
#define CALL(I, O)                                                           \
  return ForwardTyped<InT, OutT, I, O>(params, embed_width, batch_size,      \
                                       num_hots, indices, offsets, weights,  \
                                       ret, mode, fp16_math, sb, se)
  if (it == 0 && ot == 0) CALL(int32_t, int32_t);
  if (it == 0 && ot == 1) CALL(int32_t, int64_t);
  if (it == 1 && ot == 0) CALL(int64_t, int32_t);
  CALL(int64_t, int64_t);
#undef CALL
}

template <typename GradT, typename IndexT>
int BackwardTyped(const void* grad_y, int embed_width, int rows, int nnz,
                  const void* t_indices, const void* t_sids,
                  const void* t_remapped, const void* t_weights,
                  int skip_grad_init, void* grad_embedding,
                  void* inverse_mapping, int nz_begin, int nz_end) {
  if (nz_end > nnz) nz_end = nnz;
  const int n = nz_end - nz_begin;
  if (n <= 0) return 0;
  const IndexT* ti = static_cast<const IndexT*>(t_indices) + nz_begin;
  const IndexT* ts = static_cast<const IndexT*>(t_sids) + nz_begin;
  const IndexT* tr = static_cast<const IndexT*>(t_remapped);
  const GradT* tw = static_cast<const GradT*>(t_weights);
  IndexT* inv = static_cast<IndexT*>(inverse_mapping);
  if (tr != nullptr) {
    // Slices must start at a run boundary; the slice's first compressed row
    // is where its part of inverse_mapping starts.
    if (inv != nullptr) inv += tr[nz_begin];
    tr += nz_begin;
  }
  if (tw != nullptr) tw += nz_begin;
  // A slice must not re-zero rows other slices own: the caller zeroes once
  // (slice 0 with skip_grad_init == 0 zeroes the whole buffer first).
  bool skip = skip_grad_init != 0;
  if (!skip && nz_begin != 0) skip = true;
  cuembed::EmbeddingBackwardCpu<GradT, IndexT>(
      static_cast<const GradT*>(grad_y), embed_width, rows, n, ti, ts, tr, tw,
      skip, static_cast<GradT*>(grad_embedding), inv);
  return 0;
}

template <typename IndexT, typename WeightT>
void TransposeTyped(const void* rows, const void* cols, const void* weights,
                    int nnz, void* t_rows, void* t_cols, void* t_weights) {
  cuembed::TransposeCpu<IndexT, WeightT>(
      static_cast<const IndexT*>(rows), static_cast<const IndexT*>(cols),
      static_cast<const WeightT*>(weights), nnz, static_cast<IndexT*>(t_rows),
      static_cast<IndexT*>(t_cols), static_cast<WeightT*>(t_weights));
}

}  // namespace

extern "C" {

int ref_forward(const void* params, int in_dt, int embed_width,
                int batch_size, int num_hots, const void* indices, int it,
                const void* offsets, int ot, const void* weights, void* ret,
                int out_dt, int mode, int fp16_math, int sample_begin,
                int sample_end) {
#define ARGS                                                               \
  it, ot, params, embed_width, batch_size, num_hots, indices, offsets,     \
      weights, ret, mode, fp16_math, sample_begin, sample_end
  if (in_dt == 0 && out_dt == 0) return ForwardIdx<float, float>(ARGS);
  if (in_dt == 1 && out_dt == 1) return ForwardIdx<__half, __half>(ARGS);
  if (in_dt == 1 && out_dt == 0) return ForwardIdx<__half, float>(ARGS);
  if (in_dt == 2 && out_dt == 2)
    return ForwardIdx<__nv_bfloat16, __nv_bfloat16>(ARGS);
#undef ARGS
  return -100;  // combination not instantiated
}

void ref_extract_row_ids_fixed(int batch_size, int num_hots, void* row_ids,
                               int it) {
  if (it)
    cuembed::ExtractRowIdsFromFixedCpu<int64_t>(
        batch_size, num_hots, static_cast<int64_t*>(row_ids));
  else
    cuembed::ExtractRowIdsFromFixedCpu<int32_t>(
        batch_size, num_hots, static_cast<int32_t*>(row_ids));
}

void ref_extract_row_ids_csr(const void* offsets, int ot, int batch_size,
                             void* row_ids, int it) {
  if (it == 0 && ot == 0)
    cuembed::ExtractRowIdsFromCSRCpu<int32_t, int32_t>(
        static_cast<const int32_t*>(offsets), batch_size,
        static_cast<int32_t*>(row_ids));
  else if (it == 0 && ot == 1)
    cuembed::ExtractRowIdsFromCSRCpu<int32_t, int64_t>(
        static_cast<const int64_t*>(offsets), batch_size,
        static_cast<int32_t*>(row_ids));
  else if (it == 1 && ot == 0)
    cuembed::ExtractRowIdsFromCSRCpu<int64_t, int32_t>(
        static_cast<const int32_t*>(offsets), batch_size,
        static_cast<int64_t*>(row_ids));
  else
    cuembed::ExtractRowIdsFromCSRCpu<int64_t, int64_t>(
        static_cast<const int64_t*>(offsets), batch_size,
        static_cast<int64_t*>(row_ids));
}

void ref_extract_row_ids_concat(int nnz, void* row_ids, int it) {
  if (it)
    cuembed::ExtractRowIdsForConcatCpu<int64_t>(
        nnz, static_cast<int64_t*>(row_ids));
  else
    cuembed::ExtractRowIdsForConcatCpu<int32_t>(
        nnz, static_cast<int32_t*>(row_ids));
}

void ref_compressed_grad_indices(const void* indices, int it, int nnz,
                                 void* remapped) {
  if (it)
    cuembed::ComputeCompressedGradIndicesCpu<int64_t>(
        static_cast<const int64_t*>(indices), nnz,
        static_cast<int64_t*>(remapped));
  else
    cuembed::ComputeCompressedGradIndicesCpu<int32_t>(
        static_cast<const int32_t*>(indices), nnz,
        static_cast<int32_t*>(remapped));
}

int ref_transpose(const void* rows, const void* cols, const void* weights,
                  int wdt, int nnz, int it, void* t_rows, void* t_cols,
                  void* t_weights) {
#define T(I, W) \
  TransposeTyped<I, W>(rows, cols, weights, nnz, t_rows, t_cols, t_weights)
  if (it == 0 && wdt == 0) {
    T(int32_t, float);
  } else if (it == 0 && wdt == 1) {
    T(int32_t, __half);
  } else if (it == 1 && wdt == 0) {
    T(int64_t, float);
  } else if (it == 1 && wdt == 1) {
    T(int64_t, __half);
  } else if (it == 0 && wdt == 2) {
    T(int32_t, __nv_bfloat16);
  } else if (it == 1 && wdt == 2) {
    T(int64_t, __nv_bfloat16);
  } else {
    return -100;
  }
#undef T
  return 0;
}

int ref_backward(const void* grad_y, int dt, int embed_width,
                 int num_grad_embedding_rows, int nnz, int it,
                 const void* t_indices, const void* t_sample_ids,
                 const void* t_remapped, const void* t_weights,
                 int skip_grad_init, void* grad_embedding,
                 void* inverse_mapping, int acc_f32, int nz_begin,
                 int nz_end) {
  if (acc_f32) return -101;  // not a reference behaviour
#define B(G, I)                                                              \
  return BackwardTyped<G, I>(grad_y, embed_width, num_grad_embedding_rows,   \
                             nnz, t_indices, t_sample_ids, t_remapped,       \
                             t_weights, skip_grad_init, grad_embedding,      \
                             inverse_mapping, nz_begin, nz_end)
  if (dt == 0 && it == 0) B(float, int32_t);
  if (dt == 0 && it == 1) B(float, int64_t);
  if (dt == 1 && it == 0) B(__half, int32_t);
  if (dt == 1 && it == 1) B(__half, int64_t);
  if (dt == 2 && it == 0) B(__nv_bfloat16, int32_t);
  if (dt == 2 && it == 1) B(__nv_bfloat16, int64_t);
#undef B
  return -100;
}

}  // extern "C"
