// Drop-in replacement for cuembed/include/index_transforms.cuh
// (NVIDIA/cuEmbed @ 90dd8436): same host templates, B200-native kernels behind
// the C ABI (include/cuembed_b200.h).  No CUB.
//
//   ExtractRowIdsFromFixed / FromCSR / ForConcat   replace :45-93
//   Transpose                                       replaces :224-250
//   ComputeCompressedGradIndices                    replaces :278-323
// Workspace protocol as in the reference: call with work == nullptr to get the
// size in *lwork (:121-124,173-176,300-303).
#ifndef CUEMBED_INCLUDE_INDEX_TRANSFORMS_CUH_
#define CUEMBED_INCLUDE_INDEX_TRANSFORMS_CUH_

#include <cstddef>

#include "cuembed/include/embedding_lookup.cuh"

namespace cuembed {

template <typename IndexT>
void ExtractRowIdsFromFixed(const int batch_size, const int num_hots,
                            IndexT* row_ids, const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_extract_row_ids_fixed(batch_size, num_hots, row_ids,
                                    b200_detail::ITypeCode<IndexT>::value,
                                    reinterpret_cast<cuembed_stream_t>(stream)),
      "ExtractRowIdsFromFixed");
}

template <typename IndexT, typename OffsetT>
void ExtractRowIdsFromCSR(const OffsetT* offsets, const int batch_size,
                          IndexT* row_ids, const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_extract_row_ids_csr(offsets,
                                  b200_detail::ITypeCode<OffsetT>::value,
                                  batch_size, row_ids,
                                  b200_detail::ITypeCode<IndexT>::value,
                                  reinterpret_cast<cuembed_stream_t>(stream)),
      "ExtractRowIdsFromCSR");
}

template <typename IndexT>
void ExtractRowIdsForConcat(const int nnz, IndexT* row_ids,
                            const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_extract_row_ids_concat(nnz, row_ids,
                                     b200_detail::ITypeCode<IndexT>::value,
                                     reinterpret_cast<cuembed_stream_t>(stream)),
      "ExtractRowIdsForConcat");
}

// COO transpose: stable sort of (rows, cols, weights) by cols.
template <typename IndexT, typename WeightT>
void Transpose(const IndexT* rows, const IndexT* cols, const WeightT* weights,
               const int nnz, IndexT* transpose_rows, IndexT* transpose_cols,
               WeightT* transpose_weights, char* work, size_t* lwork,
               const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_transpose(rows, cols, weights,
                        b200_detail::DTypeCode<WeightT>::value, nnz,
                        b200_detail::ITypeCode<IndexT>::value, transpose_rows,
                        transpose_cols, transpose_weights, work, lwork,
                        reinterpret_cast<cuembed_stream_t>(stream)),
      "Transpose");
}

// The reference also exposes the two halves of Transpose (:95-200).
template <typename IndexT>
void TransposeUnweighted(const IndexT* rows, const IndexT* cols, const int nnz,
                         IndexT* transpose_rows, IndexT* transpose_cols,
                         char* work, size_t* lwork,
                         const cudaStream_t stream = 0) {
  Transpose<IndexT, float>(rows, cols, nullptr, nnz, transpose_rows,
                           transpose_cols, nullptr, work, lwork, stream);
}

template <typename IndexT, typename WeightT>
void TransposeWeighted(const IndexT* rows, const IndexT* cols,
                       const WeightT* weights, const int nnz,
                       IndexT* transpose_rows, IndexT* transpose_cols,
                       WeightT* transpose_weights, char* work, size_t* lwork,
                       const cudaStream_t stream = 0) {
  Transpose<IndexT, WeightT>(rows, cols, weights, nnz, transpose_rows,
                             transpose_cols, transpose_weights, work, lwork,
                             stream);
}

// Extension (not in the reference): ExtractRowIdsFromFixed + Transpose in one
// call for fixed-hotness indices; the row-id array is never materialised.
template <typename IndexT, typename WeightT>
void TransposeFromFixed(const IndexT* cols, const WeightT* weights,
                        const int batch_size, const int num_hots,
                        IndexT* transpose_rows, IndexT* transpose_cols,
                        WeightT* transpose_weights, char* work, size_t* lwork,
                        const cudaStream_t stream = 0) {
  if (work == nullptr) {  // workspace query: same size as Transpose
    Transpose<IndexT, WeightT>(nullptr, cols, weights, batch_size * num_hots,
                               transpose_rows, transpose_cols, transpose_weights,
                               nullptr, lwork, stream);
    return;
  }
  b200_detail::CheckCode(
      cuembed_transpose_fixed(cols, batch_size, num_hots, weights,
                              b200_detail::DTypeCode<WeightT>::value,
                              b200_detail::ITypeCode<IndexT>::value,
                              transpose_rows, transpose_cols, transpose_weights,
                              work, lwork,
                              reinterpret_cast<cuembed_stream_t>(stream)),
      "TransposeFromFixed");
}

// [4,4,7,8,8,8,18] -> [0,0,1,2,2,2,3]
template <typename IndexT>
void ComputeCompressedGradIndices(const IndexT* indices, const int nnz,
                                  IndexT* remapped_indices, char* work,
                                  size_t* lwork,
                                  const cudaStream_t stream = 0) {
  b200_detail::CheckCode(
      cuembed_compressed_grad_indices(
          indices, b200_detail::ITypeCode<IndexT>::value, nnz, remapped_indices,
          work, lwork, reinterpret_cast<cuembed_stream_t>(stream)),
      "ComputeCompressedGradIndices");
}

}  // namespace cuembed

#endif  // CUEMBED_INCLUDE_INDEX_TRANSFORMS_CUH_
