// ref_gpu_shim.cu -- BENCH/TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's own GPU templates (cuembed/include/*.cuh, header-only + CUB
// from the CUDA toolkit) compiled unchanged from /root/reference for sm_100,
// behind the same extern "C" signatures as include/cuembed_b200.h, so that
// scripts/bench_ref_gpu.py can time the reference's kernels and ours on the
// same B200 with the same protocol.  Built by oracle/Makefile into
// oracle/_ref/libcuembed_refgpu.so; nothing in cuembed_b200/ links or loads it.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "cuembed/include/embedding_lookup.cuh"
#include "cuembed/include/index_transforms.cuh"

namespace {
using cuembed::CombineMode;
CombineMode ToMode(int m) {
  return m == 0 ? CombineMode::kSum : (m == 1 ? CombineMode::kMean : CombineMode::kConcat);
}

template <typename T, typename IdxT, typename OffT>
void Fwd(const void* params, int w, const void* idx, const void* off, const void* wt,
         int batch, int hots, int mode, int fp16_math, void* ret, cudaStream_t s) {
  if (fp16_math)
    cuembed::EmbeddingForward<T, T, IdxT, OffT, true>(
        static_cast<const T*>(params), w, static_cast<const IdxT*>(idx),
        static_cast<const OffT*>(off), static_cast<const T*>(wt), batch, hots,
        ToMode(mode), static_cast<T*>(ret), s);
  else
    cuembed::EmbeddingForward<T, T, IdxT, OffT, false>(
        static_cast<const T*>(params), w, static_cast<const IdxT*>(idx),
        static_cast<const OffT*>(off), static_cast<const T*>(wt), batch, hots,
        ToMode(mode), static_cast<T*>(ret), s);
}
}  // namespace

extern "C" {

int refgpu_forward(const void* params, int in_dtype, int embed_width,
                   const void* indices, int idx_type, const void* offsets,
                   int off_type, const void* weights, int batch_size,
                   int num_hots, int mode, int fp16_math, void* ret,
                   int out_dtype, cudaStream_t stream) {
  if (in_dtype != out_dtype || off_type != 0 || in_dtype > 1) return -5;
#define F(T, I) Fwd<T, I, int>(params, embed_width, indices, offsets, weights, batch_size, num_hots, mode, fp16_math, ret, stream)
  if (in_dtype == 0 && idx_type == 0) F(float, int32_t);
  else if (in_dtype == 0) F(float, int64_t);
  else if (idx_type == 0) F(__half, int32_t);
  else F(__half, int64_t);
#undef F
  return 0;
}

int refgpu_extract_row_ids_fixed(int batch_size, int num_hots, void* row_ids,
                                 int idx_type, cudaStream_t stream) {
  if (idx_type)
    cuembed::ExtractRowIdsFromFixed<int64_t>(batch_size, num_hots, static_cast<int64_t*>(row_ids), stream);
  else
    cuembed::ExtractRowIdsFromFixed<int32_t>(batch_size, num_hots, static_cast<int32_t*>(row_ids), stream);
  return 0;
}

int refgpu_transpose(const void* rows, const void* cols, const void* weights,
                     int weight_dtype, int nnz, int idx_type, void* t_rows,
                     void* t_cols, void* t_weights, char* work, size_t* lwork,
                     cudaStream_t stream) {
#define T(I, W) cuembed::Transpose<I, W>(static_cast<const I*>(rows), static_cast<const I*>(cols), \
    static_cast<const W*>(weights), nnz, static_cast<I*>(t_rows), static_cast<I*>(t_cols),        \
    static_cast<W*>(t_weights), work, lwork, stream)
  if (idx_type == 0 && weight_dtype == 0) T(int32_t, float);
  else if (idx_type == 0) T(int32_t, __half);
  else if (weight_dtype == 0) T(int64_t, float);
  else T(int64_t, __half);
#undef T
  return 0;
}

int refgpu_compressed_grad_indices(const void* indices, int idx_type, int nnz,
                                   void* remapped, char* work, size_t* lwork,
                                   cudaStream_t stream) {
  if (idx_type)
    cuembed::ComputeCompressedGradIndices<int64_t>(static_cast<const int64_t*>(indices), nnz,
                                                   static_cast<int64_t*>(remapped), work, lwork, stream);
  else
    cuembed::ComputeCompressedGradIndices<int32_t>(static_cast<const int32_t*>(indices), nnz,
                                                   static_cast<int32_t*>(remapped), work, lwork, stream);
  return 0;
}

int refgpu_backward(const void* grad_y, int dtype, int embed_width, int rows,
                    int nnz, int idx_type, const void* t_idx, const void* t_sid,
                    const void* t_remapped, const void* t_w, int skip_grad_init,
                    void* grad, void* inv, cudaStream_t stream) {
#define B(G, I) cuembed::EmbeddingBackward<G, I>(static_cast<const G*>(grad_y), embed_width, rows, nnz, \
    static_cast<const I*>(t_idx), static_cast<const I*>(t_sid), static_cast<const I*>(t_remapped),      \
    static_cast<const G*>(t_w), skip_grad_init != 0, static_cast<G*>(grad), static_cast<I*>(inv), stream)
  if (dtype == 0 && idx_type == 0) B(float, int32_t);
  else if (dtype == 0) B(float, int64_t);
  else if (idx_type == 0) B(__half, int32_t);
  else B(__half, int64_t);
#undef B
  return 0;
}

}  // extern "C"
