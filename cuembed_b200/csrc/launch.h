// launch.h -- host-side launch entry points shared between the .cu files and
// the C ABI (c_api.cu).  All functions enqueue work on `stream` and return a
// CUEMBED_* code; none synchronises.
#ifndef CUEMBED_B200_CSRC_LAUNCH_H_
#define CUEMBED_B200_CSRC_LAUNCH_H_

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cuembed_b200 {

// Bumps the process-wide kernel launch counter (cuembed_launch_count()).
void CountLaunch(int n = 1);

int LaunchForward(const void* params, int in_dtype, int embed_width,
                  const void* indices, int idx_type, const void* offsets,
                  int off_type, const void* weights, int batch_size,
                  int num_hots, int mode, int fp16_math, void* ret,
                  int out_dtype, cudaStream_t stream,
                  const void* row_map = nullptr,
                  const void* cache_params = nullptr);

int LaunchDebugCheckLookup(const void* indices, int idx_type, long long nnz,
                           long long num_rows, const void* offsets, int off_type,
                           int batch_size, long long* first_bad_position,
                           cudaStream_t stream);

int LaunchForwardMulti(int num_tables, const void* const* params, int in_dtype,
                       int embed_width, const void* const* indices,
                       int idx_type, const void* const* offsets, int off_type,
                       const void* const* weights, const int* batch_sizes,
                       const int* num_hots, const int* modes, void* const* rets,
                       int out_dtype, long long out_row_stride,
                       cudaStream_t stream);

int LaunchExtractRowIdsFixed(int batch_size, int num_hots, void* row_ids,
                             int idx_type, cudaStream_t stream);
int LaunchExtractRowIdsCsr(const void* offsets, int off_type, int batch_size,
                           void* row_ids, int idx_type, cudaStream_t stream);
int LaunchExtractRowIdsConcat(int nnz, void* row_ids, int idx_type,
                              cudaStream_t stream);

int LaunchTranspose(const void* rows, const void* cols, const void* weights,
                    int weight_dtype, int nnz, int idx_type,
                    void* transpose_rows, void* transpose_cols,
                    void* transpose_weights, char* work, size_t* lwork,
                    cudaStream_t stream);

int LaunchTransposeFixed(const void* cols, int batch_size, int num_hots,
                         const void* weights, int weight_dtype, int idx_type,
                         void* transpose_rows, void* transpose_cols,
                         void* transpose_weights, char* work, size_t* lwork,
                         cudaStream_t stream);

int LaunchCompressedGradIndices(const void* indices, int idx_type, int nnz,
                                void* remapped, char* work, size_t* lwork,
                                cudaStream_t stream);

int LaunchBackward(const void* grad_y, int dtype, int embed_width,
                   int num_grad_embedding_rows, int nnz, int idx_type,
                   const void* transpose_indices,
                   const void* transpose_sample_ids,
                   const void* transpose_remapped_indices,
                   const void* transpose_weights, int skip_grad_init,
                   void* grad_embedding, void* inverse_mapping, char* work,
                   size_t* lwork, cudaStream_t stream);

int LaunchBackwardUpdate(const void* grad_y, int dtype, int embed_width, int nnz,
                         int idx_type, const void* transpose_indices,
                         const void* transpose_sample_ids,
                         const void* transpose_weights, int optimizer, float lr,
                         float eps, void* params, float* state, char* work,
                         size_t* lwork, cudaStream_t stream);

int LaunchShardSelect(const void* indices, int idx_type, const void* offsets,
                      int off_type, const void* weights, int weight_dtype,
                      int batch_size, int num_hots, long long row_lo,
                      long long row_hi, const int* counts_in,
                      int* local_offsets, void* local_indices,
                      void* local_sample_ids, void* local_weights, char* work,
                      size_t* lwork, cudaStream_t stream);

int LaunchShardFinalize(const void* partial_f32, int n_samples, int embed_width,
                        int mode, const void* offsets, int off_type,
                        int num_hots, int sample0, const void* weights,
                        int weight_dtype, void* out, int out_dtype,
                        cudaStream_t stream);

}  // namespace cuembed_b200

#endif  // CUEMBED_B200_CSRC_LAUNCH_H_
