// c_api.cu -- the extern "C" boundary declared in include/cuembed_b200.h.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "launch.h"

namespace cuembed_b200 {

// CUEMBED_NVTX=1: every entry point opens an NVTX range around its launches
// (the header-only NVTX v3 costs one branch when no tool listens).
struct NvtxScope {
  bool on;
  explicit NvtxScope(const char* name) {
    static const int enabled = EnvInt("CUEMBED_NVTX", 0);
    on = enabled != 0;
    if (on) nvtxRangePushA(name);
  }
  ~NvtxScope() {
    if (on) nvtxRangePop();
  }
};

static std::atomic<unsigned long long> g_launches{0};

void CountLaunch(int n) { g_launches.fetch_add(static_cast<unsigned>(n)); }

int EnvInt(const char* name, int default_value) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == '\0') return default_value;
  return std::atoi(v);
}

}  // namespace cuembed_b200

using namespace cuembed_b200;  // NOLINT

extern "C" {

int cuembed_version(void) { return 100; }

const char* cuembed_build_arch(void) { return "sm_100a"; }

const char* cuembed_error_string(int code) {
  switch (code) {
    case CUEMBED_OK:
      return "ok";
    case CUEMBED_ERR_WEIGHTED_CONCAT:
      return "Check failed: weights == nullptr || mode != CombineMode::kConcat";
    case CUEMBED_ERR_CSR_XOR_FIXED:
      return "Check failed: (offsets != nullptr && num_hots == 0) || "
             "(offsets == nullptr && num_hots > 0)";
    case CUEMBED_ERR_CSR_CONCAT:
      return "Check failed: offsets == nullptr || mode != CombineMode::kConcat";
    case CUEMBED_ERR_ROW_BYTES:
      return "Check failed: bytes_per_row % 4 == 0";
    case CUEMBED_ERR_DTYPE:
      return "unsupported dtype / index type / mode combination";
    case CUEMBED_ERR_WORKSPACE:
      return "Check failed: *lwork >= required_workspace";
    case CUEMBED_ERR_ARGUMENT:
      return "null pointer, negative size or misaligned buffer";
    case CUEMBED_ERR_CUDA:
      return "a CUDA runtime call or kernel launch failed";
    case CUEMBED_ERR_NNZ_LIMIT:
      return "nnz must be < 2^30 for transpose";
    case CUEMBED_ERR_INDEX_RANGE:
      return "debug check: a lookup index lies outside [0, num_rows)";
    case CUEMBED_ERR_OFFSETS:
      return "debug check: offsets are negative, not ascending or exceed nnz";
    default:
      return "unknown error";
  }
}

int cuembed_forward(const void* params, int in_dtype, int embed_width,
                    const void* indices, int idx_type, const void* offsets,
                    int off_type, const void* weights, int batch_size,
                    int num_hots, int mode, int fp16_math, void* ret,
                    int out_dtype, cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_forward");
  return LaunchForward(params, in_dtype, embed_width, indices, idx_type,
                       offsets, off_type, weights, batch_size, num_hots, mode,
                       fp16_math, ret, out_dtype,
                       reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_forward_mapped(const void* params, int in_dtype, int embed_width,
                           const void* indices, int idx_type,
                           const void* offsets, int off_type,
                           const void* weights, int batch_size, int num_hots,
                           int mode, void* ret, int out_dtype,
                           const void* row_map, const void* cache_params,
                           cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_forward_mapped");
  if (row_map == nullptr) return CUEMBED_ERR_ARGUMENT;
  return LaunchForward(params, in_dtype, embed_width, indices, idx_type,
                       offsets, off_type, weights, batch_size, num_hots, mode,
                       /*fp16_math=*/0, ret, out_dtype,
                       reinterpret_cast<cudaStream_t>(stream), row_map,
                       cache_params);
}

int cuembed_forward_multi(int num_tables, const void* const* params,
                          int in_dtype, int embed_width,
                          const void* const* indices, int idx_type,
                          const void* const* offsets, int off_type,
                          const void* const* weights, const int* batch_sizes,
                          const int* num_hots, const int* modes,
                          void* const* rets, int out_dtype,
                          long long out_row_stride, cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_forward_multi");
  return LaunchForwardMulti(num_tables, params, in_dtype, embed_width, indices,
                            idx_type, offsets, off_type, weights, batch_sizes,
                            num_hots, modes, rets, out_dtype, out_row_stride,
                            reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_extract_row_ids_fixed(int batch_size, int num_hots, void* row_ids,
                                  int idx_type, cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_extract_row_ids_fixed");
  return LaunchExtractRowIdsFixed(batch_size, num_hots, row_ids, idx_type,
                                  reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_extract_row_ids_csr(const void* offsets, int off_type,
                                int batch_size, void* row_ids, int idx_type,
                                cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_extract_row_ids_csr");
  return LaunchExtractRowIdsCsr(offsets, off_type, batch_size, row_ids,
                                idx_type,
                                reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_extract_row_ids_concat(int nnz, void* row_ids, int idx_type,
                                   cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_extract_row_ids_concat");
  return LaunchExtractRowIdsConcat(nnz, row_ids, idx_type,
                                   reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_transpose(const void* rows, const void* cols, const void* weights,
                      int weight_dtype, int nnz, int idx_type,
                      void* transpose_rows, void* transpose_cols,
                      void* transpose_weights, char* work, size_t* lwork,
                      cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_transpose");
  return LaunchTranspose(rows, cols, weights, weight_dtype, nnz, idx_type,
                         transpose_rows, transpose_cols, transpose_weights,
                         work, lwork, reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_transpose_fixed(const void* cols, int batch_size, int num_hots,
                            const void* weights, int weight_dtype, int idx_type,
                            void* transpose_rows, void* transpose_cols,
                            void* transpose_weights, char* work, size_t* lwork,
                            cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_transpose_fixed");
  return LaunchTransposeFixed(cols, batch_size, num_hots, weights, weight_dtype,
                              idx_type, transpose_rows, transpose_cols,
                              transpose_weights, work, lwork,
                              reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_compressed_grad_indices(const void* indices, int idx_type, int nnz,
                                    void* remapped_indices, char* work,
                                    size_t* lwork, cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_compressed_grad_indices");
  return LaunchCompressedGradIndices(indices, idx_type, nnz, remapped_indices,
                                     work, lwork,
                                     reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_backward_ws(const void* grad_y, int dtype, int embed_width,
                        int num_grad_embedding_rows, int nnz, int idx_type,
                        const void* transpose_indices,
                        const void* transpose_sample_ids,
                        const void* transpose_remapped_indices,
                        const void* transpose_weights, int skip_grad_init,
                        void* grad_embedding, void* inverse_mapping,
                        char* work, size_t* lwork, cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_backward_ws");
  return LaunchBackward(grad_y, dtype, embed_width, num_grad_embedding_rows,
                        nnz, idx_type, transpose_indices, transpose_sample_ids,
                        transpose_remapped_indices, transpose_weights,
                        skip_grad_init, grad_embedding, inverse_mapping, work,
                        lwork, reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_backward_update(const void* grad_y, int dtype, int embed_width,
                            int nnz, int idx_type,
                            const void* transpose_indices,
                            const void* transpose_sample_ids,
                            const void* transpose_weights, int optimizer,
                            float lr, float eps, void* params, float* state,
                            char* work, size_t* lwork, cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_backward_update");
  return LaunchBackwardUpdate(grad_y, dtype, embed_width, nnz, idx_type,
                              transpose_indices, transpose_sample_ids,
                              transpose_weights, optimizer, lr, eps, params,
                              state, work, lwork,
                              reinterpret_cast<cudaStream_t>(stream));
}

// Scratch for the drop-in signature (no workspace argument): a library-owned
// stream-ordered pool per device that keeps its memory between calls, so the
// steady state costs no driver allocation.  cudaMallocFromPoolAsync /
// cudaFreeAsync are stream-ordered and CUDA-graph capturable.
int cuembed_backward(const void* grad_y, int dtype, int embed_width,
                     int num_grad_embedding_rows, int nnz, int idx_type,
                     const void* transpose_indices,
                     const void* transpose_sample_ids,
                     const void* transpose_remapped_indices,
                     const void* transpose_weights, int skip_grad_init,
                     void* grad_embedding, void* inverse_mapping,
                     cuembed_stream_t stream_) {
  NvtxScope nvtx_scope("cuembed_backward");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  size_t lwork = 0;
  int rc = LaunchBackward(grad_y, dtype, embed_width, num_grad_embedding_rows,
                          nnz, idx_type, transpose_indices,
                          transpose_sample_ids, transpose_remapped_indices,
                          transpose_weights, skip_grad_init, grad_embedding,
                          inverse_mapping, nullptr, &lwork, stream);
  if (rc != CUEMBED_OK) return rc;
  if (nnz == 0 && skip_grad_init) return CUEMBED_OK;

  // one pool per device, created once under a lock (two host threads may make
  // their first call on the same device at the same time)
  static cudaMemPool_t pools[kMaxDevices] = {nullptr};
  static std::mutex pools_mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
    return CUEMBED_ERR_CUDA;
  cudaMemPool_t pool = nullptr;
  {
    std::lock_guard<std::mutex> lock(pools_mu);
    if (pools[dev] == nullptr) {
      cudaMemPoolProps props;
      std::memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      cudaMemPool_t created;
      if (cudaMemPoolCreate(&created, &props) != cudaSuccess)
        return CUEMBED_ERR_CUDA;
      unsigned long long threshold = ~0ull;  // never trim: reuse across calls
      cudaMemPoolSetAttribute(created, cudaMemPoolAttrReleaseThreshold,
                              &threshold);
      pools[dev] = created;
    }
    pool = pools[dev];
  }
  void* work = nullptr;
  if (cudaMallocFromPoolAsync(&work, lwork, pool, stream) != cudaSuccess)
    return CUEMBED_ERR_CUDA;
  rc = LaunchBackward(grad_y, dtype, embed_width, num_grad_embedding_rows, nnz,
                      idx_type, transpose_indices, transpose_sample_ids,
                      transpose_remapped_indices, transpose_weights,
                      skip_grad_init, grad_embedding, inverse_mapping,
                      static_cast<char*>(work), &lwork, stream);
  cudaFreeAsync(work, stream);
  return rc;
}

int cuembed_shard_select(const void* indices, int idx_type, const void* offsets,
                         int off_type, const void* weights, int weight_dtype,
                         int batch_size, int num_hots, long long row_lo,
                         long long row_hi, int* local_offsets,
                         void* local_indices, void* local_weights, char* work,
                         size_t* lwork, cuembed_stream_t stream) {
  return LaunchShardSelect(indices, idx_type, offsets, off_type, weights,
                           weight_dtype, batch_size, num_hots, row_lo, row_hi,
                           nullptr, local_offsets, local_indices, nullptr,
                           local_weights, work, lwork,
                           reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_shard_select_coo(const void* indices, int idx_type,
                             const void* offsets, int off_type,
                             const void* weights, int weight_dtype,
                             int batch_size, int num_hots, long long row_lo,
                             long long row_hi, const int* counts,
                             int* local_offsets, void* local_indices,
                             void* local_sample_ids, void* local_weights,
                             char* work, size_t* lwork,
                             cuembed_stream_t stream) {
  return LaunchShardSelect(indices, idx_type, offsets, off_type, weights,
                           weight_dtype, batch_size, num_hots, row_lo, row_hi,
                           counts, local_offsets, local_indices,
                           local_sample_ids, local_weights, work, lwork,
                           reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_shard_finalize(const void* partial_f32, int n_samples,
                           int embed_width, int mode, const void* offsets,
                           int off_type, int num_hots, int sample0,
                           const void* weights, int weight_dtype, void* out,
                           int out_dtype, cuembed_stream_t stream) {
  return LaunchShardFinalize(partial_f32, n_samples, embed_width, mode, offsets,
                             off_type, num_hots, sample0, weights, weight_dtype,
                             out, out_dtype,
                             reinterpret_cast<cudaStream_t>(stream));
}

int cuembed_debug_check_lookup(const void* indices, int idx_type, long long nnz,
                               long long num_rows, const void* offsets,
                               int off_type, int batch_size,
                               long long* first_bad_position,
                               cuembed_stream_t stream) {
  NvtxScope nvtx_scope("cuembed_debug_check_lookup");
  return LaunchDebugCheckLookup(indices, idx_type, nnz, num_rows, offsets,
                                off_type, batch_size, first_bad_position,
                                reinterpret_cast<cudaStream_t>(stream));
}

unsigned long long cuembed_launch_count(void) { return g_launches.load(); }

}  // extern "C"
