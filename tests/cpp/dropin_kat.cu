// dropin_kat.cu -- native C++ caller of the drop-in host templates
// (include/cuembed/include/*.cuh), written the way a cuEmbed user writes it:
// same include paths, same namespace, same calls, template arguments deduced
// where the reference's callers deduce them.  Runs the reference's known-answer
// vectors (tests/test_embedding_forward.cu:118-160,
// tests/test_embedding_transpose.cu:112-122,
// tests/test_embedding_backward.cu:162-202) on the GPU and exits non-zero on
// any mismatch.  Built by __graft_entry__.build(); run by tests/test_dropin_cpp.py.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "cuembed/include/embedding_lookup.cuh"
#include "cuembed/include/index_transforms.cuh"

using cuembed::CombineMode;

#define CUDA_OK(x)                                                     \
  do {                                                                 \
    cudaError_t e_ = (x);                                              \
    if (e_ != cudaSuccess) {                                           \
      std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_),  \
                  __FILE__, __LINE__);                                 \
      return 2;                                                        \
    }                                                                  \
  } while (0)

template <typename T>
T* Managed(const std::vector<float>& v) {
  T* p = nullptr;
  cudaMallocManaged(&p, sizeof(T) * (v.empty() ? 1 : v.size()));
  for (size_t i = 0; i < v.size(); ++i) p[i] = static_cast<T>(v[i]);
  return p;
}
template <typename T>
T* ManagedI(const std::vector<long long>& v) {
  T* p = nullptr;
  cudaMallocManaged(&p, sizeof(T) * (v.empty() ? 1 : v.size()));
  for (size_t i = 0; i < v.size(); ++i) p[i] = static_cast<T>(v[i]);
  return p;
}
template <typename T>
bool Equal(const T* got, const std::vector<float>& want, const char* what) {
  for (size_t i = 0; i < want.size(); ++i) {
    if (static_cast<float>(got[i]) != want[i]) {
      std::printf("FAIL %s: element %zu got %f want %f\n", what, i,
                  static_cast<float>(got[i]), want[i]);
      return false;
    }
  }
  return true;
}
template <typename T>
bool EqualI(const T* got, const std::vector<long long>& want, const char* what) {
  for (size_t i = 0; i < want.size(); ++i) {
    if (static_cast<long long>(got[i]) != want[i]) {
      std::printf("FAIL %s: element %zu got %lld want %lld\n", what, i,
                  static_cast<long long>(got[i]), want[i]);
      return false;
    }
  }
  return true;
}

template <typename ElemT, typename IndexT>
int RunAll(const char* tag) {
  int failures = 0;
  std::vector<float> table(20);
  for (int i = 0; i < 20; ++i) table[i] = static_cast<float>(i + 1);
  ElemT* params = Managed<ElemT>(table);
  IndexT* indices = ManagedI<IndexT>({1, 3, 0, 4});
  int* offsets = ManagedI<int>({0, 2, 4});
  ElemT* weights = Managed<ElemT>({1.f, 0.5f, 1.f, 0.5f});
  ElemT* ret = Managed<ElemT>(std::vector<float>(16, -1.f));
  const int* no_offsets = nullptr;
  const ElemT* no_weights = nullptr;

  // ---- forward (fixed hotness and CSR)
  cuembed::EmbeddingForward<ElemT, ElemT, IndexT, int>(
      params, 4, indices, no_offsets, no_weights, 2, 2, CombineMode::kSum, ret);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(ret, {18, 20, 22, 24, 18, 20, 22, 24}, "fixed sum");
  cuembed::EmbeddingForward<ElemT, ElemT, IndexT, int>(
      params, 4, indices, offsets, no_weights, 2, 0, CombineMode::kMean, ret);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(ret, {9, 10, 11, 12, 9, 10, 11, 12}, "csr mean");
  cuembed::EmbeddingForward<ElemT, ElemT, IndexT, int>(
      params, 4, indices, offsets, weights, 2, 0, CombineMode::kSum, ret);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(ret, {11.5, 13, 14.5, 16, 9.5, 11, 12.5, 14}, "csr weighted sum");
  cuembed::EmbeddingForward<ElemT, ElemT, IndexT, int>(
      params, 4, indices, no_offsets, no_weights, 2, 2, CombineMode::kConcat, ret);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(ret, {5, 6, 7, 8, 13, 14, 15, 16, 1, 2, 3, 4, 17, 18, 19, 20},
                     "fixed concat");

  // ---- row ids + transpose with the two-call workspace protocol
  IndexT* sample_ids = ManagedI<IndexT>({-1, -1, -1, -1});
  cuembed::ExtractRowIdsFromCSR(offsets, 2, sample_ids);  // types deduced
  CUDA_OK(cudaDeviceSynchronize());
  failures += !EqualI(sample_ids, {0, 0, 1, 1}, "row ids from CSR");
  IndexT* t_idx = ManagedI<IndexT>({-1, -1, -1, -1});
  IndexT* t_sid = ManagedI<IndexT>({-1, -1, -1, -1});
  IndexT* remapped = ManagedI<IndexT>({-1, -1, -1, -1});
  ElemT* t_w = Managed<ElemT>({0, 0, 0, 0});
  size_t lwork = 0, lwork2 = 0;
  cuembed::Transpose<IndexT, ElemT>(sample_ids, indices, weights, 4, t_idx, t_sid,
                                    t_w, nullptr, &lwork);
  cuembed::ComputeCompressedGradIndices<IndexT>(t_idx, 4, remapped, nullptr, &lwork2);
  if (lwork2 > lwork) lwork = lwork2;
  char* work = nullptr;
  CUDA_OK(cudaMalloc(&work, lwork));
  cuembed::Transpose<IndexT, ElemT>(sample_ids, indices, weights, 4, t_idx, t_sid,
                                    t_w, work, &lwork);
  cuembed::ComputeCompressedGradIndices<IndexT>(t_idx, 4, remapped, work, &lwork);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !EqualI(t_idx, {0, 1, 3, 4}, "transpose indices");
  failures += !EqualI(t_sid, {1, 0, 0, 1}, "transpose sample ids");
  failures += !Equal(t_w, {1, 1, 0.5, 0.5}, "transpose weights");
  failures += !EqualI(remapped, {0, 1, 2, 3}, "remapped");

  // ---- backward KAT (test_embedding_backward.cu:162-202)
  IndexT* b_idx = ManagedI<IndexT>({0, 1, 3, 3});
  IndexT* b_rem = ManagedI<IndexT>({0, 1, 2, 2});
  IndexT* b_sid = ManagedI<IndexT>({1, 0, 0, 1});
  ElemT* b_w = Managed<ElemT>({3.f, 1.f, 0.5f, 3.f});
  ElemT* grad_y = Managed<ElemT>({1, 2, 3, 4, 5, 6, 7, 8});
  ElemT* grad = Managed<ElemT>(std::vector<float>(20, -1.f));
  IndexT* inv = ManagedI<IndexT>({-1, -1, -1});
  const IndexT* no_remap = nullptr;
  IndexT* no_inv = nullptr;
  cuembed::EmbeddingBackward<ElemT, IndexT>(grad_y, 4, 5, 4, b_idx, b_sid, no_remap,
                                            no_weights, false, grad, no_inv);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(grad, {5, 6, 7, 8, 1, 2, 3, 4, 0, 0, 0, 0, 6, 8, 10, 12, 0, 0, 0, 0},
                     "backward full");
  cuembed::EmbeddingBackward<ElemT, IndexT>(grad_y, 4, 3, 4, b_idx, b_sid, b_rem, b_w,
                                            false, grad, inv);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(grad, {15, 18, 21, 24, 1, 2, 3, 4, 15.5, 19, 22.5, 26},
                     "backward compressed weighted");
  failures += !EqualI(inv, {0, 1, 3}, "inverse mapping");

  // ---- additions of the B200 build: fused SGD step and multi-table forward
  // table rows 0, 1, 3 receive the "backward full" sums above; lr = 0.5
  ElemT* tab2 = Managed<ElemT>(table);
  size_t lw_upd = 0;
  cuembed::EmbeddingBackwardUpdate<ElemT, IndexT>(
      grad_y, 4, 4, b_idx, b_sid, no_weights, cuembed::SparseOptimizer::kSgd, 0.5f,
      0.f, tab2, nullptr, nullptr, &lw_upd);
  char* work_upd = nullptr;
  CUDA_OK(cudaMalloc(&work_upd, lw_upd));
  cuembed::EmbeddingBackwardUpdate<ElemT, IndexT>(
      grad_y, 4, 4, b_idx, b_sid, no_weights, cuembed::SparseOptimizer::kSgd, 0.5f,
      0.f, tab2, nullptr, work_upd, &lw_upd);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(tab2, {1 - 2.5f, 2 - 3, 3 - 3.5f, 4 - 4, 5 - 0.5f, 6 - 1, 7 - 1.5f,
                            8 - 2, 9, 10, 11, 12, 13 - 3, 14 - 4, 15 - 5, 16 - 6, 17, 18,
                            19, 20},
                     "fused sgd step");
  ElemT* ret2 = Managed<ElemT>(std::vector<float>(16, -1.f));
  const ElemT* m_params[2] = {params, params};
  const IndexT* m_indices[2] = {indices, indices};
  const int* m_offsets[2] = {nullptr, offsets};
  const int m_batch[2] = {2, 2};
  const int m_hots[2] = {2, 0};
  const CombineMode m_modes[2] = {CombineMode::kSum, CombineMode::kMean};
  ElemT* m_rets[2] = {ret2, ret2 + 4};  // one [2, 2 * 4] activation matrix
  const ElemT* const* m_no_weights = nullptr;
  cuembed::EmbeddingForwardMulti<ElemT, ElemT, IndexT, int>(
      2, m_params, 4, m_indices, m_offsets, m_no_weights, m_batch, m_hots, m_modes,
      m_rets, 8);
  CUDA_OK(cudaDeviceSynchronize());
  failures += !Equal(ret2, {18, 20, 22, 24, 9, 10, 11, 12, 18, 20, 22, 24, 9, 10, 11, 12},
                     "multi-table forward");
  // ---- structured InputT: a table behind an addresser indirection
  // (cuembed::MappedTable, the reference's embedding-cache hook).  Row 1 is
  // served from slot 0 of a cache table, row 4 is remapped to ... itself
  // uncached; rows 0 and 3 stay in the backing table.
  {
    std::vector<float> cache_v = {101, 102, 103, 104};
    ElemT* cache = Managed<ElemT>(cache_v);
    IndexT* row_map = ManagedI<IndexT>({-1, 0, -1, -1, -1});
    cuembed::MappedTable<ElemT, IndexT> mapped = {params, row_map, cache};
    ElemT* ret3 = Managed<ElemT>(std::vector<float>(8, -1.f));
    cuembed::EmbeddingForward<cuembed::MappedTable<ElemT, IndexT>, ElemT, IndexT, int>(
        &mapped, 4, indices, no_offsets, no_weights, 2, 2, CombineMode::kSum, ret3);
    CUDA_OK(cudaDeviceSynchronize());
    // bag 0 = rows {1 -> cache slot 0, 3}, bag 1 = rows {0, 4}
    failures += !Equal(ret3, {101 + 13, 102 + 14, 103 + 15, 104 + 16, 18, 20, 22, 24},
                       "mapped table forward");
    // pure remapping inside the table: row 3 reads row 0
    IndexT* row_map2 = ManagedI<IndexT>({-1, -1, -1, 0, -1});
    cuembed::MappedTable<ElemT, IndexT> remap = {params, row_map2, nullptr};
    cuembed::EmbeddingForward<cuembed::MappedTable<ElemT, IndexT>, ElemT, IndexT, int>(
        &remap, 4, indices, offsets, weights, 2, 0, CombineMode::kSum, ret3);
    CUDA_OK(cudaDeviceSynchronize());
    // bag 0 = 1 * row 1 + 0.5 * row 0, bag 1 = 1 * row 0 + 0.5 * row 4
    failures += !Equal(ret3, {5.5f, 7, 8.5f, 10, 9.5f, 11, 12.5f, 14}, "remapped forward");
  }
  // ---- debug bounds check: the valid lookup passes (an invalid one aborts,
  // which the Python tests exercise through the C ABI)
  cuembed::DebugCheckLookup<IndexT, int>(indices, 4, 5, offsets, 2);

  std::printf("%s: %s\n", tag, failures == 0 ? "PASS" : "FAIL");
  return failures;
}

int main() {
  int failures = 0;
  failures += RunAll<float, int32_t>("float/int32");
  failures += RunAll<float, int64_t>("float/int64");
  failures += RunAll<__half, int32_t>("half/int32");
  failures += RunAll<__half, int64_t>("half/int64");
  failures += RunAll<__nv_bfloat16, int32_t>("bfloat16/int32");
  std::printf("dropin_kat: %d failure(s)\n", failures);
  return failures == 0 ? 0 : 1;
}
