// debug.cu -- optional input validation.  The reference removes
// every bounds check from its kernels ("For simplicity reasons, boundary checks
// are removed", cuembed/include/embedding_lookup_ops.cuh:59) and so do the
// kernels of this library; this entry point is the debug aid next to them: it
// scans a lookup's index (and offset) arrays on the device and reports the
// first violation.  It synchronises the stream -- never call it in a timed loop.
#include "common.cuh"
#include "launch.h"

namespace cuembed_b200 {

struct DebugReport {
  unsigned long long bad_indices;   // indices outside [0, num_rows)
  unsigned long long first_bad;     // smallest position of such an index
  unsigned long long bad_offsets;   // offsets[i] > offsets[i + 1], or out of [0, nnz]
  unsigned long long first_bad_off; // smallest such i
};

template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    DebugCheckIndicesKernel(const IdxT* __restrict__ indices, long long nnz,
                            long long num_rows, DebugReport* __restrict__ rep) {
  unsigned long long bad = 0, first = ~0ull;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
       i < nnz; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long v = static_cast<long long>(__ldg(indices + i));
    if (v < 0 || v >= num_rows) {
      ++bad;
      if (static_cast<unsigned long long>(i) < first) first = i;
    }
  }
  if (bad != 0) {
    atomicAdd(&rep->bad_indices, bad);
    atomicMin(&rep->first_bad, first);
  }
}

template <typename OffT>
__global__ void __launch_bounds__(kCtaThreads)
    DebugCheckOffsetsKernel(const OffT* __restrict__ offsets, int batch_size,
                            long long nnz, DebugReport* __restrict__ rep) {
  unsigned long long bad = 0, first = ~0ull;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
       i < batch_size; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long a = static_cast<long long>(__ldg(offsets + i));
    const long long b = static_cast<long long>(__ldg(offsets + i + 1));
    if (a < 0 || a > b || (nnz >= 0 && b > nnz)) {
      ++bad;
      if (static_cast<unsigned long long>(i) < first) first = i;
    }
  }
  if (bad != 0) {
    atomicAdd(&rep->bad_offsets, bad);
    atomicMin(&rep->first_bad_off, first);
  }
}

int LaunchDebugCheckLookup(const void* indices, int idx_type, long long nnz,
                           long long num_rows, const void* offsets, int off_type,
                           int batch_size, long long* first_bad_position,
                           cudaStream_t stream) {
  if (idx_type < 0 || idx_type > 1 || off_type < 0 || off_type > 1)
    return CUEMBED_ERR_DTYPE;
  if (nnz < 0 || num_rows < 0 || batch_size < 0) return CUEMBED_ERR_ARGUMENT;
  if ((nnz > 0 && indices == nullptr)) return CUEMBED_ERR_ARGUMENT;
  DebugReport init = {0ull, ~0ull, 0ull, ~0ull};
  DebugReport* rep = nullptr;
  if (cudaMallocAsync(&rep, sizeof(DebugReport), stream) != cudaSuccess)
    return CUEMBED_ERR_CUDA;
  cudaMemcpyAsync(rep, &init, sizeof(init), cudaMemcpyHostToDevice, stream);
  const int cap = GetDeviceInfo().sm_count * 8;
  if (nnz > 0) {
    const long long want = (nnz + kCtaThreads - 1) / kCtaThreads;
    const int grid = static_cast<int>(want < cap ? want : cap);
    if (idx_type == CUEMBED_I64)
      DebugCheckIndicesKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
          static_cast<const int64_t*>(indices), nnz, num_rows, rep);
    else
      DebugCheckIndicesKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
          static_cast<const int32_t*>(indices), nnz, num_rows, rep);
    CountLaunch();
  }
  if (offsets != nullptr && batch_size > 0) {
    const int want = (batch_size + kCtaThreads - 1) / kCtaThreads;
    const int grid = want < cap ? want : cap;
    if (off_type == CUEMBED_I64)
      DebugCheckOffsetsKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
          static_cast<const int64_t*>(offsets), batch_size, nnz, rep);
    else
      DebugCheckOffsetsKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
          static_cast<const int32_t*>(offsets), batch_size, nnz, rep);
    CountLaunch();
  }
  DebugReport host = init;
  cudaMemcpyAsync(&host, rep, sizeof(host), cudaMemcpyDeviceToHost, stream);
  cudaFreeAsync(rep, stream);
  if (cudaStreamSynchronize(stream) != cudaSuccess) return CUEMBED_ERR_CUDA;
  if (host.bad_offsets != 0) {
    if (first_bad_position != nullptr)
      *first_bad_position = static_cast<long long>(host.first_bad_off);
    return CUEMBED_ERR_OFFSETS;
  }
  if (host.bad_indices != 0) {
    if (first_bad_position != nullptr)
      *first_bad_position = static_cast<long long>(host.first_bad);
    return CUEMBED_ERR_INDEX_RANGE;
  }
  if (first_bad_position != nullptr) *first_bad_position = -1;
  return CUEMBED_OK;
}

}  // namespace cuembed_b200
