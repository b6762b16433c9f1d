// microbench.cu -- measured ceilings for the roofline of the gather kernels.
//
// The forward and backward kernels are ROW GATHERS: every lookup moves one
// table / grad_y row (512 B at the headline shape) from L2 (or DRAM) into the
// SM.  With power-law indices most of those rows are L2 hits, so the ceiling
// that bounds them is not the HBM copy bandwidth but the rate at which the L2
// slices + crossbar deliver scattered rows to the SMs.  This kernel measures
// that rate directly: the same access shape as the product kernels (a lane
// group per row, 16-byte loads, 8 rows in flight per lane group, persistent
// grid) with the arithmetic reduced to one XOR per loaded word, on a buffer
// and an index list chosen by the caller:
//   * 33.6 MB buffer (grad_y at C2), random rows  -> L2-resident gather ceiling
//   * multi-GB buffer, uniform random rows         -> DRAM gather ceiling
// bench.py reports the result as roofline.peak for bound "l2".
// Diagnostics only; nothing in the product path calls it.
#include "common.cuh"
#include "launch.h"

namespace cuembed_b200 {

template <int G, bool NO_L1>
__global__ void __launch_bounds__(kCtaThreads, 4)
    GatherRowsKernel(const char* __restrict__ buf, uint32_t row_bytes,
                     const int* __restrict__ rows, long long n,
                     unsigned* __restrict__ sink) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int UNROLL = 8;
  const int lane = threadIdx.x & 31;
  const int lane_g = lane & (G - 1);
  const int gw = lane / G;                     // lane group within the warp
  constexpr int GPW = 32 / G;                  // rows per warp step
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x +
                          threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const char* __restrict__ base = buf + lane_g * 16;
  uint32_t acc = 0;
  // a warp takes 32 consecutive index entries per round
  for (long long i0 = warp * 32; i0 < n; i0 += n_warps * 32) {
    const int my = (i0 + lane < n) ? __ldg(rows + i0 + lane) : 0;
#pragma unroll 1
    for (int jb = 0; jb < 32; jb += UNROLL * GPW) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int r = __shfl_sync(kFull, my, jb + u * GPW + gw);
        const char* p = base + static_cast<uint64_t>(static_cast<uint32_t>(r)) * row_bytes;
        if constexpr (NO_L1) {
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                       : "l"(p));
        } else {
          v[u] = __ldg(reinterpret_cast<const uint4*>(p));
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  }
  if (acc == 0x9e3779b9u) sink[0] = acc;  // keeps the loads alive
}

}  // namespace cuembed_b200

using namespace cuembed_b200;  // NOLINT

extern "C" int cuembed_microbench_gather(const void* buf, int row_bytes,
                                         const int* rows, long long n,
                                         int no_l1_allocate, unsigned* sink,
                                         cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (buf == nullptr || rows == nullptr || sink == nullptr || n < 0)
    return CUEMBED_ERR_ARGUMENT;
  if (row_bytes != 128 && row_bytes != 256 && row_bytes != 512)
    return CUEMBED_ERR_ROW_BYTES;
  if ((reinterpret_cast<uintptr_t>(buf) & 15) != 0) return CUEMBED_ERR_ARGUMENT;
  const int grid = GetDeviceInfo().sm_count * 4;
  const char* b = static_cast<const char*>(buf);
#define GATHER(GG)                                                              \
  do {                                                                          \
    if (no_l1_allocate)                                                         \
      GatherRowsKernel<GG, true><<<grid, kCtaThreads, 0, stream>>>(             \
          b, static_cast<uint32_t>(row_bytes), rows, n, sink);                  \
    else                                                                        \
      GatherRowsKernel<GG, false><<<grid, kCtaThreads, 0, stream>>>(            \
          b, static_cast<uint32_t>(row_bytes), rows, n, sink);                  \
  } while (0)
  if (row_bytes == 512)
    GATHER(32);
  else if (row_bytes == 256)
    GATHER(16);
  else
    GATHER(8);
#undef GATHER
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}
