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

// POLICY: rows are loaded with an L2 evict_last cache hint and the index list
// with evict_first (createpolicy + ld.global.nc.L2::cache_hint) -- the per-load
// form of an L2 access-policy window, which cannot be an address window here
// because hot rows are scattered over the table.
template <int G, bool NO_L1, int UNROLL = 8, bool POLICY = false>
__global__ void __launch_bounds__(kCtaThreads, (UNROLL <= 8 ? 4 : 2))
    GatherRowsKernel(const char* __restrict__ buf, uint32_t row_bytes,
                     const int* __restrict__ rows, long long n,
                     unsigned* __restrict__ sink) {
  constexpr unsigned kFull = 0xffffffffu;
  uint64_t pol_last = 0, pol_first = 0;
  if constexpr (POLICY) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
  }
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
    int my = 0;
    if (i0 + lane < n) {
      if constexpr (POLICY)
        asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;"
                     : "=r"(my)
                     : "l"(rows + i0 + lane), "l"(pol_first));
      else
        my = __ldg(rows + i0 + lane);
    }
#pragma unroll 1
    for (int jb = 0; jb < 32; jb += UNROLL * GPW) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int r = __shfl_sync(kFull, my, jb + u * GPW + gw);
        const char* p = base + static_cast<uint64_t>(static_cast<uint32_t>(r)) * row_bytes;
        if constexpr (POLICY) {
          asm volatile(
              "ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
              : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
              : "l"(p), "l"(pol_last));
        } else if constexpr (NO_L1) {
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

// ---- the same gather with the rows landing in SHARED MEMORY through the bulk
// copy engine (cp.async.bulk, SASS UBLKCP) instead of registers: every lane
// issues ONE 512-byte row copy for its own index (one warp instruction starts
// up to 32 row copies), completion is counted on an mbarrier per stage, and
// the lanes then read the rows from shared memory (LDS.128).  Rows in flight
// are bounded by shared memory (S stages x R rows x 512 B per warp), not by
// registers.
__device__ __forceinline__ uint32_t MbSmemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <int R, int S, int W>
__global__ void __launch_bounds__(W * 32, 1)
    GatherRowsBulkKernel(const char* __restrict__ buf,
                         const int* __restrict__ rows, long long n,
                         unsigned* __restrict__ sink) {
  constexpr int kRow = 512;
  extern __shared__ __align__(128) unsigned char mb_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  unsigned char* wbuf = mb_smem + static_cast<size_t>(warp) * S * R * kRow;
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(mb_smem + static_cast<size_t>(W) * S * R * kRow) +
      warp * S;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                       MbSmemAddr(bars + s)),
                   "r"(1)
                   : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const long long gw = static_cast<long long>(blockIdx.x) * W + warp;
  const long long nw = static_cast<long long>(gridDim.x) * W;
  const long long n_batches = (n + R - 1) / R;
  auto issue = [&](long long k, int stage) {
    const long long base = k * R;
    const int cnt = static_cast<int>(min(static_cast<long long>(R), n - base));
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                       MbSmemAddr(bars + stage)),
                   "r"(cnt * kRow)
                   : "memory");
    __syncwarp();
    if (lane < cnt) {
      const int r = __ldg(rows + base + lane);
      const char* src = buf + static_cast<uint64_t>(static_cast<uint32_t>(r)) * kRow;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
          "[%0], [%1], %2, [%3];" ::"r"(MbSmemAddr(wbuf + (stage * R + lane) * kRow)),
          "l"(src), "r"(kRow), "r"(MbSmemAddr(bars + stage))
          : "memory");
    }
  };
  // batches of this warp: gw, gw + nw, ...
  long long k_issue = gw;
  for (int s = 0; s < S && k_issue < n_batches; ++s, k_issue += nw) issue(k_issue, s);
  uint32_t acc = 0;
  int it = 0;
  for (long long k = gw; k < n_batches; k += nw, ++it) {
    const int stage = it % S;
    const uint32_t parity = (it / S) & 1;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(MbSmemAddr(bars + stage)),
        "r"(parity)
        : "memory");
    const unsigned char* st = wbuf + static_cast<size_t>(stage) * R * kRow + lane * 16;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint4 v = *reinterpret_cast<const uint4*>(st + j * kRow);
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    __syncwarp();  // every lane has its data in registers: the stage is free
    if (k_issue < n_batches) {
      issue(k_issue, stage);
      k_issue += nw;
    }
  }
  if (acc == 0x9e3779b9u) sink[0] = acc;
}

template <int R, int S, int W>
int LaunchGatherBulk(const char* buf, const int* rows, long long n,
                     unsigned* sink, cudaStream_t stream) {
  auto k = GatherRowsBulkKernel<R, S, W>;
  const size_t smem = static_cast<size_t>(W) * S * R * 512 + W * S * 8 + 128;
  static PerDeviceInt configured;
  if (configured.Get() == 0) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess)
      return CUEMBED_ERR_CUDA;
    configured.Set(1);
  }
  k<<<GetDeviceInfo().sm_count, W * 32, smem, stream>>>(buf, rows, n, sink);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}


// ---- the same gather with the rows landing in shared memory through per-lane
// 16-byte asynchronous copies (cp.async.cg, SASS LDGSTS.E.BYPASS.128): one warp
// instruction moves one 512-byte row, like LDG.128, but the destination is the
// lane's own shared-memory slot, so rows in flight are bounded by shared memory
// (D batches x 8 rows x 512 B per warp) instead of registers, and a lane only
// ever reads back what it copied itself (cp.async.wait_group, no barrier).
template <int D, int W>
__global__ void __launch_bounds__(W * 32)
    GatherRowsAsyncKernel(const char* __restrict__ buf,
                          const int* __restrict__ rows, long long n,
                          unsigned* __restrict__ sink) {
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kBatch = 8;
  extern __shared__ __align__(128) unsigned char mb_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const uint32_t slot0 = MbSmemAddr(mb_smem) +
                         static_cast<uint32_t>(warp) * D * kBatch * 512 + lane * 16;
  const long long gw = static_cast<long long>(blockIdx.x) * W + warp;
  const long long nw = static_cast<long long>(gridDim.x) * W;
  const char* __restrict__ base = buf + lane * 16;
  uint32_t acc = 0;
  // a warp takes 32 consecutive index entries per round = 4 batches of 8 rows
  auto issue = [&](int my, int j, int slot) {
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int r = __shfl_sync(kFull, my, j * kBatch + u);
      const char* p = base + static_cast<uint64_t>(static_cast<uint32_t>(r)) * 512;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                       slot0 + (slot * kBatch + u) * 512),
                   "l"(p)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int slot_i = 0, slot_c = 0, pending = 0;
  auto consume = [&]() {
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(slot0 + (slot_c * kBatch + u) * 512)
                   : "memory");
      acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    slot_c = slot_c + 1 == D ? 0 : slot_c + 1;
  };
  for (long long i0 = gw * 32; i0 < n; i0 += nw * 32) {
    int my = 0;
    if (i0 + lane < n) my = __ldg(rows + i0 + lane);
#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      if (pending == D) {
        consume();
        --pending;
      }
      issue(my, j, slot_i);
      slot_i = slot_i + 1 == D ? 0 : slot_i + 1;
      ++pending;
    }
  }
  // drain: empty groups keep the wait count uniform
  while (pending > 0) {
    asm volatile("cp.async.commit_group;" ::: "memory");
    consume();
    --pending;
  }
  if (acc == 0x9e3779b9u) sink[0] = acc;
}

template <int D, int W>
int LaunchGatherAsync(const char* buf, const int* rows, long long n, int ctas_per_sm,
                      unsigned* sink, cudaStream_t stream) {
  auto k = GatherRowsAsyncKernel<D, W>;
  const size_t smem = static_cast<size_t>(W) * D * 8 * 512;
  static PerDeviceInt configured;
  if (configured.Get() == 0) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess)
      return CUEMBED_ERR_CUDA;
    configured.Set(1);
  }
  k<<<GetDeviceInfo().sm_count * ctas_per_sm, W * 32, smem, stream>>>(buf, rows, n, sink);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

}  // namespace cuembed_b200

using namespace cuembed_b200;  // NOLINT

// variant: 0 = 32 rows x 2 stages x 6 warps, 1 = 16 x 3 x 8, 2 = 32 x 1 x 12,
// 3 = 16 x 2 x 12, 4 = 8 x 4 x 12  (rows per stage x stages x warps per SM)
extern "C" int cuembed_microbench_gather_bulk(const void* buf, int row_bytes,
                                              const int* rows, long long n,
                                              int variant, unsigned* sink,
                                              cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (buf == nullptr || rows == nullptr || sink == nullptr || n < 0)
    return CUEMBED_ERR_ARGUMENT;
  if (row_bytes != 512) return CUEMBED_ERR_ROW_BYTES;
  if ((reinterpret_cast<uintptr_t>(buf) & 15) != 0) return CUEMBED_ERR_ARGUMENT;
  const char* b = static_cast<const char*>(buf);
  switch (variant) {
    case 0: return LaunchGatherBulk<32, 2, 6>(b, rows, n, sink, stream);
    case 1: return LaunchGatherBulk<16, 3, 8>(b, rows, n, sink, stream);
    case 2: return LaunchGatherBulk<32, 1, 12>(b, rows, n, sink, stream);
    case 3: return LaunchGatherBulk<16, 2, 12>(b, rows, n, sink, stream);
    case 4: return LaunchGatherBulk<8, 4, 12>(b, rows, n, sink, stream);
    default: return CUEMBED_ERR_ARGUMENT;
  }
}

// variant: 0 = 2 batches x 4 warps x 6 CTAs / SM (192 KB in flight), 1 = 3 x 4 x 4
// (192 KB), 2 = 2 x 4 x 4 (128 KB), 3 = 4 x 4 x 3 (192 KB), 4 = 2 x 8 x 3 (192 KB)
// (batches of 8 rows in flight per warp x warps per CTA x CTAs per SM)
extern "C" int cuembed_microbench_gather_async(const void* buf, int row_bytes,
                                               const int* rows, long long n,
                                               int variant, unsigned* sink,
                                               cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (buf == nullptr || rows == nullptr || sink == nullptr || n < 0)
    return CUEMBED_ERR_ARGUMENT;
  if (row_bytes != 512) return CUEMBED_ERR_ROW_BYTES;
  if ((reinterpret_cast<uintptr_t>(buf) & 15) != 0) return CUEMBED_ERR_ARGUMENT;
  const char* b = static_cast<const char*>(buf);
  switch (variant) {
    case 0: return LaunchGatherAsync<2, 4>(b, rows, n, 6, sink, stream);
    case 1: return LaunchGatherAsync<3, 4>(b, rows, n, 4, sink, stream);
    case 2: return LaunchGatherAsync<2, 4>(b, rows, n, 4, sink, stream);
    case 3: return LaunchGatherAsync<4, 4>(b, rows, n, 3, sink, stream);
    case 4: return LaunchGatherAsync<2, 8>(b, rows, n, 3, sink, stream);
    default: return CUEMBED_ERR_ARGUMENT;
  }
}

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
  if (row_bytes == 512 && no_l1_allocate == 3) {
    // L2 evict_last on the rows, evict_first on the index list
    GatherRowsKernel<32, false, 8, true><<<grid, kCtaThreads, 0, stream>>>(
        b, static_cast<uint32_t>(row_bytes), rows, n, sink);
  } else if (row_bytes == 512 && no_l1_allocate == 2) {
    // deeper pipeline: 16 rows in flight per warp, 2 CTAs per SM
    GatherRowsKernel<32, false, 16><<<GetDeviceInfo().sm_count * 2, kCtaThreads, 0, stream>>>(
        b, static_cast<uint32_t>(row_bytes), rows, n, sink);
  } else if (row_bytes == 512)
    GATHER(32);
  else if (row_bytes == 256)
    GATHER(16);
  else
    GATHER(8);
#undef GATHER
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}
