// forward_hot.cu -- pooled lookup with a shared-memory cache of the hottest
// table rows (BASELINE.json north_star: "L2 persistence / access-policy windows
// for power-law hot rows"; SURVEY.md section 7 "power-law hot set": at the
// headline workload 377 rows receive 52 % of all lookups).
//
// The reference stages a CTA's indices in shared memory
// (cuembed/include/embedding_lookup_ops.cuh:465-486) and reads every row from
// L1 / L2.  An L2 access-policy window cannot hold the hot set because the
// generator scatters hot rows over the whole table; what does work is to keep
// the rows themselves next to the lanes:
//
//   * the caller passes a list of hot rows (cuembed_hot_rows_from_sorted builds
//     it on the device from the transposed indices of a batch: rows hit at
//     least `min_count` times -- in a training loop the list of step k - 1
//     serves step k, the hot set of a power law is stable);
//   * one CTA per SM (1024 threads) copies up to `capacity` hot rows into
//     shared memory (192 KB for 384 rows of 512 B) and builds an open-addressing
//     hash table row id -> slot (1024 entries) next to them;
//   * every lane probes the table for ITS OWN index of a round (one probe per
//     32 lookups per warp instruction) and the result travels with the index in
//     one 32-bit handle, so the row loop is the forward kernel's with one
//     warp-uniform branch per row: LDS.128 from the cache or LDG.128 from the
//     table.  Accumulation order and arithmetic are unchanged, so results are
//     bit-identical to cuembed_forward whatever the list contains.
//
// Scope: rows of 128, 256 or 512 bytes (one warp per row), int32 indices, sum
// or mean, fixed hotness or CSR, weighted or not.  Everything else:
// cuembed_forward.
#include "common.cuh"
#include "forward_kernels.cuh"
#include "launch.h"

namespace cuembed_b200 {

constexpr int kHotThreads = 1024;
constexpr int kHotTable = 1024;  // hash slots (power of two)
constexpr int kHotEmpty = -1;

struct FwdHotArgs {
  FwdArgs f;
  const int* hot_rows;   // [capacity] row ids
  const int* hot_count;  // device: number of valid entries (clamped to capacity)
  int capacity;
};

__device__ __forceinline__ uint32_t HotHash(uint32_t row) {
  return (row * 2654435761u) >> 22;  // top 10 bits
}

template <typename T, int V, bool WEIGHTED>
__global__ void __launch_bounds__(kHotThreads, 1)
    FwdHotKernel(const FwdHotArgs h) {
  using VecT = typename VecBits<V>::type;
  using AccT = Accum<T, V, false>;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int UNROLL = 8;
  const FwdArgs& a = h.f;
  extern __shared__ __align__(16) unsigned char hot_smem[];
  int* __restrict__ tab_key = reinterpret_cast<int*>(hot_smem);
  unsigned short* __restrict__ tab_slot =
      reinterpret_cast<unsigned short*>(hot_smem + kHotTable * sizeof(int));
  unsigned char* __restrict__ cache =
      hot_smem + kHotTable * (sizeof(int) + sizeof(unsigned short));

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t row_bytes = static_cast<uint32_t>(a.row_bytes);
  asm volatile("" : "+r"(row_bytes));
  const char* params = static_cast<const char*>(a.params) + lane * V;
  asm volatile("" : "+l"(params));

  // ---- prologue: table + cache
  for (int i = threadIdx.x; i < kHotTable; i += kHotThreads) tab_key[i] = kHotEmpty;
  __syncthreads();
  const int n_hot = min(max(__ldg(h.hot_count), 0), h.capacity);
  for (int s = threadIdx.x; s < n_hot; s += kHotThreads) {
    const int row = __ldg(h.hot_rows + s);
    if (row < 0) continue;
    uint32_t p = HotHash(static_cast<uint32_t>(row));
    while (true) {
      const int old = atomicCAS(&tab_key[p], kHotEmpty, row);
      if (old == kHotEmpty) {
        tab_slot[p] = static_cast<unsigned short>(s);
        break;
      }
      if (old == row) break;  // duplicate entry of the list: first one wins
      p = (p + 1) & (kHotTable - 1);
    }
  }
  for (int s = warp; s < n_hot; s += kHotThreads / 32) {
    const int row = __ldg(h.hot_rows + s);
    if (row < 0) continue;
    const VecT v = LdgVec<V>(RowAddr<int>(params, row, row_bytes));
    *reinterpret_cast<VecT*>(cache + static_cast<size_t>(s) * row_bytes + lane * V) = v;
  }
  __syncthreads();
  const unsigned char* my_cache = cache + lane * V;

  // handle of an index: 0x80000000 | slot for a cached row, else the row id
  auto probe = [&](int idx, bool valid) -> uint32_t {
    uint32_t handle = static_cast<uint32_t>(idx);
    if (valid) {
      uint32_t p = HotHash(static_cast<uint32_t>(idx));
      while (true) {
        const int k = tab_key[p];
        if (k == idx) {
          handle = 0x80000000u | tab_slot[p];
          break;
        }
        if (k == kHotEmpty) break;
        p = (p + 1) & (kHotTable - 1);
      }
    }
    return handle;
  };

  const int* __restrict__ indices = static_cast<const int*>(a.indices);
  const T* __restrict__ weights = static_cast<const T*>(a.weights);
  const int total_warps = gridDim.x * (kHotThreads / 32);
#pragma unroll 1
  for (int bag = blockIdx.x * (kHotThreads / 32) + warp; bag < a.batch;
       bag += total_warps) {
    int64_t start;
    int len;
    if (a.offsets != nullptr) {
      start = LoadOffset(a.offsets, a.off64, bag);
      len = static_cast<int>(LoadOffset(a.offsets, a.off64, bag + 1) - start);
    } else {
      start = static_cast<int64_t>(bag) * a.num_hots;
      len = a.num_hots;
    }
    const int* __restrict__ bag_idx = indices + start;
    const T* __restrict__ bag_w = weights + start;
    AccT acc;
    acc.Zero();
    float accw = 0.f;
    int idx_nxt = 0;
    T w_nxt = T();
    if (lane < len) {
      idx_nxt = __ldg(bag_idx + lane);
      if constexpr (WEIGHTED) w_nxt = __ldg(bag_w + lane);
    }
#pragma unroll 1
    for (int j0 = 0; j0 < len; j0 += 32) {
      const int cnt = min(32, len - j0);
      uint32_t hd_rot = probe(idx_nxt, lane < cnt);
      T w_rot = w_nxt;
      idx_nxt = 0;
      if (j0 + 32 + lane < len) {
        idx_nxt = __ldg(bag_idx + j0 + 32 + lane);
        if constexpr (WEIGHTED) w_nxt = __ldg(bag_w + j0 + 32 + lane);
      }
      const int rot_from = (lane + UNROLL) & 31;
#pragma unroll 1
      for (int jb = 0; jb < cnt; jb += UNROLL) {
        VecT vals[UNROLL];
        T wv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          // lanes past the end of the bag hold index 0 (a valid table row)
          const uint32_t hd = __shfl_sync(kFull, hd_rot, u);
          if constexpr (WEIGHTED) wv[u] = ShflElem<T>(kFull, w_rot, u, 32);
          if (static_cast<int>(hd) < 0)
            vals[u] = *reinterpret_cast<const VecT*>(
                my_cache + (hd & 0xffffu) * row_bytes);
          else
            vals[u] = LdgVec<V>(RowAddr<int>(params, static_cast<int>(hd), row_bytes));
        }
        hd_rot = __shfl_sync(kFull, hd_rot, rot_from);
        if constexpr (WEIGHTED) w_rot = ShflElem<T>(kFull, w_rot, rot_from, 32);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (jb + u < cnt) {  // warp-uniform
            if constexpr (WEIGHTED) {
              acc.AddWeighted(vals[u], wv[u]);
              accw = __fadd_rn(accw, Elem<T>::ToFloat(wv[u]));
            } else {
              acc.Add(vals[u]);
            }
          }
        }
      }
    }
    if (a.mean) {
      const float denom = WEIGHTED ? accw : static_cast<float>(len);
      if (denom == 0.f)
        acc.Zero();
      else
        acc.Scale(__fdiv_rn(1.0f, denom));
    }
    acc.Store(static_cast<char*>(a.out) + bag * a.out_row_bytes,
              static_cast<int64_t>(lane) * AccT::NE, a.out_dt);
  }
}

// Rows hit at least min_count times in a SORTED (grouped) key array: the thread
// at a run start checks the key min_count - 1 places further on.
template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    HotRowsKernel(const IdxT* __restrict__ keys, int nnz, int min_count,
                  int* __restrict__ out_rows, int capacity,
                  int* __restrict__ out_count) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       i + min_count - 1 < nnz; i += stride) {
    const IdxT k = __ldg(keys + i);
    if (i > 0 && __ldg(keys + i - 1) == k) continue;      // not a run start
    if (__ldg(keys + i + min_count - 1) != k) continue;   // run too short
    if (k < 0 || static_cast<long long>(k) > 0x7fffffffLL) continue;
    const int slot = atomicAdd(out_count, 1);
    if (slot < capacity) out_rows[slot] = static_cast<int>(k);
  }
}

template <typename T, int V>
int LaunchHotTyped(const FwdHotArgs& h, bool weighted, size_t smem,
                   cudaStream_t stream) {
  static PerDeviceInt configured_w, configured_u;
  const int grid = GetDeviceInfo().sm_count;
  if (weighted) {
    auto k = FwdHotKernel<T, V, true>;
    if (configured_w.Get() < static_cast<int>(smem)) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)) != cudaSuccess)
        return CUEMBED_ERR_CUDA;
      configured_w.Set(static_cast<int>(smem));
    }
    k<<<grid, kHotThreads, smem, stream>>>(h);
  } else {
    auto k = FwdHotKernel<T, V, false>;
    if (configured_u.Get() < static_cast<int>(smem)) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)) != cudaSuccess)
        return CUEMBED_ERR_CUDA;
      configured_u.Set(static_cast<int>(smem));
    }
    k<<<grid, kHotThreads, smem, stream>>>(h);
  }
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

template <typename T>
int LaunchHotVec(const FwdHotArgs& h, int v, bool weighted, size_t smem,
                 cudaStream_t stream) {
  if (v == 16) return LaunchHotTyped<T, 16>(h, weighted, smem, stream);
  if (v == 8) return LaunchHotTyped<T, 8>(h, weighted, smem, stream);
  return LaunchHotTyped<T, 4>(h, weighted, smem, stream);
}

}  // namespace cuembed_b200

using namespace cuembed_b200;  // NOLINT

extern "C" int cuembed_forward_hot_capacity(int in_dtype, int embed_width) {
  if (in_dtype < 0 || in_dtype > 2 || embed_width <= 0) return 0;
  const int64_t row_bytes = static_cast<int64_t>(embed_width) * ElemSize(in_dtype);
  if (row_bytes != 128 && row_bytes != 256 && row_bytes != 512) return 0;
  // shared memory: table + cache, leaving ~32 KB of the 228 KB to L1
  const int64_t budget = 192 * 1024;
  int64_t cap = budget / row_bytes;
  if (cap > 768) cap = 768;  // load factor of the 1024-entry table <= 0.75
  return static_cast<int>(cap);
}

extern "C" int cuembed_forward_hot(const void* params, int in_dtype,
                                   int embed_width, const void* indices,
                                   int idx_type, const void* offsets,
                                   int off_type, const void* weights,
                                   int batch_size, int num_hots, int mode,
                                   void* ret, int out_dtype,
                                   const int* hot_rows, const int* hot_count,
                                   int capacity, cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!((offsets != nullptr && num_hots == 0) ||
        (offsets == nullptr && num_hots > 0)))
    return CUEMBED_ERR_CSR_XOR_FIXED;
  if (in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2)
    return CUEMBED_ERR_DTYPE;
  if (idx_type != CUEMBED_I32 || (mode != CUEMBED_SUM && mode != CUEMBED_MEAN))
    return CUEMBED_ERR_DTYPE;
  if (batch_size < 0 || embed_width <= 0 || capacity < 0)
    return CUEMBED_ERR_ARGUMENT;
  if (batch_size == 0) return CUEMBED_OK;
  if (params == nullptr || indices == nullptr || ret == nullptr ||
      hot_rows == nullptr || hot_count == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  const int max_cap = cuembed_forward_hot_capacity(in_dtype, embed_width);
  if (max_cap == 0) return CUEMBED_ERR_ROW_BYTES;
  if (capacity > max_cap) capacity = max_cap;
  const int64_t row_bytes = static_cast<int64_t>(embed_width) * ElemSize(in_dtype);
  const int v = static_cast<int>(row_bytes / 32);
  const int64_t out_row_bytes =
      static_cast<int64_t>(embed_width) * ElemSize(out_dtype);
  const int64_t out_vec = static_cast<int64_t>(v) * ElemSize(out_dtype) / ElemSize(in_dtype);
  if (reinterpret_cast<uintptr_t>(params) % v != 0 ||
      (reinterpret_cast<uintptr_t>(ret) | static_cast<uint64_t>(out_row_bytes)) %
              (out_vec > 16 ? 16 : out_vec) != 0)
    return CUEMBED_ERR_ARGUMENT;
  FwdHotArgs h;
  h.f.params = params;
  h.f.indices = indices;
  h.f.offsets = offsets;
  h.f.weights = weights;
  h.f.out = ret;
  h.f.row_bytes = row_bytes;
  h.f.out_row_bytes = out_row_bytes;
  h.f.batch = batch_size;
  h.f.num_hots = num_hots;
  h.f.off64 = off_type == CUEMBED_I64;
  h.f.mean = mode == CUEMBED_MEAN;
  h.f.out_dt = out_dtype;
  h.f.nvec = 32;
  h.f.lanes = 32;
  h.f.log2_lanes = 5;
  h.f.col_tiles = 1;
  h.hot_rows = hot_rows;
  h.hot_count = hot_count;
  h.capacity = capacity;
  const size_t smem = kHotTable * (sizeof(int) + sizeof(unsigned short)) +
                      static_cast<size_t>(capacity) * row_bytes;
  const bool weighted = weights != nullptr;
  if (in_dtype == CUEMBED_F32) return LaunchHotVec<float>(h, v, weighted, smem, stream);
  if (in_dtype == CUEMBED_F16) return LaunchHotVec<__half>(h, v, weighted, smem, stream);
  return LaunchHotVec<__nv_bfloat16>(h, v, weighted, smem, stream);
}

extern "C" int cuembed_hot_rows_from_sorted(const void* sorted_keys, int idx_type,
                                            int nnz, int min_count, int* hot_rows,
                                            int capacity, int* hot_count,
                                            cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (idx_type < 0 || idx_type > 1) return CUEMBED_ERR_DTYPE;
  if (nnz < 0 || min_count < 1 || capacity < 0 || hot_count == nullptr ||
      (capacity > 0 && hot_rows == nullptr))
    return CUEMBED_ERR_ARGUMENT;
  if (cudaMemsetAsync(hot_count, 0, sizeof(int), stream) != cudaSuccess)
    return CUEMBED_ERR_CUDA;
  if (nnz == 0) return CUEMBED_OK;
  if (sorted_keys == nullptr) return CUEMBED_ERR_ARGUMENT;
  const int64_t ctas = (static_cast<int64_t>(nnz) + kCtaThreads - 1) / kCtaThreads;
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * 8;
  const int grid = static_cast<int>(ctas < cap ? ctas : cap);
  if (idx_type == CUEMBED_I64)
    HotRowsKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<const int64_t*>(sorted_keys), nnz, min_count, hot_rows,
        capacity, hot_count);
  else
    HotRowsKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<const int32_t*>(sorted_keys), nnz, min_count, hot_rows,
        capacity, hot_count);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}
