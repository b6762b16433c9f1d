// forward.cu -- gather-pool forward kernels (sum / mean / concat) for sm_100a.
//
// Replaces the reference's EmbeddingLookUpKernel + IndexLoader / Addresser /
// Combiner policy classes (cuembed/include/embedding_lookup_kernels.cuh:34-170,
// cuembed/include/embedding_lookup_ops.cuh:72-495) with one design:
//
//   * a lane GROUP (G = 8..32 lanes, chosen from the row width) owns one bag at
//     a time; each lane owns one V-byte vector (V = 16 where the row allows) of
//     the row and accumulates it in registers, sequentially in bag order -- the
//     same order as the reference CPU loop, so fp32 results are bit-identical;
//   * indices (and weights) are loaded one coalesced round of G at a time,
//     two rounds ahead of use, and broadcast with group-scoped shuffles: no
//     shared memory, no CTA-wide barrier, CSR bags do not serialise an index
//     load in front of every row load;
//   * UNROLL row loads (16 B each, read-only path) are issued back to back
//     before any is consumed: UNROLL x 512 B in flight per warp;
//   * a persistent grid (SM count x resident CTAs) walks the bags.
//
// HBM/L2-bound byte work: no tensor cores (nothing here is a contraction).
#include "common.cuh"
#include "forward_kernels.cuh"
#include <algorithm>

#include "launch.h"

namespace cuembed_b200 {

namespace {

int Log2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Persistent grid: SM count x resident CTAs of this kernel, capped by the
// amount of work.  `occ_cache` is a per-instantiation, per-device static of the
// caller.
int PersistentGrid(const void* kernel, PerDeviceInt* occ_cache,
                   int64_t work_ctas) {
  int occ = occ_cache->Get();
  if (occ == 0) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kCtaThreads, 0);
    occ = n > 0 ? n : 1;
    occ_cache->Set(occ);
  }
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * occ;
  int64_t g = work_ctas < cap ? work_ctas : cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

// UNROLL = 8 row loads in flight per lane group (4 and 16 were measured slower
// in round 1 and are no longer built).
template <typename T, int V, typename IdxT, bool WEIGHTED, bool LOWP>
void LaunchPool(const FwdArgs& a, cudaStream_t stream) {
  const int groups_per_cta = kCtaThreads / a.lanes;
  const int64_t work_ctas = (a.batch + groups_per_cta - 1) / groups_per_cta;
  if constexpr (!LOWP) {
    if (a.row_map != nullptr) {
      static PerDeviceInt occ_mapped;
      auto k = FwdPoolMappedKernel<T, V, IdxT, WEIGHTED>;
      const int grid = PersistentGrid(reinterpret_cast<const void*>(k),
                                      &occ_mapped, work_ctas);
      k<<<dim3(grid, a.col_tiles), kCtaThreads, 0, stream>>>(a);
      CountLaunch();
      return;
    }
  }
  static PerDeviceInt occ8;
  auto k = FwdPoolKernel<T, V, IdxT, WEIGHTED, LOWP, 8>;
  const int grid =
      PersistentGrid(reinterpret_cast<const void*>(k), &occ8, work_ctas);
  k<<<dim3(grid, a.col_tiles), kCtaThreads, 0, stream>>>(a);
  CountLaunch();
}

template <typename T, int V, typename IdxT>
void LaunchPoolFlags(const FwdArgs& a, bool weighted, bool lowp,
                     cudaStream_t stream) {
  if constexpr (sizeof(T) == 4) {
    // fp16_math is a no-op for fp32 tables
    // (cuembed/include/embedding_lookup_types.cuh:497-551).
    if (weighted)
      LaunchPool<T, V, IdxT, true, false>(a, stream);
    else
      LaunchPool<T, V, IdxT, false, false>(a, stream);
  } else {
    if (weighted && lowp)
      LaunchPool<T, V, IdxT, true, true>(a, stream);
    else if (weighted)
      LaunchPool<T, V, IdxT, true, false>(a, stream);
    else if (lowp)
      LaunchPool<T, V, IdxT, false, true>(a, stream);
    else
      LaunchPool<T, V, IdxT, false, false>(a, stream);
  }
}

template <typename T, typename IdxT>
void LaunchPoolVec(const FwdArgs& a, int vec_bytes, bool weighted, bool lowp,
                   cudaStream_t stream) {
  if (vec_bytes == 16)
    LaunchPoolFlags<T, 16, IdxT>(a, weighted, lowp, stream);
  else if (vec_bytes == 8)
    LaunchPoolFlags<T, 8, IdxT>(a, weighted, lowp, stream);
  else
    LaunchPoolFlags<T, 4, IdxT>(a, weighted, lowp, stream);
}

template <typename T>
void LaunchPoolIdx(const FwdArgs& a, int idx_type, int vec_bytes,
                   bool weighted, bool lowp, cudaStream_t stream) {
  if (idx_type == CUEMBED_I64)
    LaunchPoolVec<T, int64_t>(a, vec_bytes, weighted, lowp, stream);
  else
    LaunchPoolVec<T, int32_t>(a, vec_bytes, weighted, lowp, stream);
}

template <int V, typename IdxT>
void LaunchConcat(const ConcatArgs& a, cudaStream_t stream) {
  auto k = FwdConcatKernel<V, IdxT>;
  const int groups_per_cta = kCtaThreads / a.lanes;
  const int64_t work_ctas =
      (a.nnz + static_cast<int64_t>(groups_per_cta) * 4 - 1) /
      (static_cast<int64_t>(groups_per_cta) * 4);
  static PerDeviceInt occ;
  const int grid =
      PersistentGrid(reinterpret_cast<const void*>(k), &occ, work_ctas);
  k<<<grid, kCtaThreads, 0, stream>>>(a);
  CountLaunch();
}

// Vector width (bytes of INPUT per lane and step) for the pooled kernels:
// starts at `v0` and halves until the OR of the input addresses / pitches is a
// multiple of v and the OR of the output addresses / pitches is a multiple of
// the matching output store (v * sizeof(out) / sizeof(in), at most 16 bytes).
// Returns 0 if not even 4-byte input vectors are possible.
int PickPoolVector(int v0, uint64_t in_bits, uint64_t out_bits, int in_dtype,
                   int out_dtype) {
  for (int v = v0; v >= 4; v /= 2) {
    const int64_t out_vec =
        static_cast<int64_t>(v) * ElemSize(out_dtype) / ElemSize(in_dtype);
    if (in_bits % v == 0 && out_bits % (out_vec > 16 ? 16 : out_vec) == 0)
      return v;
  }
  return 0;
}

// Largest power-of-two <= 16 that divides every value in `bits` (an OR of
// addresses and pitches).
int CommonAlign(uint64_t bits) {
  int v = 16;
  while (v > 1 && (bits & (v - 1)) != 0) v /= 2;
  return v;
}

template <typename T, int V, typename IdxT>
void LaunchPoolMulti(const FwdMultiArgs& m, int num_tables, int col_tiles,
                     int lanes, int max_batch, bool weighted,
                     cudaStream_t stream) {
  const int groups_per_cta = kCtaThreads / lanes;
  const int64_t work_ctas = (max_batch + groups_per_cta - 1) / groups_per_cta;
  static PerDeviceInt occ_w, occ_u;
  int grid;
  if (weighted) {
    auto k = FwdPoolMultiKernel<T, V, IdxT, true>;
    grid = PersistentGrid(reinterpret_cast<const void*>(k), &occ_w, work_ctas);
    grid = std::max(1, std::min<int>(grid, (grid + num_tables - 1) / num_tables * 2));
    k<<<dim3(grid, col_tiles, num_tables), kCtaThreads, 0, stream>>>(m);
  } else {
    auto k = FwdPoolMultiKernel<T, V, IdxT, false>;
    grid = PersistentGrid(reinterpret_cast<const void*>(k), &occ_u, work_ctas);
    grid = std::max(1, std::min<int>(grid, (grid + num_tables - 1) / num_tables * 2));
    k<<<dim3(grid, col_tiles, num_tables), kCtaThreads, 0, stream>>>(m);
  }
  CountLaunch();
}

template <typename T, typename IdxT>
void LaunchPoolMultiVec(const FwdMultiArgs& m, int num_tables, int vec_bytes,
                        int col_tiles, int lanes, int max_batch, bool weighted,
                        cudaStream_t stream) {
  if (vec_bytes == 16)
    LaunchPoolMulti<T, 16, IdxT>(m, num_tables, col_tiles, lanes, max_batch,
                                 weighted, stream);
  else if (vec_bytes == 8)
    LaunchPoolMulti<T, 8, IdxT>(m, num_tables, col_tiles, lanes, max_batch,
                                weighted, stream);
  else
    LaunchPoolMulti<T, 4, IdxT>(m, num_tables, col_tiles, lanes, max_batch,
                                weighted, stream);
}

}  // namespace

// Multi-table batched forward: tables of one dtype, width, index type and
// weighted-ness; batch size, hotness / CSR and sum / mean may differ per table.
// HOST arrays of num_tables entries.  out_row_stride (elements of out_dtype,
// 0 = embed_width) lets the pooled rows of all tables land in one
// [batch, num_tables * embed_width] activation matrix.
int LaunchForwardMulti(int num_tables, const void* const* params, int in_dtype,
                       int embed_width, const void* const* indices,
                       int idx_type, const void* const* offsets, int off_type,
                       const void* const* weights, const int* batch_sizes,
                       const int* num_hots, const int* modes, void* const* rets,
                       int out_dtype, long long out_row_stride,
                       cudaStream_t stream) {
  if (num_tables < 0 || embed_width <= 0) return CUEMBED_ERR_ARGUMENT;
  if (num_tables == 0) return CUEMBED_OK;
  if (params == nullptr || indices == nullptr || batch_sizes == nullptr ||
      num_hots == nullptr || rets == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  if (in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2 ||
      idx_type < 0 || idx_type > 1)
    return CUEMBED_ERR_DTYPE;
  const int64_t row_bytes =
      static_cast<int64_t>(embed_width) * ElemSize(in_dtype);
  if (row_bytes % 4 != 0) return CUEMBED_ERR_ROW_BYTES;
  if (out_row_stride == 0) out_row_stride = embed_width;
  if (out_row_stride < embed_width) return CUEMBED_ERR_ARGUMENT;
  const int64_t out_row_bytes = out_row_stride * ElemSize(out_dtype);
  RowShape shape;
  MakeRowShape(embed_width, in_dtype, &shape);

  // One vector width for the launch: limited by every pointer and pitch.
  bool weighted = false, any = false;
  uint64_t in_bits = static_cast<uint64_t>(row_bytes);
  uint64_t out_bits = static_cast<uint64_t>(out_row_bytes);
  for (int t = 0; t < num_tables; ++t) {
    const int mode = modes != nullptr ? modes[t] : CUEMBED_SUM;
    const void* off = offsets != nullptr ? offsets[t] : nullptr;
    const void* w = weights != nullptr ? weights[t] : nullptr;
    if (mode != CUEMBED_SUM && mode != CUEMBED_MEAN) return CUEMBED_ERR_DTYPE;
    if (!((off != nullptr && num_hots[t] == 0) ||
          (off == nullptr && num_hots[t] > 0)))
      return CUEMBED_ERR_CSR_XOR_FIXED;
    if (batch_sizes[t] < 0) return CUEMBED_ERR_ARGUMENT;
    if (batch_sizes[t] == 0) continue;
    if (params[t] == nullptr || indices[t] == nullptr || rets[t] == nullptr)
      return CUEMBED_ERR_ARGUMENT;
    if (any && (w != nullptr) != weighted) return CUEMBED_ERR_ARGUMENT;
    weighted = w != nullptr;
    any = true;
    in_bits |= reinterpret_cast<uint64_t>(params[t]);
    out_bits |= reinterpret_cast<uint64_t>(rets[t]);
  }
  if (!any) return CUEMBED_OK;
  const int v =
      PickPoolVector(shape.vec_bytes, in_bits, out_bits, in_dtype, out_dtype);
  if (v == 0) return CUEMBED_ERR_ARGUMENT;

  const int nvec = static_cast<int>(row_bytes / v);
  const int lanes = Pow2Ceil(nvec) < 32 ? Pow2Ceil(nvec) : 32;
  const int col_tiles = (nvec + lanes - 1) / lanes;
  for (int t0 = 0; t0 < num_tables; t0 += kMaxTablesPerLaunch) {
    FwdMultiArgs m;
    int n = 0, max_batch = 0;
    for (int t = t0; t < num_tables && t < t0 + kMaxTablesPerLaunch; ++t) {
      if (batch_sizes[t] == 0) continue;
      FwdArgs& a = m.t[n++];
      a.row_map = nullptr;
      a.cache = nullptr;
      a.params = params[t];
      a.indices = indices[t];
      a.offsets = offsets != nullptr ? offsets[t] : nullptr;
      a.weights = weights != nullptr ? weights[t] : nullptr;
      a.out = rets[t];
      a.row_bytes = row_bytes;
      a.out_row_bytes = out_row_bytes;
      a.batch = batch_sizes[t];
      a.num_hots = num_hots[t];
      a.off64 = (off_type == CUEMBED_I64);
      a.mean = (modes != nullptr && modes[t] == CUEMBED_MEAN);
      a.out_dt = out_dtype;
      a.nvec = nvec;
      a.lanes = lanes;
      a.log2_lanes = Log2(lanes);
      a.col_tiles = col_tiles;
      if (a.batch > max_batch) max_batch = a.batch;
    }
    if (n == 0) continue;
#define MULTI_CASE(TT)                                                        \
  if (idx_type == CUEMBED_I64)                                                \
    LaunchPoolMultiVec<TT, int64_t>(m, n, v, col_tiles, lanes, max_batch,     \
                                    weighted, stream);                        \
  else                                                                        \
    LaunchPoolMultiVec<TT, int32_t>(m, n, v, col_tiles, lanes, max_batch,     \
                                    weighted, stream)
    if (in_dtype == CUEMBED_F32) {
      MULTI_CASE(float);
    } else if (in_dtype == CUEMBED_F16) {
      MULTI_CASE(__half);
    } else {
      MULTI_CASE(__nv_bfloat16);
    }
#undef MULTI_CASE
  }
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

int LaunchForward(const void* params, int in_dtype, int embed_width,
                  const void* indices, int idx_type, const void* offsets,
                  int off_type, const void* weights, int batch_size,
                  int num_hots, int mode, int fp16_math, void* ret,
                  int out_dtype, cudaStream_t stream, const void* row_map,
                  const void* cache_params) {
  // the addresser indirection exists for the pooled modes with fp32 accumulation
  if (row_map == nullptr && cache_params != nullptr) return CUEMBED_ERR_ARGUMENT;
  if (row_map != nullptr &&
      (mode == CUEMBED_CONCAT || (fp16_math != 0 && in_dtype != CUEMBED_F32)))
    return CUEMBED_ERR_ARGUMENT;
  // Same argument checks as the reference host function,
  // cuembed/include/embedding_lookup.cuh:260-267.
  if (weights != nullptr && mode == CUEMBED_CONCAT)
    return CUEMBED_ERR_WEIGHTED_CONCAT;
  if (!((offsets != nullptr && num_hots == 0) ||
        (offsets == nullptr && num_hots > 0)))
    return CUEMBED_ERR_CSR_XOR_FIXED;
  if (offsets != nullptr && mode == CUEMBED_CONCAT)
    return CUEMBED_ERR_CSR_CONCAT;
  if (in_dtype < 0 || in_dtype > 2 || out_dtype < 0 || out_dtype > 2 ||
      idx_type < 0 || idx_type > 1 || mode < 0 || mode > 2)
    return CUEMBED_ERR_DTYPE;
  if (batch_size < 0 || embed_width <= 0) return CUEMBED_ERR_ARGUMENT;
  if (batch_size == 0) return CUEMBED_OK;
  if (params == nullptr || indices == nullptr || ret == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  const int64_t row_bytes =
      static_cast<int64_t>(embed_width) * ElemSize(in_dtype);
  if (row_bytes % 4 != 0) return CUEMBED_ERR_ROW_BYTES;

  RowShape shape;
  MakeRowShape(embed_width, in_dtype, &shape);

  if (mode == CUEMBED_CONCAT) {
    if (out_dtype != in_dtype) return CUEMBED_ERR_DTYPE;
    // Vector width limited by the actual pointer alignment.
    int v = CommonAlign(reinterpret_cast<uint64_t>(params) |
                        reinterpret_cast<uint64_t>(ret) |
                        static_cast<uint64_t>(row_bytes));
    if (v > shape.vec_bytes) v = shape.vec_bytes;
    if (v < 4) return CUEMBED_ERR_ARGUMENT;
    ConcatArgs c;
    c.params = params;
    c.indices = indices;
    c.out = ret;
    c.row_bytes = row_bytes;
    c.nnz = static_cast<int64_t>(batch_size) * num_hots;
    c.nvec = static_cast<int>(row_bytes / v);
    c.lanes = Pow2Ceil(c.nvec) < 32 ? Pow2Ceil(c.nvec) : 32;
    c.log2_lanes = Log2(c.lanes);
    c.col_tiles = (c.nvec + c.lanes - 1) / c.lanes;
    if (c.nnz == 0) return CUEMBED_OK;
#define CONCAT_CASE(VV)                               \
  if (idx_type == CUEMBED_I64)                        \
    LaunchConcat<VV, int64_t>(c, stream);             \
  else                                                \
    LaunchConcat<VV, int32_t>(c, stream)
    if (v == 16) {
      CONCAT_CASE(16);
    } else if (v == 8) {
      CONCAT_CASE(8);
    } else {
      CONCAT_CASE(4);
    }
#undef CONCAT_CASE
    return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
  }

  // Sum / mean.  The input vector of V bytes holds NE elements; the output
  // vector holds the same NE elements of the output type.
  const int64_t out_row_bytes =
      static_cast<int64_t>(embed_width) * ElemSize(out_dtype);
  // Widest vector that both the input side (table pointer, row pitch) and
  // the output side (ret pointer, output pitch, NE output elements per store)
  // allow; a pointer that does not even allow 4-byte input vectors / the
  // matching output store is an argument error, not a misaligned access.
  const int v = PickPoolVector(
      shape.vec_bytes,
      reinterpret_cast<uint64_t>(params) | reinterpret_cast<uint64_t>(cache_params) |
          static_cast<uint64_t>(row_bytes),
      reinterpret_cast<uint64_t>(ret) | static_cast<uint64_t>(out_row_bytes),
      in_dtype, out_dtype);
  if (v == 0) return CUEMBED_ERR_ARGUMENT;

  FwdArgs a;
  a.row_map = row_map;
  a.cache = cache_params;
  a.params = params;
  a.indices = indices;
  a.offsets = offsets;
  a.weights = weights;
  a.out = ret;
  a.row_bytes = row_bytes;
  a.out_row_bytes = out_row_bytes;
  a.batch = batch_size;
  a.num_hots = num_hots;
  a.off64 = (off_type == CUEMBED_I64);
  a.mean = (mode == CUEMBED_MEAN);
  a.out_dt = out_dtype;
  a.nvec = static_cast<int>(row_bytes / v);
  a.lanes = Pow2Ceil(a.nvec) < 32 ? Pow2Ceil(a.nvec) : 32;
  a.log2_lanes = Log2(a.lanes);
  a.col_tiles = (a.nvec + a.lanes - 1) / a.lanes;

  const bool weighted = weights != nullptr;
  const bool lowp = fp16_math != 0 && in_dtype != CUEMBED_F32;
  if (in_dtype == CUEMBED_F32)
    LaunchPoolIdx<float>(a, idx_type, v, weighted, lowp, stream);
  else if (in_dtype == CUEMBED_F16)
    LaunchPoolIdx<__half>(a, idx_type, v, weighted, lowp, stream);
  else
    LaunchPoolIdx<__nv_bfloat16>(a, idx_type, v, weighted, lowp, stream);
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

}  // namespace cuembed_b200
