// sharded.cu -- kernels for the row-sharded multi-GPU mode (new functionality;
// the reference is single-GPU, README.md:110).
//
// A table too large for one GPU is split into contiguous row ranges, one per
// GPU.  Every GPU sees the (replicated) lookup indices of the global batch and
//   1. SELECTS the lookups that fall into its row range, producing a compact
//      local CSR (offsets, indices rebased to the shard, weights) that keeps
//      bag order -- cuembed_shard_select;
//   2. pools them with the ordinary forward kernel into an fp32 partial
//      [batch, width] (sum);
//   3. the partials are combined by an NCCL reduce-scatter (host side,
//      cuembed_b200/sharded.py);
//   4. FINALIZES its slice of samples: mean = sum / global bag length (or sum
//      of weights), cast to the output type -- cuembed_shard_finalize.
// Backward needs no new kernel: all-gather of grad_y, then the ordinary
// transpose + backward on the local CSR; gradients never leave the owner.
#include "common.cuh"
#include "launch.h"

namespace cuembed_b200 {

// --------------------------------------------------------------- select

// counts[b] = number of lookups of bag b with lo <= index < hi.  One warp per
// bag.
template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    ShardCountKernel(const IdxT* __restrict__ indices, const void* offsets,
                     int off64, int num_hots, int batch, long long lo,
                     long long hi, int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < batch; b += nwarps) {
    int64_t start, end;
    if (offsets != nullptr) {
      start = LoadOffset(offsets, off64, b);
      end = LoadOffset(offsets, off64, b + 1);
    } else {
      start = static_cast<int64_t>(b) * num_hots;
      end = start + num_hots;
    }
    int c = 0;
    for (int64_t i = start + lane; i < end; i += 32) {
      const long long v = static_cast<long long>(__ldg(indices + i));
      c += (v >= lo && v < hi) ? 1 : 0;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) counts[b] = c;
  }
}

// Exclusive scan of int32 counts[n] -> out[n + 1] (out[n] = total), two
// kernels over a fixed partition (same scheme as the compressed-index remap).
constexpr int kSelItems = 8;
constexpr int kSelChunk = kSelItems * kCtaThreads;
constexpr int kSelMaxParts = 1024;

__global__ void __launch_bounds__(kCtaThreads)
    ScanPartSumKernel(const int* __restrict__ in, int n, int chunks_per_part,
                      long long* __restrict__ part_sums) {
  __shared__ long long s_warp[kWarpsPerCta];
  const int tid = threadIdx.x;
  const int64_t begin = static_cast<int64_t>(blockIdx.x) * chunks_per_part * kSelChunk;
  const int64_t end = min(static_cast<int64_t>(n),
                          begin + static_cast<int64_t>(chunks_per_part) * kSelChunk);
  long long local = 0;
  for (int64_t i = begin + tid; i < end; i += kCtaThreads) local += in[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((tid & 31) == 0) s_warp[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int w = 0; w < kWarpsPerCta; ++w) t += s_warp[w];
    part_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kCtaThreads)
    ScanWriteKernel(const int* __restrict__ in, int n, int chunks_per_part,
                    const long long* __restrict__ part_sums, int parts,
                    int* __restrict__ out) {
  __shared__ long long s_red[kWarpsPerCta];
  __shared__ int s_warp[kWarpsPerCta];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  long long before = 0;
  for (int p = tid; p < static_cast<int>(blockIdx.x); p += kCtaThreads)
    before += part_sums[p];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) s_red[warp] = before;
  __syncthreads();
  long long carry = 0;
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) carry += s_red[w];

  const int64_t begin = static_cast<int64_t>(blockIdx.x) * chunks_per_part * kSelChunk;
  for (int c = 0; c < chunks_per_part; ++c) {
    const int64_t chunk_begin = begin + static_cast<int64_t>(c) * kSelChunk;
    if (chunk_begin >= n) break;
    const int64_t base = chunk_begin + tid * kSelItems;
    int v[kSelItems];
    int local = 0;
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      v[i] = base + i < n ? in[base + i] : 0;
      local += v[i];
    }
    int scan = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, scan, o);
      if (lane >= o) scan += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = scan;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) {
      if (w < warp) wbase += s_warp[w];
      total += s_warp[w];
    }
    long long running = carry + wbase + scan - local;
#pragma unroll
    for (int i = 0; i < kSelItems; ++i) {
      if (base + i < n) out[base + i] = static_cast<int>(running);
      running += v[i];
    }
    carry += total;
  }
  // out[n] = grand total, written by the last part
  if (static_cast<int>(blockIdx.x) == parts - 1 && tid == 0)
    out[n] = static_cast<int>(carry);
}

// Writes the selected lookups of every bag, in bag order, rebased to the shard.
template <typename IdxT, int WBYTES>
__global__ void __launch_bounds__(kCtaThreads)
    ShardFillKernel(const IdxT* __restrict__ indices, const void* offsets,
                    int off64, int num_hots, const void* __restrict__ weights,
                    int batch, long long lo, long long hi,
                    const int* __restrict__ local_offsets,
                    IdxT* __restrict__ local_indices,
                    IdxT* __restrict__ local_sample_ids,
                    void* __restrict__ local_weights) {
  using WT = typename std::conditional<WBYTES == 4, uint32_t, uint16_t>::type;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  if (offsets == nullptr) {
    // Fixed hotness: kBags bags per warp at a time, all their index loads (and
    // output offsets) requested before the first is used.  One bag at a time was
    // three dependent memory latencies per bag with nothing else in flight
    // (8 ranks scan 33.5 M indices to keep 4 M).
    constexpr int kBags = 4;
    const unsigned long long span = static_cast<unsigned long long>(hi - lo);
    for (int b0 = warp; b0 < batch; b0 += kBags * nwarps) {
      int out[kBags];
#pragma unroll
      for (int k = 0; k < kBags; ++k) {
        const int b = b0 + k * nwarps;
        out[k] = b < batch ? __ldg(local_offsets + b) : 0;
      }
      for (int i0 = 0; i0 < num_hots; i0 += 32) {
        const bool in_bag = i0 + lane < num_hots;
        long long v[kBags];
#pragma unroll
        for (int k = 0; k < kBags; ++k) {
          const int b = b0 + k * nwarps;
          v[k] = -1;
          if (b < batch && in_bag)
            v[k] = static_cast<long long>(
                __ldg(indices + static_cast<int64_t>(b) * num_hots + i0 + lane));
        }
#pragma unroll
        for (int k = 0; k < kBags; ++k) {
          const int b = b0 + k * nwarps;
          const unsigned long long rel = static_cast<unsigned long long>(v[k] - lo);
          const bool keep = b < batch && in_bag && v[k] >= 0 && rel < span;
          const unsigned m = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            const int dst = out[k] + __popc(m & lt);
            local_indices[dst] = static_cast<IdxT>(rel);
            if (local_sample_ids != nullptr) local_sample_ids[dst] = static_cast<IdxT>(b);
            if constexpr (WBYTES != 0)
              static_cast<WT*>(local_weights)[dst] = static_cast<const WT*>(
                  weights)[static_cast<int64_t>(b) * num_hots + i0 + lane];
          }
          out[k] += __popc(m);
        }
      }
    }
    return;
  }
  for (int b = warp; b < batch; b += nwarps) {
    int64_t start, end;
    if (offsets != nullptr) {
      start = LoadOffset(offsets, off64, b);
      end = LoadOffset(offsets, off64, b + 1);
    } else {
      start = static_cast<int64_t>(b) * num_hots;
      end = start + num_hots;
    }
    int out = local_offsets[b];
    for (int64_t i0 = start; i0 < end; i0 += 32) {
      const int64_t i = i0 + lane;
      long long v = -1;
      if (i < end) v = static_cast<long long>(__ldg(indices + i));
      const bool keep = i < end && v >= lo && v < hi;
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int dst = out + __popc(m & lt);
        local_indices[dst] = static_cast<IdxT>(v - lo);
        if (local_sample_ids != nullptr) local_sample_ids[dst] = static_cast<IdxT>(b);
        if constexpr (WBYTES != 0)
          static_cast<WT*>(local_weights)[dst] =
              static_cast<const WT*>(weights)[i];
      }
      out += __popc(m);
    }
  }
}

namespace {
int WarpGrid(int64_t warps) {
  const int64_t ctas = (warps * 32 + kCtaThreads - 1) / kCtaThreads;
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * 8;
  return static_cast<int>(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}
}  // namespace

int LaunchShardSelect(const void* indices, int idx_type, const void* offsets,
                      int off_type, const void* weights, int weight_dtype,
                      int batch_size, int num_hots, long long row_lo,
                      long long row_hi, const int* counts_in,
                      int* local_offsets, void* local_indices,
                      void* local_sample_ids, void* local_weights, char* work,
                      size_t* lwork, cudaStream_t stream) {
  if (lwork == nullptr || batch_size < 0) return CUEMBED_ERR_ARGUMENT;
  if (idx_type < 0 || idx_type > 1) return CUEMBED_ERR_DTYPE;
  if (!((offsets != nullptr && num_hots == 0) ||
        (offsets == nullptr && num_hots > 0)))
    return CUEMBED_ERR_CSR_XOR_FIXED;
  const int chunks = batch_size > 0 ? (batch_size + kSelChunk - 1) / kSelChunk : 0;
  const int chunks_per_part = chunks > 0 ? (chunks + kSelMaxParts - 1) / kSelMaxParts : 1;
  const int parts = chunks > 0 ? (chunks + chunks_per_part - 1) / chunks_per_part : 0;
  size_t off = 0;
  const size_t counts_off = off;
  off += AlignUp(static_cast<size_t>(batch_size) * sizeof(int), 256);
  const size_t sums_off = off;
  off += AlignUp(static_cast<size_t>(parts) * sizeof(long long), 256);
  const size_t need = off > 0 ? off : 256;
  if (work == nullptr) {
    *lwork = need;
    return CUEMBED_OK;
  }
  if (*lwork < need) return CUEMBED_ERR_WORKSPACE;
  if (local_offsets == nullptr) return CUEMBED_ERR_ARGUMENT;
  if (batch_size == 0) {
    cudaMemsetAsync(local_offsets, 0, sizeof(int), stream);
    return CUEMBED_OK;
  }
  if (indices == nullptr || local_indices == nullptr) return CUEMBED_ERR_ARGUMENT;
  if (weights != nullptr && local_weights == nullptr) return CUEMBED_ERR_ARGUMENT;
  const int* counts = counts_in != nullptr
                          ? counts_in
                          : reinterpret_cast<int*>(work + counts_off);
  long long* sums = reinterpret_cast<long long*>(work + sums_off);
  const int off64 = off_type == CUEMBED_I64;
  const int grid = WarpGrid(batch_size);
  const int wbytes = weights != nullptr ? static_cast<int>(ElemSize(weight_dtype)) : 0;
#define SELECT(IdxT)                                                            \
  if (counts_in == nullptr)                                                     \
    ShardCountKernel<IdxT><<<grid, kCtaThreads, 0, stream>>>(                   \
        static_cast<const IdxT*>(indices), offsets, off64, num_hots,            \
        batch_size, row_lo, row_hi, reinterpret_cast<int*>(work + counts_off)); \
  ScanPartSumKernel<<<parts, kCtaThreads, 0, stream>>>(counts, batch_size,      \
                                                       chunks_per_part, sums);  \
  ScanWriteKernel<<<parts, kCtaThreads, 0, stream>>>(                           \
      counts, batch_size, chunks_per_part, sums, parts, local_offsets);         \
  if (wbytes == 0)                                                              \
    ShardFillKernel<IdxT, 0><<<grid, kCtaThreads, 0, stream>>>(                 \
        static_cast<const IdxT*>(indices), offsets, off64, num_hots, weights,   \
        batch_size, row_lo, row_hi, local_offsets,                              \
        static_cast<IdxT*>(local_indices),                                      \
        static_cast<IdxT*>(local_sample_ids), local_weights);                   \
  else if (wbytes == 2)                                                         \
    ShardFillKernel<IdxT, 2><<<grid, kCtaThreads, 0, stream>>>(                 \
        static_cast<const IdxT*>(indices), offsets, off64, num_hots, weights,   \
        batch_size, row_lo, row_hi, local_offsets,                              \
        static_cast<IdxT*>(local_indices),                                      \
        static_cast<IdxT*>(local_sample_ids), local_weights);                   \
  else                                                                          \
    ShardFillKernel<IdxT, 4><<<grid, kCtaThreads, 0, stream>>>(                 \
        static_cast<const IdxT*>(indices), offsets, off64, num_hots, weights,   \
        batch_size, row_lo, row_hi, local_offsets,                              \
        static_cast<IdxT*>(local_indices),                                      \
        static_cast<IdxT*>(local_sample_ids), local_weights)
  if (idx_type == CUEMBED_I64) {
    SELECT(int64_t);
  } else {
    SELECT(int32_t);
  }
#undef SELECT
  CountLaunch(counts_in == nullptr ? 4 : 3);
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

// -------------------------------------------------------------- finalize

// out[s, :] = cast(partial[s, :] * scale(s)); scale = 1 (sum), 1 / bag length
// (mean) or 1 / sum of weights (weighted mean, zero vector if that is 0) -- the
// same rules as the single-GPU epilogue (forward_kernels.cuh), with the GLOBAL
// bag length.  One thread per 4 output elements.
template <typename WT>
__global__ void __launch_bounds__(kCtaThreads)
    ShardFinalizeKernel(const float* __restrict__ partial, int n_samples,
                        int width, int mean, const void* offsets, int off64,
                        int num_hots, int sample0, const WT* __restrict__ weights,
                        void* __restrict__ out, int out_dt) {
  const int64_t total = static_cast<int64_t>(n_samples) * width;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       e < total; e += stride) {
    const int s = static_cast<int>(e / width);
    float v = partial[e];
    if (mean) {
      int64_t start, len;
      if (offsets != nullptr) {
        start = LoadOffset(offsets, off64, sample0 + s);
        len = LoadOffset(offsets, off64, sample0 + s + 1) - start;
      } else {
        start = static_cast<int64_t>(sample0 + s) * num_hots;
        len = num_hots;
      }
      float denom;
      if (weights != nullptr) {
        denom = 0.f;
        for (int64_t j = 0; j < len; ++j)
          denom = __fadd_rn(denom, Elem<WT>::ToFloat(weights[start + j]));
      } else {
        denom = static_cast<float>(len);
      }
      v = denom == 0.f ? 0.f : __fmul_rn(v, __fdiv_rn(1.0f, denom));
    }
    if (out_dt == CUEMBED_F32)
      static_cast<float*>(out)[e] = v;
    else if (out_dt == CUEMBED_F16)
      static_cast<__half*>(out)[e] = __float2half_rn(v);
    else
      static_cast<__nv_bfloat16*>(out)[e] = __float2bfloat16_rn(v);
  }
}

int LaunchShardFinalize(const void* partial_f32, int n_samples, int embed_width,
                        int mode, const void* offsets, int off_type,
                        int num_hots, int sample0, const void* weights,
                        int weight_dtype, void* out, int out_dtype,
                        cudaStream_t stream) {
  if (n_samples < 0 || embed_width <= 0) return CUEMBED_ERR_ARGUMENT;
  if (mode != CUEMBED_SUM && mode != CUEMBED_MEAN) return CUEMBED_ERR_DTYPE;
  if (out_dtype < 0 || out_dtype > 2) return CUEMBED_ERR_DTYPE;
  if (n_samples == 0) return CUEMBED_OK;
  if (partial_f32 == nullptr || out == nullptr) return CUEMBED_ERR_ARGUMENT;
  const int64_t total = static_cast<int64_t>(n_samples) * embed_width;
  const int64_t ctas = (total + kCtaThreads - 1) / kCtaThreads;
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * 16;
  const int grid = static_cast<int>(ctas < cap ? ctas : cap);
  const int mean = mode == CUEMBED_MEAN;
  const int off64 = off_type == CUEMBED_I64;
  const float* p = static_cast<const float*>(partial_f32);
  if (weights == nullptr || weight_dtype == CUEMBED_F32)
    ShardFinalizeKernel<float><<<grid, kCtaThreads, 0, stream>>>(
        p, n_samples, embed_width, mean, offsets, off64, num_hots, sample0,
        static_cast<const float*>(weights), out, out_dtype);
  else if (weight_dtype == CUEMBED_F16)
    ShardFinalizeKernel<__half><<<grid, kCtaThreads, 0, stream>>>(
        p, n_samples, embed_width, mean, offsets, off64, num_hots, sample0,
        static_cast<const __half*>(weights), out, out_dtype);
  else
    ShardFinalizeKernel<__nv_bfloat16><<<grid, kCtaThreads, 0, stream>>>(
        p, n_samples, embed_width, mean, offsets, off64, num_hots, sample0,
        static_cast<const __nv_bfloat16*>(weights), out, out_dtype);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

}  // namespace cuembed_b200
