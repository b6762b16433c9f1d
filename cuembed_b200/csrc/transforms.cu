// transforms.cu -- index transforms for sm_100a: COO row-id extraction, the
// stable index transpose (a one-sweep LSD radix sort written here, no CUB) and
// the compressed-gradient index remap (single-pass look-back scan).
//
// Replaces cuembed/include/index_transforms.cuh:45-323 and the four helper
// kernels of cuembed/include/index_transforms_kernels.cuh:28-81; the reference
// delegates the sort and the scan to cub::DeviceRadixSort / DeviceScan /
// DeviceAdjacentDifference (index_transforms.cuh:108-136,160-195,287-322).
//
// Design of the transpose (stable sort of (key = table index, payload = sample
// id [, weight]) by key):
//   1. one histogram kernel reads the keys once and builds the 256-bin
//      histograms of EVERY byte position (4 or 8 of them);
//   2. one "one-sweep" pass per byte position.  Each pass kernel first inspects
//      the histograms: a byte position where all keys agree is skipped -- the
//      kernel returns immediately (10 M rows need 3 of the 4 / 8 passes) -- and
//      the ping-pong buffer roles are derived from the number of live passes,
//      so the host never reads anything back;
//   3. inside a pass each CTA takes a tile (dynamic tile id), ranks its keys
//      stably with warp-level match-any multisplit, publishes its per-digit
//      counts, resolves the counts of all earlier tiles by decoupled look-back,
//      reorders keys and payload through shared memory and writes runs of
//      equal digits contiguously.
//   The payload (sample id and weight) is carried as separate arrays: no
//   pack / unpack kernels (K6/K7 of the reference) and no tuple buffers.
// Keys are sorted as signed integers (the top byte is biased by 0x80), like
// cub::DeviceRadixSort::SortPairs on int32_t / int64_t.
//
// Integer / byte work bound by L2 and HBM traffic; no tensor cores.
#include "common.cuh"
#include "launch.h"

namespace cuembed_b200 {

// ---------------------------------------------------------------- row ids

template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    RowIdsFixedKernel(IdxT* __restrict__ row_ids, int64_t nnz, int num_hots,
                      bool vec_ok) {
  // nnz is an int in the API, so 32-bit unsigned arithmetic is enough.  Each
  // thread produces 16 bytes (one vector store) per iteration.
  constexpr uint32_t PER = 16 / sizeof(IdxT);
  const uint32_t n = static_cast<uint32_t>(nnz);
  const uint32_t h = static_cast<uint32_t>(num_hots);
  const uint32_t stride = gridDim.x * blockDim.x * PER;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * PER; i < n;
       i += stride) {
    uint32_t q = i / h;
    uint32_t r = i - q * h;
    IdxT v[PER];
#pragma unroll
    for (uint32_t k = 0; k < PER; ++k) {
      v[k] = static_cast<IdxT>(q);
      if (++r == h) {
        r = 0;
        ++q;
      }
    }
    if (vec_ok && i + PER <= n) {
      *reinterpret_cast<uint4*>(row_ids + i) = *reinterpret_cast<uint4*>(v);
    } else {
      for (uint32_t k = 0; i + k < n; ++k) row_ids[i + k] = v[k];
    }
  }
}

// One warp per sample: row_ids[offsets[b] .. offsets[b+1]) = b.
template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    RowIdsCsrKernel(const void* __restrict__ offsets, int off64, int batch,
                    IdxT* __restrict__ row_ids) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int b = warp; b < batch; b += nwarps) {
    const int64_t lo = LoadOffset(offsets, off64, b);
    const int64_t hi = LoadOffset(offsets, off64, b + 1);
    for (int64_t o = lo + lane; o < hi; o += 32)
      row_ids[o] = static_cast<IdxT>(b);
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    RowIdsIotaKernel(IdxT* __restrict__ row_ids, int64_t nnz) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       i < nnz; i += stride) {
    row_ids[i] = static_cast<IdxT>(i);
  }
}

namespace {
int StreamGrid(int64_t items_per_thread_total) {
  const int64_t ctas = (items_per_thread_total + kCtaThreads - 1) / kCtaThreads;
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * 8;
  return static_cast<int>(ctas < 1 ? 1 : (ctas < cap ? ctas : cap));
}
}  // namespace

int LaunchExtractRowIdsFixed(int batch_size, int num_hots, void* row_ids,
                             int idx_type, cudaStream_t stream) {
  if (batch_size < 0 || num_hots <= 0) return CUEMBED_ERR_ARGUMENT;
  const int64_t nnz = static_cast<int64_t>(batch_size) * num_hots;
  if (nnz == 0) return CUEMBED_OK;
  if (row_ids == nullptr) return CUEMBED_ERR_ARGUMENT;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(row_ids) & 15) == 0;
  const int grid = StreamGrid(nnz / 2);
  if (idx_type == CUEMBED_I64)
    RowIdsFixedKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<int64_t*>(row_ids), nnz, num_hots, vec_ok);
  else
    RowIdsFixedKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<int32_t*>(row_ids), nnz, num_hots, vec_ok);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

int LaunchExtractRowIdsCsr(const void* offsets, int off_type, int batch_size,
                           void* row_ids, int idx_type, cudaStream_t stream) {
  if (batch_size < 0) return CUEMBED_ERR_ARGUMENT;
  if (batch_size == 0) return CUEMBED_OK;
  if (offsets == nullptr || row_ids == nullptr) return CUEMBED_ERR_ARGUMENT;
  const int grid = StreamGrid(static_cast<int64_t>(batch_size) * 32);
  const int off64 = off_type == CUEMBED_I64;
  if (idx_type == CUEMBED_I64)
    RowIdsCsrKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
        offsets, off64, batch_size, static_cast<int64_t*>(row_ids));
  else
    RowIdsCsrKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
        offsets, off64, batch_size, static_cast<int32_t*>(row_ids));
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

int LaunchExtractRowIdsConcat(int nnz, void* row_ids, int idx_type,
                              cudaStream_t stream) {
  if (nnz < 0) return CUEMBED_ERR_ARGUMENT;
  if (nnz == 0) return CUEMBED_OK;
  if (row_ids == nullptr) return CUEMBED_ERR_ARGUMENT;
  const int grid = StreamGrid(nnz);
  if (idx_type == CUEMBED_I64)
    RowIdsIotaKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<int64_t*>(row_ids), nnz);
  else
    RowIdsIotaKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<int32_t*>(row_ids), nnz);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

// ------------------------------------------------------------- radix sort

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

template <typename KeyT>
__device__ __forceinline__ uint32_t DigitOf(KeyT key, int pos) {
  using U = typename std::conditional<sizeof(KeyT) == 8, uint64_t, uint32_t>::type;
  uint32_t d = static_cast<uint32_t>((static_cast<U>(key) >> (pos * kRadixBits)) &
                                     (kRadix - 1));
  // Signed order: bias the most significant byte.
  if (pos == static_cast<int>(sizeof(KeyT)) - 1) d ^= 0x80u;
  return d;
}

// Histograms of every byte position in one read of the keys.
// hist layout: [sizeof(KeyT)][256] uint32.  Each thread takes 16 keys per
// iteration (16 independent coalesced loads in flight); a byte position on
// which all 512 keys of the warp agree (the usual case for the high bytes)
// costs one aggregated shared-memory atomic instead of 512.
template <typename KeyT>
__global__ void __launch_bounds__(kCtaThreads)
    RadixHistKernel(const KeyT* __restrict__ keys, int nnz,
                    uint32_t* __restrict__ hist,
                    uint32_t* __restrict__ done_counter,
                    int* __restrict__ plan) {
  constexpr int ND = sizeof(KeyT);
  constexpr int ITEMS = 16;
  constexpr int TILE = ITEMS * kCtaThreads;
  using U = typename std::conditional<sizeof(KeyT) == 8, uint64_t, uint32_t>::type;
  __shared__ uint32_t sh[ND * kRadix];
  for (int i = threadIdx.x; i < ND * kRadix; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int num_tiles = (nnz + TILE - 1) / TILE;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int base = t * TILE + threadIdx.x;
    KeyT key[ITEMS];
    bool full = (t + 1) * static_cast<int64_t>(TILE) <= nnz;
    if (full) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) key[i] = __ldg(keys + base + i * kCtaThreads);
      // bits on which the keys of this warp differ
      U diff = 0;
#pragma unroll
      for (int i = 1; i < ITEMS; ++i)
        diff |= static_cast<U>(key[i]) ^ static_cast<U>(key[0]);
      const U k0 = static_cast<U>(
          sizeof(KeyT) == 8
              ? static_cast<U>(__shfl_sync(0xffffffffu,
                                           static_cast<long long>(key[0]), 0))
              : static_cast<U>(__shfl_sync(0xffffffffu,
                                           static_cast<int>(key[0]), 0)));
      diff |= static_cast<U>(key[0]) ^ k0;
      uint32_t dlo = static_cast<uint32_t>(diff);
      uint32_t dhi = sizeof(KeyT) == 8 ? static_cast<uint32_t>(static_cast<uint64_t>(diff) >> 32) : 0u;
      dlo = __reduce_or_sync(0xffffffffu, dlo);
      if (sizeof(KeyT) == 8) dhi = __reduce_or_sync(0xffffffffu, dhi);
#pragma unroll
      for (int p = 0; p < ND; ++p) {
        const uint32_t byte_diff =
            (p < 4 ? (dlo >> (8 * p)) : (dhi >> (8 * (p - 4)))) & 0xffu;
        if (byte_diff == 0) {
          if (lane == 0)
            atomicAdd(&sh[p * kRadix + DigitOf<KeyT>(key[0], p)], 32u * ITEMS);
        } else {
#pragma unroll
          for (int i = 0; i < ITEMS; ++i)
            atomicAdd(&sh[p * kRadix + DigitOf<KeyT>(key[i], p)], 1u);
        }
      }
    } else {
      for (int i = 0; i < ITEMS; ++i) {
        const int g = base + i * kCtaThreads;
        if (g < nnz) {
          const KeyT k = __ldg(keys + g);
#pragma unroll
          for (int p = 0; p < ND; ++p)
            atomicAdd(&sh[p * kRadix + DigitOf<KeyT>(k, p)], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ND * kRadix; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c != 0) atomicAdd(&hist[i], c);
  }

  // ---- the last CTA to finish turns the histograms into the pass plan:
  //   plan[0] = bit mask of live byte positions (a position where every key has
  //             the same byte needs no pass), plan[1] = number of live passes;
  //   hist[p][d] is replaced by its exclusive prefix sum over d (the global
  //   start of digit d in pass p).  One warp per byte position.
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int warp = threadIdx.x >> 5;
  __shared__ int s_live[ND];
  if (warp < ND) {
    const int p = warp;
    uint32_t c[8];
    uint32_t sum = 0;
    bool degenerate = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      c[k] = __ldcg(&hist[p * kRadix + lane * 8 + k]);
      degenerate |= c[k] == static_cast<uint32_t>(nnz);
      sum += c[k];
    }
    uint32_t scan = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, scan, o);
      if (lane >= o) scan += t;
    }
    uint32_t run = scan - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      hist[p * kRadix + lane * 8 + k] = run;
      run += c[k];
    }
    const bool any_deg = __any_sync(0xffffffffu, degenerate);
    if (lane == 0) s_live[p] = any_deg ? 0 : 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int mask = 0;
    for (int p = 0; p < ND; ++p) mask |= s_live[p] << p;
    plan[0] = mask;
    plan[1] = __popc(static_cast<unsigned>(mask));
  }
}

// Look-back status word: 2 flag bits + 30 value bits (nnz < 2^30).
constexpr uint32_t kFlagAgg = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

struct SortArgs {
  const void* keys_in;
  const void* vals_in;
  const void* w_in;
  void* keys_out;
  void* vals_out;
  void* w_out;
  void* keys_tmp;
  void* vals_tmp;
  void* w_tmp;
  const uint32_t* digit_base;  // [ND][256] exclusive digit starts (plan kernel)
  const int* plan;             // live mask, number of live passes
  uint32_t* lookback;          // [ND][num_tiles][256], zeroed
  int tile_pitch;              // unused
  uint32_t* tile_counters;     // [ND], zeroed
  int nnz;
  int num_tiles;
  int pass;  // byte position handled by this launch
  // > 0: the payload of the input is not read but synthesised in the first live
  // pass as position / synth_hots (fixed-hotness sample ids: the fused
  // transpose entry point, no row-id array)
  int synth_hots;
};

__device__ __forceinline__ uint32_t LdVolatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void StVolatile(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v));
}
__device__ __forceinline__ uint4 LdVolatile4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// One pass of the stable LSD sort over byte position a.pass.
//   KeyT: int32_t / int64_t (payload sample ids have the same type)
//   WBYTES: 0 (no weights), 2 or 4 (weight element size)
//   ITEMS: keys per thread; tile = ITEMS * 256 keys.
// Persistent CTAs take tiles from a counter; a tile only ever waits for tiles
// with a smaller id, which were taken earlier, so the look-back cannot
// deadlock.
template <typename KeyT, int WBYTES, int ITEMS>
__global__ void __launch_bounds__(kCtaThreads, (ITEMS <= 8 ? 4 : (ITEMS <= 12 ? 3 : 2)))
    RadixPassKernel(const SortArgs a) {
  constexpr int TILE = ITEMS * kCtaThreads;
  using WT = typename std::conditional<WBYTES == 4, uint32_t, uint16_t>::type;

  __shared__ uint32_t s_warp_cnt[kWarpsPerCta][kRadix];  // per-warp digit counts
  __shared__ uint32_t s_tile_excl[kRadix];  // digit start inside the tile
  __shared__ int32_t s_goff[kRadix];        // global start - tile start
  __shared__ uint32_t s_scan[kWarpsPerCta];
  __shared__ int s_tile;
  // Keys and sample ids travel through shared memory together, as pairs.
  struct alignas(2 * sizeof(KeyT)) Pair {
    KeyT k, v;
  };
  extern __shared__ __align__(16) unsigned char s_exch_raw[];
  Pair* s_exch = reinterpret_cast<Pair*>(s_exch_raw);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;

  const int live_mask = a.plan[0];
  const int total_live = a.plan[1];
  const bool live = ((live_mask >> a.pass) & 1) != 0;
  const int ordinal = __popc(live_mask & ((1 << a.pass) - 1));

  // No live pass at all (all keys equal): pass 0 degenerates to a copy.
  if (total_live == 0) {
    if (a.pass != 0) return;
    const KeyT* kin = static_cast<const KeyT*>(a.keys_in);
    const KeyT* vin = static_cast<const KeyT*>(a.vals_in);
    KeyT* kout = static_cast<KeyT*>(a.keys_out);
    KeyT* vout = static_cast<KeyT*>(a.vals_out);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kCtaThreads + tid;
         i < a.nnz; i += static_cast<int64_t>(gridDim.x) * kCtaThreads) {
      kout[i] = kin[i];
      vout[i] = a.synth_hots > 0 ? static_cast<KeyT>(i / a.synth_hots) : vin[i];
      if constexpr (WBYTES != 0)
        static_cast<WT*>(a.w_out)[i] = static_cast<const WT*>(a.w_in)[i];
    }
    return;
  }
  if (!live) return;

  // Ping-pong: the last live pass writes the caller's output arrays.
  const bool dst_is_out = ((total_live - 1 - ordinal) & 1) == 0;
  const KeyT* kin;
  const KeyT* vin;
  const WT* win;
  if (ordinal == 0) {
    kin = static_cast<const KeyT*>(a.keys_in);
    vin = static_cast<const KeyT*>(a.vals_in);
    win = static_cast<const WT*>(a.w_in);
  } else if (dst_is_out) {
    kin = static_cast<const KeyT*>(a.keys_tmp);
    vin = static_cast<const KeyT*>(a.vals_tmp);
    win = static_cast<const WT*>(a.w_tmp);
  } else {
    kin = static_cast<const KeyT*>(a.keys_out);
    vin = static_cast<const KeyT*>(a.vals_out);
    win = static_cast<const WT*>(a.w_out);
  }
  KeyT* kout = static_cast<KeyT*>(dst_is_out ? a.keys_out : a.keys_tmp);
  KeyT* vout = static_cast<KeyT*>(dst_is_out ? a.vals_out : a.vals_tmp);
  WT* wout = static_cast<WT*>(dst_is_out ? a.w_out : a.w_tmp);

  const uint32_t digit_base = a.digit_base[a.pass * kRadix + tid];
  uint32_t* lb = a.lookback +
                 (static_cast<size_t>(a.pass) * a.num_tiles) * kRadix + tid;
  const int warp_base = warp * (32 * ITEMS);

  // The id of the next tile is fetched while the current tile is processed
  // (the atomic's round trip is off the critical path).
  int next_tile = 0;
  if (tid == 0)
    next_tile = static_cast<int>(atomicAdd(&a.tile_counters[a.pass], 1u));
  while (true) {
    __syncthreads();  // previous tile done with the shared arrays
    if (tid == 0) s_tile = next_tile;
    for (int i = tid; i < kWarpsPerCta * kRadix; i += kCtaThreads)
      (&s_warp_cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = s_tile;
    if (tile >= a.num_tiles) break;
    if (tid == 0)
      next_tile = static_cast<int>(atomicAdd(&a.tile_counters[a.pass], 1u));
    const int tile_base = tile * TILE;
    const int tile_n = min(TILE, a.nnz - tile_base);

    // ---- load keys and sample ids together: the payload's load latency
    // hides behind the ranking instead of following the key exchange.  (A
    // software-pipelined variant that prefetched the NEXT tile was measured
    // slower because it delays the publication of the tile's digit counts,
    // profiles/r01_notes.md.)
    KeyT key[ITEMS];
    KeyT val[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int local = warp_base + i * 32 + lane;
      key[i] = local < tile_n ? kin[tile_base + local] : KeyT(0);
    }
    if (ordinal == 0 && a.synth_hots > 0) {
      // sample id = position / hotness: one division per tile and thread, then
      // 32 positions further per item
      const uint32_t hots = static_cast<uint32_t>(a.synth_hots);
      const uint32_t g0 = static_cast<uint32_t>(tile_base + warp_base + lane);
      uint32_t q = g0 / hots;
      uint32_t r = g0 - q * hots;
      const uint32_t step_q = 32u / hots;
      const uint32_t step_r = 32u - step_q * hots;
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        val[i] = static_cast<KeyT>(q);
        q += step_q;
        r += step_r;
        if (r >= hots) {
          r -= hots;
          ++q;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int local = warp_base + i * 32 + lane;
        val[i] = local < tile_n ? vin[tile_base + local] : KeyT(0);
      }
    }

    // ---- stable ranking inside the warp.  Items are warp-striped so that
    // (warp, item, lane) is input order.  The set of lanes holding the same
    // digit is built from 8 ballots (one per digit bit) -- plain vote / logic
    // instructions; match.any was measured far slower on this part
    // (profiles/r01_notes.md).
    uint32_t rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int local = warp_base + i * 32 + lane;
      const bool valid = local < tile_n;
      const uint32_t d = DigitOf<KeyT>(key[i], a.pass);
      unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int bit = 0; bit < kRadixBits; ++bit) {
        const bool set = ((d >> bit) & 1u) != 0u;
        const unsigned m = __ballot_sync(0xffffffffu, set);
        peers &= set ? m : ~m;
      }
      if (!valid) peers = 0;  // lanes past the end of the tile drop out
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && lane == leader)
        old = atomicAdd(&s_warp_cnt[warp][d], __popc(peers));
      old = __shfl_sync(0xffffffffu, old, leader < 0 ? 0 : leader);
      rank[i] = old + __popc(peers & lt_mask);
    }
    __syncthreads();

    // ---- per digit (thread tid owns digit tid): exclusive scan over warps,
    //      tile total, publish, look-back over earlier tiles
    uint32_t tile_count = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) {
      const uint32_t c = s_warp_cnt[w][tid];
      s_warp_cnt[w][tid] = tile_count;
      tile_count += c;
    }
    StVolatile(&lb[static_cast<size_t>(tile) * kRadix],
               (tile == 0 ? kFlagPrefix : kFlagAgg) | tile_count);

    // tile-local exclusive scan of the digit counts (block scan, 256 owners)
    uint32_t tscan = tile_count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, tscan, o);
      if (lane >= o) tscan += t;
    }
    if (lane == 31) s_scan[warp] = tscan;

    // Decoupled look-back, kLookWin predecessors per round: their status words
    // are loaded together (independent loads, one L2 round trip) and consumed
    // in order.  The serial one-hop-per-round-trip walk was the top stall of
    // the pass (19 hops per tile on average, half of all warp samples,
    // profiles/r01_notes.md).
    uint32_t exclusive = 0;
    if (tile > 0) {
// predecessors per look-back round.  With 12 keys per thread and three CTAs per
// SM the predecessors publish early; measured at C2 (transpose stage, ms):
// 1 -> 0.1609, 2 -> 0.1568, 3 -> 0.1592, 4 -> 0.158, 6 -> 0.1588, 8 -> 0.163,
// 16 -> 0.1715, 32 -> 0.207 (wide windows fetch status words that are not
// published yet and poll them again).
#ifndef RADIX_LOOK_WIN
#define RADIX_LOOK_WIN 2
#endif
      constexpr int kLookWin = RADIX_LOOK_WIN;
      int prev = tile - 1;
      bool done = false;
      while (!done) {
        uint32_t st[kLookWin];
#pragma unroll
        for (int u = 0; u < kLookWin; ++u) {
          const int p = prev - u;
          st[u] = LdVolatile(&lb[static_cast<size_t>(p < 0 ? 0 : p) * kRadix]);
        }
#pragma unroll
        for (int u = 0; u < kLookWin; ++u) {
          if (!done) {
            uint32_t sw = st[u];
            while ((sw & ~kValueMask) == 0)
              sw = LdVolatile(&lb[static_cast<size_t>(prev - u) * kRadix]);
            exclusive += sw & kValueMask;
            if ((sw & kFlagPrefix) != 0) done = true;
          }
        }
        prev -= kLookWin;
      }
      StVolatile(&lb[static_cast<size_t>(tile) * kRadix],
                 kFlagPrefix | (exclusive + tile_count));
    }
    __syncthreads();
    uint32_t tbase = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w)
      if (w < warp) tbase += s_scan[w];
    const uint32_t t_excl = tbase + tscan - tile_count;
    s_tile_excl[tid] = t_excl;
    s_goff[tid] = static_cast<int32_t>(digit_base + exclusive) -
                  static_cast<int32_t>(t_excl);
    __syncthreads();

    // ---- final position of every item inside the tile; pairs through smem
    uint32_t pos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int local = warp_base + i * 32 + lane;
      if (local < tile_n) {
        const uint32_t d = DigitOf<KeyT>(key[i], a.pass);
        pos[i] = s_tile_excl[d] + s_warp_cnt[warp][d] + rank[i];
        Pair pr;
        pr.k = key[i];
        pr.v = val[i];
        s_exch[pos[i]] = pr;
      } else {
        pos[i] = 0;
      }
    }
    __syncthreads();
    int32_t gaddr[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int p = tid + i * kCtaThreads;
      if (p < tile_n) {
        const Pair pr = s_exch[p];
        gaddr[i] = s_goff[DigitOf<KeyT>(pr.k, a.pass)] + p;
        kout[gaddr[i]] = pr.k;
        vout[gaddr[i]] = pr.v;
      } else {
        gaddr[i] = -1;
      }
    }

    // ---- payload: weights
    if constexpr (WBYTES != 0) {
      __syncthreads();
      WT* s_w = reinterpret_cast<WT*>(s_exch_raw);
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int local = warp_base + i * 32 + lane;
        if (local < tile_n) s_w[pos[i]] = win[tile_base + local];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int p = tid + i * kCtaThreads;
        if (p < tile_n) wout[gaddr[i]] = s_w[p];
      }
    }

  }
}

namespace {

struct SortLayout {
  int items;
  int tile;
  int num_tiles;
  size_t hist_off, counters_off, lookback_off, zero_bytes;
  size_t plan_off, keys_tmp_off, vals_tmp_off, w_tmp_off, total;
  int tile_pitch;
};

SortLayout MakeSortLayout(int nnz, int idx_type, int wbytes) {
  SortLayout L;
  const int nd = static_cast<int>(IndexSize(idx_type));
  static const int items_env = EnvInt("CUEMBED_SORT_ITEMS", 0);
  // Larger tiles once there are enough of them to keep the SMs busy: fewer
  // per-tile fixed costs (counter reset, digit scan, look-back).  Smaller
  // problems keep 2048-key tiles for parallelism.
  // 32-bit keys: 12 keys per thread (3072-key tiles, 80 registers, THREE CTAs =
  // 24 warps per SM).  Measured at C2 (graph replay, same box, transpose stage):
  // 8 keys 0.173, 10 keys 0.170, 12 keys 0.161-0.163, 14 keys 0.168, 16 keys
  // 0.172 ms -- 16 keys need 128 registers (16 warps per SM), 8 and 10 pay the
  // per-tile fixed costs more often.  64-bit keys (they would spill at 80
  // registers with 12): 8 keys per thread, measured on the C3 shape (int64
  // indices, weights, 4.2 M pairs): 8 -> 0.2915, 12 -> 0.314, 16 -> 0.3085 ms.
  L.items = items_env > 0 ? items_env : (nnz >= (2 << 20) && nd == 4 ? 12 : 8);
  if (L.items != 8 && L.items != 12 && L.items != 16) L.items = 8;
  L.tile = L.items * kCtaThreads;
  L.num_tiles = nnz > 0 ? (nnz + L.tile - 1) / L.tile : 0;
  size_t off = 0;
  L.hist_off = off;
  off += static_cast<size_t>(nd) * kRadix * sizeof(uint32_t);
  L.counters_off = off;  // nd tile counters + 1 "CTAs done" counter
  off += AlignUp(static_cast<size_t>(nd + 1) * sizeof(uint32_t), 128);
  L.lookback_off = off;
  L.tile_pitch = L.num_tiles;
  off += static_cast<size_t>(nd) * kRadix * L.tile_pitch * sizeof(uint32_t);
  L.zero_bytes = off;
  off = AlignUp(off, 256);
  L.plan_off = off;
  off += 256;
  L.keys_tmp_off = off;
  off += AlignUp(static_cast<size_t>(nnz) * nd, 256);
  L.vals_tmp_off = off;
  off += AlignUp(static_cast<size_t>(nnz) * nd, 256);
  L.w_tmp_off = off;
  off += AlignUp(static_cast<size_t>(nnz) * wbytes, 256);
  L.total = off;
  return L;
}

template <typename KeyT, int WBYTES, int ITEMS>
void LaunchPasses(const SortArgs& base, cudaStream_t stream) {
  constexpr int ND = sizeof(KeyT);
  const size_t smem = static_cast<size_t>(ITEMS) * kCtaThreads * 2 * sizeof(KeyT);
  auto kernel = RadixPassKernel<KeyT, WBYTES, ITEMS>;
  // function attributes and occupancy are per device: one cache slot per device
  static PerDeviceInt ctas_per_sm_cache;
  int ctas_per_sm = ctas_per_sm_cache.Get();
  if (ctas_per_sm == 0) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(smem));
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kCtaThreads, smem);
    ctas_per_sm = n > 0 ? n : 1;
    ctas_per_sm_cache.Set(ctas_per_sm);
  }
  const int cap = GetDeviceInfo().sm_count * ctas_per_sm;
  const int grid = base.num_tiles < cap ? base.num_tiles : cap;
  for (int p = 0; p < ND; ++p) {
    SortArgs a = base;
    a.pass = p;
    kernel<<<grid, kCtaThreads, smem, stream>>>(a);
  }
  CountLaunch(ND);
}

template <typename KeyT, int WBYTES>
void LaunchPassesItems(const SortArgs& a, int items, cudaStream_t stream) {
  if (items == 16)
    LaunchPasses<KeyT, WBYTES, 16>(a, stream);
  else if (items == 12)
    LaunchPasses<KeyT, WBYTES, 12>(a, stream);
  else
    LaunchPasses<KeyT, WBYTES, 8>(a, stream);
}

template <typename KeyT>
void LaunchSortTyped(const SortArgs& a, uint32_t* hist, int* plan, int wbytes,
                     int items, cudaStream_t stream) {
  {
    const int ctas = CeilDiv(a.nnz, kCtaThreads * 16);
    const int cap = GetDeviceInfo().sm_count * 4;
    RadixHistKernel<KeyT><<<ctas < cap ? ctas : cap, kCtaThreads, 0, stream>>>(
        static_cast<const KeyT*>(a.keys_in), a.nnz, hist,
        a.tile_counters + sizeof(KeyT), plan);
    CountLaunch();
  }
  if (wbytes == 0)
    LaunchPassesItems<KeyT, 0>(a, items, stream);
  else if (wbytes == 2)
    LaunchPassesItems<KeyT, 2>(a, items, stream);
  else
    LaunchPassesItems<KeyT, 4>(a, items, stream);
}

}  // namespace

int LaunchTransposeImpl(const void* rows, int synth_hots, const void* cols,
                        const void* weights, int weight_dtype, int nnz,
                        int idx_type, void* transpose_rows, void* transpose_cols,
                        void* transpose_weights, char* work, size_t* lwork,
                        cudaStream_t stream);

int LaunchTranspose(const void* rows, const void* cols, const void* weights,
                    int weight_dtype, int nnz, int idx_type,
                    void* transpose_rows, void* transpose_cols,
                    void* transpose_weights, char* work, size_t* lwork,
                    cudaStream_t stream) {
  return LaunchTransposeImpl(rows, 0, cols, weights, weight_dtype, nnz, idx_type,
                             transpose_rows, transpose_cols, transpose_weights,
                             work, lwork, stream);
}

// Fixed-hotness COO in one call: the sample id of position i is i / num_hots, so
// the row-id array (ExtractRowIdsFromFixed) is neither written nor read; the
// first sort pass synthesises it.
int LaunchTransposeFixed(const void* cols, int batch_size, int num_hots,
                         const void* weights, int weight_dtype, int idx_type,
                         void* transpose_rows, void* transpose_cols,
                         void* transpose_weights, char* work, size_t* lwork,
                         cudaStream_t stream) {
  if (batch_size < 0 || num_hots <= 0) return CUEMBED_ERR_ARGUMENT;
  const int64_t nnz = static_cast<int64_t>(batch_size) * num_hots;
  if (nnz >= (1 << 30)) return CUEMBED_ERR_NNZ_LIMIT;
  return LaunchTransposeImpl(nullptr, num_hots, cols, weights, weight_dtype,
                             static_cast<int>(nnz), idx_type, transpose_rows,
                             transpose_cols, transpose_weights, work, lwork,
                             stream);
}

int LaunchTransposeImpl(const void* rows, int synth_hots, const void* cols,
                        const void* weights, int weight_dtype, int nnz,
                        int idx_type, void* transpose_rows, void* transpose_cols,
                        void* transpose_weights, char* work, size_t* lwork,
                        cudaStream_t stream) {
  if (lwork == nullptr || nnz < 0) return CUEMBED_ERR_ARGUMENT;
  if (idx_type < 0 || idx_type > 1) return CUEMBED_ERR_DTYPE;
  if (weights != nullptr && (weight_dtype < 0 || weight_dtype > 2))
    return CUEMBED_ERR_DTYPE;
  if (nnz >= (1 << 30)) return CUEMBED_ERR_NNZ_LIMIT;
  const int wbytes =
      weights != nullptr ? static_cast<int>(ElemSize(weight_dtype)) : 0;
  const SortLayout L = MakeSortLayout(nnz, idx_type, wbytes);
  // Workspace query (cuembed/include/index_transforms.cuh:121-124).
  if (work == nullptr) {
    *lwork = L.total > 0 ? L.total : 256;
    return CUEMBED_OK;
  }
  if (*lwork < L.total) return CUEMBED_ERR_WORKSPACE;
  if (nnz == 0) return CUEMBED_OK;
  if ((rows == nullptr && synth_hots <= 0) || cols == nullptr ||
      transpose_rows == nullptr || transpose_cols == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  if (weights != nullptr && transpose_weights == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(work) & 15) != 0)
    return CUEMBED_ERR_ARGUMENT;

  if (cudaMemsetAsync(work, 0, L.zero_bytes, stream) != cudaSuccess)
    return CUEMBED_ERR_CUDA;

  SortArgs a;
  a.keys_in = cols;  // sort key = table index
  a.vals_in = rows;  // payload = sample id
  a.w_in = weights;
  a.keys_out = transpose_rows;
  a.vals_out = transpose_cols;
  a.w_out = transpose_weights;
  a.keys_tmp = work + L.keys_tmp_off;
  a.vals_tmp = work + L.vals_tmp_off;
  a.w_tmp = work + L.w_tmp_off;
  uint32_t* hist = reinterpret_cast<uint32_t*>(work + L.hist_off);
  int* plan = reinterpret_cast<int*>(work + L.plan_off);
  a.digit_base = hist;
  a.plan = plan;
  a.lookback = reinterpret_cast<uint32_t*>(work + L.lookback_off);
  a.tile_counters = reinterpret_cast<uint32_t*>(work + L.counters_off);
  a.nnz = nnz;
  a.num_tiles = L.num_tiles;
  a.tile_pitch = L.tile_pitch;
  a.pass = 0;
  a.synth_hots = synth_hots;
  if (idx_type == CUEMBED_I64)
    LaunchSortTyped<int64_t>(a, hist, plan, wbytes, L.items, stream);
  else
    LaunchSortTyped<int32_t>(a, hist, plan, wbytes, L.items, stream);
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

// ------------------------------------------- compressed gradient indices

// remapped[i] = number of positions j in (0, i] with idx[j] != idx[j-1].
// Two kernels over the same fixed partition of the array into `parts` ranges:
//   1. count the run starts of every range;
//   2. every CTA sums the counts of the ranges before its own (<= 2048 values,
//      one block reduction) and then scans its range.
// No cross-CTA waiting (a single-pass look-back scan was measured latency-bound
// at this size: 36 us for 4 M indices, profiles/r01_notes.md); the second read
// of the indices comes from L2.
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanItems * kCtaThreads;  // 2048
constexpr int kScanMaxParts = 2048;

// Thread-blocked: kScanItems consecutive elements starting at `base` (a
// multiple of kScanItems), fetched with 16-byte loads; the element before the
// thread's range comes from the neighbouring lane.
template <typename IdxT>
__device__ __forceinline__ uint32_t LoadFlags(const IdxT* __restrict__ idx,
                                              int64_t base, int64_t nnz,
                                              bool vec_ok, uint32_t* flag) {
  constexpr int PER = 16 / sizeof(IdxT);
  IdxT cur[kScanItems];
  if (vec_ok && base + kScanItems <= nnz) {
#pragma unroll
    for (int k = 0; k < kScanItems / PER; ++k) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(idx + base) + k);
      *reinterpret_cast<uint4*>(&cur[k * PER]) = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
      cur[i] = base + i < nnz ? __ldg(idx + base + i) : IdxT(0);
  }
  // last element of the previous thread's range
  IdxT prev;
  if constexpr (sizeof(IdxT) == 8)
    prev = static_cast<IdxT>(__shfl_up_sync(
        0xffffffffu, static_cast<long long>(cur[kScanItems - 1]), 1));
  else
    prev = __shfl_up_sync(0xffffffffu, cur[kScanItems - 1], 1);
  if ((threadIdx.x & 31) == 0)
    prev = (base > 0 && base - 1 < nnz) ? __ldg(idx + base - 1) : IdxT(0);
  uint32_t local = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const int64_t g = base + i;
    flag[i] = (g > 0 && g < nnz && cur[i] != prev) ? 1u : 0u;
    local += flag[i];
    prev = cur[i];
  }
  return local;
}

template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    CompressCountKernel(const IdxT* __restrict__ idx, int nnz, int chunks_per_part,
                        bool vec_ok, uint32_t* __restrict__ part_counts) {
  __shared__ uint32_t s_warp[kWarpsPerCta];
  const int tid = threadIdx.x;
  const int64_t part_begin =
      static_cast<int64_t>(blockIdx.x) * chunks_per_part * kScanChunk;
  uint32_t local = 0;
  for (int c = 0; c < chunks_per_part; ++c) {
    const int64_t chunk_begin = part_begin + static_cast<int64_t>(c) * kScanChunk;
    if (chunk_begin >= nnz) break;  // uniform for the CTA
    const int64_t base = chunk_begin + tid * kScanItems;
    uint32_t flag[kScanItems];
    local += LoadFlags<IdxT>(idx, base, nnz, vec_ok, flag);
  }
  local = __reduce_add_sync(0xffffffffu, local);
  if ((tid & 31) == 0) s_warp[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) total += s_warp[w];
    part_counts[blockIdx.x] = total;
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    CompressScanKernel(const IdxT* __restrict__ idx, int nnz, int chunks_per_part,
                       bool vec_ok, const uint32_t* __restrict__ part_counts,
                       IdxT* __restrict__ remapped) {
  __shared__ unsigned long long s_red[kWarpsPerCta];
  __shared__ uint32_t s_warp[kWarpsPerCta];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  // run starts in all earlier parts
  unsigned long long before = 0;
  for (int p = tid; p < static_cast<int>(blockIdx.x); p += kCtaThreads)
    before += part_counts[p];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0) s_red[warp] = before;
  __syncthreads();
  unsigned long long carry = 0;
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) carry += s_red[w];

  const int64_t part_begin =
      static_cast<int64_t>(blockIdx.x) * chunks_per_part * kScanChunk;
  for (int c = 0; c < chunks_per_part; ++c) {
    const int64_t base = part_begin + static_cast<int64_t>(c) * kScanChunk +
                         tid * kScanItems;
    if (part_begin + static_cast<int64_t>(c) * kScanChunk >= nnz) break;
    uint32_t flag[kScanItems];
    const uint32_t local = LoadFlags<IdxT>(idx, base, nnz, vec_ok, flag);
    uint32_t scan = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, scan, o);
      if (lane >= o) scan += t;
    }
    __syncthreads();  // s_warp reuse
    if (lane == 31) s_warp[warp] = scan;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) {
      if (w < warp) wbase += s_warp[w];
      total += s_warp[w];
    }
    unsigned long long running = carry + wbase + scan - local;
    IdxT outv[kScanItems];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      running += flag[i];
      outv[i] = static_cast<IdxT>(running);
    }
    if (vec_ok && base + kScanItems <= nnz) {
      constexpr int PER = 16 / sizeof(IdxT);
#pragma unroll
      for (int k = 0; k < kScanItems / PER; ++k)
        reinterpret_cast<uint4*>(remapped + base)[k] =
            *reinterpret_cast<uint4*>(&outv[k * PER]);
    } else {
#pragma unroll
      for (int i = 0; i < kScanItems; ++i)
        if (base + i < nnz) remapped[base + i] = outv[i];
    }
    carry += total;
  }
}

int LaunchCompressedGradIndices(const void* indices, int idx_type, int nnz,
                                void* remapped, char* work, size_t* lwork,
                                cudaStream_t stream) {
  if (lwork == nullptr || nnz < 0) return CUEMBED_ERR_ARGUMENT;
  if (idx_type < 0 || idx_type > 1) return CUEMBED_ERR_DTYPE;
  const int num_chunks = nnz > 0 ? (nnz + kScanChunk - 1) / kScanChunk : 0;
  const int chunks_per_part =
      num_chunks > 0 ? (num_chunks + kScanMaxParts - 1) / kScanMaxParts : 1;
  const int parts =
      num_chunks > 0 ? (num_chunks + chunks_per_part - 1) / chunks_per_part : 0;
  const size_t need = 256 + static_cast<size_t>(parts) * sizeof(uint32_t);
  if (work == nullptr) {
    *lwork = need;
    return CUEMBED_OK;
  }
  if (*lwork < need) return CUEMBED_ERR_WORKSPACE;
  if (nnz == 0) return CUEMBED_OK;
  if (indices == nullptr || remapped == nullptr) return CUEMBED_ERR_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(work) & 3) != 0) return CUEMBED_ERR_ARGUMENT;
  // 16-byte vector loads / stores only when both arrays allow them.
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(indices) |
                        reinterpret_cast<uintptr_t>(remapped)) & 15) == 0;
  uint32_t* counts = reinterpret_cast<uint32_t*>(work);
  if (idx_type == CUEMBED_I64) {
    CompressCountKernel<int64_t><<<parts, kCtaThreads, 0, stream>>>(
        static_cast<const int64_t*>(indices), nnz, chunks_per_part, vec_ok, counts);
    CompressScanKernel<int64_t><<<parts, kCtaThreads, 0, stream>>>(
        static_cast<const int64_t*>(indices), nnz, chunks_per_part, vec_ok, counts,
        static_cast<int64_t*>(remapped));
  } else {
    CompressCountKernel<int32_t><<<parts, kCtaThreads, 0, stream>>>(
        static_cast<const int32_t*>(indices), nnz, chunks_per_part, vec_ok, counts);
    CompressScanKernel<int32_t><<<parts, kCtaThreads, 0, stream>>>(
        static_cast<const int32_t*>(indices), nnz, chunks_per_part, vec_ok, counts,
        static_cast<int32_t*>(remapped));
  }
  CountLaunch(2);
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

}  // namespace cuembed_b200
