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
    RowIdsFixedKernel(IdxT* __restrict__ row_ids, int64_t nnz, int num_hots) {
  // nnz is an int in the API, so 32-bit unsigned division is enough.
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t n = static_cast<uint32_t>(nnz);
  const uint32_t h = static_cast<uint32_t>(num_hots);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    row_ids[i] = static_cast<IdxT>(i / h);
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
  const int grid = StreamGrid(nnz);
  if (idx_type == CUEMBED_I64)
    RowIdsFixedKernel<int64_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<int64_t*>(row_ids), nnz, num_hots);
  else
    RowIdsFixedKernel<int32_t><<<grid, kCtaThreads, 0, stream>>>(
        static_cast<int32_t*>(row_ids), nnz, num_hots);
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
// hist layout: [sizeof(KeyT)][256] uint32.
template <typename KeyT>
__global__ void __launch_bounds__(kCtaThreads)
    RadixHistKernel(const KeyT* __restrict__ keys, int nnz,
                    uint32_t* __restrict__ hist) {
  constexpr int ND = sizeof(KeyT);
  __shared__ uint32_t sh[ND * kRadix];
  for (int i = threadIdx.x; i < ND * kRadix; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * blockDim.x;
  const int first = blockIdx.x * blockDim.x + threadIdx.x;
  // Warp-synchronous trip count (the condition only depends on the warp's
  // first index) so that the shuffles below are convergent.
  for (int i = first; i - lane < nnz; i += stride) {
    const bool valid = i < nnz;
    const KeyT key = valid ? __ldg(keys + i) : KeyT(0);
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int p = 0; p < ND; ++p) {
      const uint32_t d = DigitOf<KeyT>(key, p);
      // Constant (or warp-uniform) digits are the common case for the high
      // bytes: one aggregated add instead of 32 same-address atomics.
      const uint32_t d0 = __shfl_sync(0xffffffffu, d, __ffs(vmask) - 1);
      const bool uniform = __all_sync(0xffffffffu, !valid || d == d0);
      if (uniform) {
        if (lane == 0) atomicAdd(&sh[p * kRadix + d0], __popc(vmask));
      } else if (valid) {
        atomicAdd(&sh[p * kRadix + d], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ND * kRadix; i += blockDim.x) {
    const uint32_t c = sh[i];
    if (c != 0) atomicAdd(&hist[i], c);
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
  const uint32_t* hist;     // [ND][256]
  uint32_t* lookback;       // [ND][num_tiles][256], zeroed
  uint32_t* tile_counters;  // [ND], zeroed
  int nnz;
  int num_tiles;
  int pass;  // byte position handled by this launch
};

__device__ __forceinline__ uint32_t LdVolatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void StVolatile(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v));
}

// One pass of the stable LSD sort over byte position a.pass.
//   KeyT: int32_t / int64_t (payload sample ids have the same type)
//   WBYTES: 0 (no weights), 2 or 4 (weight element size)
//   ITEMS: keys per thread; tile = ITEMS * 256 keys.
template <typename KeyT, int WBYTES, int ITEMS>
__global__ void __launch_bounds__(kCtaThreads)
    RadixPassKernel(const SortArgs a) {
  constexpr int ND = sizeof(KeyT);
  constexpr int TILE = ITEMS * kCtaThreads;
  using WT = typename std::conditional<WBYTES == 4, uint32_t, uint16_t>::type;

  __shared__ uint32_t s_warp_cnt[kWarpsPerCta][kRadix];  // per-warp digit counts
  __shared__ uint32_t s_tile_excl[kRadix];  // digit start inside the tile
  __shared__ int32_t s_goff[kRadix];        // global start - tile start
  __shared__ uint32_t s_scan[kWarpsPerCta];
  __shared__ int s_plan[4];  // live, ordinal, total live passes, tile id
  extern __shared__ __align__(16) unsigned char s_exch_raw[];
  KeyT* s_exch = reinterpret_cast<KeyT*>(s_exch_raw);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  // ---- plan: which byte positions are live (derived from the histograms)
  if (tid < 32) {
    int live_mask = 0;
    for (int p = 0; p < ND; ++p) {
      bool degenerate = false;
      for (int b = lane; b < kRadix; b += 32)
        degenerate |= (a.hist[p * kRadix + b] == static_cast<uint32_t>(a.nnz));
      if (!__any_sync(0xffffffffu, degenerate)) live_mask |= (1 << p);
    }
    if (lane == 0) {
      s_plan[0] = (live_mask >> a.pass) & 1;
      s_plan[1] = __popc(live_mask & ((1 << a.pass) - 1));
      s_plan[2] = __popc(live_mask);
      // Dynamic tile id: a tile only ever waits for tiles that started
      // earlier, so the look-back cannot deadlock.
      s_plan[3] = static_cast<int>(atomicAdd(&a.tile_counters[a.pass], 1u));
    }
  }
  for (int i = tid; i < kWarpsPerCta * kRadix; i += kCtaThreads)
    (&s_warp_cnt[0][0])[i] = 0;
  __syncthreads();
  const bool live = s_plan[0] != 0;
  const int ordinal = s_plan[1];
  const int total_live = s_plan[2];
  const int tile = s_plan[3];
  const int tile_base = tile * TILE;
  const int tile_n = min(TILE, a.nnz - tile_base);

  // No live pass at all (all keys equal): pass 0 degenerates to a copy.
  if (total_live == 0) {
    if (a.pass != 0) return;
    const KeyT* kin = static_cast<const KeyT*>(a.keys_in);
    const KeyT* vin = static_cast<const KeyT*>(a.vals_in);
    KeyT* kout = static_cast<KeyT*>(a.keys_out);
    KeyT* vout = static_cast<KeyT*>(a.vals_out);
    for (int i = tid; i < tile_n; i += kCtaThreads) {
      kout[tile_base + i] = kin[tile_base + i];
      vout[tile_base + i] = vin[tile_base + i];
      if constexpr (WBYTES != 0)
        static_cast<WT*>(a.w_out)[tile_base + i] =
            static_cast<const WT*>(a.w_in)[tile_base + i];
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

  // ---- load keys: warp-striped so that (warp, item, lane) is input order
  KeyT key[ITEMS];
  uint32_t rank[ITEMS];
  const int warp_base = warp * (32 * ITEMS);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int local = warp_base + i * 32 + lane;
    key[i] = local < tile_n ? kin[tile_base + local] : KeyT(0);
  }

  // ---- stable ranking inside the warp: match-any multisplit
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int local = warp_base + i * 32 + lane;
    const bool valid = local < tile_n;
    // Invalid lanes form their own peer group (digit 256) and are ignored.
    const uint32_t d = valid ? DigitOf<KeyT>(key[i], a.pass) : kRadix;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (valid && lane == leader)
      old = atomicAdd(&s_warp_cnt[warp][d], __popc(peers));
    old = __shfl_sync(0xffffffffu, old, leader);
    rank[i] = old + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();

  // ---- per digit (thread tid owns digit tid): exclusive scan over warps,
  //      tile total, look-back over earlier tiles
  uint32_t tile_count = 0;
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) {
    const uint32_t c = s_warp_cnt[w][tid];
    s_warp_cnt[w][tid] = tile_count;
    tile_count += c;
  }
  uint32_t* lb = a.lookback +
                 (static_cast<size_t>(a.pass) * a.num_tiles) * kRadix;
  if (tile == 0) {
    StVolatile(&lb[tid], kFlagPrefix | tile_count);
  } else {
    StVolatile(&lb[static_cast<size_t>(tile) * kRadix + tid],
               kFlagAgg | tile_count);
  }
  uint32_t exclusive = 0;
  if (tile > 0) {
    int prev = tile - 1;
    while (true) {
      uint32_t s;
      do {
        s = LdVolatile(&lb[static_cast<size_t>(prev) * kRadix + tid]);
      } while ((s & ~kValueMask) == 0);
      exclusive += s & kValueMask;
      if ((s & kFlagPrefix) != 0) break;
      --prev;
    }
    StVolatile(&lb[static_cast<size_t>(tile) * kRadix + tid],
               kFlagPrefix | (exclusive + tile_count));
  }

  // global start of this digit = exclusive scan of the pass histogram
  // (block-wide scan over the 256 digit owners) + earlier tiles.
  const uint32_t hcount = a.hist[a.pass * kRadix + tid];
  uint32_t hscan = hcount, tscan = tile_count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t h = __shfl_up_sync(0xffffffffu, hscan, o);
    const uint32_t t = __shfl_up_sync(0xffffffffu, tscan, o);
    if (lane >= o) {
      hscan += h;
      tscan += t;
    }
  }
  __shared__ uint32_t s_hwarp[kWarpsPerCta];
  if (lane == 31) {
    s_hwarp[warp] = hscan;
    s_scan[warp] = tscan;
  }
  __syncthreads();
  uint32_t hbase = 0, tbase = 0;
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) {
    if (w < warp) {
      hbase += s_hwarp[w];
      tbase += s_scan[w];
    }
  }
  const uint32_t h_excl = hbase + hscan - hcount;
  const uint32_t t_excl = tbase + tscan - tile_count;
  s_tile_excl[tid] = t_excl;
  s_goff[tid] = static_cast<int32_t>(h_excl + exclusive) -
                static_cast<int32_t>(t_excl);
  __syncthreads();

  // ---- final position of every item inside the tile; keys through smem
  uint32_t pos[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int local = warp_base + i * 32 + lane;
    if (local < tile_n) {
      const uint32_t d = DigitOf<KeyT>(key[i], a.pass);
      pos[i] = s_tile_excl[d] + s_warp_cnt[warp][d] + rank[i];
      s_exch[pos[i]] = key[i];
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
      const KeyT k = s_exch[p];
      gaddr[i] = s_goff[DigitOf<KeyT>(k, a.pass)] + p;
      kout[gaddr[i]] = k;
    } else {
      gaddr[i] = -1;
    }
  }
  __syncthreads();

  // ---- payload: sample ids (same type as keys) through the same buffer
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int local = warp_base + i * 32 + lane;
    if (local < tile_n) s_exch[pos[i]] = vin[tile_base + local];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int p = tid + i * kCtaThreads;
    if (p < tile_n) vout[gaddr[i]] = s_exch[p];
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

namespace {

struct SortLayout {
  int items;
  int tile;
  int num_tiles;
  size_t hist_off, counters_off, lookback_off, zero_bytes;
  size_t keys_tmp_off, vals_tmp_off, w_tmp_off, total;
};

SortLayout MakeSortLayout(int nnz, int idx_type, int wbytes) {
  SortLayout L;
  const int nd = static_cast<int>(IndexSize(idx_type));
  static const int items_env = EnvInt("CUEMBED_SORT_ITEMS", 0);
  L.items = items_env > 0 ? items_env : (idx_type == CUEMBED_I64 ? 8 : 16);
  if (L.items != 8 && L.items != 16) L.items = 8;
  L.tile = L.items * kCtaThreads;
  L.num_tiles = nnz > 0 ? (nnz + L.tile - 1) / L.tile : 0;
  size_t off = 0;
  L.hist_off = off;
  off += static_cast<size_t>(nd) * kRadix * sizeof(uint32_t);
  L.counters_off = off;
  off += AlignUp(static_cast<size_t>(nd) * sizeof(uint32_t), 128);
  L.lookback_off = off;
  off += static_cast<size_t>(nd) * L.num_tiles * kRadix * sizeof(uint32_t);
  L.zero_bytes = off;
  off = AlignUp(off, 256);
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
  const size_t smem = static_cast<size_t>(ITEMS) * kCtaThreads * sizeof(KeyT);
  auto kernel = RadixPassKernel<KeyT, WBYTES, ITEMS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>(smem));
    attr_set = true;
  }
  for (int p = 0; p < ND; ++p) {
    SortArgs a = base;
    a.pass = p;
    kernel<<<a.num_tiles, kCtaThreads, smem, stream>>>(a);
  }
  CountLaunch(ND);
}

template <typename KeyT, int WBYTES>
void LaunchPassesItems(const SortArgs& a, int items, cudaStream_t stream) {
  if (items == 16)
    LaunchPasses<KeyT, WBYTES, 16>(a, stream);
  else
    LaunchPasses<KeyT, WBYTES, 8>(a, stream);
}

template <typename KeyT>
void LaunchSortTyped(const SortArgs& a, int wbytes, int items,
                     cudaStream_t stream) {
  {
    const int ctas = CeilDiv(a.nnz, kCtaThreads * 8);
    const int cap = GetDeviceInfo().sm_count * 4;
    RadixHistKernel<KeyT><<<ctas < cap ? ctas : cap, kCtaThreads, 0, stream>>>(
        static_cast<const KeyT*>(a.keys_in), a.nnz,
        const_cast<uint32_t*>(a.hist));
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

int LaunchTranspose(const void* rows, const void* cols, const void* weights,
                    int weight_dtype, int nnz, int idx_type,
                    void* transpose_rows, void* transpose_cols,
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
  if (rows == nullptr || cols == nullptr || transpose_rows == nullptr ||
      transpose_cols == nullptr)
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
  a.hist = reinterpret_cast<const uint32_t*>(work + L.hist_off);
  a.lookback = reinterpret_cast<uint32_t*>(work + L.lookback_off);
  a.tile_counters = reinterpret_cast<uint32_t*>(work + L.counters_off);
  a.nnz = nnz;
  a.num_tiles = L.num_tiles;
  a.pass = 0;
  if (idx_type == CUEMBED_I64)
    LaunchSortTyped<int64_t>(a, wbytes, L.items, stream);
  else
    LaunchSortTyped<int32_t>(a, wbytes, L.items, stream);
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

// ------------------------------------------- compressed gradient indices

// remapped[i] = number of positions j in (0, i] with idx[j] != idx[j-1]:
// single pass, decoupled look-back over tile totals.
// status word: 2 flag bits + 62 value bits.
constexpr unsigned long long kFlagAgg64 = 1ull << 62;
constexpr unsigned long long kFlagPrefix64 = 2ull << 62;
constexpr unsigned long long kValueMask64 = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long LdVolatile64(
    const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void StVolatile64(unsigned long long* p,
                                             unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}

template <typename IdxT, int ITEMS>
__global__ void __launch_bounds__(kCtaThreads)
    CompressIndicesKernel(const IdxT* __restrict__ idx, int nnz,
                          IdxT* __restrict__ remapped,
                          unsigned long long* __restrict__ status,
                          uint32_t* __restrict__ tile_counter) {
  constexpr int TILE = ITEMS * kCtaThreads;
  __shared__ int s_tile;
  __shared__ uint32_t s_warp[kWarpsPerCta];
  __shared__ unsigned long long s_excl;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  if (tid == 0) s_tile = static_cast<int>(atomicAdd(tile_counter, 1u));
  __syncthreads();
  const int tile = s_tile;
  const int64_t base = static_cast<int64_t>(tile) * TILE + tid * ITEMS;

  // Blocked layout: thread owns ITEMS consecutive elements.
  uint32_t flag[ITEMS];
  uint32_t local = 0;
  IdxT prev = (base > 0 && base - 1 < nnz) ? __ldg(idx + base - 1) : IdxT(0);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int64_t g = base + i;
    IdxT cur = g < nnz ? __ldg(idx + g) : prev;
    flag[i] = (g > 0 && g < nnz && cur != prev) ? 1u : 0u;
    local += flag[i];
    prev = cur;
  }
  // block exclusive scan of `local`
  uint32_t scan = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, scan, o);
    if (lane >= o) scan += t;
  }
  if (lane == 31) s_warp[warp] = scan;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kWarpsPerCta; ++w) {
    if (w < warp) wbase += s_warp[w];
    total += s_warp[w];
  }
  const uint32_t thread_excl = wbase + scan - local;

  // Warp 0 resolves the totals of all earlier tiles, 32 predecessors per step.
  if (warp == 0) {
    unsigned long long exclusive = 0;
    if (tile == 0) {
      if (lane == 0) StVolatile64(&status[0], kFlagPrefix64 | total);
    } else {
      if (lane == 0) StVolatile64(&status[tile], kFlagAgg64 | total);
      int p = tile - 1;
      while (true) {
        const int q = p - lane;
        unsigned long long s = kFlagPrefix64;  // before tile 0: prefix 0
        if (q >= 0) {
          do {
            s = LdVolatile64(&status[q]);
          } while ((s & ~kValueMask64) == 0);
        }
        const unsigned pm =
            __ballot_sync(0xffffffffu, (s & kFlagPrefix64) != 0);
        const int first = __ffs(pm) - 1;  // nearest predecessor with a prefix
        unsigned long long v =
            (pm == 0 || lane <= first) ? (s & kValueMask64) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
          v += __shfl_xor_sync(0xffffffffu, v, o);
        exclusive += v;
        if (pm != 0) break;
        p -= 32;
      }
      if (lane == 0)
        StVolatile64(&status[tile], kFlagPrefix64 | (exclusive + total));
    }
    if (lane == 0) s_excl = exclusive;
  }
  __syncthreads();
  unsigned long long running = s_excl + thread_excl;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int64_t g = base + i;
    running += flag[i];
    if (g < nnz) remapped[g] = static_cast<IdxT>(running);
  }
}

int LaunchCompressedGradIndices(const void* indices, int idx_type, int nnz,
                                void* remapped, char* work, size_t* lwork,
                                cudaStream_t stream) {
  if (lwork == nullptr || nnz < 0) return CUEMBED_ERR_ARGUMENT;
  if (idx_type < 0 || idx_type > 1) return CUEMBED_ERR_DTYPE;
  constexpr int ITEMS = 8;
  constexpr int TILE = ITEMS * kCtaThreads;
  const int num_tiles = nnz > 0 ? (nnz + TILE - 1) / TILE : 0;
  const size_t need = 128 + static_cast<size_t>(num_tiles) * 8;
  if (work == nullptr) {
    *lwork = need;
    return CUEMBED_OK;
  }
  if (*lwork < need) return CUEMBED_ERR_WORKSPACE;
  if (nnz == 0) return CUEMBED_OK;
  if (indices == nullptr || remapped == nullptr) return CUEMBED_ERR_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(work) & 7) != 0) return CUEMBED_ERR_ARGUMENT;
  if (cudaMemsetAsync(work, 0, need, stream) != cudaSuccess)
    return CUEMBED_ERR_CUDA;
  uint32_t* counter = reinterpret_cast<uint32_t*>(work);
  unsigned long long* status =
      reinterpret_cast<unsigned long long*>(work + 128);
  if (idx_type == CUEMBED_I64)
    CompressIndicesKernel<int64_t, ITEMS>
        <<<num_tiles, kCtaThreads, 0, stream>>>(
            static_cast<const int64_t*>(indices), nnz,
            static_cast<int64_t*>(remapped), status, counter);
  else
    CompressIndicesKernel<int32_t, ITEMS>
        <<<num_tiles, kCtaThreads, 0, stream>>>(
            static_cast<const int32_t*>(indices), nnz,
            static_cast<int32_t*>(remapped), status, counter);
  CountLaunch();
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

}  // namespace cuembed_b200
