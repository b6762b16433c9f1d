// backward_hot.cuh -- the hot-row path of the backward (included by
// backward.cu after BwdArgs).
//
// Why: under a power law a few hundred table rows receive half of all lookups
// (C2: 204 rows with >= 2048 lookups hold 46 % of the nonzeros).  The chunk
// walker reads grad_y[sample] for every nonzero through the L2 -> SM fabric
// (L1 hit rate 2 %: inside a run every sample occurs once), and that fabric is
// what bounds it.  Reuse only exists ACROSS hot rows: they all read the same
// grad_y rows.  So the interiors of hot runs are taken out of the chunk walker
// and processed sample-tile by sample-tile:
//
//   BwdHotScanAKernel   one warp per chunk: is the chunk strictly inside one
//                       run ("through") and are its sample ids ascending?
//   BwdHotScanBKernel   maximal sequences of >= hot_min_chunks such chunks
//                       whose samples are dense enough become HOT UNITS; their
//                       chunks are marked (the chunk walker skips them and
//                       reports them as "through" chunks).
//   BwdHotKernel        CTA = (group of 64 hot units) x (range of samples).
//                       The CTA streams the grad_y rows of its sample range
//                       through shared memory ONCE (cp.async.bulk + mbarrier,
//                       3 stages of 64 KB) and every warp adds, for its 4 hot
//                       units, the rows of the tile that the unit references
//                       (sample ids are ascending inside a run, so each unit is
//                       a cursor that only moves forward).  fp32 accumulators
//                       stay in registers over the whole range; one partial row
//                       per (unit, range) is written at the end.
//   BwdHotCombineKernel adds the partials of a unit in range order and puts the
//                       sum where the fix-up expects the head partial of the
//                       unit's first chunk (zeros for its other chunks), so the
//                       existing two-level fix-up finishes the run unchanged.
//
// grad_y traffic of the hot nonzeros drops from one row per nonzero to one row
// per (sample, group of 64 units).  Every sum keeps a fixed association order
// (units are registered in arbitrary order, but a unit's arithmetic does not
// depend on its slot), so results stay bit-identical from run to run.
#ifndef CUEMBED_B200_CSRC_BACKWARD_HOT_CUH_
#define CUEMBED_B200_CSRC_BACKWARD_HOT_CUH_

namespace cuembed_b200 {

constexpr int kHotThreads = 512;
constexpr int kHotWarps = kHotThreads / 32;
constexpr int kHotStages = 3;
constexpr int kHotStageBytes = 64 * 1024;
constexpr int kHotSmemBytes = kHotStages * kHotStageBytes + 64;
constexpr int kHotMaxRanges = 64;
constexpr int kHotMinRanges = 4;
constexpr int kHotPieceBytes = 16 * 1024;  // one bulk copy
// a unit is hot only if its nonzeros cover >= 1 / kHotSparsity of the samples
// they span (otherwise streaming the tile costs more than the gathers)
constexpr int kHotSparsity = 32;

constexpr unsigned char kChunkPlain = 0;
constexpr unsigned char kChunkThrough = 1;  // strictly inside a run, ascending
constexpr unsigned char kChunkHot = 2;

__host__ __device__ inline int HotRanges(int n_groups, int sm_slots) {
  int r = n_groups > 0 ? sm_slots / n_groups : kHotMaxRanges;
  if (r > kHotMaxRanges) r = kHotMaxRanges;
  if (r < kHotMinRanges) r = kHotMinRanges;
  return r;
}

// ---------------------------------------------------------------- PTX helpers

__device__ __forceinline__ uint32_t SmemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void MbarInit(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   SmemAddr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(SmemAddr(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA, no tensor map), completion on an mbarrier
__device__ __forceinline__ void BulkLoad(void* smem_dst, const void* gsrc,
                                         uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(SmemAddr(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(SmemAddr(bar))
      : "memory");
}

// ------------------------------------------------------------------- scan A

// One warp per chunk.  state[chunk] = kChunkThrough iff the chunk lies strictly
// inside one run (the element before it and the element after it carry its
// key; equal keys are contiguous) and sample ids do not decrease from the
// element before the chunk to its last element.
template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    BwdHotScanAKernel(const BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.hot_ctr[0] = 0;  // number of hot units
    a.hot_ctr[1] = 0;  // largest sample id of a hot unit
  }
  if (chunk >= a.num_chunks) return;
  const IdxT* __restrict__ keys = static_cast<const IdxT*>(a.keys);
  const IdxT* __restrict__ sids = static_cast<const IdxT*>(a.sids);
  const int K = a.chunk_nz;
  const int64_t c0 = static_cast<int64_t>(chunk) * K;
  bool through = c0 > 0 && c0 + K < a.nnz;
  if (through) through = __ldg(keys + c0 - 1) == __ldg(keys + c0 + K);
  if (through) {
    int bad = 0;
#pragma unroll 4
    for (int j = lane; j < K; j += 32)
      bad |= static_cast<int>(__ldg(sids + c0 + j - 1) > __ldg(sids + c0 + j));
    through = __all_sync(0xffffffffu, bad == 0);
  }
  if (lane == 0) a.chunk_state[chunk] = through ? kChunkThrough : kChunkPlain;
}

// ------------------------------------------------------------------- scan B

// One warp per chunk; only the warp of the FIRST chunk of a maximal sequence
// of "through" chunks works: it measures the sequence, decides whether it is a
// hot unit, registers it and marks its chunks.
template <typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    BwdHotScanBKernel(const BwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (chunk >= a.num_chunks) return;
  // Written by scan A, upgraded to kChunkHot by other warps of this kernel:
  // volatile reads, and "through or hot" both mean "was through".
  const volatile unsigned char* state = a.chunk_state;
  if (state[chunk] == kChunkPlain) return;
  if (chunk > 0 && state[chunk - 1] != kChunkPlain) return;
  int n = 0;
  for (int base = chunk;; base += 32) {
    const int c = base + lane;
    const bool t = c < a.num_chunks && state[c] != kChunkPlain;
    const unsigned m = __ballot_sync(0xffffffffu, t);
    if (m != 0xffffffffu) {
      n += __ffs(~m) - 1;
      break;
    }
    n += 32;
  }
  if (n < a.hot_min_chunks) return;
  const IdxT* __restrict__ sids = static_cast<const IdxT*>(a.sids);
  const int K = a.chunk_nz;
  const int64_t p0 = static_cast<int64_t>(chunk) * K;
  const int64_t p1 = p0 + static_cast<int64_t>(n) * K;
  const long long s_first = static_cast<long long>(__ldg(sids + p0));
  const long long s_last = static_cast<long long>(__ldg(sids + p1 - 1));
  if (s_first < 0 || s_last >= 0x7fffffffLL) return;
  if (s_last - s_first + 1 > static_cast<long long>(kHotSparsity) * (p1 - p0))
    return;
  int slot = 0;
  if (lane == 0) {
    slot = atomicAdd(a.hot_ctr + 0, 1);
    atomicMax(a.hot_ctr + 1, static_cast<int>(s_last));
  }
  slot = __shfl_sync(0xffffffffu, slot, 0);
  if (slot >= a.hot_cap) return;  // cannot happen: cap = num_chunks / min + 1
  if (lane == 0) a.hot_units[slot] = make_int2(chunk, n);
  for (int i = lane; i < n; i += 32) a.chunk_state[chunk + i] = kChunkHot;
}

// ----------------------------------------------------------------- hot kernel

template <typename T, typename IdxT, bool WEIGHTED, int NV, int RPW>
__global__ void __launch_bounds__(kHotThreads, 1)
    BwdHotKernel(const BwdArgs a) {
  constexpr int NE = 4 * Elem<T>::kPerWord;  // elements per 16-byte vector
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int kRpc = kHotWarps * RPW;  // hot units per CTA
  extern __shared__ __align__(128) unsigned char hot_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(hot_smem +
                                               kHotStages * kHotStageBytes);

  const int n_hot = min(a.hot_ctr[0], a.hot_cap);
  const int n_groups = (n_hot + kRpc - 1) / kRpc;
  const int g = blockIdx.y;
  const int c = blockIdx.x;
  if (g >= n_groups) return;
  const int R = HotRanges(n_groups, a.sm_slots);
  if (c >= R) return;
  const int n_samples = a.hot_ctr[1] + 1;
  const int SR = (n_samples + R - 1) / R;
  const int r_lo = min(c * SR, n_samples);
  const int r_hi = min(r_lo + SR, n_samples);
  const int row_bytes = static_cast<int>(a.row_bytes);
  const int TS = kHotStageBytes / row_bytes;
  const int n_tiles = (r_hi - r_lo + TS - 1) / TS;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int G = a.lanes;              // lanes per row (32 when NV > 1)
  const int lane_g = lane & (G - 1);  // vector of the row
  const int q = lane >> a.log2_lanes;  // which of the 32 / G rows of a step
  const int spw = 32 >> a.log2_lanes;
  const IdxT* __restrict__ sids = static_cast<const IdxT*>(a.sids);
  const T* __restrict__ weights = static_cast<const T*>(a.weights);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kHotStages; ++s) MbarInit(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int t) {
    const int s = t % kHotStages;
    const int t_lo = r_lo + t * TS;
    const uint32_t bytes =
        static_cast<uint32_t>(min(TS, r_hi - t_lo)) * row_bytes;
    MbarExpectTx(full + s, bytes);
    const char* src =
        static_cast<const char*>(a.grad_y) + static_cast<int64_t>(t_lo) * row_bytes;
    unsigned char* dst = hot_smem + s * kHotStageBytes;
    for (uint32_t off = 0; off < bytes; off += kHotPieceBytes)
      BulkLoad(dst + off, src + off, min(bytes - off, (uint32_t)kHotPieceBytes),
               full + s);
  };
  if (tid == 0) {
    for (int t = 0; t < kHotStages - 1 && t < n_tiles; ++t) issue(t);
  }

  // ---- the units of this warp: cursor = first nonzero with sample >= r_lo
  int cur[RPW], end[RPW], slot[RPW];
  {
    int lo[RPW], hi[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      slot[k] = g * kRpc + warp * RPW + k;
      lo[k] = hi[k] = 0;
      if (slot[k] < n_hot) {
        const int2 u = a.hot_units[slot[k]];
        lo[k] = u.x * a.chunk_nz;
        hi[k] = lo[k] + u.y * a.chunk_nz;
      }
      end[k] = hi[k];
    }
    // 32-ary lower bound, all units of the warp in lock step
    bool more = true;
    while (more) {
      more = false;
      int v[RPW], step[RPW];
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        step[k] = (hi[k] - lo[k] + 31) >> 5;
        v[k] = 0;
        if (hi[k] - lo[k] > 32) {
          const int p = min(lo[k] + (lane + 1) * step[k] - 1, hi[k] - 1);
          v[k] = static_cast<int>(__ldg(sids + p));
        }
      }
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        if (hi[k] - lo[k] > 32) {  // warp-uniform
          const unsigned ge = __ballot_sync(kFull, v[k] >= r_lo);
          if (ge == 0u) {
            lo[k] = hi[k];
          } else {
            const int f = __ffs(ge) - 1;
            const int pf = min(lo[k] + (f + 1) * step[k] - 1, hi[k] - 1);
            if (f > 0) lo[k] = min(lo[k] + f * step[k] - 1, hi[k] - 1) + 1;
            hi[k] = pf;
          }
          more = more || (hi[k] - lo[k] > 32);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      const int p = lo[k] + lane;
      int v = 0x7fffffff;
      if (p < hi[k]) v = static_cast<int>(__ldg(sids + p));
      const unsigned ge = __ballot_sync(kFull, v >= r_lo && p < hi[k]);
      cur[k] = ge != 0u ? lo[k] + __ffs(ge) - 1 : hi[k];
    }
  }

  // Three windows of 32 sample ids per unit live in registers: w0 is being
  // consumed (its first `wo` entries are done), w1 and w2 were requested 32 and
  // 64 nonzeros ahead, so the index loads never sit in front of the row adds.
  int w0[RPW], w1[RPW], w2[RPW], wo[RPW];
  float f0[RPW], f1[RPW], f2[RPW];  // weights, widened on load (exact)
  float acc[RPW][NV][NE];
  auto load_window = [&](int k, int pos, int& w, float& f) {
    const int p = pos + lane;
    w = 0x7fffffff;
    f = 0.f;
    if (p < end[k]) {
      w = static_cast<int>(__ldg(sids + p));
      if constexpr (WEIGHTED) f = Elem<T>::ToFloat(__ldg(weights + p));
    }
  };
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
    load_window(k, cur[k], w0[k], f0[k]);
    load_window(k, cur[k] + 32, w1[k], f1[k]);
    load_window(k, cur[k] + 64, w2[k], f2[k]);
    wo[k] = 0;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < NE; ++e) acc[k][v][e] = 0.f;
  }

  for (int t = 0; t < n_tiles; ++t) {
    __syncthreads();  // every warp is done with tile t-1: its stage is free
    if (tid == 0 && t + kHotStages - 1 < n_tiles) issue(t + kHotStages - 1);
    MbarWait(full + (t % kHotStages), (t / kHotStages) & 1);
    const int t_lo = r_lo + t * TS;
    const int t_hi = min(t_lo + TS, r_hi);
    const unsigned char* tile =
        hot_smem + (t % kHotStages) * kHotStageBytes + lane_g * 16;
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      while (true) {  // all conditions are warp-uniform
        // entries wo .. wo + n - 1 of w0 belong to this tile (ids ascend)
        const unsigned m =
            __ballot_sync(kFull, w0[k] < t_hi) & (0xffffffffu << wo[k]);
        const int n = __popc(m);
        constexpr int UN = NV >= 2 ? 2 : 4;
        for (int jb = 0; jb < n; jb += spw * UN) {
          uint4 vals[UN][NV];
          float wf[UN];
          bool ok[UN];
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const int j = jb + u * spw + q;
            ok[u] = j < n;
            const int src = (wo[k] + j) & 31;
            const int s = __shfl_sync(kFull, w0[k], src);
            wf[u] = 1.f;
            if constexpr (WEIGHTED) wf[u] = __shfl_sync(kFull, f0[k], src);
            const unsigned char* row =
                tile + static_cast<int64_t>(ok[u] ? s - t_lo : 0) * row_bytes;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              if (lane_g + 32 * v < a.nvec)
                vals[u][v] = *reinterpret_cast<const uint4*>(row + v * 512);
              else
                vals[u][v] = make_uint4(0, 0, 0, 0);
            }
          }
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            if (ok[u]) {
#pragma unroll
              for (int v = 0; v < NV; ++v) {
                if constexpr (WEIGHTED)
                  AccumulateVecWeighted<T, 16>(vals[u][v], wf[u], acc[k][v]);
                else
                  AccumulateVec<T, 16>(vals[u][v], acc[k][v]);
              }
            }
          }
        }
        wo[k] += n;
        if (wo[k] < 32) break;  // the rest of w0 belongs to later tiles
        cur[k] += 32;
        w0[k] = w1[k];
        f0[k] = f1[k];
        w1[k] = w2[k];
        f1[k] = f2[k];
        load_window(k, cur[k] + 64, w2[k], f2[k]);
        wo[k] = 0;
      }
    }
  }

  // ---- one partial row per (unit, range); lane groups q > 0 hold the sums of
  // every spw-th row of a step: added in q order.
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
    for (int d = 1; d < spw; ++d) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int e = 0; e < NE; ++e) {
          const float o =
              __shfl_sync(kFull, acc[k][v][e], (lane_g + d * G) & 31);
          if (q == 0) acc[k][v][e] = __fadd_rn(acc[k][v][e], o);
        }
    }
    if (slot[k] < n_hot && q == 0) {
      float* dst = a.hot_partial +
                   (static_cast<size_t>(slot[k]) * kHotMaxRanges + c) * a.width;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int vec = lane_g + 32 * v;
        if (vec < a.nvec) StorePartial<NE>(dst + vec * NE, acc[k][v]);
      }
    }
  }
}

// -------------------------------------------------------------------- combine

__global__ void __launch_bounds__(kCtaThreads)
    BwdHotCombineKernel(const BwdArgs a, int rpc) {
  const int n_hot = min(a.hot_ctr[0], a.hot_cap);
  const int slot = blockIdx.x;
  if (slot >= n_hot) return;
  const int col = blockIdx.y * kCtaThreads + threadIdx.x;
  if (col >= a.width) return;
  const int R = HotRanges((n_hot + rpc - 1) / rpc, a.sm_slots);
  const int2 u = a.hot_units[slot];
  const float* __restrict__ p =
      a.hot_partial + static_cast<size_t>(slot) * kHotMaxRanges * a.width + col;
  float acc = p[0];
  int c = 1;
  for (; c + 8 <= R; c += 8) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = p[static_cast<size_t>(c + i) * a.width];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = __fadd_rn(acc, v[i]);
  }
  for (; c < R; ++c) acc = __fadd_rn(acc, p[static_cast<size_t>(c) * a.width]);
  const size_t pitch = static_cast<size_t>(2) * a.width;
  float* head = a.scratch + static_cast<size_t>(u.x) * pitch + col;
  head[0] = acc;
  for (int i = 1; i < u.y; ++i) head[static_cast<size_t>(i) * pitch] = 0.f;
}

}  // namespace cuembed_b200

#endif  // CUEMBED_B200_CSRC_BACKWARD_HOT_CUH_
