// backward_warp.cuh -- the chunk walker of the backward for rows that one warp
// covers exactly (row bytes 128, 256 or a multiple of 512: 32 lanes x 4 / 8 /
// 16 bytes, wider rows in column tiles).  Included by backward.cu.
//
// Same contract as BwdSegReduceKernel (chunk = K consecutive nonzeros of the
// sorted COO, walked in order with fp32 accumulators; runs inside the chunk are
// rounded once and stored; a run that crosses chunk edges leaves a head and / or
// tail partial for the fix-up), but with ONE lane group per warp every
// condition is warp-uniform, which removes most of the instructions ncu counted
// per nonzero in round 1 (~30 per nonzero on average, ~40 on cold rows):
//
//   * the row address is ONE IMAD.WIDE (sample id x row pitch + a base pointer
//     that already contains the lane's column offset and is kept opaque to the
//     compiler, which otherwise re-adds the kernel-parameter base per load);
//   * the per-lane index words of a round rotate by one batch per iteration,
//     so every shuffle has an immediate source lane although the batch loop
//     is not unrolled (a fully unrolled round was 64 KB of code and stalled on
//     instruction fetch: ncu "no_instruction" 4.0 warps per issue);
//   * run ends come from a shuffle of the keys (no second key load per lane);
//   * inverse_mapping is written by the lanes that own a run end, once per
//     round of 32 nonzeros, instead of two shuffles + a store per run end;
//   * a run of length one (74 % of the runs at the headline workload) is a row
//     COPY: the loaded vector is stored after a 16-bit "+ 0" (which gives the
//     -0 -> +0 of the oracle's `0 + x`), skipping 8 mixed-precision adds, 4
//     packs and the accumulator reset;
//   * a batch of 8 nonzeros without a run end is 8 x (shuffle, address, load)
//     + 64 adds and nothing else.
#ifndef CUEMBED_B200_CSRC_BACKWARD_WARP_CUH_
#define CUEMBED_B200_CSRC_BACKWARD_WARP_CUH_

namespace cuembed_b200 {

#ifndef BWD_WARP_MINB
#define BWD_WARP_MINB 7
#endif
#ifndef BWD_WARP_UNROLL
#define BWD_WARP_UNROLL 8
#endif
// Warps (= chunks) per CTA of the warp walker: ONE.  A CTA holds its slot on the
// SM until its slowest warp is done, and chunks differ in cost (interiors of hot
// runs take the branch-free path, cold chunks end a run at every nonzero); with
// four chunks per CTA ncu showed 30.6 % achieved occupancy against 37.5 %
// theoretical.  (The SM holds at most 32 CTAs, i.e. 32 such warps.)
#ifndef BWD_WARP_CTA_WARPS
#define BWD_WARP_CTA_WARPS 1
#endif
constexpr int kBwdWarpThreads = 32 * BWD_WARP_CTA_WARPS;
// Resident warps per SM the register allocation aims at: 4 * BWD_WARP_MINB (28
// -> at most 72 registers) for the lean instantiations (32-bit indices,
// unweighted, plain gradient), 4 fewer for the others, which would spill at 72.
// Measured at C2 (profiles/r02_notes.md): one-warp CTAs at 28 warps per SM
// 0.1938 ms, at 24 warps 0.2024, four-warp CTAs at 24 warps 0.2109; 32 warps
// (64 registers, spills) 0.2005.
constexpr int BwdWarpMinBlocks(bool lean) {
  return (lean ? BWD_WARP_MINB : BWD_WARP_MINB - 1) * 4 / BWD_WARP_CTA_WARPS;
}
// round_T(0.f + float(x)) for every 16-bit element of a vector: x itself except
// that -0 becomes +0 (what a sum that starts at +0 gives).
template <typename T, int V>
__device__ __forceinline__ typename VecBits<V>::type ZeroPlus(
    typename VecBits<V>::type v) {
  static_assert(sizeof(T) == 2, "16-bit element types only");
  constexpr int NW = V / 4;
  uint32_t w[NW];
  Unpack32(v, w);
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    if constexpr (Elem<T>::kCode == CUEMBED_F16)
      asm("add.rn.f16x2 %0, %1, %2;" : "=r"(w[i]) : "r"(w[i]), "r"(0u));
    else
      asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(w[i]) : "r"(w[i]), "r"(0u));
  }
  typename VecBits<V>::type out;
  Pack32(w, &out);
  return out;
}

// Optimisation barrier on one register value (no instruction is emitted).
__device__ __forceinline__ void PinReg(int32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void PinReg(int64_t& v) { asm volatile("" : "+l"(v)); }

template <typename IdxT>
__device__ __forceinline__ IdxT ShflDownIdx(IdxT v, int delta) {
  if constexpr (sizeof(IdxT) == 8)
    return static_cast<IdxT>(
        __shfl_down_sync(0xffffffffu, static_cast<long long>(v), delta));
  else
    return __shfl_down_sync(0xffffffffu, v, delta);
}

template <typename T>
__device__ __forceinline__ float ShflWeight(T w, int src) {
  if constexpr (sizeof(T) == 4) {
    return __shfl_sync(0xffffffffu, w, src);
  } else {
    const unsigned short b = *reinterpret_cast<const unsigned short*>(&w);
    const unsigned rb = __shfl_sync(0xffffffffu, static_cast<unsigned>(b), src);
    const unsigned short rs = static_cast<unsigned short>(rb);
    return Elem<T>::ToFloat(*reinterpret_cast<const T*>(&rs));
  }
}

template <typename T>
__device__ __forceinline__ T ShflRaw(T w, int src) {
  if constexpr (sizeof(T) == 4) {
    return __shfl_sync(0xffffffffu, w, src);
  } else {
    const unsigned short b = *reinterpret_cast<const unsigned short*>(&w);
    const unsigned short rs = static_cast<unsigned short>(
        __shfl_sync(0xffffffffu, static_cast<unsigned>(b), src));
    return *reinterpret_cast<const T*>(&rs);
  }
}

// FUSED_OPT: CUEMBED_OPT_NONE or CUEMBED_OPT_SGD (Adagrad stays on the generic
// kernel: it needs the old table rows in flight).
template <typename T, int V, typename IdxT, bool WEIGHTED, int FUSED_OPT>
__global__ void __launch_bounds__(
    kBwdWarpThreads,
    BwdWarpMinBlocks(sizeof(IdxT) == 4 && !WEIGHTED && FUSED_OPT == CUEMBED_OPT_NONE))
    BwdWarpKernel(const BwdArgs a) {
  using VecT = typename VecBits<V>::type;
  constexpr int NW = V / 4;
  constexpr int NE = NW * Elem<T>::kPerWord;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int UNROLL = BWD_WARP_UNROLL;
  constexpr bool kCopySingles =
      !WEIGHTED && FUSED_OPT == CUEMBED_OPT_NONE && sizeof(T) == 2;

  const int lane = threadIdx.x & 31;
  const int chunk = blockIdx.x * BWD_WARP_CTA_WARPS + (threadIdx.x >> 5);
  const IdxT* __restrict__ keys = static_cast<const IdxT*>(a.keys);
  const IdxT* __restrict__ sids = static_cast<const IdxT*>(a.sids);
  const T* __restrict__ weights = static_cast<const T*>(a.weights);
  const IdxT* __restrict__ tidx = static_cast<const IdxT*>(a.tidx);
  uint32_t row_bytes = static_cast<uint32_t>(a.row_bytes);
  asm volatile("" : "+r"(row_bytes));  // a plain 32-bit register: RowAddr is one IMAD.WIDE

  const int K = a.chunk_nz;
  const int64_t c0 = static_cast<int64_t>(chunk) * K;
  int n = 0;
  if (c0 < a.nnz) n = static_cast<int>(min(static_cast<int64_t>(K), a.nnz - c0));

  const int v = blockIdx.y * 32 + lane;
  const bool active = v < a.nvec;
  // base pointers that already hold this lane's column offset; opaque so that
  // a row address is one IMAD.WIDE (index x pitch + base)
  const char* gy = static_cast<const char*>(a.grad_y) +
                   static_cast<int64_t>(active ? v : a.nvec - 1) * V;
  char* out = static_cast<char*>(a.grad) + static_cast<int64_t>(active ? v : 0) * V;
  asm volatile("" : "+l"(gy), "+l"(out));
  float* __restrict__ my_head =
      a.scratch + (static_cast<size_t>(chunk) * 2 + 0) * a.width;
  float* __restrict__ my_tail =
      a.scratch + (static_cast<size_t>(chunk) * 2 + 1) * a.width;

  // Does the first run continue a run of the previous chunk; does the last
  // element end its run?  (uniform loads)
  bool cont = false;
  bool last_is_end = true;
  if (n > 0) {
    if (c0 > 0) cont = __ldg(keys + c0 - 1) == __ldg(keys + c0);
    if (c0 + n < a.nnz)
      last_is_end = __ldg(keys + c0 + n) != __ldg(keys + c0 + n - 1);
  }
  // Run ownership across chunk edges (a.own): a run that STARTS inside chunk c
  // and ends within the first kOwnWindow - 1 elements of chunk c + 1 is finished
  // by chunk c (it reads those few elements itself) and skipped by chunk c + 1.
  // Both sides evaluate the same rule from the keys alone, so no flag travels
  // between them; most two-element chains of the fix-up (~12 000 at the
  // headline workload) disappear.
  constexpr int kOwnWindow = 32;
  bool skip_head = false;
  if (a.own && cont && n >= kOwnWindow) {
    const bool short_head = __ldg(keys + c0 + kOwnWindow - 1) != __ldg(keys + c0);
    const bool prev_inside =
        c0 - K - 1 < 0 || __ldg(keys + c0 - K - 1) != __ldg(keys + c0 - 1);
    skip_head = short_head && prev_inside;
  }
  bool in_first = true;
  bool open = false;
  int head_kind = kHeadNone;
  IdxT head_row = 0;

  float acc[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) acc[e] = 0.f;

  // (key, sample, weight) of a round are requested one round ahead; lane 31
  // also requests the first key of the round after (its "next key").  Nothing
  // may touch the words of round r + 1 before round r is over: ncu's source
  // page showed 19 % of all warp samples on a shuffle of the just-requested
  // key_n (the next key of lane 31) and on a register move of sid_n that the
  // compiler had sunk into the batch loop -- one full memory latency per round
  // with no row load in flight.
  IdxT key_n = 0, sid_n = 0, kx_n = 0;
  T w_n = T();
  auto request = [&](int r) {
    key_n = 0;
    sid_n = 0;
    kx_n = 0;
    w_n = T();
    const int p = r * 32 + lane;
    if (p < n) {
      key_n = __ldg(keys + c0 + p);
      sid_n = __ldg(sids + c0 + p);
      if constexpr (WEIGHTED) w_n = __ldg(weights + c0 + p);
      if (lane == 31 && p + 1 < n) kx_n = __ldg(keys + c0 + p + 1);
    }
  };
  request(0);

  auto flush = [&](IdxT krow) {
    if (in_first && cont) {
      // Run began in an earlier chunk: partial, added by the fix-up.
      if (active) StorePartial<NE>(my_head + v * NE, acc);
      head_kind = kHeadEnds;
      head_row = krow;
    } else if (active) {
      char* dst = RowAddr<IdxT>(out, krow, row_bytes);
      if constexpr (FUSED_OPT == CUEMBED_OPT_NONE)
        StoreFloatsAs<NE>(dst, 0, Elem<T>::kCode, acc);
      else
        ApplyUpdateVec<T, V, NE>(a, dst - static_cast<int64_t>(v) * V,
                                 static_cast<int64_t>(krow),
                                 static_cast<int64_t>(v) * NE, acc, VecT());
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) acc[e] = 0.f;
    in_first = false;
    open = false;
  };

#pragma unroll 1
  for (int r = 0; r * 32 < n; ++r) {
    const int cnt = min(32, n - r * 32);
    const int p = r * 32 + lane;
    IdxT key = key_n, sid = sid_n, kx = kx_n;
    const T w = w_n;
    // the hand-over happens HERE (the words were requested a round ago), not
    // wherever the compiler would sink it to
    PinReg(key);
    PinReg(sid);
    PinReg(kx);
    request(r + 1);
    // run ends of this round: the next key is one lane to the right, the first
    // key of the next round for lane 31, `last_is_end` for the chunk's last
    IdxT knext = ShflDownIdx<IdxT>(key, 1);
    if (lane == 31) knext = kx;
    const bool end = p < n && (p == n - 1 ? last_is_end : knext != key);
    const unsigned endsw = __ballot_sync(kFull, end);
    // inverse_mapping: the lane that owns a run end writes it (load issued now,
    // store after the row loop so that nothing waits for it)
    const bool write_inv =
        end && a.inverse_mapping != nullptr && blockIdx.y == 0;
    IdxT tk = 0;
    if (write_inv) tk = __ldg(tidx + c0 + p);

    // The batch loop is NOT unrolled (a fully unrolled round was 64 KB of code
    // and stalled on instruction fetch); instead the per-lane words rotate by
    // one batch per iteration, so the shuffles still use immediate lanes 0..7.
    IdxT sid_r = sid, key_r = key;
    T w_r = w;
    unsigned ends = endsw;
    int rem0 = cnt;
    if (skip_head) {
      // the previous chunk finishes our head run: start behind it
      const IdxT k0 = ShflIdx<IdxT>(key, 0, 32);
      const unsigned same = __ballot_sync(kFull, key == k0 && p < n);
      const int e = __ffs(~same) - 1;  // 1 .. kOwnWindow - 1
      const int from = (lane + e) & 31;
      sid_r = ShflIdx<IdxT>(sid_r, from, 32);
      key_r = ShflIdx<IdxT>(key_r, from, 32);
      if constexpr (WEIGHTED) w_r = ShflRaw<T>(w_r, from);
      ends >>= e;
      rem0 -= e;
      skip_head = false;
      cont = false;
    }
#pragma unroll 1
    for (int rem = rem0; rem > 0; rem -= UNROLL) {
      // UNROLL independent row loads issued before the first add.  Lanes past
      // the end of the chunk carry sample id 0: a valid, harmless load that is
      // never accumulated.
      VecT vals[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        vals[u] = LdgVec<V>(
            RowAddr<IdxT>(gy, ShflIdx<IdxT>(sid_r, u, 32), row_bytes));
      const int from = (lane + UNROLL) & 31;
      float wf[UNROLL];
      if constexpr (WEIGHTED) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) wf[u] = ShflWeight<T>(w_r, u);
      }
      const unsigned m8 = ends & ((1u << UNROLL) - 1u);
      if (m8 == 0u && rem >= UNROLL) {
        // no run ends in this batch (the interior of a long run)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if constexpr (WEIGHTED)
            AccumulateVecWeighted<T, V>(vals[u], wf[u], acc);
          else
            AccumulateVec<T, V>(vals[u], acc);
        }
        open = true;
      } else {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          if (u < rem) {  // warp-uniform
            const bool is_end = ((m8 >> u) & 1u) != 0u;
            bool copied = false;
            if constexpr (kCopySingles) {
              if (is_end && !open && !(in_first && cont)) {
                // a run of length one: round(0 + x) == x (-0 -> +0): row copy
                const IdxT krow = ShflIdx<IdxT>(key_r, u, 32);
                if (active)
                  StcsVec<V>(RowAddr<IdxT>(out, krow, row_bytes),
                             ZeroPlus<T, V>(vals[u]));
                in_first = false;
                copied = true;
              }
            }
            if (!copied) {
              if constexpr (WEIGHTED)
                AccumulateVecWeighted<T, V>(vals[u], wf[u], acc);
              else
                AccumulateVec<T, V>(vals[u], acc);
              open = true;
              if (is_end) flush(ShflIdx<IdxT>(key_r, u, 32));
            }
          }
        }
      }
      // next batch: rotate by UNROLL lanes
      ends >>= UNROLL;
      sid_r = ShflIdx<IdxT>(sid_r, from, 32);
      key_r = ShflIdx<IdxT>(key_r, from, 32);
      if constexpr (WEIGHTED) w_r = ShflRaw<T>(w_r, from);
    }
    if (write_inv) static_cast<IdxT*>(a.inverse_mapping)[key] = tk;
  }

  // Leftover: the last run of the chunk continues into the next chunk.
  int has_tail = 0;
  IdxT tail_row = 0;
  if (open) {
    if (in_first && cont) {
      head_kind = kHeadThrough;
      if (active) StorePartial<NE>(my_head + v * NE, acc);
    } else {
      // the run started in this chunk and continues into the next one
      bool extended = false;
      const int64_t nx = c0 + K;
      if (a.own && a.nnz - nx >= kOwnWindow) {
        const IdxT tkey = __ldg(keys + nx - 1);
        if (__ldg(keys + nx + kOwnWindow - 1) != tkey) {
          // ... and ends within the next kOwnWindow - 1 elements: finish it here
          const IdxT key_x = __ldg(keys + nx + lane);
          IdxT sid_x = __ldg(sids + nx + lane);
          T w_x = T();
          if constexpr (WEIGHTED) w_x = __ldg(weights + nx + lane);
          const unsigned same = __ballot_sync(kFull, key_x == tkey);
          const int e = __ffs(~same) - 1;  // 1 .. kOwnWindow - 1
#pragma unroll 1
          for (int rem = e; rem > 0; rem -= UNROLL) {
            VecT vals[UNROLL];
            float wf[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
              const IdxT sx = ShflIdx<IdxT>(sid_x, u, 32);
              if constexpr (WEIGHTED) wf[u] = ShflWeight<T>(w_x, u);
              vals[u] = LdgVec<V>(RowAddr<IdxT>(gy, sx, row_bytes));
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
              if (u >= rem) break;  // warp-uniform
              if constexpr (WEIGHTED)
                AccumulateVecWeighted<T, V>(vals[u], wf[u], acc);
              else
                AccumulateVec<T, V>(vals[u], acc);
            }
            const int from = (lane + UNROLL) & 31;
            sid_x = ShflIdx<IdxT>(sid_x, from, 32);
            if constexpr (WEIGHTED) w_x = ShflRaw<T>(w_x, from);
          }
          flush(tkey);
          extended = true;
        }
      }
      if (!extended) {
        has_tail = 1;
        tail_row = __ldg(keys + c0 + n - 1);
        if (active) StorePartial<NE>(my_tail + v * NE, acc);
      }
    }
  }
  if (lane == 0 && blockIdx.y == 0) {
    a.meta[chunk * 2 + 0] = head_kind;
    a.meta[chunk * 2 + 1] = has_tail;
    a.meta_row[chunk * 2 + 0] = static_cast<long long>(head_row);
    a.meta_row[chunk * 2 + 1] = static_cast<long long>(tail_row);
    // work list of the fix-up (any order: every chain is summed on its own)
    if (has_tail) a.tail_list[1 + atomicAdd(a.tail_list, 1)] = chunk;
  }
}

}  // namespace cuembed_b200

#endif  // CUEMBED_B200_CSRC_BACKWARD_WARP_CUH_
