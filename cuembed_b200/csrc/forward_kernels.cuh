// forward_kernels.cuh -- device code of the forward gather-pool path.
// See forward.cu for the design notes and the launch logic.
#ifndef CUEMBED_B200_CSRC_FORWARD_KERNELS_CUH_
#define CUEMBED_B200_CSRC_FORWARD_KERNELS_CUH_

#include "common.cuh"

namespace cuembed_b200 {

struct FwdArgs {
  const void* params;
  const void* indices;
  const void* offsets;
  const void* weights;
  void* out;
  int64_t row_bytes;      // input row pitch in bytes
  int64_t out_row_bytes;  // output row pitch in bytes
  int batch;
  int num_hots;
  int off64;
  int mean;
  int out_dt;
  int nvec;
  int lanes;
  int log2_lanes;
  int col_tiles;
  // Addresser indirection (the "embedding cache" hook the reference leaves as a
  // comment, cuembed/include/embedding_lookup_kernels.cuh:114-115 and
  // embedding_lookup_ops.cuh:62-63): row_map[i] >= 0 sends lookup i to row
  // row_map[i] of `cache` (of `params` when cache is null); row_map[i] < 0 keeps
  // row i of `params`.  Null = the identity (MAPPED kernels only).
  const void* row_map;
  const void* cache;
};

template <typename T, int V, bool LOWP>
struct Accum;

// fp32 accumulation (the default, and the only mode for fp32 tables).
template <typename T, int V>
struct Accum<T, V, false> {
  static constexpr int NW = V / 4;
  static constexpr int NE = NW * Elem<T>::kPerWord;
  float acc[NE];
  __device__ __forceinline__ void Zero() {
#pragma unroll
    for (int i = 0; i < NE; ++i) acc[i] = 0.f;
  }
  __device__ __forceinline__ void Add(typename VecBits<V>::type v) {
    AccumulateVec<T, V>(v, acc);
  }
  // Separate multiply and add (no FMA contraction): matches the reference CPU
  // loop, utils/include/embedding_lookup_cpu.hpp:73-75, for any weight value.
  __device__ __forceinline__ void AddWeighted(typename VecBits<V>::type v,
                                              T weight) {
    AccumulateVecWeighted<T, V>(v, Elem<T>::ToFloat(weight), acc);
  }
  __device__ __forceinline__ void Scale(float s) {
#pragma unroll
    for (int i = 0; i < NE; ++i) acc[i] = __fmul_rn(acc[i], s);
  }
  __device__ __forceinline__ void Store(void* out_row, int64_t elem_off,
                                        int out_dt) const {
    StoreFloatsAs<NE>(out_row, elem_off, out_dt, acc);
  }
};

// Accumulation in the 16-bit input type (fp16_math == true in the reference:
// SumT = ElemT, embedding_lookup_cpu.hpp:59): every multiply and add rounds to
// the element type.
template <typename T, int V>
struct Accum<T, V, true> {
  static constexpr int NW = V / 4;
  static constexpr int NE = NW * 2;
  using Pair = typename Elem<T>::Pair;
  Pair acc[NW];
  __device__ __forceinline__ static Pair FromWord(uint32_t w) {
    return *reinterpret_cast<Pair*>(&w);
  }
  __device__ __forceinline__ static Pair Splat(T v) {
    Pair p;
    p.x = v;
    p.y = v;
    return p;
  }
  __device__ __forceinline__ void Zero() {
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = FromWord(0u);
  }
  __device__ __forceinline__ void Add(typename VecBits<V>::type v) {
    uint32_t w[NW];
    Unpack32(v, w);
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = __hadd2_rn(acc[i], FromWord(w[i]));
  }
  __device__ __forceinline__ void AddWeighted(typename VecBits<V>::type v,
                                              T weight) {
    const Pair w2 = Splat(weight);
    uint32_t w[NW];
    Unpack32(v, w);
#pragma unroll
    for (int i = 0; i < NW; ++i)
      acc[i] = __hadd2_rn(acc[i], __hmul2_rn(FromWord(w[i]), w2));
  }
  __device__ __forceinline__ void Scale(float s) {
    T st;
    if constexpr (Elem<T>::kCode == CUEMBED_F16)
      st = __float2half_rn(s);
    else
      st = __float2bfloat16_rn(s);
    const Pair s2 = Splat(st);
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = __hmul2_rn(acc[i], s2);
  }
  __device__ __forceinline__ void Store(void* out_row, int64_t elem_off,
                                        int out_dt) const {
    if (out_dt == Elem<T>::kCode) {
      uint32_t w[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i)
        w[i] = *reinterpret_cast<const uint32_t*>(&acc[i]);
      typename VecBits<V>::type v;
      Pack32(w, &v);
      StcsVec<V>(reinterpret_cast<uint16_t*>(out_row) + elem_off, v);
    } else {
      float f[NE];
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        f[2 * i] = Elem<T>::ToFloat(acc[i].x);
        f[2 * i + 1] = Elem<T>::ToFloat(acc[i].y);
      }
      StoreFloatsAs<NE>(out_row, elem_off, out_dt, f);
    }
  }
};

template <typename T>
__device__ __forceinline__ T ShflElem(unsigned mask, T v, int src, int width) {
  if constexpr (sizeof(T) == 4) {
    return __shfl_sync(mask, v, src, width);
  } else {
    unsigned short b = *reinterpret_cast<unsigned short*>(&v);
    unsigned r = __shfl_sync(mask, static_cast<unsigned>(b), src, width);
    unsigned short rb = static_cast<unsigned short>(r);
    return *reinterpret_cast<T*>(&rb);
  }
}

template <typename IdxT>
__device__ __forceinline__ IdxT ShflIndex(unsigned mask, IdxT v, int src,
                                          int width) {
  if constexpr (sizeof(IdxT) == 8) {
    return static_cast<IdxT>(
        __shfl_sync(mask, static_cast<long long>(v), src, width));
  } else {
    return __shfl_sync(mask, v, src, width);
  }
}

// Byte offset of a table row.  Indices are non-negative, so the 32-bit case is
// one IMAD.WIDE.U32.
template <typename IdxT>
__device__ __forceinline__ uint64_t RowOffset(IdxT row, uint32_t row_bytes) {
  if constexpr (sizeof(IdxT) == 4) {
    return static_cast<uint64_t>(static_cast<uint32_t>(row)) * row_bytes;
  } else {
    return static_cast<uint64_t>(row) * row_bytes;
  }
}

#ifndef FWD_MINB
#define FWD_MINB(UNROLL) ((UNROLL) <= 4 ? 5 : ((UNROLL) <= 8 ? 4 : 2))
#endif

template <typename T, int V, typename IdxT, bool WEIGHTED, bool LOWP,
          int UNROLL, bool MAPPED = false>
__device__ __forceinline__ void FwdPoolBody(const FwdArgs& a) {
  using VecT = typename VecBits<V>::type;
  using AccT = Accum<T, V, LOWP>;
  constexpr unsigned kFull = 0xffffffffu;
  const int G = a.lanes;
  const int lane_g = threadIdx.x & (G - 1);
  const int groups_per_cta = kCtaThreads >> a.log2_lanes;
  const int group = blockIdx.x * groups_per_cta + (threadIdx.x >> a.log2_lanes);
  const int total_groups = gridDim.x * groups_per_cta;
  // Column tile (rows wider than 32 vectors): one grid.y slice per tile.  Lanes
  // past the end of the row read a duplicate of the last vector and never
  // store, so no load in the hot loop is predicated.
  const int v = blockIdx.y * G + lane_g;
  const bool active = v < a.nvec;
  const char* params = static_cast<const char*>(a.params) +
                       static_cast<int64_t>(active ? v : a.nvec - 1) * V;
  // table base + this lane's column offset in one (opaque) register pair, so a
  // row address is a single IMAD.WIDE (common.cuh RowAddr)
  asm volatile("" : "+l"(params));
  // MAPPED: lookups are translated when their index round is requested; a
  // translated word t >= 0 is row t of the cache table, t < 0 is row ~t of the
  // backing table (so one register still carries a lookup through the shuffles).
  const IdxT* __restrict__ row_map = static_cast<const IdxT*>(a.row_map);
  const char* cache = params;
  if constexpr (MAPPED) {
    if (a.cache != nullptr) {
      cache = static_cast<const char*>(a.cache) +
              static_cast<int64_t>(active ? v : a.nvec - 1) * V;
      asm volatile("" : "+l"(cache));
    }
  }
  auto translate = [&](IdxT raw) -> IdxT {
    if constexpr (MAPPED) {
      const IdxT slot = __ldg(row_map + raw);
      return slot >= 0 ? slot : static_cast<IdxT>(~raw);
    } else {
      return raw;
    }
  };
  const IdxT* __restrict__ indices = static_cast<const IdxT*>(a.indices);
  const T* __restrict__ weights = static_cast<const T*>(a.weights);
  uint32_t row_bytes = static_cast<uint32_t>(a.row_bytes);
  asm volatile("" : "+r"(row_bytes));  // a plain 32-bit register: RowAddr is one IMAD.WIDE
  // All 32 lanes of a warp run the loops below in lockstep (bounds are
  // warp-wide maxima) so every shuffle uses the constant full mask; lane
  // groups with shorter bags idle through predicates.
  const int warp_first_bag = group - ((threadIdx.x & 31) >> a.log2_lanes);

#pragma unroll 1
  for (int bag0 = warp_first_bag; bag0 < a.batch; bag0 += total_groups) {
    const int bag = bag0 + ((threadIdx.x & 31) >> a.log2_lanes);
    const bool bag_ok = bag < a.batch;
    int64_t start = 0;
    int len = 0;
    if (bag_ok) {
      if (a.offsets != nullptr) {
        start = LoadOffset(a.offsets, a.off64, bag);
        len = static_cast<int>(LoadOffset(a.offsets, a.off64, bag + 1) - start);
      } else {
        start = static_cast<int64_t>(bag) * a.num_hots;
        len = a.num_hots;
      }
    }
    const int len_max = (G == 32) ? len : __reduce_max_sync(kFull, len);
    const IdxT* __restrict__ bag_idx = indices + start;
    const T* __restrict__ bag_w = weights + start;

    AccT acc;
    acc.Zero();
    float accw = 0.f;

    // Index rounds: G indices per round, loaded coalesced one round ahead.
    // Lanes past the end of the bag hold index 0 (a valid row).
    IdxT idx_cur = 0, idx_nxt = 0;
    T w_cur = T(), w_nxt = T();
    if (lane_g < len) {
      idx_cur = __ldg(bag_idx + lane_g);
      if constexpr (WEIGHTED) w_cur = __ldg(bag_w + lane_g);
    }
    if (G + lane_g < len) {
      idx_nxt = __ldg(bag_idx + G + lane_g);
      if constexpr (WEIGHTED) w_nxt = __ldg(bag_w + G + lane_g);
    }
    // MAPPED: one more round of raw indices in flight, so that a translation is
    // only ever requested for an index that arrived a round ago
    IdxT raw_nn = 0;
    if constexpr (MAPPED) {
      if (2 * G + lane_g < len) raw_nn = __ldg(bag_idx + 2 * G + lane_g);
      idx_cur = translate(idx_cur);
      idx_nxt = translate(idx_nxt);
    }
#pragma unroll 1
    for (int j0 = 0; j0 < len_max; j0 += G) {
      const int cnt = min(G, len - j0);  // may be <= 0 for a finished group
      // The index (and weight) words rotate by one batch per iteration inside
      // the lane group, so every broadcast shuffle has an immediate source lane
      // (G >= UNROLL or G == 8 == UNROLL: rotating by UNROLL stays inside the
      // group).
      IdxT idx_rot = idx_cur;
      T w_rot = w_cur;
      const int rot_from = (lane_g + UNROLL) & (G - 1);
#pragma unroll 1
      for (int jb = 0; jb < G && j0 + jb < len_max; jb += UNROLL) {
        VecT vals[UNROLL];
        T wv[UNROLL];
        // UNROLL independent row loads issued before the first accumulate.
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const IdxT row = ShflIndex<IdxT>(kFull, idx_rot, u, G);
          if constexpr (WEIGHTED) wv[u] = ShflElem<T>(kFull, w_rot, u, G);
          if constexpr (MAPPED) {
            const bool direct = row < 0;
            vals[u] = LdgVec<V>(RowAddr<IdxT>(direct ? params : cache,
                                              direct ? static_cast<IdxT>(~row) : row,
                                              row_bytes));
          } else {
            vals[u] = LdgVec<V>(RowAddr<IdxT>(params, row, row_bytes));
          }
        }
        idx_rot = ShflIndex<IdxT>(kFull, idx_rot, rot_from, G);
        if constexpr (WEIGHTED) w_rot = ShflElem<T>(kFull, w_rot, rot_from, G);
        if (__all_sync(kFull, jb + UNROLL <= cnt)) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if constexpr (WEIGHTED) {
              acc.AddWeighted(vals[u], wv[u]);
              accw = __fadd_rn(accw, Elem<T>::ToFloat(wv[u]));
            } else {
              acc.Add(vals[u]);
            }
          }
        } else {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if (jb + u < cnt) {
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
      idx_cur = idx_nxt;
      idx_nxt = 0;
      if constexpr (WEIGHTED) w_cur = w_nxt;
      if constexpr (MAPPED) {
        idx_nxt = translate(raw_nn);  // requested a round ago
        raw_nn = 0;
        if (j0 + 3 * G + lane_g < len) raw_nn = __ldg(bag_idx + j0 + 3 * G + lane_g);
        if constexpr (WEIGHTED) {
          if (j0 + 2 * G + lane_g < len) w_nxt = __ldg(bag_w + j0 + 2 * G + lane_g);
        }
      } else if (j0 + 2 * G + lane_g < len) {
        idx_nxt = __ldg(bag_idx + j0 + 2 * G + lane_g);
        if constexpr (WEIGHTED) w_nxt = __ldg(bag_w + j0 + 2 * G + lane_g);
      }
    }

    if (a.mean) {
      // cuembed/include/embedding_lookup_ops.cuh:273-285 (sum * (1/accw), zero
      // vector when the accumulated weight is zero) and
      // utils/include/embedding_lookup_cpu.hpp:82-90 for the unweighted case.
      const float denom = WEIGHTED ? accw : static_cast<float>(len);
      if (denom == 0.f) {
        acc.Zero();
      } else {
        acc.Scale(__fdiv_rn(1.0f, denom));
      }
    }
    if (active && bag_ok)
      acc.Store(static_cast<char*>(a.out) + bag * a.out_row_bytes,
                static_cast<int64_t>(v) * AccT::NE, a.out_dt);
  }
}

template <typename T, int V, typename IdxT, bool WEIGHTED, bool LOWP,
          int UNROLL>
__global__ void __launch_bounds__(kCtaThreads, FWD_MINB(UNROLL))
    FwdPoolKernel(const FwdArgs a) {
  FwdPoolBody<T, V, IdxT, WEIGHTED, LOWP, UNROLL>(a);
}

// The pooled kernel with the addresser indirection (FwdArgs::row_map / cache).
template <typename T, int V, typename IdxT, bool WEIGHTED>
__global__ void __launch_bounds__(kCtaThreads, FWD_MINB(8))
    FwdPoolMappedKernel(const FwdArgs a) {
  FwdPoolBody<T, V, IdxT, WEIGHTED, false, 8, true>(a);
}

// Multi-table batched lookup (the reference is "single
// table", README.md:110): up to kMaxTablesPerLaunch pooled lookups of the same
// row shape in ONE launch, grid.z = table.  The descriptors travel as a
// __grid_constant__ kernel parameter (constant bank, no device-side descriptor
// buffer, CUDA-graph capturable); each table gets an equal share of a
// persistent grid, so many small tables cost one launch instead of one each.
constexpr int kMaxTablesPerLaunch = 32;
struct FwdMultiArgs {
  FwdArgs t[kMaxTablesPerLaunch];
};

template <typename T, int V, typename IdxT, bool WEIGHTED>
__global__ void __launch_bounds__(kCtaThreads, FWD_MINB(8))
    FwdPoolMultiKernel(const __grid_constant__ FwdMultiArgs m) {
  FwdPoolBody<T, V, IdxT, WEIGHTED, false, 8>(m.t[blockIdx.z]);
}

// Concat: out[nz, :] = params[indices[nz], :], a pure row gather
// (cuembed/include/embedding_lookup_ops.cuh:300-322).  Four rows per lane group
// in flight.
struct ConcatArgs {
  const void* params;
  const void* indices;
  void* out;
  int64_t row_bytes;
  int64_t nnz;
  int nvec;
  int lanes;
  int log2_lanes;
  int col_tiles;
};

template <int V, typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    FwdConcatKernel(const ConcatArgs a) {
  using VecT = typename VecBits<V>::type;
  constexpr int R = 4;
  const int G = a.lanes;
  const int lane_g = threadIdx.x & (G - 1);
  const int groups_per_cta = kCtaThreads >> a.log2_lanes;
  const int64_t group =
      static_cast<int64_t>(blockIdx.x) * groups_per_cta +
      (threadIdx.x >> a.log2_lanes);
  const int64_t total_groups = static_cast<int64_t>(gridDim.x) * groups_per_cta;
  const char* __restrict__ params = static_cast<const char*>(a.params);
  const IdxT* __restrict__ indices = static_cast<const IdxT*>(a.indices);
  char* __restrict__ out = static_cast<char*>(a.out);

  for (int64_t base = group; base < a.nnz; base += total_groups * R) {
    IdxT row[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t nz = base + r * total_groups;
      row[r] = nz < a.nnz ? __ldg(indices + nz) : IdxT(0);
    }
    for (int ct = 0; ct < a.col_tiles; ++ct) {
      const int v = ct * G + lane_g;
      if (v >= a.nvec) continue;
      const int64_t voff = static_cast<int64_t>(v) * V;
      VecT vals[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t nz = base + r * total_groups;
        if (nz < a.nnz)
          vals[r] = LdgVec<V>(params +
                              static_cast<int64_t>(row[r]) * a.row_bytes + voff);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t nz = base + r * total_groups;
        if (nz < a.nnz) StcsVec<V>(out + nz * a.row_bytes + voff, vals[r]);
      }
    }
  }
}

}  // namespace cuembed_b200

#endif  // CUEMBED_B200_CSRC_FORWARD_KERNELS_CUH_
