// backward.cu -- deterministic segmented-reduction backward for sm_100a.
//
// Replaces EmbeddingBackwardKernel + GradIndexLoader / GradAddresser /
// GradCombiner + CompactSparseIndicesKernel of the reference
// (cuembed/include/embedding_lookup_kernels.cuh:175-302,
//  cuembed/include/embedding_lookup_ops.cuh:499-666), which accumulate in the
// gradient type and resolve block edges with float atomics (order varies from
// run to run).  Here:
//
//   * the sorted COO is cut into fixed chunks of K nonzeros; a lane group (G
//     lanes, one V-byte vector of the row per lane) walks its chunk in order,
//     accumulating weight * grad_y[sample] in fp32 registers, with UNROLL row
//     loads in flight;
//   * a run of equal rows that lies inside one chunk is rounded once and
//     stored directly; a run that crosses chunk edges leaves one head and / or
//     one tail partial per chunk in an fp32 scratch (written only then);
//     lane groups never wait for each other -- no shared memory, no barrier
//     (the first version stitched the chunks of a CTA through shared memory
//     and lost a quarter of its warp time at that barrier: cold rows cost a
//     load and a store per nonzero, hot runs only a load);
//   * a second kernel adds, in chunk order, tail + following heads of every
//     run that crosses chunk edges.
//   Every sum therefore has a fixed association order: results are identical
//   from run to run, with one rounding to the gradient type per element.
//   * inverse_mapping (compressed gradients) is written where each run ends,
//     so the separate compaction kernel of the reference disappears.
//   * 64-bit row offsets throughout (the reference's `int` arithmetic,
//     embedding_lookup_ops.cuh:610-618, overflows at rows*width >= 2^31).
//
// L2/HBM-bound gather + stream-out: no tensor cores.
#include <algorithm>
#include <atomic>

#include "common.cuh"
#include "launch.h"

namespace cuembed_b200 {

struct BwdArgs {
  const void* grad_y;
  const void* keys;  // remapped indices if compressed, else table indices
  const void* tidx;  // table indices (for inverse_mapping) or nullptr
  const void* sids;
  const void* weights;
  void* grad;
  void* inverse_mapping;
  float* scratch;       // [num_chunks][2][width]: head partial, tail partial
  float* group_part;    // [num_chunks / kFixGroup][width]: sums of through groups
  int* tail_list;       // [1 + num_chunks]: count, then the chunks with a tail
  int* meta;            // [num_chunks][2] : head kind, has tail
  long long* meta_row;  // [num_chunks][2] : head row, tail row
  int64_t row_bytes;
  int width;
  int nnz;
  int nvec;
  int lanes;
  int log2_lanes;
  int rounds;      // rounds of G nonzeros per lane group (K = G * rounds)
  int cta_nz;      // nonzeros per CTA = kBwdThreads * rounds
  int num_ctas;
  int num_chunks;  // = num_ctas * lane groups per CTA
  int chunk_nz;  // nonzeros per chunk = lanes * rounds
  int sm_slots;
  // fused sparse optimizer step: opt_kind != 0 makes `grad`
  // the TABLE and every finished row sum an in-place update of its table row
  int warp_path;     // rows that one warp covers: BwdWarpKernel (backward_warp.cuh)
  int own;           // warp path: run ownership across chunk edges
  int opt_kind;      // CUEMBED_OPT_NONE / SGD / ADAGRAD
  float opt_lr;
  float opt_eps;
  float* opt_state;  // Adagrad accumulator [rows][width] fp32
};

constexpr int kHeadNone = 0;
constexpr int kHeadEnds = 1;     // first run of the chunk started earlier, ends here
constexpr int kHeadThrough = 2;  // whole chunk lies inside one earlier run

template <typename IdxT>
__device__ __forceinline__ uint64_t GradRowOffset(IdxT row, uint32_t row_bytes) {
  if constexpr (sizeof(IdxT) == 4) {
    return static_cast<uint64_t>(static_cast<uint32_t>(row)) * row_bytes;
  } else {
    return static_cast<uint64_t>(row) * row_bytes;
  }
}

template <typename T>
__device__ __forceinline__ void StoreOneAs(T* p, float v);
template <>
__device__ __forceinline__ void StoreOneAs<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void StoreOneAs<__half>(__half* p, float v) {
  *p = __float2half_rn(v);
}
template <>
__device__ __forceinline__ void StoreOneAs<__nv_bfloat16>(__nv_bfloat16* p,
                                                          float v) {
  *p = __float2bfloat16_rn(v);
}

constexpr int kBwdThreads = 128;
#ifndef BWD_MINB
#define BWD_MINB 6
#endif

template <int NE>
__device__ __forceinline__ void StorePartial(float* dst, const float* acc) {
  if constexpr (NE % 4 == 0) {
#pragma unroll
    for (int e = 0; e < NE; e += 4)
      *reinterpret_cast<float4*>(dst + e) =
          make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
  } else if constexpr (NE % 2 == 0) {
#pragma unroll
    for (int e = 0; e < NE; e += 2)
      *reinterpret_cast<float2*>(dst + e) = make_float2(acc[e], acc[e + 1]);
  } else {
#pragma unroll
    for (int e = 0; e < NE; ++e) dst[e] = acc[e];
  }
}

// ---- fused optimizer step ---------------------------------------------------
// SGD is a pure "add -lr * g to the row": it goes to the L2 atomic unit as a
// vector reduction (REDG.E.ADD.F16x8 / BF16x8 / F32x4), fire and forget, so the
// SM never waits for the cold table row (a read-modify-write in the SM was
// measured 1.9x slower: every run end waited for DRAM).  Each table row is
// updated by exactly one reduction (one run per row), so the result does not
// depend on timing:
//     p <- p (+) round_T(-(lr * g))      (+) = one add in the table's type, rn
// (fp32 tables: p - lr * g, subnormal results flushed to zero by the unit).
// Adagrad needs sqrt of the NEW state and stays a read-modify-write:
//     s <- s + g * g;  p <- round_T(p - (lr * g) / (sqrt(s) + eps))
// every operation rounded separately in fp32, so a numpy float32 restatement
// reproduces it bit for bit.
template <typename T, int NW>
__device__ __forceinline__ void RedAddWords(void* addr, const uint32_t* w) {
  if constexpr (sizeof(T) == 4) {
    if constexpr (NW == 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr),
                   "f"(__uint_as_float(w[0])), "f"(__uint_as_float(w[1])),
                   "f"(__uint_as_float(w[2])), "f"(__uint_as_float(w[3]))
                   : "memory");
    else if constexpr (NW == 2)
      asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(addr),
                   "f"(__uint_as_float(w[0])), "f"(__uint_as_float(w[1]))
                   : "memory");
    else
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr),
                   "f"(__uint_as_float(w[0]))
                   : "memory");
  } else if constexpr (Elem<T>::kCode == CUEMBED_F16) {
    if constexpr (NW == 4)
      asm volatile("red.global.add.noftz.v4.f16x2 [%0], {%1,%2,%3,%4};" ::"l"(
                       addr),
                   "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                   : "memory");
    else if constexpr (NW == 2)
      asm volatile("red.global.add.noftz.v2.f16x2 [%0], {%1,%2};" ::"l"(addr),
                   "r"(w[0]), "r"(w[1])
                   : "memory");
    else
      asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(w[0])
                   : "memory");
  } else {
    if constexpr (NW == 4)
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(
                       addr),
                   "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                   : "memory");
    else if constexpr (NW == 2)
      asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1,%2};" ::"l"(addr),
                   "r"(w[0]), "r"(w[1])
                   : "memory");
    else
      asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(addr),
                   "r"(w[0])
                   : "memory");
  }
}

// One element through the same unit (fix-up kernel: runs that cross chunks).
template <typename T>
__device__ __forceinline__ void RedAddOne(T* addr, float v) {
  if constexpr (sizeof(T) == 4) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
  } else if constexpr (Elem<T>::kCode == CUEMBED_F16) {
    const __half h = __float2half_rn(v);
    asm volatile("red.global.add.noftz.f16 [%0], %1;" ::"l"(addr),
                 "h"(*reinterpret_cast<const unsigned short*>(&h))
                 : "memory");
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    asm volatile("red.global.add.noftz.bf16 [%0], %1;" ::"l"(addr),
                 "h"(*reinterpret_cast<const unsigned short*>(&h))
                 : "memory");
  }
}

__device__ __forceinline__ float AdagradStep(float lr, float eps, float p,
                                             float g, float* state) {
  const float s = __fadd_rn(*state, __fmul_rn(g, g));
  *state = s;
  return __fsub_rn(p,
                   __fdiv_rn(__fmul_rn(lr, g), __fadd_rn(__fsqrt_rn(s), eps)));
}

// Update of NE consecutive elements of table row `row` with the finished
// gradient sums g[NE] (one V-byte vector of the row per lane).  old_vec is the
// current content of the vector (Adagrad only).
template <typename T, int V, int NE>
__device__ __forceinline__ void ApplyUpdateVec(
    const BwdArgs& a, char* table_row, int64_t row, int64_t elem_off,
    const float* g, typename VecBits<V>::type old_vec) {
  using VecT = typename VecBits<V>::type;
  constexpr int NW = V / 4;
  VecT* pv = reinterpret_cast<VecT*>(table_row + elem_off * sizeof(T));
  uint32_t w[NW];
  if (a.opt_kind == CUEMBED_OPT_SGD) {
    const float neg_lr = -a.opt_lr;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      float f[Elem<T>::kPerWord];
#pragma unroll
      for (int k = 0; k < Elem<T>::kPerWord; ++k)
        f[k] = __fmul_rn(neg_lr, g[i * Elem<T>::kPerWord + k]);
      w[i] = Elem<T>::FloatToWord(f);
    }
    RedAddWords<T, NW>(pv, w);
    return;
  }
  Unpack32(old_vec, w);
  float st[NE];
  float* sp = a.opt_state + row * a.width + elem_off;
#pragma unroll
  for (int e = 0; e < NE; ++e) st[e] = sp[e];
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    float f[Elem<T>::kPerWord];
    Elem<T>::WordToFloat(w[i], f);
#pragma unroll
    for (int k = 0; k < Elem<T>::kPerWord; ++k) {
      const int e = i * Elem<T>::kPerWord + k;
      f[k] = AdagradStep(a.opt_lr, a.opt_eps, f[k], g[e], &st[e]);
    }
    w[i] = Elem<T>::FloatToWord(f);
  }
  VecT out;
  Pack32(w, &out);
  *pv = out;
  StorePartial<NE>(sp, st);
}

}  // namespace cuembed_b200
#include "backward_warp.cuh"
namespace cuembed_b200 {

// FUSED_OPT: CUEMBED_OPT_NONE (write the gradient), _SGD (vector reductions
// into the table, nothing read), _ADAGRAD (old table rows requested with the
// gradient rows).
template <typename T, int V, typename IdxT, bool WEIGHTED, int UNROLL,
          int FUSED_OPT>
__global__ void __launch_bounds__(kBwdThreads, BWD_MINB)
    BwdSegReduceKernel(const BwdArgs a) {
  using VecT = typename VecBits<V>::type;
  constexpr int NW = V / 4;
  constexpr int NE = NW * Elem<T>::kPerWord;
  constexpr unsigned kFull = 0xffffffffu;

  const int tid = threadIdx.x;
  const int G = a.lanes;
  const int lane_g = tid & (G - 1);
  const int g = tid >> a.log2_lanes;
  const int groups_per_cta = kBwdThreads >> a.log2_lanes;
  const int chunk = blockIdx.x * groups_per_cta + g;
  float* __restrict__ my_head =
      a.scratch + (static_cast<size_t>(chunk) * 2 + 0) * a.width;
  float* __restrict__ my_tail =
      a.scratch + (static_cast<size_t>(chunk) * 2 + 1) * a.width;
  const int gl0 = (tid & 31) & ~(G - 1);  // first lane of this group in the warp
  // One bit per lane group of the warp (bit k*G), to test "does any group end
  // a run at position j" with a warp-uniform branch.
  unsigned patt = 0;
  for (int k = 0; k < 32; k += G) patt |= 1u << k;

  const IdxT* __restrict__ keys = static_cast<const IdxT*>(a.keys);
  const IdxT* __restrict__ sids = static_cast<const IdxT*>(a.sids);
  const T* __restrict__ weights = static_cast<const T*>(a.weights);
  const IdxT* __restrict__ tidx = static_cast<const IdxT*>(a.tidx);
  const uint32_t row_bytes = static_cast<uint32_t>(a.row_bytes);

  const int K = G * a.rounds;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * a.cta_nz +
                     static_cast<int64_t>(g) * K;
  int n_g = 0;
  if (c0 < a.nnz)
    n_g = static_cast<int>(min(static_cast<int64_t>(K), a.nnz - c0));

  const int v = blockIdx.y * G + lane_g;
  const bool active = v < a.nvec;
  const char* __restrict__ gy = static_cast<const char*>(a.grad_y) +
                                static_cast<int64_t>(active ? v : a.nvec - 1) * V;

  // Does the first run of this chunk continue a run of the previous chunk?
  bool cont = false;
  if (n_g > 0 && c0 > 0) cont = __ldg(keys + c0 - 1) == __ldg(keys + c0);
  bool in_first = true;
  bool open = false;
  int head_kind = kHeadNone;
  IdxT head_row = 0;

  float acc[NE];
#pragma unroll
  for (int e = 0; e < NE; ++e) acc[e] = 0.f;

  // The (key, sample, weight, next key, table row) of a round are requested one
  // round ahead, so no index load sits in front of the row loads of a round
  // (ncu: 8 % of the warp samples waited for keys[i + 1], another 8 % for the
  // table row that inverse_mapping receives at every run end).
  IdxT n_key = 0, n_sid = 0, n_knext = 0, n_tk = 0;
  T n_w = T();
  auto request = [&](int r) {
    const int cnt = max(0, min(G, n_g - r * G));
    const int64_t i = c0 + r * G + lane_g;
    n_key = 0;
    n_sid = 0;
    n_knext = 0;
    n_tk = 0;
    n_w = T();
    if (lane_g < cnt) {
      n_key = __ldg(keys + i);
      n_sid = __ldg(sids + i);
      if constexpr (WEIGHTED) n_w = __ldg(weights + i);
      if (i + 1 < a.nnz) n_knext = __ldg(keys + i + 1);
      if (tidx != nullptr) n_tk = __ldg(tidx + i);
    }
  };
  request(0);

#pragma unroll 1
  for (int r = 0; r < a.rounds; ++r) {
    const int cnt = max(0, min(G, n_g - r * G));
    const int64_t i = c0 + r * G + lane_g;
    const IdxT key = n_key, sid = n_sid, tk = n_tk;
    const T w = n_w;
    const bool end =
        lane_g < cnt && ((i == a.nnz - 1) || (n_knext != key));
    if (r + 1 < a.rounds) request(r + 1);
    // Fused optimizer: the table rows this round will update are cold (DRAM);
    // ask for them now so that the read-modify-write at the run end finds them
    // in L2 (without this every batch with a run end waited for DRAM).
    if constexpr (FUSED_OPT == CUEMBED_OPT_ADAGRAD) {
      if (end && blockIdx.y == 0) {
        const char* trow_ptr = static_cast<const char*>(a.grad) +
                               GradRowOffset<IdxT>(key, row_bytes);
        for (uint32_t off = 0; off < row_bytes; off += 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(trow_ptr + off));
      }
    }
    const unsigned endsw = __ballot_sync(kFull, end);
    const int cnt_max = (G == 32) ? cnt : __reduce_max_sync(kFull, cnt);

#pragma unroll 1
    for (int jb = 0; jb < cnt_max; jb += UNROLL) {
      VecT vals[UNROLL];
      T wv[UNROLL];
      VecT pold[FUSED_OPT == CUEMBED_OPT_ADAGRAD ? UNROLL : 1];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int src = (jb + u) & (G - 1);
        IdxT s;
        if constexpr (sizeof(IdxT) == 8)
          s = static_cast<IdxT>(
              __shfl_sync(kFull, static_cast<long long>(sid), src, G));
        else
          s = __shfl_sync(kFull, sid, src, G);
        if constexpr (WEIGHTED) {
          if constexpr (sizeof(T) == 4) {
            wv[u] = __shfl_sync(kFull, w, src, G);
          } else {
            const unsigned short b = *reinterpret_cast<const unsigned short*>(&w);
            unsigned rb = __shfl_sync(kFull, static_cast<unsigned>(b), src, G);
            unsigned short rs = static_cast<unsigned short>(rb);
            wv[u] = *reinterpret_cast<T*>(&rs);
          }
        }
        // Positions past the end of the chunk carry sample id 0: a valid,
        // harmless load that is never accumulated.
        vals[u] = LdgVec<V>(gy + GradRowOffset<IdxT>(s, row_bytes));
        // Fused optimizer: the table row that a run ending at this position
        // will update is requested together with the gradient rows, so the
        // read-modify-write does not wait for DRAM at every run end.
        if constexpr (FUSED_OPT == CUEMBED_OPT_ADAGRAD) {
          const IdxT kr = ShflIdx<IdxT>(key, src, G);
          pold[u] = VecT();
          if (active && jb + u < G && ((endsw >> (gl0 + jb + u)) & 1u) != 0u)
            pold[u] = *reinterpret_cast<const VecT*>(
                static_cast<const char*>(a.grad) +
                GradRowOffset<IdxT>(kr, row_bytes) + static_cast<int64_t>(v) * V);
        }
      }
      // Fast path: a full batch in which no lane group of the warp ends a run
      // (the interior of long runs, ~half of all nonzeros under a power law).
      if constexpr (UNROLL <= 8) {
        const unsigned batch_bits = (patt * ((1u << UNROLL) - 1u)) << jb;
        if (G >= UNROLL && (endsw & batch_bits) == 0u &&
            __all_sync(kFull, jb + UNROLL <= cnt)) {
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if constexpr (WEIGHTED)
              AccumulateVecWeighted<T, V>(vals[u], Elem<T>::ToFloat(wv[u]), acc);
            else
              AccumulateVec<T, V>(vals[u], acc);
          }
          open = true;
          continue;
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int j = jb + u;
        if (j < cnt) {
          if constexpr (WEIGHTED)
            AccumulateVecWeighted<T, V>(vals[u], Elem<T>::ToFloat(wv[u]), acc);
          else
            AccumulateVec<T, V>(vals[u], acc);
          open = true;
        }
        // Warp-uniform test: does any lane group of this warp end a run at j?
        if (j < G && (endsw & (patt << j)) != 0u) {
          const IdxT krow = ShflIdx<IdxT>(key, j, G);
          const IdxT trow = ShflIdx<IdxT>(tk, j, G);
          if (((endsw >> (gl0 + j)) & 1u) != 0u) {
            if (in_first && cont) {
              // Run began in an earlier chunk: partial, added by the fix-up.
              if (active) StorePartial<NE>(my_head + v * NE, acc);
              head_kind = kHeadEnds;
              head_row = krow;
            } else if (active) {
              char* out_row = static_cast<char*>(a.grad) +
                              GradRowOffset<IdxT>(krow, row_bytes);
              if constexpr (FUSED_OPT == CUEMBED_OPT_NONE)
                StoreFloatsAs<NE>(out_row, static_cast<int64_t>(v) * NE,
                                  Elem<T>::kCode, acc);
              else
                ApplyUpdateVec<T, V, NE>(a, out_row, static_cast<int64_t>(krow),
                                         static_cast<int64_t>(v) * NE, acc,
                                         pold[FUSED_OPT == CUEMBED_OPT_ADAGRAD ? u : 0]);
            }
            if (a.inverse_mapping != nullptr && lane_g == 0 &&
                blockIdx.y == 0) {
              static_cast<IdxT*>(a.inverse_mapping)[krow] = trow;
            }
#pragma unroll
            for (int e = 0; e < NE; ++e) acc[e] = 0.f;
            in_first = false;
            open = false;
          }
        }
      }
    }
  }

  // Leftover: the last run of the chunk continues into the next chunk.
  int has_tail = 0;
  IdxT tail_row = 0;
  if (open) {
    if (in_first && cont) {
      head_kind = kHeadThrough;
      if (active) StorePartial<NE>(my_head + v * NE, acc);
    } else {
      has_tail = 1;
      tail_row = __ldg(keys + c0 + n_g - 1);
      if (active) StorePartial<NE>(my_tail + v * NE, acc);
    }
  }
  if (lane_g == 0 && blockIdx.y == 0) {
    a.meta[chunk * 2 + 0] = head_kind;
    a.meta[chunk * 2 + 1] = has_tail;
    a.meta_row[chunk * 2 + 0] = static_cast<long long>(head_row);
    a.meta_row[chunk * 2 + 1] = static_cast<long long>(tail_row);
    // work list of the fix-up (any order: every chain is summed on its own)
    if (has_tail) a.tail_list[1 + atomicAdd(a.tail_list, 1)] = chunk;
  }
}

// Fix-up, level 1: a GROUP of kFixGroup consecutive chunks that all lie inside
// one long run ("through" chunks) is summed, in chunk order, into one partial
// row.  Only the interiors of very hot rows produce such groups; they make the
// chains of level 2 up to kFixGroup times shorter (a 262144-sample batch puts
// the hottest row into 2048 consecutive chunks).
constexpr int kFixGroup = 32;

__global__ void __launch_bounds__(kCtaThreads)
    BwdGroupKernel(const BwdArgs a) {
  __shared__ int s_all;
  const int g = blockIdx.x;
  const int c_first = g * kFixGroup;
  if (c_first + kFixGroup > a.num_chunks) return;
  const int tid = threadIdx.x;
  if (tid < 32) {
    const bool through = a.meta[(c_first + tid) * 2 + 0] == kHeadThrough;
    const bool all = __all_sync(0xffffffffu, through);
    if (tid == 0) s_all = all ? 1 : 0;
  }
  __syncthreads();
  if (s_all == 0) return;
  const int col = blockIdx.y * kCtaThreads + tid;
  if (col >= a.width) return;
  const size_t pitch = static_cast<size_t>(2) * a.width;
  const float* __restrict__ p = a.scratch + static_cast<size_t>(c_first) * pitch + col;
  float acc = 0.f;
#pragma unroll
  for (int h = 0; h < kFixGroup; h += 16) {
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = p[static_cast<size_t>(h + u) * pitch];
#pragma unroll
    for (int u = 0; u < 16; ++u)
      acc = (h + u == 0) ? v[u] : __fadd_rn(acc, v[u]);
  }
  a.group_part[static_cast<size_t>(g) * a.width + col] = acc;
}

// Fix-up, level 2: adds, in chunk order, the partials of every run that crosses
// chunk edges: tail of the chunk where the run starts + heads of the following
// chunks, whole "through" groups taken from level 1.  Small CTAs (64 threads,
// 32 per SM) walk the list of chunks with a tail that the main kernel wrote:
// most chains have two elements, so what matters is how many of them an SM
// holds while each waits for its three or four dependent round trips, and not
// launching a CTA for every chunk that has nothing to do (one 256-thread CTA
// per chunk took 30 us at C2, one 64-thread CTA per chunk 24 us, one warp per
// chunk 38 us).  A thread owns every 64th column, four columns and
// eight chain elements in flight.  Every sum has a fixed association.
constexpr int kFixThreads = 64;

template <typename T>
__global__ void __launch_bounds__(kFixThreads)
    BwdFixupKernel(const BwdArgs a) {
  __shared__ int s_len;
  const int tid = threadIdx.x;
  const int n_tails = a.tail_list[0];
  for (int li = blockIdx.x; li < n_tails; li += gridDim.x) {
  const int c0 = a.tail_list[1 + li];
  // chain = chunks c0+1 .. c0+len; the last one has kind "ends".
  if (tid == 0) s_len = 0x7fffffff;
  __syncthreads();
  for (int base = c0 + 1; base < a.num_chunks; base += kFixThreads) {
    const int c = base + tid;
    if (c < a.num_chunks && a.meta[c * 2 + 0] != kHeadThrough)
      atomicMin(&s_len, c - c0);
    __syncthreads();
    const int found = s_len;
    __syncthreads();  // everyone has read s_len before anyone updates it
    if (found != 0x7fffffff) break;
  }
  int len = s_len;
  int last_through = c0 + len - 1;  // chunks c0+1 .. last_through are "through"
  if (len == 0x7fffffff) {          // malformed input guard
    len = a.num_chunks - 1 - c0;
    last_through = c0 + len;
  }
  // a "none" head at the end of the chain means the run ended exactly at the
  // chunk edge (cannot happen for a tail, kept as a guard): exclude it.
  if (c0 + len < a.num_chunks && a.meta[(c0 + len) * 2 + 0] == kHeadNone) --len;
  const int end = c0 + len;  // inclusive
  // whole groups inside the "through" part of the chain: [g0, g1)
  int g0 = (c0 + 1 + kFixGroup - 1) / kFixGroup;
  int g1 = (last_through + 1) / kFixGroup;
  if (g1 <= g0) g0 = g1 = 0x3fffffff / kFixGroup;  // none
  const int lead_end = min(end, g0 * kFixGroup - 1);  // singles before the groups

  const size_t pitch = static_cast<size_t>(2) * a.width;
  const long long row = a.meta_row[c0 * 2 + 1];
  constexpr int kCols = 4;
  constexpr int kRows = 8;
  for (int cb = tid; cb < a.width; cb += kFixThreads * kCols) {
    float acc[kCols];
    bool live[kCols];
#pragma unroll
    for (int i = 0; i < kCols; ++i) {
      live[i] = cb + kFixThreads * i < a.width;
      acc[i] = live[i] ? a.scratch[static_cast<size_t>(c0) * pitch + a.width + cb +
                                   kFixThreads * i]
                       : 0.f;
    }
    // acc += rows p[0], p[stride], ... (count rows), in order
    auto add_rows = [&](const float* __restrict__ p, size_t stride, int count) {
      int r = 0;
      for (; r + kRows <= count; r += kRows) {
        float v[kRows][kCols];
#pragma unroll
        for (int u = 0; u < kRows; ++u)
#pragma unroll
          for (int i = 0; i < kCols; ++i)
            v[u][i] = live[i] ? p[static_cast<size_t>(r + u) * stride + kFixThreads * i]
                              : 0.f;
#pragma unroll
        for (int u = 0; u < kRows; ++u)
#pragma unroll
          for (int i = 0; i < kCols; ++i) acc[i] = __fadd_rn(acc[i], v[u][i]);
      }
      for (; r < count; ++r) {
        float v[kCols];
#pragma unroll
        for (int i = 0; i < kCols; ++i)
          v[i] = live[i] ? p[static_cast<size_t>(r) * stride + kFixThreads * i] : 0.f;
#pragma unroll
        for (int i = 0; i < kCols; ++i)
          if (live[i]) acc[i] = __fadd_rn(acc[i], v[i]);
      }
    };
    int c = c0 + 1;
    if (lead_end >= c) {
      add_rows(a.scratch + static_cast<size_t>(c) * pitch + cb, pitch,
               lead_end - c + 1);
      c = lead_end + 1;
    }
    if (c <= end && g1 > g0) {
      add_rows(a.group_part + static_cast<size_t>(g0) * a.width + cb,
               static_cast<size_t>(a.width), g1 - g0);
      c = g1 * kFixGroup;
      if (c <= end)
        add_rows(a.scratch + static_cast<size_t>(c) * pitch + cb, pitch,
                 end - c + 1);
    }
#pragma unroll
    for (int i = 0; i < kCols; ++i) {
      if (!live[i]) continue;
      T* dst = static_cast<T*>(a.grad) + row * a.width + cb + kFixThreads * i;
      if (a.opt_kind == CUEMBED_OPT_SGD) {
        RedAddOne<T>(dst, __fmul_rn(-a.opt_lr, acc[i]));
        continue;
      }
      float val = acc[i];
      if (a.opt_kind == CUEMBED_OPT_ADAGRAD) {
        float* sp = a.opt_state + row * a.width + cb + kFixThreads * i;
        float st = *sp;
        val = AdagradStep(a.opt_lr, a.opt_eps, Elem<T>::ToFloat(*dst), val, &st);
        *sp = st;
      }
      StoreOneAs<T>(dst, val);
    }
  }
  __syncthreads();  // s_len is reused for the next chain of this CTA
  }  // chains of this CTA
}

namespace {

int Log2i(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

struct BwdLayout {
  int rounds;
  int cta_nz;
  int num_ctas;
  int num_chunks;
  size_t scratch_off, group_off, meta_off, row_off, tail_off, total;
};

BwdLayout MakeBwdLayout(int nnz, int embed_width, int lanes, int dtype) {
  BwdLayout L;
  static const int rounds_env = EnvInt("CUEMBED_BWD_ROUNDS", 0);
  int rounds = rounds_env;
  if (rounds <= 0) {
    // Nonzeros per CTA: as large as possible (short cross-CTA chains, little
    // scratch) while leaving >= ~8 CTAs per SM.
    const int64_t target = static_cast<int64_t>(nnz) /
                           (static_cast<int64_t>(GetDeviceInfo().sm_count) * 8);
    rounds = 1;
    while (rounds < 8 && static_cast<int64_t>(rounds) * 2 * kCtaThreads <= target)
      rounds *= 2;
    // the same ~256 nonzeros per lane group whatever the row width (narrow
    // rows use fewer lanes per group): keeps the chains of the fix-up short
    if (rounds == 8) rounds = 8 * (32 / lanes);
  }
  L.rounds = rounds;
  L.cta_nz = rounds * kBwdThreads;
  L.num_ctas = nnz > 0 ? (nnz + L.cta_nz - 1) / L.cta_nz : 0;
  L.num_chunks = L.num_ctas * (kBwdThreads / lanes);
  size_t off = 0;
  L.scratch_off = off;
  off += AlignUp(static_cast<size_t>(L.num_chunks) * 2 * embed_width * sizeof(float),
                 256);
  L.group_off = off;
  off += AlignUp(static_cast<size_t>(L.num_chunks / kFixGroup + 1) * embed_width * sizeof(float),
                 256);
  L.meta_off = off;
  off += AlignUp(static_cast<size_t>(L.num_chunks) * 2 * sizeof(int), 256);
  L.row_off = off;
  off += AlignUp(static_cast<size_t>(L.num_chunks) * 2 * sizeof(long long), 256);
  L.tail_off = off;
  off += AlignUp(static_cast<size_t>(L.num_chunks + 1) * sizeof(int), 256);
  L.total = off > 0 ? off : 256;
  return L;
}

template <typename T, int V, typename IdxT, bool WEIGHTED>
void LaunchSegReduce(const BwdArgs& a, int col_tiles, cudaStream_t stream) {
  dim3 grid(a.num_ctas, col_tiles);
  // the warp walker: one chunk per warp, BWD_WARP_CTA_WARPS chunks per CTA
  const dim3 wgrid(a.num_chunks / BWD_WARP_CTA_WARPS, col_tiles);
  if (a.warp_path && a.opt_kind == CUEMBED_OPT_SGD)
    BwdWarpKernel<T, V, IdxT, WEIGHTED, CUEMBED_OPT_SGD>
        <<<wgrid, kBwdWarpThreads, 0, stream>>>(a);
  else if (a.warp_path && a.opt_kind == CUEMBED_OPT_NONE)
    BwdWarpKernel<T, V, IdxT, WEIGHTED, CUEMBED_OPT_NONE>
        <<<wgrid, kBwdWarpThreads, 0, stream>>>(a);
  else if (a.opt_kind == CUEMBED_OPT_ADAGRAD)
    BwdSegReduceKernel<T, V, IdxT, WEIGHTED, 4, CUEMBED_OPT_ADAGRAD>
        <<<grid, kBwdThreads, 0, stream>>>(a);  // 4: room for the old rows
  else if (a.opt_kind == CUEMBED_OPT_SGD)
    BwdSegReduceKernel<T, V, IdxT, WEIGHTED, 8, CUEMBED_OPT_SGD>
        <<<grid, kBwdThreads, 0, stream>>>(a);
  else
    BwdSegReduceKernel<T, V, IdxT, WEIGHTED, 8, CUEMBED_OPT_NONE>
        <<<grid, kBwdThreads, 0, stream>>>(a);
  const int wtiles = (a.width + kCtaThreads - 1) / kCtaThreads;
  const int groups = a.num_chunks / kFixGroup;
  if (groups > 0)
    BwdGroupKernel<<<dim3(groups, wtiles), kCtaThreads, 0, stream>>>(a);
  // persistent over the list of chunks with a tail (written by the main kernel)
  const int fix_ctas = std::min(a.num_chunks, a.sm_slots * 32);
  BwdFixupKernel<T><<<fix_ctas, kFixThreads, 0, stream>>>(a);
  CountLaunch(groups > 0 ? 3 : 2);
}

template <typename T, int V>
void LaunchSegReduceIdx(const BwdArgs& a, int idx_type, bool weighted,
                        int col_tiles, cudaStream_t stream) {
  if (idx_type == CUEMBED_I64) {
    if (weighted)
      LaunchSegReduce<T, V, int64_t, true>(a, col_tiles, stream);
    else
      LaunchSegReduce<T, V, int64_t, false>(a, col_tiles, stream);
  } else {
    if (weighted)
      LaunchSegReduce<T, V, int32_t, true>(a, col_tiles, stream);
    else
      LaunchSegReduce<T, V, int32_t, false>(a, col_tiles, stream);
  }
}

template <typename T>
void LaunchSegReduceVec(const BwdArgs& a, int vec_bytes, int idx_type,
                        bool weighted, int col_tiles, cudaStream_t stream) {
  if (vec_bytes == 16)
    LaunchSegReduceIdx<T, 16>(a, idx_type, weighted, col_tiles, stream);
  else if (vec_bytes == 8)
    LaunchSegReduceIdx<T, 8>(a, idx_type, weighted, col_tiles, stream);
  else
    LaunchSegReduceIdx<T, 4>(a, idx_type, weighted, col_tiles, stream);
}

}  // namespace

namespace {
struct OptParams {
  int kind;
  float lr, eps;
  float* state;
};

int LaunchBackwardImpl(const void* grad_y, int dtype, int embed_width,
                       int num_grad_embedding_rows, int nnz, int idx_type,
                       const void* transpose_indices,
                       const void* transpose_sample_ids,
                       const void* transpose_remapped_indices,
                       const void* transpose_weights, int skip_grad_init,
                       void* grad_embedding, void* inverse_mapping, char* work,
                       size_t* lwork, const OptParams& opt, cudaStream_t stream);
}  // namespace

int LaunchBackward(const void* grad_y, int dtype, int embed_width,
                   int num_grad_embedding_rows, int nnz, int idx_type,
                   const void* transpose_indices,
                   const void* transpose_sample_ids,
                   const void* transpose_remapped_indices,
                   const void* transpose_weights, int skip_grad_init,
                   void* grad_embedding, void* inverse_mapping, char* work,
                   size_t* lwork, cudaStream_t stream) {
  const OptParams none = {CUEMBED_OPT_NONE, 0.f, 0.f, nullptr};
  return LaunchBackwardImpl(grad_y, dtype, embed_width, num_grad_embedding_rows,
                            nnz, idx_type, transpose_indices,
                            transpose_sample_ids, transpose_remapped_indices,
                            transpose_weights, skip_grad_init, grad_embedding,
                            inverse_mapping, work, lwork, none, stream);
}

// Fused backward + sparse optimizer step (the reference
// lists "optimizer" as a future kernel type, README.md:119): the row sums of
// the backward are applied to the table in place instead of being written out
// as a gradient, so neither the compressed-index pass nor the gradient round
// trip (293 MB at C2) exists.
int LaunchBackwardUpdate(const void* grad_y, int dtype, int embed_width, int nnz,
                         int idx_type, const void* transpose_indices,
                         const void* transpose_sample_ids,
                         const void* transpose_weights, int optimizer, float lr,
                         float eps, void* params, float* state, char* work,
                         size_t* lwork, cudaStream_t stream) {
  if (optimizer != CUEMBED_OPT_SGD && optimizer != CUEMBED_OPT_ADAGRAD)
    return CUEMBED_ERR_ARGUMENT;
  if (work != nullptr && nnz > 0 &&
      (params == nullptr ||
       (optimizer == CUEMBED_OPT_ADAGRAD && state == nullptr)))
    return CUEMBED_ERR_ARGUMENT;
  const OptParams opt = {optimizer, lr, eps, state};
  return LaunchBackwardImpl(grad_y, dtype, embed_width, 0, nnz, idx_type,
                            transpose_indices, transpose_sample_ids, nullptr,
                            transpose_weights, /*skip_grad_init=*/1, params,
                            nullptr, work, lwork, opt, stream);
}

namespace {
int LaunchBackwardImpl(const void* grad_y, int dtype, int embed_width,
                       int num_grad_embedding_rows, int nnz, int idx_type,
                       const void* transpose_indices,
                       const void* transpose_sample_ids,
                       const void* transpose_remapped_indices,
                       const void* transpose_weights, int skip_grad_init,
                       void* grad_embedding, void* inverse_mapping, char* work,
                       size_t* lwork, const OptParams& opt,
                       cudaStream_t stream) {
  if (lwork == nullptr || nnz < 0 || embed_width <= 0 ||
      num_grad_embedding_rows < 0)
    return CUEMBED_ERR_ARGUMENT;
  if (dtype < 0 || dtype > 2 || idx_type < 0 || idx_type > 1)
    return CUEMBED_ERR_DTYPE;
  const int64_t row_bytes = static_cast<int64_t>(embed_width) * ElemSize(dtype);
  if (row_bytes % 4 != 0) return CUEMBED_ERR_ROW_BYTES;
  // Scratch is sized for the aligned case (fewest lanes per row = most chunks);
  // a misaligned call uses narrower vectors, i.e. fewer, wider lane groups.
  RowShape shape;
  MakeRowShape(embed_width, dtype, &shape);
  // Rows that a warp covers exactly (32 lanes x 4 / 8 / 16 bytes, wider rows in
  // column tiles) take the warp-uniform walker (backward_warp.cuh).
  static const int warp_env = EnvInt("CUEMBED_BWD_WARP", 1);
  const int warp_vec =
      (warp_env == 0 || opt.kind == CUEMBED_OPT_ADAGRAD)
          ? 0
          : (row_bytes == 128 ? 4
                                        : (row_bytes == 256
                                               ? 8
                                               : (row_bytes % 512 == 0 ? 16 : 0)));
  const BwdLayout L =
      MakeBwdLayout(nnz, embed_width, warp_vec != 0 ? 32 : shape.lanes, dtype);
  if (work == nullptr) {
    *lwork = L.total;
    return CUEMBED_OK;
  }
  if (*lwork < L.total) return CUEMBED_ERR_WORKSPACE;
  if (grad_embedding == nullptr && (nnz > 0 || !skip_grad_init))
    return CUEMBED_ERR_ARGUMENT;

  // Zero-fill unless told otherwise (cuembed/include/embedding_lookup.cuh:455-461).
  if (!skip_grad_init && num_grad_embedding_rows > 0) {
    if (cudaMemsetAsync(grad_embedding, 0,
                        static_cast<size_t>(num_grad_embedding_rows) *
                            static_cast<size_t>(row_bytes),
                        stream) != cudaSuccess)
      return CUEMBED_ERR_CUDA;
  }
  if (nnz == 0) return CUEMBED_OK;
  if (grad_y == nullptr || transpose_indices == nullptr ||
      transpose_sample_ids == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  if (transpose_remapped_indices != nullptr && inverse_mapping == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(work) & 15) != 0)
    return CUEMBED_ERR_ARGUMENT;

  // Vector width limited by the actual pointer alignment.
  int v = shape.vec_bytes;
  const uint64_t bits = reinterpret_cast<uint64_t>(grad_y) |
                        reinterpret_cast<uint64_t>(grad_embedding) |
                        static_cast<uint64_t>(row_bytes);
  while (v > 4 && (bits % v) != 0) v /= 2;
  if ((bits % v) != 0) return CUEMBED_ERR_ARGUMENT;
  // warp path: its vector width must be allowed by the pointers (otherwise the
  // generic kernel runs with narrower vectors, i.e. 32 lanes as well, so the
  // layout sized above still fits)
  const bool warp_path = warp_vec != 0 && (bits % warp_vec) == 0;
  if (warp_path) v = warp_vec;

  BwdArgs a;
  a.grad_y = grad_y;
  a.keys = transpose_remapped_indices != nullptr ? transpose_remapped_indices
                                                 : transpose_indices;
  a.tidx = transpose_remapped_indices != nullptr ? transpose_indices : nullptr;
  a.sids = transpose_sample_ids;
  a.weights = transpose_weights;
  a.grad = grad_embedding;
  a.inverse_mapping =
      transpose_remapped_indices != nullptr ? inverse_mapping : nullptr;
  a.scratch = reinterpret_cast<float*>(work + L.scratch_off);
  a.group_part = reinterpret_cast<float*>(work + L.group_off);
  a.meta = reinterpret_cast<int*>(work + L.meta_off);
  a.meta_row = reinterpret_cast<long long*>(work + L.row_off);
  a.tail_list = reinterpret_cast<int*>(work + L.tail_off);
  if (cudaMemsetAsync(a.tail_list, 0, sizeof(int), stream) != cudaSuccess)
    return CUEMBED_ERR_CUDA;
  a.row_bytes = row_bytes;
  a.width = embed_width;
  a.nnz = nnz;
  a.nvec = static_cast<int>(row_bytes / v);
  a.lanes = Pow2Ceil(a.nvec) < 32 ? Pow2Ceil(a.nvec) : 32;
  a.log2_lanes = Log2i(a.lanes);
  a.rounds = L.rounds;
  a.cta_nz = L.cta_nz;
  a.num_ctas = L.num_ctas;
  a.num_chunks = L.num_ctas * (kBwdThreads / a.lanes);
  a.warp_path = warp_path ? 1 : 0;
  static const int own_env = EnvInt("CUEMBED_BWD_OWN", 1);
  a.own = own_env != 0 ? 1 : 0;
  a.opt_kind = opt.kind;
  a.opt_lr = opt.lr;
  a.opt_eps = opt.eps;
  a.opt_state = opt.state;
  a.chunk_nz = a.lanes * a.rounds;
  a.sm_slots = GetDeviceInfo().sm_count;
  const int col_tiles = (a.nvec + a.lanes - 1) / a.lanes;
  const bool weighted = transpose_weights != nullptr;

  if (dtype == CUEMBED_F32)
    LaunchSegReduceVec<float>(a, v, idx_type, weighted, col_tiles, stream);
  else if (dtype == CUEMBED_F16)
    LaunchSegReduceVec<__half>(a, v, idx_type, weighted, col_tiles, stream);
  else
    LaunchSegReduceVec<__nv_bfloat16>(a, v, idx_type, weighted, col_tiles,
                                      stream);
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}
}  // namespace

}  // namespace cuembed_b200
