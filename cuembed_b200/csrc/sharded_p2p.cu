// sharded_p2p.cu -- the row-sharded forward / backward exchange fused with the
// kernels over NVLink 5 / NVSwitch PEER MEMORY (new functionality; the
// reference is single-GPU, README.md:110; specified by BASELINE.json
// `north_star` and SURVEY.md 8(e)).
//
// One process per GPU.  Every rank maps the exchange buffers of every other
// rank into its address space (CUDA IPC, cuembed_peer_*), after which the data
// path needs no collective library call:
//
//   forward   ShardPoolPushKernel: a lane group walks one bag of the GLOBAL
//             batch, keeps the lookups that fall into the rows this rank owns
//             (warp-ballot compaction into a shared-memory queue, so that row
//             loads stay UNROLL deep), pools them in bag order in fp32 and
//             STORES the partial row straight into the slot
//             [this rank][bag - owner * per] of the rank that owns the bag.
//             The NVLink transfer therefore overlaps the gather bag by bag, and
//             there is neither a local partial buffer nor a select pass.
//             Consecutive bags belong to consecutive owners (starting at
//             rank + 1), so pushes are spread over the kernel and the receivers.
//   signal    one release-store per peer (after the kernel boundary, which
//             makes the pushes visible system-wide).
//   finalize  ShardReduceFinalizeKernel: waits (acquire loads) for the flag of
//             every rank, then sums the `world` slots of each of its bags IN
//             RANK ORDER (deterministic, unlike a ring), applies the mean by
//             the GLOBAL bag length and casts.
//   backward  the adjoint of that exchange is an all-gather of grad_y:
//             cuembed_shard_allgather_push copies this rank's slice into every
//             peer's gather buffer with the copy engines (no SM is taken from
//             the transpose kernels that run meanwhile) and signals;
//             cuembed_shard_wait holds the stream until all slices arrived.
//
// Exchange buffers are double-buffered by the host (epoch parity): a rank can
// only overwrite a peer's slot two steps after the peer consumed it, and it
// cannot get there before having seen that peer's next signal.
// A wait that sees no signal for kWaitTimeoutNs gives up and raises the status
// word instead of hanging the device.
#include <algorithm>
#include <atomic>
#include <cstring>

#include "common.cuh"
#include "forward_kernels.cuh"
#include "launch.h"

namespace cuembed_b200 {

constexpr int kMaxWorld = CUEMBED_MAX_WORLD;
// How long a wait spins before it gives up (a stalled peer: first-iteration
// compile, data-loader hiccup, checkpoint).  Default 60 s; CUEMBED_PEER_TIMEOUT_MS
// or cuembed_shard_set_timeout_ms() change it.  A wait that gives up raises the
// status word AND poisons what depends on it (NaN output / NaN gather slice),
// so stale peer data can never pass for a result.
std::atomic<long long> g_wait_timeout_ms{-1};

unsigned long long WaitTimeoutNs() {
  long long ms = g_wait_timeout_ms.load();
  if (ms < 0) {
    ms = EnvInt("CUEMBED_PEER_TIMEOUT_MS", 60000);
    if (ms <= 0) ms = 60000;
    g_wait_timeout_ms.store(ms);
  }
  return static_cast<unsigned long long>(ms) * 1000000ull;
}

struct PeerPtrs {
  void* p[kMaxWorld];
};

// ------------------------------------------------------------ flag helpers

__device__ __forceinline__ void StoreReleaseSys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v)
               : "memory");
}

__device__ __forceinline__ unsigned LoadAcquireSys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p)
               : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long GlobalTimerNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The first `world` threads of the CTA each wait for one rank's flag to reach
// `epoch` (wrap-around safe), then the CTA proceeds.  Returns a bit mask of
// the ranks whose signal did not arrive within `timeout_ns` (0 = all arrived);
// the status word keeps 1 + the last such rank until the host clears it.
__device__ __forceinline__ unsigned WaitForPeers(const unsigned* flags,
                                                 int world, unsigned epoch,
                                                 unsigned* status,
                                                 unsigned long long timeout_ns) {
  __shared__ unsigned s_missing;
  if (threadIdx.x == 0) s_missing = 0u;
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < world) {
    const unsigned long long t0 = GlobalTimerNs();
    while (static_cast<int>(LoadAcquireSys(flags + threadIdx.x) - epoch) < 0) {
      if (GlobalTimerNs() - t0 > timeout_ns) {
        atomicExch(status, 1u + threadIdx.x);
        atomicOr(&s_missing, 1u << threadIdx.x);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  return s_missing;
}

__global__ void ShardSignalKernel(PeerPtrs flags, int world, int rank,
                                  unsigned epoch) {
  const int j = threadIdx.x;
  if (j < world) {
    __threadfence_system();
    StoreReleaseSys(static_cast<unsigned*>(flags.p[j]) + rank, epoch);
  }
}

// Holds the stream until every rank has signalled `epoch`.  If a rank never
// does, its slice of `poison` ([world] slices of poison_bytes, may be null) is
// filled with 0xff bytes -- NaN in fp32 / fp16 / bf16 -- so that whatever is
// computed from the incomplete gather buffer is visibly invalid.
__global__ void ShardWaitKernel(const unsigned* flags, int world,
                                unsigned epoch, unsigned* status,
                                unsigned long long timeout_ns,
                                unsigned char* poison, size_t poison_bytes) {
  const unsigned missing = WaitForPeers(flags, world, epoch, status, timeout_ns);
  if (missing == 0u || poison == nullptr) return;
  for (int r = 0; r < world; ++r) {
    if (((missing >> r) & 1u) == 0u) continue;
    unsigned char* p = poison + static_cast<size_t>(r) * poison_bytes;
    for (size_t i = threadIdx.x; i < poison_bytes; i += blockDim.x) p[i] = 0xff;
  }
}

// ---------------------------------------------------------- pool and push

struct PushArgs {
  const void* params;   // this rank's rows [lo, hi)
  const void* indices;  // global batch, replicated
  const void* offsets;
  const void* weights;
  PeerPtrs slots;  // slots.p[o]: slot [this rank] in owner o's exchange buffer
  int* counts;     // optional: lookups of each bag that fell into [lo, hi)
  int64_t row_bytes;
  int64_t out_row_bytes;
  long long lo, hi;
  int batch;
  int per;  // bags per owner
  int rank;
  int world;
  int num_hots;
  int off64;
  int out_dt;
  int nvec;
  int lanes;
  int log2_lanes;
};

constexpr int kQueuePerWarp = 128;

#ifndef PUSH_MINB
#define PUSH_MINB 4
#endif
template <typename T, int V, typename IdxT, bool WEIGHTED>
__global__ void __launch_bounds__(kCtaThreads, PUSH_MINB)
    ShardPoolPushKernel(const PushArgs a) {
  using VecT = typename VecBits<V>::type;
  using AccT = Accum<T, V, false>;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int UNROLL = 8;
  __shared__ uint32_t s_rows[kWarpsPerCta * kQueuePerWarp];
  __shared__ T s_w[WEIGHTED ? kWarpsPerCta * kQueuePerWarp : 1];

  const int G = a.lanes;
  const int lane = threadIdx.x & 31;
  const int lane_g = threadIdx.x & (G - 1);
  const int gw = lane >> a.log2_lanes;  // lane group within the warp
  const int groups_per_cta = kCtaThreads >> a.log2_lanes;
  const int group = blockIdx.x * groups_per_cta + (threadIdx.x >> a.log2_lanes);
  const int total_groups = gridDim.x * groups_per_cta;
  const int qcap = kQueuePerWarp >> (5 - a.log2_lanes);  // entries per group
  const int qbase = (threadIdx.x >> 5) * kQueuePerWarp + gw * qcap;
  uint32_t* __restrict__ q = s_rows + qbase;
  T* __restrict__ qw = s_w + (WEIGHTED ? qbase : 0);
  const unsigned gshift = gw << a.log2_lanes;
  const unsigned gmask = G == 32 ? kFull : ((1u << G) - 1u);
  const unsigned lt = (1u << lane_g) - 1u;

  const int v = blockIdx.y * G + lane_g;
  const bool active = v < a.nvec;
  // table base + this lane's column offset in one opaque register pair and a
  // plain 32-bit pitch: a row address is a single IMAD.WIDE (common.cuh RowAddr;
  // ncu had counted 576 warp instructions per bag at 8 ranks, 16 of them
  // reloading the table pointer from the constant bank for every batch)
  const char* params = static_cast<const char*>(a.params) +
                       static_cast<int64_t>(active ? v : a.nvec - 1) * V;
  asm volatile("" : "+l"(params));
  const IdxT* __restrict__ indices = static_cast<const IdxT*>(a.indices);
  const T* __restrict__ weights = static_cast<const T*>(a.weights);
  uint32_t row_bytes = static_cast<uint32_t>(a.row_bytes);
  asm volatile("" : "+r"(row_bytes));
  const int warp_first = group - gw;
  const unsigned long long span = static_cast<unsigned long long>(a.hi - a.lo);

  // seq = smod + world * sdiv, advanced by total_groups per bag without a
  // division (two runtime divisions per bag were ~40 instructions)
  const int step_div = total_groups / a.world;
  const int step_mod = total_groups - step_div * a.world;
  int sdiv = (warp_first + gw) / a.world;
  int smod = (warp_first + gw) - sdiv * a.world;

#pragma unroll 1
  for (int seq0 = warp_first; seq0 < a.batch; seq0 += total_groups) {
    const int seq = seq0 + gw;
    const bool bag_ok = seq < a.batch;
    // Consecutive bags go to consecutive owners, starting at rank + 1: the
    // NVLink pushes are spread evenly over the whole kernel (they overlap the
    // bags pooled for the rank itself) and over the receivers.
    int owner = a.rank + 1 + smod;
    if (owner >= a.world) owner -= a.world;
    const int bag = owner * a.per + sdiv;
    smod += step_mod;
    sdiv += step_div;
    if (smod >= a.world) {
      smod -= a.world;
      ++sdiv;
    }
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
    const bool fits = len_max <= qcap;
    const IdxT* __restrict__ bag_idx = indices + start;
    const T* __restrict__ bag_w = weights + start;

    AccT acc;
    acc.Zero();
    int qn = 0;    // queued, not yet pooled
    int kept = 0;  // lookups of this bag owned by this rank

    IdxT idx_cur = 0;
    T w_cur = T();
    if (lane_g < len) {
      idx_cur = __ldg(bag_idx + lane_g);
      if constexpr (WEIGHTED) w_cur = __ldg(bag_w + lane_g);
    }
#pragma unroll 1
    for (int j0 = 0;; j0 += G) {
      const bool more = j0 < len_max;
      if (more) {
        // One round of G indices: keep the owned ones, in bag order.
        IdxT idx_nxt = 0;
        T w_nxt = T();
        if (j0 + G + lane_g < len) {
          idx_nxt = __ldg(bag_idx + j0 + G + lane_g);
          if constexpr (WEIGHTED) w_nxt = __ldg(bag_w + j0 + G + lane_g);
        }
        // owned <=> 0 <= row - lo < hi - lo: one unsigned compare (the shard has
        // fewer than 2^32 rows, checked by the launcher)
        const unsigned long long rel =
            static_cast<unsigned long long>(static_cast<long long>(idx_cur) - a.lo);
        const bool keep = (j0 + lane_g < len) && rel < span;
        const unsigned m = (__ballot_sync(kFull, keep) >> gshift) & gmask;
        if (keep) {
          const int slot = qn + __popc(m & lt);
          q[slot] = static_cast<uint32_t>(rel);
          if constexpr (WEIGHTED) qw[slot] = w_cur;
        }
        qn += __popc(m);
        kept += __popc(m);
        idx_cur = idx_nxt;
        if constexpr (WEIGHTED) w_cur = w_nxt;
      }
      // a bag that fits the queue as a whole is pooled once, at its end
      if (!more || (!fits && __any_sync(kFull, qn + G > qcap))) {
        __syncwarp();  // the queue entries of all lanes are visible
        // Pool the queued rows in queue (= bag) order, UNROLL loads in flight.
        const int qmax = (G == 32) ? qn : __reduce_max_sync(kFull, qn);
#pragma unroll 1
        for (int jb = 0; jb < qmax; jb += UNROLL) {
          VecT vals[UNROLL];
          T wv[UNROLL];
          const bool whole = (G == 32) ? (jb + UNROLL <= qn)
                                       : __all_sync(kFull, jb + UNROLL <= qn);
          if (whole) {  // a full batch in every lane group: nothing predicated
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
              vals[u] = LdgVec<V>(RowAddr<uint32_t>(params, q[jb + u], row_bytes));
              if constexpr (WEIGHTED) wv[u] = qw[jb + u];
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
              if constexpr (WEIGHTED)
                acc.AddWeighted(vals[u], wv[u]);
              else
                acc.Add(vals[u]);
            }
            continue;
          }
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if (jb + u < qn) {
              vals[u] = LdgVec<V>(RowAddr<uint32_t>(params, q[jb + u], row_bytes));
              if constexpr (WEIGHTED) wv[u] = qw[jb + u];
            }
          }
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if (jb + u < qn) {
              if constexpr (WEIGHTED)
                acc.AddWeighted(vals[u], wv[u]);
              else
                acc.Add(vals[u]);
            }
          }
        }
        qn = 0;
        __syncwarp();
      }
      if (!more) break;
    }

    if (bag_ok) {
      if (active) {
        char* dst = static_cast<char*>(a.slots.p[owner]) +
                    static_cast<int64_t>(bag - owner * a.per) * a.out_row_bytes;
        acc.Store(dst, static_cast<int64_t>(v) * AccT::NE, a.out_dt);
      }
      if (a.counts != nullptr && blockIdx.y == 0 && lane_g == 0)
        a.counts[bag] = kept;
    }
  }
}

// ------------------------------------------------------ reduce + finalize

template <typename PT>
__device__ __forceinline__ float PartialToFloat(PT v);
template <>
__device__ __forceinline__ float PartialToFloat<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float PartialToFloat<__half>(__half v) {
  return __half2float(v);
}
template <>
__device__ __forceinline__ float PartialToFloat<__nv_bfloat16>(
    __nv_bfloat16 v) {
  return __bfloat162float(v);
}

template <typename PT, int N>
struct alignas(sizeof(PT) * N) PartialVec {
  PT e[N];
};

// out[s, :] = cast(scale(s) * (slot[0][s, :] + slot[1][s, :] + ...)), slots in
// rank order.  VEC elements per thread.
template <typename PT, typename WT, int VEC>
__global__ void __launch_bounds__(kCtaThreads)
    ShardReduceFinalizeKernel(const PT* __restrict__ slots, int world,
                              const unsigned* flags, unsigned epoch,
                              unsigned* status, unsigned long long timeout_ns,
                              int n_samples, int width,
                              int mean, const void* offsets, int off64,
                              int num_hots, int sample0,
                              const WT* __restrict__ weights,
                              void* __restrict__ out, int out_dt) {
  // a missing peer poisons the whole output (NaN), it is never summed stale
  const bool poisoned =
      WaitForPeers(flags, world, epoch, status, timeout_ns) != 0u;
  using PV = PartialVec<PT, VEC>;
  const int64_t total = static_cast<int64_t>(n_samples) * width;
  const int64_t nvec = total / VEC;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const PV* __restrict__ src = reinterpret_cast<const PV*>(slots);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
       i < nvec; i += stride) {
    float acc[VEC];
    {
      const PV p0 = src[i];
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = PartialToFloat<PT>(p0.e[k]);
    }
    for (int r = 1; r < world; ++r) {
      const PV p = src[i + static_cast<int64_t>(r) * nvec];
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        acc[k] = __fadd_rn(acc[k], PartialToFloat<PT>(p.e[k]));
    }
    if (mean) {
      const int s = static_cast<int>(i * VEC / width);
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
      const float scale = denom == 0.f ? 0.f : __fdiv_rn(1.0f, denom);
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        acc[k] = denom == 0.f ? 0.f : __fmul_rn(acc[k], scale);
    }
    if (poisoned) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] = __int_as_float(0x7fc00000);
    }
    StoreFloatsAs<VEC>(out, i * VEC, out_dt, acc);
  }
}

// ----------------------------------------------------------- concat push

// Sharded concat: every lookup has exactly one owner, so the owner copies the
// row straight to its final place in the bag owner's output -- an all-to-all
// made of peer stores, bit-exact, no reduction.
struct ConcatPushArgs {
  const void* params;
  const void* indices;
  PeerPtrs outs;  // outs.p[o]: output [per * num_hots, width] of rank o
  int64_t row_bytes;
  int64_t nnz;
  int64_t nnz_per;  // per * num_hots
  int64_t first_nz;
  long long lo, hi;
  int nvec;
  int lanes;
  int log2_lanes;
  int col_tiles;
};

template <int V, typename IdxT>
__global__ void __launch_bounds__(kCtaThreads)
    ShardConcatPushKernel(const ConcatPushArgs a) {
  using VecT = typename VecBits<V>::type;
  constexpr int R = 4;
  const int G = a.lanes;
  const int lane_g = threadIdx.x & (G - 1);
  const int groups_per_cta = kCtaThreads >> a.log2_lanes;
  const int64_t group = static_cast<int64_t>(blockIdx.x) * groups_per_cta +
                        (threadIdx.x >> a.log2_lanes);
  const int64_t total_groups = static_cast<int64_t>(gridDim.x) * groups_per_cta;
  const char* __restrict__ params = static_cast<const char*>(a.params);
  const IdxT* __restrict__ indices = static_cast<const IdxT*>(a.indices);
  for (int64_t base = group; base < a.nnz; base += total_groups * R) {
    long long row[R];
    int64_t nz[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t seq = base + r * total_groups;
      int64_t z = seq + a.first_nz;
      if (z >= a.nnz) z -= a.nnz;
      nz[r] = z;
      row[r] = seq < a.nnz ? static_cast<long long>(__ldg(indices + z)) : -1;
      if (row[r] < a.lo || row[r] >= a.hi) row[r] = -1;
    }
    for (int ct = 0; ct < a.col_tiles; ++ct) {
      const int v = ct * G + lane_g;
      if (v >= a.nvec) continue;
      const int64_t voff = static_cast<int64_t>(v) * V;
      VecT vals[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (row[r] >= 0)
          vals[r] = LdgVec<V>(params + (row[r] - a.lo) * a.row_bytes + voff);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row[r] >= 0) {
          const int owner = static_cast<int>(nz[r] / a.nnz_per);
          char* dst = static_cast<char*>(a.outs.p[owner]) +
                      (nz[r] - owner * a.nnz_per) * a.row_bytes + voff;
          *reinterpret_cast<VecT*>(dst) = vals[r];
        }
      }
    }
  }
}

// ------------------------------------------------------------ host side

namespace {

int Log2i(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

int ResidentGrid(const void* kernel, PerDeviceInt* occ_cache,
                 int64_t work_ctas) {
  int occ = occ_cache->Get();
  if (occ == 0) {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kCtaThreads, 0);
    occ = n > 0 ? n : 1;
    occ_cache->Set(occ);
  }
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * occ;
  const int64_t g = work_ctas < cap ? work_ctas : cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

template <typename T, int V, typename IdxT, bool WEIGHTED>
void LaunchPush(const PushArgs& a, int col_tiles, cudaStream_t stream) {
  auto k = ShardPoolPushKernel<T, V, IdxT, WEIGHTED>;
  static PerDeviceInt occ;
  const int groups_per_cta = kCtaThreads / a.lanes;
  const int64_t work = (a.batch + groups_per_cta - 1) / groups_per_cta;
  const int grid = ResidentGrid(reinterpret_cast<const void*>(k), &occ, work);
  k<<<dim3(grid, col_tiles), kCtaThreads, 0, stream>>>(a);
  CountLaunch();
}

template <typename T, int V>
void LaunchPushIdx(const PushArgs& a, int col_tiles, int idx_type,
                   bool weighted, cudaStream_t stream) {
  if (idx_type == CUEMBED_I64) {
    if (weighted)
      LaunchPush<T, V, int64_t, true>(a, col_tiles, stream);
    else
      LaunchPush<T, V, int64_t, false>(a, col_tiles, stream);
  } else {
    if (weighted)
      LaunchPush<T, V, int32_t, true>(a, col_tiles, stream);
    else
      LaunchPush<T, V, int32_t, false>(a, col_tiles, stream);
  }
}

template <typename T>
void LaunchPushVec(const PushArgs& a, int col_tiles, int v, int idx_type,
                   bool weighted, cudaStream_t stream) {
  if (v == 16)
    LaunchPushIdx<T, 16>(a, col_tiles, idx_type, weighted, stream);
  else if (v == 8)
    LaunchPushIdx<T, 8>(a, col_tiles, idx_type, weighted, stream);
  else
    LaunchPushIdx<T, 4>(a, col_tiles, idx_type, weighted, stream);
}

template <typename PT, typename WT>
void LaunchReduceVec(const void* slots, int world, const unsigned* flags,
                     unsigned epoch, unsigned* status, int n, int width,
                     int mean, const void* offsets, int off64, int num_hots,
                     int sample0, const void* weights, void* out, int out_dt,
                     cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(n) * width;
  const int vec = width % 4 == 0 ? 4 : (width % 2 == 0 ? 2 : 1);
  const int64_t ctas = (total / vec + kCtaThreads - 1) / kCtaThreads;
  const int64_t cap = static_cast<int64_t>(GetDeviceInfo().sm_count) * 8;
  const int grid = static_cast<int>(ctas < cap ? (ctas < 1 ? 1 : ctas) : cap);
#define REDUCE(VEC)                                                          \
  ShardReduceFinalizeKernel<PT, WT, VEC><<<grid, kCtaThreads, 0, stream>>>(  \
      static_cast<const PT*>(slots), world, flags, epoch, status,            \
      WaitTimeoutNs(), n, width, mean, offsets, off64, num_hots, sample0,    \
      static_cast<const WT*>(weights), out, out_dt)
  if (vec == 4) {
    REDUCE(4);
  } else if (vec == 2) {
    REDUCE(2);
  } else {
    REDUCE(1);
  }
#undef REDUCE
  CountLaunch();
}

template <typename PT>
void LaunchReduceW(const void* slots, int world, const unsigned* flags,
                   unsigned epoch, unsigned* status, int n, int width, int mean,
                   const void* offsets, int off64, int num_hots, int sample0,
                   const void* weights, int weight_dtype, void* out, int out_dt,
                   cudaStream_t stream) {
  if (weights == nullptr || weight_dtype == CUEMBED_F32)
    LaunchReduceVec<PT, float>(slots, world, flags, epoch, status, n, width,
                               mean, offsets, off64, num_hots, sample0, weights,
                               out, out_dt, stream);
  else if (weight_dtype == CUEMBED_F16)
    LaunchReduceVec<PT, __half>(slots, world, flags, epoch, status, n, width,
                                mean, offsets, off64, num_hots, sample0,
                                weights, out, out_dt, stream);
  else
    LaunchReduceVec<PT, __nv_bfloat16>(slots, world, flags, epoch, status, n,
                                       width, mean, offsets, off64, num_hots,
                                       sample0, weights, out, out_dt, stream);
}

bool FillPeers(void* const* host_ptrs, int world, PeerPtrs* out) {
  if (host_ptrs == nullptr || world < 1 || world > kMaxWorld) return false;
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < world; ++i) {
    if (host_ptrs[i] == nullptr) return false;
    out->p[i] = host_ptrs[i];
  }
  return true;
}

int CudaRc() {
  return cudaPeekAtLastError() == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

}  // namespace
}  // namespace cuembed_b200

using namespace cuembed_b200;  // NOLINT

extern "C" {

// ------------------------------------------------------------ peer memory

int cuembed_peer_alloc(size_t bytes, void** ptr) {
  if (ptr == nullptr || bytes == 0) return CUEMBED_ERR_ARGUMENT;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return CUEMBED_ERR_CUDA;
  }
  if (cudaMemset(p, 0, bytes) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(p);
    return CUEMBED_ERR_CUDA;
  }
  *ptr = p;
  return CUEMBED_OK;
}

int cuembed_peer_free(void* ptr) {
  return cudaFree(ptr) == cudaSuccess ? CUEMBED_OK : CUEMBED_ERR_CUDA;
}

int cuembed_peer_export(void* ptr, unsigned char* handle) {
  static_assert(sizeof(cudaIpcMemHandle_t) == CUEMBED_PEER_HANDLE_BYTES,
                "handle size");
  if (ptr == nullptr || handle == nullptr) return CUEMBED_ERR_ARGUMENT;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) {
    cudaGetLastError();
    return CUEMBED_ERR_CUDA;
  }
  std::memcpy(handle, &h, sizeof(h));
  return CUEMBED_OK;
}

int cuembed_peer_open(const unsigned char* handle, void** ptr) {
  if (ptr == nullptr || handle == nullptr) return CUEMBED_ERR_ARGUMENT;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) !=
      cudaSuccess) {
    cudaGetLastError();
    return CUEMBED_ERR_CUDA;
  }
  *ptr = p;
  return CUEMBED_OK;
}

int cuembed_peer_close(void* ptr) {
  return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? CUEMBED_OK
                                                   : CUEMBED_ERR_CUDA;
}

// ------------------------------------------------------------- forward

int cuembed_shard_pool_push(const void* local_params, int in_dtype,
                            int embed_width, const void* indices, int idx_type,
                            const void* offsets, int off_type,
                            const void* weights, int batch_size, int num_hots,
                            long long row_lo, long long row_hi,
                            void* const* slot_ptrs, int world, int rank,
                            int partial_dtype, int* counts,
                            cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!((offsets != nullptr && num_hots == 0) ||
        (offsets == nullptr && num_hots > 0)))
    return CUEMBED_ERR_CSR_XOR_FIXED;
  if (in_dtype < 0 || in_dtype > 2 || partial_dtype < 0 || partial_dtype > 2 ||
      idx_type < 0 || idx_type > 1)
    return CUEMBED_ERR_DTYPE;
  if (batch_size < 0 || embed_width <= 0 || world < 1 || world > kMaxWorld ||
      rank < 0 || rank >= world || batch_size % world != 0 ||
      row_hi < row_lo || row_hi - row_lo > 0xffffffffll)
    return CUEMBED_ERR_ARGUMENT;
  if (batch_size == 0) return CUEMBED_OK;
  if (indices == nullptr || (row_hi > row_lo && local_params == nullptr))
    return CUEMBED_ERR_ARGUMENT;
  const int64_t row_bytes =
      static_cast<int64_t>(embed_width) * ElemSize(in_dtype);
  if (row_bytes % 4 != 0) return CUEMBED_ERR_ROW_BYTES;
  PushArgs a;
  if (!FillPeers(slot_ptrs, world, &a.slots)) return CUEMBED_ERR_ARGUMENT;
  const int64_t out_row_bytes =
      static_cast<int64_t>(embed_width) * ElemSize(partial_dtype);
  RowShape shape;
  MakeRowShape(embed_width, in_dtype, &shape);
  // widest vector every input AND every peer slot pointer allows (same rule
  // as the single-GPU forward; a misaligned pointer is an argument error)
  int v = 0;
  {
    const uint64_t in_bits = reinterpret_cast<uint64_t>(local_params) |
                             static_cast<uint64_t>(row_bytes);
    uint64_t out_bits = static_cast<uint64_t>(out_row_bytes);
    for (int i = 0; i < world; ++i)
      out_bits |= reinterpret_cast<uint64_t>(slot_ptrs[i]);
    for (int c = shape.vec_bytes; c >= 4 && v == 0; c /= 2) {
      const int64_t out_vec =
          static_cast<int64_t>(c) * ElemSize(partial_dtype) / ElemSize(in_dtype);
      if (in_bits % c == 0 && out_bits % (out_vec > 16 ? 16 : out_vec) == 0)
        v = c;
    }
  }
  if (v == 0) return CUEMBED_ERR_ARGUMENT;
  a.params = local_params;
  a.indices = indices;
  a.offsets = offsets;
  a.weights = weights;
  a.counts = counts;
  a.row_bytes = row_bytes;
  a.out_row_bytes = out_row_bytes;
  a.lo = row_lo;
  a.hi = row_hi;
  a.batch = batch_size;
  a.per = batch_size / world;
  a.rank = rank;
  a.world = world;
  a.num_hots = num_hots;
  a.off64 = off_type == CUEMBED_I64;
  a.out_dt = partial_dtype;
  a.nvec = static_cast<int>(row_bytes / v);
  a.lanes = Pow2Ceil(a.nvec) < 32 ? Pow2Ceil(a.nvec) : 32;
  a.log2_lanes = Log2i(a.lanes);
  const int col_tiles = (a.nvec + a.lanes - 1) / a.lanes;
  const bool weighted = weights != nullptr;
  if (in_dtype == CUEMBED_F32)
    LaunchPushVec<float>(a, col_tiles, v, idx_type, weighted, stream);
  else if (in_dtype == CUEMBED_F16)
    LaunchPushVec<__half>(a, col_tiles, v, idx_type, weighted, stream);
  else
    LaunchPushVec<__nv_bfloat16>(a, col_tiles, v, idx_type, weighted, stream);
  return CudaRc();
}

int cuembed_shard_concat_push(const void* local_params, int dtype,
                              int embed_width, const void* indices,
                              int idx_type, int batch_size, int num_hots,
                              long long row_lo, long long row_hi,
                              void* const* out_ptrs, int world, int rank,
                              cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (dtype < 0 || dtype > 2 || idx_type < 0 || idx_type > 1)
    return CUEMBED_ERR_DTYPE;
  if (batch_size < 0 || num_hots <= 0 || embed_width <= 0 || world < 1 ||
      world > kMaxWorld || rank < 0 || rank >= world ||
      batch_size % world != 0 || row_hi < row_lo)
    return CUEMBED_ERR_ARGUMENT;
  if (batch_size == 0) return CUEMBED_OK;
  if (indices == nullptr || (row_hi > row_lo && local_params == nullptr))
    return CUEMBED_ERR_ARGUMENT;
  const int64_t row_bytes = static_cast<int64_t>(embed_width) * ElemSize(dtype);
  if (row_bytes % 4 != 0) return CUEMBED_ERR_ROW_BYTES;
  ConcatPushArgs c;
  if (!FillPeers(out_ptrs, world, &c.outs)) return CUEMBED_ERR_ARGUMENT;
  uint64_t bits = reinterpret_cast<uint64_t>(local_params) |
                  static_cast<uint64_t>(row_bytes);
  for (int i = 0; i < world; ++i) bits |= reinterpret_cast<uint64_t>(out_ptrs[i]);
  int v = 16;
  while (v > 1 && (bits & (v - 1)) != 0) v /= 2;
  if (v < 4) return CUEMBED_ERR_ARGUMENT;
  c.params = local_params;
  c.indices = indices;
  c.row_bytes = row_bytes;
  c.nnz = static_cast<int64_t>(batch_size) * num_hots;
  c.nnz_per = c.nnz / world;
  c.first_nz = ((rank + 1) % world) * c.nnz_per;
  c.lo = row_lo;
  c.hi = row_hi;
  c.nvec = static_cast<int>(row_bytes / v);
  c.lanes = Pow2Ceil(c.nvec) < 32 ? Pow2Ceil(c.nvec) : 32;
  c.log2_lanes = Log2i(c.lanes);
  c.col_tiles = (c.nvec + c.lanes - 1) / c.lanes;
  const int groups_per_cta = kCtaThreads / c.lanes;
  const int64_t work = (c.nnz + groups_per_cta * 4 - 1) / (groups_per_cta * 4);
#define CONCAT_PUSH(VV, IdxT)                                                \
  {                                                                          \
    auto k = ShardConcatPushKernel<VV, IdxT>;                                \
    static PerDeviceInt occ;                                                 \
    const int grid = ResidentGrid(reinterpret_cast<const void*>(k), &occ,    \
                                  work);                                     \
    k<<<grid, kCtaThreads, 0, stream>>>(c);                                  \
  }
  if (idx_type == CUEMBED_I64) {
    if (v == 16) CONCAT_PUSH(16, int64_t) else if (v == 8)
        CONCAT_PUSH(8, int64_t) else CONCAT_PUSH(4, int64_t)
  } else {
    if (v == 16) CONCAT_PUSH(16, int32_t) else if (v == 8)
        CONCAT_PUSH(8, int32_t) else CONCAT_PUSH(4, int32_t)
  }
#undef CONCAT_PUSH
  CountLaunch();
  return CudaRc();
}

int cuembed_shard_signal(void* const* flag_ptrs, int world, int rank,
                         int channel, unsigned epoch,
                         cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (channel < 0 || channel >= CUEMBED_PEER_CHANNELS || rank < 0 ||
      rank >= world)
    return CUEMBED_ERR_ARGUMENT;
  PeerPtrs f;
  if (!FillPeers(flag_ptrs, world, &f)) return CUEMBED_ERR_ARGUMENT;
  for (int i = 0; i < world; ++i)
    f.p[i] = static_cast<unsigned*>(f.p[i]) + channel * kMaxWorld;
  ShardSignalKernel<<<1, 32, 0, stream>>>(f, world, rank, epoch);
  CountLaunch();
  return CudaRc();
}

int cuembed_shard_set_timeout_ms(long long timeout_ms) {
  if (timeout_ms <= 0) return CUEMBED_ERR_ARGUMENT;
  g_wait_timeout_ms.store(timeout_ms);
  return CUEMBED_OK;
}

int cuembed_shard_wait(const void* flags, int world, int channel,
                       unsigned epoch, void* poison, size_t poison_bytes,
                       cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (flags == nullptr || channel < 0 || channel >= CUEMBED_PEER_CHANNELS ||
      world < 1 || world > kMaxWorld)
    return CUEMBED_ERR_ARGUMENT;
  const unsigned* f = static_cast<const unsigned*>(flags);
  unsigned* status =
      const_cast<unsigned*>(f) + CUEMBED_PEER_CHANNELS * kMaxWorld;
  ShardWaitKernel<<<1, 256, 0, stream>>>(
      f + channel * kMaxWorld, world, epoch, status, WaitTimeoutNs(),
      static_cast<unsigned char*>(poison), poison_bytes);
  CountLaunch();
  return CudaRc();
}

int cuembed_shard_reduce_finalize(const void* slots, int partial_dtype,
                                  int world, const void* flags, int channel,
                                  unsigned epoch, int n_samples,
                                  int embed_width, int mode,
                                  const void* offsets, int off_type,
                                  int num_hots, int sample0,
                                  const void* weights, int weight_dtype,
                                  void* out, int out_dtype,
                                  cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (n_samples < 0 || embed_width <= 0 || world < 1 || world > kMaxWorld ||
      channel < 0 || channel >= CUEMBED_PEER_CHANNELS)
    return CUEMBED_ERR_ARGUMENT;
  if (mode != CUEMBED_SUM && mode != CUEMBED_MEAN) return CUEMBED_ERR_DTYPE;
  if (out_dtype < 0 || out_dtype > 2 || partial_dtype < 0 || partial_dtype > 2)
    return CUEMBED_ERR_DTYPE;
  if (slots == nullptr || flags == nullptr || (n_samples > 0 && out == nullptr))
    return CUEMBED_ERR_ARGUMENT;
  const unsigned* f = static_cast<const unsigned*>(flags);
  unsigned* status =
      const_cast<unsigned*>(f) + CUEMBED_PEER_CHANNELS * kMaxWorld;
  f += channel * kMaxWorld;
  const int mean = mode == CUEMBED_MEAN;
  const int off64 = off_type == CUEMBED_I64;
  if (partial_dtype == CUEMBED_F32)
    LaunchReduceW<float>(slots, world, f, epoch, status, n_samples, embed_width,
                         mean, offsets, off64, num_hots, sample0, weights,
                         weight_dtype, out, out_dtype, stream);
  else if (partial_dtype == CUEMBED_F16)
    LaunchReduceW<__half>(slots, world, f, epoch, status, n_samples,
                          embed_width, mean, offsets, off64, num_hots, sample0,
                          weights, weight_dtype, out, out_dtype, stream);
  else
    LaunchReduceW<__nv_bfloat16>(slots, world, f, epoch, status, n_samples,
                                 embed_width, mean, offsets, off64, num_hots,
                                 sample0, weights, weight_dtype, out, out_dtype,
                                 stream);
  return CudaRc();
}

// ------------------------------------------------------------- backward

int cuembed_shard_allgather_push(const void* src, size_t bytes,
                                 void* const* gather_ptrs, int world, int rank,
                                 cuembed_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world ||
      gather_ptrs == nullptr)
    return CUEMBED_ERR_ARGUMENT;
  if (bytes == 0) return CUEMBED_OK;
  if (src == nullptr) return CUEMBED_ERR_ARGUMENT;
  for (int k = 1; k <= world; ++k) {
    const int o = (rank + k) % world;  // own copy last
    if (gather_ptrs[o] == nullptr) return CUEMBED_ERR_ARGUMENT;
    char* dst = static_cast<char*>(gather_ptrs[o]) + static_cast<size_t>(rank) * bytes;
    if (dst == src) continue;
    if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream) !=
        cudaSuccess)
      return CUEMBED_ERR_CUDA;
  }
  return CUEMBED_OK;
}

}  // extern "C"
