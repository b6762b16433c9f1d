// common.cuh -- shared device/host helpers for the sm_100a embedding kernels.
//
// Everything in csrc/ is written for B200 (sm_100a) only: 32-lane warps, 16-byte
// vector loads through the read-only path, 148 SMs.  No tensor-core code lives
// here because no stage of the path is a dense contraction.
#ifndef CUEMBED_B200_CSRC_COMMON_CUH_
#define CUEMBED_B200_CSRC_COMMON_CUH_

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>

#include "../../include/cuembed_b200.h"

namespace cuembed_b200 {

constexpr int kCtaThreads = 256;
constexpr int kWarpsPerCta = kCtaThreads / 32;

// ------------------------------------------------------------------ host utils

struct DeviceInfo {
  int sm_count;
  int max_smem_optin;
};

constexpr int kMaxDevices = 64;

inline int CurrentDeviceSlot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError();
  return (dev < 0 || dev >= kMaxDevices) ? 0 : dev;
}

// Cached per-device properties.  Each slot is published with release /
// acquire ordering, so a thread that sees `ready` also sees the values.
inline const DeviceInfo& GetDeviceInfo() {
  static DeviceInfo info[kMaxDevices];
  static std::atomic<bool> ready[kMaxDevices];
  static std::mutex mu;
  const int dev = CurrentDeviceSlot();
  if (!ready[dev].load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(mu);
    if (!ready[dev].load(std::memory_order_relaxed)) {
      DeviceInfo d;
      d.sm_count = 0;
      d.max_smem_optin = 0;
      cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
      cudaDeviceGetAttribute(&d.max_smem_optin,
                             cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      // Workspace queries must also work where no device is visible (build
      // hosts): size them for the part this library is written for.
      if (d.sm_count <= 0) {
        d.sm_count = 148;
        cudaGetLastError();
      }
      info[dev] = d;
      ready[dev].store(true, std::memory_order_release);
    }
  }
  return info[dev];
}

// One int per device for values that are device / context properties (kernel
// occupancy, "function attribute already raised"): a process that drives
// several GPUs must not reuse what it learned on the first one.  0 = unset.
// Racing first calls on one device compute the same value (idempotent).
struct PerDeviceInt {
  std::atomic<int> v[kMaxDevices];
  PerDeviceInt() {
    for (auto& x : v) x.store(0, std::memory_order_relaxed);
  }
  int Get() const {
    return v[CurrentDeviceSlot()].load(std::memory_order_acquire);
  }
  void Set(int x) { v[CurrentDeviceSlot()].store(x, std::memory_order_release); }
};

// Environment-variable tuning knob (read once per name per process by callers
// that cache the result).
int EnvInt(const char* name, int default_value);

inline int CeilDiv(int64_t a, int64_t b) {
  return static_cast<int>((a + b - 1) / b);
}

inline int Pow2Ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

inline size_t ElemSize(int dtype) { return dtype == CUEMBED_F32 ? 4 : 2; }
inline size_t IndexSize(int itype) { return itype == CUEMBED_I64 ? 8 : 4; }
inline size_t AlignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Row -> (vector bytes, vectors per row, lanes per row).  A row is split into
// V-byte vectors (V = 16, 8 or 4); G = lanes cooperating on one row (power of
// two, <= 32).  Rows wider than 32 vectors are walked in column tiles.
// Prefers the widest vector that still leaves >= 8 lanes per row so that one
// index-load round feeds at least 8 row loads per lane group.
struct RowShape {
  int vec_bytes;
  int nvec;
  int lanes;      // G
  int col_tiles;  // ceil(nvec / G)
};

inline bool MakeRowShape(int embed_width, int dtype, RowShape* s) {
  const int64_t row_bytes = static_cast<int64_t>(embed_width) * ElemSize(dtype);
  if (embed_width <= 0 || row_bytes % 4 != 0) return false;
  int v = 16;
  while (row_bytes % v != 0) v /= 2;
  while (v > 4 && row_bytes / v < 8) v /= 2;
  s->vec_bytes = v;
  s->nvec = static_cast<int>(row_bytes / v);
  s->lanes = Pow2Ceil(s->nvec) < 32 ? Pow2Ceil(s->nvec) : 32;
  s->col_tiles = (s->nvec + s->lanes - 1) / s->lanes;
  return true;
}

// -------------------------------------------------------------- device helpers

#ifdef __CUDACC__

template <int BYTES>
struct VecBits;
template <>
struct VecBits<16> {
  using type = uint4;
};
template <>
struct VecBits<8> {
  using type = uint2;
};
template <>
struct VecBits<4> {
  using type = uint32_t;
};

// Read-only (non-coherent) vector load; the table / grad_y are never written
// by the kernel that reads them.
template <int BYTES>
__device__ __forceinline__ typename VecBits<BYTES>::type LdgVec(
    const void* p) {
  return __ldg(reinterpret_cast<const typename VecBits<BYTES>::type*>(p));
}

// Streaming (evict-first) store for outputs that are written once.
template <int BYTES>
__device__ __forceinline__ void StcsVec(void* p,
                                        typename VecBits<BYTES>::type v) {
  __stcs(reinterpret_cast<typename VecBits<BYTES>::type*>(p), v);
}

__device__ __forceinline__ void Unpack32(uint32_t w, uint32_t* out) {
  out[0] = w;
}
__device__ __forceinline__ void Unpack32(uint2 w, uint32_t* out) {
  out[0] = w.x;
  out[1] = w.y;
}
__device__ __forceinline__ void Unpack32(uint4 w, uint32_t* out) {
  out[0] = w.x;
  out[1] = w.y;
  out[2] = w.z;
  out[3] = w.w;
}
__device__ __forceinline__ void Pack32(const uint32_t* in, uint32_t* w) {
  *w = in[0];
}
__device__ __forceinline__ void Pack32(const uint32_t* in, uint2* w) {
  w->x = in[0];
  w->y = in[1];
}
__device__ __forceinline__ void Pack32(const uint32_t* in, uint4* w) {
  w->x = in[0];
  w->y = in[1];
  w->z = in[2];
  w->w = in[3];
}

// Element traits: NE elements of type T live in a V-byte vector, i.e. in
// V/4 32-bit words.
template <typename T>
struct Elem;

template <>
struct Elem<float> {
  static constexpr int kPerWord = 1;
  static constexpr int kCode = CUEMBED_F32;
  __device__ __forceinline__ static void WordToFloat(uint32_t w, float* f) {
    f[0] = __uint_as_float(w);
  }
  __device__ __forceinline__ static uint32_t FloatToWord(const float* f) {
    return __float_as_uint(f[0]);
  }
  __device__ __forceinline__ static float ToFloat(float v) { return v; }
};

template <>
struct Elem<__half> {
  static constexpr int kPerWord = 2;
  static constexpr int kCode = CUEMBED_F16;
  using Pair = __half2;
  __device__ __forceinline__ static void WordToFloat(uint32_t w, float* f) {
    __half2 h = *reinterpret_cast<__half2*>(&w);
    float2 v = __half22float2(h);
    f[0] = v.x;
    f[1] = v.y;
  }
  __device__ __forceinline__ static uint32_t FloatToWord(const float* f) {
    __half2 h = __floats2half2_rn(f[0], f[1]);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static float ToFloat(__half v) {
    return __half2float(v);
  }
};

template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kPerWord = 2;
  static constexpr int kCode = CUEMBED_BF16;
  using Pair = __nv_bfloat162;
  __device__ __forceinline__ static void WordToFloat(uint32_t w, float* f) {
    // bf16 -> f32 is a 16-bit shift.
    f[0] = __uint_as_float(w << 16);
    f[1] = __uint_as_float(w & 0xffff0000u);
  }
  __device__ __forceinline__ static uint32_t FloatToWord(const float* f) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __device__ __forceinline__ static float ToFloat(__nv_bfloat16 v) {
    return __bfloat162float(v);
  }
};

// acc0 += lo half, acc1 += hi half of a packed 16-bit pair, in fp32 with one
// rounding per add.  sm_100a has a mixed-precision add (SASS FHADD), so this is
// ONE instruction per element instead of convert + add; the result is
// bit-identical because widening to fp32 is exact.
template <typename T>
__device__ __forceinline__ void AddPairF32(uint32_t w, float& acc0, float& acc1);
template <>
__device__ __forceinline__ void AddPairF32<__half>(uint32_t w, float& acc0,
                                                   float& acc1) {
  asm("{.reg .f16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %2;\n\t"
      "add.rn.f32.f16 %0, lo, %0;\n\t"
      "add.rn.f32.f16 %1, hi, %1;}"
      : "+f"(acc0), "+f"(acc1)
      : "r"(w));
}
template <>
__device__ __forceinline__ void AddPairF32<__nv_bfloat16>(uint32_t w,
                                                          float& acc0,
                                                          float& acc1) {
  asm("{.reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %2;\n\t"
      "add.rn.f32.bf16 %0, lo, %0;\n\t"
      "add.rn.f32.bf16 %1, hi, %1;}"
      : "+f"(acc0), "+f"(acc1)
      : "r"(w));
}

// acc[NE] += the NE elements packed in the V-byte vector `v` (fp32 adds).
template <typename T, int V>
__device__ __forceinline__ void AccumulateVec(typename VecBits<V>::type v,
                                              float* acc) {
  constexpr int NW = V / 4;
  uint32_t w[NW];
  Unpack32(v, w);
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = __fadd_rn(acc[i], __uint_as_float(w[i]));
  } else {
#pragma unroll
    for (int i = 0; i < NW; ++i) AddPairF32<T>(w[i], acc[2 * i], acc[2 * i + 1]);
  }
}

// acc[NE] += weight * element, multiply and add rounded separately (matches the
// reference CPU loops for any weight value; no FMA contraction).
template <typename T, int V>
__device__ __forceinline__ void AccumulateVecWeighted(
    typename VecBits<V>::type v, float wf, float* acc) {
  constexpr int NW = V / 4;
  uint32_t w[NW];
  Unpack32(v, w);
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    float f[Elem<T>::kPerWord];
    Elem<T>::WordToFloat(w[i], f);
#pragma unroll
    for (int k = 0; k < Elem<T>::kPerWord; ++k) {
      float& a = acc[i * Elem<T>::kPerWord + k];
      a = __fadd_rn(a, __fmul_rn(f[k], wf));
    }
  }
}

// Convert NW accumulator words' worth of floats to an output vector of dtype
// `out_dt` (runtime) and store it.  `acc` holds NE floats where the INPUT type
// packs NE elements in V_IN bytes; the output vector has NE elements of the
// output type, i.e. NE * sizeof(out) bytes.
template <int NE>
__device__ __forceinline__ void StoreFloatsAs(void* out_row, int64_t elem_off,
                                              int out_dt, const float* acc) {
  if (out_dt == CUEMBED_F32) {
    float* o = reinterpret_cast<float*>(out_row) + elem_off;
    if constexpr (NE % 4 == 0) {
#pragma unroll
      for (int i = 0; i < NE; i += 4)
        __stcs(reinterpret_cast<float4*>(o + i),
               make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]));
    } else if constexpr (NE % 2 == 0) {
#pragma unroll
      for (int i = 0; i < NE; i += 2)
        __stcs(reinterpret_cast<float2*>(o + i),
               make_float2(acc[i], acc[i + 1]));
    } else {
#pragma unroll
      for (int i = 0; i < NE; ++i) __stcs(o + i, acc[i]);
    }
    return;
  }
  // 16-bit outputs.
  uint16_t* o = reinterpret_cast<uint16_t*>(out_row) + elem_off;
  if constexpr (NE == 1) {
    uint16_t b;
    if (out_dt == CUEMBED_F16) {
      __half h = __float2half_rn(acc[0]);
      b = *reinterpret_cast<uint16_t*>(&h);
    } else {
      __nv_bfloat16 h = __float2bfloat16_rn(acc[0]);
      b = *reinterpret_cast<uint16_t*>(&h);
    }
    o[0] = b;
  } else {
    uint32_t w[NE / 2];
#pragma unroll
    for (int i = 0; i < NE / 2; ++i) {
      w[i] = (out_dt == CUEMBED_F16)
                 ? Elem<__half>::FloatToWord(acc + 2 * i)
                 : Elem<__nv_bfloat16>::FloatToWord(acc + 2 * i);
    }
    if constexpr (NE == 8) {
      __stcs(reinterpret_cast<uint4*>(o), make_uint4(w[0], w[1], w[2], w[3]));
    } else if constexpr (NE == 4) {
      __stcs(reinterpret_cast<uint2*>(o), make_uint2(w[0], w[1]));
    } else {
      __stcs(reinterpret_cast<uint32_t*>(o), w[0]);
    }
  }
}

// Address of row `row` of a row-major array: base + row * pitch.  Indices are
// non-negative; with 32-bit indices and a 32-bit pitch this compiles to exactly
// ONE instruction (IMAD.WIDE.U32 with the 64-bit base as addend) PROVIDED that
// `base` is an opaque register pair (asm volatile("" : "+l"(base)) after the
// lane's column offset has been added) -- otherwise the compiler re-adds the
// kernel-parameter base with two more adds per row.  (An explicit PTX
// mad.wide.u32 is split by ptxas into a multiply and two adds; the plain C++
// expression is not.)
template <typename IdxT>
__device__ __forceinline__ const char* RowAddr(const char* base, IdxT row,
                                               uint32_t pitch) {
  if constexpr (sizeof(IdxT) == 4)
    return base + static_cast<uint64_t>(static_cast<uint32_t>(row)) * pitch;
  else
    return base + static_cast<uint64_t>(row) * pitch;
}
template <typename IdxT>
__device__ __forceinline__ char* RowAddr(char* base, IdxT row, uint32_t pitch) {
  return const_cast<char*>(
      RowAddr<IdxT>(static_cast<const char*>(base), row, pitch));
}

template <typename IdxT>
__device__ __forceinline__ IdxT ShflIdx(IdxT v, int src, int width) {
  if constexpr (sizeof(IdxT) == 8) {
    long long r = __shfl_sync(0xffffffffu, static_cast<long long>(v), src, width);
    return static_cast<IdxT>(r);
  } else {
    return __shfl_sync(0xffffffffu, v, src, width);
  }
}

__device__ __forceinline__ int64_t LoadOffset(const void* offsets, int off64,
                                              int64_t i) {
  return off64 ? __ldg(reinterpret_cast<const long long*>(offsets) + i)
               : static_cast<int64_t>(
                     __ldg(reinterpret_cast<const int*>(offsets) + i));
}

#endif  // __CUDACC__

}  // namespace cuembed_b200

#endif  // CUEMBED_B200_CSRC_COMMON_CUH_
