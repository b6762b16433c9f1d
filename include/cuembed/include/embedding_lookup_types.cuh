// Drop-in replacement for cuembed/include/embedding_lookup_types.cuh
// (NVIDIA/cuEmbed @ 90dd8436) -- public types only.
//
// The reference header mixes the public surface (CombineMode :29, GetElemT
// :576-586) with the vector structs / casts / atomics its own kernels use.
// The B200 kernels live behind a C ABI (include/cuembed_b200.h), so only the
// public part is kept, plus the scalar VecCast and `float * __half` helpers
// that the reference's CPU oracle and harness pull from this header
// (utils/include/embedding_lookup_cpu.hpp:75,81-89).
#ifndef CUEMBED_INCLUDE_EMBEDDING_LOOKUP_TYPES_CUH_
#define CUEMBED_INCLUDE_EMBEDDING_LOOKUP_TYPES_CUH_

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

namespace cuembed {

// Same enumerators, same order as the reference (:29).
enum class CombineMode { kSum, kMean, kConcat };

// Scalar casts used by host-side reference code; identity by default.
template <typename ToType, typename FromType>
__device__ __host__ __forceinline__ ToType VecCast(const FromType& value) {
  return value;
}
template <>
__device__ __host__ __forceinline__ __half VecCast<__half, float>(
    const float& value) {
  return __float2half(value);
}
template <>
__device__ __host__ __forceinline__ float VecCast<float, __half>(
    const __half& value) {
  return __half2float(value);
}
template <>
__device__ __host__ __forceinline__ __nv_bfloat16
VecCast<__nv_bfloat16, float>(const float& value) {
  return __float2bfloat16(value);
}
template <>
__device__ __host__ __forceinline__ float VecCast<float, __nv_bfloat16>(
    const __nv_bfloat16& value) {
  return __bfloat162float(value);
}

__device__ __host__ __forceinline__ float operator*(
    const float& lhs, const __half rhs) {  // NOLINT(runtime/references)
  return lhs * __half2float(rhs);
}
__device__ __host__ __forceinline__ float operator*(
    const float& lhs, const __nv_bfloat16 rhs) {  // NOLINT(runtime/references)
  return lhs * __bfloat162float(rhs);
}

// Customisation point: element type of a (possibly structured) InputT; the
// weights of EmbeddingForward are `const GetElemT<InputT>*` (:576-586).
template <typename T>
struct GetElemType {
  using Type = T;
};
template <typename T>
using GetElemT = typename GetElemType<T>::Type;

// A structured InputT (new): a table seen through an addresser indirection --
// what the reference's comments reserve for an embedding cache ("float with
// cache", :576-578; "templatize the addresser with the cache",
// embedding_lookup_kernels.cuh:114-115).  Lookup i reads row row_map[i] of
// `cache` (of `rows` itself when cache is null) if row_map[i] >= 0, else row i
// of `rows`.  EmbeddingForward<MappedTable<ElemT, IndexT>, ...> takes a HOST
// pointer to one such descriptor as `params`; weights stay `const ElemT*`
// through the GetElemType specialisation below.  Sum / mean, fp32 accumulation.
template <typename ElemT, typename IndexT>
struct MappedTable {
  const ElemT* rows;        // backing table [num_rows, width], device-accessible
  const IndexT* row_map;    // [num_rows], device
  const ElemT* cache;       // cache table [slots, width] or nullptr, device
};
template <typename ElemT, typename IndexT>
struct GetElemType<MappedTable<ElemT, IndexT>> {
  using Type = ElemT;
};
namespace b200_detail {
template <typename T>
struct IsMappedTable {
  static constexpr bool value = false;
};
template <typename ElemT, typename IndexT>
struct IsMappedTable<MappedTable<ElemT, IndexT>> {
  static constexpr bool value = true;
};
}  // namespace b200_detail

// Element / index type -> C ABI code (include/cuembed_b200.h).
namespace b200_detail {
template <typename T>
struct DTypeCode;
template <>
struct DTypeCode<float> {
  static constexpr int value = 0;
};
template <>
struct DTypeCode<__half> {
  static constexpr int value = 1;
};
template <>
struct DTypeCode<__nv_bfloat16> {
  static constexpr int value = 2;
};
template <typename T>
struct ITypeCode {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8,
                "index / offset types must be 32- or 64-bit integers");
  static constexpr int value = sizeof(T) == 8 ? 1 : 0;
};
}  // namespace b200_detail

}  // namespace cuembed

#endif  // CUEMBED_INCLUDE_EMBEDDING_LOOKUP_TYPES_CUH_
