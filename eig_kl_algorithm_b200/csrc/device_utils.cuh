// device_utils.cuh -- small device helpers shared by the kernels (sm_100a)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace eigkl {

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
// max of a 64-bit key over the warp, all lanes get it: two redux.sync (hardware warp reductions) -- the high
// words first, then the low words of the lanes that hold the winning high word -- instead of a five-round
// butterfly of 64-bit shuffles (10 SHFL + 20 ALU on the KL swap loop's critical path)
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mhi = __reduce_max_sync(FULL_MASK, hi);
  const unsigned mlo = __reduce_max_sync(FULL_MASK, hi == mhi ? lo : 0u);
  return ((unsigned long long)mhi << 32) | mlo;
}

// monotone float -> uint32 map (a < b  <=>  ord(a) < ord(b)); -0.0f is folded onto +0.0f first so
// that the two zeros tie exactly as they do under the reference's float comparisons
__device__ __forceinline__ uint32_t float_orderable(float f) {
  f = __fadd_rn(f, 0.0f);
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long double_orderable(double d) {
  unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double double_from_orderable(unsigned long long u) {
  u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
  return __longlong_as_double((long long)u);
}

}  // namespace eigkl
