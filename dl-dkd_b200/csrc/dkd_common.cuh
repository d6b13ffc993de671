// Shared device helpers for the dkd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/dkd_b200.h"

#define DKD_CUDA_TRY(expr)                      \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

#define DKD_LAUNCH_CHECK()                      \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

namespace dkd {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// (value desc, index asc) "better" predicate: the torch.max / stable-rank tie rule.
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  return (v > bv) || (v == bv && i < bi);
}

// Proposal index of window (length w >= 1, start s) among T clips: SURVEY §8 N2.
__host__ __device__ __forceinline__ int prop_index(int w, int s, int T) {
  return (w - 1) * T - ((w - 1) * (w - 2)) / 2 + s;
}

// Monotone map float -> uint32 (larger float => larger key); -0.0 < +0.0 is harmless here.
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}
// 64-bit sort key: score desc then id asc  <=>  key desc.
__device__ __forceinline__ unsigned long long pack_key(float s, int id) {
  return ((unsigned long long)float_key(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)id);
}
__device__ __forceinline__ float key_score(unsigned long long k) { return key_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ int key_id(unsigned long long k) { return (int)(0xffffffffu - (uint32_t)(k & 0xffffffffu)); }

}  // namespace dkd
