// Shared helpers for the hugs_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hugs_b200.h"

#ifndef HUGS_API
#define HUGS_API extern "C" __attribute__((visibility("default")))
#endif

namespace hugs {

constexpr float kF32Eps = 1.1920928955078125e-07f;          // jnp.finfo(jnp.float32).eps
constexpr float kF32EpsSq = kF32Eps * kF32Eps;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);

int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define HUGS_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::hugs::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define HUGS_REQUIRE(cond, ...)                                                \
  do {                                                                         \
    if (!(cond)) { ::hugs::set_error(__VA_ARGS__); return HUGS_ERR_INVALID; }  \
  } while (0)

extern long long g_launch_count;
#define HUGS_LAUNCH_CHECK()                 \
  do {                                      \
    ++::hugs::g_launch_count;               \
    HUGS_CUDA(cudaGetLastError());          \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// Canonical in-place inclusive cumsum over a[0..n) held in shared memory by one warp.
// Order (fixed, so the oracle can mirror it): lane L owns the contiguous chunk
// [L*c, min((L+1)*c, n)), c = ceil(n/32), summed left to right; chunk totals are then
// accumulated left to right across lanes; result = lane_offset + local_prefix.
__device__ __forceinline__ void warp_cumsum_inplace(float* a, int n, int lane) {
  const int c = (n + 31) >> 5;
  const int b = min(lane * c, n), e = min(b + c, n);
  float s = 0.f;
  for (int i = b; i < e; ++i) { s = s + a[i]; a[i] = s; }
  float off = 0.f;
  for (int j = 0; j < 31; ++j) {
    float v = __shfl_sync(kFull, s, j);
    if (j < lane) off = off + v;
  }
  for (int i = b; i < e; ++i) a[i] = off + a[i];
  __syncwarp();
}

}  // namespace hugs
