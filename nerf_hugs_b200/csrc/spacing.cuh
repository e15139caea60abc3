// Normalised ray distance s in [0, 1] -> metric distance t (MipNeRF360/internal/coord.py:63-99 construct_ray_warps;
// nerfacto/models/nerf.py:209-226 and nerfacto.py:231-248 are the same maps under the names uniform / reciprocal / piecewise).
#pragma once
#include "common.cuh"

namespace hugs {

__device__ __forceinline__ float s_to_t(int fn, float s, float near, float far) {
  // coord.py:96-98: fn_inv(s * s_far + (1 - s) * s_near)
  switch (fn) {
    case HUGS_RAYDIST_RECIPROCAL: {
      float sn = 1.0f / near, sf = 1.0f / far;
      return 1.0f / (s * sf + (1.0f - s) * sn);
    }
    case HUGS_RAYDIST_LOG: {
      float sn = logf(near), sf = logf(far);
      return expf(s * sf + (1.0f - s) * sn);
    }
    case HUGS_RAYDIST_PIECEWISE: {
      float sn = near < 1.f ? .5f * near : 1.f - .5f / near;
      float sf = far < 1.f ? .5f * far : 1.f - .5f / far;
      float x = s * sf + (1.0f - s) * sn;
      return x < .5f ? 2.f * x : .5f / (1.f - x);
    }
    default:
      return s * far + (1.0f - s) * near;
  }
}

}  // namespace hugs
