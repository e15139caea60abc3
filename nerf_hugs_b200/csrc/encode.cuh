// Conical-frustum Gaussians, contraction push-forward and integrated positional encoding.
//
// Reference semantics (paths under /root/reference/MipNeRF360/internal):
//   render.py:44-78,81-100  conical_frustum_to_gaussian (stable) / cylinder_to_gaussian
//   render.py:21-41         lift_gaussian(diag=False):  cov = t_var d d^T + r_var (I - d d^T/|d|^2)
//   coord.py:21-27,39-60    contract + track_linearize: cov' = J cov J^T
//   coord.py:129-133        lift_and_diagonalize: mu_b = mean.p_b, var_b = p_b^T cov' p_b
//   coord.py:102-126        integrated_pos_enc, math.py:26-38 safe_sin
// The 3x3 covariance is never formed: var_b = t_var (d.v_b)^2 + r_var (|v_b|^2 - (d.v_b)^2/|d|^2)
// with v_b = J p_b (J symmetric; closed form in SURVEY.md App. A / DESIGN.md).
#pragma once
#include "common.cuh"

namespace hugs {

struct SampleGauss {
  float x[3];      // mean = o + d * t_mean (un-contracted)
  float z[3];      // contract(mean) (== x when not contracted)
  float t_var, r_var;
  float dd;        // max(1e-10, |d|^2)
  // contraction (valid when contract): z = sz * x ; J v = js * v + jc * (x.v) * x
  float sz, js, jc;
};

__device__ __forceinline__ void frustum_gaussian(const float o[3], const float d[3], float radius,
                                                 float t0, float t1, int ray_shape, int contract,
                                                 SampleGauss& g) {
  float t_mean, t_var, r_var;
  if (ray_shape == HUGS_RAY_CONE) {
    float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
    float denom = fmaxf(kF32Eps, 3.f * mu * mu + hw * hw);
    t_mean = mu + (2.f * mu * hw * hw) / denom;
    float hw4 = hw * hw * hw * hw;
    t_var = (hw * hw) / 3.f - (4.f / 15.f) * hw4 * (12.f * mu * mu - hw * hw) / (denom * denom);
    r_var = (mu * mu) / 4.f + (5.f / 12.f) * hw * hw - (4.f / 15.f) * hw4 / denom;
    r_var *= radius * radius;
  } else {
    t_mean = (t0 + t1) / 2.f;
    r_var = radius * radius / 4.f;
    t_var = (t1 - t0) * (t1 - t0) / 12.f;
  }
  g.t_var = t_var; g.r_var = r_var;
  g.dd = fmaxf(1e-10f, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  g.x[0] = o[0] + d[0] * t_mean; g.x[1] = o[1] + d[1] * t_mean; g.x[2] = o[2] + d[2] * t_mean;
  g.sz = 1.f; g.js = 1.f; g.jc = 0.f;
  g.z[0] = g.x[0]; g.z[1] = g.x[1]; g.z[2] = g.x[2];
  if (contract) {
    float m = fmaxf(kF32Eps, g.x[0] * g.x[0] + g.x[1] * g.x[1] + g.x[2] * g.x[2]);
    if (m > 1.f) {
      float sq = sqrtf(m);
      g.sz = (2.f * sq - 1.f) / m;
      g.js = g.sz;
      g.jc = 2.f * (1.f / (m * m) - 1.f / (m * sq));
      g.z[0] = g.sz * g.x[0]; g.z[1] = g.sz * g.x[1]; g.z[2] = g.sz * g.x[2];
    }
  }
}

// lifted mean / variance along basis direction p
__device__ __forceinline__ void lift_basis(const SampleGauss& g, const float d[3], const float p[3],
                                           float& mu_b, float& var_b) {
  float xp = g.x[0] * p[0] + g.x[1] * p[1] + g.x[2] * p[2];
  mu_b = g.z[0] * p[0] + g.z[1] * p[1] + g.z[2] * p[2];
  float v0 = g.js * p[0] + g.jc * xp * g.x[0];
  float v1 = g.js * p[1] + g.jc * xp * g.x[1];
  float v2 = g.js * p[2] + g.jc * xp * g.x[2];
  float dv = d[0] * v0 + d[1] * v1 + d[2] * v2;
  float vv = v0 * v0 + v1 * v1 + v2 * v2;
  var_b = g.t_var * dv * dv + g.r_var * (vv - dv * dv / g.dd);
}

__device__ __forceinline__ float safe_sin_ref(float x) {
  const float t = 314.159271240234375f;           // fp32(100*pi)
  if (!(fabsf(x) < t)) {
    float r = fmodf(x, t);
    if (r != 0.f && r < 0.f) r += t;              // Python/NumPy remainder: sign of the divisor
    x = r;
  }
  return sinf(x);
}

}  // namespace hugs
