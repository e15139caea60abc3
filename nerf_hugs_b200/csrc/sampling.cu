// Hierarchical interval sampling: weight dilation + anneal/softmax/CDF + inverse-CDF resampling.
//
// One warp owns one ray; all per-ray arrays live in shared memory.  Compiled with -fmad=false so
// every a*b+c rounds twice like the CPU oracle (and like XLA:CPU).
//
// Reference semantics (paths under /root/reference/MipNeRF360/internal):
//   stepfun.py:89-128  weight_to_pdf / max_dilate / max_dilate_weights (+ [1:-1] trim, models.py:178-179)
//   models.py:182-193  anneal + logits
//   stepfun.py:131-161 integrate_weights / invert_cdf, math.py:108-127 sorted_interp
//   stepfun.py:214-263 sample_intervals
//   coord.py:63-99     s_to_t
#include "common.cuh"
#include "kernels.h"
#include "spacing.cuh"

namespace hugs {

namespace {

constexpr int kWarpsPerBlock = 4;

// number of elements x in a sorted sequence f(0..n) with f(i) < v (strict) or <= v
template <bool kStrict, class F>
__device__ __forceinline__ int count_before(F f, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    float x = f(mid);
    bool before = kStrict ? (x < v) : (x <= v);
    if (before) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One uniform draw in [0, 1) per (key, ray): splitmix64 finaliser, top 24 bits (the reference draws jax.random.uniform per
// level and ray, stepfun.py:206-209; threefry is not reproducible without JAX, so the stream is this library's own).
__device__ __forceinline__ float jitter_draw(uint64_t key, int ray) {
  uint64_t z = key + 0x9E3779B97F4A7C15ull * (uint64_t)(ray + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) resample_kernel(ResampleArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= a.n_rays) return;
  const int np = a.np, ns = a.ns;
  const int per_warp = (np + 1) + np + (3 * np + 1) + 3 * np + (3 * np + 1) + ns;
  float* T = smem + warp * per_warp;   // [np+1]
  float* W = T + (np + 1);             // [np]
  float* TD = W + np;                  // [3np+1]
  float* WD = TD + (3 * np + 1);       // [3np]
  float* CW = WD + 3 * np;             // [3np+1]
  float* C = CW + (3 * np + 1);        // [ns]

  // ---- load -------------------------------------------------------------------------------
  if (a.t_in) {
    for (int i = lane; i <= np; i += 32) T[i] = a.t_in[(size_t)ray * (np + 1) + i];
    for (int i = lane; i < np; i += 32) W[i] = a.w_in[(size_t)ray * np + i];
  } else {  // level 0: a single interval [s_near, s_far] of weight 1 (models.py:145-150)
    if (lane == 0) { T[0] = a.dom_lo; T[1] = a.dom_hi; W[0] = 1.0f; }
  }
  __syncwarp();

  float* t_cur = T;
  float* w_cur = W;
  int nb = np;

  // ---- max_dilate_weights(renormalize=True) + trim -----------------------------------------
  if (a.dilate) {
    const float d = a.dilation;
    // weight_to_pdf
    for (int i = lane; i < np; i += 32) W[i] = W[i] / fmaxf(kF32EpsSq, T[i + 1] - T[i]);
    __syncwarp();
    auto f_t = [&](int i) { return T[i]; };            // np+1 values
    auto f_t0 = [&](int i) { return T[i] - d; };       // np values
    auto f_t1 = [&](int i) { return T[i + 1] + d; };   // np values
    // merge the three sorted lists (priority t0 < t < t1 on ties); jnp.sort returns values only.
    for (int i = lane; i < np; i += 32) {
      float v = f_t0(i);
      int r = i + count_before<true>(f_t, np + 1, v) + count_before<true>(f_t1, np, v);
      TD[r] = fminf(fmaxf(v, a.dom_lo), a.dom_hi);
    }
    for (int i = lane; i <= np; i += 32) {
      float v = f_t(i);
      int r = i + count_before<false>(f_t0, np, v) + count_before<true>(f_t1, np, v);
      TD[r] = fminf(fmaxf(v, a.dom_lo), a.dom_hi);
    }
    for (int i = lane; i < np; i += 32) {
      float v = f_t1(i);
      int r = i + count_before<false>(f_t0, np, v) + count_before<false>(f_t, np + 1, v);
      TD[r] = fminf(fmaxf(v, a.dom_lo), a.dom_hi);
    }
    __syncwarp();
    // max-pool: p_d[i] = max{p_j : t0_j <= td_i < t1_j}, then pdf_to_weight
    float part = 0.f;
    for (int i = lane; i < 3 * np; i += 32) {
      float x = TD[i];
      int jhi = count_before<false>(f_t0, np, x) - 1;   // last j with t0_j <= x
      int jlo = count_before<false>(f_t1, np, x);       // first j with t1_j > x
      float m = 0.f;
      for (int j = jlo; j <= jhi; ++j) m = fmaxf(m, W[j]);
      float wv = m * (TD[i + 1] - x);
      WD[i] = wv;
      part += wv;
    }
    float tot = warp_sum(part);
    float denom = fmaxf(kF32EpsSq, tot);
    for (int i = lane; i < 3 * np; i += 32) WD[i] = WD[i] / denom;
    __syncwarp();
    t_cur = TD + 1;
    w_cur = WD + 1;
    nb = 3 * np - 2;
    if (a.td_out) {
      for (int i = lane; i <= nb; i += 32) a.td_out[(size_t)ray * (nb + 1) + i] = t_cur[i];
      for (int i = lane; i < nb; i += 32) a.wd_out[(size_t)ray * nb + i] = w_cur[i];
    }
  }

  if (a.ns > 0) {
    // ---- logits -> softmax -> CDF --------------------------------------------------------
    if (!a.cw_in) {
      float mx = -INFINITY;
      for (int i = lane; i < nb; i += 32) {
        float l;
        if (a.w_is_logits) l = w_cur[i];
        else l = (t_cur[i + 1] > t_cur[i]) ? a.anneal * logf(w_cur[i] + a.padding) : -INFINITY;
        CW[i + 1] = l;
        mx = fmaxf(mx, l);
      }
      mx = warp_max(mx);
      if (a.torch_twin && mx == -INFINITY) {   // quirk B5 (ray_utils.py:143-144): weights_logit[all -inf rows] = 1
        for (int i = lane; i < nb; i += 32) CW[i + 1] = 1.0f;
        mx = 1.0f;
      }
      float part = 0.f;
      for (int i = lane; i < nb; i += 32) {
        float e = expf(CW[i + 1] - mx);
        CW[i + 1] = e;
        part += e;
      }
      float tot = warp_sum(part);
      for (int i = lane; i < nb; i += 32) CW[i + 1] = CW[i + 1] / tot;
      __syncwarp();
      // integrate_weights: cw = [0, min(1, cumsum(w[:-1])), 1]
      warp_cumsum_inplace(CW + 1, nb, lane);
      for (int i = lane; i < nb; i += 32) CW[i + 1] = fminf(1.0f, CW[i + 1]);
      if (lane == 0) { CW[0] = 0.f; CW[nb] = 1.0f; }
      __syncwarp();
    } else {
      for (int i = lane; i <= nb; i += 32) CW[i] = a.cw_in[(size_t)ray * (nb + 1) + i];
      __syncwarp();
    }
    // ---- sorted_interp -------------------------------------------------------------------
    const bool jittered = a.jitter != nullptr || a.jitter_key != 0;
    const float jit = a.jitter ? a.jitter[(size_t)ray * a.jitter_stride] * a.max_jitter
                               : (a.jitter_key ? jitter_draw(a.jitter_key, ray) * a.max_jitter : 0.f);
    for (int j = lane; j < ns; j += 32) {
      float u;
      if (a.u_in) u = a.u_in[(size_t)ray * ns + j];
      else if (a.jitter && a.jitter_stride > 1) u = a.u_base[j] + a.jitter[(size_t)ray * a.jitter_stride + j] * a.max_jitter;
      else u = jittered ? a.u_base[j] + jit : a.u_base[j];
      // i0 = max{k : u >= cw_k} (0 if none); i1 = min{k : u < cw_k} (nb if none)
      int cnt = count_before<false>([&](int k) { return CW[k]; }, nb + 1, u);  // #{cw_k <= u}
      int i0 = max(cnt - 1, 0), i1 = min(cnt, nb);
      float xp0 = CW[i0], xp1 = CW[i1], fp0 = t_cur[i0], fp1 = t_cur[i1];
      float off = (u - xp0) / (xp1 - xp0);
      if (off != off) off = 0.f;                     // nan_to_num(nan -> 0); +-inf clip below
      off = fminf(fmaxf(off, 0.f), 1.f);
      C[j] = fp0 + off * (fp1 - fp0);
      if (a.idx_out) a.idx_out[(size_t)ray * ns + j] = i0;
    }
    __syncwarp();
    if (a.centers_out) {
      for (int j = lane; j < ns; j += 32) a.centers_out[(size_t)ray * ns + j] = C[j];
    }
    if (a.s_out) {
      // ---- sample_intervals: midpoints, reflected + clamped end posts -------------------
      const float near = a.near ? a.near[ray] : 0.f, far = a.far ? a.far[ray] : 1.f;
      for (int j = lane; j <= ns; j += 32) {
        float s;
        if (j == 0) {
          float mid = (C[1] + C[0]) / 2.f;
          s = fmaxf(a.dom_lo, 2.f * C[0] - mid);
        } else if (j == ns) {
          float mid = (C[ns - 1] + C[ns - 2]) / 2.f;
          s = fminf(a.dom_hi, 2.f * C[ns - 1] - mid);
        } else {
          s = (C[j] + C[j - 1]) / 2.f;
        }
        a.s_out[(size_t)ray * (ns + 1) + j] = s;
        if (a.t_out) a.t_out[(size_t)ray * (ns + 1) + j] = s_to_t(a.raydist_fn, s, near, far);
      }
    }
  }
}

}  // namespace

int launch_resample(const ResampleArgs& a, cudaStream_t stream) {
  const int np = a.np, ns = a.ns;
  size_t per_warp = (size_t)((np + 1) + np + (3 * np + 1) + 3 * np + (3 * np + 1) + ns) * sizeof(float);
  size_t smem = per_warp * kWarpsPerBlock;
  HUGS_REQUIRE(smem <= 200 * 1024, "resample: %d bins / %d samples per ray exceed shared memory", np, ns);
  static bool attr_set[64] = {};   // per device ordinal: the shared-memory opt-in is a per-device attribute
  int dev = 0;
  HUGS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    HUGS_CUDA(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  if (a.n_rays <= 0) return HUGS_OK;
  int blocks = (a.n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  resample_kernel<<<blocks, kWarpsPerBlock * 32, smem, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs
