// Alpha compositing, depth statistics, the three losses and their backward pass.
//
// One warp owns one ray (transmittance is a prefix scan along the ray; the warp keeps the
// whole ray in shared memory).  Compiled with -fmad=false (see sampling.cu).
//
// Reference semantics (paths under /root/reference/MipNeRF360/internal):
//   render.py:130-151   compute_alpha_weights
//   render.py:185-244   volumetric_rendering (+ stepfun.py:298-308 weighted_percentile)
//   train_utils.py:72-111   compute_data_loss (quirk B1 kept: see launch_lossmult_sum)
//   train_utils.py:228-248  interlevel_loss / distortion_loss
//   stepfun.py:30-86, 266-276  searchsorted / inner_outer / lossfun_outer / lossfun_distortion
// Backward formulas are derived in DESIGN.md §"Compositing backward".
#include "common.cuh"
#include "kernels.h"

namespace hugs {
namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ float softplus_f(float x) {  // jax.nn.softplus = logaddexp(x, 0)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// Fills X (density*delta), EX (exclusive cumsum of X), WT (alpha*trans) for one ray.  TMP is scratch [S].
__device__ __forceinline__ void alpha_weights(const float* __restrict__ raw_density, int raw_stride,
                                              const float* __restrict__ tdist, float dnorm, int S,
                                              int opaque, float density_bias, int lane,
                                              float* X, float* EX, float* WT) {
  for (int i = lane; i < S; i += 32) {
    float delta = (tdist[i + 1] - tdist[i]) * dnorm;
    float x = softplus_f(raw_density[(size_t)i * raw_stride] + density_bias) * delta;
    if (opaque && i == S - 1) x = INFINITY;
    X[i] = x;
    EX[i] = (i < S - 1) ? x : 0.f;
  }
  __syncwarp();
  warp_cumsum_inplace(EX, S - 1, lane);   // inclusive cumsum of X[0..S-2]
  // shift to exclusive: read before write, one barrier in between
  float keep[8];
  int cnt = 0;
  for (int i = lane; i < S; i += 32) keep[cnt++] = (i == 0) ? 0.f : EX[i - 1];
  __syncwarp();
  cnt = 0;
  for (int i = lane; i < S; i += 32) {
    float ex = keep[cnt++];
    EX[i] = ex;
    WT[i] = (1.0f - expf(-X[i])) * expf(-ex);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kWarps * 32) composite_kernel(CompositeArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int S = a.S;
  float* X = smem + warp * (4 * S + 8);
  float* EX = X + S;
  float* WT = EX + S;
  float* CW = WT + S;  // [S+2]
  const float* td = a.tdist + (size_t)ray * (S + 1);
  const float dx = a.directions[ray * 3], dy = a.directions[ray * 3 + 1], dz = a.directions[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  alpha_weights(a.raw_density + (size_t)ray * S * a.raw_stride, a.raw_stride, td, dnorm, S,
                a.opaque_background, a.density_bias, lane, X, EX, WT);

  float acc = 0.f, r = 0.f, g = 0.f, b = 0.f, lt = 0.f;
  for (int i = lane; i < S; i += 32) {
    float w = WT[i];
    acc += w;
    if (a.raw_rgb) {
      const float* c = a.raw_rgb + ((size_t)ray * S + i) * a.rgb_stride;
      float s = 1.f + 2.f * a.rgb_padding;
      float cr = sigmoid_f(a.rgb_premult * c[0] + a.rgb_bias) * s - a.rgb_padding;
      float cg = sigmoid_f(a.rgb_premult * c[1] + a.rgb_bias) * s - a.rgb_padding;
      float cb = sigmoid_f(a.rgb_premult * c[2] + a.rgb_bias) * s - a.rgb_padding;
      r += w * cr; g += w * cg; b += w * cb;
      if (a.out.rgbs) {
        float* o = a.out.rgbs + ((size_t)ray * S + i) * 3;
        o[0] = cr; o[1] = cg; o[2] = cb;
      }
    }
    if (a.compute_extras) lt += w * logf(0.5f * (td[i] + td[i + 1]));
    if (a.out.weights) a.out.weights[(size_t)ray * S + i] = w;
    if (a.out.density) a.out.density[(size_t)ray * S + i] =
        softplus_f(a.raw_density[((size_t)ray * S + i) * a.raw_stride] + a.density_bias);
  }
  acc = warp_sum(acc); r = warp_sum(r); g = warp_sum(g); b = warp_sum(b);
  const float bg_w = fmaxf(0.f, 1.f - acc);
  if (lane == 0) {
    if (a.out.rgb) {
      a.out.rgb[ray * 3 + 0] = r + bg_w * a.bg;
      a.out.rgb[ray * 3 + 1] = g + bg_w * a.bg;
      a.out.rgb[ray * 3 + 2] = b + bg_w * a.bg;
    }
    if (a.out.acc) a.out.acc[ray] = acc;
  }
  if (!a.compute_extras) return;
  lt = warp_sum(lt);
  if (lane == 0 && a.out.distance_mean) {
    float dm = expf(lt / fmaxf(kF32Eps, acc));
    if (dm != dm) dm = 0.f;   // jnp.nan_to_num(x, jnp.inf): inf lands in `copy`, NaN -> 0 (render.py:222)
    a.out.distance_mean[ray] = fminf(fmaxf(dm, td[0]), td[S]);
  }
  // weighted percentiles over (t ++ far, w ++ bg_w): cw = [0, min(1, cumsum(w_aug[:-1])), 1]
  for (int i = lane; i < S; i += 32) CW[i + 1] = WT[i];
  __syncwarp();
  warp_cumsum_inplace(CW + 1, S, lane);
  for (int i = lane; i < S; i += 32) CW[i + 1] = fminf(1.0f, CW[i + 1]);
  if (lane == 0) { CW[0] = 0.f; CW[S + 1] = 1.0f; }
  __syncwarp();
  if (lane < 3) {
    const float p = lane == 0 ? 0.05f : (lane == 1 ? 0.5f : 0.95f);
    const int n = S + 2;
    int lo = 0, hi = n;  // searchsorted(side='right')
    while (lo < hi) { int m = (lo + hi) >> 1; if (CW[m] <= p) lo = m + 1; else hi = m; }
    int i1 = min(max(lo, 1), n - 1), i0 = i1 - 1;
    auto tq = [&](int k) { return k <= S ? td[k] : a.far[ray]; };
    float dxp = CW[i1] - CW[i0];
    float f = (fabsf(dxp) <= 1.4210855e-14f) ? tq(i0)   /* jnp.interp: np.spacing(finfo(f32).eps) */ : tq(i0) + ((p - CW[i0]) / dxp) * (tq(i1) - tq(i0));
    float* dst = lane == 0 ? a.out.distance_p5 : (lane == 1 ? a.out.distance_median : a.out.distance_p95);
    if (dst) dst[ray] = f;
  }
}

// dL/dx_k from dL/dw (G): dL/dx_k = G_k e^{-x_k} T_k - sum_{i>k} G_i w_i.   H is scratch [S].
__device__ __forceinline__ void alpha_backward(const float* X, const float* EX, const float* WT,
                                               const float* G, float* H, int S, int lane, float* DX) {
  for (int i = lane; i < S; i += 32) H[i] = G[i] * WT[i];
  __syncwarp();
  warp_cumsum_inplace(H, S, lane);
  const float tot = H[S - 1];
  for (int i = lane; i < S; i += 32) {
    float suffix = tot - H[i];
    float x = X[i];
    float own = (x == INFINITY) ? 0.f : G[i] * expf(-x) * expf(-EX[i]);
    DX[i] = own - suffix;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kWarps * 32) final_loss_bwd_kernel(LossBwdArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int S = a.S;
  float* X = smem + warp * (8 * S);
  float* EX = X + S; float* WT = EX + S; float* G = WT + S; float* H = G + S; float* DX = H + S;
  float* PW = DX + S; float* PWU = PW + S;
  const float* td = a.tdist + (size_t)ray * (S + 1);
  const float* sd = a.sdist + (size_t)ray * (S + 1);
  const float* raw = a.raw + (size_t)ray * S * 4;
  const float dx = a.directions[ray * 3], dy = a.directions[ray * 3 + 1], dz = a.directions[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  alpha_weights(raw, 4, td, dnorm, S, a.opaque_background, a.density_bias, lane, X, EX, WT);

  // ---- rendered colour ---------------------------------------------------------------------
  const float cs = 1.f + 2.f * a.rgb_padding;
  float acc = 0.f, r = 0.f, g = 0.f, b = 0.f;
  for (int i = lane; i < S; i += 32) {
    float w = WT[i];
    acc += w;
    r += w * (sigmoid_f(a.rgb_premult * raw[i * 4 + 1] + a.rgb_bias) * cs - a.rgb_padding);
    g += w * (sigmoid_f(a.rgb_premult * raw[i * 4 + 2] + a.rgb_bias) * cs - a.rgb_padding);
    b += w * (sigmoid_f(a.rgb_premult * raw[i * 4 + 3] + a.rgb_bias) * cs - a.rgb_padding);
    if (a.weights) a.weights[(size_t)ray * S + i] = w;
  }
  acc = warp_sum(acc); r = warp_sum(r); g = warp_sum(g); b = warp_sum(b);
  const float bg_w = fmaxf(0.f, 1.f - acc);
  const float bg_on = (1.f - acc > 0.f) ? 1.f : 0.f;
  float rgb[3] = {r + bg_w * a.bg, g + bg_w * a.bg, b + bg_w * a.bg};

  // ---- data loss (train_utils.py:72-111) ---------------------------------------------------
  float lm;
  if (a.loss.use_static_mask) {
    float m = a.static_mask ? (a.static_mask[ray] >= 0.5f ? 1.f : 0.f) : 1.f;
    lm = m + (1.f - m) * a.loss.withmask_transient_weight;
  } else {
    lm = (a.loss.disable_multiscale_loss || !a.lossmult) ? 1.f : a.lossmult[ray];
  }
  const float denom = fmaxf(a.denom[0], kF32Eps);
  float gc[3], data_num = 0.f, sq_num = 0.f;
  for (int c = 0; c < 3; ++c) {
    float res = rgb[c] - a.rgb_gt[ray * 3 + c];
    float rs = res * res;
    sq_num += lm * rs;
    if (a.loss.data_loss_type == HUGS_LOSS_MSE) {
      data_num += lm * rs;
      gc[c] = a.loss.data_loss_mult * lm * 2.f * res / denom;
    } else {
      float ch = sqrtf(rs + a.loss.charb_padding * a.loss.charb_padding);
      data_num += lm * ch;
      gc[c] = a.loss.data_loss_mult * lm * (res / ch) / denom;
    }
  }

  // ---- distortion loss (stepfun.py:266-276) in O(S) with prefix sums -----------------------
  for (int i = lane; i < S; i += 32) {
    float u = (sd[i + 1] + sd[i]) / 2.f;
    PW[i] = WT[i];
    PWU[i] = WT[i] * u;
  }
  __syncwarp();
  warp_cumsum_inplace(PW, S, lane);
  warp_cumsum_inplace(PWU, S, lane);
  const float wtot = PW[S - 1], wutot = PWU[S - 1];
  const float sdist_scale = a.loss.distortion_loss_mult / (float)a.n_rays;
  float dist = 0.f;
  for (int i = lane; i < S; i += 32) {
    float w = WT[i];
    float u = (sd[i + 1] + sd[i]) / 2.f, dl = sd[i + 1] - sd[i];
    float wlt = PW[i] - w, wult = PWU[i] - w * u;
    float wgt = wtot - PW[i], wugt = wutot - PWU[i];
    dist += 2.f * w * (u * wlt - wult) + w * w * dl / 3.f;
    float dD = 2.f * (u * (wlt - wgt) - wult + wugt) + (2.f / 3.f) * w * dl;
    // dL/dw_i: data term through rgb = sum w c + max(0,1-acc) bg, plus distortion
    float cr = sigmoid_f(a.rgb_premult * raw[i * 4 + 1] + a.rgb_bias) * cs - a.rgb_padding;
    float cg = sigmoid_f(a.rgb_premult * raw[i * 4 + 2] + a.rgb_bias) * cs - a.rgb_padding;
    float cb = sigmoid_f(a.rgb_premult * raw[i * 4 + 3] + a.rgb_bias) * cs - a.rgb_padding;
    float bgt = bg_on * a.bg;
    G[i] = gc[0] * (cr - bgt) + gc[1] * (cg - bgt) + gc[2] * (cb - bgt) + sdist_scale * dD;
  }
  dist = warp_sum(dist);
  __syncwarp();
  alpha_backward(X, EX, WT, G, H, S, lane, DX);

  for (int i = lane; i < S; i += 32) {
    float delta = (td[i + 1] - td[i]) * dnorm;
    float pre = raw[i * 4] + a.density_bias;
    float dden = (a.opaque_background && i == S - 1) ? 0.f : DX[i] * delta;
    float4 o;
    o.x = dden * sigmoid_f(pre);
    float w = WT[i];
    float sr = sigmoid_f(a.rgb_premult * raw[i * 4 + 1] + a.rgb_bias);
    float sg = sigmoid_f(a.rgb_premult * raw[i * 4 + 2] + a.rgb_bias);
    float sb = sigmoid_f(a.rgb_premult * raw[i * 4 + 3] + a.rgb_bias);
    o.y = w * gc[0] * cs * sr * (1.f - sr) * a.rgb_premult;
    o.z = w * gc[1] * cs * sg * (1.f - sg) * a.rgb_premult;
    o.w = w * gc[2] * cs * sb * (1.f - sb) * a.rgb_premult;
    reinterpret_cast<float4*>(a.d_raw)[(size_t)ray * S + i] = o;
  }
  if (lane == 0 && a.ray_stats) {
    float* st = a.ray_stats + (size_t)ray * 4;
    st[0] = data_num; st[1] = sq_num; st[2] = dist; st[3] = 0.f;
  }
}

__global__ void __launch_bounds__(kWarps * 32) prop_loss_bwd_kernel(PropLossBwdArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int Sp = a.Sp, S = a.S;
  float* X = smem + warp * (7 * Sp + 1 + 3 * S);
  float* EX = X + Sp; float* WT = EX + Sp; float* G = WT + Sp; float* H = G + Sp; float* DX = H + Sp;
  float* CY = DX + Sp;                 // [Sp+1]
  float* HH = CY + Sp + 1;             // [S]
  int* LO = reinterpret_cast<int*>(HH + S);
  int* HI = LO + S;
  const float* td = a.tdist + (size_t)ray * (Sp + 1);
  const float* cp = a.sdist + (size_t)ray * (Sp + 1);
  const float* c = a.sdist_final + (size_t)ray * (S + 1);
  const float* wf = a.w_final + (size_t)ray * S;
  const float dx = a.directions[ray * 3], dy = a.directions[ray * 3 + 1], dz = a.directions[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  alpha_weights(a.raw_density + (size_t)ray * Sp, 1, td, dnorm, Sp, a.opaque_background,
                a.density_bias, lane, X, EX, WT);
  // cy = [0, cumsum(wp)]
  for (int i = lane; i < Sp; i += 32) CY[i + 1] = WT[i];
  __syncwarp();
  warp_cumsum_inplace(CY + 1, Sp, lane);
  if (lane == 0) CY[0] = 0.f;
  __syncwarp();
  // inner_outer (stepfun.py:64-77) outer measure + lossfun_outer (:80-86)
  float loss = 0.f;
  for (int i = lane; i < S; i += 32) {
    // idx_lo(c_i) = max{k : c_i >= cp_k} (0 if none); idx_hi(c_{i+1}) = min{k : c_{i+1} < cp_k} (Sp if none)
    float v0 = c[i], v1 = c[i + 1];
    int lo = 0, hi = Sp + 1;
    while (lo < hi) { int m = (lo + hi) >> 1; if (cp[m] <= v0) lo = m + 1; else hi = m; }
    int ilo = max(lo - 1, 0);
    lo = 0; hi = Sp + 1;
    while (lo < hi) { int m = (lo + hi) >> 1; if (cp[m] <= v1) lo = m + 1; else hi = m; }
    int ihi = min(lo, Sp);
    float wo = CY[ihi] - CY[ilo];
    float w = wf[i];
    float e = fmaxf(0.f, w - wo);
    loss += e * e / (w + kF32Eps);
    HH[i] = -2.f * e / (w + kF32Eps) * a.scale;
    LO[i] = ilo; HI[i] = ihi;
  }
  loss = warp_sum(loss);
  __syncwarp();
  for (int j = lane; j < Sp; j += 32) {
    float gsum = 0.f;
    for (int i = 0; i < S; ++i) if (LO[i] <= j && j < HI[i]) gsum += HH[i];
    G[j] = gsum;
  }
  __syncwarp();
  alpha_backward(X, EX, WT, G, H, Sp, lane, DX);
  for (int i = lane; i < Sp; i += 32) {
    float delta = (td[i + 1] - td[i]) * dnorm;
    float pre = a.raw_density[(size_t)ray * Sp + i] + a.density_bias;
    float dden = (a.opaque_background && i == Sp - 1) ? 0.f : DX[i] * delta;
    a.d_raw[(size_t)ray * Sp + i] = dden * sigmoid_f(pre);
  }
  if (lane == 0 && a.ray_stats) a.ray_stats[ray] = loss;
  if (a.sq_stats) {
    float acc = 0.f;
    for (int i = lane; i < Sp; i += 32) acc += WT[i];
    acc = warp_sum(acc);
    if (lane == 0) {
      float lm;
      if (a.loss.use_static_mask) {
        float m = a.static_mask ? (a.static_mask[ray] >= 0.5f ? 1.f : 0.f) : 1.f;
        lm = m + (1.f - m) * a.loss.withmask_transient_weight;
      } else {
        lm = (a.loss.disable_multiscale_loss || !a.lossmult) ? 1.f : a.lossmult[ray];
      }
      const float c = fmaxf(0.f, 1.f - acc) * a.bg;
      float sq = 0.f;
      for (int ch = 0; ch < 3; ++ch) { const float r = c - a.rgb_gt[ray * 3 + ch]; sq += lm * (r * r); }
      a.sq_stats[ray] = sq;
    }
  }
}

__global__ void lossmult_sum_kernel(const float* lossmult, const float* static_mask, int use_mask,
                                    float transient_w, int disable_multiscale, int n, float* out) {
  // single block, deterministic order
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v;
    if (use_mask) {
      float m = static_mask ? (static_mask[i] >= 0.5f ? 1.f : 0.f) : 1.f;
      v = m + (1.f - m) * transient_w;          // quirk B1: [n,1] -> counted once per ray
    } else {
      float l = (disable_multiscale || !lossmult) ? 1.f : lossmult[i];
      v = 3.f * l;                              // broadcast to [n,3] before the sum
    }
    s += v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[0] = v;
  }
}

__global__ void column_sums_kernel(const float* in, int n_rows, int stride, int n_cols, float* out) {
  __shared__ float red[32];
  const int col = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < n_rows; i += blockDim.x) s += in[(size_t)i * stride + col];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[col] = v;
  }
}

// the opt-in for > 48 KB of dynamic shared memory is a per-device attribute: one flag per device ordinal
template <class K>
int set_smem_once(K kernel, bool* flags) {
  int dev = 0;
  HUGS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !flags[dev]) {
    HUGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    if (dev >= 0 && dev < 64) flags[dev] = true;
  }
  return HUGS_OK;
}

}  // namespace

int launch_composite(const CompositeArgs& a, cudaStream_t stream) {
  HUGS_REQUIRE(a.S >= 2 && a.S <= 256, "composite: samples per ray must be in [2,256], got %d", a.S);
  static bool f[64] = {};
  int rc = set_smem_once(composite_kernel, f);
  if (rc) return rc;
  if (a.n_rays <= 0) return HUGS_OK;
  size_t smem = (size_t)kWarps * (4 * a.S + 8) * sizeof(float);
  composite_kernel<<<(a.n_rays + kWarps - 1) / kWarps, kWarps * 32, smem, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_final_loss_bwd(const LossBwdArgs& a, cudaStream_t stream) {
  HUGS_REQUIRE(a.S >= 2 && a.S <= 256, "loss: samples per ray must be in [2,256], got %d", a.S);
  static bool f[64] = {};
  int rc = set_smem_once(final_loss_bwd_kernel, f);
  if (rc) return rc;
  if (a.n_rays <= 0) return HUGS_OK;
  size_t smem = (size_t)kWarps * 8 * a.S * sizeof(float);
  final_loss_bwd_kernel<<<(a.n_rays + kWarps - 1) / kWarps, kWarps * 32, smem, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_prop_loss_bwd(const PropLossBwdArgs& a, cudaStream_t stream) {
  HUGS_REQUIRE(a.S >= 2 && a.S <= 256 && a.Sp >= 2 && a.Sp <= 256, "interlevel: samples per ray must be in [2,256]");
  static bool f[64] = {};
  int rc = set_smem_once(prop_loss_bwd_kernel, f);
  if (rc) return rc;
  if (a.n_rays <= 0) return HUGS_OK;
  size_t smem = (size_t)kWarps * (7 * a.Sp + 1 + 3 * a.S) * sizeof(float);
  prop_loss_bwd_kernel<<<(a.n_rays + kWarps - 1) / kWarps, kWarps * 32, smem, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_lossmult_sum(const float* lossmult, const float* static_mask, int use_mask, float transient_w,
                        int disable_multiscale, int n, float* out, cudaStream_t stream) {
  lossmult_sum_kernel<<<1, 1024, 0, stream>>>(lossmult, static_mask, use_mask, transient_w,
                                             disable_multiscale, n, out);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_column_sums(const float* in, int n_rows, int stride, int n_cols, float* out, cudaStream_t stream) {
  column_sums_kernel<<<n_cols, 1024, 0, stream>>>(in, n_rows, stride, n_cols, out);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs
