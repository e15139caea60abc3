// Operators of the torch twins (nerfacto/): fencepost merge, compositing with its backward, point positional encoding,
// table-driven parameter copies.  One warp owns one ray.  Compiled with -fmad=false so that a*b+c rounds twice like the
// reference's eager torch kernels.
//
// Reference semantics (paths under /root/reference/nerfacto):
//   models/nerf.py:287-295                 fine-level fenceposts around the sorted union of coarse and fine centres
//   utils/ray_utils.py:226-249             density_to_weight (quirk B2: deltas from the first fencepost; B6: nan_to_num)
//   utils/ray_utils.py:295-312,336-346     render_features, render_depth (quirk B7: clip with the batch-wide maximum)
//   models/custom_functions.py:15-21,37-63 spatial_distortion_norm2, trunc_exp, pos_enc
//   models/nerf.py:682-698,832             density / rgb activations
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "spacing.cuh"

namespace hugs {
namespace {

constexpr int kWarps = 4;

// ------------------------------------------------------------------------------------------ merge
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

__global__ void __launch_bounds__(kWarps * 32) nf_merge_kernel(NfMergeArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int na = a.na, nb = a.nb, nm = na + nb;
  float* A = smem + warp * (2 * nm);   // centres of a [na], centres of b [nb]
  float* B = A + na;
  float* M = B + nb;                   // merged [nm]
  const float* ba = a.bins_a + (size_t)ray * (na + 1);
  const float* bb = a.bins_b + (size_t)ray * (nb + 1);
  for (int i = lane; i < na; i += 32) A[i] = (ba[i + 1] + ba[i]) / 2.f;
  for (int i = lane; i < nb; i += 32) B[i] = (bb[i + 1] + bb[i]) / 2.f;
  __syncwarp();
  // both lists are sorted: rank by counting (ties: a first); torch.sort returns the same values
  for (int i = lane; i < na; i += 32) M[i + count_before<true>([&](int k) { return B[k]; }, nb, A[i])] = A[i];
  for (int i = lane; i < nb; i += 32) M[i + count_before<false>([&](int k) { return A[k]; }, na, B[i])] = B[i];
  __syncwarp();
  const float near = a.near ? a.near[ray] : 0.f, far = a.far ? a.far[ray] : 1.f;
  for (int j = lane; j <= nm; j += 32) {
    float s;
    if (j == 0) s = fmaxf(a.dom_lo, 2.f * M[0] - (M[1] + M[0]) / 2.f);
    else if (j == nm) s = fminf(a.dom_hi, 2.f * M[nm - 1] - (M[nm - 1] + M[nm - 2]) / 2.f);
    else s = (M[j] + M[j - 1]) / 2.f;
    a.bins_out[(size_t)ray * (nm + 1) + j] = s;
    if (a.t_out) a.t_out[(size_t)ray * (nm + 1) + j] = s_to_t(a.spacing_fn, s, near, far);
  }
}

// ------------------------------------------------------------------------------------------ compositing
__device__ __forceinline__ float sigmoid_t(float x) { return 1.0f / (1.0f + expf(-x)); }

// density activation and its derivative (F.softplus: beta 1, threshold 20; trunc_exp: backward clamps the exponent)
__device__ __forceinline__ float density_act(int act, float x) {
  if (act == HUGS_DENSITY_TRUNC_EXP) return expf(x);
  if (act == HUGS_DENSITY_RELU) return fmaxf(x, 0.f);
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float density_act_grad(int act, float x) {
  if (act == HUGS_DENSITY_TRUNC_EXP) return expf(fminf(fmaxf(x, -15.f), 15.f));
  if (act == HUGS_DENSITY_RELU) return x > 0.f ? 1.f : 0.f;
  if (x > 20.f) return 1.f;
  const float z = expf(x);
  return z / (z + 1.f);
}

// X = density * delta (delta from the FIRST fencepost, quirk B2), EX = inclusive cumsum of X[0..S-2],
// WT = nan_to_num((1 - e^-X) * e^-excl) for one ray
__device__ __forceinline__ void nf_alpha_weights(const NfCompositeArgs& a, const float* raw, const float* td, float dnorm,
                                                 int lane, float* X, float* EX, float* WT) {
  const int S = a.S;
  for (int i = lane; i < S; i += 32) {
    const float delta = (td[i + 1] - td[0]) * dnorm;
    float x = density_act(a.cfg.density_activation, raw[(size_t)i * a.C] + a.cfg.density_bias) * delta;
    if (a.cfg.opaque_background && i == S - 1) x = INFINITY;
    X[i] = x;
    EX[i] = x;
  }
  __syncwarp();
  warp_cumsum_inplace(EX, S - 1, lane);
  for (int i = lane; i < S; i += 32) {
    const float ex = i == 0 ? 0.f : EX[i - 1];
    float w = (1.0f - expf(-X[i])) * expf(-ex);
    if (w != w) w = 0.f;                      // torch.nan_to_num (quirk B6)
    WT[i] = w;
  }
  __syncwarp();
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <bool kBackward>
__global__ void __launch_bounds__(kWarps * 32) nf_composite_kernel(NfCompositeArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int S = a.S, C = a.C;
  float* X = smem + warp * (kBackward ? 5 : 3) * S;
  float* EX = X + S; float* WT = EX + S;
  const float* td = a.tdist + (size_t)ray * (S + 1);
  const float* raw = a.raw + (size_t)ray * S * C;
  const float dx = a.directions[ray * 3], dy = a.directions[ray * 3 + 1], dz = a.directions[ray * 3 + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);     // torch.linalg.norm
  nf_alpha_weights(a, raw, td, dnorm, lane, X, EX, WT);

  const float cs = 1.f + 2.f * a.cfg.rgb_padding;
  auto colour = [&](float r) { return sigmoid_t(a.cfg.rgb_premultiplier * r + a.cfg.rgb_bias) * cs - a.cfg.rgb_padding; };
  float acc = 0.f, r = 0.f, g = 0.f, b = 0.f, num = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float w = WT[i];
    acc += w;
    num += w * ((td[i + 1] + td[i]) / 2.f);
    if (C == 4) { r += w * colour(raw[i * 4 + 1]); g += w * colour(raw[i * 4 + 2]); b += w * colour(raw[i * 4 + 3]); }
    if (!kBackward && a.weights) a.weights[(size_t)ray * S + i] = w;
  }
  acc = warp_sum(acc); num = warp_sum(num); r = warp_sum(r); g = warp_sum(g); b = warp_sum(b);
  const float acc_safe = acc > 0.f ? acc : kF32Eps;
  const float depth = num / acc_safe;
  if (!kBackward) {
    if (lane == 0) {
      const float bg_acc = fmaxf(1.f - acc, 0.f);
      if (a.rgb) {
        float bgc[3] = {0.f, 0.f, 0.f};
        if (a.bg_rgb) { bgc[0] = a.bg_rgb[ray * 3]; bgc[1] = a.bg_rgb[ray * 3 + 1]; bgc[2] = a.bg_rgb[ray * 3 + 2]; }
        a.rgb[ray * 3 + 0] = a.bg_rgb ? r + bgc[0] * bg_acc : r;
        a.rgb[ray * 3 + 1] = a.bg_rgb ? g + bgc[1] * bg_acc : g;
        a.rgb[ray * 3 + 2] = a.bg_rgb ? b + bgc[2] * bg_acc : b;
      }
      if (a.acc) a.acc[ray] = acc;
      if (a.depth) a.depth[ray] = depth;
      if (a.steps_max) atomic_max_float(a.steps_max, (td[S] + td[S - 1]) / 2.f);
    }
    return;
  }
  // ---- backward: G = dL/dw, then dL/dx_k = G_k e^{-x_k} T_k - sum_{i>k} G_i w_i ------------------------------
  float* G = WT + S; float* H = G + S;
  float drgb[3] = {0.f, 0.f, 0.f};
  if (a.d_rgb) { drgb[0] = a.d_rgb[ray * 3]; drgb[1] = a.d_rgb[ray * 3 + 1]; drgb[2] = a.d_rgb[ray * 3 + 2]; }
  float bgterm = 0.f;    // d rgb / d acc through the background: -bg * [1 - acc >= 0]  (torch.clamp_min backward)
  if (a.bg_rgb && (1.f - acc) >= 0.f)
    bgterm = drgb[0] * a.bg_rgb[ray * 3] + drgb[1] * a.bg_rgb[ray * 3 + 1] + drgb[2] * a.bg_rgb[ray * 3 + 2];
  float dd = a.d_depth ? a.d_depth[ray] : 0.f;
  if (a.steps_max_in && !(depth >= 0.f && depth <= a.steps_max_in[0])) dd = 0.f;    // torch.clip backward
  const float dacc = (a.d_acc ? a.d_acc[ray] : 0.f) - bgterm - (acc > 0.f ? dd * num / (acc_safe * acc_safe) : 0.f);
  for (int i = lane; i < S; i += 32) {
    float gi = dacc + dd * ((td[i + 1] + td[i]) / 2.f) / acc_safe;
    if (a.d_weights) gi += a.d_weights[(size_t)ray * S + i];
    if (C == 4) gi += drgb[0] * colour(raw[i * 4 + 1]) + drgb[1] * colour(raw[i * 4 + 2]) + drgb[2] * colour(raw[i * 4 + 3]);
    G[i] = gi;
    H[i] = gi * WT[i];
  }
  __syncwarp();
  warp_cumsum_inplace(H, S, lane);
  const float tot = H[S - 1];
  for (int i = lane; i < S; i += 32) {
    const float x = X[i];
    const float ex = i == 0 ? 0.f : EX[i - 1];
    const float own = (x == INFINITY) ? 0.f : G[i] * expf(-x) * expf(-ex);
    const float dxk = own - (tot - H[i]);
    const float delta = (td[i + 1] - td[0]) * dnorm;
    const float pre = raw[(size_t)i * C] + a.cfg.density_bias;
    float dden = (a.cfg.opaque_background && i == S - 1) ? 0.f : dxk * delta;
    float draw = dden * density_act_grad(a.cfg.density_activation, pre);
    if (draw != draw) draw = 0.f;
    if (C == 4) {
      const float w = WT[i];
      float4 o;
      o.x = draw;
      const float sr = sigmoid_t(a.cfg.rgb_premultiplier * raw[i * 4 + 1] + a.cfg.rgb_bias);
      const float sg = sigmoid_t(a.cfg.rgb_premultiplier * raw[i * 4 + 2] + a.cfg.rgb_bias);
      const float sb = sigmoid_t(a.cfg.rgb_premultiplier * raw[i * 4 + 3] + a.cfg.rgb_bias);
      o.y = w * drgb[0] * cs * sr * (1.f - sr) * a.cfg.rgb_premultiplier;
      o.z = w * drgb[1] * cs * sg * (1.f - sg) * a.cfg.rgb_premultiplier;
      o.w = w * drgb[2] * cs * sb * (1.f - sb) * a.cfg.rgb_premultiplier;
      reinterpret_cast<float4*>(a.d_raw)[(size_t)ray * S + i] = o;
    } else {
      a.d_raw[(size_t)ray * S + i] = draw;
    }
  }
}

__global__ void nf_clip_depth_kernel(float* depth, const float* steps_max, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) depth[i] = fminf(fmaxf(depth[i], 0.f), steps_max[0]);
}

// ------------------------------------------------------------------------------------------ photometric loss
__global__ void __launch_bounds__(256) nf_rgb_loss_kernel(const float* pred, const float* gt, const float* mask,
                                                          float transient_w, int loss_type, float padding, int n,
                                                          float* sums, float* dl) {
  __shared__ float red[3][8];
  float sl = 0.f, se = 0.f, sm = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * 3; i += gridDim.x * blockDim.x) {
    const int ray = i / 3;
    float m = 1.f;
    if (mask) { const float st = mask[ray] >= 0.5f ? 1.f : 0.f; m = st + (1.f - st) * transient_w; }
    const float r = pred[i] - gt[i];
    const float e = r * r;
    float l, g;
    if (loss_type == HUGS_LOSS_MSE) { l = e; g = 2.f * r; }
    else { l = sqrtf(e + padding * padding); g = r / l; }
    sl += m * l; se += m * e; sm += m;
    dl[i] = m * g;
  }
  sl = warp_sum(sl); se = warp_sum(se); sm = warp_sum(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = sl; red[1][warp] = se; red[2][warp] = sm; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, v);
  }
}

__global__ void nf_rgb_loss_bwd_kernel(const float* dl, const float* sums, const float* upstream, float scale, int n,
                                       float* d_pred) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  d_pred[i] = upstream[0] * scale / fmaxf(sums[2], kF32Eps) * dl[i];
}

// ------------------------------------------------------------------------------------------ proposal losses
// utils/loss_utils.py:48-84.  One warp per ray; value per ray summed into out[0]; gradients w.r.t. the weights written
// unscaled (the mean's 1 / count and the upstream gradient are applied by nf_scale_kernel in the autograd backward).
struct NfLossArgs {
  const float* c; const float* w; int S;        // final level: spacing fenceposts [n, S+1], weights [n, S]
  const float* cp; const float* wp; int Sp;     // proposal level (interlevel only)
  int n_rays;
  float* out;                                   // [1] sum over rays and samples
  float* grad;                                  // distortion: [n, S] dL/dw; interlevel: [n, Sp] dL/dwp
};

// lossfun_distortion (loss_utils.py:66-77) in O(S): the interval midpoints are sorted, so
// sum_ij w_i w_j |u_i - u_j| = 2 sum_i w_i (u_i W_<i - WU_<i) with exclusive prefix sums
__global__ void __launch_bounds__(kWarps * 32) nf_distortion_kernel(NfLossArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int S = a.S;
  float* PW = smem + warp * 2 * S; float* PWU = PW + S;
  const float* c = a.c + (size_t)ray * (S + 1);
  const float* w = a.w + (size_t)ray * S;
  for (int i = lane; i < S; i += 32) { const float u = (c[i + 1] + c[i]) / 2.f; PW[i] = w[i]; PWU[i] = w[i] * u; }
  __syncwarp();
  warp_cumsum_inplace(PW, S, lane);
  warp_cumsum_inplace(PWU, S, lane);
  const float wtot = PW[S - 1], wutot = PWU[S - 1];
  float loss = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float wi = w[i], u = (c[i + 1] + c[i]) / 2.f, dl = c[i + 1] - c[i];
    const float wlt = PW[i] - wi, wult = PWU[i] - wi * u, wgt = wtot - PW[i], wugt = wutot - PWU[i];
    loss += 2.f * wi * (u * wlt - wult) + wi * wi * dl / 3.f;
    a.grad[(size_t)ray * S + i] = 2.f * (u * (wlt - wgt) - wult + wugt) + (2.f / 3.f) * wi * dl;
  }
  loss = warp_sum(loss);
  if (lane == 0) atomicAdd(a.out, loss);
}

// lossfun_outer / outer (loss_utils.py:7-45): w_outer_i = sum of wp over [idx_lo_i, idx_hi_i] with
// idx_lo = clamp(searchsorted(cp[:-1], c_i, right) - 1, 0, Sp - 1), idx_hi = clamp(searchsorted(cp[1:], c_{i+1}, right), 0, Sp - 1)
__global__ void __launch_bounds__(kWarps * 32) nf_interlevel_kernel(NfLossArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + warp;
  if (ray >= a.n_rays) return;
  const int S = a.S, Sp = a.Sp;
  float* CY = smem + warp * (2 * Sp + 2);      // [Sp+1] = [0, cumsum(wp)]
  float* D = CY + Sp + 1;                       // [Sp+1] difference array of the gradient
  const float* c = a.c + (size_t)ray * (S + 1);
  const float* w = a.w + (size_t)ray * S;
  const float* cp = a.cp + (size_t)ray * (Sp + 1);
  const float* wp = a.wp + (size_t)ray * Sp;
  for (int j = lane; j < Sp; j += 32) { CY[j + 1] = wp[j]; D[j] = 0.f; }
  if (lane == 0) { CY[0] = 0.f; D[Sp] = 0.f; }
  __syncwarp();
  warp_cumsum_inplace(CY + 1, Sp, lane);
  float loss = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float v0 = c[i], v1 = c[i + 1];
    int lo = 0, hi = Sp;                       // #{k < Sp : cp[k] <= v0}
    while (lo < hi) { const int m = (lo + hi) >> 1; if (cp[m] <= v0) lo = m + 1; else hi = m; }
    const int ilo = min(max(lo - 1, 0), Sp - 1);
    lo = 0; hi = Sp;                           // #{k < Sp : cp[k + 1] <= v1}
    while (lo < hi) { const int m = (lo + hi) >> 1; if (cp[m + 1] <= v1) lo = m + 1; else hi = m; }
    const int ihi = min(max(lo, 0), Sp - 1);
    const float wo = CY[ihi + 1] - CY[ilo];
    const float wi = w[i];
    const float e = fmaxf(wi - wo, 0.f);
    loss += e * e / (wi + 1.0e-7f);
    const float hgrad = -2.f * e / (wi + 1.0e-7f);
    if (hgrad != 0.f && ihi >= ilo) { atomicAdd(&D[ilo], hgrad); atomicAdd(&D[ihi + 1], -hgrad); }
  }
  loss = warp_sum(loss);
  __syncwarp();
  warp_cumsum_inplace(D, Sp, lane);
  for (int j = lane; j < Sp; j += 32) a.grad[(size_t)ray * Sp + j] = D[j];
  if (lane == 0) atomicAdd(a.out, loss);
}

__global__ void nf_scale_kernel(const float* src, const float* upstream, float mult, long long n, float* dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * (upstream[0] * mult);
}

// ------------------------------------------------------------------------------------------ parameter copies
__global__ void params_copy_kernel(const hugs_tensor_copy* table, float* flat, int direction, float* base) {
  hugs_tensor_copy t = table[blockIdx.y];
  if (base) t.ptr = reinterpret_cast<float*>(reinterpret_cast<char*>(base) + reinterpret_cast<size_t>(t.ptr));
  const long long n = (long long)t.rows * t.cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / t.cols), j = (int)(e % t.cols);
    const long long q = t.transpose ? (long long)j * (t.ld > 0 ? t.ld : t.rows) + i : (long long)i * (t.ld > 0 ? t.ld : t.cols) + j;
    if (direction == 0) flat[t.flat_off + e] = t.ptr[q];
    else t.ptr[q] = flat[t.flat_off + e];
  }
}

// ------------------------------------------------------------------------------------------ point positional encoding
__device__ __forceinline__ void sample_point(const PointPeArgs& a, int s, float x[3]) {
  const int ray = s / a.S, i = s % a.S;
  const float t0 = a.tdist[(size_t)ray * (a.S + 1) + i], t1 = a.tdist[(size_t)ray * (a.S + 1) + i + 1];
  const float tm = (t1 + t0) / 2.f;                                       // nerf.py:299
  for (int c = 0; c < 3; ++c) x[c] = a.origins[ray * 3 + c] + a.directions[ray * 3 + c] * tm;   // nerf.py:300
  if (a.contract) {                                                       // spatial_distortion_norm2
    const float m = fmaxf(x[0] * x[0] + x[1] * x[1] + x[2] * x[2], kF32Eps);
    if (!(m <= 1.f)) {
      const float sc = (2.f * sqrtf(m) - 1.f) / m;
      for (int c = 0; c < 3; ++c) x[c] = sc * x[c];
    }
  }
}

// column `col` of pos_enc(x, min_deg, min_deg + ndeg, append_identity=True): [x, sin(2^k x_c) k-major, sin(2^k x_c + pi/2)]
__device__ __forceinline__ float pe_value(const float x[3], int col, int min_deg, int ndeg) {
  if (col < 3) return x[col];
  int j = col - 3;
  const int shifted = j >= 3 * ndeg;
  if (shifted) j -= 3 * ndeg;
  const int k = j / 3, c = j % 3;
  float v = x[c] * exp2f((float)(min_deg + k));
  if (shifted) v = v + 1.57079637050628662109375f;    // fp32(0.5 * pi)
  return sinf(v);
}

__global__ void __launch_bounds__(256) point_pe_f32_kernel(PointPeArgs a) {
  const int fd = 3 + 6 * a.ndeg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.n_rays * a.S * fd) return;
  const int s = (int)(idx / fd), col = (int)(idx % fd);
  float x[3];
  sample_point(a, s, x);
  a.features[idx] = pe_value(x, col, a.min_deg, a.ndeg);
}

// one thread = 8 consecutive columns (16 bytes) of one row
__global__ void __launch_bounds__(256) point_pe_bf16_kernel(PointPeArgs a) {
  const int cpr = a.zero_cols / 8;                       // 16-byte chunks per row
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.rows_pad * cpr) return;
  const int s = (int)(idx / cpr), ch = (int)(idx % cpr);
  const int fd = 3 + 6 * a.ndeg;
  uint32_t wh[4] = {0u, 0u, 0u, 0u}, wl[4] = {0u, 0u, 0u, 0u};
  if (s < a.n_rays * a.S && ch * 8 < fd) {
    float x[3];
    sample_point(a, s, x);
    for (int q = 0; q < 4; ++q) {
      float f[2];
      for (int e = 0; e < 2; ++e) { const int col = ch * 8 + q * 2 + e; f[e] = col < fd ? pe_value(x, col, a.min_deg, a.ndeg) : 0.f; }
      const __nv_bfloat16 h0 = __float2bfloat16(f[0]), h1 = __float2bfloat16(f[1]);
      wh[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      if (a.lo) {
        const __nv_bfloat16 l0 = __float2bfloat16(f[0] - __bfloat162float(h0)), l1 = __float2bfloat16(f[1] - __bfloat162float(h1));
        wl[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
    }
  }
  reinterpret_cast<uint4*>(a.hi + (size_t)s * a.ld)[ch] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
  if (a.lo) reinterpret_cast<uint4*>(a.lo + (size_t)s * a.ld)[ch] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
}

template <class K>
int opt_in_smem(K kernel, bool* flags, int bytes) {
  int dev = 0;
  HUGS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !flags[dev]) {
    HUGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) flags[dev] = true;
  }
  return HUGS_OK;
}

}  // namespace

int launch_nf_merge(const NfMergeArgs& a, cudaStream_t stream) {
  HUGS_REQUIRE(a.na >= 1 && a.nb >= 1 && a.na + a.nb >= 2 && a.na + a.nb <= 4096, "nf_merge: bad bin counts %d + %d", a.na, a.nb);
  static bool f[64] = {};
  int rc = opt_in_smem(nf_merge_kernel, f, 160 * 1024);
  if (rc) return rc;
  if (a.n_rays <= 0) return HUGS_OK;
  const size_t smem = (size_t)kWarps * 2 * (a.na + a.nb) * sizeof(float);
  nf_merge_kernel<<<(a.n_rays + kWarps - 1) / kWarps, kWarps * 32, smem, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_composite(const NfCompositeArgs& a, bool backward, cudaStream_t stream) {
  HUGS_REQUIRE(a.S >= 2 && a.S <= 2048, "nf_composite: samples per ray must be in [2,2048], got %d", a.S);
  HUGS_REQUIRE(a.C == 1 || a.C == 4, "nf_composite: raw_channels must be 1 or 4");
  static bool ff[64] = {}, fb[64] = {};
  int rc = backward ? opt_in_smem(nf_composite_kernel<true>, fb, 160 * 1024) : opt_in_smem(nf_composite_kernel<false>, ff, 160 * 1024);
  if (rc) return rc;
  if (a.n_rays <= 0) return HUGS_OK;
  const size_t smem = (size_t)kWarps * (backward ? 5 : 3) * a.S * sizeof(float);
  const int grid = (a.n_rays + kWarps - 1) / kWarps;
  if (backward) nf_composite_kernel<true><<<grid, kWarps * 32, smem, stream>>>(a);
  else nf_composite_kernel<false><<<grid, kWarps * 32, smem, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_clip_depth(float* depth, const float* steps_max, int n, cudaStream_t stream) {
  if (n <= 0) return HUGS_OK;
  nf_clip_depth_kernel<<<(n + 255) / 256, 256, 0, stream>>>(depth, steps_max, n);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_rgb_loss(const float* pred, const float* gt, const float* mask, float transient_w, int loss_type,
                       float padding, int n, float* sums, float* dl, cudaStream_t stream) {
  HUGS_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(float), stream));
  if (n <= 0) return HUGS_OK;
  const int grid = std::min((n * 3 + 255) / 256, 296);
  nf_rgb_loss_kernel<<<grid, 256, 0, stream>>>(pred, gt, mask, transient_w, loss_type, padding, n, sums, dl);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_rgb_loss_bwd(const float* dl, const float* sums, const float* upstream, float scale, int n, float* d_pred,
                           cudaStream_t stream) {
  if (n <= 0) return HUGS_OK;
  nf_rgb_loss_bwd_kernel<<<(n * 3 + 255) / 256, 256, 0, stream>>>(dl, sums, upstream, scale, n, d_pred);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_distortion(const float* c, const float* w, int n_rays, int S, float* out, float* grad, cudaStream_t stream) {
  HUGS_REQUIRE(S >= 1 && S <= 4096, "nf_distortion: bad sample count %d", S);
  static bool f[64] = {};
  int rc = opt_in_smem(nf_distortion_kernel, f, 160 * 1024);
  if (rc) return rc;
  HUGS_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  if (n_rays <= 0) return HUGS_OK;
  NfLossArgs a{c, w, S, nullptr, nullptr, 0, n_rays, out, grad};
  nf_distortion_kernel<<<(n_rays + kWarps - 1) / kWarps, kWarps * 32, (size_t)kWarps * 2 * S * sizeof(float), stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_interlevel(const float* c, const float* w, int S, const float* cp, const float* wp, int Sp, int n_rays,
                         float* out, float* grad, cudaStream_t stream) {
  HUGS_REQUIRE(S >= 1 && Sp >= 1 && Sp <= 4096, "nf_interlevel: bad sample counts %d / %d", S, Sp);
  static bool f[64] = {};
  int rc = opt_in_smem(nf_interlevel_kernel, f, 160 * 1024);
  if (rc) return rc;
  HUGS_CUDA(cudaMemsetAsync(out, 0, sizeof(float), stream));
  if (n_rays <= 0) return HUGS_OK;
  NfLossArgs a{c, w, S, cp, wp, Sp, n_rays, out, grad};
  nf_interlevel_kernel<<<(n_rays + kWarps - 1) / kWarps, kWarps * 32, (size_t)kWarps * (2 * Sp + 2) * sizeof(float), stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_nf_scale(const float* src, const float* upstream, float mult, long long n, float* dst, cudaStream_t stream) {
  if (n <= 0) return HUGS_OK;
  nf_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, upstream, mult, n, dst);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_params_copy(const hugs_tensor_copy* table, int n, float* flat, int direction, float* base, cudaStream_t stream) {
  if (n <= 0) return HUGS_OK;
  params_copy_kernel<<<dim3(32, n), 256, 0, stream>>>(table, flat, direction, base);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_point_pe(const PointPeArgs& a, cudaStream_t stream) {
  if (a.features) {
    const long long tot = (long long)a.n_rays * a.S * (3 + 6 * a.ndeg);
    if (tot > 0) {
      point_pe_f32_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(a);
      HUGS_LAUNCH_CHECK();
    }
  }
  if (a.hi) {
    HUGS_REQUIRE(a.zero_cols % 8 == 0 && a.zero_cols >= 3 + 6 * a.ndeg && a.zero_cols <= a.ld, "point_pe: bad padding");
    const long long tot = (long long)a.rows_pad * (a.zero_cols / 8);
    if (tot > 0) {
      point_pe_bf16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(a);
      HUGS_LAUNCH_CHECK();
    }
  }
  return HUGS_OK;
}

}  // namespace hugs
