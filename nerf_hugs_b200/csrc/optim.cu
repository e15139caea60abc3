// Per-module gradient clipping + nan_to_num + Adam (one fused elementwise pass over the flat
// parameter buffer), and the small reductions that feed it.
//
// Reference semantics (paths under /root/reference/MipNeRF360/internal):
//   train_utils.py:351-369  clip_gradients: per top-level module, value clip then norm clip
//   train_utils.py:464-468  nan_to_num, state.apply_gradients (optax.adam, b1/b2/eps from configs.py:99-107)
//   train_utils.py:461-462  stats['grad_norms'], stats['grad_maxes']
#include "handle.h"
#include "tc.h"

namespace hugs {
namespace {

constexpr int kRedBlocks = 128;

__device__ __forceinline__ float clipv(float g, float max_val) {
  return max_val > 0.f ? fminf(fmaxf(g, -max_val), max_val) : g;
}

// partial[m][b] = {sum g^2 (raw), max |g| (raw), sum clip(g)^2}
__global__ void __launch_bounds__(256) grad_norm_partial_kernel(const float* grad, int64_t b0, int64_t e0, int64_t b1,
                                                                int64_t e1, int64_t b2, int64_t e2, float max_val,
                                                                float gscale, float* partial) {
  __shared__ float red[3][8];
  const int m = blockIdx.y;
  const int64_t beg = m == 0 ? b0 : (m == 1 ? b1 : b2), end = m == 0 ? e0 : (m == 1 ? e1 : e2);
  float s = 0.f, mx = 0.f, sc = 0.f;
  for (int64_t i = beg + (int64_t)blockIdx.x * 256 + threadIdx.x; i < end; i += (int64_t)kRedBlocks * 256) {
    float g = grad[i] * gscale;
    s += g * g;
    mx = fmaxf(mx, fabsf(g));
    float c = clipv(g, max_val);
    sc += c * c;
  }
  s = warp_sum(s); sc = warp_sum(sc); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = mx; red[2][threadIdx.x >> 5] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b = fmaxf(b, red[1][w]); c += red[2][w]; }
    float* p = partial + ((size_t)m * kRedBlocks + blockIdx.x) * 3;
    p[0] = a; p[1] = b; p[2] = c;
  }
}

// out[m] = {norm_raw, max_raw, clip multiplier}
__global__ void grad_norm_final_kernel(const float* partial, float max_norm, float* out) {
  const int m = threadIdx.x;
  if (m >= 3) return;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int i = 0; i < kRedBlocks; ++i) {
    const float* p = partial + ((size_t)m * kRedBlocks + i) * 3;
    a += p[0]; b = fmaxf(b, p[1]); c += p[2];
  }
  out[m * 3 + 0] = sqrtf(a);
  out[m * 3 + 1] = b;
  out[m * 3 + 2] = max_norm > 0.f ? fminf(1.f, max_norm / (kF32Eps + sqrtf(c))) : 1.f;
}

constexpr int kAdamPerThread = 8;            // elements per thread: a block covers 2048 consecutive parameters

// kStats: additionally accumulates, per parameter tensor, {sum w^2 (before the update), sum g^2, max |g| (the averaged,
// unclipped gradient), sum delta^2, max |delta| (the applied update)} -> tstats[tensor][5]: the stats['weight_l2s'],
// ['grad_norms'], ['grad_maxes'], ['opt_update_norms'], ['opt_update_maxes'] trees of train_utils.py:442,461-462,470-473
// come out of the pass that already touches every parameter.  A block that lies inside one tensor (almost all do)
// reduces in shared memory and issues one set of five atomics.
template <bool kStats>
__global__ void __launch_bounds__(256) adam_kernel(float* params, const float* grad, float* mu, float* nu,
                                                   int64_t n, int64_t e0, int64_t e1, const float* norms,
                                                   hugs_adam_cfg cfg, float bc1, float bc2,
                                                   const int64_t* tensor_ends, int n_tensors, float* tstats) {
  __shared__ int64_t ends[128];
  __shared__ float red[5][8];
  if (kStats) {
    for (int k = threadIdx.x; k < n_tensors; k += 256) ends[k] = tensor_ends[k];
    __syncthreads();
  }
  const int64_t base = (int64_t)blockIdx.x * (256 * kAdamPerThread);
  auto tensor_of = [&](int64_t key) {
    int lo = 0, hi = n_tensors - 1;                        // first k with key < ends[k]
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (key < ends[mid]) hi = mid; else lo = mid + 1; }
    return lo;
  };
  const int64_t last = (base + 256 * kAdamPerThread <= n ? base + 256 * kAdamPerThread : n) - 1;
  const bool uniform = kStats && tensor_of(base) == tensor_of(last);
  float w2 = 0.f, g2 = 0.f, ga = 0.f, d2 = 0.f, da = 0.f;
#pragma unroll
  for (int r = 0; r < kAdamPerThread; ++r) {
    const int64_t i = base + r * 256 + threadIdx.x;
    if (i >= n) break;
    const int m = i < e0 ? 0 : (i < e1 ? 1 : 2);
    const float gs = grad[i] * cfg.grad_scale;
    float g = clipv(gs, cfg.grad_max_val) * norms[m * 3 + 2];
    if (g != g) g = 0.f;                                  // jnp.nan_to_num
    else if (isinf(g)) g = g > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    float m1 = cfg.beta1 * mu[i] + (1.f - cfg.beta1) * g;
    float m2 = cfg.beta2 * nu[i] + (1.f - cfg.beta2) * g * g;
    mu[i] = m1; nu[i] = m2;
    float mhat = m1 / bc1, vhat = m2 / bc2;
    const float p_old = params[i];
    const float p_new = p_old - cfg.lr * mhat / (sqrtf(vhat) + cfg.eps);
    params[i] = p_new;
    if (kStats) {
      const float d = p_new - p_old;
      if (uniform) {
        w2 += p_old * p_old; g2 += gs * gs; ga = fmaxf(ga, fabsf(gs)); d2 += d * d; da = fmaxf(da, fabsf(d));
      } else {                                             // a block that straddles tensors (biases): per warp / lane
        const unsigned act = __activemask();
        const int t = tensor_of(i);
        const int t0 = __shfl_sync(act, t, __ffs(act) - 1);
        float a0 = p_old * p_old, a1 = gs * gs, a2 = fabsf(gs), a3 = d * d, a4 = fabsf(d);
        float* dst = tstats + (size_t)t * 5;
        if (act == kFull && __all_sync(kFull, t == t0)) {
          a0 = warp_sum(a0); a1 = warp_sum(a1); a3 = warp_sum(a3); a2 = warp_max(a2); a4 = warp_max(a4);
          if ((threadIdx.x & 31) != 0) continue;
        }
        atomicAdd(dst + 0, a0); atomicAdd(dst + 1, a1); atomicAdd(dst + 3, a3);
        atomicMax(reinterpret_cast<int*>(dst + 2), __float_as_int(a2));
        atomicMax(reinterpret_cast<int*>(dst + 4), __float_as_int(a4));
      }
    }
  }
  if (kStats && uniform) {
    w2 = warp_sum(w2); g2 = warp_sum(g2); d2 = warp_sum(d2); ga = warp_max(ga); da = warp_max(da);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][w] = w2; red[1][w] = g2; red[2][w] = ga; red[3][w] = d2; red[4][w] = da; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < 8; ++k) {
        red[0][0] += red[0][k]; red[1][0] += red[1][k]; red[3][0] += red[3][k];
        red[2][0] = fmaxf(red[2][0], red[2][k]); red[4][0] = fmaxf(red[4][0], red[4][k]);
      }
      float* dst = tstats + (size_t)tensor_of(base) * 5;
      atomicAdd(dst + 0, red[0][0]); atomicAdd(dst + 1, red[1][0]); atomicAdd(dst + 3, red[3][0]);
      atomicMax(reinterpret_cast<int*>(dst + 2), __float_as_int(red[2][0]));
      atomicMax(reinterpret_cast<int*>(dst + 4), __float_as_int(red[4][0]));
    }
  }
}

// One block: the per-ray loss partials (ray_stats: [n, 4] of the final level, then one [n] column per proposal level
// for the interlevel term and one for the level's squared error) are summed column by column in a fixed order, then
// the scalars of stats_out are formed (train_utils.py:93-111, 228-248): deterministic, one launch.
__global__ void __launch_bounds__(1024) reduce_finalize_stats_kernel(hugs_loss_cfg loss, int n, int L, int S_final,
                                                                    const float* denom, const float* ray_stats,
                                                                    float* stats) {
  __shared__ float red[32];
  __shared__ float sums[16];
  // column c: c < 3 -> ray_stats[r * 4 + c]; 4 + l -> interlevel of level l; 8 + l -> squared error of level l
  for (int c = 0; c < 12; ++c) {
    const bool final_col = c < 3;
    const int l = c >= 8 ? c - 8 : c - 4;
    if (!final_col && (c == 3 || l < 0 || l >= L - 1)) { if (threadIdx.x == 0) sums[c] = 0.f; continue; }
    const float* src = final_col ? ray_stats + c : ray_stats + (size_t)n * c;
    const int stride = final_col ? 4 : 1;
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += src[(size_t)i * stride];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      v = warp_sum(v);
      if (threadIdx.x == 0) sums[c] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const float dn = fmaxf(denom[0], kF32Eps);
  const float data = loss.data_loss_mult * sums[0] / dn;
  const float mse = sums[1] / dn;
  const float dist = loss.distortion_loss_mult > 0.f ? loss.distortion_loss_mult * sums[2] / (float)n : 0.f;
  float inter = 0.f;
  if (loss.interlevel_loss_mult > 0.f)
    for (int l = 0; l < L - 1; ++l) inter += sums[4 + l] / ((float)n * (float)S_final);
  inter *= loss.interlevel_loss_mult;
  for (int i = 0; i < 16; ++i) stats[i] = 0.f;
  stats[0] = data + inter + dist; stats[1] = data; stats[2] = inter; stats[3] = dist;
  stats[4 + L - 1] = mse;
  // proposal levels render rgb = max(0, 1 - acc) * bg (their MLP has no colour, models.py:469-470): stats['mses'][l]
  for (int l = 0; l < L - 1; ++l) stats[4 + l] = sums[8 + l] / dn;
}

}  // namespace

int launch_finalize_stats(hugs_handle* h, const hugs_loss_cfg& loss, int n, const float* denom, const float* ray_stats,
                          float* stats_out, cudaStream_t st) {
  reduce_finalize_stats_kernel<<<1, 1024, 0, st>>>(loss, n, h->d.num_levels, h->samples(h->d.num_levels - 1), denom,
                                                   ray_stats, stats_out);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs

using namespace hugs;

namespace {
int adam_step_impl(hugs_handle* h, float* params, const float* grad, float* mu, float* nu, const hugs_adam_cfg* cfg,
                   float* norms_out, float* tensor_stats_out, cudaStream_t st) {
  HUGS_REQUIRE(h && params && grad && mu && nu && cfg, "hugs_adam_step: null argument");
  HUGS_REQUIRE(cfg->step >= 0, "hugs_adam_step: step must be >= 0");
  HUGS_REQUIRE(cfg->grad_scale > 0.f, "hugs_adam_step: grad_scale must be > 0 (1 on a single GPU)");
  ProfScope ps(h, HUGS_K_ADAM_PACK, st);
  float* partial = h->scalars + 64;   // see hugs_create: scalars has room for 64 + 3*kRedBlocks*3
  float* norms = h->scalars + 32;
  dim3 grid(kRedBlocks, 3);
  grad_norm_partial_kernel<<<grid, 256, 0, st>>>(grad, h->module_begin[0], h->module_end[0], h->module_begin[1],
                                                 h->module_end[1], h->module_begin[2], h->module_end[2],
                                                 cfg->grad_max_val, cfg->grad_scale, partial);
  HUGS_LAUNCH_CHECK();
  grad_norm_final_kernel<<<1, 32, 0, st>>>(partial, cfg->grad_max_norm, norms);
  HUGS_LAUNCH_CHECK();
  const double t = (double)cfg->step + 1.0;
  const float bc1 = (float)(1.0 - pow((double)cfg->beta1, t)), bc2 = (float)(1.0 - pow((double)cfg->beta2, t));
  const int64_t n = h->n_params;
  const int n_tensors = (int)h->tensors.size();
  if (tensor_stats_out) {
    HUGS_REQUIRE(n_tensors <= 128, "too many parameter tensors for the per-tensor statistics");
    HUGS_CUDA(cudaMemsetAsync(tensor_stats_out, 0, sizeof(float) * 5 * n_tensors, st));
    adam_kernel<true><<<(unsigned)((n + 256 * kAdamPerThread - 1) / (256 * kAdamPerThread)), 256, 0, st>>>(params, grad, mu, nu, n, h->module_end[0],
                                                                   h->module_end[1], norms, *cfg, bc1, bc2,
                                                                   h->tensor_ends, n_tensors, tensor_stats_out);
  } else {
    adam_kernel<false><<<(unsigned)((n + 256 * kAdamPerThread - 1) / (256 * kAdamPerThread)), 256, 0, st>>>(params, grad, mu, nu, n, h->module_end[0],
                                                                    h->module_end[1], norms, *cfg, bc1, bc2, nullptr,
                                                                    0, nullptr);
  }
  HUGS_LAUNCH_CHECK();
  if (norms_out) HUGS_CUDA(cudaMemcpyAsync(norms_out, norms, sizeof(float) * 9, cudaMemcpyDeviceToDevice, st));
  if (h->d.precision != HUGS_PRECISION_FP32) return tc_pack_params(h, params, st);
  return HUGS_OK;
}
}  // namespace

HUGS_API int hugs_adam_step(hugs_handle* h, float* params, const float* grad, float* mu, float* nu,
                            const hugs_adam_cfg* cfg, float* norms_out, void* stream) {
  return adam_step_impl(h, params, grad, mu, nu, cfg, norms_out, nullptr, (cudaStream_t)stream);
}

HUGS_API int hugs_adam_step_stats(hugs_handle* h, float* params, const float* grad, float* mu, float* nu,
                                  const hugs_adam_cfg* cfg, float* norms_out, float* tensor_stats_out, void* stream) {
  HUGS_REQUIRE(tensor_stats_out, "hugs_adam_step_stats: tensor_stats_out is required");
  return adam_step_impl(h, params, grad, mu, nu, cfg, norms_out, tensor_stats_out, (cudaStream_t)stream);
}
