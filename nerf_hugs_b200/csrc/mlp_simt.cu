// fp32 CUDA-core path ("parity mode", HUGS_PRECISION_FP32): exact-arithmetic IPE features and a
// tiled fp32 Dense kernel.  Slow by design (no tensor cores); it exists so that renders can be
// compared with the fp32 reference at 1e-4, which bf16 tensor-core operands cannot reach.
//
// Reference semantics: models.py:437-519 (MLP.__call__), coord.py:102-147.
#include "common.cuh"
#include "encode.cuh"
#include "kernels.h"
#include "mlp.h"
#include "tc_internal.h"

namespace hugs {
namespace {

// One warp per sample; lanes stride over (degree, basis) pairs so stores are coalesced.
__global__ void __launch_bounds__(256) ipe_features_kernel(IpeArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.n_rays * a.S) return;
  const int ray = warp / a.S, i = warp % a.S;
  float o[3], d[3];
  for (int c = 0; c < 3; ++c) { o[c] = a.origins[ray * 3 + c]; d[c] = a.directions[ray * 3 + c]; }
  const float t0 = a.tdist[(size_t)ray * (a.S + 1) + i], t1 = a.tdist[(size_t)ray * (a.S + 1) + i + 1];
  SampleGauss g;
  frustum_gaussian(o, d, a.radii[ray], t0, t1, a.ray_shape, a.contract, g);
  const int nb = a.num_basis, ndeg = a.max_deg - a.min_deg, half = nb * ndeg;
  float* out = a.features + (size_t)warp * (2 * half);
  for (int idx = lane; idx < half; idx += 32) {
    const int k = idx / nb, b = idx % nb;
    float p[3] = {a.basis[b], a.basis[nb + b], a.basis[2 * nb + b]};
    float mu, var;
    lift_basis(g, d, p, mu, var);
    const float scale = exp2f((float)(a.min_deg + k));
    const float sm = mu * scale, sv = var * scale * scale;
    const float e = expf(-0.5f * sv);
    out[idx] = e * safe_sin_ref(sm);
    out[half + idx] = e * safe_sin_ref(sm + 1.57079637050628662109375f);  // fp32(0.5*pi)
  }
}

// dir_enc(viewdirs) ++ glo_vec per ray: [n, 3 + 6*deg_view + glo]  (models.py:399-403,488-501)
__global__ void view_inputs_kernel(const float* viewdirs, const int32_t* embed_idx, const float* glo_table,
                                   int n_rays, int deg_view, int glo, int zero_glo, int num_embeddings, float* out) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n_rays) return;
  const int width = 3 + 6 * deg_view + glo;
  float* o = out + (size_t)ray * width;
  float v[3] = {viewdirs[ray * 3], viewdirs[ray * 3 + 1], viewdirs[ray * 3 + 2]};
  for (int c = 0; c < 3; ++c) o[c] = v[c];
  for (int k = 0; k < deg_view; ++k)
    for (int c = 0; c < 3; ++c) {
      float x = v[c] * exp2f((float)k);
      o[3 + k * 3 + c] = sinf(x);
      o[3 + 3 * deg_view + k * 3 + c] = sinf(x + 1.57079637050628662109375f);
    }
  // an out-of-range row index never leaves the table (the host surface validates embed_idx; nn.Embed never corrupts memory)
  const int row = glo > 0 && !zero_glo ? min(max(embed_idx[ray], 0), num_embeddings - 1) : 0;
  for (int j = 0; j < glo; ++j)
    o[3 + 6 * deg_view + j] = zero_glo ? 0.f : glo_table[(size_t)row * glo + j];
}

// Split-precision tensor-core mode: the exact features above (same arithmetic, compiled with -fmad=false), each written
// as bf16 hi + bf16 residual in the engine's column order f' = (b * ndeg + k) * 2 + {sin, shifted sin}; one warp per row.
__global__ void __launch_bounds__(256) encode_split_kernel(EncSplitArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.n_rows_pad) return;
  uint32_t* hi = reinterpret_cast<uint32_t*>(a.feat_hi + (size_t)warp * kFeatPad);
  uint32_t* lo = reinterpret_cast<uint32_t*>(a.feat_lo + (size_t)warp * kFeatPad);
  const int nb = a.nb, ndeg = a.ndeg, half = nb * ndeg;
  if (warp >= a.n_samples) {   // padding rows of the last tile: finite (zero) features
    for (int w = lane; w < kFeatPad / 2; w += 32) { hi[w] = 0u; lo[w] = 0u; }
    return;
  }
  const int ray = warp / a.S, i = warp % a.S;
  float o[3], d[3];
  for (int c = 0; c < 3; ++c) { o[c] = a.origins[ray * 3 + c]; d[c] = a.directions[ray * 3 + c]; }
  const float t0 = a.tdist[(size_t)ray * (a.S + 1) + i], t1 = a.tdist[(size_t)ray * (a.S + 1) + i + 1];
  SampleGauss g;
  frustum_gaussian(o, d, a.radii[ray], t0, t1, a.ray_shape, a.contract, g);
  for (int idx = lane; idx < kFeatPad / 2; idx += 32) {      // idx = b * ndeg + k: one (sin, shifted sin) column pair
    uint32_t wh = 0u, wl = 0u;
    if (idx < half) {
      const int b = idx / ndeg, k = idx % ndeg;
      float p[3] = {a.basis[b], a.basis[nb + b], a.basis[2 * nb + b]};
      float mu, var;
      lift_basis(g, d, p, mu, var);
      const float scale = exp2f((float)(a.min_deg + k));
      const float sm = mu * scale, sv = var * scale * scale;
      const float e = expf(-0.5f * sv);
      const float f0 = e * safe_sin_ref(sm);
      const float f1 = e * safe_sin_ref(sm + 1.57079637050628662109375f);
      const __nv_bfloat16 h0 = __float2bfloat16(f0), h1 = __float2bfloat16(f1);
      const __nv_bfloat16 l0 = __float2bfloat16(f0 - __bfloat162float(h0)), l1 = __float2bfloat16(f1 - __bfloat162float(h1));
      wh = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      wl = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    hi[idx] = wh; lo[idx] = wl;
  }
}

constexpr int TM = 64, TN = 64, TK = 16;

// y[m, n] = act(sum_seg x_seg[m / row_div, :] . W[koff_seg + :, n] + b[n])
__global__ void __launch_bounds__(256) dense_simt_kernel(DenseArgs a) {
  __shared__ float Xs[TK][TM + 1];
  __shared__ float Ws[TK][TN + 1];
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4x4 outputs each
  float acc[4][4] = {};
  int koff = 0;
  for (int s = 0; s < a.nseg; ++s) {
    const DenseSeg seg = a.seg[s];
    for (int k0 = 0; k0 < seg.k; k0 += TK) {
      for (int e = threadIdx.x; e < TM * TK; e += 256) {
        int mm = e / TK, kk = e % TK;
        int m = m0 + mm, k = k0 + kk;
        Xs[kk][mm] = (m < a.M && k < seg.k) ? seg.x[(size_t)(m / seg.row_div) * seg.ld + k] : 0.f;
      }
      for (int e = threadIdx.x; e < TK * TN; e += 256) {
        int kk = e / TN, nn = e % TN;
        int k = k0 + kk, n = n0 + nn;
        Ws[kk][nn] = (k < seg.k && n < a.N) ? a.W[(size_t)(koff + k) * a.N + n] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float xv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = Xs[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = Ws[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
      }
      __syncthreads();
    }
    koff += seg.k;
  }
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
      if (a.relu) v = fmaxf(v, 0.f);
      a.y[(size_t)m * a.ldy + n] = v;
    }
  }
}

}  // namespace

int launch_ipe_features(const IpeArgs& a, cudaStream_t stream) {
  long long warps = (long long)a.n_rays * a.S;
  if (warps <= 0) return HUGS_OK;
  long long blocks = (warps * 32 + 255) / 256;
  ipe_features_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_encode_split(const EncSplitArgs& a, cudaStream_t stream) {
  if (a.n_rows_pad <= 0) return HUGS_OK;
  const long long blocks = ((long long)a.n_rows_pad * 32 + 255) / 256;
  encode_split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_view_inputs(const float* viewdirs, const int32_t* embed_idx, const float* glo_table, int n_rays,
                       int deg_view, int glo, int zero_glo, int num_embeddings, float* out, cudaStream_t stream) {
  if (n_rays <= 0) return HUGS_OK;
  view_inputs_kernel<<<(n_rays + 127) / 128, 128, 0, stream>>>(viewdirs, embed_idx, glo_table, n_rays,
                                                               deg_view, glo, zero_glo, num_embeddings, out);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int launch_dense_simt(const DenseArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return HUGS_OK;
  dim3 grid((a.M + TM - 1) / TM, (a.N + TN - 1) / TN);
  dense_simt_kernel<<<grid, 256, 0, stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs
