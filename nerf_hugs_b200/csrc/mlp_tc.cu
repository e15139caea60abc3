// Tensor-core MLP path (HUGS_PRECISION_BF16_TC / HUGS_PRECISION_TC_SPLIT), host side + the small kernels around the
// chain kernel (mlp_pp.cu): bf16 IPE feature encoder, per-ray view bias, parameter packing.
//
// Layer schedule, packing and the feature-column permutation are documented in DESIGN.md.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <vector>

#include "encode.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"
#include "dense_tc.h"
#include "tc_internal.h"

namespace hugs {

namespace {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// bf16 feature encoder (throughput mode): one thread = (sample, half of the basis directions)
// ------------------------------------------------------------------------------------------
struct EncArgs {
  const float* origins; const float* directions; const float* radii; const float* tdist;
  const float* basis;     // [3][nb]
  int n_samples, n_rows_pad, S, nb, min_deg, ndeg, ray_shape, contract;
  __nv_bfloat16* feat;    // [rows, 512] (already offset to the level's first row)
};

#ifndef HUGS_ENC_ROWS
#define HUGS_ENC_ROWS 8
#endif
#ifndef HUGS_ENC_SPLIT
#define HUGS_ENC_SPLIT 4
#endif
constexpr int kEncRows = HUGS_ENC_ROWS;
constexpr int kEncSplit = HUGS_ENC_SPLIT;     // threads per sample (each takes a contiguous range of basis directions)
// kNdeg > 0: number of IPE degrees known at compile time (12 for every shipped config): the per-direction word buffer
// stays in registers instead of a dynamically indexed local array.
template <int kNdeg>
__global__ void __launch_bounds__(kEncSplit * kEncRows) encode_bf16_kernel(EncArgs a) {
  __shared__ __align__(16) uint4 tile[kEncRows * 64];   // rows x 64 chunks of 16 B, chunk index swizzled
  const int tid = threadIdx.x, rl = tid % kEncRows, part = tid / kEncRows;
  const int s = blockIdx.x * kEncRows + rl;
  const int b_beg = (a.nb * part) / kEncSplit, b_end = (a.nb * (part + 1)) / kEncSplit;
  const bool half = part == kEncSplit - 1;      // the last part also clears the padding columns
  const int ndeg = kNdeg > 0 ? kNdeg : a.ndeg;
  const int cpb = ndeg >> 2;                      // 16-byte chunks per basis direction
  if (s < a.n_samples) {
    const int ray = s / a.S, i = s % a.S;
    float o[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { o[c] = a.origins[ray * 3 + c]; d[c] = a.directions[ray * 3 + c]; }
    const float t0 = a.tdist[(size_t)ray * (a.S + 1) + i], t1 = a.tdist[(size_t)ray * (a.S + 1) + i + 1];
    SampleGauss g;
    frustum_gaussian(o, d, a.radii[ray], t0, t1, a.ray_shape, a.contract, g);
    const float sc0 = exp2f((float)a.min_deg);
    for (int b = b_beg; b < b_end; ++b) {
      float p[3] = {a.basis[b], a.basis[a.nb + b], a.basis[2 * a.nb + b]};
      float mu, var;
      lift_basis(g, d, p, mu, var);
      // phase in turns, split hi/lo so that 2^k * phase (mod 1) stays accurate at degree 11
      const float kInv2PiHi = 0.15915494309189535f, kInv2PiLo = -1.2154201256553420e-10f;   // 1/(2 pi)
      const float r_hi = mu * kInv2PiHi;
      const float r_lo = fmaf(mu, kInv2PiHi, -r_hi) + mu * kInv2PiLo;
      const float av = -0.5f * 1.4426950408889634f * var;                                  // exp(x)=2^(x log2e)
      float sc = sc0;
      uint32_t w[16];
#pragma unroll
      for (int k = 0; k < (kNdeg > 0 ? kNdeg : 16); k += 2) {
        if (kNdeg == 0 && k >= ndeg) break;
        float t = r_hi * sc;
        float f = (t - rintf(t)) + r_lo * sc;
        float x = f * 6.283185307179586f;
        float sn = __sinf(x), cs = __cosf(x);
        float e = exp2f(av * sc * sc);
        w[k] = ptx::pack_bf16x2(e * sn, e * cs);
        float sn2 = 2.f * sn * cs, cs2 = 1.f - 2.f * sn * sn;
        float e2 = e * e; e2 = e2 * e2;
        w[k + 1] = ptx::pack_bf16x2(e2 * sn2, e2 * cs2);
        sc *= 4.f;
      }
#pragma unroll
      for (int c = 0; c < (kNdeg > 0 ? kNdeg / 4 : 4); ++c) {
        if (kNdeg == 0 && c >= cpb) break;
        uint4 v = make_uint4(w[c * 4], w[c * 4 + 1], w[c * 4 + 2], w[c * 4 + 3]);
        tile[rl * 64 + swz_chunk(rl, b * cpb + c)] = v;
      }
    }
    if (half) {
      for (int c = a.nb * cpb; c < 64; ++c) tile[rl * 64 + swz_chunk(rl, c)] = make_uint4(0, 0, 0, 0);
    }
  } else if (s < a.n_rows_pad) {
    // rows between n_samples and the 128-row tile boundary: finite (zero) features so that the saved
    // activations of padding rows can never poison the weight-gradient reduction
    for (int c = part * (64 / kEncSplit); c < (part + 1) * (64 / kEncSplit); ++c)
      tile[rl * 64 + swz_chunk(rl, c)] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  // coalesced copy-out of the block's consecutive rows
  const int rows = min(kEncRows, a.n_rows_pad - blockIdx.x * kEncRows);
  uint4* dst = reinterpret_cast<uint4*>(a.feat) + (size_t)blockIdx.x * kEncRows * 64;
  for (int e = tid; e < rows * 64; e += kEncSplit * kEncRows) {
    int r = e >> 6, c = e & 63;
    dst[e] = tile[r * 64 + swz_chunk(r, c)];
  }
}

// viewbias[ray][c] = sum_j bf16(view_in[ray][j]) * bf16(W_view[256 + j][c]) + b_view[c]
__global__ void viewbias_kernel(const float* view_in, int view_in_dim, const float* params, long long koff,
                                long long boff, int bott_w, int out, int n_rays, int exact, float* vb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * out) return;
  const int ray = idx / out, c = idx % out;
  float acc = 0.f;
  for (int j = 0; j < view_in_dim; ++j) {
    float x = view_in[(size_t)ray * view_in_dim + j];
    float w = params[koff + (long long)(bott_w + j) * out + c];
    if (!exact) { x = __bfloat162float(__float2bfloat16(x)); w = __bfloat162float(__float2bfloat16(w)); }
    acc = fmaf(x, w, acc);
  }
  vb[idx] = acc + params[boff + c];
}

// ------------------------------------------------------------------------------------------
// parameter packing: flat fp32 (flax layout) -> bf16 operand tensors
// ------------------------------------------------------------------------------------------
struct PackArgs {
  TcMlp::PackLayer layers[16];
  int n_layers, rows_f, rows_b, nb, ndeg, feat_dim, bias_floats;
  const float* params;
  __nv_bfloat16* wt; __nv_bfloat16* wn; float* bias;
  int w_dens_off, w_rgb_off; long long dens_koff, rgb_koff; int dens_in, rgb_in;
  __nv_bfloat16* bias_img; int n_bias_layers; int bias_layer[4 * kBiasChunks];   // dense-layer index of biased layer j
  int part;         // 0: bf16(w) (+ bias tables); 1: bf16(w - bf16(w)) into the lo half of wt / wn (split-precision mode)
  int exact_heads;  // split-precision mode: the fp32 head-weight table is not rounded to bf16
};

// operand word of weight value v for pack part `part`
__device__ __forceinline__ __nv_bfloat16 pack_part(float v, int part) {
  const __nv_bfloat16 hi = __float2bfloat16(v);
  return part == 0 ? hi : __float2bfloat16(v - __bfloat162float(hi));
}

__global__ void pack_params_kernel(PackArgs a) {
  const long long nf = (long long)a.rows_f * kKP, nbk = (long long)a.rows_b * kW;
  const long long nimg = 2LL * kBiasChunks * kBiasChunkElems;
  const long long total = a.part == 0 ? nf + nbk + a.bias_floats + nimg : nf + nbk;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    if (e >= nf + nbk + a.bias_floats) {
      // bias image of the CTA-pair kernel: element (n, k) of chunk c lives at (n / 8) * 64 + (n % 8) * 8 + k
      // (8 x 16-byte core matrices, no swizzle); K columns 2j', 2j'+1 = bf16 hi / lo parts of layer 4c + j'
      const int q = (int)(e - nf - nbk - a.bias_floats);
      const int rank = q / (kBiasChunks * kBiasChunkElems), w = q % (kBiasChunks * kBiasChunkElems);
      const int c = w / kBiasChunkElems, o = w % kBiasChunkElems;
      const int n = (o / 64) * 8 + (o % 64) / 8, k = o % 8;
      const int j = c * 4 + k / 2;
      float v = 0.f;
      if (j < a.n_bias_layers) {
        const auto& L = a.layers[a.bias_layer[j]];
        const int f = rank * 128 + n;
        if (f < L.out) {
          const float b = a.params[L.boff + f];
          const float hi = __bfloat162float(__float2bfloat16(b));
          v = (k & 1) ? (b - hi) : hi;
        }
      }
      a.bias_img[q] = __float2bfloat16(v);
    } else if (e < nf) {
      const int r = (int)(e / kKP), k = (int)(e % kKP);
      float v = 0.f;
      for (int l = 0; l < a.n_layers; ++l) {
        const auto& L = a.layers[l];
        if (r < L.row0 || r >= L.row0 + L.rows_pad) continue;
        const int n = r - L.row0;
        if (n >= L.out) break;
        int i = -1;
        if (k < L.x_in) i = k;
        else if (L.feat_in > 0 && k < L.x_in + kFeatPad) {
          int fp = k - L.x_in;
          if (fp < a.feat_dim) i = L.x_in + ref_feature_col(fp, a.nb, a.ndeg);
        }
        if (i >= 0 && i < L.in) v = a.params[L.koff + (long long)i * L.out + n];
        break;
      }
      a.wt[e + a.part * nf] = pack_part(v, a.part);
    } else if (e < nf + nbk) {
      const long long q = e - nf;
      const int r = (int)(q / kW), c = (int)(q % kW);
      float v = 0.f;
      for (int l = 0; l < a.n_layers; ++l) {
        const auto& L = a.layers[l];
        if (L.brow0 < 0 || r < L.brow0 || r >= L.brow0 + kW) continue;
        const int i = r - L.brow0;
        if (c < L.out && i < L.x_in) v = a.params[L.koff + (long long)i * L.out + c];
        break;
      }
      a.wn[q + a.part * nbk] = pack_part(v, a.part);
    } else {
      const int q = (int)(e - nf - nbk);
      float v = 0.f;
      for (int l = 0; l < a.n_layers; ++l) {
        const auto& L = a.layers[l];
        if (q >= L.bias_off && q < L.bias_off + L.out) { v = a.params[L.boff + (q - L.bias_off)]; break; }
      }
      if (q >= a.w_dens_off && q < a.w_dens_off + a.dens_in)     // density kernel [in,1], bf16-rounded
        v = a.exact_heads ? a.params[a.dens_koff + (q - a.w_dens_off)]
                          : __bfloat162float(__float2bfloat16(a.params[a.dens_koff + (q - a.w_dens_off)]));
      if (a.w_rgb_off >= 0 && q >= a.w_rgb_off && q < a.w_rgb_off + a.rgb_in * 3)  // rgb kernel [in,3]
        v = a.exact_heads ? a.params[a.rgb_koff + (q - a.w_rgb_off)]
                          : __bfloat162float(__float2bfloat16(a.params[a.rgb_koff + (q - a.w_rgb_off)]));
      a.bias[q] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_map_impl(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return HUGS_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return HUGS_ERR_CUDA; }
  return HUGS_OK;
}

template <class T>
int tc_alloc(hugs_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return HUGS_ERR_NOMEM;
  }
  h->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return HUGS_OK;
}

// Build packing tables and the forward/backward layer schedules of one MLP.
int build_mlp_schedule(hugs_handle* h, const MlpViews& mv, TcMlp* m) {
  const hugs_model_desc& d = h->d;
  m->present = true; m->has_rgb = mv.has_rgb; m->depth = mv.depth;
  int row_f = 0, row_b = 0, boff = 0;
  // ---- packing ----
  for (size_t li = 0; li < mv.dense.size(); ++li) {
    const DenseView& v = mv.dense[li];
    TcMlp::PackLayer P{};
    const bool trunk = (int)li < mv.depth;
    const bool is_density = (int)li == mv.depth;
    const bool is_view = mv.has_rgb && (int)li == mv.depth + 2;
    P.out = v.out; P.in = v.in; P.koff = v.kernel_off; P.boff = v.bias_off;
    if (is_view) { P.x_in = d.bottleneck_width; P.feat_in = 0; }
    else if (v.in == h->feat_dim) { P.x_in = 0; P.feat_in = h->feat_dim; }
    else if (v.in == kW + h->feat_dim) { P.x_in = kW; P.feat_in = h->feat_dim; }
    else { P.x_in = v.in; P.feat_in = 0; }
    P.rows_pad = v.out >= 128 ? ((v.out + 127) / 128) * 128 : 16;
    P.row0 = row_f; row_f += P.rows_pad;
    P.bias_off = boff; boff += ((v.out + 3) / 4) * 4;
    // natural-layout copy for dgrad: every layer whose *input* carries a gradient
    const bool needs_dgrad = (trunk && li > 0) || (mv.has_rgb && ((int)li == mv.depth + 1 || is_view));
    P.brow0 = -1;
    if (needs_dgrad) { P.brow0 = row_b; row_b += kW; }
    (void)is_density;
    m->pack.push_back(P);
  }
  m->rows_f = row_f; m->rows_b = std::max(row_b, kW);
  // head weights for the CUDA-core parts of the backward chain
  const DenseView& dens = mv.dense[mv.depth];
  m->w_dens_off = boff; boff += ((dens.in + 3) / 4) * 4;
  m->w_rgb_off = -1;
  if (mv.has_rgb) { m->w_rgb_off = boff; boff += mv.dense[mv.depth + 3].in * 3 + 4; }
  m->bias_floats = boff;

  return pp_build(h, mv, m);
}

int fill_pack_args(hugs_handle* h, const MlpViews& mv, const TcMlp& m, const float* params, PackArgs* a) {
  memset(a, 0, sizeof(*a));
  HUGS_REQUIRE(m.pack.size() <= 16, "too many layers to pack");
  for (size_t i = 0; i < m.pack.size(); ++i) a->layers[i] = m.pack[i];
  a->n_layers = (int)m.pack.size(); a->rows_f = m.rows_f; a->rows_b = m.rows_b;
  a->nb = h->perm_nb; a->ndeg = h->d.max_deg_point - h->d.min_deg_point; a->feat_dim = h->feat_dim;
  a->bias_floats = m.bias_floats; a->params = params; a->wt = m.wt; a->wn = m.wn; a->bias = m.bias;
  a->w_dens_off = m.w_dens_off; a->w_rgb_off = m.w_rgb_off;
  a->dens_koff = mv.dense[mv.depth].kernel_off; a->dens_in = mv.dense[mv.depth].in;
  if (mv.has_rgb) { a->rgb_koff = mv.dense[mv.depth + 3].kernel_off; a->rgb_in = mv.dense[mv.depth + 3].in; }
  // biased layers of the CTA-pair kernel: trunk layer i -> j = i, bottleneck -> j = depth (see pp_build)
  a->bias_img = m.bias_img;
  a->n_bias_layers = mv.depth + (mv.has_rgb ? 1 : 0);
  HUGS_REQUIRE(a->n_bias_layers <= 4 * kBiasChunks, "too many biased layers (%d)", a->n_bias_layers);
  for (int i = 0; i < mv.depth; ++i) a->bias_layer[i] = i;
  if (mv.has_rgb) a->bias_layer[mv.depth] = mv.depth + 1;
  return HUGS_OK;
}

}  // namespace

int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
  return make_map_impl(m, base, rows, cols, box_rows);
}

// ------------------------------------------------------------------------------------------
int tc_create(hugs_handle* h) {
  const hugs_model_desc& d = h->d;
  const int ndeg = d.max_deg_point - d.min_deg_point;
  // NerfMLP: 256 wide -> chain kernel; 512 / 768 / 1024 ... -> layer-at-a-time GEMMs.  PropMLP: 256 (every shipped gin).
  const bool nerf_layered = d.nerf_width != kW;
  const bool pe = d.encoding == HUGS_ENC_POINT_PE;
  if (pe && nerf_layered) {
    set_error("point positional encoding is supported with net_width 256 on the tensor-core path (got %d)", d.nerf_width);
    return HUGS_ERR_UNSUPPORTED;
  }
  if ((nerf_layered && (d.nerf_width % 256 != 0 || d.nerf_width > 2048)) ||
      (d.num_levels > 1 && d.prop_width != kW) || d.bottleneck_width != kW ||
      d.view_width != 128 || h->feat_dim > kFeatPad || (!pe && ((ndeg % 4) != 0 || ndeg > 16))) {
    set_error("tensor-core path supports NerfMLP.net_width 256 (chain kernel) or 512 / 768 / 1024 "
              "(layer-at-a-time kernels), PropMLP.net_width 256, bottleneck 256, view width 128 and <= 512 IPE "
              "features with a degree count divisible by 4 (got widths %d/%d/%d/%d, %d features); use HUGS_PRECISION_FP32",
              d.nerf_width, d.prop_width, d.bottleneck_width, d.view_width, h->feat_dim);
    return HUGS_ERR_UNSUPPORTED;
  }
  cudaDeviceProp prop;
  HUGS_CUDA(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    set_error("tensor-core path needs an sm_100 device (found sm_%d%d)", prop.major, prop.minor);
    return HUGS_ERR_UNSUPPORTED;
  }
  TcState* tc = new TcState();
  h->tc = tc;
  tc->num_sms = prop.multiProcessorCount;
  tc->split = d.precision == HUGS_PRECISION_TC_SPLIT;
  const int parts = tc->split ? 2 : 1;       // hi (+ lo) halves of every bf16 operand tensor, stacked along the rows
  int rc;
  if (!nerf_layered && (rc = build_mlp_schedule(h, h->nerf, &tc->nerf))) return rc;
  if (d.num_levels > 1 && (rc = build_mlp_schedule(h, h->prop, &tc->prop))) return rc;
  for (TcMlp* m : {&tc->nerf, &tc->prop}) {
    if (!m->present) continue;
    if ((rc = tc_alloc(h, &m->wt, (size_t)parts * m->rows_f * kKP)) ||
        (rc = tc_alloc(h, &m->wn, (size_t)parts * m->rows_b * kW)) ||
        (rc = tc_alloc(h, &m->bias, (size_t)m->bias_floats)) ||
        (rc = tc_alloc(h, &m->bias_img, (size_t)2 * kBiasChunks * kBiasChunkElems)))
      return rc;
    if ((rc = make_map(&m->map_wt128, m->wt, (long long)parts * m->rows_f, kKP, 128)) ||
        (rc = make_map(&m->map_wn128, m->wn, (long long)parts * m->rows_b, kW, 128)) ||
        (rc = make_map(&m->map_wt64, m->wt, (long long)parts * m->rows_f, kKP, 64)))
      return rc;
  }
  // per-level feature / saved-activation regions
  const int L = d.num_levels;
  tc->cap.resize(L); tc->feat_row0.resize(L); tc->save_row0.resize(L);
  int frow = 0, srow = 0;
  for (int l = 0; l < L; ++l) {
    const int S = h->samples(l);
    tc->cap[l] = (int)((((long long)d.max_rays * S + 255) / 256) * 256);     // whole 256-row GEMM tiles
    tc->feat_row0[l] = frow; frow += tc->cap[l];
    const int n_saved = (l == L - 1) ? (nerf_layered ? 0 : d.nerf_depth + 2) : d.prop_depth;
    tc->save_row0[l] = srow; srow += n_saved * tc->cap[l];
  }
  tc->total_feat_rows = frow; tc->total_save_rows = srow;
  if ((rc = tc_alloc(h, &tc->feat, (size_t)parts * frow * kFeatPad))) return rc;
  for (int l = 0; l < L; ++l) tc->drgb_rows = std::max(tc->drgb_rows, tc->cap[l]);
  if ((rc = tc_alloc(h, &tc->viewbias, (size_t)d.max_rays * 128))) return rc;
  if ((rc = make_map(&tc->map_feat, tc->feat, (long long)parts * frow, kFeatPad, 128))) return rc;
  HUGS_CUDA(cudaMemset(tc->feat, 0, (size_t)parts * frow * kFeatPad * 2));
  if (nerf_layered && (rc = layered_create(h, h->nerf, L - 1, &tc->nerf_layered))) return rc;
  return pp_init(h);
}

// Saved activations / dZ / ReLU gates / head gradients (1,056 B per saved row: 10.6 KB per NeRF sample; twice that in
// the split-precision mode) and the weight-gradient state exist only on handles that train: allocated by the first
// hugs_loss_and_grad, so a render-only handle with a large max_rays (render_chunk_size) does not reserve tens of GB it
// never touches.
int tc_ensure_training(hugs_handle* h) {
  TcState* tc = h->tc;
  HUGS_REQUIRE(tc, "tensor-core state missing");
  if (tc->train_ready) return HUGS_OK;
  const size_t parts = tc->split ? 2 : 1;
  const size_t srow = (size_t)tc->total_save_rows;
  int rc;
  if ((rc = tc_alloc(h, &tc->act, parts * srow * kW)) || (rc = tc_alloc(h, &tc->dz, parts * srow * kW)) ||
      (rc = tc_alloc(h, &tc->gate, srow * 4)) ||
      (rc = tc_alloc(h, &tc->drgb, parts * (size_t)tc->drgb_rows * kHeadCols)))
    return rc;
  // unused columns (head gradients beyond col 3, view-activation columns 128..255) must read as zero
  HUGS_CUDA(cudaMemset(tc->gate, 0, srow * 4 * sizeof(uint2)));
  HUGS_CUDA(cudaMemset(tc->drgb, 0, parts * (size_t)tc->drgb_rows * kHeadCols * 2));
  HUGS_CUDA(cudaMemset(tc->act, 0, parts * srow * kW * 2));
  HUGS_CUDA(cudaMemset(tc->dz, 0, parts * srow * kW * 2));
  HUGS_CUDA(cudaDeviceSynchronize());      // the caller's stream may be non-blocking w.r.t. the default stream
  if ((rc = make_map(&tc->map_act, tc->act, (long long)(parts * srow), kW, 128)) ||
      (rc = make_map(&tc->map_dz, tc->dz, (long long)(parts * srow), kW, 128)))
    return rc;
  if ((rc = wgrad_create(h))) return rc;
  if (tc->nerf_layered && (rc = layered_ensure_training(h, tc->nerf_layered))) return rc;
  tc->train_ready = true;
  return HUGS_OK;
}

void tc_destroy(hugs_handle* h) {
  if (h->tc) {
    wgrad_destroy(h);
    if (h->tc->nerf_layered) layered_destroy(h->tc->nerf_layered);
    delete h->tc; h->tc = nullptr;
  }
}

int tc_pack_params(hugs_handle* h, const float* params, cudaStream_t st) {
  TcState* tc = h->tc;
  HUGS_REQUIRE(tc, "tensor-core state missing");
  PackArgs a;
  int rc;
  if (tc->nerf_layered && (rc = layered_pack(h, tc->nerf_layered, params, st))) return rc;
  for (int which = 0; which < 2; ++which) {
    const TcMlp& m = which == 0 ? tc->nerf : tc->prop;
    if (!m.present) continue;
    if ((rc = fill_pack_args(h, which == 0 ? h->nerf : h->prop, m, params, &a))) return rc;
    a.exact_heads = tc->split ? 1 : 0;
    for (int part = 0; part < (tc->split ? 2 : 1); ++part) {
      a.part = part;
      pack_params_kernel<<<512, 256, 0, st>>>(a);
      HUGS_LAUNCH_CHECK();
    }
  }
  return HUGS_OK;
}

int tc_mlp_forward(hugs_handle* h, int level, const hugs_rays* rays, int n_rays, bool training, cudaStream_t st) {
  TcState* tc = h->tc;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int S = h->samples(level);
  const int n_samples = n_rays * S;
  const bool layered = !is_prop && tc->nerf_layered != nullptr;
  // rows of the feature tensor written (zeros beyond n_samples): whole chain tiles / whole 256-row GEMM tiles
  const int n_tiles = layered ? 2 * ((n_samples + 255) / 256) : (n_samples + kTileM - 1) / kTileM;
  const int contract = is_prop ? d.prop_contract : d.nerf_contract;
  __nv_bfloat16* feat = tc->feat + (size_t)tc->feat_row0[level] * kFeatPad;
  // 1. bf16 IPE features (own column order) -> feat[level]
  {
    ProfScope ps(h, HUGS_K_ENCODE, st);
    if (d.encoding == HUGS_ENC_POINT_PE) {
      // pos_enc of the interval midpoints in the reference's arithmetic and column order, bf16 (+ residual half)
      PointPeArgs pa{rays->origins, rays->directions, h->tdist[level], n_rays, S, d.min_deg_point,
                     d.max_deg_point - d.min_deg_point, contract, nullptr, feat,
                     tc->split ? feat + (size_t)tc->total_feat_rows * kFeatPad : nullptr, kFeatPad, h->feat_panels * 64,
                     n_tiles * kTileM};
      int rc = launch_point_pe(pa, st);
      if (rc) return rc;
    } else if (tc->split) {
      // exact reference arithmetic (safe_sin quirk B12 included), split into hi / lo halves
      EncSplitArgs ea{rays->origins, rays->directions, rays->radii, h->tdist[level], h->basis, n_samples,
                      n_tiles * kTileM, S, d.num_basis, d.min_deg_point, d.max_deg_point - d.min_deg_point,
                      d.ray_shape, contract, feat, feat + (size_t)tc->total_feat_rows * kFeatPad};
      int rc = launch_encode_split(ea, st);
      if (rc) return rc;
    } else {
      EncArgs ea{rays->origins, rays->directions, rays->radii, h->tdist[level], h->basis, n_samples, n_tiles * kTileM,
                 S, d.num_basis, d.min_deg_point, d.max_deg_point - d.min_deg_point, d.ray_shape, contract, feat};
      const int eg = (n_tiles * kTileM + kEncRows - 1) / kEncRows;
      if (ea.ndeg == 12) encode_bf16_kernel<12><<<eg, kEncSplit * kEncRows, 0, st>>>(ea);
      else encode_bf16_kernel<0><<<eg, kEncSplit * kEncRows, 0, st>>>(ea);
      HUGS_LAUNCH_CHECK();
    }
  }
  // 2. per-ray view bias (direction encoding + GLO folded through the view layer)
  if (!is_prop) {
    const DenseView& vv = mv.dense[mv.depth + 2];
    viewbias_kernel<<<(n_rays * 128 + 255) / 256, 256, 0, st>>>(h->view_in, h->view_in_dim, h->cur_params,
                                                               vv.kernel_off, vv.bias_off, d.bottleneck_width,
                                                               128, n_rays, tc->split ? 1 : 0, tc->viewbias);
    HUGS_LAUNCH_CHECK();
  }
  // 3. fused chain (or one GEMM per layer)
  ProfScope ps(h, is_prop ? HUGS_K_CHAIN_FWD_PROP : HUGS_K_CHAIN_FWD_NERF, st);
  if (layered) return layered_forward(h, tc->nerf_layered, level, n_rays, training, st);
  return pp_launch(h, level, n_rays, training ? 1 : 0, st);
}

int tc_mlp_backward(hugs_handle* h, int level, const hugs_rays* rays, int n_rays, float* grad, cudaStream_t st) {
  (void)rays;
  const bool is_prop = level < h->d.num_levels - 1;
  if (!is_prop && h->tc->nerf_layered) return layered_backward(h, h->tc->nerf_layered, level, n_rays, grad, st);
  {
    ProfScope ps(h, is_prop ? HUGS_K_CHAIN_BWD_PROP : HUGS_K_CHAIN_BWD_NERF, st);
    int rc = pp_launch(h, level, n_rays, 2, st);
    if (rc) return rc;
  }
  return wgrad_run(h, level, n_rays, grad, st);
}

// development / test hook: the throughput-mode bf16 encoder on its own (tests/test_gpu_ref_golden.py)
int tc_debug_encode(hugs_handle* h, const hugs_rays* rays, const float* tdist, int n_rays, int S, int contract,
                    __nv_bfloat16* out, cudaStream_t st) {
  const hugs_model_desc& d = h->d;
  const int n_samples = n_rays * S;
  EncArgs ea{rays->origins, rays->directions, rays->radii, tdist, h->basis, n_samples, n_samples, S, d.num_basis,
             d.min_deg_point, d.max_deg_point - d.min_deg_point, d.ray_shape, contract, out};
  const int eg = (n_samples + kEncRows - 1) / kEncRows;
  if (ea.ndeg == 12) encode_bf16_kernel<12><<<eg, kEncSplit * kEncRows, 0, st>>>(ea);
  else encode_bf16_kernel<0><<<eg, kEncSplit * kEncRows, 0, st>>>(ea);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs
