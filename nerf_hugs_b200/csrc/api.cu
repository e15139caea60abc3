// C ABI of libhugs_b200.so (see include/hugs_b200.h).  Host-side orchestration only: every
// numeric step is a kernel in sampling.cu / composite.cu / mlp_simt.cu / mlp_tc.cu / optim.cu.
#include <stdarg.h>
#include <math.h>

#include <algorithm>
#include <new>

#include "handle.h"
#include "tc.h"

namespace hugs {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %s (%s) at %s:%d in `%s`", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
  return HUGS_ERR_CUDA;
}

namespace {

template <class T>
int dev_alloc(hugs_handle* h, T** p, size_t count) {
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

void add_tensor(hugs_handle* h, const std::string& name, int rows, int cols, int module, int64_t* off) {
  hugs_tensor_desc t{};
  snprintf(t.name, sizeof(t.name), "%s", name.c_str());
  t.offset = *off; t.rows = rows; t.cols = cols; t.module = module;
  h->tensors.push_back(t);
  *off += (int64_t)rows * cols;
}

// Dense shapes in flax creation order (models.py:449-515).
void build_mlp(hugs_handle* h, MlpViews* m, const char* mod, int module, int depth, int width, bool rgb,
               int glo, int64_t* off) {
  const hugs_model_desc& d = h->d;
  m->depth = depth; m->width = width; m->has_rgb = rgb; m->module = module;
  auto add = [&](int in, int out) {
    DenseView v{};
    v.in = in; v.out = out;
    std::string base = std::string(mod) + "/Dense_" + std::to_string(m->dense.size());
    v.kernel_off = *off; add_tensor(h, base + "/kernel", in, out, module, off);
    v.bias_off = *off;   add_tensor(h, base + "/bias", 1, out, module, off);
    m->dense.push_back(v);
  };
  int in = h->feat_dim;
  for (int i = 0; i < depth; ++i) {
    add(in, width);
    in = width;
    if (i % d.skip_layer == 0 && i > 0) in = width + h->feat_dim;
  }
  add(in, 1);
  if (rgb) {
    add(in, d.bottleneck_width);
    add(d.bottleneck_width + 3 + 6 * d.deg_view + glo, d.view_width);
    add(d.view_width, 3);
  }
}

int validate(const hugs_model_desc& d) {
  HUGS_REQUIRE(d.num_levels >= 1 && d.num_levels <= 4, "num_levels must be in [1,4], got %d", d.num_levels);
  HUGS_REQUIRE(d.num_nerf_samples >= 2 && d.num_nerf_samples <= 256, "num_nerf_samples must be in [2,256]");
  HUGS_REQUIRE(d.num_levels == 1 || (d.num_prop_samples >= 2 && d.num_prop_samples <= 256),
               "num_prop_samples must be in [2,256]");
  HUGS_REQUIRE(d.num_basis >= 1 && d.num_basis <= 32, "num_basis must be in [1,32]");
  HUGS_REQUIRE(d.encoding == HUGS_ENC_IPE || d.encoding == HUGS_ENC_POINT_PE, "unknown encoding %d", d.encoding);
  HUGS_REQUIRE(d.max_deg_point > d.min_deg_point && d.max_deg_point - d.min_deg_point <= (d.encoding == HUGS_ENC_IPE ? 16 : 40),
               "bad positional-encoding degrees");
  HUGS_REQUIRE(d.nerf_depth >= 1 && d.prop_depth >= 1 && d.nerf_width >= 1 && d.prop_width >= 1, "bad MLP shape");
  HUGS_REQUIRE(d.skip_layer >= 1, "skip_layer must be >= 1");
  HUGS_REQUIRE(d.deg_view >= 0 && d.deg_view <= 8, "deg_view must be in [0,8]");
  HUGS_REQUIRE(d.bottleneck_width >= 1 && d.view_width >= 1, "bottleneck/view widths must be >= 1");
  HUGS_REQUIRE(d.num_glo_features >= 0 && d.num_glo_features <= 64, "num_glo_features must be in [0,64]");
  HUGS_REQUIRE(d.num_glo_features == 0 || d.num_embeddings > 0, "num_embeddings must be > 0 with GLO");
  HUGS_REQUIRE(d.max_rays >= 1, "max_rays must be >= 1");
  HUGS_REQUIRE(d.raydist_fn >= 0 && d.raydist_fn <= 3, "unknown raydist_fn %d", d.raydist_fn);
  HUGS_REQUIRE(d.ray_shape == HUGS_RAY_CONE || d.ray_shape == HUGS_RAY_CYLINDER, "unknown ray_shape");
  HUGS_REQUIRE(d.precision == HUGS_PRECISION_FP32 || d.precision == HUGS_PRECISION_BF16_TC ||
               d.precision == HUGS_PRECISION_TC_SPLIT, "unknown precision");
  return HUGS_OK;
}

// The `u` grids of stepfun.sample (stepfun.py:188-209): numpy-style float64 linspace
// (i * step + start, last point exact) rounded once to fp32 — identical to oracle.sample_u.
void linspace_f32(double lo, double hi, int n, std::vector<float>* u) {
  u->resize(n);
  const double step = n > 1 ? (hi - lo) / (n - 1) : 0.0;
  for (int i = 0; i < n; ++i) (*u)[i] = (float)(i * step + lo);
  if (n > 1) (*u)[n - 1] = (float)hi;
}

void make_u(int ns, bool train, std::vector<float>* u, float* max_jitter) {
  const double eps = (double)kF32Eps;
  if (!train) {
    const double pad = 1.0 / (2.0 * ns);
    linspace_f32(pad, 1.0 - pad - eps, ns, u);
    *max_jitter = 0.f;
  } else {
    const double u_max = eps + (1.0 - eps) / ns;
    *max_jitter = (float)((1.0 - u_max) / (ns - 1) - eps);
    linspace_f32(0.0, 1.0 - u_max, ns, u);
  }
}

}  // namespace
}  // namespace hugs

using namespace hugs;

HUGS_API const char* hugs_last_error(void) { return g_err; }
HUGS_API int hugs_abi_version(void) { return HUGS_ABI_VERSION; }
HUGS_API int64_t hugs_launch_count(void) { return g_launch_count; }

HUGS_API int hugs_profile_enable(hugs_handle* h, int32_t enable) {
  HUGS_REQUIRE(h, "hugs_profile_enable: null handle");
  h->prof_on = enable != 0;
  return HUGS_OK;
}

HUGS_API int hugs_profile_read(hugs_handle* h, float* ms_out, int32_t* count_out) {
  HUGS_REQUIRE(h && ms_out && count_out, "hugs_profile_read: null argument");
  HUGS_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < HUGS_K_COUNT; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
  for (auto& r : h->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.cls >= 0 && r.cls < HUGS_K_COUNT) {
      ms_out[r.cls] += ms; count_out[r.cls] += 1;
    }
    h->prof_pool.push_back(r.a); h->prof_pool.push_back(r.b);
  }
  h->prof_recs.clear();
  return HUGS_OK;
}

HUGS_API int hugs_create(const hugs_model_desc* desc, hugs_handle** out) {
  HUGS_REQUIRE(desc && out, "hugs_create: null argument");
  *out = nullptr;
  int rc = validate(*desc);
  if (rc) return rc;
  int ndev = 0;
  HUGS_CUDA(cudaGetDeviceCount(&ndev));
  HUGS_REQUIRE(ndev > 0, "no CUDA device: this library has no CPU path");
  hugs_handle* h = new (std::nothrow) hugs_handle();
  if (!h) { set_error("out of host memory"); return HUGS_ERR_NOMEM; }
  h->d = *desc;
  const hugs_model_desc& d = h->d;
  rc = HUGS_OK;
  auto fail = [&](int code) { hugs_destroy(h); return code; };
  if (cudaGetDevice(&h->device) != cudaSuccess) return fail(HUGS_ERR_CUDA);
  if (d.encoding == HUGS_ENC_POINT_PE) {
    h->feat_dim = 3 + 6 * (d.max_deg_point - d.min_deg_point);    // pos_enc(..., append_identity=True)
    h->perm_nb = 0;
  } else {
    h->feat_dim = 2 * d.num_basis * (d.max_deg_point - d.min_deg_point);
    h->perm_nb = d.num_basis;
  }
  h->feat_panels = (h->feat_dim + 63) / 64;
  h->max_nerf_samples = d.num_nerf_samples;
  h->view_in_dim = 3 + 6 * d.deg_view + d.num_glo_features;

  int64_t off = 0;
  h->module_begin[0] = off;
  build_mlp(h, &h->nerf, "NerfMLP_0", 0, d.nerf_depth, d.nerf_width, true, d.num_glo_features, &off);
  h->module_end[0] = off;
  h->module_begin[1] = off;
  if (d.num_levels > 1) build_mlp(h, &h->prop, "PropMLP_0", 1, d.prop_depth, d.prop_width, false, 0, &off);
  h->module_end[1] = off;
  h->module_begin[2] = off;
  if (d.num_glo_features > 0) {
    h->glo_off = off;
    add_tensor(h, "GloEmbed_0/embedding", d.num_embeddings, d.num_glo_features, 2, &off);
  }
  h->module_end[2] = off;
  h->n_params = off;

  const int L = d.num_levels;
  const size_t n = (size_t)d.max_rays;
  if ((rc = dev_alloc(h, &h->basis, 3 * 32))) return fail(rc);
  if (cudaMemcpy(h->basis, d.basis, sizeof(float) * 3 * d.num_basis, cudaMemcpyHostToDevice) != cudaSuccess)
    return fail(cuda_fail(cudaGetLastError(), "basis upload", __FILE__, __LINE__));
  h->u_det.resize(L); h->u_train.resize(L); h->max_jitter.resize(L);
  h->sdist.resize(L); h->tdist.resize(L); h->weights.resize(L); h->raw.resize(L); h->d_raw.resize(L);
  int smax = 0;
  for (int l = 0; l < L; ++l) {
    const int S = h->samples(l);
    smax = std::max(smax, S);
    std::vector<float> u;
    float mj;
    if ((rc = dev_alloc(h, &h->u_det[l], S)) || (rc = dev_alloc(h, &h->u_train[l], S))) return fail(rc);
    make_u(S, false, &u, &mj);
    cudaMemcpy(h->u_det[l], u.data(), sizeof(float) * S, cudaMemcpyHostToDevice);
    make_u(S, true, &u, &mj);
    cudaMemcpy(h->u_train[l], u.data(), sizeof(float) * S, cudaMemcpyHostToDevice);
    h->max_jitter[l] = mj;
    const int c = (l == L - 1) ? 4 : 1;
    if ((rc = dev_alloc(h, &h->sdist[l], n * (S + 1))) || (rc = dev_alloc(h, &h->tdist[l], n * (S + 1))) ||
        (rc = dev_alloc(h, &h->weights[l], n * S)) || (rc = dev_alloc(h, &h->raw[l], n * S * c)) ||
        (rc = dev_alloc(h, &h->d_raw[l], n * S * c)))
      return fail(rc);
  }
  if ((rc = dev_alloc(h, &h->view_in, n * h->view_in_dim))) return fail(rc);
  if ((rc = dev_alloc(h, &h->ray_stats, n * 12))) return fail(rc);
  {
    std::vector<int64_t> ends;
    for (const auto& t : h->tensors) ends.push_back(t.offset + (int64_t)t.rows * t.cols);
    if ((rc = dev_alloc(h, &h->tensor_ends, ends.size()))) return fail(rc);
    if (cudaMemcpy(h->tensor_ends, ends.data(), sizeof(int64_t) * ends.size(), cudaMemcpyHostToDevice) != cudaSuccess)
      return fail(cuda_fail(cudaGetLastError(), "tensor table upload", __FILE__, __LINE__));
  }
  if ((rc = dev_alloc(h, &h->scalars, 2048))) return fail(rc);
  if (d.precision == HUGS_PRECISION_FP32) {
    const int wmax = std::max({d.nerf_width, d.prop_width, d.bottleneck_width, d.view_width});
    if ((rc = dev_alloc(h, &h->feat, n * smax * h->feat_dim))) return fail(rc);
    for (int i = 0; i < 3; ++i)
      if ((rc = dev_alloc(h, &h->act[i], n * smax * wmax))) return fail(rc);
  } else {
    if ((rc = tc_create(h))) return fail(rc);
  }
  if (cudaDeviceSynchronize() != cudaSuccess)
    return fail(cuda_fail(cudaGetLastError(), "hugs_create sync", __FILE__, __LINE__));
  *out = h;
  return HUGS_OK;
}

HUGS_API int hugs_destroy(hugs_handle* h) {
  if (!h) return HUGS_OK;
  tc_destroy(h);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return HUGS_OK;
}

HUGS_API int64_t hugs_param_count(const hugs_handle* h) { return h ? h->n_params : -1; }

HUGS_API int hugs_param_layout(const hugs_handle* h, hugs_tensor_desc* out, int32_t capacity, int32_t* count) {
  HUGS_REQUIRE(h && count, "hugs_param_layout: null argument");
  *count = (int32_t)h->tensors.size();
  if (!out) return HUGS_OK;
  HUGS_REQUIRE(capacity >= *count, "hugs_param_layout: capacity %d < %d tensors", capacity, *count);
  memcpy(out, h->tensors.data(), sizeof(hugs_tensor_desc) * h->tensors.size());
  return HUGS_OK;
}

HUGS_API int hugs_params_changed(hugs_handle* h, const float* params, void* stream) {
  HUGS_REQUIRE(h && params, "hugs_params_changed: null argument");
  if (h->d.precision != HUGS_PRECISION_FP32) return tc_pack_params(h, params, (cudaStream_t)stream);
  return HUGS_OK;
}

// ------------------------------------------------------------------ operator-level entry points

HUGS_API int hugs_sample_intervals(const float* t, const float* w_logits, const float* u_base,
                                   const float* jitter, float max_jitter, int32_t n_rays, int32_t n_bins,
                                   int32_t n_samples, float dom_lo, float dom_hi, float* t_out,
                                   int32_t* idx_out, void* stream) {
  HUGS_REQUIRE(n_samples > 1, "num_samples must be > 1, is %d.", n_samples);   // stepfun.py:240-241
  HUGS_REQUIRE(n_bins >= 1 && n_rays >= 0, "hugs_sample_intervals: bad sizes");
  if (n_rays == 0) return HUGS_OK;                                             // empty batch: nothing to do
  HUGS_REQUIRE(t && w_logits && u_base && t_out, "hugs_sample_intervals: null argument");
  ResampleArgs a;
  a.t_in = t; a.w_in = w_logits; a.w_is_logits = 1; a.n_rays = n_rays; a.np = n_bins; a.ns = n_samples;
  a.dom_lo = dom_lo; a.dom_hi = dom_hi; a.u_base = u_base; a.jitter = jitter; a.max_jitter = max_jitter;
  a.s_out = t_out; a.idx_out = idx_out;
  return launch_resample(a, (cudaStream_t)stream);
}

HUGS_API int hugs_invert_cdf(const float* t, const float* cw, const float* u, int32_t n_rays, int32_t n_bins,
                             int32_t n_samples, float* centers_out, int32_t* idx_out, void* stream) {
  HUGS_REQUIRE(t && cw && u && centers_out, "hugs_invert_cdf: null argument");
  HUGS_REQUIRE(n_bins >= 1 && n_samples >= 1 && n_rays >= 0, "hugs_invert_cdf: bad sizes");
  ResampleArgs a;
  a.t_in = t; a.w_in = t; a.cw_in = cw; a.u_in = u; a.n_rays = n_rays; a.np = n_bins; a.ns = n_samples;
  a.centers_out = centers_out; a.idx_out = idx_out;
  return launch_resample(a, (cudaStream_t)stream);
}

HUGS_API int hugs_max_dilate_weights(const float* t, const float* w, int32_t n_rays, int32_t n_bins,
                                     float dilation, float dom_lo, float dom_hi, float* t_out, float* w_out,
                                     void* stream) {
  HUGS_REQUIRE(t && w && t_out && w_out, "hugs_max_dilate_weights: null argument");
  HUGS_REQUIRE(n_bins >= 1 && n_rays >= 0, "hugs_max_dilate_weights: bad sizes");
  ResampleArgs a;
  a.t_in = t; a.w_in = w; a.n_rays = n_rays; a.np = n_bins; a.ns = 0; a.dilate = 1; a.dilation = dilation;
  a.dom_lo = dom_lo; a.dom_hi = dom_hi; a.td_out = t_out; a.wd_out = w_out;
  return launch_resample(a, (cudaStream_t)stream);
}

HUGS_API int hugs_alpha_composite(const hugs_handle* h, const float* raw_density, const float* raw_rgb,
                                  const float* tdist, const float* sdist, const float* directions,
                                  const float* far, int32_t n_rays, int32_t n_samples, int32_t compute_extras,
                                  const hugs_level_out* out, void* stream) {
  (void)sdist;
  HUGS_REQUIRE(h && raw_density && tdist && directions && out, "hugs_alpha_composite: null argument");
  HUGS_REQUIRE(!compute_extras || far, "hugs_alpha_composite: far is required with compute_extras");
  CompositeArgs a;
  a.raw_density = raw_density; a.raw_rgb = raw_rgb; a.raw_stride = 1; a.rgb_stride = 3;
  a.tdist = tdist; a.directions = directions; a.far = far; a.n_rays = n_rays; a.S = n_samples;
  a.opaque_background = h->d.opaque_background; a.compute_extras = compute_extras; a.bg = h->d.bg_intensity;
  a.density_bias = h->d.density_bias; a.rgb_premult = h->d.rgb_premultiplier; a.rgb_bias = h->d.rgb_bias;
  a.rgb_padding = h->d.rgb_padding; a.out = *out;
  return launch_composite(a, (cudaStream_t)stream);
}

HUGS_API int hugs_ipe_features(const hugs_handle* h, const hugs_rays* rays, const float* tdist, int32_t n_rays,
                               int32_t n_samples, int32_t contract, float* features, void* stream) {
  HUGS_REQUIRE(h && rays && tdist && features, "hugs_ipe_features: null argument");
  if (h->d.encoding == HUGS_ENC_POINT_PE) {
    PointPeArgs pa{rays->origins, rays->directions, tdist, n_rays, n_samples, h->d.min_deg_point,
                   h->d.max_deg_point - h->d.min_deg_point, contract, features, nullptr, nullptr, 0, 0, 0};
    return launch_point_pe(pa, (cudaStream_t)stream);
  }
  IpeArgs a{rays->origins, rays->directions, rays->radii, tdist, h->basis, n_rays, n_samples,
            h->d.num_basis, h->d.min_deg_point, h->d.max_deg_point, h->d.ray_shape, contract, features};
  return launch_ipe_features(a, (cudaStream_t)stream);
}

HUGS_API int hugs_debug_encode_bf16(hugs_handle* h, const hugs_rays* rays, const float* tdist, int32_t n_rays,
                                    int32_t n_samples, int32_t contract, void* features_bf16, void* stream) {
  HUGS_REQUIRE(h && rays && tdist && features_bf16, "hugs_debug_encode_bf16: null argument");
  HUGS_REQUIRE(h->d.precision != HUGS_PRECISION_FP32, "hugs_debug_encode_bf16 needs a tensor-core handle");
  return tc_debug_encode(h, rays, tdist, n_rays, n_samples, contract, static_cast<__nv_bfloat16*>(features_bf16),
                         (cudaStream_t)stream);
}

// ------------------------------------------------------------------ model-level entry points
namespace hugs {
namespace {

int run_level_sampling(hugs_handle* h, int l, const hugs_rays* rays, int n, float train_frac,
                       const float* jitter, uint64_t rng_key, float s_near, float* prod_samples, cudaStream_t st) {
  const hugs_model_desc& d = h->d;
  const int S = h->samples(l);
  ResampleArgs a;
  a.n_rays = n; a.ns = S; a.dom_lo = s_near; a.dom_hi = 1.f;
  const float dilation = d.dilation_bias + d.dilation_multiplier * (1.f - s_near) / *prod_samples;
  *prod_samples *= (float)S;
  if (l > 0) {
    a.t_in = h->sdist[l - 1]; a.w_in = h->weights[l - 1]; a.np = h->samples(l - 1);
    a.dilate = (d.dilation_bias > 0 || d.dilation_multiplier > 0) ? 1 : 0;
    a.dilation = dilation;
  } else {
    a.np = 1;
  }
  a.anneal = d.anneal_slope > 0 ? (d.anneal_slope * train_frac) / ((d.anneal_slope - 1.f) * train_frac + 1.f) : 1.f;
  a.padding = d.resample_padding;
  if (jitter) { a.u_base = h->u_train[l]; a.jitter = jitter + (size_t)l * n; a.max_jitter = h->max_jitter[l]; }
  else if (rng_key) {
    a.u_base = h->u_train[l]; a.max_jitter = h->max_jitter[l];
    a.jitter_key = rng_key * 0xD1342543DE82EF95ull + (uint64_t)(l + 1);      // one independent stream per level
    if (a.jitter_key == 0) a.jitter_key = 1;
  } else { a.u_base = h->u_det[l]; }
  a.s_out = h->sdist[l]; a.t_out = h->tdist[l];
  a.raydist_fn = d.raydist_fn; a.near = rays->near; a.far = rays->far;
  ProfScope ps(h, HUGS_K_SAMPLE, st);
  return launch_resample(a, st);
}

// fp32 CUDA-core MLP for level l: fills h->raw[l].
int run_level_mlp_fp32(hugs_handle* h, int l, const float* params, const hugs_rays* rays, int n, cudaStream_t st) {
  const hugs_model_desc& d = h->d;
  const bool is_prop = l < d.num_levels - 1;
  const MlpViews& m = is_prop ? h->prop : h->nerf;
  const int S = h->samples(l);
  const int M = n * S;
  ProfScope ps(h, HUGS_K_MLP_FP32, st);
  int rc;
  if (d.encoding == HUGS_ENC_POINT_PE) {
    PointPeArgs pa{rays->origins, rays->directions, h->tdist[l], n, S, d.min_deg_point, d.max_deg_point - d.min_deg_point,
                   is_prop ? d.prop_contract : d.nerf_contract, h->feat, nullptr, nullptr, 0, 0, 0};
    rc = launch_point_pe(pa, st);
  } else {
    IpeArgs ia{rays->origins, rays->directions, rays->radii, h->tdist[l], h->basis, n, S, d.num_basis,
               d.min_deg_point, d.max_deg_point, d.ray_shape, is_prop ? d.prop_contract : d.nerf_contract, h->feat};
    rc = launch_ipe_features(ia, st);
  }
  if (rc) return rc;
  const float* x = h->feat;
  int xk = h->feat_dim;
  bool cat = false;
  int li = 0, buf = 0;
  auto dense = [&](const DenseView& v, bool relu, float* y, int ldy) {
    DenseArgs a{};
    a.nseg = 0;
    a.seg[a.nseg++] = DenseSeg{x, xk, xk, 1};
    if (cat) a.seg[a.nseg++] = DenseSeg{h->feat, h->feat_dim, h->feat_dim, 1};
    a.W = params + v.kernel_off; a.bias = params + v.bias_off; a.M = M; a.N = v.out; a.relu = relu;
    a.y = y; a.ldy = ldy;
    return launch_dense_simt(a, st);
  };
  for (int i = 0; i < m.depth; ++i) {
    float* y = h->act[buf];
    if ((rc = dense(m.dense[li++], true, y, m.width))) return rc;
    x = y; xk = m.width; buf ^= 1;
    cat = (i % d.skip_layer == 0 && i > 0);
  }
  const int c = is_prop ? 1 : 4;
  if ((rc = dense(m.dense[li++], false, h->raw[l], c))) return rc;       // density head
  if (!is_prop) {
    float* bott = h->act[buf];
    if ((rc = dense(m.dense[li++], false, bott, d.bottleneck_width))) return rc;
    DenseArgs a{};
    const DenseView& vv = m.dense[li++];
    a.nseg = 2;
    a.seg[0] = DenseSeg{bott, d.bottleneck_width, d.bottleneck_width, 1};
    a.seg[1] = DenseSeg{h->view_in, h->view_in_dim, h->view_in_dim, S};
    a.W = params + vv.kernel_off; a.bias = params + vv.bias_off; a.M = M; a.N = vv.out; a.relu = 1;
    a.y = h->act[2]; a.ldy = d.view_width;
    if ((rc = launch_dense_simt(a, st))) return rc;
    x = h->act[2]; xk = d.view_width; cat = false;
    if ((rc = dense(m.dense[li++], false, h->raw[l] + 1, 4))) return rc;  // rgb head
  }
  return HUGS_OK;
}

int check_rays(const hugs_handle* h, const hugs_rays* rays, int n, bool field_only = false) {
  HUGS_REQUIRE(rays && rays->origins && rays->directions && rays->viewdirs, "rays: origins/directions/viewdirs are required");
  HUGS_REQUIRE(field_only || (rays->near && rays->far), "rays: near/far are required");
  HUGS_REQUIRE(h->d.encoding == HUGS_ENC_POINT_PE || rays->radii, "rays: radii are required by the integrated encoding");
  HUGS_REQUIRE(n >= 0 && n <= h->d.max_rays, "n_rays %d exceeds max_rays %d of this handle", n, h->d.max_rays);
  HUGS_REQUIRE(h->d.num_glo_features == 0 || rays->embed_idx, "rays: embed_idx is required with GLO features");
  return HUGS_OK;
}

float s_near_of(const hugs_model_desc& d, float train_frac) {
  if (!(d.near_anneal_rate > 0.f)) return 0.f;
  return fminf(fmaxf(1.f - train_frac / d.near_anneal_rate, 0.f), d.near_anneal_init);
}

// Levels 0..L-1 forward: sampling -> MLP -> (composite: weights for the next level + outputs).
int forward_levels(hugs_handle* h, const float* params, const hugs_rays* rays, int n, float train_frac,
                   const float* jitter, int compute_extras, int zero_glo, const hugs_level_out* out,
                   bool training, cudaStream_t st) {
  // training without a jitter tensor: draws come from the handle's counter-based stream (hugs_set_train_rng), if any
  const uint64_t rng_key = (training && !jitter) ? h->train_rng_key : 0;
  const hugs_model_desc& d = h->d;
  int rc;
  const float s_near = s_near_of(d, train_frac);
  float prod = 1.f;
  h->cur_params = params;
  h->cur_embed_idx = rays->embed_idx;
  if ((rc = launch_view_inputs(rays->viewdirs, rays->embed_idx, h->glo_off >= 0 ? params + h->glo_off : nullptr,
                               n, d.deg_view, d.num_glo_features, zero_glo, d.num_embeddings, h->view_in, st)))
    return rc;
  for (int l = 0; l < d.num_levels; ++l) {
    const bool is_prop = l < d.num_levels - 1;
    const int S = h->samples(l);
    if ((rc = run_level_sampling(h, l, rays, n, train_frac, jitter, rng_key, s_near, &prod, st))) return rc;
    if (d.precision == HUGS_PRECISION_FP32) rc = run_level_mlp_fp32(h, l, params, rays, n, st);
    else rc = tc_mlp_forward(h, l, rays, n, training, st);
    if (rc) return rc;
    // the final level's compositing is folded into the loss kernel when training
    if (training && !is_prop) break;
    CompositeArgs a;
    a.raw_density = h->raw[l]; a.raw_stride = is_prop ? 1 : 4;
    a.raw_rgb = is_prop ? nullptr : h->raw[l] + 1; a.rgb_stride = 4;
    a.tdist = h->tdist[l]; a.directions = rays->directions; a.far = rays->far; a.n_rays = n; a.S = S;
    a.opaque_background = d.opaque_background; a.compute_extras = compute_extras; a.bg = d.bg_intensity;
    a.density_bias = d.density_bias; a.rgb_premult = d.rgb_premultiplier; a.rgb_bias = d.rgb_bias;
    a.rgb_padding = d.rgb_padding;
    if (out) a.out = out[l];
    float* user_w = a.out.weights;
    a.out.weights = h->weights[l];
    {
      ProfScope ps(h, HUGS_K_COMPOSITE_LOSS, st);
      if ((rc = launch_composite(a, st))) return rc;
    }
    if (out) {
      if (user_w) HUGS_CUDA(cudaMemcpyAsync(user_w, h->weights[l], sizeof(float) * n * S, cudaMemcpyDeviceToDevice, st));
      if (out[l].sdist) HUGS_CUDA(cudaMemcpyAsync(out[l].sdist, h->sdist[l], sizeof(float) * n * (S + 1), cudaMemcpyDeviceToDevice, st));
    }
  }
  return HUGS_OK;
}

}  // namespace
}  // namespace hugs

HUGS_API int hugs_forward(hugs_handle* h, const float* params, const hugs_rays* rays, int32_t n_rays,
                          float train_frac, const float* jitter, int32_t compute_extras, int32_t zero_glo,
                          const hugs_level_out* out, void* stream) {
  HUGS_REQUIRE(h && params && out, "hugs_forward: null argument");
  int rc = check_rays(h, rays, n_rays);
  if (rc) return rc;
  if (n_rays == 0) return HUGS_OK;
  return forward_levels(h, params, rays, n_rays, train_frac, jitter, compute_extras, zero_glo, out, false,
                        (cudaStream_t)stream);
}

HUGS_API int hugs_render_frame(hugs_handle* h, const float* params, const hugs_camera_set* cams, int32_t cam, int32_t width,
                               int32_t height, int32_t row0, int32_t row1, float train_frac, int32_t zero_glo,
                               const hugs_frame_out* out, void* stream) {
  HUGS_REQUIRE(h && params && cams && out && out->rgb, "hugs_render_frame: null argument (rgb is required)");
  HUGS_REQUIRE(cams->pixtocams && cams->camtoworlds && cams->heights && cams->widths && cams->pixel_offset,
               "camera set needs pixtocams, camtoworlds, heights, widths and pixel_offset");
  HUGS_REQUIRE(cams->camtype == 0 || cams->camtype == 1, "camtype must be 0 (perspective) or 1 (fisheye), got %d", cams->camtype);
  HUGS_REQUIRE(cam >= 0 && width >= 1 && height >= 1 && row0 >= 0 && row0 <= row1 && row1 <= height,
               "hugs_render_frame: bad stripe rows [%d, %d) of a %d x %d image", row0, row1, width, height);
  HUGS_REQUIRE(!out->sse || cams->images || cams->images_u8, "hugs_render_frame: sse requested but the camera set holds no images");
  cudaStream_t st = (cudaStream_t)stream;
  const int L = h->d.num_levels, chunk = h->d.max_rays;
  const long long n = (long long)(row1 - row0) * width;
  if (n == 0) return HUGS_OK;
  int rc;
  if (!h->frame_ws && (rc = dev_alloc(h, &h->frame_ws, (size_t)chunk * 15))) return rc;
  float* w = h->frame_ws;
  hugs_ray_batch rb{};
  rb.origins = w; rb.directions = w + (size_t)chunk * 3; rb.viewdirs = w + (size_t)chunk * 6;
  rb.radii = w + (size_t)chunk * 9; rb.near = w + (size_t)chunk * 10; rb.far = w + (size_t)chunk * 11;
  rb.lossmult = w + (size_t)chunk * 12; rb.static_mask = w + (size_t)chunk * 13;
  rb.embed_idx = reinterpret_cast<int32_t*>(w + (size_t)chunk * 14);
  hugs_rays rays{};
  rays.origins = rb.origins; rays.directions = rb.directions; rays.viewdirs = rb.viewdirs; rays.radii = rb.radii;
  rays.near = rb.near; rays.far = rb.far; rays.lossmult = rb.lossmult; rays.embed_idx = rb.embed_idx;
  const int extras = (out->distance_mean || out->distance_median) ? 1 : 0;
  std::vector<hugs_level_out> outs((size_t)L);
  for (long long p0 = 0; p0 < n; p0 += chunk) {
    const int m = (int)std::min<long long>(chunk, n - p0);
    if ((rc = launch_frame_rays(*cams, cam, width, (long long)row0 * width + p0, m, rb, st))) return rc;
    hugs_level_out& o = outs[(size_t)L - 1];
    o = hugs_level_out{};
    o.rgb = out->rgb + p0 * 3;
    if (out->acc) o.acc = out->acc + p0;
    if (out->distance_mean) o.distance_mean = out->distance_mean + p0;
    if (out->distance_median) o.distance_median = out->distance_median + p0;
    if ((rc = forward_levels(h, params, &rays, m, train_frac, nullptr, extras, zero_glo, outs.data(), false, st))) return rc;
  }
  return launch_frame_finish(out->rgb, n * 3, *cams, cam, (long long)row0 * width * 3, out->rgb_u8, out->sse, st);
}

HUGS_API int hugs_loss_and_grad(hugs_handle* h, const float* params, const hugs_rays* rays, const float* rgb_gt,
                                int32_t n_rays, float train_frac, const float* jitter, const hugs_loss_cfg* loss,
                                float* grad_out, float* stats_out, void* stream) {
  HUGS_REQUIRE(h && params && rgb_gt && loss && grad_out && stats_out, "hugs_loss_and_grad: null argument");
  int rc = check_rays(h, rays, n_rays);
  if (rc) return rc;
  HUGS_REQUIRE(n_rays > 0, "hugs_loss_and_grad: empty batch");
  const hugs_model_desc& d = h->d;
  if (d.precision == HUGS_PRECISION_FP32) {
    set_error("hugs_loss_and_grad needs a tensor-core precision mode (HUGS_PRECISION_BF16_TC for throughput, "
              "HUGS_PRECISION_TC_SPLIT for fp32-level parity); the fp32 CUDA-core path is render-only");
    return HUGS_ERR_UNSUPPORTED;
  }
  if (loss->data_coarse_loss_mult != 0.f) {
    set_error("data_coarse_loss_mult != 0 is not supported (shipped configs use 0, configs.py:88)");
    return HUGS_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int L = d.num_levels, n = n_rays;
  if ((rc = tc_ensure_training(h))) return rc;
  if ((rc = forward_levels(h, params, rays, n, train_frac, jitter, 0, 0, nullptr, true, st))) return rc;

  float* denom = h->scalars;          // [0] loss normaliser
  ProfScope* ps_loss = new ProfScope(h, HUGS_K_COMPOSITE_LOSS, st);
  struct Guard { ProfScope** p; ~Guard() { delete *p; *p = nullptr; } } guard{&ps_loss};
  if ((rc = launch_lossmult_sum(rays->lossmult, rays->static_mask, loss->use_static_mask,
                                loss->withmask_transient_weight, loss->disable_multiscale_loss, n, denom, st)))
    return rc;
  {
    LossBwdArgs a;
    const int S = h->samples(L - 1);
    a.raw = h->raw[L - 1]; a.tdist = h->tdist[L - 1]; a.sdist = h->sdist[L - 1]; a.directions = rays->directions;
    a.rgb_gt = rgb_gt; a.lossmult = rays->lossmult; a.static_mask = rays->static_mask; a.denom = denom;
    a.n_rays = n; a.S = S; a.opaque_background = d.opaque_background; a.bg = d.bg_intensity;
    a.density_bias = d.density_bias; a.rgb_premult = d.rgb_premultiplier; a.rgb_bias = d.rgb_bias;
    a.rgb_padding = d.rgb_padding; a.loss = *loss; a.d_raw = h->d_raw[L - 1]; a.weights = h->weights[L - 1];
    a.ray_stats = h->ray_stats;
    if ((rc = launch_final_loss_bwd(a, st))) return rc;
  }
  for (int l = 0; l < L - 1; ++l) {
    PropLossBwdArgs a;
    a.raw_density = h->raw[l]; a.tdist = h->tdist[l]; a.sdist = h->sdist[l]; a.directions = rays->directions;
    a.sdist_final = h->sdist[L - 1]; a.w_final = h->weights[L - 1]; a.n_rays = n; a.Sp = h->samples(l);
    a.S = h->samples(L - 1); a.opaque_background = d.opaque_background; a.density_bias = d.density_bias;
    a.scale = loss->interlevel_loss_mult / ((float)n * (float)a.S);
    a.d_raw = h->d_raw[l]; a.ray_stats = h->ray_stats + (size_t)n * (4 + l);
    a.rgb_gt = rgb_gt; a.lossmult = rays->lossmult; a.static_mask = rays->static_mask; a.loss = *loss;
    a.bg = d.bg_intensity; a.sq_stats = h->ray_stats + (size_t)n * (8 + l);
    if ((rc = launch_prop_loss_bwd(a, st))) return rc;
  }
  if ((rc = launch_finalize_stats(h, *loss, n, denom, h->ray_stats, stats_out, st))) return rc;
  delete ps_loss; ps_loss = nullptr;

  HUGS_CUDA(cudaMemsetAsync(grad_out, 0, sizeof(float) * h->n_params, st));
  for (int l = L - 1; l >= 0; --l) {
    if ((rc = tc_mlp_backward(h, l, rays, n, grad_out, st))) return rc;
    // NerfMLP_0 and GloEmbed_0 gradients are final here: a caller-provided event lets the all-reduce of that part of the
    // gradient start while the proposal levels are still in their backward pass
    if (l == L - 1 && h->grad_ready_event) HUGS_CUDA(cudaEventRecord(h->grad_ready_event, st));
  }
  return HUGS_OK;
}

HUGS_API int hugs_set_train_rng(hugs_handle* h, uint64_t seed, uint64_t counter) {
  HUGS_REQUIRE(h, "hugs_set_train_rng: null handle");
  // seed == 0 switches the in-kernel draws off again (deterministic sampling when no jitter tensor is passed)
  h->train_rng_key = seed == 0 ? 0 : (seed ^ 0xA0761D6478BD642Full) * 0xE7037ED1A0B428DBull + counter * 0x8EBC6AF09C88C6E3ull + 1;
  if (seed != 0 && h->train_rng_key == 0) h->train_rng_key = 1;
  return HUGS_OK;
}

HUGS_API int hugs_set_grad_ready_event(hugs_handle* h, void* cuda_event) {
  HUGS_REQUIRE(h, "hugs_set_grad_ready_event: null handle");
  h->grad_ready_event = (cudaEvent_t)cuda_event;
  return HUGS_OK;
}

// ------------------------------------------------------------------ torch twins (nerfacto/)

namespace hugs {
namespace {
// Runs `fn` with the last level of the handle re-pointed at caller buffers and sized for n_samples.
template <class F>
int with_field_buffers(hugs_handle* h, int n_samples, const float* tdist, float* raw, const float* d_raw, F fn) {
  const int l = h->d.num_levels - 1;
  float* keep_t = h->tdist[l]; float* keep_r = h->raw[l]; float* keep_d = h->d_raw[l];
  const int keep_s = h->d.num_nerf_samples;
  if (tdist) h->tdist[l] = const_cast<float*>(tdist);
  if (raw) h->raw[l] = raw;
  if (d_raw) h->d_raw[l] = const_cast<float*>(d_raw);
  h->d.num_nerf_samples = n_samples;
  const int rc = fn(l);
  h->tdist[l] = keep_t; h->raw[l] = keep_r; h->d_raw[l] = keep_d; h->d.num_nerf_samples = keep_s;
  return rc;
}
}  // namespace
}  // namespace hugs

HUGS_API int hugs_field_forward(hugs_handle* h, const float* params, const hugs_rays* rays, const float* tdist,
                                int32_t n_rays, int32_t n_samples, int32_t training, int32_t zero_glo, float* raw_out,
                                void* stream) {
  HUGS_REQUIRE(h && params && tdist && raw_out, "hugs_field_forward: null argument");
  int rc = check_rays(h, rays, n_rays, true);
  if (rc) return rc;
  HUGS_REQUIRE(n_samples >= 2 && n_samples <= h->max_nerf_samples, "hugs_field_forward: n_samples %d outside [2, %d]",
               n_samples, h->max_nerf_samples);
  if (n_rays == 0) return HUGS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const hugs_model_desc& d = h->d;
  if (training) {
    if (d.precision == HUGS_PRECISION_FP32) {
      set_error("hugs_field_forward(training) needs a tensor-core precision mode; the fp32 CUDA-core path is render-only");
      return HUGS_ERR_UNSUPPORTED;
    }
    if ((rc = tc_ensure_training(h))) return rc;
  }
  h->cur_params = params;
  h->cur_embed_idx = rays->embed_idx;
  if ((rc = launch_view_inputs(rays->viewdirs, rays->embed_idx, h->glo_off >= 0 ? params + h->glo_off : nullptr, n_rays,
                               d.deg_view, d.num_glo_features, zero_glo, d.num_embeddings, h->view_in, st)))
    return rc;
  return with_field_buffers(h, n_samples, tdist, raw_out, nullptr, [&](int l) {
    if (d.precision == HUGS_PRECISION_FP32) return run_level_mlp_fp32(h, l, params, rays, n_rays, st);
    return tc_mlp_forward(h, l, rays, n_rays, training != 0, st);
  });
}

HUGS_API int hugs_field_backward(hugs_handle* h, const float* params, const hugs_rays* rays, int32_t n_rays,
                                 int32_t n_samples, const float* d_raw, float* grad_out, void* stream) {
  HUGS_REQUIRE(h && params && d_raw && grad_out, "hugs_field_backward: null argument");
  int rc = check_rays(h, rays, n_rays, true);
  if (rc) return rc;
  HUGS_REQUIRE(h->d.precision != HUGS_PRECISION_FP32 && h->tc, "hugs_field_backward needs a tensor-core precision mode");
  HUGS_REQUIRE(n_samples >= 2 && n_samples <= h->max_nerf_samples, "hugs_field_backward: n_samples %d outside [2, %d]",
               n_samples, h->max_nerf_samples);
  HUGS_REQUIRE(n_rays > 0, "hugs_field_backward: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_params = params;
  h->cur_embed_idx = rays->embed_idx;
  HUGS_CUDA(cudaMemsetAsync(grad_out, 0, sizeof(float) * h->n_params, st));
  return with_field_buffers(h, n_samples, nullptr, nullptr, d_raw,
                            [&](int l) { return tc_mlp_backward(h, l, rays, n_rays, grad_out, st); });
}

HUGS_API int hugs_nf_sample_intervals(const float* bins, const float* weights, const float* u_base, const float* jitter,
                                      int32_t jitter_stride, float max_jitter, float anneal, float padding, int32_t n_rays,
                                      int32_t n_bins, int32_t n_samples, float dom_lo, float dom_hi, int32_t spacing_fn,
                                      const float* near, const float* far, float* bins_out, float* t_out, void* stream) {
  HUGS_REQUIRE(n_samples > 1, "num_samples must be > 1, is %d.", n_samples);
  HUGS_REQUIRE(n_bins >= 1 && n_rays >= 0, "hugs_nf_sample_intervals: bad sizes");
  if (n_rays == 0) return HUGS_OK;
  HUGS_REQUIRE(bins && weights && u_base && bins_out, "hugs_nf_sample_intervals: null argument");
  HUGS_REQUIRE(!t_out || (near && far), "hugs_nf_sample_intervals: near / far are required with t_out");
  HUGS_REQUIRE(!jitter || jitter_stride == 1 || jitter_stride == n_samples, "hugs_nf_sample_intervals: jitter_stride must be 1 or n_samples");
  ResampleArgs a;
  a.t_in = bins; a.w_in = weights; a.n_rays = n_rays; a.np = n_bins; a.ns = n_samples; a.dom_lo = dom_lo; a.dom_hi = dom_hi;
  a.anneal = anneal; a.padding = padding; a.u_base = u_base; a.jitter = jitter; a.jitter_stride = jitter ? jitter_stride : 1;
  a.max_jitter = max_jitter; a.torch_twin = 1; a.s_out = bins_out; a.t_out = t_out; a.raydist_fn = spacing_fn;
  a.near = near; a.far = far;
  return launch_resample(a, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_merge_bins(const float* bins_a, int32_t n_a, const float* bins_b, int32_t n_b, int32_t n_rays,
                                float dom_lo, float dom_hi, int32_t spacing_fn, const float* near, const float* far,
                                float* bins_out, float* t_out, void* stream) {
  HUGS_REQUIRE(bins_a && bins_b && bins_out, "hugs_nf_merge_bins: null argument");
  HUGS_REQUIRE(!t_out || (near && far), "hugs_nf_merge_bins: near / far are required with t_out");
  NfMergeArgs a;
  a.bins_a = bins_a; a.na = n_a; a.bins_b = bins_b; a.nb = n_b; a.n_rays = n_rays; a.dom_lo = dom_lo; a.dom_hi = dom_hi;
  a.spacing_fn = spacing_fn; a.near = near; a.far = far; a.bins_out = bins_out; a.t_out = t_out;
  return launch_nf_merge(a, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_composite(const hugs_nf_render_cfg* cfg, const float* raw, int32_t raw_channels, const float* tdist,
                               const float* directions, const float* bg_rgb, int32_t n_rays, int32_t n_samples,
                               float* weights_out, float* rgb_out, float* depth_out, float* acc_out, float* steps_max,
                               void* stream) {
  HUGS_REQUIRE(cfg && raw && tdist && directions, "hugs_nf_composite: null argument");
  HUGS_REQUIRE(raw_channels == 4 || !rgb_out, "hugs_nf_composite: rgb needs raw_channels == 4");
  NfCompositeArgs a;
  a.cfg = *cfg; a.raw = raw; a.C = raw_channels; a.tdist = tdist; a.directions = directions; a.bg_rgb = bg_rgb;
  a.n_rays = n_rays; a.S = n_samples; a.weights = weights_out; a.rgb = rgb_out; a.depth = depth_out; a.acc = acc_out;
  a.steps_max = steps_max;
  return launch_nf_composite(a, false, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_clip_depth(float* depth, const float* steps_max, int32_t n_rays, void* stream) {
  HUGS_REQUIRE(depth && steps_max, "hugs_nf_clip_depth: null argument");
  return launch_nf_clip_depth(depth, steps_max, n_rays, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_composite_bwd(const hugs_nf_render_cfg* cfg, const float* raw, int32_t raw_channels,
                                   const float* tdist, const float* directions, const float* bg_rgb, int32_t n_rays,
                                   int32_t n_samples, const float* d_weights, const float* d_rgb, const float* d_depth,
                                   const float* d_acc, const float* steps_max, float* d_raw, void* stream) {
  HUGS_REQUIRE(cfg && raw && tdist && directions && d_raw, "hugs_nf_composite_bwd: null argument");
  HUGS_REQUIRE(!d_depth || steps_max, "hugs_nf_composite_bwd: steps_max is required with d_depth");
  NfCompositeArgs a;
  a.cfg = *cfg; a.raw = raw; a.C = raw_channels; a.tdist = tdist; a.directions = directions; a.bg_rgb = bg_rgb;
  a.n_rays = n_rays; a.S = n_samples; a.d_weights = d_weights; a.d_rgb = raw_channels == 4 ? d_rgb : nullptr;
  a.d_depth = d_depth; a.d_acc = d_acc; a.steps_max_in = steps_max; a.d_raw = d_raw;
  return launch_nf_composite(a, true, (cudaStream_t)stream);
}

HUGS_API int hugs_params_copy(const hugs_tensor_copy* table, int32_t n, float* flat, int32_t direction, float* tensor_base,
                              void* stream) {
  HUGS_REQUIRE(table && flat && n >= 0, "hugs_params_copy: null argument");
  HUGS_REQUIRE(direction == 0 || direction == 1, "hugs_params_copy: direction must be 0 (tensors -> flat) or 1");
  return launch_params_copy(table, n, flat, direction, tensor_base, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_rgb_loss(const float* pred, const float* gt, const float* static_mask, float transient_weight,
                              int32_t loss_type, float charb_padding, int32_t n_rays, float* sums_out, float* dl_out,
                              void* stream) {
  HUGS_REQUIRE(pred && gt && sums_out && dl_out, "hugs_nf_rgb_loss: null argument");
  HUGS_REQUIRE(loss_type == HUGS_LOSS_MSE || loss_type == HUGS_LOSS_CHARB, "hugs_nf_rgb_loss: unknown loss type %d", loss_type);
  return launch_nf_rgb_loss(pred, gt, static_mask, transient_weight, loss_type, charb_padding, n_rays, sums_out, dl_out,
                            (cudaStream_t)stream);
}

HUGS_API int hugs_nf_rgb_loss_bwd(const float* dl, const float* sums, const float* upstream, float scale, int32_t n_rays,
                                  float* d_pred, void* stream) {
  HUGS_REQUIRE(dl && sums && upstream && d_pred, "hugs_nf_rgb_loss_bwd: null argument");
  return launch_nf_rgb_loss_bwd(dl, sums, upstream, scale, n_rays, d_pred, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_distortion_loss(const float* spacing_bins, const float* weights, int32_t n_rays, int32_t n_samples,
                                     float* sum_out, float* grad_out, void* stream) {
  HUGS_REQUIRE(spacing_bins && weights && sum_out && grad_out, "hugs_nf_distortion_loss: null argument");
  return launch_nf_distortion(spacing_bins, weights, n_rays, n_samples, sum_out, grad_out, (cudaStream_t)stream);
}

HUGS_API int hugs_nf_interlevel_loss(const float* spacing_bins, const float* weights, int32_t n_samples, const float* prop_bins,
                                     const float* prop_weights, int32_t n_prop, int32_t n_rays, float* sum_out,
                                     float* grad_out, void* stream) {
  HUGS_REQUIRE(spacing_bins && weights && prop_bins && prop_weights && sum_out && grad_out, "hugs_nf_interlevel_loss: null argument");
  return launch_nf_interlevel(spacing_bins, weights, n_samples, prop_bins, prop_weights, n_prop, n_rays, sum_out, grad_out,
                              (cudaStream_t)stream);
}

HUGS_API int hugs_nf_scale(const float* src, const float* upstream, float mult, int64_t n, float* dst, void* stream) {
  HUGS_REQUIRE(src && upstream && dst, "hugs_nf_scale: null argument");
  return launch_nf_scale(src, upstream, mult, n, dst, (cudaStream_t)stream);
}
