/* hugs_b200.h — C ABI of the B200-native NeRF-HuGS volume-rendering path.
 *
 * The reference (cnhaox/NeRF-HuGS) has no FFI: its boundary is the Python call surface
 * MipNeRF360/internal/{train_utils,models}.py exposes to MipNeRF360/{train,eval,render}.py.
 * Each entry point below names the reference function(s) it replaces (paths relative to
 * /root/reference).  The Python host (nerf_hugs_b200/) binds these with ctypes and re-creates
 * the reference call surface on top.
 *
 * Conventions
 *  - every function returns 0 on success, a negative hugs_status otherwise; the message is
 *    available (thread-local) from hugs_last_error().  Nothing throws or aborts across the ABI.
 *  - all tensor arguments are caller-owned DEVICE pointers (fp32 unless noted) with explicit
 *    element counts; `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *  - a handle is bound to the device that was current at hugs_create and is not thread-safe.
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with
 *    HUGS_ERR_CUDA.
 */
#ifndef HUGS_B200_H_
#define HUGS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HUGS_ABI_VERSION 2

typedef enum {
  HUGS_OK = 0,
  HUGS_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  HUGS_ERR_CUDA = -2,      /* CUDA runtime / driver error (message has the cudaError string) */
  HUGS_ERR_NOMEM = -3,
  HUGS_ERR_UNSUPPORTED = -4
} hugs_status;

typedef enum { HUGS_RAYDIST_NONE = 0, HUGS_RAYDIST_RECIPROCAL = 1, HUGS_RAYDIST_LOG = 2,
               HUGS_RAYDIST_PIECEWISE = 3 } hugs_raydist_fn;   /* coord.py:63-99 */
typedef enum { HUGS_RAY_CONE = 0, HUGS_RAY_CYLINDER = 1 } hugs_ray_shape; /* render.py:103-127 */
typedef enum { HUGS_PRECISION_FP32 = 0,    /* CUDA-core fp32 MLP (render only): independent cross-check              */
               HUGS_PRECISION_BF16_TC = 1, /* tcgen05 bf16 x bf16 -> fp32: throughput mode                            */
               HUGS_PRECISION_TC_SPLIT = 2 /* the SAME tcgen05 kernels with every operand split into bf16 hi + lo
                                              halves (4 products, fp32 accumulate, fp32 epilogues): forward and
                                              backward at fp32-level accuracy (render 1e-4, gradients 1e-3)          */
} hugs_precision;
typedef enum { HUGS_LOSS_CHARB = 0, HUGS_LOSS_MSE = 1 } hugs_data_loss;  /* train_utils.py:96-103 */
/* Input encoding of the MLPs.  HUGS_ENC_IPE: integrated positional encoding of contracted conical-frustum Gaussians
 * (MipNeRF360/internal/coord.py:102-133).  HUGS_ENC_POINT_PE: the torch twin's pos_enc(x, min_deg, max_deg,
 * append_identity=True) of the interval midpoint o + d * (t0 + t1) / 2, optionally contracted by
 * spatial_distortion_norm2 (nerfacto/models/custom_functions.py:15-21,55-63; nerfacto/models/nerf.py:299-300,788-794). */
typedef enum { HUGS_ENC_IPE = 0, HUGS_ENC_POINT_PE = 1 } hugs_encoding;

/* Model description == the gin-bound fields of models.Model / NerfMLP / PropMLP
 * (MipNeRF360/internal/models.py:47-71, :360-391) the hot path reads. */
typedef struct {
  int32_t num_levels;          /* Model.num_levels            */
  int32_t num_prop_samples;    /* Model.num_prop_samples      */
  int32_t num_nerf_samples;    /* Model.num_nerf_samples      */
  int32_t nerf_depth;          /* NerfMLP.net_depth           */
  int32_t nerf_width;          /* NerfMLP.net_width           */
  int32_t prop_depth;          /* PropMLP.net_depth           */
  int32_t prop_width;          /* PropMLP.net_width           */
  int32_t bottleneck_width;    /* MLP.bottleneck_width        */
  int32_t view_width;          /* MLP.net_width_viewdirs (depth 1) */
  int32_t skip_layer;          /* MLP.skip_layer              */
  int32_t min_deg_point, max_deg_point, deg_view;
  int32_t num_basis;           /* columns of pos_basis_t (21 for icosahedron/2) */
  float   basis[3 * 32];       /* pos_basis_t, row-major [3][num_basis] (geopoly.py:78) */
  int32_t raydist_fn;          /* hugs_raydist_fn: Model.raydist_fn */
  int32_t ray_shape;           /* hugs_ray_shape                    */
  int32_t nerf_contract;       /* NerfMLP.warp_fn == @coord.contract */
  int32_t prop_contract;       /* PropMLP.warp_fn == @coord.contract */
  int32_t opaque_background;
  float   bg_intensity;        /* Model.bg_intensity_range (min==max) */
  float   anneal_slope, dilation_multiplier, dilation_bias, resample_padding;
  float   near_anneal_rate;    /* <= 0: disabled (None) */
  float   near_anneal_init;
  int32_t num_glo_features, num_embeddings;
  float   density_bias, rgb_premultiplier, rgb_bias, rgb_padding;
  int32_t precision;           /* hugs_precision */
  int32_t max_rays;            /* workspace is sized for this many rays per call */
  int32_t encoding;            /* hugs_encoding (0 = IPE: every Mip-NeRF 360 config) */
  int32_t reserved_[3];
} hugs_model_desc;

/* utils.Rays (MipNeRF360/internal/utils.py:44-57) as struct-of-arrays device pointers,
 * each [n_rays, C] contiguous.  pix_coords / cam_idx are not read by this path. */
typedef struct {
  const float* origins;      /* [n,3] */
  const float* directions;   /* [n,3] */
  const float* viewdirs;     /* [n,3] */
  const float* radii;        /* [n,1] */
  const float* near;         /* [n,1] */
  const float* far;          /* [n,1] */
  const float* lossmult;     /* [n,1] (train only; may be NULL => 1) */
  const float* static_mask;  /* [n,1] HuGS mask gathered per pixel (datasets.py:473); NULL => 1 */
  const int32_t* embed_idx;  /* [n,1] (GLO; may be NULL when num_glo_features == 0) */
} hugs_rays;

/* Loss / optimiser fields of configs.Config (configs.py:84-107,131-132). */
typedef struct {
  int32_t data_loss_type;          /* hugs_data_loss */
  float   charb_padding, data_loss_mult, data_coarse_loss_mult;
  float   interlevel_loss_mult, distortion_loss_mult;
  int32_t use_static_mask;         /* transient_type == 'withmask' (train_utils.py:80-82, quirk B1) */
  float   withmask_transient_weight;
  int32_t disable_multiscale_loss;
} hugs_loss_cfg;

typedef struct {
  float lr;                        /* math.learning_rate_decay(step) evaluated by the host */
  float beta1, beta2, eps;         /* optax.adam (train_utils.py:489-510) */
  float grad_max_norm, grad_max_val; /* clip_gradients (train_utils.py:351-369) */
  int32_t step;                    /* optax count before this update (0-based) */
  float grad_scale;                /* gradient pre-multiplier: 1/world_size after an all-reduce(SUM) == pmean
                                      (train_utils.py:457-458); use 1 on a single GPU */
} hugs_adam_cfg;

/* One named view into the flat fp32 parameter buffer (flax names, e.g. "NerfMLP_0/Dense_3/kernel"). */
typedef struct {
  char    name[64];
  int64_t offset;                  /* in floats */
  int32_t rows, cols;              /* kernel: [in,out]; bias: [1,out]; embedding: [num,feat] */
  int32_t module;                  /* 0 NerfMLP_0, 1 PropMLP_0, 2 GloEmbed_0 (clip groups) */
} hugs_tensor_desc;

/* Per-level rendering outputs (render.volumetric_rendering, render.py:185-244). NULL = skip. */
typedef struct {
  float* rgb;            /* [n,3] */
  float* acc;            /* [n]   */
  float* distance_mean, *distance_median, *distance_p5, *distance_p95; /* [n] (compute_extras) */
  float* sdist;          /* [n, S+1] ray_history['sdist']   */
  float* weights;        /* [n, S]   ray_history['weights'] */
  float* density;        /* [n, S]   */
  float* rgbs;           /* [n, S, 3] (NeRF level only) */
} hugs_level_out;

typedef struct hugs_handle hugs_handle;

const char* hugs_last_error(void);
int hugs_abi_version(void);

/* models.construct_model + train_utils.setup_model (models.py:333-357, train_utils.py:579-596) */
int hugs_create(const hugs_model_desc* desc, hugs_handle** out);
int hugs_destroy(hugs_handle* h);
int64_t hugs_param_count(const hugs_handle* h);
int hugs_param_layout(const hugs_handle* h, hugs_tensor_desc* out, int32_t capacity, int32_t* count);
/* Re-derive the packed bf16 operand copies after the caller wrote the fp32 parameters. */
int hugs_params_changed(hugs_handle* h, const float* params, void* stream);

/* ---- operator-level entry points (parity tests bind these one-to-one) ---- */

/* stepfun.sample_intervals (stepfun.py:214-263) incl. softmax/integrate_weights/sorted_interp
 * (stepfun.py:131-161, math.py:108-127).  u = u_base[j] + jitter[ray]*max_jitter (jitter may be NULL).
 * idx_out (optional) receives the selected CDF interval of every sample centre. */
int hugs_sample_intervals(const float* t, const float* w_logits, const float* u_base,
                          const float* jitter, float max_jitter, int32_t n_rays, int32_t n_bins,
                          int32_t n_samples, float dom_lo, float dom_hi,
                          float* t_out, int32_t* idx_out, void* stream);
/* math.sorted_interp on a caller-provided CDF: bit-exact index + value contract. */
int hugs_invert_cdf(const float* t, const float* cw, const float* u, int32_t n_rays, int32_t n_bins,
                    int32_t n_samples, float* centers_out, int32_t* idx_out, void* stream);
/* stepfun.max_dilate_weights(renormalize=True) + the [1:-1] trim of models.py:171-179.
 * t_out [n, 3*n_bins-1], w_out [n, 3*n_bins-2]. */
int hugs_max_dilate_weights(const float* t, const float* w, int32_t n_rays, int32_t n_bins,
                            float dilation, float dom_lo, float dom_hi,
                            float* t_out, float* w_out, void* stream);
/* render.compute_alpha_weights + render.volumetric_rendering (render.py:130-151,185-244).
 * raw_density [n,S] is pre-activation (softplus(raw + density_bias) applied inside);
 * raw_rgb [n,S,3] pre-sigmoid or NULL (proposal levels: rgb = 0). */
int hugs_alpha_composite(const hugs_handle* h, const float* raw_density, const float* raw_rgb,
                         const float* tdist, const float* sdist, const float* directions,
                         const float* far, int32_t n_rays, int32_t n_samples, int32_t compute_extras,
                         const hugs_level_out* out, void* stream);
/* coord.track_linearize(contract) + lift_and_diagonalize + integrated_pos_enc on cast_rays
 * Gaussians (render.py:103-127, coord.py:39-60,102-133): features [n*S, 2*num_basis*degs],
 * reference column order.  exact != 0 uses the reference's safe_sin arithmetic. */
int hugs_ipe_features(const hugs_handle* h, const hugs_rays* rays, const float* tdist,
                      int32_t n_rays, int32_t n_samples, int32_t contract, float* features,
                      void* stream);

/* Test hook: the throughput-mode bf16 IPE encoder on its own ([n*S, 512] bf16, engine column order
 * f' = (b * degs + k) * 2 + {sin, shifted sin}; columns >= 2*num_basis*degs are zero).  Needs a tensor-core handle. */
int hugs_debug_encode_bf16(hugs_handle* h, const hugs_rays* rays, const float* tdist, int32_t n_rays,
                           int32_t n_samples, int32_t contract, void* features_bf16, void* stream);

/* ---- model-level entry points ---- */

/* Model.__call__ with rng=None or caller-provided jitter (models.py:74-330);
 * render_eval_pfn's payload (train_utils.py:558-575).  jitter: [num_levels, n_rays] or NULL.
 * out: array of num_levels hugs_level_out. */
int hugs_forward(hugs_handle* h, const float* params, const hugs_rays* rays, int32_t n_rays,
                 float train_frac, const float* jitter, int32_t compute_extras, int32_t zero_glo,
                 const hugs_level_out* out, void* stream);

/* loss_fn + value_and_grad of train_utils.train_step (train_utils.py:407-455) on this rank's
 * rays.  grad_out: flat fp32 [param_count] (overwritten).  stats_out: fp32[16]:
 * [0] loss, [1] data, [2] interlevel, [3] distortion, [4..4+L) mse per level (proposal levels: of their
 * background-only rendering, as the reference reports them). */
int hugs_loss_and_grad(hugs_handle* h, const float* params, const hugs_rays* rays,
                       const float* rgb_gt, int32_t n_rays, float train_frac, const float* jitter,
                       const hugs_loss_cfg* loss, float* grad_out, float* stats_out, void* stream);

/* Train-mode sampling randomness without a jitter tensor: when `jitter` is NULL in hugs_loss_and_grad and a non-zero
 * seed was set, the sampling kernel draws one uniform per (level, ray) from a counter-based hash of (seed, counter)
 * (the role of `rng, key = random.split(rng)` in train_step, train_utils.py:408; the stream is this library's own —
 * threefry cannot be reproduced without JAX).  seed == 0 switches it off. */
int hugs_set_train_rng(hugs_handle* h, uint64_t seed, uint64_t counter);
/* Multi-GPU overlap: `cuda_event` (a cudaEvent_t, or NULL to clear) is recorded on the call's stream inside
 * hugs_loss_and_grad as soon as the NerfMLP_0 / GloEmbed_0 part of grad_out is final, i.e. before the proposal levels'
 * backward pass, so that the all-reduce of that part (jax.lax.pmean, train_utils.py:457-458) can start early. */
int hugs_set_grad_ready_event(hugs_handle* h, void* cuda_event);

/* clip_gradients + nan_to_num + optax.adam apply (train_utils.py:351-369,464-468) on the
 * (already all-reduced) flat gradient.  norms_out (optional) fp32[9]: {grad norm, abs-max, clip multiplier} per module. */
int hugs_adam_step(hugs_handle* h, float* params, const float* grad, float* mu, float* nu,
                   const hugs_adam_cfg* cfg, float* norms_out, void* stream);

/* hugs_adam_step + the per-tensor statistics trees of train_step (train_utils.py:442,461-462,470-473) from the same pass:
 * tensor_stats_out fp32[5 * n_tensors] (tensor order of hugs_param_layout), per tensor
 * {sum w^2 before the update, sum g^2, max |g| (averaged, unclipped gradient), sum delta^2, max |delta| (applied update)}. */
int hugs_adam_step_stats(hugs_handle* h, float* params, const float* grad, float* mu, float* nu,
                         const hugs_adam_cfg* cfg, float* norms_out, float* tensor_stats_out, void* stream);

/* ---- the torch twins: nerfacto/{train,eval}.py call surface (SURVEY.md §8(a) last row, §8(b) last row) ----
 *
 * nerfacto/models/nerf.py::Model.forward_rays is, per field ('coarse', 'fine'): sample_intervals under no_grad ->
 * s_to_t -> field MLP on point encodings -> density_to_weight -> render_features / render_depth.  The host mirror
 * (nerf_hugs_b200/nerfacto/models/nerf.py) strings the operators below together and exposes them to torch autograd;
 * one hugs_handle (num_levels = 1, encoding = HUGS_ENC_POINT_PE) per field carries the MLP. */

/* MLP.forward of one field on caller-provided interval fenceposts (nerf.py:299-318,788-860 == the NerfMLP of
 * MipNeRF360/internal/models.py:405-550 with a point encoding): raw_out [n, S, 4] = pre-activation density and rgb.
 * Uses the handle's last (NeRF) level; n_samples <= num_nerf_samples.  rays: origins, directions, viewdirs
 * (+ embed_idx with appearance embeddings; radii with HUGS_ENC_IPE).  training != 0 saves what hugs_field_backward needs. */
int hugs_field_forward(hugs_handle* h, const float* params, const hugs_rays* rays, const float* tdist,
                       int32_t n_rays, int32_t n_samples, int32_t training, int32_t zero_glo, float* raw_out,
                       void* stream);
/* loss.backward() through the field of the preceding hugs_field_forward(training = 1): d_raw [n, S, 4] ->
 * grad_out (flat fp32 [param_count], flax layout, overwritten). */
int hugs_field_backward(hugs_handle* h, const float* params, const hugs_rays* rays, int32_t n_rays,
                        int32_t n_samples, const float* d_raw, float* grad_out, void* stream);

/* utils/ray_utils.py:113-223 sample + sample_intervals (torch.searchsorted(side='right') + gather; quirk B5: a ray
 * whose bins all have zero width gets uniform logits).  u = u_base[j] + jitter[ray * jitter_stride + j * (stride > 1)]
 * * max_jitter; jitter NULL = deterministic.  spacing_fn (hugs_raydist_fn: NONE = 'uniform', RECIPROCAL, PIECEWISE) with
 * near / far [n] turns the new fenceposts into euclidean ones (nerf.py:221-226) when t_out != NULL. */
int hugs_nf_sample_intervals(const float* bins, const float* weights, const float* u_base, const float* jitter,
                             int32_t jitter_stride, float max_jitter, float anneal, float padding, int32_t n_rays,
                             int32_t n_bins, int32_t n_samples, float dom_lo, float dom_hi, int32_t spacing_fn,
                             const float* near, const float* far, float* bins_out, float* t_out, void* stream);
/* nerf.py:287-295: fenceposts around the sorted union of the centres of two fencepost sets. */
int hugs_nf_merge_bins(const float* bins_a, int32_t n_a, const float* bins_b, int32_t n_b, int32_t n_rays,
                       float dom_lo, float dom_hi, int32_t spacing_fn, const float* near, const float* far,
                       float* bins_out, float* t_out, void* stream);

typedef enum { HUGS_DENSITY_SOFTPLUS = 0, HUGS_DENSITY_TRUNC_EXP = 1, HUGS_DENSITY_RELU = 2 } hugs_density_act;
typedef struct {
  int32_t opaque_background;
  int32_t density_activation;    /* hugs_density_act (nerf.py:682-691, nerfacto.py:702-710) */
  float   density_bias;
  float   rgb_premultiplier, rgb_bias, rgb_padding;   /* nerf.py:696-698,832 */
  int32_t reserved_[2];
} hugs_nf_render_cfg;

/* utils/ray_utils.py:226-249 density_to_weight (quirk B2: deltas from the FIRST fencepost; B6: nan_to_num),
 * :295-312 render_features (per-ray background colour), :336-346 render_depth before its clip.
 * raw [n, S, C]: C = 4 (density, r, g, b pre-activation) or C = 1 (density only: rgb_out must be NULL).
 * steps_max: device scalar, atomically raised to max over rays of the last interval midpoint (caller zeroes it);
 * hugs_nf_clip_depth applies the batch-wide clip of quirk B7. */
int hugs_nf_composite(const hugs_nf_render_cfg* cfg, const float* raw, int32_t raw_channels, const float* tdist,
                      const float* directions, const float* bg_rgb, int32_t n_rays, int32_t n_samples,
                      float* weights_out, float* rgb_out, float* depth_out, float* acc_out, float* steps_max,
                      void* stream);
int hugs_nf_clip_depth(float* depth, const float* steps_max, int32_t n_rays, void* stream);
/* Backward of hugs_nf_composite (+ the clip): upstream gradients d_weights [n,S], d_rgb [n,3], d_depth [n], d_acc [n]
 * (each may be NULL) -> d_raw [n, S, C]. */
int hugs_nf_composite_bwd(const hugs_nf_render_cfg* cfg, const float* raw, int32_t raw_channels, const float* tdist,
                          const float* directions, const float* bg_rgb, int32_t n_rays, int32_t n_samples,
                          const float* d_weights, const float* d_rgb, const float* d_depth, const float* d_acc,
                          const float* steps_max, float* d_raw, void* stream);

/* nerf.py:404-461 / nerfacto.py:428-490 compute_data_loss / compute_withmask_loss of one rendering:
 * e = (pred - gt)^2, l = e (HUGS_LOSS_MSE) or sqrt(e + padding^2) (HUGS_LOSS_CHARB), lossmult m per ray (static_mask >= 0.5
 * ? 1 : transient_weight; NULL mask: 1) broadcast over the channels BEFORE the sums (no quirk B1 on the torch side).
 * sums_out fp32[3] (zeroed here) = {sum m*l, sum m*e, sum m (x channels)}; dl_out [n,3] = m * dl/dpred for the backward. */
int hugs_nf_rgb_loss(const float* pred, const float* gt, const float* static_mask, float transient_weight,
                     int32_t loss_type, float charb_padding, int32_t n_rays, float* sums_out, float* dl_out, void* stream);
/* d_pred = upstream[0] * scale / max(sums[2], eps) * dl   (upstream: device scalar dL/dloss) */
int hugs_nf_rgb_loss_bwd(const float* dl, const float* sums, const float* upstream, float scale, int32_t n_rays,
                         float* d_pred, void* stream);

/* utils/loss_utils.py:66-84 distortion_loss on the final level: sum_out[0] = sum over rays of lossfun_distortion(c, w)
 * (the caller divides by n_rays: torch.mean), grad_out [n, S] = d lossfun_distortion / d w per ray. */
int hugs_nf_distortion_loss(const float* spacing_bins, const float* weights, int32_t n_rays, int32_t n_samples,
                            float* sum_out, float* grad_out, void* stream);
/* utils/loss_utils.py:7-63 interlevel_loss term of ONE proposal level: sum_out[0] = sum of lossfun_outer(c, w, cp, wp) over
 * rays and final-level samples (the caller divides by n_rays * n_samples), grad_out [n, n_prop] = d sum / d wp. */
int hugs_nf_interlevel_loss(const float* spacing_bins, const float* weights, int32_t n_samples, const float* prop_bins,
                            const float* prop_weights, int32_t n_prop, int32_t n_rays, float* sum_out, float* grad_out,
                            void* stream);
/* dst[i] = src[i] * upstream[0] * mult (upstream: device scalar): the chain rule of the two losses above. */
int hugs_nf_scale(const float* src, const float* upstream, float mult, int64_t n, float* dst, void* stream);

/* torch parameters <-> the flat flax-layout buffer of a handle: one table-driven copy instead of one per tensor.
 * table (DEVICE pointer) of n entries; direction 0: flat[flat_off + i*cols + j] = transpose ? ptr[j*ld + i]
 * : ptr[i*ld + j] (ld defaults to rows resp. cols); direction 1: the reverse (flat -> tensors).  nn.Linear.weight is [out, in] (transpose = 1 against
 * the [in, out] kernel of hugs_param_layout). */
typedef struct {
  float*  ptr;
  int64_t flat_off;
  int32_t rows, cols;            /* shape of the FLAT (flax) view */
  int32_t transpose;
  int32_t ld;                    /* elements between consecutive rows of the tensor behind `ptr` (0: dense) */
} hugs_tensor_copy;
int hugs_params_copy(const hugs_tensor_copy* table, int32_t n, float* flat, int32_t direction, float* tensor_base,
                     void* stream);   /* tensor_base != NULL: every `ptr` of the table is a BYTE OFFSET from it (a constant table
                                         for buffers that move, e.g. a fresh gradient buffer per backward pass) */

/* ---- hash-grid fields of the nerfacto twin (nerfacto/models/nerfacto.py:643-1008) ----
 *
 * tcnn.Encoding('HashGrid') / ('SphericalHarmonics', degree 4) are a third-party dependency that is not under the reference
 * tree (tiny-cuda-nn, unpinned git HEAD): nerf_hugs_b200/csrc/hashfield.cu restates its published algorithm; `grid` below has
 * tcnn's parameter layout (levels back to back, features innermost), so a reference state_dict's `params` tensor is used
 * in place.  geo_feat_dim == 0: HashMLPDensityField (grid -> 64 -> raw density, fused CUDA-core kernel, raw_out [n, S]);
 * geo_feat_dim == 64: NerfactoField (grid -> 256 -> [density | 64]; [SH4 | geometry | appearance] -> 256 -> 256 -> rgb, tcgen05
 * GEMMs, raw_out [n, S, 4]).  Samples whose normalised position leaves [0, 1]^3 are evaluated at 0 and get raw density -inf
 * (density * selector, nerfacto.py:822-836). */
typedef struct {
  int32_t n_levels, features_per_level, log2_hashmap_size, base_res;
  float   per_level_scale;
  int32_t hidden_dim, geo_feat_dim, hidden_dim_color;
  int32_t appearance_dim, num_embeddings;
  float   bound;               /* positions are mapped by (x + bound) / (2 bound) ... */
  int32_t contract;            /* ... or, with scene contraction, (spatial_distortion_norm2(x) + 2) / 4 */
  int32_t max_samples;         /* n_rays * n_samples the workspace is sized for */
  int32_t max_rays;
  int32_t precision;           /* NerfactoField only: 0 or HUGS_PRECISION_BF16_TC = throughput, HUGS_PRECISION_TC_SPLIT = the same
                                  tcgen05 GEMMs with bf16 hi + lo operands (fp32-level parity, forward and backward) */
  int32_t reserved_;
} hugs_hashfield_desc;
typedef struct hugs_hashfield hugs_hashfield;

int hugs_hashfield_create(const hugs_hashfield_desc* desc, hugs_hashfield** out);
int hugs_hashfield_destroy(hugs_hashfield* h);
int64_t hugs_hashfield_grid_floats(const hugs_hashfield* h);   /* == tcnn Encoding.params.numel() */
int64_t hugs_hashfield_mlp_floats(const hugs_hashfield* h);    /* flat fp32 MLP buffer ([in, out] kernels, hugs_hashfield_layout) */
int hugs_hashfield_layout(const hugs_hashfield* h, hugs_tensor_desc* out, int32_t capacity, int32_t* count);
int hugs_hashfield_level_info(const hugs_hashfield* h, int32_t level, float* scale, uint32_t* resolution, uint32_t* offset,
                              uint32_t* entries);
/* Re-derive the packed bf16 operands after the caller wrote the flat MLP buffer (NerfactoField only). */
int hugs_hashfield_params_changed(hugs_hashfield* h, const float* mlp, void* stream);
/* Operator-level hook: the hash encoding alone, fp32 features [n * S, n_levels * 2] of the interval midpoints. */
int hugs_hashfield_encode(hugs_hashfield* h, const float* grid, const hugs_rays* rays, const float* tdist, int32_t n_rays,
                          int32_t n_samples, float* features, void* stream);
/* field(positions, viewdirs, embedded_appearance, None) of nerfacto.py:838-875 / :991-1008 on the interval midpoints. */
int hugs_hashfield_forward(hugs_hashfield* h, const float* grid, const float* mlp, const hugs_rays* rays, const float* tdist,
                           int32_t n_rays, int32_t n_samples, int32_t training, int32_t zero_app, float* raw_out, void* stream);
/* Its backward: d_raw -> grid_grad (ACCUMULATED: the caller zeroes it) and mlp_grad (overwritten). */
int hugs_hashfield_backward(hugs_hashfield* h, const float* grid, const float* mlp, const hugs_rays* rays, const float* tdist,
                            int32_t n_rays, int32_t n_samples, const float* d_raw, float* grid_grad, float* mlp_grad,
                            void* stream);

/* ---- batch assembly on the device (the caller of the path: SURVEY.md §8f item 2) ---- */

/* Device-resident dataset: Dataset.{pixtocams, camtoworlds, heights, widths, images, static_masks, nears, fars,
 * embed_idxs} (MipNeRF360/internal/datasets.py:310-383,446-482).  Images of different sizes are packed back to back;
 * camera c's pixel (y, x) is element pixel_offset[c] + y * widths[c] + x of every per-pixel store. */
typedef struct {
  const float* pixtocams;      /* [n_cams, 3, 3] inverse intrinsics (camera_utils.get_pixtocam) */
  const float* camtoworlds;    /* [n_cams, 3, 4] */
  const int32_t* heights;      /* [n_cams] */
  const int32_t* widths;       /* [n_cams] */
  const int64_t* pixel_offset; /* [n_cams] */
  const float* images;         /* packed [pixels, 3] fp32 in [0,1], or NULL */
  const uint8_t* images_u8;    /* packed [pixels, 3] uint8 (rgb = u8 / 255), or NULL; takes precedence */
  const float* static_masks;   /* packed [pixels] HuGS static masks (datasets.py:473), NULL => 1 */
  const float* nears;          /* packed [pixels] per-pixel near (datasets.py:474), NULL => `near` */
  const float* fars;           /* packed [pixels], NULL => `far` */
  const int32_t* embed_idxs;   /* [n_cams], NULL => camera index */
  float near, far;
  const float* distortion;     /* [6] k1 k2 k3 k4 p1 p2 (Dataset.distortion_params, camera_utils.py:460-494), NULL => none */
  int32_t camtype;             /* 0 perspective, 1 fisheye (camera_utils.ProjectionType) */
  int32_t reserved_;
} hugs_camera_set;

/* utils.Rays + Batch.rgb as writable struct-of-arrays device pointers, each [n_rays, C] contiguous. */
typedef struct {
  float* origins; float* directions; float* viewdirs;   /* [n,3] */
  float* radii; float* near; float* far; float* lossmult; float* static_mask;   /* [n,1] */
  int32_t* embed_idx;          /* [n,1] */
  int32_t* cam_idx;            /* [n,1] or NULL */
  float* pix_coords;           /* [n,2] or NULL */
  float* rgb;                  /* [n,3] or NULL (render paths have no images) */
} hugs_ray_batch;

/* Dataset._make_ray_batch (datasets.py:446-482) -> camera_utils.cast_ray_batch -> pixels_to_rays
 * (camera_utils.py:503-607,610-669) for perspective and fisheye cameras with optional lens distortion (no NDC):
 * rays, HuGS static mask, near/far and ground-truth colours of pixels (cam_idx[i], pix_y[i], pix_x[i]). */
int hugs_make_ray_batch(const hugs_camera_set* cams, const int32_t* cam_idx, const int32_t* pix_x,
                        const int32_t* pix_y, int32_t n_rays, const hugs_ray_batch* out, void* stream);

/* ---- full-frame render pipeline ---- */

/* Frame-sized outputs of one pixel stripe, row-major [rows, width, C]; NULL = skip. */
typedef struct {
  float*   rgb;               /* [rows, W, 3] rendering['rgb'] of the final level */
  float*   acc;               /* [rows, W] */
  float*   distance_mean;     /* [rows, W] (needs compute_extras, like render_image) */
  float*   distance_median;   /* [rows, W] */
  uint8_t* rgb_u8;            /* [rows, W, 3] utils.save_img_u8's quantisation: (clip(nan_to_num(rgb), 0, 1) * 255) truncated */
  double*  sse;               /* [2], ACCUMULATED (the caller zeroes it): sum over the stripe's pixels and channels of
                                 (rgb - gt)^2 and of (round(rgb * 255) / 255 - gt)^2 (eval.py:139-143 eval_quantize_metrics)
                                 against the camera set's image: psnr = -10 log10(sse / (3 H W)) (image.mse_to_psnr) */
} hugs_frame_out;

/* models.render_image (models.py:568-649) driven by eval.py:104-160 / render.py:164-187 for camera `cam` of a device-resident
 * dataset, pixel rows [row0, row1) of a width x height image: rays are generated on the device (hugs_make_ray_batch's
 * arithmetic) and rendered chunk by chunk (max_rays of the handle per chunk) with the deterministic path (rng = None); every
 * chunk writes straight into the frame-sized outputs.  One call per frame and rank: no host loop, no per-chunk gather.
 * `width` / `height` must be the camera's own (cams->widths[cam], cams->heights[cam]: device arrays the host cannot check). */
int hugs_render_frame(hugs_handle* h, const float* params, const hugs_camera_set* cams, int32_t cam, int32_t width,
                      int32_t height, int32_t row0, int32_t row1, float train_frac, int32_t zero_glo,
                      const hugs_frame_out* out, void* stream);

/* ---- measurement hooks (bench.py): no reference counterpart beyond train.py:162-168 wall-clock ---- */

/* Number of kernels this library has launched in this process (every launch site counts itself). */
int64_t hugs_launch_count(void);

typedef enum {
  HUGS_K_SAMPLE = 0, HUGS_K_ENCODE = 1, HUGS_K_CHAIN_FWD_PROP = 2, HUGS_K_CHAIN_FWD_NERF = 3,
  HUGS_K_COMPOSITE_LOSS = 4, HUGS_K_CHAIN_BWD_NERF = 5, HUGS_K_CHAIN_BWD_PROP = 6, HUGS_K_WGRAD_NERF = 7,
  HUGS_K_WGRAD_PROP = 8, HUGS_K_REDUCTIONS = 9, HUGS_K_ADAM_PACK = 10, HUGS_K_MLP_FP32 = 11, HUGS_K_COUNT = 12
} hugs_kernel_class;
/* Enable/disable CUDA-event timing of kernel classes on the launching stream. */
int hugs_profile_enable(hugs_handle* h, int32_t enable);
/* Synchronises, then returns accumulated milliseconds and launch-group counts per class (arrays of
 * HUGS_K_COUNT) since the last read, and resets them. */
int hugs_profile_read(hugs_handle* h, float* ms_out, int32_t* count_out);

#ifdef __cplusplus
}
#endif
#endif  /* HUGS_B200_H_ */
