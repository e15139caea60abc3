// Internal launch interfaces between the translation units of libhugs_b200.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hugs_b200.h"

namespace hugs {

// ---------------------------------------------------------------- sampling.cu
struct ResampleArgs {
  const float* t_in = nullptr;    // [n, np+1]; nullptr => single interval [dom_lo, dom_hi], weight 1
  const float* w_in = nullptr;    // [n, np] weights (or logits when w_is_logits)
  const float* cw_in = nullptr;   // optional caller-provided CDF [n, nb+1] (skips softmax/cumsum)
  const float* u_in = nullptr;    // optional per-ray u [n, ns] (overrides u_base/jitter)
  int n_rays = 0, np = 1, ns = 0; // ns == 0: dilation only
  int dilate = 0;
  float dilation = 0.f, dom_lo = 0.f, dom_hi = 1.f;
  int w_is_logits = 0;
  float anneal = 1.f, padding = 0.f;
  const float* u_base = nullptr;  // [ns]
  const float* jitter = nullptr;  // [n] uniform draws in [0,1) or nullptr
  float max_jitter = 0.f;
  uint64_t jitter_key = 0;        // != 0 (and jitter == nullptr): one draw per ray from hash(jitter_key, ray) in the kernel
  int jitter_stride = 1;          // floats per ray in `jitter`: 1 = one draw per ray, ns = one per sample (torch twin)
  int torch_twin = 0;             // utils/ray_utils.py:143-144 (quirk B5): a ray whose logits are all -inf gets uniform logits
  float* s_out = nullptr;         // [n, ns+1]
  float* t_out = nullptr;         // [n, ns+1] metric distances (optional)
  float* centers_out = nullptr;   // [n, ns] (optional)
  int32_t* idx_out = nullptr;     // [n, ns] (optional)
  int raydist_fn = 0;
  const float* near = nullptr;
  const float* far = nullptr;
  float* td_out = nullptr;        // dilated + trimmed t [n, 3np-1] (optional)
  float* wd_out = nullptr;        // dilated + trimmed w [n, 3np-2] (optional)
};
int launch_resample(const ResampleArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- composite.cu
struct CompositeArgs {
  const float* raw_density = nullptr;  // [n, S]
  const float* raw_rgb = nullptr;      // [n, S, 3] or nullptr
  int raw_stride = 1;                  // floats between consecutive samples of raw_density
  int rgb_stride = 3;                  // floats between consecutive samples of raw_rgb
  const float* tdist = nullptr;        // [n, S+1]
  const float* directions = nullptr;   // [n, 3]
  const float* far = nullptr;          // [n]
  int n_rays = 0, S = 0;
  int opaque_background = 0, compute_extras = 0;
  float bg = 1.f, density_bias = -1.f, rgb_premult = 1.f, rgb_bias = 0.f, rgb_padding = 0.001f;
  hugs_level_out out{};
};
int launch_composite(const CompositeArgs& a, cudaStream_t stream);

// raygen.cu: rays of pixels [pix0, pix0 + n) (row-major) of one camera; uint8 quantisation + squared error of a rendered stripe
int launch_frame_rays(const hugs_camera_set& cams, int cam, int width, long long pix0, int n, const hugs_ray_batch& out,
                      cudaStream_t st);
int launch_frame_finish(const float* rgb, long long n_values, const hugs_camera_set& cams, int cam, long long value0,
                        uint8_t* rgb_u8, double* sse, cudaStream_t st);

struct LossBwdArgs {
  // final (NeRF) level
  const float* raw = nullptr;          // [n, S, 4] (raw_density, raw_r, raw_g, raw_b)
  const float* tdist = nullptr;        // [n, S+1]
  const float* sdist = nullptr;        // [n, S+1]
  const float* directions = nullptr;   // [n, 3]
  const float* rgb_gt = nullptr;       // [n, 3]
  const float* lossmult = nullptr;     // [n] or nullptr
  const float* static_mask = nullptr;  // [n] or nullptr
  const float* denom = nullptr;        // device scalar: sum of loss multipliers (pre-clamp)
  int n_rays = 0, S = 0;
  int opaque_background = 0;
  float bg = 1.f, density_bias = -1.f, rgb_premult = 1.f, rgb_bias = 0.f, rgb_padding = 0.001f;
  hugs_loss_cfg loss{};
  float* d_raw = nullptr;              // [n, S, 4] out
  float* weights = nullptr;            // [n, S] out (final-level weights, reused by interlevel)
  float* ray_stats = nullptr;          // [n, 4] out: data-loss numerator, sq-err numerator, distortion, unused
};
int launch_final_loss_bwd(const LossBwdArgs& a, cudaStream_t stream);

struct PropLossBwdArgs {
  const float* raw_density = nullptr;  // [n, Sp]
  const float* tdist = nullptr;        // [n, Sp+1]
  const float* sdist = nullptr;        // [n, Sp+1]
  const float* directions = nullptr;
  const float* sdist_final = nullptr;  // [n, S+1]
  const float* w_final = nullptr;      // [n, S]
  int n_rays = 0, Sp = 0, S = 0;
  int opaque_background = 0;
  float density_bias = -1.f;
  float scale = 0.f;                   // interlevel_loss_mult / (n_rays * S)
  float* d_raw = nullptr;              // [n, Sp] out
  float* ray_stats = nullptr;          // [n] out: per-ray sum of lossfun_outer
  // stats['mses'] of a proposal level (train_utils.py:94): its rendering is max(0, 1 - acc) * bg
  const float* rgb_gt = nullptr;       // [n, 3]
  const float* lossmult = nullptr;     // [n] or nullptr
  const float* static_mask = nullptr;  // [n] or nullptr
  hugs_loss_cfg loss{};
  float bg = 1.f;
  float* sq_stats = nullptr;           // [n] out: lossmult-weighted squared error of the level's rendering (optional)
};
int launch_prop_loss_bwd(const PropLossBwdArgs& a, cudaStream_t stream);

// ---------------------------------------------------------------- nerfacto_ops.cu (torch twins)
struct NfMergeArgs {
  const float* bins_a = nullptr; int na = 0;   // [n, na+1]
  const float* bins_b = nullptr; int nb = 0;   // [n, nb+1]
  int n_rays = 0;
  float dom_lo = 0.f, dom_hi = 1.f;
  int spacing_fn = 0;
  const float* near = nullptr; const float* far = nullptr;
  float* bins_out = nullptr;                   // [n, na+nb+1]
  float* t_out = nullptr;                      // optional
};
int launch_nf_merge(const NfMergeArgs& a, cudaStream_t stream);

struct NfCompositeArgs {
  hugs_nf_render_cfg cfg{};
  const float* raw = nullptr; int C = 4;       // [n, S, C]
  const float* tdist = nullptr;                // [n, S+1] euclidean fenceposts
  const float* directions = nullptr;           // [n, 3]
  const float* bg_rgb = nullptr;               // [n, 3] or nullptr
  int n_rays = 0, S = 0;
  // forward outputs (each optional)
  float* weights = nullptr; float* rgb = nullptr; float* depth = nullptr; float* acc = nullptr; float* steps_max = nullptr;
  // backward inputs (each optional) / output
  const float* d_weights = nullptr; const float* d_rgb = nullptr; const float* d_depth = nullptr; const float* d_acc = nullptr;
  const float* steps_max_in = nullptr;
  float* d_raw = nullptr;
};
int launch_nf_composite(const NfCompositeArgs& a, bool backward, cudaStream_t stream);
int launch_nf_clip_depth(float* depth, const float* steps_max, int n, cudaStream_t stream);
int launch_nf_rgb_loss(const float* pred, const float* gt, const float* mask, float transient_w, int loss_type,
                       float padding, int n, float* sums, float* dl, cudaStream_t stream);
int launch_nf_rgb_loss_bwd(const float* dl, const float* sums, const float* upstream, float scale, int n, float* d_pred,
                           cudaStream_t stream);
int launch_nf_distortion(const float* c, const float* w, int n_rays, int S, float* out, float* grad, cudaStream_t stream);
int launch_nf_interlevel(const float* c, const float* w, int S, const float* cp, const float* wp, int Sp, int n_rays,
                         float* out, float* grad, cudaStream_t stream);
int launch_nf_scale(const float* src, const float* upstream, float mult, long long n, float* dst, cudaStream_t stream);
int launch_params_copy(const hugs_tensor_copy* table, int n, float* flat, int direction, float* base, cudaStream_t stream);

// pos_enc of interval midpoints (custom_functions.py:55-63): fp32 features [n*S, 3 + 6*ndeg] in the reference's column order,
// and / or bf16 rows of `ld` columns (hi, optional residual lo; columns beyond the features zero up to `zero_cols`)
struct PointPeArgs {
  const float* origins; const float* directions; const float* tdist;
  int n_rays, S, min_deg, ndeg, contract;
  float* features;                             // fp32 [n*S, feat_dim] or nullptr
  __nv_bfloat16* hi; __nv_bfloat16* lo;        // bf16 [rows_pad, ld] or nullptr
  int ld, zero_cols, rows_pad;
};
int launch_point_pe(const PointPeArgs& a, cudaStream_t stream);

// sums `n` floats (optionally thresholded at 0.5 like the static mask) into out[0]
int launch_lossmult_sum(const float* lossmult, const float* static_mask, int use_mask, float transient_w,
                        int disable_multiscale, int n, float* out, cudaStream_t stream);
// out[k] = scale_k * sum_r in[r*stride + k]  (deterministic single-block reduction)
int launch_column_sums(const float* in, int n_rows, int stride, int n_cols, float* out, cudaStream_t stream);

}  // namespace hugs
