// The opaque handle behind the C ABI: model description, parameter layout, device workspace.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "mlp.h"

namespace hugs {

struct DenseView {
  int64_t kernel_off, bias_off;   // float offsets into the flat parameter buffer
  int in, out;
};

struct MlpViews {
  std::vector<DenseView> dense;   // flax creation order: trunk..., density, [bottleneck, view, rgb]
  int depth = 0, width = 0;
  bool has_rgb = false;
  int module = 0;
};

struct TcState;                   // tensor-core path state (mlp_tc.cu)

}  // namespace hugs

struct hugs_handle {
  hugs_model_desc d{};
  int device = 0;
  int feat_dim = 0;               // IPE: 2 * num_basis * (max_deg - min_deg); point PE: 3 + 6 * (max_deg - min_deg)
  int feat_panels = 0;            // 64-column K panels that carry features (tensor-core path)
  int perm_nb = 0;                // basis count of the engine's feature-column permutation (0: reference order, point PE)
  int max_nerf_samples = 0;       // d.num_nerf_samples at creation (hugs_field_forward may run fewer)
  int view_in_dim = 0;            // 3 + 6*deg_view + glo
  int64_t n_params = 0;
  int64_t glo_off = -1;
  std::vector<hugs_tensor_desc> tensors;
  hugs::MlpViews nerf, prop;
  int64_t module_begin[3] = {0, 0, 0}, module_end[3] = {0, 0, 0};

  // ---- device workspace (sized for d.max_rays) ----
  float* basis = nullptr;                     // [3][num_basis]
  std::vector<float*> u_det, u_train;         // per level [S]
  std::vector<float> max_jitter;              // per level
  std::vector<float*> sdist, tdist, weights;  // per level [n,S+1], [n,S+1], [n,S]
  std::vector<float*> raw;                    // per level: prop [n,S]; nerf [n,S,4]
  std::vector<float*> d_raw;                  // same shapes (training)
  float* view_in = nullptr;                   // [n, view_in_dim]
  float* feat = nullptr;                      // fp32 path: [n*Smax, feat_dim]
  float* act[3] = {nullptr, nullptr, nullptr};// fp32 path: [n*Smax, max width]
  float* ray_stats = nullptr;                 // [n, 4 + L]
  float* scalars = nullptr;                   // [64] device scalars (denominators, norms, ...)
  int64_t* tensor_ends = nullptr;             // [tensors.size()] end offset of every parameter tensor (per-tensor statistics)
  float* frame_ws = nullptr;                  // hugs_render_frame: ray workspace [max_rays, 15] (allocated by its first call)
  std::vector<void*> allocs;
  hugs::TcState* tc = nullptr;
  // ---- optional CUDA-event profiling of kernel classes (hugs_profile_enable / hugs_profile_read) ----
  bool prof_on = false;
  struct ProfRec { int cls; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }

  uint64_t train_rng_key = 0;                 // hugs_set_train_rng: in-kernel jitter draws when no jitter tensor is passed
  cudaEvent_t grad_ready_event = nullptr;     // hugs_set_grad_ready_event: recorded after the NeRF level's backward
  const float* cur_params = nullptr;          // parameters of the call in flight
  const int32_t* cur_embed_idx = nullptr;

  int samples(int level) const { return level < d.num_levels - 1 ? d.num_prop_samples : d.num_nerf_samples; }
};

namespace hugs {
// Brackets the kernels launched in its scope with two events on `st` when profiling is enabled.
struct ProfScope {
  hugs_handle* h; cudaStream_t st; int cls; cudaEvent_t a{}, b{};
  ProfScope(hugs_handle* h_, int cls_, cudaStream_t st_) : h(h_), st(st_), cls(cls_) {
    if (h->prof_on) { a = h->prof_event(); b = h->prof_event(); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (h->prof_on) { cudaEventRecord(b, st); h->prof_recs.push_back({cls, a, b}); }
  }
};
}  // namespace hugs
