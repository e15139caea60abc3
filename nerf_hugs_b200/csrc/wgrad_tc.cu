// Weight / bias gradients of the tensor-core path.
//
//   dW_l[k_in, n_out] = sum_s A_l[s, k_in] * dZ_l[s, n_out]
//
// A_l (saved forward activations / IPE features) and dZ_l (saved by the backward chain) are
// row-major [samples, columns] bf16 tensors in HBM, i.e. *MN-major* operands for a GEMM whose
// reduction dimension is the sample index.  TMA loads 64-sample x 64-column boxes (128B swizzle),
// tcgen05.mma consumes them through MN-major shared-memory descriptors, and a 256 x N fp32
// accumulator (two M=128 blocks, all 512 TMEM columns for N = 256) lives in TMEM for the whole
// sample range of a work item; it is flushed once with red.global.add.f32 into the flat gradient.
//
// The remaining tiny reductions (biases = column sums of dZ, the N<=3 head kernels, the per-ray
// view-direction / GLO inputs of the view layer) run on CUDA cores.
//
// Reference semantics: jax.value_and_grad of train_utils.py:407-455 restricted to the Dense layers of
// models.py:449-519.
#include <algorithm>
#include <vector>

#include "ptx.cuh"
#include "tc_internal.h"

namespace hugs {

struct WgItem {
  int a_map;        // 0: saved activations, 1: features
  int a_row0;       // first row of this level in the A tensor
  int a_col0;       // first A column of the 256-wide superblock
  int b_row0;       // first row of the dZ slot
  int n;            // dZ columns (256 | 128)
  int st0, st1;     // [st0, st1) 64-sample stages
  int out;          // kernel columns
  long long koff;   // kernel offset in the flat gradient
  int in_base;      // kernel row of A column a_col0 (feature mode: first feature row)
  int feat_mode;    // 1: A columns are features in engine order -> permute rows on flush
  int pad;
};

struct WgState {
  CUtensorMap map_act64, map_feat64, map_dz64;
  std::vector<std::vector<WgItem>> host;   // per level
  std::vector<WgItem*> dev;                // per level
  std::vector<int> built_for;              // n_samples the list was built for
  float* dzv_ray = nullptr;                // [max_rays, 128]
};

namespace {

constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;       // A0 16K | A1 16K | B 32K
constexpr int kWgThreads = 192;
constexpr int kWgSmem = 1024 + kWgStages * kWgStageBytes + 256;

struct alignas(64) WgParams {
  CUtensorMap map_act64, map_feat64, map_dz64;
  const WgItem* items;
  int n_items, nb, ndeg, feat_dim;
  float* grad;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kWgStages * kWgStageBytes);
  uint64_t* full = bars; uint64_t* empty = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages; uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.map_act64); ptx::prefetch_tmap(&p.map_feat64); ptx::prefetch_tmap(&p.map_dz64);
    for (int i = 0; i < kWgStages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(acc_full, 1); ptx::mbar_init(acc_empty, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const WgItem w = p.items[it];
        const CUtensorMap* amap = w.a_map ? &p.map_feat64 : &p.map_act64;
        const int nb_atoms = w.n >> 6;
        for (int st = w.st0; st < w.st1; ++st) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* s = base + stage * kWgStageBytes;
          ptx::mbar_expect_tx(&full[stage], (4 + nb_atoms) * 8192);
          for (int a = 0; a < 4; ++a)
            ptx::tma_load_2d(s + a * 8192, amap, &full[stage], w.a_col0 + a * 64, w.a_row0 + st * 64);
          for (int a = 0; a < nb_atoms; ++a)
            ptx::tma_load_2d(s + 32768 + a * 8192, &p.map_dz64, &full[stage], a * 64, w.b_row0 + st * 64);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, ae_phase = 0;
      bool first_item = true;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const WgItem w = p.items[it];
        const uint32_t idesc = ptx::make_idesc_bf16(128, w.n, 1, 1);
        if (!first_item) { ptx::mbar_wait(acc_empty, ae_phase); ae_phase ^= 1; }
        first_item = false;
        ptx::tc_fence_after();
        for (int st = w.st0; st < w.st1; ++st) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t s = ptx::smem_u32(base + stage * kWgStageBytes);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint64_t db = ptx::make_desc_sw128(s + 32768 + k16 * 2048, 8192, 1024);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint64_t da = ptx::make_desc_sw128(s + mb * 16384 + k16 * 2048, 8192, 1024);
              ptx::mma_bf16_ss(tmem_base + mb * 256, da, db, idesc, (st > w.st0 || k16 > 0) ? 1u : 0u);
            }
          }
          ptx::mma_commit(&empty[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(acc_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t af_phase = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const WgItem w = p.items[it];
      ptx::mbar_wait(acc_full, af_phase); af_phase ^= 1;
      ptx::tc_fence_after();
      if (w.st1 > w.st0) {
#pragma unroll 1
        for (int mb = 0; mb < 2; ++mb) {
          const int m = mb * 128 + row;
          int krow;
          if (w.feat_mode) {
            const int fp = w.a_col0 + m;
            krow = fp < p.feat_dim ? w.in_base + ref_feature_col(fp, p.nb, p.ndeg) : -1;
          } else {
            krow = w.in_base + m;
          }
#pragma unroll 1
          for (int c = 0; c < w.n; c += 32) {
            uint32_t r[32];
            ptx::tmem_ld32(lane_addr + (uint32_t)(mb * 256 + c), r);
            ptx::tmem_ld_wait();
            if (krow >= 0) {
              float* dst = p.grad + w.koff + (long long)krow * w.out + c;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c + j < w.out) atomicAdd(dst + j, __uint_as_float(r[j]));
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(acc_empty);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------- CUDA-core reductions
struct ColsumJob { int row0; int cols; long long boff; };
struct ColsumArgs { ColsumJob jobs[16]; int n_jobs; int n_rows; const __nv_bfloat16* dz; float* grad; };

// grad[boff + c] += sum_s dz[row0 + s][c]     (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(ColsumArgs a) {
  const ColsumJob j = a.jobs[blockIdx.y];
  const int c = threadIdx.x;
  const int chunk = (a.n_rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * chunk, r1 = min(r0 + chunk, a.n_rows);
  if (c >= j.cols) return;
  float acc = 0.f;
  const __nv_bfloat16* src = a.dz + (size_t)j.row0 * kW + c;
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(src[(size_t)r * kW]);
  if (r1 > r0) atomicAdd(a.grad + j.boff + c, acc);
}

struct HeadArgs {
  const __nv_bfloat16* a_last;   // [n, 256] last trunk activation (density head input)
  const __nv_bfloat16* v_act;    // [n, 256] view activation in cols [0,128) (nullptr: proposal MLP)
  const __nv_bfloat16* dhead;    // [n, 16]: (d_r, d_g, d_b, d_density, 0...)
  int n_rows;
  long long dens_koff, dens_boff, rgb_koff, rgb_boff;
  float* grad;
};

__global__ void __launch_bounds__(256) head_wgrad_kernel(HeadArgs a) {
  const int c = threadIdx.x;
  const int chunk = (a.n_rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * chunk, r1 = min(r0 + chunk, a.n_rows);
  float wd = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f, bd = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const __nv_bfloat16* dh = a.dhead + (size_t)r * 16;
    const float dd = __bfloat162float(dh[3]);
    wd += __bfloat162float(a.a_last[(size_t)r * kW + c]) * dd;
    if (c == 0) bd += dd;
    if (a.v_act) {
      const float d0 = __bfloat162float(dh[0]), d1 = __bfloat162float(dh[1]), d2 = __bfloat162float(dh[2]);
      if (c < 128) {
        const float v = __bfloat162float(a.v_act[(size_t)r * kW + c]);
        w0 += v * d0; w1 += v * d1; w2 += v * d2;
      }
      if (c == 0) { b0 += d0; b1 += d1; b2 += d2; }
    }
  }
  if (r1 <= r0) return;
  atomicAdd(a.grad + a.dens_koff + c, wd);
  if (c == 0) atomicAdd(a.grad + a.dens_boff, bd);
  if (a.v_act) {
    if (c < 128) {
      atomicAdd(a.grad + a.rgb_koff + c * 3 + 0, w0);
      atomicAdd(a.grad + a.rgb_koff + c * 3 + 1, w1);
      atomicAdd(a.grad + a.rgb_koff + c * 3 + 2, w2);
    }
    if (c == 0) {
      atomicAdd(a.grad + a.rgb_boff + 0, b0); atomicAdd(a.grad + a.rgb_boff + 1, b1);
      atomicAdd(a.grad + a.rgb_boff + 2, b2);
    }
  }
}

// dzv_ray[ray][c] = sum over the ray's samples of dZ_view[s][c]
__global__ void __launch_bounds__(128) ray_sum_kernel(const __nv_bfloat16* dzv, int n_rays, int S, float* out) {
  const int ray = blockIdx.x, c = threadIdx.x;
  if (ray >= n_rays) return;
  float acc = 0.f;
  const __nv_bfloat16* src = dzv + (size_t)ray * S * kW + c;
  for (int s = 0; s < S; ++s) acc += __bfloat162float(src[(size_t)s * kW]);
  out[(size_t)ray * 128 + c] = acc;
}

// dW_view[bott + j][c] += sum_ray bf16(view_in[ray][j]) * dzv_ray[ray][c]; block = input row j, thread = c
__global__ void __launch_bounds__(128) view_extra_wgrad_kernel(const float* view_in, int view_in_dim,
                                                               const float* dzv_ray, int n_rays, long long koff,
                                                               int bott_w, float* grad) {
  const int j = blockIdx.x, c = threadIdx.x;
  float acc = 0.f;
  for (int r = 0; r < n_rays; ++r)
    acc = fmaf(__bfloat162float(__float2bfloat16(view_in[(size_t)r * view_in_dim + j])), dzv_ray[(size_t)r * 128 + c], acc);
  atomicAdd(grad + koff + (long long)(bott_w + j) * 128 + c, acc);
}

// GLO embedding rows: d_embed[idx[ray]][g] += sum_c dzv_ray[ray][c] * bf16(W_view[bott + dir_dim + g][c])
__global__ void glo_grad_kernel(const float* dzv_ray, const int32_t* embed_idx, const float* params,
                                long long view_koff, int bott_w, int dir_dim, int glo, int n_rays,
                                long long glo_off, float* grad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * glo) return;
  const int ray = idx / glo, g = idx % glo;
  const float* wrow = params + view_koff + (long long)(bott_w + dir_dim + g) * 128;
  float acc = 0.f;
  for (int c = 0; c < 128; ++c)
    acc = fmaf(dzv_ray[(size_t)ray * 128 + c], __bfloat162float(__float2bfloat16(wrow[c])), acc);
  atomicAdd(grad + glo_off + (long long)embed_idx[ray] * glo + g, acc);
}

}  // namespace

// ------------------------------------------------------------------------------------------
int wgrad_create(hugs_handle* h) {
  TcState* tc = h->tc;
  WgState* w = new WgState();
  tc->wg = w;
  int rc;
  if ((rc = make_map(&w->map_act64, tc->act, tc->total_save_rows, kW, 64)) ||
      (rc = make_map(&w->map_dz64, tc->dz, tc->total_save_rows, kW, 64)) ||
      (rc = make_map(&w->map_feat64, tc->feat, tc->total_feat_rows, kFeatPad, 64)))
    return rc;
  const int L = h->d.num_levels;
  w->host.resize(L); w->dev.assign(L, nullptr); w->built_for.assign(L, -1);
  for (int l = 0; l < L; ++l) {
    void* q = nullptr;
    HUGS_CUDA(cudaMalloc(&q, sizeof(WgItem) * 1024));
    h->allocs.push_back(q);
    w->dev[l] = static_cast<WgItem*>(q);
  }
  void* q = nullptr;
  HUGS_CUDA(cudaMalloc(&q, sizeof(float) * (size_t)h->d.max_rays * 128));
  h->allocs.push_back(q);
  w->dzv_ray = static_cast<float*>(q);
  HUGS_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
  return HUGS_OK;
}

void wgrad_destroy(hugs_handle* h) {
  if (h->tc && h->tc->wg) { delete h->tc->wg; h->tc->wg = nullptr; }
}

static void build_items(hugs_handle* h, int level, int n_tiles, std::vector<WgItem>* items) {
  TcState* tc = h->tc;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int D = mv.depth;
  const int cap = tc->cap[level], srow = tc->save_row0[level], frow = tc->feat_row0[level];
  const int T = n_tiles * 2;   // 64-sample stages
  struct Unit { WgItem w; float cost; };
  std::vector<Unit> units;
  auto add = [&](int a_map, int a_row0, int a_col0, int b_slot, int n, const DenseView& v, int in_base, int feat_mode) {
    WgItem w{};
    w.a_map = a_map; w.a_row0 = a_row0; w.a_col0 = a_col0; w.b_row0 = srow + b_slot * cap; w.n = n;
    w.out = v.out; w.koff = v.kernel_off; w.in_base = in_base; w.feat_mode = feat_mode;
    units.push_back({w, n / 256.f});
  };
  bool cat = false;
  for (int l = 0; l < D; ++l) {
    const DenseView& v = mv.dense[l];
    if (l == 0) {
      for (int sb = 0; sb < kFeatPad / 256; ++sb) add(1, frow, sb * 256, 0, 256, v, 0, 1);
    } else {
      add(0, srow + (l - 1) * cap, 0, l, 256, v, 0, 0);
      if (cat) for (int sb = 0; sb < kFeatPad / 256; ++sb) add(1, frow, sb * 256, l, 256, v, kW, 1);
    }
    cat = (l % d.skip_layer == 0 && l > 0);
  }
  if (mv.has_rgb) {
    add(0, srow + (D - 1) * cap, 0, D, 256, mv.dense[D + 1], 0, 0);        // bottleneck
    add(0, srow + D * cap, 0, D + 1, 128, mv.dense[D + 2], 0, 0);          // view layer (bottleneck rows)
  }
  float total = 0.f;
  for (auto& u : units) total += u.cost;
  items->clear();
  for (auto& u : units) {
    int splits = std::max(1, (int)(tc->num_sms * u.cost / total));
    splits = std::min(splits, std::max(1, T / 4));
    for (int k = 0; k < splits; ++k) {
      WgItem w = u.w;
      w.st0 = (int)((long long)T * k / splits);
      w.st1 = (int)((long long)T * (k + 1) / splits);
      if (w.st1 > w.st0) items->push_back(w);
    }
  }
}

int wgrad_run(hugs_handle* h, int level, int n_rays, float* grad, cudaStream_t st) {
  TcState* tc = h->tc;
  WgState* w = tc->wg;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int D = mv.depth, S = h->samples(level);
  const int n_samples = n_rays * S;
  const int n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int n_rows = n_tiles * kTileM;
  const int cap = tc->cap[level], srow = tc->save_row0[level];
  if (w->built_for[level] != n_samples) {
    build_items(h, level, n_tiles, &w->host[level]);
    HUGS_REQUIRE(w->host[level].size() <= 1024, "wgrad: too many work items");
    HUGS_CUDA(cudaMemcpyAsync(w->dev[level], w->host[level].data(), sizeof(WgItem) * w->host[level].size(),
                              cudaMemcpyHostToDevice, st));
    HUGS_CUDA(cudaStreamSynchronize(st));
    w->built_for[level] = n_samples;
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.map_act64 = w->map_act64; p.map_feat64 = w->map_feat64; p.map_dz64 = w->map_dz64;
  p.items = w->dev[level]; p.n_items = (int)w->host[level].size();
  p.nb = d.num_basis; p.ndeg = d.max_deg_point - d.min_deg_point; p.feat_dim = h->feat_dim; p.grad = grad;
  {
    ProfScope ps(h, is_prop ? HUGS_K_WGRAD_PROP : HUGS_K_WGRAD_NERF, st);
    wgrad_kernel<<<std::min(p.n_items, tc->num_sms), kWgThreads, kWgSmem, st>>>(p);
    HUGS_LAUNCH_CHECK();
  }
  ProfScope ps_red(h, HUGS_K_REDUCTIONS, st);

  // biases: column sums of the saved dZ slots
  ColsumArgs ca;
  memset(&ca, 0, sizeof(ca));
  for (int l = 0; l < D; ++l) ca.jobs[ca.n_jobs++] = ColsumJob{srow + l * cap, 256, mv.dense[l].bias_off};
  if (mv.has_rgb) {
    ca.jobs[ca.n_jobs++] = ColsumJob{srow + D * cap, 256, mv.dense[D + 1].bias_off};
    ca.jobs[ca.n_jobs++] = ColsumJob{srow + (D + 1) * cap, 128, mv.dense[D + 2].bias_off};
  }
  ca.n_rows = n_rows; ca.dz = tc->dz; ca.grad = grad;
  const int chunks = std::max(1, std::min(64, n_rows / 256));
  colsum_kernel<<<dim3(chunks, ca.n_jobs), 256, 0, st>>>(ca);
  HUGS_LAUNCH_CHECK();

  // density / rgb heads
  HeadArgs ha;
  memset(&ha, 0, sizeof(ha));
  ha.a_last = tc->act + (size_t)(srow + (D - 1) * cap) * kW;
  ha.v_act = mv.has_rgb ? tc->act + (size_t)(srow + (D + 1) * cap) * kW : nullptr;
  ha.dhead = tc->drgb; ha.n_rows = n_samples;
  ha.dens_koff = mv.dense[D].kernel_off; ha.dens_boff = mv.dense[D].bias_off;
  if (mv.has_rgb) { ha.rgb_koff = mv.dense[D + 3].kernel_off; ha.rgb_boff = mv.dense[D + 3].bias_off; }
  ha.grad = grad;
  head_wgrad_kernel<<<std::max(1, std::min(296, n_samples / 128)), 256, 0, st>>>(ha);
  HUGS_LAUNCH_CHECK();

  if (mv.has_rgb) {
    const DenseView& vv = mv.dense[D + 2];
    ray_sum_kernel<<<n_rays, 128, 0, st>>>(tc->dz + (size_t)(srow + (D + 1) * cap) * kW, n_rays, S, w->dzv_ray);
    HUGS_LAUNCH_CHECK();
    view_extra_wgrad_kernel<<<h->view_in_dim, 128, 0, st>>>(h->view_in, h->view_in_dim, w->dzv_ray, n_rays,
                                                           vv.kernel_off, d.bottleneck_width, grad);
    HUGS_LAUNCH_CHECK();
    if (d.num_glo_features > 0) {
      const int tot = n_rays * d.num_glo_features;
      glo_grad_kernel<<<(tot + 127) / 128, 128, 0, st>>>(w->dzv_ray, h->cur_embed_idx, h->cur_params, vv.kernel_off,
                                                         d.bottleneck_width, 3 + 6 * d.deg_view, d.num_glo_features,
                                                         n_rays, h->glo_off, grad);
      HUGS_LAUNCH_CHECK();
    }
  }
  return HUGS_OK;
}

}  // namespace hugs
