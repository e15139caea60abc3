// Layer-at-a-time tensor-core path for the NerfMLP at widths the chain kernel cannot keep in shared memory
// (NerfMLP.net_width = 512 / 1024: MipNeRF360/configs/360.gin:14-16 and every other non-debug gin).
//
// Every Dense layer of models.py:449-519 is one launch of dense_tc_kernel (dense_tc.cu); activations travel through
// HBM / L2 as bf16 [samples, width] tensors (1024 FLOP per byte at width 1024: still tensor-bound).  The backward pass
// interleaves one dgrad GEMM and one weight-gradient launch (wgrad_kernel, wgrad_tc.cu) per layer, so only two dZ
// buffers exist.  The PropMLP (width 256 in every gin) stays on the chain kernel.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "dense_tc.h"

namespace hugs {

enum { LW_ACT = 0, LW_FEAT = 1, LW_DZ0 = 2, LW_DZ1 = 3, LW_BOTT = 4, LW_VACT = 5, LW_DZB = 6, LW_DZV = 7, LW_DH = 8,
       LW_MAPS = 9 };

struct LayeredMlp {
  int W = 0, D = 0, level = 0, cap = 0, skip = 4;
  bool split = false;             // HUGS_PRECISION_TC_SPLIT: every bf16 tensor holds a hi half and, `lo_*` rows further down, a lo half
  int parts = 1;
  int rows_f = 0, rows_b = 0, kmax = 0, tab_floats = 0;
  std::vector<int> row_f, row_b, bias_off;        // per dense index (flax order: trunk..., density, bottleneck, view, rgb)
  int heads_row_f = 0, heads_bias_off = 0, w_dens_off = 0, w_rgb_off = 0;
  __nv_bfloat16 *wt = nullptr, *wn = nullptr;
  float* tab = nullptr;
  CUtensorMap map_wt128, map_wt64, map_wt8, map_wn128;
  __nv_bfloat16 *act = nullptr, *bott = nullptr, *vact = nullptr;
  int act_slots = 0;
  __nv_bfloat16 *dz[2] = {nullptr, nullptr}, *dz_bott = nullptr, *dz_view = nullptr, *drgb = nullptr;
  uint32_t* gate = nullptr;       // ReLU gate bit masks of the trunk activations, [D][cap][W / 32] (training)
  CUtensorMap map_act, map_bott, map_vact, map_dz[2], map_dzb, map_dzv;
  CUtensorMap wg_maps[LW_MAPS];
  bool train_ready = false;
  // weight-gradient work items of one training step, per launch
  WgItem* items_dev = nullptr;
  std::vector<WgItem> items_host;
  std::vector<std::pair<int, int>> launches;      // (first item, count) in backward order
  std::vector<bool> launch_pairs;                 // the launch runs on CTA pairs (wgrad2_kernel)
  bool wg_pairs = getenv("HUGS_WGRAD_PAIRS") ? atoi(getenv("HUGS_WGRAD_PAIRS")) != 0 : true;   // development switch
  int built_for = -1;
  float* dzv_ray = nullptr;
};

namespace {

template <class T>
int lalloc(hugs_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return HUGS_ERR_NOMEM;
  }
  h->allocs.push_back(q);
  *p = static_cast<T*>(q);
  HUGS_CUDA(cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(T)));
  return HUGS_OK;
}

struct LPackEntry {
  int row0, rows, out, out_stride, x_in, feat_in;   // forward rows [row0, row0 + rows): output units; K = [x | features]
  long long koff, boff;
  int bias_off;
  int brow0, b_in, b_out;                            // backward block: rows = inputs [0, b_in), cols = outputs; -1: none
};

struct LPackArgs {
  LPackEntry e[20];
  int n, kmax, W, nb, ndeg, feat_dim, rows_f, rows_b, tab_floats, w_dens_off, w_rgb_off, dens_in, rgb_in;
  long long dens_koff, rgb_koff;
  const float* params;
  __nv_bfloat16 *wt, *wn;
  float* tab;
  int part;         // 0: bf16(w) (+ tables); 1: bf16(w - bf16(w)) into the lo half of wt / wn (split-precision mode)
  int exact_heads;  // split-precision mode: the fp32 head-weight tables are not rounded to bf16
};

__device__ __forceinline__ __nv_bfloat16 lpack_part(float v, int part) {
  const __nv_bfloat16 hi = __float2bfloat16(v);
  return part == 0 ? hi : __float2bfloat16(v - __bfloat162float(hi));
}

// Forward pack Wt[r][k] = W[in(k)][n(r)]: a transpose of the flax kernels.  32 x 32 tiles through shared memory: the reads run
// along the output unit n (contiguous in the [in, out] kernel), the writes along k.
__global__ void __launch_bounds__(256) layered_pack_wt_kernel(LPackArgs a) {
  __shared__ float tile[32][33];
  const int kt = blockIdx.x * 32, rt = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long nf = (long long)a.rows_f * a.kmax;
  {
    const int r = rt + tx;
    int li = -1;
    for (int l = 0; l < a.n; ++l)
      if (r >= a.e[l].row0 && r < a.e[l].row0 + a.e[l].rows) { li = l; break; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kt + ty + 8 * j;
      float v = 0.f;
      if (li >= 0 && k < a.kmax) {
        const LPackEntry& L = a.e[li];
        const int n = r - L.row0;
        if (n < L.out) {
          int in = -1;
          if (k < L.x_in) in = k;
          else if (L.feat_in > 0 && k < L.x_in + kFeatPad) {
            const int fp = k - L.x_in;
            if (fp < a.feat_dim) in = L.x_in + ref_feature_col(fp, a.nb, a.ndeg);
          }
          if (in >= 0) v = a.params[L.koff + (long long)in * L.out_stride + n];
        }
      }
      tile[ty + 8 * j][tx] = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = rt + ty + 8 * j, k = kt + tx;
    if (r < a.rows_f && k < a.kmax) a.wt[(long long)r * a.kmax + k + a.part * nf] = lpack_part(tile[tx][ty + 8 * j], a.part);
  }
}

// backward pack Wn (the kernels' own [in, out] layout) and the fp32 tables
__global__ void layered_pack_kernel(LPackArgs a) {
  const long long nf = (long long)a.rows_f * a.kmax, nbk = (long long)a.rows_b * a.W;
  const long long total = a.part == 0 ? nf + nbk + a.tab_floats : nf + nbk;
  for (long long i = nf + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (i < nf + nbk) {
      const long long q = i - nf;
      const int r = (int)(q / a.W), c = (int)(q % a.W);
      float v = 0.f;
      for (int l = 0; l < a.n; ++l) {
        const LPackEntry& L = a.e[l];
        if (L.brow0 < 0 || r < L.brow0 || r >= L.brow0 + L.b_in) continue;
        if (c < L.b_out) v = a.params[L.koff + (long long)(r - L.brow0) * L.out_stride + c];
        break;
      }
      a.wn[q + a.part * nbk] = lpack_part(v, a.part);
    } else {
      const int q = (int)(i - nf - nbk);
      float v = 0.f;
      for (int l = 0; l < a.n; ++l) {
        const LPackEntry& L = a.e[l];
        if (q >= L.bias_off && q < L.bias_off + L.out) { v = a.params[L.boff + (q - L.bias_off)]; break; }
      }
      if (q >= a.w_dens_off && q < a.w_dens_off + a.dens_in) {
        v = a.params[a.dens_koff + (q - a.w_dens_off)];
        if (!a.exact_heads) v = __bfloat162float(__float2bfloat16(v));
      }
      if (q >= a.w_rgb_off && q < a.w_rgb_off + a.rgb_in * 3) {
        v = a.params[a.rgb_koff + (q - a.w_rgb_off)];
        if (!a.exact_heads) v = __bfloat162float(__float2bfloat16(v));
      }
      a.tab[q] = v;
    }
  }
}

}  // namespace

int layered_create(hugs_handle* h, const MlpViews& mv, int level, LayeredMlp** out) {
  const hugs_model_desc& d = h->d;
  TcState* tc = h->tc;
  HUGS_REQUIRE(mv.has_rgb, "layer-at-a-time path: NerfMLP only");
  HUGS_REQUIRE(mv.width % 256 == 0 && mv.width >= 256 && mv.width <= 256 * kMaxNTiles,
               "layer-at-a-time path: net_width must be a multiple of 256 up to %d (got %d)", 256 * kMaxNTiles, mv.width);
  HUGS_REQUIRE(d.bottleneck_width == 256 && d.view_width == 128, "layer-at-a-time path: bottleneck 256 / view width 128");
  LayeredMlp* m = new LayeredMlp();
  *out = m;
  m->W = mv.width; m->D = mv.depth; m->level = level; m->cap = tc->cap[level]; m->skip = d.skip_layer;
  m->split = tc->split; m->parts = tc->split ? 2 : 1;
  const size_t parts = (size_t)m->parts;
  m->kmax = m->W + kFeatPad;
  const int D = m->D, W = m->W;
  const int nd = (int)mv.dense.size();
  m->row_f.assign(nd, 0); m->row_b.assign(nd, -1); m->bias_off.assign(nd, 0);
  int rf = 0, rb = 0, bo = 0;
  for (int l = 0; l < D; ++l) {
    m->row_f[l] = rf; rf += W;
    m->bias_off[l] = bo; bo += W;
    if (l > 0) { m->row_b[l] = rb; rb += W; }
  }
  // heads block: bottleneck rows [0, 256), density row 256 (+ 15 zero rows); one bias table of 272 entries
  m->heads_row_f = rf; m->row_f[D + 1] = rf; m->row_f[D] = rf + 256; rf += 272;
  m->heads_bias_off = bo; m->bias_off[D + 1] = bo; m->bias_off[D] = bo + 256; bo += 272;
  m->row_b[D + 1] = rb; rb += W;                      // bottleneck dgrad: rows = its W inputs, cols = 256 outputs
  m->row_f[D + 2] = rf; rf += 128;                    // view layer
  m->bias_off[D + 2] = bo; bo += 128;
  m->row_b[D + 2] = rb; rb += 256;                    // view dgrad: rows = the 256 bottleneck inputs, cols = 128 outputs
  m->row_f[D + 3] = rf; rf += 16;                     // rgb head
  m->bias_off[D + 3] = bo; bo += 16;
  m->w_dens_off = bo; bo += W;
  m->w_rgb_off = bo; bo += 128 * 3 + 16;
  m->rows_f = ((rf + 127) / 128) * 128; m->rows_b = rb; m->tab_floats = bo;
  int rc;
  if ((rc = lalloc(h, &m->wt, parts * m->rows_f * m->kmax)) || (rc = lalloc(h, &m->wn, parts * m->rows_b * W)) ||
      (rc = lalloc(h, &m->tab, (size_t)m->tab_floats)))
    return rc;
  const long long pf = (long long)parts * m->rows_f, pb = (long long)parts * m->rows_b;
  if ((rc = make_map(&m->map_wt128, m->wt, pf, m->kmax, 128)) || (rc = make_map(&m->map_wt64, m->wt, pf, m->kmax, 64)) ||
      (rc = make_map(&m->map_wt8, m->wt, pf, m->kmax, 8)) || (rc = make_map(&m->map_wn128, m->wn, pb, W, 128)))
    return rc;
  // inference buffers: two ping-pong activation slots, bottleneck, view activation (hi half, then the lo half)
  m->act_slots = 2;
  if ((rc = lalloc(h, &m->act, parts * m->act_slots * m->cap * W)) || (rc = lalloc(h, &m->bott, parts * m->cap * 256)) ||
      (rc = lalloc(h, &m->vact, parts * m->cap * 128)))
    return rc;
  if ((rc = make_map(&m->map_act, m->act, (long long)parts * m->act_slots * m->cap, W, 128)) ||
      (rc = make_map(&m->map_bott, m->bott, (long long)parts * m->cap, 256, 128)) ||
      (rc = make_map(&m->map_vact, m->vact, (long long)parts * m->cap, 128, 128)))
    return rc;
  return dense_tc_init();
}

void layered_destroy(LayeredMlp* m) { delete m; }

int layered_pack(hugs_handle* h, LayeredMlp* m, const float* params, cudaStream_t st) {
  const MlpViews& mv = h->nerf;
  const int D = m->D, W = m->W;
  LPackArgs a;
  memset(&a, 0, sizeof(a));
  HUGS_REQUIRE((int)mv.dense.size() <= 20, "too many layers to pack");
  bool cat = false;
  for (int li = 0; li < (int)mv.dense.size(); ++li) {
    const DenseView& v = mv.dense[li];
    LPackEntry& e = a.e[a.n++];
    e.row0 = m->row_f[li]; e.out = v.out; e.out_stride = v.out; e.koff = v.kernel_off; e.boff = v.bias_off;
    e.bias_off = m->bias_off[li]; e.brow0 = m->row_b[li]; e.b_in = 0; e.b_out = v.out;
    if (li < D) {
      e.rows = W;
      if (li == 0) { e.x_in = 0; e.feat_in = h->feat_dim; }
      else { e.x_in = W; e.feat_in = cat ? h->feat_dim : 0; e.b_in = W; }
      cat = (li % m->skip == 0 && li > 0);
    } else if (li == D) { e.rows = 16; e.x_in = W; e.feat_in = 0; }                       // density head
    else if (li == D + 1) { e.rows = 256; e.x_in = W; e.feat_in = 0; e.b_in = W; }       // bottleneck
    else if (li == D + 2) { e.rows = 128; e.x_in = 256; e.feat_in = 0; e.b_in = 256; }   // view layer (bottleneck rows)
    else { e.rows = 16; e.x_in = 128; e.feat_in = 0; }                                    // rgb head
  }
  a.kmax = m->kmax; a.W = W; a.nb = h->perm_nb; a.ndeg = h->d.max_deg_point - h->d.min_deg_point;
  a.feat_dim = h->feat_dim; a.rows_f = m->rows_f; a.rows_b = m->rows_b; a.tab_floats = m->tab_floats;
  a.w_dens_off = m->w_dens_off; a.w_rgb_off = m->w_rgb_off; a.dens_in = W; a.rgb_in = 128;
  a.dens_koff = mv.dense[D].kernel_off; a.rgb_koff = mv.dense[D + 3].kernel_off;
  a.params = params; a.wt = m->wt; a.wn = m->wn; a.tab = m->tab;
  a.exact_heads = m->split ? 1 : 0;
  for (int part = 0; part < m->parts; ++part) {
    a.part = part;
    layered_pack_wt_kernel<<<dim3((m->kmax + 31) / 32, (m->rows_f + 31) / 32), 256, 0, st>>>(a);
    HUGS_LAUNCH_CHECK();
    layered_pack_kernel<<<1024, 256, 0, st>>>(a);
    HUGS_LAUNCH_CHECK();
  }
  return HUGS_OK;
}

int layered_ensure_training(hugs_handle* h, LayeredMlp* m) {
  if (m->train_ready) return HUGS_OK;
  TcState* tc = h->tc;
  const int W = m->W, D = m->D;
  const size_t parts = (size_t)m->parts;
  const long long pc = (long long)parts * m->cap;
  int rc;
  // every trunk activation is kept for the backward pass
  m->act_slots = D;
  if ((rc = lalloc(h, &m->act, parts * m->act_slots * m->cap * W))) return rc;
  if ((rc = make_map(&m->map_act, m->act, (long long)parts * m->act_slots * m->cap, W, 128))) return rc;
  for (int i = 0; i < 2; ++i) {
    if ((rc = lalloc(h, &m->dz[i], parts * m->cap * W))) return rc;
    if ((rc = make_map(&m->map_dz[i], m->dz[i], pc, W, 128))) return rc;
  }
  if ((rc = lalloc(h, &m->gate, (size_t)D * m->cap * (W / 32)))) return rc;
  if ((rc = lalloc(h, &m->dz_bott, parts * m->cap * 256)) || (rc = lalloc(h, &m->dz_view, parts * m->cap * 128)) ||
      (rc = lalloc(h, &m->drgb, parts * m->cap * kHeadCols)) || (rc = lalloc(h, &m->items_dev, 4096 * 4)))
    return rc;
  if ((rc = make_map(&m->map_dzb, m->dz_bott, pc, 256, 128)) || (rc = make_map(&m->map_dzv, m->dz_view, pc, 128, 128)))
    return rc;
  // 64-sample boxes for the weight-gradient kernel
  if ((rc = make_map(&m->wg_maps[LW_ACT], m->act, (long long)parts * m->act_slots * m->cap, W, 64)) ||
      (rc = make_map(&m->wg_maps[LW_FEAT], tc->feat, (long long)parts * tc->total_feat_rows, kFeatPad, 64)) ||
      (rc = make_map(&m->wg_maps[LW_DZ0], m->dz[0], pc, W, 64)) || (rc = make_map(&m->wg_maps[LW_DZ1], m->dz[1], pc, W, 64)) ||
      (rc = make_map(&m->wg_maps[LW_BOTT], m->bott, pc, 256, 64)) || (rc = make_map(&m->wg_maps[LW_VACT], m->vact, pc, 128, 64)) ||
      (rc = make_map(&m->wg_maps[LW_DZB], m->dz_bott, pc, 256, 64)) || (rc = make_map(&m->wg_maps[LW_DZV], m->dz_view, pc, 128, 64)) ||
      (rc = make_map(&m->wg_maps[LW_DH], m->drgb, pc, kHeadCols, 64)))
    return rc;
  HUGS_CUDA(cudaDeviceSynchronize());
  m->train_ready = true;
  return HUGS_OK;
}

namespace {

void fill_common(DenseParams* p, const LayeredMlp* m, int n_samples) {
  memset(p, 0, sizeof(*p));
  p->b_map = m->map_wt128; p->b_map_64 = m->map_wt64; p->b_map_8 = m->map_wt8;
  p->m_rows = n_samples; p->m_tiles = (n_samples + 255) / 256;
  p->split = m->split ? 1 : 0;
}

// One source of A columns: `kp` K panels of tensor `map` from (row0, col0), multiplied with weight K columns from wcol0;
// `lo_rows` = row distance of the tensor's lo half (split-precision mode).
struct SegSrc { const CUtensorMap* map; int kp, row0, col0, wcol0, lo_rows; };

// A . W as K-concatenated segments; split-precision mode: A_hi W_hi + A_lo W_hi + A_hi W_lo (the lo . lo term is below fp32
// resolution), `w_lo_rows` = row distance of the lo half of the weight pack.
void set_segs(DenseParams* p, const LayeredMlp* m, const SegSrc* srcs, int n_src, int w_lo_rows) {
  int n = 0;
  const int combos = m->split ? 3 : 1;
  for (int c = 0; c < combos; ++c) {
    const bool a_lo = c == 1, w_lo = c == 2;
    for (int i = 0; i < n_src; ++i) {
      const SegSrc& s = srcs[i];
      p->a_map[n] = *s.map; p->a_kp[n] = s.kp; p->a_row0[n] = s.row0 + (a_lo ? s.lo_rows : 0); p->a_col0[n] = s.col0;
      p->w_col0[n] = s.wcol0; p->w_row_off[n] = w_lo ? w_lo_rows : 0;
      ++n;
    }
  }
  p->n_seg = n;
}

}  // namespace

int layered_forward(hugs_handle* h, LayeredMlp* m, int level, int n_rays, bool training, cudaStream_t st) {
  TcState* tc = h->tc;
  const int S = h->samples(level), M = n_rays * S, W = m->W, D = m->D, cap = m->cap;
  HUGS_REQUIRE(!training || m->train_ready, "layered path: training buffers missing");
  auto slot = [&](int l) { return training ? l : (l & 1); };
  const int lo_act = m->act_slots * cap, lo_feat = tc->total_feat_rows;
  int rc;
  bool cat = false;
  for (int l = 0; l < D; ++l) {
    DenseParams p;
    fill_common(&p, m, M);
    SegSrc src[2]; int ns = 0;
    if (l > 0) src[ns++] = SegSrc{&m->map_act, W / 64, slot(l - 1) * cap, 0, 0, lo_act};
    if (l == 0 || cat) src[ns++] = SegSrc{&tc->map_feat, kFeatPad / 64, tc->feat_row0[level], 0, l == 0 ? 0 : W, lo_feat};
    set_segs(&p, m, src, ns, m->rows_f);
    p.b_row0 = m->row_f[l]; p.b_col0 = 0;
    p.n_tiles = W / 256;
    for (int j = 0; j < p.n_tiles; ++j) { p.tile_n0[j] = j * 256; p.tile_bn[j] = 256; p.tile_epi[j] = DE_RELU; }
    p.bias = m->tab + m->bias_off[l];
    p.out_map = m->map_act; p.out_row0 = slot(l) * cap; p.out_col0 = 0; p.out_lo_row_off = lo_act;
    if (training) { p.gate_out = m->gate; p.gate_ld = W / 32; p.gate_row0 = l * cap; }
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
    cat = (l % m->skip == 0 && l > 0);
  }
  {  // heads: bottleneck (linear, bf16) + raw density (fp32 column)
    DenseParams p;
    fill_common(&p, m, M);
    SegSrc src[1] = {SegSrc{&m->map_act, W / 64, slot(D - 1) * cap, 0, 0, lo_act}};
    set_segs(&p, m, src, 1, m->rows_f);
    p.b_row0 = m->heads_row_f; p.n_tiles = 2;
    p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_LINEAR;
    p.tile_n0[1] = 256; p.tile_bn[1] = 16; p.tile_epi[1] = DE_HEAD_F32;
    p.bias = m->tab + m->heads_bias_off;
    p.out_map = m->map_bott; p.out_row0 = 0; p.out_col0 = 0; p.out_lo_row_off = cap;
    p.raw_out = h->raw[level]; p.raw_c = 4; p.raw_chan0 = 0; p.raw_nchan = 1;
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
  }
  {  // view layer: K = bottleneck; direction encoding / GLO terms arrive as the per-ray bias
    DenseParams p;
    fill_common(&p, m, M);
    SegSrc src[1] = {SegSrc{&m->map_bott, 4, 0, 0, 0, cap}};
    set_segs(&p, m, src, 1, m->rows_f);
    p.b_row0 = m->row_f[D + 2]; p.n_tiles = 1;
    p.tile_n0[0] = 0; p.tile_bn[0] = 128; p.tile_epi[0] = DE_VIEW;
    p.viewbias = tc->viewbias; p.view_ld = 128; p.S = S;
    p.out_map = m->map_vact; p.out_lo_row_off = cap;
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
  }
  {  // rgb head (fp32 columns 1..3 of raw)
    DenseParams p;
    fill_common(&p, m, M);
    SegSrc src[1] = {SegSrc{&m->map_vact, 2, 0, 0, 0, cap}};
    set_segs(&p, m, src, 1, m->rows_f);
    p.b_row0 = m->row_f[D + 3]; p.n_tiles = 1;
    p.tile_n0[0] = 0; p.tile_bn[0] = 16; p.tile_epi[0] = DE_HEAD_F32;
    p.bias = m->tab + m->bias_off[D + 3];
    p.out_map = m->map_vact;
    p.raw_out = h->raw[level]; p.raw_c = 4; p.raw_chan0 = 1; p.raw_nchan = 3;
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
  }
  return HUGS_OK;
}

namespace {

// weight-gradient work of one training step: one launch per layer, in backward order
void build_wgrad(hugs_handle* h, LayeredMlp* m, int level, int n_samples) {
  TcState* tc = h->tc;
  const MlpViews& mv = h->nerf;
  const int W = m->W, D = m->D, cap = m->cap;
  const int T = ((n_samples + 255) / 256) * 4;          // 64-sample stages (rows padded to the 256-row GEMM tiles)
  m->items_host.clear(); m->launches.clear(); m->launch_pairs.clear();
  // row distance of the lo half of every tensor the weight-gradient kernel reads (split-precision mode)
  const int lo_rows[LW_MAPS] = {m->act_slots * cap, tc->total_feat_rows, cap, cap, cap, cap, cap, cap, cap};
  auto flush = [&](std::vector<WgUnit>& units) {
    // launches made of 256 x 256 kernel blocks only run on CTA pairs (wgrad2_kernel: half the L2 -> SM bytes per FLOP)
    bool pairs = m->wg_pairs;
    for (const WgUnit& u : units) pairs = pairs && u.w.n == 256 && u.w.flush_mode == 0;
    m->launch_pairs.push_back(pairs);
    // ... whose bias gradients come from the dgrad GEMM that produces dZ (DenseParams::colsum); items count 128-sample stages
    if (pairs) for (WgUnit& u : units) u.w.bias_mode = 0;
    std::vector<WgItem> items;
    wgrad_plan(units, pairs ? T / 2 : T, pairs ? tc->num_sms / 2 : tc->num_sms, &items);
    const size_t first = m->items_host.size();
    // split-precision mode: (A_hi + A_lo)^T (dZ_hi + dZ_lo) as four items; the bias column sums ride on the A_hi items only
    for (const WgItem& base : items)
      for (int pa = 0; pa < m->parts; ++pa)
        for (int pb = 0; pb < m->parts; ++pb) {
          WgItem w = base;
          w.a_row0 += pa * lo_rows[w.a_map];
          w.b_row0 += pb * lo_rows[w.b_map];
          if (pa > 0) w.bias_mode = 0;
          m->items_host.push_back(w);
        }
    m->launches.push_back({(int)first, (int)(m->items_host.size() - first)});
    units.clear();
  };
  auto unit = [&](int a_map, int a_row0, int a_col0, int b_map, int b_col0, int n, const DenseView& v, int in_base,
                  int feat_mode, int bias_mode, int group) {
    WgItem w{};
    w.a_map = a_map; w.a_row0 = a_row0; w.a_col0 = a_col0; w.b_map = b_map; w.b_row0 = 0; w.b_col0 = b_col0; w.n = n;
    w.out = v.out; w.koff = v.kernel_off; w.in_base = in_base; w.feat_mode = feat_mode; w.bias_mode = bias_mode;
    w.boff = v.bias_off;
    return WgUnit{w, (512.f + 2.f * n) / 1024.f, group};
  };
  std::vector<WgUnit> units;
  // launch 0: heads that only need the start op: rgb head, density head, view layer
  {
    WgItem w{};   // rgb head: A = view activation (128 columns; the box beyond them is zero-filled), B = head gradients
    w.a_map = LW_VACT; w.b_map = LW_DH; w.n = kHeadCols; w.out = 3; w.koff = mv.dense[D + 3].kernel_off;
    w.flush_mode = 2; w.bias_mode = 3; w.boff = mv.dense[D + 3].bias_off;
    units.push_back({w, 0.6f, 100});
    for (int ab = 0; ab < W / 256; ++ab) {   // density head: A = last trunk activation, B = head gradients (column 3)
      WgItem q{};
      q.a_map = LW_ACT; q.a_row0 = (D - 1) * cap; q.a_col0 = ab * 256; q.b_map = LW_DH; q.n = kHeadCols; q.out = 1;
      q.koff = mv.dense[D].kernel_off + ab * 256; q.flush_mode = 1; q.bias_mode = ab == 0 ? 2 : 0; q.boff = mv.dense[D].bias_off;
      units.push_back({q, 0.6f, 101 + ab});
    }
    units.push_back(unit(LW_BOTT, 0, 0, LW_DZV, 0, 128, mv.dense[D + 2], 0, 0, 1, 99));
    flush(units);
  }
  // launch 1: bottleneck (after dZ_bott)
  for (int ab = 0; ab < W / 256; ++ab)
    units.push_back(unit(LW_ACT, (D - 1) * cap, ab * 256, LW_DZB, 0, 256, mv.dense[D + 1], ab * 256, 0, ab == 0 ? 1 : 0, 0));
  flush(units);
  // launches 2..: trunk layers D-1 .. 0; dZ_l lives in dz[(D - 1 - l) & 1].  All (A block, dZ block) units of a layer form
  // one group: the 16 (24 with the skip features) items of a sample range are queued next to each other, so every A
  // and dZ block is streamed from HBM once and read by its other three consumers out of L2
  for (int l = D - 1; l >= 0; --l) {
    const int bmap = ((D - 1 - l) & 1) ? LW_DZ1 : LW_DZ0;
    const bool cat_in = l > 0 && ((l - 1) % m->skip == 0 && (l - 1) > 0);
    for (int bb = 0; bb < W / 256; ++bb) {
      if (l > 0)
        for (int ab = 0; ab < W / 256; ++ab)
          units.push_back(unit(LW_ACT, (l - 1) * cap, ab * 256, bmap, bb * 256, 256, mv.dense[l], ab * 256, 0,
                               ab == 0 ? 1 : 0, 0));
      if (l == 0 || cat_in)
        for (int sb = 0; sb < kFeatPad / 256; ++sb)
          units.push_back(unit(LW_FEAT, tc->feat_row0[level], sb * 256, bmap, bb * 256, 256, mv.dense[l], l == 0 ? 0 : W, 1,
                               (l == 0 && sb == 0) ? 1 : 0, 0));
    }
    flush(units);
  }
}

}  // namespace

int layered_backward(hugs_handle* h, LayeredMlp* m, int level, int n_rays, float* grad, cudaStream_t st) {
  TcState* tc = h->tc;
  const int S = h->samples(level), M = n_rays * S, W = m->W, D = m->D, cap = m->cap;
  const int rows_pad = ((M + 255) / 256) * 256;
  HUGS_REQUIRE(m->train_ready, "layered path: training buffers missing");
  int rc;
  if (m->built_for != M) {
    build_wgrad(h, m, level, M);
    HUGS_REQUIRE(m->items_host.size() <= 4096 * 4, "layered wgrad: too many work items");
    HUGS_CUDA(cudaMemcpyAsync(m->items_dev, m->items_host.data(), sizeof(WgItem) * m->items_host.size(),
                              cudaMemcpyHostToDevice, st));
    HUGS_CUDA(cudaStreamSynchronize(st));
    m->built_for = M;
  }
  int launch = 0;
  auto wgrad = [&]() {
    ProfScope ps(h, HUGS_K_WGRAD_NERF, st);
    const bool pairs = m->launch_pairs[launch];
    const auto& L = m->launches[launch++];
    if (pairs) {
      const CUtensorMap maps128[LW_MAPS] = {m->map_act, tc->map_feat, m->map_dz[0], m->map_dz[1], m->map_bott, m->map_vact,
                                            m->map_dzb, m->map_dzv, m->wg_maps[LW_DH]};
      return wgrad2_launch_raw(tc->num_sms, h->perm_nb, h->d.max_deg_point - h->d.min_deg_point, h->feat_dim, maps128, LW_MAPS,
                               m->items_dev + L.first, L.second, grad, st);
    }
    return wgrad_launch(h, m->wg_maps, LW_MAPS, m->items_dev + L.first, L.second, grad, st);
  };
  const int lo_act = m->act_slots * cap;
  const float* w_rgb = m->tab + m->w_rgb_off;
  {
    ProfScope ps(h, HUGS_K_CHAIN_BWD_NERF, st);
    if ((rc = launch_bwd_start(h->d_raw[level], m->vact, 128, w_rgb, M, rows_pad, m->dz_view, 128, m->drgb,
                               m->split ? m->dz_view + (size_t)cap * 128 : nullptr,
                               m->split ? m->drgb + (size_t)cap * kHeadCols : nullptr, st)))
      return rc;
  }
  if ((rc = wgrad())) return rc;
  {
    ProfScope ps(h, HUGS_K_REDUCTIONS, st);
    if ((rc = wgrad_view_extras(h, m->dz_view, m->split ? m->dz_view + (size_t)cap * 128 : nullptr, 128, n_rays, S, grad, st)))
      return rc;
  }
  auto bwd_common = [&](DenseParams* p) {
    fill_common(p, m, M);
    p->b_map = m->map_wn128;
    p->exact_rank1 = m->split ? 1 : 0;
  };
  {  // dZ_bott = dZ_view . W_view[bottleneck rows]^T
    ProfScope ps(h, HUGS_K_CHAIN_BWD_NERF, st);
    DenseParams p;
    bwd_common(&p);
    SegSrc src[1] = {SegSrc{&m->map_dzv, 2, 0, 0, 0, cap}};
    set_segs(&p, m, src, 1, m->rows_b);
    p.b_row0 = m->row_b[D + 2]; p.n_tiles = 1;
    p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_BWD_LINEAR;
    p.out_map = m->map_dzb; p.out_lo_row_off = cap;
    if (m->wg_pairs) p.colsum = grad + h->nerf.dense[D + 1].bias_off;
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
  }
  if ((rc = wgrad())) return rc;
  int cur = 0;
  {  // dZ_{D-1} = (dZ_bott . W_bott^T + d_density (x) w_density) * [act_{D-1} > 0]
    ProfScope ps(h, HUGS_K_CHAIN_BWD_NERF, st);
    DenseParams p;
    bwd_common(&p);
    SegSrc src[1] = {SegSrc{&m->map_dzb, 4, 0, 0, 0, cap}};
    set_segs(&p, m, src, 1, m->rows_b);
    p.b_row0 = m->row_b[D + 1]; p.n_tiles = W / 256;
    for (int j = 0; j < p.n_tiles; ++j) { p.tile_n0[j] = j * 256; p.tile_bn[j] = 256; p.tile_epi[j] = DE_BWD_RELU; }
    p.mask_act = m->act; p.mask_ld = W; p.mask_row0 = (D - 1) * cap;
    p.gate_in = m->gate; p.gate_ld = W / 32; p.gate_row0 = (D - 1) * cap;
    p.rank1_row = h->d_raw[level]; p.rank1_stride = 4; p.rank1_col = m->tab + m->w_dens_off;
    p.out_map = m->map_dz[cur]; p.out_lo_row_off = cap;
    if (m->wg_pairs) p.colsum = grad + h->nerf.dense[D - 1].bias_off;
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
  }
  for (int l = D - 1; l >= 0; --l) {
    if ((rc = wgrad())) return rc;
    if (l == 0) break;
    ProfScope ps(h, HUGS_K_CHAIN_BWD_NERF, st);
    DenseParams p;   // dZ_{l-1} = (dZ_l . W_l[x rows]^T) * [act_{l-1} > 0]
    bwd_common(&p);
    SegSrc src[1] = {SegSrc{&m->map_dz[cur], W / 64, 0, 0, 0, cap}};
    set_segs(&p, m, src, 1, m->rows_b);
    p.b_row0 = m->row_b[l]; p.n_tiles = W / 256;
    for (int j = 0; j < p.n_tiles; ++j) { p.tile_n0[j] = j * 256; p.tile_bn[j] = 256; p.tile_epi[j] = DE_BWD_RELU; }
    p.mask_act = m->act; p.mask_ld = W; p.mask_row0 = (l - 1) * cap;
    p.gate_in = m->gate; p.gate_ld = W / 32; p.gate_row0 = (l - 1) * cap;
    p.out_map = m->map_dz[cur ^ 1]; p.out_lo_row_off = cap;
    if (m->wg_pairs) p.colsum = grad + h->nerf.dense[l - 1].bias_off;
    if ((rc = dense_tc_launch(p, tc->num_sms, st))) return rc;
    cur ^= 1;
  }
  (void)lo_act;
  return HUGS_OK;
}

}  // namespace hugs
