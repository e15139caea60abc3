// Internal declarations shared by the tensor-core translation units (mlp_tc.cu, mlp_pp.cu, wgrad_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

#include "tc.h"

namespace hugs {

// ------------------------------------------------------------------------------------------
// constants / schedule description
// ------------------------------------------------------------------------------------------
constexpr int kTileM = 128;             // samples per tile (= TMEM lanes)
constexpr int kPanelBytes = 16384;      // 128 rows x 64 bf16, SWIZZLE_128B
constexpr int kNumPanels = 8;           // two activation buffers of 4 K-panels (256 columns) each
constexpr int kStages = 5;              // TMA ring depth
constexpr int kW = 256;                 // trunk / bottleneck width of this path
constexpr int kFeatPad = 512;           // IPE features padded to 8 K-panels
constexpr int kKP = kW + kFeatPad;      // K extent of the packed forward weights
constexpr int kEpiGroups = 4;           // epilogue groups of 128 threads; group q owns output columns [64q, 64q+64)
constexpr int kHeadCols = 64;           // columns of the head-gradient tensor (d_r, d_g, d_b, d_density, 0...)

enum Epi : int {
  EPI_RELU = 0,      // bias + ReLU -> bf16 panels            (trunk)
  EPI_LINEAR = 1,    // bias         -> bf16 panels            (bottleneck)
  EPI_VIEW = 3,      // per-ray view bias + ReLU -> panels 0,1 (N = 128) + rgb head
  // backward chain
  EPI_BWD_START = 5, // no MMA: d_raw -> dZ_view panels (+ rgb head dgrad on CUDA cores)
  EPI_BWD_LINEAR = 6,// dA -> bf16 panels                      (through the linear bottleneck)
  EPI_BWD_RELU = 7,  // dA * [A > 0] -> bf16 panels
  EPI_BWD_RELU_D = 8,// (dA + d_density * w_density) * [A > 0] -> bf16 panels
  EPI_BWD_START_PROP = 9,  // no MMA: d_raw_density * w_density * [A > 0] -> panels
};

constexpr int EPI_NONE = -1;
constexpr int kMaxSegs = 32;

// One segment of the chain kernel's per-tile program (mlp_pp.cu): <= 4 K-panels of one layer.
struct PpSeg {
  int kps;          // K panels (64 columns each); 0: epilogue-only start op (backward)
  int a_feat;       // 1: A = IPE feature columns [feat_col0, feat_col0 + 64*kps), TMA-loaded into the tile's panels
  int feat_col0;
  int n_halves;     // 2: N = 256, 1: N = 128
  int w_row, w_col0;
  int accumulate;   // 1: keep accumulating into the tile's TMEM accumulator
  int epi;          // EPI_NONE: more segments of the same layer follow
  int bias_off;
  int save_row;     // slot index in the schedule, row base in the launch parameters; -1: do not save
  int mask_row;     // same convention; saved forward activation gating a backward epilogue
  int no_signal;    // the produced panels feed no MMA
  int head;         // forward: this epilogue also evaluates the density head on CUDA cores
  int last_epi;     // last epilogue of the tile: release the panels for the next pair's features
  int feat_next;    // the next MMA segment of the program (cyclically) refills this tile's panels with features:
                    // signal `consumed` per K panel so that the refill can start panel by panel
  int bias_idx;     // >= 0 on the first segment of a layer whose bias is applied by a K = 16 MMA
                    // (ones x [bias_hi, bias_lo]) that also initialises the accumulator; -1: none
  int no_wait;      // split-precision programs: the A panels were already waited for by the preceding segment of the layer
};

// biases as bf16 (hi, lo) column pairs of no-swizzle K-major B tiles, 4 layers per 8-column chunk
constexpr int kBiasChunks = 3;            // <= 12 biased layers
constexpr int kBiasChunkElems = 1024;     // 128 rows (this CTA's half of the outputs) x 8 K columns
constexpr int kBiasTailFloats = 1280;     // fp32 table kept in shared memory (head weights + head biases)

struct alignas(64) PpParams {
  CUtensorMap map_w, map_feat, map_save;
  CUtensorMap map_w_half;            // N = 128 layers: box of 64 weight rows per CTA
  PpSeg segs[kMaxSegs];
  int n_segs, any_feat;
  int n_tiles, n_units, n_samples, S;   // unit = 4 tiles of a CTA pair (2 in the split-precision mode)
  int feat_row0;
  const float* bias; int bias_floats;
  const float* viewbias;
  float* raw_out; int raw_c;
  const float* d_raw;
  __nv_bfloat16* drgb_out;
  int w_dens_off, w_rgb_off, dens_bias_off, rgb_bias_off;
  uint2* gate;                       // ReLU gate bits [save row base * 4 + group * cap + sample] (64 columns each)
  int cap;                           // rows per save slot of this level
  int epi_seg[kMaxSegs]; int n_epi;  // segments that carry an epilogue, in program order
  const __nv_bfloat16* bias_img;     // [2 ranks][kBiasChunks][kBiasChunkElems]
  int bias_tail0;                    // first float of `bias` staged in shared memory
  // split-precision mode: row distance between the hi and the lo half of the feature / saved / head-gradient tensors
  int lo_feat_rows, lo_save_rows, lo_drgb_rows;
};

struct TcMlp {
  bool present = false, has_rgb = false;
  int depth = 0;
  __nv_bfloat16* wt = nullptr; int rows_f = 0;   // forward pack  [rows_f (x2: hi, lo), kKP]   (K-major rows = outputs)
  __nv_bfloat16* wn = nullptr; int rows_b = 0;   // backward pack [rows_b (x2), kW]             (rows = inputs, cols = outputs)
  float* bias = nullptr; int bias_floats = 0;
  __nv_bfloat16* bias_img = nullptr;             // see PpParams::bias_img
  int w_dens_off = 0, w_rgb_off = 0;
  std::vector<PpSeg> pp_fwd, pp_bwd;
  CUtensorMap map_wt128, map_wn128, map_wt64;
  // packing tables
  struct PackLayer { int row0, rows_pad, out, in, x_in, feat_in; long long koff, boff; int bias_off, brow0, b_out_pad; };
  std::vector<PackLayer> pack;
};

struct WgState;
struct LayeredMlp;

struct TcState {
  TcMlp nerf, prop;
  LayeredMlp* nerf_layered = nullptr;   // NerfMLP.net_width != 256: layer-at-a-time path (layered.cu) instead of the chain
  WgState* wg = nullptr;
  bool split = false;                // HUGS_PRECISION_TC_SPLIT: every bf16 operand tensor holds a hi and a lo half
  __nv_bfloat16* feat = nullptr;     // per level region [cap_l, 512]
  __nv_bfloat16* act = nullptr;      // saved forward activations
  __nv_bfloat16* dz = nullptr;       // saved backward dZ
  uint2* gate = nullptr;             // ReLU gate bit masks of the saved activations, 32 B per save row
  __nv_bfloat16* drgb = nullptr;     // [max cap, kHeadCols]
  int drgb_rows = 0;
  float* viewbias = nullptr;
  CUtensorMap map_feat, map_act, map_dz;
  std::vector<int> cap, feat_row0, save_row0;   // per level
  int total_feat_rows = 0, total_save_rows = 0;
  int num_sms = 148;
  bool train_ready = false;          // training buffers + weight-gradient state allocated (tc_ensure_training)
};


// row-major bf16 [rows, cols] tensor map, box = [box_rows, 64 cols], 128B swizzle
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows);

// reference feature column of my column f' = (b*ndeg + k)*2 + s   ->   s*(nb*ndeg) + k*nb + b
// (nb <= 0: the engine keeps the reference's order, point positional encoding)
__host__ __device__ inline int ref_feature_col(int fp, int nb, int ndeg) {
  if (nb <= 0) return fp;
  int s = fp & 1, bk = fp >> 1, b = bk / ndeg, k = bk % ndeg;
  return s * (nb * ndeg) + k * nb + b;
}

// mlp_pp.cu
int pp_build(hugs_handle* h, const MlpViews& mv, TcMlp* m);
int pp_init(hugs_handle* h);
int pp_launch(hugs_handle* h, int level, int n_rays, int direction, cudaStream_t st);

// wgrad_tc.cu
struct WgItem {
  int a_map;        // index into WgParams::maps of the A tensor (saved activations, features, ...)
  int a_row0;       // first row of this level in the A tensor
  int a_col0;       // first A column of the 256-wide superblock
  int b_row0;       // first row of the dZ slot
  int b_col0;       // first dZ column of this item (layer-at-a-time path: 256-column blocks of a wider dZ)
  int n;            // dZ columns (256 | 128 | 64)
  int st0, st1;     // [st0, st1) 64-sample stages
  int out;          // kernel columns
  long long koff;   // kernel offset in the flat gradient
  int in_base;      // kernel row of A column a_col0 (feature mode: first feature row)
  int feat_mode;    // 1: A columns are features in engine order -> permute rows on flush
  int b_map;        // index into WgParams::maps of the B tensor (dZ, head gradients, ...)
  int flush_mode;   // 0: kernel tile; 1: density head (column 3 -> [in,1]); 2: rgb head (columns 0..2 -> [in,3])
  int bias_mode;    // 0: none; 1: all n columns -> boff + c; 2: column 3 -> boff; 3: columns 0..2 -> boff + c
  long long boff;   // bias offset in the flat gradient
  int in_rows;      // > 0: only A columns [a_col0, a_col0 + in_rows) are kernel rows (narrow inputs inside a 256-wide superblock)
  int head_rows;    // flush_mode 2: kernel rows of the rgb head (0: 128, the view layer's width)
};

struct WgUnit { WgItem w; float cost; int group; };
constexpr int kWgMaxMaps = 12;
void wgrad_plan(const std::vector<WgUnit>& units, int T, int num_sms, std::vector<WgItem>* items);
int wgrad_launch(hugs_handle* h, const CUtensorMap* maps, int n_maps, const WgItem* dev_items, int n_items, float* grad,
                 cudaStream_t st);
// the same launch without a model handle (hash-grid fields): feature-permutation parameters passed explicitly
int wgrad_launch_raw(int num_sms, int perm_nb, int ndeg, int feat_dim, const CUtensorMap* maps, int n_maps,
                     const WgItem* dev_items, int n_items, float* grad, cudaStream_t st);
// CTA-pair variant for launches whose items are all 256 x 256 kernel blocks (n == 256, flush_mode == 0, no bias sums;
// tensor maps with 128-row boxes, st0 / st1 in 128-sample stages): layered path
int wgrad2_launch_raw(int num_sms, int perm_nb, int ndeg, int feat_dim, const CUtensorMap* maps, int n_maps,
                      const WgItem* dev_items, int n_items, float* grad, cudaStream_t st);
int wgrad_kernel_init();
// view-layer extras shared by the chain and the layered path: dW rows of the direction / GLO inputs, GLO embedding rows
int wgrad_view_extras(hugs_handle* h, const __nv_bfloat16* dzv, const __nv_bfloat16* dzv_lo, int dz_ld, int n_rays, int S,
                      float* grad, cudaStream_t st);
int wgrad_create(hugs_handle* h);
void wgrad_destroy(hugs_handle* h);
int wgrad_run(hugs_handle* h, int level, int n_rays, float* grad, cudaStream_t st);

// mlp_simt.cu: exact-arithmetic IPE features split into bf16 hi / lo halves, engine column order (split-precision mode)
struct EncSplitArgs {
  const float* origins; const float* directions; const float* radii; const float* tdist; const float* basis;
  int n_samples, n_rows_pad, S, nb, min_deg, ndeg, ray_shape, contract;
  __nv_bfloat16* feat_hi; __nv_bfloat16* feat_lo;    // [rows, 512]
};
int launch_encode_split(const EncSplitArgs& a, cudaStream_t stream);

}  // namespace hugs
