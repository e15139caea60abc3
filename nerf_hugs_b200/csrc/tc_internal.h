// Internal declarations shared by the tensor-core translation units (mlp_tc.cu, wgrad_tc.cu).
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
constexpr int kThreads = 64 + kEpiGroups * 128;   // producer warp + MMA warp + epilogue warps
constexpr int kHeadCols = 64;           // columns of the head-gradient tensor (d_r, d_g, d_b, d_density, 0...)
constexpr int kBiasTab = 3136;          // fp32 bias / head-weight table staged in shared memory
constexpr int kMaxSteps = 64;           // MMA issue steps per tile (precomputed list in shared memory)
constexpr int kSmemBytes = 1024 + (kNumPanels + kStages) * kPanelBytes + kBiasTab * 4 + kMaxSteps * 16 + 512;

enum Epi : int {
  EPI_RELU = 0,      // bias + ReLU -> bf16 panels            (trunk)
  EPI_LINEAR = 1,    // bias         -> bf16 panels            (bottleneck)
  EPI_DENSITY = 2,   // column 0 + bias -> raw density         (group 0)
  EPI_VIEW = 3,      // per-ray view bias + ReLU -> panels 0,1 (N = 128)
  EPI_RGB = 4,       // columns 0..2 + bias -> raw rgb         (group 0)
  // backward chain
  EPI_BWD_START = 5, // no MMA: d_raw -> dZ_view panels (+ rgb head dgrad on CUDA cores)
  EPI_BWD_LINEAR = 6,// dA -> bf16 panels                      (through the linear bottleneck)
  EPI_BWD_RELU = 7,  // dA * [A > 0] -> bf16 panels
  EPI_BWD_RELU_D = 8,// (dA + d_density * w_density) * [A > 0] -> bf16 panels
  EPI_BWD_START_PROP = 9,  // no MMA: d_raw_density * w_density * [A > 0] -> panels
};

struct TcLayer {
  int a_res, a_str, a_buf, wait_panels;
  int a_feat;        // 1: the resident A panels are the tile's IPE features, TMA-loaded into panels 0..7 (layer 0)
  int n_halves, n_mma, acc_col, acc_bar;
  int w_row, w_map;
  int epi, dst_buf, bias_off;
  int save_row;      // base row in the save tensor (activations fwd / dZ bwd), -1 = do not save
  int mask_row;      // bwd: base row of the saved forward activation whose sign gates this epilogue
  int no_signal;     // 1: the produced panels feed no later MMA (last backward op): do not arrive on panel_ready
};

constexpr int kMaxLayers = 14;

struct alignas(64) TcParams {
  CUtensorMap map_w128, map_w16, map_feat, map_save;
  TcLayer layers[kMaxLayers];
  int n_layers;
  int n_tiles, n_samples, S;
  int feat_row0;
  const float* bias;                 // packed fp32 biases (+ head weights, see TcMlp)
  const float* viewbias;             // [n_rays, 128]
  float* raw_out; int raw_c;         // [n_samples, raw_c]
  const float* d_raw;                // bwd: [n_samples, raw_c]
  const __nv_bfloat16* act;          // bwd: saved forward activations [rows, 256]
  __nv_bfloat16* drgb_out;           // bwd: [rows, kHeadCols] bf16 head gradients for the head wgrad GEMMs
  int w_dens_off, w_rgb_off;         // float offsets of head weights inside `bias`
  int bias_floats;                   // size of the bias table
  long long* dbg;                    // optional [gridDim.x][16] cycle counters (development instrumentation)
};

constexpr int EPI_NONE = -1;
constexpr int kMaxSegs = 20;

// One segment of the ping-pong kernel's per-tile program (mlp_pp.cu): <= 4 K-panels of one layer.
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
  int bias_idx;     // CTA-pair kernel: >= 0 on the first segment of a layer whose bias is applied by a K = 16 MMA
                    // (ones x [bias_hi, bias_lo]) that also initialises the accumulator; -1: none
};

// CTA-pair kernel: biases as bf16 (hi, lo) column pairs of no-swizzle K-major B tiles, 4 layers per 8-column chunk
constexpr int kBiasChunks = 3;            // <= 12 biased layers
constexpr int kBiasChunkElems = 1024;     // 128 rows (this CTA's half of the outputs) x 8 K columns
constexpr int kBiasTailFloats = 1280;     // fp32 table kept in shared memory by the CTA-pair kernel (heads + head biases)

struct alignas(64) PpParams {
  CUtensorMap map_w, map_feat, map_save;
  CUtensorMap map_w_half;            // CTA-pair kernel, N = 128 layers: box of 64 weight rows per CTA
  PpSeg segs[kMaxSegs];
  int n_segs, any_feat;
  int n_tiles, n_units, n_samples, S;   // unit = 2 tiles (one CTA) or 4 tiles (CTA pair)
  int feat_row0;
  const float* bias; int bias_floats;
  const float* viewbias;
  float* raw_out; int raw_c;
  const float* d_raw;
  const __nv_bfloat16* act;
  __nv_bfloat16* drgb_out;
  int w_dens_off, w_rgb_off, dens_bias_off, rgb_bias_off;
  uint2* gate;                       // CTA-pair kernel: ReLU gate bits [save row base * 4 + group * cap + sample] (64 columns each)
  int cap;                           // rows per save slot of this level
  int epi_seg[kMaxSegs]; int n_epi;  // segments that carry an epilogue, in program order
  const __nv_bfloat16* bias_img;     // CTA-pair kernel: [2 ranks][kBiasChunks][kBiasChunkElems]
  int bias_tail0;                    // CTA-pair kernel: first float of `bias` staged in shared memory
  long long* dbg;                    // optional [gridDim.x][16] cycle counters (development instrumentation)
};

struct TcMlp {
  bool present = false, has_rgb = false;
  int depth = 0;
  __nv_bfloat16* wt = nullptr; int rows_f = 0;   // forward pack  [rows_f, kKP]   (K-major rows = outputs)
  __nv_bfloat16* wn = nullptr; int rows_b = 0;   // backward pack [rows_b, kW]    (rows = inputs, cols = outputs)
  float* bias = nullptr; int bias_floats = 0;
  __nv_bfloat16* bias_img = nullptr;             // see PpParams::bias_img
  int w_dens_off = 0, w_rgb_off = 0, view_bias_off = 0, view_w_row = 0;
  std::vector<TcLayer> fwd, bwd;
  std::vector<PpSeg> pp_fwd, pp_bwd;
  CUtensorMap map_wt128, map_wt16, map_wn128, map_wt64;
  // packing tables
  struct PackLayer { int row0, rows_pad, out, in, x_in, feat_in; long long koff, boff; int bias_off, brow0, b_out_pad; };
  std::vector<PackLayer> pack;
};

struct WgState;

struct TcState {
  TcMlp nerf, prop;
  WgState* wg = nullptr;
  __nv_bfloat16* feat = nullptr;     // per level region [cap_l, 512]
  __nv_bfloat16* act = nullptr;      // saved forward activations
  __nv_bfloat16* dz = nullptr;       // saved backward dZ
  uint2* gate = nullptr;             // ReLU gate bit masks of the saved activations (CTA-pair kernel), 32 B per save row
  __nv_bfloat16* drgb = nullptr;     // [max cap, kHeadCols]
  int drgb_rows = 0;
  float* viewbias = nullptr;
  CUtensorMap map_feat, map_act, map_dz;
  std::vector<int> cap, feat_row0, save_row0;   // per level
  int total_feat_rows = 0, total_save_rows = 0;
  int num_sms = 148;
  bool train_ready = false;          // training buffers + weight-gradient state allocated (tc_ensure_training)
  bool use_pp = true;                // two-tile ping-pong chain kernel (HUGS_CHAIN=single selects the older one)
  bool use_cg2 = true;               // ... on CTA pairs with tcgen05 cta_group::2 (HUGS_CHAIN=pp selects one CTA per unit)
  void* pack_tables = nullptr;
};


// row-major bf16 [rows, cols] tensor map, box = [box_rows, 64 cols], 128B swizzle
int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows);

// reference feature column of my column f' = (b*ndeg + k)*2 + s   ->   s*(nb*ndeg) + k*nb + b
__host__ __device__ inline int ref_feature_col(int fp, int nb, int ndeg) {
  int s = fp & 1, bk = fp >> 1, b = bk / ndeg, k = bk % ndeg;
  return s * (nb * ndeg) + k * nb + b;
}

// mlp_pp.cu
int pp_build(hugs_handle* h, const MlpViews& mv, TcMlp* m);
int pp_init(hugs_handle* h);
int pp_launch(hugs_handle* h, int level, int n_rays, int direction, cudaStream_t st);

// wgrad_tc.cu
int wgrad_create(hugs_handle* h);
void wgrad_destroy(hugs_handle* h);
int wgrad_run(hugs_handle* h, int level, int n_rays, float* grad, cudaStream_t st);

}  // namespace hugs
