// Layer-at-a-time tensor-core path (dense_tc.cu + layered.cu): MLP widths other than 256.
#pragma once
#include "tc_internal.h"

namespace hugs {

enum DenseEpi : int {
  DE_RELU = 0,        // bf16(relu(acc + bias[n]))
  DE_LINEAR = 1,      // bf16(acc + bias[n])
  DE_VIEW = 2,        // bf16(relu(acc + viewbias[row / S][n]))
  DE_HEAD_F32 = 3,    // raw_out[row * raw_c + raw_chan0 + c] = acc[c] + bias[n0 + c]   (N tile of 16)
  DE_BWD_RELU = 4,    // bf16((acc + rank1_row[row] * rank1_col[n]) * [saved activation > 0])
  DE_BWD_LINEAR = 5,  // bf16(acc)
};

constexpr int kMaxNTiles = 8;
constexpr int kDMaxSegs = 6;   // A segments of one GEMM (K concatenation): [x | features] x {hi.hi, lo.hi, hi.lo} in the split mode

struct alignas(64) DenseParams {
  CUtensorMap a_map[kDMaxSegs];// A segments: bf16 [rows, cols], box 128 rows x 64 cols
  CUtensorMap b_map;           // weights, K-major rows = GEMM output columns: box 128 rows x 64 (BN = 256)
  CUtensorMap b_map_64;        // box 64 rows (BN = 128)
  CUtensorMap b_map_8;         // box 8 rows  (BN = 16)
  CUtensorMap out_map;         // bf16 output [rows, cols], box 128 x 64
  int n_seg;                   // 0: the two-segment form (a_kp[0], a_kp[1], weight columns contiguous from b_col0)
  int a_kp[kDMaxSegs];         // K panels (64 columns) of each A segment
  int a_row0[kDMaxSegs], a_col0[kDMaxSegs];
  int w_col0[kDMaxSegs];       // first weight K column of the segment (n_seg > 0)
  int w_row_off[kDMaxSegs];    // added to the weight row of the segment (the lo half of a split pack)
  int b_row0, b_col0;          // first weight row / K column of this GEMM inside the weight tensor
  // split-precision mode (HUGS_PRECISION_TC_SPLIT): bf16 outputs are written as hi + residual lo, the lo half
  // `out_lo_row_off` rows further down in the output tensor; the ring is one stage shorter to make room for its panels
  int split, out_lo_row_off, exact_rank1;
  int m_tiles, n_tiles, m_rows;
  int tile_n0[kMaxNTiles], tile_bn[kMaxNTiles], tile_epi[kMaxNTiles];
  int S;                       // samples per ray (DE_VIEW)
  const float* bias;           // fp32, indexed by GEMM output column
  const float* viewbias; int view_ld;
  int out_row0, out_col0;
  float* raw_out; int raw_c, raw_chan0, raw_nchan;
  const __nv_bfloat16* mask_act; int mask_ld, mask_row0;
  const float* rank1_row; int rank1_stride; const float* rank1_col;
  // backward GEMMs: colsum[n] += sum over rows of the bf16 output (hi + lo in the split mode) = the bias gradient of the layer
  // whose dZ this GEMM produces (column sums of dZ), taken from the staged panels before they are stored; nullptr: off
  float* colsum;
  // ReLU gate bit masks, [rows][gate_ld] 32-bit words, bit k of word j = column 32 j + k is open (output > 0):
  // gate_out != nullptr: a DE_RELU / DE_VIEW launch with 256-column tiles also writes them (training forward);
  // gate_in  != nullptr: a DE_BWD_RELU launch reads them instead of the saved activation (16 instead of 256 bytes per thread and
  //                      tile half: the kernel is bound by the shared-memory / L1 data pipe)
  uint32_t* gate_out; const uint32_t* gate_in; int gate_ld, gate_row0;
};

int dense_tc_init();
int dense_tc_launch(const DenseParams& p, int num_sms, cudaStream_t st);
// lo pointers != nullptr: split-precision mode (unrounded head gradients, hi + lo outputs)
int launch_bwd_start(const float* d_raw, const __nv_bfloat16* view_act, int view_ld, const float* w_rgb, int n_samples,
                     int n_rows_pad, __nv_bfloat16* dz_view, int dz_ld, __nv_bfloat16* drgb, __nv_bfloat16* dz_view_lo,
                     __nv_bfloat16* drgb_lo, cudaStream_t st);

// layered.cu: per-MLP state of the layer-at-a-time path
struct LayeredMlp;
int layered_create(hugs_handle* h, const MlpViews& mv, int level_lo, LayeredMlp** out);
void layered_destroy(LayeredMlp* m);
int layered_pack(hugs_handle* h, LayeredMlp* m, const float* params, cudaStream_t st);
int layered_ensure_training(hugs_handle* h, LayeredMlp* m);
int layered_forward(hugs_handle* h, LayeredMlp* m, int level, int n_rays, bool training, cudaStream_t st);
int layered_backward(hugs_handle* h, LayeredMlp* m, int level, int n_rays, float* grad, cudaStream_t st);

}  // namespace hugs
