// Forward chain kernel of the nerfacto field (field_chain.cu).
#pragma once
#include "dense_tc.h"

namespace hugs {

constexpr int kFcMaxLinks = 6;

struct FieldChainLink {
  int kp;                       // K panels (64 columns) of the link's A operand (link 0: the TMA-loaded feature panel)
  int b_row0;                   // first weight row of the link inside the forward pack (K-major rows = output units)
  int n_tiles;                  // 1 or 2 N tiles
  int tile_n0[2], tile_bn[2], tile_epi[2];   // DenseEpi: DE_RELU / DE_LINEAR / DE_VIEW / DE_HEAD_F32
  int bias_off;                 // offset of the link's bias row in the fp32 table (indexed by output column)
  int store;                    // TMA-store the output panels to out_map[link] (training: the saved activation)
  int gate_row0;                // >= 0: write ReLU gate bits of the output at this row base of gate_out
  int raw_chan0, raw_nchan;     // DE_HEAD_F32 tiles: channels of raw_out
  int gate_in_row0;             // DE_BWD_RELU tiles: row base of the layer's ReLU gate bits in gate_in
  int rank1;                    // DE_BWD_RELU tiles: add d_density[row] * rank1_col[n] before the gate (density head dgrad)
};

struct alignas(64) FieldChainParams {
  CUtensorMap a_map;            // features, bf16 [rows, 64], box 128 x 64
  CUtensorMap b_map, b_map_64, b_map_8;   // forward weight pack, boxes of 128 / 64 / 8 rows x 64 K
  CUtensorMap out_map[kFcMaxLinks];       // per link: bf16 [rows, cols], box 128 x 64 (used when link.store)
  FieldChainLink link[kFcMaxLinks];
  int n_links, m_tiles, m_rows, S;
  const float* bias; int n_bias;   // fp32 table (biases, head weights): copied to shared memory by every CTA (<= 1792 floats)
  const float* viewbias; int view_ld;
  float* raw_out; int raw_c;
  uint32_t* gate_out; int gate_ld;
  // backward program (start_mode = 1): no feature load; the chain starts from dZ of the last colour layer, computed by the slot's
  // epilogue group from d_raw = (d_density, d_r, d_g, d_b) per sample: dZ[c] = (d_rgb . w_rgb[c]) * [h1[c] > 0]
  int start_mode;
  const float* d_raw;           // [M, 4]
  const uint8_t* inside;        // [M] density * selector mask (no density gradient out of range)
  int w_rgb_off;                // offset of the rgb-head weights [256][3] (bf16-rounded fp32) in the table
  int rank1_off;                // offset of the density-head weights [256] in the table
  const uint32_t* gate_in;      // gate bits written by the forward chain
  int start_gate_row0;          // row base of the last colour layer's gate bits
  CUtensorMap start_map;        // dZ of the last colour layer [rows, 256] (saved for the weight gradients)
  __nv_bfloat16* dh_out;        // [M, 64] head-gradient rows (d_r, d_g, d_b, d_density, 0 ...) for the head weight gradients
};

int field_chain_init();
int field_chain_launch(const FieldChainParams& p, int num_sms, cudaStream_t st);

}  // namespace hugs
