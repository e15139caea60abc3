// Launch interfaces of the MLP paths (fp32 CUDA-core parity path, bf16 tcgen05 path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hugs_b200.h"

namespace hugs {

struct IpeArgs {
  const float* origins; const float* directions; const float* radii;
  const float* tdist;            // [n, S+1]
  const float* basis;            // device [3][num_basis]
  int n_rays, S, num_basis, min_deg, max_deg, ray_shape, contract;
  float* features;               // [n*S, 2*num_basis*(max_deg-min_deg)] reference column order
};
int launch_ipe_features(const IpeArgs& a, cudaStream_t stream);

int launch_view_inputs(const float* viewdirs, const int32_t* embed_idx, const float* glo_table, int n_rays,
                       int deg_view, int glo, int zero_glo, int num_embeddings, float* out, cudaStream_t stream);

struct DenseSeg { const float* x; int k; int ld; int row_div; };
struct DenseArgs {
  DenseSeg seg[3]; int nseg;
  const float* W;                // [sum k, N] (flax kernel layout)
  const float* bias;             // [N] or nullptr
  int M, N, relu;
  float* y; int ldy;
};
int launch_dense_simt(const DenseArgs& a, cudaStream_t stream);

}  // namespace hugs
