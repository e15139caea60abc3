// Tensor-core (tcgen05) MLP path: host interface used by api.cu / optim.cu.
#pragma once
#include <cuda_bf16.h>

#include "handle.h"

namespace hugs {

int tc_create(hugs_handle* h);
void tc_destroy(hugs_handle* h);
// fp32 flat params -> packed bf16 operand tensors (+ fp32 bias packs)
int tc_pack_params(hugs_handle* h, const float* params, cudaStream_t st);
// allocate the training-only buffers of the handle on first use (saved activations, dZ, gates, weight-gradient state)
int tc_ensure_training(hugs_handle* h);
// encode + fused MLP chain for level l; fills h->raw[l]; saves activations when `training`
int tc_mlp_forward(hugs_handle* h, int level, const hugs_rays* rays, int n_rays, bool training, cudaStream_t st);
// dgrad chain + wgrad for level l from h->d_raw[l]; accumulates into grad (flat fp32, flax layout)
int tc_mlp_backward(hugs_handle* h, int level, const hugs_rays* rays, int n_rays, float* grad, cudaStream_t st);

// test hook: the throughput-mode bf16 feature encoder on its own ([n*S, 512] bf16, engine column order)
int tc_debug_encode(hugs_handle* h, const hugs_rays* rays, const float* tdist, int n_rays, int S, int contract,
                    __nv_bfloat16* out, cudaStream_t st);

int launch_finalize_stats(hugs_handle* h, const hugs_loss_cfg& loss, int n, const float* denom, const float* ray_stats,
                          float* stats_out, cudaStream_t st);

}  // namespace hugs
