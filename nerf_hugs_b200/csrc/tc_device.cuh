// Device helpers shared by the tensor-core kernels (mlp_tc.cu, mlp_pp.cu).
#pragma once
#include "ptx.cuh"

namespace hugs {

// 16-byte chunk index of a 128-byte row under the 128B swizzle (chunk ^= row % 8)
__device__ __forceinline__ uint32_t swz_chunk(uint32_t row, uint32_t chunk) { return chunk ^ (row & 7u); }

// One 32-column half of a chunk: TMEM -> registers (fp32).
__device__ __forceinline__ void load_acc32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  ptx::tmem_ld32(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 values of one row -> bf16 -> four 16-byte chunks (chunk0 .. chunk0+3) of a swizzled panel row.
template <bool kRelu>
__device__ __forceinline__ void store_half32(uint8_t* panel, int row, int chunk0, const float (&v)[32]) {
  uint4* prow = reinterpret_cast<uint4*>(panel + row * 128);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 q;
    if (kRelu) {
      q.x = ptx::pack_bf16x2_relu(v[c * 8 + 0], v[c * 8 + 1]); q.y = ptx::pack_bf16x2_relu(v[c * 8 + 2], v[c * 8 + 3]);
      q.z = ptx::pack_bf16x2_relu(v[c * 8 + 4], v[c * 8 + 5]); q.w = ptx::pack_bf16x2_relu(v[c * 8 + 6], v[c * 8 + 7]);
    } else {
      q.x = ptx::pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); q.y = ptx::pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
      q.z = ptx::pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); q.w = ptx::pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    }
    prow[swz_chunk(row, chunk0 + c)] = q;
  }
}

// ReLU gate from four 16-byte words of the saved (post-ReLU, hence >= 0) bf16 activation row.
__device__ __forceinline__ void apply_mask32(const uint4 (&mk)[4], float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t w[4] = {mk[c].x, mk[c].y, mk[c].z, mk[c].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((w[j] & 0xFFFFu) == 0u) v[c * 8 + j * 2] = 0.f;
      if ((w[j] >> 16) == 0u) v[c * 8 + j * 2 + 1] = 0.f;
    }
  }
}

// ReLU gate bits of 16 packed bf16x2 words (non-negative halves): bit j = low half of word j is non-zero,
// bit 16 + j = high half.  (h + 0x7FFF sets bit 15 of a half-word iff h >= 1; no carry between the halves.)
__device__ __forceinline__ uint32_t gate_bits16(const uint32_t* pk) {
  uint32_t g = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) g |= ((pk[j] + 0x7FFF7FFFu) >> (15 - j)) & (0x00010001u << j);
  return g;
}
// zero the halves of 16 packed words whose gate bit is clear
__device__ __forceinline__ void apply_gate16(uint32_t g, uint32_t* pk) {
#pragma unroll
  for (int j = 0; j < 16; ++j) pk[j] &= ((g >> j) & 0x00010001u) * 0xFFFFu;
}

}  // namespace hugs
