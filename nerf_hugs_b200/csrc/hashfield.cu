// Hash-grid fields of the nerfacto twin: multiresolution hash encoding + small MLPs.
//
// Reference call sites (paths under /root/reference/nerfacto/models): nerfacto.py:643-875 NerfactoField (hash grid -> 256
// -> [raw density | 64 geometry features]; [SH4(viewdir) | geometry | appearance] -> 256 -> 256 -> rgb), nerfacto.py:878-1008
// HashMLPDensityField (hash grid -> 64 -> raw density, the proposal networks).  The encodings themselves live in a
// third-party dependency that is NOT under /root/reference: tiny-cuda-nn (unpinned git HEAD, requirements_torch.txt:8).
// Its published algorithm is restated here (and in oracle/hashgrid.py): per level l, scale_l = exp2(l * log2(per_level_scale))
// * base_res - 1, resolution ceil(scale_l) + 1, pos = x * scale_l + 0.5, trilinear weights of frac(pos), corner index
// = x + y * res + z * res^2 while the stride fits the level's table, else the coherent prime hash x ^ y * 2654435761 ^ z *
// 805459861, modulo the level's entry count; levels are stored back to back (entry counts rounded up to 8, capped at
// 2^log2_hashmap_size), features innermost; spherical harmonics of degree 4 on 2 * v - 1.  Parity is therefore pinned by
// the reference's call sites run on this restatement (tests/golden/make_golden_nerfacto.py), not by tcnn outputs.
//
// B200 design: the gathers are L2 / HBM-sector bound, not tensor bound: one thread per sample walks all levels (8
// independent 8-byte gathers per level in flight), consecutive threads = consecutive samples of a ray, so neighbouring
// lanes hit the same or adjacent cells of the coarse levels.  The 64-wide density fields (proposal networks: 98 % of the
// samples, 2 % of the FLOPs) are ONE fused CUDA-core kernel per direction - encode, hidden layer and head in registers,
// weights in shared memory, nothing but the raw density leaves the SM.  The 256-wide main field runs its five Dense layers
// on the tcgen05 GEMM kernel of dense_tc.cu (bf16 x bf16 -> fp32 in TMEM) and its weight gradients on wgrad_kernel.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "tc_device.cuh"
#include "dense_tc.h"
#include "field_chain.h"

namespace hugs {

struct HashGridCfg {
  int L, F, in_dim;                 // levels, features per level (2), L * F
  float scale[24];
  uint32_t res[24], off[25];        // entries (not floats) before each level
  float bound; int contract;
};

namespace {

constexpr int kPropHidden = 64;
constexpr int kMaxIn = 48;          // L * F of a fused density field
constexpr int kH = 256;             // hidden width of the main field (both MLPs)
constexpr int kG = 64;              // geometry features
constexpr int kSH = 16;

// ---------------------------------------------------------------------------------------------------- encoding
__device__ __forceinline__ uint32_t grid_index(uint32_t T, uint32_t res, uint32_t x, uint32_t y, uint32_t z) {
  uint32_t stride = 1, index = 0;
  index += x * stride; stride *= res;
  if (stride <= T) { index += y * stride; stride *= res; }
  if (stride <= T) { index += z * stride; stride *= res; }
  if (T < stride) index = x ^ (y * 2654435761u) ^ (z * 805459861u);
  return index % T;
}

struct FieldRays {
  const float* origins; const float* directions; const float* tdist;
  int n_rays, S;
};

// normalised position of sample s and its in-range flag (nerfacto.py:816-827): the position is zeroed when out of range
__device__ __forceinline__ bool sample_unit_pos(const FieldRays& r, const HashGridCfg& g, int s, float x[3]) {
  const int ray = s / r.S, i = s % r.S;
  const float t0 = r.tdist[(size_t)ray * (r.S + 1) + i], t1 = r.tdist[(size_t)ray * (r.S + 1) + i + 1];
  const float tm = __fdiv_rn(__fadd_rn(t1, t0), 2.f);
  for (int c = 0; c < 3; ++c) x[c] = __fadd_rn(r.origins[ray * 3 + c], __fmul_rn(r.directions[ray * 3 + c], tm));
  if (g.contract) {
    const float m = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(x[0], x[0]), __fmul_rn(x[1], x[1])), __fmul_rn(x[2], x[2])), kF32Eps);
    if (!(m <= 1.f)) {
      const float sc = __fdiv_rn(__fadd_rn(__fmul_rn(2.f, sqrtf(m)), -1.f), m);
      for (int c = 0; c < 3; ++c) x[c] = __fmul_rn(sc, x[c]);
    }
    for (int c = 0; c < 3; ++c) x[c] = __fdiv_rn(__fadd_rn(x[c], 2.0f), 4.0f);
  } else {
    for (int c = 0; c < 3; ++c) x[c] = __fdiv_rn(__fadd_rn(x[c], g.bound), __fmul_rn(2.f, g.bound));
  }
  const bool in = x[0] >= 0.f && x[0] <= 1.f && x[1] >= 0.f && x[1] <= 1.f && x[2] >= 0.f && x[2] <= 1.f;
  if (!in) { x[0] = 0.f; x[1] = 0.f; x[2] = 0.f; }
  return in;
}

struct Corner { uint32_t idx[8]; float w[8]; };

__device__ __forceinline__ void level_corners(const HashGridCfg& g, int l, const float x[3], Corner& c) {
  const float sc = g.scale[l];
  const uint32_t res = g.res[l], T = g.off[l + 1] - g.off[l];
  float f[3]; uint32_t p[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float pos = fmaf(sc, x[d], 0.5f);
    const float fl = floorf(pos);
    p[d] = (uint32_t)(int)fl;
    f[d] = pos - fl;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
    c.w[k] = (dx ? f[0] : 1.f - f[0]) * (dy ? f[1] : 1.f - f[1]) * (dz ? f[2] : 1.f - f[2]);
    c.idx[k] = g.off[l] + grid_index(T, res, p[0] + dx, p[1] + dy, p[2] + dz);
  }
}

// features of one sample, level-major: out[l * 2 + f]   (F == 2)
template <int kMaxL>
__device__ __forceinline__ void encode_sample(const HashGridCfg& g, const float2* __restrict__ grid, const float x[3], float* out) {
  constexpr int kUnroll = kMaxL <= 16 ? kMaxL : 2;    // few levels: fully unrolled, `out` stays in registers
#pragma unroll kUnroll
  for (int l = 0; l < kMaxL; ++l) {
    if (l < g.L) {
      Corner c;
      level_corners(g, l, x, c);
      float2 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(grid + c.idx[k]);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { a = fmaf(c.w[k], v[k].x, a); b = fmaf(c.w[k], v[k].y, b); }
      out[2 * l] = a; out[2 * l + 1] = b;
    }
  }
}

// d_features of one sample -> grid gradient (trilinear weights transposed).  MUST be called by all 32 lanes of a warp
// (`live` = this lane has a sample): consecutive lanes are consecutive samples of a ray, so on the coarse levels whole
// runs of lanes hit the same table entry; those runs are summed with a segmented shuffle reduction and only the head of
// a run issues the atomic (the level-0 table has 4096 entries for millions of samples: without this the atomics of one
// warp serialise on a handful of addresses).
// Measured on config 4 (profiles/r02_nerfacto.md): no aggregation 15.9 ms backward, levels up to resolution 64 only 15.5 ms,
// up to 512 8.7 ms, every level 8.7 ms - runs of equal entries are common well beyond the coarsest levels.
constexpr int kAggregateRes = 512;
template <int kMaxL>
__device__ __forceinline__ void scatter_sample(const HashGridCfg& g, float2* grid_grad, const float x[3], const float* df,
                                               bool live) {
  const int lane = threadIdx.x & 31;
  constexpr int kUnroll = kMaxL <= 16 ? kMaxL : 1;
#pragma unroll kUnroll
  for (int l = 0; l < kMaxL; ++l) {
    if (l >= g.L) continue;
    const float a = live ? df[2 * l] : 0.f, b = live ? df[2 * l + 1] : 0.f;
    const bool aggregate = g.res[l] <= (uint32_t)kAggregateRes;
    if (!aggregate && a == 0.f && b == 0.f) continue;
    Corner c;
    level_corners(g, l, x, c);
    if (!aggregate) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(grid_grad + c.idx[k], make_float2(c.w[k] * a, c.w[k] * b));
      continue;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t key = live ? c.idx[k] : (0xFFFFFF00u + (uint32_t)lane);      // idle lanes never join a run
      float va = c.w[k] * a, vb = c.w[k] * b;
      const uint32_t prev = __shfl_up_sync(kFull, key, 1);
      const bool head = lane == 0 || prev != key;
      // run id = number of run heads at or below this lane: equal ids <=> same contiguous run (equal keys further apart
      // belong to different runs and issue their own atomics)
      const uint32_t rid = __popc(__ballot_sync(kFull, head) & (0xFFFFFFFFu >> (31 - lane)));
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const uint32_t ro = __shfl_down_sync(kFull, rid, dlt);
        const float ao = __shfl_down_sync(kFull, va, dlt), bo = __shfl_down_sync(kFull, vb, dlt);
        if (lane + dlt < 32 && ro == rid) { va += ao; vb += bo; }
      }
      if (head && live && (va != 0.f || vb != 0.f)) atomicAdd(grid_grad + key, make_float2(va, vb));
    }
  }
}

// ---------------------------------------------------------------------------------------------------- main-field encode
struct EncodeArgs {
  FieldRays r; HashGridCfg g;
  const float2* grid;
  __nv_bfloat16* feats;   // [rows_pad, 64] (columns >= L * F zero)
  float* feats_f32;       // optional [n_samples, L * F] (operator-level test hook)
  int rows_pad;
  __nv_bfloat16* feats_lo;// split-precision mode: residual half (same shape), or nullptr
};

__global__ void __launch_bounds__(128) hash_encode_kernel(EncodeArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.rows_pad) return;
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = 0.f;
  if (s < a.r.n_rays * a.r.S) {
    float x[3];
    sample_unit_pos(a.r, a.g, s, x);
    encode_sample<16>(a.g, a.grid, x, f);
    if (a.feats_f32)
      for (int i = 0; i < a.g.in_dim; ++i) a.feats_f32[(size_t)s * a.g.in_dim + i] = f[i];
  }
  if (a.feats) {
    uint4* dst = reinterpret_cast<uint4*>(a.feats + (size_t)s * 64);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_uint4(ptx::pack_bf16x2(f[q * 8], f[q * 8 + 1]), ptx::pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                          ptx::pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), ptx::pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
#pragma unroll
    for (int q = 4; q < 8; ++q) dst[q] = make_uint4(0u, 0u, 0u, 0u);
    if (a.feats_lo) {
      uint4* dl = reinterpret_cast<uint4*>(a.feats_lo + (size_t)s * 64);
      auto lo2 = [&](int i) {
        const uint32_t hw = ptx::pack_bf16x2(f[i], f[i + 1]);
        return ptx::pack_bf16x2(f[i] - __uint_as_float(hw << 16), f[i + 1] - __uint_as_float(hw & 0xFFFF0000u));
      };
#pragma unroll
      for (int q = 0; q < 4; ++q) dl[q] = make_uint4(lo2(q * 8), lo2(q * 8 + 2), lo2(q * 8 + 4), lo2(q * 8 + 6));
#pragma unroll
      for (int q = 4; q < 8; ++q) dl[q] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

struct ScatterArgs {
  FieldRays r; HashGridCfg g;
  const __nv_bfloat16* d_feats; int ld;   // [rows, ld] bf16
  float2* grid_grad;
  const __nv_bfloat16* d_feats_lo;        // split-precision mode: residual half, or nullptr
};

__global__ void __launch_bounds__(128) hash_scatter_kernel(ScatterArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = s < a.r.n_rays * a.r.S;
  float x[3] = {0.f, 0.f, 0.f};
  float df[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) df[i] = 0.f;
  if (live) {
    sample_unit_pos(a.r, a.g, s, x);
    const uint4* src = reinterpret_cast<const uint4*>(a.d_feats + (size_t)s * a.ld);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = __ldg(src + q);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { df[q * 8 + 2 * j] = __uint_as_float(w[j] << 16); df[q * 8 + 2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u); }
    }
    if (a.d_feats_lo) {
      const uint4* sl = reinterpret_cast<const uint4*>(a.d_feats_lo + (size_t)s * a.ld);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = __ldg(sl + q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { df[q * 8 + 2 * j] += __uint_as_float(w[j] << 16); df[q * 8 + 2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u); }
      }
    }
  }
  scatter_sample<16>(a.g, a.grid_grad, x, df, live);
}

// ---------------------------------------------------------------------------------------------------- fused density field
// flat MLP layout of a density field: W1 [in, 64] | b1 [64] | w2 [64] | b2 [1]
struct PropArgs {
  FieldRays r; HashGridCfg g;
  const float2* grid;
  const float* mlp;
  float* raw;               // [n_samples] pre-activation density (-inf where the sample is out of range)
  // backward
  const float* d_raw;       // [n_samples]
  float2* grid_grad;
  float* mlp_grad;
};

// two independent fp32 FMAs in one instruction (sm_100 FFMA2): the fused proposal-network kernels are bound by instruction
// issue, and every lane is an exact fp32 fma, so results do not change
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

constexpr int kPropAcc = (kMaxIn * kPropHidden + 127) / 128;
__host__ __device__ inline int prop_ldf(int in) { return in | 1; }     // odd row stride: conflict-free staging
inline size_t prop_smem_bytes(int in, bool backward) {
  size_t fl = (size_t)in * kPropHidden + 2 * kPropHidden + 4;
  if (backward) fl += 128 * (size_t)prop_ldf(in) + 128 * 65;
  return fl * sizeof(float);
}

// kIn > 0: the input width (2 x levels: 10 and 14 in the shipped configs) is a compile-time constant, so the feature /
// gradient vectors live in registers and every loop unrolls; kIn == 0 is the generic version (runtime width, local arrays).
template <bool kBackward, int kIn>
__global__ void __launch_bounds__(128) prop_field_kernel(PropArgs a) {
  extern __shared__ float sm[];
  constexpr int kCap = kIn > 0 ? kIn : kMaxIn;
  const int in = kIn > 0 ? kIn : a.g.in_dim, ldf = prop_ldf(in);
  float* W1 = sm;                      // [in][64]
  float* b1 = W1 + in * kPropHidden;   // [64]
  float* w2 = b1 + kPropHidden;        // [64]
  float* b2 = w2 + kPropHidden;        // [1] (+3 pad)
  float* stF = b2 + 4;                 // backward staging: [128][ldf] features
  float* stZ = stF + 128 * ldf;        //                   [128][65] dZ of the hidden layer
  const int n_w = in * kPropHidden + 2 * kPropHidden + 1;
  for (int i = threadIdx.x; i < n_w; i += blockDim.x) sm[i] = a.mlp[i];
  __syncthreads();
  const int n = a.r.n_rays * a.r.S;
  const int n_out = in * kPropHidden;
  // backward: per-thread partial sums of dW1 (outputs o = threadIdx.x + 128 q), db1 (threads 0..63: one column each),
  // dw2 (lane l of every warp: columns l and l + 32 of the warp's samples), db2
  constexpr int kAcc = (kCap * kPropHidden + 127) / 128;
  float acc_w[kAcc];
  float acc_col = 0.f, acc_b2 = 0.f, acc_w2a = 0.f, acc_w2b = 0.f;
  const int lane_id = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < kAcc; ++q) acc_w[q] = 0.f;
  for (int base = blockIdx.x * 128; base < n; base += gridDim.x * 128) {
    const int s = base + threadIdx.x;
    const bool live = s < n;
    float f[kCap];
    float x[3] = {0.f, 0.f, 0.f};
    bool inside = false;
#pragma unroll
    for (int i = 0; i < kCap; ++i) f[i] = 0.f;
    if (live) {
      inside = sample_unit_pos(a.r, a.g, s, x);
      encode_sample<kCap / 2>(a.g, a.grid, x, f);
    }
    float2 h2[kPropHidden / 2];         // hidden pre-activations, columns (2 jj, 2 jj + 1)
#pragma unroll
    for (int jj = 0; jj < kPropHidden / 2; ++jj) h2[jj] = *reinterpret_cast<const float2*>(b1 + 2 * jj);
#pragma unroll
    for (int i = 0; i < kCap; ++i) {
      if (kIn == 0 && i >= in) break;
      const float2 fi = make_float2(f[i], f[i]);
#pragma unroll
      for (int jj = 0; jj < kPropHidden / 2; ++jj)
        h2[jj] = ffma2(fi, *reinterpret_cast<const float2*>(W1 + i * kPropHidden + 2 * jj), h2[jj]);
    }
#define HUGS_H(j) (((j) & 1) ? h2[(j) >> 1].y : h2[(j) >> 1].x)
    if (!kBackward) {
      float o = b2[0];
#pragma unroll
      for (int j = 0; j < kPropHidden; ++j) o = fmaf(fmaxf(HUGS_H(j), 0.f), w2[j], o);
      if (live) a.raw[s] = inside ? o : -INFINITY;      // density * selector (nerfacto.py:985-989)
      continue;
    }
    // ---- backward: dZ = d_raw * w2 * [h > 0]; the selector zeroes the density gradient of out-of-range samples ----
    const float dr = (live && inside) ? a.d_raw[s] : 0.f;
    float2 df2[kCap];                   // feature gradients: partial sums over the even / odd hidden columns
#pragma unroll
    for (int i = 0; i < kCap; ++i) df2[i] = make_float2(0.f, 0.f);
    __syncthreads();                    // the previous chunk's staging has been consumed
#pragma unroll
    for (int i = 0; i < kCap; ++i) { if (kIn == 0 && i >= in) break; stF[threadIdx.x * ldf + i] = f[i]; }
#pragma unroll
    for (int jj = 0; jj < kPropHidden / 2; ++jj) {
      float2 dz2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * jj + e;
        const float hj = HUGS_H(j);
        const float dz = hj > 0.f ? dr * w2[j] : 0.f;
        stZ[threadIdx.x * 65 + j] = dz;
        if (e == 0) dz2.x = dz; else dz2.y = dz;
        // dw2[j] += sum over the warp's samples of d_raw * relu(h_j): butterfly sum, kept by lane j % 32
        const float hs = warp_sum(dr * fmaxf(hj, 0.f));
        if (lane_id == (j & 31)) { if (j < 32) acc_w2a += hs; else acc_w2b += hs; }
      }
#pragma unroll
      for (int i = 0; i < kCap; ++i) {
        if (kIn == 0 && i >= in) break;
        df2[i] = ffma2(*reinterpret_cast<const float2*>(W1 + i * kPropHidden + 2 * jj), dz2, df2[i]);
      }
    }
    float df[kCap];
#pragma unroll
    for (int i = 0; i < kCap; ++i) df[i] = df2[i].x + df2[i].y;
#undef HUGS_H
    __syncthreads();
    scatter_sample<kCap / 2>(a.g, a.grid_grad, x, df, dr != 0.f);
    // dW1[i][j] += sum_t f[t][i] * dZ[t][j]: output o = i * 64 + j (a warp: one i, 32 consecutive j)
    // o = threadIdx.x + 128 q  =>  j = threadIdx.x % 64 for every q and i = threadIdx.x / 64 + 2 q: one dZ value per sample
    // serves all of this thread's outputs (the feature values are warp-wide broadcasts)
    {
      static_assert(kPropHidden == 64, "output mapping of the weight-gradient loop");
      const int j = threadIdx.x & 63, i0 = threadIdx.x >> 6;
      float sacc[kAcc];
#pragma unroll
      for (int q = 0; q < kAcc; ++q) sacc[q] = 0.f;
#pragma unroll 4
      for (int t = 0; t < 128; ++t) {
        const float z = stZ[t * 65 + j];
#pragma unroll
        for (int q = 0; q < kAcc; ++q)
          if (i0 + 2 * q < in) sacc[q] = fmaf(stF[t * ldf + i0 + 2 * q], z, sacc[q]);
      }
#pragma unroll
      for (int q = 0; q < kAcc; ++q) acc_w[q] += sacc[q];
    }
    if (threadIdx.x < kPropHidden) {   // db1[j] = sum_t dZ[t][j]
      float sacc = 0.f;
#pragma unroll 8
      for (int t = 0; t < 128; ++t) sacc += stZ[t * 65 + threadIdx.x];
      acc_col += sacc;
    }
    const float pb = warp_sum(dr);
    if ((threadIdx.x & 31) == 0) acc_b2 += pb;
  }
  if (kBackward) {
#pragma unroll
    for (int q = 0; q < kAcc; ++q) {
      const int o = threadIdx.x + 128 * q;
      if (o < n_out) atomicAdd(a.mlp_grad + o, acc_w[q]);
    }
    if (threadIdx.x < kPropHidden) atomicAdd(a.mlp_grad + n_out + threadIdx.x, acc_col);
    atomicAdd(a.mlp_grad + n_out + kPropHidden + lane_id, acc_w2a);
    atomicAdd(a.mlp_grad + n_out + kPropHidden + 32 + lane_id, acc_w2b);
    if (lane_id == 0) atomicAdd(a.mlp_grad + n_out + 2 * kPropHidden, acc_b2);
  }
}

template <bool kBackward>
int launch_prop_field(const PropArgs& a, int blocks, cudaStream_t st) {
  const size_t smem = prop_smem_bytes(a.g.in_dim, kBackward);
  if (a.g.in_dim == 10) prop_field_kernel<kBackward, 10><<<blocks, 128, smem, st>>>(a);
  else if (a.g.in_dim == 14) prop_field_kernel<kBackward, 14><<<blocks, 128, smem, st>>>(a);
  else prop_field_kernel<kBackward, 0><<<blocks, 128, smem, st>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

// ---------------------------------------------------------------------------------------------------- main-field helpers
// per-ray inputs of the colour MLP: [SH4(2 * ((v + 1) / 2) - 1) (16) | appearance embedding]   (nerfacto.py:849-855)
__global__ void sh_inputs_kernel(const float* viewdirs, const int32_t* embed_idx, const float* emb, int n_rays, int app,
                                 int num_emb, int zero_app, float* out) {
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= n_rays) return;
  float* o = out + (size_t)ray * (kSH + app);
  // tcnn's SphericalHarmonics maps its [0, 1] input back with x * 2 - 1
  const float x = ((viewdirs[ray * 3] + 1.0f) / 2.0f) * 2.f - 1.f, y = ((viewdirs[ray * 3 + 1] + 1.0f) / 2.0f) * 2.f - 1.f,
              z = ((viewdirs[ray * 3 + 2] + 1.0f) / 2.0f) * 2.f - 1.f;
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
  const int row = app > 0 && !zero_app ? min(max(embed_idx[ray], 0), num_emb - 1) : 0;
  for (int j = 0; j < app; ++j) o[kSH + j] = zero_app ? 0.f : emb[(size_t)row * app + j];
}

// raybias[ray][c] = sum_j bf16(inp[ray][j]) * bf16(W[(row0 + j) * out + c]) + b[c]
__global__ void ray_bias_kernel(const float* inp, int in_dim, const float* W, int row0, const float* bias, int out, int n_rays,
                                int exact, float* rb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * out) return;
  const int ray = idx / out, c = idx % out;
  float acc = 0.f;
  for (int j = 0; j < in_dim; ++j) {
    float x = inp[(size_t)ray * in_dim + j], w = W[(size_t)(row0 + j) * out + c];
    if (!exact) { x = __bfloat162float(__float2bfloat16(x)); w = __bfloat162float(__float2bfloat16(w)); }
    acc = fmaf(x, w, acc);
  }
  rb[idx] = acc + bias[c];
}

// dZ_head1 = (d_rgb . W_rgb^T) * [head activation > 0] (bf16) + the head-gradient rows (d_r, d_g, d_b, d_density).
// Block = 8 samples x 32 lanes, a lane owns 8 consecutive columns (one 16-byte load / store).
__global__ void __launch_bounds__(256) field_bwd_start_kernel(const float* d_raw, const __nv_bfloat16* hact, const float* w_rgb,
                                                              const uint8_t* inside, int n_samples, int n_rows_pad,
                                                              __nv_bfloat16* dz, __nv_bfloat16* dh, float* d_dens,
                                                              __nv_bfloat16* dz_lo, __nv_bfloat16* dh_lo,
                                                              const uint32_t* gate /* [rows][8] bit masks of hact, or nullptr */) {
  __shared__ float wsm[kH * 3];
  for (int i = threadIdx.x; i < kH * 3; i += blockDim.x) wsm[i] = w_rgb[i];
  __syncthreads();
  const int s = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (s >= n_rows_pad) return;
  float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s < n_samples) dr = __ldg(reinterpret_cast<const float4*>(d_raw) + s);
  if (s < n_samples && !inside[s]) dr.x = 0.f;            // density * selector: no density gradient out of range
  const bool split = dz_lo != nullptr;
  if (lane == 0) {
    const uint32_t h01 = ptx::pack_bf16x2(dr.y, dr.z), h23 = ptx::pack_bf16x2(dr.w, dr.x);
    reinterpret_cast<uint4*>(dh + (size_t)s * kHeadCols)[0] = make_uint4(h01, h23, 0u, 0u);
    if (split)
      reinterpret_cast<uint4*>(dh_lo + (size_t)s * kHeadCols)[0] =
          make_uint4(ptx::pack_bf16x2(dr.y - __uint_as_float(h01 << 16), dr.z - __uint_as_float(h01 & 0xFFFF0000u)),
                     ptx::pack_bf16x2(dr.w - __uint_as_float(h23 << 16), dr.x - __uint_as_float(h23 & 0xFFFF0000u)), 0u, 0u);
    d_dens[s] = dr.x;
  }
  float d0 = dr.y, d1 = dr.z, d2 = dr.w;
  if (!split) {
    d0 = __bfloat162float(__float2bfloat16(d0)); d1 = __bfloat162float(__float2bfloat16(d1)); d2 = __bfloat162float(__float2bfloat16(d2));
  }
  // ReLU gates of this lane's 8 columns: bit masks written by the forward GEMM (32 bytes per sample), else the saved
  // activation itself (512 bytes per sample); gb bit 2 j / 2 j + 1 = column lane * 8 + 2 j / + 1 is open
  uint32_t gb = 0u;
  if (s < n_samples) {
    if (gate) {
      gb = (__ldg(gate + (size_t)s * (kH / 32) + (lane >> 2)) >> ((lane & 3) * 8)) & 0xFFu;
    } else {
      const uint4 hv = __ldg(reinterpret_cast<const uint4*>(hact + (size_t)s * kH) + lane);
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) gb |= ((hw[j] & 0xFFFFu) ? 1u : 0u) << (2 * j) | ((hw[j] >> 16) ? 1u : 0u) << (2 * j + 1);
    }
  }
  uint32_t o[4], ol[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c0 = lane * 8 + 2 * j;
    float g0 = d0 * wsm[c0 * 3] + d1 * wsm[c0 * 3 + 1] + d2 * wsm[c0 * 3 + 2];
    float g1 = d0 * wsm[c0 * 3 + 3] + d1 * wsm[c0 * 3 + 4] + d2 * wsm[c0 * 3 + 5];
    // the saved activation is post-ReLU (>= 0): a non-zero bf16 pattern (of the hi half) means the gate is open
    if (!((gb >> (2 * j)) & 1u)) g0 = 0.f;
    if (!((gb >> (2 * j + 1)) & 1u)) g1 = 0.f;
    o[j] = ptx::pack_bf16x2(g0, g1);
    ol[j] = ptx::pack_bf16x2(g0 - __uint_as_float(o[j] << 16), g1 - __uint_as_float(o[j] & 0xFFFF0000u));
  }
  reinterpret_cast<uint4*>(dz + (size_t)s * kH)[lane] = make_uint4(o[0], o[1], o[2], o[3]);
  if (split) reinterpret_cast<uint4*>(dz_lo + (size_t)s * kH)[lane] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}

__global__ void inside_mask_kernel(FieldRays r, HashGridCfg g, uint8_t* inside, float* raw, int C) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= r.n_rays * r.S) return;
  float x[3];
  const bool in = sample_unit_pos(r, g, s, x);
  if (inside) inside[s] = in ? 1 : 0;
  if (raw && !in) raw[(size_t)s * C] = -INFINITY;
}

// dzsum[ray][c] = sum over the ray's samples of dZ[s][c]
__global__ void __launch_bounds__(kH) ray_colsum_kernel(const __nv_bfloat16* dz, const __nv_bfloat16* dz_lo, int S, int n_rays,
                                                        float* out) {
  const int ray = blockIdx.x, c = threadIdx.x;
  if (ray >= n_rays) return;
  float acc = 0.f;
  const __nv_bfloat16* src = dz + (size_t)ray * S * kH + c;
  for (int s = 0; s < S; ++s) acc += __bfloat162float(src[(size_t)s * kH]);
  if (dz_lo) {
    const __nv_bfloat16* sl = dz_lo + (size_t)ray * S * kH + c;
    for (int s = 0; s < S; ++s) acc += __bfloat162float(sl[(size_t)s * kH]);
  }
  out[(size_t)ray * kH + c] = acc;
}

// dW[row0 + j][c] += sum_ray bf16(inp[ray][j]) * dzsum[ray][c]; block.x = j, block.y = ray chunk, thread = c
__global__ void __launch_bounds__(kH) ray_input_wgrad_kernel(const float* inp, int in_dim, const float* dzsum, int n_rays,
                                                             int row0, int exact, float* dW) {
  const int j = blockIdx.x, c = threadIdx.x;
  const int chunk = (n_rays + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * chunk, r1 = min(r0 + chunk, n_rays);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) {
    float x = inp[(size_t)r * in_dim + j];
    if (!exact) x = __bfloat162float(__float2bfloat16(x));
    acc = fmaf(x, dzsum[(size_t)r * kH + c], acc);
  }
  if (r1 > r0) atomicAdd(dW + (size_t)(row0 + j) * kH + c, acc);
}

// d embedding[row][g] += sum_c bf16(W[(row0 + g) * 256 + c]) * dzsum[ray][c]: one warp per ray, lanes over c (coalesced)
__global__ void __launch_bounds__(256) app_embed_grad_kernel(const float* dzsum, const int32_t* embed_idx, const float* W,
                                                             int row0, int app, int n_rays, int num_emb, int exact, float* d_emb) {
  const int ray = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (ray >= n_rays) return;
  float z[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) z[q] = dzsum[(size_t)ray * kH + q * 32 + lane];
  const int row = embed_idx[ray];
  const bool ok = row >= 0 && row < num_emb;
  for (int g = 0; g < app; ++g) {
    const float* wrow = W + (size_t)(row0 + g) * kH;
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float w = __ldg(wrow + q * 32 + lane);
      if (!exact) w = __bfloat162float(__float2bfloat16(w));
      acc = fmaf(z[q], w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0 && ok) atomicAdd(d_emb + (size_t)row * app + g, acc);
  }
}

// bf16 operand packs of the main field.  Source kernels are flax-style [in, out_stride]; a forward block holds element
// (r = output unit, k = input unit), a dgrad block (r = input unit, k = output unit); rows / columns beyond the block's
// extent are zero.
struct PackBlock { int dst, row0, rows, fwd, in0, n_in, out0, n_out, out_stride; long long koff; };
struct FieldPackArgs {
  PackBlock b[12]; int n;
  const float* params; __nv_bfloat16* wt; __nv_bfloat16* wn; int ldk;
  int part; long long lo_wt, lo_wn;      // part 1: bf16(w - bf16(w)) into the lo halves (split-precision mode)
};
__global__ void field_pack_kernel(FieldPackArgs a) {
  const PackBlock B = a.b[blockIdx.y];
  __nv_bfloat16* dst = B.dst == 0 ? a.wt : a.wn;
  const long long tot = (long long)B.rows * a.ldk;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / a.ldk), k = (int)(e % a.ldk);
    float v = 0.f;
    if (B.fwd) { if (r < B.n_out && k < B.n_in) v = a.params[B.koff + (long long)(B.in0 + k) * B.out_stride + B.out0 + r]; }
    else       { if (r < B.n_in && k < B.n_out) v = a.params[B.koff + (long long)(B.in0 + r) * B.out_stride + B.out0 + k]; }
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const long long lo_off = a.part == 0 ? 0 : (B.dst == 0 ? a.lo_wt : a.lo_wn);
    dst[lo_off + (size_t)(B.row0 + r) * a.ldk + k] = a.part == 0 ? hi : __float2bfloat16(v - __bfloat162float(hi));
  }
}

// fp32 tables: biases per launch and the bf16-rounded head kernels of the CUDA-core backward start
__global__ void field_table_kernel(const float* params, float* tab, long long b_base0, long long b_dens, long long b_geo,
                                   long long b_head1, long long b_rgb, long long k_dens, long long k_rgb, int exact) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // [0,256) base0 bias | [256,400) heads: geo bias (64) + zeros, density bias at 256 + 128 | [400,656) head1 bias |
  // [656,672) rgb bias | [672,928) w_dens (bf16-rounded) | [928, 928 + 768) w_rgb [256][3] (bf16-rounded)
  if (i >= 928 + 768) return;
  float v = 0.f;
  if (i < 256) v = params[b_base0 + i];
  else if (i < 400) { const int j = i - 256; if (j < kG) v = params[b_geo + j]; else if (j == 128) v = params[b_dens]; }
  else if (i < 656) v = params[b_head1 + (i - 400)];
  else if (i < 672) { const int j = i - 656; if (j < 3) v = params[b_rgb + j]; }
  else if (i < 928) { v = params[k_dens + (i - 672)]; if (!exact) v = __bfloat162float(__float2bfloat16(v)); }
  else { v = params[k_rgb + (i - 928)]; if (!exact) v = __bfloat162float(__float2bfloat16(v)); }
  tab[i] = v;
}

}  // namespace
}  // namespace hugs

// ====================================================================================================== host side
using namespace hugs;

constexpr int kMaxWgItems = 4096;
enum { HF_FEAT = 0, HF_ACT0, HF_GEO, HF_H0, HF_H1, HF_DZH1, HF_DZH0, HF_DGEO, HF_DZA0, HF_DFEAT, HF_DH, HF_MAPS };

struct hugs_hashfield {
  hugs_hashfield_desc d{};
  HashGridCfg g{};
  int device = 0, num_sms = 148;
  bool is_prop = false;
  bool split = false; int parts = 1;   // HUGS_PRECISION_TC_SPLIT: every bf16 tensor has a lo half `cap` (weights: rows_f / rows_b) rows further down
  int64_t grid_floats = 0, mlp_floats = 0;
  std::vector<hugs_tensor_desc> tensors;
  std::vector<void*> allocs;
  int cap = 0;                       // rows (samples) the workspace holds, multiple of 256
  // ---- main field ----
  int64_t o_base0_k = 0, o_base0_b = 0, o_dens_k = 0, o_dens_b = 0, o_geo_k = 0, o_geo_b = 0, o_head0_k = 0, o_head0_b = 0,
          o_head1_k = 0, o_head1_b = 0, o_rgb_k = 0, o_rgb_b = 0, o_emb = -1;
  int head0_in = 0;
  __nv_bfloat16 *wt = nullptr, *wn = nullptr; float* tab = nullptr;
  int rf_base0 = 0, rf_heads = 0, rf_head0 = 0, rf_head1 = 0, rf_rgb = 0, rows_f = 0;
  int rb_base0 = 0, rb_geo = 0, rb_head0 = 0, rb_head1 = 0, rows_b = 0;
  CUtensorMap map_wt128, map_wt64, map_wt8, map_wn128, map_wn64;
  __nv_bfloat16* buf[HF_MAPS] = {};
  int buf_cols[HF_MAPS] = {64, 256, 128, 256, 256, 256, 256, 128, 256, 128, 64};
  CUtensorMap map128[HF_MAPS], map64[HF_MAPS];
  float *ray_in = nullptr, *ray_bias = nullptr, *dzsum = nullptr, *d_dens = nullptr;
  uint8_t* inside = nullptr;
  uint32_t* gate = nullptr;          // ReLU gate bit masks of ACT0 | H0 | H1 (H1: chain kernels only), [3][cap][8] 32-bit words (training)
  bool use_gate = getenv("HUGS_NF_GATE") ? atoi(getenv("HUGS_NF_GATE")) != 0 : true;   // development switch
  bool use_chain = getenv("HUGS_NF_CHAIN") ? atoi(getenv("HUGS_NF_CHAIN")) != 0 : true;  // forward chain kernel (bf16 mode); 0: five dense_tc launches
  bool train_ready = false;
  WgItem* items_dev = nullptr; std::vector<WgItem> items_host; std::vector<std::pair<int, int>> launches; int built_for = -1;
  int max_rays = 0;
};

namespace {

template <class T>
int hf_alloc(hugs_hashfield* h, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return HUGS_ERR_NOMEM;
  }
  h->allocs.push_back(q);
  *p = static_cast<T*>(q);
  if (zero) HUGS_CUDA(cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(T)));
  return HUGS_OK;
}

void hf_tensor(hugs_hashfield* h, const char* name, int rows, int cols, int64_t* off, int64_t* where) {
  hugs_tensor_desc t{};
  snprintf(t.name, sizeof(t.name), "%s", name);
  t.offset = *off; t.rows = rows; t.cols = cols; t.module = 0;
  h->tensors.push_back(t);
  if (where) *where = *off;
  *off += (int64_t)rows * cols;
}

int hf_check_rays(const hugs_hashfield* h, const hugs_rays* rays, int n_rays, int S) {
  HUGS_REQUIRE(rays && rays->origins && rays->directions, "hash field: rays need origins and directions");
  HUGS_REQUIRE(h->is_prop || rays->viewdirs, "hash field: viewdirs are required by the colour MLP");
  HUGS_REQUIRE(h->is_prop || h->d.appearance_dim == 0 || rays->embed_idx, "hash field: embed_idx is required with appearance embeddings");
  HUGS_REQUIRE(n_rays >= 0 && S >= 1 && (long long)n_rays * S <= h->d.max_samples, "hash field: %d x %d samples exceed max_samples %d",
               n_rays, S, h->d.max_samples);
  return HUGS_OK;
}

FieldRays field_rays(const hugs_rays* rays, const float* tdist, int n_rays, int S) {
  return FieldRays{rays->origins, rays->directions, tdist, n_rays, S};
}

void dense_common(DenseParams* p, const hugs_hashfield* h, int M) {
  memset(p, 0, sizeof(*p));
  p->b_map = h->map_wt128; p->b_map_64 = h->map_wt64; p->b_map_8 = h->map_wt8;
  p->m_rows = M; p->m_tiles = (M + 255) / 256;
  p->split = h->split ? 1 : 0; p->out_lo_row_off = h->cap; p->exact_rank1 = h->split ? 1 : 0;
}

// A = `kp` K panels of workspace tensor `buf`; split-precision mode: A_hi W_hi + A_lo W_hi + A_hi W_lo as one accumulation chain
// (`w_lo_rows` = row distance of the lo half of the weight pack: rows_f forward, rows_b backward)
void dense_a(DenseParams* p, const hugs_hashfield* h, int buf, int kp, int w_lo_rows) {
  const int combos = h->split ? 3 : 1;
  for (int c = 0; c < combos; ++c) {
    p->a_map[c] = h->map128[buf]; p->a_kp[c] = kp; p->a_row0[c] = c == 1 ? h->cap : 0; p->a_col0[c] = 0;
    p->w_col0[c] = 0; p->w_row_off[c] = c == 2 ? w_lo_rows : 0;
  }
  p->n_seg = combos;
}

int hf_ensure_training(hugs_hashfield* h) {
  if (h->train_ready || h->is_prop) return HUGS_OK;
  int rc;
  for (int i = HF_DZH1; i < HF_MAPS; ++i) {
    if ((rc = hf_alloc(h, &h->buf[i], (size_t)h->cap * h->parts * h->buf_cols[i]))) return rc;
    if ((rc = make_map(&h->map128[i], h->buf[i], h->cap * h->parts, h->buf_cols[i], 128)) ||
        (rc = make_map(&h->map64[i], h->buf[i], h->cap * h->parts, h->buf_cols[i], 64)))
      return rc;
  }
  if ((rc = hf_alloc(h, &h->gate, (size_t)3 * h->cap * (kH / 32)))) return rc;
  if ((rc = hf_alloc(h, &h->dzsum, (size_t)h->max_rays * kH)) || (rc = hf_alloc(h, &h->d_dens, (size_t)h->cap)) ||
      (rc = hf_alloc(h, &h->items_dev, kMaxWgItems)))
    return rc;
  if ((rc = wgrad_kernel_init())) return rc;
  HUGS_CUDA(cudaDeviceSynchronize());
  h->train_ready = true;
  return HUGS_OK;
}

void hf_build_wgrad(hugs_hashfield* h, int M) {
  const int T = ((M + 255) / 256) * 4;     // 64-sample stages
  h->items_host.clear(); h->launches.clear();
  std::vector<WgUnit> units;
  auto flush = [&]() {
    std::vector<WgItem> items;
    wgrad_plan(units, T, h->num_sms, &items);
    const size_t first = h->items_host.size();
    // split-precision mode: (A_hi + A_lo)^T (dZ_hi + dZ_lo) as four items; the bias column sums ride on the A_hi items only
    for (const WgItem& base : items)
      for (int pa = 0; pa < h->parts; ++pa)
        for (int pb = 0; pb < h->parts; ++pb) {
          WgItem w = base;
          w.a_row0 += pa * h->cap; w.b_row0 += pb * h->cap;
          if (pa > 0) w.bias_mode = 0;
          h->items_host.push_back(w);
        }
    h->launches.push_back({(int)first, (int)(h->items_host.size() - first)});
    units.clear();
  };
  int grp = 10;
  auto unit = [&](int a_map, int b_map, int n, int out, int64_t koff, int in_base, int in_rows, int bias_mode, int64_t boff,
                  int flush_mode, int head_rows) {
    WgItem w{};
    w.a_map = a_map; w.b_map = b_map; w.n = n; w.out = out; w.koff = koff; w.in_base = in_base; w.in_rows = in_rows;
    w.bias_mode = bias_mode; w.boff = boff; w.flush_mode = flush_mode; w.head_rows = head_rows;
    units.push_back({w, (512.f + 2.f * n) / 1024.f, grp++});
  };
  // launch 0 (after the start op): rgb head, head1
  unit(HF_H1, HF_DH, kHeadCols, 3, h->o_rgb_k, 0, 0, 3, h->o_rgb_b, 2, kH);
  unit(HF_H0, HF_DZH1, 256, kH, h->o_head1_k, 0, 0, 1, h->o_head1_b, 0, 0);
  flush();
  // launch 1 (after dZ_head0): geometry rows of head0
  unit(HF_GEO, HF_DZH0, 256, kH, h->o_head0_k, 0, kG, 1, h->o_head0_b, 0, 0);
  flush();
  // launch 2 (after d_geo): geometry head + density head of the base MLP
  unit(HF_ACT0, HF_DGEO, 64, kG, h->o_geo_k, 0, 0, 1, h->o_geo_b, 0, 0);
  unit(HF_ACT0, HF_DH, kHeadCols, 1, h->o_dens_k, 0, 0, 2, h->o_dens_b, 1, 0);
  flush();
  // launch 3 (after dZ_act0): first layer (hash features)
  unit(HF_FEAT, HF_DZA0, 256, kH, h->o_base0_k, 0, h->g.in_dim, 1, h->o_base0_b, 0, 0);
  flush();
}

}  // namespace

HUGS_API int hugs_hashfield_create(const hugs_hashfield_desc* desc, hugs_hashfield** out) {
  HUGS_REQUIRE(desc && out, "hugs_hashfield_create: null argument");
  *out = nullptr;
  const hugs_hashfield_desc& d = *desc;
  HUGS_REQUIRE(d.n_levels >= 1 && d.n_levels <= 24, "hash field: n_levels must be in [1,24], got %d", d.n_levels);
  HUGS_REQUIRE(d.features_per_level == 2, "hash field: features_per_level must be 2 (every shipped config), got %d", d.features_per_level);
  HUGS_REQUIRE(d.log2_hashmap_size >= 4 && d.log2_hashmap_size <= 24, "hash field: log2_hashmap_size out of range");
  HUGS_REQUIRE(d.base_res >= 1 && d.per_level_scale >= 1.f, "hash field: bad resolutions");
  HUGS_REQUIRE(d.max_samples >= 1, "hash field: max_samples must be >= 1");
  const bool is_prop = d.geo_feat_dim == 0;
  if (is_prop) {
    HUGS_REQUIRE(d.hidden_dim == kPropHidden, "density field: hidden_dim must be 64 (every shipped config), got %d", d.hidden_dim);
    HUGS_REQUIRE(d.n_levels * 2 <= kMaxIn, "density field: at most %d levels", kMaxIn / 2);
  } else {
    HUGS_REQUIRE(d.hidden_dim == kH && d.hidden_dim_color == kH && d.geo_feat_dim == kG,
                 "nerfacto field: hidden_dim / hidden_dim_color / geo_feat_dim must be 256 / 256 / 64 (every shipped config), got %d / %d / %d",
                 d.hidden_dim, d.hidden_dim_color, d.geo_feat_dim);
    HUGS_REQUIRE(d.n_levels <= 16, "nerfacto field: at most 16 levels");
    HUGS_REQUIRE(d.appearance_dim >= 0 && d.appearance_dim <= 64, "nerfacto field: appearance_dim must be in [0,64]");
    HUGS_REQUIRE(d.appearance_dim == 0 || d.num_embeddings > 0, "nerfacto field: num_embeddings must be > 0");
  }
  int ndev = 0;
  HUGS_CUDA(cudaGetDeviceCount(&ndev));
  HUGS_REQUIRE(ndev > 0, "no CUDA device: this library has no CPU path");
  hugs_hashfield* h = new (std::nothrow) hugs_hashfield();
  if (!h) { set_error("out of host memory"); return HUGS_ERR_NOMEM; }
  auto fail = [&](int code) { hugs_hashfield_destroy(h); return code; };
  h->d = d; h->is_prop = is_prop;
  HUGS_REQUIRE(d.precision == 0 || d.precision == HUGS_PRECISION_BF16_TC || d.precision == HUGS_PRECISION_TC_SPLIT,
               "hash field: precision must be HUGS_PRECISION_BF16_TC or HUGS_PRECISION_TC_SPLIT, got %d", d.precision);
  h->split = !is_prop && d.precision == HUGS_PRECISION_TC_SPLIT; h->parts = h->split ? 2 : 1;
  if (cudaGetDevice(&h->device) != cudaSuccess) return fail(HUGS_ERR_CUDA);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) return fail(HUGS_ERR_CUDA);
  h->num_sms = prop.multiProcessorCount;
  if (!is_prop && prop.major != 10) { set_error("the nerfacto field needs an sm_100 device (found sm_%d%d)", prop.major, prop.minor); return fail(HUGS_ERR_UNSUPPORTED); }
  // ---- grid geometry (tiny-cuda-nn GridEncoding: float arithmetic throughout) ----
  HashGridCfg& g = h->g;
  g.L = d.n_levels; g.F = 2; g.in_dim = 2 * d.n_levels; g.bound = d.bound; g.contract = d.contract;
  const float l2s = log2f(d.per_level_scale);
  uint32_t off = 0;
  for (int l = 0; l < g.L; ++l) {
    const float scale = exp2f((float)l * l2s) * (float)d.base_res - 1.0f;
    const uint32_t res = (uint32_t)ceilf(scale) + 1;
    g.scale[l] = scale; g.res[l] = res; g.off[l] = off;
    const double dense = (double)res * res * res;
    uint32_t cnt = dense > (double)(0xFFFFFFFFu / 2) ? 0xFFFFFFFFu / 2 : (uint32_t)dense;
    cnt = ((cnt + 7) / 8) * 8;
    cnt = std::min(cnt, 1u << d.log2_hashmap_size);
    off += cnt;
  }
  g.off[g.L] = off;
  h->grid_floats = (int64_t)off * 2;
  // ---- flat MLP layout ([in, out] kernels) ----
  int64_t o = 0;
  if (is_prop) {
    hf_tensor(h, "base0/kernel", g.in_dim, kPropHidden, &o, nullptr);
    hf_tensor(h, "base0/bias", 1, kPropHidden, &o, nullptr);
    hf_tensor(h, "density/kernel", kPropHidden, 1, &o, nullptr);
    hf_tensor(h, "density/bias", 1, 1, &o, nullptr);
  } else {
    h->head0_in = kG + kSH + d.appearance_dim;
    hf_tensor(h, "base0/kernel", g.in_dim, kH, &o, &h->o_base0_k);
    hf_tensor(h, "base0/bias", 1, kH, &o, &h->o_base0_b);
    hf_tensor(h, "density/kernel", kH, 1, &o, &h->o_dens_k);
    hf_tensor(h, "density/bias", 1, 1, &o, &h->o_dens_b);
    hf_tensor(h, "geo/kernel", kH, kG, &o, &h->o_geo_k);
    hf_tensor(h, "geo/bias", 1, kG, &o, &h->o_geo_b);
    hf_tensor(h, "head0/kernel", h->head0_in, kH, &o, &h->o_head0_k);      // rows: [geometry | SH | appearance]
    hf_tensor(h, "head0/bias", 1, kH, &o, &h->o_head0_b);
    hf_tensor(h, "head1/kernel", kH, kH, &o, &h->o_head1_k);
    hf_tensor(h, "head1/bias", 1, kH, &o, &h->o_head1_b);
    hf_tensor(h, "rgb/kernel", kH, 3, &o, &h->o_rgb_k);
    hf_tensor(h, "rgb/bias", 1, 3, &o, &h->o_rgb_b);
    if (d.appearance_dim > 0) hf_tensor(h, "embedding", d.num_embeddings, d.appearance_dim, &o, &h->o_emb);
  }
  h->mlp_floats = o;
  h->cap = ((d.max_samples + 255) / 256) * 256;
  int rc;
  if (!is_prop) {
    // forward rows: base0 256 | heads: geo 128 (64 valid) + density 16 | head0 256 | head1 256 | rgb 16
    int rf = 0;
    h->rf_base0 = rf; rf += 256; h->rf_heads = rf; rf += 144; h->rf_head0 = rf; rf += 256; h->rf_head1 = rf; rf += 256;
    h->rf_rgb = rf; rf += 16;
    h->rows_f = ((rf + 127) / 128) * 128;
    // dgrad rows (= inputs): base0 128 (32 valid) | geo 256 | head0 geometry rows 128 (64 valid) | head1 256
    int rb = 0;
    h->rb_base0 = rb; rb += 128; h->rb_geo = rb; rb += 256; h->rb_head0 = rb; rb += 128; h->rb_head1 = rb; rb += 256;
    h->rows_b = rb;
    const int P = h->parts;
    if ((rc = hf_alloc(h, &h->wt, (size_t)h->rows_f * P * 256)) || (rc = hf_alloc(h, &h->wn, (size_t)h->rows_b * P * 256)) ||
        (rc = hf_alloc(h, &h->tab, 2048)))
      return fail(rc);
    if ((rc = make_map(&h->map_wt128, h->wt, h->rows_f * P, 256, 128)) || (rc = make_map(&h->map_wt64, h->wt, h->rows_f * P, 256, 64)) ||
        (rc = make_map(&h->map_wt8, h->wt, h->rows_f * P, 256, 8)) || (rc = make_map(&h->map_wn128, h->wn, h->rows_b * P, 256, 128)) ||
        (rc = make_map(&h->map_wn64, h->wn, h->rows_b * P, 256, 64)))
      return fail(rc);
    for (int i = HF_FEAT; i <= HF_H1; ++i) {
      if ((rc = hf_alloc(h, &h->buf[i], (size_t)h->cap * P * h->buf_cols[i]))) return fail(rc);
      if ((rc = make_map(&h->map128[i], h->buf[i], h->cap * P, h->buf_cols[i], 128)) ||
          (rc = make_map(&h->map64[i], h->buf[i], h->cap * P, h->buf_cols[i], 64)))
        return fail(rc);
    }
    h->max_rays = d.max_rays > 0 ? d.max_rays : std::max(1, d.max_samples / 16);
    if ((rc = hf_alloc(h, &h->ray_in, (size_t)h->max_rays * (kSH + 64))) || (rc = hf_alloc(h, &h->ray_bias, (size_t)h->max_rays * kH)) ||
        (rc = hf_alloc(h, &h->inside, (size_t)h->cap)))
      return fail(rc);
    if ((rc = dense_tc_init()) || (rc = field_chain_init())) return fail(rc);
  }
  {
    cudaError_t e = cudaSuccess;
    auto opt = [&](auto kernel) { if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); };
    opt(prop_field_kernel<false, 10>); opt(prop_field_kernel<false, 14>); opt(prop_field_kernel<false, 0>);
    opt(prop_field_kernel<true, 10>); opt(prop_field_kernel<true, 14>); opt(prop_field_kernel<true, 0>);
    if (e != cudaSuccess) return fail(cuda_fail(e, "hash field smem opt-in", __FILE__, __LINE__));
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(cuda_fail(cudaGetLastError(), "hugs_hashfield_create sync", __FILE__, __LINE__));
  *out = h;
  return HUGS_OK;
}

HUGS_API int hugs_hashfield_destroy(hugs_hashfield* h) {
  if (!h) return HUGS_OK;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return HUGS_OK;
}

HUGS_API int64_t hugs_hashfield_grid_floats(const hugs_hashfield* h) { return h ? h->grid_floats : -1; }
HUGS_API int64_t hugs_hashfield_mlp_floats(const hugs_hashfield* h) { return h ? h->mlp_floats : -1; }

HUGS_API int hugs_hashfield_layout(const hugs_hashfield* h, hugs_tensor_desc* out, int32_t capacity, int32_t* count) {
  HUGS_REQUIRE(h && count, "hugs_hashfield_layout: null argument");
  *count = (int32_t)h->tensors.size();
  if (!out) return HUGS_OK;
  HUGS_REQUIRE(capacity >= *count, "hugs_hashfield_layout: capacity %d < %d tensors", capacity, *count);
  memcpy(out, h->tensors.data(), sizeof(hugs_tensor_desc) * h->tensors.size());
  return HUGS_OK;
}

HUGS_API int hugs_hashfield_level_info(const hugs_hashfield* h, int32_t level, float* scale, uint32_t* resolution, uint32_t* offset,
                                       uint32_t* entries) {
  HUGS_REQUIRE(h && level >= 0 && level < h->g.L, "hugs_hashfield_level_info: bad level");
  if (scale) *scale = h->g.scale[level];
  if (resolution) *resolution = h->g.res[level];
  if (offset) *offset = h->g.off[level];
  if (entries) *entries = h->g.off[level + 1] - h->g.off[level];
  return HUGS_OK;
}

HUGS_API int hugs_hashfield_params_changed(hugs_hashfield* h, const float* mlp, void* stream) {
  HUGS_REQUIRE(h && mlp, "hugs_hashfield_params_changed: null argument");
  if (h->is_prop) return HUGS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  FieldPackArgs a;
  memset(&a, 0, sizeof(a));
  a.params = mlp; a.wt = h->wt; a.wn = h->wn; a.ldk = 256;
  const int in = h->g.in_dim;
  auto blk = [&](int dst, int row0, int rows, int fwd, int in0, int n_in, int out0, int n_out, int stride, int64_t koff) {
    a.b[a.n++] = PackBlock{dst, row0, rows, fwd, in0, n_in, out0, n_out, stride, (long long)koff};
  };
  blk(0, h->rf_base0, 256, 1, 0, in, 0, kH, kH, h->o_base0_k);
  blk(0, h->rf_heads, 128, 1, 0, kH, 0, kG, kG, h->o_geo_k);
  blk(0, h->rf_heads + 128, 16, 1, 0, kH, 0, 1, 1, h->o_dens_k);
  blk(0, h->rf_head0, 256, 1, 0, kG, 0, kH, kH, h->o_head0_k);
  blk(0, h->rf_head1, 256, 1, 0, kH, 0, kH, kH, h->o_head1_k);
  blk(0, h->rf_rgb, 16, 1, 0, kH, 0, 3, 3, h->o_rgb_k);
  blk(1, h->rb_base0, 128, 0, 0, in, 0, kH, kH, h->o_base0_k);
  blk(1, h->rb_geo, 256, 0, 0, kH, 0, kG, kG, h->o_geo_k);
  blk(1, h->rb_head0, 128, 0, 0, kG, 0, kH, kH, h->o_head0_k);
  blk(1, h->rb_head1, 256, 0, 0, kH, 0, kH, kH, h->o_head1_k);
  a.lo_wt = (long long)h->rows_f * a.ldk; a.lo_wn = (long long)h->rows_b * a.ldk;
  for (int part = 0; part < h->parts; ++part) {
    a.part = part;
    field_pack_kernel<<<dim3(64, a.n), 256, 0, st>>>(a);
    HUGS_LAUNCH_CHECK();
  }
  field_table_kernel<<<(928 + 768 + 255) / 256, 256, 0, st>>>(mlp, h->tab, h->o_base0_b, h->o_dens_b, h->o_geo_b, h->o_head1_b,
                                                             h->o_rgb_b, h->o_dens_k, h->o_rgb_k, h->split ? 1 : 0);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

HUGS_API int hugs_hashfield_encode(hugs_hashfield* h, const float* grid, const hugs_rays* rays, const float* tdist, int32_t n_rays,
                                   int32_t n_samples, float* features, void* stream) {
  HUGS_REQUIRE(h && grid && tdist && features, "hugs_hashfield_encode: null argument");
  int rc = hf_check_rays(h, rays, n_rays, n_samples);
  if (rc) return rc;
  HUGS_REQUIRE(h->g.L <= 16, "hugs_hashfield_encode: at most 16 levels");
  const int M = n_rays * n_samples;
  if (M == 0) return HUGS_OK;
  EncodeArgs a{field_rays(rays, tdist, n_rays, n_samples), h->g, reinterpret_cast<const float2*>(grid), nullptr, features, M, nullptr};
  hash_encode_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

HUGS_API int hugs_hashfield_forward(hugs_hashfield* h, const float* grid, const float* mlp, const hugs_rays* rays,
                                    const float* tdist, int32_t n_rays, int32_t n_samples, int32_t training, int32_t zero_app,
                                    float* raw_out, void* stream) {
  HUGS_REQUIRE(h && grid && mlp && tdist && raw_out, "hugs_hashfield_forward: null argument");
  int rc = hf_check_rays(h, rays, n_rays, n_samples);
  if (rc) return rc;
  const int M = n_rays * n_samples;
  if (M == 0) return HUGS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const FieldRays fr = field_rays(rays, tdist, n_rays, n_samples);
  if (h->is_prop) {
    PropArgs a{fr, h->g, reinterpret_cast<const float2*>(grid), mlp, raw_out, nullptr, nullptr, nullptr};
    return launch_prop_field<false>(a, std::min((M + 127) / 128, h->num_sms * 8), st);
  }
  HUGS_REQUIRE(n_rays <= h->max_rays, "nerfacto field: %d rays exceed the per-ray workspace (%d)", n_rays, h->max_rays);
  if (training && (rc = hf_ensure_training(h))) return rc;
  const int rows_pad = ((M + 255) / 256) * 256;
  const int app = h->d.appearance_dim;
  // 1. hash features (bf16 [rows, 64]) and the in-range mask
  {
    EncodeArgs a{fr, h->g, reinterpret_cast<const float2*>(grid), h->buf[HF_FEAT], nullptr, rows_pad,
                 h->split ? h->buf[HF_FEAT] + (size_t)h->cap * 64 : nullptr};
    hash_encode_kernel<<<(rows_pad + 127) / 128, 128, 0, st>>>(a);
    HUGS_LAUNCH_CHECK();
  }
  // 2. per-ray inputs of the colour MLP folded into a per-ray bias of its first layer
  sh_inputs_kernel<<<(n_rays + 127) / 128, 128, 0, st>>>(rays->viewdirs, rays->embed_idx, h->o_emb >= 0 ? mlp + h->o_emb : nullptr,
                                                         n_rays, app, h->d.num_embeddings, zero_app, h->ray_in);
  HUGS_LAUNCH_CHECK();
  ray_bias_kernel<<<(n_rays * kH + 255) / 256, 256, 0, st>>>(h->ray_in, kSH + app, mlp + h->o_head0_k, kG, mlp + h->o_head0_b, kH,
                                                             n_rays, h->split ? 1 : 0, h->ray_bias);
  HUGS_LAUNCH_CHECK();
  if (h->use_chain && !h->split) {
    // 3-7. all five Dense layers in one launch: the activations stay in shared memory between the layers (field_chain.cu)
    FieldChainParams c;
    memset(&c, 0, sizeof(c));
    c.a_map = h->map128[HF_FEAT]; c.b_map = h->map_wt128; c.b_map_64 = h->map_wt64; c.b_map_8 = h->map_wt8;
    c.out_map[0] = h->map128[HF_ACT0]; c.out_map[1] = h->map128[HF_GEO]; c.out_map[2] = h->map128[HF_H0];
    c.out_map[3] = h->map128[HF_H1];
    auto link = [&](int l, int kp, int b_row0, int bias_off, int store, int gate_row0) {
      FieldChainLink& L = c.link[l];
      L.kp = kp; L.b_row0 = b_row0; L.bias_off = bias_off; L.store = store; L.gate_row0 = gate_row0; L.n_tiles = 1;
      return &L;
    };
    const int gate0 = (training && h->use_gate) ? 0 : -1, gate1 = (training && h->use_gate) ? h->cap : -1,
              gate2 = (training && h->use_gate) ? 2 * h->cap : -1;      // (the backward chain starts from the gate bits of h1)
    FieldChainLink* L = link(0, 1, h->rf_base0, 0, training ? 1 : 0, gate0);
    L->tile_n0[0] = 0; L->tile_bn[0] = 256; L->tile_epi[0] = DE_RELU;
    L = link(1, 4, h->rf_heads, 256, training ? 1 : 0, -1);
    L->n_tiles = 2; L->tile_n0[0] = 0; L->tile_bn[0] = 128; L->tile_epi[0] = DE_LINEAR;
    L->tile_n0[1] = 128; L->tile_bn[1] = 16; L->tile_epi[1] = DE_HEAD_F32; L->raw_chan0 = 0; L->raw_nchan = 1;
    L = link(2, 1, h->rf_head0, 0, training ? 1 : 0, gate1);
    L->tile_n0[0] = 0; L->tile_bn[0] = 256; L->tile_epi[0] = DE_VIEW;
    L = link(3, 4, h->rf_head1, 400, training ? 1 : 0, gate2);
    L->tile_n0[0] = 0; L->tile_bn[0] = 256; L->tile_epi[0] = DE_RELU;
    L = link(4, 4, h->rf_rgb, 656, 0, -1);
    L->tile_n0[0] = 0; L->tile_bn[0] = 16; L->tile_epi[0] = DE_HEAD_F32; L->raw_chan0 = 1; L->raw_nchan = 3;
    c.n_links = 5; c.m_rows = M; c.m_tiles = (M + 255) / 256; c.S = n_samples;
    c.bias = h->tab; c.n_bias = 928 + 768; c.viewbias = h->ray_bias; c.view_ld = kH;
    c.raw_out = raw_out; c.raw_c = 4;
    c.gate_out = (training && h->use_gate) ? h->gate : nullptr; c.gate_ld = kH / 32;
    if ((rc = field_chain_launch(c, h->num_sms, st))) return rc;
    inside_mask_kernel<<<(M + 255) / 256, 256, 0, st>>>(fr, h->g, h->inside, raw_out, 4);
    HUGS_LAUNCH_CHECK();
    return HUGS_OK;
  }
  DenseParams p;
  // 3. base MLP layer 0: features -> 256 (ReLU)
  dense_common(&p, h, M);
  dense_a(&p, h, HF_FEAT, 1, h->rows_f);
  p.b_row0 = h->rf_base0; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_RELU;
  p.bias = h->tab; p.out_map = h->map128[HF_ACT0];
  if (training && h->use_gate) { p.gate_out = h->gate; p.gate_ld = kH / 32; p.gate_row0 = 0; }
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  // 4. heads of the base MLP: geometry features (linear, bf16) + raw density (fp32 column 0 of raw)
  dense_common(&p, h, M);
  dense_a(&p, h, HF_ACT0, 4, h->rows_f);
  p.b_row0 = h->rf_heads; p.n_tiles = 2;
  p.tile_n0[0] = 0; p.tile_bn[0] = 128; p.tile_epi[0] = DE_LINEAR;
  p.tile_n0[1] = 128; p.tile_bn[1] = 16; p.tile_epi[1] = DE_HEAD_F32;
  p.bias = h->tab + 256; p.out_map = h->map128[HF_GEO];
  p.raw_out = raw_out; p.raw_c = 4; p.raw_chan0 = 0; p.raw_nchan = 1;
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  // 5. colour MLP layer 0: geometry features (K = 64) + per-ray bias -> 256 (ReLU)
  dense_common(&p, h, M);
  dense_a(&p, h, HF_GEO, 1, h->rows_f);
  p.b_row0 = h->rf_head0; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_VIEW;
  p.viewbias = h->ray_bias; p.view_ld = kH; p.S = n_samples; p.out_map = h->map128[HF_H0];
  if (training && h->use_gate) { p.gate_out = h->gate; p.gate_ld = kH / 32; p.gate_row0 = h->cap; }
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  // 6. colour MLP layer 1
  dense_common(&p, h, M);
  dense_a(&p, h, HF_H0, 4, h->rows_f);
  p.b_row0 = h->rf_head1; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_RELU;
  p.bias = h->tab + 400; p.out_map = h->map128[HF_H1];
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  // 7. rgb head (fp32 columns 1..3 of raw)
  dense_common(&p, h, M);
  dense_a(&p, h, HF_H1, 4, h->rows_f);
  p.b_row0 = h->rf_rgb; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 16; p.tile_epi[0] = DE_HEAD_F32;
  p.bias = h->tab + 656; p.out_map = h->map128[HF_H1];
  p.raw_out = raw_out; p.raw_c = 4; p.raw_chan0 = 1; p.raw_nchan = 3;
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  // 8. density * selector (nerfacto.py:836): out-of-range samples get a raw density of -inf
  inside_mask_kernel<<<(M + 255) / 256, 256, 0, st>>>(fr, h->g, h->inside, raw_out, 4);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

HUGS_API int hugs_hashfield_backward(hugs_hashfield* h, const float* grid, const float* mlp, const hugs_rays* rays,
                                     const float* tdist, int32_t n_rays, int32_t n_samples, const float* d_raw,
                                     float* grid_grad, float* mlp_grad, void* stream) {
  HUGS_REQUIRE(h && grid && mlp && tdist && d_raw && grid_grad && mlp_grad, "hugs_hashfield_backward: null argument");
  int rc = hf_check_rays(h, rays, n_rays, n_samples);
  if (rc) return rc;
  const int M = n_rays * n_samples;
  HUGS_REQUIRE(M > 0, "hugs_hashfield_backward: empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  const FieldRays fr = field_rays(rays, tdist, n_rays, n_samples);
  HUGS_CUDA(cudaMemsetAsync(mlp_grad, 0, sizeof(float) * h->mlp_floats, st));
  if (h->is_prop) {
    PropArgs a{fr, h->g, reinterpret_cast<const float2*>(grid), mlp, nullptr, d_raw, reinterpret_cast<float2*>(grid_grad), mlp_grad};
    return launch_prop_field<true>(a, std::min((M + 127) / 128, h->num_sms * 3), st);
  }
  HUGS_REQUIRE(h->train_ready, "nerfacto field: backward without a training forward");
  const int rows_pad = ((M + 255) / 256) * 256;
  const int app = h->d.appearance_dim;
  if (h->built_for != M) {
    hf_build_wgrad(h, M);
    HUGS_REQUIRE(h->items_host.size() <= (size_t)kMaxWgItems, "nerfacto field: too many weight-gradient items");
    HUGS_CUDA(cudaMemcpyAsync(h->items_dev, h->items_host.data(), sizeof(WgItem) * h->items_host.size(), cudaMemcpyHostToDevice, st));
    HUGS_CUDA(cudaStreamSynchronize(st));
    h->built_for = M;
  }
  int launch = 0;
  auto wgrad = [&]() {
    const auto& L = h->launches[launch++];
    return wgrad_launch_raw(h->num_sms, 0, 1, 1 << 30, h->map64, HF_MAPS, h->items_dev + L.first, L.second, mlp_grad, st);
  };
  if (h->use_chain && h->use_gate && !h->split) {
    // the four dgrad GEMMs and the head-gradient start op as ONE chain launch (field_chain.cu, backward program): dZ stays in
    // shared memory between the layers; every dZ the weight gradients need is written once by TMA store
    FieldChainParams c;
    memset(&c, 0, sizeof(c));
    c.b_map = h->map_wn128; c.b_map_64 = h->map_wn64; c.b_map_8 = h->map_wn64; c.a_map = h->map128[HF_DZH1];
    c.out_map[0] = h->map128[HF_DZH0]; c.out_map[1] = h->map128[HF_DGEO]; c.out_map[2] = h->map128[HF_DZA0];
    c.out_map[3] = h->map128[HF_DFEAT];
    auto link = [&](int l, int kp, int b_row0, int bn, int epi, int gate_in_row0, int rank1) {
      FieldChainLink& L = c.link[l];
      L.kp = kp; L.b_row0 = b_row0; L.n_tiles = 1; L.tile_n0[0] = 0; L.tile_bn[0] = bn; L.tile_epi[0] = epi;
      L.store = 1; L.gate_row0 = -1; L.gate_in_row0 = gate_in_row0; L.rank1 = rank1;
    };
    link(0, 4, h->rb_head1, 256, DE_BWD_RELU, h->cap, 0);      // dZ_head0 = (dZ_head1 . W_head1^T) * [h0 > 0]
    link(1, 4, h->rb_head0, 128, DE_BWD_LINEAR, 0, 0);         // d_geo = dZ_head0 . W_head0[geometry rows]^T
    link(2, 1, h->rb_geo, 256, DE_BWD_RELU, 0, 1);             // dZ_act0 = (d_geo . W_geo^T + d_density (x) w_density) * [act0 > 0]
    link(3, 4, h->rb_base0, 128, DE_BWD_LINEAR, 0, 0);         // d_features = dZ_act0 . W_base0^T
    c.n_links = 4; c.m_rows = M; c.m_tiles = (M + 255) / 256; c.S = n_samples;
    c.start_mode = 1; c.d_raw = d_raw; c.inside = h->inside; c.bias = h->tab; c.n_bias = 928 + 768; c.w_rgb_off = 928;
    c.rank1_off = 672;
    c.gate_in = h->gate; c.gate_ld = kH / 32; c.start_gate_row0 = 2 * h->cap; c.start_map = h->map128[HF_DZH1];
    c.dh_out = h->buf[HF_DH];
    if ((rc = field_chain_launch(c, h->num_sms, st))) return rc;
    for (int i = 0; i < 4; ++i)
      if ((rc = wgrad())) return rc;
    ray_colsum_kernel<<<n_rays, kH, 0, st>>>(h->buf[HF_DZH0], nullptr, n_samples, n_rays, h->dzsum);
    HUGS_LAUNCH_CHECK();
    ray_input_wgrad_kernel<<<dim3(kSH + app, 32), kH, 0, st>>>(h->ray_in, kSH + app, h->dzsum, n_rays, kG, 0, mlp_grad + h->o_head0_k);
    HUGS_LAUNCH_CHECK();
    if (app > 0) {
      app_embed_grad_kernel<<<(n_rays + 7) / 8, 256, 0, st>>>(h->dzsum, rays->embed_idx, mlp + h->o_head0_k, kG + kSH, app, n_rays,
                                                              h->d.num_embeddings, 0, mlp_grad + h->o_emb);
      HUGS_LAUNCH_CHECK();
    }
    ScatterArgs sa{fr, h->g, h->buf[HF_DFEAT], 128, reinterpret_cast<float2*>(grid_grad), nullptr};
    hash_scatter_kernel<<<(M + 127) / 128, 128, 0, st>>>(sa);
    HUGS_LAUNCH_CHECK();
    return HUGS_OK;
  }
  // start: dZ_head1 from d_rgb, head-gradient rows, masked density gradient
  field_bwd_start_kernel<<<(rows_pad + 7) / 8, 256, 0, st>>>(d_raw, h->buf[HF_H1], h->tab + 928, h->inside, M, rows_pad, h->buf[HF_DZH1],
                                                  h->buf[HF_DH], h->d_dens,
                                                  h->split ? h->buf[HF_DZH1] + (size_t)h->cap * kH : nullptr,
                                                  h->split ? h->buf[HF_DH] + (size_t)h->cap * kHeadCols : nullptr,
                                                  nullptr);   // gate bits of H1 measured no faster here (the write in the forward GEMM costs 0.1 ms)
  HUGS_LAUNCH_CHECK();
  if ((rc = wgrad())) return rc;
  DenseParams p;
  // dZ_head0 = (dZ_head1 . W_head1^T) * [hact0 > 0]
  dense_common(&p, h, M);
  dense_a(&p, h, HF_DZH1, 4, h->rows_b);
  p.b_map = h->map_wn128; p.b_row0 = h->rb_head1; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_BWD_RELU;
  p.mask_act = h->buf[HF_H0]; p.mask_ld = kH; p.mask_row0 = 0; p.out_map = h->map128[HF_DZH0];
  if (h->use_gate) { p.gate_in = h->gate; p.gate_ld = kH / 32; p.gate_row0 = h->cap; }
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  if ((rc = wgrad())) return rc;
  // per-ray inputs of head0: SH / appearance rows of its kernel and the appearance embedding rows
  ray_colsum_kernel<<<n_rays, kH, 0, st>>>(h->buf[HF_DZH0], h->split ? h->buf[HF_DZH0] + (size_t)h->cap * kH : nullptr, n_samples,
                                           n_rays, h->dzsum);
  HUGS_LAUNCH_CHECK();
  ray_input_wgrad_kernel<<<dim3(kSH + app, 32), kH, 0, st>>>(h->ray_in, kSH + app, h->dzsum, n_rays, kG, h->split ? 1 : 0,
                                                             mlp_grad + h->o_head0_k);
  HUGS_LAUNCH_CHECK();
  if (app > 0) {
    app_embed_grad_kernel<<<(n_rays + 7) / 8, 256, 0, st>>>(h->dzsum, rays->embed_idx, mlp + h->o_head0_k, kG + kSH, app,
                                                                      n_rays, h->d.num_embeddings, h->split ? 1 : 0,
                                                                      mlp_grad + h->o_emb);
    HUGS_LAUNCH_CHECK();
  }
  // d_geo = dZ_head0 . W_head0[geometry rows]^T  (64 columns, linear)
  dense_common(&p, h, M);
  dense_a(&p, h, HF_DZH0, 4, h->rows_b);
  p.b_map_64 = h->map_wn64; p.b_row0 = h->rb_head0; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 128; p.tile_epi[0] = DE_BWD_LINEAR;
  p.out_map = h->map128[HF_DGEO];
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  if ((rc = wgrad())) return rc;
  // dZ_act0 = (d_geo . W_geo^T + d_density (x) w_density) * [act0 > 0]
  dense_common(&p, h, M);
  dense_a(&p, h, HF_DGEO, 1, h->rows_b);
  p.b_map = h->map_wn128; p.b_row0 = h->rb_geo; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 256; p.tile_epi[0] = DE_BWD_RELU;
  p.mask_act = h->buf[HF_ACT0]; p.mask_ld = kH; p.mask_row0 = 0;
  if (h->use_gate) { p.gate_in = h->gate; p.gate_ld = kH / 32; p.gate_row0 = 0; }
  p.rank1_row = h->d_dens; p.rank1_stride = 1; p.rank1_col = h->tab + 672; p.out_map = h->map128[HF_DZA0];
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  if ((rc = wgrad())) return rc;
  // d_features = dZ_act0 . W_base0^T (32 columns), then the trilinear scatter into the grid gradient
  dense_common(&p, h, M);
  dense_a(&p, h, HF_DZA0, 4, h->rows_b);
  p.b_map_64 = h->map_wn64; p.b_row0 = h->rb_base0; p.n_tiles = 1; p.tile_n0[0] = 0; p.tile_bn[0] = 128; p.tile_epi[0] = DE_BWD_LINEAR;
  p.out_map = h->map128[HF_DFEAT];
  if ((rc = dense_tc_launch(p, h->num_sms, st))) return rc;
  ScatterArgs sa{fr, h->g, h->buf[HF_DFEAT], 128, reinterpret_cast<float2*>(grid_grad),
                 h->split ? h->buf[HF_DFEAT] + (size_t)h->cap * 128 : nullptr};
  hash_scatter_kernel<<<(M + 127) / 128, 128, 0, st>>>(sa);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}
