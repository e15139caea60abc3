// Tensor-core MLP path (HUGS_PRECISION_BF16_TC): bf16 x bf16 -> fp32 on tcgen05, sm_100a only.
//
// One persistent CTA per SM walks 128-sample tiles.  For every tile the whole MLP of
// models.py:437-519 runs as a *chain* of GEMMs without leaving the SM:
//
//   TMA (weights K-panels, streamed feature panels) -> 6-stage shared-memory ring
//   tcgen05.mma (M=128, N=128|16, K=16, fp32 accumulators in TMEM, one issuing thread)
//   epilogue warps: tcgen05.ld -> bias/ReLU -> bf16 -> 128B-swizzled shared-memory panels that are
//   directly the next layer's A operand (activations never touch HBM in inference;
//   in training each panel is additionally TMA-stored for the weight-gradient pass).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-9 = two epilogue groups of 128 threads (group g owns output columns [128g, 128g+128)).
// TMEM: 512 columns = two 256-column accumulators, ping-ponged by layer so that the epilogue of
// layer l overlaps the MMAs of layer l+1 (n-half-major issue order: columns [0,128) complete first).
//
// Layer schedule, packing and the feature-column permutation are documented in DESIGN.md.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>
#include <vector>

#include "encode.cuh"
#include "ptx.cuh"
#include "tc_device.cuh"
#include "tc_internal.h"

namespace hugs {

namespace {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------
// bf16 feature encoder (throughput mode): one thread = (sample, half of the basis directions)
// ------------------------------------------------------------------------------------------
struct EncArgs {
  const float* origins; const float* directions; const float* radii; const float* tdist;
  const float* basis;     // [3][nb]
  int n_samples, n_rows_pad, S, nb, min_deg, ndeg, ray_shape, contract;
  __nv_bfloat16* feat;    // [rows, 512] (already offset to the level's first row)
};

#ifndef HUGS_ENC_ROWS
#define HUGS_ENC_ROWS 8
#endif
#ifndef HUGS_ENC_SPLIT
#define HUGS_ENC_SPLIT 4
#endif
constexpr int kEncRows = HUGS_ENC_ROWS;
constexpr int kEncSplit = HUGS_ENC_SPLIT;     // threads per sample (each takes a contiguous range of basis directions)
// kNdeg > 0: number of IPE degrees known at compile time (12 for every shipped config): the per-direction word buffer
// stays in registers instead of a dynamically indexed local array.
template <int kNdeg>
__global__ void __launch_bounds__(kEncSplit * kEncRows) encode_bf16_kernel(EncArgs a) {
  __shared__ __align__(16) uint4 tile[kEncRows * 64];   // rows x 64 chunks of 16 B, chunk index swizzled
  const int tid = threadIdx.x, rl = tid % kEncRows, part = tid / kEncRows;
  const int s = blockIdx.x * kEncRows + rl;
  const int b_beg = (a.nb * part) / kEncSplit, b_end = (a.nb * (part + 1)) / kEncSplit;
  const bool half = part == kEncSplit - 1;      // the last part also clears the padding columns
  const int ndeg = kNdeg > 0 ? kNdeg : a.ndeg;
  const int cpb = ndeg >> 2;                      // 16-byte chunks per basis direction
  if (s < a.n_samples) {
    const int ray = s / a.S, i = s % a.S;
    float o[3], d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { o[c] = a.origins[ray * 3 + c]; d[c] = a.directions[ray * 3 + c]; }
    const float t0 = a.tdist[(size_t)ray * (a.S + 1) + i], t1 = a.tdist[(size_t)ray * (a.S + 1) + i + 1];
    SampleGauss g;
    frustum_gaussian(o, d, a.radii[ray], t0, t1, a.ray_shape, a.contract, g);
    const float sc0 = exp2f((float)a.min_deg);
    for (int b = b_beg; b < b_end; ++b) {
      float p[3] = {a.basis[b], a.basis[a.nb + b], a.basis[2 * a.nb + b]};
      float mu, var;
      lift_basis(g, d, p, mu, var);
      // phase in turns, split hi/lo so that 2^k * phase (mod 1) stays accurate at degree 11
      const float kInv2PiHi = 0.15915494309189535f, kInv2PiLo = -1.2154201256553420e-10f;   // 1/(2 pi)
      const float r_hi = mu * kInv2PiHi;
      const float r_lo = fmaf(mu, kInv2PiHi, -r_hi) + mu * kInv2PiLo;
      const float av = -0.5f * 1.4426950408889634f * var;                                  // exp(x)=2^(x log2e)
      float sc = sc0;
      uint32_t w[16];
#pragma unroll
      for (int k = 0; k < (kNdeg > 0 ? kNdeg : 16); k += 2) {
        if (kNdeg == 0 && k >= ndeg) break;
        float t = r_hi * sc;
        float f = (t - rintf(t)) + r_lo * sc;
        float x = f * 6.283185307179586f;
        float sn = __sinf(x), cs = __cosf(x);
        float e = exp2f(av * sc * sc);
        w[k] = ptx::pack_bf16x2(e * sn, e * cs);
        float sn2 = 2.f * sn * cs, cs2 = 1.f - 2.f * sn * sn;
        float e2 = e * e; e2 = e2 * e2;
        w[k + 1] = ptx::pack_bf16x2(e2 * sn2, e2 * cs2);
        sc *= 4.f;
      }
#pragma unroll
      for (int c = 0; c < (kNdeg > 0 ? kNdeg / 4 : 4); ++c) {
        if (kNdeg == 0 && c >= cpb) break;
        uint4 v = make_uint4(w[c * 4], w[c * 4 + 1], w[c * 4 + 2], w[c * 4 + 3]);
        tile[rl * 64 + swz_chunk(rl, b * cpb + c)] = v;
      }
    }
    if (half) {
      for (int c = a.nb * cpb; c < 64; ++c) tile[rl * 64 + swz_chunk(rl, c)] = make_uint4(0, 0, 0, 0);
    }
  } else if (s < a.n_rows_pad) {
    // rows between n_samples and the 128-row tile boundary: finite (zero) features so that the saved
    // activations of padding rows can never poison the weight-gradient reduction
    for (int c = part * (64 / kEncSplit); c < (part + 1) * (64 / kEncSplit); ++c)
      tile[rl * 64 + swz_chunk(rl, c)] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  // coalesced copy-out of the block's consecutive rows
  const int rows = min(kEncRows, a.n_rows_pad - blockIdx.x * kEncRows);
  uint4* dst = reinterpret_cast<uint4*>(a.feat) + (size_t)blockIdx.x * kEncRows * 64;
  for (int e = tid; e < rows * 64; e += kEncSplit * kEncRows) {
    int r = e >> 6, c = e & 63;
    dst[e] = tile[r * 64 + swz_chunk(r, c)];
  }
}

// viewbias[ray][c] = sum_j bf16(view_in[ray][j]) * bf16(W_view[256 + j][c]) + b_view[c]
__global__ void viewbias_kernel(const float* view_in, int view_in_dim, const float* params, long long koff,
                                long long boff, int bott_w, int out, int n_rays, float* vb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * out) return;
  const int ray = idx / out, c = idx % out;
  float acc = 0.f;
  for (int j = 0; j < view_in_dim; ++j) {
    float x = __bfloat162float(__float2bfloat16(view_in[(size_t)ray * view_in_dim + j]));
    float w = __bfloat162float(__float2bfloat16(params[koff + (long long)(bott_w + j) * out + c]));
    acc = fmaf(x, w, acc);
  }
  vb[idx] = acc + params[boff + c];
}

// ------------------------------------------------------------------------------------------
// parameter packing: flat fp32 (flax layout) -> bf16 operand tensors
// ------------------------------------------------------------------------------------------
struct PackArgs {
  TcMlp::PackLayer layers[16];
  int n_layers, rows_f, rows_b, nb, ndeg, feat_dim, bias_floats;
  const float* params;
  __nv_bfloat16* wt; __nv_bfloat16* wn; float* bias;
  int w_dens_off, w_rgb_off; long long dens_koff, rgb_koff; int dens_in, rgb_in;
  __nv_bfloat16* bias_img; int n_bias_layers; int bias_layer[4 * kBiasChunks];   // dense-layer index of biased layer j
};

__global__ void pack_params_kernel(PackArgs a) {
  const long long nf = (long long)a.rows_f * kKP, nbk = (long long)a.rows_b * kW;
  const long long nimg = 2LL * kBiasChunks * kBiasChunkElems;
  const long long total = nf + nbk + a.bias_floats + nimg;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    if (e >= nf + nbk + a.bias_floats) {
      // bias image of the CTA-pair kernel: element (n, k) of chunk c lives at (n / 8) * 64 + (n % 8) * 8 + k
      // (8 x 16-byte core matrices, no swizzle); K columns 2j', 2j'+1 = bf16 hi / lo parts of layer 4c + j'
      const int q = (int)(e - nf - nbk - a.bias_floats);
      const int rank = q / (kBiasChunks * kBiasChunkElems), w = q % (kBiasChunks * kBiasChunkElems);
      const int c = w / kBiasChunkElems, o = w % kBiasChunkElems;
      const int n = (o / 64) * 8 + (o % 64) / 8, k = o % 8;
      const int j = c * 4 + k / 2;
      float v = 0.f;
      if (j < a.n_bias_layers) {
        const auto& L = a.layers[a.bias_layer[j]];
        const int f = rank * 128 + n;
        if (f < L.out) {
          const float b = a.params[L.boff + f];
          const float hi = __bfloat162float(__float2bfloat16(b));
          v = (k & 1) ? (b - hi) : hi;
        }
      }
      a.bias_img[q] = __float2bfloat16(v);
    } else if (e < nf) {
      const int r = (int)(e / kKP), k = (int)(e % kKP);
      float v = 0.f;
      for (int l = 0; l < a.n_layers; ++l) {
        const auto& L = a.layers[l];
        if (r < L.row0 || r >= L.row0 + L.rows_pad) continue;
        const int n = r - L.row0;
        if (n >= L.out) break;
        int i = -1;
        if (k < L.x_in) i = k;
        else if (L.feat_in > 0 && k < L.x_in + kFeatPad) {
          int fp = k - L.x_in;
          if (fp < a.feat_dim) i = L.x_in + ref_feature_col(fp, a.nb, a.ndeg);
        }
        if (i >= 0 && i < L.in) v = a.params[L.koff + (long long)i * L.out + n];
        break;
      }
      a.wt[e] = __float2bfloat16(v);
    } else if (e < nf + nbk) {
      const long long q = e - nf;
      const int r = (int)(q / kW), c = (int)(q % kW);
      float v = 0.f;
      for (int l = 0; l < a.n_layers; ++l) {
        const auto& L = a.layers[l];
        if (L.brow0 < 0 || r < L.brow0 || r >= L.brow0 + kW) continue;
        const int i = r - L.brow0;
        if (c < L.out && i < L.x_in) v = a.params[L.koff + (long long)i * L.out + c];
        break;
      }
      a.wn[q] = __float2bfloat16(v);
    } else {
      const int q = (int)(e - nf - nbk);
      float v = 0.f;
      for (int l = 0; l < a.n_layers; ++l) {
        const auto& L = a.layers[l];
        if (q >= L.bias_off && q < L.bias_off + L.out) { v = a.params[L.boff + (q - L.bias_off)]; break; }
      }
      if (q >= a.w_dens_off && q < a.w_dens_off + a.dens_in)     // density kernel [in,1], bf16-rounded
        v = __bfloat162float(__float2bfloat16(a.params[a.dens_koff + (q - a.w_dens_off)]));
      if (a.w_rgb_off >= 0 && q >= a.w_rgb_off && q < a.w_rgb_off + a.rgb_in * 3)  // rgb kernel [in,3]
        v = __bfloat162float(__float2bfloat16(a.params[a.rgb_koff + (q - a.w_rgb_off)]));
      a.bias[q] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// the fused MLP chain kernel
// ------------------------------------------------------------------------------------------
struct Smem {
  uint8_t* panels;          // [kNumPanels][kPanelBytes]
  uint8_t* ring;            // [kStages][kPanelBytes]
  float* bias;              // [kBiasTab]
  uint4* steps;             // [kMaxSteps] precomputed MMA issue list
  uint64_t* full;           // [kStages]
  uint64_t* empty;          // [kStages]
  uint64_t* panel_ready;    // [kNumPanels]  (count 128: one epilogue group per panel)
  uint64_t* feat_ready;     // [kNumPanels]  (count 1 + tx: layer-0 features TMA-loaded into the panels)
  uint64_t* acc_full;       // [8]           ([7] = tile_done)
  uint64_t* panels_free;    // [1]           (count kEpiGroups: no TMA store still reads the panels)
  uint32_t* tmem_ptr;
};

__device__ __forceinline__ Smem carve(uint8_t* raw) {
  Smem s;
  // offset arithmetic on the __shared__ symbol (not an integer round trip) keeps the shared address space, so the
  // accesses below compile to LDS/STS instead of generic loads and stores
  uint8_t* base = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  s.panels = base;
  s.ring = base + kNumPanels * kPanelBytes;
  s.bias = reinterpret_cast<float*>(s.ring + kStages * kPanelBytes);
  s.steps = reinterpret_cast<uint4*>(s.bias + kBiasTab);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.steps + kMaxSteps);
  s.full = bars; s.empty = bars + kStages; s.panel_ready = bars + 2 * kStages;
  s.feat_ready = s.panel_ready + kNumPanels;
  s.acc_full = s.feat_ready + kNumPanels;
  s.panels_free = s.acc_full + 8;
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.panels_free + 1);
  return s;
}

__device__ __forceinline__ bool layer_has_mma(const TcLayer& L) {
  return L.epi != EPI_BWD_START && L.epi != EPI_BWD_START_PROP;
}

template <bool kTrain>
__global__ void __launch_bounds__(kThreads, 1) mlp_chain_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  Smem sm = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const uint32_t panels_u32 = ptx::smem_u32(sm.panels), ring_u32 = ptx::smem_u32(sm.ring);
  const uint32_t full_u32 = ptx::smem_u32(sm.full), empty_u32 = ptx::smem_u32(sm.empty);
  const uint32_t waitbar_u32 = ptx::smem_u32(sm.panel_ready);   // panel_ready[0..7] ++ feat_ready[0..7]
  const uint32_t featready_u32 = ptx::smem_u32(sm.feat_ready), accfull_u32 = ptx::smem_u32(sm.acc_full);
  const uint32_t pfree_u32 = ptx::smem_u32(sm.panels_free);

  constexpr int kProducerWarp = kEpiGroups * 4, kMmaWarp = kEpiGroups * 4 + 1;
  if (warp == kProducerWarp && lane == 0) {
    ptx::prefetch_tmap(&p.map_w128); ptx::prefetch_tmap(&p.map_w16);
    ptx::prefetch_tmap(&p.map_feat); ptx::prefetch_tmap(&p.map_save);
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&sm.full[i], 1); ptx::mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < kNumPanels; ++i) { ptx::mbar_init(&sm.panel_ready[i], 128); ptx::mbar_init(&sm.feat_ready[i], 1); }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&sm.acc_full[i], 1);
    ptx::mbar_init(sm.panels_free, kEpiGroups);
    ptx::fence_mbar_init();
  }
  int n_steps = 0;
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(sm.tmem_ptr, 512);
    // Precompute the per-tile MMA issue list (identical for every tile):
    //   x = smem address of the resident A panel (0: A is streamed through the ring)
    //   y = TMEM column of the accumulator        z = instruction descriptor
    //   w = [0,5) barrier to wait for (1..8 panel_ready, 9..16 feat_ready, 0 none) | bit 5 accumulate
    //       | bit 6 two N-halves | [8,12) acc_full barrier to commit to + 1 (0: none)
    for (int l = 0; l < p.n_layers; ++l) {
      const TcLayer& L = p.layers[l];
      if (!layer_has_mma(L)) continue;
      const int kps = L.a_res + L.a_str;
      for (int kp = 0; kp < kps; ++kp, ++n_steps) {
        if (lane != 0) continue;
        uint4 e;
        const int pi = L.a_buf * 4 + kp;
        e.x = kp < L.a_res ? panels_u32 + pi * kPanelBytes : 0u;
        e.y = (uint32_t)L.acc_col;
        e.z = ptx::make_idesc_bf16(128, L.n_mma, 0, 0);
        uint32_t wi = 0;
        if (kp < L.a_res && L.wait_panels) wi = (L.a_feat ? 9 : 1) + pi;
        e.w = wi | (kp > 0 ? 32u : 0u) | (L.n_halves == 2 ? 64u : 0u) |
              (kp == kps - 1 ? (uint32_t)(L.acc_bar + 1) << 8 : 0u);
        sm.steps[n_steps] = e;
      }
    }
  }
  for (int i = threadIdx.x; i < p.bias_floats; i += kThreads) sm.bias[i] = p.bias[i];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;
  const bool feat_resident = p.layers[0].a_feat != 0;

  if (warp == kProducerWarp) {
    // =============================== TMA producer (one lane) ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int tile_iter = 0;
      long long c_empty = 0, c_tile = 0;
      const long long c_start = clock64();
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_iter) {
        const int feat_row = p.feat_row0 + tile * kTileM;
        if (feat_resident) {
          // the tile's 512 IPE feature columns go straight into the eight activation panels, which are
          // free between tiles: every MMA of the previous tile has completed and no TMA store reads them
          if (tile_iter > 0) {
            const long long c0 = clock64();
            ptx::mbar_wait_u32(accfull_u32 + 7 * 8, (uint32_t)((tile_iter - 1) & 1));
            ptx::mbar_wait_u32(pfree_u32, (uint32_t)((tile_iter - 1) & 1));
            c_tile += clock64() - c0;
          }
          for (int kp = 0; kp < kNumPanels; ++kp) {
            ptx::mbar_expect_tx_u32(featready_u32 + kp * 8, kPanelBytes);
            ptx::tma_load_2d_u32(panels_u32 + kp * kPanelBytes, &p.map_feat, featready_u32 + kp * 8, kp * 64, feat_row);
          }
        }
        for (int l = 0; l < p.n_layers; ++l) {
          const TcLayer& L = p.layers[l];
          if (!layer_has_mma(L)) continue;
          const int a_res = L.a_res, kps = L.a_res + L.a_str, n_halves = L.n_halves;
          const uint32_t w_bytes = (uint32_t)L.n_mma * 128u;
          const CUtensorMap* wmap = L.w_map ? &p.map_w16 : &p.map_w128;
          const int w_row = L.w_row;
          for (int kp = 0; kp < kps; ++kp) {
            if (kp >= a_res) {
              ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
              ptx::mbar_expect_tx_u32(full_u32 + stage * 8, kPanelBytes);
              ptx::tma_load_2d_u32(ring_u32 + stage * kPanelBytes, &p.map_feat, full_u32 + stage * 8,
                                   (kp - a_res) * 64, feat_row);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            for (int h = 0; h < n_halves; ++h) {
              const long long c0 = clock64();
              ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
              c_empty += clock64() - c0;
              ptx::mbar_expect_tx_u32(full_u32 + stage * 8, w_bytes);
              ptx::tma_load_2d_u32(ring_u32 + stage * kPanelBytes, wmap, full_u32 + stage * 8, kp * 64,
                                   w_row + h * 128);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      if (p.dbg) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[0] = clock64() - c_start; d[1] = c_empty; d[2] = c_tile;
      }
    }
  } else if (warp == kMmaWarp) {
    // =============================== MMA issuer (one lane walks the step list) ===============================
    if (lane == 0) {
      constexpr uint32_t kDescHi = ptx::desc_hi_sw128(1024);
      int stage = 0; uint32_t phase = 0;
      uint32_t wait_phase = 0;    // bit i: parity to wait for on barrier i of {panel_ready, feat_ready}
      long long c_panel = 0, c_full = 0, c_issue = 0;
      const long long c_start = clock64();
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int i = 0; i < n_steps; ++i) {
          const uint4 e = sm.steps[i];
          const uint32_t wi = e.w & 31u;
          if (wi) {
            const uint32_t idx = wi - 1;
            const long long c0 = clock64();
            ptx::mbar_wait_u32(waitbar_u32 + idx * 8, (wait_phase >> idx) & 1u);
            c_panel += clock64() - c0;
            wait_phase ^= 1u << idx;
          }
          uint32_t a_addr = e.x, a_empty = 0;
          if (a_addr == 0) {
            ptx::mbar_wait_u32(full_u32 + stage * 8, phase);
            a_addr = ring_u32 + stage * kPanelBytes;
            a_empty = empty_u32 + stage * 8;
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          long long c1 = clock64();
          ptx::mbar_wait_u32(full_u32 + stage * 8, phase);
          long long c2 = clock64();
          c_full += c2 - c1;
          ptx::tc_fence_after();
          const uint64_t da = ptx::desc_from(kDescHi, a_addr);
          const uint32_t d_tmem = tmem_base + e.y;
          {
            // advancing K by 16 bf16 (32 bytes) inside the 128B swizzle atom = +2 in the address field
            const uint64_t db = ptx::desc_from(kDescHi, ring_u32 + stage * kPanelBytes);
            ptx::mma_bf16_ss(d_tmem, da, db, e.z, (e.w >> 5) & 1u);
            ptx::mma_bf16_ss(d_tmem, da + 2, db + 2, e.z, 1u);
            ptx::mma_bf16_ss(d_tmem, da + 4, db + 4, e.z, 1u);
            ptx::mma_bf16_ss(d_tmem, da + 6, db + 6, e.z, 1u);
            ptx::mma_commit_u32(empty_u32 + stage * 8);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          c_issue += clock64() - c2;
          if (e.w & 64u) {
            c1 = clock64();
            ptx::mbar_wait_u32(full_u32 + stage * 8, phase);
            c2 = clock64();
            c_full += c2 - c1;
            ptx::tc_fence_after();
            const uint64_t db = ptx::desc_from(kDescHi, ring_u32 + stage * kPanelBytes);
            ptx::mma_bf16_ss(d_tmem + 128, da, db, e.z, (e.w >> 5) & 1u);
            ptx::mma_bf16_ss(d_tmem + 128, da + 2, db + 2, e.z, 1u);
            ptx::mma_bf16_ss(d_tmem + 128, da + 4, db + 4, e.z, 1u);
            ptx::mma_bf16_ss(d_tmem + 128, da + 6, db + 6, e.z, 1u);
            ptx::mma_commit_u32(empty_u32 + stage * 8);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            c_issue += clock64() - c2;
          }
          if (a_empty) ptx::mma_commit_u32(a_empty);
          const uint32_t ci = (e.w >> 8) & 15u;
          if (ci) ptx::mma_commit_u32(accfull_u32 + (ci - 1) * 8);
        }
        ptx::mma_commit_u32(accfull_u32 + 7 * 8);   // tile_done: every MMA of this tile has completed
      }
      if (p.dbg) {
        long long* d = p.dbg + blockIdx.x * 16;
        d[4] = clock64() - c_start; d[5] = c_panel; d[6] = c_full; d[7] = c_issue;
      }
    }
  } else {
    // =============================== epilogue groups ===============================
    const int ew = warp;                   // 0..15 (the scheduler favours high warp ids: issuer warps come last)
    const int q = ew >> 2;                 // group: owns output columns [64q, 64q+64) of a 256-wide layer
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;   // tile row == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool group_leader = (ew & 3) == 0 && lane == 0;
    const int bar_id = 1 + q;
    uint32_t acc_phase = 0;
    int tile_iter = 0;
    float v[32];
    long long c_acc = 0, c_guard = 0, c_ld = 0, c_math = 0, c_pub = 0;
    const long long c_epi_start = clock64();

    // Publish a finished panel: optional TMA store (saved activations / dZ), then signal the MMA issuer.
    auto publish = [&](const TcLayer& L, int pi, int col, int tile) {
      ptx::fence_proxy_async();
      if (kTrain && L.save_row >= 0) {
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (group_leader) {
          ptx::tma_store_2d(&p.map_save, sm.panels + pi * kPanelBytes, col, L.save_row + tile * kTileM);
          ptx::tma_commit_group();
        }
      }
      ptx::tc_fence_before();
      if (!L.no_signal) ptx::mbar_arrive(&sm.panel_ready[pi]);
    };
    // The panel this group is about to overwrite was TMA-stored two layers ago: that read must be done.
    auto guard_panel = [&](const TcLayer& L) {
      if (kTrain && L.save_row >= 0) {
        const long long c0 = clock64();
        if (group_leader) ptx::tma_wait_group_read<1>();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        c_guard += clock64() - c0;
      }
    };
    auto wait_acc = [&](int bar) {
      const long long c0 = clock64();
      ptx::mbar_wait(&sm.acc_full[bar], (acc_phase >> bar) & 1u);
      c_acc += clock64() - c0;
      acc_phase ^= 1u << bar;
      ptx::tc_fence_after();
    };

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_iter) {
      const int s = tile * kTileM + row;   // global sample index
      const bool valid = s < p.n_samples;
      float raw_d = 0.f;
      for (int l = 0; l < p.n_layers; ++l) {
        const TcLayer& L = p.layers[l];
        if (!layer_has_mma(L) && tile_iter > 0) {
          // the start op of a backward tile overwrites panels the previous tile's last MMAs may still read
          ptx::mbar_wait(&sm.acc_full[7], (uint32_t)((tile_iter - 1) & 1));
        }
        switch (L.epi) {
          case EPI_RELU: case EPI_LINEAR: {
            const int col = q * 64, pi = L.dst_buf * 4 + q;
            uint8_t* panel = sm.panels + pi * kPanelBytes;
            guard_panel(L);
            wait_acc(L.acc_bar);
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
              const long long c0 = clock64();
              load_acc32(lane_addr + (uint32_t)(L.acc_col + col + hf * 32), v);
              const long long c1 = clock64();
              c_ld += c1 - c0;
              const float4* b4 = reinterpret_cast<const float4*>(sm.bias + L.bias_off + col + hf * 32);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 b = b4[c];
                v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
              }
              if (L.epi == EPI_RELU) store_half32<true>(panel, row, hf * 4, v);
              else store_half32<false>(panel, row, hf * 4, v);
              c_math += clock64() - c1;
            }
            const long long c2 = clock64();
            publish(L, pi, col, tile);
            c_pub += clock64() - c2;
            break;
          }
          case EPI_BWD_LINEAR: case EPI_BWD_RELU: case EPI_BWD_RELU_D: {
            const int col = q * 64, pi = L.dst_buf * 4 + q;
            uint8_t* panel = sm.panels + pi * kPanelBytes;
            // side inputs do not depend on the accumulator: fetch them before waiting for the MMAs
            uint4 mk[8];
            float dd = 0.f;
            if (L.epi != EPI_BWD_LINEAR) {
              if (valid) {
                const uint4* src = reinterpret_cast<const uint4*>(p.act + ((size_t)L.mask_row + s) * kW + col);
#pragma unroll
                for (int c = 0; c < 8; ++c) mk[c] = __ldg(src + c);
              } else {
#pragma unroll
                for (int c = 0; c < 8; ++c) mk[c] = make_uint4(0u, 0u, 0u, 0u);
              }
              if (L.epi == EPI_BWD_RELU_D)
                dd = valid ? __bfloat162float(__float2bfloat16(p.d_raw[(size_t)s * p.raw_c])) : 0.f;
            }
            guard_panel(L);
            wait_acc(L.acc_bar);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              load_acc32(lane_addr + (uint32_t)(L.acc_col + col + hf * 32), v);
              if (L.epi == EPI_BWD_RELU_D) {
                const float4* w4 = reinterpret_cast<const float4*>(sm.bias + p.w_dens_off + col + hf * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 w = w4[c];
                  v[c * 4 + 0] = fmaf(dd, w.x, v[c * 4 + 0]); v[c * 4 + 1] = fmaf(dd, w.y, v[c * 4 + 1]);
                  v[c * 4 + 2] = fmaf(dd, w.z, v[c * 4 + 2]); v[c * 4 + 3] = fmaf(dd, w.w, v[c * 4 + 3]);
                }
              }
              if (L.epi != EPI_BWD_LINEAR) {
                const uint4 (&half)[4] = *reinterpret_cast<const uint4 (*)[4]>(&mk[hf * 4]);
                apply_mask32(half, v);
              }
              store_half32<false>(panel, row, hf * 4, v);
            }
            publish(L, pi, col, tile);
            break;
          }
          case EPI_VIEW: {
            if (q >= 2) break;                      // N = 128: groups 0 and 1
            const int col = q * 64, pi = L.dst_buf * 4 + q;
            uint8_t* panel = sm.panels + pi * kPanelBytes;
            guard_panel(L);
            wait_acc(L.acc_bar);
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
              load_acc32(lane_addr + (uint32_t)(L.acc_col + col + hf * 32), v);
              if (valid) {
                const float4* b4 = reinterpret_cast<const float4*>(p.viewbias + (size_t)(s / p.S) * 128 + col + hf * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 b = __ldg(b4 + c);
                  v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
                }
              }
              store_half32<true>(panel, row, hf * 4, v);
            }
            publish(L, pi, col, tile);
            break;
          }
          case EPI_DENSITY: case EPI_RGB: {
            if (q != 0) break;
            wait_acc(L.acc_bar);
            uint32_t r4[4];
            ptx::tmem_ld4(lane_addr + (uint32_t)L.acc_col, r4);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            if (L.epi == EPI_DENSITY) {
              raw_d = __uint_as_float(r4[0]) + sm.bias[L.bias_off];
              if (p.raw_c == 1 && valid) p.raw_out[s] = raw_d;
            } else if (valid) {
              float4 o;
              o.x = raw_d;
              o.y = __uint_as_float(r4[0]) + sm.bias[L.bias_off + 0];
              o.z = __uint_as_float(r4[1]) + sm.bias[L.bias_off + 1];
              o.w = __uint_as_float(r4[2]) + sm.bias[L.bias_off + 2];
              reinterpret_cast<float4*>(p.raw_out)[s] = o;
            }
            break;
          }
          case EPI_BWD_START: {
            // dV = W_rgb^T d_rgb (CUDA cores), gated by the saved view activation; 128 columns: groups 0, 1
            if (q >= 2) break;
            const int col = q * 64, pi = L.dst_buf * 4 + q;
            uint8_t* panel = sm.panels + pi * kPanelBytes;
            const float4 dr = valid ? reinterpret_cast<const float4*>(p.d_raw)[s] : make_float4(0, 0, 0, 0);
            // bf16-round the head gradient once so that dgrad (here) and wgrad (tensor cores) agree
            const float d0 = __bfloat162float(__float2bfloat16(dr.y)), d1 = __bfloat162float(__float2bfloat16(dr.z)),
                        d2 = __bfloat162float(__float2bfloat16(dr.w));
            if (q == 0) {   // padding rows of the tile get zeros (dr == 0)
              uint4* dst = reinterpret_cast<uint4*>(p.drgb_out + (size_t)s * kHeadCols);
              dst[0] = make_uint4(ptx::pack_bf16x2(dr.y, dr.z), ptx::pack_bf16x2(dr.w, dr.x), 0u, 0u);
            }
            guard_panel(L);
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
              const float* wr = sm.bias + p.w_rgb_off + (col + hf * 32) * 3;
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = d0 * wr[c * 3] + d1 * wr[c * 3 + 1] + d2 * wr[c * 3 + 2];
              uint4 mk[4];
              if (valid) {
                const uint4* src = reinterpret_cast<const uint4*>(p.act + ((size_t)L.mask_row + s) * kW + col + hf * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) mk[c] = __ldg(src + c);
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) mk[c] = make_uint4(0u, 0u, 0u, 0u);
              }
              apply_mask32(mk, v);
              store_half32<false>(panel, row, hf * 4, v);
            }
            publish(L, pi, col, tile);
            break;
          }
          case EPI_BWD_START_PROP: {
            // dZ_last = d_raw_density * w_density gated by the last trunk activation; 256 columns: all groups
            const int col = q * 64, pi = L.dst_buf * 4 + q;
            uint8_t* panel = sm.panels + pi * kPanelBytes;
            const float dd0 = valid ? p.d_raw[s] : 0.f;
            const float dd = __bfloat162float(__float2bfloat16(dd0));
            if (q == 0) {
              uint4* dst = reinterpret_cast<uint4*>(p.drgb_out + (size_t)s * kHeadCols);
              dst[0] = make_uint4(0u, ptx::pack_bf16x2(0.f, dd0), 0u, 0u);
            }
            guard_panel(L);
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
              const float* wd = sm.bias + p.w_dens_off + col + hf * 32;
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = dd * wd[c];
              uint4 mk[4];
              if (valid) {
                const uint4* src = reinterpret_cast<const uint4*>(p.act + ((size_t)L.mask_row + s) * kW + col + hf * 32);
#pragma unroll
                for (int c = 0; c < 4; ++c) mk[c] = __ldg(src + c);
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) mk[c] = make_uint4(0u, 0u, 0u, 0u);
              }
              apply_mask32(mk, v);
              store_half32<false>(panel, row, hf * 4, v);
            }
            publish(L, pi, col, tile);
            break;
          }
          default: break;
        }
      }
      if (feat_resident && group_leader) {
        // the producer refills the panels with the next tile's features once no TMA store reads them
        if (kTrain) ptx::tma_wait_group_read<0>();
        ptx::mbar_arrive(sm.panels_free);
      }
    }
    if (group_leader) ptx::tma_wait_group<0>();
    if (p.dbg && (threadIdx.x == 0 || threadIdx.x == 3 * 128)) {
      long long* d = p.dbg + blockIdx.x * 16 + (threadIdx.x == 0 ? 8 : 12);
      d[0] = clock64() - c_epi_start; d[1] = c_acc; d[2] = c_guard;
      if (threadIdx.x == 0) { long long* e = p.dbg + blockIdx.x * 16; e[3] = c_ld; e[11] = c_math; e[15] = c_pub; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_map_impl(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return HUGS_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return HUGS_ERR_CUDA; }
  return HUGS_OK;
}

template <class T>
int tc_alloc(hugs_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return HUGS_ERR_NOMEM;
  }
  h->allocs.push_back(q);
  *p = static_cast<T*>(q);
  return HUGS_OK;
}

// Build packing tables and the forward/backward layer schedules of one MLP.
int build_mlp_schedule(hugs_handle* h, const MlpViews& mv, TcMlp* m, int save_layers_base) {
  (void)save_layers_base;
  const hugs_model_desc& d = h->d;
  m->present = true; m->has_rgb = mv.has_rgb; m->depth = mv.depth;
  int row_f = 0, row_b = 0, boff = 0;
  std::vector<int> layer_w_row(mv.dense.size()), layer_b_row(mv.dense.size(), -1), layer_bias(mv.dense.size());
  // ---- packing ----
  for (size_t li = 0; li < mv.dense.size(); ++li) {
    const DenseView& v = mv.dense[li];
    TcMlp::PackLayer P{};
    const bool trunk = (int)li < mv.depth;
    const bool is_density = (int)li == mv.depth;
    const bool is_view = mv.has_rgb && (int)li == mv.depth + 2;
    P.out = v.out; P.in = v.in; P.koff = v.kernel_off; P.boff = v.bias_off;
    if (is_view) { P.x_in = d.bottleneck_width; P.feat_in = 0; }
    else if (v.in == h->feat_dim) { P.x_in = 0; P.feat_in = h->feat_dim; }
    else if (v.in == kW + h->feat_dim) { P.x_in = kW; P.feat_in = h->feat_dim; }
    else { P.x_in = v.in; P.feat_in = 0; }
    P.rows_pad = v.out >= 128 ? ((v.out + 127) / 128) * 128 : 16;
    P.row0 = row_f; row_f += P.rows_pad;
    P.bias_off = boff; boff += ((v.out + 3) / 4) * 4;
    // natural-layout copy for dgrad: every layer whose *input* carries a gradient
    const bool needs_dgrad = (trunk && li > 0) || (mv.has_rgb && ((int)li == mv.depth + 1 || is_view));
    P.brow0 = -1;
    if (needs_dgrad) { P.brow0 = row_b; row_b += kW; }
    (void)is_density;
    layer_w_row[li] = P.row0; layer_b_row[li] = P.brow0; layer_bias[li] = P.bias_off;
    m->pack.push_back(P);
  }
  m->rows_f = row_f; m->rows_b = std::max(row_b, kW);
  // head weights for the CUDA-core parts of the backward chain
  const DenseView& dens = mv.dense[mv.depth];
  m->w_dens_off = boff; boff += ((dens.in + 3) / 4) * 4;
  m->w_rgb_off = -1;
  if (mv.has_rgb) { m->w_rgb_off = boff; boff += mv.dense[mv.depth + 3].in * 3 + 4; }
  m->bias_floats = boff;

  // ---- forward schedule ----
  auto trunk_layer = [&](int i, bool first, bool cat) {
    TcLayer L{};
    L.a_res = first ? kFeatPad / 64 : 4; L.a_str = (!first && cat) ? kFeatPad / 64 : 0;
    L.a_feat = first ? 1 : 0;
    L.a_buf = first ? 0 : i % 2; L.wait_panels = 1; L.n_halves = 2; L.n_mma = 128;
    L.acc_col = (i % 2) * 256; L.acc_bar = i % 2;
    L.w_row = layer_w_row[i]; L.w_map = 0; L.epi = EPI_RELU; L.dst_buf = (i + 1) % 2; L.bias_off = layer_bias[i];
    L.save_row = -1; L.mask_row = -1;
    return L;
  };
  bool cat = false;
  for (int i = 0; i < mv.depth; ++i) {
    m->fwd.push_back(trunk_layer(i, i == 0, cat));
    cat = (i % d.skip_layer == 0 && i > 0);
  }
  HUGS_REQUIRE(!cat, "tensor-core path: a skip connection into the heads is not supported (depth %d, skip %d)",
               mv.depth, d.skip_layer);
  const int D = mv.depth;
  const int head_buf = D % 2;             // buffer holding the last trunk activation
  if (!mv.has_rgb) {
    TcLayer L{};
    L.a_res = 4; L.a_buf = head_buf; L.wait_panels = 1; L.n_halves = 1; L.n_mma = 16;
    L.acc_col = (D % 2) * 256; L.acc_bar = 4; L.w_row = layer_w_row[D]; L.w_map = 1; L.epi = EPI_DENSITY;
    L.bias_off = layer_bias[D]; L.save_row = -1; L.mask_row = -1;
    m->fwd.push_back(L);
  } else {
    const int other = (D + 1) % 2;        // accumulator / buffer parity not used by the bottleneck
    TcLayer B{};                          // bottleneck (linear)
    B.a_res = 4; B.a_buf = head_buf; B.wait_panels = 1; B.n_halves = 2; B.n_mma = 128;
    B.acc_col = (D % 2) * 256; B.acc_bar = D % 2; B.w_row = layer_w_row[D + 1]; B.w_map = 0;
    B.epi = EPI_LINEAR; B.dst_buf = other; B.bias_off = layer_bias[D + 1]; B.save_row = -1; B.mask_row = -1;
    m->fwd.push_back(B);
    TcLayer Dn{};                         // density head reads the same activation (already waited for)
    Dn.a_res = 4; Dn.a_buf = head_buf; Dn.wait_panels = 0; Dn.n_halves = 1; Dn.n_mma = 16;
    Dn.acc_col = other * 256; Dn.acc_bar = 4; Dn.w_row = layer_w_row[D]; Dn.w_map = 1; Dn.epi = EPI_DENSITY;
    Dn.bias_off = layer_bias[D]; Dn.save_row = -1; Dn.mask_row = -1;
    m->fwd.push_back(Dn);
    TcLayer V{};                          // view layer: K = bottleneck (dir/GLO terms live in viewbias)
    V.a_res = 4; V.a_buf = other; V.wait_panels = 1; V.n_halves = 1; V.n_mma = 128;
    V.acc_col = other * 256 + 128; V.acc_bar = 5; V.w_row = layer_w_row[D + 2]; V.w_map = 0; V.epi = EPI_VIEW;
    V.dst_buf = head_buf; V.bias_off = 0; V.save_row = -1; V.mask_row = -1;
    m->fwd.push_back(V);
    TcLayer R{};                          // rgb head, K = 128 (2 panels)
    R.a_res = 2; R.a_buf = head_buf; R.wait_panels = 1; R.n_halves = 1; R.n_mma = 16;
    R.acc_col = other * 256 + 16; R.acc_bar = 6; R.w_row = layer_w_row[D + 3]; R.w_map = 1; R.epi = EPI_RGB;
    R.bias_off = layer_bias[D + 3]; R.save_row = -1; R.mask_row = -1;
    m->fwd.push_back(R);
    m->view_w_row = layer_w_row[D + 2];
  }
  HUGS_REQUIRE((int)m->fwd.size() <= kMaxLayers, "tensor-core path: too many layers (%zu)", m->fwd.size());

  // ---- backward (dgrad) schedule; save_row / mask_row hold *slot* indices, resolved per call ----
  // forward slots: j in [0,D) = output of trunk layer j, D = bottleneck output, D+1 = view activation.
  // dZ slots use the same indexing (gradient w.r.t. the pre-activation of that layer).
  auto mma_op = [&](int k, int a_res, int w_row, int epi, int save_slot, int mask_slot) {
    TcLayer L{};
    L.a_res = a_res; L.a_str = 0; L.a_buf = (k - 1) % 2; L.wait_panels = 1; L.n_halves = 2; L.n_mma = 128;
    L.acc_col = ((k - 1) % 2) * 256; L.acc_bar = (k - 1) % 2; L.w_row = w_row; L.w_map = 0; L.epi = epi;
    L.dst_buf = k % 2; L.bias_off = 0; L.save_row = save_slot; L.mask_row = mask_slot;
    return L;
  };
  int k = 0;
  if (mv.has_rgb) {
    TcLayer S0{};
    S0.epi = EPI_BWD_START; S0.dst_buf = 0; S0.save_row = D + 1; S0.mask_row = D + 1;
    m->bwd.push_back(S0); k = 1;
    m->bwd.push_back(mma_op(k++, 2, layer_b_row[D + 2], EPI_BWD_LINEAR, D, -1));        // through the view layer
    m->bwd.push_back(mma_op(k++, 4, layer_b_row[D + 1], EPI_BWD_RELU_D, D - 1, D - 1)); // through the bottleneck
  } else {
    TcLayer S0{};
    S0.epi = EPI_BWD_START_PROP; S0.dst_buf = 0; S0.save_row = D - 1; S0.mask_row = D - 1;
    m->bwd.push_back(S0); k = 1;
  }
  for (int l = D - 1; l >= 1; --l) m->bwd.push_back(mma_op(k++, 4, layer_b_row[l], EPI_BWD_RELU, l - 1, l - 1));
  m->bwd.back().no_signal = 1;   // dZ of the first layer feeds only the weight-gradient pass
  HUGS_REQUIRE((int)m->bwd.size() <= kMaxLayers, "tensor-core path: too many backward ops (%zu)", m->bwd.size());
  return pp_build(h, mv, m);
}

int fill_pack_args(hugs_handle* h, const MlpViews& mv, const TcMlp& m, const float* params, PackArgs* a) {
  memset(a, 0, sizeof(*a));
  HUGS_REQUIRE(m.pack.size() <= 16, "too many layers to pack");
  for (size_t i = 0; i < m.pack.size(); ++i) a->layers[i] = m.pack[i];
  a->n_layers = (int)m.pack.size(); a->rows_f = m.rows_f; a->rows_b = m.rows_b;
  a->nb = h->d.num_basis; a->ndeg = h->d.max_deg_point - h->d.min_deg_point; a->feat_dim = h->feat_dim;
  a->bias_floats = m.bias_floats; a->params = params; a->wt = m.wt; a->wn = m.wn; a->bias = m.bias;
  a->w_dens_off = m.w_dens_off; a->w_rgb_off = m.w_rgb_off;
  a->dens_koff = mv.dense[mv.depth].kernel_off; a->dens_in = mv.dense[mv.depth].in;
  if (mv.has_rgb) { a->rgb_koff = mv.dense[mv.depth + 3].kernel_off; a->rgb_in = mv.dense[mv.depth + 3].in; }
  // biased layers of the CTA-pair kernel: trunk layer i -> j = i, bottleneck -> j = depth (see pp_build)
  a->bias_img = m.bias_img;
  a->n_bias_layers = mv.depth + (mv.has_rgb ? 1 : 0);
  HUGS_REQUIRE(a->n_bias_layers <= 4 * kBiasChunks, "too many biased layers (%d)", a->n_bias_layers);
  for (int i = 0; i < mv.depth; ++i) a->bias_layer[i] = i;
  if (mv.has_rgb) a->bias_layer[mv.depth] = mv.depth + 1;
  return HUGS_OK;
}

}  // namespace

int make_map(CUtensorMap* m, const void* base, long long rows, long long cols, int box_rows) {
  return make_map_impl(m, base, rows, cols, box_rows);
}

// ------------------------------------------------------------------------------------------
int tc_create(hugs_handle* h) {
  const hugs_model_desc& d = h->d;
  const int ndeg = d.max_deg_point - d.min_deg_point;
  if (d.nerf_width != kW || (d.num_levels > 1 && d.prop_width != kW) || d.bottleneck_width != kW ||
      d.view_width != 128 || h->feat_dim > kFeatPad || (ndeg % 4) != 0 || ndeg > 16) {
    set_error("tensor-core path supports net_width 256, bottleneck 256, view width 128 and <= 512 IPE features "
              "with a degree count divisible by 4 (got widths %d/%d/%d/%d, %d features); use HUGS_PRECISION_FP32",
              d.nerf_width, d.prop_width, d.bottleneck_width, d.view_width, h->feat_dim);
    return HUGS_ERR_UNSUPPORTED;
  }
  cudaDeviceProp prop;
  HUGS_CUDA(cudaGetDeviceProperties(&prop, h->device));
  if (prop.major != 10) {
    set_error("tensor-core path needs an sm_100 device (found sm_%d%d)", prop.major, prop.minor);
    return HUGS_ERR_UNSUPPORTED;
  }
  TcState* tc = new TcState();
  h->tc = tc;
  tc->num_sms = prop.multiProcessorCount;
  int rc;
  if ((rc = build_mlp_schedule(h, h->nerf, &tc->nerf, 0))) return rc;
  if (d.num_levels > 1 && (rc = build_mlp_schedule(h, h->prop, &tc->prop, 0))) return rc;
  for (TcMlp* m : {&tc->nerf, &tc->prop}) {
    if (!m->present) continue;
    if ((rc = tc_alloc(h, &m->wt, (size_t)m->rows_f * kKP)) || (rc = tc_alloc(h, &m->wn, (size_t)m->rows_b * kW)) ||
        (rc = tc_alloc(h, &m->bias, (size_t)m->bias_floats)) ||
        (rc = tc_alloc(h, &m->bias_img, (size_t)2 * kBiasChunks * kBiasChunkElems)))
      return rc;
    if ((rc = make_map(&m->map_wt128, m->wt, m->rows_f, kKP, 128)) ||
        (rc = make_map(&m->map_wt16, m->wt, m->rows_f, kKP, 16)) ||
        (rc = make_map(&m->map_wn128, m->wn, m->rows_b, kW, 128)) ||
        (rc = make_map(&m->map_wt64, m->wt, m->rows_f, kKP, 64)))
      return rc;
  }
  // per-level feature / saved-activation regions
  const int L = d.num_levels;
  tc->cap.resize(L); tc->feat_row0.resize(L); tc->save_row0.resize(L);
  int frow = 0, srow = 0;
  for (int l = 0; l < L; ++l) {
    const int S = h->samples(l);
    tc->cap[l] = (int)((((long long)d.max_rays * S + kTileM - 1) / kTileM) * kTileM);
    tc->feat_row0[l] = frow; frow += tc->cap[l];
    const int n_saved = (l == L - 1) ? d.nerf_depth + 2 : d.prop_depth;
    tc->save_row0[l] = srow; srow += n_saved * tc->cap[l];
  }
  tc->total_feat_rows = frow; tc->total_save_rows = srow;
  if ((rc = tc_alloc(h, &tc->feat, (size_t)frow * kFeatPad))) return rc;
  for (int l = 0; l < L; ++l) tc->drgb_rows = std::max(tc->drgb_rows, tc->cap[l]);
  if ((rc = tc_alloc(h, &tc->viewbias, (size_t)d.max_rays * 128))) return rc;
  if ((rc = make_map(&tc->map_feat, tc->feat, frow, kFeatPad, 128))) return rc;
  HUGS_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  HUGS_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  {
    const char* e = getenv("HUGS_CHAIN");
    tc->use_pp = !(e && strcmp(e, "single") == 0);
    tc->use_cg2 = tc->use_pp && !(e && strcmp(e, "pp") == 0);
  }
  return pp_init(h);
}

// Saved activations / dZ / ReLU gates / head gradients (1,056 B per saved row: 10.6 KB per NeRF sample) and the
// weight-gradient state exist only on handles that train: allocated by the first hugs_loss_and_grad, so a render-only
// handle with a large max_rays (render_chunk_size) does not reserve tens of GB it never touches.
int tc_ensure_training(hugs_handle* h) {
  TcState* tc = h->tc;
  HUGS_REQUIRE(tc, "tensor-core state missing");
  if (tc->train_ready) return HUGS_OK;
  const size_t srow = (size_t)tc->total_save_rows;
  int rc;
  if ((rc = tc_alloc(h, &tc->act, srow * kW)) || (rc = tc_alloc(h, &tc->dz, srow * kW)) ||
      (rc = tc_alloc(h, &tc->gate, srow * 4)) || (rc = tc_alloc(h, &tc->drgb, (size_t)tc->drgb_rows * kHeadCols)))
    return rc;
  // unused columns (head gradients beyond col 3, view-activation columns 128..255) must read as zero
  HUGS_CUDA(cudaMemset(tc->gate, 0, srow * 4 * sizeof(uint2)));
  HUGS_CUDA(cudaMemset(tc->drgb, 0, (size_t)tc->drgb_rows * kHeadCols * 2));
  HUGS_CUDA(cudaMemset(tc->act, 0, srow * kW * 2));
  HUGS_CUDA(cudaMemset(tc->dz, 0, srow * kW * 2));
  HUGS_CUDA(cudaDeviceSynchronize());      // the caller's stream may be non-blocking w.r.t. the default stream
  if ((rc = make_map(&tc->map_act, tc->act, (long long)srow, kW, 128)) ||
      (rc = make_map(&tc->map_dz, tc->dz, (long long)srow, kW, 128)))
    return rc;
  if ((rc = wgrad_create(h))) return rc;
  tc->train_ready = true;
  return HUGS_OK;
}

void tc_destroy(hugs_handle* h) {
  if (h->tc) { wgrad_destroy(h); delete h->tc; h->tc = nullptr; }
}

int tc_pack_params(hugs_handle* h, const float* params, cudaStream_t st) {
  TcState* tc = h->tc;
  HUGS_REQUIRE(tc, "tensor-core state missing");
  PackArgs a;
  int rc;
  if ((rc = fill_pack_args(h, h->nerf, tc->nerf, params, &a))) return rc;
  pack_params_kernel<<<512, 256, 0, st>>>(a);
  HUGS_LAUNCH_CHECK();
  if (tc->prop.present) {
    if ((rc = fill_pack_args(h, h->prop, tc->prop, params, &a))) return rc;
    pack_params_kernel<<<512, 256, 0, st>>>(a);
    HUGS_LAUNCH_CHECK();
  }
  return HUGS_OK;
}

int tc_mlp_forward(hugs_handle* h, int level, const hugs_rays* rays, int n_rays, bool training, cudaStream_t st) {
  TcState* tc = h->tc;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const TcMlp& m = is_prop ? tc->prop : tc->nerf;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int S = h->samples(level);
  const int n_samples = n_rays * S;
  const int n_tiles = (n_samples + kTileM - 1) / kTileM;
  // 1. bf16 IPE features (own column order) -> feat[level]
  EncArgs ea{rays->origins, rays->directions, rays->radii, h->tdist[level], h->basis, n_samples, n_tiles * kTileM, S, d.num_basis,
             d.min_deg_point, d.max_deg_point - d.min_deg_point, d.ray_shape,
             is_prop ? d.prop_contract : d.nerf_contract, tc->feat + (size_t)tc->feat_row0[level] * kFeatPad};
  {
    ProfScope ps(h, HUGS_K_ENCODE, st);
    const int eg = (n_tiles * kTileM + kEncRows - 1) / kEncRows;
    if (ea.ndeg == 12) encode_bf16_kernel<12><<<eg, kEncSplit * kEncRows, 0, st>>>(ea);
    else encode_bf16_kernel<0><<<eg, kEncSplit * kEncRows, 0, st>>>(ea);
    HUGS_LAUNCH_CHECK();
  }
  // 2. per-ray view bias (direction encoding + GLO folded through the view layer)
  if (!is_prop) {
    const DenseView& vv = mv.dense[mv.depth + 2];
    viewbias_kernel<<<(n_rays * 128 + 255) / 256, 256, 0, st>>>(h->view_in, h->view_in_dim, h->cur_params,
                                                               vv.kernel_off, vv.bias_off, d.bottleneck_width,
                                                               128, n_rays, tc->viewbias);
    HUGS_LAUNCH_CHECK();
  }
  // 3. fused chain
  if (tc->use_pp) {
    ProfScope ps(h, is_prop ? HUGS_K_CHAIN_FWD_PROP : HUGS_K_CHAIN_FWD_NERF, st);
    return pp_launch(h, level, n_rays, training ? 1 : 0, st);
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.map_w128 = m.map_wt128; p.map_w16 = m.map_wt16; p.map_feat = tc->map_feat; p.map_save = tc->map_act;
  p.n_layers = (int)m.fwd.size();
  for (int i = 0; i < p.n_layers; ++i) {
    p.layers[i] = m.fwd[i];
    if (training) {
      // save every panel-producing layer's output: slot i of this level's region
      const int e = p.layers[i].epi;
      if (e == EPI_RELU || e == EPI_LINEAR || e == EPI_VIEW) {
        int slot = i;
        if (e == EPI_VIEW) slot = mv.depth + 1;       // after trunk (0..D-1) and bottleneck (D)
        if (e == EPI_LINEAR) slot = mv.depth;
        p.layers[i].save_row = tc->save_row0[level] + slot * tc->cap[level];
      }
    }
  }
  p.n_tiles = n_tiles; p.n_samples = n_samples; p.S = S; p.feat_row0 = tc->feat_row0[level];
  p.bias = m.bias; p.viewbias = tc->viewbias; p.raw_out = h->raw[level]; p.raw_c = is_prop ? 1 : 4;
  p.w_dens_off = m.w_dens_off; p.w_rgb_off = m.w_rgb_off; p.bias_floats = m.bias_floats;
  p.dbg = (!is_prop) ? h->dbg_counters : nullptr;
  const int grid = std::min(n_tiles, tc->num_sms);
  ProfScope ps(h, is_prop ? HUGS_K_CHAIN_FWD_PROP : HUGS_K_CHAIN_FWD_NERF, st);
  if (training) mlp_chain_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(p);
  else mlp_chain_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(p);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int tc_mlp_backward(hugs_handle* h, int level, const hugs_rays* rays, int n_rays, float* grad, cudaStream_t st) {
  (void)rays;
  TcState* tc = h->tc;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const TcMlp& m = is_prop ? tc->prop : tc->nerf;
  const int S = h->samples(level);
  const int n_samples = n_rays * S;
  const int n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int cap = tc->cap[level], srow = tc->save_row0[level];
  if (tc->use_pp) {
    {
      ProfScope ps(h, is_prop ? HUGS_K_CHAIN_BWD_PROP : HUGS_K_CHAIN_BWD_NERF, st);
      int rc = pp_launch(h, level, n_rays, 2, st);
      if (rc) return rc;
    }
    return wgrad_run(h, level, n_rays, grad, st);
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.map_w128 = m.map_wn128; p.map_w16 = m.map_wn128; p.map_feat = tc->map_feat; p.map_save = tc->map_dz;
  p.n_layers = (int)m.bwd.size();
  for (int i = 0; i < p.n_layers; ++i) {
    p.layers[i] = m.bwd[i];
    if (p.layers[i].save_row >= 0) p.layers[i].save_row = srow + p.layers[i].save_row * cap;
    if (p.layers[i].mask_row >= 0) p.layers[i].mask_row = srow + p.layers[i].mask_row * cap;
  }
  p.n_tiles = n_tiles; p.n_samples = n_samples; p.S = S; p.feat_row0 = tc->feat_row0[level];
  p.bias = m.bias; p.viewbias = tc->viewbias; p.raw_out = nullptr; p.raw_c = is_prop ? 1 : 4;
  p.d_raw = h->d_raw[level]; p.act = tc->act; p.drgb_out = tc->drgb;
  p.w_dens_off = m.w_dens_off; p.w_rgb_off = m.w_rgb_off; p.bias_floats = m.bias_floats;
  const int grid = std::min(n_tiles, tc->num_sms);
  {
    ProfScope ps(h, is_prop ? HUGS_K_CHAIN_BWD_PROP : HUGS_K_CHAIN_BWD_NERF, st);
    mlp_chain_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(p);
    HUGS_LAUNCH_CHECK();
  }
  return wgrad_run(h, level, n_rays, grad, st);
}

}  // namespace hugs
