// Chain kernel of the nerfacto field (nerfacto.py:838-875): all five Dense layers of a 256-sample tile - or, in its backward
// program, the head-gradient start op and the four dgrad GEMMs of its backward pass - on one CTA pair without
// the activations (dZ) leaving the SMs between layers.
//
//   features [M, 64] --W_base0--> act0 (256, ReLU) --W_geo | w_density--> geo (64, linear) , raw density
//                    geo --W_head0 + per-ray bias--> h0 (256, ReLU) --W_head1--> h1 (256, ReLU) --W_rgb--> raw rgb
//
// The layer-at-a-time path (five dense_tc_kernel launches) is bound by HBM: every layer writes its output and the next launch
// reads it back (3.7 KB per sample).  Here a link's epilogue leaves the bf16 output in shared memory as the swizzled K-major
// panels the next link's MMAs read (the layout a TMA load would have produced), so only the hash features are read (128 B per
// sample) and - in training - the activations the backward pass needs are written once by TMA store (1.8 KB per sample, with 32 B
// of ReLU gate bits per gated layer).  Structure per CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256 = 128 rows per CTA):
//
//   shared memory  X[2 slots][4 panels]  a slot's activation, updated in place by the link epilogues          128 KB
//                  F[2 slots]            hash-feature panel of the slot's tile (TMA)                          32 KB
//                  ring[3]               weight stages, <= 128 rows x 64 K per CTA (TMA, half of the N rows)  48 KB
//   tensor memory  slot s accumulates in columns [256 s, 256 s + 256)
//   warps          0 TMA producer, 1 MMA issuer (leader CTA), 2..9 epilogue group of slot 0, 10..17 of slot 1
//
// Two tiles are in flight per pair and the issuer alternates between them link by link: one slot's epilogue runs under the
// other slot's MMAs (with a single tile in flight the chain is latency-bound: measured no faster than five launches).
// The table the epilogues read per 32-column chunk (biases, head weights) is staged in shared memory: an L1 / L2 round trip
// there is a serial ~500 cycles on the chain's critical path (measured: 11.53 -> 11.13 ms per step of config 4).
// Backward program (start_mode = 1): no feature load; the slot's epilogue group computes dZ of the last colour layer from d_raw,
// the rgb weights and h1's gate bits (K = 3, CUDA cores), then links run dZ . W^T with the ReLU gates from the forward's bit
// masks; every dZ the weight gradients need leaves by TMA store.
// bf16 mode only (the split-precision mode keeps the layer-at-a-time path: its hi + lo panels do not fit).
#include <algorithm>

#include "tc_device.cuh"
#include "field_chain.h"

namespace hugs {
namespace {

constexpr int kFcStages = 3;
constexpr int kFcStageBytes = 16384;
constexpr int kFcGroupWarps = 8;                       // epilogue warps per tile slot
constexpr int kFcThreads = (2 + 2 * kFcGroupWarps) * 32;
constexpr int kFcTabFloats = 1792;                      // bias / head-weight table staged in shared memory
constexpr int kFcSmem = 1024 + 8 * kPanelBytes + 2 * kPanelBytes + kFcStages * kFcStageBytes + 512 + kFcTabFloats * 4;
static_assert(kFcSmem <= 232448, "shared memory budget");

// Two tiles (slots) are in flight per CTA pair: slot s keeps its activation IN PLACE in X[s] (a link's epilogue starts only after
// all MMAs of the link have completed, so its output may overwrite its input), accumulates in TMEM columns [256 s, 256 s + 256)
// and has its own group of epilogue warps; the issuer alternates the slots link by link, so slot 0's epilogue runs under slot
// 1's MMAs and vice versa.
__global__ void __launch_bounds__(kFcThreads, 1) field_chain_kernel(const __grid_constant__ FieldChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* xbuf = base;                                   // X[2 slots][4 panels]
  uint8_t* fbuf = base + 8 * kPanelBytes;                 // F[2 slots]
  uint8_t* ring = fbuf + 2 * kPanelBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kFcStages * kFcStageBytes);
  uint64_t* full = bars;                    // [3] leader: both CTAs' weight bytes of a stage have landed
  uint64_t* empty = bars + 3;               // [3] per CTA: the MMAs that read the stage have completed
  uint64_t* f_full = bars + 6;              // [2] leader: both CTAs' feature panels of slot s have landed
  uint64_t* f_free = bars + 8;              // [2] per CTA: link 0's MMAs have read F[s]
  uint64_t* a_ready = bars + 10;            // [2] leader: both CTAs' epilogue groups of slot s are done with the link
  uint64_t* acc_full = bars + 12;           // [2] per CTA: the link's MMAs of slot s have completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
  // the table the epilogues read for every 32-column chunk: a shared-memory broadcast instead of an L1 / L2 round trip
  float* btab = reinterpret_cast<float*>(ring + kFcStages * kFcStageBytes + 512);
  for (int i = threadIdx.x; i < p.n_bias; i += kFcThreads) btab[i] = p.bias[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  const int n_units = (p.m_tiles + 1) / 2;                // unit u = tiles 2 u (slot 0) and 2 u + 1 (slot 1)
  const uint32_t x_u32 = ptx::smem_u32(xbuf), f_u32 = ptx::smem_u32(fbuf), ring_u32 = ptx::smem_u32(ring);
  const uint32_t full_u32 = ptx::smem_u32(full), empty_u32 = ptx::smem_u32(empty);
  const uint32_t ffull_u32 = ptx::smem_u32(f_full), ffree_u32 = ptx::smem_u32(f_free);
  const uint32_t aready_u32 = ptx::smem_u32(a_ready), accfull_u32 = ptx::smem_u32(acc_full);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.a_map); ptx::prefetch_tmap(&p.b_map); ptx::prefetch_tmap(&p.b_map_64); ptx::prefetch_tmap(&p.b_map_8);
    for (int i = 0; i < kFcStages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&f_full[i], 1); ptx::mbar_init(&f_free[i], 1);
      ptx::mbar_init(&a_ready[i], 2); ptx::mbar_init(&acc_full[i], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_cg2(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      uint32_t ff_phase = 0;     // bit s: parity of the next f_free phase of slot s
      auto load_features = [&](int u, int k) {     // k-th unit of this pair
        for (int sl = 0; sl < 2; ++sl) {
          const int t = 2 * u + sl;
          if (t >= p.m_tiles) break;
          if (k >= 1) { ptx::mbar_wait_u32(ffree_u32 + sl * 8, (ff_phase >> sl) & 1u); ff_phase ^= 1u << sl; }
          if (rank == 0) ptx::mbar_expect_tx_u32(ffull_u32 + sl * 8, 2 * kPanelBytes);
          ptx::tma_load_2d_cg2(f_u32 + sl * kPanelBytes, &p.a_map, ptx::mapa_u32(ffull_u32 + sl * 8, 0), 0, t * 256 + rank * 128);
        }
      };
      int k = 0;
      if (pair < n_units && !p.start_mode) load_features(pair, 0);
      for (int u = pair; u < n_units; u += n_pairs, ++k) {
        const int n_slots = (2 * u + 1 < p.m_tiles) ? 2 : 1;
        for (int l = 0; l < p.n_links; ++l) {
          const FieldChainLink& L = p.link[l];
          for (int sl = 0; sl < n_slots; ++sl) {
            for (int nt = 0; nt < L.n_tiles; ++nt) {
              const int bn = L.tile_bn[nt], n0 = L.tile_n0[nt];
              const CUtensorMap* bmap = bn == 256 ? &p.b_map : (bn == 128 ? &p.b_map_64 : &p.b_map_8);
              const uint32_t bytes = (uint32_t)(bn / 2) * 128u;
              for (int kp = 0; kp < L.kp; ++kp) {
                ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
                if (rank == 0) ptx::mbar_expect_tx_u32(full_u32 + stage * 8, 2 * bytes);
                ptx::tma_load_2d_cg2(ring_u32 + stage * kFcStageBytes, bmap, ptx::mapa_u32(full_u32 + stage * 8, 0), kp * 64,
                                     L.b_row0 + n0 + rank * (bn / 2));
                if (++stage == kFcStages) { stage = 0; phase ^= 1; }
              }
            }
          }
          // the next unit's features travel under this unit's remaining links (link 0 has released F by then)
          if (l == 1 && u + n_pairs < n_units && !p.start_mode) load_features(u + n_pairs, k + 1);
        }
        if (p.n_links < 2 && u + n_pairs < n_units && !p.start_mode) load_features(u + n_pairs, k + 1);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t kDescHi = ptx::desc_hi_sw128(1024);
      int stage = 0; uint32_t phase = 0;
      uint32_t ff_par = 0, ar_par = 0;
      bool first = true;
      for (int u = pair; u < n_units; u += n_pairs) {
        const int n_slots = (2 * u + 1 < p.m_tiles) ? 2 : 1;
        for (int l = 0; l < p.n_links; ++l) {
          const FieldChainLink& L = p.link[l];
          for (int sl = 0; sl < n_slots; ++sl) {
            // the slot's epilogue group is done with the previous link (its accumulator is drained, its panels are staged)
            if (!(first && l == 0) || p.start_mode) {
              ptx::mbar_wait_u32(aready_u32 + sl * 8, (ar_par >> sl) & 1u);
              ar_par ^= 1u << sl;
            }
            uint32_t a_base = x_u32 + sl * 4 * kPanelBytes;
            if (l == 0 && !p.start_mode) {
              ptx::mbar_wait_u32(ffull_u32 + sl * 8, (ff_par >> sl) & 1u);
              ff_par ^= 1u << sl;
              a_base = f_u32 + sl * kPanelBytes;
            }
            ptx::tc_fence_after();
            int tcol = 0;
            for (int nt = 0; nt < L.n_tiles; ++nt) {
              const int bn = L.tile_bn[nt];
              const uint32_t idesc = ptx::make_idesc_bf16(256, bn, 0, 0);
              const uint32_t d_tmem = tmem_base + (uint32_t)(sl * 256 + tcol);
              for (int kp = 0; kp < L.kp; ++kp) {
                ptx::mbar_wait_u32(full_u32 + stage * 8, phase);
                ptx::tc_fence_after();
                const uint64_t da = ptx::desc_from(kDescHi, a_base + kp * kPanelBytes);
                const uint64_t db = ptx::desc_from(kDescHi, ring_u32 + stage * kFcStageBytes);
                ptx::mma_bf16_ss_cg2(d_tmem, da, db, idesc, kp > 0 ? 1u : 0u);
                ptx::mma_bf16_ss_cg2(d_tmem, da + 2, db + 2, idesc, 1u);
                ptx::mma_bf16_ss_cg2(d_tmem, da + 4, db + 4, idesc, 1u);
                ptx::mma_bf16_ss_cg2(d_tmem, da + 6, db + 6, idesc, 1u);
                ptx::mma_commit_mc2_u32(empty_u32 + stage * 8);
                if (++stage == kFcStages) { stage = 0; phase ^= 1; }
              }
              tcol += bn;
            }
            ptx::mma_commit_mc2_u32(accfull_u32 + sl * 8);
            if (l == 0 && !p.start_mode) ptx::mma_commit_mc2_u32(ffree_u32 + sl * 8);   // F[slot] has been read once these MMAs complete
          }
        }
        first = false;
      }
    }
  } else {
    // =============================== epilogue groups: group g owns tile slot g ===============================
    const int sl = (warp - 2) / kFcGroupWarps;
    const int ew = (warp - 2) % kFcGroupWarps;
    const int quarter = warp & 3, half = ew >> 2;       // TMEM lane quarter follows the hardware warp id
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(sl * 256);
    const bool leader = ew == 0 && lane == 0;
    uint8_t* xs = xbuf + sl * 4 * kPanelBytes;
    uint32_t af_phase = 0;
    float v[32];
    auto group_sync = [&]() {
      if (sl == 0) asm volatile("bar.sync 1, %0;" ::"n"(kFcGroupWarps * 32) : "memory");
      else asm volatile("bar.sync 2, %0;" ::"n"(kFcGroupWarps * 32) : "memory");
    };
    for (int u = pair; u < n_units; u += n_pairs) {
      const int t = 2 * u + sl;
      if (t >= p.m_tiles) break;                         // (only the last unit can lack its second tile)
      const int grow = t * 256 + rank * 128 + row;
      const bool valid = grow < p.m_rows;
      float d_dens = 0.f;
      if (p.start_mode) {
        // ---- backward start: dZ of the last colour layer from the head gradients (CUDA cores: K = 3) ----
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) dr = __ldg(reinterpret_cast<const float4*>(p.d_raw) + grow);
        if (valid && !p.inside[grow]) dr.x = 0.f;         // density * selector: no density gradient out of range
        d_dens = dr.x;
        const uint32_t h01 = ptx::pack_bf16x2(dr.y, dr.z), h23 = ptx::pack_bf16x2(dr.w, dr.x);
        if (half == 0) reinterpret_cast<uint4*>(p.dh_out + (size_t)grow * kHeadCols)[0] = make_uint4(h01, h23, 0u, 0u);   // (zeros on padding rows)
        // bf16-round the head gradient once so that dgrad (here) and wgrad (tensor cores, from dh_out) agree
        const float d0 = __uint_as_float(h01 << 16), d1 = __uint_as_float(h01 & 0xFFFF0000u), d2 = __uint_as_float(h23 << 16);
        uint4 gq = make_uint4(0u, 0u, 0u, 0u);
        if (valid) gq = __ldg(reinterpret_cast<const uint4*>(p.gate_in + ((size_t)p.start_gate_row0 + grow) * p.gate_ld + half * 4));
        // the previous unit's last store read the panels this op overwrites
        if (leader) ptx::tma_wait_group_read<0>();
        group_sync();
        for (int c0 = 0; c0 < 128; c0 += 32) {
          const int col = half * 128 + c0;
          const uint32_t g = c0 == 0 ? gq.x : (c0 == 32 ? gq.y : (c0 == 64 ? gq.z : gq.w));
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float* w = btab + p.w_rgb_off + (col + k) * 3;
            const float z = d0 * w[0] + d1 * w[1] + d2 * w[2];
            v[k] = ((g >> k) & 1u) ? z : 0.f;
          }
          store_half32<false>(xs + (col >> 6) * kPanelBytes, row, (col & 63) >> 3, v);
        }
        ptx::fence_proxy_async();
        group_sync();
        if (leader) {
          ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(aready_u32 + sl * 8, 0));
          for (int pn = 0; pn < 4; ++pn) ptx::tma_store_2d(&p.start_map, xs + pn * kPanelBytes, pn * 64, t * 256 + rank * 128);
          ptx::tma_commit_group();
        }
      }
      for (int l = 0; l < p.n_links; ++l) {
        const FieldChainLink& L = p.link[l];
        ptx::mbar_wait_u32(accfull_u32 + sl * 8, af_phase);
        af_phase ^= 1u;
        ptx::tc_fence_after();
        // the TMA store of the previous link read the panels this epilogue overwrites in place
        if (leader) ptx::tma_wait_group_read<0>();
        group_sync();
        int out_panels = 0, tcol = 0;
        for (int nt = 0; nt < L.n_tiles; ++nt) {
          const int bn = L.tile_bn[nt], n0 = L.tile_n0[nt], epi = L.tile_epi[nt];
          const uint32_t acc_addr = lane_addr + (uint32_t)tcol;
          tcol += bn;
          if (epi == DE_HEAD_F32) {
            if (half == 0) {
              uint32_t r4[4];
              ptx::tmem_ld4(acc_addr, r4);
              ptx::tmem_ld_wait();
              if (valid)
                for (int c = 0; c < L.raw_nchan; ++c)
                  p.raw_out[(size_t)grow * p.raw_c + L.raw_chan0 + c] = __uint_as_float(r4[c]) + btab[L.bias_off + n0 + c];
            }
            continue;
          }
          out_panels = bn / 64;
          const int cols_per_half = bn / 2;             // 128 | 64
          uint32_t gw[4] = {0u, 0u, 0u, 0u};
          for (int c0 = 0; c0 < cols_per_half; c0 += 32) {
            const int col = half * cols_per_half + c0;
            load_acc32(acc_addr + (uint32_t)col, v);
            const int n = n0 + col;
            if (epi == DE_BWD_RELU || epi == DE_BWD_LINEAR) {
              if (epi == DE_BWD_RELU) {
                if (L.rank1) {   // density head: d_density (bf16-rounded like the head-gradient rows) x w_density
                  const float dd = __bfloat162float(__float2bfloat16(d_dens));
                  const float4* w4 = reinterpret_cast<const float4*>(btab + p.rank1_off + n);
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 w = w4[c];
                    v[c * 4 + 0] = fmaf(dd, w.x, v[c * 4 + 0]); v[c * 4 + 1] = fmaf(dd, w.y, v[c * 4 + 1]);
                    v[c * 4 + 2] = fmaf(dd, w.z, v[c * 4 + 2]); v[c * 4 + 3] = fmaf(dd, w.w, v[c * 4 + 3]);
                  }
                }
                const uint32_t g = valid ? __ldg(p.gate_in + ((size_t)L.gate_in_row0 + grow) * p.gate_ld + (n >> 5)) : 0u;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                  if (((g >> k) & 1u) == 0u) v[k] = 0.f;
              } else if (!valid) {
#pragma unroll
                for (int k = 0; k < 32; ++k) v[k] = 0.f;
              }
            } else if (epi == DE_VIEW) {
              if (valid) {
                const float4* b4 = reinterpret_cast<const float4*>(p.viewbias + (size_t)(grow / p.S) * p.view_ld + n);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 b = __ldg(b4 + c);
                  v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
                }
              }
            } else {
              const float4* b4 = reinterpret_cast<const float4*>(btab + L.bias_off + n);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 b = b4[c];
                v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
              }
            }
            if (p.gate_out && L.gate_row0 >= 0) {
              uint32_t g = 0u;
#pragma unroll
              for (int k = 0; k < 32; ++k) g |= (v[k] > 0.f ? 1u : 0u) << k;
              if (c0 == 0) gw[0] = g; else if (c0 == 32) gw[1] = g; else if (c0 == 64) gw[2] = g; else gw[3] = g;
            }
            uint8_t* panel = xs + (col >> 6) * kPanelBytes;
            const int chunk0 = (col & 63) >> 3;
            if (epi == DE_RELU || epi == DE_VIEW) store_half32<true>(panel, row, chunk0, v);
            else store_half32<false>(panel, row, chunk0, v);
          }
          if (p.gate_out && L.gate_row0 >= 0 && valid && bn == 256)
            *reinterpret_cast<uint4*>(p.gate_out + ((size_t)L.gate_row0 + grow) * p.gate_ld + ((n0 + half * 128) >> 5)) =
                make_uint4(gw[0], gw[1], gw[2], gw[3]);
        }
        // accumulator drained and output panels staged: the next link's MMAs (async proxy) may start; in training the panels
        // are also the saved activation
        ptx::tc_fence_before();
        ptx::fence_proxy_async();
        group_sync();
        if (leader) {
          // (backward program: the next unit's start op signals the issuer instead of the last link)
          if (!p.start_mode || l + 1 < p.n_links) ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(aready_u32 + sl * 8, 0));
          if (L.store && out_panels > 0) {
            for (int pn = 0; pn < out_panels; ++pn)
              ptx::tma_store_2d(&p.out_map[l], xs + pn * kPanelBytes, pn * 64, t * 256 + rank * 128);
            ptx::tma_commit_group();
          }
        }
      }
    }
    if (leader) ptx::tma_wait_group<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc_cg2(tmem_base, 512);
}

}  // namespace

int field_chain_init() {
  HUGS_CUDA(cudaFuncSetAttribute(field_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFcSmem));
  return HUGS_OK;
}

int field_chain_launch(const FieldChainParams& p, int num_sms, cudaStream_t st) {
  if (p.m_tiles <= 0) return HUGS_OK;
  HUGS_REQUIRE(p.n_links >= 1 && p.n_links <= kFcMaxLinks, "field chain: bad link count %d", p.n_links);
  HUGS_REQUIRE(p.bias && p.n_bias >= 0 && p.n_bias <= kFcTabFloats, "field chain: bad table size %d", p.n_bias);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * std::min(p.m_tiles, num_sms / 2));
  cfg.blockDim = dim3(kFcThreads);
  cfg.dynamicSmemBytes = kFcSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  HUGS_CUDA(cudaLaunchKernelEx(&cfg, field_chain_kernel, p));
  ++g_launch_count;
  return HUGS_OK;
}

}  // namespace hugs
