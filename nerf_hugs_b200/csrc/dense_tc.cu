// Layer-at-a-time tensor-core Dense kernels for MLP widths the chain kernel (mlp_pp.cu) does not hold in shared memory:
// NerfMLP.net_width = 512 / 1024, i.e. every non-debug gin of the reference (MipNeRF360/configs/360.gin:14-16).
//
//   Y[M, N] = epilogue( [A0 | A1][M, K] . W[N, K]^T )          (bf16 x bf16 -> fp32 in TMEM)
//
// One persistent CTA pair (cluster of 2, tcgen05 cta_group::2) per SM pair walks output tiles of 256 rows (128 per CTA) x
// BN <= 256 columns, N fastest so that the pairs working on one row block re-read its A panels from L2.  Per K panel of
// 64: TMA loads this CTA's 128 A rows and its half of the BN weight rows into a 4-stage ring (completion on the leader's
// barrier), the leader issues 4 MMAs (M = 256, N = BN, K = 16), two 256-column accumulators alternate so that the
// epilogue of tile i runs under the MMAs of tile i + 1.  Epilogues: bias + ReLU / linear / per-ray view bias -> bf16 ->
// swizzled staging panels -> TMA store; fp32 head columns (raw density / raw rgb); backward: ReLU gate from the saved
// activation (+ the rank-1 density-head term) -> bf16 dZ.
//
// Reference semantics: models.py:437-519 (MLP.__call__) and its jax.value_and_grad (train_utils.py:454-455).
#include <algorithm>

#include "tc_device.cuh"
#include "dense_tc.h"

namespace hugs {
namespace {

constexpr int kDStages = 4;                // ring depth; 3 in the split-precision mode (the staging area doubles)
constexpr int kDStageBytes = 32768;        // A: 128 rows x 64 (16 KB) | B: <= 128 rows x 64 (16 KB)
constexpr int kDOutPanels = 4;             // staging: this CTA's 128 rows x 256 output columns (x 2: hi and lo, split mode)
constexpr int kDEpiWarps = 8;
constexpr int kDThreads = (2 + kDEpiWarps) * 32;
constexpr int kDSmem = 1024 + 3 * kDStageBytes + 2 * kDOutPanels * kPanelBytes + 256;      // >= the 4-stage / 4-panel layout
static_assert(kDSmem <= 232448 && kDSmem >= 1024 + kDStages * kDStageBytes + kDOutPanels * kPanelBytes + 256,
              "shared memory budget");

// 32 fp32 values of one row -> bf16 hi words into `hi_panel`, residual words into `lo_panel` (chunks chunk0 .. chunk0 + 3)
__device__ __forceinline__ void store_split_half32(uint8_t* hi_panel, uint8_t* lo_panel, int row, int chunk0, const float (&v)[32]) {
  uint4* ph = reinterpret_cast<uint4*>(hi_panel + row * 128);
  uint4* pl = reinterpret_cast<uint4*>(lo_panel + row * 128);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = v[c * 8 + 2 * j], b = v[c * 8 + 2 * j + 1];
      h[j] = ptx::pack_bf16x2(a, b);
      l[j] = ptx::pack_bf16x2(a - __uint_as_float(h[j] << 16), b - __uint_as_float(h[j] & 0xFFFF0000u));
    }
    ph[swz_chunk(row, chunk0 + c)] = make_uint4(h[0], h[1], h[2], h[3]);
    pl[swz_chunk(row, chunk0 + c)] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(kDThreads, 1) dense_tc_kernel(const __grid_constant__ DenseParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int n_stages = p.split ? 3 : kDStages;
  uint8_t* ring = base;
  uint8_t* stage_out = base + n_stages * kDStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_out + (p.split ? 2 : 1) * kDOutPanels * kPanelBytes);
  uint64_t* full = bars; uint64_t* empty = bars + kDStages;
  uint64_t* acc_full = bars + 2 * kDStages; uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  const uint32_t ring_u32 = ptx::smem_u32(ring);
  const uint32_t full_u32 = ptx::smem_u32(full), empty_u32 = ptx::smem_u32(empty);
  const uint32_t accfull_u32 = ptx::smem_u32(acc_full), accempty_u32 = ptx::smem_u32(acc_empty);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.a_map[0]); ptx::prefetch_tmap(&p.b_map); ptx::prefetch_tmap(&p.out_map);
    for (int i = 0; i < kDStages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 2 * kDEpiWarps); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc_cg2(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_tiles = p.m_tiles * p.n_tiles;
  int kp_total = 0;
  for (int seg = 0; seg < p.n_seg; ++seg) kp_total += p.a_kp[seg];

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = pair; t < n_tiles; t += n_pairs) {
        const int mt = t / p.n_tiles, nt = t % p.n_tiles;
        const int bn = p.tile_bn[nt], n0 = p.tile_n0[nt];
        const CUtensorMap* bmap = bn == 256 ? &p.b_map : (bn == 128 ? &p.b_map_64 : &p.b_map_8);
        const uint32_t bytes = 16384u + (uint32_t)(bn / 2) * 128u;
        const int row = mt * 256 + rank * 128;
        for (int seg = 0; seg < p.n_seg; ++seg) {
          for (int kp = 0; kp < p.a_kp[seg]; ++kp) {
            ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
            if (rank == 0) ptx::mbar_expect_tx_u32(full_u32 + stage * 8, 2 * bytes);
            const uint32_t bar = ptx::mapa_u32(full_u32 + stage * 8, 0);
            ptx::tma_load_2d_cg2(ring_u32 + stage * kDStageBytes, &p.a_map[seg], bar, p.a_col0[seg] + kp * 64,
                                 p.a_row0[seg] + row);
            ptx::tma_load_2d_cg2(ring_u32 + stage * kDStageBytes + 16384, bmap, bar, p.w_col0[seg] + kp * 64,
                                 p.b_row0 + p.w_row_off[seg] + n0 + rank * (bn / 2));
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t kDescHi = ptx::desc_hi_sw128(1024);
      int stage = 0; uint32_t phase = 0;
      uint32_t ae_phase = 0;     // bit s: parity of the next acc_empty phase of accumulator s
      int it = 0;
      for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
        const int nt = t % p.n_tiles;
        const int bn = p.tile_bn[nt];
        const uint32_t idesc = ptx::make_idesc_bf16(256, bn, 0, 0);
        const int as = it & 1;
        if (it >= 2) {           // the epilogue of the tile that used this accumulator two tiles ago has drained it
          ptx::mbar_wait_u32(accempty_u32 + as * 8, (ae_phase >> as) & 1u);
          ae_phase ^= 1u << as;
          ptx::tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int kp = 0; kp < kp_total; ++kp) {
          ptx::mbar_wait_u32(full_u32 + stage * 8, phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::desc_from(kDescHi, ring_u32 + stage * kDStageBytes);
          const uint64_t db = ptx::desc_from(kDescHi, ring_u32 + stage * kDStageBytes + 16384);
          ptx::mma_bf16_ss_cg2(d_tmem, da, db, idesc, kp > 0 ? 1u : 0u);
          ptx::mma_bf16_ss_cg2(d_tmem, da + 2, db + 2, idesc, 1u);
          ptx::mma_bf16_ss_cg2(d_tmem, da + 4, db + 4, idesc, 1u);
          ptx::mma_bf16_ss_cg2(d_tmem, da + 6, db + 6, idesc, 1u);
          ptx::mma_commit_mc2_u32(empty_u32 + stage * 8);
          if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_mc2_u32(accfull_u32 + as * 8);
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int ew = warp - 2;
    // a warp may only touch the TMEM lanes [32 * (warp % 4), +32): the lane quarter follows the hardware warp id
    const int quarter = warp & 3, half = ew >> 2;
    const int row = quarter * 32 + lane;                // row inside this CTA's 128-row block == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool store_leader = ew == 0 && lane == 0;
    uint32_t af_phase = 0;
    int it = 0;
    float v[32];
    // column sums (p.colsum): tile t = pair + it * n_pairs has N index t % n_tiles, which repeats with this period
    int cs_period = 1;
    { const int step = n_pairs % p.n_tiles; for (int x = step; x % p.n_tiles != 0; x += step) ++cs_period; }
    float cs0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cs1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      const int bn = p.tile_bn[nt], n0 = p.tile_n0[nt], epi = p.tile_epi[nt];
      const int as = it & 1;
      const int grow = mt * 256 + rank * 128 + row;     // row of the GEMM (sample index relative to the level)
      const bool valid = grow < p.m_rows;
      // ReLU gates of a backward tile: the saved activation comes from HBM (microseconds under load) and does not depend on
      // the accumulator, so all of this thread's gate words are fetched BEFORE the accumulator wait and kept as bit masks
      // (bit j / 16 + j = low / high half of packed word j is non-zero); padding rows carry no gradient
      uint32_t gate[4] = {0u, 0u, 0u, 0u};
      if (epi == DE_BWD_RELU && valid && p.gate_in) {
        // bit masks written by the forward pass: four words = this thread's 128 columns (bn == 256)
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.gate_in + ((size_t)p.gate_row0 + grow) * p.gate_ld +
                                                             ((n0 + half * (bn / 2)) >> 5)));
        gate[0] = q.x; gate[1] = q.y; gate[2] = q.z; gate[3] = q.w;
      } else if (epi == DE_BWD_RELU && valid) {
        const int cph = bn / 2;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i * 32 < cph) {
            const uint4* src = reinterpret_cast<const uint4*>(p.mask_act + ((size_t)p.mask_row0 + grow) * p.mask_ld + n0 +
                                                              half * cph + i * 32);
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 q = __ldg(src + c);
              pk[c * 4] = q.x; pk[c * 4 + 1] = q.y; pk[c * 4 + 2] = q.z; pk[c * 4 + 3] = q.w;
            }
            // to the column order of the bit-mask format: bit 2 j = low half of word j, bit 2 j + 1 = high half
            const uint32_t gb = gate_bits16(pk);
            uint32_t g = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) g |= (((gb >> j) & 1u) << (2 * j)) | (((gb >> (16 + j)) & 1u) << (2 * j + 1));
            gate[i] = g;
          }
        }
      }
      ptx::mbar_wait_u32(accfull_u32 + as * 8, (af_phase >> as) & 1u);
      af_phase ^= 1u << as;
      ptx::tc_fence_after();
      const uint32_t acc_addr = lane_addr + (uint32_t)(as * 256);
      if (epi == DE_HEAD_F32) {
        // fp32 head columns: raw_out[row * raw_c + raw_chan0 + c] = acc[c] + bias[n0 + c], c < raw_nchan
        if (half == 0) {
          uint32_t r4[4];
          ptx::tmem_ld4(acc_addr, r4);
          ptx::tmem_ld_wait();
          if (valid)
            for (int c = 0; c < p.raw_nchan; ++c)
              p.raw_out[(size_t)grow * p.raw_c + p.raw_chan0 + c] = __uint_as_float(r4[c]) + p.bias[n0 + c];
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(accempty_u32 + as * 8, 0));
        continue;
      }
      // the previous tile's TMA stores must have read the staging panels before they are rewritten
      if (store_leader) ptx::tma_wait_group_read<0>();
      asm volatile("bar.sync 1, %0;" ::"n"(kDEpiWarps * 32) : "memory");
      const int cols_per_half = bn / 2;                 // 128 | 64
      uint32_t gw[4] = {0u, 0u, 0u, 0u};                // ReLU gate words of this thread's columns (gate_out)
      for (int c0 = 0; c0 < cols_per_half; c0 += 32) {
        const int col = half * cols_per_half + c0;      // column inside the tile
        load_acc32(acc_addr + (uint32_t)col, v);
        const int n = n0 + col;                         // output column of the GEMM
        if (epi == DE_RELU || epi == DE_LINEAR) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 b = __ldg(b4 + c);
            v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
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
        } else if (epi == DE_BWD_RELU) {
          if (p.rank1_row) {
            float dd = valid ? p.rank1_row[(size_t)grow * p.rank1_stride] : 0.f;
            if (!p.exact_rank1) dd = __bfloat162float(__float2bfloat16(dd));
            const float4* w4 = reinterpret_cast<const float4*>(p.rank1_col + n);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 w = __ldg(w4 + c);
              v[c * 4 + 0] = fmaf(dd, w.x, v[c * 4 + 0]); v[c * 4 + 1] = fmaf(dd, w.y, v[c * 4 + 1]);
              v[c * 4 + 2] = fmaf(dd, w.z, v[c * 4 + 2]); v[c * 4 + 3] = fmaf(dd, w.w, v[c * 4 + 3]);
            }
          }
          const uint32_t g = c0 == 0 ? gate[0] : (c0 == 32 ? gate[1] : (c0 == 64 ? gate[2] : gate[3]));
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (((g >> k) & 1u) == 0u) v[k] = 0.f;
        } else if (epi == DE_BWD_LINEAR) {
          if (!valid) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = 0.f;
          }
        }
        if (p.gate_out && (epi == DE_RELU || epi == DE_VIEW)) {
          uint32_t g = 0u;
#pragma unroll
          for (int k = 0; k < 32; ++k) g |= (v[k] > 0.f ? 1u : 0u) << k;
          if (c0 == 0) gw[0] = g; else if (c0 == 32) gw[1] = g; else if (c0 == 64) gw[2] = g; else gw[3] = g;
        }
        uint8_t* panel = stage_out + (col >> 6) * kPanelBytes;
        const int chunk0 = (col & 63) >> 3;
        if (p.split) {
          if (epi == DE_RELU || epi == DE_VIEW) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
          }
          store_split_half32(panel, panel + kDOutPanels * kPanelBytes, row, chunk0, v);
        } else if (epi == DE_RELU || epi == DE_VIEW) store_half32<true>(panel, row, chunk0, v);
        else store_half32<false>(panel, row, chunk0, v);
      }
      if (p.gate_out && (epi == DE_RELU || epi == DE_VIEW) && valid && bn == 256)
        *reinterpret_cast<uint4*>(p.gate_out + ((size_t)p.gate_row0 + grow) * p.gate_ld + ((n0 + half * 128) >> 5)) =
            make_uint4(gw[0], gw[1], gw[2], gw[3]);
      // accumulator drained: the MMA issuer may reuse it (tile it + 2)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(accempty_u32 + as * 8, 0));
      ptx::fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(kDEpiWarps * 32) : "memory");
      if (store_leader) {
        for (int pn = 0; pn < bn / 64; ++pn) {
          ptx::tma_store_2d(&p.out_map, stage_out + pn * kPanelBytes, p.out_col0 + n0 + pn * 64,
                            p.out_row0 + mt * 256 + rank * 128);
          if (p.split)
            ptx::tma_store_2d(&p.out_map, stage_out + (kDOutPanels + pn) * kPanelBytes, p.out_col0 + n0 + pn * 64,
                              p.out_row0 + p.out_lo_row_off + mt * 256 + rank * 128);
        }
        ptx::tma_commit_group();
      }
      if (p.colsum) {
        // (after the TMA store has been issued: both only read the panels)
        // column sums of the staged tile: warp g reads rows g, g + 8, ... (16 x 16 bytes per thread: lane = 16-byte chunk of
        // the 256-column row, conflict-free), 8 partial sums per thread stay in registers across the tiles of this pair
        const int n_chunks = bn >> 3;                 // 16-byte chunks per row
        if (lane < n_chunks) {
          const uint8_t* pan = stage_out + (lane >> 3) * kPanelBytes;
          const uint32_t chunk = (uint32_t)lane & 7u;
          float a[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = 0.f;
          for (int part = 0; part < (p.split ? 2 : 1); ++part) {
            const uint8_t* q = pan + part * kDOutPanels * kPanelBytes;
            // the loads of a batch are issued back to back: a shared-memory load waits hundreds of cycles behind the tensor
            // core's operand reads, so the latency is paid once per batch, not once per row
#pragma unroll
            for (int batch = 0; batch < 2; ++batch) {
              uint4 w[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int r = ew + (batch * 8 + k) * kDEpiWarps;
                w[k] = *reinterpret_cast<const uint4*>(q + r * 128 + (swz_chunk(r, chunk) << 4));
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                a[0] += __uint_as_float(w[k].x << 16); a[1] += __uint_as_float(w[k].x & 0xFFFF0000u);
                a[2] += __uint_as_float(w[k].y << 16); a[3] += __uint_as_float(w[k].y & 0xFFFF0000u);
                a[4] += __uint_as_float(w[k].z << 16); a[5] += __uint_as_float(w[k].z & 0xFFFF0000u);
                a[6] += __uint_as_float(w[k].w << 16); a[7] += __uint_as_float(w[k].w & 0xFFFF0000u);
              }
            }
          }
          // a pair revisits the same N tile every cs_period tiles: one atomic per thread and column at the end of the kernel
          if (cs_period == 1 || (cs_period == 2 && (it & 1) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) cs0[j] += a[j];
          } else if (cs_period == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) cs1[j] += a[j];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(p.colsum + n0 + lane * 8 + j, a[j]);
          }
        }
      }
    }
    if (store_leader) ptx::tma_wait_group<0>();
    if (p.colsum && cs_period <= 2) {
      const int nt0 = pair % p.n_tiles, nt1 = (pair + n_pairs) % p.n_tiles;
      if (pair < n_tiles && lane * 8 < p.tile_bn[nt0]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(p.colsum + p.tile_n0[nt0] + lane * 8 + j, cs0[j]);
      }
      if (cs_period == 2 && pair + n_pairs < n_tiles && lane * 8 < p.tile_bn[nt1]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(p.colsum + p.tile_n0[nt1] + lane * 8 + j, cs1[j]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) ptx::tmem_dealloc_cg2(tmem_base, 512);
}

// dZ_view = (d_rgb . W_rgb^T) * [view activation > 0] (bf16), plus the head-gradient rows (d_r, d_g, d_b, d_density)
// of the head weight-gradient GEMMs.  CUDA cores: K = 3.
__global__ void __launch_bounds__(128) bwd_start_kernel(const float* d_raw, const __nv_bfloat16* view_act, int view_ld,
                                                        const float* w_rgb /* [128][3] fp32, bf16-rounded */, int n_samples,
                                                        int n_rows_pad, __nv_bfloat16* dz_view, int dz_ld,
                                                        __nv_bfloat16* drgb, __nv_bfloat16* dz_view_lo,
                                                        __nv_bfloat16* drgb_lo) {
  const int s = blockIdx.x, c = threadIdx.x;
  if (s >= n_rows_pad) return;
  const bool split = dz_view_lo != nullptr;
  float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s < n_samples) dr = reinterpret_cast<const float4*>(d_raw)[s];
  if (c == 0) {
    const uint32_t h01 = ptx::pack_bf16x2(dr.y, dr.z), h23 = ptx::pack_bf16x2(dr.w, dr.x);
    reinterpret_cast<uint4*>(drgb + (size_t)s * kHeadCols)[0] = make_uint4(h01, h23, 0u, 0u);
    if (split) {
      const uint32_t l01 = ptx::pack_bf16x2(dr.y - __uint_as_float(h01 << 16), dr.z - __uint_as_float(h01 & 0xFFFF0000u));
      const uint32_t l23 = ptx::pack_bf16x2(dr.w - __uint_as_float(h23 << 16), dr.x - __uint_as_float(h23 & 0xFFFF0000u));
      reinterpret_cast<uint4*>(drgb_lo + (size_t)s * kHeadCols)[0] = make_uint4(l01, l23, 0u, 0u);
    }
  }
  float d0 = dr.y, d1 = dr.z, d2 = dr.w;
  if (!split) {   // bf16-round the head gradient once so that dgrad (here) and wgrad (tensor cores) agree
    d0 = __bfloat162float(__float2bfloat16(d0)); d1 = __bfloat162float(__float2bfloat16(d1)); d2 = __bfloat162float(__float2bfloat16(d2));
  }
  float g = d0 * w_rgb[c * 3] + d1 * w_rgb[c * 3 + 1] + d2 * w_rgb[c * 3 + 2];
  const bool on = s < n_samples && __bfloat162float(view_act[(size_t)s * view_ld + c]) > 0.f;
  if (!on) g = 0.f;
  const __nv_bfloat16 hi = __float2bfloat16(g);
  dz_view[(size_t)s * dz_ld + c] = hi;
  if (split) dz_view_lo[(size_t)s * dz_ld + c] = __float2bfloat16(g - __bfloat162float(hi));
}

}  // namespace

int dense_tc_init() {
  HUGS_CUDA(cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDSmem));
  return HUGS_OK;
}

int dense_tc_launch(const DenseParams& p_in, int num_sms, cudaStream_t st) {
  DenseParams p = p_in;
  if (p.n_seg == 0) {     // the two-segment form: weight columns run on from b_col0
    p.n_seg = 2;
    p.w_col0[0] = p.b_col0; p.w_col0[1] = p.b_col0 + 64 * p.a_kp[0];
    p.w_row_off[0] = p.w_row_off[1] = 0;
  }
  HUGS_REQUIRE(p.n_seg >= 1 && p.n_seg <= kDMaxSegs, "dense_tc: bad segment count %d", p.n_seg);
  const int n_tiles = p.m_tiles * p.n_tiles;
  if (n_tiles <= 0) return HUGS_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * std::min(n_tiles, num_sms / 2));
  cfg.blockDim = dim3(kDThreads);
  cfg.dynamicSmemBytes = kDSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  HUGS_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_kernel, p));
  ++g_launch_count;
  return HUGS_OK;
}

int launch_bwd_start(const float* d_raw, const __nv_bfloat16* view_act, int view_ld, const float* w_rgb, int n_samples,
                     int n_rows_pad, __nv_bfloat16* dz_view, int dz_ld, __nv_bfloat16* drgb, __nv_bfloat16* dz_view_lo,
                     __nv_bfloat16* drgb_lo, cudaStream_t st) {
  if (n_rows_pad <= 0) return HUGS_OK;
  bwd_start_kernel<<<n_rows_pad, 128, 0, st>>>(d_raw, view_act, view_ld, w_rgb, n_samples, n_rows_pad, dz_view, dz_ld, drgb,
                                               dz_view_lo, drgb_lo);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs
