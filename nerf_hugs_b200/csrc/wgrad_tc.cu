// Weight / bias gradients of the tensor-core path.
//
//   dW_l[k_in, n_out] = sum_s A_l[s, k_in] * dZ_l[s, n_out]
//
// A_l (saved forward activations / IPE features) and dZ_l (saved by the backward chain) are
// row-major [samples, columns] bf16 tensors in HBM, i.e. *MN-major* operands for a GEMM whose
// reduction dimension is the sample index.  TMA loads 64-sample x 64-column boxes (128B swizzle),
// tcgen05.mma consumes them through MN-major shared-memory descriptors, and a 256 x N fp32
// accumulator (two M=128 blocks, all 512 TMEM columns for N = 256) lives in TMEM for the whole
// sample range of a work item; it is flushed once with red.global.add.f32 into the flat gradient.
//
// The remaining tiny reductions (biases = column sums of dZ, the N<=3 head kernels, the per-ray
// view-direction / GLO inputs of the view layer) run on CUDA cores.
//
// Split-precision mode (HUGS_PRECISION_TC_SPLIT): A and dZ are stored as hi + lo bf16 halves (lo `lo_rows` further down in
// the same tensors); every work item is issued four times (A_hi/A_lo x dZ_hi/dZ_lo) through the unchanged kernel and the
// four fp32 partial products meet in the gradient buffer.
//
// Reference semantics: jax.value_and_grad of train_utils.py:407-455 restricted to the Dense layers of
// models.py:449-519.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"
#include "tc_internal.h"

namespace hugs {

enum { WG_MAP_ACT = 0, WG_MAP_FEAT = 1, WG_MAP_DZ = 2, WG_MAP_DH = 3 };   // chain path; the layered path adds its own

struct WgState {
  CUtensorMap map_act64, map_feat64, map_dz64, map_dh64;
  std::vector<std::vector<WgItem>> host;   // per level
  std::vector<WgItem*> dev;                // per level
  std::vector<int> built_for;              // n_samples the list was built for
  float* dzv_ray = nullptr;                // [max_rays, 128]
};

namespace {

constexpr int kMaxWgItems = 4096;
constexpr int kWgStages = 3;
constexpr int kWgStageBytes = 65536;       // A0 16K | A1 16K | B 32K
constexpr int kWgThreads = 192;
constexpr int kWgScratchFloats = 4 * 32 * 33;   // per epilogue warp: 32 x 32 accumulator block (row stride 33) for the transposed flush
constexpr int kWgSmem = 1024 + kWgStages * kWgStageBytes + 256 + kWgScratchFloats * 4;

struct alignas(64) WgParams {
  CUtensorMap maps[kWgMaxMaps];
  const WgItem* items;
  int n_items, nb, ndeg, feat_dim;
  float* grad;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic on the __shared__ symbol (not an integer round trip) keeps the shared address space, so the
  // accesses below compile to LDS/STS instead of generic loads and stores
  uint8_t* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kWgStages * kWgStageBytes);
  uint64_t* full = bars; uint64_t* empty = bars + kWgStages;
  uint64_t* acc_full = bars + 2 * kWgStages; uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* scratch = reinterpret_cast<float*>(base + kWgStages * kWgStageBytes + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 4 && lane == 0) {
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.maps[i]);
    // a stage is released by the MMA commit and by the four epilogue warps (bias column sums read B)
    for (int i = 0; i < kWgStages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 5); }
    ptx::mbar_init(acc_full, 1); ptx::mbar_init(acc_empty, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 5) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const WgItem w = p.items[it];
        const CUtensorMap* amap = &p.maps[w.a_map];
        const int nb_atoms = w.n >> 6;
        for (int st = w.st0; st < w.st1; ++st) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* s = base + stage * kWgStageBytes;
          ptx::mbar_expect_tx(&full[stage], (4 + nb_atoms) * 8192);
          for (int a = 0; a < 4; ++a)
            ptx::tma_load_2d(s + a * 8192, amap, &full[stage], w.a_col0 + a * 64, w.a_row0 + st * 64);
          const CUtensorMap* bmap = &p.maps[w.b_map];
          for (int a = 0; a < nb_atoms; ++a)
            ptx::tma_load_2d(s + 32768 + a * 8192, bmap, &full[stage], w.b_col0 + a * 64, w.b_row0 + st * 64);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, ae_phase = 0;
      bool first_item = true;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const WgItem w = p.items[it];
        const uint32_t idesc = ptx::make_idesc_bf16(128, w.n, 1, 1);
        if (!first_item) { ptx::mbar_wait(acc_empty, ae_phase); ae_phase ^= 1; }
        first_item = false;
        ptx::tc_fence_after();
        for (int st = w.st0; st < w.st1; ++st) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t s = ptx::smem_u32(base + stage * kWgStageBytes);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            const uint64_t db = ptx::make_desc_sw128(s + 32768 + k16 * 2048, 8192, 1024);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
              const uint64_t da = ptx::make_desc_sw128(s + mb * 16384 + k16 * 2048, 8192, 1024);
              ptx::mma_bf16_ss(tmem_base + mb * 256, da, db, idesc, (st > w.st0 || k16 > 0) ? 1u : 0u);
            }
          }
          ptx::mma_commit(&empty[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit(acc_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int t = warp * 32 + lane;                // 0..127: owns B columns 2t, 2t+1 for the bias sums
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t af_phase = 0;
    int stage = 0; uint32_t phase = 0;
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const WgItem w = p.items[it];
      // ---- main loop: bias gradients = column sums of the dZ stage while the tensor core consumes it ----
      float s0 = 0.f, s1 = 0.f;
      const int cc = (2 * t) & 63, atom = (2 * t) >> 6;
      const bool sum_cols = w.bias_mode != 0 && 2 * t < w.n;
      for (int st = w.st0; st < w.st1; ++st) {
        if (w.bias_mode == 0) {
          // nothing to read from this stage: one lane observes `full` and releases the warp's share
          if (lane == 0) { ptx::mbar_wait(&full[stage], phase); ptx::mbar_arrive(&empty[stage]); }
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          continue;
        }
        if (lane == 0) ptx::mbar_wait(&full[stage], phase);
        __syncwarp();
        if (sum_cols) {
          const uint8_t* bs = base + stage * kWgStageBytes + 32768 + atom * 8192 + (cc & 7) * 2;
#pragma unroll 8
          for (int k = 0; k < 64; ++k) {
            const uint32_t pr = *reinterpret_cast<const uint32_t*>(bs + k * 128 + (((cc >> 3) ^ (k & 7)) << 4));
            s0 += __uint_as_float(pr << 16);
            s1 += __uint_as_float(pr & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      if (sum_cols && w.st1 > w.st0) {
        const int c = 2 * t;
        if (w.bias_mode == 1) { atomicAdd(p.grad + w.boff + w.b_col0 + c, s0); atomicAdd(p.grad + w.boff + w.b_col0 + c + 1, s1); }
        else if (w.bias_mode == 2) { if (c == 2) atomicAdd(p.grad + w.boff, s1); }
        else if (w.bias_mode == 3) {
          if (c == 0) { atomicAdd(p.grad + w.boff, s0); atomicAdd(p.grad + w.boff + 1, s1); }
          if (c == 2) atomicAdd(p.grad + w.boff + 2, s0);
        }
      }
      // ---- flush the accumulators ----
      ptx::mbar_wait(acc_full, af_phase); af_phase ^= 1;
      ptx::tc_fence_after();
      if (w.st1 > w.st0) {
#pragma unroll 1
        for (int mb = 0; mb < 2; ++mb) {
          const int m = mb * 128 + row;
          if (w.flush_mode == 0) {
            int krow;
            if (w.feat_mode) {
              const int fp = w.a_col0 + m;
              krow = fp < p.feat_dim ? w.in_base + ref_feature_col(fp, p.nb, p.ndeg) : -1;
            } else {
              krow = (w.in_rows > 0 && m >= w.in_rows) ? -1 : w.in_base + m;
            }
            // thread = accumulator row, but the gradient rows are `out` floats apart: transpose each 32 x 32 block
            // through shared memory so that one warp instruction adds 32 consecutive floats of one row (one 128-byte
            // L2 transaction instead of 32 scattered ones)
            float* sc = scratch + warp * (32 * 33);
#pragma unroll 1
            for (int c = 0; c < w.n; c += 32) {
              uint32_t r[32];
              ptx::tmem_ld32(lane_addr + (uint32_t)(mb * 256 + c), r);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = __uint_as_float(r[j]);
              __syncwarp();
              const bool col_ok = w.b_col0 + c + lane < w.out;
#pragma unroll 4
              for (int rr = 0; rr < 32; ++rr) {
                const int kr = __shfl_sync(0xffffffffu, krow, rr);
                if (kr >= 0 && col_ok)
                  atomicAdd(p.grad + w.koff + (long long)kr * w.out + w.b_col0 + c + lane, sc[rr * 33 + lane]);
              }
              __syncwarp();
            }
          } else {
            uint32_t r4[4];
            ptx::tmem_ld4(lane_addr + (uint32_t)(mb * 256), r4);
            ptx::tmem_ld_wait();
            if (w.flush_mode == 1) {
              atomicAdd(p.grad + w.koff + m, __uint_as_float(r4[3]));
            } else if (m < (w.head_rows > 0 ? w.head_rows : 128)) {
              atomicAdd(p.grad + w.koff + m * 3 + 0, __uint_as_float(r4[0]));
              atomicAdd(p.grad + w.koff + m * 3 + 1, __uint_as_float(r4[1]));
              atomicAdd(p.grad + w.koff + m * 3 + 2, __uint_as_float(r4[2]));
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(acc_empty);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 5) ptx::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------- CTA-pair weight-gradient kernel (wide layers)
// The layer-at-a-time path (NerfMLP width 512 / 1024) cuts dW[W, W] into 256 x 256 blocks.  This variant runs an item on a
// CTA *pair* (cluster of 2, tcgen05 cta_group::2, M = 256 over the pair): CTA r stages A columns [128 r, 128 r + 128) and dZ
// columns [128 r, 128 r + 128) of the item.  Measured on B200 (profiles/r02_wgrad_pairs.md): at 64-sample stages the two
// single threads that drive the ring (one barrier wait per stage each, ~400 cycles under load) cannot keep up with 512
// tensor-core cycles per stage, so a stage is 128 samples (four 16 KB TMA boxes per CTA issued by four lanes, eight MMAs =
// 1024 cycles per barrier round trip), and the flush of item i runs under the MMAs of item i + 1 (two 256-column
// accumulators).  Bias gradients (column sums of dZ) are NOT computed here - any reader of the dZ stages next to the ring
// slowed the kernel down by 1.5 - 2 x - but in the epilogue of the dgrad GEMM that produces dZ (DenseParams::colsum).
// Items: n == 256, flush_mode == 0, bias_mode == 0; st0 / st1 count 128-sample stages.
constexpr int kW2Stages = 3;
constexpr int kW2StageBytes = 65536;      // A: 2 atoms of 128 samples x 64 columns (32 KB) | dZ: 2 atoms (32 KB)
constexpr int kW2Threads = 192;           // warps 0-3 accumulator flush, 4 TMA producer, 5 MMA issuer
constexpr int kW2Smem = 1024 + kW2Stages * kW2StageBytes + 512 + kWgScratchFloats * 4;
static_assert(kW2Smem <= 232448, "shared memory budget");

__global__ void __launch_bounds__(kW2Threads, 1) wgrad2_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kW2Stages * kW2StageBytes);
  uint64_t* full = bars;                       // leader: both CTAs' TMA bytes of a stage have landed
  uint64_t* empty = bars + kW2Stages;          // per CTA: the MMAs that read the stage have completed (multicast commit)
  uint64_t* acc_full = bars + 2 * kW2Stages;   // [2] per CTA (multicast commit)
  uint64_t* acc_empty = acc_full + 2;          // [2] leader: both CTAs' flush threads have drained accumulator s
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* scratch = reinterpret_cast<float*>(base + kW2Stages * kW2StageBytes + 512);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);
  const uint32_t ring_u32 = ptx::smem_u32(base);
  const uint32_t full_u32 = ptx::smem_u32(full), empty_u32 = ptx::smem_u32(empty);
  const uint32_t accfull_u32 = ptx::smem_u32(acc_full), accempty_u32 = ptx::smem_u32(acc_empty);
  if (warp == 4 && lane == 0) {
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.maps[i]);
    for (int i = 0; i < kW2Stages; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 256); }
    ptx::fence_mbar_init();
  }
  if (warp == 5) ptx::tmem_alloc_cg2(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    // =============================== TMA producer (both CTAs): lane 0 waits, lanes 0..3 issue one box each ==============
    int stage = 0; uint32_t phase = 0;
    const uint32_t bar0 = ptx::mapa_u32(full_u32, 0);
    for (int it = pair; it < p.n_items; it += n_pairs) {
      const WgItem w = p.items[it];
      // lane 0, 1: A atoms; lane 2, 3: dZ atoms
      const CUtensorMap* map = &p.maps[lane < 2 ? w.a_map : w.b_map];
      const int col = (lane < 2 ? w.a_col0 : w.b_col0) + rank * 128 + (lane & 1) * 64;
      const int row0 = lane < 2 ? w.a_row0 : w.b_row0;
      for (int st = w.st0; st < w.st1; ++st) {
        if (lane == 0) {
          ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
          if (rank == 0) ptx::mbar_expect_tx_u32(full_u32 + stage * 8, 2 * kW2StageBytes);
        }
        __syncwarp();
        if (lane < 4)
          ptx::tma_load_2d_cg2(ring_u32 + stage * kW2StageBytes + lane * 16384, map, bar0 + stage * 8, col, row0 + st * 128);
        if (++stage == kW2Stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0, ae_phase = 0;
      const uint32_t idesc = ptx::make_idesc_bf16(256, 256, 1, 1);
      int i = 0;
      for (int it = pair; it < p.n_items; it += n_pairs, ++i) {
        const WgItem w = p.items[it];
        const int as = i & 1;
        if (i >= 2) {
          ptx::mbar_wait_u32(accempty_u32 + as * 8, (ae_phase >> as) & 1u);
          ae_phase ^= 1u << as;
          ptx::tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int st = w.st0; st < w.st1; ++st) {
          ptx::mbar_wait_u32(full_u32 + stage * 8, phase);
          ptx::tc_fence_after();
          const uint32_t s = ring_u32 + stage * kW2StageBytes;
#pragma unroll
          for (int k16 = 0; k16 < 8; ++k16) {
            // MN-major operands: 16 samples = 2048 bytes along K, 64-column atoms 16 KB apart, 8-row groups 1 KB apart
            const uint64_t da = ptx::make_desc_sw128(s + k16 * 2048, 16384, 1024);
            const uint64_t db = ptx::make_desc_sw128(s + 32768 + k16 * 2048, 16384, 1024);
            ptx::mma_bf16_ss_cg2(d_tmem, da, db, idesc, (st > w.st0 || k16 > 0) ? 1u : 0u);
          }
          ptx::mma_commit_mc2_u32(empty_u32 + stage * 8);
          if (++stage == kW2Stages) { stage = 0; phase ^= 1; }
        }
        ptx::mma_commit_mc2_u32(accfull_u32 + as * 8);
      }
    }
  } else {
    // =============================== flush warps 0..3 (both CTAs) ===============================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // TMEM lane = accumulator row of this CTA's half of M
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t af_phase = 0;
    int i = 0;
    for (int it = pair; it < p.n_items; it += n_pairs, ++i) {
      const WgItem w = p.items[it];
      const int as = i & 1;
      const int nst = w.st1 - w.st0;
      ptx::mbar_wait_u32(accfull_u32 + as * 8, (af_phase >> as) & 1u);
      af_phase ^= 1u << as;
      ptx::tc_fence_after();
      if (nst > 0) {
        const int m = rank * 128 + row;
        int krow;
        if (w.feat_mode) {
          const int fp = w.a_col0 + m;
          krow = fp < p.feat_dim ? w.in_base + ref_feature_col(fp, p.nb, p.ndeg) : -1;
        } else {
          krow = (w.in_rows > 0 && m >= w.in_rows) ? -1 : w.in_base + m;
        }
        // thread = accumulator row, but the gradient rows are `out` floats apart: transpose each 32 x 32 block through
        // shared memory so that one warp instruction adds 32 consecutive floats of one row
        float* sc = scratch + warp * (32 * 33);
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
          uint32_t r[32];
          ptx::tmem_ld32(lane_addr + (uint32_t)(as * 256 + c), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sc[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          const bool col_ok = w.b_col0 + c + lane < w.out;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int kr = __shfl_sync(0xffffffffu, krow, rr);
            if (kr >= 0 && col_ok)
              atomicAdd(p.grad + w.koff + (long long)kr * w.out + w.b_col0 + c + lane, sc[rr * 33 + lane]);
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(accempty_u32 + as * 8, 0));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 5) ptx::tmem_dealloc_cg2(tmem_base, 512);
}

// ---------------------------------------------------------------- CUDA-core reductions (view layer extras)
// dzv_ray[ray][c] = sum over the ray's samples of dZ_view[s][c]
__global__ void __launch_bounds__(128) ray_sum_kernel(const __nv_bfloat16* dzv, const __nv_bfloat16* dzv_lo, int ld,
                                                      int n_rays, int S, float* out) {
  const int ray = blockIdx.x, c = threadIdx.x;
  if (ray >= n_rays) return;
  float acc = 0.f;
  const __nv_bfloat16* src = dzv + (size_t)ray * S * ld + c;
  for (int s = 0; s < S; ++s) acc += __bfloat162float(src[(size_t)s * ld]);
  if (dzv_lo) {   // split-precision mode: dZ = hi + lo
    const __nv_bfloat16* lo = dzv_lo + (size_t)ray * S * ld + c;
    for (int s = 0; s < S; ++s) acc += __bfloat162float(lo[(size_t)s * ld]);
  }
  out[(size_t)ray * 128 + c] = acc;
}

// dW_view[bott + j][c] += sum_ray bf16(view_in[ray][j]) * dzv_ray[ray][c]; block = input row j, thread = c
__global__ void __launch_bounds__(128) view_extra_wgrad_kernel(const float* view_in, int view_in_dim,
                                                               const float* dzv_ray, int n_rays, long long koff,
                                                               int bott_w, int exact, float* grad) {
  const int j = blockIdx.x, c = threadIdx.x;
  const int chunk = (n_rays + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * chunk, r1 = min(r0 + chunk, n_rays);
  if (r1 <= r0) return;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) {
    float x = view_in[(size_t)r * view_in_dim + j];
    if (!exact) x = __bfloat162float(__float2bfloat16(x));
    acc = fmaf(x, dzv_ray[(size_t)r * 128 + c], acc);
  }
  atomicAdd(grad + koff + (long long)(bott_w + j) * 128 + c, acc);
}

// GLO embedding rows: d_embed[idx[ray]][g] += sum_c dzv_ray[ray][c] * bf16(W_view[bott + dir_dim + g][c])
__global__ void glo_grad_kernel(const float* dzv_ray, const int32_t* embed_idx, const float* params,
                                long long view_koff, int bott_w, int dir_dim, int glo, int n_rays,
                                long long glo_off, int num_embeddings, int exact, float* grad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * glo) return;
  const int ray = idx / glo, g = idx % glo;
  const float* wrow = params + view_koff + (long long)(bott_w + dir_dim + g) * 128;
  float acc = 0.f;
  for (int c = 0; c < 128; ++c)
    acc = fmaf(dzv_ray[(size_t)ray * 128 + c], exact ? wrow[c] : __bfloat162float(__float2bfloat16(wrow[c])), acc);
  const int row = embed_idx[ray];
  if (row < 0 || row >= num_embeddings) return;     // out-of-range rows never touch memory (hugs_forward reports them)
  atomicAdd(grad + glo_off + (long long)row * glo + g, acc);
}

}  // namespace

// ------------------------------------------------------------------------------------------
int wgrad_create(hugs_handle* h) {
  TcState* tc = h->tc;
  WgState* w = new WgState();
  tc->wg = w;
  int rc;
  const long long parts = tc->split ? 2 : 1;
  if ((rc = make_map(&w->map_act64, tc->act, parts * tc->total_save_rows, kW, 64)) ||
      (rc = make_map(&w->map_dz64, tc->dz, parts * tc->total_save_rows, kW, 64)) ||
      (rc = make_map(&w->map_feat64, tc->feat, parts * tc->total_feat_rows, kFeatPad, 64)) ||
      (rc = make_map(&w->map_dh64, tc->drgb, parts * tc->drgb_rows, kHeadCols, 64)))
    return rc;
  const int L = h->d.num_levels;
  w->host.resize(L); w->dev.assign(L, nullptr); w->built_for.assign(L, -1);
  for (int l = 0; l < L; ++l) {
    void* q = nullptr;
    HUGS_CUDA(cudaMalloc(&q, sizeof(WgItem) * kMaxWgItems));
    h->allocs.push_back(q);
    w->dev[l] = static_cast<WgItem*>(q);
  }
  void* q = nullptr;
  HUGS_CUDA(cudaMalloc(&q, sizeof(float) * (size_t)h->d.max_rays * 128));
  h->allocs.push_back(q);
  w->dzv_ray = static_cast<float*>(q);
  HUGS_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
  HUGS_CUDA(cudaFuncSetAttribute(wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2Smem));
  return HUGS_OK;
}

void wgrad_destroy(hugs_handle* h) {
  if (h->tc && h->tc->wg) { delete h->tc->wg; h->tc->wg = nullptr; }
}

static void build_items(hugs_handle* h, int level, int n_tiles, std::vector<WgItem>* items) {
  TcState* tc = h->tc;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int D = mv.depth;
  const int cap = tc->cap[level], srow = tc->save_row0[level], frow = tc->feat_row0[level];
  const int T = n_tiles * 2;   // 64-sample stages
  // Units that stream a common operand (the IPE features against dZ of layer 0 and of the skip layer; the last trunk
  // activation against the bottleneck and the density-head gradients) form a *group*: they are cut into the same sample
  // ranges and queued next to each other, so that neighbouring CTAs stream the shared rows at the same time and the
  // second reader hits L2 instead of HBM (the kernel is HBM-bound; this removes up to 20 % of its DRAM reads).
  std::vector<WgUnit> units;
  int next_group = 16;
  constexpr int kGroupFeat = 1, kGroupLast = 2;
  auto add = [&](int a_map, int a_row0, int a_col0, int b_slot, int n, const DenseView& v, int in_base, int feat_mode,
                 bool first_of_layer, int group) {
    WgItem w{};
    w.a_map = a_map ? WG_MAP_FEAT : WG_MAP_ACT; w.a_row0 = a_row0; w.a_col0 = a_col0; w.b_map = WG_MAP_DZ;
    w.b_row0 = srow + b_slot * cap; w.n = n;
    w.out = v.out; w.koff = v.kernel_off; w.in_base = in_base; w.feat_mode = feat_mode;
    w.bias_mode = first_of_layer ? 1 : 0; w.boff = v.bias_off;
    // the kernel is HBM-bound: cost = bytes streamed per sample
    units.push_back({w, (512.f + 2.f * n) / 1024.f, group > 0 ? group : next_group++});
  };
  bool cat = false;
  const int n_feat_sb = (h->feat_panels + 3) / 4;    // 256-column superblocks of the feature tensor that carry features
  for (int l = 0; l < D; ++l) {
    const DenseView& v = mv.dense[l];
    if (l == 0) {
      for (int sb = 0; sb < n_feat_sb; ++sb) add(1, frow, sb * 256, 0, 256, v, 0, 1, sb == 0, kGroupFeat);
    } else {
      add(0, srow + (l - 1) * cap, 0, l, 256, v, 0, 0, true, cat ? kGroupFeat : 0);
      if (cat) for (int sb = 0; sb < n_feat_sb; ++sb) add(1, frow, sb * 256, l, 256, v, kW, 1, false, kGroupFeat);
    }
    cat = (l % d.skip_layer == 0 && l > 0);
  }
  if (mv.has_rgb) {
    add(0, srow + (D - 1) * cap, 0, D, 256, mv.dense[D + 1], 0, 0, true, kGroupLast);     // bottleneck
    add(0, srow + D * cap, 0, D + 1, 128, mv.dense[D + 2], 0, 0, true, 0);               // view layer (bottleneck rows)
  }
  {  // density head: A = last trunk activation, B = head gradients (column 3)
    WgItem w{};
    w.a_map = WG_MAP_ACT; w.a_row0 = srow + (D - 1) * cap; w.a_col0 = 0; w.b_map = WG_MAP_DH; w.b_row0 = 0; w.n = kHeadCols;
    w.out = 1; w.koff = mv.dense[D].kernel_off; w.flush_mode = 1; w.bias_mode = 2; w.boff = mv.dense[D].bias_off;
    units.push_back({w, (512.f + 2.f * kHeadCols) / 1024.f, kGroupLast});
  }
  if (mv.has_rgb) {  // rgb head: A = view activation (columns 128..255 are zero), B = head gradients (columns 0..2)
    WgItem w{};
    w.a_map = WG_MAP_ACT; w.a_row0 = srow + (D + 1) * cap; w.a_col0 = 0; w.b_map = WG_MAP_DH; w.b_row0 = 0; w.n = kHeadCols;
    w.out = 3; w.koff = mv.dense[D + 3].kernel_off; w.flush_mode = 2; w.bias_mode = 3; w.boff = mv.dense[D + 3].bias_off;
    units.push_back({w, (512.f + 2.f * kHeadCols) / 1024.f, next_group++});
  }
  std::vector<WgItem> planned;
  wgrad_plan(units, T, tc->num_sms, &planned);
  // split-precision mode: (A_hi + A_lo)^T (dZ_hi + dZ_lo) as four items; the bias column sums ride on the A_hi items only,
  // so that each dZ half is summed exactly once
  const int n_a = tc->split ? 2 : 1, n_b = tc->split ? 2 : 1;
  items->clear();
  for (const WgItem& base : planned)
    for (int pa = 0; pa < n_a; ++pa)
      for (int pb = 0; pb < n_b; ++pb) {
        WgItem w = base;
        w.a_row0 += pa * (w.a_map == WG_MAP_FEAT ? tc->total_feat_rows : tc->total_save_rows);
        w.b_row0 += pb * (w.b_map == WG_MAP_DH ? tc->drgb_rows : tc->total_save_rows);
        if (pa > 0) w.bias_mode = 0;
        items->push_back(w);
      }
}

// Cuts the units of one weight-gradient launch into work items (sample ranges).  Units of a *group* stream a common
// operand: they are cut into the same sample ranges and queued next to each other, so that neighbouring CTAs stream the
// shared rows at the same time and the second reader hits L2 instead of HBM.  T = number of 64-sample stages.
void wgrad_plan(const std::vector<WgUnit>& units, int T, int num_sms, std::vector<WgItem>* items) {
  float total = 0.f;
  for (auto& u : units) total += u.cost;
  items->clear();
  std::vector<int> group_ids;   // in first-appearance order
  for (auto& u : units)
    if (std::find(group_ids.begin(), group_ids.end(), u.group) == group_ids.end()) group_ids.push_back(u.group);
  // Several items per CTA even out the differences between items (measured: NerfMLP 1.28 -> 1.17 ms with 3 items per
  // CTA), but every item pays one accumulator flush: only when an item still streams >= ~250 stages.
  const float stages_per_sm = (float)T * total / (float)num_sms;
  const int waves = std::min(4, std::max(1, (int)(stages_per_sm / 250.f)));
  const int target = num_sms * std::max(1, waves);
  // sample ranges per group: proportional to the mean cost of its units (at least 1, at most one range per 4 stages)
  const int max_splits = std::max(1, T / 4);
  std::vector<int> splits(group_ids.size()), members(group_ids.size(), 0);
  std::vector<float> gcost(group_ids.size(), 0.f);
  for (auto& u : units) {
    const size_t g = std::find(group_ids.begin(), group_ids.end(), u.group) - group_ids.begin();
    ++members[g]; gcost[g] += u.cost;
  }
  int used = 0;
  for (size_t g = 0; g < group_ids.size(); ++g) {
    splits[g] = std::max(1, (int)(target * gcost[g] / total / members[g]));
    splits[g] = std::min(splits[g], max_splits);
    used += splits[g] * members[g];
  }
  for (size_t i = 0; i < group_ids.size() * 4; ++i) {   // hand out the remainder round-robin
    const size_t g = i % group_ids.size();
    if (used + members[g] <= target && splits[g] < max_splits) { ++splits[g]; used += members[g]; }
  }
  for (size_t g = 0; g < group_ids.size(); ++g)
    for (int k = 0; k < splits[g]; ++k)
      for (auto& u : units) {
        if (u.group != group_ids[g]) continue;
        WgItem w = u.w;
        w.st0 = (int)((long long)T * k / splits[g]);
        w.st1 = (int)((long long)T * (k + 1) / splits[g]);
        if (w.st1 > w.st0) items->push_back(w);
      }
}

// One launch of the weight-gradient kernel over `n_items` device-resident work items.
int wgrad_launch(hugs_handle* h, const CUtensorMap* maps, int n_maps, const WgItem* dev_items, int n_items, float* grad,
                 cudaStream_t st) {
  return wgrad_launch_raw(h->tc->num_sms, h->perm_nb, h->d.max_deg_point - h->d.min_deg_point, h->feat_dim, maps, n_maps,
                          dev_items, n_items, grad, st);
}

int wgrad_kernel_init() {
  HUGS_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
  HUGS_CUDA(cudaFuncSetAttribute(wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kW2Smem));
  return HUGS_OK;
}

// The CTA-pair kernel: every item must be a 256 x 256 kernel block (n == 256, flush_mode == 0, bias_mode == 0); `maps` have
// boxes of 128 rows and the items count 128-sample stages.
int wgrad2_launch_raw(int num_sms, int perm_nb, int ndeg, int feat_dim, const CUtensorMap* maps, int n_maps,
                      const WgItem* dev_items, int n_items, float* grad, cudaStream_t st) {
  if (n_items <= 0) return HUGS_OK;
  HUGS_REQUIRE(n_maps <= kWgMaxMaps, "wgrad: too many tensor maps");
  WgParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n_maps; ++i) p.maps[i] = maps[i];
  for (int i = n_maps; i < kWgMaxMaps; ++i) p.maps[i] = maps[0];
  p.items = dev_items; p.n_items = n_items;
  p.nb = perm_nb; p.ndeg = ndeg; p.feat_dim = feat_dim; p.grad = grad;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * std::min(n_items, num_sms / 2));
  cfg.blockDim = dim3(kW2Threads);
  cfg.dynamicSmemBytes = kW2Smem;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  HUGS_CUDA(cudaLaunchKernelEx(&cfg, wgrad2_kernel, p));
  ++g_launch_count;
  return HUGS_OK;
}

int wgrad_launch_raw(int num_sms, int perm_nb, int ndeg, int feat_dim, const CUtensorMap* maps, int n_maps,
                     const WgItem* dev_items, int n_items, float* grad, cudaStream_t st) {
  if (n_items <= 0) return HUGS_OK;
  HUGS_REQUIRE(n_maps <= kWgMaxMaps, "wgrad: too many tensor maps");
  WgParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n_maps; ++i) p.maps[i] = maps[i];
  for (int i = n_maps; i < kWgMaxMaps; ++i) p.maps[i] = maps[0];
  p.items = dev_items; p.n_items = n_items;
  p.nb = perm_nb; p.ndeg = ndeg; p.feat_dim = feat_dim; p.grad = grad;
  wgrad_kernel<<<std::min(n_items, num_sms), kWgThreads, kWgSmem, st>>>(p);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

int wgrad_run(hugs_handle* h, int level, int n_rays, float* grad, cudaStream_t st) {
  TcState* tc = h->tc;
  WgState* w = tc->wg;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int D = mv.depth, S = h->samples(level);
  const int n_samples = n_rays * S;
  const int n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int cap = tc->cap[level], srow = tc->save_row0[level];
  if (w->built_for[level] != n_samples) {
    build_items(h, level, n_tiles, &w->host[level]);
    HUGS_REQUIRE(w->host[level].size() <= (size_t)kMaxWgItems, "wgrad: too many work items");
    HUGS_CUDA(cudaMemcpyAsync(w->dev[level], w->host[level].data(), sizeof(WgItem) * w->host[level].size(),
                              cudaMemcpyHostToDevice, st));
    HUGS_CUDA(cudaStreamSynchronize(st));
    w->built_for[level] = n_samples;
  }
  {
    ProfScope ps(h, is_prop ? HUGS_K_WGRAD_PROP : HUGS_K_WGRAD_NERF, st);
    const CUtensorMap maps[4] = {w->map_act64, w->map_feat64, w->map_dz64, w->map_dh64};
    int rc = wgrad_launch(h, maps, 4, w->dev[level], (int)w->host[level].size(), grad, st);
    if (rc) return rc;
  }
  ProfScope ps_red(h, HUGS_K_REDUCTIONS, st);
  if (mv.has_rgb) {
    const __nv_bfloat16* dzv = tc->dz + (size_t)(srow + (D + 1) * cap) * kW;
    return wgrad_view_extras(h, dzv, tc->split ? dzv + (size_t)tc->total_save_rows * kW : nullptr, kW, n_rays, S, grad, st);
  }
  return HUGS_OK;
}

// view layer: weight rows of the per-ray inputs (direction encoding, GLO vector) and the GLO embedding rows, from the
// per-ray sums of dZ_view (rows of `dz_ld` bf16 elements, 128 valid columns)
int wgrad_view_extras(hugs_handle* h, const __nv_bfloat16* dzv, const __nv_bfloat16* dzv_lo, int dz_ld, int n_rays, int S,
                      float* grad, cudaStream_t st) {
  TcState* tc = h->tc;
  WgState* w = tc->wg;
  const hugs_model_desc& d = h->d;
  const DenseView& vv = h->nerf.dense[h->nerf.depth + 2];
  const int exact = tc->split ? 1 : 0;
  ray_sum_kernel<<<n_rays, 128, 0, st>>>(dzv, dzv_lo, dz_ld, n_rays, S, w->dzv_ray);
  HUGS_LAUNCH_CHECK();
  view_extra_wgrad_kernel<<<dim3(h->view_in_dim, 32), 128, 0, st>>>(h->view_in, h->view_in_dim, w->dzv_ray, n_rays,
                                                                    vv.kernel_off, d.bottleneck_width, exact, grad);
  HUGS_LAUNCH_CHECK();
  if (d.num_glo_features > 0) {
    const int tot = n_rays * d.num_glo_features;
    glo_grad_kernel<<<(tot + 127) / 128, 128, 0, st>>>(w->dzv_ray, h->cur_embed_idx, h->cur_params, vv.kernel_off,
                                                       d.bottleneck_width, 3 + 6 * d.deg_view, d.num_glo_features, n_rays,
                                                       h->glo_off, d.num_embeddings, exact, grad);
    HUGS_LAUNCH_CHECK();
  }
  return HUGS_OK;
}

}  // namespace hugs
