// MLP chain kernel (sm_100a): the whole MLP of models.py:437-519 (and its dgrad chain) on a CTA *pair*
// (cluster of 2, tcgen05 cta_group::2) without leaving the SMs.
//
//   shared memory  panels[8]  : tile t owns panels 4t..4t+3 = its 128 x 256 bf16 activation, updated IN PLACE
//                               (a layer's epilogue starts only after all MMAs of that layer have completed)
//                  ring[5]    : 16 KB weight stages ([128 out x 64 in] bf16, SWIZZLE_128B), TMA-fed
//                  head table, head partial sums, bias MMA operands, mbarriers
//   tensor memory  tile t accumulates in columns [256t, 256t+256)
//   warps          0-15 epilogue (group q = warp/4 owns output columns [64q, 64q+64)),
//                  16 weight producer, 17 feature producer, 18 MMA issuer
//
// Two 128-sample tiles are in flight per CTA (ping-pong): while the epilogue warps turn tile A's accumulator into the
// next layer's A operand, the tensor core already runs tile B's MMAs of the same layer.  Every MMA is M = 256 (128 rows
// per CTA) x N = 256, each CTA stages only its half of the weight rows; a weight stage (one K panel) is fetched once
// per unit and used by both tiles.  The leader CTA issues all MMAs; barriers the issuer waits on live in the leader and
// receive the peer's TMA completions / epilogue arrivals through the cluster address space, barriers signalled by the
// tensor core are multicast to both CTAs.  Layer biases are applied by a K = 16 MMA (ones x [bias_hi, bias_lo]) that
// initialises the accumulator; the epilogue is one TMEM wait + pack + eight 16-byte stores; the forward stores 64-bit
// ReLU gate masks which the backward program loads one epilogue ahead.
//
// A layer is a sequence of *segments* of <= 4 K-panels (K <= 256).  IPE features (layer 0 and the skip connection) are
// TMA-loaded straight into the tile's own panels once the previous segment has been consumed; the N = 1 / N = 3 heads
// are evaluated on CUDA cores inside the producing epilogue.
//
// kSplit = true (HUGS_PRECISION_TC_SPLIT) runs the SAME program - same producers, issuer, barriers, descriptors, TMEM
// and panel layout - with every bf16 operand split into a hi and a lo half (x = hi + lo, |x - hi - lo| <= 2^-16 |x|):
// a CTA then owns ONE tile whose hi half lives in panels 0..3 and whose lo half lives in panels 4..7 (the slot of the
// second tile), the "tile" loop of the issuer becomes the loop over the two A halves, every segment is issued once
// against the hi weights and once against the lo weights, and all four products accumulate into one fp32 accumulator.
// The epilogue keeps fp32 values end to end (heads, gates, view bias), writes hi and lo panels and saves both.
// Renders then match the fp32 oracle to ~1e-5 and gradients to ~1e-4 through the tensor-core path itself.
// DESIGN.md ("Tensor-core kernel structure") has the rationale and profiles/r01_ab_experiments.md the measurements.
#include <algorithm>

#include "tc_device.cuh"
#include "tc_internal.h"

namespace hugs {
namespace {

constexpr int kPpThreads = (kEpiGroups * 4 + 3) * 32;
// ONE barrier per tile (panel_ready[4t]) collects the elected arrivals of all 4 epilogue groups of both CTAs.  A try_wait
// costs the issuing thread ~350 cycles under load even on a completed phase (shared-memory pipe queue), and four of
// them in a row left the tensor pipe idle for ~1 k cycles per tile and layer (event trace, profiles/r01_ab_experiments.md).
constexpr int kPanelArrivals = 8;
constexpr int kPartFloats = 768;     // head partial sums: [2 tiles][3][128]
constexpr int kBiasImgBytes = kBiasChunks * kBiasChunkElems * 2;
constexpr int kOnesBytes = 4 * 256;   // variant v: core matrix with 1.0 in K columns 2v, 2v+1, then a zero core matrix
constexpr int kPpSmemBytes = 1024 + (kNumPanels + kStages) * kPanelBytes + kBiasTailFloats * 4 + kPartFloats * 4 +
                             kBiasImgBytes + kOnesBytes + 512;
static_assert(kPpSmemBytes <= 232448, "shared memory budget");

struct PpSmem {
  uint8_t* panels; uint8_t* ring; float* bias; float* part;
  uint8_t* bias_img; uint8_t* ones;
  uint64_t *full, *empty, *panel_ready, *feat_ready, *acc_full, *consumed, *epi_done;
  uint32_t* tmem_ptr;
};

__device__ __forceinline__ PpSmem pp_carve(uint8_t* raw) {
  PpSmem s;
  // offset arithmetic on the __shared__ symbol (not an integer round trip) keeps the shared address space, so the
  // accesses below compile to LDS/STS instead of generic loads and stores
  uint8_t* base = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  s.panels = base;
  s.ring = base + kNumPanels * kPanelBytes;
  s.bias = reinterpret_cast<float*>(s.ring + kStages * kPanelBytes);
  s.part = s.bias + kBiasTailFloats;
  s.bias_img = reinterpret_cast<uint8_t*>(s.part + kPartFloats);
  s.ones = s.bias_img + kBiasImgBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.ones + kOnesBytes);
  s.full = bars; s.empty = bars + kStages;
  s.panel_ready = bars + 2 * kStages;        // [8]
  s.feat_ready = s.panel_ready + 8;          // [8]
  s.acc_full = s.feat_ready + 8;             // [2]
  s.consumed = s.acc_full + 2;               // [8]  per (tile, K panel): the MMAs that read the panel have completed
  s.epi_done = s.consumed + 8;               // [2]
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.epi_done + 2);
  return s;
}

// fp32 pair -> packed bf16 hi word and packed bf16 residual word (x = hi + lo up to 2^-16 |x|)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = ptx::pack_bf16x2(a, b);
  lo = ptx::pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}

template <bool kTrain, bool kSplit>
__global__ void __launch_bounds__(kPpThreads, 1) mlp_pp_kernel(const __grid_constant__ PpParams p) {
  extern __shared__ uint8_t smem_raw[];
  PpSmem sm = pp_carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWProd = kEpiGroups * 4, kFProd = kWProd + 1, kMma = kWProd + 2;
  // work distribution: a *unit* is one pass of the segment program over the pair's tiles (2 per CTA; 1 when kSplit)
  const int rank = (int)ptx::cluster_ctarank();
  const int unit0 = (int)(blockIdx.x >> 1);
  const int unit_stride = (int)(gridDim.x >> 1);
  // tile (of 128 samples) that index t of this CTA works on in `unit`; kSplit: t is the operand half of the one tile
  auto tile_of = [&](int unit, int t) { return kSplit ? unit * 2 + rank : (unit * 2 + t) * 2 + rank; };

  const uint32_t panels_u32 = ptx::smem_u32(sm.panels), ring_u32 = ptx::smem_u32(sm.ring);
  const uint32_t full_u32 = ptx::smem_u32(sm.full), empty_u32 = ptx::smem_u32(sm.empty);
  const uint32_t pready_u32 = ptx::smem_u32(sm.panel_ready), fready_u32 = ptx::smem_u32(sm.feat_ready);
  const uint32_t accfull_u32 = ptx::smem_u32(sm.acc_full), consumed_u32 = ptx::smem_u32(sm.consumed);
  const uint32_t epidone_u32 = ptx::smem_u32(sm.epi_done);
  const uint32_t ones_u32 = ptx::smem_u32(sm.ones), biasimg_u32 = ptx::smem_u32(sm.bias_img);

  if (warp == kWProd && lane == 0) {
    ptx::prefetch_tmap(&p.map_w); ptx::prefetch_tmap(&p.map_feat); ptx::prefetch_tmap(&p.map_save);
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&sm.full[i], 1); ptx::mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(&sm.panel_ready[i], kPanelArrivals); ptx::mbar_init(&sm.feat_ready[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&sm.acc_full[i], 1); ptx::mbar_init(&sm.epi_done[i], kEpiGroups); }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&sm.consumed[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == kMma) ptx::tmem_alloc_cg2(sm.tmem_ptr, 512);
  {
    // sm.bias holds table entries [bias_tail0, bias_floats): head weights and head biases; the layer biases are
    // staged as MMA operands (this CTA's half of the output rows)
    for (int i = p.bias_tail0 + threadIdx.x; i < p.bias_floats; i += kPpThreads) sm.bias[i - p.bias_tail0] = p.bias[i];
    const uint4* src = reinterpret_cast<const uint4*>(p.bias_img + (size_t)rank * kBiasChunks * kBiasChunkElems);
    for (int i = threadIdx.x; i < kBiasImgBytes / 16; i += kPpThreads) reinterpret_cast<uint4*>(sm.bias_img)[i] = src[i];
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kPpThreads) {
      // 32-bit word i: variant v = i / 64, word w = i % 64; core matrix 0 = words 0..31 (row r = w / 4, K pair w % 4)
      const int v = i >> 6, w = i & 63;
      reinterpret_cast<uint32_t*>(sm.ones)[i] = (w < 32 && (w & 3) == v) ? 0x3F803F80u : 0u;
    }
    ptx::fence_proxy_async();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kWProd) {
    // =============================== weight producer ===============================
    // this CTA stages rows [rank * N/2, (rank + 1) * N/2) of every weight tile; the transaction bytes of both CTAs are
    // accounted on the leader's `full` barrier.  One stage per K panel, shared by the unit's two tiles (the issuer
    // releases it after tile 1).
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = unit0; unit < p.n_units; unit += unit_stride) {
        for (int si = 0; si < p.n_segs; ++si) {
          const PpSeg& S = p.segs[si];
          if (S.kps == 0) continue;
          const int kps = S.kps, n_halves = S.n_halves, w_col0 = S.w_col0;
          const int w_row = S.w_row + rank * n_halves * 64;
          const CUtensorMap* map = n_halves == 2 ? &p.map_w : &p.map_w_half;
          const uint32_t bytes = (uint32_t)n_halves * (kPanelBytes / 2);
          for (int kp = 0; kp < kps; ++kp) {
            ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
            if (rank == 0) ptx::mbar_expect_tx_u32(full_u32 + stage * 8, 2 * bytes);
            ptx::tma_load_2d_cg2(ring_u32 + stage * kPanelBytes, map, ptx::mapa_u32(full_u32 + stage * 8, 0),
                                 w_col0 + kp * 64, w_row);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kFProd) {
    // =============================== feature producer ===============================
    // Walks the MMA segments in issue order and waits for every `consumed` phase exactly once (no phase is
    // ever skipped, so parity waits cannot alias); refills a tile's panels with IPE feature columns as soon
    // as the segment that last read those panels has completed.
    if (lane == 0 && p.any_feat) {
      uint32_t cons_phase = 0;   // bit t*4+kp: parity of the next `consumed` phase of that panel to wait for
      int prev_kps0 = 0, prev_kps1 = 0;   // K panels of the previous MMA segment on tile 0 / 1 (0: none yet)
      int unit_iter = 0;
      for (int unit = unit0; unit < p.n_units; unit += unit_stride, ++unit_iter) {
        bool first_in_unit = true;
        for (int si = 0; si < p.n_segs; ++si) {
          const PpSeg& S = p.segs[si];
          if (S.kps == 0) continue;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int pk = t == 0 ? prev_kps0 : prev_kps1;
            const int row = p.feat_row0 + tile_of(unit, t) * kTileM + (kSplit ? t * p.lo_feat_rows : 0);
            const int n_it = S.kps > pk ? S.kps : pk;
            for (int kp = 0; kp < n_it; ++kp) {
              const uint32_t idx = (uint32_t)(t * 4 + kp);
              if (kp < pk) {   // panel kp is free once the previous segment's MMAs on it have completed
                ptx::mbar_wait_u32(consumed_u32 + idx * 8, (cons_phase >> idx) & 1u);
                cons_phase ^= 1u << idx;
              }
              if (S.a_feat && kp < S.kps) {
                if (first_in_unit && unit_iter > 0 && kp == 0)   // previous unit's last epilogue (and its TMA store) is done
                  ptx::mbar_wait_u32(epidone_u32 + t * 8, (uint32_t)((unit_iter - 1) & 1));
                const uint32_t bar = fready_u32 + idx * 8;
                if (rank == 0) ptx::mbar_expect_tx_u32(bar, 2 * kPanelBytes);
                ptx::tma_load_2d_cg2(panels_u32 + idx * kPanelBytes, &p.map_feat, ptx::mapa_u32(bar, 0),
                                     S.feat_col0 + kp * 64, row);
              }
            }
            const int sig = S.feat_next ? S.kps : 0;
            if (t == 0) prev_kps0 = sig; else prev_kps1 = sig;
          }
          first_in_unit = false;
        }
      }
    }
  } else if (warp == kMma) {
    // =============================== MMA issuer ===============================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t kDescHi = ptx::desc_hi_sw128(1024);
      const uint32_t idesc128 = ptx::make_idesc_bf16(256, 128, 0, 0);
      const uint32_t idesc256 = ptx::make_idesc_bf16(256, 256, 0, 0);
      // the issuing thread never reads the data behind these barriers itself (the tensor core does, behind
      // tcgen05.fence::after_thread_sync), so CTA-scope waits suffice also for the peer's arrivals
      // (a polling wait with a short suspend hint instead of the hardware-suspended one measured the same and only adds
      //  shared-memory traffic: profiles/r01_ab_experiments.md)
      auto wait = [](uint32_t bar, uint32_t parity) { ptx::mbar_wait_u32(bar, parity); };
      auto commit = [](uint32_t bar) { ptx::mma_commit_mc2_u32(bar); };
      int stage = 0; uint32_t phase = 0;
      uint32_t wait_phase = 0;   // bits 0-7 panel_ready, 8-15 feat_ready
      for (int unit = unit0; unit < p.n_units; unit += unit_stride) {
        for (int si = 0; si < p.n_segs; ++si) {
          const PpSeg& S = p.segs[si];
          if (S.kps == 0) continue;
          const int kps = S.kps, n_halves = S.n_halves, a_feat = S.a_feat, acc0 = S.accumulate;
          const bool has_epi = S.epi != EPI_NONE;
          for (int t = 0; t < 2; ++t) {
            // kSplit: both operand halves (t = 0 hi panels, t = 1 lo panels) accumulate into the one accumulator
            const uint32_t d_tmem = tmem_base + (uint32_t)(kSplit ? 0 : t * 256);
            if (!a_feat && !S.no_wait) {
              // in-place accumulator: every epilogue group must have drained the previous layer before the
              // first MMA of this one overwrites it, so wait for the tile's panels up front
              const uint32_t idx = (uint32_t)(t * 4);
              wait(pready_u32 + idx * 8, (wait_phase >> idx) & 1u);
              wait_phase ^= 1u << idx;
              ptx::tc_fence_after();
            }
            bool bias_pending = S.bias_idx >= 0 && (!kSplit || t == 0);
            for (int kp = 0; kp < kps; ++kp) {
              if (a_feat) {
                const uint32_t idx = (uint32_t)(8 + t * 4 + kp);
                wait(fready_u32 + (t * 4 + kp) * 8, (wait_phase >> idx) & 1u);
                wait_phase ^= 1u << idx;
                ptx::tc_fence_after();
              }
              const uint64_t da = ptx::desc_from(kDescHi, panels_u32 + (t * 4 + kp) * kPanelBytes);
              uint32_t accum = (acc0 || kp > 0 || (kSplit && t > 0)) ? 1u : 0u;
              if (bias_pending) {
                // acc = ones[256 x 16] * [bias_hi, bias_lo, ...]^T: initialises the accumulator with the layer bias.
                // No-swizzle K-major tiles: A = one 8-row core matrix for every row group (SBO 0) followed by a zero
                // core matrix for K 8..15 (LBO 128); B = 16 row groups of this CTA's outputs (SBO 128), K 8..15
                // aliasing K 0..7 (LBO 0, multiplied by zeros).
                const int j = S.bias_idx;
                const uint64_t oa = ptx::make_desc_nosw(ones_u32 + (uint32_t)(j & 3) * 256u, 128u, 0u);
                const uint64_t ob = ptx::make_desc_nosw(biasimg_u32 + (uint32_t)(j >> 2) * (kBiasChunkElems * 2), 0u, 128u);
                ptx::mma_bf16_ss_cg2(d_tmem, oa, ob, idesc256, 0u);
                accum = 1u;
                bias_pending = false;
              }
              // one stage = this K panel of all N columns (each CTA holds its half of the rows); tile 0 waits
              // for it, tile 1 reuses it and releases it
              int st_k = stage + kp; uint32_t ph_k = phase;
              if (st_k >= kStages) { st_k -= kStages; ph_k ^= 1; }
              if (t == 0) {
                wait(full_u32 + st_k * 8, ph_k);
                ptx::tc_fence_after();
              }
              const uint64_t db = ptx::desc_from(kDescHi, ring_u32 + st_k * kPanelBytes);
              const uint32_t idesc = n_halves == 2 ? idesc256 : idesc128;
              ptx::mma_bf16_ss_cg2(d_tmem, da, db, idesc, accum);
              ptx::mma_bf16_ss_cg2(d_tmem, da + 2, db + 2, idesc, 1u);
              ptx::mma_bf16_ss_cg2(d_tmem, da + 4, db + 4, idesc, 1u);
              ptx::mma_bf16_ss_cg2(d_tmem, da + 6, db + 6, idesc, 1u);
              if (t == 1) commit(empty_u32 + st_k * 8);
              if (S.feat_next) commit(consumed_u32 + (t * 4 + kp) * 8);
            }
            if (has_epi && (!kSplit || t == 1)) commit(accfull_u32 + (kSplit ? 0 : t) * 8);
          }
          stage += kps;
          if (stage >= kStages) { stage -= kStages; phase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue groups ===============================
    const int q = warp >> 2;               // group: output columns [64q, 64q+64)
    const int quarter = warp & 3;          // TMEM lane quarter of this warp
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool group_leader = (warp & 3) == 0 && lane == 0;
    const int bar_id = 1 + q;
    uint32_t acc_phase = 0;
    float v[32];
    float raw_keep[2] = {0.f, 0.f};
    const int col = q * 64;
    const int tail0 = p.bias_tail0;   // sm.bias[i - tail0] == bias table entry i
    // epilogue e of a unit: segment p.epi_seg[e >> 1] on tile e & 1 (kSplit: segment p.epi_seg[e] on the one tile)
    auto seg_of = [&](int e) -> const PpSeg& { return p.segs[p.epi_seg[kSplit ? e : (e >> 1)]]; };
    auto t_of = [&](int e) { return kSplit ? 0 : (e & 1); };

    // make the freshly written panel visible to the async proxy, optionally TMA-store it (its smem read is
    // complete before panel_ready completes, so later in-place overwrites / feature refills are safe),
    // then hand it to the MMA issuer: one elected arrival per group (on the leader's barrier) after the group
    // has synchronised
    auto publish = [&](const PpSeg& S, int pi, int tile, bool tile_ok) {
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (group_leader) {
        const bool save = kTrain && S.save_row >= 0 && tile_ok;
        if (save) {
          ptx::tma_store_2d(&p.map_save, sm.panels + pi * kPanelBytes, col, S.save_row + tile * kTileM);
          if (kSplit)
            ptx::tma_store_2d(&p.map_save, sm.panels + (pi + 4) * kPanelBytes, col,
                              S.save_row + p.lo_save_rows + tile * kTileM);
          ptx::tma_commit_group();
        }
        if (!S.no_signal) {
          ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(pready_u32 + (pi & ~3) * 8, 0));
          if (kSplit) ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(pready_u32 + 4 * 8, 0));
        }
        // the store's shared-memory read must be over before the panel is rewritten; this group's next
        // write to it is behind the bar.sync of the other tile's publish (or of last_epi), which this
        // thread only reaches after the wait
        if (save) ptx::tma_wait_group_read<0>();
      }
    };

    // ReLU gate bits of epilogue (unit, e): loaded one epilogue ahead (backward programs)
    auto load_gate = [&](int unit, int e) -> uint2 {
      const PpSeg& G = seg_of(e);
      if (G.mask_row < 0 || ((G.epi == EPI_BWD_START) && q >= 2)) return make_uint2(0u, 0u);
      const int s2 = tile_of(unit, t_of(e)) * kTileM + row;
      if (s2 >= p.n_samples) return make_uint2(0u, 0u);      // padding rows carry no gradient
      return __ldg(p.gate + (size_t)G.mask_row * 4 + (size_t)q * p.cap + s2);
    };
    const int n_e = (kSplit ? 1 : 2) * p.n_epi;
    uint2 gate_next = make_uint2(0u, 0u);
    if (unit0 < p.n_units) gate_next = load_gate(unit0, 0);

    // 16 packed words (32 columns) -> chunks chunk0 .. chunk0 + 3 of the swizzled panel row
    auto store_pk16 = [&](uint8_t* panel, int chunk0, const uint32_t* pk) {
      uint4* prow = reinterpret_cast<uint4*>(panel + row * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        prow[swz_chunk(row, chunk0 + c)] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    };
    // split mode: 32 fp32 values -> hi words into `panel`, residual words into the panel 4 slots further
    auto store_split32 = [&](uint8_t* panel, int chunk0, const float* x) {
      uint32_t ph[16], pl[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) split_pair(x[2 * j], x[2 * j + 1], ph[j], pl[j]);
      store_pk16(panel, chunk0, ph);
      store_pk16(panel + 4 * kPanelBytes, chunk0, pl);
    };
    // gate bits of 32 fp32 values (bit j = value 2j is positive, bit 16 + j = value 2j + 1), same layout as gate_bits16
    auto gate_bits32f = [&](const float* x) {
      uint32_t g = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) g |= ((x[2 * j] > 0.f) ? (1u << j) : 0u) | ((x[2 * j + 1] > 0.f) ? (0x10000u << j) : 0u);
      return g;
    };
    auto apply_gate32f = [&](uint32_t g, float* x) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (!((g >> j) & 1u)) x[2 * j] = 0.f;
        if (!((g >> (16 + j)) & 1u)) x[2 * j + 1] = 0.f;
      }
    };

    for (int unit = unit0; unit < p.n_units; unit += unit_stride) {
      for (int e = 0; e < n_e; ++e) {
        const PpSeg& S = seg_of(e);
        const int t = t_of(e);
        const uint2 gate = gate_next;
        if (e + 1 < n_e) gate_next = load_gate(unit, e + 1);
        else if (unit + unit_stride < p.n_units) gate_next = load_gate(unit + unit_stride, 0);
        const int tile = tile_of(unit, t);
        const bool tile_ok = tile < p.n_tiles;
        const int s = tile * kTileM + row;
        const bool valid = s < p.n_samples;
        const int pi = t * 4 + q;
        uint8_t* panel = sm.panels + pi * kPanelBytes;
        const uint32_t acc_addr = lane_addr + (uint32_t)(t * 256 + col);

        // side inputs that do not depend on the accumulator are fetched before waiting for the MMAs
        float dd = 0.f;
        if (S.epi == EPI_BWD_RELU_D) {
          dd = valid ? p.d_raw[(size_t)s * p.raw_c] : 0.f;
          if (!kSplit) dd = __bfloat162float(__float2bfloat16(dd));
        }

        if (S.kps > 0) {
          // every group waits for every phase, also when it has no columns in this layer: a parity wait
          // that skipped a phase could be satisfied by an older phase of the same parity
          ptx::mbar_wait_u32(accfull_u32 + t * 8, (acc_phase >> t) & 1u);
          acc_phase ^= 1u << t;
          ptx::tc_fence_after();
        }

        switch (S.epi) {
          case EPI_RELU: case EPI_LINEAR: {
            float head = 0.f;
            if (kSplit) {
              uint32_t g01[2] = {0u, 0u};
#pragma unroll 1
              for (int hf = 0; hf < 2; ++hf) {
                load_acc32(acc_addr + (uint32_t)(hf * 32), v);
                if (S.epi == EPI_RELU) {
#pragma unroll
                  for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
                  g01[hf] = gate_bits32f(v);
                }
                if (S.head) {   // density head in fp32 (unrounded kernel)
                  const float* wd = sm.bias + (p.w_dens_off - tail0) + col + hf * 32;
#pragma unroll
                  for (int c = 0; c < 32; ++c) head = fmaf(v[c], wd[c], head);
                }
                store_split32(panel, hf * 4, v);
              }
              if (kTrain && S.epi == EPI_RELU && S.save_row >= 0 && tile_ok)
                p.gate[(size_t)S.save_row * 4 + (size_t)q * p.cap + s] = make_uint2(g01[0], g01[1]);
            } else {
              // bias already in the accumulator: TMEM -> registers -> bf16 -> swizzled panel, one LDTM wait
              uint32_t r0[32], r1[32];
              ptx::tmem_ld32(acc_addr, r0);
              ptx::tmem_ld32(acc_addr + 32u, r1);
              ptx::tmem_ld_wait();
              uint32_t pk[32];
              if (S.epi == EPI_RELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  pk[j] = ptx::pack_bf16x2_relu(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                  pk[16 + j] = ptx::pack_bf16x2_relu(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  pk[j] = ptx::pack_bf16x2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                  pk[16 + j] = ptx::pack_bf16x2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                }
              }
              store_pk16(panel, 0, pk); store_pk16(panel, 4, pk + 16);
              if (kTrain && S.epi == EPI_RELU && S.save_row >= 0 && tile_ok)   // gates of the backward ReLU
                p.gate[(size_t)S.save_row * 4 + (size_t)q * p.cap + s] = make_uint2(gate_bits16(pk), gate_bits16(pk + 16));
              if (S.head) {   // density head: dot of the bf16 activation with the bf16-rounded kernel
                const float4* wd4 = reinterpret_cast<const float4*>(sm.bias + (p.w_dens_off - tail0) + col);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const float4 w = wd4[c];
                  head = fmaf(__uint_as_float(pk[2 * c] << 16), w.x, head);
                  head = fmaf(__uint_as_float(pk[2 * c] & 0xFFFF0000u), w.y, head);
                  head = fmaf(__uint_as_float(pk[2 * c + 1] << 16), w.z, head);
                  head = fmaf(__uint_as_float(pk[2 * c + 1] & 0xFFFF0000u), w.w, head);
                }
              }
            }
            if (S.head) {
              float* part = sm.part + t * 384;
              if (q > 0) part[(q - 1) * 128 + row] = head;
              asm volatile("bar.sync 5, 512;" ::: "memory");
              if (q == 0) {
                const float rd = ((head + part[row]) + part[128 + row]) + part[256 + row] + sm.bias[p.dens_bias_off - tail0];
                raw_keep[t] = rd;
                if (p.raw_c == 1 && valid) p.raw_out[s] = rd;
              }
            }
            publish(S, pi, tile, tile_ok);
            break;
          }
          case EPI_VIEW: {
            if (q < 2) {
              float h0 = 0.f, h1 = 0.f, h2 = 0.f;
              uint32_t g01[2] = {0u, 0u};
#pragma unroll 1
              for (int hf = 0; hf < 2; ++hf) {
                load_acc32(acc_addr + (uint32_t)(hf * 32), v);
                if (valid) {
                  // (prefetching this row before the accumulator wait costs more in registers than the exposed
                  //  L2 latency: measured, profiles/r01_ab_experiments.md)
                  const float4* b4 = reinterpret_cast<const float4*>(p.viewbias + (size_t)(s / p.S) * 128 + col + hf * 32);
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 b = __ldg(b4 + c);
                    v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
                  }
                }
                const float4* wr4 = reinterpret_cast<const float4*>(sm.bias + (p.w_rgb_off - tail0) + (col + hf * 32) * 3);
                if (kSplit) {
#pragma unroll
                  for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {   // 4 columns x 3 channels = 3 float4 per step
                    const float4 wa = wr4[3 * j], wb = wr4[3 * j + 1], wc = wr4[3 * j + 2];
                    const float a0 = v[4 * j], a1 = v[4 * j + 1], a2 = v[4 * j + 2], a3 = v[4 * j + 3];
                    h0 = fmaf(a0, wa.x, h0); h1 = fmaf(a0, wa.y, h1); h2 = fmaf(a0, wa.z, h2);
                    h0 = fmaf(a1, wa.w, h0); h1 = fmaf(a1, wb.x, h1); h2 = fmaf(a1, wb.y, h2);
                    h0 = fmaf(a2, wb.z, h0); h1 = fmaf(a2, wb.w, h1); h2 = fmaf(a2, wc.x, h2);
                    h0 = fmaf(a3, wc.y, h0); h1 = fmaf(a3, wc.z, h1); h2 = fmaf(a3, wc.w, h2);
                  }
                  if (kTrain) { store_split32(panel, hf * 4, v); g01[hf] = gate_bits32f(v); }
                } else {
                  uint32_t pk[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16x2_relu(v[2 * j], v[2 * j + 1]);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {   // 4 columns x 3 channels = 3 float4 per step
                    const float4 wa = wr4[3 * j], wb = wr4[3 * j + 1], wc = wr4[3 * j + 2];
                    const float a0 = __uint_as_float(pk[2 * j] << 16), a1 = __uint_as_float(pk[2 * j] & 0xFFFF0000u);
                    const float a2 = __uint_as_float(pk[2 * j + 1] << 16), a3 = __uint_as_float(pk[2 * j + 1] & 0xFFFF0000u);
                    h0 = fmaf(a0, wa.x, h0); h1 = fmaf(a0, wa.y, h1); h2 = fmaf(a0, wa.z, h2);
                    h0 = fmaf(a1, wa.w, h0); h1 = fmaf(a1, wb.x, h1); h2 = fmaf(a1, wb.y, h2);
                    h0 = fmaf(a2, wb.z, h0); h1 = fmaf(a2, wb.w, h1); h2 = fmaf(a2, wc.x, h2);
                    h0 = fmaf(a3, wc.y, h0); h1 = fmaf(a3, wc.z, h1); h2 = fmaf(a3, wc.w, h2);
                  }
                  if (kTrain) { store_pk16(panel, hf * 4, pk); g01[hf] = gate_bits16(pk); }
                }
              }
              if (kTrain && S.save_row >= 0 && tile_ok)
                p.gate[(size_t)S.save_row * 4 + (size_t)q * p.cap + s] = make_uint2(g01[0], g01[1]);
              float* part = sm.part + t * 384;
              if (q == 1) { part[row * 3] = h0; part[row * 3 + 1] = h1; part[row * 3 + 2] = h2; }
              asm volatile("bar.sync 6, 256;" ::: "memory");
              if (q == 0 && valid) {
                float4 o;
                o.x = raw_keep[t];
                o.y = h0 + part[row * 3] + sm.bias[p.rgb_bias_off - tail0];
                o.z = h1 + part[row * 3 + 1] + sm.bias[p.rgb_bias_off + 1 - tail0];
                o.w = h2 + part[row * 3 + 2] + sm.bias[p.rgb_bias_off + 2 - tail0];
                reinterpret_cast<float4*>(p.raw_out)[s] = o;
              }
              publish(S, pi, tile, tile_ok);
            }
            break;
          }
          case EPI_BWD_LINEAR: case EPI_BWD_RELU: case EPI_BWD_RELU_D: {
            if (kSplit) {
#pragma unroll 1
              for (int hf = 0; hf < 2; ++hf) {
                load_acc32(acc_addr + (uint32_t)(hf * 32), v);
                if (S.epi == EPI_BWD_RELU_D) {
                  const float* wd = sm.bias + (p.w_dens_off - tail0) + col + hf * 32;
#pragma unroll
                  for (int c = 0; c < 32; ++c) v[c] = fmaf(dd, wd[c], v[c]);
                }
                if (S.epi != EPI_BWD_LINEAR) apply_gate32f(hf == 0 ? gate.x : gate.y, v);
                store_split32(panel, hf * 4, v);
              }
              publish(S, pi, tile, tile_ok);
              break;
            }
            uint32_t r0[32], r1[32];
            ptx::tmem_ld32(acc_addr, r0);
            ptx::tmem_ld32(acc_addr + 32u, r1);
            ptx::tmem_ld_wait();
            if (S.epi == EPI_BWD_RELU_D) {
              const float4* w4 = reinterpret_cast<const float4*>(sm.bias + (p.w_dens_off - tail0) + col);
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 wa = w4[c], wb = w4[8 + c];
                r0[c * 4 + 0] = __float_as_uint(fmaf(dd, wa.x, __uint_as_float(r0[c * 4 + 0])));
                r0[c * 4 + 1] = __float_as_uint(fmaf(dd, wa.y, __uint_as_float(r0[c * 4 + 1])));
                r0[c * 4 + 2] = __float_as_uint(fmaf(dd, wa.z, __uint_as_float(r0[c * 4 + 2])));
                r0[c * 4 + 3] = __float_as_uint(fmaf(dd, wa.w, __uint_as_float(r0[c * 4 + 3])));
                r1[c * 4 + 0] = __float_as_uint(fmaf(dd, wb.x, __uint_as_float(r1[c * 4 + 0])));
                r1[c * 4 + 1] = __float_as_uint(fmaf(dd, wb.y, __uint_as_float(r1[c * 4 + 1])));
                r1[c * 4 + 2] = __float_as_uint(fmaf(dd, wb.z, __uint_as_float(r1[c * 4 + 2])));
                r1[c * 4 + 3] = __float_as_uint(fmaf(dd, wb.w, __uint_as_float(r1[c * 4 + 3])));
              }
            }
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pk[j] = ptx::pack_bf16x2(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
              pk[16 + j] = ptx::pack_bf16x2(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
            }
            if (S.epi != EPI_BWD_LINEAR) { apply_gate16(gate.x, pk); apply_gate16(gate.y, pk + 16); }
            store_pk16(panel, 0, pk); store_pk16(panel, 4, pk + 16);
            publish(S, pi, tile, tile_ok);
            break;
          }
          case EPI_BWD_START: {
            // dV = W_rgb^T d_rgb (CUDA cores), gated by the saved view activation; 128 columns: groups 0, 1
            if (q < 2) {
              const float4 dr = valid ? reinterpret_cast<const float4*>(p.d_raw)[s] : make_float4(0, 0, 0, 0);
              float d0 = dr.y, d1 = dr.z, d2 = dr.w;
              if (!kSplit) {   // bf16-round the head gradient once so that dgrad (here) and wgrad (tensor cores) agree
                d0 = __bfloat162float(__float2bfloat16(d0)); d1 = __bfloat162float(__float2bfloat16(d1));
                d2 = __bfloat162float(__float2bfloat16(d2));
              }
              if (q == 0 && tile_ok) {   // padding rows of a real tile get zeros
                uint4* dst = reinterpret_cast<uint4*>(p.drgb_out + (size_t)s * kHeadCols);
                if (kSplit) {
                  uint32_t h01, l01, h23, l23;
                  split_pair(dr.y, dr.z, h01, l01); split_pair(dr.w, dr.x, h23, l23);
                  dst[0] = make_uint4(h01, h23, 0u, 0u);
                  reinterpret_cast<uint4*>(p.drgb_out + ((size_t)p.lo_drgb_rows + s) * kHeadCols)[0] = make_uint4(l01, l23, 0u, 0u);
                } else {
                  dst[0] = make_uint4(ptx::pack_bf16x2(dr.y, dr.z), ptx::pack_bf16x2(dr.w, dr.x), 0u, 0u);
                }
              }
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const float* wr = sm.bias + (p.w_rgb_off - tail0) + (col + hf * 32) * 3;
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = d0 * wr[c * 3] + d1 * wr[c * 3 + 1] + d2 * wr[c * 3 + 2];
                if (kSplit) {
                  apply_gate32f(hf == 0 ? gate.x : gate.y, v);
                  store_split32(panel, hf * 4, v);
                } else {
                  uint32_t pk[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
                  apply_gate16(hf == 0 ? gate.x : gate.y, pk);
                  store_pk16(panel, hf * 4, pk);
                }
              }
              publish(S, pi, tile, tile_ok);
            } else if (!S.no_signal && group_leader) {
              // the tile barrier counts every group of both CTAs: groups without columns in this op arrive at once
              ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(pready_u32 + (pi & ~3) * 8, 0));
              if (kSplit) ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(pready_u32 + 4 * 8, 0));
            }
            break;
          }
          case EPI_BWD_START_PROP: {
            const float dd0 = valid ? p.d_raw[s] : 0.f;
            const float ddq = kSplit ? dd0 : __bfloat162float(__float2bfloat16(dd0));
            if (q == 0 && tile_ok) {
              uint4* dst = reinterpret_cast<uint4*>(p.drgb_out + (size_t)s * kHeadCols);
              if (kSplit) {
                uint32_t hh, ll;
                split_pair(0.f, dd0, hh, ll);
                dst[0] = make_uint4(0u, hh, 0u, 0u);
                reinterpret_cast<uint4*>(p.drgb_out + ((size_t)p.lo_drgb_rows + s) * kHeadCols)[0] = make_uint4(0u, ll, 0u, 0u);
              } else {
                dst[0] = make_uint4(0u, ptx::pack_bf16x2(0.f, dd0), 0u, 0u);
              }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const float* wd = sm.bias + (p.w_dens_off - tail0) + col + hf * 32;
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = ddq * wd[c];
              if (kSplit) {
                apply_gate32f(hf == 0 ? gate.x : gate.y, v);
                store_split32(panel, hf * 4, v);
              } else {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
                apply_gate16(hf == 0 ? gate.x : gate.y, pk);
                store_pk16(panel, hf * 4, pk);
              }
            }
            publish(S, pi, tile, tile_ok);
            break;
          }
          default: break;
        }
        if (S.last_epi) {
          // every warp of the group is past its TMEM reads / panel writes before the panels are released
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (group_leader) {
            ptx::mbar_arrive(&sm.epi_done[t]);
            if (kSplit) ptx::mbar_arrive(&sm.epi_done[1]);
          }
        }
      }
    }
    if (group_leader) ptx::tma_wait_group<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();   // neither CTA retires (or frees TMEM) while the pair's MMAs / arrivals are in flight
  if (warp == kMma) ptx::tmem_dealloc_cg2(tmem_base, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side: segment programs
// ------------------------------------------------------------------------------------------
namespace {

// split-precision program: every MMA segment is issued against the hi weights and, reading the same A panels, against
// the lo weights (`lo_rows` further down in the packed weight tensor); the epilogue moves to the second issue
std::vector<PpSeg> expand_split(const std::vector<PpSeg>& prog, int lo_rows) {
  std::vector<PpSeg> out;
  for (const PpSeg& s : prog) {
    if (s.kps == 0) { out.push_back(s); continue; }
    PpSeg hi = s, lo = s;
    hi.epi = EPI_NONE; hi.head = 0; hi.last_epi = 0; hi.no_signal = 0; hi.save_row = -1; hi.mask_row = -1;
    lo.w_row = s.w_row + lo_rows; lo.accumulate = 1; lo.bias_idx = -1; lo.a_feat = 0; lo.no_wait = 1;
    out.push_back(hi); out.push_back(lo);
  }
  return out;
}

int finish_program(std::vector<PpSeg>* prog, const char* what) {
  for (size_t i = 0; i < prog->size(); ++i) {
    size_t j = (i + 1) % prog->size();
    while ((*prog)[j].kps == 0) j = (j + 1) % prog->size();
    (*prog)[i].feat_next = (*prog)[j].a_feat;
  }
  HUGS_REQUIRE((int)prog->size() <= kMaxSegs, "chain schedule (%s): too many segments (%zu)", what, prog->size());
  // a feature refill must never directly follow an epilogue of the same tile (the epilogue writes the panels)
  for (size_t i = 1; i < prog->size(); ++i)
    HUGS_REQUIRE(!((*prog)[i].a_feat && (*prog)[i - 1].epi != EPI_NONE), "unsupported segment order at %zu (%s)", i, what);
  return HUGS_OK;
}

}  // namespace

int pp_build(hugs_handle* h, const MlpViews& mv, TcMlp* m) {
  const hugs_model_desc& d = h->d;
  const int D = mv.depth;
  auto base_seg = [] {
    PpSeg s{};
    s.kps = 4; s.n_halves = 2; s.epi = EPI_NONE; s.save_row = -1; s.mask_row = -1; s.bias_idx = -1;
    return s;
  };
  // ---- forward ----
  // features enter in segments of <= 4 K panels: 8 panels for the 504 IPE columns, 2 for a 93-column point encoding
  const int feat_parts = (h->feat_panels + 3) / 4;
  bool cat = false;
  for (int i = 0; i < D; ++i) {
    const auto& P = m->pack[i];
    const bool last = i == D - 1;
    auto finish = [&](PpSeg& s) {
      s.epi = EPI_RELU; s.bias_off = P.bias_off; s.save_row = i;
      if (last) { s.head = 1; if (!mv.has_rgb) s.no_signal = 1; }
    };
    if (i == 0) {
      for (int part = 0; part < feat_parts; ++part) {
        PpSeg s = base_seg();
        s.kps = std::min(4, h->feat_panels - 4 * part);
        s.a_feat = 1; s.feat_col0 = part * 256; s.w_row = P.row0; s.w_col0 = part * 256; s.accumulate = part > 0;
        if (part == 0) s.bias_idx = i;
        if (part == feat_parts - 1) finish(s);
        m->pp_fwd.push_back(s);
      }
    } else {
      PpSeg s = base_seg();
      s.w_row = P.row0; s.w_col0 = 0; s.bias_idx = i;
      if (!cat) finish(s);
      m->pp_fwd.push_back(s);
      if (cat) {
        for (int part = 0; part < feat_parts; ++part) {
          PpSeg f = base_seg();
          f.kps = std::min(4, h->feat_panels - 4 * part);
          f.a_feat = 1; f.feat_col0 = part * 256; f.w_row = P.row0; f.w_col0 = kW + part * 256; f.accumulate = 1;
          if (part == feat_parts - 1) finish(f);
          m->pp_fwd.push_back(f);
        }
      }
    }
    cat = (i % d.skip_layer == 0 && i > 0);
  }
  HUGS_REQUIRE(!cat, "tensor-core path: a skip connection into the heads is not supported (depth %d, skip %d)", D,
               d.skip_layer);
  if (mv.has_rgb) {
    PpSeg b = base_seg();                       // bottleneck (linear)
    b.w_row = m->pack[D + 1].row0; b.epi = EPI_LINEAR; b.bias_off = m->pack[D + 1].bias_off; b.save_row = D;
    b.bias_idx = D;
    m->pp_fwd.push_back(b);
    PpSeg v = base_seg();                       // view layer (N = 128) + rgb head
    v.n_halves = 1; v.w_row = m->pack[D + 2].row0; v.epi = EPI_VIEW; v.save_row = D + 1; v.no_signal = 1;
    m->pp_fwd.push_back(v);
  }
  m->pp_fwd.back().last_epi = 1;

  // ---- backward (dgrad chain) ----
  auto mma_seg = [&](int kps, int w_row, int epi, int save_slot, int mask_slot) {
    PpSeg s = base_seg();
    s.kps = kps; s.w_row = w_row; s.epi = epi; s.save_row = save_slot; s.mask_row = mask_slot;
    return s;
  };
  PpSeg st = base_seg();
  st.kps = 0;
  if (mv.has_rgb) {
    st.epi = EPI_BWD_START; st.save_row = D + 1; st.mask_row = D + 1;
    m->pp_bwd.push_back(st);
    m->pp_bwd.push_back(mma_seg(2, m->pack[D + 2].brow0, EPI_BWD_LINEAR, D, -1));
    m->pp_bwd.push_back(mma_seg(4, m->pack[D + 1].brow0, EPI_BWD_RELU_D, D - 1, D - 1));
  } else {
    st.epi = EPI_BWD_START_PROP; st.save_row = D - 1; st.mask_row = D - 1;
    m->pp_bwd.push_back(st);
  }
  for (int l = D - 1; l >= 1; --l) m->pp_bwd.push_back(mma_seg(4, m->pack[l].brow0, EPI_BWD_RELU, l - 1, l - 1));
  m->pp_bwd.back().no_signal = 1;
  m->pp_bwd.back().last_epi = 1;

  if (h->tc->split) {
    m->pp_fwd = expand_split(m->pp_fwd, m->rows_f);
    m->pp_bwd = expand_split(m->pp_bwd, m->rows_b);
  }
  int rc;
  if ((rc = finish_program(&m->pp_fwd, "forward")) || (rc = finish_program(&m->pp_bwd, "backward"))) return rc;
  return HUGS_OK;
}

int pp_init(hugs_handle* h) {
  (void)h;
  HUGS_CUDA(cudaFuncSetAttribute(mlp_pp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpSmemBytes));
  HUGS_CUDA(cudaFuncSetAttribute(mlp_pp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpSmemBytes));
  HUGS_CUDA(cudaFuncSetAttribute(mlp_pp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpSmemBytes));
  HUGS_CUDA(cudaFuncSetAttribute(mlp_pp_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPpSmemBytes));
  return HUGS_OK;
}

// direction: 0 forward (render), 1 forward (training, saves activations), 2 backward dgrad chain
int pp_launch(hugs_handle* h, int level, int n_rays, int direction, cudaStream_t st) {
  TcState* tc = h->tc;
  const hugs_model_desc& d = h->d;
  const bool is_prop = level < d.num_levels - 1;
  const TcMlp& m = is_prop ? tc->prop : tc->nerf;
  const MlpViews& mv = is_prop ? h->prop : h->nerf;
  const int S = h->samples(level);
  const int n_samples = n_rays * S;
  const int n_tiles = (n_samples + kTileM - 1) / kTileM;
  const int cap = tc->cap[level], srow = tc->save_row0[level];
  const std::vector<PpSeg>& prog = direction == 2 ? m.pp_bwd : m.pp_fwd;
  PpParams p;
  memset(&p, 0, sizeof(p));
  p.map_w = direction == 2 ? m.map_wn128 : m.map_wt128;
  p.map_w_half = m.map_wt64;         // only forward programs contain N = 128 segments
  p.map_feat = tc->map_feat;
  p.map_save = direction == 2 ? tc->map_dz : tc->map_act;
  p.n_segs = (int)prog.size();
  for (int i = 0; i < p.n_segs; ++i) {
    p.segs[i] = prog[i];
    if (direction == 0) p.segs[i].save_row = -1;
    if (p.segs[i].save_row >= 0) p.segs[i].save_row = srow + p.segs[i].save_row * cap;
    if (p.segs[i].mask_row >= 0) p.segs[i].mask_row = srow + p.segs[i].mask_row * cap;
    if (p.segs[i].a_feat) p.any_feat = 1;
    HUGS_REQUIRE(direction != 2 || p.segs[i].n_halves == 2, "backward program must be N = 256 throughout");
  }
  const int tiles_per_unit = tc->split ? 2 : 4;
  p.n_tiles = n_tiles; p.n_units = (n_tiles + tiles_per_unit - 1) / tiles_per_unit; p.n_samples = n_samples; p.S = S;
  p.feat_row0 = tc->feat_row0[level];
  p.bias = m.bias; p.bias_floats = m.bias_floats; p.viewbias = tc->viewbias;
  p.raw_out = h->raw[level]; p.raw_c = is_prop ? 1 : 4;
  p.d_raw = h->d_raw[level]; p.drgb_out = tc->drgb;
  p.w_dens_off = m.w_dens_off; p.w_rgb_off = m.w_rgb_off;
  p.dens_bias_off = m.pack[mv.depth].bias_off;
  p.gate = tc->gate; p.cap = cap;
  for (int i = 0; i < p.n_segs; ++i)
    if (p.segs[i].epi != EPI_NONE) p.epi_seg[p.n_epi++] = i;
  p.bias_img = m.bias_img;
  p.bias_tail0 = m.pack[mv.depth].bias_off;
  HUGS_REQUIRE(m.bias_floats - p.bias_tail0 <= kBiasTailFloats, "head table too large for the chain kernel");
  p.rgb_bias_off = mv.has_rgb ? m.pack[mv.depth + 3].bias_off : 0;
  p.lo_feat_rows = tc->total_feat_rows; p.lo_save_rows = tc->total_save_rows; p.lo_drgb_rows = tc->drgb_rows;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * std::min(p.n_units, tc->num_sms / 2));
  cfg.blockDim = dim3(kPpThreads);
  cfg.dynamicSmemBytes = kPpSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  if (tc->split) {
    if (direction == 0) HUGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_pp_kernel<false, true>, p));
    else HUGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_pp_kernel<true, true>, p));
  } else {
    if (direction == 0) HUGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_pp_kernel<false, false>, p));
    else HUGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_pp_kernel<true, false>, p));
  }
  ++g_launch_count;
  return HUGS_OK;
}

}  // namespace hugs
