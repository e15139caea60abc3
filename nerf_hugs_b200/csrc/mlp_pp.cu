// Ping-pong MLP chain kernel (sm_100a): two 128-sample tiles in flight per CTA.
//
// While the epilogue warps turn tile A's accumulator into the next layer's A operand, the tensor core
// already runs tile B's MMAs of the same layer, and vice versa — MMA, epilogue and TMA latencies of one
// tile hide behind the other tile (profiles/r01_microbench_and_counters.md shows they were serialised
// in the single-tile kernel).
//
//   shared memory  panels[8]  : tile t owns panels 4t..4t+3 = its 128 x 256 bf16 activation, updated IN PLACE
//                               (a layer's epilogue starts only after all MMAs of that layer have completed)
//                  ring[5]    : 16 KB weight stages ([128 out x 64 in] bf16, SWIZZLE_128B), TMA-fed
//                  bias table, head partial sums, mbarriers
//   tensor memory  tile t accumulates in columns [256t, 256t+256)
//   warps          0-15 epilogue (group q = warp/4 owns output columns [64q, 64q+64)),
//                  16 weight producer, 17 feature producer, 18 MMA issuer
//
// A layer is a sequence of *segments* of <= 4 K-panels (K <= 256).  IPE features (layer 0 and the skip
// connection) are TMA-loaded straight into the tile's own panels once the previous segment has been
// consumed; the N = 1 / N = 3 heads are evaluated on CUDA cores inside the producing epilogue.
//
// kCg2 = true runs the same program on a CTA *pair* (cluster of 2, tcgen05 cta_group::2): every MMA is
// M = 256 (128 rows per CTA) x N = 256, each CTA stages only its half of the weight rows, so per-SM weight
// traffic (L2 -> shared memory, and shared memory -> tensor core) is halved.  The leader CTA issues all MMAs;
// barriers the issuer waits on live in the leader and receive the peer's TMA completions / epilogue arrivals
// through the cluster address space, barriers signalled by the tensor core are multicast to both CTAs.
// In this mode additionally: a weight stage (one K panel) is fetched once per unit and used by both tiles; layer
// biases are applied by a K = 16 MMA (ones x [bias_hi, bias_lo]) that initialises the accumulator; the epilogue is
// one TMEM wait + pack + eight 16-byte stores; the forward stores 64-bit ReLU gate masks which the backward program
// loads one epilogue ahead; one barrier per tile collects the elected arrivals of all epilogue groups of both CTAs.
// DESIGN.md ("Tensor-core kernel structure") has the rationale and profiles/r01_ab_experiments.md the measurements.
#include <algorithm>

#include "tc_device.cuh"
#include "tc_internal.h"

namespace hugs {
namespace {

constexpr int kPpThreads = (kEpiGroups * 4 + 3) * 32;
#ifdef HUGS_PP_COUNTERS       // per-role cycle counters (scripts/chain_counters.py); off by default: they cost registers
constexpr bool kCounters = true;
#else
constexpr bool kCounters = false;
#endif
// CTA-pair kernel: ONE barrier per tile (panel_ready[4t]) collects the elected arrivals of all 4 epilogue groups of both
// CTAs.  A try_wait costs the issuing thread ~350 cycles under load even on a completed phase (shared-memory pipe queue),
// and four of them in a row left the tensor pipe idle for ~1 k cycles per tile and layer (event trace,
// profiles/r01_ab_experiments.md).
constexpr int kPanelArrivals = 8;
#ifdef HUGS_EXP_MMA_ONLY      // timing experiment (results invalid): weight producer + MMA issuer only, no epilogue dependency
constexpr bool kMmaOnly = true;
#else
constexpr bool kMmaOnly = false;
#endif
constexpr int kPartFloats = 768;     // head partial sums: [2 tiles][3][128]
constexpr int kPpSmemBytes = 1024 + (kNumPanels + kStages) * kPanelBytes + kBiasTab * 4 + kPartFloats * 4 + 512;
// CTA-pair kernel: the fp32 table shrinks to the head weights / head biases, the layer biases become MMA operands
constexpr int kBiasImgBytes = kBiasChunks * kBiasChunkElems * 2;
constexpr int kOnesBytes = 4 * 256;   // variant v: core matrix with 1.0 in K columns 2v, 2v+1, then a zero core matrix
constexpr int kCg2SmemBytes = 1024 + (kNumPanels + kStages) * kPanelBytes + kBiasTailFloats * 4 + kPartFloats * 4 +
                              kBiasImgBytes + kOnesBytes + 512;
static_assert(kCg2SmemBytes <= kPpSmemBytes, "CTA-pair layout must fit the common allocation");
static_assert(kPpSmemBytes <= 232448, "shared memory budget");

struct PpSmem {
  uint8_t* panels; uint8_t* ring; float* bias; float* part;
  uint8_t* bias_img; uint8_t* ones;     // CTA-pair kernel only
  uint64_t *full, *empty, *panel_ready, *feat_ready, *acc_full, *consumed, *epi_done;
  uint32_t* tmem_ptr;
};

template <bool kCg2>
__device__ __forceinline__ PpSmem pp_carve(uint8_t* raw) {
  PpSmem s;
  // offset arithmetic on the __shared__ symbol (not an integer round trip) keeps the shared address space, so the
  // accesses below compile to LDS/STS instead of generic loads and stores
  uint8_t* base = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  s.panels = base;
  s.ring = base + kNumPanels * kPanelBytes;
  s.bias = reinterpret_cast<float*>(s.ring + kStages * kPanelBytes);
  s.part = s.bias + (kCg2 ? kBiasTailFloats : kBiasTab);
  s.bias_img = reinterpret_cast<uint8_t*>(s.part + kPartFloats);
  s.ones = s.bias_img + (kCg2 ? kBiasImgBytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.ones + (kCg2 ? kOnesBytes : 0));
  s.full = bars; s.empty = bars + kStages;
  s.panel_ready = bars + 2 * kStages;        // [8]
  s.feat_ready = s.panel_ready + 8;          // [8]
  s.acc_full = s.feat_ready + 8;             // [2]
  s.consumed = s.acc_full + 2;               // [8]  per (tile, K panel): the MMAs that read the panel have completed
  s.epi_done = s.consumed + 8;               // [2]
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.epi_done + 2);
  return s;
}

template <bool kTrain, bool kCg2>
__global__ void __launch_bounds__(kPpThreads, 1) mlp_pp_kernel(const __grid_constant__ PpParams p) {
  extern __shared__ uint8_t smem_raw[];
  PpSmem sm = pp_carve<kCg2>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWProd = kEpiGroups * 4, kFProd = kWProd + 1, kMma = kWProd + 2;
  // work distribution: a *unit* is one pass of the segment program over `kCtas` x 2 tiles
  constexpr int kCtas = kCg2 ? 2 : 1;
  const int rank = kCg2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit0 = kCg2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = kCg2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  const uint32_t panels_u32 = ptx::smem_u32(sm.panels), ring_u32 = ptx::smem_u32(sm.ring);
  const uint32_t full_u32 = ptx::smem_u32(sm.full), empty_u32 = ptx::smem_u32(sm.empty);
  const uint32_t pready_u32 = ptx::smem_u32(sm.panel_ready), fready_u32 = ptx::smem_u32(sm.feat_ready);
  const uint32_t accfull_u32 = ptx::smem_u32(sm.acc_full), consumed_u32 = ptx::smem_u32(sm.consumed);
  const uint32_t epidone_u32 = ptx::smem_u32(sm.epi_done);
  const uint32_t ones_u32 = ptx::smem_u32(sm.ones), biasimg_u32 = ptx::smem_u32(sm.bias_img);
  // counters build: event trace of cluster 0 (leader CTA): (code, clock) pairs; code = ev << 16 | si << 8 | t << 4 | q
  __shared__ unsigned int trace_n;
  if (threadIdx.x == 0) trace_n = 0;
  auto trace = [&](int ev, int si_, int t_, int q_) {
    if (kCounters && p.dbg != nullptr && blockIdx.x < 2) {      // both CTAs of cluster 0, %globaltimer (ns) as common clock
      const unsigned int i = atomicAdd(&trace_n, 1u);
      if (i < 2048u) {
        long long* e = p.dbg + 148 * 16 + 74 * 60 + 2 * (i + 2048u * blockIdx.x);
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        e[0] = (long long)((blockIdx.x << 24) | (ev << 16) | (si_ << 8) | (t_ << 4) | q_); e[1] = (long long)gt;
      }
    }
  };
  __shared__ long long ts_acc[2];     // counters build: clock at which epilogue group 0 saw acc_full of tile t
  __shared__ long long ts_pub[2][4];  // counters build: clock at which group q's leader arrived on panel_ready of tile t

  if (warp == kWProd && lane == 0) {
    ptx::prefetch_tmap(&p.map_w); ptx::prefetch_tmap(&p.map_feat); ptx::prefetch_tmap(&p.map_save);
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(&sm.full[i], 1); ptx::mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(&sm.panel_ready[i], kCg2 ? kPanelArrivals : 128); ptx::mbar_init(&sm.feat_ready[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&sm.acc_full[i], 1); ptx::mbar_init(&sm.epi_done[i], kEpiGroups); }
    for (int i = 0; i < 8; ++i) ptx::mbar_init(&sm.consumed[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == kMma) { if (kCg2) ptx::tmem_alloc_cg2(sm.tmem_ptr, 512); else ptx::tmem_alloc(sm.tmem_ptr, 512); }
  if (kCg2) {
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
  } else {
    for (int i = threadIdx.x; i < p.bias_floats; i += kPpThreads) sm.bias[i] = p.bias[i];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (kCg2) ptx::cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA signal
  ptx::tc_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kWProd) {
    // =============================== weight producer ===============================
    // cg2: this CTA stages rows [rank * N/2, (rank + 1) * N/2) of every weight tile; the transaction bytes of
    // both CTAs are accounted on the leader's `full` barrier
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = unit0; unit < p.n_units; unit += unit_stride) {
        for (int si = 0; si < p.n_segs; ++si) {
          const PpSeg& S = p.segs[si];
          if (S.kps == 0) continue;
          const int kps = S.kps, n_halves = S.n_halves, w_col0 = S.w_col0;
          if (kCg2) {
            // one stage per K panel, shared by the unit's two tiles (the issuer releases it after tile 1)
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
          } else {
            const int w_row = S.w_row;
            for (int t = 0; t < 2; ++t)
              for (int kp = 0; kp < kps; ++kp)
                for (int h = 0; h < n_halves; ++h) {
                  ptx::mbar_wait_u32(empty_u32 + stage * 8, phase ^ 1);
                  ptx::mbar_expect_tx_u32(full_u32 + stage * 8, kPanelBytes);
                  ptx::tma_load_2d_u32(ring_u32 + stage * kPanelBytes, &p.map_w, full_u32 + stage * 8,
                                       w_col0 + kp * 64, w_row + h * 128);
                  if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
          }
        }
      }
    }
  } else if (warp == kFProd) {
    // =============================== feature producer ===============================
    // Walks the MMA segments in issue order and waits for every `consumed` phase exactly once (no phase is
    // ever skipped, so parity waits cannot alias); refills a tile's panels with IPE feature columns as soon
    // as the segment that last read those panels has completed.
    if (lane == 0 && p.any_feat && !kMmaOnly) {
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
            const int row = p.feat_row0 + ((unit * 2 + t) * kCtas + rank) * kTileM;
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
                if (kCg2) {
                  if (rank == 0) ptx::mbar_expect_tx_u32(bar, 2 * kPanelBytes);
                  ptx::tma_load_2d_cg2(panels_u32 + idx * kPanelBytes, &p.map_feat, ptx::mapa_u32(bar, 0),
                                       S.feat_col0 + kp * 64, row);
                } else {
                  ptx::mbar_expect_tx_u32(bar, kPanelBytes);
                  ptx::tma_load_2d_u32(panels_u32 + idx * kPanelBytes, &p.map_feat, bar, S.feat_col0 + kp * 64, row);
                }
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
      const uint32_t idesc128 = ptx::make_idesc_bf16(kCg2 ? 256 : 128, 128, 0, 0);
      const uint32_t idesc256 = ptx::make_idesc_bf16(kCg2 ? 256 : 128, 256, 0, 0);
      // the issuing thread never reads the data behind these barriers itself (the tensor core does, behind
      // tcgen05.fence::after_thread_sync), so CTA-scope waits suffice also for the peer's arrivals
      // (a polling wait with a short suspend hint instead of the hardware-suspended one measured the same and only adds
      //  shared-memory traffic: profiles/r01_ab_experiments.md)
      auto wait = [](uint32_t bar, uint32_t parity) { ptx::mbar_wait_u32(bar, parity); };
      auto commit = [](uint32_t bar) { if (kCg2) ptx::mma_commit_mc2_u32(bar); else ptx::mma_commit_u32(bar); };
      int stage = 0; uint32_t phase = 0;
      uint32_t wait_phase = 0;   // bits 0-7 panel_ready, 8-15 feat_ready
      long long c_panel = 0, c_feat = 0, c_full = 0;
      long long c_elat = 0, n_elat = 0;   // acc_full seen by epilogue group 0 -> all panels ready (as seen by the issuer)
      long long c_pubq[4] = {0, 0, 0, 0}, c_after = 0;   // ... -> local group q arrived; last local arrival -> issuer proceeds
      long long c_seg[kMaxSegs][3];     // counters build: per-segment waits (panel, full, feat)
      if (kCounters) for (int i = 0; i < kMaxSegs; ++i) c_seg[i][0] = c_seg[i][1] = c_seg[i][2] = 0;
      const long long c_start = clock64();
      const bool dbg = kCounters && p.dbg != nullptr;
      for (int unit = unit0; unit < p.n_units; unit += unit_stride) {
        for (int si = 0; si < p.n_segs; ++si) {
          const PpSeg& S = p.segs[si];
          if (S.kps == 0) continue;
          const int kps = S.kps, n_halves = S.n_halves, a_feat = S.a_feat, acc0 = S.accumulate;
          const bool has_epi = S.epi != EPI_NONE;
          for (int t = 0; t < 2; ++t) {
            trace(1, si, t, 0);       // issuer reaches (segment, tile)
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * 256);
            if (!a_feat && !kMmaOnly) {
              // in-place accumulator: every epilogue group must have drained the previous layer before the
              // first MMA of this one overwrites it, so wait for all consumed panels up front
              const long long c0 = dbg ? clock64() : 0;
              for (int kp = 0; kp < (kCg2 ? 1 : kps); ++kp) {
                const uint32_t idx = (uint32_t)(t * 4 + kp);
                wait(pready_u32 + idx * 8, (wait_phase >> idx) & 1u);
                wait_phase ^= 1u << idx;
                trace(6, si, t, kp);     // panel kp ready
              }
              if (dbg) {
                const long long now = clock64(), dt = now - c0;
                c_panel += dt; c_seg[si][0] += dt;
                const long long ta = *(volatile long long*)&ts_acc[t];
                c_elat += now - ta; ++n_elat;
                long long mx = 0;
                for (int g = 0; g < 4; ++g) {
                  const long long tp = *(volatile long long*)&ts_pub[t][g];
                  c_pubq[g] += tp - ta; if (tp > mx) mx = tp;
                }
                c_after += now - mx;
              }
              ptx::tc_fence_after();
            }
            trace(2, si, t, 0);       // panel waits done
            bool bias_pending = kCg2 && S.bias_idx >= 0;
            for (int kp = 0; kp < kps; ++kp) {
              if (a_feat && !kMmaOnly) {
                const uint32_t idx = (uint32_t)(8 + t * 4 + kp);
                const long long c0 = dbg ? clock64() : 0;
                wait(fready_u32 + (t * 4 + kp) * 8, (wait_phase >> idx) & 1u);
                if (dbg) { const long long dt = clock64() - c0; c_feat += dt; c_seg[si][2] += dt; }
                wait_phase ^= 1u << idx;
                ptx::tc_fence_after();
              }
              const uint64_t da = ptx::desc_from(kDescHi, panels_u32 + (t * 4 + kp) * kPanelBytes);
              uint32_t accum = (acc0 || kp > 0) ? 1u : 0u;
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
              if (kCg2) {
                // one stage = this K panel of all N columns (each CTA holds its half of the rows); tile 0 waits
                // for it, tile 1 reuses it and releases it
                int st_k = stage + kp; uint32_t ph_k = phase;
                if (st_k >= kStages) { st_k -= kStages; ph_k ^= 1; }
                if (t == 0) {
                  const long long c0 = dbg ? clock64() : 0;
                  wait(full_u32 + st_k * 8, ph_k);
                  if (dbg) { const long long dt = clock64() - c0; c_full += dt; c_seg[si][1] += dt; }
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
              } else {
                for (int h = 0; h < n_halves; ++h) {
                  const long long c0 = dbg ? clock64() : 0;
                  wait(full_u32 + stage * 8, phase);
                  if (dbg) c_full += clock64() - c0;
                  ptx::tc_fence_after();
                  const uint64_t db = ptx::desc_from(kDescHi, ring_u32 + stage * kPanelBytes);
                  const uint32_t d = d_tmem + (uint32_t)(h * 128);
                  ptx::mma_bf16_ss(d, da, db, idesc128, accum);
                  ptx::mma_bf16_ss(d, da + 2, db + 2, idesc128, 1u);
                  ptx::mma_bf16_ss(d, da + 4, db + 4, idesc128, 1u);
                  ptx::mma_bf16_ss(d, da + 6, db + 6, idesc128, 1u);
                  commit(empty_u32 + stage * 8);
                  if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (S.feat_next) commit(consumed_u32 + (t * 4 + kp) * 8);
              }
            }
            if (has_epi) commit(accfull_u32 + t * 8);
            trace(3, si, t, 0);       // all MMAs of (segment, tile) issued
          }
          if (kCg2) {
            stage += kps;
            if (stage >= kStages) { stage -= kStages; phase ^= 1; }
          }
        }
      }
      if (dbg) {
        long long* d = p.dbg + unit0 * 16;
        d[4] = clock64() - c_start; d[5] = c_panel; d[6] = c_full; d[7] = c_feat;
        if (kCounters) {
          long long* e = p.dbg + 148 * 16 + unit0 * (kMaxSegs * 3);
          for (int i = 0; i < kMaxSegs - 3; ++i) { e[i * 3] = c_seg[i][0]; e[i * 3 + 1] = c_seg[i][1]; e[i * 3 + 2] = c_seg[i][2]; }
          e[57] = c_elat; e[58] = n_elat; e[59] = c_after;
          for (int g = 0; g < 4; ++g) e[53 + g] = c_pubq[g];
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
    const int tail0 = kCg2 ? p.bias_tail0 : 0;   // sm.bias[i - tail0] == bias table entry i
    // development counters (group 0 / group 3, first lane of the leader CTA)
    const bool dbg_t = kCounters && p.dbg != nullptr && rank == 0 && (threadIdx.x == 0 || threadIdx.x == 3 * 128);
    long long c_acc = 0, c_work = 0, c_pub = 0, c_t0 = 0, c_t1 = 0, c_ld = 0, c_st = 0, c_view = 0, c_head = 0;
    const long long c_epi_start = clock64();

    // make the freshly written panel visible to the async proxy, optionally TMA-store it (its smem read is
    // complete before panel_ready completes, so later in-place overwrites / feature refills are safe),
    // then hand it to the MMA issuer
    auto publish = [&](const PpSeg& S, int pi, int tile, bool tile_ok) {
      const long long c_p0 = dbg_t ? clock64() : 0;
      ptx::fence_proxy_async();
      if (kCg2) {
        // one elected arrival per CTA (on the leader's barrier) after the group has synchronised
        ptx::tc_fence_before();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (group_leader) {
          const bool save = kTrain && S.save_row >= 0 && tile_ok;
          if (save) {
            ptx::tma_store_2d(&p.map_save, sm.panels + pi * kPanelBytes, col, S.save_row + tile * kTileM);
            ptx::tma_commit_group();
          }
          if (kCounters && rank == 0) *(volatile long long*)&ts_pub[pi >> 2][pi & 3] = clock64();
          trace(5, 0, pi >> 2, pi & 3);                              // group arrives on panel_ready
          if (!S.no_signal) {
            ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(pready_u32 + (kCg2 ? (pi & ~3) : pi) * 8, 0));
          }
          // the store's shared-memory read must be over before the panel is rewritten; this group's next
          // write to it is behind the bar.sync of the other tile's publish (or of last_epi), which this
          // thread only reaches after the wait
          if (save) ptx::tma_wait_group_read<0>();
        }
        if (dbg_t) c_pub += clock64() - c_p0;
        return;
      }
      if (kTrain && S.save_row >= 0) {
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (group_leader && tile_ok) {
          ptx::tma_store_2d(&p.map_save, sm.panels + pi * kPanelBytes, col, S.save_row + tile * kTileM);
          ptx::tma_commit_group();
          ptx::tma_wait_group_read<0>();
        }
      }
      ptx::tc_fence_before();
      if (!S.no_signal) ptx::mbar_arrive(&sm.panel_ready[pi]);
      if (dbg_t) c_pub += clock64() - c_p0;
    };

    // ReLU gate bits of epilogue (unit, e): loaded one epilogue ahead (CTA-pair backward programs)
    auto load_gate = [&](int unit, int e) -> uint2 {
      const PpSeg& G = p.segs[p.epi_seg[e >> 1]];
      if (G.mask_row < 0 || ((G.epi == EPI_BWD_START) && q >= 2)) return make_uint2(0u, 0u);
      const int s2 = ((unit * 2 + (e & 1)) * kCtas + rank) * kTileM + row;
      if (s2 >= p.n_samples) return make_uint2(0u, 0u);      // padding rows carry no gradient
      return __ldg(p.gate + (size_t)G.mask_row * 4 + (size_t)q * p.cap + s2);
    };
    const int n_e = kMmaOnly ? 0 : 2 * p.n_epi;
    uint2 gate_next = make_uint2(0u, 0u);
    if (kCg2 && unit0 < p.n_units) gate_next = load_gate(unit0, 0);

    // 16 packed words (32 columns) -> chunks chunk0 .. chunk0 + 3 of the swizzled panel row
    auto store_pk16 = [&](uint8_t* panel, int chunk0, const uint32_t* pk) {
      uint4* prow = reinterpret_cast<uint4*>(panel + row * 128);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        prow[swz_chunk(row, chunk0 + c)] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
    };

    for (int unit = unit0; unit < p.n_units; unit += unit_stride) {
      for (int e = 0; e < n_e; ++e) {
        {
          const PpSeg& S = p.segs[p.epi_seg[e >> 1]];
          const int t = e & 1;
          const uint2 gate = gate_next;
          if (kCg2) {
            if (e + 1 < n_e) gate_next = load_gate(unit, e + 1);
            else if (unit + unit_stride < p.n_units) gate_next = load_gate(unit + unit_stride, 0);
          }
          const int tile = (unit * 2 + t) * kCtas + rank;
          const bool tile_ok = tile < p.n_tiles;
          const int s = tile * kTileM + row;
          const bool valid = s < p.n_samples;
          const int pi = t * 4 + q;
          uint8_t* panel = sm.panels + pi * kPanelBytes;
          const uint32_t acc_addr = lane_addr + (uint32_t)(t * 256 + col);
          const bool participates = !((S.epi == EPI_VIEW || S.epi == EPI_BWD_START) && q >= 2);

          // side inputs that do not depend on the accumulator are fetched before waiting for the MMAs
          uint4 mk[8];
          float dd = 0.f;
          const bool need_mask = S.epi == EPI_BWD_RELU || S.epi == EPI_BWD_RELU_D || S.epi == EPI_BWD_START ||
                                 S.epi == EPI_BWD_START_PROP;
          if (!kCg2 && need_mask && participates) {
            if (valid) {
              const uint4* src = reinterpret_cast<const uint4*>(p.act + ((size_t)S.mask_row + s) * kW + col);
#pragma unroll
              for (int c = 0; c < 8; ++c) mk[c] = __ldg(src + c);
            } else {
#pragma unroll
              for (int c = 0; c < 8; ++c) mk[c] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
          if (S.epi == EPI_BWD_RELU_D) dd = valid ? __bfloat162float(__float2bfloat16(p.d_raw[(size_t)s * p.raw_c])) : 0.f;

          if (S.kps > 0) {
            // every group waits for every phase, also when it has no columns in this layer: a parity wait
            // that skipped a phase could be satisfied by an older phase of the same parity
            if (dbg_t) c_t0 = clock64();
            ptx::mbar_wait_u32(accfull_u32 + t * 8, (acc_phase >> t) & 1u);
            acc_phase ^= 1u << t;
            ptx::tc_fence_after();
            if (dbg_t) { c_t1 = clock64(); c_acc += c_t1 - c_t0; if (threadIdx.x == 0) *(volatile long long*)&ts_acc[t] = c_t1; }
            if (group_leader) trace(4, p.epi_seg[e >> 1], t, q);     // acc_full seen
          }

          switch (S.epi) {
            case EPI_RELU: case EPI_LINEAR: {
              float head = 0.f;
              if (kCg2) {
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
                  const float4* wd4 = reinterpret_cast<const float4*>(sm.bias + (p.w_dens_off - p.bias_tail0) + col);
#pragma unroll
                  for (int c = 0; c < 16; ++c) {
                    const float4 w = wd4[c];
                    head = fmaf(__uint_as_float(pk[2 * c] << 16), w.x, head);
                    head = fmaf(__uint_as_float(pk[2 * c] & 0xFFFF0000u), w.y, head);
                    head = fmaf(__uint_as_float(pk[2 * c + 1] << 16), w.z, head);
                    head = fmaf(__uint_as_float(pk[2 * c + 1] & 0xFFFF0000u), w.w, head);
                  }
                }
              } else {
#pragma unroll 1
              for (int hf = 0; hf < 2; ++hf) {
                load_acc32(acc_addr + (uint32_t)(hf * 32), v);
                const float4* b4 = reinterpret_cast<const float4*>(sm.bias + S.bias_off + col + hf * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 b = b4[c];
                  v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
                }
                if (S.head) {   // density head: dot of the bf16-rounded activation with the bf16-rounded kernel
                  const float* wd = sm.bias + p.w_dens_off + col + hf * 32;
#pragma unroll
                  for (int c = 0; c < 32; ++c)
                    head = fmaf(__bfloat162float(__float2bfloat16(fmaxf(v[c], 0.f))), wd[c], head);
                }
                if (S.epi == EPI_RELU) store_half32<true>(panel, row, hf * 4, v);
                else store_half32<false>(panel, row, hf * 4, v);
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
                if (kCg2) {
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
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16x2_relu(v[2 * j], v[2 * j + 1]);
                    const float4* wr4 = reinterpret_cast<const float4*>(sm.bias + (p.w_rgb_off - tail0) + (col + hf * 32) * 3);
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
                  if (kTrain && S.save_row >= 0 && tile_ok)
                    p.gate[(size_t)S.save_row * 4 + (size_t)q * p.cap + s] = make_uint2(g01[0], g01[1]);
                } else {
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                  load_acc32(acc_addr + (uint32_t)(hf * 32), v);
                  if (valid) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.viewbias + (size_t)(s / p.S) * 128 + col + hf * 32);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                      const float4 b = __ldg(b4 + c);
                      v[c * 4 + 0] += b.x; v[c * 4 + 1] += b.y; v[c * 4 + 2] += b.z; v[c * 4 + 3] += b.w;
                    }
                  }
                  const float* wr = sm.bias + (p.w_rgb_off - tail0) + (col + hf * 32) * 3;
#pragma unroll
                  for (int c = 0; c < 32; ++c) {
                    const float a = __bfloat162float(__float2bfloat16(fmaxf(v[c], 0.f)));
                    h0 = fmaf(a, wr[c * 3], h0); h1 = fmaf(a, wr[c * 3 + 1], h1); h2 = fmaf(a, wr[c * 3 + 2], h2);
                  }
                  if (kTrain) store_half32<true>(panel, row, hf * 4, v);
                }
                }
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
              if (kCg2) {
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
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                load_acc32(acc_addr + (uint32_t)(hf * 32), v);
                if (S.epi == EPI_BWD_RELU_D) {
                  const float4* w4 = reinterpret_cast<const float4*>(sm.bias + (p.w_dens_off - tail0) + col + hf * 32);
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 w = w4[c];
                    v[c * 4 + 0] = fmaf(dd, w.x, v[c * 4 + 0]); v[c * 4 + 1] = fmaf(dd, w.y, v[c * 4 + 1]);
                    v[c * 4 + 2] = fmaf(dd, w.z, v[c * 4 + 2]); v[c * 4 + 3] = fmaf(dd, w.w, v[c * 4 + 3]);
                  }
                }
                if (S.epi != EPI_BWD_LINEAR) {
                  const uint4 (&half)[4] = *reinterpret_cast<const uint4 (*)[4]>(&mk[hf * 4]);
                  apply_mask32(half, v);
                }
                store_half32<false>(panel, row, hf * 4, v);
              }
              publish(S, pi, tile, tile_ok);
              break;
            }
            case EPI_BWD_START: {
              // dV = W_rgb^T d_rgb (CUDA cores), gated by the saved view activation; 128 columns: groups 0, 1
              if (q < 2) {
                const float4 dr = valid ? reinterpret_cast<const float4*>(p.d_raw)[s] : make_float4(0, 0, 0, 0);
                const float d0 = __bfloat162float(__float2bfloat16(dr.y)), d1 = __bfloat162float(__float2bfloat16(dr.z)),
                            d2 = __bfloat162float(__float2bfloat16(dr.w));
                if (q == 0 && tile_ok) {   // padding rows of a real tile get zeros
                  uint4* dst = reinterpret_cast<uint4*>(p.drgb_out + (size_t)s * kHeadCols);
                  dst[0] = make_uint4(ptx::pack_bf16x2(dr.y, dr.z), ptx::pack_bf16x2(dr.w, dr.x), 0u, 0u);
                }
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  const float* wr = sm.bias + (p.w_rgb_off - tail0) + (col + hf * 32) * 3;
#pragma unroll
                  for (int c = 0; c < 32; ++c) v[c] = d0 * wr[c * 3] + d1 * wr[c * 3 + 1] + d2 * wr[c * 3 + 2];
                  if (kCg2) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
                    apply_gate16(hf == 0 ? gate.x : gate.y, pk);
                    store_pk16(panel, hf * 4, pk);
                  } else {
                  const uint4 (&half)[4] = *reinterpret_cast<const uint4 (*)[4]>(&mk[hf * 4]);
                  apply_mask32(half, v);
                  store_half32<false>(panel, row, hf * 4, v);
                  }
                }
                publish(S, pi, tile, tile_ok);
              } else if (kCg2 && !S.no_signal && group_leader) {
                // the tile barrier counts every group of both CTAs: groups without columns in this op arrive at once
                ptx::mbar_arrive_cluster_u32(ptx::mapa_u32(pready_u32 + (pi & ~3) * 8, 0));
              }
              break;
            }
            case EPI_BWD_START_PROP: {
              const float dd0 = valid ? p.d_raw[s] : 0.f;
              const float ddq = __bfloat162float(__float2bfloat16(dd0));
              if (q == 0 && tile_ok) {
                uint4* dst = reinterpret_cast<uint4*>(p.drgb_out + (size_t)s * kHeadCols);
                dst[0] = make_uint4(0u, ptx::pack_bf16x2(0.f, dd0), 0u, 0u);
              }
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const float* wd = sm.bias + (p.w_dens_off - tail0) + col + hf * 32;
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = ddq * wd[c];
                if (kCg2) {
                  uint32_t pk[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16x2(v[2 * j], v[2 * j + 1]);
                  apply_gate16(hf == 0 ? gate.x : gate.y, pk);
                  store_pk16(panel, hf * 4, pk);
                } else {
                const uint4 (&half)[4] = *reinterpret_cast<const uint4 (*)[4]>(&mk[hf * 4]);
                apply_mask32(half, v);
                store_half32<false>(panel, row, hf * 4, v);
                }
              }
              publish(S, pi, tile, tile_ok);
              break;
            }
            default: break;
          }
          if (dbg_t && S.kps > 0) {
            const long long dt = clock64() - c_t1;
            c_work += dt;
            if (S.epi == EPI_VIEW) c_view += dt; else if (S.head) c_head += dt;
          }
          if (S.last_epi) {
            // every warp of the group is past its TMEM reads / panel writes before the panels are released
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (group_leader) ptx::mbar_arrive(&sm.epi_done[t]);
          }
        }
      }
    }
    if (group_leader) ptx::tma_wait_group<0>();
    if (dbg_t) {
      long long* d = p.dbg + unit0 * 16 + (threadIdx.x == 0 ? 8 : 12);
      d[0] = clock64() - c_epi_start; d[1] = c_acc; d[2] = c_work; d[3] = c_pub;
      if (threadIdx.x == 0) { long long* e = p.dbg + unit0 * 16; e[0] = c_ld; e[1] = c_st; e[2] = c_view; e[3] = c_head; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (kCg2) ptx::cluster_sync_all();   // neither CTA retires (or frees TMEM) while the pair's MMAs / arrivals are in flight
  if (warp == kMma) { if (kCg2) ptx::tmem_dealloc_cg2(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512); }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side: segment programs
// ------------------------------------------------------------------------------------------
int pp_build(hugs_handle* h, const MlpViews& mv, TcMlp* m) {
  const hugs_model_desc& d = h->d;
  const int D = mv.depth;
  auto base_seg = [] {
    PpSeg s{};
    s.kps = 4; s.n_halves = 2; s.epi = EPI_NONE; s.save_row = -1; s.mask_row = -1; s.bias_idx = -1;
    return s;
  };
  // ---- forward ----
  bool cat = false;
  for (int i = 0; i < D; ++i) {
    const auto& P = m->pack[i];
    const bool last = i == D - 1;
    auto finish = [&](PpSeg& s) {
      s.epi = EPI_RELU; s.bias_off = P.bias_off; s.save_row = i;
      if (last) { s.head = 1; if (!mv.has_rgb) s.no_signal = 1; }
    };
    if (i == 0) {
      for (int part = 0; part < kFeatPad / 256; ++part) {
        PpSeg s = base_seg();
        s.a_feat = 1; s.feat_col0 = part * 256; s.w_row = P.row0; s.w_col0 = part * 256; s.accumulate = part > 0;
        if (part == 0) s.bias_idx = i;
        if (part == kFeatPad / 256 - 1) finish(s);
        m->pp_fwd.push_back(s);
      }
    } else {
      PpSeg s = base_seg();
      s.w_row = P.row0; s.w_col0 = 0; s.bias_idx = i;
      if (!cat) finish(s);
      m->pp_fwd.push_back(s);
      if (cat) {
        for (int part = 0; part < kFeatPad / 256; ++part) {
          PpSeg f = base_seg();
          f.a_feat = 1; f.feat_col0 = part * 256; f.w_row = P.row0; f.w_col0 = kW + part * 256; f.accumulate = 1;
          if (part == kFeatPad / 256 - 1) finish(f);
          m->pp_fwd.push_back(f);
        }
      }
    }
    cat = (i % d.skip_layer == 0 && i > 0);
  }
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
  for (size_t i = 0; i < m->pp_fwd.size(); ++i) {
    size_t j = (i + 1) % m->pp_fwd.size();
    while (m->pp_fwd[j].kps == 0) j = (j + 1) % m->pp_fwd.size();
    m->pp_fwd[i].feat_next = m->pp_fwd[j].a_feat;
  }
  HUGS_REQUIRE((int)m->pp_fwd.size() <= kMaxSegs, "ping-pong schedule: too many segments (%zu)", m->pp_fwd.size());
  // a feature refill must never directly follow an epilogue of the same tile (the epilogue writes the panels)
  for (size_t i = 1; i < m->pp_fwd.size(); ++i)
    HUGS_REQUIRE(!(m->pp_fwd[i].a_feat && m->pp_fwd[i - 1].epi != EPI_NONE), "unsupported segment order at %zu", i);

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
  HUGS_REQUIRE((int)m->pp_bwd.size() <= kMaxSegs, "ping-pong schedule: too many backward segments");
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
  }
  const bool cg2 = tc->use_cg2;
  const int tiles_per_unit = cg2 ? 4 : 2;
  p.n_tiles = n_tiles; p.n_units = (n_tiles + tiles_per_unit - 1) / tiles_per_unit; p.n_samples = n_samples; p.S = S;
  p.feat_row0 = tc->feat_row0[level];
  p.bias = m.bias; p.bias_floats = m.bias_floats; p.viewbias = tc->viewbias;
  p.raw_out = h->raw[level]; p.raw_c = is_prop ? 1 : 4;
  p.d_raw = h->d_raw[level]; p.act = tc->act; p.drgb_out = tc->drgb;
  p.w_dens_off = m.w_dens_off; p.w_rgb_off = m.w_rgb_off;
  p.dbg = (!is_prop && direction != 2) ? h->dbg_counters : nullptr;
  p.dens_bias_off = m.pack[mv.depth].bias_off;
  p.gate = tc->gate; p.cap = cap;
  for (int i = 0; i < p.n_segs; ++i)
    if (p.segs[i].epi != EPI_NONE) p.epi_seg[p.n_epi++] = i;
  p.bias_img = m.bias_img;
  p.bias_tail0 = cg2 ? m.pack[mv.depth].bias_off : 0;
  HUGS_REQUIRE(!cg2 || m.bias_floats - p.bias_tail0 <= kBiasTailFloats, "head table too large for the CTA-pair kernel");
  p.rgb_bias_off = mv.has_rgb ? m.pack[mv.depth + 3].bias_off : 0;
  if (cg2) {
    for (int i = 0; i < p.n_segs; ++i)
      HUGS_REQUIRE(direction != 2 || p.segs[i].n_halves == 2, "CTA-pair backward program must be N = 256 throughout");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * std::min(p.n_units, tc->num_sms / 2));
    cfg.blockDim = dim3(kPpThreads);
    cfg.dynamicSmemBytes = kPpSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    if (direction == 0) HUGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_pp_kernel<false, true>, p));
    else HUGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_pp_kernel<true, true>, p));
    return HUGS_OK;
  }
  const int grid = std::min(p.n_units, tc->num_sms);
  if (direction == 0) mlp_pp_kernel<false, false><<<grid, kPpThreads, kPpSmemBytes, st>>>(p);
  else mlp_pp_kernel<true, false><<<grid, kPpThreads, kPpSmemBytes, st>>>(p);
  HUGS_LAUNCH_CHECK();
  return HUGS_OK;
}

}  // namespace hugs
