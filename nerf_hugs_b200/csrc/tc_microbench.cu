// Micro-benchmark of the tcgen05.mma issue/execute rate (development evidence for DESIGN.md):
// one CTA, one issuing thread, operands fixed in shared memory, no TMA and no epilogue.
#include "ptx.cuh"
#include "common.cuh"
#include "handle.h"

namespace hugs {
namespace {

// mode bit 0: alternate between two accumulators; bit 1: commit after every 4 MMAs (like the chain kernel);
// bit 2: a second warp hammers shared memory with stores while the MMAs run
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n_cols, int n_mmas, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 6 * 16384);
  uint64_t* bar2 = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 3 && lane == 0) {
    constexpr uint32_t kDescHi = ptx::desc_hi_sw128(1024);
    const uint32_t a_addr = ptx::smem_u32(base), b_addr = ptx::smem_u32(base + 16384);
    const uint32_t idesc = ptx::make_idesc_bf16(128, n_cols, 0, 0);
    const uint64_t da = ptx::desc_from(kDescHi, a_addr), db = ptx::desc_from(kDescHi, b_addr);
    const long long t0 = clock64();
    for (int i = 0; i < n_mmas; ++i) {
      const uint32_t d = tmem_base + (((mode & 1) && (i & 1)) ? 256u : 0u);
      ptx::mma_bf16_ss(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, i > 1 ? 1u : 0u);
      if ((mode & 2) && (i & 3) == 3) ptx::mma_commit_u32(ptx::smem_u32(bar2));
    }
    const long long t1 = clock64();
    ptx::mma_commit(bar);
    ptx::mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;   // issue time
    out[1] = t2 - t0;   // until all MMAs completed
  } else if ((mode & 4) && warp == 1) {
    // shared-memory store traffic into an unrelated region
    uint4* dst = reinterpret_cast<uint4*>(base + 2 * 16384);
    for (int it = 0; it < n_mmas * 4; ++it) dst[(it * 32 + lane) & 4095] = make_uint4(it, it, it, it);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

// CTA-pair variant: M = 256 (cta_group::2), N = n_cols, issued by the leader CTA; mode bit 2: four warps per CTA
// hammer shared memory with 16-byte stores (epilogue-like traffic) while the MMAs run
__global__ void __launch_bounds__(192, 1) mma_rate_cg2_kernel(int n_cols, int n_mmas, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 6 * 16384);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc_cg2(tmem_ptr, 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 5 && lane == 0 && rank == 0) {
    constexpr uint32_t kDescHi = ptx::desc_hi_sw128(1024);
    const uint32_t a_addr = ptx::smem_u32(base), b_addr = ptx::smem_u32(base + 16384);
    const uint32_t idesc = ptx::make_idesc_bf16(256, n_cols, 0, 0);
    const uint64_t da = ptx::desc_from(kDescHi, a_addr), db = ptx::desc_from(kDescHi, b_addr);
    const long long t0 = clock64();
    for (int i = 0; i < n_mmas; ++i) {
      const uint32_t d = tmem_base + (((mode & 1) && (i & 4)) ? 256u : 0u);
      ptx::mma_bf16_ss_cg2(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, i > 7 ? 1u : 0u);
    }
    const long long t1 = clock64();
    ptx::mma_commit_mc2_u32(ptx::smem_u32(bar));
    ptx::mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  } else if ((mode & 4) && warp < 4) {
    uint4* dst = reinterpret_cast<uint4*>(base + 2 * 16384);
    for (int it = 0; it < n_mmas * 2; ++it) dst[(it * 128 + threadIdx.x) & 4095] = make_uint4(it, it, it, it);
  }
  if (!(warp == 5 && lane == 0 && rank == 0) && threadIdx.x == 160 + 1) {
    // the peer CTA must not leave before the MMAs that read its shared memory have completed
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 0) ptx::tmem_dealloc_cg2(tmem_base, 512);
}

}  // namespace
}  // namespace hugs

using namespace hugs;

HUGS_API int hugs_debug_mma_rate_cg2(int32_t n_cols, int32_t n_mmas, int32_t mode, int64_t* out_host) {
  HUGS_REQUIRE(out_host && (n_cols == 128 || n_cols == 256) && n_mmas > 0, "bad arguments");
  long long* d = nullptr;
  HUGS_CUDA(cudaMalloc(&d, 16));
  const int smem = 1024 + 6 * 16384 + 64;
  HUGS_CUDA(cudaFuncSetAttribute(mma_rate_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  HUGS_CUDA(cudaLaunchKernelEx(&cfg, mma_rate_cg2_kernel, (int)n_cols, (int)n_mmas, (int)mode, d));
  HUGS_CUDA(cudaDeviceSynchronize());
  long long h[2];
  HUGS_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  out_host[0] = h[0]; out_host[1] = h[1];
  return HUGS_OK;
}

namespace hugs { namespace {
// tcgen05.ld throughput: `n_warps` warps (4 per lane quarter group) each read `cols` fp32 columns of their
// 32-lane slice `iters` times with the 32x32b.x32 shape (the chain-kernel epilogue's access pattern).
__global__ void __launch_bounds__(512, 1) ldtm_rate_kernel(int cols, int iters, long long* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) ptx::tmem_alloc(&tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < cols; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld32(base + (uint32_t)((c + (warp >> 2) * 32) & 511), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_slot, 512);
}
}}  // namespace

HUGS_API int hugs_debug_ldtm_rate(int32_t n_warps, int32_t cols, int32_t iters, int64_t* out_host) {
  HUGS_REQUIRE(out_host && n_warps >= 4 && n_warps <= 16 && (n_warps % 4) == 0 && cols % 32 == 0, "bad arguments");
  long long* d = nullptr;
  HUGS_CUDA(cudaMalloc(&d, 16));
  ldtm_rate_kernel<<<1, n_warps * 32>>>(cols, iters, d);
  HUGS_LAUNCH_CHECK();
  HUGS_CUDA(cudaDeviceSynchronize());
  long long h[2];
  HUGS_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  out_host[0] = h[0]; out_host[1] = h[1];
  return HUGS_OK;
}

// Debug entry point (not part of the drop-in boundary): cycles to issue / complete n_mmas MMAs of
// shape M=128, N=n_cols, K=16 (bf16) from one thread.  out: int64[2] on the host.
HUGS_API int hugs_debug_mma_rate(int32_t n_cols, int32_t n_mmas, int32_t mode, int64_t* out_host) {
  HUGS_REQUIRE(out_host && (n_cols == 64 || n_cols == 128 || n_cols == 256) && n_mmas > 0, "bad arguments");
  long long* d = nullptr;
  HUGS_CUDA(cudaMalloc(&d, 16));
  const int smem = 1024 + 6 * 16384 + 64;
  HUGS_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mma_rate_kernel<<<1, 128, smem>>>(n_cols, n_mmas, mode, d);
  HUGS_LAUNCH_CHECK();
  HUGS_CUDA(cudaDeviceSynchronize());
  long long h[2];
  HUGS_CUDA(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  cudaFree(d);
  out_host[0] = h[0]; out_host[1] = h[1];
  return HUGS_OK;
}

// Debug: enable per-role cycle counters of the NerfMLP forward chain kernel; returns [148][16] int64 on read.
HUGS_API int hugs_debug_counters(hugs_handle* h, int32_t enable, int64_t* out_host) {
  HUGS_REQUIRE(h, "null handle");
  if (enable && !h->dbg_counters) {
    void* q = nullptr;
    HUGS_CUDA(cudaMalloc(&q, 148 * 16 * 8 + 74 * 60 * 8 + 8192 * 8));      // + [74 clusters][20 segments][3] per-segment waits
    HUGS_CUDA(cudaMemset(q, 0, 148 * 16 * 8 + 74 * 60 * 8 + 8192 * 8));
    h->allocs.push_back(q);
    h->dbg_counters = static_cast<long long*>(q);
  }
  if (out_host && h->dbg_counters) {
    HUGS_CUDA(cudaDeviceSynchronize());
    HUGS_CUDA(cudaMemcpy(out_host, h->dbg_counters, 148 * 16 * 8 + 74 * 60 * 8 + 8192 * 8, cudaMemcpyDeviceToHost));
  }
  if (!enable) h->dbg_counters = nullptr;
  return HUGS_OK;
}
