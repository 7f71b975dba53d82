// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// small vector load/store helpers, and the host-side error plumbing of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda.h>  // CUtensorMap type + enums only; the driver entry point is resolved at run time
#include <stdint.h>
#include <stdio.h>

namespace gd {

// ----------------------------------------------------------------------------------------------
// host-side error state (thread local), see gd_last_error() in api.cu
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define GD_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      gd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                    __LINE__);                                                           \
      return -2;                                                                         \
    }                                                                                    \
  } while (0)
#define GD_REQUIRE(cond, ...)     \
  do {                            \
    if (!(cond)) {                \
      gd::set_error(__VA_ARGS__); \
      return -1;                  \
    }                             \
  } while (0)

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched through launch_pdl() may be SCHEDULED while its predecessor in the
// stream is still running (its CTAs become resident as SM resources free up) and must call pdl_wait() before its
// first access to global memory; pdl_wait() returns once every prerequisite grid has completed and its writes are
// visible.  pdl_trigger() lets the NEXT kernel in the stream be scheduled.  Every kernel calls wait first and then
// trigger, so at most one successor is parked behind a running kernel.  The idea: hide the launch latency and the
// CTA-scheduling ramp between the ~725 dependent launches of a sampling step (also inside a captured CUDA graph,
// where the attribute becomes a programmatic edge).  MEASURED (profiles/pdl_ab_r01.log, same box, alternating runs):
// the step gets SLOWER with the attribute on -- 166.6 -> 169.8 ms at batch 64, 22.0 -> 22.4 ms at batch 8 -- so it is
// OFF by default; GD_B200_PDL=1 (or gd_debug_set(6, 1)) turns it on.  Without the attribute griddepcontrol.* are no-ops.
bool pdl_enabled();
bool pdl_small_enabled();
// Function attributes (opt-in dynamic shared memory) are per DEVICE: true the first time this call site runs on the
// current device (one process may drive several GPUs), false afterwards.  `seen` is a zero-initialised static bitmap.
inline bool first_use_on_device(unsigned long long (&seen)[2]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (seen[dev >> 6] & bit) return false;
  seen[dev >> 6] |= bit;
  return true;
}

#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_wait();
  pdl_trigger();
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool enabled, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = enabled ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  return launch_pdl_if(pdl_enabled(), kernel, grid, block, smem, st, args...);
}
// tiny kernels between two convolutions (see pdl_small_enabled)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_small(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args... args) {
  return launch_pdl_if(pdl_small_enabled(), kernel, grid, block, smem, st, args...);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// x * sigmoid(x) with the approximate-division intrinsic (2 ulp): the IEEE division's slow path costs a branch per
// element in the bandwidth-bound GroupNorm kernels
// One MUFU op per element: sigmoid(v) = 0.5 * tanh(v / 2) + 0.5 with tanh.approx.f32 (max rel. error 2^-11, the
// same size as the fp16 rounding of the result).  The exp + reciprocal form costs two MUFU ops and made the fused
// GroupNorm+SiLU kernels MUFU-bound instead of HBM-bound (profiles/gn_sweep_r01: bwd 64% -> with/without SiLU).
__device__ __forceinline__ float tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float sigmoid_f(float v) { return fmaf(0.5f, tanh_approx(0.5f * v), 0.5f); }
__device__ __forceinline__ float silu_f(float v) {
  const float h = 0.5f * v;
  return fmaf(h, tanh_approx(h), h);
}

// GroupNorm32 (+FiLM) folded into one per-channel affine y = x * a + b:
//   a = rstd * gamma [* (1 + scale)],  b = (beta - mean * rstd * gamma) [* (1 + scale) + shift]
// (nn.py:17-19 + unet.py:248-252; scale = first half of the FiLM row).  ONE definition with pinned roundings, shared by
// gd_groupnorm_apply, by the coefficient table of gd_groupnorm_coef / gd_groupnorm_finalize_partials and through it by
// the convolution that normalises its operand on the fly, so all paths produce the same bits.
__device__ __forceinline__ void gn_affine(float mean, float rstd, float ga, float be, bool film, float scale, float shift,
                                          float& a, float& b) {
  float aa = __fmul_rn(rstd, ga);
  float bb = __fmaf_rn(-mean, aa, be);
  if (film) {
    const float sc = __fadd_rn(1.0f, scale);
    aa = __fmul_rn(aa, sc);
    bb = __fmaf_rn(bb, sc, shift);
  }
  a = aa;
  b = bb;
}
__device__ __forceinline__ void gn_load_affine(const float* mean_rstd, const float* gamma, const float* beta,
                                               const float* film, int film_ld, int n, int c, int ch0, float (&a)[8],
                                               float (&b)[8]) {
  const int cpg = c / 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j;
    const int g = ch / cpg;
    const float mean = mean_rstd[(static_cast<size_t>(n) * 32 + g) * 2];
    const float rstd = mean_rstd[(static_cast<size_t>(n) * 32 + g) * 2 + 1];
    const bool has_film = film != nullptr;
    gn_affine(mean, rstd, gamma[ch], beta[ch], has_film, has_film ? film[static_cast<size_t>(n) * film_ld + ch] : 0.f,
              has_film ? film[static_cast<size_t>(n) * film_ld + c + ch] : 0.f, a[j], b[j]);
  }
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// The same with a suspend-time hint (ns): the thread may sleep in hardware until the phase completes OR the hint
// expires, instead of returning after the (short) default time limit.  Completion still wakes it at once; what the hint
// removes is the polling itself — in the fused conv a quarter of all issued warp instructions were try_wait / clock /
// compare / branch of spinning producer, MMA and epilogue warps (ncu source page), issue slots and power the
// transform warps of the same SM partitions need.
constexpr uint32_t kMbarSuspendNs = 20000;
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
// acquire at cluster scope: the data the barrier guards was written by the OTHER CTA of a pair (generic-proxy stores
// followed by a release.cluster arrive)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.  The clock is only read
// every 1024 unsuccessful polls (with the hint a poll is rare; without it the loop must stay cheap).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = 0;
  for (uint32_t spins = 1; !mbar_try_wait_hint(bar, parity); ++spins) {
    if ((spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {  // ~2 s at 2 GHz
        printf("gd: mbarrier wait timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  long long t0 = 0;
  for (uint32_t spins = 1; !mbar_try_wait_cluster(bar, parity); ++spins) {
    if ((spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {
        printf("gd: mbarrier (cluster) wait timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_arrive_count(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// One lane of a converged warp (elect.sync); used so that single-thread issue sits inside warp-uniform control flow.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- TMA (cp.async.bulk.tensor, tiled mode) --------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, fp16 operands, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA pairs (cta_group::2): two SMs of a cluster cooperate on one 256-row tile -------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of `p` translated into the shared::cluster window of CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the only thing ordered through this arrive is TMEM traffic, which the tcgen05 fences on both sides
  // order; a .release here costs a MEMBAR that waits for every outstanding global store of the thread (7-10% of
  // the epilogue's stall samples in profiles/ncu_conv_r01f)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// release at cluster scope: publishes this thread's (and, after __syncwarp, its warp's) shared-memory stores to the
// CTA that waits on the barrier
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// default semantics (release at CTA scope) on a barrier of another CTA of the cluster: what a producer warp uses after
// fence.proxy.async to hand shared-memory tiles to the pair's MMA issuer (the tiles are read by the async proxy of the
// CTA that wrote them; a cluster-scope release / acquire pair costs a full fence on both sides — measured: the MMA
// issuer lost ~800 cycles per slot to the acquire)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads issued by either CTA of a pair; completion bytes are signalled on the barrier at `bar_cluster_addr`
// (the leader CTA's barrier, in the shared::cluster window)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// the same with multicast: the box lands at the same CTA-relative address in every CTA of `cta_mask`, and each
// destination's bytes are signalled on the barrier at the same offset in the leader of ITS pair (cta_group::2)
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                   int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T over the CTA pair: M = 256 (128 rows per CTA), B rows split across the pair
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `cta_mask` once the pair's MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major operand tile laid out by TMA with 128-byte swizzle:
// rows of 128 B (64 fp16), 8-row groups 1024 B apart.  Field layout follows the sm_100 UMMA
// descriptor (start>>4 | LBO>>4 @16 | SBO>>4 @32 | version=1 @46 | layout=SWIZZLE_128B(2) @61).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;          // leading byte offset (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;  // stride byte offset between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;          // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;          // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: fp16 A/B (K-major both), fp32 D, M x N tile.
__device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- 16-byte vector helpers ------------------------------------------------------------------
struct alignas(16) Half8 {
  __half2 h[4];
};
__device__ __forceinline__ void half8_to_float(const Half8& v, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(v.h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ Half8 float_to_half8(const float (&f)[8]) {
  Half8 v;
#pragma unroll
  for (int i = 0; i < 4; ++i) v.h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
// 16-byte accesses go through uint4: copying the struct itself is member-wise (__half2 has user-provided copy
// operations), which the compiler lowers to FOUR 32-bit accesses — for the 64-byte-pitch staging tiles in shared memory
// that is a 4-way bank conflict per store (ncu: 134 M wavefronts instead of 34 M in the dominant conv), in global memory
// four partial-sector stores per lane.
__device__ __forceinline__ Half8 ld_half8(const __half* p) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  Half8 v;
  *reinterpret_cast<uint4*>(&v) = t;
  return v;
}
// streaming 16-byte load: read-only path, do not allocate in L1 (the GroupNorm kernels touch every byte exactly once)
__device__ __forceinline__ Half8 ld_half8_stream(const __half* p) {
  uint32_t a, b, c, d;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "l"(p));
  Half8 v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v);
  w[0] = a; w[1] = b; w[2] = c; w[3] = d;
  return v;
}
__device__ __forceinline__ void st_half8(__half* p, const Half8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}
// the same for any 16-byte aligned address (shared-memory staging tiles)
__device__ __forceinline__ void st_half8_at(void* p, const Half8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}
__device__ __forceinline__ Half8 ld_half8_at(const void* p) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  Half8 v;
  *reinterpret_cast<uint4*>(&v) = t;
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

}  // namespace gd
