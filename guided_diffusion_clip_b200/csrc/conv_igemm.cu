// Implicit-GEMM convolution (3x3 pad 1 / 1x1) on tcgen05 tensor cores for sm_100a.
//
// Replaces every nn.Conv2d / nn.Conv1d(k=1) call site of the reference torso
// (guided_diffusion/unet.py:185,211,222,286,294,483,616 via nn.py:22-32) and, with flipped/transposed
// packed weights, their data-gradient (conv backward-data) used by the guidance gradient.
//
// Formulation:  D[m, n] = sum_k A[m, k] * B[n, k]
//   m  : output pixel.  One CTA tile = 128 pixels = a BI x BH x BW patch (images x rows x cols).
//   k  : (tap, channel) of source 0, followed by the channels of an optional 1x1 "skip" source 1
//        (this fuses ResBlock.skip_connection + out_layers conv into one accumulation, unet.py:256).
//   n  : output channel.
// Activations are NHWC fp16 "channel views" (pointer, C, ld) so concatenations are free: producers
// write straight into channel slices of the concat buffer.  For a K block (tap, 64 channels) the A tile
// is ONE 4-D TMA box {64ch, BW, BH, BI} at (c, x0+dx, y0+dy, n0): the conv halo and the image border are
// the TMA's out-of-bounds zero fill, and the box lands in shared memory exactly as the 128-row x 128-byte
// K-major SWIZZLE_128B tile tcgen05.mma wants.  Weights are a plain 2-D TMA box {64, BN}.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over tiles):
//   warp 0   : TMA producer (one lane)           smem ring: full[s]/empty[s] mbarriers
//   warp 1   : TMEM alloc + tcgen05.mma issuer   accumulators: 2 x 256 TMEM columns (double buffered)
//   warps 2-5: epilogue.  Fast path (fp16 NHWC output, Cout % 64 == 0): per 64-column sub-tile
//              tcgen05.ld -> + bias (smem) + residual (TMA-loaded tile) -> fp16 -> 128B-swizzled smem tile ->
//              ONE 4-D TMA store {64ch, BW, BH, BI}: full-line writes, clipping of overhanging patches for free.
//              Slow path (fp32 NCHW output / Cout not a multiple of 64): direct per-thread stores.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KiB
constexpr int kMaxA = 8;   // activation-ring slots (barrier array size)
constexpr int kMaxB = 32;  // weight-ring slots
constexpr int kThreads = 320;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
constexpr int kEpiThreads = 256;
// Fused-GroupNorm variant (kGn): warps 2..7 normalise the main operand on its way into shared memory, the epilogue
// moves to warps 8..15 (four whole warpgroups, so setmaxnreg can shift registers from the light roles to the epilogue)
constexpr int kThreadsGn = 512;
constexpr int kGnWarps = 6;
constexpr int kGnHaloW = 18, kGnHaloH = 10;                    // (BW + 2) x (BH + 2) pixels of a 16 x 8 tile
constexpr int kGnRows = kGnHaloW * kGnHaloH * 8 / 32;          // 45 warp-rows of 32 16-byte chunks per channel block
constexpr int kGnIters = (kGnRows + kGnWarps - 1) / kGnWarps;  // 8 chunks per lane
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBarrierBytes = 1024;
constexpr int kBiasBytes = 1024;            // 256 floats
constexpr int kEpiCols = 32;                          // columns per epilogue sub-tile
constexpr int kEpiTileBytes = kBM * kEpiCols * 2;     // 8 KiB: 128 rows x 32 fp16, SWIZZLE_64B
constexpr int kEpiBytes = 4 * kEpiTileBytes;          // per epilogue group: output staging + residual staging

struct ConvArgs {
  // geometry
  int n_img, h, w;
  int bi, bh, bw;              // patch
  int patches_x, patches_y;    // per image (group)
  int m_tiles, n_tiles;
  int bn;                      // N tile
  // mainloop rings: activation slots and weight slots are pipelined independently
  int halo;          // 1: one activation slot = (BH+2) x BW pixel halo tile shared by the 3 ky taps of one kx
  int a_slot_bytes;  // halo: (BH+2)*BW*128, else 16 KiB
  int n_a, n_b;      // ring depths
  // K blocks
  int kb0_per_tap;             // C0 / 64
  int taps;                    // 9 or 1
  int kb0;                     // taps * C0/64
  int kb_total;                // kb0 + C1/64
  // epilogue
  int cout;
  const float* bias;
  const __half* res;
  int ld_res;
  int res_mode;
  void* out;
  int ld_out;
  int out_mode;
  float out_scale;
  int tma_epi;  // 1 = staged TMA-store epilogue
  float* stats;  // optional GroupNorm partials [m_tiles*4][stats_ld][2] (sum, sum of squares per 4-channel chunk)
  int stats_ld;  // n_pad / 4
  int stats_per_tile;  // 1: one partial row per 128-pixel tile (tile inside one image); 4: one per 32-pixel quarter
  int debug;    // 0 = normal; 1 = epilogue skipped (barriers only); 2 = TMEM loads only (timing experiments)
  int quad;     // pair kernels: 1 = clusters of FOUR CTAs (two pairs on consecutive pixel-tile pairs of the same N tile):
                // every CTA loads a QUARTER of the weight tile and multicasts it to its counterpart in the other pair
  // fused GroupNorm (+FiLM, +SiLU, + nearest x2) on the main operand (kGn kernels only)
  int gn_mode, gn_silu;
  const __half* gn_src;  // raw a0
  int gn_ld;
  const __half* gn_src1;  // a1 (fused 1x1-skip operand), read by the same warps
  int gn_ld1;
  const float* gn_coef;  // [n][c0/8][16]: a[8], b[8] per 8-channel chunk
  // split-K (few pixel tiles, long K: the 8x8 / 16x16 layers at small batch): work item = (pixel unit, n tile, split);
  // split s accumulates main channel blocks [s*cb_split, (s+1)*cb_split) (with their share of the skip operand) and
  // stores its raw fp32 accumulators to ws[s][m_tiles*128][ws_ld]; conv_splitk_reduce_kernel finishes the epilogue
  int ksplit, cb_split;
  float* ws;
  int ws_ld;
};

__device__ __forceinline__ void tile_coords(const ConvArgs& p, int m_tile, int& n0, int& y0, int& x0) {
  const int px = m_tile % p.patches_x;
  const int t = m_tile / p.patches_x;
  const int py = t % p.patches_y;
  const int ig = t / p.patches_y;
  n0 = ig * p.bi;
  y0 = py * p.bh;
  x0 = px * p.bw;
}

// Output tiles are written once and not re-read by this kernel: store them with an evict-first L2 policy so they do
// not displace the activation lines the other 8 taps are about to re-read.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
      ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barrier of one epilogue group (4 warps); ids 1 and 2
__device__ __forceinline__ void epi_barrier(int group) { asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory"); }

// predicated 16-byte shared-memory store at a compile-time byte offset from `addr` (the offset folds into the instruction)
template <int kOff>
__device__ __forceinline__ void st_shared_v4_pred(uint32_t addr, const uint32_t (&v)[4], uint32_t pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "@p st.shared.v4.b32 [%0 + %6], {%1, %2, %3, %4};\n\t}"
      ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(pred), "n"(kOff)
      : "memory");
}

// 16-byte chunk `chunk` (0..3) of row `row` inside a [rows][64 B] tile laid out with the 64-byte swizzle
__device__ __forceinline__ uint8_t* sw64(uint8_t* tile, int row, int chunk) {
  return tile + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}

// Fast-epilogue arithmetic of one thread's 32 columns: accumulator + bias (+ residual) -> fp16 -> swizzled staging
// row; optional GroupNorm partial sums per 4-channel chunk (taken from the fp32 values: the difference to the
// fp16-rounded ones is O(2^-12 / sqrt(count)) of the mean, far below the fp16 activation noise).  Straight-line
// per variant: runtime flags inside the unrolled loop cost instruction-cache misses and branch stalls.
template <bool kRes, bool kStats>
__device__ __forceinline__ void epi_compute32(const uint32_t (&v)[32], const float4 (&bv)[8], const Half8 (&rv)[4],
                                              uint8_t* so, int row, float (&cs)[16]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float f[8];
    f[0] = __uint_as_float(v[8 * q + 0]) + bv[2 * q].x;
    f[1] = __uint_as_float(v[8 * q + 1]) + bv[2 * q].y;
    f[2] = __uint_as_float(v[8 * q + 2]) + bv[2 * q].z;
    f[3] = __uint_as_float(v[8 * q + 3]) + bv[2 * q].w;
    f[4] = __uint_as_float(v[8 * q + 4]) + bv[2 * q + 1].x;
    f[5] = __uint_as_float(v[8 * q + 5]) + bv[2 * q + 1].y;
    f[6] = __uint_as_float(v[8 * q + 6]) + bv[2 * q + 1].z;
    f[7] = __uint_as_float(v[8 * q + 7]) + bv[2 * q + 1].w;
    if (kRes) {
      float t[8];
      half8_to_float(rv[q], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += t[j];
    }
    st_half8_at(sw64(so, row, q), float_to_half8(f));  // one STS.128: conflict-free with the 64-byte swizzle
    if (kStats) {
      cs[2 * q] = (f[0] + f[1]) + (f[2] + f[3]);
      cs[2 * q + 1] = (f[4] + f[5]) + (f[6] + f[7]);
      cs[8 + 2 * q] = fmaf(f[3], f[3], fmaf(f[2], f[2], fmaf(f[1], f[1], f[0] * f[0])));
      cs[8 + 2 * q + 1] = fmaf(f[7], f[7], fmaf(f[6], f[6], fmaf(f[5], f[5], f[4] * f[4])));
    }
  }
}

struct EpiCtx {
  int img, y, x;
  bool valid;
  size_t pix;
};

// ---- slow-path epilogue: direct per-thread stores (fp32 NCHW outputs and Cout not a multiple of 64) ----------
template <int kCols>
__device__ __forceinline__ void epi_chunk_direct(const ConvArgs& p, uint32_t taddr, int col, const EpiCtx& e) {
  uint32_t v[kCols];
  if constexpr (kCols == 32) {
    tmem_ld_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
  } else {
    tmem_ld_x16(taddr, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
  }
  const bool full = col + kCols <= p.cout;
  const bool use_res = p.res_mode != GD_RES_NONE && e.valid && full && p.debug == 0;
  float racc[kCols];
  if (use_res) {
#pragma unroll
    for (int j = 0; j < kCols; ++j) racc[j] = 0.f;
    const int nsrc = p.res_mode == GD_RES_AVGPOOL2 ? 4 : 1;
    for (int s4 = 0; s4 < nsrc; ++s4) {
      size_t rpix;
      if (p.res_mode == GD_RES_SAME) rpix = e.pix;
      else if (p.res_mode == GD_RES_UPSAMPLE2)
        rpix = (static_cast<size_t>(e.img) * (p.h >> 1) + (e.y >> 1)) * (p.w >> 1) + (e.x >> 1);
      else
        rpix = (static_cast<size_t>(e.img) * (p.h * 2) + (2 * e.y + (s4 >> 1))) * (p.w * 2) + (2 * e.x + (s4 & 1));
      const __half* rp = p.res + rpix * p.ld_res + col;
#pragma unroll
      for (int q = 0; q < kCols / 8; ++q) {
        float t[8];
        half8_to_float(ld_half8(rp + 8 * q), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) racc[8 * q + j] += t[j];
      }
    }
  }
  tmem_ld_wait();
  if (!e.valid || p.debug != 0) return;
  const float rs = p.res_mode == GD_RES_AVGPOOL2 ? 0.25f : 1.0f;
  float f[kCols];
#pragma unroll
  for (int j = 0; j < kCols; ++j) {
    f[j] = __uint_as_float(v[j]);
    if (p.bias != nullptr && col + j < p.cout) f[j] += __ldg(p.bias + col + j);
    if (use_res) f[j] += rs * racc[j];
    f[j] *= p.out_scale;
  }
  if (p.out_mode == GD_OUT_NHWC_F16) {
    __half* op = reinterpret_cast<__half*>(p.out) + e.pix * p.ld_out + col;
    if (full) {
#pragma unroll
      for (int q = 0; q < kCols / 8; ++q)
        st_half8(op + 8 * q, float_to_half8(*reinterpret_cast<float(*)[8]>(&f[8 * q])));
    } else {
#pragma unroll
      for (int j = 0; j < kCols; ++j)
        if (col + j < p.cout) op[j] = __float2half_rn(f[j]);
    }
  } else {  // GD_OUT_NCHW_F32: adjacent threads are adjacent x -> coalesced per channel plane
    float* op = reinterpret_cast<float*>(p.out);
    const size_t plane = static_cast<size_t>(p.h) * p.w;
#pragma unroll
    for (int j = 0; j < kCols; ++j)
      if (col + j < p.cout)
        op[(static_cast<size_t>(e.img) * p.cout + col + j) * plane + static_cast<size_t>(e.y) * p.w + e.x] = f[j];
  }
}

// kTwo = CTA-pair mode (cluster of 2, tcgen05 cta_group::2): the pair computes a 256-pixel x BN tile — rank r owns
// pixel tile 2*pair + r and loads its own A tile plus HALF of the weight tile; the leader's single thread issues
// M=256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM.  Halving the B traffic per SM is what
// lifts the mainloop off the shared-memory / L2->SM bandwidth limit.
template <bool kTwo, bool kGn>
__global__ void __launch_bounds__(kGn ? kThreadsGn : kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ CUtensorMap map_res, const ConvArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [barriers 1 KiB][spare 1 KiB][activation ring][weight ring][out staging 2 x 8 KiB][residual staging 2 x 8 KiB]
  // align by OFFSETTING the shared-space pointer (an integer round trip would demote every later access to
  // generic LD/ST instead of LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_a = full_a + kMaxA;
  uint64_t* full_b = empty_a + kMaxA;
  uint64_t* empty_b = full_b + kMaxB;
  uint64_t* tmem_full = empty_b + kMaxB;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;  // one per epilogue group
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_bar + 2);
  float* s_stat = reinterpret_cast<float*>(smem + kBarrierBytes);  // [2 groups][4 quarters][16] parked GN partial sums
  uint8_t* a_ring = smem + kBarrierBytes + kBiasBytes;
  const int b_rows = kTwo ? (p.bn >> 1) : p.bn;  // weight rows this CTA stages per K block
  const int b_tile_bytes = b_rows * kBK * 2;     // multiple of 1024 (host checks b_rows % 8 == 0)
  uint8_t* b_ring = a_ring + p.n_a * p.a_slot_bytes;
  uint8_t* s_out = b_ring + p.n_b * b_tile_bytes;
  uint8_t* s_res = s_out + 2 * kEpiTileBytes;

  // warp index via shuffle: the compiler then knows it is warp-uniform and keeps the role loops on the uniform path
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base_u32 = smem_u32(smem);
  const uint32_t crank = kTwo ? cluster_ctarank() : 0u;  // rank in the cluster (2 CTAs, or 4 in quad mode)
  const uint32_t rank = crank & 1u;                      // rank within the CTA pair
  const uint32_t pair_in_cluster = crank >> 1;           // 0, or 0 / 1 in quad mode
  const uint32_t lead_rank = crank & ~1u;                // cluster rank of this pair's leader (MMA issuer)
  const bool lead_cta = rank == 0;
  const bool quad = kTwo && p.quad != 0;
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2u * pair_in_cluster));
  // work items: (pixel tile | pixel-tile pair | two pixel-tile pairs, n tile); all CTAs of a cluster walk the same list
  const int m_per = kTwo ? (quad ? 4 : 2) : 1;
  const int m_units = (p.m_tiles + m_per - 1) / m_per;
  const int mn_tiles = m_units * p.n_tiles;
  const int total_tiles = mn_tiles * p.ksplit;  // split-K: the split index is the slowest digit of the work item
  const int work0 = static_cast<int>(blockIdx.x) / m_per;
  const int work_step = static_cast<int>(gridDim.x) / m_per;
  // pixel tile of this CTA inside work item `tile`
  auto m_tile_of = [&](int tile) {
    const int unit = (tile / p.n_tiles) % m_units;
    return kTwo ? m_per * unit + 2 * static_cast<int>(pair_in_cluster) + static_cast<int>(rank) : unit;
  };
  const int cb_split = p.cb_split;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    if (p.tma_epi) {
      tma_prefetch_desc(&map_out);
      tma_prefetch_desc(&map_res);
    }
    for (int s = 0; s < p.n_a; ++s) {
      // kGn: every transform warp (of both CTAs of a pair) arrives once per slot
      mbar_init(&full_a[s], kGn ? (kTwo ? 2 * kGnWarps : kGnWarps) : 1);
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < p.n_b; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], quad ? 2 : 1);  // quad: a weight slot is refilled for BOTH pairs -> both must have read it
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kTwo ? 2 * kEpiThreads : kEpiThreads);  // pair mode: both CTAs' epilogue threads release the leader
    }
    mbar_init(&res_bar[0], 1);
    mbar_init(&res_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (kTwo) tmem_alloc_2sm(tmem_base_slot, 512);
    else tmem_alloc(tmem_base_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (kTwo) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  constexpr int kEpiWarp0 = kGn ? 8 : 2;  // first epilogue warp
  // everything above touched only shared memory, TMEM and the kernel parameters: it overlaps the predecessor's tail
  pdl_enter();

  // K schedule shared by the producer and the MMA issuer.  Per 64-channel block of the main operand:
  //   halo mode (3x3, one image per tile): 3 activation groups (kx = 0..2), each ONE (BH+2) x BW halo tile that
  //     serves the 3 ky taps as row-shifted views (ky * BW * 128 bytes) -> 3 activation copies instead of 9;
  //   plain 3x3: 9 groups of one tap; 1x1: one group.  Then the fused 1x1-skip operand: one group per block.
  // Every tap of every group consumes one weight slot.
  const int n_a = p.n_a, n_b = p.n_b, a_slot = p.a_slot_bytes;
  const int kb0_per_tap = p.kb0_per_tap, ntap0 = p.taps, nblk1 = p.kb_total - p.kb0, c0_total = p.kb0_per_tap * kBK;
  const bool halo = p.halo != 0;
  const int ngrp0 = halo ? 3 : ntap0;  // groups per channel block of the main operand
  const int nt0 = halo ? 3 : 1;        // taps per group
  const uint32_t tap_stride = static_cast<uint32_t>(p.bw) * 128u;  // smem bytes between ky views of a halo tile

  if (warp < kEpiWarp0) {
  // kGn: 512 threads start with 128 registers each; the producer / MMA / transform warpgroups (warps 0..7) hand 40 of
  // them to the two epilogue warpgroups, which need ~168 for a 32-column sub-tile with residual and statistics
  if (kGn) asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp walks the loop (warp-uniform control flow keeps addresses and coordinates in uniform
    // registers); one elected lane issues the copies.
    const int n_tiles = p.n_tiles, bn = p.bn;
    const uint32_t full_a0_cluster = kTwo ? mapa_u32(&full_a[0], lead_rank) : 0u;  // the pair leader's barriers
    const uint32_t full_b0_cluster = kTwo ? mapa_u32(&full_b[0], lead_rank) : 0u;
    const uint32_t halo_bytes = static_cast<uint32_t>(a_slot);
    // quad mode: this CTA fetches rows [q_row, q_row + b_rows/2) of its pair-half of the weight tile and multicasts
    // them to the CTA of the same pair rank in the other pair (cluster ranks rank and rank + 2)
    const int q_rows = b_rows >> 1;
    const uint16_t mc_mask = static_cast<uint16_t>((1u << rank) | (1u << (rank + 2u)));
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = work0; tile < total_tiles; tile += work_step) {
      const int n_tile = tile % n_tiles;
      const int m_tile = m_tile_of(tile);
      int n0, y0, x0;
      tile_coords(p, m_tile, n0, y0, x0);
      const int b_row0 = n_tile * bn + (kTwo ? static_cast<int>(rank) * b_rows : 0) +
                         (quad ? static_cast<int>(pair_in_cluster) * q_rows : 0);
      // K schedule (shared with the MMA issuer and the transform warps): per 64-channel block of the main operand its
      // groups (3 kx halo groups, 9 taps, or the single tap of a 1x1), then THIS block's share of the fused 1x1-skip
      // operand's blocks — the skip operand is spread evenly over the main blocks instead of trailing the tile: a run of
      // one-tap skip blocks needs 16 KB of activations + 16 KB of weights per 512 tensor cycles (64 B/clk per SM, more
      // than L2 delivers to 148 SMs), interleaved the stream stays below 48 B/clk (K = 2816: 1 294 -> see DESIGN 4.1.1).
      const int cb_lo = (tile / mn_tiles) * cb_split;
      const int cb_hi = cb_lo + cb_split < kb0_per_tap ? cb_lo + cb_split : kb0_per_tap;
      for (int cb = cb_lo; cb < cb_hi; ++cb) {
        const int s_lo = cb * nblk1 / kb0_per_tap, s_hi = (cb + 1) * nblk1 / kb0_per_tap;
        const int n_items = ngrp0 + (s_hi - s_lo);
        for (int it = 0; it < n_items; ++it) {
          const bool main_src = it < ngrp0;
          const int g = it;                      // main operand: group index
          const int skb = s_lo + (it - ngrp0);   // skip operand: block index
          const CUtensorMap* ma = main_src ? &map_a0 : &map_a1;
          const int nt = main_src ? nt0 : 1;
          const uint32_t a_bytes = (main_src && halo) ? halo_bytes : static_cast<uint32_t>(kATileBytes);
          const int a_ch = (main_src ? cb : skb) * kBK;
          {
            int cx = x0, cy = y0, kcol, kstep = 0;
            if (main_src && halo) {  // group = kx; taps ky = 0..2 -> weight columns (ky*3 + kx)*C0
              cx = x0 + g - 1;
              cy = y0 - 1;
              kcol = g * c0_total + cb * kBK;
              kstep = 3 * c0_total;
            } else if (main_src) {  // group = tap (or the only tap of a 1x1)
              if (ntap0 == 9) {
                cx = x0 + g % 3 - 1;
                cy = y0 + g / 3 - 1;
              }
              kcol = g * c0_total + cb * kBK;
            } else {
              kcol = p.kb0 * kBK + skb * kBK;
            }
            // kGn: the transform warps own the WHOLE activation ring — they write the main operand's slots and issue
            // the TMA loads of the fused 1x1-skip operand's slots themselves, in ring order.  (Two producers taking
            // turns on one ring can pass each other: a parity wait that falls two phases behind a barrier returns
            // early and a slot is overwritten before it was consumed.  Hand-shake barriers between this warp and the
            // transform warps fixed that but tied the transform's look-ahead to this warp's, which the weight ring
            // keeps at 4 taps.)  This warp then only streams weights.
            if (!kGn) {
              mbar_wait(&empty_a[sa], pa ^ 1);
              __syncwarp();
              if (elect_one()) {
                uint8_t* dst = a_ring + sa * a_slot;
                if (kTwo) {
                  // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of the whole pair
                  if (lead_cta) mbar_arrive_expect_tx(&full_a[sa], 2 * a_bytes);
                  tma_load_4d_2sm(dst, ma, full_a0_cluster + static_cast<uint32_t>(sa) * 8u, a_ch, cx, cy, n0);
                } else {
                  mbar_arrive_expect_tx(&full_a[sa], a_bytes);
                  tma_load_4d(dst, ma, &full_a[sa], a_ch, cx, cy, n0);
                }
              }
            }
            if (++sa == n_a) {
              sa = 0;
              pa ^= 1;
            }
            for (int t = 0; t < nt; ++t) {
              mbar_wait(&empty_b[sb], pb ^ 1);
              __syncwarp();
              if (elect_one()) {
                uint8_t* dst = b_ring + sb * b_tile_bytes;
                if (kTwo) {
                  // the pair leader's barrier expects the bytes landing in BOTH CTAs of the pair, whoever sends them
                  if (lead_cta) mbar_arrive_expect_tx(&full_b[sb], static_cast<uint32_t>(2 * b_tile_bytes));
                  if (quad)
                    tma_load_2d_2sm_mc(dst + static_cast<int>(pair_in_cluster) * q_rows * (kBK * 2), &map_b,
                                       full_b0_cluster + static_cast<uint32_t>(sb) * 8u, kcol + t * kstep, b_row0, mc_mask);
                  else
                    tma_load_2d_2sm(dst, &map_b, full_b0_cluster + static_cast<uint32_t>(sb) * 8u, kcol + t * kstep, b_row0);
                } else {
                  mbar_arrive_expect_tx(&full_b[sb], static_cast<uint32_t>(b_tile_bytes));
                  tma_load_2d(dst, &map_b, &full_b[sb], kcol + t * kstep, b_row0);
                }
              }
              if (++sb == n_b) {
                sb = 0;
                pb ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Warp-uniform loop, one elected lane issues: descriptors stay in uniform registers, so each tcgen05.mma is
    // a single predicated instruction instead of a register->uniform "waterfall" loop (which made the issue
    // thread, not the tensor pipe, the bound of every N<=256 tile).
    if (lead_cta) {
      const uint32_t idesc = umma_idesc_f16(kTwo ? 2 * kBM : kBM, static_cast<uint32_t>(p.bn));
      const uint32_t a_ring_u32 = smem_base_u32 + static_cast<uint32_t>(kBarrierBytes + kBiasBytes);
      const uint32_t b_ring_u32 = a_ring_u32 + static_cast<uint32_t>(n_a * a_slot);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = work0; tile < total_tiles; tile += work_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        // the K schedule of the producer: per main channel block its groups, then its share of the skip blocks
        const int cb_lo = (tile / mn_tiles) * cb_split;
        const int cb_hi = cb_lo + cb_split < kb0_per_tap ? cb_lo + cb_split : kb0_per_tap;
        for (int cb = cb_lo; cb < cb_hi; ++cb) {
        const int n_items = ngrp0 + ((cb + 1) * nblk1 / kb0_per_tap - cb * nblk1 / kb0_per_tap);
        for (int it = 0; it < n_items; ++it) {
          const int nt = it < ngrp0 ? nt0 : 1;
          const bool last_item = cb == cb_hi - 1 && it == n_items - 1;
          mbar_wait(&full_a[sa], pa);
          const uint32_t a_base = a_ring_u32 + static_cast<uint32_t>(sa * a_slot);
          for (int t = 0; t < nt; ++t) {
            mbar_wait(&full_b[sb], pb);
            __syncwarp();
            tc_fence_after();
            const uint64_t a_desc = umma_smem_desc_sw128(a_base + static_cast<uint32_t>(t) * tap_stride);
            const uint64_t b_desc = umma_smem_desc_sw128(b_ring_u32 + static_cast<uint32_t>(sb * b_tile_bytes));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle atom: +2 in (addr >> 4) units
                const uint32_t accumulate = static_cast<uint32_t>(((cb - cb_lo) | it | t | k) != 0);
                if (kTwo)
                  umma_f16_2sm(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                               accumulate);
                else
                  umma_f16(d_tmem, a_desc + static_cast<uint64_t>(2 * k), b_desc + static_cast<uint64_t>(2 * k), idesc,
                           accumulate);
              }
              // free the weight slot (in both CTAs of a pair) once these MMAs have read it; after a group's last
              // tap also the activation slot, after the tile's last tap signal the accumulator
              if (kTwo) umma_commit_2sm(&empty_b[sb], quad ? static_cast<uint16_t>(0xF) : pair_mask);
              else umma_commit(&empty_b[sb]);
              if (t == nt - 1) {
                if (kTwo) umma_commit_2sm(&empty_a[sa], pair_mask);
                else umma_commit(&empty_a[sa]);
                if (last_item) {
                  if (kTwo) umma_commit_2sm(&tmem_full[acc], pair_mask);
                  else umma_commit(&tmem_full[acc]);
                }
              }
            }
            if (++sb == n_b) {
              sb = 0;
              pb ^= 1;
            }
          }
          if (++sa == n_a) {
            sa = 0;
            pa ^= 1;
          }
        }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (kGn) {
    // ------------------------------------------------------------------ GroupNorm transform (warps 2..7, kGn only)
    // A producer of the MAIN operand.  Per (tile, 64-channel block) the six warps read the RAW (BH+2) x (BW+2) halo of
    // the tile straight from global memory (16-byte chunks, one per lane and iteration, all in flight while the warp
    // waits for free slots), apply y = act(x * a + b) — the same arithmetic as gd_groupnorm_apply — ONCE per element
    // and write the result into the three kx-shifted (BH+2) x BW halo slots the MMA consumes as row-shifted ky views,
    // in the 128-byte-swizzled K-major layout a TMA box would have produced.  Pixels outside the image stay exactly 0:
    // the reference pads the NORMALISED tensor (unet.py:184-185 -> conv padding=1).  GD_CONV_GN_UPSAMPLE2 reads source
    // pixel (y/2, x/2): nearest x2 of the activated tensor (unet.py:191-195) without materialising it.
    // The loop is issue-bound (six warps, ~1.5 per scheduler), so everything that does not depend on the tile is
    // hoisted: per iteration a lane keeps its halo coordinates and its three destination offsets' ingredients packed
    // in one register, validity is a mask (no branches), and the per-channel affine comes from a precomputed table.
    const int tw = warp - 2;
    const int cch = lane & 7, sub = lane >> 3;
    const uint32_t a_ring_u32 = smem_base_u32 + static_cast<uint32_t>(kBarrierBytes + kBiasBytes);
    const uint32_t full_a0_cluster = kTwo ? mapa_u32(&full_a[0], lead_rank) : 0u;
    const bool up = p.gn_mode == GD_CONV_GN_UPSAMPLE2;
    const int hs = up ? (p.h >> 1) : p.h, ws = up ? (p.w >> 1) : p.w;
    const size_t img_stride = static_cast<size_t>(hs) * ws * p.gn_ld;
    constexpr float act_scale = 0.5f;  // SiLU(z) = h + h * tanh(h), h = z / 2: the 1/2 is folded into a and b (exact)
    // Lane geometry.  Iteration i of a lane covers halo pixel pidx0 + 24 i; 24 = 18 + 6, so the pixel's column repeats
    // with period 3 (xx_{i+3} = xx_i) while its row advances by 4: with j = i % 3 and m = i / 3 everything that depends
    // on the pixel splits into a per-lane constant of j (3 values, computed once per kernel) plus m times a constant —
    // source offsets, destination offsets (m * 4 rows * 16 pixels * 128 B = m * 8192, an immediate), swizzle terms and
    // store predicates; the per-pass address arithmetic shrinks to a handful of adds.
    // (y0 is a multiple of 8 and x0 of 16, so (y0 - 1 + yy) >> 1 == (y0 >> 1) + ((yy - 1) >> 1) for the x2 mode.)
    static_assert(kGnWarps * 4 == 24 && kGnHaloW == 18, "period-3 decomposition of the halo walk");
    const int pidx0 = tw * 4 + sub;  // halo pixel of iteration 0
    int relj[3];            // source element offset of (yy_j, xx_j) relative to the tile's first pixel
    int xxj[3], yyj[3];
    uint32_t dstj[3][3];    // [j][k]: byte offset inside the kx = k slot incl. the swizzle term, before + m * 8192
    uint32_t stmask = 0;    // bit 3 j + k: column xx_j - k lies inside the kx = k slot
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int pidx = pidx0 + 24 * j;
      const int yy = pidx / kGnHaloW, xx = pidx - yy * kGnHaloW;
      yyj[j] = yy;
      xxj[j] = xx;
      relj[j] = up ? (((yy - 1) >> 1) * ws + ((xx - 1) >> 1)) * p.gn_ld : ((yy - 1) * ws + (xx - 1)) * p.gn_ld;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int xk = xx - k;
        dstj[j][k] = static_cast<uint32_t>((yy * 16 + xk) * 128 + ((cch ^ (xk & 7)) << 4));
        if (xk >= 0 && xk < 16) stmask |= 1u << (3 * j + k);
      }
    }
    const int rel_m = (up ? 2 : 4) * ws * p.gn_ld;  // source offset of 4 halo rows (2 source rows in the x2 mode)
    const bool has7 = pidx0 + 24 * 7 < kGnHaloW * kGnHaloH;  // only the last iteration can fall off the 180-pixel halo
    int sa = 0;
    uint32_t pa = 0;
    uint8_t* a_ring_ptr = a_ring;
    // Software pipeline at HALF-pass granularity with one set of registers: a lane's 8 chunks form two halves; while
    // half A of channel block cb is normalised and stored, the loads of half B are in flight, and while B is processed
    // the loads of half A of the NEXT channel block (or of the next tile's first block) are in flight — the L2 latency
    // of the raw loads hides behind the other half's arithmetic instead of adding to every pass.
    uint32_t raw[kGnIters][4];
    constexpr int kHalf = kGnIters / 2;
    // source pointer (this lane's channel chunk of the tile's first pixel), affine table row and validity mask of a tile
    auto tile_state = [&](int tile, const __half*& img, const float4*& coef, uint32_t& vmask, int& n0, int& y0,
                          int& x0) {
      const int m_tile = m_tile_of(tile);
      tile_coords(p, m_tile, n0, y0, x0);
      const bool img_ok = n0 < p.n_img;  // (a tile count that is not a multiple of the cluster leaves CTAs without a tile)
      const int n_c = img_ok ? n0 : p.n_img - 1;
      // (only lanes whose pixel lies inside the image dereference img + rel[i])
      img = p.gn_src + static_cast<size_t>(n_c) * img_stride + cch * 8 +
            static_cast<ptrdiff_t>(up ? ((y0 >> 1) * ws + (x0 >> 1)) : (y0 * ws + x0)) * p.gn_ld;
      coef = reinterpret_cast<const float4*>(p.gn_coef) + (static_cast<size_t>(n_c) * (c0_total >> 3) + cch) * 4;
      vmask = 0;
#pragma unroll
      for (int i = 0; i < kGnIters; ++i) {
        const int y = y0 - 1 + yyj[i % 3] + 4 * (i / 3), x = x0 - 1 + xxj[i % 3];
        const bool ok = img_ok && (i < 7 || has7) && static_cast<unsigned>(y) < static_cast<unsigned>(p.h) &&
                        static_cast<unsigned>(x) < static_cast<unsigned>(p.w);
        vmask |= (ok ? 1u : 0u) << i;
      }
    };
    // raw chunks of one half: predicated loads, zeros where the pixel lies outside the image
    auto issue_half = [&](int first, const __half* img, uint32_t vmask, int cb) {
#pragma unroll
      for (int u = 0; u < kHalf; ++u) {
        const int i = first + u;
        const __half* src = img + (relj[i % 3] + (i / 3) * rel_m + cb * kBK);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
            "@p ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
            : "=r"(raw[i][0]), "=r"(raw[i][1]), "=r"(raw[i][2]), "=r"(raw[i][3])
            : "l"(src), "r"((vmask >> i) & 1u));
      }
    };
    // normalise one half in registers and store it into the three kx-shifted slots
    auto process_half = [&](int first, uint32_t vmask, const float (&ga)[8], const float (&gb)[8], uint32_t slot0,
                            uint32_t slot1, uint32_t slot2) {
#pragma unroll
      for (int u = 0; u < kHalf; ++u) {
        const int i = first + u;
        const uint32_t keep = 0u - ((vmask >> i) & 1u);  // all ones for pixels inside the image
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&raw[i][q]));
          // h = z / 2 (exact: a and b were halved); SiLU(z) = h + h * tanh(h) — gd_groupnorm_apply's silu_f
          const float h0 = fmaf(f.x, ga[2 * q], gb[2 * q]), h1 = fmaf(f.y, ga[2 * q + 1], gb[2 * q + 1]);
          const float r0 = fmaf(h0, tanh_approx(h0), h0);
          const float r1 = fmaf(h1, tanh_approx(h1), h1);
          const __half2 o = __floats2half2_rn(r0, r1);
          raw[i][q] = *reinterpret_cast<const uint32_t*>(&o) & keep;
        }
        const bool exists = i < 7 || has7;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          // row (yy*16 + xx - k) of the kx = k slot, 16-byte chunk cch ^ (row & 7): dstj[j][k] + m * (4 rows of 2 KiB)
          const uint32_t slot = k == 0 ? slot0 : (k == 1 ? slot1 : slot2);
          const uint32_t dst = slot + dstj[i % 3][k];
          const uint32_t pred = static_cast<uint32_t>(exists) & (stmask >> (3 * (i % 3) + k)) & 1u;
          if (i / 3 == 0) st_shared_v4_pred<0>(dst, raw[i], pred);
          else if (i / 3 == 1) st_shared_v4_pred<8192>(dst, raw[i], pred);
          else st_shared_v4_pred<16384>(dst, raw[i], pred);
        }
      }
    };
    const __half* img = p.gn_src;
    const float4* coef = reinterpret_cast<const float4*>(p.gn_coef);
    uint32_t vmask = 0;
    int t_n0 = 0, t_y0 = 0, t_x0 = 0;  // origin of the current tile (for the skip operand's TMA boxes)
    float4 cn0, cn1, cn2, cn3;  // affine (a[8], b[8]) of the NEXT channel block, prefetched
    cn0 = cn1 = cn2 = cn3 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (work0 < total_tiles) {
      tile_state(work0, img, coef, vmask, t_n0, t_y0, t_x0);
      issue_half(0, img, vmask, 0);
      cn0 = __ldg(coef), cn1 = __ldg(coef + 1), cn2 = __ldg(coef + 2), cn3 = __ldg(coef + 3);
    }
    for (int tile = work0; tile < total_tiles; tile += work_step) {
      for (int cb = 0; cb < kb0_per_tap; ++cb) {
        // ---- half A of this block: its loads were issued one half-step ago; half B's loads go out now
        issue_half(kHalf, img, vmask, cb);
        const float ga[8] = {cn0.x * act_scale, cn0.y * act_scale, cn0.z * act_scale, cn0.w * act_scale,
                             cn1.x * act_scale, cn1.y * act_scale, cn1.z * act_scale, cn1.w * act_scale};
        const float gb[8] = {cn2.x * act_scale, cn2.y * act_scale, cn2.z * act_scale, cn2.w * act_scale,
                             cn3.x * act_scale, cn3.y * act_scale, cn3.z * act_scale, cn3.w * act_scale};
        // the three slots of this channel block (kx = 0, 1, 2), in ring order
        const int s0 = sa;
        mbar_wait(&empty_a[s0], pa ^ 1);
        if (++sa == n_a) {
          sa = 0;
          pa ^= 1;
        }
        const int s1 = sa;
        mbar_wait(&empty_a[s1], pa ^ 1);
        if (++sa == n_a) {
          sa = 0;
          pa ^= 1;
        }
        const int s2 = sa;
        mbar_wait(&empty_a[s2], pa ^ 1);
        if (++sa == n_a) {
          sa = 0;
          pa ^= 1;
        }
        const uint32_t slot0 = a_ring_u32 + static_cast<uint32_t>(s0 * a_slot);
        const uint32_t slot1 = a_ring_u32 + static_cast<uint32_t>(s1 * a_slot);
        const uint32_t slot2 = a_ring_u32 + static_cast<uint32_t>(s2 * a_slot);
        process_half(0, vmask, ga, gb, slot0, slot1, slot2);
        // ---- half B: meanwhile half A of the next channel block (or of the next tile's first block) is fetched
        const __half* img_n = img;
        const float4* coef_n = coef;
        uint32_t vmask_n = vmask;
        int n0_n = t_n0, y0_n = t_y0, x0_n = t_x0;
        int cb_n = cb + 1;
        if (cb_n == kb0_per_tap) {
          cb_n = 0;
          if (tile + work_step < total_tiles) tile_state(tile + work_step, img_n, coef_n, vmask_n, n0_n, y0_n, x0_n);
          else vmask_n = 0;  // nothing left: no loads (the table row stays a valid address)
        }
        issue_half(0, img_n, vmask_n, cb_n);
        {
          const float4* cp = coef_n + cb_n * 32;
          cn0 = __ldg(cp), cn1 = __ldg(cp + 1), cn2 = __ldg(cp + 2), cn3 = __ldg(cp + 3);
        }
        process_half(kHalf, vmask, ga, gb, slot0, slot1, slot2);
        fence_proxy_async();  // generic-proxy stores -> visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (kTwo) {
            mbar_arrive_remote(full_a0_cluster + static_cast<uint32_t>(s0) * 8u);
            mbar_arrive_remote(full_a0_cluster + static_cast<uint32_t>(s1) * 8u);
            mbar_arrive_remote(full_a0_cluster + static_cast<uint32_t>(s2) * 8u);
          } else {
            mbar_arrive(&full_a[s0]);
            mbar_arrive(&full_a[s1]);
            mbar_arrive(&full_a[s2]);
          }
        }
        // This block's share of the fused 1x1-skip operand (raw, one tap per block): plain TMA boxes into the next ring
        // slots, issued by the first transform warp in ring order — the transform warps are the ring's only producer.
        const int s_lo = cb * nblk1 / kb0_per_tap, s_hi = (cb + 1) * nblk1 / kb0_per_tap;
        for (int skb = s_lo; skb < s_hi; ++skb) {
          if (tw == 0) {
            mbar_wait(&empty_a[sa], pa ^ 1);
            __syncwarp();
            if (elect_one()) {
              uint8_t* dst = a_ring_ptr + sa * a_slot;
              if (kTwo) {
                // both CTAs' boxes complete on the pair leader's barrier; its other arrivals stand in for the warps
                if (lead_cta) {
                  mbar_arrive_expect_tx(&full_a[sa], 2 * kATileBytes);
                  mbar_arrive_count(&full_a[sa], 2 * kGnWarps - 1);
                }
                tma_load_4d_2sm(dst, &map_a1, full_a0_cluster + static_cast<uint32_t>(sa) * 8u, skb * kBK, t_x0, t_y0, t_n0);
              } else {
                mbar_arrive_expect_tx(&full_a[sa], kATileBytes);
                mbar_arrive_count(&full_a[sa], kGnWarps - 1);
                tma_load_4d(dst, &map_a1, &full_a[sa], skb * kBK, t_x0, t_y0, t_n0);
              }
            }
          }
          if (++sa == n_a) {
            sa = 0;
            pa ^= 1;
          }
        }
        img = img_n;
        coef = coef_n;
        vmask = vmask_n;
        t_n0 = n0_n;
        t_y0 = y0_n;
        t_x0 = x0_n;
      }
    }
  }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9 / kGn: 8..15)
    if (kGn) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int quarter = warp & 3;          // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;   // row of the 128-row tile == pixel within the patch
    // Two independent epilogue groups (warps 2..5 and 6..9, one warp per TMEM lane quarter each) take alternate
    // 32-column sub-tiles of every accumulator, each with its own staging buffers, residual barrier and named
    // barrier: while one group waits (TMEM load, residual, TMA store read-out) the other computes.
    const int group = (warp - kEpiWarp0) >> 2;
    const bool leader = (static_cast<int>(threadIdx.x) - kEpiWarp0 * 32 - group * 128) == 0;
    const bool first_warp = ((warp - kEpiWarp0) & 3) == 0;  // the warp that holds the group's leader thread
    uint8_t* so = s_out + group * kEpiTileBytes;
    uint8_t* sr = s_res + group * kEpiTileBytes;
    uint64_t* rbar = &res_bar[group];
    // hand a drained accumulator back to the MMA issuer (pair mode: the issuer lives in the leader CTA)
    auto release_accumulator = [&](uint64_t* bar) {
      if (kTwo) mbar_arrive_cluster(mapa_u32(bar, lead_rank));
      else mbar_arrive(bar);
    };
    const int patch_px = p.bh * p.bw;
    const int i_local = row / patch_px;
    const int rem = row - i_local * patch_px;
    const int y_local = rem / p.bw;
    const int x_local = rem - y_local * p.bw;
    // row of this pixel's source inside the residual staging tile
    const bool res_tma = p.tma_epi && (p.res_mode == GD_RES_SAME || p.res_mode == GD_RES_UPSAMPLE2) && p.debug == 0;
    const int res_row = p.res_mode == GD_RES_UPSAMPLE2
                            ? (i_local * (p.bh >> 1) + (y_local >> 1)) * (p.bw >> 1) + (x_local >> 1)
                            : row;
    const uint32_t res_bytes = p.res_mode == GD_RES_UPSAMPLE2 ? kEpiTileBytes / 4 : kEpiTileBytes;
    const uint64_t store_policy = l2_policy_evict_first();
    uint32_t res_phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = work0; tile < total_tiles; tile += work_step) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = m_tile_of(tile);
      int n0, y0, x0;
      tile_coords(p, m_tile, n0, y0, x0);
      EpiCtx e;
      e.img = n0 + i_local;
      e.y = y0 + y_local;
      e.x = x0 + x_local;
      e.valid = (e.img < p.n_img) && (e.y < p.h) && (e.x < p.w);
      e.pix = (static_cast<size_t>(e.img) * p.h + e.y) * p.w + e.x;
      const int col_base = n_tile * p.bn;
      const int rx0 = p.res_mode == GD_RES_UPSAMPLE2 ? (x0 >> 1) : x0;
      const int ry0 = p.res_mode == GD_RES_UPSAMPLE2 ? (y0 >> 1) : y0;
      const int nsub = p.bn >> 5;  // 32-column sub-tiles; this group takes sub = group, group + 2, ...
      if (leader && res_tma) {
        // prefetch this group's first residual sub-tile while the mainloop is still running (the group consumed
        // its residual buffer before the last barrier of the previous tile)
        mbar_arrive_expect_tx(rbar, res_bytes);
        tma_load_4d(sr, &map_res, rbar, col_base + group * kEpiCols, rx0, ry0, n0);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256);

      if (p.debug == 1) {
        // timing experiment: no epilogue work at all
      } else if (p.ksplit > 1) {
        // split-K: raw fp32 accumulators -> this split's slab of the workspace (row = pixel-tile row, 128 B per thread
        // and sub-tile: whole sectors); bias, residual, fp16 rounding and statistics happen in the reduce kernel
        float* wrow = p.ws + (static_cast<size_t>((tile / mn_tiles) * p.m_tiles + m_tile) * kBM + row) * p.ws_ld + col_base;
        for (int j = group; j < nsub; j += 2) {
          uint32_t v[32];
          tmem_ld_x32(t_row + static_cast<uint32_t>(j * 32), v);
          tmem_ld_wait();
          if (m_tile < p.m_tiles) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(wrow + j * 32 + 4 * q) =
                  make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                              __uint_as_float(v[4 * q + 3]));
          }
        }
      } else if (!p.tma_epi) {
        for (int j = group; j < nsub; j += 2)
          epi_chunk_direct<32>(p, t_row + static_cast<uint32_t>(j * 32), col_base + j * 32, e);
        if ((p.bn & 16) != 0 && group == (nsub & 1))
          epi_chunk_direct<16>(p, t_row + static_cast<uint32_t>(nsub * 32), col_base + nsub * 32, e);
      } else {
        for (int sub = group; sub < nsub; sub += 2) {
          const int ch = sub * kEpiCols;  // first column of this sub-tile within the N tile
          uint32_t v[32];
          tmem_ld_x32(t_row + static_cast<uint32_t>(ch), v);
          // bias: the same 128 bytes for every thread -> broadcast loads through L1
          float4 bv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
            bv[q] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + col_base + ch) + q)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          Half8 rv[4];
          if (res_tma) {
            mbar_wait(rbar, res_phase);
            res_phase ^= 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) rv[q] = ld_half8_at(sw64(sr, res_row, q));
          } else if (p.res_mode == GD_RES_AVGPOOL2 && e.valid && p.debug == 0) {
            // residual lives at double resolution (unet.py:136,241): direct global reads, fp32 average -> fp16
            float racc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) racc[j] = 0.f;
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              const size_t rpix = (static_cast<size_t>(e.img) * (p.h * 2) + (2 * e.y + (s4 >> 1))) * (p.w * 2) +
                                  (2 * e.x + (s4 & 1));
              const __half* rp = p.res + rpix * p.ld_res + col_base + ch;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float t[8];
                half8_to_float(ld_half8(rp + 8 * q), t);
#pragma unroll
                for (int j = 0; j < 8; ++j) racc[8 * q + j] += t[j];
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float t[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) t[j] = 0.25f * racc[8 * q + j];
              rv[q] = float_to_half8(t);
            }
          }
          tmem_ld_wait();
          if (sub + 2 >= nsub) {
            // this thread has read its last columns of the accumulator: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            release_accumulator(&tmem_empty[acc]);
          }
          // the staging tile is free once the group's previous TMA store has finished READING it
          if (leader) bulk_wait_read0();
          epi_barrier(group);  // (A) staging tile reusable; every thread of the group has consumed the residual tile
          if (leader && res_tma && sub + 2 < nsub) {
            mbar_arrive_expect_tx(rbar, res_bytes);
            tma_load_4d(sr, &map_res, rbar, col_base + ch + 2 * kEpiCols, rx0, ry0, n0);
          }
          if (p.debug == 0 || p.debug == 3) {  // 3 = everything but the TMA store (timing experiments)
            float cs[16];  // [0,8): per-4-channel-chunk sums of this row, [8,16): sums of squares
            const bool has_res = p.res_mode != GD_RES_NONE;
            if (p.stats != nullptr) {
              if (has_res) epi_compute32<true, true>(v, bv, rv, so, row, cs);
              else epi_compute32<false, true>(v, bv, rv, so, row, cs);
            } else {
              if (has_res) epi_compute32<true, false>(v, bv, rv, so, row, cs);
              else epi_compute32<false, false>(v, bv, rv, so, row, cs);
            }
            if (p.stats != nullptr) {
              if (!e.valid) {  // rows outside the image contribute 0
#pragma unroll
                for (int i = 0; i < 16; ++i) cs[i] = 0.f;
              }
              // reduce the 16 values over the warp's 32 rows: a halving butterfly (8+4+2+1 shuffles) leaves value
              // (lane >> 1) in each lane pair, one more exchange completes it; even lanes store
#pragma unroll
              for (int off = 16, n = 8; off >= 2; off >>= 1, n >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < n; ++i) {
                  const float send = hi ? cs[i] : cs[i + n];
                  const float keep = hi ? cs[i + n] : cs[i];
                  cs[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
              }
              cs[0] += __shfl_xor_sync(0xffffffffu, cs[0], 1);
              const int idx = lane >> 1;  // 0..7: chunk sums, 8..15: chunk sums of squares
              if (p.stats_per_tile == 4) {
                if (m_tile < p.m_tiles && (lane & 1) == 0) {
                  float* sp = p.stats + (static_cast<size_t>(m_tile) * 4 + quarter) * p.stats_ld * 2;
                  sp[(((col_base + ch) >> 2) + (idx & 7)) * 2 + (idx >> 3)] = cs[0];
                }
              } else if ((lane & 1) == 0) {
                // the tile lies inside one image: park the quarter's sums, the group's first warp adds the four
                // after the barrier below -> 4x fewer partial rows for gd_groupnorm_finalize_partials to read
                s_stat[(group * 4 + quarter) * 16 + idx] = cs[0];
              }
            }
            fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          }
          epi_barrier(group);  // (B) staging tile complete
          if (p.stats != nullptr && p.stats_per_tile == 1 && p.debug == 0 && first_warp && lane < 16 &&
              m_tile < p.m_tiles) {
            // fixed summation order over the quarters -> bitwise reproducible; the next sub-tile's parked values are
            // written only after this warp has passed the group's next barrier (A)
            const float* q4 = s_stat + group * 64 + lane;
            const float tot = (q4[0] + q4[16]) + (q4[32] + q4[48]);
            float* sp = p.stats + static_cast<size_t>(m_tile) * p.stats_ld * 2;
            sp[(((col_base + ch) >> 2) + (lane & 7)) * 2 + (lane >> 3)] = tot;
          }
          if (leader && p.debug == 0 && m_tile < p.m_tiles) {  // (an odd tile count leaves the last pair half empty)
            tma_store_4d(&map_out, so, col_base + ch, x0, y0, n0, store_policy);
            bulk_commit();
          }
        }
      }
      if (!p.tma_epi || p.debug == 1) {
        tc_fence_before();
        release_accumulator(&tmem_empty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (leader && p.tma_epi) bulk_wait0();  // all stores complete before the CTA (and its smem) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (kTwo) cluster_sync_all();  // the peer may still be arriving on this CTA's barriers / reading its TMEM half
  if (warp == 1) {
    tc_fence_after();
    if (kTwo) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// Second half of a split-K convolution: out = fp16(sum over splits of the fp32 partial accumulators + bias (+ residual))
// and, where the conv was asked for them, the GroupNorm partial sums in exactly the row layout the fused epilogue emits
// (one row per 32-pixel quarter, or one per 128-pixel tile when the tile lies inside one image), taken from the fp32
// values like epi_compute32.  grid = (row blocks, 64-channel slabs), 8 warps: warp = 8-channel chunk, lane = pixel row.
// Fixed summation order (splits ascending, butterfly over lanes, quarters ascending) -> bitwise reproducible.
__global__ void __launch_bounds__(256) conv_splitk_reduce_kernel(const ConvArgs p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = p.stats_per_tile == 1 ? 4 : 1;  // quarters handled by one block
  const int m_tile = nq == 4 ? static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x >> 2);
  const int q0 = nq == 4 ? 0 : static_cast<int>(blockIdx.x & 3);
  const int col = static_cast<int>(blockIdx.y) * 64 + warp * 8;
  int n0, y0, x0;
  tile_coords(p, m_tile, n0, y0, x0);
  const int patch_px = p.bh * p.bw;
  const size_t split_stride = static_cast<size_t>(p.m_tiles) * kBM * p.ws_ld;
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (p.bias != nullptr) {
    b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + 1);
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f};  // sums of the two 4-channel chunks, then their sums of squares
  for (int qi = 0; qi < nq; ++qi) {
    const int row = (q0 + qi) * 32 + lane;
    const int i_local = row / patch_px;
    const int rem = row - i_local * patch_px;
    const int y_local = rem / p.bw;
    const int img = n0 + i_local, y = y0 + y_local, x = x0 + (rem - y_local * p.bw);
    const bool valid = img < p.n_img && y < p.h && x < p.w;
    const float* wp = p.ws + (static_cast<size_t>(m_tile) * kBM + row) * p.ws_ld + col;
    float4 a0 = b0, a1 = b1;
    for (int s = 0; s < p.ksplit; ++s) {
      const float4 u0 = *reinterpret_cast<const float4*>(wp + s * split_stride);
      const float4 u1 = *reinterpret_cast<const float4*>(wp + s * split_stride + 4);
      a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
      a1.x += u1.x; a1.y += u1.y; a1.z += u1.z; a1.w += u1.w;
    }
    float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float qs[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const size_t pix = (static_cast<size_t>(img) * p.h + y) * p.w + x;
      if (p.res_mode == GD_RES_SAME) {
        float t[8];
        half8_to_float(ld_half8(p.res + pix * p.ld_res + col), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += t[j];
      }
      st_half8(reinterpret_cast<__half*>(p.out) + pix * p.ld_out + col, float_to_half8(f));
      qs[0] = (f[0] + f[1]) + (f[2] + f[3]);
      qs[1] = (f[4] + f[5]) + (f[6] + f[7]);
      qs[2] = fmaf(f[3], f[3], fmaf(f[2], f[2], fmaf(f[1], f[1], f[0] * f[0])));
      qs[3] = fmaf(f[7], f[7], fmaf(f[6], f[6], fmaf(f[5], f[5], f[4] * f[4])));
    }
    if (p.stats != nullptr) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) qs[i] += __shfl_xor_sync(0xffffffffu, qs[i], off);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) cs[i] += qs[i];
    }
  }
  if (p.stats != nullptr && lane < 4) {
    const int prow = nq == 4 ? m_tile : m_tile * 4 + q0;
    float* sp = p.stats + static_cast<size_t>(prow) * p.stats_ld * 2;
    // lane 0/1: sums of chunk 0/1, lane 2/3: their sums of squares
    const float val = lane == 0 ? cs[0] : (lane == 1 ? cs[1] : (lane == 2 ? cs[2] : cs[3]));
    sp[((col >> 2) + (lane & 1)) * 2 + (lane >> 1)] = val;
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// box_c = 64 (mainloop operand tiles, SWIZZLE_128B) or 32 (epilogue staging tiles, SWIZZLE_64B)
int encode_act_map(CUtensorMap* m, const void* base, int c, int ld, int n, int h, int w, int bi, int bh, int bw,
                   int box_c = kBK) {
  EncodeTiledFn enc = get_encode_fn();
  GD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bi};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_c == kBK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: %d (c=%d ld=%d n=%d h=%d w=%d)", (int)r, c,
             ld, n, h, w);
  return 0;
}

int encode_weight_map(CUtensorMap* m, const void* base, int k_total, int n_pad, int bn) {
  EncodeTiledFn enc = get_encode_fn();
  GD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)n_pad};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: %d (k=%d n=%d bn=%d)", (int)r, k_total, n_pad,
             bn);
  return 0;
}

int g_num_sms = 0;
int g_debug_epilogue = 0;
int g_force_bn = 0;
int g_disable_tma_epi = 0;
int g_two_cta_mode = 1;
int g_halo_mode = 1;
int g_gn_na = 0;  // experiments: force the activation-ring depth of the fused-GroupNorm kernels
// Clusters of four CTAs with multicast weight tiles (GD_B200_QUAD=1 / gd_debug_set(8, 1)).  OFF by default: it is
// correct (the conv parity tests pass with it) but MEASURED slower — alone 1 498 -> 1 451 TFLOP/s (3x3 256->256 @256^2),
// 1 069 -> 1 027 (128->128), 1 555 -> 1 499 (512->512 @64^2), and 155.5 -> 156.4 ms per batch-64 step
// (profiles/quad_ab_r02.log): halving the L2 -> SM weight stream buys nothing because the mainloop is bound by the
// tensor pipe at the power-capped clock (ncu: 99.3 % tensor-pipe active), while the lock-step of two pairs on one
// weight ring and the 4-CTA cluster placement cost a few percent.
// Split-K is OPT-IN (GD_B200_SPLITK=1, devtools key 9): the number of splits follows the number of pixel tiles, i.e. the
// batch, and with it the fp32 summation order — a sample's bits would depend on the batch it is computed in, which every
// other kernel here avoids (tests/test_gpu_models.py::test_full_size_guided_step_properties) — for a measured 0-2.6 %
// of the batch-8 step (profiles/splitk_ab_r02.log).
int g_splitk_mode = -1;  // -1: read GD_B200_SPLITK
int g_quad_mode = [] {
  const char* e = getenv("GD_B200_QUAD");
  return e ? atoi(e) : 0;
}();
int g_max_quads[2] = {-1, -1};  // co-resident 4-CTA clusters per kernel variant (cudaOccupancyMaxActiveClusters)

}  // namespace

void conv_debug_set(int key, int value) {
  if (key == 0) g_debug_epilogue = value;
  if (key == 1) g_force_bn = value;
  if (key == 2) g_disable_tma_epi = value;
  if (key == 3) g_two_cta_mode = value;
  if (key == 4) g_halo_mode = value;
  if (key == 7) g_gn_na = value;
  if (key == 8) g_quad_mode = value;
  if (key == 9) g_splitk_mode = value ? 1 : 0;
}

// N tile: the largest divisor of n_pad (multiple of 16, <= 256) that still yields enough tiles to fill the SMs;
// small-spatial layers (8x8 / 16x16 with C = 1024) otherwise run on 4-16 CTAs.  Below 32 columns the MMA is
// shared-memory bound, so 32 is the floor unless n_pad itself is smaller.
int conv_pick_bn(int n_pad, int m_tiles, int num_sms, int step) {
  // Measured (profiles/conv_sweep.py, r01): narrower N tiles never pay off, even when they would fill more SMs —
  // every extra N tile re-streams the whole A operand from L2 and drops the MMA below its shared-memory-bound rate
  // (32x32x512ch: 37 us at BN=256 vs 99 us at BN=64).  So: the widest tile that divides n_pad.
  // One exception (profiles/conv_small_probe.py, r02): when the widest tile leaves three quarters of the SM pairs
  // without work — 16x16 layers at batch 8: 8 pixel-tile pairs x 4 N tiles — halving the tile to 128 columns doubles
  // the CTAs at the same per-CTA load rate: 54 -> 35 us (C_in 1024), 93 -> 60 us (C_in 2048).  With even fewer tiles
  // (8x8 at batch 8) the K loop of a single CTA is the floor either way; with more, 256 columns win (see above).
  const int pairs = (m_tiles + 1) / 2;
  if (n_pad % 256 == 0 && num_sms > 0 && pairs * (n_pad / 256) <= num_sms / 4) return 128;
  for (int bn = 256; bn >= step; bn -= step)
    if (n_pad % bn == 0) return bn;
  return 16;
}

// Output patch of one 128-row tile: BW x BH pixels of BI images, BW*BH*BI == 128 (powers of two).
bool patch_shape(int h, int w, int& bw, int& bh, int& bi) {
  bw = w < 16 ? w : 16;
  int p2 = 1;
  while (p2 * 2 <= bw) p2 *= 2;
  bw = p2;
  bh = 128 / bw;
  if (bh > h) {
    int q = 1;
    while (q * 2 <= h) q *= 2;
    bh = q;
  }
  bi = 128 / (bw * bh);
  return bi >= 1 && bi <= 256 && bw * bh * bi == 128;
}

// split-K applies to 3x3 convs with the aligned fp16 NHWC output whose (pixel unit, n tile) work items leave at least
// half of the GPU idle; the caller lends a workspace (gd_conv_desc.splitk_ws)
// SMs a persistent conv grid leaves free (GD_B200_CONV_SM_RESERVE): a conv CTA owns its SM outright (all shared memory,
// kGn: all registers), so while a full-width conv runs nothing of another stream can start.  In a guided step the
// classifier and the UNet are two branches of one graph; a few free SMs let one branch's latency-bound small kernels
// proceed under the other's long convolutions.
int g_sm_reserve = -1;
int conv_sms() {
  if (g_sm_reserve < 0) {
    const char* e = getenv("GD_B200_CONV_SM_RESERVE");
    g_sm_reserve = e != nullptr ? atoi(e) : 0;
    if (g_sm_reserve < 0 || g_sm_reserve > g_num_sms / 2) g_sm_reserve = 0;
  }
  return g_num_sms - g_sm_reserve;
}
bool splitk_static_ok(const gd_conv_desc* d) {
  if (g_splitk_mode < 0) {
    const char* e = getenv("GD_B200_SPLITK");
    g_splitk_mode = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_splitk_mode != 0 && d->gn_mode == GD_CONV_GN_OFF && d->taps == 9 && d->c0 >= 128 &&
         d->out_mode == GD_OUT_NHWC_F16 && d->cout % 64 == 0 && d->n_pad == d->cout && d->ld_out % 8 == 0 &&
         reinterpret_cast<uintptr_t>(d->out) % 16 == 0 && (d->out_scale == 1.0f || d->out_scale == 0.0f) &&
         (d->res_mode == GD_RES_NONE ||
          (d->res_mode == GD_RES_SAME && d->ld_res % 8 == 0 && reinterpret_cast<uintptr_t>(d->res) % 16 == 0)) &&
         reinterpret_cast<uintptr_t>(d->bias) % 16 == 0;
}
// number of splits for `items` work items on `slots` CTA (pair) slots, at most kmax
int splitk_pick(int items, int slots, int kb0_per_tap, long long per_split_bytes, long long ws_bytes) {
  int ks = items > 0 ? slots / items : 1;
  if (ks > 8) ks = 8;
  if (ks > kb0_per_tap) ks = kb0_per_tap;
  if (per_split_bytes > 0 && ks > ws_bytes / per_split_bytes) ks = static_cast<int>(ws_bytes / per_split_bytes);
  if (ks < 2) return 1;
  const int cbs = (kb0_per_tap + ks - 1) / ks;
  return (kb0_per_tap + cbs - 1) / cbs;  // no empty split
}

}  // namespace gd

extern "C" int64_t gd_conv_splitk_ws_bytes(const gd_conv_desc* d) {
  using namespace gd;
  if (d == nullptr || d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c0 <= 0 || !splitk_static_ok(d)) return 0;
  int bw, bh, bi;
  if (!patch_shape(d->h, d->w, bw, bh, bi)) return 0;
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return 0;
  }
  const int m_tiles = ((d->w + bw - 1) / bw) * ((d->h + bh - 1) / bh) * ((d->n + bi - 1) / bi);
  const bool two = m_tiles >= 2;
  const int units = two ? (m_tiles + 1) / 2 : m_tiles;
  const int slots = two ? g_num_sms / 2 : g_num_sms;
  const long long per_split = static_cast<long long>(m_tiles) * kBM * d->n_pad * 4;
  // upper bound over the N tilings the launch may choose (at least one n tile)
  const int ks = splitk_pick(units, slots, d->c0 / 64, per_split, 1ll << 62);
  return ks >= 2 ? ks * per_split : 0;
}

extern "C" int gd_conv_igemm(const gd_conv_desc* d, void* stream) {
  using namespace gd;
  GD_REQUIRE(d != nullptr, "gd_conv_igemm: null descriptor");
  GD_REQUIRE(d->a0 != nullptr && d->wpack != nullptr && d->out != nullptr, "gd_conv_igemm: null tensor pointer");
  GD_REQUIRE(d->taps == 9 || d->taps == 1, "gd_conv_igemm: taps must be 9 or 1, got %d", d->taps);
  GD_REQUIRE(d->c0 > 0 && d->c0 % 64 == 0, "gd_conv_igemm: C0 must be a positive multiple of 64, got %d", d->c0);
  GD_REQUIRE(d->ld0 >= d->c0 && d->ld0 % 8 == 0, "gd_conv_igemm: bad ld0 %d", d->ld0);
  const int c1 = d->a1 ? d->c1 : 0;
  GD_REQUIRE(c1 % 64 == 0, "gd_conv_igemm: C1 must be a multiple of 64, got %d", c1);
  if (d->a1) GD_REQUIRE(d->ld1 >= c1 && d->ld1 % 8 == 0, "gd_conv_igemm: bad ld1 %d", d->ld1);
  GD_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "gd_conv_igemm: bad geometry");
  const int k_total = d->taps * d->c0 + c1;
  GD_REQUIRE(d->k_total == k_total, "gd_conv_igemm: k_total %d != taps*C0+C1 = %d", d->k_total, k_total);
  GD_REQUIRE(d->n_pad % 16 == 0 && d->n_pad >= d->cout && d->cout > 0, "gd_conv_igemm: n_pad %d / cout %d invalid",
             d->n_pad, d->cout);
  if (d->bn > 0)
    GD_REQUIRE(d->bn % 16 == 0 && d->bn <= 256 && d->n_pad % d->bn == 0, "gd_conv_igemm: bad N tile %d for n_pad %d",
               d->bn, d->n_pad);
  GD_REQUIRE(d->out_mode == GD_OUT_NHWC_F16 || d->out_mode == GD_OUT_NCHW_F32, "gd_conv_igemm: bad out_mode");
  if (d->out_mode == GD_OUT_NHWC_F16)
    GD_REQUIRE(d->ld_out >= d->cout && (d->ld_out % 8 == 0 || d->cout % 16 != 0), "gd_conv_igemm: bad ld_out %d", d->ld_out);
  GD_REQUIRE(d->res_mode >= GD_RES_NONE && d->res_mode <= GD_RES_AVGPOOL2, "gd_conv_igemm: bad res_mode");
  if (d->res_mode != GD_RES_NONE) {
    GD_REQUIRE(d->res != nullptr && d->ld_res % 8 == 0 && d->cout % 16 == 0,
               "gd_conv_igemm: residual needs a pointer, ld%%8==0 and cout%%16==0");
    if (d->res_mode == GD_RES_UPSAMPLE2) GD_REQUIRE(d->h % 2 == 0 && d->w % 2 == 0, "gd_conv_igemm: odd upsample target");
  }

  int bw, bh, bi;
  GD_REQUIRE(patch_shape(d->h, d->w, bw, bh, bi), "gd_conv_igemm: cannot tile %dx%d into 128-pixel patches", d->h, d->w);
  const bool gn = d->gn_mode != GD_CONV_GN_OFF;
  if (gn) {
    GD_REQUIRE(d->gn_mode == GD_CONV_GN_SAME || d->gn_mode == GD_CONV_GN_UPSAMPLE2, "gd_conv_igemm: bad gn_mode %d",
               d->gn_mode);
    GD_REQUIRE(d->taps == 9 && gd_conv_gn_fusable(d->h, d->w) && g_halo_mode != 0,
               "gd_conv_igemm: a fused GroupNorm operand needs a 3x3 conv over images gd_conv_gn_fusable() accepts (%dx%d)",
               d->h, d->w);
    GD_REQUIRE(d->gn_silu == 1, "gd_conv_igemm: the fused GroupNorm operand always ends in SiLU (gn_silu must be 1): "
               "every GroupNorm that feeds a 3x3 conv in the reference is followed by one (unet.py:184, 208)");
    GD_REQUIRE(d->gn_coef != nullptr && reinterpret_cast<uintptr_t>(d->gn_coef) % 16 == 0,
               "gd_conv_igemm: fused GroupNorm needs the 16-byte aligned affine table gn_coef (gd_groupnorm_coef)");
    GD_REQUIRE(reinterpret_cast<uintptr_t>(d->a0) % 16 == 0, "gd_conv_igemm: fused GroupNorm operand must be 16-byte aligned");
    if (d->gn_mode == GD_CONV_GN_UPSAMPLE2)
      GD_REQUIRE(d->h % 2 == 0 && d->w % 2 == 0, "gd_conv_igemm: fused upsample needs even output size");
    const long long src_px = static_cast<long long>(d->gn_mode == GD_CONV_GN_UPSAMPLE2 ? d->h / 2 : d->h) *
                             (d->gn_mode == GD_CONV_GN_UPSAMPLE2 ? d->w / 2 : d->w);
    GD_REQUIRE(src_px * d->ld0 < (1ll << 31), "gd_conv_igemm: fused GroupNorm operand image too large for 32-bit offsets");
    if (d->a1)
      GD_REQUIRE(reinterpret_cast<uintptr_t>(d->a1) % 16 == 0 && static_cast<long long>(d->h) * d->w * d->ld1 < (1ll << 31),
                 "gd_conv_igemm: fused GroupNorm: the 1x1 source must be 16-byte aligned and < 2^31 elements per image");
  }
  const bool want_stats = d->stats_out != nullptr;
  if (want_stats)
    GD_REQUIRE(d->out_mode == GD_OUT_NHWC_F16 && d->cout % 64 == 0 && d->n_pad == d->cout && bw * bh >= 32 &&
                   (d->bn == 0 || d->bn % 64 == 0),
               "gd_conv_igemm: fused GroupNorm statistics need fp16 NHWC output, cout %% 64 == 0 and >= 32 pixels/image");
  if (want_stats)
    GD_REQUIRE(bi == 1 || ((d->w + bw - 1) / bw) * ((d->h + bh - 1) / bh) == 1,
               "gd_conv_igemm: fused GroupNorm statistics are not defined for %dx%d images (several images per tile and "
               "several tiles per image); gd_conv_stats_rows returns 0 for this geometry", d->h, d->w);

  if (g_num_sms == 0) {
    int dev = 0;
    GD_CHECK_CUDA(cudaGetDevice(&dev));
    GD_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int m_tiles = ((d->w + bw - 1) / bw) * ((d->h + bh - 1) / bh) * ((d->n + bi - 1) / bi);
  int bn = d->bn > 0 ? d->bn
                     : (g_force_bn > 0 && d->n_pad % g_force_bn == 0 && !want_stats
                            ? g_force_bn
                            : conv_pick_bn(d->n_pad, m_tiles, g_num_sms, want_stats ? 64 : 16));
  GD_REQUIRE(bn % 16 == 0 && bn >= 16 && bn <= 256 && d->n_pad % bn == 0, "gd_conv_igemm: bad N tile %d for n_pad %d", bn,
             d->n_pad);

  ConvArgs p;
  p.debug = g_debug_epilogue;
  p.quad = 0;
  p.n_img = d->n;
  p.h = d->h;
  p.w = d->w;
  p.bi = bi;
  p.bh = bh;
  p.bw = bw;
  p.patches_x = (d->w + bw - 1) / bw;
  p.patches_y = (d->h + bh - 1) / bh;
  p.m_tiles = m_tiles;
  p.n_tiles = d->n_pad / bn;
  p.bn = bn;
  // CTA-pair mode whenever there are at least two pixel tiles and half a weight tile is a whole number of 8-row
  // swizzle groups (bn % 16 == 0 always holds)
  const bool two_cta = g_two_cta_mode != 0 && m_tiles >= 2 && (bn / 2) % 8 == 0 && g_num_sms >= 2;
  // halo mode: a 3x3 conv whose 128-pixel tile is BH whole rows x BW (multiple of 8) pixels of ONE image
  const bool halo = g_halo_mode != 0 && d->taps == 9 && bi == 1 && bw % 8 == 0 && bw * bh == 128;
  const int b_tile = (two_cta ? bn / 2 : bn) * kBK * 2;
  const int a_slot = halo ? (bh + 2) * bw * kBK * 2 : kATileBytes;
  const int ring_budget = kSmemBudget - kBarrierBytes - kBiasBytes - kEpiBytes - 1024;
  int n_a, n_b;
  if (gn) {
    // the transform warps fill the three kx slots of a channel block at once: two whole groups (6 slots) decouple them
    // from the MMA completely (measured at N = 256: 1480 / 1500 / 1532 TFLOP/s for 4 / 5 / 6 slots, with only 4 weight
    // slots left in the last case); fewer slots only where not even 4 weight slots would fit
    n_a = 4;
    n_b = (ring_budget - n_a * a_slot) / b_tile;
    for (int na = 6; na > 4; --na) {
      const int nb = (ring_budget - na * a_slot) / b_tile;
      if (nb >= 4) {
        n_a = na;
        n_b = nb;
        break;
      }
    }
    if (g_gn_na >= 4 && g_gn_na <= kMaxA && (ring_budget - g_gn_na * a_slot) / b_tile >= 2) {
      n_a = g_gn_na;
      n_b = (ring_budget - n_a * a_slot) / b_tile;
    }
  } else if (halo) {
    // every activation slot feeds 3 taps: pick the split that keeps the most taps in flight (narrow N tiles are
    // latency-bound on the 20 KiB halo copies unless many of them are outstanding)
    int best = -1;
    n_a = 2;
    n_b = 2;
    for (int na = 2; na <= kMaxA; ++na) {
      int nb = (ring_budget - na * a_slot) / b_tile;
      if (nb > kMaxB) nb = kMaxB;
      if (nb < 2) break;
      const int in_flight = 3 * na < nb ? 3 * na : nb;
      if (in_flight > best) {
        best = in_flight;
        n_a = na;
        n_b = nb;
      }
    }
  } else {
    n_a = n_b = ring_budget / (a_slot + b_tile);
  }
  if (n_a > kMaxA) n_a = kMaxA;
  if (n_b > kMaxB) n_b = kMaxB;
  GD_REQUIRE(n_a >= 2 && n_b >= 2, "gd_conv_igemm: not enough shared memory for the operand rings");
  p.halo = halo ? 1 : 0;
  p.a_slot_bytes = a_slot;
  p.n_a = n_a;
  p.n_b = n_b;
  p.kb0_per_tap = d->c0 / 64;
  p.taps = d->taps;
  p.kb0 = d->taps * p.kb0_per_tap;
  p.kb_total = p.kb0 + c1 / 64;
  p.cout = d->cout;
  p.bias = d->bias;
  p.res = reinterpret_cast<const __half*>(d->res);
  p.ld_res = d->ld_res;
  p.res_mode = d->res_mode;
  p.out = d->out;
  p.ld_out = d->ld_out;
  p.out_mode = d->out_mode;
  p.out_scale = d->out_scale == 0.0f ? 1.0f : d->out_scale;
  // staged TMA-store epilogue: fp16 NHWC, whole 64-channel sub-tiles, 16-byte aligned rows; an upsampled residual
  // additionally needs an even patch
  p.tma_epi = (!g_disable_tma_epi && d->out_scale == 1.0f && d->out_mode == GD_OUT_NHWC_F16 && d->cout % 64 == 0 && bn % 64 == 0 &&
               d->ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out) % 16 == 0) &&
               !(d->res_mode == GD_RES_UPSAMPLE2 && (bw % 2 || bh % 2)))
                  ? 1
                  : 0;
  GD_REQUIRE(!want_stats || p.tma_epi, "gd_conv_igemm: fused statistics need the staged epilogue (aligned fp16 output)");
  p.stats = d->stats_out;
  p.stats_ld = d->n_pad / 4;
  p.stats_per_tile = bi == 1 ? 1 : 4;
  p.gn_mode = d->gn_mode;
  p.gn_silu = d->gn_silu;
  p.gn_src = reinterpret_cast<const __half*>(d->a0);
  p.gn_ld = d->ld0;
  p.gn_src1 = reinterpret_cast<const __half*>(d->a1);
  p.gn_ld1 = d->a1 ? d->ld1 : 0;
  p.gn_coef = d->gn_coef;
  p.ksplit = 1;
  p.cb_split = p.kb0_per_tap;
  p.ws = nullptr;
  p.ws_ld = d->n_pad;
  ConvArgs p_reduce;
  if (d->splitk_ws != nullptr && d->splitk_ws_bytes > 0 && bn % 32 == 0 && g_debug_epilogue == 0 &&
      reinterpret_cast<uintptr_t>(d->splitk_ws) % 16 == 0 && splitk_static_ok(d)) {
    const int units = two_cta ? (m_tiles + 1) / 2 : m_tiles;
    const long long per_split = static_cast<long long>(m_tiles) * kBM * d->n_pad * 4;
    const int ks = splitk_pick(units * p.n_tiles, two_cta ? g_num_sms / 2 : g_num_sms, p.kb0_per_tap, per_split,
                               d->splitk_ws_bytes);
    if (ks >= 2) {
      p.ksplit = ks;
      p.cb_split = (p.kb0_per_tap + ks - 1) / ks;
      p.ws = reinterpret_cast<float*>(d->splitk_ws);
      p_reduce = p;  // the reduce kernel finishes the epilogue: bias, residual, fp16 output, statistics
      p.tma_epi = 0;
      p.stats = nullptr;
      p.bias = nullptr;
      p.res_mode = GD_RES_NONE;
    }
  }

  CUtensorMap ma0, ma1, mb, mout, mres;
  // (a fused-GroupNorm operand is read with plain loads; its tensor map is only a placeholder, encoded over the source
  //  geometry so that it stays valid)
  const int a0_h = d->gn_mode == GD_CONV_GN_UPSAMPLE2 ? d->h / 2 : d->h;
  const int a0_w = d->gn_mode == GD_CONV_GN_UPSAMPLE2 ? d->w / 2 : d->w;
  int rc = gn ? encode_act_map(&ma0, d->a0, d->c0, d->ld0, d->n, a0_h, a0_w, 1, 1, 8)
              : encode_act_map(&ma0, d->a0, d->c0, d->ld0, d->n, d->h, d->w, bi, halo ? bh + 2 : bh, bw);
  if (rc) return rc;
  if (d->a1) {
    rc = encode_act_map(&ma1, d->a1, c1, d->ld1, d->n, d->h, d->w, bi, bh, bw);
    if (rc) return rc;
  } else {
    ma1 = ma0;
  }
  // quad mode: clusters of 4 = two CTA pairs on consecutive pixel-tile pairs of the same N tile; every CTA fetches a
  // quarter of the weight tile and multicasts it to the other pair, which halves the L2 -> SM weight stream (the
  // largest operand stream of the mainloop: 590 KB of weights against 245 KB of activations per 128 x 256 tile).
  // Needs whole 8-row swizzle groups per quarter and at least one work item per cluster slot to be worth the lock-step.
  const bool quad = two_cta && p.ksplit == 1 && g_quad_mode != 0 && (bn / 4) % 8 == 0 && m_tiles >= 8 &&
                    ((m_tiles + 3) / 4) * (d->n_pad / bn) >= g_num_sms / 8;
  p.quad = quad ? 1 : 0;
  rc = encode_weight_map(&mb, d->wpack, k_total, d->n_pad, quad ? bn / 4 : (two_cta ? bn / 2 : bn));
  if (rc) return rc;
  mout = ma0;
  mres = ma0;
  if (p.tma_epi) {
    rc = encode_act_map(&mout, d->out, d->cout, d->ld_out, d->n, d->h, d->w, bi, bh, bw, kEpiCols);
    if (rc) return rc;
    if (d->res_mode == GD_RES_SAME) {
      rc = encode_act_map(&mres, d->res, d->cout, d->ld_res, d->n, d->h, d->w, bi, bh, bw, kEpiCols);
      if (rc) return rc;
    } else if (d->res_mode == GD_RES_UPSAMPLE2) {
      rc = encode_act_map(&mres, d->res, d->cout, d->ld_res, d->n, d->h / 2, d->w / 2, bi, bh / 2, bw / 2, kEpiCols);
      if (rc) return rc;
    }
  }

  // Always request (almost) the full shared memory so exactly one CTA owns an SM and its 512 TMEM columns.
  const int smem_bytes = kSmemBudget;
  {
    // function attributes are per device: remember which devices of this process have them
    static bool attr_set[64] = {};
    int dev = 0;
    GD_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      GD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      GD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      GD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      GD_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  const int threads = gn ? kThreadsGn : kThreads;
  if (two_cta) {
    // one cluster of 2 (or 4) CTAs per pixel-tile pair (or two pairs); persistent over (m unit, n tile) work items
    const int csize = quad ? 4 : 2;
    const int items = ((p.m_tiles + csize - 1) / csize) * p.n_tiles * p.ksplit;
    int max_clusters = conv_sms() / csize;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    if (quad) {
      // how many 4-CTA clusters the GPU can hold at once (GPC boundaries may leave SMs unused): ask the driver once
      int& cached = g_max_quads[gn ? 1 : 0];
      if (cached < 0) {
        cfg.gridDim = dim3(4 * max_clusters);
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        int n = 0;
        cudaError_t e = gn ? cudaOccupancyMaxActiveClusters(&n, conv_igemm_kernel<true, true>, &cfg)
                           : cudaOccupancyMaxActiveClusters(&n, conv_igemm_kernel<true, false>, &cfg);
        cached = (e == cudaSuccess && n > 0) ? n : max_clusters;
        (void)cudaGetLastError();
      }
      if (cached < max_clusters) max_clusters = cached;
    }
    const int clusters = items < max_clusters ? items : max_clusters;
    cfg.gridDim = dim3(csize * clusters);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    if (gn) GD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, true>, ma0, ma1, mb, mout, mres, p));
    else GD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true, false>, ma0, ma1, mb, mout, mres, p));
  } else {
    const int total_tiles = p.m_tiles * p.n_tiles * p.ksplit;
    const int grid = total_tiles < conv_sms() ? total_tiles : conv_sms();
    if (gn)
      GD_CHECK_CUDA(launch_pdl(conv_igemm_kernel<false, true>, dim3(grid), dim3(threads), smem_bytes,
                               reinterpret_cast<cudaStream_t>(stream), ma0, ma1, mb, mout, mres, p));
    else
      GD_CHECK_CUDA(launch_pdl(conv_igemm_kernel<false, false>, dim3(grid), dim3(threads), smem_bytes,
                               reinterpret_cast<cudaStream_t>(stream), ma0, ma1, mb, mout, mres, p));
  }
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  if (p.ksplit > 1) {
    const dim3 rgrid(static_cast<unsigned>(p.m_tiles * (p_reduce.stats_per_tile == 1 ? 1 : 4)),
                     static_cast<unsigned>(d->cout / 64));
    conv_splitk_reduce_kernel<<<rgrid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p_reduce);
    GD_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
  }
  return 0;
}

extern "C" int gd_conv_gn_fusable(int32_t h, int32_t w) {
  int bw, bh, bi;
  if (h <= 0 || w <= 0 || !gd::patch_shape(h, w, bw, bh, bi)) return 0;
  return (bi == 1 && bw == 16 && bh == 8) ? 1 : 0;
}

extern "C" int64_t gd_conv_stats_rows(int32_t n, int32_t h, int32_t w, int32_t* rows_per_image) {
  int bw, bh, bi;
  if (n <= 0 || h <= 0 || w <= 0 || !gd::patch_shape(h, w, bw, bh, bi) || bw * bh < 32) {
    if (rows_per_image) *rows_per_image = 0;
    return 0;
  }
  const int tiles_per_group = ((w + bw - 1) / bw) * ((h + bh - 1) / bh);
  // Several images per tile AND several tiles per image (small images that are not a power of two, e.g. 12x12: two
  // images per 8x8 tile, four tiles per image): the per-quarter rows of one image would interleave with its tile
  // mate's, but gd_groupnorm_finalize_partials reads a contiguous block of rows per image -> no fused statistics
  // for that geometry (the caller falls back to gd_groupnorm_stats).
  if (bi > 1 && tiles_per_group > 1) {
    if (rows_per_image) *rows_per_image = 0;
    return 0;
  }
  const int groups = (n + bi - 1) / bi;
  // a tile that lies inside one image (bi == 1) emits ONE partial row, otherwise one per 32-pixel quarter
  const int per_tile = bi == 1 ? 1 : 4;
  if (rows_per_image) *rows_per_image = tiles_per_group * per_tile / bi;
  return static_cast<int64_t>(groups) * tiles_per_group * per_tile;
}
