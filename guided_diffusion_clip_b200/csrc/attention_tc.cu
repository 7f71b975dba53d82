// Fused multi-head attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64.
// Same contract as attn_fwd_kernel in attention.cu (QKVAttentionLegacy.forward unet.py:337-354, QKVAttention.forward
// unet.py:370-389: softmax((q*s)(k*s)^T) v with s = 64^-1/4, softmax in fp32) for sequence lengths that are a
// multiple of 128 (the 32x32 and 16x16 attention resolutions of the 256x256 models, and the 256-padded ViT sequence).
//
// One CTA = 128 queries of one (sample, head); two CTAs per SM (256 TMEM columns, 113 KiB shared memory each).
//   warp 4      TMA producer: the Q tile once, then one K and one V stage (128 keys x 64, SWIZZLE_128B boxes cut
//               straight out of the token-major qkv tensor -- no repacking for either channel order); a stage is
//               refilled as soon as the MMAs that read it have completed, a full step before it is needed again.
//   warp 5      MMA issuer: S = Q K^T  (M=128, N=128, 4 k-steps, both operands K-major) into TMEM columns [0,128);
//               O += P V (M=128, N=64, 8 k-steps; P from shared memory K-major, V as an MN-MAJOR operand -- the key
//               dimension is the strided one in the V tile) ACCUMULATED in TMEM columns [128,192) over all key tiles.
//   warps 0-3   softmax: thread r owns query row r = TMEM lane r.  The whole 128-column row of S is pulled into
//               registers with one batch of tcgen05.ld (S is released to the MMA warp at once, so S_{j+1} is computed
//               while step j does its exponentials), row maximum, 2^(s c - m c) with packed FFMA2 / FADD2, fp16 P
//               written into the swizzled A-operand layout (double buffered).  The softmax warpgroup runs with 208
//               registers per thread (setmaxnreg; warps 4-7 drop to 40).  The output accumulator never lives in
//               registers: the
//               reference maximum m is only raised when the row maximum has grown by more than 2^8 (P stays <= 256,
//               exact after the final division by the row sum, which uses the same m), and only then does the thread
//               rescale its TMEM row of O (tcgen05.ld / st) -- on real data a handful of times per row.
// The dominant cost is the 128 x 128 exponentials per tile (MUFU, 16/clk/SM), not the MMAs: at d = 64 the kernel's
// ceiling is the exp roofline (~2x the tensor-pipe time of the two GEMMs).
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

constexpr int kD = 64;
constexpr int kBQ = 128;
constexpr int kBKV = 128;
constexpr int kTcThreads = 256;                            // warps 0-3 softmax, 4 TMA, 5 MMA, 6-7 idle (register donors)
constexpr uint32_t kTile = kBKV * kD * 2;                  // 16 KiB: one 128 x 64 fp16 tile
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + kTile;                  // 1 stage: free again as soon as S_j has been computed
constexpr uint32_t kOffV = kOffK + kTile;                  // 1 stage: free again as soon as PV_j has been computed
constexpr uint32_t kOffP = kOffV + kTile;                  // 2 buffers of 128 x 128 fp16 (two K-atoms of 16 KiB each)
constexpr uint32_t kOffBar = kOffP + 4 * kTile;            // 114688
constexpr uint32_t kSmemBytes = kOffBar + 1024;            // 2 CTAs/SM: 2 * (115712 + 1024 reserved) = 233472 = 228 KiB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// packed fp32 pairs (FFMA2 / FADD2 / FMUL2 on sm_100): halve the issue slots of the per-score scale, the row sum and
// the accumulator rescale -- the softmax warps are issue-bound next to the MUFU
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 ra, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// 2^a for a PAIR of scores on the FMA / integer pipes instead of the MUFU (which is the bound of the softmax: one
// ex2 per score at 16 per clock and SM).  Cody-Waite with the round-to-nearest magic number: x = a + 1.5*2^23 carries
// n = rint(a) in its low mantissa bits, r = a - n lies in [-0.5, 0.5], 2^r is a degree-4 polynomial (relative error
// 4.5e-5, an order of magnitude below the fp16 rounding of P) and 2^n is added to the exponent field with one shift-add
// (the magic number's own bits shift out).  a is clamped at -120: 2^-120 rounds to the same fp16 zero as ex2(-inf).
__device__ __forceinline__ void ex2_poly2(float a0, float a1, float& p0, float& p1) {
  constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23
  a0 = fmaxf(a0, -120.0f);
  a1 = fmaxf(a1, -120.0f);
  float x0, x1, n0, n1, r0, r1;
  fadd2(x0, x1, a0, a1, kMagic, kMagic);
  fadd2(n0, n1, x0, x1, -kMagic, -kMagic);
  ffma2(r0, r1, n0, n1, -1.0f, -1.0f, a0, a1);
  float q0, q1;
  ffma2(q0, q1, r0, r1, 0.009618129f, 0.009618129f, 0.055504109f, 0.055504109f);
  ffma2(q0, q1, q0, q1, r0, r1, 0.240226507f, 0.240226507f);
  ffma2(q0, q1, q0, q1, r0, r1, 0.693147181f, 0.693147181f);
  ffma2(q0, q1, q0, q1, r0, r1, 1.0f, 1.0f);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(x0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(x1) << 23));
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// kPolyMask: which of the 16 score pairs of every 32-column chunk take the polynomial exponential (bit i = pair i)
template <uint32_t kPolyMask>
__global__ void __launch_bounds__(kTcThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, __half* __restrict__ out, int ld_out,
                   float* __restrict__ lse, int t, int t_valid, int heads, int order) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   // K_j landed
  uint64_t* k_empty = bars + 2;  // S_j computed: the K stage is free
  uint64_t* v_full = bars + 3;   // V_j landed
  uint64_t* v_empty = bars + 4;  // PV_j computed: the V stage is free
  uint64_t* s_full = bars + 5;   // S_j complete in TMEM
  uint64_t* s_free = bars + 6;   // S_j copied to registers by all 128 rows: S_{j+1} may be issued
  uint64_t* p_full = bars + 7;   // P_j in shared memory (and any rescale of O done): PV_j may be issued
  uint64_t* o_full = bars + 8;   // PV_j complete: P buffer j & 1 reusable, O readable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int qb = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  int qcol, kcol, vcol;
  if (order == GD_QKV_LEGACY) {
    qcol = head * 3 * kD;
    kcol = qcol + kD;
    vcol = qcol + 2 * kD;
  } else {
    qcol = head * kD;
    kcol = qcol + heads * kD;
    vcol = qcol + 2 * heads * kD;
  }
  const int row0 = n * t;
  const int nkv = (t_valid + kBKV - 1) / kBKV;  // key tiles beyond the valid length are never touched

  if (threadIdx.x == 0 && (sbase & 1023u) != 0) {
    printf("gd: attn_fwd_tc shared memory base %u is not 1024-byte aligned\n", sbase);
    __trap();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;        // S: columns [0,128)
  const uint32_t tmem_o = tmem_base + 128;  // O accumulator: columns [128,192)
  pdl_enter();  // barrier init and the TMEM allocation above overlap the predecessor kernel's tail

  if (warp >= 4) {
    // the non-softmax warpgroup needs few registers; the softmax warpgroup takes them (128 fp32 scores per thread)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 4) {
      // ---------------- TMA producer (loads in the order their stages become free) ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, kTile);
        tma_load_2d(smem + kOffQ, &map_qkv, q_full, qcol, row0 + qb * kBQ);
        mbar_arrive_expect_tx(k_full, kTile);
        tma_load_2d(smem + kOffK, &map_qkv, k_full, kcol, row0);
        mbar_arrive_expect_tx(v_full, kTile);
        tma_load_2d(smem + kOffV, &map_qkv, v_full, vcol, row0);
        // K_{j+1} can go once S_j is complete (beginning of step j-1); V_{j+1} once PV_j is complete (beginning of
        // step j+1).  In time order: K_1, K_2, then V_{i+1}, K_{i+3} for i = 0, 1, ...
        for (int j = 1; j <= 2 && j < nkv; ++j) {
          mbar_wait(k_empty, static_cast<uint32_t>((j - 1) & 1));
          mbar_arrive_expect_tx(k_full, kTile);
          tma_load_2d(smem + kOffK, &map_qkv, k_full, kcol, row0 + j * kBKV);
        }
        for (int i = 0; i + 1 < nkv; ++i) {
          mbar_wait(v_empty, static_cast<uint32_t>(i & 1));
          mbar_arrive_expect_tx(v_full, kTile);
          tma_load_2d(smem + kOffV, &map_qkv, v_full, vcol, row0 + (i + 1) * kBKV);
          if (i + 3 < nkv) {
            mbar_wait(k_empty, static_cast<uint32_t>((i + 2) & 1));
            mbar_arrive_expect_tx(k_full, kTile);
            tma_load_2d(smem + kOffK, &map_qkv, k_full, kcol, row0 + (i + 3) * kBKV);
          }
        }
      }
      __syncwarp();
    } else if (warp == 5) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc_s = umma_idesc_f16(128, 128);
      const uint32_t idesc_pv = umma_idesc_f16(128, 64) | (1u << 16);  // B (= V tile) is MN-major
      const uint64_t q_desc = umma_smem_desc_sw128(sbase + kOffQ);
      const uint64_t k_desc = umma_smem_desc_sw128(sbase + kOffK);
      const uint64_t v_desc = umma_smem_desc_sw128(sbase + kOffV);
      mbar_wait(q_full, 0);
      for (int j = 0; j <= nkv; ++j) {
        // S_j: at j = 0 straight away, otherwise as soon as S_{j-1} sits in the softmax warps' registers
        if (j < nkv) {
          if (j >= 1) mbar_wait(s_free, static_cast<uint32_t>((j - 1) & 1));
          mbar_wait(k_full, static_cast<uint32_t>(j & 1));
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem_s, q_desc + static_cast<uint64_t>(2 * k), k_desc + static_cast<uint64_t>(2 * k), idesc_s,
                       static_cast<uint32_t>(k != 0));
            umma_commit(s_full);
            umma_commit(k_empty);
          }
          __syncwarp();
        }
        // O (+)= P_{j-1} V_{j-1}.  k-step kk covers keys [16 kk, 16 kk + 16): P advances 32 B inside its 64-key atom
        // (atoms 16 KiB apart), V advances 16 rows of 128 B (a whole number of 1024-byte swizzle groups)
        if (j >= 1) {
          const int i = j - 1;
          mbar_wait(p_full, static_cast<uint32_t>(i & 1));
          mbar_wait(v_full, static_cast<uint32_t>(i & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint64_t p_desc = umma_smem_desc_sw128(sbase + kOffP + static_cast<uint32_t>(i & 1) * 2u * kTile);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t a = p_desc + static_cast<uint64_t>(((kk >> 2) * kTile + (kk & 3) * 32) >> 4);
              const uint64_t b = v_desc + static_cast<uint64_t>((kk * 2048) >> 4);
              umma_f16(tmem_o, a, b, idesc_pv, static_cast<uint32_t>((i | kk) != 0));
            }
            umma_commit(o_full);
            umma_commit(v_empty);
          }
          __syncwarp();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ---------------- softmax: thread = query row = TMEM lane ----------------
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float sc = 0.125f * kLog2e;  // (64^-1/4)^2 folded with log2(e)
    const float kRaise = 8.0f / sc;    // raise the reference maximum only when the row maximum grew by > 2^8
    float m_ref = 0.f, l_run = 0.f;
    const uint32_t sw = static_cast<uint32_t>(row & 7);

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, static_cast<uint32_t>(j & 1));
      tc_fence_after();
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tmem_s + lane_addr + c * 32, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);  // S_{j+1} overlaps everything below
      const int valid = t_valid - j * kBKV;  // >= 1; < 128 only in a masked last tile
      if (valid < kBKV) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) v[c][i] = 0xff800000u;  // -inf: probability 0
      }
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mx[c] = fmaxf(__uint_as_float(v[c][0]), __uint_as_float(v[c][1]));
#pragma unroll
        for (int i = 2; i < 32; i += 2)
          mx[c] = fmaxf(mx[c], fmaxf(__uint_as_float(v[c][i]), __uint_as_float(v[c][i + 1])));
      }
      const float mxr = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float corr = 1.f;
      if (j == 0) {
        m_ref = mxr;
      } else if (mxr > m_ref + kRaise) {
        corr = ex2_approx((m_ref - mxr) * sc);
        m_ref = mxr;
      }
      if (__any_sync(0xffffffffu, corr != 1.f)) {  // rare: rescale this warp's 32 rows of O in TMEM
        mbar_wait(o_full, static_cast<uint32_t>((j - 1) & 1));  // PV_{j-1} has been accumulated
        tc_fence_after();
#pragma unroll
        for (int hcol = 0; hcol < 2; ++hcol) {
          uint32_t ov[32];
          tmem_ld_x32(tmem_o + lane_addr + hcol * 32, ov);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
          tmem_st_x32(tmem_o + lane_addr + hcol * 32, ov);
        }
        tmem_st_wait();
        l_run *= corr;
      }
      // P_j -> buffer j & 1 (PV_{j-2}, its last reader, was waited for at the end of step j-1)
      const uint32_t p_row = sbase + kOffP + static_cast<uint32_t>(j & 1) * 2u * kTile + static_cast<uint32_t>(row) * 128u;
      const float nmsc = -m_ref * sc;
      float rs[4][2];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        rs[c][0] = rs[c][1] = 0.f;
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float a0, a1;
          ffma2(a0, a1, __uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1]), sc, sc, nmsc, nmsc);
          float p0, p1;
          if ((kPolyMask >> i) & 1u) {
            ex2_poly2(a0, a1, p0, p1);
          } else {
            p0 = ex2_approx(a0);
            p1 = ex2_approx(a1);
          }
          fadd2(rs[c][0], rs[c][1], rs[c][0], rs[c][1], p0, p1);
          h[i] = pack_h2(p0, p1);
        }
        const uint32_t atom = p_row + static_cast<uint32_t>(c >> 1) * kTile;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(atom + ((static_cast<uint32_t>((c & 1) * 4 + q) ^ sw) << 4), h[4 * q], h[4 * q + 1], h[4 * q + 2],
                 h[4 * q + 3]);
      }
      l_run += ((rs[0][0] + rs[0][1]) + (rs[1][0] + rs[1][1])) + ((rs[2][0] + rs[2][1]) + (rs[3][0] + rs[3][1]));
      tc_fence_before();
      fence_proxy_async_smem();
      // Observe PV_{j-1}'s completion (issued a whole step ago: no stall) BEFORE releasing P_j: once p_full is
      // complete PV_j may finish too, and a parity wait that is two phases behind would never return.  It also tells
      // the next step that P buffer (j+1) & 1 is free.
      if (j > 0) mbar_wait(o_full, static_cast<uint32_t>((j - 1) & 1));
      mbar_arrive(p_full);
    }
    // all key tiles accumulated
    mbar_wait(o_full, static_cast<uint32_t>((nkv - 1) & 1));
    tc_fence_after();
    float o[kD];
    {
      uint32_t v0[32], v1[32];
      tmem_ld_x32(tmem_o + lane_addr, v0);
      tmem_ld_x32(tmem_o + lane_addr + 32, v1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = __uint_as_float(v0[i]);
        o[32 + i] = __uint_as_float(v1[i]);
      }
    }
    const float inv = 1.0f / l_run;
#pragma unroll
    for (int i = 0; i < kD; i += 2) fmul2(o[i], o[i + 1], o[i], o[i + 1], inv, inv);
    // stage the row in the (now idle) Q tile, swizzled so that both the row-wise writes and the coalesced read-back
    // are bank-conflict free; each warp only touches its own 32 rows
    const uint32_t stage_row = sbase + kOffQ + static_cast<uint32_t>(row) * 128u;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      sts128(stage_row + ((static_cast<uint32_t>(q) ^ sw) << 4), pack_h2(o[8 * q], o[8 * q + 1]),
             pack_h2(o[8 * q + 2], o[8 * q + 3]), pack_h2(o[8 * q + 4], o[8 * q + 5]), pack_h2(o[8 * q + 6], o[8 * q + 7]));
    __syncwarp();
    __half* o_base = out + (static_cast<size_t>(row0) + static_cast<size_t>(qb) * kBQ) * ld_out + head * kD;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rr = warp * 32 + i * 4 + (lane >> 3);
      const uint32_t ch = static_cast<uint32_t>(lane & 7);
      const uint4 v = lds128(sbase + kOffQ + static_cast<uint32_t>(rr) * 128u + ((ch ^ static_cast<uint32_t>(rr & 7)) << 4));
      *reinterpret_cast<uint4*>(o_base + static_cast<size_t>(rr) * ld_out + ch * 8) = v;
    }
    if (lse != nullptr)  // natural-log LSE of the scaled scores: max/8 + ln(sum)
      lse[(static_cast<size_t>(n) * heads + head) * t + qb * kBQ + row] = m_ref * 0.125f + logf(l_run);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

int g_attn_tc_enabled = 1;
// Share of the exponentials computed with the polynomial on the FMA pipe (index into the instantiated masks).  Measured
// (profiles/attn_sweep.py, r02, 8 heads x T = 1024 at batch 64): 0/16 637.5, 4/16 663.3, 6/16 638.0, 8/16 609.0 TFLOP/s —
// the softmax is co-limited by issue slots (the polynomial costs ~5.5 slots per score against 1.5 for the MUFU path),
// so only a quarter of the scores is worth moving.
int g_attn_poly = 1;

}  // namespace

// 0 / 1: mma.sync or tcgen05 forward; 10 + k: select the polynomial-exponential share k (0: none, 1: 4/16, 2: 6/16, 3: 8/16)
void attn_debug_set(int value) {
  if (value >= 10) g_attn_poly = value - 10;
  else g_attn_tc_enabled = value;
}

// true when the tcgen05 kernel covers this call (otherwise the caller runs the mma.sync kernel)
bool attn_fwd_tc_applicable(const void* qkv, int ld_qkv, const void* out, int ld_out, int t) {
  return g_attn_tc_enabled && t % kBKV == 0 && ld_out % 8 == 0 && ld_qkv % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
}

int attn_fwd_tc_launch(const void* qkv, int ld_qkv, void* out, int ld_out, float* lse, int n, int t, int t_valid,
                       int heads, int order, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn();
  GD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * heads * kD), (cuuint64_t)n * (cuuint64_t)t};
  cuuint64_t strides[1] = {(cuuint64_t)ld_qkv * 2};
  cuuint32_t box[2] = {(cuuint32_t)kD, (cuuint32_t)kBKV};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(qkv) failed: %d (heads=%d ld=%d n=%d t=%d)", (int)r, heads, ld_qkv,
             n, t);
  auto k0 = attn_fwd_tc_kernel<0x0000u>;
  auto k1 = attn_fwd_tc_kernel<0x1111u>;   // 4 of 16 pairs
  auto k2 = attn_fwd_tc_kernel<0x4925u>;   // 6 of 16
  auto k3 = attn_fwd_tc_kernel<0x5555u>;   // 8 of 16
  static unsigned long long configured_on[2] = {0, 0};
  if (gd::first_use_on_device(configured_on)) {
    for (auto k : {k0, k1, k2, k3}) {
      GD_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
      GD_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
  }
  dim3 grid(t / kBQ, heads, n);
  auto kern = g_attn_poly == 0 ? k0 : g_attn_poly == 1 ? k1 : g_attn_poly == 3 ? k3 : k2;
  GD_CHECK_CUDA(launch_pdl(kern, grid, dim3(kTcThreads), kSmemBytes, stream, map,
                           reinterpret_cast<__half*>(out), ld_out, lse, t, t_valid, heads, order));
  count_launch(1);
  return 0;
}

}  // namespace gd
