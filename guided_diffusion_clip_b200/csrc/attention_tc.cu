// Fused multi-head attention forward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), head dim 64.
// Same contract as attn_fwd_kernel in attention.cu (QKVAttentionLegacy.forward unet.py:337-354, QKVAttention.forward
// unet.py:370-389: softmax((q*s)(k*s)^T) v with s = 64^-1/4, softmax in fp32) for sequence lengths that are a
// multiple of 128 (the 32x32 and 16x16 attention resolutions of the 256x256 models, and the 256-padded ViT sequence).
//
// One CTA = 128 queries of one (sample, head); two CTAs per SM (256 TMEM columns, 113 KiB shared memory each).
//   warp 4      TMA producer: the Q tile once, then a 2-stage ring of K / V tiles (128 keys x 64, SWIZZLE_128B boxes
//               cut straight out of the token-major qkv tensor -- no repacking for either channel order).
//   warp 5      MMA issuer: S = Q K^T  (M=128, N=128, 4 k-steps, both operands K-major) into TMEM columns [0,128);
//               PV = P V (M=128, N=64, 8 k-steps; P from shared memory K-major, V as an MN-MAJOR operand -- the key
//               dimension is the strided one in the V tile) into TMEM columns [128,192), not accumulated.
//   warps 0-3   softmax: thread r owns query row r = TMEM lane r.  Two passes over S with tcgen05.ld (row maximum,
//               then exp2 -> fp16 P written into the swizzled A-operand layout), running sum and the output
//               accumulator O (64 fp32) live in registers:  O <- O * 2^((m_old - m_new) c) + PV_tile.  The PV tile of
//               step j is folded in during step j+1 (between the two passes), so the tensor pipe and the MUFU-bound
//               softmax overlap without a second S buffer.
// The dominant cost is the 128 x 128 exponentials per tile (MUFU, 16/clk/SM), not the MMAs: at d = 64 the kernel's
// ceiling is the exp roofline, ~2x above the tensor-pipe time of the two GEMMs.
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

constexpr int kD = 64;
constexpr int kBQ = 128;
constexpr int kBKV = 128;
constexpr int kTcThreads = 192;
constexpr uint32_t kTile = kBKV * kD * 2;                  // 16 KiB: one 128 x 64 fp16 tile
constexpr uint32_t kOffQ = 0;
constexpr uint32_t kOffK = kOffQ + kTile;                  // 2 stages
constexpr uint32_t kOffV = kOffK + 2 * kTile;              // 2 stages
constexpr uint32_t kOffP = kOffV + 2 * kTile;              // 128 x 128 fp16 = two K-atoms of 16 KiB
constexpr uint32_t kOffBar = kOffP + 2 * kTile;            // 114688
constexpr uint32_t kSmemBytes = kOffBar + 1024;            // 2 CTAs/SM: 2 * (115712 + 1024 reserved) = 233472 = 228 KiB
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
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
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kTcThreads, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, __half* __restrict__ out, int ld_out,
                   float* __restrict__ lse, int t, int t_valid, int heads, int order) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

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
    mbar_init(&kv_full[0], 1);
    mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1);
    mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;         // S: columns [0,128)
  const uint32_t tmem_pv = tmem_base + 128;  // PV tile: columns [128,192)

  if (warp == 4) {
    // ---------------- TMA producer ----------------
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kTile);
      tma_load_2d(smem + kOffQ, &map_qkv, q_full, qcol, row0 + qb * kBQ);
      for (int j = 0; j < nkv; ++j) {
        const int st = j & 1;
        if (j >= 2) mbar_wait(&kv_empty[st], static_cast<uint32_t>(((j >> 1) - 1) & 1));
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTile);
        tma_load_2d(smem + kOffK + st * kTile, &map_qkv, &kv_full[st], kcol, row0 + j * kBKV);
        tma_load_2d(smem + kOffV + st * kTile, &map_qkv, &kv_full[st], vcol, row0 + j * kBKV);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc_s = umma_idesc_f16(128, 128);
    const uint32_t idesc_pv = umma_idesc_f16(128, 64) | (1u << 16);  // B (= V tile) is MN-major
    const uint64_t q_desc = umma_smem_desc_sw128(sbase + kOffQ);
    const uint64_t p_desc = umma_smem_desc_sw128(sbase + kOffP);
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t k_desc = umma_smem_desc_sw128(sbase + kOffK);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16(tmem_s, q_desc + static_cast<uint64_t>(2 * k), k_desc + static_cast<uint64_t>(2 * k), idesc_s,
                 static_cast<uint32_t>(k != 0));
      umma_commit(s_full);
    }
    __syncwarp();
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1;
      mbar_wait(p_full, static_cast<uint32_t>(j & 1));  // P_j is in shared memory; S and the PV tile have been read
      if (j + 1 < nkv) mbar_wait(&kv_full[(j + 1) & 1], static_cast<uint32_t>(((j + 1) >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        if (j + 1 < nkv) {  // the softmax warps wait for this one first
          const uint64_t k_desc = umma_smem_desc_sw128(sbase + kOffK + static_cast<uint32_t>((j + 1) & 1) * kTile);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_s, q_desc + static_cast<uint64_t>(2 * k), k_desc + static_cast<uint64_t>(2 * k), idesc_s,
                     static_cast<uint32_t>(k != 0));
          umma_commit(s_full);
        }
        // PV = P V: k-step kk covers keys [16 kk, 16 kk + 16): P advances 32 B inside its 64-key atom (atoms 16 KiB
        // apart), V advances 16 rows of 128 B (a whole number of 1024-byte swizzle groups)
        const uint64_t v_desc = umma_smem_desc_sw128(sbase + kOffV + static_cast<uint32_t>(st) * kTile);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a = p_desc + static_cast<uint64_t>(((kk >> 2) * kTile + (kk & 3) * 32) >> 4);
          const uint64_t b = v_desc + static_cast<uint64_t>((kk * 2048) >> 4);
          umma_f16(tmem_pv, a, b, idesc_pv, static_cast<uint32_t>(kk != 0));
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ---------------- softmax / accumulate: thread = query row = TMEM lane ----------------
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const float sc = 0.125f * kLog2e;  // (64^-1/4)^2 folded with log2(e)
    float o[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, corr_pend = 1.f;
    const uint32_t p_row = sbase + kOffP + static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, static_cast<uint32_t>(j & 1));
      tc_fence_after();
      const int valid = t_valid - j * kBKV;  // >= 1; < 128 only in a masked last tile
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_s + lane_addr + c * 32, v);
        tmem_ld_wait();
        if (valid >= kBKV) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = ex2_approx((m_run - m_new) * sc);  // first tile: 2^-inf = 0
      m_run = m_new;
      const float msc = m_new * sc;
      // fold in the PV tile of the previous step (also frees the PV columns and the P buffer for this step)
      if (j > 0) {
        mbar_wait(o_full, static_cast<uint32_t>((j - 1) & 1));
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_x32(tmem_pv + lane_addr, v0);
        tmem_ld_x32(tmem_pv + lane_addr + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o[i] = fmaf(o[i], corr_pend, __uint_as_float(v0[i]));
          o[32 + i] = fmaf(o[32 + i], corr_pend, __uint_as_float(v1[i]));
        }
      }
      corr_pend = corr;
      // pass 2: P = 2^(S c - m c) -> fp16, into the K-major SWIZZLE_128B A-operand layout
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_s + lane_addr + c * 32, v);
        tmem_ld_wait();
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), sc, -msc));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), sc, -msc));
          if (valid < kBKV) {
            if (c * 32 + 2 * i >= valid) p0 = 0.f;
            if (c * 32 + 2 * i + 1 >= valid) p1 = 0.f;
          }
          rs += p0 + p1;
          h[i] = pack_h2(p0, p1);
        }
        const uint32_t atom = p_row + static_cast<uint32_t>(c >> 1) * kTile;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t chunk = static_cast<uint32_t>((c & 1) * 4 + q);
          sts128(atom + ((chunk ^ sw) << 4), h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        }
      }
      l_run = fmaf(l_run, corr, rs);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }
    // last PV tile
    mbar_wait(o_full, static_cast<uint32_t>((nkv - 1) & 1));
    tc_fence_after();
    {
      uint32_t v0[32], v1[32];
      tmem_ld_x32(tmem_pv + lane_addr, v0);
      tmem_ld_x32(tmem_pv + lane_addr + 32, v1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o[i] = fmaf(o[i], corr_pend, __uint_as_float(v0[i]));
        o[32 + i] = fmaf(o[32 + i], corr_pend, __uint_as_float(v1[i]));
      }
    }
    const float inv = 1.0f / l_run;
    // stage the row in the (now idle) Q tile, swizzled so that both the row-wise writes and the coalesced read-back
    // are bank-conflict free; each warp only touches its own 32 rows
    const uint32_t stage_row = sbase + kOffQ + static_cast<uint32_t>(row) * 128u;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      sts128(stage_row + ((static_cast<uint32_t>(q) ^ sw) << 4), pack_h2(o[8 * q] * inv, o[8 * q + 1] * inv),
             pack_h2(o[8 * q + 2] * inv, o[8 * q + 3] * inv), pack_h2(o[8 * q + 4] * inv, o[8 * q + 5] * inv),
             pack_h2(o[8 * q + 6] * inv, o[8 * q + 7] * inv));
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
      lse[(static_cast<size_t>(n) * heads + head) * t + qb * kBQ + row] = m_run * 0.125f + logf(l_run);
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

}  // namespace

void attn_debug_set(int value) { g_attn_tc_enabled = value; }

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
  static bool configured = false;
  if (!configured) {
    GD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    GD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  dim3 grid(t / kBQ, heads, n);
  attn_fwd_tc_kernel<<<grid, kTcThreads, kSmemBytes, stream>>>(map, reinterpret_cast<__half*>(out), ld_out, lse, t, t_valid,
                                                               heads, order);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace gd
