// Fused multi-head attention forward for head dimensions other than 64 (16 .. 256, multiple of 16).
//
// The reference's factory default is num_heads=4 / num_head_channels=-1 (script_util.py:54-56), so the head width
// follows the channel count of the level: 96 / 128 for the default 64x64 model, 128 / 192 / 256 for the 128x128
// model of README.md:56.  Same math as attention.cu (QKVAttentionLegacy.forward unet.py:337-354, QKVAttention.forward
// unet.py:370-389): softmax((q*s)(k*s)^T) v with s = d^-1/4, fp32 online softmax, the [T,T] weights never exist.
//
// CTA = 64 queries of one (sample, head), 4 warps x 16 query rows, key/value tiles of 64 tokens double buffered
// with cp.async; rows are padded by 16 bytes (row stride 2*D+16: ldmatrix phases of 8 rows hit 8 distinct bank
// groups for every D that is a multiple of 16 — the XOR swizzle of the D = 64 kernel needs 8 chunks per row).
// Q fragments are re-read from shared memory per key tile (they would cost D/4 registers), the output accumulator
// (D/2 registers per thread) stays in registers.  Forward only: the classifier, the only model that is
// differentiated, hard-codes 64-wide heads (script_util.py:265).
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

constexpr int kHQ = 64, kHKV = 64;
constexpr float kLog2eH = 1.4426950408889634f;

__device__ __forceinline__ void hd_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void hd_ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void hd_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hd_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int D>
__device__ __forceinline__ void hd_load_tile(uint32_t smem_base, const __half* g, int ld, int rows_valid, int tid) {
  constexpr int kRS = 2 * D + 16, kChunks = D / 8;
  for (int idx = tid; idx < 64 * kChunks; idx += 128) {
    const int row = idx / kChunks, chunk = idx - row * kChunks;
    const uint32_t dst = smem_base + static_cast<uint32_t>(row * kRS + chunk * 16);
    if (row < rows_valid) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + static_cast<size_t>(row) * ld + chunk * 8)
                   : "memory");
    } else {
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128)
attn_fwd_hd_kernel(const __half* __restrict__ qkv, int ld_qkv, __half* __restrict__ out, int ld_out,
                   float* __restrict__ lse, int t, int heads, int order, float scale) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int kRS = 2 * D + 16;       // bytes per shared-memory row
  constexpr int kTile = 64 * kRS;
  constexpr int kKS = D / 16;           // k-steps of Q K^T = pairs of 8-wide output column tiles of P V
  const int qb = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sQ = smem_u32(smem), sK0 = sQ + kTile, sV0 = sK0 + 2 * kTile;
  const __half* qkv_n = qkv + static_cast<size_t>(n) * t * ld_qkv;
  const __half *gq, *gk, *gv;
  if (order == GD_QKV_LEGACY) {  // [head][q,k,v][D]
    gq = qkv_n + head * 3 * D;
    gk = gq + D;
    gv = gq + 2 * D;
  } else {                       // [q,k,v][head][D]
    gq = qkv_n + head * D;
    gk = gq + heads * D;
    gv = gq + 2 * heads * D;
  }
  const int q0 = qb * kHQ;
  hd_load_tile<D>(sQ, gq + static_cast<size_t>(q0) * ld_qkv, ld_qkv, min(kHQ, t - q0), tid);
  hd_load_tile<D>(sK0, gk, ld_qkv, min(kHKV, t), tid);
  hd_load_tile<D>(sV0, gv, ld_qkv, min(kHKV, t), tid);
  asm volatile("cp.async.commit_group;" ::: "memory");

  const int nkv = (t + kHKV - 1) / kHKV;
  float o[2 * kKS][4];
#pragma unroll
  for (int j = 0; j < 2 * kKS; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sc = scale * kLog2eH;

  for (int kv = 0; kv < nkv; ++kv) {
    const int buf = kv & 1;
    if (kv + 1 < nkv) {
      const int k1 = (kv + 1) * kHKV;
      hd_load_tile<D>(sK0 + (buf ^ 1) * kTile, gk + static_cast<size_t>(k1) * ld_qkv, ld_qkv, min(kHKV, t - k1), tid);
      hd_load_tile<D>(sV0 + (buf ^ 1) * kTile, gv + static_cast<size_t>(k1) * ld_qkv, ld_qkv, min(kHKV, t - k1), tid);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // S = Q K^T (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
    const uint32_t kbase = sK0 + buf * kTile;
#pragma unroll
    for (int kk = 0; kk < kKS; ++kk) {
      uint32_t qf[4];
      hd_ldmatrix_x4(qf, sQ + static_cast<uint32_t>((warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kRS +
                                                     (kk * 2 + (lane >> 4)) * 16));
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t kf[4];
        hd_ldmatrix_x4(kf, kbase + static_cast<uint32_t>((jp * 16 + (lane & 7) + 8 * (lane >> 4)) * kRS +
                                                         (kk * 2 + ((lane >> 3) & 1)) * 16));
        hd_mma16816(s[2 * jp], qf, kf[0], kf[1]);
        hd_mma16816(s[2 * jp + 1], qf, kf[2], kf[3]);
      }
    }
    if ((kv + 1) * kHKV > t) {  // ragged last key tile: keys >= t get probability 0
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kv * kHKV + j * 8 + (lane & 3) * 2 + (e & 1) >= t) s[j][e] = -INFINITY;
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r]);
      corr[r] = exp2f((m_run[r] - m_new[r]) * sc);
      m_run[r] = m_new[r];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = exp2f((s[j][0] - m_new[0]) * sc);
      s[j][1] = exp2f((s[j][1] - m_new[0]) * sc);
      s[j][2] = exp2f((s[j][2] - m_new[1]) * sc);
      s[j][3] = exp2f((s[j][3] - m_new[1]) * sc);
      rs[0] += s[j][0] + s[j][1];
      rs[1] += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int j = 0; j < 2 * kKS; ++j) {
      o[j][0] *= corr[0];
      o[j][1] *= corr[0];
      o[j][2] *= corr[1];
      o[j][3] *= corr[1];
    }
    // O += P V
    const uint32_t vbase = sV0 + buf * kTile;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4];
      pf[0] = hd_pack(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = hd_pack(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = hd_pack(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = hd_pack(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < kKS; ++jp) {
        uint32_t vf[4];
        hd_ldmatrix_x4_trans(vf, vbase + static_cast<uint32_t>((kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kRS +
                                                               (jp * 2 + (lane >> 4)) * 16));
        hd_mma16816(o[2 * jp], pf, vf[0], vf[1]);
        hd_mma16816(o[2 * jp + 1], pf, vf[2], vf[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int row0 = q0 + warp * 16 + (lane >> 2);
  __half* o_n = out + static_cast<size_t>(n) * t * ld_out + head * D;
#pragma unroll
  for (int j = 0; j < 2 * kKS; ++j) {
    const int col = j * 8 + (lane & 3) * 2;
    if (row0 < t)
      *reinterpret_cast<__half2*>(o_n + static_cast<size_t>(row0) * ld_out + col) =
          __floats2half2_rn(o[j][0] * inv0, o[j][1] * inv0);
    if (row0 + 8 < t)
      *reinterpret_cast<__half2*>(o_n + static_cast<size_t>(row0 + 8) * ld_out + col) =
          __floats2half2_rn(o[j][2] * inv1, o[j][3] * inv1);
  }
  if (lse != nullptr && (lane & 3) == 0) {
    float* l = lse + (static_cast<size_t>(n) * heads + head) * t;
    if (row0 < t) l[row0] = m_run[0] * scale + logf(l_run[0]);
    if (row0 + 8 < t) l[row0 + 8] = m_run[1] * scale + logf(l_run[1]);
  }
}

template <int D>
int launch_hd(const void* qkv, int ld_qkv, void* out, int ld_out, float* lse, int n, int t, int heads, int order,
              cudaStream_t st) {
  constexpr int kSmem = 5 * 64 * (2 * D + 16);
  static unsigned long long configured_on[2] = {0, 0};
  if (gd::first_use_on_device(configured_on)) {
    GD_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_hd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  attn_fwd_hd_kernel<D><<<dim3((t + kHQ - 1) / kHQ, heads, n), 128, kSmem, st>>>(
      reinterpret_cast<const __half*>(qkv), ld_qkv, reinterpret_cast<__half*>(out), ld_out, lse, t, heads, order, scale);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int gd_attention_fwd_hd(const void* qkv, int32_t ld_qkv, void* out, int32_t ld_out, float* lse, int32_t n,
                                   int32_t t, int32_t heads, int32_t head_dim, int32_t order, void* stream) {
  GD_REQUIRE(qkv && out, "gd_attention_fwd_hd: null pointer");
  GD_REQUIRE(n > 0 && t > 0 && heads > 0, "gd_attention_fwd_hd: bad n/t/heads");
  GD_REQUIRE(order == GD_QKV_LEGACY || order == GD_QKV_NEW, "gd_attention_fwd_hd: bad qkv order %d", order);
  GD_REQUIRE(ld_qkv >= 3 * heads * head_dim && ld_qkv % 8 == 0 && ld_out >= heads * head_dim && ld_out % 2 == 0 &&
                 (reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0,
             "gd_attention_fwd_hd: rows must hold all heads and be 16-byte aligned (ld_qkv %d, ld_out %d)", ld_qkv, ld_out);
  GD_REQUIRE(heads <= 65535 && n <= 65535, "gd_attention_fwd_hd: grid too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define GD_HD_CASE(DD) \
  case DD: return launch_hd<DD>(qkv, ld_qkv, out, ld_out, lse, n, t, heads, order, st)
  switch (head_dim) {
    GD_HD_CASE(16);
    GD_HD_CASE(32);
    GD_HD_CASE(48);
    GD_HD_CASE(64);
    GD_HD_CASE(80);
    GD_HD_CASE(96);
    GD_HD_CASE(112);
    GD_HD_CASE(128);
    GD_HD_CASE(160);
    GD_HD_CASE(192);
    GD_HD_CASE(224);
    GD_HD_CASE(256);
    default: break;
  }
#undef GD_HD_CASE
  GD_REQUIRE(false, "gd_attention_fwd_hd: head_dim %d has no kernel (16..128 step 16, 160, 192, 224, 256)", head_dim);
  return -1;
}
