// Fused multi-head attention, head dim 64, fp16 operands, fp32 online softmax (flash-style), forward and
// data-gradient.  Replaces QKVAttentionLegacy.forward (unet.py:337-354) and QKVAttention.forward
// (unet.py:370-389): weight = softmax((q*s)(k*s)^T) with s = 64^-1/4, a = weight @ v — without ever
// materialising the [T,T] weights (the reference writes B*H*T^2 fp16 + an fp32 copy per block).
//
// Layout: qkv is the NHWC output of the qkv 1x1 conv, i.e. token-major [n][t][3*H*64]; a head's q/k/v
// rows are contiguous 64-vectors, which is exactly the K-major operand layout mma.sync wants.
// Channel order: legacy = [head][q,k,v][64], new = [q,k,v][head][64].
//
// Attention is 0.5 % of the UNet FLOPs, so this kernel uses warp-level mma.sync
// (m16n8k16) rather than tcgen05: T in {64,256,1024}, tiles of 64 queries x 64 keys, 4 warps per CTA.
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
// attention_tc.cu: tcgen05 / TMEM / TMA forward for sequence lengths that are a multiple of 128
bool attn_fwd_tc_applicable(const void* qkv, int ld_qkv, const void* out, int ld_out, int t);
int attn_fwd_tc_launch(const void* qkv, int ld_qkv, void* out, int ld_out, float* lse, int n, int t, int t_valid,
                       int heads, int order, cudaStream_t stream);
namespace {

constexpr int kD = 64;
constexpr int kBQ = 64;
constexpr int kBKV = 64;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// A [64 rows][64 fp16] tile in shared memory, 128 B per row, 16-byte chunks XOR-swizzled by (row & 7).
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
// cooperative load of a 64x64 fp16 tile: rows `row0..row0+63` of a [.., ld] matrix starting at column col0
__device__ __forceinline__ void load_tile_async(__half* smem_tile, const __half* g, int ld, int tid) {
  const uint32_t base = smem_u32(smem_tile);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 128;  // 512 chunks
    const int row = idx >> 3, chunk = idx & 7;
    const uint32_t dst = tile_addr(base, row, chunk);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + static_cast<size_t>(row) * ld + chunk * 8)
                 : "memory");
  }
}

struct HeadPtrs {
  const __half* q;
  const __half* k;
  const __half* v;
};
__device__ __forceinline__ HeadPtrs head_ptrs(const __half* qkv_n, int head, int heads, int order) {
  HeadPtrs p;
  if (order == GD_QKV_LEGACY) {
    p.q = qkv_n + head * 3 * kD;
    p.k = p.q + kD;
    p.v = p.q + 2 * kD;
  } else {
    p.q = qkv_n + head * kD;
    p.k = p.q + heads * kD;
    p.v = p.q + 2 * heads * kD;
  }
  return p;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __half* __restrict__ qkv, int ld_qkv, __half* __restrict__ out, int ld_out,
                float* __restrict__ lse, int t, int t_valid, int heads, int order) {
  pdl_enter();
  __shared__ __align__(128) __half sQ[kBQ * kD];
  __shared__ __align__(128) __half sK[2][kBKV * kD];
  __shared__ __align__(128) __half sV[2][kBKV * kD];
  const int qb = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* qkv_n = qkv + static_cast<size_t>(n) * t * ld_qkv;
  const HeadPtrs hp = head_ptrs(qkv_n, head, heads, order);

  load_tile_async(sQ, hp.q + static_cast<size_t>(qb) * kBQ * ld_qkv, ld_qkv, tid);
  load_tile_async(sK[0], hp.k, ld_qkv, tid);
  load_tile_async(sV[0], hp.v, ld_qkv, tid);
  cp_async_commit();

  const int nkv = (t_valid + kBKV - 1) / kBKV;  // key blocks beyond the valid length are never touched
  // Q fragments (A operand) for the 4 k-steps, loaded once
  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[j][e] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sc = 0.125f * kLog2e;  // (64^-1/4)^2 = 1/8, folded with log2(e) for exp2f

  for (int kv = 0; kv < nkv; ++kv) {
    const int buf = kv & 1;
    if (kv + 1 < nkv) {
      load_tile_async(sK[buf ^ 1], hp.k + static_cast<size_t>(kv + 1) * kBKV * ld_qkv, ld_qkv, tid);
      load_tile_async(sV[buf ^ 1], hp.v + static_cast<size_t>(kv + 1) * kBKV * ld_qkv, ld_qkv, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kv == 0) {
      const uint32_t qbase = smem_u32(sQ);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int chunk = kk * 2 + (lane >> 4);
        ldmatrix_x4(qf[kk], tile_addr(qbase, row, chunk));
      }
    }
    // S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
    const uint32_t kbase = smem_u32(sK[buf]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key n-tiles
        uint32_t kf[4];
        const int row = jp * 16 + (lane & 7) + 8 * (lane >> 4);
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kf, tile_addr(kbase, row, chunk));
        mma16816(s[2 * jp], qf[kk], kf[0], kf[1]);
        mma16816(s[2 * jp + 1], qf[kk], kf[2], kf[3]);
      }
    }
    if ((kv + 1) * kBKV > t_valid) {  // padded keys (sequence padded to a multiple of 64): score -inf
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kv * kBKV + j * 8 + (lane & 3) * 2 + (e & 1) >= t_valid) s[j][e] = -INFINITY;
    }
    // online softmax over this 64-key block; rows r0 = lane/4 and r0+8
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
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
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= corr[0];
      o[j][1] *= corr[0];
      o[j][2] *= corr[1];
      o[j][3] *= corr[1];
    }
    // O += P V
    const uint32_t vbase = smem_u32(sV[buf]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
      uint32_t pf[4];
      pf[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-wide d tiles
        uint32_t vf[4];
        const int row = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int chunk = jp * 2 + (lane >> 4);
        ldmatrix_x4_trans(vf, tile_addr(vbase, row, chunk));
        mma16816(o[2 * jp], pf, vf[0], vf[1]);
        mma16816(o[2 * jp + 1], pf, vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  // finalise: divide by the row sums (reduce across the 4 lanes sharing a row)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int row0 = qb * kBQ + warp * 16 + (lane >> 2);
  __half* o_n = out + static_cast<size_t>(n) * t * ld_out + head * kD;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + (lane & 3) * 2;
    *reinterpret_cast<__half2*>(o_n + static_cast<size_t>(row0) * ld_out + col) =
        __floats2half2_rn(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<__half2*>(o_n + static_cast<size_t>(row0 + 8) * ld_out + col) =
        __floats2half2_rn(o[j][2] * inv1, o[j][3] * inv1);
  }
  if (lse != nullptr && (lane & 3) == 0) {
    // natural-log LSE of the scaled scores: max*scale + ln(sum)
    float* l = lse + (static_cast<size_t>(n) * heads + head) * t;
    l[row0] = m_run[0] * 0.125f + logf(l_run[0]);
    l[row0 + 8] = m_run[1] * 0.125f + logf(l_run[1]);
  }
}

// ---------------------------------------------------------------------------------------------
// backward
//   P = exp(S/8 - lse);  dV = P^T dO;  dP = dO V^T;  dS = P o (dP - delta) / 8 ... (scale folded below)
//   dQ = dS K;  dK = dS^T Q;   delta_t = sum_j dO[t,j] O[t,j]
// One CTA per (64-key block, head, sample): loops over query blocks, accumulates dK/dV in registers and
// writes them once; dQ contributions are accumulated with fp32 atomics into a workspace-free fp32 pass:
// to stay deterministic and atomic-free we instead run a second kernel with the roles swapped
// (one CTA per query block looping over key blocks) that produces dQ.
// ---------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __half* __restrict__ out, int ld_out, const __half* __restrict__ dout,
                                  int ld_dout, float* __restrict__ delta, int t, int heads) {
  pdl_enter();
  // one warp per (n, head, token): 64 channels -> 2 per lane
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int total = gridDim.y * heads * t;
  (void)total;
  const int n = blockIdx.y;
  if (gw >= heads * t) return;
  const int head = gw / t, tok = gw - head * t;
  const __half2 a = *reinterpret_cast<const __half2*>(out + (static_cast<size_t>(n) * t + tok) * ld_out + head * kD + lane * 2);
  const __half2 b = *reinterpret_cast<const __half2*>(dout + (static_cast<size_t>(n) * t + tok) * ld_dout + head * kD + lane * 2);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  float v = fa.x * fb.x + fa.y * fb.y;
  v = warp_sum(v);
  if (lane == 0) delta[(static_cast<size_t>(n) * heads + head) * t + tok] = v;
}

// dK, dV: CTA owns 64 keys; warp w owns keys [16w, 16w+16); loops over all query blocks.
//   S^T tile (keys x queries) = K Q^T  -> P^T;  dV += P^T dO;  dP^T = V dO^T;  dS^T = P^T o (dP^T - delta);  dK += dS^T Q
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const __half* __restrict__ qkv, int ld_qkv, const __half* __restrict__ dout, int ld_dout,
                    const float* __restrict__ lse, const float* __restrict__ delta, __half* __restrict__ dqkv,
                    int ld_dqkv, int t, int t_valid, int heads, int order) {
  pdl_enter();
  // 32 KiB of tiles: buffer 1 of the Q/dO ring first stages this CTA's K and V (read once into registers).
  __shared__ __align__(128) __half sQ[2][kBQ * kD];
  __shared__ __align__(128) __half sdO[2][kBQ * kD];
  __shared__ float sLse[2][kBQ];
  __shared__ float sDelta[2][kBQ];
  const int kb = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* qkv_n = qkv + static_cast<size_t>(n) * t * ld_qkv;
  const HeadPtrs hp = head_ptrs(qkv_n, head, heads, order);
  const __half* do_n = dout + static_cast<size_t>(n) * t * ld_dout + head * kD;
  const float* lse_h = lse + (static_cast<size_t>(n) * heads + head) * t;
  const float* delta_h = delta + (static_cast<size_t>(n) * heads + head) * t;

  load_tile_async(sQ[1], hp.k + static_cast<size_t>(kb) * kBKV * ld_qkv, ld_qkv, tid);
  load_tile_async(sdO[1], hp.v + static_cast<size_t>(kb) * kBKV * ld_qkv, ld_qkv, tid);
  cp_async_commit();
  load_tile_async(sQ[0], hp.q, ld_qkv, tid);
  load_tile_async(sdO[0], do_n, ld_dout, tid);
  cp_async_commit();
  if (tid < kBQ) {
    sLse[0][tid] = lse_h[tid];
    sDelta[0][tid] = delta_h[tid];
  }

  float dk[8][4], dv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[j][e] = dv[j][e] = 0.f;
  uint32_t kf[4][4], vf[4][4];  // K and V rows of this warp as A operands (16 keys x 64 d)
  cp_async_wait<1>();
  __syncthreads();
  {
    const uint32_t kbase = smem_u32(sQ[1]), vbase = smem_u32(sdO[1]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int chunk = kk * 2 + (lane >> 4);
      ldmatrix_x4(kf[kk], tile_addr(kbase, row, chunk));
      ldmatrix_x4(vf[kk], tile_addr(vbase, row, chunk));
    }
  }
  __syncthreads();  // K/V staging buffers may now be overwritten by the Q/dO prefetch
  const int nq = t / kBQ;
  for (int qb = 0; qb < nq; ++qb) {
    const int buf = qb & 1;
    if (qb + 1 < nq) {
      load_tile_async(sQ[buf ^ 1], hp.q + static_cast<size_t>(qb + 1) * kBQ * ld_qkv, ld_qkv, tid);
      load_tile_async(sdO[buf ^ 1], do_n + static_cast<size_t>(qb + 1) * kBQ * ld_dout, ld_dout, tid);
      cp_async_commit();
      if (tid < kBQ) {
        sLse[buf ^ 1][tid] = lse_h[(qb + 1) * kBQ + tid];
        sDelta[buf ^ 1][tid] = delta_h[(qb + 1) * kBQ + tid];
      }
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t qbase = smem_u32(sQ[buf]), dobase = smem_u32(sdO[buf]);
    // S^T = K Q^T (16 keys x 64 queries), dP^T = V dO^T
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) st[j][e] = dpt[j][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t bq[4], bdo[4];
        const int row = jp * 16 + (lane & 7) + 8 * (lane >> 4);
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bq, tile_addr(qbase, row, chunk));
        ldmatrix_x4(bdo, tile_addr(dobase, row, chunk));
        mma16816(st[2 * jp], kf[kk], bq[0], bq[1]);
        mma16816(st[2 * jp + 1], kf[kk], bq[2], bq[3]);
        mma16816(dpt[2 * jp], vf[kk], bdo[0], bdo[1]);
        mma16816(dpt[2 * jp + 1], vf[kk], bdo[2], bdo[3]);
      }
    }
    // P^T and dS^T (columns are queries: col = j*8 + (lane&3)*2 + {0,1})
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int q0 = j * 8 + (lane & 3) * 2;
      const float l0 = sLse[buf][q0], l1 = sLse[buf][q0 + 1];
      const float d0 = sDelta[buf][q0], d1 = sDelta[buf][q0 + 1];
      // rows of S^T are keys: (lane >> 2) and +8 within this warp's 16; padded keys have probability 0
      const int key0 = kb * kBKV + warp * 16 + (lane >> 2);
      const bool v0 = key0 < t_valid, v1 = key0 + 8 < t_valid;
      const float p0 = v0 ? __expf(st[j][0] * 0.125f - l0) : 0.f, p1 = v0 ? __expf(st[j][1] * 0.125f - l1) : 0.f;
      const float p2 = v1 ? __expf(st[j][2] * 0.125f - l0) : 0.f, p3 = v1 ? __expf(st[j][3] * 0.125f - l1) : 0.f;
      st[j][0] = p0; st[j][1] = p1; st[j][2] = p2; st[j][3] = p3;
      dpt[j][0] = p0 * (dpt[j][0] - d0) * 0.125f;
      dpt[j][1] = p1 * (dpt[j][1] - d1) * 0.125f;
      dpt[j][2] = p2 * (dpt[j][2] - d0) * 0.125f;
      dpt[j][3] = p3 * (dpt[j][3] - d1) * 0.125f;
    }
    // dV += P^T dO ; dK += dS^T Q     (k dimension = queries; B operands need [k=query][n=d] -> ldmatrix.trans)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pf[4], dsf[4];
      pf[0] = pack_half2(st[2 * kk][0], st[2 * kk][1]);
      pf[1] = pack_half2(st[2 * kk][2], st[2 * kk][3]);
      pf[2] = pack_half2(st[2 * kk + 1][0], st[2 * kk + 1][1]);
      pf[3] = pack_half2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
      dsf[0] = pack_half2(dpt[2 * kk][0], dpt[2 * kk][1]);
      dsf[1] = pack_half2(dpt[2 * kk][2], dpt[2 * kk][3]);
      dsf[2] = pack_half2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]);
      dsf[3] = pack_half2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t bdo[4], bq[4];
        const int row = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int chunk = jp * 2 + (lane >> 4);
        ldmatrix_x4_trans(bdo, tile_addr(dobase, row, chunk));
        ldmatrix_x4_trans(bq, tile_addr(qbase, row, chunk));
        mma16816(dv[2 * jp], pf, bdo[0], bdo[1]);
        mma16816(dv[2 * jp + 1], pf, bdo[2], bdo[3]);
        mma16816(dk[2 * jp], dsf, bq[0], bq[1]);
        mma16816(dk[2 * jp + 1], dsf, bq[2], bq[3]);
      }
    }
    __syncthreads();
  }
  // write dK, dV into the dqkv tensor at the k / v channel positions of this head
  __half* dq_n = dqkv + static_cast<size_t>(n) * t * ld_dqkv;
  const HeadPtrs dp = head_ptrs(dq_n, head, heads, order);
  __half* dkp = const_cast<__half*>(dp.k);
  __half* dvp = const_cast<__half*>(dp.v);
  const int row0 = kb * kBKV + warp * 16 + (lane >> 2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + (lane & 3) * 2;
    *reinterpret_cast<__half2*>(dkp + static_cast<size_t>(row0) * ld_dqkv + col) = __floats2half2_rn(dk[j][0], dk[j][1]);
    *reinterpret_cast<__half2*>(dkp + static_cast<size_t>(row0 + 8) * ld_dqkv + col) = __floats2half2_rn(dk[j][2], dk[j][3]);
    *reinterpret_cast<__half2*>(dvp + static_cast<size_t>(row0) * ld_dqkv + col) = __floats2half2_rn(dv[j][0], dv[j][1]);
    *reinterpret_cast<__half2*>(dvp + static_cast<size_t>(row0 + 8) * ld_dqkv + col) = __floats2half2_rn(dv[j][2], dv[j][3]);
  }
}

// dQ: CTA owns 64 queries; warp w owns 16 of them; loops over key blocks.
//   S = Q K^T -> P;  dP = dO V^T;  dS = P o (dP - delta)/8;  dQ += dS K
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const __half* __restrict__ qkv, int ld_qkv, const __half* __restrict__ dout, int ld_dout,
                   const float* __restrict__ lse, const float* __restrict__ delta, __half* __restrict__ dqkv,
                   int ld_dqkv, int t, int t_valid, int heads, int order) {
  pdl_enter();
  __shared__ __align__(128) __half sQ[kBQ * kD];
  __shared__ __align__(128) __half sdO[kBQ * kD];
  __shared__ __align__(128) __half sK[2][kBKV * kD];
  __shared__ __align__(128) __half sV[2][kBKV * kD];
  const int qb = blockIdx.x, head = blockIdx.y, n = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __half* qkv_n = qkv + static_cast<size_t>(n) * t * ld_qkv;
  const HeadPtrs hp = head_ptrs(qkv_n, head, heads, order);
  const __half* do_n = dout + static_cast<size_t>(n) * t * ld_dout + head * kD;
  const float* lse_h = lse + (static_cast<size_t>(n) * heads + head) * t;
  const float* delta_h = delta + (static_cast<size_t>(n) * heads + head) * t;

  load_tile_async(sQ, hp.q + static_cast<size_t>(qb) * kBQ * ld_qkv, ld_qkv, tid);
  load_tile_async(sdO, do_n + static_cast<size_t>(qb) * kBQ * ld_dout, ld_dout, tid);
  load_tile_async(sK[0], hp.k, ld_qkv, tid);
  load_tile_async(sV[0], hp.v, ld_qkv, tid);
  cp_async_commit();

  const int r0 = qb * kBQ + warp * 16 + (lane >> 2);
  const float lse0 = lse_h[r0], lse1 = lse_h[r0 + 8];
  const float del0 = delta_h[r0], del1 = delta_h[r0 + 8];
  uint32_t qf[4][4], dof[4][4];
  float dq[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[j][e] = 0.f;
  const int nkv = (t_valid + kBKV - 1) / kBKV;
  for (int kv = 0; kv < nkv; ++kv) {
    const int buf = kv & 1;
    if (kv + 1 < nkv) {
      load_tile_async(sK[buf ^ 1], hp.k + static_cast<size_t>(kv + 1) * kBKV * ld_qkv, ld_qkv, tid);
      load_tile_async(sV[buf ^ 1], hp.v + static_cast<size_t>(kv + 1) * kBKV * ld_qkv, ld_qkv, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (kv == 0) {
      const uint32_t qbase = smem_u32(sQ), dobase = smem_u32(sdO);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int row = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int chunk = kk * 2 + (lane >> 4);
        ldmatrix_x4(qf[kk], tile_addr(qbase, row, chunk));
        ldmatrix_x4(dof[kk], tile_addr(dobase, row, chunk));
      }
    }
    const uint32_t kbase = smem_u32(sK[buf]), vbase = smem_u32(sV[buf]);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = dp[j][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t bk[4], bv[4];
        const int row = jp * 16 + (lane & 7) + 8 * (lane >> 4);
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(bk, tile_addr(kbase, row, chunk));
        ldmatrix_x4(bv, tile_addr(vbase, row, chunk));
        mma16816(s[2 * jp], qf[kk], bk[0], bk[1]);
        mma16816(s[2 * jp + 1], qf[kk], bk[2], bk[3]);
        mma16816(dp[2 * jp], dof[kk], bv[0], bv[1]);
        mma16816(dp[2 * jp + 1], dof[kk], bv[2], bv[3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int key = kv * kBKV + j * 8 + (lane & 3) * 2;  // columns are keys; padded keys have probability 0
      const bool v0 = key < t_valid, v1 = key + 1 < t_valid;
      const float p0 = v0 ? __expf(s[j][0] * 0.125f - lse0) : 0.f, p1 = v1 ? __expf(s[j][1] * 0.125f - lse0) : 0.f;
      const float p2 = v0 ? __expf(s[j][2] * 0.125f - lse1) : 0.f, p3 = v1 ? __expf(s[j][3] * 0.125f - lse1) : 0.f;
      dp[j][0] = p0 * (dp[j][0] - del0) * 0.125f;
      dp[j][1] = p1 * (dp[j][1] - del0) * 0.125f;
      dp[j][2] = p2 * (dp[j][2] - del1) * 0.125f;
      dp[j][3] = p3 * (dp[j][3] - del1) * 0.125f;
    }
    // dQ += dS K   (k dimension = keys; B operand [k=key][n=d] -> ldmatrix.trans on K)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t dsf[4];
      dsf[0] = pack_half2(dp[2 * kk][0], dp[2 * kk][1]);
      dsf[1] = pack_half2(dp[2 * kk][2], dp[2 * kk][3]);
      dsf[2] = pack_half2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      dsf[3] = pack_half2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t bk[4];
        const int row = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int chunk = jp * 2 + (lane >> 4);
        ldmatrix_x4_trans(bk, tile_addr(kbase, row, chunk));
        mma16816(dq[2 * jp], dsf, bk[0], bk[1]);
        mma16816(dq[2 * jp + 1], dsf, bk[2], bk[3]);
      }
    }
    __syncthreads();
  }
  __half* dq_n = dqkv + static_cast<size_t>(n) * t * ld_dqkv;
  const HeadPtrs dpz = head_ptrs(dq_n, head, heads, order);
  __half* dqp = const_cast<__half*>(dpz.q);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + (lane & 3) * 2;
    *reinterpret_cast<__half2*>(dqp + static_cast<size_t>(r0) * ld_dqkv + col) = __floats2half2_rn(dq[j][0], dq[j][1]);
    *reinterpret_cast<__half2*>(dqp + static_cast<size_t>(r0 + 8) * ld_dqkv + col) = __floats2half2_rn(dq[j][2], dq[j][3]);
  }
}

int check_attn(const char* who, int ld_qkv, int n, int t, int heads, int order) {
  GD_REQUIRE(n > 0 && heads > 0, "%s: bad n/heads", who);
  GD_REQUIRE(t > 0 && t % 64 == 0, "%s: sequence length must be a multiple of 64, got %d", who, t);
  GD_REQUIRE(ld_qkv >= 3 * heads * kD && ld_qkv % 8 == 0, "%s: bad ld_qkv %d", who, ld_qkv);
  GD_REQUIRE(order == GD_QKV_LEGACY || order == GD_QKV_NEW, "%s: bad qkv order %d", who, order);
  return 0;
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int gd_attention_fwd_masked(const void* qkv, int32_t ld_qkv, void* out, int32_t ld_out, float* lse, int32_t n,
                                       int32_t t, int32_t t_valid, int32_t heads, int32_t order, void* stream) {
  GD_REQUIRE(qkv && out, "gd_attention_fwd: null pointer");
  if (int rc = check_attn("gd_attention_fwd", ld_qkv, n, t, heads, order)) return rc;
  GD_REQUIRE(t_valid > 0 && t_valid <= t, "gd_attention_fwd: valid length %d outside (0, %d]", t_valid, t);
  GD_REQUIRE(ld_out >= heads * kD && ld_out % 2 == 0, "gd_attention_fwd: bad ld_out %d", ld_out);
  if (attn_fwd_tc_applicable(qkv, ld_qkv, out, ld_out, t))
    return attn_fwd_tc_launch(qkv, ld_qkv, out, ld_out, lse, n, t, t_valid, heads, order,
                              reinterpret_cast<cudaStream_t>(stream));
  dim3 grid(t / kBQ, heads, n);  // 64-token sequences (8x8 resolution): warp-level mma.sync kernel
  GD_CHECK_CUDA(launch_pdl(attn_fwd_kernel, grid, dim3(128), 0, reinterpret_cast<cudaStream_t>(stream),
                           reinterpret_cast<const __half*>(qkv), ld_qkv, reinterpret_cast<__half*>(out), ld_out, lse, t,
                           t_valid, heads, order));
  count_launch(1);
  return 0;
}

extern "C" int gd_attention_fwd(const void* qkv, int32_t ld_qkv, void* out, int32_t ld_out, float* lse, int32_t n,
                                int32_t t, int32_t heads, int32_t order, void* stream) {
  return gd_attention_fwd_masked(qkv, ld_qkv, out, ld_out, lse, n, t, t, heads, order, stream);
}

extern "C" int gd_attention_bwd_masked(const void* qkv, int32_t ld_qkv, const void* out, int32_t ld_out, const void* dout,
                                       int32_t ld_dout, const float* lse, float* delta_ws, void* dqkv, int32_t ld_dqkv,
                                       int32_t n, int32_t t, int32_t t_valid, int32_t heads, int32_t order, void* stream) {
  GD_REQUIRE(qkv && out && dout && lse && delta_ws && dqkv, "gd_attention_bwd: null pointer");
  if (int rc = check_attn("gd_attention_bwd", ld_qkv, n, t, heads, order)) return rc;
  GD_REQUIRE(t_valid > 0 && t_valid <= t, "gd_attention_bwd: valid length %d outside (0, %d]", t_valid, t);
  GD_REQUIRE(ld_dqkv >= 3 * heads * kD && ld_dqkv % 8 == 0 && ld_dout % 8 == 0 && ld_out % 2 == 0,
             "gd_attention_bwd: bad strides");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int warps = heads * t;
  dim3 dgrid((warps * 32 + 255) / 256, n);
  GD_CHECK_CUDA(launch_pdl(attn_delta_kernel, dgrid, dim3(256), 0, st, reinterpret_cast<const __half*>(out), ld_out,
                           reinterpret_cast<const __half*>(dout), ld_dout, delta_ws, t, heads));
  dim3 grid(t / 64, heads, n);
  GD_CHECK_CUDA(launch_pdl(attn_bwd_dkv_kernel, grid, dim3(128), 0, st, reinterpret_cast<const __half*>(qkv), ld_qkv,
                           reinterpret_cast<const __half*>(dout), ld_dout, lse, delta_ws, reinterpret_cast<__half*>(dqkv),
                           ld_dqkv, t, t_valid, heads, order));
  GD_CHECK_CUDA(launch_pdl(attn_bwd_dq_kernel, grid, dim3(128), 0, st, reinterpret_cast<const __half*>(qkv), ld_qkv,
                           reinterpret_cast<const __half*>(dout), ld_dout, lse, delta_ws, reinterpret_cast<__half*>(dqkv),
                           ld_dqkv, t, t_valid, heads, order));
  count_launch(3);
  return 0;
}

extern "C" int gd_attention_bwd(const void* qkv, int32_t ld_qkv, const void* out, int32_t ld_out, const void* dout,
                                int32_t ld_dout, const float* lse, float* delta_ws, void* dqkv, int32_t ld_dqkv,
                                int32_t n, int32_t t, int32_t heads, int32_t order, void* stream) {
  return gd_attention_bwd_masked(qkv, ld_qkv, out, ld_out, dout, ld_dout, lse, delta_ws, dqkv, ld_dqkv, n, t, t, heads,
                                 order, stream);
}
