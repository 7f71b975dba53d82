// Attention-pool head of the EncoderUNet classifier (AttentionPool2d, unet.py:22-51, built at :833-841),
// forward and data-gradient, in fp32 like the reference (the pool head is never converted to fp16,
// unet.py:858-863).  Only token 0 of the output is used (unet.py:51), so only ONE query per head is
// evaluated: k, v for all T = hw+1 tokens, q for token 0.
//
// Workspace (floats), T = hw + 1:
//   x    [n][T][c]     tokens (mean token first) + positional embedding
//   kv   [n][T][2c]    k | v   ("new" qkv order: [q,k,v][head][64], unet.py:380-388)
//   q0   [n][c]
//   p    [n][heads][T] softmax probabilities of query 0
//   a    [n][c]        attention output of token 0
//   da   [n][c]
//   dqkv [n][T][3c]
//   dx   [n][T][c]
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

struct PoolWs {
  float *x, *kv, *q0, *p, *a, *da, *dqkv, *dx;
};
PoolWs carve(float* ws, int n, int T, int c, int heads) {
  PoolWs w;
  size_t o = 0;
  w.x = ws + o;    o += static_cast<size_t>(n) * T * c;
  w.kv = ws + o;   o += static_cast<size_t>(n) * T * 2 * c;
  w.q0 = ws + o;   o += static_cast<size_t>(n) * c;
  w.p = ws + o;    o += static_cast<size_t>(n) * heads * T;
  w.a = ws + o;    o += static_cast<size_t>(n) * c;
  w.da = ws + o;   o += static_cast<size_t>(n) * c;
  w.dqkv = ws + o; o += static_cast<size_t>(n) * T * 3 * c;
  w.dx = ws + o;
  return w;
}

// x[n][0][ch] = mean_p h[n][p][ch] + pos[ch][0];  x[n][1+p][ch] = h[n][p][ch] + pos[ch][1+p]
__global__ void pool_tokens_kernel(const __half* __restrict__ h, int ld, const float* __restrict__ pos,
                                   float* __restrict__ x, int hw, int c) {
  const int n = blockIdx.y;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const int T = hw + 1;
  const __half* hn = h + static_cast<size_t>(n) * hw * ld + ch;
  float* xn = x + static_cast<size_t>(n) * T * c + ch;
  float s = 0.f;
  for (int p = 0; p < hw; ++p) {
    const float v = __half2float(hn[static_cast<size_t>(p) * ld]);
    s += v;
    xn[static_cast<size_t>(1 + p) * c] = v + pos[static_cast<size_t>(ch) * T + 1 + p];
  }
  xn[0] = s / static_cast<float>(hw) + pos[static_cast<size_t>(ch) * T];
}

// one CTA (64 threads) per (head, n): scores, softmax, weighted sum of v
__global__ void pool_attn_fwd_kernel(const float* __restrict__ q0, const float* __restrict__ kv, float* __restrict__ p,
                                     float* __restrict__ a, int T, int c, int heads) {
  extern __shared__ float s_p[];  // [T]
  __shared__ float s_red[2];
  const int head = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const float* q = q0 + static_cast<size_t>(n) * c + head * 64;
  const float* kvn = kv + static_cast<size_t>(n) * T * 2 * c;
  for (int s = tid; s < T; s += blockDim.x) {
    const float* k = kvn + static_cast<size_t>(s) * 2 * c + head * 64;
    float d = 0.f;
#pragma unroll 8
    for (int j = 0; j < 64; ++j) d = fmaf(q[j], k[j], d);
    s_p[s] = d * 0.125f;
  }
  __syncthreads();
  if (tid == 0) {
    float mx = -INFINITY;
    for (int s = 0; s < T; ++s) mx = fmaxf(mx, s_p[s]);
    float sum = 0.f;
    for (int s = 0; s < T; ++s) sum += expf(s_p[s] - mx);
    s_red[0] = mx;
    s_red[1] = 1.0f / sum;
  }
  __syncthreads();
  for (int s = tid; s < T; s += blockDim.x) {
    const float pr = expf(s_p[s] - s_red[0]) * s_red[1];
    s_p[s] = pr;
    p[(static_cast<size_t>(n) * heads + head) * T + s] = pr;
  }
  __syncthreads();
  if (tid < 64) {
    float acc = 0.f;
    for (int s = 0; s < T; ++s) acc = fmaf(s_p[s], kvn[static_cast<size_t>(s) * 2 * c + c + head * 64 + tid], acc);
    a[static_cast<size_t>(n) * c + head * 64 + tid] = acc;
  }
}

// dqkv[n][s] = [dq (token 0 only) | dk | dv]
__global__ void pool_attn_bwd_kernel(const float* __restrict__ q0, const float* __restrict__ kv,
                                     const float* __restrict__ p, const float* __restrict__ da,
                                     float* __restrict__ dqkv, int T, int c, int heads) {
  extern __shared__ float s_ds[];  // [T]
  __shared__ float s_dot;
  const int head = blockIdx.x, n = blockIdx.y, tid = threadIdx.x;
  const float* q = q0 + static_cast<size_t>(n) * c + head * 64;
  const float* kvn = kv + static_cast<size_t>(n) * T * 2 * c;
  const float* pn = p + (static_cast<size_t>(n) * heads + head) * T;
  const float* dan = da + static_cast<size_t>(n) * c + head * 64;
  float* dn = dqkv + static_cast<size_t>(n) * T * 3 * c;
  // dp_s = da . v_s
  for (int s = tid; s < T; s += blockDim.x) {
    const float* v = kvn + static_cast<size_t>(s) * 2 * c + c + head * 64;
    float d = 0.f;
#pragma unroll 8
    for (int j = 0; j < 64; ++j) d = fmaf(dan[j], v[j], d);
    s_ds[s] = d;
  }
  __syncthreads();
  if (tid == 0) {
    float dot = 0.f;
    for (int s = 0; s < T; ++s) dot = fmaf(pn[s], s_ds[s], dot);
    s_dot = dot;
  }
  __syncthreads();
  for (int s = tid; s < T; s += blockDim.x) s_ds[s] = pn[s] * (s_ds[s] - s_dot) * 0.125f;
  __syncthreads();
  if (tid < 64) {
    const int j = tid;
    float dq = 0.f;
    for (int s = 0; s < T; ++s) {
      const float ds = s_ds[s];
      dq = fmaf(ds, kvn[static_cast<size_t>(s) * 2 * c + head * 64 + j], dq);
      float* row = dn + static_cast<size_t>(s) * 3 * c;
      row[c + head * 64 + j] = ds * q[j];            // dk_s
      row[2 * c + head * 64 + j] = pn[s] * dan[j];   // dv_s
      if (s > 0) row[head * 64 + j] = 0.f;           // dq only exists for token 0
    }
    dn[head * 64 + j] = dq;
  }
}

// dh[n][p][ch] = scale * (dx[n][1+p][ch] + dx[n][0][ch] / hw)
__global__ void pool_tokens_bwd_kernel(const float* __restrict__ dx, __half* __restrict__ dh, int ld, int hw, int c,
                                       float scale) {
  const int n = blockIdx.y;
  const int T = hw + 1;
  const size_t total = static_cast<size_t>(hw) * c;
  const float* dxn = dx + static_cast<size_t>(n) * T * c;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i / c), ch = static_cast<int>(i - static_cast<size_t>(p) * c);
    const float v = dxn[static_cast<size_t>(1 + p) * c + ch] + dxn[ch] / static_cast<float>(hw);
    dh[(static_cast<size_t>(n) * hw + p) * ld + ch] = __float2half_rn(v * scale);
  }
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int64_t gd_attnpool_ws_floats(int32_t n, int32_t tokens, int32_t c) {
  const int64_t T = tokens;
  return static_cast<int64_t>(n) * T * c * 7 + static_cast<int64_t>(n) * c * 3 + static_cast<int64_t>(n) * (c / 64) * T + 64;
}

extern "C" int gd_attnpool_fwd(const void* h, int32_t ld, const float* pos_emb, const float* w_qkv, const float* b_qkv,
                               const float* w_c, const float* b_c, float* logits, float* ws, int32_t n, int32_t hw,
                               int32_t c, int32_t heads, int32_t n_out, void* stream) {
  GD_REQUIRE(h && pos_emb && w_qkv && b_qkv && w_c && b_c && logits && ws, "gd_attnpool_fwd: null pointer");
  GD_REQUIRE(c == heads * 64, "gd_attnpool_fwd: head dim must be 64 (c=%d heads=%d)", c, heads);
  const int T = hw + 1;
  const PoolWs w = carve(ws, n, T, c, heads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  pool_tokens_kernel<<<dim3((c + 127) / 128, n), 128, 0, st>>>(reinterpret_cast<const __half*>(h), ld, pos_emb, w.x, hw, c);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  // k | v for every token: rows c..3c of w_qkv
  int rc = gd_linear_f32(w.x, c, w_qkv + static_cast<size_t>(c) * c, b_qkv + c, nullptr, 0, w.kv, 2 * c, n * T, c, 2 * c,
                         0, 0, stream);
  if (rc) return rc;
  // q for token 0 only (row stride T*c picks token 0 of each sample)
  rc = gd_linear_f32(w.x, T * c, w_qkv, b_qkv, nullptr, 0, w.q0, c, n, c, c, 0, 0, stream);
  if (rc) return rc;
  pool_attn_fwd_kernel<<<dim3(heads, n), 64, T * sizeof(float), st>>>(w.q0, w.kv, w.p, w.a, T, c, heads);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return gd_linear_f32(w.a, c, w_c, b_c, nullptr, 0, logits, n_out, n, c, n_out, 0, 0, stream);
}

extern "C" int gd_attnpool_bwd(const float* dlogits, const float* w_qkv_t, const float* w_c_t, float* ws, void* dh,
                               int32_t ld_dh, int32_t n, int32_t hw, int32_t c, int32_t heads, int32_t n_out,
                               float out_scale, void* stream) {
  GD_REQUIRE(dlogits && w_qkv_t && w_c_t && ws && dh, "gd_attnpool_bwd: null pointer");
  GD_REQUIRE(c == heads * 64, "gd_attnpool_bwd: head dim must be 64 (c=%d heads=%d)", c, heads);
  const int T = hw + 1;
  const PoolWs w = carve(ws, n, T, c, heads);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // da = dlogits W_c  (w_c_t is [c][n_out])
  int rc = gd_linear_f32(dlogits, n_out, w_c_t, nullptr, nullptr, 0, w.da, c, n, n_out, c, 0, 0, stream);
  if (rc) return rc;
  pool_attn_bwd_kernel<<<dim3(heads, n), 64, T * sizeof(float), st>>>(w.q0, w.kv, w.p, w.da, w.dqkv, T, c, heads);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  // dx = dqkv W_qkv  (w_qkv_t is [c][3c])
  rc = gd_linear_f32(w.dqkv, 3 * c, w_qkv_t, nullptr, nullptr, 0, w.dx, c, n * T, 3 * c, c, 0, 0, stream);
  if (rc) return rc;
  const size_t total = static_cast<size_t>(hw) * c;
  int gx = static_cast<int>((total + 255) / 256);
  if (gx > 1024) gx = 1024;
  pool_tokens_bwd_kernel<<<dim3(gx, n), 256, 0, st>>>(w.dx, reinterpret_cast<__half*>(dh), ld_dh, hw, c,
                                                      out_scale == 0.f ? 1.f : out_scale);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
