// Small fp32 / bandwidth-bound kernels of the sampling step: the fused posterior + noise update, the
// timestep-embedding MLP pieces, the tiny-C_in first conv, uint8 packing and layout helpers.
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

// ---------------------------------------------------------------------------------------------
// Posterior / noise update.  Mirrors the reference op by op in fp32 (each product and sum rounded
// separately, like the chain of ATen pointwise ops in gaussian_diffusion.py:262-326,356-393,430-438,
// 575-593), so only expf/sqrtf implementations separate it from the PyTorch result.
// ---------------------------------------------------------------------------------------------
struct PostArgs {
  const float* x;
  const float* model_out;
  const float* grad;
  const float* noise;
  float* sample;
  float* pred_xstart;
  float* mean_out;
  float* var_out;
  float* logvar_out;
  const float* coef;
  const int64_t* t;
  int n, c, hw;
  int var_type, mean_type, clip, ddim;
  float eta;
  int num_timesteps;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

// One element of the update.  `b` = sample index (selects the coefficient row), `i` = flat index into x-shaped tensors,
// `m_out` / `v_out` = the model's mean / variance channels at this element.  Returns through `r`.
struct PostElem {
  float mean, var, logvar, x0, out;
};
__device__ __forceinline__ void posterior_elem(const PostArgs& p, const float* __restrict__ co, float x, float m_out,
                                               float v_out, float g, float z, PostElem& r) {
  const float sr = co[GD_COEF_SQRT_RECIP_ACP], srm1 = co[GD_COEF_SQRT_RECIPM1_ACP];
  const float c1 = co[GD_COEF_POST_MEAN1], c2 = co[GD_COEF_POST_MEAN2];
  const float max_log = co[GD_COEF_LOG_BETA], min_log = co[GD_COEF_POST_LOGVAR];
  const float acp = co[GD_COEF_ACP], acp_prev = co[GD_COEF_ACP_PREV];
  const float nonzero = co[GD_COEF_NONZERO];
  const float acp_next = co[GD_COEF_ACP_NEXT];
  float var, logvar;
  if (p.var_type == GD_VAR_LEARNED_RANGE) {
    const float frac = mul(add(v_out, 1.0f), 0.5f);  // (v + 1) / 2
    logvar = add(mul(frac, max_log), mul(sub(1.0f, frac), min_log));
    var = expf(logvar);
  } else if (p.var_type == GD_VAR_LEARNED) {
    logvar = v_out;
    var = expf(logvar);
  } else {
    var = co[GD_COEF_FIXED_VAR];
    logvar = co[GD_COEF_FIXED_LOGVAR];
  }
  float x0;
  if (p.mean_type == GD_MEAN_EPSILON) {
    x0 = sub(mul(sr, x), mul(srm1, m_out));
  } else {
    x0 = m_out;
  }
  if (p.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  float mean = add(mul(c1, x0), mul(c2, x));
  r.mean = mean;
  r.var = var;
  r.logvar = logvar;
  r.out = 0.f;
  if (p.ddim == GD_DDIM_REVERSE) {  // ddim_reverse_sample (gaussian_diffusion.py:596-632): deterministic, x_t -> x_{t+1}
    const float e = __fdiv_rn(sub(mul(sr, x), x0), srm1);
    r.out = add(mul(x0, sqrtf(acp_next)), mul(sqrtf(sub(1.0f, acp_next)), e));
  } else if (p.noise != nullptr) {
    if (!p.ddim) {
      if (p.grad != nullptr) mean = add(mean, mul(var, g));
      r.out = add(mean, mul(mul(nonzero, expf(mul(0.5f, logvar))), z));
    } else {
      if (p.grad != nullptr) {
        float e = __fdiv_rn(sub(mul(sr, x), x0), srm1);
        e = sub(e, mul(sqrtf(sub(1.0f, acp)), g));
        x0 = sub(mul(sr, x), mul(srm1, e));
      }
      const float e2 = __fdiv_rn(sub(mul(sr, x), x0), srm1);
      const float sigma = mul(mul(p.eta, sqrtf(__fdiv_rn(sub(1.0f, acp_prev), sub(1.0f, acp)))),
                              sqrtf(sub(1.0f, __fdiv_rn(acp, acp_prev))));
      const float mean_pred =
          add(mul(x0, sqrtf(acp_prev)), mul(sqrtf(sub(sub(1.0f, acp_prev), mul(sigma, sigma))), e2));
      r.out = add(mean_pred, mul(mul(nonzero, sigma), z));
    }
  }
  r.x0 = x0;
}

// kVec = 4: every tensor is read and written as float4 (host guarantees c*hw % 4 == 0 and 16-byte aligned pointers, so a
// vector never straddles two samples or the eps / variance halves of the model output); kVec = 1: any shape.
template <int kVec>
__global__ void posterior_kernel(const PostArgs p) {
  pdl_enter();
  const size_t chw = static_cast<size_t>(p.c) * p.hw;
  const size_t total = static_cast<size_t>(p.n) * chw / kVec;
  const int out_c = (p.var_type == GD_VAR_FIXED) ? p.c : 2 * p.c;
  const bool learned = p.var_type != GD_VAR_FIXED;
  const bool writes_sample = p.ddim == GD_DDIM_REVERSE || p.noise != nullptr;
  for (size_t iv = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; iv < total;
       iv += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t i = iv * kVec;
    const size_t b = i / chw, r = i - b * chw;
    // per-sample timestep index into the coefficient table (_extract_into_tensor, gaussian_diffusion.py:904-917)
    // An index outside [0, num_timesteps) raises IndexError in the reference; a kernel cannot raise, so it poisons
    // this sample's outputs with NaN instead of reading past the table (the host checks user-supplied t eagerly).
    const long long tb = p.t[b];
    const bool t_ok = tb >= 0 && tb < p.num_timesteps;
    const float* co = p.coef + static_cast<size_t>(t_ok ? tb : 0) * GD_COEF_STRIDE;
    const size_t mo = b * static_cast<size_t>(out_c) * p.hw + r;
    float x[kVec], m[kVec], v[kVec], g[kVec], z[kVec];
    if (kVec == 4) {
      *reinterpret_cast<float4*>(x) = __ldcs(reinterpret_cast<const float4*>(p.x + i));
      *reinterpret_cast<float4*>(m) = __ldcs(reinterpret_cast<const float4*>(p.model_out + mo));
      if (learned) *reinterpret_cast<float4*>(v) = __ldcs(reinterpret_cast<const float4*>(p.model_out + mo + chw));
      if (p.grad != nullptr) *reinterpret_cast<float4*>(g) = __ldcs(reinterpret_cast<const float4*>(p.grad + i));
      if (p.noise != nullptr) *reinterpret_cast<float4*>(z) = __ldcs(reinterpret_cast<const float4*>(p.noise + i));
    } else {
      x[0] = p.x[i];
      m[0] = p.model_out[mo];
      if (learned) v[0] = p.model_out[mo + chw];
      if (p.grad != nullptr) g[0] = p.grad[i];
      if (p.noise != nullptr) z[0] = p.noise[i];
    }
    float o_mean[kVec], o_var[kVec], o_logvar[kVec], o_x0[kVec], o_out[kVec];
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
      PostElem e;
      posterior_elem(p, co, x[j], m[j], learned ? v[j] : 0.f, p.grad != nullptr ? g[j] : 0.f,
                     p.noise != nullptr ? z[j] : 0.f, e);
      o_mean[j] = e.mean;
      o_var[j] = e.var;
      o_logvar[j] = e.logvar;
      o_x0[j] = e.x0;
      o_out[j] = e.out;
      if (!t_ok) o_mean[j] = o_var[j] = o_logvar[j] = o_x0[j] = o_out[j] = __int_as_float(0x7fc00000);
    }
    if (kVec == 4) {
      if (p.mean_out != nullptr) *reinterpret_cast<float4*>(p.mean_out + i) = *reinterpret_cast<const float4*>(o_mean);
      if (p.var_out != nullptr) *reinterpret_cast<float4*>(p.var_out + i) = *reinterpret_cast<const float4*>(o_var);
      if (p.logvar_out != nullptr) *reinterpret_cast<float4*>(p.logvar_out + i) = *reinterpret_cast<const float4*>(o_logvar);
      if (writes_sample) *reinterpret_cast<float4*>(p.sample + i) = *reinterpret_cast<const float4*>(o_out);
      if (p.pred_xstart != nullptr) *reinterpret_cast<float4*>(p.pred_xstart + i) = *reinterpret_cast<const float4*>(o_x0);
    } else {
      if (p.mean_out != nullptr) p.mean_out[i] = o_mean[0];
      if (p.var_out != nullptr) p.var_out[i] = o_var[0];
      if (p.logvar_out != nullptr) p.logvar_out[i] = o_logvar[0];
      if (writes_sample) p.sample[i] = o_out[0];
      if (p.pred_xstart != nullptr) p.pred_xstart[i] = o_x0[0];
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int b = i / half, k = i - b * half;
  // freqs = exp(-ln(10000) * k / half) computed in fp32 like nn.py:113-116
  const float freq = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, static_cast<float>(k)), static_cast<float>(half)));
  const float arg = __fmul_rn(t[b], freq);
  out[static_cast<size_t>(b) * dim + k] = cosf(arg);
  out[static_cast<size_t>(b) * dim + half + k] = sinf(arg);
  if ((dim & 1) && k == 0) out[static_cast<size_t>(b) * dim + dim - 1] = 0.f;
}

// y[m][n] = act_out( sum_k act_in(x[m][k]) W[n][k] + b[n] + add[m][n] ); one warp per output column,
// rows processed in blocks of 8 so each weight row is streamed once per row block (weights dominate).
template <bool kSiluIn>
__global__ void linear_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w,
                                  const float* __restrict__ b, const float* __restrict__ addp, int ld_add,
                                  float* __restrict__ y, int ldy, int m, int k, int n, int silu_out) {
  pdl_enter();
  const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (col >= n) return;
  const float* wr = w + static_cast<size_t>(col) * k;
  for (int m0 = 0; m0 < m; m0 += 8) {
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    for (int kk = lane; kk < k; kk += 32) {
      const float wv = __ldg(wr + kk);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (m0 + r < m) {
          float xv = x[static_cast<size_t>(m0 + r) * ldx + kk];
          if (kSiluIn) xv = xv / (1.0f + expf(-xv));
          acc[r] = fmaf(xv, wv, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (m0 + r < m) {
          float v = acc[r] + (b ? b[col] : 0.f);
          if (addp) v += addp[static_cast<size_t>(m0 + r) * ld_add + col];
          if (silu_out) v = v / (1.0f + expf(-v));
          y[static_cast<size_t>(m0 + r) * ldy + col] = v;
        }
      }
    }
  }
}

// Tiled fp32 GEMM for many rows (the classifier pool head: M = batch * 65 tokens): 64x64 output tile per CTA,
// K stepped by 16 through shared memory, 4x4 register tile per thread.  Same contract as linear_f32_kernel.
__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ b,
                const float* __restrict__ addp, int ld_add, float* __restrict__ y, int ldy, int m, int k, int n,
                int silu_in, int silu_out) {
  pdl_enter();
  __shared__ float sx[16][64 + 4];
  __shared__ float sw[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // each thread loads 4 consecutive k of one row of each operand
  for (int k0 = 0; k0 < k; k0 += 16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kk = k0 + lk + q;
      float xv = 0.f, wv = 0.f;
      if (kk < k) {
        if (m0 + lr < m) {
          xv = x[static_cast<size_t>(m0 + lr) * ldx + kk];
          if (silu_in) xv = xv / (1.0f + expf(-xv));
        }
        if (n0 + lr < n) wv = __ldg(w + static_cast<size_t>(n0 + lr) * k + kk);
      }
      sx[lk + q][lr] = xv;
      sw[lk + q][lr] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sx[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sw[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= n) continue;
      float v = acc[i][j] + (b ? b[c] : 0.f);
      if (addp) v += addp[static_cast<size_t>(r) * ld_add + c];
      if (silu_out) v = v / (1.0f + expf(-v));
      y[static_cast<size_t>(r) * ldy + c] = v;
    }
  }
}

__global__ void embedding_gather_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx,
                                        float* __restrict__ out, int n, int dim, int num_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * dim) return;
  const int b = i / dim, j = i - b * dim;
  // nn.Embedding raises IndexError for a label outside [0, num_rows); here the row becomes NaN (never a clamped,
  // plausible-looking embedding) and the host validates labels eagerly (unet.UNetModel.forward)
  const long long r = idx[b];
  out[i] = (r >= 0 && r < num_rows) ? table[static_cast<size_t>(r) * dim + j] : __int_as_float(0x7fc00000);
}

// ---------------------------------------------------------------------------------------------
// First conv (input_blocks.0.0, unet.py:483,741): C_in is 3 (or 6 for super-resolution), far too thin for a K block.
// The fp32 NCHW input is expanded to an fp16 NHWC im2col tensor with 64 channels, k = (ky*3+kx)*cin + ci (zero padded),
// and the conv then runs as a 1x1 problem with K = 64 on the tcgen05 kernel (engine.pack_conv_in packs the weights the
// same way).
// ---------------------------------------------------------------------------------------------
__global__ void im2col3x3_small_cin_kernel(const float* __restrict__ x, __half* __restrict__ out, int ld_out, int n,
                                           int cin, int h, int w) {
  // 8 threads per pixel, one 16-byte chunk (8 of the 64 K columns) each: a warp writes 4 whole 128-byte rows, so
  // the stores are fully coalesced; the fp32 NCHW reads of neighbouring pixels hit L1.  The grid stride is a
  // multiple of 8, so a thread keeps its chunk: the (tap, channel) decode of its 8 columns is hoisted out of the
  // loop and the loop body is 32-bit index arithmetic only (the host bounds n*h*w*8 < 2^31).
  const unsigned hw = static_cast<unsigned>(h) * static_cast<unsigned>(w);
  const unsigned total = static_cast<unsigned>(n) * hw * 8u;
  const unsigned tid0 = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = static_cast<int>(tid0 & 7u);
  int dy[8], dx[8];
  unsigned plane[8];  // channel plane offset of column j; dy = 2 marks a zero (padding) column
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = 8 * q + j;
    if (k < 9 * cin) {
      const int tap = k / cin, ci = k - tap * cin;
      dy[j] = tap / 3 - 1;
      dx[j] = tap - (tap / 3) * 3 - 1;
      plane[j] = static_cast<unsigned>(ci) * hw;
    } else {
      dy[j] = 2;
      dx[j] = 0;
      plane[j] = 0;
    }
  }
  for (unsigned i = tid0; i < total; i += gridDim.x * blockDim.x) {
    const unsigned p = i >> 3;
    const unsigned img = p / hw;
    const unsigned rem = p - img * hw;
    const int y = static_cast<int>(rem / static_cast<unsigned>(w));
    const int xx = static_cast<int>(rem) - y * w;
    const float* xi = x + static_cast<size_t>(img) * cin * hw;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int yy = y + dy[j], xc = xx + dx[j];
      const bool ok = dy[j] != 2 && yy >= 0 && yy < h && xc >= 0 && xc < w;
      f[j] = ok ? __ldg(xi + plane[j] + static_cast<unsigned>(yy * w + xc)) : 0.f;
    }
    st_half8(out + static_cast<size_t>(p) * ld_out + 8 * q, float_to_half8(f));
  }
}

__global__ void to_uint8_nhwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int n, int c, int h, int w) {
  const size_t hw = static_cast<size_t>(h) * w;
  const size_t total = static_cast<size_t>(n) * c * hw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // i indexes the OUTPUT (NHWC) so the stores are coalesced
    const size_t pix = i / c;
    const int ch = static_cast<int>(i - pix * c);
    const size_t b = pix / hw, r = pix - b * hw;
    float v = __fmul_rn(__fadd_rn(x[(b * c + ch) * hw + r], 1.0f), 127.5f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    out[i] = static_cast<uint8_t>(v);  // truncation, like .to(th.uint8)
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __half* __restrict__ out, int ld_out, int n, int c,
                                    int hw) {
  const size_t total = static_cast<size_t>(n) * hw * c;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = i / c;
    const int ch = static_cast<int>(i - pix * c);
    const size_t b = pix / hw, r = pix - b * hw;
    out[pix * ld_out + ch] = __float2half_rn(x[(b * c + ch) * hw + r]);
  }
}
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ x, int ld, float* __restrict__ out, int n, int c, int hw) {
  const size_t total = static_cast<size_t>(n) * hw * c;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // i indexes the OUTPUT (NCHW)
    const size_t r = i % hw;
    const size_t bc = i / hw;
    const int ch = static_cast<int>(bc % c);
    const size_t b = bc / c;
    out[i] = __half2float(x[(b * hw + r) * ld + ch]);
  }
}

// F.interpolate(mode="bilinear", align_corners=False) as used by SuperResModel.forward (unet.py:679)
__global__ void bilinear_kernel(const float* __restrict__ x, float* __restrict__ out, int n, int c, int hi, int wi,
                                int ho, int wo, int out_c_total, int out_c_offset) {
  const size_t total = static_cast<size_t>(n) * c * ho * wo;
  const float sy = static_cast<float>(hi) / static_cast<float>(ho);
  const float sx = static_cast<float>(wi) / static_cast<float>(wo);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % wo);
    const int oy = static_cast<int>((i / wo) % ho);
    const int ch = static_cast<int>((i / (static_cast<size_t>(wo) * ho)) % c);
    const int b = static_cast<int>(i / (static_cast<size_t>(wo) * ho * c));
    float fy = (oy + 0.5f) * sy - 0.5f;
    float fx = (ox + 0.5f) * sx - 0.5f;
    if (fy < 0.f) fy = 0.f;
    if (fx < 0.f) fx = 0.f;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    const float* xp = x + (static_cast<size_t>(b) * c + ch) * hi * wi;
    const float v = (1.f - ly) * ((1.f - lx) * xp[y0 * wi + x0] + lx * xp[y0 * wi + x1]) +
                    ly * ((1.f - lx) * xp[y1 * wi + x0] + lx * xp[y1 * wi + x1]);
    out[((static_cast<size_t>(b) * out_c_total + out_c_offset + ch) * ho + oy) * wo + ox] = v;
  }
}

// dlogits = scale * (onehot(y) - softmax(logits)); one warp per row
__global__ void logsoftmax_select_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ y,
                                             float* __restrict__ dlogits, int n, int classes, float scale) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* l = logits + static_cast<size_t>(row) * classes;
  float mx = -INFINITY;
  for (int j = lane; j < classes; j += 32) mx = fmaxf(mx, l[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < classes; j += 32) s += expf(l[j] - mx);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  const long long yy = y[row];
  for (int j = lane; j < classes; j += 32) {
    const float pr = expf(l[j] - mx) * inv;
    dlogits[static_cast<size_t>(row) * classes + j] = scale * ((j == yy ? 1.0f : 0.0f) - pr);
  }
}

inline int grid_for(size_t total, int block, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  if (g > static_cast<size_t>(cap)) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int gd_posterior_step(const gd_posterior_desc* d, void* stream) {
  GD_REQUIRE(d != nullptr, "gd_posterior_step: null descriptor");
  GD_REQUIRE(d->x && d->model_out && d->coef && d->t, "gd_posterior_step: null pointer");
  GD_REQUIRE(d->noise == nullptr || d->sample != nullptr, "gd_posterior_step: noise given but no sample output");
  GD_REQUIRE(d->ddim >= 0 && d->ddim <= GD_DDIM_REVERSE, "gd_posterior_step: ddim must be 0, 1 or GD_DDIM_REVERSE");
  GD_REQUIRE(d->ddim != GD_DDIM_REVERSE || (d->sample != nullptr && d->eta == 0.0f && d->grad == nullptr),
             "gd_posterior_step: the reverse ODE step needs a sample output, eta == 0 and takes no guidance gradient");
  GD_REQUIRE(d->n > 0 && d->c > 0 && d->hw > 0, "gd_posterior_step: bad shape");
  GD_REQUIRE(d->num_timesteps > 0, "gd_posterior_step: num_timesteps (rows of the coefficient table) must be > 0");
  GD_REQUIRE(d->var_type >= GD_VAR_LEARNED_RANGE && d->var_type <= GD_VAR_LEARNED, "gd_posterior_step: bad var_type");
  GD_REQUIRE(d->mean_type == GD_MEAN_EPSILON || d->mean_type == GD_MEAN_START_X,
             "gd_posterior_step: model_mean_type PREVIOUS_X is not on the sampling path built here");
  PostArgs p;
  p.x = d->x; p.model_out = d->model_out; p.grad = d->grad; p.noise = d->noise;
  p.sample = d->sample; p.pred_xstart = d->pred_xstart; p.coef = d->coef; p.t = d->t;
  p.mean_out = d->mean_out; p.var_out = d->var_out; p.logvar_out = d->logvar_out;
  p.n = d->n; p.c = d->c; p.hw = d->hw;
  p.var_type = d->var_type; p.mean_type = d->mean_type; p.clip = d->clip_denoised; p.ddim = d->ddim; p.eta = d->eta;
  p.num_timesteps = d->num_timesteps;
  const size_t total = static_cast<size_t>(d->n) * d->c * d->hw;
  uintptr_t align = 0;
  for (const void* q : {static_cast<const void*>(d->x), static_cast<const void*>(d->model_out),
                        static_cast<const void*>(d->grad), static_cast<const void*>(d->noise),
                        static_cast<const void*>(d->sample), static_cast<const void*>(d->pred_xstart),
                        static_cast<const void*>(d->mean_out), static_cast<const void*>(d->var_out),
                        static_cast<const void*>(d->logvar_out)})
    align |= reinterpret_cast<uintptr_t>(q);
  const bool vec4 = (static_cast<size_t>(d->c) * d->hw) % 4 == 0 && (align & 15u) == 0;
  if (vec4)
    GD_CHECK_CUDA(launch_pdl(posterior_kernel<4>, dim3(grid_for(total / 4, 256, 148 * 8)), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), p));
  else
    GD_CHECK_CUDA(launch_pdl(posterior_kernel<1>, dim3(grid_for(total, 256)), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), p));
  count_launch(1);
  return 0;
}

extern "C" int gd_timestep_embedding(const float* t, float* out, int32_t n, int32_t dim, void* stream) {
  GD_REQUIRE(t && out && n > 0 && dim >= 2, "gd_timestep_embedding: bad arguments");
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, out, n, dim);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_linear_f32(const float* x, int32_t ldx, const float* w, const float* b, const float* add,
                             int32_t ld_add, float* y, int32_t ldy, int32_t m, int32_t k, int32_t n, int32_t silu_in,
                             int32_t silu_out, void* stream) {
  GD_REQUIRE(x && w && y && m > 0 && k > 0 && n > 0, "gd_linear_f32: bad arguments");
  GD_REQUIRE(ldx >= k && ldy >= n, "gd_linear_f32: bad leading dimensions");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (m > 16) {  // many rows: tiled GEMM (weights are re-used across rows through shared memory)
    dim3 grid2((n + 63) / 64, (m + 63) / 64);
    GD_CHECK_CUDA(launch_pdl(sgemm_nt_kernel, grid2, dim3(256), 0, st, x, ldx, w, b, add, ld_add, y, ldy, m, k, n, silu_in,
                             silu_out));
    count_launch(1);
    return 0;
  }
  const int warps = 8;
  const int grid = (n + warps - 1) / warps;
  if (silu_in)
    GD_CHECK_CUDA(launch_pdl(linear_f32_kernel<true>, dim3(grid), dim3(warps * 32), 0, st, x, ldx, w, b, add, ld_add, y, ldy,
                             m, k, n, silu_out));
  else
    GD_CHECK_CUDA(launch_pdl(linear_f32_kernel<false>, dim3(grid), dim3(warps * 32), 0, st, x, ldx, w, b, add, ld_add, y,
                             ldy, m, k, n, silu_out));
  count_launch(1);
  return 0;
}

extern "C" int gd_embedding_gather(const float* table, const int64_t* idx, float* out, int32_t n, int32_t dim,
                                   int32_t num_rows, void* stream) {
  GD_REQUIRE(table && idx && out && n > 0 && dim > 0 && num_rows > 0, "gd_embedding_gather: bad arguments");
  const int total = n * dim;
  embedding_gather_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table, idx, out, n,
                                                                                                 dim, num_rows);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_im2col3x3_small_cin(const float* x, void* out, int32_t ld_out, int32_t n, int32_t cin, int32_t h,
                                      int32_t w, void* stream) {
  GD_REQUIRE(x && out, "gd_im2col3x3_small_cin: null pointer");
  GD_REQUIRE(cin >= 1 && cin * 9 <= 64, "gd_im2col3x3_small_cin: cin*9 must fit one 64-wide K block, got cin=%d", cin);
  GD_REQUIRE(ld_out >= 64 && ld_out % 8 == 0, "gd_im2col3x3_small_cin: bad ld_out %d", ld_out);
  const size_t total = static_cast<size_t>(n) * h * w;
  GD_REQUIRE(total * 8 < (1ull << 31), "gd_im2col3x3_small_cin: n*h*w too large (%zu)", total);
  im2col3x3_small_cin_kernel<<<grid_for(total * 8, 256, 148 * 32), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__half*>(out), ld_out, n, cin, h, w);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

namespace gd {
namespace {
// CTA = 8 x 32 output pixels of one image.  Its (8+2) x (32+2) halo of ytap rows (64 fp16 = 32 words each) is staged in
// shared memory with 16-byte global loads (rows padded to 33 words: the word-wise stores of a warp and the later
// reads of 32 horizontally adjacent pixels are both bank-conflict free), then every thread sums its 9 taps x cout
// values from there; pixels outside the image are staged as zeros = the conv's zero padding.  fp32 NCHW stores are
// coalesced along x.
constexpr int kTgW = 32, kTgH = 8, kTgRow = 33;
template <int kCout>
__global__ void __launch_bounds__(kTgW * kTgH)
tap_gather3x3_kernel(const __half* __restrict__ ytap, int ld, const float* __restrict__ bias, float* __restrict__ out,
                     int h, int w, float out_scale) {
  pdl_enter();
  __shared__ uint32_t tile[(kTgH + 2) * (kTgW + 2) * kTgRow];
  const int n = blockIdx.z, y0 = blockIdx.y * kTgH, x0 = blockIdx.x * kTgW;
  const size_t plane = static_cast<size_t>(h) * w;
  constexpr int kChunks = (9 * kCout * 2 + 15) / 16;  // 16-byte chunks that hold the 9*cout useful halves
  for (int i = threadIdx.x; i < (kTgH + 2) * (kTgW + 2) * kChunks; i += kTgW * kTgH) {
    const int r = i / kChunks, c = i - r * kChunks;
    const int yy = y0 + r / (kTgW + 2) - 1, xx = x0 + r % (kTgW + 2) - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < h && xx >= 0 && xx < w)
      v = __ldg(reinterpret_cast<const uint4*>(ytap + (static_cast<size_t>(n) * plane + static_cast<size_t>(yy) * w + xx) * ld) + c);
    uint32_t* dst = tile + r * kTgRow + 4 * c;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  __syncthreads();
  const int tx = threadIdx.x % kTgW, ty = threadIdx.x / kTgW;
  const int x = x0 + tx, y = y0 + ty;
  if (x >= w || y >= h) return;
  float acc[kCout];
#pragma unroll
  for (int c = 0; c < kCout; ++c) acc[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const uint32_t* src = tile + ((ty + t / 3) * (kTgW + 2) + tx + t % 3) * kTgRow;
#pragma unroll
    for (int c = 0; c < kCout; ++c) {
      const int hidx = t * kCout + c;
      const uint32_t wv = src[hidx >> 1];
      const __half hv = __ushort_as_half(static_cast<unsigned short>((hidx & 1) ? (wv >> 16) : (wv & 0xffffu)));
      acc[c] += __half2float(hv);
    }
  }
  float* o = out + static_cast<size_t>(n) * kCout * plane + static_cast<size_t>(y) * w + x;
#pragma unroll
  for (int c = 0; c < kCout; ++c) o[c * plane] = acc[c] * out_scale;
}
}  // namespace
}  // namespace gd

extern "C" int gd_tap_gather3x3(const void* ytap, int32_t ld, const float* bias, float* out, int32_t n, int32_t cout,
                                int32_t h, int32_t w, float out_scale, void* stream) {
  GD_REQUIRE(ytap && out && n > 0 && h > 0 && w > 0, "gd_tap_gather3x3: bad arguments");
  GD_REQUIRE(cout >= 1 && cout <= 7 && ld >= 64 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(ytap) & 15u) == 0,
             "gd_tap_gather3x3: cout %d outside [1,7], or ytap rows (ld %d) not 16-byte aligned / shorter than 64", cout, ld);
  const __half* y = reinterpret_cast<const __half*>(ytap);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid((w + kTgW - 1) / kTgW, (h + kTgH - 1) / kTgH, n);
  switch (cout) {
#define GD_TAP_CASE(K) \
  case K: GD_CHECK_CUDA(launch_pdl(tap_gather3x3_kernel<K>, grid, dim3(kTgW * kTgH), 0, st, y, ld, bias, out, h, w, out_scale)); break;
    GD_TAP_CASE(1) GD_TAP_CASE(2) GD_TAP_CASE(3) GD_TAP_CASE(4) GD_TAP_CASE(5) GD_TAP_CASE(6) GD_TAP_CASE(7)
#undef GD_TAP_CASE
  }
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_to_uint8_nhwc(const float* x, uint8_t* out, int32_t n, int32_t c, int32_t h, int32_t w, void* stream) {
  GD_REQUIRE(x && out && n > 0 && c > 0 && h > 0 && w > 0, "gd_to_uint8_nhwc: bad arguments");
  const size_t total = static_cast<size_t>(n) * c * h * w;
  to_uint8_nhwc_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, n, c, h, w);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_nchw_f32_to_nhwc_f16(const float* x, void* out, int32_t ld_out, int32_t n, int32_t c, int32_t h,
                                       int32_t w, void* stream) {
  GD_REQUIRE(x && out && ld_out >= c, "gd_nchw_f32_to_nhwc_f16: bad arguments");
  const size_t total = static_cast<size_t>(n) * c * h * w;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__half*>(out), ld_out, n, c, h * w);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_nhwc_f16_to_nchw_f32(const void* x, int32_t ld, float* out, int32_t n, int32_t c, int32_t h, int32_t w,
                                       void* stream) {
  GD_REQUIRE(x && out && ld >= c, "gd_nhwc_f16_to_nchw_f32: bad arguments");
  const size_t total = static_cast<size_t>(n) * c * h * w;
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, out, n, c, h * w);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_bilinear_upsample_nchw(const float* x, float* out, int32_t n, int32_t c, int32_t h_in, int32_t w_in,
                                         int32_t h_out, int32_t w_out, int32_t out_c_total, int32_t out_c_offset,
                                         void* stream) {
  GD_REQUIRE(x && out && n > 0 && c > 0 && out_c_offset + c <= out_c_total, "gd_bilinear_upsample_nchw: bad arguments");
  const size_t total = static_cast<size_t>(n) * c * h_out * w_out;
  bilinear_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, out, n, c, h_in, w_in, h_out, w_out, out_c_total, out_c_offset);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_logsoftmax_select_bwd(const float* logits, const int64_t* y, float* dlogits, int32_t n,
                                        int32_t classes, float scale, void* stream) {
  GD_REQUIRE(logits && y && dlogits && n > 0 && classes > 0, "gd_logsoftmax_select_bwd: bad arguments");
  const int warps = 4;
  logsoftmax_select_bwd_kernel<<<(n + warps - 1) / warps, warps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, y, dlogits, n, classes, scale);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// conv_resample layers of models built with resblock_updown=False (the reference factory default,
// script_util.py:60): Downsample = 3x3 stride-2 conv (unet.py:125-136), Upsample = nearest x2 then 3x3 conv
// (unet.py:91-110); ResBlocks without FiLM add the embedding before the second GroupNorm (unet.py:253-255).
// All three are single bandwidth-bound passes over fp16 NHWC with 16-byte vectors; the contractions run on
// gd_conv_igemm (the strided conv as a taps = 1 GEMM over the gathered [.., 9*C] rows).
// ---------------------------------------------------------------------------------------------
namespace gd {
namespace {

__global__ void __launch_bounds__(256)
im2col3x3_s2_kernel(const __half* __restrict__ x, int ld, __half* __restrict__ out, int ld_out, int n, int h, int w,
                    int c8, int ho, int wo) {
  const size_t total = static_cast<size_t>(n) * ho * wo * 9 * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    size_t r = i / c8;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int xo = static_cast<int>(r % wo);
    r /= wo;
    const int yo = static_cast<int>(r % ho);
    const int img = static_cast<int>(r / ho);
    const int yy = 2 * yo + tap / 3 - 1, xx = 2 * xo + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (yy >= 0 && yy < h && xx >= 0 && xx < w)
      v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(img) * h + yy) * w + xx) * ld) + cc);
    *(reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * ho + yo) * wo + xo) * ld_out + tap * c8 * 8) + cc) = v;
  }
}

__global__ void __launch_bounds__(256)
upsample2_nhwc_kernel(const __half* __restrict__ x, int ld, __half* __restrict__ out, int ld_out, int n, int h, int w,
                      int c8) {
  const size_t total = static_cast<size_t>(n) * (2 * h) * (2 * w) * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    size_t r = i / c8;
    const int xo = static_cast<int>(r % (2 * w));
    r /= 2 * w;
    const int yo = static_cast<int>(r % (2 * h));
    const int img = static_cast<int>(r / (2 * h));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(img) * h + (yo >> 1)) * w + (xo >> 1)) * ld) + cc);
    *(reinterpret_cast<uint4*>(out + ((static_cast<size_t>(img) * 2 * h + yo) * 2 * w + xo) * ld_out) + cc) = v;
  }
}

__global__ void __launch_bounds__(256)
add_emb_nhwc_kernel(__half* __restrict__ x, int ld, const float* __restrict__ emb, int ld_emb, int n, int hw, int c8) {
  const size_t total = static_cast<size_t>(n) * hw * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    const size_t px = i / c8;
    const int img = static_cast<int>(px / hw);
    uint4* p = reinterpret_cast<uint4*>(x + px * ld) + cc;
    uint4 v = *p;
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(emb + static_cast<size_t>(img) * ld_emb + cc * 8));
    const float4 e1 = __ldg(reinterpret_cast<const float4*>(emb + static_cast<size_t>(img) * ld_emb + cc * 8) + 1);
    __half2* hv = reinterpret_cast<__half2*>(&v);
    float2 f;
    f = __half22float2(hv[0]); hv[0] = __floats2half2_rn(f.x + e0.x, f.y + e0.y);
    f = __half22float2(hv[1]); hv[1] = __floats2half2_rn(f.x + e0.z, f.y + e0.w);
    f = __half22float2(hv[2]); hv[2] = __floats2half2_rn(f.x + e1.x, f.y + e1.y);
    f = __half22float2(hv[3]); hv[3] = __floats2half2_rn(f.x + e1.z, f.y + e1.w);
    *p = v;
  }
}

// transpose of im2col3x3_s2_kernel: dx[n][y][x][c] = sum over the (ky, kx) with y = 2*yo + ky - 1, x = 2*xo + kx - 1 of
// dcols[n][yo][xo][(ky*3+kx)*C + c]  (1, 2 or 4 terms per pixel); fp32 sums, fp16 result
__global__ void __launch_bounds__(256)
col2im3x3_s2_kernel(const __half* __restrict__ dcols, int ld, __half* __restrict__ dx, int ld_dx, int n, int h, int w,
                    int c8, int ho, int wo) {
  const size_t total = static_cast<size_t>(n) * h * w * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    size_t r = i / c8;
    const int x = static_cast<int>(r % w);
    r /= w;
    const int y = static_cast<int>(r % h);
    const int img = static_cast<int>(r / h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ty = y + 1 - ky;
      if (ty < 0 || (ty & 1) != 0 || (ty >> 1) >= ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int tx = x + 1 - kx;
        if (tx < 0 || (tx & 1) != 0 || (tx >> 1) >= wo) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(
                                  dcols + ((static_cast<size_t>(img) * ho + (ty >> 1)) * wo + (tx >> 1)) * ld +
                                  (ky * 3 + kx) * c8 * 8) + cc);
        const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(hv[j]);
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
    uint4 o;
    __half2* ho2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ho2[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    *(reinterpret_cast<uint4*>(dx + ((static_cast<size_t>(img) * h + y) * w + x) * ld_dx) + cc) = o;
  }
}

int check_nhwc(const char* who, const void* x, int ld, const void* out, int ld_out, int n, int h, int w, int c, int c_out) {
  GD_REQUIRE(x && out, "%s: null pointer", who);
  GD_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 8 == 0, "%s: bad shape n=%d h=%d w=%d c=%d (c must be a multiple of 8)",
             who, n, h, w, c);
  GD_REQUIRE(ld >= c && ld % 8 == 0 && ld_out >= c_out && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
             "%s: rows must be 16-byte aligned (ld %d, ld_out %d)", who, ld, ld_out);
  return 0;
}

}  // namespace
}  // namespace gd

extern "C" int gd_im2col3x3_s2_nhwc(const void* x, int32_t ld, void* out, int32_t ld_out, int32_t n, int32_t h, int32_t w,
                                    int32_t c, void* stream) {
  if (int rc = check_nhwc("gd_im2col3x3_s2_nhwc", x, ld, out, ld_out, n, h, w, c, 9 * c)) return rc;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t total = static_cast<size_t>(n) * ho * wo * 9 * (c / 8);
  im2col3x3_s2_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, reinterpret_cast<__half*>(out), ld_out, n, h, w, c / 8, ho, wo);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_upsample2_nhwc(const void* x, int32_t ld, void* out, int32_t ld_out, int32_t n, int32_t h, int32_t w,
                                 int32_t c, void* stream) {
  if (int rc = check_nhwc("gd_upsample2_nhwc", x, ld, out, ld_out, n, h, w, c, c)) return rc;
  const size_t total = static_cast<size_t>(n) * 4 * h * w * (c / 8);
  upsample2_nhwc_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, reinterpret_cast<__half*>(out), ld_out, n, h, w, c / 8);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_add_emb_nhwc(void* x, int32_t ld, const float* emb, int32_t ld_emb, int32_t n, int32_t hw, int32_t c,
                               void* stream) {
  if (int rc = check_nhwc("gd_add_emb_nhwc", x, ld, x, ld, n, hw, 1, c, c)) return rc;
  GD_REQUIRE(emb != nullptr && ld_emb >= c && ld_emb % 4 == 0 && (reinterpret_cast<uintptr_t>(emb) & 15u) == 0,
             "gd_add_emb_nhwc: embedding rows must be 16-byte aligned fp32 (ld_emb %d)", ld_emb);
  const size_t total = static_cast<size_t>(n) * hw * (c / 8);
  add_emb_nhwc_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__half*>(x), ld, emb, ld_emb, n, hw, c / 8);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_col2im3x3_s2_nhwc(const void* dcols, int32_t ld, void* dx, int32_t ld_dx, int32_t n, int32_t h, int32_t w,
                                    int32_t c, void* stream) {
  if (int rc = check_nhwc("gd_col2im3x3_s2_nhwc", dx, ld_dx, dcols, ld, n, h, w, c, 9 * c)) return rc;
  const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
  const size_t total = static_cast<size_t>(n) * h * w * (c / 8);
  col2im3x3_s2_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(dcols), ld, reinterpret_cast<__half*>(dx), ld_dx, n, h, w, c / 8, ho, wo);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
