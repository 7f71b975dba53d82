// Kernels of the CLIP ViT image-encoder guidance path (spec §8c) that are not convolutions or
// attention: LayerNorm forward / data-gradient, QuickGELU forward / backward, the fused preprocessing
// (x in [-1,1] -> (x+1)/2 -> bilinear resize, align_corners=False -> CLIP normalisation -> 16x16 patches) and its
// transpose, and the similarity head (projection, L2 normalisation, s*<e_img,e_txt>, and its gradient).
// Token tensors are fp16 [rows][C] views with a row stride (sequence padded to a multiple of 64 tokens); statistics
// and the head run in fp32.  All bandwidth-/latency-bound, < 3 % of a guided step (the ViT is ~35 GFLOP per sample
// against 2 240 for the UNet), so the design goal is "one pass, vector loads", not a roofline.
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

constexpr int kLnMaxChunks = 8;  // per lane -> C <= 32 * 8 * 8 = 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row
__global__ void layernorm_fwd_kernel(const __half* __restrict__ x, int ld, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, __half* __restrict__ out, int ld_out,
                                     float* __restrict__ stats, int rows, int c) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int c8 = c >> 3;
  const __half* xr = x + static_cast<size_t>(row) * ld;
  float f[kLnMaxChunks][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < c8) {
      half8_to_float(ld_half8(xr + ch * 8), f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(c);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    if (lane + 32 * i < c8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        v = fmaf(d, d, v);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(v) / static_cast<float>(c) + eps);
  if (stats != nullptr && lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
  __half* orow = out + static_cast<size_t>(row) * ld_out;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < c8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf((f[i][j] - mean) * rstd, gamma[ch * 8 + j], beta[ch * 8 + j]);
      st_half8(orow + ch * 8, float_to_half8(o));
    }
  }
}

// dx = rstd * (g - mean(g) - xh * mean(g * xh)) (+ add),  g = dy * gamma, xh = (x - mean) * rstd
__global__ void layernorm_bwd_kernel(const __half* __restrict__ x, int ld, const float* __restrict__ stats,
                                     const float* __restrict__ gamma, const __half* __restrict__ dy, int ld_dy,
                                     const __half* __restrict__ add, int ld_add, __half* __restrict__ dx, int ld_dx,
                                     int rows, int c) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int c8 = c >> 3;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  const __half* xr = x + static_cast<size_t>(row) * ld;
  const __half* dr = dy + static_cast<size_t>(row) * ld_dy;
  float xh[kLnMaxChunks][8], g[kLnMaxChunks][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < c8) {
      float xf[8], df[8];
      half8_to_float(ld_half8(xr + ch * 8), xf);
      half8_to_float(ld_half8(dr + ch * 8), df);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xf[j] - mean) * rstd;
        g[i][j] = df[j] * gamma[ch * 8 + j];
        s1 += g[i][j];
        s2 = fmaf(g[i][j], xh[i][j], s2);
      }
    }
  }
  const float m1 = warp_sum(s1) / static_cast<float>(c);
  const float m2 = warp_sum(s2) / static_cast<float>(c);
  __half* orow = dx + static_cast<size_t>(row) * ld_dx;
#pragma unroll
  for (int i = 0; i < kLnMaxChunks; ++i) {
    const int ch = lane + 32 * i;
    if (ch < c8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[i][j] - m1 - xh[i][j] * m2);
      if (add != nullptr) {
        float a[8];
        half8_to_float(ld_half8(add + static_cast<size_t>(row) * ld_add + ch * 8), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += a[j];
      }
      st_half8(orow + ch * 8, float_to_half8(o));
    }
  }
}

// QuickGELU: y = x * sigmoid(1.702 x);  dy/dx = s + 1.702 x s (1 - s)
template <bool kBwd>
__global__ void quickgelu_kernel(const __half* __restrict__ x, int ld, const __half* __restrict__ dy, int ld_dy,
                                 __half* __restrict__ out, int ld_out, int rows, int c8) {
  const size_t total = static_cast<size_t>(rows) * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t row = i / c8;
    const int ch = static_cast<int>(i - row * c8);
    float f[8], o[8];
    half8_to_float(ld_half8(x + row * ld + ch * 8), f);
    if (kBwd) {
      float d[8];
      half8_to_float(ld_half8(dy + row * ld_dy + ch * 8), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = sigmoid_f(1.702f * f[j]);
        o[j] = d[j] * (s + 1.702f * f[j] * s * (1.0f - s));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = f[j] * sigmoid_f(1.702f * f[j]);
    }
    st_half8(out + row * ld_out + ch * 8, float_to_half8(o));
  }
}

__constant__ float kClipMean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
__constant__ float kClipStd[3] = {0.26862954f, 0.26130258f, 0.27577711f};

// source coordinate of output index o (PyTorch bilinear, align_corners=False): max(0, (o + 0.5) * scale - 0.5)
__device__ __forceinline__ void bilinear_src(int o, float scale, int in_size, int& i0, int& i1, float& lam) {
  float src = (static_cast<float>(o) + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + 1 < in_size ? i0 + 1 : in_size - 1;
  lam = src - static_cast<float>(i0);
}

// patches[n][token][k], token 1 + py*g + px, k = c*P*P + dy*P + dx (the layout nn.Conv2d(3, H, P, stride=P) flattens
// to); token 0 (class-token slot) and the padding tokens are zero.  One thread per 8 consecutive dx.
__global__ void clip_preprocess_fwd_kernel(const float* __restrict__ x, __half* __restrict__ patches, int ld, int n,
                                           int hin, int win, int size, int patch, int t_pad) {
  const int g = size / patch;
  const int kdim = 3 * patch * patch;
  const int k8 = kdim >> 3;
  const size_t total = static_cast<size_t>(n) * t_pad * k8;
  const float sy = static_cast<float>(hin) / static_cast<float>(size);
  const float sx = static_cast<float>(win) / static_cast<float>(size);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(i % k8);
    const size_t rt = i / k8;
    const int tok = static_cast<int>(rt % t_pad);
    const int img = static_cast<int>(rt / t_pad);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    if (tok >= 1 && tok <= g * g) {
      const int py = (tok - 1) / g, px = (tok - 1) - py * g;
      const int k0 = kc * 8;
      const int c = k0 / (patch * patch);
      const int rem = k0 - c * patch * patch;
      const int dy = rem / patch, dx0 = rem - dy * patch;
      const int oy = py * patch + dy;
      int y0, y1;
      float ly;
      bilinear_src(oy, sy, hin, y0, y1, ly);
      const float* xc = x + (static_cast<size_t>(img) * 3 + c) * hin * win;
      const float inv_std = 1.0f / kClipStd[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = px * patch + dx0 + j;
        int x0, x1;
        float lx;
        bilinear_src(ox, sx, win, x0, x1, lx);
        const float top = __ldg(xc + y0 * win + x0) * (1.f - lx) + __ldg(xc + y0 * win + x1) * lx;
        const float bot = __ldg(xc + y1 * win + x0) * (1.f - lx) + __ldg(xc + y1 * win + x1) * lx;
        const float v = top * (1.f - ly) + bot * ly;           // bilinear(x); (x+1)/2 commutes with it
        o[j] = ((v + 1.0f) * 0.5f - kClipMean[c]) * inv_std;
      }
    }
    st_half8(patches + (static_cast<size_t>(img) * t_pad + tok) * ld + kc * 8, float_to_half8(o));
  }
}

// transpose of the above: dx[n][c][iy][ix] = out_scale * 0.5 / std_c * sum over the resized pixels that read (iy, ix)
__global__ void clip_preprocess_bwd_kernel(const __half* __restrict__ dpatches, int ld, float* __restrict__ dx, int n,
                                           int hin, int win, int size, int patch, int t_pad, float out_scale) {
  const int g = size / patch;
  const size_t total = static_cast<size_t>(n) * 3 * hin * win;
  const float sy = static_cast<float>(hin) / static_cast<float>(size);
  const float sx = static_cast<float>(win) / static_cast<float>(size);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ix = static_cast<int>(i % win);
    size_t r = i / win;
    const int iy = static_cast<int>(r % hin);
    r /= hin;
    const int c = static_cast<int>(r % 3);
    const int img = static_cast<int>(r / 3);
    // candidate resized rows / columns: those whose two source taps can include iy / ix
    int oy_lo = static_cast<int>(floorf((static_cast<float>(iy) - 0.5f) / sy - 0.5f)) - 1;
    int ox_lo = static_cast<int>(floorf((static_cast<float>(ix) - 0.5f) / sx - 0.5f)) - 1;
    if (oy_lo < 0) oy_lo = 0;
    if (ox_lo < 0) ox_lo = 0;
    const int span_y = static_cast<int>(2.0f / sy) + 4, span_x = static_cast<int>(2.0f / sx) + 4;
    float acc = 0.f;
    for (int oy = oy_lo; oy < oy_lo + span_y && oy < size; ++oy) {
      int y0, y1;
      float ly;
      bilinear_src(oy, sy, hin, y0, y1, ly);
      const float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f);
      if (wy == 0.f) continue;
      const int py = oy / patch, dyy = oy - py * patch;
      for (int ox = ox_lo; ox < ox_lo + span_x && ox < size; ++ox) {
        int x0, x1;
        float lx;
        bilinear_src(ox, sx, win, x0, x1, lx);
        const float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
        if (wx == 0.f) continue;
        const int px = ox / patch, dxx = ox - px * patch;
        const int tok = 1 + py * g + px;
        const int k = c * patch * patch + dyy * patch + dxx;
        acc += wy * wx * __half2float(dpatches[(static_cast<size_t>(img) * t_pad + tok) * ld + k]);
      }
    }
    dx[i] = acc * out_scale * 0.5f / kClipStd[c];
  }
}

// One CTA per sample: e = Wp f, sim = s <e/|e|, t>, df = grad_scale * Wp^T (s (t - ehat <ehat,t>) / |e|)
__global__ void __launch_bounds__(256)
clip_head_kernel(const __half* __restrict__ f, int ld_f, const float* __restrict__ wproj, const float* __restrict__ text,
                 int text_stride, float scale, float grad_scale, float* __restrict__ sim, __half* __restrict__ df,
                 int ld_df, int h, int p) {
  extern __shared__ float sm[];  // [h] f | [p] e (then de) | [2 * 8] reductions
  float* sf = sm;
  float* se = sm + h;
  float* red = se + p;
  const int img = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < h; i += blockDim.x) sf[i] = __half2float(f[static_cast<size_t>(img) * ld_f + i]);
  __syncthreads();
  for (int j = warp; j < p; j += 8) {
    const float* w = wproj + static_cast<size_t>(j) * h;
    float a = 0.f;
    for (int i = lane; i < h; i += 32) a = fmaf(w[i], sf[i], a);
    a = warp_sum(a);
    if (lane == 0) se[j] = a;
  }
  __syncthreads();
  const float* t = text + static_cast<size_t>(img) * text_stride;
  float n2 = 0.f, dt = 0.f;
  for (int j = tid; j < p; j += blockDim.x) {
    n2 = fmaf(se[j], se[j], n2);
    dt = fmaf(se[j], t[j], dt);
  }
  n2 = warp_sum(n2);
  dt = warp_sum(dt);
  if (lane == 0) {
    red[warp] = n2;
    red[8 + warp] = dt;
  }
  __syncthreads();
  float norm2 = 0.f, dot = 0.f;
#pragma unroll
  for (int w8 = 0; w8 < 8; ++w8) {
    norm2 += red[w8];
    dot += red[8 + w8];
  }
  const float inv = rsqrtf(norm2);
  const float cosv = dot * inv;
  if (tid == 0 && sim != nullptr) sim[img] = scale * cosv;
  __syncthreads();
  for (int j = tid; j < p; j += blockDim.x) se[j] = scale * inv * (t[j] - se[j] * inv * cosv);  // d sim / d e_j
  __syncthreads();
  if (df != nullptr) {
    for (int i = tid; i < h; i += blockDim.x) {
      float a = 0.f;
      for (int j = 0; j < p; ++j) a = fmaf(wproj[static_cast<size_t>(j) * h + i], se[j], a);
      df[static_cast<size_t>(img) * ld_df + i] = __float2half_rn(a * grad_scale);
    }
  }
}

int check_rows(const char* who, const void* x, int ld, int rows, int c) {
  GD_REQUIRE(x != nullptr, "%s: null pointer", who);
  GD_REQUIRE(rows > 0 && c > 0 && c % 8 == 0 && ld >= c && ld % 8 == 0, "%s: bad rows/c/ld (%d, %d, %d)", who, rows, c, ld);
  return 0;
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int gd_layernorm_fwd(const void* x, int32_t ld, const float* gamma, const float* beta, float eps, void* out,
                                int32_t ld_out, float* mean_rstd, int32_t rows, int32_t c, void* stream) {
  if (int rc = check_rows("gd_layernorm_fwd", x, ld, rows, c)) return rc;
  GD_REQUIRE(gamma && beta && out && ld_out >= c && ld_out % 8 == 0, "gd_layernorm_fwd: bad arguments");
  GD_REQUIRE(c <= 32 * 8 * kLnMaxChunks, "gd_layernorm_fwd: c %d > %d", c, 32 * 8 * kLnMaxChunks);
  layernorm_fwd_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, gamma, beta, eps, reinterpret_cast<__half*>(out), ld_out, mean_rstd, rows, c);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_layernorm_bwd(const void* x, int32_t ld, const float* mean_rstd, const float* gamma, const void* dy,
                                int32_t ld_dy, const void* add, int32_t ld_add, void* dx, int32_t ld_dx, int32_t rows,
                                int32_t c, void* stream) {
  if (int rc = check_rows("gd_layernorm_bwd", x, ld, rows, c)) return rc;
  GD_REQUIRE(mean_rstd && gamma && dy && dx && ld_dy >= c && ld_dx >= c && ld_dy % 8 == 0 && ld_dx % 8 == 0,
             "gd_layernorm_bwd: bad arguments");
  if (add) GD_REQUIRE(ld_add >= c && ld_add % 8 == 0, "gd_layernorm_bwd: bad ld_add %d", ld_add);
  GD_REQUIRE(c <= 32 * 8 * kLnMaxChunks, "gd_layernorm_bwd: c %d > %d", c, 32 * 8 * kLnMaxChunks);
  layernorm_bwd_kernel<<<(rows + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, mean_rstd, gamma, reinterpret_cast<const __half*>(dy), ld_dy,
      reinterpret_cast<const __half*>(add), ld_add, reinterpret_cast<__half*>(dx), ld_dx, rows, c);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_quickgelu_fwd(const void* x, int32_t ld, void* out, int32_t ld_out, int32_t rows, int32_t c,
                                void* stream) {
  if (int rc = check_rows("gd_quickgelu_fwd", x, ld, rows, c)) return rc;
  GD_REQUIRE(out && ld_out >= c && ld_out % 8 == 0, "gd_quickgelu_fwd: bad output");
  const size_t total = static_cast<size_t>(rows) * (c / 8);
  const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  quickgelu_kernel<false><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, nullptr, 0, reinterpret_cast<__half*>(out), ld_out, rows, c / 8);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_quickgelu_bwd(const void* x, int32_t ld, const void* dy, int32_t ld_dy, void* dx, int32_t ld_dx,
                                int32_t rows, int32_t c, void* stream) {
  if (int rc = check_rows("gd_quickgelu_bwd", x, ld, rows, c)) return rc;
  GD_REQUIRE(dy && dx && ld_dy >= c && ld_dx >= c && ld_dy % 8 == 0 && ld_dx % 8 == 0, "gd_quickgelu_bwd: bad arguments");
  const size_t total = static_cast<size_t>(rows) * (c / 8);
  const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  quickgelu_kernel<true><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(x), ld, reinterpret_cast<const __half*>(dy), ld_dy, reinterpret_cast<__half*>(dx),
      ld_dx, rows, c / 8);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_clip_preprocess_fwd(const float* x, void* patches, int32_t ld, int32_t n, int32_t hin, int32_t win,
                                      int32_t size, int32_t patch, int32_t t_pad, void* stream) {
  GD_REQUIRE(x && patches && n > 0 && hin > 0 && win > 0, "gd_clip_preprocess_fwd: bad arguments");
  GD_REQUIRE(patch > 0 && patch % 8 == 0 && size % patch == 0, "gd_clip_preprocess_fwd: size %d / patch %d", size, patch);
  const int g = size / patch;
  GD_REQUIRE(t_pad >= 1 + g * g && ld >= 3 * patch * patch && ld % 8 == 0, "gd_clip_preprocess_fwd: t_pad %d / ld %d", t_pad, ld);
  const size_t total = static_cast<size_t>(n) * t_pad * (3 * patch * patch / 8);
  const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  clip_preprocess_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, reinterpret_cast<__half*>(patches), ld, n, hin, win, size, patch, t_pad);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_clip_preprocess_bwd(const void* dpatches, int32_t ld, float* dx, int32_t n, int32_t hin, int32_t win,
                                      int32_t size, int32_t patch, int32_t t_pad, float out_scale, void* stream) {
  GD_REQUIRE(dpatches && dx && n > 0 && hin > 0 && win > 0, "gd_clip_preprocess_bwd: bad arguments");
  GD_REQUIRE(patch > 0 && size % patch == 0, "gd_clip_preprocess_bwd: size %d / patch %d", size, patch);
  const int g = size / patch;
  GD_REQUIRE(t_pad >= 1 + g * g && ld >= 3 * patch * patch, "gd_clip_preprocess_bwd: t_pad %d / ld %d", t_pad, ld);
  const size_t total = static_cast<size_t>(n) * 3 * hin * win;
  const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  clip_preprocess_bwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(dpatches), ld, dx, n, hin, win, size, patch, t_pad, out_scale);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_clip_head(const void* f, int32_t ld_f, const float* wproj, const float* text, int32_t text_stride,
                            float scale, float grad_scale, float* sim, void* df, int32_t ld_df, int32_t n, int32_t h,
                            int32_t p, void* stream) {
  GD_REQUIRE(f && wproj && text && n > 0 && h > 0 && p > 0, "gd_clip_head: bad arguments");
  GD_REQUIRE(ld_f >= h && (df == nullptr || ld_df >= h), "gd_clip_head: bad strides");
  const size_t smem = sizeof(float) * (static_cast<size_t>(h) + p + 16);
  GD_REQUIRE(smem <= 48 * 1024, "gd_clip_head: h + p too large (%d + %d)", h, p);
  clip_head_kernel<<<n, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(f), ld_f, wproj, text, text_stride, scale, grad_scale, sim,
      reinterpret_cast<__half*>(df), ld_df, h, p);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
