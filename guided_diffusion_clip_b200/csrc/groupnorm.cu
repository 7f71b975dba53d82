// GroupNorm32 (32 groups, biased variance, fp32 math on fp16 NHWC storage) fused with SiLU, FiLM
// scale/shift and the 2x avg-pool / nearest-upsample of up/down ResBlocks, plus its data-gradient.
//
// Reference call sites: nn.py:17-19,93-100 (GroupNorm32), unet.py:184,208 (SiLU), unet.py:248-252 (FiLM:
// out_norm(h)*(1+scale)+shift, scale = first half of emb_out), unet.py:191-195,237-242 (h_upd),
// unet.py:285,302 (attention norm, no activation), unet.py:614-615 (out head).
//
// Bandwidth-bound: every kernel streams 16-byte (8 x fp16) vectors with the channel chunk fixed per thread
// (blockDim is a multiple of C/8) so per-channel affine terms live in registers; the statistics pass
// reduces per-CTA partials deterministically (no float atomics in global memory).
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

constexpr int kGroups = 32;
constexpr int kMaxChunks = 128;

struct Geo {
  int c8;        // C / 8
  int rep;       // pixel lanes per CTA
  int threads;   // c8 * rep
  int chunks;    // CTAs per image
  int px_per_chunk;
};

// CTAs per image: ~64 pixels per thread lane on big tensors (the per-CTA affine set-up is ~70 scalar instructions
// and a dozen dependent loads; at 16 pixels per lane the sustained, power-capped bandwidth was 10-25 % lower:
// profiles/gn_geo_probe.py), but never fewer CTAs than ~4 per SM in total (small 8x8 / 16x16 layers are
// latency-bound otherwise) as long as every lane still gets >= 2 pixels.
constexpr int kPxPerLane = 64;
Geo make_geo(int c, int hw, int n, int max_chunks) {
  Geo g;
  g.c8 = c / 8;
  g.rep = 256 / g.c8;
  if (g.rep < 1) g.rep = 1;
  g.threads = g.c8 * g.rep;
  int chunks = (hw + g.rep * kPxPerLane - 1) / (g.rep * kPxPerLane);
  const int want = (4 * 148 + n - 1) / n;
  if (chunks < want) chunks = want;
  const int max_by_work = hw / (g.rep * 2);
  if (chunks > max_by_work) chunks = max_by_work;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  g.px_per_chunk = (hw + chunks - 1) / chunks;
  g.chunks = (hw + g.px_per_chunk - 1) / g.px_per_chunk;
  return g;
}

// Deterministic block reduction of two per-thread [8]-channel partials into per-group sums:
// every thread parks its values in shared memory, then thread (which, g) adds its group's
// cpg channels x rep pixel lanes in a fixed order (no atomics -> bitwise reproducible statistics).
// s_part: dynamic smem [2][blockDim.x][8] floats; s_acc: [2][32] result.
__device__ __forceinline__ void block_group_reduce(float* s_part, float* s_acc, const float (&a)[8],
                                                   const float (&b)[8], int c8, int rep, int cpg) {
  const int T = blockDim.x, tid = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_part[tid * 8 + j] = a[j];
    s_part[(T + tid) * 8 + j] = b[j];
  }
  __syncthreads();
  if (tid < 2 * kGroups) {
    const int which = tid / kGroups, g = tid % kGroups;
    const float* src = s_part + static_cast<size_t>(which) * T * 8;
    float acc = 0.f;
    for (int ch = g * cpg; ch < (g + 1) * cpg; ++ch) {
      const int cch = ch >> 3, j = ch & 7;
      for (int pl = 0; pl < rep; ++pl) acc += src[(pl * c8 + cch) * 8 + j];
    }
    s_acc[tid] = acc;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// forward statistics: partial[n][chunk][32][2] = (sum x, sum x^2) over this CTA's pixels
// ---------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __half* __restrict__ x, int ld, int hw, int c, int c8, int rep, int px_per_chunk,
                                float* __restrict__ partial) {
  pdl_enter();
  __shared__ float s_acc[2 * kGroups];
  const int n = blockIdx.y, chunk = blockIdx.x;
  extern __shared__ float s_part[];
  __syncthreads();
  const int cch = threadIdx.x % c8, pl = threadIdx.x / c8;
  const int p0 = chunk * px_per_chunk;
  const int p1 = min(hw, p0 + px_per_chunk);
  const __half* base = x + static_cast<size_t>(n) * hw * ld + cch * 8;
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
  int p = p0 + pl;
  // 4 independent 16-byte loads in flight per thread
  for (; p + 3 * rep < p1; p += 4 * rep) {
    Half8 v0 = ld_half8(base + static_cast<size_t>(p) * ld);
    Half8 v1 = ld_half8(base + static_cast<size_t>(p + rep) * ld);
    Half8 v2 = ld_half8(base + static_cast<size_t>(p + 2 * rep) * ld);
    Half8 v3 = ld_half8(base + static_cast<size_t>(p + 3 * rep) * ld);
    float f[8];
    half8_to_float(v0, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
    half8_to_float(v1, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
    half8_to_float(v2, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
    half8_to_float(v3, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
  }
  for (; p < p1; p += rep) {
    float f[8];
    half8_to_float(ld_half8(base + static_cast<size_t>(p) * ld), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += f[j]; ss[j] += f[j] * f[j]; }
  }
  const int cpg = c / kGroups;
  block_group_reduce(s_part, s_acc, s, ss, c8, rep, cpg);
  __syncthreads();
  if (threadIdx.x < kGroups) {
    float* o = partial + ((static_cast<size_t>(n) * gridDim.x + chunk) * kGroups + threadIdx.x) * 2;
    o[0] = s_acc[threadIdx.x];
    o[1] = s_acc[kGroups + threadIdx.x];
  }
}

// mode 0: (mean, rstd); mode 1: (sum0/count, sum1/count)
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int chunks, float inv_count, float eps, int mode,
                                   float* __restrict__ out) {
  pdl_enter();
  const int n = blockIdx.x, g = threadIdx.x;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < chunks; ++k) {
    const float* pp = partial + ((static_cast<size_t>(n) * chunks + k) * kGroups + g) * 2;
    a += pp[0];
    b += pp[1];
  }
  a *= inv_count;
  b *= inv_count;
  float* o = out + (static_cast<size_t>(n) * kGroups + g) * 2;
  if (mode == 0) {
    double var = b - a * a;
    if (var < 0.0) var = 0.0;
    o[0] = static_cast<float>(a);
    o[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  } else {
    o[0] = static_cast<float>(a);
    o[1] = static_cast<float>(b);
  }
}

// per-thread affine: y = x * a + b for the thread's 8 channels (shared with the fused conv operand path)
__device__ __forceinline__ void load_affine(const float* mean_rstd, const float* gamma, const float* beta,
                                            const float* film, int film_ld, int n, int c, int ch0, float (&a)[8],
                                            float (&b)[8]) {
  gn_load_affine(mean_rstd, gamma, beta, film, film_ld, n, c, ch0, a, b);
}

template <bool kSilu>
__device__ __forceinline__ void norm_act(const Half8& v, const float (&a)[8], const float (&b)[8], float (&o)[8]) {
  float f[8];
  half8_to_float(v, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float z = fmaf(f[j], a[j], b[j]);
    o[j] = kSilu ? silu_f(z) : z;
  }
}

// ---------------------------------------------------------------------------------------------
// forward apply.  Iterates over INPUT pixels for SAME / UPSAMPLE2 and over OUTPUT pixels for AVGPOOL2.
// ---------------------------------------------------------------------------------------------
template <bool kSilu, int kMode>
__global__ void gn_apply_kernel(const __half* __restrict__ x, int ld, const float* __restrict__ mean_rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ film, int film_ld, __half* __restrict__ out, int ld_out,
                                __half* __restrict__ aux, int ld_aux, int h, int w, int c, int c8, int rep,
                                int px_per_chunk) {
  pdl_enter();
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int cch = threadIdx.x % c8, pl = threadIdx.x / c8;
  float a[8], b[8];
  load_affine(mean_rstd, gamma, beta, film, film_ld, n, c, cch * 8, a, b);
  const int hw_in = h * w;
  const __half* xin = x + static_cast<size_t>(n) * hw_in * ld + cch * 8;
  if (kMode == GD_GN_SAME) {
    __half* o = out + static_cast<size_t>(n) * hw_in * ld_out + cch * 8;
    const int p0 = chunk * px_per_chunk, p1 = min(hw_in, p0 + px_per_chunk);
    int p = p0 + pl;
    for (; p + 7 * rep < p1; p += 8 * rep) {  // 8 independent 16-byte streaming loads in flight per thread
      Half8 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ld_half8_stream(xin + static_cast<size_t>(p + u * rep) * ld);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float r[8];
        norm_act<kSilu>(v[u], a, b, r);
        st_half8(o + static_cast<size_t>(p + u * rep) * ld_out, float_to_half8(r));
      }
    }
    for (; p + 3 * rep < p1; p += 4 * rep) {
      Half8 v0 = ld_half8(xin + static_cast<size_t>(p) * ld);
      Half8 v1 = ld_half8(xin + static_cast<size_t>(p + rep) * ld);
      Half8 v2 = ld_half8(xin + static_cast<size_t>(p + 2 * rep) * ld);
      Half8 v3 = ld_half8(xin + static_cast<size_t>(p + 3 * rep) * ld);
      float r[8];
      norm_act<kSilu>(v0, a, b, r);
      st_half8(o + static_cast<size_t>(p) * ld_out, float_to_half8(r));
      norm_act<kSilu>(v1, a, b, r);
      st_half8(o + static_cast<size_t>(p + rep) * ld_out, float_to_half8(r));
      norm_act<kSilu>(v2, a, b, r);
      st_half8(o + static_cast<size_t>(p + 2 * rep) * ld_out, float_to_half8(r));
      norm_act<kSilu>(v3, a, b, r);
      st_half8(o + static_cast<size_t>(p + 3 * rep) * ld_out, float_to_half8(r));
    }
    for (; p < p1; p += rep) {
      float r[8];
      norm_act<kSilu>(ld_half8(xin + static_cast<size_t>(p) * ld), a, b, r);
      st_half8(o + static_cast<size_t>(p) * ld_out, float_to_half8(r));
    }
  } else if (kMode == GD_GN_AVGPOOL2) {
    const int ho = h / 2, wo = w / 2, hw_out = ho * wo;
    __half* o = out + static_cast<size_t>(n) * hw_out * ld_out + cch * 8;
    const int p0 = chunk * px_per_chunk, p1 = min(hw_out, p0 + px_per_chunk);
    for (int p = p0 + pl; p < p1; p += rep) {
      const int oy = p / wo, ox = p - oy * wo;
      const size_t i00 = static_cast<size_t>(2 * oy) * w + 2 * ox;
      Half8 v0 = ld_half8(xin + i00 * ld);
      Half8 v1 = ld_half8(xin + (i00 + 1) * ld);
      Half8 v2 = ld_half8(xin + (i00 + w) * ld);
      Half8 v3 = ld_half8(xin + (i00 + w + 1) * ld);
      float r[8], acc[8];
      norm_act<kSilu>(v0, a, b, acc);
      norm_act<kSilu>(v1, a, b, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += r[j];
      norm_act<kSilu>(v2, a, b, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += r[j];
      norm_act<kSilu>(v3, a, b, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = (acc[j] + r[j]) * 0.25f;
      st_half8(o + static_cast<size_t>(p) * ld_out, float_to_half8(acc));
      if (aux != nullptr) {
        // side output: avg-pool of the RAW input = x_upd(x) of a down ResBlock (unet.py:195,241), the identity
        // residual the block's second conv adds; costs one extra quarter-size write instead of a 4-way gather there
        float x0[8], x1[8], x2[8], x3[8];
        half8_to_float(v0, x0);
        half8_to_float(v1, x1);
        half8_to_float(v2, x2);
        half8_to_float(v3, x3);
#pragma unroll
        for (int j = 0; j < 8; ++j) x0[j] = 0.25f * (x0[j] + x1[j] + x2[j] + x3[j]);
        st_half8(aux + (static_cast<size_t>(n) * hw_out + p) * ld_aux + cch * 8, float_to_half8(x0));
      }
    }
  } else {  // GD_GN_UPSAMPLE2
    const int wo = w * 2;
    __half* o = out + static_cast<size_t>(n) * hw_in * 4 * ld_out + cch * 8;
    const int p0 = chunk * px_per_chunk, p1 = min(hw_in, p0 + px_per_chunk);
    for (int p = p0 + pl; p < p1; p += rep) {
      const int iy = p / w, ix = p - iy * w;
      float r[8];
      norm_act<kSilu>(ld_half8(xin + static_cast<size_t>(p) * ld), a, b, r);
      const Half8 hv = float_to_half8(r);
      const size_t o00 = static_cast<size_t>(2 * iy) * wo + 2 * ix;
      st_half8(o + o00 * ld_out, hv);
      st_half8(o + (o00 + 1) * ld_out, hv);
      st_half8(o + (o00 + wo) * ld_out, hv);
      st_half8(o + (o00 + wo + 1) * ld_out, hv);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  With xh = (x-mean)*rstd, z = xh*g' + b', y = act(z), g' = gamma*(1+scale):
//   dz = dy_in * act'(z);  dxh = dz * g';  dx = rstd * (dxh - mean_g(dxh) - xh * mean_g(dxh*xh))
// dy_in is dy pulled back through the spatial op (avg-pool: dy/4 at the parent; upsample: sum of 4).
// ---------------------------------------------------------------------------------------------
template <int kMode>
__device__ __forceinline__ void load_dy(const __half* dy, int ld_dy, int n, int h, int w, int p, int ch0, float (&d)[8]) {
  if (kMode == GD_GN_SAME) {
    half8_to_float(ld_half8(dy + (static_cast<size_t>(n) * h * w + p) * ld_dy + ch0), d);
  } else if (kMode == GD_GN_AVGPOOL2) {
    const int iy = p / w, ix = p - iy * w;
    const int wo = w / 2;
    const size_t q = static_cast<size_t>(n) * (h / 2) * wo + static_cast<size_t>(iy / 2) * wo + ix / 2;
    half8_to_float(ld_half8(dy + q * ld_dy + ch0), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] *= 0.25f;
  } else {
    const int iy = p / w, ix = p - iy * w;
    const int wo = w * 2;
    const size_t q = static_cast<size_t>(n) * h * 2 * wo + static_cast<size_t>(2 * iy) * wo + 2 * ix;
    float t[8];
    half8_to_float(ld_half8(dy + q * ld_dy + ch0), d);
    half8_to_float(ld_half8(dy + (q + 1) * ld_dy + ch0), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] += t[j];
    half8_to_float(ld_half8(dy + (q + wo) * ld_dy + ch0), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] += t[j];
    half8_to_float(ld_half8(dy + (q + wo + 1) * ld_dy + ch0), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] += t[j];
  }
}

// address of the dy row that pixel p of an h x w image pulls its gradient from (same resolution, or the parent pixel
// of a 2x2 average pool at half resolution: every parent row is read by its 4 children, so it stays in L1 / L2)
template <int kMode>
__device__ __forceinline__ const __half* dy_row(const __half* dy_n, int ld_dy, int w, int p) {
  if (kMode == GD_GN_SAME) return dy_n + static_cast<size_t>(p) * ld_dy;
  const int iy = p / w, ix = p - iy * w;
  return dy_n + (static_cast<size_t>(iy >> 1) * (w >> 1) + (ix >> 1)) * ld_dy;
}

template <bool kSilu>
__device__ __forceinline__ float act_grad(float z) {
  if (!kSilu) return 1.0f;
  const float s = sigmoid_f(z);
  return s * (1.0f + z * (1.0f - s));
}

// xa/xb: xh = x*xa + xb (normalised value); ga: g'; z = xh*ga + bz
struct BwdAffine {
  float xa[8], xb[8], ga[8], bz[8];
};
__device__ __forceinline__ void load_bwd_affine(const float* mean_rstd, const float* gamma, const float* beta,
                                                const float* film, int film_ld, int n, int c, int ch0, BwdAffine& A) {
  const int cpg = c / kGroups;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j;
    const int g = ch / cpg;
    const float mean = mean_rstd[(static_cast<size_t>(n) * kGroups + g) * 2];
    const float rstd = mean_rstd[(static_cast<size_t>(n) * kGroups + g) * 2 + 1];
    A.xa[j] = rstd;
    A.xb[j] = -mean * rstd;
    float ga = gamma[ch], be = beta[ch];
    if (film != nullptr) {
      const float sc = 1.0f + film[static_cast<size_t>(n) * film_ld + ch];
      const float sh = film[static_cast<size_t>(n) * film_ld + c + ch];
      ga *= sc;
      be = be * sc + sh;
    }
    A.ga[j] = ga;
    A.bz[j] = be;
  }
}

template <bool kSilu, int kMode>
__global__ void __launch_bounds__(256, 2) gn_bwd_stats_kernel(const __half* __restrict__ x, int ld, const float* __restrict__ mean_rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ film, int film_ld, const __half* __restrict__ dy, int ld_dy,
                                    int h, int w, int c, int c8, int rep, int px_per_chunk,
                                    float* __restrict__ partial) {
  pdl_enter();
  __shared__ float s_acc[2 * kGroups];
  const int n = blockIdx.y, chunk = blockIdx.x;
  extern __shared__ float s_part[];
  __syncthreads();
  const int cch = threadIdx.x % c8, pl = threadIdx.x / c8;
  BwdAffine A;
  load_bwd_affine(mean_rstd, gamma, beta, film, film_ld, n, c, cch * 8, A);
  const int hw = h * w;
  const __half* xin = x + static_cast<size_t>(n) * hw * ld + cch * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const int p0 = chunk * px_per_chunk, p1 = min(hw, p0 + px_per_chunk);
  auto accumulate = [&](const float (&xf)[8], const float (&d)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xf[j], A.xa[j], A.xb[j]);
      const float z = fmaf(xh, A.ga[j], A.bz[j]);
      const float dxh = d[j] * act_grad<kSilu>(z) * A.ga[j];
      s1[j] += dxh;
      s2[j] += dxh * xh;
    }
  };
  int p = p0 + pl;
  if (kMode == GD_GN_SAME || kMode == GD_GN_AVGPOOL2) {
    // 4 pixels per iteration: all eight 16-byte loads are issued before any arithmetic
    const int hw_dy = kMode == GD_GN_SAME ? hw : (h >> 1) * (w >> 1);
    const __half* dyn = dy + static_cast<size_t>(n) * hw_dy * ld_dy + cch * 8;
    const float dsc = kMode == GD_GN_SAME ? 1.0f : 0.25f;
    for (; p + 3 * rep < p1; p += 4 * rep) {
      Half8 xv[4], dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xv[u] = ld_half8_stream(xin + static_cast<size_t>(p + u * rep) * ld);
        const __half* dp = dy_row<kMode>(dyn, ld_dy, w, p + u * rep);
        dv[u] = kMode == GD_GN_SAME ? ld_half8_stream(dp) : ld_half8(dp);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xf[8], d[8];
        half8_to_float(xv[u], xf);
        half8_to_float(dv[u], d);
        if (kMode != GD_GN_SAME) {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= dsc;
        }
        accumulate(xf, d);
      }
    }
  }
  for (; p < p1; p += rep) {
    float xf[8], d[8];
    half8_to_float(ld_half8(xin + static_cast<size_t>(p) * ld), xf);
    load_dy<kMode>(dy, ld_dy, n, h, w, p, cch * 8, d);
    accumulate(xf, d);
  }
  const int cpg = c / kGroups;
  block_group_reduce(s_part, s_acc, s1, s2, c8, rep, cpg);
  __syncthreads();
  if (threadIdx.x < kGroups) {
    float* o = partial + ((static_cast<size_t>(n) * gridDim.x + chunk) * kGroups + threadIdx.x) * 2;
    o[0] = s_acc[threadIdx.x];
    o[1] = s_acc[kGroups + threadIdx.x];
  }
}

template <bool kSilu, int kMode>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_kernel(const __half* __restrict__ x, int ld, const float* __restrict__ mean_rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ film, int film_ld, const __half* __restrict__ dy, int ld_dy,
                                    const float* __restrict__ gsum, const __half* __restrict__ add, int ld_add,
                                    int add_mode, __half* __restrict__ dx, int ld_dx, int h, int w, int c, int c8,
                                    int rep, int px_per_chunk) {
  pdl_enter();
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int cch = threadIdx.x % c8, pl = threadIdx.x / c8;
  BwdAffine A;
  load_bwd_affine(mean_rstd, gamma, beta, film, film_ld, n, c, cch * 8, A);
  const int cpg = c / kGroups;
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (cch * 8 + j) / cpg;
    m1[j] = gsum[(static_cast<size_t>(n) * kGroups + g) * 2];
    m2[j] = gsum[(static_cast<size_t>(n) * kGroups + g) * 2 + 1];
  }
  const int hw = h * w;
  const __half* xin = x + static_cast<size_t>(n) * hw * ld + cch * 8;
  const int p0 = chunk * px_per_chunk, p1 = min(hw, p0 + px_per_chunk);
  auto grad_x = [&](const float (&xf)[8], const float (&d)[8], float (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = fmaf(xf[j], A.xa[j], A.xb[j]);
      const float z = fmaf(xh, A.ga[j], A.bz[j]);
      const float dxh = d[j] * act_grad<kSilu>(z) * A.ga[j];
      r[j] = A.xa[j] * (dxh - m1[j] - xh * m2[j]);
    }
  };
  int p = p0 + pl;
  if (kMode == GD_GN_SAME || kMode == GD_GN_AVGPOOL2) {
    // 4 pixels per iteration with every load (x, dy, add) issued up front
    const int hw_dy = kMode == GD_GN_SAME ? hw : (h >> 1) * (w >> 1);
    const __half* dyn = dy + static_cast<size_t>(n) * hw_dy * ld_dy + cch * 8;
    const float dsc = kMode == GD_GN_SAME ? 1.0f : 0.25f;
    const bool add_same = add_mode == GD_GN_SAME;
    const __half* addn = add != nullptr
        ? add + static_cast<size_t>(n) * (add_same ? hw : (h >> 1) * (w >> 1)) * ld_add + cch * 8 : nullptr;
    const float asc = add_same ? 1.0f : 0.25f;
    __half* dxn = dx + static_cast<size_t>(n) * hw * ld_dx + cch * 8;
    for (; p + 3 * rep < p1; p += 4 * rep) {
      Half8 xv[4], dv[4], av[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p + u * rep;
        xv[u] = ld_half8_stream(xin + static_cast<size_t>(pp) * ld);
        const __half* dp = dy_row<kMode>(dyn, ld_dy, w, pp);
        dv[u] = kMode == GD_GN_SAME ? ld_half8_stream(dp) : ld_half8(dp);
        if (addn != nullptr)
          av[u] = add_same ? ld_half8_stream(addn + static_cast<size_t>(pp) * ld_add)
                           : ld_half8(dy_row<GD_GN_AVGPOOL2>(addn, ld_add, w, pp));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xf[8], d[8], r[8];
        half8_to_float(xv[u], xf);
        half8_to_float(dv[u], d);
        if (kMode != GD_GN_SAME) {
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] *= dsc;
        }
        grad_x(xf, d, r);
        if (addn != nullptr) {
          float t[8];
          half8_to_float(av[u], t);
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = fmaf(asc, t[j], r[j]);
        }
        st_half8(dxn + static_cast<size_t>(p + u * rep) * ld_dx, float_to_half8(r));
      }
    }
  }
  for (; p < p1; p += rep) {
    float xf[8], d[8], r[8];
    half8_to_float(ld_half8(xin + static_cast<size_t>(p) * ld), xf);
    load_dy<kMode>(dy, ld_dy, n, h, w, p, cch * 8, d);
    grad_x(xf, d, r);
    const size_t off = (static_cast<size_t>(n) * hw + p);
    if (add != nullptr) {
      float t[8];
      if (add_mode == GD_GN_SAME) {
        half8_to_float(ld_half8(add + off * ld_add + cch * 8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += t[j];
      } else {  // gradient of the avg-pooled copy of x, stored at half resolution
        const int iy = p / w, ix = p - iy * w;
        const size_t q = (static_cast<size_t>(n) * (h / 2) + iy / 2) * (w / 2) + ix / 2;
        half8_to_float(ld_half8(add + q * ld_add + cch * 8), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += 0.25f * t[j];
      }
    }
    st_half8(dx + off * ld_dx + cch * 8, float_to_half8(r));
  }
}

// ---------------------------------------------------------------------------------------------
// mean / rstd from the partials a producing conv's epilogue emitted (gd_conv_desc.stats_out): per 32-pixel row
// block and 4-channel chunk (sum, sum of squares).  The normalised tensor may be the concatenation of two conv
// outputs (c0 channels from p0, then c1 from p1); a group may straddle the boundary.
// One warp per (group, image); fixed-order reduction in fp64 -> bitwise reproducible.
// ---------------------------------------------------------------------------------------------
// kW warps share one (group, image): kW = 1 for layers with few partial rows (one warp each, four per CTA, no block
// barrier), kW = 4 for the full-resolution layers (512 / 128 rows per image) so that every lane still has only a
// handful of independent loads in flight.
template <int kW>
__global__ void __launch_bounds__(128)
gn_finalize_partials_kernel(const float* __restrict__ p0, int c0, int ld0, const float* __restrict__ p1, int c1,
                            int ld1, int rows_per_image, int n_img, double inv_count, float eps, float* __restrict__ out,
                            const float* __restrict__ gamma, const float* __restrict__ beta,
                            const float* __restrict__ film, int film_ld, float* __restrict__ coef) {
  pdl_enter();
  __shared__ double sh[2][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = kW == 1 ? blockIdx.x * 4 + warp : blockIdx.x;
  const bool active = item < kGroups * n_img;
  const int g = item % kGroups, n = item / kGroups;
  const int cpg = (c0 + c1) / kGroups;
  const int ch_lo = g * cpg, ch_hi = ch_lo + cpg;
  double s = 0.0, ss = 0.0;
  if (active) {
#pragma unroll 4
    for (int r = (kW == 1 ? lane : threadIdx.x); r < rows_per_image; r += 32 * kW) {
      const size_t row = static_cast<size_t>(n) * rows_per_image + r;
      for (int ch = ch_lo; ch < ch_hi; ch += 4) {
        const float* src = ch < c0 ? p0 + (row * ld0 + (ch >> 2)) * 2 : p1 + (row * ld1 + ((ch - c0) >> 2)) * 2;
        const float2 v = __ldcg(reinterpret_cast<const float2*>(src));
        s += v.x;
        ss += v.y;
      }
    }
  }
  // fixed-order fp64 reduction -> bitwise reproducible
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (kW == 4) {
    if (lane == 0) {
      sh[0][warp] = s;
      sh[1][warp] = ss;
    }
    __syncthreads();
    s = (sh[0][0] + sh[0][1]) + (sh[0][2] + sh[0][3]);
    ss = (sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3]);
    if (warp != 0) return;
  }
  if (active) {
    // (every lane holds the full sums after the butterfly)
    const double mean = s * inv_count;
    double var = ss * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mean_f = static_cast<float>(mean);
    const float rstd_f = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (lane == 0) {
      float* o = out + (static_cast<size_t>(n) * kGroups + g) * 2;
      o[0] = mean_f;
      o[1] = rstd_f;
    }
    if (coef != nullptr) {
      // optional: the per-channel affine of this group, consumed by a conv that normalises its operand on the fly
      const int c = c0 + c1;
      for (int ch = ch_lo + lane; ch < ch_hi; ch += 32) {
        float a, b;
        const bool has_film = film != nullptr;
        gn_affine(mean_f, rstd_f, gamma[ch], beta[ch], has_film,
                  has_film ? film[static_cast<size_t>(n) * film_ld + ch] : 0.f,
                  has_film ? film[static_cast<size_t>(n) * film_ld + c + ch] : 0.f, a, b);
        float* o = coef + (static_cast<size_t>(n) * (c >> 3) + (ch >> 3)) * 16 + (ch & 7);
        o[0] = a;
        o[8] = b;
      }
    }
  }
}

// coef[n][c/8][16] = (a[8], b[8]) per 8-channel chunk from finished statistics (the path without fused partials)
__global__ void gn_coef_kernel(const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ film, int film_ld, int n_img,
                               int c, float* __restrict__ coef) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img * c) return;
  const int n = i / c, ch = i - n * c;
  const int g = ch / (c / kGroups);
  float a, b;
  const bool has_film = film != nullptr;
  gn_affine(mean_rstd[(static_cast<size_t>(n) * kGroups + g) * 2], mean_rstd[(static_cast<size_t>(n) * kGroups + g) * 2 + 1],
            gamma[ch], beta[ch], has_film, has_film ? film[static_cast<size_t>(n) * film_ld + ch] : 0.f,
            has_film ? film[static_cast<size_t>(n) * film_ld + c + ch] : 0.f, a, b);
  float* o = coef + (static_cast<size_t>(n) * (c >> 3) + (ch >> 3)) * 16 + (ch & 7);
  o[0] = a;
  o[8] = b;
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int gd_groupnorm_finalize_partials(const float* p0, int32_t c0, int32_t ld0, const float* p1, int32_t c1,
                                              int32_t ld1, int32_t rows_per_image, int32_t n, int32_t hw, float eps,
                                              float* mean_rstd, const float* gamma, const float* beta, const float* film,
                                              int32_t film_ld, float* coef_out, void* stream) {
  GD_REQUIRE(p0 && mean_rstd && n > 0 && hw > 0 && rows_per_image > 0, "gd_groupnorm_finalize_partials: bad arguments");
  if (coef_out != nullptr)
    GD_REQUIRE(gamma && beta && (film == nullptr || film_ld >= 2 * (c0 + (p1 ? c1 : 0))),
               "gd_groupnorm_finalize_partials: coef_out needs gamma, beta and (if FiLM) film_ld >= 2*C");
  if (p1 == nullptr) c1 = 0;
  const int c = c0 + c1;
  GD_REQUIRE(c0 > 0 && c0 % 4 == 0 && c1 % 4 == 0 && c % 32 == 0 && (c / 32) % 4 == 0,
             "gd_groupnorm_finalize_partials: needs channels-per-group %% 4 == 0 (c0=%d c1=%d)", c0, c1);
  GD_REQUIRE(ld0 >= c0 / 4 && (p1 == nullptr || ld1 >= c1 / 4), "gd_groupnorm_finalize_partials: bad leading dimension");
  const double inv_count = 1.0 / (static_cast<double>(hw) * static_cast<double>(c / kGroups));
  if (rows_per_image >= 128)
    GD_CHECK_CUDA(launch_pdl_small(gn_finalize_partials_kernel<4>, dim3(kGroups * n), dim3(128), 0,
                             reinterpret_cast<cudaStream_t>(stream), p0, c0, ld0, p1, c1, ld1, rows_per_image, n,
                             inv_count, eps, mean_rstd, gamma, beta, film, film_ld, coef_out));
  else
    GD_CHECK_CUDA(launch_pdl_small(gn_finalize_partials_kernel<1>, dim3((kGroups * n + 3) / 4), dim3(128), 0,
                             reinterpret_cast<cudaStream_t>(stream), p0, c0, ld0, p1, c1, ld1, rows_per_image, n,
                             inv_count, eps, mean_rstd, gamma, beta, film, film_ld, coef_out));
  count_launch(1);
  return 0;
}

extern "C" int gd_groupnorm_coef(const float* mean_rstd, const float* gamma, const float* beta, const float* film,
                                 int32_t film_ld, int32_t n, int32_t c, float* coef_out, void* stream) {
  GD_REQUIRE(mean_rstd && gamma && beta && coef_out && n > 0, "gd_groupnorm_coef: bad arguments");
  GD_REQUIRE(c > 0 && c % 32 == 0 && (film == nullptr || film_ld >= 2 * c), "gd_groupnorm_coef: bad channel count %d / film stride %d",
             c, film_ld);
  GD_CHECK_CUDA(launch_pdl_small(gn_coef_kernel, dim3((n * c + 255) / 256), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                           mean_rstd, gamma, beta, film, film_ld, n, c, coef_out));
  count_launch(1);
  return 0;
}

extern "C" int64_t gd_groupnorm_ws_floats(int32_t n, int32_t hw, int32_t c) {
  (void)hw;
  (void)c;
  return static_cast<int64_t>(n) * (kMaxChunks + 1) * kGroups * 2;
}

static int check_gn_common(const char* who, const void* x, int ld, int n, int hw, int c) {
  GD_REQUIRE(x != nullptr, "%s: null input", who);
  GD_REQUIRE(n > 0 && hw > 0, "%s: bad n/hw", who);
  GD_REQUIRE(c > 0 && c % 32 == 0 && c <= 2048, "%s: channels must be a multiple of 32 (GroupNorm32), got %d", who, c);
  GD_REQUIRE(ld >= c && ld % 8 == 0, "%s: bad ld %d for c %d", who, ld, c);
  return 0;
}

extern "C" int gd_groupnorm_stats(const void* x, int32_t ld, int32_t n, int32_t hw, int32_t c, float eps,
                                  float* partial_ws, float* mean_rstd, void* stream) {
  if (int rc = check_gn_common("gd_groupnorm_stats", x, ld, n, hw, c)) return rc;
  GD_REQUIRE(partial_ws != nullptr && mean_rstd != nullptr, "gd_groupnorm_stats: null workspace/output");
  const Geo g = make_geo(c, hw, n, kMaxChunks);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GD_CHECK_CUDA(launch_pdl(gn_stats_kernel, dim3(g.chunks, n), dim3(g.threads), g.threads * 16 * sizeof(float), st,
                           reinterpret_cast<const __half*>(x), ld, hw, c, g.c8, g.rep, g.px_per_chunk, partial_ws));
  const float inv_count = 1.0f / (static_cast<float>(hw) * static_cast<float>(c / kGroups));
  GD_CHECK_CUDA(launch_pdl_small(gn_finalize_kernel, dim3(n), dim3(kGroups), 0, st, partial_ws, g.chunks, inv_count, eps, 0,
                           mean_rstd));
  count_launch(2);
  return 0;
}

#define GD_GN_DISPATCH(KERNEL, silu, mode, ...)                                                                    \
  do {                                                                                                             \
    cudaError_t _le;                                                                                               \
    if (silu) {                                                                                                    \
      if (mode == GD_GN_SAME) _le = launch_pdl(KERNEL<true, GD_GN_SAME>, __VA_ARGS__);                             \
      else if (mode == GD_GN_AVGPOOL2) _le = launch_pdl(KERNEL<true, GD_GN_AVGPOOL2>, __VA_ARGS__);                \
      else _le = launch_pdl(KERNEL<true, GD_GN_UPSAMPLE2>, __VA_ARGS__);                                           \
    } else {                                                                                                       \
      if (mode == GD_GN_SAME) _le = launch_pdl(KERNEL<false, GD_GN_SAME>, __VA_ARGS__);                            \
      else if (mode == GD_GN_AVGPOOL2) _le = launch_pdl(KERNEL<false, GD_GN_AVGPOOL2>, __VA_ARGS__);               \
      else _le = launch_pdl(KERNEL<false, GD_GN_UPSAMPLE2>, __VA_ARGS__);                                          \
    }                                                                                                              \
    GD_CHECK_CUDA(_le);                                                                                            \
  } while (0)

extern "C" int gd_groupnorm_apply(const void* x, int32_t ld, const float* mean_rstd, const float* gamma,
                                  const float* beta, const float* film, int32_t film_ld, void* out, int32_t ld_out,
                                  int32_t n, int32_t h, int32_t w, int32_t c, int32_t silu, int32_t spatial_mode,
                                  void* aux_out, int32_t ld_aux, void* stream) {
  if (int rc = check_gn_common("gd_groupnorm_apply", x, ld, n, h * w, c)) return rc;
  GD_REQUIRE(mean_rstd && gamma && beta && out, "gd_groupnorm_apply: null pointer");
  GD_REQUIRE(ld_out >= c && ld_out % 8 == 0, "gd_groupnorm_apply: bad ld_out %d", ld_out);
  GD_REQUIRE(spatial_mode >= GD_GN_SAME && spatial_mode <= GD_GN_UPSAMPLE2, "gd_groupnorm_apply: bad spatial mode");
  if (spatial_mode == GD_GN_AVGPOOL2) GD_REQUIRE(h % 2 == 0 && w % 2 == 0, "gd_groupnorm_apply: avgpool needs even h,w");
  if (film) GD_REQUIRE(film_ld >= 2 * c, "gd_groupnorm_apply: film_ld %d < 2*c", film_ld);
  if (aux_out)
    GD_REQUIRE(spatial_mode == GD_GN_AVGPOOL2 && ld_aux >= c && ld_aux % 8 == 0,
               "gd_groupnorm_apply: aux output only exists for the avg-pool mode (ld_aux %d)", ld_aux);
  const int hw_iter = spatial_mode == GD_GN_AVGPOOL2 ? (h / 2) * (w / 2) : h * w;
  const Geo g = make_geo(c, hw_iter, n, 4096);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GD_GN_DISPATCH(gn_apply_kernel, silu, spatial_mode,
                 dim3(g.chunks, n), dim3(g.threads), 0, st, reinterpret_cast<const __half*>(x), ld, mean_rstd, gamma, beta,
                 film, film_ld, reinterpret_cast<__half*>(out), ld_out, reinterpret_cast<__half*>(aux_out), ld_aux, h, w,
                 c, g.c8, g.rep, g.px_per_chunk);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int gd_groupnorm_bwd(const void* x, int32_t ld, const float* mean_rstd, const float* gamma, const float* beta,
                                const float* film, int32_t film_ld, const void* dy, int32_t ld_dy, const void* add,
                                int32_t ld_add, int32_t add_mode, void* dx, int32_t ld_dx, float* partial_ws, int32_t n,
                                int32_t h, int32_t w, int32_t c, int32_t silu, int32_t spatial_mode, void* stream) {
  if (int rc = check_gn_common("gd_groupnorm_bwd", x, ld, n, h * w, c)) return rc;
  GD_REQUIRE(mean_rstd && gamma && beta && dy && dx && partial_ws, "gd_groupnorm_bwd: null pointer");
  GD_REQUIRE(ld_dy % 8 == 0 && ld_dx % 8 == 0 && (add == nullptr || ld_add % 8 == 0), "gd_groupnorm_bwd: bad strides");
  GD_REQUIRE(spatial_mode >= GD_GN_SAME && spatial_mode <= GD_GN_UPSAMPLE2, "gd_groupnorm_bwd: bad spatial mode");
  GD_REQUIRE(add_mode == GD_GN_SAME || add_mode == GD_GN_AVGPOOL2, "gd_groupnorm_bwd: bad add_mode");
  if (spatial_mode == GD_GN_AVGPOOL2) GD_REQUIRE(h % 2 == 0 && w % 2 == 0, "gd_groupnorm_bwd: avgpool needs even h,w");
  const Geo g = make_geo(c, h * w, n, kMaxChunks);   // statistics pass (partials buffer bounds the CTA count)
  const Geo ga = make_geo(c, h * w, n, 4096);        // apply pass
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // partial_ws layout: [n][kMaxChunks][32][2] partials, then [n][32][2] group means
  float* gsum = partial_ws + static_cast<size_t>(n) * kMaxChunks * kGroups * 2;
  GD_GN_DISPATCH(gn_bwd_stats_kernel, silu, spatial_mode,
                 dim3(g.chunks, n), dim3(g.threads), g.threads * 16 * sizeof(float), st,
                 reinterpret_cast<const __half*>(x), ld, mean_rstd, gamma, beta, film, film_ld,
                 reinterpret_cast<const __half*>(dy), ld_dy, h, w, c, g.c8, g.rep, g.px_per_chunk, partial_ws);
  GD_CHECK_CUDA(cudaGetLastError());
  const float inv_count = 1.0f / (static_cast<float>(h * w) * static_cast<float>(c / kGroups));
  GD_CHECK_CUDA(launch_pdl_small(gn_finalize_kernel, dim3(n), dim3(kGroups), 0, st, partial_ws, g.chunks, inv_count, 0.f, 1,
                           gsum));
  GD_GN_DISPATCH(gn_bwd_apply_kernel, silu, spatial_mode,
                 dim3(ga.chunks, n), dim3(ga.threads), 0, st, reinterpret_cast<const __half*>(x), ld, mean_rstd, gamma, beta,
                 film, film_ld, reinterpret_cast<const __half*>(dy), ld_dy, gsum, reinterpret_cast<const __half*>(add),
                 ld_add, add_mode, reinterpret_cast<__half*>(dx), ld_dx, h, w, c, ga.c8, ga.rep, ga.px_per_chunk);
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(3);
  return 0;
}
