// First layer of the UNet / classifier / upsampler: conv_nd(2, C_in, C, 3, padding=1) with C_in = 3 or 6
// (input_blocks.0.0, unet.py:483, 741) straight from the fp32 NCHW network input to the fp16 NHWC activation, with
// the GroupNorm partial statistics of the stored output (same format as gd_conv_igemm's stats_out).
//
// With K = 9 * C_in = 27 (54) the layer does 27 MACs per output: it is bound by WRITING its output (512 B per pixel
// at 256 channels), not by arithmetic.  On the tcgen05 kernel the same layer (im2col to a 64-wide K block + GEMM) is
// bound by reading a 128 x 256 fp32 accumulator tile out of TMEM (64 B/clk/SM) plus an epilogue that cannot overlap
// with a 4-MMA mainloop: 1.1 ms + 0.3 ms im2col at batch 64.  Here the accumulators never leave registers:
// warp-level mma.sync m16n8k16 (the tensor-pipe choice is irrelevant at 27 MACs per output), A fragments gathered
// directly from the input image (L1-resident: every input value is reused by 9 taps x C outputs), weights in shared
// memory, output staged per warp and stored with coalesced 16-byte rows.
//
// CTA = 4 warps = 128 consecutive pixels of one image (h*w % 128 == 0), warp = 32 pixels, C_out in passes of 64.
#include "common.cuh"
#include "../../include/gd_b200.h"

namespace gd {
void count_launch(int n = 1);
namespace {

constexpr int kCiThreads = 128;

__device__ __forceinline__ void ci_ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void ci_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ci_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// KS = number of 16-wide k-steps that cover 9 * cin (2 for cin <= 3, 4 for cin <= 7)
template <int KS>
__global__ void __launch_bounds__(kCiThreads, KS == 2 ? 4 : 2)
conv_in3x3_kernel(const float* __restrict__ x, const __half* __restrict__ wpack, const float* __restrict__ bias,
                  __half* __restrict__ out, int ld_out, float* __restrict__ stats, int rows_per_image, int n_img, int cin,
                  int h, int w, int cout) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int kRowB = KS * 32 + 16;  // weight row: KS*16 halves + 16 B pad -> the 8 rows of an ldmatrix phase hit 8 bank groups
  uint8_t* sW = smem;
  uint8_t* sStage = smem + cout * kRowB;                            // 4 warps x [32 pixels][64 channels] fp16
  float* sStat = reinterpret_cast<float*>(sStage + 4 * 4096);      // [4 warps][16 chunks][2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  pdl_enter();

  for (int i = tid; i < cout * KS * 2; i += kCiThreads) {  // wpack rows are 64 halves (k zero padded)
    const int row = i / (KS * 2), ch = i - row * (KS * 2);
    *reinterpret_cast<uint4*>(sW + row * kRowB + ch * 16) = __ldg(reinterpret_cast<const uint4*>(wpack + row * 64) + ch);
  }
  __syncthreads();

  // the 4 * KS im2col columns this thread supplies: k = 16 ks + 8 kk + 2 t + e  ->  (tap, ci) = (k / cin, k % cin)
  const int hw = h * w;
  int k_off[KS * 4], k_dy[KS * 4], k_dx[KS * 4];
#pragma unroll
  for (int i = 0; i < KS * 4; ++i) {
    const int k = 16 * (i >> 2) + 8 * ((i >> 1) & 1) + 2 * t + (i & 1);
    if (k < 9 * cin) {
      const int tap = k / cin, ci = k - tap * cin;
      k_dy[i] = tap / 3 - 1;
      k_dx[i] = tap % 3 - 1;
      k_off[i] = ci * hw + k_dy[i] * w + k_dx[i];
    } else {
      k_dy[i] = 1 << 20;  // never in bounds
      k_dx[i] = 0;
      k_off[i] = 0;
    }
  }

  const int tiles_per_image = hw / 128;
  const int tiles = n_img * tiles_per_image;
  const uint32_t sW_u32 = smem_u32(sW);
  const uint32_t stage_u32 = smem_u32(sStage) + static_cast<uint32_t>(warp) * 4096u;
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int img = tile / tiles_per_image;
    const int rem0 = (tile - img * tiles_per_image) * 128 + warp * 32;  // this warp's first pixel inside the image
    const float* xb = x + static_cast<size_t>(img) * cin * hw;
    // A fragments: rows = pixels (g, g+8 of the two 16-pixel m-tiles), columns = im2col k
    uint32_t a[2][KS][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int rem = rem0 + 16 * m + 8 * r + g;
        const int y = rem / w, xx = rem - y * w;
        const float* xp = xb + rem;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = ks * 4 + kk * 2 + e;
              const int yy = y + k_dy[i], xc = xx + k_dx[i];
              v[e] = (yy >= 0 && yy < h && xc >= 0 && xc < w) ? __ldg(xp + k_off[i]) : 0.f;
            }
            a[m][ks][r + 2 * kk] = ci_pack(v[0], v[1]);
          }
      }

    for (int pass = 0; pass < cout / 64; ++pass) {
      float acc[2][8][4];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-channel n-tiles
          uint32_t bf[4];
          const int row = pass * 64 + jp * 16 + (lane & 7) + 8 * (lane >> 4);
          const int chunk = ks * 2 + ((lane >> 3) & 1);
          ci_ldmatrix_x4(bf, sW_u32 + static_cast<uint32_t>(row * kRowB + chunk * 16));
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            ci_mma16816(acc[m][2 * jp], a[m][ks], bf[0], bf[1]);
            ci_mma16816(acc[m][2 * jp + 1], a[m][ks], bf[2], bf[3]);
          }
        }
      // epilogue: + bias -> fp16 -> per-warp staging (16-byte chunks XOR-swizzled by the row) ; statistics of the
      // ROUNDED values (what the consumer GroupNorm normalises)
      float s1[8], s2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(bias + pass * 64 + 8 * j + 2 * t));
        s1[j] = s2[j] = 0.f;
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const __half2 hv = __floats2half2_rn(acc[m][j][2 * r] + b.x, acc[m][j][2 * r + 1] + b.y);
            const int row = 16 * m + 8 * r + g;
            const uint32_t addr = stage_u32 + static_cast<uint32_t>(row * 128 + ((j ^ (row & 7)) << 4) + t * 4);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&hv)) : "memory");
            const float2 f = __half22float2(hv);
            s1[j] += f.x + f.y;
            s2[j] = fmaf(f.x, f.x, fmaf(f.y, f.y, s2[j]));
          }
      }
      if (stats != nullptr) {
        // 4-channel chunk = columns {2t, 2t+1} of lanes t = 0,1 (chunk 2j) or t = 2,3 (chunk 2j+1); then over the 8 row lanes
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int o = 1; o <= 16; o = (o == 1 ? 4 : o * 2)) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
          }
          if (g == 0 && (t & 1) == 0) {
            float* d = sStat + (warp * 16 + 2 * j + (t >> 1)) * 2;
            d[0] = s1[j];
            d[1] = s2[j];
          }
        }
      }
      __syncwarp();
      __half* o_base = out + (static_cast<size_t>(img) * hw + rem0) * ld_out + pass * 64;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), ch = lane & 7;
        uint4 v;
        const uint32_t addr = stage_u32 + static_cast<uint32_t>(rr * 128 + ((ch ^ (rr & 7)) << 4));
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
        *reinterpret_cast<uint4*>(o_base + static_cast<size_t>(rr) * ld_out + ch * 8) = v;
      }
      __syncwarp();
      if (stats != nullptr) {  // one partial row per 128-pixel tile: sum the 4 warps
        __syncthreads();
        if (tid < 32) {
          const int c = tid >> 1, q = tid & 1;
          const float v = (sStat[(0 * 16 + c) * 2 + q] + sStat[(1 * 16 + c) * 2 + q]) +
                          (sStat[(2 * 16 + c) * 2 + q] + sStat[(3 * 16 + c) * 2 + q]);
          const size_t row = static_cast<size_t>(img) * rows_per_image + (tile - img * tiles_per_image);
          stats[(row * (cout / 4) + pass * 16 + c) * 2 + q] = v;
        }
        __syncthreads();
      }
    }
  }
}

}  // namespace
}  // namespace gd

using namespace gd;

extern "C" int gd_conv_in3x3(const gd_conv_in_desc* d, void* stream) {
  GD_REQUIRE(d != nullptr && d->x && d->wpack && d->bias && d->out, "gd_conv_in3x3: null pointer");
  GD_REQUIRE(d->n > 0 && d->cin >= 1 && d->cin <= 7, "gd_conv_in3x3: cin %d outside [1,7] (9*cin must fit 64)", d->cin);
  GD_REQUIRE(d->cout > 0 && d->cout % 64 == 0 && d->cout <= 1024, "gd_conv_in3x3: cout %d must be a multiple of 64 (<= 1024)",
             d->cout);
  GD_REQUIRE(d->h > 0 && d->w > 0 && (d->h * d->w) % 128 == 0, "gd_conv_in3x3: h*w = %d must be a multiple of 128",
             d->h * d->w);
  GD_REQUIRE(d->ld_out >= d->cout && d->ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(d->out) & 15u) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->wpack) & 15u) == 0,
             "gd_conv_in3x3: output rows / packed weights must be 16-byte aligned (ld_out %d)", d->ld_out);
  int rpi = 0;
  if (d->stats_out != nullptr) {
    const int64_t rows = gd_conv_stats_rows(d->n, d->h, d->w, &rpi);
    GD_REQUIRE(rows > 0 && rpi >= d->h * d->w / 128,
               "gd_conv_in3x3: fused statistics need the gd_conv_stats_rows geometry with >= h*w/128 rows per image");
  }
  const int ks = d->cin * 9 <= 32 ? 2 : 4;
  const int smem_bytes = d->cout * (ks * 32 + 16) + 4 * 4096 + 4 * 16 * 2 * 4;
  const int tiles = d->n * (d->h * d->w / 128);
  const int grid = tiles < 148 * 4 ? tiles : 148 * 4;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half* wp = reinterpret_cast<const __half*>(d->wpack);
  __half* op = reinterpret_cast<__half*>(d->out);
  if (ks == 2) {
    static unsigned long long configured_on[2] = {0, 0};
    if (gd::first_use_on_device(configured_on)) {
      GD_CHECK_CUDA(cudaFuncSetAttribute(conv_in3x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    GD_CHECK_CUDA(launch_pdl(conv_in3x3_kernel<2>, dim3(grid), dim3(kCiThreads), smem_bytes, st, d->x, wp, d->bias, op,
                             d->ld_out, d->stats_out, rpi, d->n, d->cin, d->h, d->w, d->cout));
  } else {
    static unsigned long long configured_on[2] = {0, 0};
    if (gd::first_use_on_device(configured_on)) {
      GD_CHECK_CUDA(cudaFuncSetAttribute(conv_in3x3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
    }
    GD_CHECK_CUDA(launch_pdl(conv_in3x3_kernel<4>, dim3(grid), dim3(kCiThreads), smem_bytes, st, d->x, wp, d->bias, op,
                             d->ld_out, d->stats_out, rpi, d->n, d->cin, d->h, d->w, d->cout));
  }
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
