// STATIC instruction-budget probe for the round-2 plan "GroupNorm-apply inside the conv's operand path" (DESIGN.md 7.1):
// the in-place shared-memory transform of one 64-channel x 160-pixel halo slot (a*x + b -> SiLU, border rows zeroed,
// 128-byte swizzle preserved) by four warps.  Not part of the library; compiled and read as SASS only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cubin -o /tmp/xform.cubin profiles/exp_gn_transform_static.cu
//   cuobjdump -sass /tmp/xform.cubin
// Result (CUDA 12.9): the loop body handles two 16-byte chunks in ~120 instructions (32 FFMA, 16 FMUL, 16 MUFU.TANH,
// 16 HADD2 + 8 F2FP conversions, 2 LDS.128 + 2 STS.128, ~30 integer / predicate) = ~60 per chunk -> 1 280 chunks per
// slot = ~2 400 warp-instructions and 10 240 MUFU operations (640 cycles of the 16/clk pipe) per slot, against the
// 1 536 tensor-pipe cycles the slot's 12 MMAs occupy: ~40 % of the issue slots of each of the 4 schedulers.
#include <cuda_fp16.h>
#include <stdint.h>
__device__ __forceinline__ float tanh_approx(float v) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float silu_f(float v) { const float h = 0.5f * v; return fmaf(h, tanh_approx(h), h); }
// one 64-channel x 160-pixel halo slot (20 KB), 128 threads: thread -> (pixel row p = i / 8, 16-byte chunk q = i % 8)
// a/b: per-channel scale/shift of this image (64 floats each) staged in shared memory by the caller
__global__ void __launch_bounds__(128) xform(uint8_t* gbuf, const float* ga, const float* gb, int y0, int x0, int h, int w, int bw) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sa = reinterpret_cast<float*>(smem + 20480);
  float* sb = sa + 64;
  for (int i = threadIdx.x; i < 1280; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<uint4*>(gbuf)[i];
  if (threadIdx.x < 64) { sa[threadIdx.x] = ga[threadIdx.x]; sb[threadIdx.x] = gb[threadIdx.x]; }
  __syncthreads();
  const int q = threadIdx.x & 7;           // logical chunk = 8 channels
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = sa[q * 8 + j]; b[j] = sb[q * 8 + j]; }
#pragma unroll 2
  for (int p = threadIdx.x >> 3; p < 160; p += 16) {
    const int yy = y0 + p / bw, xx = x0 + p % bw;
    const bool inside = yy >= 0 && yy < h && xx >= 0 && xx < w;
    uint4* ptr = reinterpret_cast<uint4*>(smem + p * 128 + ((q ^ (p & 7)) << 4));
    uint4 v = *ptr;
    __half2* hv = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(hv[j]);
      f.x = silu_f(fmaf(f.x, a[2 * j], b[2 * j]));
      f.y = silu_f(fmaf(f.y, a[2 * j + 1], b[2 * j + 1]));
      hv[j] = __floats2half2_rn(f.x, f.y);
    }
    if (!inside) v = make_uint4(0u, 0u, 0u, 0u);
    *ptr = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1280; i += 128) reinterpret_cast<uint4*>(gbuf)[i] = reinterpret_cast<uint4*>(smem)[i];
}
