// Bandwidth probe (profiles/bw_probe.py only): streaming kernels of different structure and arithmetic weight over the
// same buffers, to separate "DRAM-bound" from "SM-side bound" behaviour when the board sits at its power cap.
#include "common.cuh"
#include "../../include/gd_b200.h"
#include "../../include/gd_b200_devtools.h"
#ifdef GD_B200_DEVTOOLS

namespace gd {
void count_launch(int n = 1);
namespace {

template <int kMath>
__device__ __forceinline__ Half8 probe_math(const Half8& v) {
  if (kMath == 0) return v;
  float f[8], o[8];
  half8_to_float(v, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float z = fmaf(f[j], 0.4999f, 0.001f);
    o[j] = kMath == 1 ? z : fmaf(z, tanh_approx(z), z);
  }
  return float_to_half8(o);
}

// one-shot flat: every thread moves 4 x 16 bytes, 128 threads per CTA own a contiguous 8 KiB
template <int kMath>
__global__ void probe_flat_kernel(const __half* __restrict__ src, __half* __restrict__ dst, size_t chunks) {
  const size_t base = (static_cast<size_t>(blockIdx.x) * blockDim.x) * 4 + threadIdx.x;
  Half8 v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (base + u * blockDim.x < chunks) v[u] = ld_half8_stream(src + (base + u * blockDim.x) * 8);
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (base + u * blockDim.x < chunks) st_half8(dst + (base + u * blockDim.x) * 8, probe_math<kMath>(v[u]));
}

// persistent grid-stride: 8 loads in flight per thread
template <int kMath>
__global__ void probe_stride_kernel(const __half* __restrict__ src, __half* __restrict__ dst, size_t chunks) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < chunks; i += 8 * stride) {
    Half8 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ld_half8_stream(src + (i + u * stride) * 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) st_half8(dst + (i + u * stride) * 8, probe_math<kMath>(v[u]));
  }
  for (; i < chunks; i += stride) st_half8(dst + i * 8, probe_math<kMath>(ld_half8_stream(src + i * 8)));
}

// GroupNorm-like: 256-thread CTA owns a contiguous region and walks it in `iters` rounds of 8 loads -> 8 stores per
// thread (thread t of round r touches chunk base + (r*8 + u)*256 + t)
template <int kMath>
__global__ void probe_chunked_kernel(const __half* __restrict__ src, __half* __restrict__ dst, size_t chunks, int iters) {
  const size_t base = static_cast<size_t>(blockIdx.x) * 256 * 8 * iters + threadIdx.x;
  for (int r = 0; r < iters; ++r) {
    const size_t b = base + static_cast<size_t>(r) * 8 * 256;
    if (b + 7 * 256 >= chunks) return;
    Half8 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ld_half8_stream(src + (b + u * 256) * 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) st_half8(dst + (b + u * 256) * 8, probe_math<kMath>(v[u]));
  }
}

}  // namespace
}  // namespace gd

extern "C" int gd_bw_probe(int32_t structure, int32_t math, const void* src, void* dst, int64_t bytes, void* stream) {
  using namespace gd;
  GD_REQUIRE(src && dst && bytes > 0 && bytes % 16 == 0, "gd_bw_probe: bad arguments");
  const size_t chunks = static_cast<size_t>(bytes) / 16;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __half* s = reinterpret_cast<const __half*>(src);
  __half* d = reinterpret_cast<__half*>(dst);
  if (structure < 0) {
    const int iters = -structure;
    const unsigned grid = static_cast<unsigned>((chunks + 256 * 8 * iters - 1) / (256 * 8 * iters));
    if (math == 0) probe_chunked_kernel<0><<<grid, 256, 0, st>>>(s, d, chunks, iters);
    else if (math == 1) probe_chunked_kernel<1><<<grid, 256, 0, st>>>(s, d, chunks, iters);
    else probe_chunked_kernel<2><<<grid, 256, 0, st>>>(s, d, chunks, iters);
  } else if (structure == 0) {
    const unsigned grid = static_cast<unsigned>((chunks + 511) / 512);
    if (math == 0) probe_flat_kernel<0><<<grid, 128, 0, st>>>(s, d, chunks);
    else if (math == 1) probe_flat_kernel<1><<<grid, 128, 0, st>>>(s, d, chunks);
    else probe_flat_kernel<2><<<grid, 128, 0, st>>>(s, d, chunks);
  } else {
    const unsigned grid = 148u * static_cast<unsigned>(structure);  // structure = CTAs per SM
    if (math == 0) probe_stride_kernel<0><<<grid, 256, 0, st>>>(s, d, chunks);
    else if (math == 1) probe_stride_kernel<1><<<grid, 256, 0, st>>>(s, d, chunks);
    else probe_stride_kernel<2><<<grid, 256, 0, st>>>(s, d, chunks);
  }
  GD_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}
#endif  // GD_B200_DEVTOOLS
