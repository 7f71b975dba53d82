/* Development / measurement hooks of libgd_b200.so — NOT part of the product ABI (include/gd_b200.h).
 * Present only in libraries built with -DGD_B200_DEVTOOLS (the default of the in-tree build, which the profiling
 * scripts under profiles/ and the kernel-selection parity test need; GD_B200_NO_DEVTOOLS=1 builds without them). */
#ifndef GD_B200_DEVTOOLS_H_
#define GD_B200_DEVTOOLS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Measurement hooks (profiles/ only): key 0 = conv epilogue mode (0 normal, 1 barriers only, 2 TMEM loads only,
 * 3 everything but the TMA store), key 1 = force the conv N tile (0 = heuristic), key 2 = 1 disables the staged
 * TMA-store epilogue, key 3 = 0 disables CTA-pair (cta_group::2) mode, key 4 = 0 disables halo reuse (every tap
 * loads its own activation tile), key 5 = 0 disables the tcgen05 attention forward (mma.sync kernel for every length),
 * key 6 = 1 launches the frequent kernels with programmatic dependent launch (default 0: measured slower, DESIGN.md 4.2;
 * also GD_B200_PDL=1), key 7 = force the activation-ring depth of the fused-GroupNorm conv kernels (0 = heuristic), key 8 = 0 disables
 * the 4-CTA-cluster mode (weight tiles multicast to two CTA pairs), key 9 = 1 enables split-K for convs that were lent a
 * workspace (gd_conv_desc.splitk_ws; default 0, also GD_B200_SPLITK=1). */
void gd_debug_set(int key, int value);
/* Measurement hook (profiles/bw_probe.py): stream `bytes` from src to dst. structure 0 = one-shot flat grid, -k = 256-thread
 * CTAs owning a contiguous region walked in k rounds of 8 loads/stores per thread, k>0 =
 * persistent grid-stride with k CTAs per SM; math 0 = copy, 1 = fp16->fp32 FMA->fp16, 2 = + SiLU (GroupNorm's arithmetic). */
int gd_bw_probe(int32_t structure, int32_t math, const void* src, void* dst, int64_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GD_B200_DEVTOOLS_H_ */
