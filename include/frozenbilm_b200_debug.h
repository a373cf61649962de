/* frozenbilm_b200_debug.h -- TEST-ONLY entry points of libfrozenbilm_sm100_dbg.so.
 *
 * The debug library is the product library's sources compiled with -DFB_DEBUG_SURFACE plus fb_probe.cu; it is
 * loaded only by scripts/ and tests/ (hardware layout probes, phase timelines).  None of these symbols exist in
 * libfrozenbilm_sm100.so. */
#ifndef FROZENBILM_B200_DEBUG_H
#define FROZENBILM_B200_DEBUG_H
#include "frozenbilm_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Profiling aid: when non-NULL (device int64 [pairs * 32]), CTA (0,0) of every following fb_attn_bwd launch records
 * clock64() stamps of its phases (see FB_STAMP in fb_attn_bwd.cu).  NULL (the default) disables it. */
int fb_attn_bwd_set_debug(long long* stamps);

/* Hardware test bench (tests/ and scripts/ only, not on the training path): runs n_mma tcgen05.mma instructions of
 * one CTA on caller-supplied raw shared-memory images of A and B, with caller-supplied descriptor templates
 * (a_desc / b_desc: every field except the start address) and per-instruction start offsets, and returns the TMEM
 * accumulator as fp32 [128, N].  Used to pin operand-layout semantics (SWIZZLE_NONE strides, skewed SBO) on a B200
 * before a kernel depends on them. */
typedef struct fb_umma_probe_args {
  const void* a_image; int a_bytes;          /* device pointers; sizes are multiples of 16, <= 128 KB / 64 KB */
  const void* b_image; int b_bytes;
  int n_mma; const int* a_off; const int* b_off;   /* host arrays [n_mma], byte offsets of the descriptor start */
  unsigned long long a_desc, b_desc;
  unsigned int idesc;
  int N;
  float* out;                                /* device fp32 [128, N] */
  long long* cycles;                         /* optional device int64 [2]: SM cycles (all issued, all complete) */
} fb_umma_probe_args;
int fb_umma_probe(const fb_umma_probe_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
