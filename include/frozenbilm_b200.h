/* frozenbilm_b200 -- C-ABI of libfrozenbilm_sm100.so
 *
 * The reference (antoyang/FrozenBiLM) has no FFI: its "operator interface" for the masked-LM
 * hot path is the Python surface of package `model` (model/__init__.py:14-112,
 * model/deberta.py:1414-1501).  This header is the drop-in boundary UNDER that surface: the
 * Python mirror in frozenbilm_b200/model binds exactly these entry points with ctypes
 * (see INTEGRATION.md for the stub).  Conventions for every function:
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless noted;
 *   - `stream` is a cudaStream_t passed as void* (e.g. torch.cuda.current_stream().cuda_stream);
 *   - the library never allocates or frees caller tensors and never synchronises the host;
 *   - returns 0 on success, <0 on error; fb_last_error() returns a static description;
 *   - bf16 tensors are row-major with an explicit leading dimension in ELEMENTS;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 * Each entry point cites the reference code it replaces.
 */
#ifndef FROZENBILM_B200_H
#define FROZENBILM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 1

const char* fb_last_error(void);
int fb_abi_version(void);
/* Number of kernels this library has launched since load (bench.py reports it as gpu_launches). */
long long fb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * fb_gemm: D[b] = epilogue( alpha * sum_k A[b](m,k) * B[b](n,k) )   bf16 x bf16 -> fp32 accumulate
 * tcgen05.mma (TMEM accumulators) fed by TMA, persistent over 128 x block_n output tiles.
 * Replaces every nn.Linear / bmm on the path: deberta.py:757-765 (Q/K/V), :255 (attn out),
 * :311 (FFN1 + GELU), :329 (FFN2), adapter.py:37-42 (down/up), deberta.py:397 (conv as a
 * K=3H GEMM), :1545-1556 (head dense + vocab GEMM), :993 (linear_video), and their dX/dW in
 * the selective backward (torch autograd in the reference, main.py:81).
 * ------------------------------------------------------------------------------------------ */
enum { FB_ACT_NONE = 0, FB_ACT_GELU = 1, FB_ACT_RELU = 2 };
enum { FB_AUX_NONE = 0, FB_AUX_MUL_GELU_GRAD = 1, FB_AUX_MUL_RELU_MASK = 2, FB_AUX_ADD = 3 };
enum { FB_DT_BF16 = 0, FB_DT_F32 = 1 };

typedef struct fb_gemm_args {
  int M, N, K, batch;
  /* A: a_mn_major = 0 -> stored [M, K] (lda = row pitch); 1 -> stored [K, M] */
  const void* A; long long lda; long long stride_a; int a_mn_major;
  /* B: b_mn_major = 0 -> stored [N, K] (nn.Linear weight layout); 1 -> stored [K, N] */
  const void* B; long long ldb; long long stride_b; int b_mn_major;
  void* D; long long ldd; long long stride_d; int d_dtype;   /* FB_DT_BF16 | FB_DT_F32 */
  float alpha;
  const float* bias;          /* [N] fp32 or NULL: v += bias[n] */
  void* D_pre; long long ld_pre;  /* optional bf16 [M, N]: receives v BEFORE the activation */
  int act;                    /* FB_ACT_* */
  const void* aux; long long ld_aux; int aux_op;  /* optional bf16 [M, N], FB_AUX_* */
  const float* resid; long long ld_resid;         /* optional fp32 [M, N]: v += resid */
  int accumulate;             /* fp32 D only: D += v  */
  int split_k;                /* >1: K is split over CTAs, results reduced with fp32 atomics
                                 (requires d_dtype = F32 and accumulate semantics; bias/aux/resid
                                 applied by split 0 only) */
  int block_n;                /* 0 = auto, else 64 | 128 | 192 | 256 */
} fb_gemm_args;

int fb_gemm(const fb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FROZENBILM_B200_H */
