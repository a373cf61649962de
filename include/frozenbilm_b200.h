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

/* ------------------------------------------------------------------------------------------
 * fb_attn_fwd: fused disentangled self-attention forward for one layer pass (head_dim = 64).
 * Replaces DisentangledSelfAttention.forward (deberta.py:717-818) + disentangled_attention_bias
 * (:820-947) + XSoftmax (:100-138) + attention-prob dropout (:796):
 *   s[b,h,i,j] = (Q_i.K_j + Q_i.PKrel[h, i-j] + K_j.PQrel[h, i-j]) * scale ; masked softmax ; P.V
 * Q/K/V: bf16, token row (b*L + i) with pitch ld*, head hh at columns [hh*64, hh*64+64).
 * PKrel/PQrel: bf16 [h, rel_rows, 64] tables indexed by relative position, row = (i - j) + rel_off
 * (fb_build_rel_tables); mask: uint8 [B, L] (cat(video_mask, attention_mask)); M[b,i,j] = m_i*m_j.
 * ctx: bf16 [B*L, ldc]; lse: fp32 [B, h, L] = log2-domain log-sum-exp of the scaled scores
 * (+inf for fully masked rows) for the backward.  Dropout masks come from Philox(seed, offset).
 * ------------------------------------------------------------------------------------------ */
typedef struct fb_attn_args {
  int B, L, h;
  const void* Q; long long ldq;
  const void* K; long long ldk;
  const void* V; long long ldv;
  const void* PKrel; const void* PQrel; int rel_rows; int rel_off;
  const unsigned char* mask;
  void* ctx; long long ldc;
  float* lse;
  float scale;
  float dropout_p; unsigned long long seed; unsigned long long offset;
} fb_attn_args;

int fb_attn_fwd(const fb_attn_args* args, void* stream);

/* Expands pqk = Linear_{q|k}(rel_emb) ([2S, 2H] bf16: PQ in cols [0,H), PK in [H,2H); deberta.py:847-853)
 * into the per-relative-position tables used above: out[hh][row] = pqk[bucket_of_rel[row - rel_off + 511]].
 * bucket_of_rel: int32 [1023] = clamp(bucket(rel) + S, 0, 2S-1) for rel in [-511, 511]
 * (make_log_bucket_position, deberta.py:578-589).  Rows with |rel| >= L are zero. */
int fb_build_rel_tables(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                        int rel_off, int L, void* out_pq, void* out_pk, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FROZENBILM_B200_H */
