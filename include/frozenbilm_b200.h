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

#define FB_ABI_VERSION 3   /* 3: fb_gemm_args.stride_bias, the *_batched entry points */

const char* fb_last_error(void);
int fb_abi_version(void);
/* Programmatic dependent launch for the launches that follow (process-wide; returns the previous setting; FB_PDL=0 in
 * the environment disables it for good).  Every kernel of the library blocks on griddepcontrol.wait before its first
 * global access, so a dependent launch only overlaps launch latency and prologues.  The host side switches it off
 * while the low-priority side stream has work (the backward of a training step): early-resident kernels that sit in
 * griddepcontrol.wait keep the side stream's CTAs off the SMs they were meant to fill (measured: +1.5 ms/step). */
int fb_set_pdl(int on);
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
/* One dropout site: masks are Philox4x32-10(seed; row, col/8, offset) at 16-bit resolution, so the
 * backward regenerates them from (seed, offset) instead of storing them.  p = 0 disables. */
typedef struct fb_dropout { float p; unsigned long long seed; unsigned long long offset; } fb_dropout;

enum { FB_ACT_NONE = 0, FB_ACT_GELU = 1, FB_ACT_RELU = 2 };
enum { FB_AUX_NONE = 0, FB_AUX_MUL_GELU_GRAD = 1, FB_AUX_MUL_RELU_MASK = 2, FB_AUX_ADD = 3 };
enum { FB_DT_BF16 = 0, FB_DT_F32 = 1 };

typedef struct fb_gemm_args {
  int M, N, K, batch;         /* batch = number of OUTER batches (stride_*) */
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
  /* optional second (inner) batch level, e.g. heads inside a token row: total = batch * batch_inner */
  int batch_inner; long long stride_a_in, stride_b_in, stride_d_in;
  const unsigned char* rowmask;  /* optional uint8 [M]: v = 0 for rows with rowmask == 0 (after bias) */
  fb_dropout drop; int drop_pos; /* 0 none, 1 = after bias/rowmask (before D_pre and act),
                                    2 = after act/aux (before resid); element id = (row, col) */
  long long stride_bias;         /* batched mode: outer batch o uses bias + o * stride_bias (0 = one shared bias) */
} fb_gemm_args;

int fb_gemm(const fb_gemm_args* args, void* stream);
/* Resident 2-CTA clusters of the multicast GEMM kernel on the current device (0: cluster kernel unavailable). */
int fb_gemm_cluster_capacity(void);

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
  void* s_out;    /* optional fp16 [B, h, L, ceil128(L)]: masked scaled scores in the log2 domain (-inf = masked),
                     kept for the backward (training only) */
  int key_mask_only; /* 0: DeBERTa mask M[i,j] = m_i*m_j (deberta.py:480-490); 1: BERT additive key mask
                        (bert.py:165-169): padded QUERY rows are computed normally.  PKrel = PQrel = NULL selects
                        plain attention (BertSelfAttention, bert.py:141-191). */
} fb_attn_args;

int fb_attn_fwd(const fb_attn_args* args, void* stream);

/* Attention backward (autograd through deberta.py:769-808 in the reference), staged so that every
 * contraction is an fb_gemm:
 *   dP  = dO . V^T                                  (fb_gemm, batched over (b, h))
 *   fb_attn_bwd_prep: P = exp2(s2 - lse); dS = P * (dropout_bwd(dP) - rowsum(dO*O)) * scale, written as
 *        ds [B,h,L,lp], pd [B,h,L,lp] (dropped probabilities), z1 / z2 [h, B*L, rel_rows] = dS regrouped by
 *        relative position for the c2p / p2c terms
 *   dV = pd^T . dO ; dQ = ds . K + z1 . PKrel ; dK = ds^T . Q + z2 . PQrel ;
 *   dPKrel = z1^T . Q ; dPQrel = z2^T . K          (fb_gemm), then fb_rel_to_bucket folds them onto the
 *   2S bucket rows of d(Linear_{q|k}(rel_emb)). */
typedef struct fb_attn_bwd_prep_args {
  int B, L, h, lp, rel_rows, rel_off;
  const void* s2; const float* lse;         /* from fb_attn_fwd (s_out: fp16, lse: fp32) */
  const void* dp;                            /* bf16 [B,h,L,lp] */
  const void* ctx; const void* dctx; long long ldc;   /* bf16 [B*L, ldc] */
  void* ds; void* pd; void* z1; void* z2;    /* bf16 outputs */
  float scale;
  float dropout_p; unsigned long long seed; unsigned long long offset;
} fb_attn_bwd_prep_args;
int fb_attn_bwd_prep(const fb_attn_bwd_prep_args* args, void* stream);
int fb_rel_to_bucket(const float* gpq, const float* gpk, int H, int h, const int* bucket_of_rel, int rel_rows,
                     int rel_off, int L, float* out, long long ldo, void* stream);
/* The same for nb layer passes in one launch: gpq / gpk are [nb, h, rel_rows, 64], pass b accumulates into
 * out + b * stride_out. */
int fb_rel_to_bucket_batched(const float* gpq, const float* gpk, int H, int h, const int* bucket_of_rel, int rel_rows,
                             int rel_off, int L, float* out, long long ldo, int nb, long long stride_out, void* stream);

/* Fused attention backward (same autograd, deberta.py:769-808 / 820-947 / XSoftmax :134-138 / XDropout :188-194),
 * one tcgen05 kernel per layer pass instead of the staged schedule above: one CTA per (batch, head) recomputes
 * P = exp2(s2 - lse) and dS' = P * (dropout_bwd(dP) - rowsum(dO*O)) tile by tile, keeps dQ / dK / dV in TMEM and
 * writes them once as bf16:
 *   dV = Pd^T.dO ; dQ = scale (dS'.K + Z1.PKrel) ; dK = scale (dS'^T.Q + Z2.PQrel)   (Z1/Z2 = dS' regrouped by
 *   relative position inside shared memory, never in HBM).  ds_out (optional, bf16 [B,h,L,lp]) receives the
 *   UNSCALED dS' and feeds fb_attn_bwd_rel.  PKrev is the row-reversed PKrel table of fb_build_rel_tables.
 * PKrev = PQrel = NULL selects plain attention (BERT, bert.py:141-191).  O and dO share the row pitch ldo.
 * Every operand and output is moved by TMA: base pointers 16-byte aligned, row pitches multiples of 8 elements;
 * rows >= L of the last 128-row tile are clipped by the tensor maps (nothing is written past L). */
typedef struct fb_attn_bwd_args {
  int B, L, h, lp;                           /* lp = L rounded up to 128 = row pitch of s2 / ds_out */
  const void* Q; long long ldq;              /* bf16 [B*L, >= h*64] views, as in fb_attn_fwd */
  const void* K; long long ldk;
  const void* V; long long ldv;
  const void* O; const void* dO; long long ldo;   /* bf16 context and its gradient */
  const void* s2; const float* lse;          /* fp16 [B,h,L,lp] scores (log2 domain, -inf = masked), fp32 [B,h,L] */
  const void* PKrev; const void* PQrel; int rel_rows, rel_off;
  void* dQ; long long lddq;                  /* bf16 outputs */
  void* dK; long long lddk;
  void* dV; long long lddv;
  void* ds_out;
  float scale;
  float dropout_p; unsigned long long seed; unsigned long long offset;   /* must equal the forward's */
} fb_attn_bwd_args;
int fb_attn_bwd(const fb_attn_bwd_args* args, void* stream);

/* Gradients of the per-relative-position tables (fp32 [h, rel_rows, 64], OVERWRITTEN):
 *   gpk[h, (i-j)+rel_off] = scale sum_{b,i,j} dS'[b,h,i,j] Q[b,h,i]   (c2p, deberta.py:870-881)
 *   gpq[h, (i-j)+rel_off] = scale sum_{b,i,j} dS'[b,h,i,j] K[b,h,j]   (p2c, deberta.py:884-918)
 * One CTA per (head, tile diagonal, batch slice) accumulates its 256-row band in TMEM and flushes once.
 * units_per_cta = 0 picks the slice size. */
typedef struct fb_attn_bwd_rel_args {
  int B, L, h, lp, rel_rows, rel_off;
  const void* ds;                            /* bf16 [B,h,L,lp]: dS' from fb_attn_bwd (unscaled) */
  const void* Q; long long ldq;
  const void* K; long long ldk;
  float* gpk; float* gpq;
  int units_per_cta;
  float scale;
} fb_attn_bwd_rel_args;
int fb_attn_bwd_rel(const fb_attn_bwd_rel_args* args, void* stream);

/* Expands pqk = Linear_{q|k}(rel_emb) ([2S, 2H] bf16: PQ in cols [0,H), PK in [H,2H); deberta.py:847-853)
 * into the per-relative-position tables used above: out[hh][row] = pqk[bucket_of_rel[row - rel_off + 511]].
 * bucket_of_rel: int32 [1023] = clamp(bucket(rel) + S, 0, 2S-1) for rel in [-511, 511]
 * (make_log_bucket_position, deberta.py:578-589).  Rows with |rel| >= L are zero.  out_pk_rev (nullable) receives
 * out_pk with its rows reversed (row r <-> rel_rows - 1 - r): the operand fb_attn_bwd wants. */
int fb_build_rel_tables(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                        int rel_off, int L, void* out_pq, void* out_pk, void* out_pk_rev, void* stream);
/* The tables of nb layer passes in one launch (the positional projections do not depend on the activations, so the
 * host side builds them for the whole forward up front): pass b reads pqk + b * stride_pqk (elements) and writes
 * out_* + b * h * rel_rows * 64. */
int fb_build_rel_tables_batched(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                                int rel_off, int L, void* out_pq, void* out_pk, void* out_pk_rev, int nb,
                                long long stride_pqk, void* stream);

/* ------------------------------------------------------------------------------------------
 * HBM-bound row kernels (fb_elementwise.cu).  All matrices are dense row-major [M, H].
 * ------------------------------------------------------------------------------------------ */
/* y = dropout(LayerNorm(x) * rowmask): torch LayerNorm (+ "* mask" + StableDropout) at
 * deberta.py:1043-1054 (embeddings), :259/:333 (sub-layer outputs), :405-417 (conv), :474-478 (relative
 * embeddings), :1547 (head).  Writes fp32 and/or bf16 copies plus the row statistics for the backward. */
int fb_layernorm_fwd(const float* x, int M, int H, const float* gamma, const float* beta, float eps,
                     const unsigned char* rowmask, const fb_dropout* out_drop, float* y_f32, void* y_bf16,
                     float* mean, float* rstd, void* stream);
/* Backward of the above (autograd in the reference).  dx_f32: gradient of the LN input; dx_bf16: the same with
 * the upstream hidden dropout (in_drop) backward applied, ready as a GEMM operand; dgamma/dbeta ACCUMULATE;
 * dx_bf16_colsum (nullable, [H], accumulates) = column sums of dx_bf16 = the adapter up-projection bias gradient. */
int fb_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     int M, int H, const unsigned char* rowmask, const fb_dropout* out_drop,
                     const fb_dropout* in_drop, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                     float* dx_bf16_colsum, void* stream);
/* out[idx_out[r],:] = in[idx_in[r],:] with fp32<->bf16 conversion (null index array = identity, <0 skips):
 * word_embeddings lookup (deberta.py:1013) and the labelled-position gather/scatter of the MLM head. */
int fb_move_rows(const void* in, long long ld_in, int in_bf16, const long long* idx_in, const long long* idx_out,
                 int R, int ncols, void* out, long long ld_out, int out_bf16, void* stream);
/* out_bf16 = dy * gelu'(pre) [* dropout_bwd] [* rowmask]: GELU backward of the conv branch (deberta.py:401-403)
 * and of the MLM head transform (deberta.py:1545-1546). */
int fb_gelu_bwd(const float* dy, const void* pre_bf16, int M, int N, const unsigned char* rowmask,
                const fb_dropout* drop, void* out_bf16, void* stream);
/* out[b,l,:] = a[b,l,:] + pos[l,:]: z = position_embeddings + h[-2] of the enhanced mask decoder
 * (deberta.py:1394-1396); fp32 and bf16 copies. */
int fb_add_rowbcast(const float* a, const float* b, int B, int L, int H, float* out_f32, void* out_bf16,
                    void* stream);
/* Conv1d(H, H, k=3, pad=1) over the sequence (deberta.py:395-400) as a K = 3H GEMM: im2col / col2im. */
int fb_conv_im2col(const void* x_bf16, int B, int L, int H, void* out_bf16, void* stream);
int fb_conv_col2im(const void* dcol_bf16, int B, int L, int H, float* dx, int accumulate, void* stream);
/* out[n] += sum_m x[m, n]: bias gradients of the adapters / linear_video. */
int fb_colsum_bf16(const void* x, long long ld, int M, int N, float* out, void* stream);
/* CrossEntropyLoss over fp32 logits [P, V] (deberta.py:1481-1488): *loss_sum += loss_scale * sum_rows CE;
 * dlogits (bf16, nullable) = (softmax - onehot) * grad_scale. */
int fb_cross_entropy(const float* logits, long long ld, int P, int V, const long long* labels, float* loss_sum,
                     float loss_scale, void* dlogits_bf16, long long ldd, float grad_scale, void* stream);
/* Answer scoring at the [MASK] rows: softmax over the answer vocabulary + top-k
 * (videoqa.py:164-170 logits.softmax(-1) / torch.topk; mc.py:168-172 softmax(-1)[:, 0]).  logits fp32 [P, V] (pitch
 * ld); probs (nullable) fp32 [P, V] (pitch ldp); top_p fp32 [P, k] descending, top_i int64 [P, k] (ties -> lower id). */
int fb_softmax_topk(const float* logits, long long ld, int P, int V, int k, float* probs, long long ldp, float* top_p,
                    long long* top_i, void* stream);
/* out = dropout(x) as fp32 (optionally accumulating) and/or bf16: relative-embedding dropout, deberta.py:779. */
int fb_dropout_apply(const float* x, int M, int N, const fb_dropout* drop, float* out_f32, int accumulate,
                     void* out_bf16, void* stream);
/* fb_dropout_apply for nb layer passes in one launch; pass b uses the site offset drop->offset + b * offset_stride and
 * element ids (row, col) (bit-identical to nb separate calls).  Exactly one output:
 *   out_bf16 [nb, M, N]     = dropout_b(x [M, N])                      (x shared by the passes)
 *   out_f32_accum [M, N]   += sum_b dropout_b(x [nb, M, N])            (backward: dR accumulates over the passes) */
int fb_dropout_batched(const float* x, int M, int N, int nb, const fb_dropout* drop, unsigned long long offset_stride,
                       float* out_f32_accum, void* out_bf16, void* stream);
/* clip_grad_norm_ + Adam over a flat fp32 bucket (main.py:82-84, args.py:246-275), refreshing the bf16 copy. */
int fb_sumsq(const float* g, long long n, float* out, void* stream);
/* Same sum of squares, OVERWRITING *out, with a fixed summation order (bit-identical on every data-parallel rank for
 * identical gradients).  scratch: fp32 [scratch_len >= 2], zero-initialised ONCE by the caller and then owned by this
 * function (per-CTA partials + a ticket counter that the kernel resets itself). */
int fb_sumsq_det(const float* g, long long n, float* out, float* scratch, int scratch_len, void* stream);
int fb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, int step, const float* gnorm_sq, float max_norm, float grad_scale, void* p_bf16,
                 void* stream);
int fb_cast_f32_bf16(const float* x, long long n, void* out, void* stream);



/* ------------------------------------------------------------------------------------------
 * fp32-class instantiation of the path (fb_precise.cu), selected by the host side with FB_PRECISION=fp32x3 /
 * model.set_precision("fp32x3"): the parity gate of BASELINE.json configs[0] (north_star: logits within
 * rtol 1e-3 / atol 1e-4 of the fp32 reference), which bf16 tensor-core operands cannot meet.
 * ------------------------------------------------------------------------------------------ */
/* x fp32 [rows, cols] (pitch ld) -> three bf16 pieces p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) laid
 * out as SIX blocks along the GEMM K axis: A side (b_mode = 0): p0 p0 p1 p1 p0 p2,  B side (b_mode = 1): p0 p1 p0 p1
 * p2 p0, so that fb_gemm over the 6x longer K accumulates a0b0 + a0b1 + a1b0 + a1b1 + a0b2 + a2b0 in fp32: every
 * product above 2^-24 |a||b| (fp32-class nn.Linear / bmm, same call sites as fb_gemm).  along_rows = 0: K = columns,
 * block t starts at column t * block_stride of out [rows, >= 6 block_stride]; along_rows = 1: K = rows (MN-major
 * operand), block t starts at row t * block_stride.  Padding must be zeroed by the caller. */
int fb_split3(const float* x, long long ld, int rows, int cols, int along_rows, int b_mode, long long block_stride,
              void* out_bf16, long long ld_out, void* stream);
/* fb_gemm's epilogue with every operand in fp32 (same order: alpha, bias, rowmask, dropout pos 1, D_pre, act, aux,
 * dropout pos 2, resid, accumulate); GELU / GELU' use erff. */
typedef struct fb_epilogue_args {
  int M, N;
  const float* in; long long ld_in; float alpha;
  const float* bias; const unsigned char* rowmask; fb_dropout drop; int drop_pos;
  float* D_pre; long long ld_pre; int act;
  const float* aux; long long ld_aux; int aux_op;
  const float* resid; long long ld_resid; int accumulate;
  float* D; long long ldd;
} fb_epilogue_args;
int fb_epilogue_f32(const fb_epilogue_args* args, void* stream);
/* fp32-typed variants of fb_colsum_bf16 / fb_conv_im2col / fb_conv_col2im / fb_cross_entropy (fp32 dlogits). */
int fb_colsum_f32(const float* x, long long ld, int M, int N, float* out, void* stream);
int fb_conv_im2col_f32(const float* x, int B, int L, int H, float* out, void* stream);
int fb_conv_col2im_f32(const float* dcol, int B, int L, int H, float* dx, int accumulate, void* stream);
int fb_cross_entropy_f32(const float* logits, long long ld, int P, int V, const long long* labels, float* loss_sum,
                         float loss_scale, float* dlogits_f32, long long ldd, float grad_scale, void* stream);
/* Disentangled attention in fp32 on the CUDA cores (deberta.py:717-947; plain attention when pqk = NULL, bert.py:
 * 141-191).  pqk: fp32 [2S, >= 2H] = Linear_{q|k}(dropout(LN(rel_embeddings))), PQ in columns [0, H), PK in [H, 2H);
 * bucket_of_rel as in fb_build_rel_tables.  lse is the natural-log log-sum-exp (+inf = dead row).  The backward
 * overwrites dQ / dK / dV, ACCUMULATES into dpqk ([2S, >= 2H]) and needs two fp32 [B, h, L, L] workspaces. */
typedef struct fb_attn_f32_args {
  int B, L, h;
  const float* Q; long long ldq; const float* K; long long ldk; const float* V; long long ldv;
  const float* pqk; long long ld_pqk; const int* bucket_of_rel;
  const unsigned char* mask; int key_mask_only;
  float scale; float dropout_p; unsigned long long seed; unsigned long long offset;
  float* ctx; long long ldc; float* lse;
  /* backward only */
  const float* dctx; float* dQ; long long lddq; float* dK; long long lddk; float* dV; long long lddv;
  float* dpqk; long long ld_dpqk; float* ws_ds; float* ws_pd;
} fb_attn_f32_args;
int fb_attn_fwd_f32(const fb_attn_f32_args* args, void* stream);
int fb_attn_bwd_f32(const fb_attn_f32_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input side of the path (fb_data.cu).
 * ------------------------------------------------------------------------------------------ */
/* Masked-LM corruption on the device (reference util/misc.py:14-56 mask_tokens, called from main.py:52-58):
 * a token is selected with probability mlm_probability unless special_mask[i] != 0 (nullable) or ids[i] == pad_id
 * (pad_id < 0: no pad token); selected tokens become mask_id (80 %), a uniform random id in [0, vocab) (10 %) or stay
 * (10 %); labels_out = original id where selected, -100 elsewhere.  Decisions are a pure function of (seed, i):
 * oracle/data_oracle.py reproduces them bit for bit.  ids / outputs: int64 [n] (ids_out may alias ids). */
int fb_mlm_corrupt(const long long* ids, const unsigned char* special_mask, long long n, long long pad_id,
                   long long mask_id, long long vocab, float mlm_probability, unsigned long long seed,
                   long long* ids_out, long long* labels_out, void* stream);
/* Frame subsampling / padding of CLIP features (reference datasets/videotext_dataset.py:23-44 + collate :47-57):
 * shard_f16 holds the [n_b, D] fp16 feature matrices of all videos back to back, sample b starts at row offsets[b]
 * and has nframes[b] rows (0 = missing video).  out[b, j] = frame (j * n) / max_feats when n > max_feats, frame j when
 * j < n, zeros otherwise; video_len[b] = min(n, max_feats).  out_f32 and/or out_bf16: [B, max_feats, D]. */
int fb_gather_frames(const void* shard_f16, const long long* offsets, const int* nframes, int B, int max_feats, int D,
                     float* out_f32, void* out_bf16, long long* video_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FROZENBILM_B200_H */
