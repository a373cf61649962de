// fb_precise.cu -- the fp32-class instantiation of the path ("split-bf16", selected with FB_PRECISION=fp32x3).
//
// north_star asks for MLM logits within rtol 1e-3 / atol 1e-4 of the fp32 reference.  One bf16 rounding of a GEMM
// operand is 2^-9 relative, so the bf16 training path cannot meet that bound; this file supplies what the SAME kernel
// schedule needs to run every contraction at fp32-class (~2^-24 relative) precision instead:
//
//   * fb_split3:       x (fp32) -> three bf16 pieces p0 + p1 + p2 (24 mantissa bits), laid out as six blocks along K
//                      (A side: p0 p0 p1 p1 p0 p2, B side: p0 p1 p0 p1 p2 p0).  The ordinary tcgen05 GEMM (fb_gemm) over
//                      the 6x longer K then accumulates the six products larger than 2^-24 |a||b| in fp32 TMEM: an
//                      fp32-class product on the bf16 tensor pipe.  No new tensor-core kernel: the product GEMM is the
//                      one measured.  (Two pieces / three products leave 2^-17 per operand: measured 1.1e-4 on the
//                      configs[0] logits after 12 layers, just over north_star's atol.)
//   * fb_epilogue_f32: the GEMM epilogue (alpha, bias, row mask, dropout, pre-activation copy, GELU/ReLU, aux op,
//                      residual, accumulate) with every operand in fp32 (the bf16 path rounds aux / D_pre).
//   * fb_attn_fwd_f32 / fb_attn_bwd_f32: disentangled attention (deberta.py:717-947) in fp32 on the CUDA cores,
//                      straight from the un-expanded [2S, 2H] positional projections and the bucket table.
//   * fp32 variants of the small bf16-typed helpers (column sums, im2col / col2im).
// These are correctness instantiations (parity gate of BASELINE configs[0]); they are not tuned.
#include <cstring>

#include "../../include/frozenbilm_b200.h"
#include "fb_common.cuh"

namespace fb {

// ---------------------------------------------------------------------------------------------
// split3
// ---------------------------------------------------------------------------------------------
__global__ void split3_kernel(const float* __restrict__ x, long long ld, int rows, int cols, int along_rows, int b_mode,
                              long long block_stride, __nv_bfloat16* __restrict__ out, long long ld_out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int r = (int)(idx / cols), c = (int)(idx % cols);
  const float v = x[(long long)r * ld + c];
  // three bf16 pieces, 24 mantissa bits in total: v = p0 + p1 + p2 (+ <= 2^-25 |v|)
  const __nv_bfloat16 p0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(p0);
  const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
  const __nv_bfloat16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
  // the six products larger than 2^-24 |a||b|:  a0 b0 + a0 b1 + a1 b0 + a1 b1 + a0 b2 + a2 b0
  //   A side blocks: a0 a0 a1 a1 a0 a2      B side blocks: b0 b1 b0 b1 b2 b0
  const __nv_bfloat16 blk[6] = {p0, b_mode ? p1 : p0, b_mode ? p0 : p1, p1, b_mode ? p2 : p0, b_mode ? p0 : p2};
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    if (along_rows) out[(t * block_stride + r) * ld_out + c] = blk[t];
    else out[(long long)r * ld_out + t * block_stride + c] = blk[t];
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 epilogue (same order of operations as fb_gemm's: see fb_gemm.cu epi_pre / epi_post)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_exact_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

struct EpiParams {
  int M, N;
  const float* in; long long ld_in; float alpha;
  const float* bias; const uint8_t* rowmask; DropSpec drop; int drop_pos;
  float* D_pre; long long ld_pre; int act;
  const float* aux; long long ld_aux; int aux_op;
  const float* resid; long long ld_resid; int accumulate;
  float* D; long long ldd;
};

__global__ void epilogue_f32_kernel(const EpiParams p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * p.N) return;
  const int r = (int)(idx / p.N), c = (int)(idx % p.N);
  float v = p.in[(long long)r * p.ld_in + c] * p.alpha;
  if (p.bias) v += p.bias[c];
  if (p.rowmask && p.rowmask[r] == 0) v = 0.0f;
  float keep = 1.0f;
  if (p.drop_pos) keep = ((drop_keep8(p.drop, (uint32_t)r, (uint32_t)c >> 3) >> (c & 7)) & 1u) ? p.drop.inv_keep : 0.0f;
  if (p.drop_pos == 1) v *= keep;
  if (p.D_pre) p.D_pre[(long long)r * p.ld_pre + c] = v;
  if (p.act == FB_ACT_GELU) v = gelu_exact(v);
  else if (p.act == FB_ACT_RELU) v = fmaxf(v, 0.0f);
  if (p.aux_op != FB_AUX_NONE) {
    const float a = p.aux[(long long)r * p.ld_aux + c];
    if (p.aux_op == FB_AUX_MUL_GELU_GRAD) v *= gelu_exact_grad(a);
    else if (p.aux_op == FB_AUX_MUL_RELU_MASK) v = (a > 0.0f) ? v : 0.0f;
    else v += a;
  }
  if (p.drop_pos == 2) v *= keep;
  if (p.resid) v += p.resid[(long long)r * p.ld_resid + c];
  float* d = p.D + (long long)r * p.ldd + c;
  *d = p.accumulate ? *d + v : v;
}

// ---------------------------------------------------------------------------------------------
// small fp32 helpers
// ---------------------------------------------------------------------------------------------
__global__ void colsum_f32_kernel(const float* __restrict__ x, long long ld, int M, int N, int rows_per_cta,
                                  float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float a = 0.0f;
  for (int r = r0; r < r1; ++r) a += x[(long long)r * ld + c];
  atomicAdd(out + c, a);
}
__global__ void im2col3_f32_kernel(const float* __restrict__ x, int L, int H, float* __restrict__ out) {
  const int row = blockIdx.x, t = row % L;
  for (int c = threadIdx.x; c < 3 * H; c += blockDim.x) {
    const int k = c / H, cc = c % H, tt = t + k - 1;
    out[(long long)row * 3 * H + c] = (tt >= 0 && tt < L) ? x[(long long)(row + k - 1) * H + cc] : 0.0f;
  }
}
__global__ void col2im3_f32_kernel(const float* __restrict__ dcol, int L, int H, float* __restrict__ dx, int accumulate) {
  const int row = blockIdx.x, t = row % L;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float v = dcol[(long long)row * 3 * H + H + c];
    if (t + 1 < L) v += dcol[(long long)(row + 1) * 3 * H + c];
    if (t - 1 >= 0) v += dcol[(long long)(row - 1) * 3 * H + 2 * H + c];
    float* d = dx + (long long)row * H + c;
    *d = accumulate ? *d + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 disentangled attention (head_dim = 64)
//   s[i,j] = scale (Q_i.K_j + Q_i.PK[bkt(i-j)] + K_j.PQ[bkt(i-j)])   (deberta.py:769-777, 870-918; one gather index,
//   SURVEY 0.5);  P = masked softmax (XSoftmax :100-138);  O = dropout(P).V (:796-808)
// ---------------------------------------------------------------------------------------------
struct AttnF32Params {
  int B, L, h;
  const float* Q; long long ldq; const float* K; long long ldk; const float* V; long long ldv;
  const float* pqk; long long ld_pqk; const int* bucket; const uint8_t* mask; int mask_queries;
  float scale; DropSpec drop;
  float* ctx; long long ldc; float* lse;
  const float* dctx; float* dQ; long long lddq; float* dK; long long lddk; float* dV; long long lddv;
  float* dpqk; long long ld_dpqk; float* ds; float* pd;
};

// scores of query row i against key j (this lane's j), natural-log domain; -inf where masked
__device__ __forceinline__ float attn_score(const AttnF32Params& p, int b, int hh, int i, int j, const float* q_s) {
  const float* k = p.K + ((long long)b * p.L + j) * p.ldk + hh * 64;
  float qk = 0.0f, c2p = 0.0f, p2c = 0.0f;
  if (p.pqk) {
    const int bk = p.bucket[i - j + 511];
    const float* pq = p.pqk + (long long)bk * p.ld_pqk + hh * 64;
    const float* pk = pq + (long long)p.h * 64;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      const float kd = k[d], qd = q_s[d];
      qk = fmaf(qd, kd, qk); c2p = fmaf(qd, pk[d], c2p); p2c = fmaf(kd, pq[d], p2c);
    }
  } else {
#pragma unroll 8
    for (int d = 0; d < 64; ++d) qk = fmaf(q_s[d], k[d], qk);
  }
  return (qk + c2p + p2c) * p.scale;
}
__device__ __forceinline__ float attn_keep(const AttnF32Params& p, uint32_t rowid, int j) {
  if (!p.drop.thresh16) return 1.0f;
  return ((drop_keep8(p.drop, rowid, (uint32_t)j >> 3) >> (j & 7)) & 1u) ? p.drop.inv_keep : 0.0f;
}

// one warp per (b, head, query row); 4 warps per CTA; smem: per warp 64 (q) + 64 (dO) + L (p) floats
template <bool BWD>
__global__ void __launch_bounds__(128) attn_f32_row_kernel(const AttnF32Params p) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + warp;   // (b * h + hh) * L + i
  if (row >= (long long)p.B * p.h * p.L) return;
  const int i = (int)(row % p.L), hh = (int)((row / p.L) % p.h), b = (int)(row / ((long long)p.L * p.h));
  float* q_s = sm + warp * (128 + p.L);
  float* do_s = q_s + 64;
  float* p_s = q_s + 128;
  const float* q = p.Q + ((long long)b * p.L + i) * p.ldq + hh * 64;
  q_s[lane] = q[lane]; q_s[lane + 32] = q[lane + 32];
  __syncwarp();
  const bool valid_q = !p.mask_queries || p.mask[(long long)b * p.L + i] != 0;
  float m = -INFINITY;
  for (int j = lane; j < p.L; j += 32) {
    const bool ok = valid_q && p.mask[(long long)b * p.L + j] != 0;
    const float s = ok ? attn_score(p, b, hh, i, j, q_s) : -INFINITY;
    p_s[j] = s;
    m = fmaxf(m, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float lsum = 0.0f;
  if (!BWD) {
    const float mu = (m == -INFINITY) ? 0.0f : m;
    for (int j = lane; j < p.L; j += 32) {
      const float e = expf(p_s[j] - mu);   // exp(-inf) = 0
      p_s[j] = e;
      lsum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    const float inv = lsum > 0.0f ? 1.0f / lsum : 0.0f;
    if (lane == 0) p.lse[row] = lsum > 0.0f ? m + logf(lsum) : INFINITY;
    __syncwarp();
    float o0 = 0.0f, o1 = 0.0f;
    for (int j = 0; j < p.L; ++j) {
      const float w = p_s[j] * inv * attn_keep(p, (uint32_t)row, j);
      if (w != 0.0f) {
        const float* v = p.V + ((long long)b * p.L + j) * p.ldv + hh * 64;
        o0 = fmaf(w, v[lane], o0); o1 = fmaf(w, v[lane + 32], o1);
      }
    }
    float* c = p.ctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
    c[lane] = o0; c[lane + 32] = o1;
  } else {
    // backward, query-row part: P from the saved lse; dS = P (dropout_bwd(dO.V_j) - rowsum(dO O)) scale
    const float lse = p.lse[row];
    const float* g = p.dctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
    const float* o = p.ctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
    do_s[lane] = g[lane]; do_s[lane + 32] = g[lane + 32];
    float D = g[lane] * o[lane] + g[lane + 32] * o[lane + 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) D += __shfl_xor_sync(0xffffffffu, D, off);
    __syncwarp();
    float* ds_row = p.ds + row * p.L;
    float* pd_row = p.pd + row * p.L;
    for (int j = lane; j < p.L; j += 32) {
      const float s = p_s[j];
      const float pr = (s == -INFINITY || lse == INFINITY) ? 0.0f : expf(s - lse);
      const float f = attn_keep(p, (uint32_t)row, j);
      const float* v = p.V + ((long long)b * p.L + j) * p.ldv + hh * 64;
      float dp = 0.0f;
#pragma unroll 8
      for (int d = 0; d < 64; ++d) dp = fmaf(do_s[d], v[d], dp);
      const float dsv = pr * (dp * f - D) * p.scale;
      p_s[j] = dsv;
      ds_row[j] = dsv;
      pd_row[j] = pr * f;
    }
    __syncwarp();
    // dQ_i = sum_j dS_ij (K_j + PK[bkt(i-j)])
    float a0 = 0.0f, a1 = 0.0f;
    for (int j = 0; j < p.L; ++j) {
      const float w = p_s[j];
      if (w != 0.0f) {
        const float* k = p.K + ((long long)b * p.L + j) * p.ldk + hh * 64;
        float k0 = k[lane], k1 = k[lane + 32];
        if (p.pqk) {
          const float* pk = p.pqk + (long long)p.bucket[i - j + 511] * p.ld_pqk + (long long)p.h * 64 + hh * 64;
          k0 += pk[lane]; k1 += pk[lane + 32];
        }
        a0 = fmaf(w, k0, a0); a1 = fmaf(w, k1, a1);
      }
    }
    float* dq = p.dQ + ((long long)b * p.L + i) * p.lddq + hh * 64;
    dq[lane] = a0; dq[lane + 32] = a1;
  }
}

// backward, key-row part: dK_j = sum_i dS_ij (Q_i + PQ[bkt(i-j)]),  dV_j = sum_i Pd_ij dO_i
__global__ void __launch_bounds__(128) attn_f32_bwd_key_kernel(const AttnF32Params p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 4 + warp;   // (b * h + hh) * L + j
  if (row >= (long long)p.B * p.h * p.L) return;
  const int j = (int)(row % p.L), hh = (int)((row / p.L) % p.h), b = (int)(row / ((long long)p.L * p.h));
  const long long base = ((long long)b * p.h + hh) * p.L;
  float k0 = 0.0f, k1 = 0.0f, v0 = 0.0f, v1 = 0.0f;
  for (int i = 0; i < p.L; ++i) {
    const float w = p.ds[(base + i) * p.L + j], u = p.pd[(base + i) * p.L + j];
    if (w != 0.0f) {
      const float* q = p.Q + ((long long)b * p.L + i) * p.ldq + hh * 64;
      float q0 = q[lane], q1 = q[lane + 32];
      if (p.pqk) {
        const float* pq = p.pqk + (long long)p.bucket[i - j + 511] * p.ld_pqk + hh * 64;
        q0 += pq[lane]; q1 += pq[lane + 32];
      }
      k0 = fmaf(w, q0, k0); k1 = fmaf(w, q1, k1);
    }
    if (u != 0.0f) {
      const float* g = p.dctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
      v0 = fmaf(u, g[lane], v0); v1 = fmaf(u, g[lane + 32], v1);
    }
  }
  float* dk = p.dK + ((long long)b * p.L + j) * p.lddk + hh * 64;
  float* dv = p.dV + ((long long)b * p.L + j) * p.lddv + hh * 64;
  dk[lane] = k0; dk[lane + 32] = k1; dv[lane] = v0; dv[lane + 32] = v1;
}

// backward, positional projections: one warp per (head, relative position r = i - j):
//   dPK[bkt(r)] += sum_{b,i} dS[b,i,i-r] Q[b,i],   dPQ[bkt(r)] += sum_{b,i} dS[b,i,i-r] K[b,i-r]
__global__ void __launch_bounds__(128) attn_f32_bwd_rel_kernel(const AttnF32Params p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int idx = blockIdx.x * 4 + warp;
  const int nrel = 2 * p.L - 1;
  if (idx >= nrel * p.h) return;
  const int hh = idx / nrel, r = idx % nrel - (p.L - 1);
  float gk0 = 0.f, gk1 = 0.f, gq0 = 0.f, gq1 = 0.f;
  for (int b = 0; b < p.B; ++b) {
    const long long base = ((long long)b * p.h + hh) * p.L;
    for (int i = max(0, r); i < min(p.L, p.L + r); ++i) {
      const int j = i - r;
      const float w = p.ds[(base + i) * p.L + j];
      if (w != 0.0f) {
        const float* q = p.Q + ((long long)b * p.L + i) * p.ldq + hh * 64;
        const float* k = p.K + ((long long)b * p.L + j) * p.ldk + hh * 64;
        gk0 = fmaf(w, q[lane], gk0); gk1 = fmaf(w, q[lane + 32], gk1);
        gq0 = fmaf(w, k[lane], gq0); gq1 = fmaf(w, k[lane + 32], gq1);
      }
    }
  }
  float* o = p.dpqk + (long long)p.bucket[r + 511] * p.ld_dpqk + hh * 64;
  atomicAdd(o + lane, gq0); atomicAdd(o + lane + 32, gq1);
  o += (long long)p.h * 64;
  atomicAdd(o + lane, gk0); atomicAdd(o + lane + 32, gk1);
}

static int fill_attn(const fb_attn_f32_args* a, AttnF32Params& p) {
  FB_REQUIRE(a && a->Q && a->K && a->V && a->mask && a->ctx && a->lse, "fb_attn_f32: null arg");
  FB_REQUIRE(a->B > 0 && a->h > 0 && a->L > 0 && a->L <= 512, "fb_attn_f32: need 0 < L <= 512 (got %d)", a->L);
  FB_REQUIRE(!a->pqk || a->bucket_of_rel, "fb_attn_f32: pqk needs the bucket table");
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.h = a->h;
  p.Q = a->Q; p.ldq = a->ldq; p.K = a->K; p.ldk = a->ldk; p.V = a->V; p.ldv = a->ldv;
  p.pqk = a->pqk; p.ld_pqk = a->ld_pqk; p.bucket = a->bucket_of_rel; p.mask = a->mask;
  p.mask_queries = a->key_mask_only ? 0 : 1;
  p.scale = a->scale; p.drop = make_drop(a->dropout_p, a->seed, a->offset);
  p.ctx = a->ctx; p.ldc = a->ldc; p.lse = a->lse;
  p.dctx = a->dctx; p.dQ = a->dQ; p.lddq = a->lddq; p.dK = a->dK; p.lddk = a->lddk; p.dV = a->dV; p.lddv = a->lddv;
  p.dpqk = a->dpqk; p.ld_dpqk = a->ld_dpqk; p.ds = a->ws_ds; p.pd = a->ws_pd;
  return 0;
}

}  // namespace fb

using namespace fb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int fb_split3(const float* x, long long ld, int rows, int cols, int along_rows, int b_mode, long long block_stride,
              void* out_bf16, long long ld_out, void* stream) {
  FB_REQUIRE(x && out_bf16 && rows > 0 && cols > 0, "fb_split3: bad args");
  const long long n = (long long)rows * cols;
  split3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, ld, rows, cols, along_rows, b_mode, block_stride,
                                                                    (__nv_bfloat16*)out_bf16, ld_out);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fb_epilogue_f32(const fb_epilogue_args* a, void* stream) {
  FB_REQUIRE(a && a->in && a->D && a->M > 0 && a->N > 0, "fb_epilogue_f32: bad args");
  EpiParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.in = a->in; p.ld_in = a->ld_in; p.alpha = a->alpha;
  p.bias = a->bias; p.rowmask = a->rowmask;
  p.drop = make_drop(a->drop.p, a->drop.seed, a->drop.offset);
  p.drop_pos = (a->drop.p > 0.0f) ? a->drop_pos : 0;
  p.D_pre = a->D_pre; p.ld_pre = a->ld_pre; p.act = a->act;
  p.aux = a->aux; p.ld_aux = a->ld_aux; p.aux_op = a->aux ? a->aux_op : FB_AUX_NONE;
  p.resid = a->resid; p.ld_resid = a->ld_resid; p.accumulate = a->accumulate;
  p.D = a->D; p.ldd = a->ldd;
  const long long n = (long long)a->M * a->N;
  epilogue_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fb_colsum_f32(const float* x, long long ld, int M, int N, float* out, void* stream) {
  FB_REQUIRE(x && out && M > 0 && N > 0, "fb_colsum_f32: bad args");
  const int rows_per_cta = 64;
  dim3 grid((N + 127) / 128, (M + rows_per_cta - 1) / rows_per_cta);
  colsum_f32_kernel<<<grid, 128, 0, ST(stream)>>>(x, ld, M, N, rows_per_cta, out);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_conv_im2col_f32(const float* x, int B, int L, int H, float* out, void* stream) {
  FB_REQUIRE(x && out && B > 0 && L > 0 && H > 0, "fb_conv_im2col_f32: bad args");
  im2col3_f32_kernel<<<B * L, 256, 0, ST(stream)>>>(x, L, H, out);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_conv_col2im_f32(const float* dcol, int B, int L, int H, float* dx, int accumulate, void* stream) {
  FB_REQUIRE(dcol && dx && B > 0 && L > 0 && H > 0, "fb_conv_col2im_f32: bad args");
  col2im3_f32_kernel<<<B * L, 256, 0, ST(stream)>>>(dcol, L, H, dx, accumulate);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fb_attn_fwd_f32(const fb_attn_f32_args* a, void* stream) {
  AttnF32Params p;
  if (int rc = fill_attn(a, p)) return rc;
  const long long rows = (long long)p.B * p.h * p.L;
  const size_t smem = 4 * (128 + p.L) * sizeof(float);
  attn_f32_row_kernel<false><<<(unsigned)((rows + 3) / 4), 128, smem, ST(stream)>>>(p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fb_attn_bwd_f32(const fb_attn_f32_args* a, void* stream) {
  AttnF32Params p;
  if (int rc = fill_attn(a, p)) return rc;
  FB_REQUIRE(p.dctx && p.dQ && p.dK && p.dV && p.ds && p.pd, "fb_attn_bwd_f32: null gradient / workspace pointer");
  FB_REQUIRE(!p.pqk || p.dpqk, "fb_attn_bwd_f32: dpqk required with relative attention");
  const long long rows = (long long)p.B * p.h * p.L;
  const size_t smem = 4 * (128 + p.L) * sizeof(float);
  attn_f32_row_kernel<true><<<(unsigned)((rows + 3) / 4), 128, smem, ST(stream)>>>(p);
  count_launch();
  attn_f32_bwd_key_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, ST(stream)>>>(p);
  count_launch();
  if (p.pqk) {
    const int n = (2 * p.L - 1) * p.h;
    attn_f32_bwd_rel_kernel<<<(n + 3) / 4, 128, 0, ST(stream)>>>(p);
    count_launch();
  }
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
