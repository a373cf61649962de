// HBM-bound row kernels of the FrozenBiLM hot path: LayerNorm forward/backward (with the row-mask,
// dropout and bf16-operand copies fused), embedding/row gathers, conv im2col/col2im, bias-gradient
// column sums, cross-entropy over the vocabulary and the fused clip+Adam update.
// Every kernel is a single pass over its operands with 16-byte vector accesses; the roofline that
// bounds them is HBM bandwidth (bytes listed per kernel in DESIGN.md).
#include <algorithm>
#include <cstring>

#include "../../include/frozenbilm_b200.h"
#include "fb_common.cuh"

namespace fb {

static inline DropSpec drop_from(const fb_dropout* d) {
  if (!d) return make_drop(0.0f, 0, 0);
  return make_drop(d->p, d->seed, d->offset);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// keep/scale factor for the 4 consecutive columns starting at col (col % 4 == 0)
__device__ __forceinline__ void drop_factors4(const DropSpec& d, uint32_t row, int col, float (&f)[4]) {
  const uint32_t keep = drop_keep8(d, row, (uint32_t)col >> 3);
  const int sh = col & 4;
#pragma unroll
  for (int e = 0; e < 4; ++e) f[e] = ((keep >> (sh + e)) & 1) ? d.inv_keep : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row, the row lives in registers (H = 128 * NV, NV <= 16).
// y = LN(x) * rowmask, then dropout.  Replaces torch LayerNorm + "* mask" + StableDropout
// (deberta.py:1043-1054, 259, 333, 405-417, 474-478, 1547).
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x, int M, const float* __restrict__ gamma, const float* __restrict__ beta,
              float eps, const uint8_t* __restrict__ rowmask, DropSpec drop, float* __restrict__ y32,
              __nv_bfloat16* __restrict__ y16, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_prologue();
  constexpr int H = NV * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * H);
  float4 v[NV];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k] = __ldg(xr + lane + 32 * k);
    s += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  const float mean = warp_sum(s) * (1.0f / H);
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / H) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float rm = rowmask ? (rowmask[row] ? 1.0f : 0.0f) : 1.0f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int col = (lane + 32 * k) * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + col));
    float o[4] = {((v[k].x - mean) * rstd * g.x + bt.x) * rm, ((v[k].y - mean) * rstd * g.y + bt.y) * rm,
                  ((v[k].z - mean) * rstd * g.z + bt.z) * rm, ((v[k].w - mean) * rstd * g.w + bt.w) * rm};
    if (drop.thresh16) {
      float f[4];
      drop_factors4(drop, (uint32_t)row, col, f);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] *= f[e];
    }
    if (y32) *reinterpret_cast<float4*>(y32 + (long long)row * H + col) = make_float4(o[0], o[1], o[2], o[3]);
    if (y16)
      *reinterpret_cast<uint2*>(y16 + (long long)row * H + col) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  dy_eff = dropout_bwd(dy) * rowmask; dx = rstd * (g - mean(g) - xhat * mean(g xhat)),
// g = dy_eff * gamma.  dgamma/dbeta are accumulated per CTA in registers over a grid-stride loop and
// flushed with one atomic per column per CTA.  dx is written as fp32 (residual stream) and/or as a bf16
// copy with the UPSTREAM hidden dropout's backward applied (the operand of the next dX/dW GEMMs).
// ---------------------------------------------------------------------------------------------
template <int NV, int WPR>
__global__ void __launch_bounds__(256, WPR == 4 ? 2 : 1)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ gamma, int M,
              const uint8_t* __restrict__ rowmask, DropSpec out_drop, DropSpec in_drop, float* __restrict__ dx32,
              __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma, float* __restrict__ dbeta,
              float* __restrict__ dx16_colsum) {
  pdl_prologue();
  // WPR warps per row (32 WPR lanes x NV/WPR float4 each), persistent CTAs, the row read ONCE: x and dy of the
  // current row live in registers (reused for xhat and g = dy_eff * gamma), the NEXT row of the group is already in
  // flight in a second register set while this one is reduced and stored, so HBM keeps streaming across the row
  // reduction.  WPR = 4 (H a multiple of 512): ~110 registers, two CTAs = 16 warps per SM -- the kernel is bound by
  // instruction latency (dropout hash + ~25 FP32 operations per float4), not by the second read of the row that the
  // two-pass version did (measured: 2 warps per row at 255 registers, 8 warps per SM, ran no faster than two-pass).
  constexpr int H = NV * 128;
  constexpr int NK = NV / WPR;          // float4 per thread and operand
  constexpr int LPR = 32 * WPR;         // lanes per row
  constexpr int GROUPS = 8 / WPR;       // rows per CTA and step
  static_assert(NV % WPR == 0 && (WPR == 2 || WPR == 4), "row split");
  __shared__ float red[GROUPS][LPR * 4];
  __shared__ float rowsum[2][GROUPS][WPR][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp / WPR, part = warp % WPR;
  const int lrow = part * 32 + lane;
  float4 ag[NK], ab[NK], ac[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    ag[k] = make_float4(0, 0, 0, 0); ab[k] = make_float4(0, 0, 0, 0); ac[k] = make_float4(0, 0, 0, 0);
  }
  const int stride = gridDim.x * GROUPS;
  float4 xn[NK], dn[NK];
  auto fetch = [&](int r) {
    if (r < M) {
      const float4* xr = reinterpret_cast<const float4*>(x + (long long)r * H);
      const float4* dr = reinterpret_cast<const float4*>(dy + (long long)r * H);
#pragma unroll
      for (int k = 0; k < NK; ++k) { xn[k] = __ldg(xr + lrow + LPR * k); dn[k] = __ldg(dr + lrow + LPR * k); }
    }
  };
  int row = blockIdx.x * GROUPS + grp;
  fetch(row);
  for (int it = 0; row < M; row += stride, ++it) {     // the warps of a group share the row: the loop is group-uniform
    float4 xv[NK], dv[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) { xv[k] = xn[k]; dv[k] = dn[k]; }
    fetch(row + stride);
    const float mean = mean_in[row], rstd = rstd_in[row];
    const float rm = rowmask ? (rowmask[row] ? 1.0f : 0.0f) : 1.0f;
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int col = (lrow + LPR * k) * 4;
      float4 d = dv[k];
      if (out_drop.thresh16) {
        float f[4];
        drop_factors4(out_drop, (uint32_t)row, col, f);
        d.x *= f[0]; d.y *= f[1]; d.z *= f[2]; d.w *= f[3];
      }
      d.x *= rm; d.y *= rm; d.z *= rm; d.w *= rm;
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 xh = make_float4((xv[k].x - mean) * rstd, (xv[k].y - mean) * rstd, (xv[k].z - mean) * rstd,
                                    (xv[k].w - mean) * rstd);
      ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
      ag[k].x += d.x * xh.x; ag[k].y += d.y * xh.y; ag[k].z += d.z * xh.z; ag[k].w += d.w * xh.w;
      const float4 g = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += g.x + g.y + g.z + g.w;
      s2 += g.x * xh.x + g.y * xh.y + g.z * xh.z + g.w * xh.w;
      xv[k] = xh; dv[k] = g;          // keep xhat and g for the output pass
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    float (*rs)[2] = rowsum[it & 1][grp];              // double-buffered: one barrier per row
    if (lane == 0) { rs[part][0] = s1; rs[part][1] = s2; }
    named_bar_sync(1 + grp, LPR);
    float m1 = 0.0f, m2 = 0.0f;
#pragma unroll
    for (int w = 0; w < WPR; ++w) { m1 += rs[w][0]; m2 += rs[w][1]; }
    m1 *= (1.0f / H); m2 *= (1.0f / H);
    if (!dx32 && !dx16) continue;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int col = (lrow + LPR * k) * 4;
      float o[4] = {rstd * (dv[k].x - m1 - xv[k].x * m2), rstd * (dv[k].y - m1 - xv[k].y * m2),
                    rstd * (dv[k].z - m1 - xv[k].z * m2), rstd * (dv[k].w - m1 - xv[k].w * m2)};
      if (dx32) *reinterpret_cast<float4*>(dx32 + (long long)row * H + col) = make_float4(o[0], o[1], o[2], o[3]);
      if (dx16) {
        if (in_drop.thresh16) {
          float f[4];
          drop_factors4(in_drop, (uint32_t)row, col, f);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] *= f[e];
        }
        const uint2 q = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
        *reinterpret_cast<uint2*>(dx16 + (long long)row * H + col) = q;
        if (dx16_colsum) {   // sums of the ROUNDED values, i.e. exactly colsum(dx16)
          const float2 r0 = unpack_bf16(q.x), r1 = unpack_bf16(q.y);
          ac[k].x += r0.x; ac[k].y += r0.y; ac[k].z += r1.x; ac[k].w += r1.y;
        }
      }
    }
  }
  // cross-group reduction of the per-lane column partials: 4 LPR columns (one k of the row group) at a time
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    float* dst = pass == 0 ? dgamma : (pass == 1 ? dbeta : dx16_colsum);
    if (!dst) continue;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const float4 a = pass == 0 ? ag[k] : (pass == 1 ? ab[k] : ac[k]);
      __syncthreads();
      red[grp][lrow * 4 + 0] = a.x; red[grp][lrow * 4 + 1] = a.y;
      red[grp][lrow * 4 + 2] = a.z; red[grp][lrow * 4 + 3] = a.w;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < LPR * 4 / 256; ++j) {
        float v = 0.0f;
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) v += red[gi][j * 256 + threadIdx.x];
        atomicAdd(dst + k * LPR * 4 + j * 256 + threadIdx.x, v);
      }
    }
  }
}

#define FB_DISPATCH_NV(H, CALL)                                         \
  switch ((H) / 128) {                                                  \
    case 2: { constexpr int NV = 2; CALL; } break;                      \
    case 6: { constexpr int NV = 6; CALL; } break;                      \
    case 8: { constexpr int NV = 8; CALL; } break;                      \
    case 12: { constexpr int NV = 12; CALL; } break;                    \
    case 16: { constexpr int NV = 16; CALL; } break;                    \
    default:                                                            \
      fb::set_last_error("hidden size %d unsupported (need 256/768/1024/1536/2048)", (H)); \
      return -1;                                                        \
  }

int layernorm_fwd(const float* x, int M, int H, const float* gamma, const float* beta, float eps,
                  const uint8_t* rowmask, const fb_dropout* out_drop, float* y32, void* y16, float* mean,
                  float* rstd, cudaStream_t st) {
  FB_REQUIRE(x && gamma && beta && M > 0 && H % 128 == 0, "fb_layernorm_fwd: bad args");
  const DropSpec d = drop_from(out_drop);
  const int grid = (M + 7) / 8;
  cudaError_t le = cudaSuccess;
  FB_DISPATCH_NV(H, (le = launch_pdl(ln_fwd_kernel<NV>, dim3(grid), dim3(256), 0, st, x, M, gamma, beta, eps, rowmask, d, y32,
                                     reinterpret_cast<__nv_bfloat16*>(y16), mean, rstd)));
  FB_CHECK_CUDA(le);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, int M,
                  int H, const uint8_t* rowmask, const fb_dropout* out_drop, const fb_dropout* in_drop, float* dx32,
                  void* dx16, float* dgamma, float* dbeta, float* dx16_colsum, cudaStream_t st) {
  FB_REQUIRE(dy && x && mean && rstd && gamma && M > 0 && H % 128 == 0, "fb_layernorm_bwd: bad args");
  const DropSpec od = drop_from(out_drop), id = drop_from(in_drop);
  cudaError_t le = cudaSuccess;
  if (H % 512 == 0) {
    // four warps per row, two rows per CTA and step, two persistent CTAs per SM
    const int grid = std::max(1, std::min((M + 1) / 2, 2 * num_sms()));
    FB_DISPATCH_NV(H, (le = launch_pdl(ln_bwd_kernel<(NV % 4 == 0 ? NV : 4), 4>, dim3(grid), dim3(256), 0, st, dy, x, mean, rstd,
                                       gamma, M, rowmask, od, id, dx32, reinterpret_cast<__nv_bfloat16*>(dx16), dgamma,
                                       dbeta, dx16_colsum)));
  } else {
    const int grid = std::max(1, std::min((M + 3) / 4, num_sms()));
    FB_DISPATCH_NV(H, (le = launch_pdl(ln_bwd_kernel<NV, 2>, dim3(grid), dim3(256), 0, st, dy, x, mean, rstd, gamma, M, rowmask,
                                       od, id, dx32, reinterpret_cast<__nv_bfloat16*>(dx16), dgamma, dbeta, dx16_colsum)));
  }
  FB_CHECK_CUDA(le);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// row move with dtype conversion: out[idx_out[r], :] = in[idx_in[r], :]  (a null index array means r).
// Negative indices skip the row.  Used for the word-embedding lookup (deberta.py:1013), the
// labelled positions of the MLM head, and the video rows of the embedding gradient.
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void rows_kernel(const TI* __restrict__ in, long long ld_in, const long long* __restrict__ idx_in,
                            const long long* __restrict__ idx_out, int ncols, TO* __restrict__ out,
                            long long ld_out) {
  pdl_prologue();
  const int r = blockIdx.x;
  const long long ji = idx_in ? idx_in[r] : r, jo = idx_out ? idx_out[r] : r;
  if (ji < 0 || jo < 0) return;
  const TI* src = in + ji * ld_in;
  TO* dst = out + jo * ld_out;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) {
    float v;
    if constexpr (sizeof(TI) == 2) v = __bfloat162float(src[c]); else v = src[c];
    if constexpr (sizeof(TO) == 2) dst[c] = __float2bfloat16_rn(v); else dst[c] = v;
  }
}

int move_rows(const void* in, long long ld_in, int in_bf16, const long long* idx_in, const long long* idx_out, int R,
              int ncols, void* out, long long ld_out, int out_bf16, cudaStream_t st) {
  FB_REQUIRE(in && out && R > 0 && ncols > 0, "fb_move_rows: bad args");
  const int threads = 256;
  if (in_bf16 && out_bf16)
    FB_CHECK_CUDA(launch_pdl(rows_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(R), dim3(threads), 0, st, (const __nv_bfloat16*)in,
                             ld_in, idx_in, idx_out, ncols, (__nv_bfloat16*)out, ld_out));
  else if (in_bf16 && !out_bf16)
    FB_CHECK_CUDA(launch_pdl(rows_kernel<__nv_bfloat16, float>, dim3(R), dim3(threads), 0, st, (const __nv_bfloat16*)in, ld_in,
                             idx_in, idx_out, ncols, (float*)out, ld_out));
  else if (!in_bf16 && out_bf16)
    FB_CHECK_CUDA(launch_pdl(rows_kernel<float, __nv_bfloat16>, dim3(R), dim3(threads), 0, st, (const float*)in, ld_in, idx_in,
                             idx_out, ncols, (__nv_bfloat16*)out, ld_out));
  else
    FB_CHECK_CUDA(launch_pdl(rows_kernel<float, float>, dim3(R), dim3(threads), 0, st, (const float*)in, ld_in, idx_in, idx_out,
                             ncols, (float*)out, ld_out));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Conv1d(k=3, pad=1) over the sequence as a K = 3H GEMM (deberta.py:395-400): im2col of the bf16
// embedding output, zero padded per sample; col2im is its transpose for the backward.
// ---------------------------------------------------------------------------------------------
__global__ void im2col3_kernel(const __nv_bfloat16* __restrict__ x, int L, int H, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const int row = blockIdx.x;  // b * L + t
  const int t = row % L;
  const int hv = H / 8;
  for (int c = threadIdx.x; c < 3 * hv; c += blockDim.x) {
    const int k = c / hv, cc = c % hv;
    const int tt = t + k - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tt >= 0 && tt < L) v = __ldg(reinterpret_cast<const uint4*>(x + (long long)(row + k - 1) * H) + cc);
    reinterpret_cast<uint4*>(out + (long long)row * 3 * H + (long long)k * H)[cc] = v;
  }
}
__global__ void col2im3_kernel(const __nv_bfloat16* __restrict__ dcol, int L, int H, float* __restrict__ dx,
                               int accumulate) {
  pdl_prologue();
  const int row = blockIdx.x;
  const int t = row % L;
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    // dx[t] = dcol[t+1][k=0] + dcol[t][k=1] + dcol[t-1][k=2]
    float v = __bfloat162float(dcol[(long long)row * 3 * H + H + c]);
    if (t + 1 < L) v += __bfloat162float(dcol[(long long)(row + 1) * 3 * H + c]);
    if (t - 1 >= 0) v += __bfloat162float(dcol[(long long)(row - 1) * 3 * H + 2 * H + c]);
    float* d = dx + (long long)row * H + c;
    *d = accumulate ? *d + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// column sums of a bf16 [M, N] matrix into fp32 [N] (accumulating): bias gradients.
// ---------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, int M, int N, int rows_per_cta,
                              float* __restrict__ out) {
  pdl_prologue();
  const int c2 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (c2 >= N) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float a = 0.0f, b = 0.0f;
  if (c2 + 1 < N) {
    for (int r = r0; r < r1; ++r) {
      const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (long long)r * ld + c2));
      a += f.x; b += f.y;
    }
    atomicAdd(out + c2, a);
    atomicAdd(out + c2 + 1, b);
  } else {
    for (int r = r0; r < r1; ++r) a += __bfloat162float(x[(long long)r * ld + c2]);
    atomicAdd(out + c2, a);
  }
}

// ---------------------------------------------------------------------------------------------
// cross-entropy over fp32 logits [P, V] (CrossEntropyLoss, deberta.py:1481-1488): per-row
// log-sum-exp; loss_sum += sum_rows (lse - logit[label]) * inv_count; dlogits (bf16, optional)
// = (softmax - onehot) * grad_scale.  One CTA per row.
// ---------------------------------------------------------------------------------------------
// Two passes per row: (1) online max / sum-of-exp with 16-byte loads, (2) gradient.  One 1024-thread CTA per SM at a
// time (the launch asks for > half of the shared memory), so the <= 148 rows in flight (0.5 MB each at V = 128100)
// stay in the 126 MB L2 and pass 2 does not go back to HBM.  (The first version made three scalar passes with four
// rows per SM in flight: 2.9 GB of HBM traffic, 1.14 ms; this one moves ~1.3 GB.)
static constexpr int CE_THREADS = 1024;
static constexpr int CE_SMEM = 120 * 1024;   // occupancy limiter only
__device__ __forceinline__ void ce_combine(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  const float a = (m == -INFINITY) ? 0.0f : __expf(m - mn), b = (m2 == -INFINITY) ? 0.0f : __expf(m2 - mn);
  s = s * a + s2 * b;
  m = mn;
}
__device__ __forceinline__ void ce_store4(__nv_bfloat16* d, const float (&g)[4]) {
  *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]));
}
__device__ __forceinline__ void ce_store4(float* d, const float (&g)[4]) {
  d[0] = g[0]; d[1] = g[1]; d[2] = g[2]; d[3] = g[3];   // rows need not be 16-byte aligned
}
__device__ __forceinline__ void ce_store1(__nv_bfloat16* d, float g) { *d = __float2bfloat16_rn(g); }
__device__ __forceinline__ void ce_store1(float* d, float g) { *d = g; }
// GT = gradient type: bf16 (training path) or float (fp32-class instantiation, fb_cross_entropy_f32)
template <typename GT>
__global__ void __launch_bounds__(CE_THREADS, 1)
ce_kernel(const float* __restrict__ logits, long long ld, int V, const long long* __restrict__ labels,
          float* __restrict__ loss_sum, float loss_scale, GT* __restrict__ dlogits, long long ldd,
          float grad_scale) {
  pdl_prologue();
  __shared__ float sm[32], ss[32];
  __shared__ float sbc[2];
  const int row = blockIdx.x;
  const float* lr = logits + (long long)row * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = ((reinterpret_cast<uintptr_t>(lr) & 15) == 0);
  const int V4 = vec ? (V >> 2) : 0;
  float m = -INFINITY, s = 0.0f;
  for (int c4 = threadIdx.x; c4 < V4; c4 += CE_THREADS) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(lr) + c4);
    const float mx = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
    if (mx > m) { s *= __expf(m - mx); m = mx; }   // exp(-inf) = 0 on the first chunk
    s += __expf(x.x - m) + __expf(x.y - m) + __expf(x.z - m) + __expf(x.w - m);
  }
  for (int c = V4 * 4 + threadIdx.x; c < V; c += CE_THREADS) {
    const float x = lr[c];
    if (x > m) { s *= __expf(m - x); m = x; }
    s += __expf(x - m);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, s, off);
    ce_combine(m, s, m2, s2);
  }
  if (lane == 0) { sm[warp] = m; ss[warp] = s; }
  __syncthreads();
  if (warp == 0) {
    m = sm[lane]; s = ss[lane];   // CE_THREADS / 32 = 32 warps
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, s, off);
      ce_combine(m, s, m2, s2);
    }
    if (lane == 0) { sbc[0] = m; sbc[1] = s; }
  }
  __syncthreads();
  const float mx = sbc[0], total = sbc[1];
  const long long lab = labels[row];
  if (threadIdx.x == 0) atomicAdd(loss_sum, (mx + logf(total) - lr[lab]) * loss_scale);
  if (dlogits) {
    const float inv = grad_scale / total;
    GT* dr = dlogits + (long long)row * ldd;
    const bool vecd = vec && ((reinterpret_cast<uintptr_t>(dr) & 7) == 0);
    const int V4d = vecd ? (V >> 2) : 0;
    for (int c4 = threadIdx.x; c4 < V4d; c4 += CE_THREADS) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(lr) + c4);
      float g[4] = {__expf(x.x - mx) * inv, __expf(x.y - mx) * inv, __expf(x.z - mx) * inv, __expf(x.w - mx) * inv};
      const long long d = lab - 4LL * c4;
      if (d >= 0 && d < 4) g[d] -= grad_scale;
      ce_store4(dr + 4 * c4, g);
    }
    for (int c = V4d * 4 + threadIdx.x; c < V; c += CE_THREADS) {
      float g = __expf(lr[c] - mx) * inv;
      if (c == lab) g -= grad_scale;
      ce_store1(dr + c, g);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dropout of an fp32 matrix into bf16 and/or fp32 (optionally accumulating): the per-layer-pass
// dropout of the relative embeddings (deberta.py:779) and its backward.
// ---------------------------------------------------------------------------------------------
__global__ void dropout_kernel(const float* __restrict__ x, int M, int N, DropSpec d, float* __restrict__ out32,
                               int accumulate, __nv_bfloat16* __restrict__ out16) {
  pdl_prologue();
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= (long long)M * N) return;
  const int row = (int)(i4 / N), col = (int)(i4 % N);
  float4 v = *reinterpret_cast<const float4*>(x + i4);
  float o[4] = {v.x, v.y, v.z, v.w};
  if (d.thresh16) {
    float f[4];
    drop_factors4(d, (uint32_t)row, col, f);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] *= f[e];
  }
  if (out32) {
    float4* p = reinterpret_cast<float4*>(out32 + i4);
    float4 w = make_float4(o[0], o[1], o[2], o[3]);
    if (accumulate) { const float4 q = *p; w.x += q.x; w.y += q.y; w.z += q.z; w.w += q.w; }
    *p = w;
  }
  if (out16) *reinterpret_cast<uint2*>(out16 + i4) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
}

// The same dropout for ALL layer passes of a step in one launch (the per-pass positional chain, engine.py):
//   expand (out16 != null): out16[b][r][c] = dropout_b(x[r][c])            x shared by the nb passes
//   reduce (out32 != null): out32[r][c]   += sum_b dropout_b(x[b][r][c])   (the backward: dR accumulates over passes)
// pass b uses the site offset d.off_lo + b * off_stride and element id (r, c): bit-identical to nb calls of
// dropout_kernel with those offsets.
__global__ void dropout_batched_kernel(const float* __restrict__ x, int M, int N, int nb, DropSpec d, uint32_t off_stride,
                                       float* __restrict__ out32, __nv_bfloat16* __restrict__ out16) {
  pdl_prologue();
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long MN = (long long)M * N;
  if (i4 >= MN) return;
  const int row = (int)(i4 / N), col = (int)(i4 % N);
  const bool reduce = out32 != nullptr;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!reduce) v = *reinterpret_cast<const float4*>(x + i4);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < nb; ++b) {
    if (reduce) v = *reinterpret_cast<const float4*>(x + b * MN + i4);
    float o[4] = {v.x, v.y, v.z, v.w};
    if (d.thresh16) {
      DropSpec db = d;
      db.off_lo = d.off_lo + (uint32_t)b * off_stride;
      float f[4];
      drop_factors4(db, (uint32_t)row, col, f);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] *= f[e];
    }
    if (reduce) {
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e] += o[e];
    } else {
      *reinterpret_cast<uint2*>(out16 + b * MN + i4) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
    }
  }
  if (reduce) {
    float4* p = reinterpret_cast<float4*>(out32 + i4);
    float4 q = *p;
    q.x += acc[0]; q.y += acc[1]; q.z += acc[2]; q.w += acc[3];
    *p = q;
  }
}

__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ pre, int M, int N,
                                const uint8_t* __restrict__ rowmask, DropSpec d, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= (long long)M * N) return;
  const int row = (int)(i4 / N), col = (int)(i4 % N);
  const float4 g = *reinterpret_cast<const float4*>(dy + i4);
  const uint2 pr = *reinterpret_cast<const uint2*>(pre + i4);
  const float2 p0 = unpack_bf16(pr.x), p1 = unpack_bf16(pr.y);
  float o[4] = {g.x * gelu_erf_grad(p0.x), g.y * gelu_erf_grad(p0.y), g.z * gelu_erf_grad(p1.x),
                g.w * gelu_erf_grad(p1.y)};
  if (d.thresh16) {
    float f[4];
    drop_factors4(d, (uint32_t)row, col, f);
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] *= f[e];
  }
  if (rowmask && rowmask[row] == 0) { o[0] = o[1] = o[2] = o[3] = 0.0f; }
  *reinterpret_cast<uint2*>(out + i4) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
}
__global__ void add_rowbcast_kernel(const float* __restrict__ a, const float* __restrict__ b, long long period,
                                    long long n4, float* __restrict__ out32, __nv_bfloat16* __restrict__ out16) {
  pdl_prologue();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = reinterpret_cast<const float4*>(a)[i];
  const float4 y = reinterpret_cast<const float4*>(b)[(i * 4 % period) / 4];
  const float4 o = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  reinterpret_cast<float4*>(out32)[i] = o;
  if (out16) reinterpret_cast<uint2*>(out16)[i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
}

// ---------------------------------------------------------------------------------------------
// fused clip_grad_norm_ + Adam over the flat trainable bucket (main.py:82-84; betas/eps from
// args.py:246-262), also refreshing the bf16 operand copy of the updated parameters.
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float sred[8];
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    atomicAdd(out, t);
  }
}
// Deterministic variant: every CTA writes its partial to scratch[blockIdx.x]; the LAST CTA to finish (ticket counter
// at scratch[gridDim.x], reset for the next launch) adds them in index order.  Identical gradients therefore give a
// bit-identical norm on every data-parallel rank, so the clip factor -- and with it every parameter -- stays
// bit-identical across replicas (float atomics would add the partials in arrival order).
__global__ void sumsq_det_kernel(const float* __restrict__ g, long long n, float* __restrict__ out,
                                 float* __restrict__ scratch) {
  pdl_prologue();
  __shared__ float sred[8];
  __shared__ bool last;
  float s = 0.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    scratch[blockIdx.x] = t;
    __threadfence();
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + gridDim.x);
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.0f;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(scratch + b);
    *out = t;
    *reinterpret_cast<unsigned int*>(scratch + gridDim.x) = 0u;
  }
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2, const float* __restrict__ gnorm_sq, float max_norm, float grad_scale,
                            __nv_bfloat16* __restrict__ p16) {
  pdl_prologue();
  float coef = grad_scale;
  if (gnorm_sq && max_norm > 0.0f) {
    const float norm = sqrtf(*gnorm_sq) * grad_scale;
    coef *= fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float upd = (mi / bc1) / (sqrtf(vi / bc2) + eps);
    const float pi = p[i] - lr * upd;
    p[i] = pi;
    if (p16) p16[i] = __float2bfloat16_rn(pi);
  }
}
__global__ void cast_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ out) {
  pdl_prologue();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i]);
}

// ---------------------------------------------------------------------------------------------
// softmax over the answer vocabulary + top-k at the [MASK] rows (videoqa.py:164-170: logits.softmax(-1),
// torch.topk(logits, k); mc.py:168-172: softmax(-1)[:, 0]).  One CTA per row; V up to a few thousand answers.
// probs_out (nullable): full softmax [P, V].  top_p / top_i: the k largest probabilities (descending; ties -> the
// lower index, like torch.topk on a strictly ordered input) and their answer ids.
// ---------------------------------------------------------------------------------------------
static constexpr int STK_THREADS = 256;
__global__ void __launch_bounds__(STK_THREADS)
softmax_topk_kernel(const float* __restrict__ logits, long long ld, int V, int k, float* __restrict__ probs_out,
                    long long ldp, float* __restrict__ top_p, long long* __restrict__ top_i) {
  pdl_prologue();
  extern __shared__ float srow[];                 // V probabilities
  __shared__ float rv[STK_THREADS / 32];
  __shared__ int ri[STK_THREADS / 32];
  __shared__ float bc;
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* lr = logits + (long long)row * ld;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < V; c += STK_THREADS) { const float x = lr[c]; srow[c] = x; m = fmaxf(m, x); }
  m = warp_max(m);
  if (lane == 0) rv[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) { float t = rv[0]; for (int w = 1; w < STK_THREADS / 32; ++w) t = fmaxf(t, rv[w]); bc = t; }
  __syncthreads();
  m = bc;
  float s = 0.0f;
  for (int c = threadIdx.x; c < V; c += STK_THREADS) { const float e = expf(srow[c] - m); srow[c] = e; s += e; }
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) rv[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.0f; for (int w = 0; w < STK_THREADS / 32; ++w) t += rv[w]; bc = 1.0f / t; }
  __syncthreads();
  const float inv = bc;
  for (int c = threadIdx.x; c < V; c += STK_THREADS) {
    const float pr = srow[c] * inv;
    srow[c] = pr;
    if (probs_out) probs_out[(long long)row * ldp + c] = pr;
  }
  __syncthreads();
  for (int t = 0; t < k; ++t) {
    float bv = -1.0f; int bi = 0x7fffffff;
    for (int c = threadIdx.x; c < V; c += STK_THREADS) {
      const float x = srow[c];
      if (x > bv) { bv = x; bi = c; }              // strided scan: lower index wins inside a thread
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { rv[warp] = bv; ri[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float fv = rv[0]; int fi = ri[0];
      for (int w = 1; w < STK_THREADS / 32; ++w)
        if (rv[w] > fv || (rv[w] == fv && ri[w] < fi)) { fv = rv[w]; fi = ri[w]; }
      top_p[(long long)row * k + t] = fv;
      top_i[(long long)row * k + t] = fi;
      srow[fi] = -1.0f;                            // remove it from the next round
    }
    __syncthreads();
  }
}

}  // namespace fb

using namespace fb;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int fb_layernorm_fwd(const float* x, int M, int H, const float* gamma, const float* beta, float eps,
                     const unsigned char* rowmask, const fb_dropout* out_drop, float* y_f32, void* y_bf16,
                     float* mean, float* rstd, void* stream) {
  return layernorm_fwd(x, M, H, gamma, beta, eps, rowmask, out_drop, y_f32, y_bf16, mean, rstd, ST(stream));
}
int fb_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     int M, int H, const unsigned char* rowmask, const fb_dropout* out_drop,
                     const fb_dropout* in_drop, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                     float* dx_bf16_colsum, void* stream) {
  return layernorm_bwd(dy, x, mean, rstd, gamma, M, H, rowmask, out_drop, in_drop, dx_f32, dx_bf16, dgamma, dbeta,
                       dx_bf16_colsum, ST(stream));
}
int fb_move_rows(const void* in, long long ld_in, int in_bf16, const long long* idx_in, const long long* idx_out,
                 int R, int ncols, void* out, long long ld_out, int out_bf16, void* stream) {
  return move_rows(in, ld_in, in_bf16, idx_in, idx_out, R, ncols, out, ld_out, out_bf16, ST(stream));
}
/* out_bf16 = dy * gelu'(pre) [* dropout_bwd(drop)] [* rowmask]: backward of GELU at the conv branch
 * (deberta.py:401-403) and the MLM head transform (deberta.py:1545-1546). */
int fb_gelu_bwd(const float* dy, const void* pre_bf16, int M, int N, const unsigned char* rowmask,
                const fb_dropout* drop, void* out_bf16, void* stream) {
  FB_REQUIRE(dy && pre_bf16 && out_bf16 && M > 0 && N % 8 == 0, "fb_gelu_bwd: bad args");
  const long long n4 = (long long)M * N / 4;
  FB_CHECK_CUDA(launch_pdl(gelu_bwd_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, ST(stream), dy,
                           (const __nv_bfloat16*)pre_bf16, M, N, rowmask, drop_from(drop), (__nv_bfloat16*)out_bf16));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
/* out = a + b (row-broadcast of b over the batch): z = position_embeddings + h[-2] of the enhanced mask
 * decoder (deberta.py:1394-1396); writes fp32 and bf16 copies.  a: [B*L, H], b: [L, H]. */
int fb_add_rowbcast(const float* a, const float* b, int B, int L, int H, float* out_f32, void* out_bf16,
                    void* stream) {
  FB_REQUIRE(a && b && out_f32 && B > 0 && L > 0 && H % 4 == 0, "fb_add_rowbcast: bad args");
  const long long n4 = (long long)B * L * H / 4;
  FB_CHECK_CUDA(launch_pdl(add_rowbcast_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, ST(stream), a, b,
                           (long long)L * H, n4, out_f32, (__nv_bfloat16*)out_bf16));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_conv_im2col(const void* x_bf16, int B, int L, int H, void* out_bf16, void* stream) {
  FB_REQUIRE(x_bf16 && out_bf16 && B > 0 && L > 0 && H % 8 == 0, "fb_conv_im2col: bad args");
  FB_CHECK_CUDA(launch_pdl(im2col3_kernel, dim3(B * L), dim3(256), 0, ST(stream), (const __nv_bfloat16*)x_bf16, L, H,
                           (__nv_bfloat16*)out_bf16));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_conv_col2im(const void* dcol_bf16, int B, int L, int H, float* dx, int accumulate, void* stream) {
  FB_REQUIRE(dcol_bf16 && dx && B > 0 && L > 0, "fb_conv_col2im: bad args");
  FB_CHECK_CUDA(launch_pdl(col2im3_kernel, dim3(B * L), dim3(256), 0, ST(stream), (const __nv_bfloat16*)dcol_bf16, L, H, dx,
                           accumulate));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_colsum_bf16(const void* x, long long ld, int M, int N, float* out, void* stream) {
  FB_REQUIRE(x && out && M > 0 && N > 0 && (ld % 2 == 0), "fb_colsum_bf16: bad args");
  const int rows_per_cta = 64;
  dim3 grid(((N + 1) / 2 + 127) / 128, (M + rows_per_cta - 1) / rows_per_cta);
  FB_CHECK_CUDA(launch_pdl(colsum_kernel, grid, dim3(128), 0, ST(stream), (const __nv_bfloat16*)x, ld, M, N, rows_per_cta, out));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_cross_entropy(const float* logits, long long ld, int P, int V, const long long* labels, float* loss_sum,
                     float loss_scale, void* dlogits_bf16, long long ldd, float grad_scale, void* stream) {
  FB_REQUIRE(logits && labels && loss_sum && P > 0 && V > 0, "fb_cross_entropy: bad args");
  static bool configured = false;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(fb::ce_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       fb::CE_SMEM));
    configured = true;
  }
  FB_CHECK_CUDA(launch_pdl(fb::ce_kernel<__nv_bfloat16>, dim3(P), dim3(fb::CE_THREADS), fb::CE_SMEM, ST(stream), logits, ld, V,
                           labels, loss_sum, loss_scale, (__nv_bfloat16*)dlogits_bf16, ldd, grad_scale));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_cross_entropy_f32(const float* logits, long long ld, int P, int V, const long long* labels, float* loss_sum,
                         float loss_scale, float* dlogits_f32, long long ldd, float grad_scale, void* stream) {
  FB_REQUIRE(logits && labels && loss_sum && P > 0 && V > 0, "fb_cross_entropy_f32: bad args");
  static bool configured = false;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(fb::ce_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb::CE_SMEM));
    configured = true;
  }
  fb::ce_kernel<float><<<P, fb::CE_THREADS, fb::CE_SMEM, ST(stream)>>>(logits, ld, V, labels, loss_sum, loss_scale,
                                                                       dlogits_f32, ldd, grad_scale);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_softmax_topk(const float* logits, long long ld, int P, int V, int k, float* probs, long long ldp, float* top_p,
                    long long* top_i, void* stream) {
  FB_REQUIRE(logits && P > 0 && V > 0 && k >= 0 && k <= V && (k == 0 || (top_p && top_i)), "fb_softmax_topk: bad args");
  FB_REQUIRE((size_t)V * 4 <= 200 * 1024, "fb_softmax_topk: answer vocabulary too large for one CTA (V = %d)", V);
  const size_t smem = (size_t)V * sizeof(float);
  if (smem > 48 * 1024)
    FB_CHECK_CUDA(cudaFuncSetAttribute(fb::softmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fb::softmax_topk_kernel<<<P, fb::STK_THREADS, smem, ST(stream)>>>(logits, ld, V, k, probs, ldp, top_p, top_i);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_dropout_apply(const float* x, int M, int N, const fb_dropout* drop, float* out_f32, int accumulate,
                     void* out_bf16, void* stream) {
  FB_REQUIRE(x && M > 0 && N > 0 && N % 8 == 0, "fb_dropout_apply: bad args (N must be a multiple of 8)");
  const long long n4 = (long long)M * N / 4;
  FB_CHECK_CUDA(launch_pdl(dropout_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, ST(stream), x, M, N,
                           drop_from(drop), out_f32, accumulate, (__nv_bfloat16*)out_bf16));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_dropout_batched(const float* x, int M, int N, int nb, const fb_dropout* drop, unsigned long long offset_stride,
                       float* out_f32_accum, void* out_bf16, void* stream) {
  FB_REQUIRE(x && M > 0 && N > 0 && N % 8 == 0 && nb > 0, "fb_dropout_batched: bad args (N must be a multiple of 8)");
  FB_REQUIRE((out_f32_accum != nullptr) != (out_bf16 != nullptr), "fb_dropout_batched: exactly one of the outputs");
  const long long n4 = (long long)M * N / 4;
  FB_CHECK_CUDA(launch_pdl(fb::dropout_batched_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, ST(stream), x, M, N,
                           nb, drop_from(drop), (uint32_t)offset_stride, out_f32_accum, (__nv_bfloat16*)out_bf16));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_sumsq(const float* g, long long n, float* out, void* stream) {
  FB_REQUIRE(g && out && n > 0, "fb_sumsq: bad args");
  const int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 4LL * num_sms()));
  sumsq_kernel<<<grid, 256, 0, ST(stream)>>>(g, n, out);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_sumsq_det(const float* g, long long n, float* out, float* scratch, int scratch_len, void* stream) {
  FB_REQUIRE(g && out && scratch && n > 0 && scratch_len >= 2, "fb_sumsq_det: bad args");
  const int grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>((n + 255) / 256, 4LL * num_sms()),
                                                                   scratch_len - 1));
  FB_CHECK_CUDA(launch_pdl(sumsq_det_kernel, dim3(grid), dim3(256), 0, ST(stream), g, n, out, scratch));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, int step, const float* gnorm_sq, float max_norm, float grad_scale, void* p_bf16,
                 void* stream) {
  FB_REQUIRE(p && g && m && v && n > 0 && step >= 1, "fb_adam_step: bad args");
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  const int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 8LL * num_sms()));
  FB_CHECK_CUDA(launch_pdl(adam_kernel, dim3(grid), dim3(256), 0, ST(stream), p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2,
                           gnorm_sq, max_norm, grad_scale, (__nv_bfloat16*)p_bf16));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int fb_cast_f32_bf16(const float* x, long long n, void* out, void* stream) {
  FB_REQUIRE(x && out && n > 0, "fb_cast_f32_bf16: bad args");
  const int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 8LL * num_sms()));
  FB_CHECK_CUDA(launch_pdl(cast_kernel, dim3(grid), dim3(256), 0, ST(stream), x, n, (__nv_bfloat16*)out));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
