// Fused disentangled attention (DeBERTa-V2, share_att_key, c2p|p2c) for sm_100a.
//
// Replaces DisentangledSelfAttention.forward + disentangled_attention_bias + XSoftmax + dropout
// (reference model/deberta.py:717-947, 100-138).  One CTA = one (batch, head, 128-query tile);
// it walks the keys in blocks of 128 with an online softmax:
//
//   S   = Q_tile . K_blk^T                      tcgen05.mma 128x128x64   -> TMEM cols [0,128)
//   X2  = K_blk  . PQband^T  (lane = key  b)     tcgen05.mma 128x256x64   -> TMEM cols [128,384)
//   X1  = Q_tile . PKband^T  (lane = query a)    tcgen05.mma 128x256x64   -> same columns, after X2
//   s[a,b] = (S[a,b] + X1[a, a-b+127] + X2[b, a-b+127]) * scale      (Toeplitz skew through smem)
//   O_blk = P . V_blk                            tcgen05.mma 128x64x128   -> TMEM cols [384,448)
//
// PKband / PQband are 256 consecutive rows of per-head tables indexed by RELATIVE POSITION
// (row = (i-j) + rel_off, already mapped through the log-bucket function), so the band of a tile is
// one TMA box and the exact and log-bucketed zones are handled uniformly.
// The positional terms are staged through shared memory in fp16 (11-bit mantissa, finer than the
// bf16 operands that produced them) because the skew needs a per-lane column shift that TMEM
// loads cannot express; scores, softmax statistics and the output accumulate in fp32.
#include <cuda_fp16.h>
#include <cstring>

#include "../../include/frozenbilm_b200.h"
#include "fb_common.cuh"

namespace fb {

static constexpr int AT_THREADS = 192;   // warps 0-3 softmax (thread = query row), 4 = TMA+MMA, 5 = TMEM alloc
static constexpr int T_PITCH = 136;      // halves per staging row: 272 B -> conflict-free LDS.128
static constexpr uint32_t OFF_Q = 0, OFF_K = 16384, OFF_V = 32768, OFF_PK = 49152, OFF_PQ = 81920,
                          OFF_P = 114688, OFF_T1 = 147456, OFF_T2 = OFF_T1 + 128 * T_PITCH * 2,
                          OFF_KMASK = OFF_T2 + 128 * T_PITCH * 2, OFF_BARS = OFF_KMASK + 128,
                          AT_SMEM = OFF_BARS + 128 + 1024;
static constexpr uint32_t TM_S = 0, TM_X = 128, TM_O = 384;

struct AttnParams {
  int B, L, h, nkb, rel_off;
  const uint8_t* mask;
  __nv_bfloat16* ctx; long long ldc;
  float* lse;
  float scale_log2;
  float inv_keep; uint32_t drop_thresh16;  // keep iff rnd16 >= thresh
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
};

// 8 Bernoulli(keep) decisions for columns [8*jgroup, 8*jgroup+8) of score row `rowid`.
__device__ __forceinline__ uint32_t dropout_keep8(uint32_t rowid, uint32_t jgroup, uint32_t thresh16,
                                                  uint32_t seed_lo, uint32_t seed_hi, uint32_t off_lo,
                                                  uint32_t off_hi) {
  const uint4 r = philox4x32(rowid, jgroup, off_lo, off_hi, seed_lo, seed_hi);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFF) >= thresh16) << 0; m |= ((r.x >> 16) >= thresh16) << 1;
  m |= ((r.y & 0xFFFF) >= thresh16) << 2; m |= ((r.y >> 16) >= thresh16) << 3;
  m |= ((r.z & 0xFFFF) >= thresh16) << 4; m |= ((r.z >> 16) >= thresh16) << 5;
  m |= ((r.w & 0xFFFF) >= thresh16) << 6; m |= ((r.w >> 16) >= thresh16) << 7;
  return m;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_pk,
                const __grid_constant__ CUtensorMap map_pq, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_mma1 = bars + 3;
  uint64_t* bar_xfree = bars + 4;
  uint64_t* bar_mma2 = bars + 5;
  uint64_t* bar_pready = bars + 6;
  uint64_t* bar_mma3 = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  __half* T1 = reinterpret_cast<__half*>(smem + OFF_T1);
  __half* T2 = reinterpret_cast<__half*>(smem + OFF_T2);
  uint8_t* kmask = smem + OFF_KMASK;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, hh = blockIdx.y, b = blockIdx.z;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_pk); tma_prefetch_desc(&map_pq);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    mbar_init(bar_mma1, 1); mbar_init(bar_xfree, 4); mbar_init(bar_mma2, 1);
    mbar_init(bar_pready, 4); mbar_init(bar_mma3, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (elect_one_sync()) {
      // ======================= TMA + MMA control thread =======================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_band = make_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t sQ = smem_u32(smem + OFF_Q), sK = smem_u32(smem + OFF_K), sV = smem_u32(smem + OFF_V),
                     sPK = smem_u32(smem + OFF_PK), sPQ = smem_u32(smem + OFF_PQ), sP = smem_u32(smem + OFF_P);
      auto load_k = [&](int kb) {
        const int j0 = kb * 128;
        mbar_arrive_expect_tx(bar_k, 16384 + 2 * 32768);
        tma_load_4d(smem + OFF_K, &map_k, bar_k, 0, hh, j0, b);
        const int band0 = i0 - j0 - 127 + p.rel_off;
        tma_load_3d(smem + OFF_PK, &map_pk, bar_k, 0, band0, hh);
        tma_load_3d(smem + OFF_PQ, &map_pq, bar_k, 0, band0, hh);
      };
      auto load_v = [&](int kb) {
        mbar_arrive_expect_tx(bar_v, 16384);
        tma_load_4d(smem + OFF_V, &map_v, bar_v, 0, hh, kb * 128, b);
      };
      mbar_arrive_expect_tx(bar_q, 16384);
      tma_load_4d(smem + OFF_Q, &map_q, bar_q, 0, hh, i0, b);
      load_k(0);
      load_v(0);
      mbar_wait(bar_q, 0);
      for (int kb = 0; kb < p.nkb; ++kb) {
        const uint32_t ph = kb & 1;
        mbar_wait(bar_k, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + TM_S, make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(sK + k * 32, 16, 1024), idesc_qk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + TM_X, make_smem_desc_sw128(sK + k * 32, 16, 1024),
                    make_smem_desc_sw128(sPQ + k * 32, 16, 1024), idesc_band, k > 0);
        umma_commit(bar_mma1);
        if (kb > 0) {
          // the previous block's P.V must retire before its V tile is overwritten
          mbar_wait(bar_mma3, (kb - 1) & 1);
          load_v(kb);
        }
        mbar_wait(bar_xfree, ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + TM_X, make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(sPK + k * 32, 16, 1024), idesc_band, k > 0);
        umma_commit(bar_mma2);
        mbar_wait(bar_mma2, ph);  // K / PK / PQ tiles are free again
        if (kb + 1 < p.nkb) load_k(kb + 1);
        mbar_wait(bar_pready, ph);
        tc_fence_after();
        mbar_wait(bar_v, ph);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem + TM_O, make_smem_desc_sw128(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(sV + k * 2048, 16, 1024), idesc_pv, k > 0);
        umma_commit(bar_mma3);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ======================= softmax warps: thread = query row =======================
    const int a = threadIdx.x;           // row in tile == TMEM lane
    const int i = i0 + a;
    const bool valid_q = (i < p.L) && (p.mask[(long long)b * p.L + i] != 0);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t rowid = static_cast<uint32_t>((b * p.h + hh) * p.L + i);
    const bool do_drop = p.drop_thresh16 != 0;
    float m_run = -INFINITY, l_run = 0.0f;
    float o[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) o[k] = 0.0f;

    for (int kb = 0; kb < p.nkb; ++kb) {
      const uint32_t ph = kb & 1;
      const int j0 = kb * 128;
      mbar_wait(bar_mma1, ph);
      tc_fence_after();
      {
        const int j = j0 + a;
        kmask[a] = (j < p.L) ? p.mask[(long long)b * p.L + j] : 0;
      }
      // ---- p2c: this thread owns KEY row a of X2; scatter its window into T2[query][key]
#pragma unroll 1
      for (int q = 0; q < 5; ++q) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + TM_X + (96 - 32 * warp + 32 * q), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int aa = 32 * q + e + lane - 31;
          if (aa >= 0 && aa < 128) T2[aa * T_PITCH + a] = __float2half_rn(__uint_as_float(v[e]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_xfree);
      mbar_wait(bar_mma2, ph);
      tc_fence_after();
      // ---- c2p: this thread owns QUERY row a of X1; un-skew its window into T1[query][key]
#pragma unroll 1
      for (int q = 0; q < 5; ++q) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + TM_X + (32 * warp + 32 * q), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int bb = lane + 127 - 32 * q - e;
          if (bb >= 0 && bb < 128) T1[a * T_PITCH + bb] = __float2half_rn(__uint_as_float(v[e]));
        }
      }
      named_bar_sync(1, 128);  // T2 / kmask were written by other threads
      // ---- pass A: combine the three terms, mask, row max; park s in the S columns of TMEM
      float mblk = -INFINITY;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + TM_S + 32 * q, v);
        tmem_ld_wait();
        const uint4* t1 = reinterpret_cast<const uint4*>(T1 + a * T_PITCH + 32 * q);
        const uint4* t2 = reinterpret_cast<const uint4*>(T2 + a * T_PITCH + 32 * q);
        const uint4 km0 = *reinterpret_cast<const uint4*>(kmask + 32 * q);
        const uint4 km1 = *reinterpret_cast<const uint4*>(kmask + 32 * q + 16);
        const uint32_t kmw[8] = {km0.x, km0.y, km0.z, km0.w, km1.x, km1.y, km1.z, km1.w};
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 c1 = t1[g], c2 = t2[g];
          const uint32_t w1[4] = {c1.x, c1.y, c1.z, c1.w}, w2[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&w1[u]));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w2[u]));
            const int e = g * 8 + u * 2;
            float s0 = (__uint_as_float(v[e]) + f1.x + f2.x) * p.scale_log2;
            float s1 = (__uint_as_float(v[e + 1]) + f1.y + f2.y) * p.scale_log2;
            if (!valid_q || ((kmw[e >> 2] >> (8 * (e & 3))) & 0xFF) == 0) s0 = -INFINITY;
            if (!valid_q || ((kmw[(e + 1) >> 2] >> (8 * ((e + 1) & 3))) & 0xFF) == 0) s1 = -INFINITY;
            mblk = fmaxf(mblk, fmaxf(s0, s1));
            v[e] = __float_as_uint(s0);
            v[e + 1] = __float_as_uint(s1);
          }
        }
        tmem_st_32x32(lane_base + TM_S + 32 * q, v);
      }
      tmem_st_wait();
      const float m_new = fmaxf(m_run, mblk);
      const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
      const float alpha = exp2f(m_run - m_use);
      // ---- pass B: probabilities -> bf16 A-operand tile (128B-swizzled, K-major), row sum
      float rsum = 0.0f;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + TM_S + 32 * q, v);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          pr[e] = exp2f(__uint_as_float(v[e]) - m_use);
          rsum += pr[e];
        }
        if (do_drop) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t keep = dropout_keep8(rowid, (j0 + 32 * q + 8 * g) >> 3, p.drop_thresh16, p.seed_lo,
                                                p.seed_hi, p.off_lo, p.off_hi);
#pragma unroll
            for (int e = 0; e < 8; ++e) pr[8 * g + e] = ((keep >> e) & 1) ? pr[8 * g + e] * p.inv_keep : 0.0f;
          }
        }
        // columns 32q .. 32q+31 -> chunk (q>>1), 16-byte units 4*(q&1) .. +3 of row a
        uint8_t* prow = smem + OFF_P + (q >> 1) * 16384 + a * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 pk = make_uint4(pack_bf16(pr[8 * g], pr[8 * g + 1]), pack_bf16(pr[8 * g + 2], pr[8 * g + 3]),
                                      pack_bf16(pr[8 * g + 4], pr[8 * g + 5]), pack_bf16(pr[8 * g + 6], pr[8 * g + 7]));
          const int unit = ((q & 1) * 4 + g) ^ (a & 7);
          *reinterpret_cast<uint4*>(prow + unit * 16) = pk;
        }
      }
      l_run = l_run * alpha + rsum;
      m_run = m_new;
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the MMA (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pready);
      // ---- O accumulation in registers (rescaled), O_blk from TMEM
      mbar_wait(bar_mma3, ph);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + TM_O + 32 * q, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) o[32 * q + e] = o[32 * q + e] * alpha + __uint_as_float(v[e]);
      }
    }
    // ---- epilogue: normalise, write ctx row and the log2-domain LSE
    if (i < p.L) {
      const float inv_l = (l_run > 0.0f) ? 1.0f / l_run : 0.0f;
      __nv_bfloat16* dst = p.ctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 pk = make_uint4(pack_bf16(o[8 * g] * inv_l, o[8 * g + 1] * inv_l),
                                    pack_bf16(o[8 * g + 2] * inv_l, o[8 * g + 3] * inv_l),
                                    pack_bf16(o[8 * g + 4] * inv_l, o[8 * g + 5] * inv_l),
                                    pack_bf16(o[8 * g + 6] * inv_l, o[8 * g + 7] * inv_l));
        reinterpret_cast<uint4*>(dst)[g] = pk;
      }
      p.lse[((long long)b * p.h + hh) * p.L + i] = (l_run > 0.0f) ? (m_run + log2f(l_run)) : INFINITY;
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Expand the per-bucket positional projections into per-relative-position tables:
//   out[t][hh][row][0..63] = PQK[bucket_of_rel[row - rel_off + 511]][t * H + hh * 64 + ...]
// where PQK = Linear_{q|k}(dropout(LayerNorm(rel_embeddings)))  [2S, 2H]  (deberta.py:847-853).
__global__ void build_rel_tables_kernel(const __nv_bfloat16* __restrict__ pqk, long long ld, int H, int h,
                                        const int* __restrict__ bucket_of_rel, int rel_rows, int rel_off, int L,
                                        __nv_bfloat16* __restrict__ out_pq, __nv_bfloat16* __restrict__ out_pk) {
  const int row = blockIdx.x, hh = blockIdx.y, t = blockIdx.z;
  const int rel = row - rel_off;
  uint32_t val = 0;
  if (rel > -L && rel < L) {
    const int bucket = bucket_of_rel[rel + 511];
    val = *reinterpret_cast<const uint32_t*>(pqk + (long long)bucket * ld + t * H + hh * 64 + threadIdx.x * 2);
  }
  __nv_bfloat16* out = (t == 0 ? out_pq : out_pk) + ((long long)hh * rel_rows + row) * 64;
  *reinterpret_cast<uint32_t*>(out + threadIdx.x * 2) = val;
}

static int make_qkv_map(CUtensorMap* m, const void* ptr, long long ld, int B, int L, int h) {
  const uint64_t dims[4] = {64, (uint64_t)h, (uint64_t)L, (uint64_t)B};
  const uint64_t strides[3] = {64, (uint64_t)ld, (uint64_t)ld * L};
  const uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16_nd(m, ptr, 4, dims, strides, box);
}
static int make_rel_map(CUtensorMap* m, const void* ptr, int rel_rows, int h) {
  const uint64_t dims[3] = {64, (uint64_t)rel_rows, (uint64_t)h};
  const uint64_t strides[2] = {64, (uint64_t)rel_rows * 64};
  const uint32_t box[3] = {64, 256, 1};
  return make_tmap_bf16_nd(m, ptr, 3, dims, strides, box);
}

int attn_fwd(const fb_attn_args* a, cudaStream_t st) {
  FB_REQUIRE(a && a->Q && a->K && a->V && a->PKrel && a->PQrel && a->mask && a->ctx && a->lse, "fb_attn_fwd: null arg");
  FB_REQUIRE(a->B > 0 && a->h > 0 && a->L > 0 && a->L <= 512, "fb_attn_fwd: need 0 < L <= 512 (got %d)", a->L);
  const int lpad = (a->L + 127) / 128 * 128;
  FB_REQUIRE(a->rel_off >= lpad && a->rel_rows >= a->rel_off + lpad + 1,
             "fb_attn_fwd: relative tables too small (rel_rows=%d rel_off=%d L=%d)", a->rel_rows, a->rel_off, a->L);
  CUtensorMap mq, mk, mv, mpk, mpq;
  int rc;
  if ((rc = make_qkv_map(&mq, a->Q, a->ldq, a->B, a->L, a->h))) return rc;
  if ((rc = make_qkv_map(&mk, a->K, a->ldk, a->B, a->L, a->h))) return rc;
  if ((rc = make_qkv_map(&mv, a->V, a->ldv, a->B, a->L, a->h))) return rc;
  if ((rc = make_rel_map(&mpk, a->PKrel, a->rel_rows, a->h))) return rc;
  if ((rc = make_rel_map(&mpq, a->PQrel, a->rel_rows, a->h))) return rc;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.h = a->h; p.nkb = lpad / 128; p.rel_off = a->rel_off;
  p.mask = a->mask;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(a->ctx); p.ldc = a->ldc;
  p.lse = a->lse;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  if (a->dropout_p > 0.0f) {
    FB_REQUIRE(a->dropout_p < 1.0f, "fb_attn_fwd: dropout_p must be < 1");
    p.drop_thresh16 = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    p.inv_keep = 1.0f / (1.0f - a->dropout_p);
  } else {
    p.drop_thresh16 = 0; p.inv_keep = 1.0f;
  }
  p.seed_lo = (uint32_t)a->seed; p.seed_hi = (uint32_t)(a->seed >> 32);
  p.off_lo = (uint32_t)a->offset; p.off_hi = (uint32_t)(a->offset >> 32);
  static bool configured = false;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    configured = true;
  }
  dim3 grid(lpad / 128, a->h, a->B);
  attn_fwd_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(mq, mk, mv, mpk, mpq, p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int build_rel_tables(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                     int rel_off, int L, void* out_pq, void* out_pk, cudaStream_t st) {
  FB_REQUIRE(pqk && bucket_of_rel && out_pq && out_pk, "fb_build_rel_tables: null arg");
  FB_REQUIRE(H == h * 64, "fb_build_rel_tables: head_dim must be 64");
  dim3 grid(rel_rows, h, 2);
  build_rel_tables_kernel<<<grid, 32, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(pqk), ld, H, h, bucket_of_rel,
                                               rel_rows, rel_off, L, reinterpret_cast<__nv_bfloat16*>(out_pq),
                                               reinterpret_cast<__nv_bfloat16*>(out_pk));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb

extern "C" {
int fb_attn_fwd(const fb_attn_args* args, void* stream) { return fb::attn_fwd(args, reinterpret_cast<cudaStream_t>(stream)); }
int fb_build_rel_tables(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                        int rel_off, int L, void* out_pq, void* out_pk, void* stream) {
  return fb::build_rel_tables(pqk, ld, H, h, bucket_of_rel, rel_rows, rel_off, L, out_pq, out_pk,
                              reinterpret_cast<cudaStream_t>(stream));
}
}
