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
#include <cstdlib>
#include <cstring>

#include "../../include/frozenbilm_b200.h"
#include "fb_common.cuh"

namespace fb {

static constexpr int AT_THREADS = 384;   // warps 0-7: two softmax warpgroups (thread = query row, warpgroup =
                                         // half of the keys), warp 8 = TMA+MMA control, warp 9 = TMEM alloc,
                                         // warps 10-11 idle (they complete the third warpgroup for setmaxnreg)
static constexpr int AT_REGS_SOFTMAX = 224, AT_REGS_CONTROL = 56;   // 2*128*224 + 128*56 = 64512 <= 65536
static constexpr int T_PITCH = 136;      // halves per staging row: 272 B -> conflict-free LDS.128
static constexpr uint32_t OFF_Q = 0, OFF_K = 16384, OFF_V = 32768, OFF_PK = 49152, OFF_PQ = 81920,
                          OFF_P = 114688, OFF_T1 = 147456, OFF_T2 = OFF_T1 + 128 * T_PITCH * 2,
                          OFF_KMASK = OFF_T2 + 128 * T_PITCH * 2, OFF_BARS = OFF_KMASK + 128,
                          AT_SMEM = OFF_BARS + 128 + 1024;
static constexpr uint32_t TM_S = 0, TM_X = 128, TM_O = 384;   // O: two 64-column halves (one per warpgroup)

struct AttnParams {
  int B, L, h, nkb, rel_off;
  const uint8_t* mask;
  __nv_bfloat16* ctx; long long ldc;
  float* lse;
  __half* s_out; int lp;  // optional [B, h, L, lp] fp16 copy of the masked, scaled scores (log2 domain)
  float scale_log2;
  int mask_queries;                        // DeBERTa: M[i,j] = m_i * m_j;  BERT: key mask only
  float inv_keep; uint32_t drop_thresh16;  // keep iff rnd16 >= thresh
  uint32_t seed_lo, seed_hi, off_lo, off_hi;
};

// Staging stores of the Toeplitz skew (one 32-column TMEM chunk of this thread's lane).
//   MODE 0: elements e >= thr   MODE 1: every element   MODE 2: elements e < thr
// p2c: element e -> dst[e * T_PITCH]     c2p: element e -> dst[-e]
template <int MODE>
__device__ __forceinline__ void stage_p2c(__half* dst, const uint32_t (&v)[32], int thr) {
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const __half hv = __float2half_rn(__uint_as_float(v[e]));
    if (MODE == 1 || (MODE == 0 && e >= thr) || (MODE == 2 && e < thr)) dst[e * T_PITCH] = hv;
  }
}
template <int MODE>
__device__ __forceinline__ void stage_c2p(__half* dst, const uint32_t (&v)[32], int thr) {
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const __half hv = __float2half_rn(__uint_as_float(v[e]));
    if (MODE == 1 || (MODE == 0 && e >= thr) || (MODE == 2 && e < thr)) dst[-e] = hv;
  }
}

// Pipeline (per 128-key block kb; the tensor core and the softmax warps overlap across blocks):
//   control thread                                   softmax warps (8 warps, thread = query row)
//   wait mma1(kb)   -> TMA K/PQ(kb+1)                wait mma1(kb): p2c staging of X2 -> arrive xfree
//   wait xfree(kb)  -> X1(kb) MMA -> commit mma2     O += P.V of block kb-1 (deferred, in the shadow of X1)
//   wait mma2(kb)   -> TMA PK(kb+1)                  wait mma2(kb): c2p staging of X1 -> bar.sync
//   wait sfree(kb)  -> S/X2(kb+1) MMA -> commit mma1 combine + mask + max in REGISTERS -> arrive sfree
//   wait pready(kb) -> P.V(kb) MMA -> commit mma3    exp2 / dropout / P -> smem -> arrive pready
// so the next block's S and X2 are computed while the current block is exponentiated, and the P.V result is
// consumed one block later; only the X1 MMA (which shares TMEM columns with X2) is waited for in place.
template <bool REL>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_pk,
                const __grid_constant__ CUtensorMap map_pq, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms) by pointer arithmetic on the __shared__ array, so that the compiler
  // keeps the shared address space (STS/LDS instead of generic ST/LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;        // TMA: K block + PQ band (operands of S and X2)
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_mma1 = bars + 3;     // S, X2 complete
  uint64_t* bar_xfree = bars + 4;    // X2 staged by all 8 warps
  uint64_t* bar_mma2 = bars + 5;     // X1 complete
  uint64_t* bar_pready = bars + 6;   // [2], one per softmax warpgroup
  uint64_t* bar_mma3 = bars + 8;     // [2]
  uint64_t* bar_pk = bars + 10;      // TMA: PK band (operand of X1)
  uint64_t* bar_sfree = bars + 11;   // S / X1 / staging buffers consumed by all 8 warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  __half* T1 = reinterpret_cast<__half*>(smem + OFF_T1);
  __half* T2 = reinterpret_cast<__half*>(smem + OFF_T2);
  uint8_t* kmask = smem + OFF_KMASK;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, hh = blockIdx.y, b = blockIdx.z;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_pk); tma_prefetch_desc(&map_pq);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    mbar_init(bar_mma1, 1); mbar_init(bar_xfree, 8); mbar_init(bar_mma2, 1);
    mbar_init(&bar_pready[0], 4); mbar_init(&bar_pready[1], 4);
    mbar_init(&bar_mma3[0], 1); mbar_init(&bar_mma3[1], 1);
    mbar_init(bar_pk, 1); mbar_init(bar_sfree, 8);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // register reallocation between the warpgroups: the softmax threads keep 64 output accumulators AND the 64
  // scores of a block in registers (no TMEM round trip between the max and the exponentiation)
  if (warp >= 8) {
    // the whole third warpgroup (warps 8-11) releases registers at the same instruction
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AT_REGS_CONTROL));
    if (warp == 8 && elect_one_sync()) {
      // ======================= TMA + MMA control thread =======================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_band = make_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t sQ = smem_u32(smem + OFF_Q), sK = smem_u32(smem + OFF_K), sV = smem_u32(smem + OFF_V),
                     sPK = smem_u32(smem + OFF_PK), sPQ = smem_u32(smem + OFF_PQ), sP = smem_u32(smem + OFF_P);
      auto load_k = [&](int kb) {
        const int j0 = kb * 128;
        mbar_arrive_expect_tx(bar_k, REL ? 16384 + 32768 : 16384);
        tma_load_4d(smem + OFF_K, &map_k, bar_k, 0, hh, j0, b);
        if constexpr (REL) tma_load_3d(smem + OFF_PQ, &map_pq, bar_k, 0, i0 - j0 - 127 + p.rel_off, hh);
      };
      auto load_pk = [&](int kb) {
        if constexpr (REL) {
          mbar_arrive_expect_tx(bar_pk, 32768);
          tma_load_3d(smem + OFF_PK, &map_pk, bar_pk, 0, i0 - kb * 128 - 127 + p.rel_off, hh);
        }
      };
      auto load_v = [&](int kb) {
        mbar_arrive_expect_tx(bar_v, 16384);
        tma_load_4d(smem + OFF_V, &map_v, bar_v, 0, hh, kb * 128, b);
      };
      auto issue_mma1 = [&]() {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + TM_S, make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(sK + k * 32, 16, 1024), idesc_qk, k > 0);
        if constexpr (REL) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TM_X, make_smem_desc_sw128(sK + k * 32, 16, 1024),
                      make_smem_desc_sw128(sPQ + k * 32, 16, 1024), idesc_band, k > 0);
        }
        umma_commit(bar_mma1);
      };
      mbar_arrive_expect_tx(bar_q, 16384);
      tma_load_4d(smem + OFF_Q, &map_q, bar_q, 0, hh, i0, b);
      load_k(0);
      load_pk(0);
      load_v(0);
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      issue_mma1();
      for (int kb = 0; kb < p.nkb; ++kb) {
        const uint32_t ph = kb & 1;
        const bool has_next = kb + 1 < p.nkb;
        mbar_wait(bar_mma1, ph);      // S / X2 of this block are complete: the K and PQ tiles are free
        if (has_next) load_k(kb + 1);
        if (kb > 0) {
          // the previous block's P.V (both halves) must retire before its V tile is overwritten
          mbar_wait(&bar_mma3[0], (kb - 1) & 1);
          mbar_wait(&bar_mma3[1], (kb - 1) & 1);
          load_v(kb);
        }
        if constexpr (REL) mbar_wait(bar_pk, ph);
        mbar_wait(bar_xfree, ph);     // X2 has been staged: its TMEM columns take X1
        tc_fence_after();
        if constexpr (REL) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TM_X, make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                      make_smem_desc_sw128(sPK + k * 32, 16, 1024), idesc_band, k > 0);
        }
        umma_commit(bar_mma2);
        if (has_next) {
          if constexpr (REL) {
            mbar_wait(bar_mma2, ph);  // the PK tile is free
            load_pk(kb + 1);
          }
          // next block's S / X2 as soon as every warp has pulled this block's S and X1 out of TMEM
          mbar_wait(bar_sfree, ph);
          mbar_wait(bar_k, (kb + 1) & 1);
          tc_fence_after();
          issue_mma1();
        }
        mbar_wait(bar_v, ph);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          // O_g = P[:, 64g:64g+64] . V[64g:64g+64, :]   (each softmax warpgroup owns one half of the keys)
          mbar_wait(&bar_pready[g], ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TM_O + 64 * g, make_smem_desc_sw128(sP + g * 16384 + k * 32, 16, 1024),
                      make_smem_desc_sw128(sV + (4 * g + k) * 2048, 16, 1024), idesc_pv, k > 0);
          umma_commit(&bar_mma3[g]);
        }
      }
    }
    __syncwarp();
  } else {
    // ============ softmax warpgroups: thread = query row, warpgroup g = key columns [64g, 64g+64) ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AT_REGS_SOFTMAX));
    const int g = warp >> 2, wq = warp & 3;
    const int a = wq * 32 + lane;        // row in tile == TMEM lane
    const int i = i0 + a;
    const bool valid_q = (i < p.L) && (!p.mask_queries || p.mask[(long long)b * p.L + i] != 0);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t rowid = static_cast<uint32_t>((b * p.h + hh) * p.L + i);
    const bool do_drop = p.drop_thresh16 != 0;
    const uint32_t rowkey = drop_row_key_raw(rowid, p.seed_lo, p.seed_hi, p.off_lo, p.off_hi);
    float m_run = -INFINITY, l_run = 0.0f, alpha_pend = 1.0f;
    float o[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) o[k] = 0.0f;
    // o = o * alpha + (P.V of block kb): issued one block late so that the MMA never has to be waited for
    auto accumulate_o = [&](int kb) {
      mbar_wait(&bar_mma3[g], kb & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld_32x32(lane_base + TM_O + 64 * g, va);
      tmem_ld_32x32(lane_base + TM_O + 64 * g + 32, vb);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        o[e] = fmaf(o[e], alpha_pend, __uint_as_float(va[e]));
        o[32 + e] = fmaf(o[32 + e], alpha_pend, __uint_as_float(vb[e]));
      }
      tc_fence_before();
    };

    for (int kb = 0; kb < p.nkb; ++kb) {
      const uint32_t ph = kb & 1;
      const int j0 = kb * 128;
      mbar_wait(bar_mma1, ph);
      tc_fence_after();
      if (g == 0) {
        // key-validity bits of this block: word wq covers keys [32 wq, 32 wq + 32)
        const int j = j0 + a;
        const bool ok = (j < p.L) && (p.mask[(long long)b * p.L + j] != 0);
        const uint32_t bits = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) reinterpret_cast<uint32_t*>(kmask)[wq] = bits;
      }
      // ---- p2c: lane = KEY row a of X2; scatter its 160-column window into T2[query][key].
      //      Window offset t = 32q + e is valid for 31 - lane <= t < 159 - lane: chunk 0 needs e >= 31 - lane,
      //      chunks 1-3 are valid for every lane, chunk 4 needs e < 31 - lane.  (g = 0: chunks 0-2, g = 1: 3-4.)
      //      The TMEM load of the next chunk is in flight while the current one is stored.
      if constexpr (REL) {
        uint32_t va[32], vb[32];
        const uint32_t xbase = lane_base + TM_X + (96 - 32 * wq);
        __half* dst = T2 + (lane - 31) * T_PITCH + a;   // chunk q, element e -> dst[(32 q + e) * T_PITCH]
        const int thr = 31 - lane;
        if (g == 0) {
          tmem_ld_32x32(xbase, va);
          tmem_ld_wait();
          tmem_ld_32x32(xbase + 32, vb);
          stage_p2c<0>(dst, va, thr);
          tmem_ld_wait();
          tmem_ld_32x32(xbase + 64, va);
          stage_p2c<1>(dst + 32 * T_PITCH, vb, thr);
          tmem_ld_wait();
          stage_p2c<1>(dst + 64 * T_PITCH, va, thr);
        } else {
          tmem_ld_32x32(xbase + 96, va);
          tmem_ld_wait();
          tmem_ld_32x32(xbase + 128, vb);
          stage_p2c<1>(dst + 96 * T_PITCH, va, thr);
          tmem_ld_wait();
          stage_p2c<2>(dst + 128 * T_PITCH, vb, thr);
        }
      }
      tc_fence_before();
      __syncwarp();
      if constexpr (REL) {          // (plain attention has no X2 / X1 hand-over: nobody waits on this barrier)
        if (lane == 0) mbar_arrive(bar_xfree);
      }
      if (kb > 0) accumulate_o(kb - 1);
      mbar_wait(bar_mma2, ph);
      tc_fence_after();
      // ---- c2p: lane = QUERY row a of X1; keys b = a + 127 - c.  Warpgroup g needs b in [64g, 64g+64):
      //      3 chunks starting at column 32*wq + 64*(1-g); key of element e in chunk q: 64g + lane + 63 - 32q - e.
      //      chunk 0 needs e >= lane, chunk 1 is valid for every lane, chunk 2 needs e < lane.
      if constexpr (REL) {
        uint32_t va[32], vb[32];
        const uint32_t xbase = lane_base + TM_X + 32 * wq + 64 * (1 - g);
        __half* dst = T1 + a * T_PITCH + 64 * g + lane + 63;   // chunk q, element e -> dst[-32 q - e]
        tmem_ld_32x32(xbase, va);
        tmem_ld_wait();
        tmem_ld_32x32(xbase + 32, vb);
        stage_c2p<0>(dst, va, lane);
        tmem_ld_wait();
        tmem_ld_32x32(xbase + 64, va);
        stage_c2p<1>(dst - 32, vb, lane);
        tmem_ld_wait();
        stage_c2p<2>(dst - 64, va, lane);
      }
      // ---- this warpgroup's 64 scores of the row, straight into registers (both TMEM loads in flight at once)
      uint32_t sv[2][32];
      tmem_ld_32x32(lane_base + TM_S + 64 * g, sv[0]);
      tmem_ld_32x32(lane_base + TM_S + 64 * g + 32, sv[1]);
      named_bar_sync(1, 256);  // T2 / kmask were written by other threads
      tmem_ld_wait();
      // ---- pass A: combine the three terms, mask, row max over this warpgroup's 64 keys
      float mblk = -INFINITY;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c0 = 64 * g + 32 * q;
        uint32_t (&v)[32] = sv[q];
        const uint32_t kbits = valid_q ? reinterpret_cast<const uint32_t*>(kmask)[2 * g + q] : 0u;
        if constexpr (REL) {
          const uint4* t1 = reinterpret_cast<const uint4*>(T1 + a * T_PITCH + c0);
          const uint4* t2 = reinterpret_cast<const uint4*>(T2 + a * T_PITCH + c0);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            const uint4 c1 = t1[gg], c2 = t2[gg];
            const uint32_t w1[4] = {c1.x, c1.y, c1.z, c1.w}, w2[4] = {c2.x, c2.y, c2.z, c2.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              // positional terms: fp16 pair sum (one HADD2), then widen
              const __half2 hs = __hadd2(*reinterpret_cast<const __half2*>(&w1[u]),
                                         *reinterpret_cast<const __half2*>(&w2[u]));
              const float2 f = __half22float2(hs);
              const int e = gg * 8 + u * 2;
              v[e] = __float_as_uint(__uint_as_float(v[e]) + f.x);
              v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) + f.y);
            }
          }
        }
        if (kbits == 0xffffffffu) {
          // every key of the chunk is attended (the common case): no per-element select
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float sc = __uint_as_float(v[e]) * p.scale_log2;
            mblk = fmaxf(mblk, sc);
            v[e] = __float_as_uint(sc);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float sc = ((kbits >> e) & 1u) ? __uint_as_float(v[e]) * p.scale_log2 : -INFINITY;
            mblk = fmaxf(mblk, sc);
            v[e] = __float_as_uint(sc);
          }
        }
      }
      if (p.s_out != nullptr) {
        // Saved scores for the backward, as fp16 (11-bit mantissa: |error| <= 2^-7 in the log2 domain for |s| < 16,
        // i.e. <= 0.55 % on a probability -- the size of the bf16 rounding of P that feeds P.V; the fp32 dump this
        // replaces cost 415 MB of writes + 415 MB of reads per layer pass at B = 32).  A direct store would touch 32
        // different 128-byte lines per instruction (thread = row): instead each warp transposes its 32 rows x 128
        // bytes through the staging bytes its own lanes have just consumed (T1 rows a, halves [64g, 64g+64): private
        // to the warp, so __syncwarp suffices) and writes four complete lines per instruction.
        const int rowb = (T_PITCH * 2);
        uint8_t* mine = reinterpret_cast<uint8_t*>(T1) + a * rowb + 128 * g;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 h2 = __floats2half2_rn(__uint_as_float(sv[q][8 * u + 2 * e]), __uint_as_float(sv[q][8 * u + 2 * e + 1]));
              w[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(mine + 64 * q + 16 * u) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        __syncwarp();
        const int rr = lane >> 3, uu = lane & 7;
        __half* gbase = p.s_out + (((long long)b * p.h + hh) * p.L + i0 + 32 * wq) * p.lp + j0 + 64 * g + 8 * uu;
        const uint8_t* tb = reinterpret_cast<const uint8_t*>(T1) + (32 * wq) * rowb + 128 * g + 16 * uu;
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const int r = 4 * st + rr;
          const uint4 x = *reinterpret_cast<const uint4*>(tb + r * rowb);
          if (i0 + 32 * wq + r < p.L) *reinterpret_cast<uint4*>(gbase + (long long)r * p.lp) = x;
        }
      }
      // S, X1 and the staging buffers of this block are consumed: the next block's S / X2 may be issued
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree);
      const float m_new = fmaxf(m_run, mblk);
      const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
      const float alpha = ex2_approx(m_run - m_use);   // ex2(-inf) = +0
      // ---- pass B: probabilities -> bf16 A-operand chunk g (128B-swizzled, K-major), partial row sum
      float rsum = 0.0f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c0 = 64 * g + 32 * q;
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          pr[e] = ex2_approx(__uint_as_float(sv[q][e]) - m_use);   // ex2(-inf) = +0
          rsum += pr[e];
        }
        if (do_drop) {
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            float f[8];
            drop_factors8(rowkey, static_cast<uint32_t>(j0 + c0 + 8 * gg) >> 3, p.drop_thresh16, p.inv_keep, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) pr[8 * gg + e] *= f[e];
          }
        }
        uint8_t* prow = smem + OFF_P + g * 16384 + a * 128;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const uint4 pk = make_uint4(pack_bf16(pr[8 * gg], pr[8 * gg + 1]), pack_bf16(pr[8 * gg + 2], pr[8 * gg + 3]),
                                      pack_bf16(pr[8 * gg + 4], pr[8 * gg + 5]), pack_bf16(pr[8 * gg + 6], pr[8 * gg + 7]));
          const int unit = (q * 4 + gg) ^ (a & 7);
          *reinterpret_cast<uint4*>(prow + unit * 16) = pk;
        }
      }
      l_run = l_run * alpha + rsum;
      m_run = m_new;
      alpha_pend = alpha;
      fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the MMA (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pready[g]);
    }
    accumulate_o(p.nkb - 1);
    // ---- merge the two key halves (flash-decoding style) through shared memory, then write ctx / LSE
    named_bar_sync(1, 256);                      // all staging reads are done: reuse T1/T2 as exchange space
    float* xch = reinterpret_cast<float*>(smem + OFF_T1);   // [128][66] floats = 33.8 KB <= 2 * 34.8 KB
    if (g == 1) {
      float* dstx = xch + a * 66;
      dstx[0] = m_run; dstx[1] = l_run;
#pragma unroll
      for (int k = 0; k < 64; ++k) dstx[2 + k] = o[k];
    }
    named_bar_sync(1, 256);
    if (g == 0 && i < p.L) {
      const float* srcx = xch + a * 66;
      const float m1 = srcx[0], l1 = srcx[1];
      const float m_all = fmaxf(m_run, m1);
      const float m_use = (m_all == -INFINITY) ? 0.0f : m_all;
      const float a0 = exp2f(m_run - m_use), a1 = exp2f(m1 - m_use);
      const float l_all = l_run * a0 + l1 * a1;
      const float inv_l = (l_all > 0.0f) ? 1.0f / l_all : 0.0f;
      __nv_bfloat16* dst = p.ctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) {
        float r[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = (o[8 * gg + e] * a0 + srcx[2 + 8 * gg + e] * a1) * inv_l;
        reinterpret_cast<uint4*>(dst)[gg] = make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]),
                                                        pack_bf16(r[4], r[5]), pack_bf16(r[6], r[7]));
      }
      p.lse[((long long)b * p.h + hh) * p.L + i] = (l_all > 0.0f) ? (m_all + log2f(l_all)) : INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================
// Forward, second generation (default; FB_ATTN_FWD=1 selects the kernel above).
//
// The first kernel regroups the two positional terms through fp16 staging buffers that the softmax threads then
// READ BACK and add to S (per 128 x 128 block and row: 352 two-byte scatter stores with 2-way bank conflicts, 32
// 16-byte loads, ~190 ALU instructions).  Here the staging buffers are MMA OPERANDS and the tensor core does the
// regrouped addition itself:
//
//   Wc[a][k] = X1[a][a + 127 - k]        (thread = query row a stores its own row, k = key index inside the block)
//   Wp[b][x] = X2[b][x - b + 127]        (thread = key row b stores its own row, x = query index inside the tile)
//   S += Wc . I            8 x tcgen05.mma 128x16x16, A = Wc K-major   (adds the c2p term,  Q_a . PK[a - b])
//   S += Wp^T . I          8 x tcgen05.mma 128x16x16, A = Wp MN-major  (adds the p2c term,  K_b . PQ[a - b], transposed
//                                                                       by the operand view: no transposition in smem)
//
// with I the 16 x 16 identity (the trick fb_attn_bwd uses to transpose dS').  Wc / Wp are dense 128 x 128 fp16
// matrices in the SWIZZLE_NONE core-matrix layout with LB = 128 (column groups) and SB = 2080 (row groups): for every
// store instruction the 32 lanes write (row r0 + lane, column x0 + lane), and with these strides the 32 two-byte
// stores fall into 32 different banks (simulated for every (warp, x0): 1.0 wavefronts; the dense SB = 2048 gives
// 3.4, a 128B-swizzled tile 1.8).  Per row and block that leaves 256 conflict-free stores, no loads, no adds; the 16
// extra MMAs cost ~0.8 K tensor-pipe cycles of the ~2.7 K the block needs, on a pipe that was idle 90 % of the time.
// fp16 operands x 1.0 with fp32 accumulation: the positional terms reach S with the same precision as before.
// =============================================================================================
static constexpr uint32_t W_LB = 128, W_SB = 2080, W_BYTES = 15 * W_SB + 16 * W_LB;   // 33248 B per staging matrix
static constexpr uint32_t F2_OFF_WC = OFF_T1, F2_OFF_WP = F2_OFF_WC + 33280, F2_OFF_KMASK = F2_OFF_WP + 33280,
                          F2_OFF_I16 = F2_OFF_KMASK + 128, F2_OFF_BARS = F2_OFF_I16 + 512,
                          AT2_SMEM = F2_OFF_BARS + 256 + 1024;
static_assert(W_BYTES <= 33280 && AT2_SMEM <= 232448, "shared memory budget (forward, generation 2)");

// Staging stores: 32-bit shared-space addresses (8 per thread, one per column residue, computed ONCE before the key
// loop) + an immediate, so that every element is a single STS.U16 [R + imm].
__device__ __forceinline__ void sts16(uint32_t addr, float v) {
  const __half hv = __float2half_rn(v);
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(*reinterpret_cast<const unsigned short*>(&hv)) : "memory");
}
__device__ __forceinline__ uint32_t w_row_base(int r) { return (uint32_t)(r >> 3) * W_SB + (uint32_t)(r & 7) * 16; }
__device__ __forceinline__ int w_col_off(int x) { return (x >> 3) * (int)W_LB + (x & 7) * 2; }   // x may be < 0 (masked)

// NC consecutive TMEM columns of this thread's row, first element index e0 inside its 32-column chunk (for the
// validity threshold); element e -> staging column (first + e) for DIR = +1 (Wp), (first - e) for DIR = -1 (Wc), where
// `adr[r]` is the address of column (first +- r) and every further 8 columns are +-W_LB bytes away.
//   MODE 0: elements with (e0 + e) >= thr   MODE 1: every element   MODE 2: elements with (e0 + e) < thr
template <int MODE, int DIR, int NC>
__device__ __forceinline__ void stage_w(const uint32_t (&adr)[8], int unit0, const uint32_t* v, int e0, int thr) {
#pragma unroll
  for (int e = 0; e < NC; ++e) {
    const int ee = e0 + e;
    if (MODE == 1 || (MODE == 0 && ee >= thr) || (MODE == 2 && ee < thr))
      sts16(adr[ee & 7] + (uint32_t)(DIR * (unit0 + (ee >> 3)) * (int)W_LB), __uint_as_float(v[e]));
  }
}

template <bool REL>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_pk,
                 const __grid_constant__ CUtensorMap map_pq, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F2_OFF_BARS);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;        // TMA: K block + PQ band (operands of S and X2)
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_mma1 = bars + 3;     // S (content term), X2 complete
  uint64_t* bar_xfree = bars + 4;    // X2 is out of TMEM and Wp is staged (all 8 warps)
  uint64_t* bar_mma2 = bars + 5;     // X1 complete
  uint64_t* bar_pready = bars + 6;   // [2], one per softmax warpgroup
  uint64_t* bar_mma3 = bars + 8;     // [2]
  uint64_t* bar_pk = bars + 10;      // TMA: PK band (operand of X1)
  uint64_t* bar_sfree = bars + 11;   // S consumed by all 8 warps (and the staging matrices are free)
  uint64_t* bar_cready = bars + 12;  // [2] X1 is out of TMEM and the Wc columns of key half g are staged (warpgroup g)
  uint64_t* bar_mmas = bars + 14;    // [2] S columns of key half g are final: both positional terms added
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint8_t* wc = smem + F2_OFF_WC;
  uint8_t* wp = smem + F2_OFF_WP;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * 128, hh = blockIdx.y, b = blockIdx.z;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    tma_prefetch_desc(&map_pk); tma_prefetch_desc(&map_pq);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    mbar_init(bar_mma1, 1); mbar_init(bar_xfree, 8); mbar_init(bar_mma2, 1);
    mbar_init(&bar_pready[0], 4); mbar_init(&bar_pready[1], 4);
    mbar_init(&bar_mma3[0], 1); mbar_init(&bar_mma3[1], 1);
    mbar_init(bar_pk, 1); mbar_init(bar_sfree, 8);
    mbar_init(&bar_cready[0], 4); mbar_init(&bar_cready[1], 4); mbar_init(&bar_mmas[0], 1); mbar_init(&bar_mmas[1], 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (REL && threadIdx.x < 256) {
    // I16[n][k] (fp16) as a K-major SWIZZLE_NONE B operand: core matrices (n >> 3, k >> 3), LBO (K) = 128, SBO (N) = 256
    const int n = threadIdx.x >> 4, k = threadIdx.x & 15;
    *reinterpret_cast<uint16_t*>(smem + F2_OFF_I16 + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
        (n == k) ? static_cast<uint16_t>(0x3C00) : static_cast<uint16_t>(0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_prologue();

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AT_REGS_CONTROL));
    if (warp == 8 && elect_one_sync()) {
      // ======================= TMA + MMA control thread =======================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_band = make_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t idesc_addc = make_idesc_f16(128, 16, 0, 0);   // S[:, 16j..] += Wc[:, 16j..] . I
      constexpr uint32_t idesc_addp = make_idesc_f16(128, 16, 1, 0);   // S[:, 16j..] += Wp[16j.., :]^T . I
      const uint32_t sQ = smem_u32(smem + OFF_Q), sK = smem_u32(smem + OFF_K), sV = smem_u32(smem + OFF_V),
                     sPK = smem_u32(smem + OFF_PK), sPQ = smem_u32(smem + OFF_PQ), sP = smem_u32(smem + OFF_P),
                     sWc = smem_u32(wc), sWp = smem_u32(wp);
      const uint64_t idb = make_smem_desc_none(smem_u32(smem + F2_OFF_I16), 128, 256);
      auto load_k = [&](int kb) {
        const int j0 = kb * 128;
        mbar_arrive_expect_tx(bar_k, REL ? 16384 + 32768 : 16384);
        tma_load_4d(smem + OFF_K, &map_k, bar_k, 0, hh, j0, b);
        if constexpr (REL) tma_load_3d(smem + OFF_PQ, &map_pq, bar_k, 0, i0 - j0 - 127 + p.rel_off, hh);
      };
      auto load_pk = [&](int kb) {
        if constexpr (REL) {
          mbar_arrive_expect_tx(bar_pk, 32768);
          tma_load_3d(smem + OFF_PK, &map_pk, bar_pk, 0, i0 - kb * 128 - 127 + p.rel_off, hh);
        }
      };
      auto load_v = [&](int kb) {
        mbar_arrive_expect_tx(bar_v, 16384);
        tma_load_4d(smem + OFF_V, &map_v, bar_v, 0, hh, kb * 128, b);
      };
      auto issue_mma1 = [&]() {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + TM_S, make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(sK + k * 32, 16, 1024), idesc_qk, k > 0);
        if constexpr (REL) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TM_X, make_smem_desc_sw128(sK + k * 32, 16, 1024),
                      make_smem_desc_sw128(sPQ + k * 32, 16, 1024), idesc_band, k > 0);
        }
        umma_commit(bar_mma1);
      };
      mbar_arrive_expect_tx(bar_q, 16384);
      tma_load_4d(smem + OFF_Q, &map_q, bar_q, 0, hh, i0, b);
      load_k(0);
      load_pk(0);
      load_v(0);
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      issue_mma1();
      for (int kb = 0; kb < p.nkb; ++kb) {
        const uint32_t ph = kb & 1;
        const bool has_next = kb + 1 < p.nkb;
        mbar_wait(bar_mma1, ph);      // S / X2 of this block are complete: the K and PQ tiles are free
        if (has_next) load_k(kb + 1);
        if (kb > 0) {
          mbar_wait(&bar_mma3[0], (kb - 1) & 1);
          mbar_wait(&bar_mma3[1], (kb - 1) & 1);
          load_v(kb);
        }
        if constexpr (REL) {
          mbar_wait(bar_pk, ph);
          mbar_wait(bar_xfree, ph);   // X2 is out of TMEM (its columns take X1) and Wp is staged
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TM_X, make_smem_desc_sw128(sQ + k * 32, 16, 1024),
                      make_smem_desc_sw128(sPK + k * 32, 16, 1024), idesc_band, k > 0);
          umma_commit(bar_mma2);
          // p2c term: S[a][16 j + n] += sum_{b in 16 j ..} Wp[b][a] I[b - 16 j][n]   (A = Wp^T: MN-major view, K = key rows)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16(tmem + TM_S + 16 * j, make_smem_desc_none(sWp + j * 2 * W_SB, W_SB, W_LB), idb, idesc_addp, 1u);
          if (has_next) {
            mbar_wait(bar_mma2, ph);  // the PK tile is free
            load_pk(kb + 1);
          }
          // c2p term, one key half per softmax warpgroup (each half is released to its warpgroup on its own, so the
          // two warpgroups drift apart and fill each other's waits):
          //   S[a][16 j + n] += sum_{k in 16 j ..} Wc[a][k] I[k - 16 j][n],   j = 4 g .. 4 g + 3
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&bar_cready[g], ph);   // X1 is out of TMEM (for g = 1: for both halves) and the half is staged
            tc_fence_after();
#pragma unroll
            for (int j = 4 * g; j < 4 * g + 4; ++j)
              umma_bf16(tmem + TM_S + 16 * j, make_smem_desc_none(sWc + j * 2 * W_LB, W_LB, W_SB), idb, idesc_addc, 1u);
            umma_commit(&bar_mmas[g]);
          }
        }
        if (has_next) {
          // next block's S / X2 as soon as every warp has pulled this block's S out of TMEM
          mbar_wait(bar_sfree, ph);
          mbar_wait(bar_k, (kb + 1) & 1);
          tc_fence_after();
          issue_mma1();
        }
        mbar_wait(bar_v, ph);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&bar_pready[g], ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + TM_O + 64 * g, make_smem_desc_sw128(sP + g * 16384 + k * 32, 16, 1024),
                      make_smem_desc_sw128(sV + (4 * g + k) * 2048, 16, 1024), idesc_pv, k > 0);
          umma_commit(&bar_mma3[g]);
        }
      }
    }
    __syncwarp();
  } else {
    // ============ softmax warpgroups: thread = tile row, warpgroup g = key columns [64g, 64g+64) ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AT_REGS_SOFTMAX));
    const int g = warp >> 2, wq = warp & 3;
    const int a = wq * 32 + lane;        // row in tile == TMEM lane
    const int i = i0 + a;
    const bool valid_q = (i < p.L) && (!p.mask_queries || p.mask[(long long)b * p.L + i] != 0);
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t rowid = static_cast<uint32_t>((b * p.h + hh) * p.L + i);
    const bool do_drop = p.drop_thresh16 != 0;
    const uint32_t rowkey = drop_row_key_raw(rowid, p.seed_lo, p.seed_hi, p.off_lo, p.off_hi);
    float m_run = -INFINITY, l_run = 0.0f, alpha_pend = 1.0f;
    float o[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) o[k] = 0.0f;
    auto accumulate_o = [&](int kb) {
      mbar_wait(&bar_mma3[g], kb & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld_32x32(lane_base + TM_O + 64 * g, va);
      tmem_ld_32x32(lane_base + TM_O + 64 * g + 32, vb);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        o[e] = fmaf(o[e], alpha_pend, __uint_as_float(va[e]));
        o[32 + e] = fmaf(o[32 + e], alpha_pend, __uint_as_float(vb[e]));
      }
      tc_fence_before();
    };
    // staging addresses of this thread's row (the same row index a in Wp -- as a key -- and in Wc -- as a query):
    //   Wp column of (chunk q, element e): x = lane - 31 + 32 q + e        -> adrp[e & 7] + (4 q + (e >> 3)) W_LB
    //   Wc column of (chunk q, element e): k = 64 g + lane + 63 - 32 q - e -> adrc[e & 7] - (4 q + (e >> 3)) W_LB
    uint32_t adrp[8], adrc[8];
    {
      const uint32_t rowp = smem_u32(wp) + w_row_base(a), rowc = smem_u32(wc) + w_row_base(a);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        adrp[r] = rowp + (uint32_t)w_col_off(lane - 31 + r);
        adrc[r] = rowc + (uint32_t)w_col_off(64 * g + lane + 63 - r);
      }
    }

    for (int kb = 0; kb < p.nkb; ++kb) {
      const uint32_t ph = kb & 1;
      const int j0 = kb * 128;
      mbar_wait(bar_mma1, ph);
      tc_fence_after();
      // key-validity bits of the two 32-key chunks this warp will score (each warp computes its own: no shared
      // memory, no block-level barrier)
      uint32_t kbits2[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = j0 + 64 * g + 32 * q + lane;
        const bool ok = (j < p.L) && (p.mask[(long long)b * p.L + j] != 0);
        kbits2[q] = __ballot_sync(0xffffffffu, ok);
      }
      // ---- p2c: lane = KEY row a of X2.  Its 128 values for the queries x = 0..127 sit at band columns 127 - a + x:
      //      the warp loads the 160-column window starting at 96 - 32 wq, chunk q / element e is query
      //      x = lane - 31 + 32 q + e (chunk 0 valid for e >= 31 - lane, chunks 1-3 for every lane, chunk 4 for
      //      e < 31 - lane); warpgroup 0 takes chunks 0-2, warpgroup 1 chunks 3-4.
      if constexpr (REL) {
        // (the 160-column window is split 80 / 80 between the two warpgroups: chunks 0, 1 and the first half of chunk 2
        //  for warpgroup 0, the rest for warpgroup 1)
        uint32_t va[32], vb[32], vh[16];
        const uint32_t xbase = lane_base + TM_X + (96 - 32 * wq);
        const int thr = 31 - lane;
        if (g == 0) {
          tmem_ld_32x32(xbase, va);
          tmem_ld_wait();
          tmem_ld_32x32(xbase + 32, vb);
          stage_w<0, 1, 32>(adrp, 0, va, 0, thr);
          tmem_ld_wait();
          tmem_ld_32x16(xbase + 64, vh);
          stage_w<1, 1, 32>(adrp, 4, vb, 0, thr);
          tmem_ld_wait();
          stage_w<1, 1, 16>(adrp, 8, vh, 0, thr);
        } else {
          tmem_ld_32x16(xbase + 80, vh);
          tmem_ld_wait();
          tmem_ld_32x32(xbase + 96, va);
          stage_w<1, 1, 16>(adrp, 8, vh, 16, thr);
          tmem_ld_wait();
          tmem_ld_32x32(xbase + 128, vb);
          stage_w<1, 1, 32>(adrp, 12, va, 0, thr);
          tmem_ld_wait();
          stage_w<2, 1, 32>(adrp, 16, vb, 0, thr);
        }
        fence_proxy_async_smem();    // generic-proxy stores -> visible to the tensor core (async proxy)
      }
      tc_fence_before();
      __syncwarp();
      if constexpr (REL) {          // (plain attention has no X2 / X1 hand-over: nobody waits on this barrier)
        if (lane == 0) mbar_arrive(bar_xfree);
      }
      if (kb > 0) accumulate_o(kb - 1);
      if constexpr (REL) {
        mbar_wait(bar_mma2, ph);
        tc_fence_after();
        // ---- c2p: lane = QUERY row a of X1.  Key k of this block sits at band column a + 127 - k; warpgroup g stages
        //      its own key half [64g, 64g+64): 3 chunks starting at column 32 wq + 64 (1 - g), chunk q / element e is
        //      key 64 g + lane + 63 - 32 q - e (chunk 0 valid for e >= lane, chunk 1 for every lane, chunk 2 for e < lane).
        uint32_t va[32], vb[32];
        const uint32_t xbase = lane_base + TM_X + 32 * wq + 64 * (1 - g);
        tmem_ld_32x32(xbase, va);
        tmem_ld_wait();
        tmem_ld_32x32(xbase + 32, vb);
        stage_w<0, -1, 32>(adrc, 0, va, 0, lane);
        tmem_ld_wait();
        tmem_ld_32x32(xbase + 64, va);
        stage_w<1, -1, 32>(adrc, 4, vb, 0, lane);
        tmem_ld_wait();
        stage_w<2, -1, 32>(adrc, 8, va, 0, lane);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_cready[g]);
        mbar_wait(&bar_mmas[g], ph);   // both positional terms have been added to this key half of S
        tc_fence_after();
      }
      // ---- this warpgroup's 64 final scores of the row
      uint32_t sv[2][32];
      tmem_ld_32x32(lane_base + TM_S + 64 * g, sv[0]);
      tmem_ld_32x32(lane_base + TM_S + 64 * g + 32, sv[1]);
      tmem_ld_wait();
      // ---- pass A: scale, mask, row max over this warpgroup's 64 keys
      float mblk = -INFINITY;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t (&v)[32] = sv[q];
        const uint32_t kbits = valid_q ? kbits2[q] : 0u;
        if (kbits == 0xffffffffu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float sc = __uint_as_float(v[e]) * p.scale_log2;
            mblk = fmaxf(mblk, sc);
            v[e] = __float_as_uint(sc);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float sc = ((kbits >> e) & 1u) ? __uint_as_float(v[e]) * p.scale_log2 : -INFINITY;
            mblk = fmaxf(mblk, sc);
            v[e] = __float_as_uint(sc);
          }
        }
      }
      if (p.s_out != nullptr) {
        // Saved scores for the backward (fp16, see the first kernel): the warp transposes its 32 rows x 128 bytes
        // through a private 4 KB slice of the Wp staging matrix (free: the tensor core has finished reading it, the
        // next block's staging starts only after bar_sfree) and writes four complete lines per instruction.
        uint8_t* xb = wp + warp * 4096;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 h2 = __floats2half2_rn(__uint_as_float(sv[q][8 * u + 2 * e]), __uint_as_float(sv[q][8 * u + 2 * e + 1]));
              w[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(xb + lane * 128 + (((4 * q + u) ^ (lane & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        __syncwarp();
        const int rr = lane >> 3, uu = lane & 7;
        __half* gbase = p.s_out + (((long long)b * p.h + hh) * p.L + i0 + 32 * wq) * p.lp + j0 + 64 * g + 8 * uu;
#pragma unroll
        for (int st = 0; st < 8; ++st) {
          const int r = 4 * st + rr;
          const uint4 x = *reinterpret_cast<const uint4*>(xb + r * 128 + ((uu ^ (r & 7)) << 4));
          if (i0 + 32 * wq + r < p.L) *reinterpret_cast<uint4*>(gbase + (long long)r * p.lp) = x;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree);
      const float m_new = fmaxf(m_run, mblk);
      const float m_use = (m_new == -INFINITY) ? 0.0f : m_new;
      const float alpha = ex2_approx(m_run - m_use);   // ex2(-inf) = +0
      // ---- pass B: probabilities -> bf16 A-operand chunk g (128B-swizzled, K-major), partial row sum
      float rsum = 0.0f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c0 = 64 * g + 32 * q;
        float pr[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          pr[e] = ex2_approx(__uint_as_float(sv[q][e]) - m_use);   // ex2(-inf) = +0
          rsum += pr[e];
        }
        if (do_drop) {
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            float f[8];
            drop_factors8(rowkey, static_cast<uint32_t>(j0 + c0 + 8 * gg) >> 3, p.drop_thresh16, p.inv_keep, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) pr[8 * gg + e] *= f[e];
          }
        }
        uint8_t* prow = smem + OFF_P + g * 16384 + a * 128;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const uint4 pk = make_uint4(pack_bf16(pr[8 * gg], pr[8 * gg + 1]), pack_bf16(pr[8 * gg + 2], pr[8 * gg + 3]),
                                      pack_bf16(pr[8 * gg + 4], pr[8 * gg + 5]), pack_bf16(pr[8 * gg + 6], pr[8 * gg + 7]));
          const int unit = (q * 4 + gg) ^ (a & 7);
          *reinterpret_cast<uint4*>(prow + unit * 16) = pk;
        }
      }
      l_run = l_run * alpha + rsum;
      m_run = m_new;
      alpha_pend = alpha;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pready[g]);
    }
    accumulate_o(p.nkb - 1);
    // ---- merge the two key halves through shared memory (the staging matrices are free), then write ctx / LSE
    named_bar_sync(1, 256);
    float* xch = reinterpret_cast<float*>(wc);   // [128][66] floats = 33.8 KB <= the two staging matrices (66.5 KB)
    if (g == 1) {
      float* dstx = xch + a * 66;
      dstx[0] = m_run; dstx[1] = l_run;
#pragma unroll
      for (int k = 0; k < 64; ++k) dstx[2 + k] = o[k];
    }
    named_bar_sync(1, 256);
    if (g == 0 && i < p.L) {
      const float* srcx = xch + a * 66;
      const float m1 = srcx[0], l1 = srcx[1];
      const float m_all = fmaxf(m_run, m1);
      const float m_use = (m_all == -INFINITY) ? 0.0f : m_all;
      const float a0 = exp2f(m_run - m_use), a1 = exp2f(m1 - m_use);
      const float l_all = l_run * a0 + l1 * a1;
      const float inv_l = (l_all > 0.0f) ? 1.0f / l_all : 0.0f;
      __nv_bfloat16* dst = p.ctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) {
        float r[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = (o[8 * gg + e] * a0 + srcx[2 + 8 * gg + e] * a1) * inv_l;
        reinterpret_cast<uint4*>(dst)[gg] = make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]),
                                                        pack_bf16(r[4], r[5]), pack_bf16(r[6], r[7]));
      }
      p.lse[((long long)b * p.h + hh) * p.L + i] = (l_all > 0.0f) ? (m_all + log2f(l_all)) : INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Expand the per-bucket positional projections into per-relative-position tables:
//   out[t][hh][row][0..63] = PQK[bucket_of_rel[row - rel_off + 511]][t * H + hh * 64 + ...]
// where PQK = Linear_{q|k}(dropout(LayerNorm(rel_embeddings)))  [2S, 2H]  (deberta.py:847-853).
// 8 table rows per 256-thread CTA (one warp per row, 4 bytes per lane): the one-warp-per-CTA geometry spent its
// 20 us on scheduling 37 000 tiny CTAs.
__global__ void __launch_bounds__(256)
build_rel_tables_kernel(const __nv_bfloat16* __restrict__ pqk, long long ld, int H, int h,
                        const int* __restrict__ bucket_of_rel, int rel_rows, int rel_off, int L,
                        __nv_bfloat16* __restrict__ out_pq, __nv_bfloat16* __restrict__ out_pk,
                        __nv_bfloat16* __restrict__ out_pk_rev, long long stride_pqk) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int hh = blockIdx.y, t = blockIdx.z & 1;
  if (row >= rel_rows) return;
  {
    // batched over layer passes: blockIdx.z = 2 * pass + t, the tables of a pass are [h, rel_rows, 64] back to back
    const long long pass = blockIdx.z >> 1, tab = (long long)h * rel_rows * 64;
    pqk += pass * stride_pqk;
    out_pq += pass * tab; out_pk += pass * tab;
    if (out_pk_rev != nullptr) out_pk_rev += pass * tab;
  }
  pdl_prologue();
  const int rel = row - rel_off;
  uint32_t val = 0;
  if (rel > -L && rel < L) {
    const int bucket = bucket_of_rel[rel + 511];
    val = *reinterpret_cast<const uint32_t*>(pqk + (long long)bucket * ld + t * H + hh * 64 + lane * 2);
  }
  __nv_bfloat16* out = (t == 0 ? out_pq : out_pk) + ((long long)hh * rel_rows + row) * 64;
  *reinterpret_cast<uint32_t*>(out + lane * 2) = val;
  if (t == 1 && out_pk_rev != nullptr)   // row-reversed copy of PKrel for the fused backward (fb_attn_bwd)
    *reinterpret_cast<uint32_t*>(out_pk_rev + ((long long)hh * rel_rows + (rel_rows - 1 - row)) * 64 + lane * 2) = val;
}


// ---------------------------------------------------------------------------------------------
// Attention backward, stage 1 (one CTA per score row): from the saved scores s2 (log2 domain), the
// row log-sum-exp and dP = dO.V^T, rebuild P and form
//   dS = P * (dropout_bwd(dP) - D) * scale,  D = rowsum(dO * O)         (softmax + XDropout backward,
//   deberta.py:134-138, 188-194), written as bf16 in three layouts: dS[b,h,i,j], the dropped
//   probabilities Pd[b,h,i,j] (for dV), and Z1[h, b*L+i, (i-j)+rel_off] = dS (the c2p term regrouped by
//   relative position, so that dQ += Z1 . PKrel and dPKrel = Z1^T . Q are plain GEMMs).
// ---------------------------------------------------------------------------------------------
struct AttnBwdParams {
  int B, L, h, lp, rel_rows, rel_off;
  const __half* s2; const float* lse; const __nv_bfloat16* dp; const __nv_bfloat16* ctx; const __nv_bfloat16* dctx;
  long long ldc;
  __nv_bfloat16* ds; __nv_bfloat16* pd; __nv_bfloat16* z1;
  float scale; DropSpec drop;
};

__global__ void __launch_bounds__(256) attn_bwd_ds_kernel(const AttnBwdParams p) {
  // one warp per score row; 16-byte accesses throughout (the kernel is HBM-bound: 14 B/element read+written)
  __shared__ __align__(16) __nv_bfloat16 srow[8][512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;   // (b * h + hh) * L + i
  const long long nrows = (long long)p.B * p.h * p.L;
  if (row >= nrows) return;
  const int i = (int)(row % p.L);
  const int hh = (int)((row / p.L) % p.h);
  const int b = (int)(row / ((long long)p.L * p.h));
  float d;
  {
    const __nv_bfloat16* o = p.ctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
    const __nv_bfloat16* g = p.dctx + ((long long)b * p.L + i) * p.ldc + hh * 64;
    const float2 o2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(o + 2 * lane));
    const float2 g2 = unpack_bf16(*reinterpret_cast<const uint32_t*>(g + 2 * lane));
    d = o2.x * g2.x + o2.y * g2.y;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
  }
  const float D = d, lse = p.lse[row];
  const uint32_t rowid = (uint32_t)row;
  for (int j0 = lane * 4; j0 < p.lp; j0 += 128) {
    float dsv[4] = {0.f, 0.f, 0.f, 0.f}, pdv[4] = {0.f, 0.f, 0.f, 0.f};
    if (j0 < p.L) {   // L is a multiple of 4 or the tail columns are masked scores (-inf) anyway
      const uint2 s4 = *reinterpret_cast<const uint2*>(p.s2 + row * p.lp + j0);
      const uint2 dp2 = *reinterpret_cast<const uint2*>(p.dp + row * p.lp + j0);
      const float2 s01 = __half22float2(*reinterpret_cast<const __half2*>(&s4.x));
      const float2 s23 = __half22float2(*reinterpret_cast<const __half2*>(&s4.y));
      const float sv[4] = {s01.x, s01.y, s23.x, s23.y};
      const float2 d01 = unpack_bf16(dp2.x), d23 = unpack_bf16(dp2.y);
      float dpv[4] = {d01.x, d01.y, d23.x, d23.y};
      uint32_t keep = 0xF;
      if (p.drop.thresh16) keep = (drop_keep8(p.drop, rowid, (uint32_t)j0 >> 3) >> (j0 & 4)) & 0xF;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool in = (j0 + e) < p.L;
        const float pr = in ? exp2f(sv[e] - lse) : 0.0f;   // s = -inf (masked) or lse = +inf (dead row) -> 0
        const float f = ((keep >> e) & 1) ? p.drop.inv_keep : 0.0f;
        pdv[e] = pr * f;
        dsv[e] = (pr > 0.0f) ? pr * ((in ? dpv[e] : 0.0f) * f - D) * p.scale : 0.0f;
      }
    }
    const uint2 dsq = make_uint2(pack_bf16(dsv[0], dsv[1]), pack_bf16(dsv[2], dsv[3]));
    *reinterpret_cast<uint2*>(p.ds + row * p.lp + j0) = dsq;
    *reinterpret_cast<uint2*>(p.pd + row * p.lp + j0) = make_uint2(pack_bf16(pdv[0], pdv[1]), pack_bf16(pdv[2], pdv[3]));
    *reinterpret_cast<uint2*>(&srow[warp][j0]) = dsq;
  }
  if (p.z1 == nullptr) return;   // plain attention (BERT): no positional terms
  __syncwarp();
  // Z1 row: z[r] = dS[i, j = i - (r - rel_off)], eight entries (16 bytes) per store
  __nv_bfloat16* z = p.z1 + ((long long)hh * p.B * p.L + (long long)b * p.L + i) * p.rel_rows;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.0f);
  for (int r0 = lane * 8; r0 < p.rel_rows; r0 += 256) {
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = i - (r0 + e - p.rel_off);
      v[e] = (j >= 0 && j < p.L) ? srow[warp][j] : zero;
    }
    *reinterpret_cast<uint4*>(z + r0) = *reinterpret_cast<const uint4*>(v);
  }
}

// Stage 2: Z2[h, b*L+j, (i-j)+rel_off] = dS[b,h,i,j]  (the p2c term regrouped by relative position:
// dK += Z2 . PQrel, dPQrel = Z2^T . K).  One CTA per (b, h, 32-key stripe); the stripe is transposed
// through shared memory so that both the dS reads and the Z2 writes are contiguous.
__global__ void __launch_bounds__(256) attn_bwd_z2_kernel(const __nv_bfloat16* __restrict__ ds, int B, int L, int h,
                                                          int lp, int rel_rows, int rel_off,
                                                          __nv_bfloat16* __restrict__ z2) {
  extern __shared__ __nv_bfloat16 tile[];   // [L][34]
  const int j0 = blockIdx.x * 32, hh = blockIdx.y, b = blockIdx.z;
  const long long base = ((long long)b * h + hh) * L;
  for (int idx = threadIdx.x; idx < L * 32; idx += 256) {
    const int i = idx >> 5, jj = idx & 31;
    tile[i * 34 + jj] = (j0 + jj < L) ? ds[(base + i) * lp + j0 + jj] : __float2bfloat16_rn(0.0f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int jj = warp; jj < 32; jj += 8) {
    const int j = j0 + jj;
    if (j >= L) break;
    __nv_bfloat16* z = z2 + ((long long)hh * B * L + (long long)b * L + j) * rel_rows;
    for (int r = lane; r < rel_rows; r += 32) {
      const int i = j + (r - rel_off);
      z[r] = (i >= 0 && i < L) ? tile[i * 34 + jj] : __float2bfloat16_rn(0.0f);
    }
  }
}

// Stage 3: fold the per-relative-position gradients back onto the 2S bucket rows:
//   out[bucket_of_rel[rel], t*H + hh*64 + k] += g_t[hh, rel + rel_off, k]   (t = 0: dPQ, t = 1: dPK)
__global__ void __launch_bounds__(256)
rel_to_bucket_kernel(const float* __restrict__ gpq, const float* __restrict__ gpk, int H, int h,
                     const int* __restrict__ bucket_of_rel, int rel_rows, int rel_off, int L,
                     float* __restrict__ out, long long ldo, long long stride_out) {
  // 4 table rows per CTA (64 threads = 64 columns each)
  const int row = blockIdx.x * 4 + (threadIdx.x >> 6), col = threadIdx.x & 63;
  const int hh = blockIdx.y, t = blockIdx.z & 1;
  if (row >= rel_rows) return;
  {
    // batched over layer passes: blockIdx.z = 2 * pass + t
    const long long pass = blockIdx.z >> 1, tab = (long long)h * rel_rows * 64;
    gpq += pass * tab; gpk += pass * tab;
    out += pass * stride_out;
  }
  pdl_prologue();
  const int rel = row - rel_off;
  if (rel <= -L || rel >= L) return;
  const int bucket = bucket_of_rel[rel + 511];
  const float* g = (t == 0 ? gpq : gpk) + ((long long)hh * rel_rows + row) * 64;
  atomicAdd(out + (long long)bucket * ldo + t * H + hh * 64 + col, g[col]);
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
  FB_REQUIRE(a && a->Q && a->K && a->V && a->mask && a->ctx && a->lse, "fb_attn_fwd: null arg");
  FB_REQUIRE(a->B > 0 && a->h > 0 && a->L > 0 && a->L <= 512, "fb_attn_fwd: need 0 < L <= 512 (got %d)", a->L);
  const int lpad = (a->L + 127) / 128 * 128;
  const bool rel = a->PKrel != nullptr;
  FB_REQUIRE(!rel || a->PQrel, "fb_attn_fwd: PKrel and PQrel must be given together");
  FB_REQUIRE(!rel || (a->rel_off >= lpad && a->rel_rows >= a->rel_off + lpad + 1),
             "fb_attn_fwd: relative tables too small (rel_rows=%d rel_off=%d L=%d)", a->rel_rows, a->rel_off, a->L);
  CUtensorMap mq, mk, mv, mpk, mpq;
  int rc;
  if ((rc = make_qkv_map(&mq, a->Q, a->ldq, a->B, a->L, a->h))) return rc;
  if ((rc = make_qkv_map(&mk, a->K, a->ldk, a->B, a->L, a->h))) return rc;
  if ((rc = make_qkv_map(&mv, a->V, a->ldv, a->B, a->L, a->h))) return rc;
  if (rel) {
    if ((rc = make_rel_map(&mpk, a->PKrel, a->rel_rows, a->h))) return rc;
    if ((rc = make_rel_map(&mpq, a->PQrel, a->rel_rows, a->h))) return rc;
  } else {
    mpk = mk; mpq = mk;   // never dereferenced by the plain-attention instantiation
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.h = a->h; p.nkb = lpad / 128; p.rel_off = a->rel_off;
  p.mask = a->mask;
  p.ctx = reinterpret_cast<__nv_bfloat16*>(a->ctx); p.ldc = a->ldc;
  p.lse = a->lse;
  p.s_out = reinterpret_cast<__half*>(a->s_out); p.lp = lpad;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.mask_queries = a->key_mask_only ? 0 : 1;
  if (a->dropout_p > 0.0f) {
    FB_REQUIRE(a->dropout_p < 1.0f, "fb_attn_fwd: dropout_p must be < 1");
    p.drop_thresh16 = (uint32_t)(a->dropout_p * 65536.0f + 0.5f);
    if (p.drop_thresh16 > 65535u) p.drop_thresh16 = 65535u;
    p.inv_keep = 1.0f / (1.0f - a->dropout_p);
  } else {
    p.drop_thresh16 = 0; p.inv_keep = 1.0f;
  }
  p.seed_lo = (uint32_t)a->seed; p.seed_hi = (uint32_t)(a->seed >> 32);
  p.off_lo = (uint32_t)a->offset; p.off_hi = (uint32_t)(a->offset >> 32);
  static bool configured = false;
  static int generation = 2;   // FB_ATTN_FWD=1: the first-generation kernel (positional terms combined by the threads)
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT2_SMEM));
    const char* e = getenv("FB_ATTN_FWD");
    if (e != nullptr && atoi(e) == 1) generation = 1;
    configured = true;
  }
  dim3 grid(lpad / 128, a->h, a->B);
  if (generation == 2) {
    if (rel)
      FB_CHECK_CUDA(launch_pdl(attn_fwd2_kernel<true>, grid, dim3(AT_THREADS), AT2_SMEM, st, mq, mk, mv, mpk, mpq, p));
    else
      FB_CHECK_CUDA(launch_pdl(attn_fwd2_kernel<false>, grid, dim3(AT_THREADS), AT2_SMEM, st, mq, mk, mv, mpk, mpq, p));
  } else if (rel)
    attn_fwd_kernel<true><<<grid, AT_THREADS, AT_SMEM, st>>>(mq, mk, mv, mpk, mpq, p);
  else
    attn_fwd_kernel<false><<<grid, AT_THREADS, AT_SMEM, st>>>(mq, mk, mv, mpk, mpq, p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int build_rel_tables(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                     int rel_off, int L, void* out_pq, void* out_pk, void* out_pk_rev, int nb, long long stride_pqk,
                     cudaStream_t st) {
  FB_REQUIRE(pqk && bucket_of_rel && out_pq && out_pk, "fb_build_rel_tables: null arg");
  FB_REQUIRE(H == h * 64, "fb_build_rel_tables: head_dim must be 64");
  FB_REQUIRE(nb >= 1 && nb <= 32767, "fb_build_rel_tables: bad batch %d", nb);
  dim3 grid((rel_rows + 7) / 8, h, 2 * nb);
  FB_CHECK_CUDA(launch_pdl(build_rel_tables_kernel, grid, dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(pqk), ld,
                           H, h, bucket_of_rel, rel_rows, rel_off, L, reinterpret_cast<__nv_bfloat16*>(out_pq),
                           reinterpret_cast<__nv_bfloat16*>(out_pk), reinterpret_cast<__nv_bfloat16*>(out_pk_rev),
                           stride_pqk));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attn_bwd_prep(const fb_attn_bwd_prep_args* a, cudaStream_t st) {
  FB_REQUIRE(a && a->s2 && a->lse && a->dp && a->ctx && a->dctx && a->ds && a->pd, "fb_attn_bwd_prep: null arg");
  FB_REQUIRE((a->z1 == nullptr) == (a->z2 == nullptr), "fb_attn_bwd_prep: z1 and z2 go together");
  FB_REQUIRE(a->L > 0 && a->L <= 512 && a->lp >= a->L && a->lp <= 512, "fb_attn_bwd_prep: bad L/lp");
  AttnBwdParams p;
  p.B = a->B; p.L = a->L; p.h = a->h; p.lp = a->lp; p.rel_rows = a->rel_rows; p.rel_off = a->rel_off;
  p.s2 = (const __half*)a->s2; p.lse = a->lse; p.dp = (const __nv_bfloat16*)a->dp; p.ctx = (const __nv_bfloat16*)a->ctx;
  p.dctx = (const __nv_bfloat16*)a->dctx; p.ldc = a->ldc;
  p.ds = (__nv_bfloat16*)a->ds; p.pd = (__nv_bfloat16*)a->pd; p.z1 = (__nv_bfloat16*)a->z1;
  p.scale = a->scale; p.drop = make_drop(a->dropout_p, a->seed, a->offset);
  FB_REQUIRE((a->z1 == nullptr || a->rel_rows % 8 == 0) && a->lp % 4 == 0, "fb_attn_bwd_prep: rel_rows %% 8 / lp %% 4");
  const long long nrows = (long long)a->B * a->h * a->L;
  attn_bwd_ds_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, st>>>(p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  if (a->z2 == nullptr) return 0;
  const int smem = a->L * 34 * 2;
  attn_bwd_z2_kernel<<<dim3((a->L + 31) / 32, a->h, a->B), 256, smem, st>>>(
      (const __nv_bfloat16*)a->ds, a->B, a->L, a->h, a->lp, a->rel_rows, a->rel_off, (__nv_bfloat16*)a->z2);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int rel_to_bucket(const float* gpq, const float* gpk, int H, int h, const int* bucket_of_rel, int rel_rows,
                  int rel_off, int L, float* out, long long ldo, int nb, long long stride_out, cudaStream_t st) {
  FB_REQUIRE(gpq && gpk && bucket_of_rel && out && H == h * 64, "fb_rel_to_bucket: bad args");
  FB_REQUIRE(nb >= 1 && nb <= 32767, "fb_rel_to_bucket: bad batch %d", nb);
  FB_CHECK_CUDA(launch_pdl(rel_to_bucket_kernel, dim3((rel_rows + 3) / 4, h, 2 * nb), dim3(256), 0, st, gpq, gpk, H, h,
                           bucket_of_rel, rel_rows, rel_off, L, out, ldo, stride_out));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb

extern "C" {
int fb_attn_bwd_prep(const fb_attn_bwd_prep_args* args, void* stream) {
  return fb::attn_bwd_prep(args, reinterpret_cast<cudaStream_t>(stream));
}
int fb_rel_to_bucket(const float* gpq, const float* gpk, int H, int h, const int* bucket_of_rel, int rel_rows,
                     int rel_off, int L, float* out, long long ldo, void* stream) {
  return fb::rel_to_bucket(gpq, gpk, H, h, bucket_of_rel, rel_rows, rel_off, L, out, ldo, 1, 0,
                           reinterpret_cast<cudaStream_t>(stream));
}
int fb_rel_to_bucket_batched(const float* gpq, const float* gpk, int H, int h, const int* bucket_of_rel, int rel_rows,
                             int rel_off, int L, float* out, long long ldo, int nb, long long stride_out, void* stream) {
  return fb::rel_to_bucket(gpq, gpk, H, h, bucket_of_rel, rel_rows, rel_off, L, out, ldo, nb, stride_out,
                           reinterpret_cast<cudaStream_t>(stream));
}
int fb_attn_fwd(const fb_attn_args* args, void* stream) { return fb::attn_fwd(args, reinterpret_cast<cudaStream_t>(stream)); }
int fb_build_rel_tables(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                        int rel_off, int L, void* out_pq, void* out_pk, void* out_pk_rev, void* stream) {
  return fb::build_rel_tables(pqk, ld, H, h, bucket_of_rel, rel_rows, rel_off, L, out_pq, out_pk, out_pk_rev, 1, 0,
                              reinterpret_cast<cudaStream_t>(stream));
}
int fb_build_rel_tables_batched(const void* pqk, long long ld, int H, int h, const int* bucket_of_rel, int rel_rows,
                                int rel_off, int L, void* out_pq, void* out_pk, void* out_pk_rev, int nb,
                                long long stride_pqk, void* stream) {
  return fb::build_rel_tables(pqk, ld, H, h, bucket_of_rel, rel_rows, rel_off, L, out_pq, out_pk, out_pk_rev, nb,
                              stride_pqk, reinterpret_cast<cudaStream_t>(stream));
}
}
