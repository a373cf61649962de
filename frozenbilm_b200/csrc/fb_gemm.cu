// fb_gemm: persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   D[b] = epilogue(alpha * A[b] (M x K) * B[b]^T (N x K)),  fp32 accumulation in TMEM.
//
// Roles inside one 256-thread CTA (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer: cp.async.bulk.tensor (128B-swizzled boxes) -> smem ring
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers
//   warp 2      TMEM allocator (2 accumulator stages so the epilogue of tile i overlaps tile i+1)
//   warps 4-11  epilogue (two warpgroups, each half of the tile columns): tcgen05.ld (lane = output row)
//               -> bias/activation/aux/dropout/residual -> global
//
// Both operands may be K-major (nn.Linear layout) or MN-major (the transposed view), selected in
// the UMMA instruction/smem descriptors, so dX = dY.W and dW = dY^T.X need no explicit transposes.
// Replaces the nn.Linear/bmm call sites listed in include/frozenbilm_b200.h.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "../../include/frozenbilm_b200.h"
#include "fb_common.cuh"

namespace fb {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int A_STAGE_BYTES = BM * BK * 2;
// warps 0-2: TMA / MMA / TMEM alloc, warp 3 idle, then EG epilogue warpgroups (EG = 2: warps 4-11).  EG = 3 (BN = 192
// only: 6 chunks of 32 columns, two per group) gives ALU-heavy epilogues (erf-GELU + second output, GELU' x aux, fp32
// residual) three warps per scheduler instead of two: they are bound by the issue rate of the epilogue warps.
static constexpr int NUM_THREADS = 384;
__host__ __device__ constexpr int gemm_threads(int eg) { return 128 + 128 * eg; }

template <int BN>
struct GemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 192) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int STORE_STAGE_BYTES = 16384;  // one 128 x 32 chunk (fp32) per epilogue warpgroup
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STORE_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int smem_bytes(int eg) { return STAGES * STAGE_BYTES + eg * STORE_STAGE_BYTES + 1024 + 256; }
};

struct GemmParams {
  int M, N, K, batch, batch_inner;
  int a_bi_mul, a_bo_mul, b_bi_mul, b_bo_mul;  // 0 = operand broadcast over that batch level
  long long stride_d_in;
  const uint8_t* rowmask;
  DropSpec drop; int drop_pos;
  int mblocks, nblocks, kblocks, split_k, kblocks_per_split, total_tiles;
  void* D; long long ldd, stride_d; int d_f32;
  float alpha;
  const float* bias; long long stride_bias;   // stride_bias: bias offset per OUTER batch (batched mode only)
  __nv_bfloat16* D_pre; long long ld_pre;
  int act;
  const __nv_bfloat16* aux; long long ld_aux; int aux_op;
  const float* resid; long long ld_resid;
  int accumulate;
  int tma_store;   // epilogue stages 128x32 chunks in smem and writes them with cp.async.bulk.tensor stores
  int m_fast;      // tile order: M block fastest (see decode_tile)
};

__device__ __forceinline__ void decode_tile(const GemmParams& p, int tile, int& split, int& b,
                                            int& mb, int& nb) {
  int t;
  if (p.m_fast) {
    // consecutive tiles (= the CTAs of one wave) share the N block: a B operand that is much larger than L2 (the
    // 128100-row vocabulary table of the MLM head) is then streamed from HBM once instead of once per M block
    mb = tile % p.mblocks;
    t = tile / p.mblocks;
    nb = t % p.nblocks;
    t /= p.nblocks;
  } else {
    nb = tile % p.nblocks;
    t = tile / p.nblocks;
    mb = t % p.mblocks;
    t /= p.mblocks;
  }
  b = t % p.batch;   // combined batch index: outer * batch_inner + inner
  split = t / p.batch;
}

__device__ __forceinline__ void apply_dropout32(const DropSpec& d, int row, int col0, float (&v)[32]) {
  const uint32_t rowkey = drop_row_key(d, (uint32_t)row);   // same decisions as drop_keep8(d, row, group)
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float f[8];
    drop_factors8(rowkey, (uint32_t)(col0 + 8 * g) >> 3, d.thresh16, d.inv_keep, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[8 * g + e] *= f[e];
  }
}

template <int BN>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int b, int row, int col0,
                                               const uint32_t (&acc)[32], bool first_split) {
  if (row >= p.M || col0 >= p.N) return;
  const bool full = (col0 + 32 <= p.N);
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * p.alpha;
  if (first_split) {
    if (p.bias) {
      const float* bias = p.bias + (long long)(b / p.batch_inner) * p.stride_bias;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (full || col0 + i < p.N) v[i] += __ldg(bias + col0 + i);
    }
    if (p.rowmask && p.rowmask[row] == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.0f;
    }
    if (p.drop_pos == 1) apply_dropout32(p.drop, row, col0, v);
    if (p.D_pre) {
      __nv_bfloat16* dst = p.D_pre + (long long)row * p.ld_pre + col0;
      if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                               pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
          reinterpret_cast<uint4*>(dst)[i] = q;
        }
      } else {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < p.N) dst[i] = __float2bfloat16_rn(v[i]);
      }
    }
    if (p.act == FB_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    } else if (p.act == FB_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
    }
    if (p.aux_op != FB_AUX_NONE) {
      const __nv_bfloat16* src = p.aux + (long long)row * p.ld_aux + col0;
      float a[32];
      if (full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
          float2 f0 = unpack_bf16(q.x), f1 = unpack_bf16(q.y), f2 = unpack_bf16(q.z), f3 = unpack_bf16(q.w);
          a[8 * i] = f0.x; a[8 * i + 1] = f0.y; a[8 * i + 2] = f1.x; a[8 * i + 3] = f1.y;
          a[8 * i + 4] = f2.x; a[8 * i + 5] = f2.y; a[8 * i + 6] = f3.x; a[8 * i + 7] = f3.y;
        }
      } else {
        for (int i = 0; i < 32; ++i) a[i] = (col0 + i < p.N) ? __bfloat162float(src[i]) : 0.0f;
      }
      if (p.aux_op == FB_AUX_MUL_GELU_GRAD) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= gelu_erf_grad(a[i]);
      } else if (p.aux_op == FB_AUX_MUL_RELU_MASK) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (a[i] > 0.0f) ? v[i] : 0.0f;
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += a[i];
      }
    }
    if (p.drop_pos == 2) apply_dropout32(p.drop, row, col0, v);
    if (p.resid) {
      const float* src = p.resid + (long long)row * p.ld_resid + col0;
      if (full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 q = __ldg(reinterpret_cast<const float4*>(src) + i);
          v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
        }
      } else {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < p.N) v[i] += src[i];
      }
    }
  }
  if (p.d_f32) {
    float* dst = reinterpret_cast<float*>(p.D) + (long long)(b / p.batch_inner) * p.stride_d +
                 (long long)(b % p.batch_inner) * p.stride_d_in + (long long)row * p.ldd + col0;
    if (p.split_k > 1) {
      for (int i = 0; i < 32; ++i)
        if (full || col0 + i < p.N) atomicAdd(dst + i, v[i]);
    } else if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 q = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (p.accumulate) {
          float4 o = reinterpret_cast<float4*>(dst)[i];
          q.x += o.x; q.y += o.y; q.z += o.z; q.w += o.w;
        }
        reinterpret_cast<float4*>(dst)[i] = q;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) dst[i] = p.accumulate ? dst[i] + v[i] : v[i];
    }
  } else {
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.D) + (long long)(b / p.batch_inner) * p.stride_d +
                         (long long)(b % p.batch_inner) * p.stride_d_in + (long long)row * p.ldd + col0;
    if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                             pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        reinterpret_cast<uint4*>(dst)[i] = q;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) dst[i] = __float2bfloat16_rn(v[i]);
    }
  }
}


// value pipeline of the epilogue for one 32-column chunk (no stores), in two halves so that the pre-activation
// can be stored between them: epi_pre = alpha, bias, row mask, early dropout; epi_post = activation, aux,
// late dropout, residual.
__device__ __forceinline__ void epi_pre(const GemmParams& p, int row, int col0, const uint32_t (&acc)[32],
                                        float (&v)[32]) {
  const bool in_rows = row < p.M;
  const bool full = (col0 + 32 <= p.N);
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * p.alpha;
  if (p.bias) {
    if (full && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 q = __ldg(b4 + i);
        v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) v[i] += __ldg(p.bias + col0 + i);
    }
  }
  if (p.rowmask && in_rows && p.rowmask[row] == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.0f;
  }
  if (p.drop_pos == 1) apply_dropout32(p.drop, row, col0, v);
}
__device__ __forceinline__ void epi_post(const GemmParams& p, int row, int col0, float (&v)[32]) {
  const bool in_rows = row < p.M;
  const bool full = (col0 + 32 <= p.N);
  if (p.act == FB_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  } else if (p.act == FB_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  }
  if (p.aux_op != FB_AUX_NONE && in_rows) {
    const __nv_bfloat16* src = p.aux + (long long)row * p.ld_aux + col0;
    float a[32];
    if (full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 q = __ldg(reinterpret_cast<const uint4*>(src) + i);
        float2 f0 = unpack_bf16(q.x), f1 = unpack_bf16(q.y), f2 = unpack_bf16(q.z), f3 = unpack_bf16(q.w);
        a[8 * i] = f0.x; a[8 * i + 1] = f0.y; a[8 * i + 2] = f1.x; a[8 * i + 3] = f1.y;
        a[8 * i + 4] = f2.x; a[8 * i + 5] = f2.y; a[8 * i + 6] = f3.x; a[8 * i + 7] = f3.y;
      }
    } else {
      for (int i = 0; i < 32; ++i) a[i] = (col0 + i < p.N) ? __bfloat162float(src[i]) : 0.0f;
    }
    if (p.aux_op == FB_AUX_MUL_GELU_GRAD) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= gelu_erf_grad(a[i]);
    } else if (p.aux_op == FB_AUX_MUL_RELU_MASK) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (a[i] > 0.0f) ? v[i] : 0.0f;
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += a[i];
    }
  }
  if (p.drop_pos == 2) apply_dropout32(p.drop, row, col0, v);
  if (p.resid && in_rows) {
    const float* src = p.resid + (long long)row * p.ld_resid + col0;
    if (full && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 q = __ldg(reinterpret_cast<const float4*>(src) + i);
        v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
      }
    } else {
      for (int i = 0; i < 32; ++i)
        if (col0 + i < p.N) v[i] += src[i];
    }
  }
}

// stage one row segment of a 128 x 32 chunk into the warpgroup's smem buffer (TMA swizzle layouts):
// bf16: 64-byte rows, 16B unit u -> u ^ ((r >> 1) & 3)   (CU_TENSOR_MAP_SWIZZLE_64B)
// fp32: 128-byte rows, unit u -> u ^ (r & 7)              (CU_TENSOR_MAP_SWIZZLE_128B)
__device__ __forceinline__ void stage_row_bf16(uint8_t* buf, int r, const float (&v)[32]) {
  uint8_t* rowp = buf + r * 64;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 q = make_uint4(pack_bf16(v[8 * u], v[8 * u + 1]), pack_bf16(v[8 * u + 2], v[8 * u + 3]),
                               pack_bf16(v[8 * u + 4], v[8 * u + 5]), pack_bf16(v[8 * u + 6], v[8 * u + 7]));
    *reinterpret_cast<uint4*>(rowp + ((u ^ ((r >> 1) & 3)) << 4)) = q;
  }
}
__device__ __forceinline__ void stage_row_f32(uint8_t* buf, int r, const float (&v)[32]) {
  uint8_t* rowp = buf + r * 128;
#pragma unroll
  for (int u = 0; u < 8; ++u)
    *reinterpret_cast<float4*>(rowp + ((u ^ (r & 7)) << 4)) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
}

// CL = 2: thread-block cluster of two CTAs working on M-adjacent tiles of the same N block.  Each CTA fetches its own
// A tile and HALF of the shared B tile, multicast into both CTAs' shared memory, so the L2 -> SM traffic per
// 128 x BN x 64 step drops from 16 + BN/8 KB to 16 + BN/16 KB (the 1-CTA kernel sits at the ~42 B/clk/SM L2 fabric
// limit: 87 FLOP/B for BN = 256).  A pipeline stage is released to BOTH producers (multicast commit), the accumulator
// / epilogue side is unchanged.  Host side: only for batch = 1, split_k = 1, an even number of M blocks, K-major A.
template <int CL>
__device__ __forceinline__ void decode_tile_cl(const GemmParams& p, int t, uint32_t crank, int& split, int& b, int& mb,
                                               int& nb) {
  if constexpr (CL == 1) {
    decode_tile(p, t, split, b, mb, nb);
  } else {
    nb = t % p.nblocks;
    mb = (t / p.nblocks) * CL + (int)crank;
    b = 0;
    split = 0;
  }
}

template <int BN, bool A_MN, bool B_MN, bool COAL, int CL, int EG = 2>
__global__ void __launch_bounds__(gemm_threads(EG), 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_pre,
            const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms) by pointer arithmetic on the __shared__ array, so that the compiler
  // keeps the shared address space (STS/LDS instead of generic ST/LD)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + Cfg::STAGES * A_STAGE_BYTES;
  uint8_t* smem_store = smem + Cfg::STAGES * Cfg::STAGE_BYTES;   // 2 x 16 KB epilogue staging (TMA stores)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + EG * Cfg::STORE_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);
  volatile int* pf_started = reinterpret_cast<volatile int*>(bars + 2 * Cfg::STAGES + 5);   // epilogues started (COAL)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  // persistent tile sequence of this CTA (CL = 2: of its cluster; both CTAs walk the same sequence of tile pairs)
  const int tile_first = (CL > 1) ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const int tile_step = (CL > 1) ? (int)(gridDim.x / CL) : (int)gridDim.x;
  const int tile_total = (CL > 1) ? p.total_tiles / CL : p.total_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL);   // released by the MMA warps of every CTA that multicasts into this stage
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * EG);
    }
    *pf_started = 0;
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_prologue();   // operands / epilogue inputs come from the previous kernel of the stream

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < tile_total; tile += tile_step) {
        int split, b, mb, nb;
        decode_tile_cl<CL>(p, tile, crank, split, b, mb, nb);
        const int m0 = mb * BM, n0 = nb * BN;
        const int k_begin = split * p.kblocks_per_split;
        const int k_end = min(k_begin + p.kblocks_per_split, p.kblocks);
        for (int kb = k_begin; kb < k_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES;
          const int bi = b % p.batch_inner, bo = b / p.batch_inner;
          const int abi = bi * p.a_bi_mul, abo = bo * p.a_bo_mul, bbi = bi * p.b_bi_mul, bbo = bo * p.b_bo_mul;
          if constexpr (!A_MN) {
            tma_load_4d(sa, &map_a, &full_bar[stage], kb * BK, m0, abi, abo);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_4d(sa + c * (64 * BK * 2), &map_a, &full_bar[stage], m0 + c * 64, kb * BK, abi, abo);
          }
          if constexpr (CL == 1) {
            if constexpr (!B_MN) {
              tma_load_4d(sb, &map_b, &full_bar[stage], kb * BK, n0, bbi, bbo);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_4d(sb + c * (64 * BK * 2), &map_b, &full_bar[stage], n0 + c * 64, kb * BK, bbi, bbo);
            }
          } else {
            // this CTA's share of the B tile, multicast to both CTAs (the peer sends the rest)
            constexpr uint16_t kAll = (1u << CL) - 1u;
            if constexpr (!B_MN) {
              constexpr int ROWS = BN / CL;   // map_b was built with a box of BN / CL rows
              tma_load_4d_mc(sb + crank * (ROWS * BK * 2), &map_b, &full_bar[stage], kb * BK, n0 + (int)crank * ROWS, bbi,
                             bbo, kAll);
            } else {
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                if ((c % CL) == (int)crank)
                  tma_load_4d_mc(sb + c * (64 * BK * 2), &map_b, &full_bar[stage], n0 + c * 64, kb * BK, bbi, bbo, kAll);
            }
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      // K-major: rows of 128 B, 8-row groups 1024 B apart (SBO); one UMMA_K = 32 B inside the row.
      // MN-major: [64 mn x 64 k] chunks of 8192 B (LBO between chunks), 8-k groups 1024 B apart
      // (SBO); one UMMA_K = 16 k-rows = 2048 B.
      constexpr uint32_t A_LBO = A_MN ? 64 * BK * 2 : 16, B_LBO = B_MN ? 64 * BK * 2 : 16;
      constexpr uint32_t A_KSTEP = A_MN ? 2048 : 32, B_KSTEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile_first; tile < tile_total; tile += tile_step, ++it) {
        int split, b, mb, nb;
        decode_tile_cl<CL>(p, tile, crank, split, b, mb, nb);
        const int k_begin = split * p.kblocks_per_split;
        const int k_end = min(k_begin + p.kblocks_per_split, p.kblocks);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = k_begin; kb < k_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(sa + k * A_KSTEP, A_LBO, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sb + k * B_KSTEP, B_LBO, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > k_begin || k > 0) ? 1u : 0u);
          }
          // frees the smem slot once these MMAs retire (CL = 2: in both CTAs, whose producers both write into it)
          if constexpr (CL == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << CL) - 1u));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== epilogue-operand prefetcher (coalesced-operand instantiation) =====================
    // The aux / residual rows of a tile are pulled into L2 ahead of the epilogue warps, whose loads are the critical
    // path of the K = 192 adapter GEMMs (one dependent HBM round trip per 32-column chunk and warp).  Tile `it` is
    // requested once the epilogue of tile it - 1 has started (a monotonic counter in shared memory written by the
    // first epilogue warp: no barrier phase can be missed).  Plain per-line L2 prefetches: per-row bulk prefetches
    // (256 per tile through the TMA unit) could not keep up with 10 us tiles.  A hint only: correctness does not
    // depend on it and the wait is bounded.
    if constexpr (COAL) {
      const bool has_aux = p.aux_op != FB_AUX_NONE, has_res = p.resid != nullptr;
      int it = 0;
      for (int tile = tile_first; tile < tile_total; tile += tile_step, ++it) {
        int split, b, mb, nb;
        decode_tile_cl<CL>(p, tile, crank, split, b, mb, nb);
        for (uint32_t spins = 0; spins < (1u << 18) && *pf_started < it; ++spins) __nanosleep(64);
        if (split != 0) continue;
        const int n0 = nb * BN;
        const int ncols = min(BN, p.N - n0);
        const int rows = min(BM, p.M - mb * BM);
        if (has_res) {
          const int lpr = (ncols * 4 + 127) >> 7;                    // 128-byte lines per row segment
          const char* base = reinterpret_cast<const char*>(p.resid + (long long)mb * BM * p.ld_resid + n0);
          for (int i = lane; i < rows * lpr; i += 32) {
            const int r = i / lpr, l = i - r * lpr;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)r * p.ld_resid * 4 + l * 128));
          }
        }
        if (has_aux) {
          const int lpr = (ncols * 2 + 127) >> 7;
          const char* base = reinterpret_cast<const char*>(p.aux + (long long)mb * BM * p.ld_aux + n0);
          for (int i = lane; i < rows * lpr; i += 32) {
            const int r = i / lpr, l = i - r * lpr;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (long long)r * p.ld_aux * 2 + l * 128));
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int wq = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;    // which share of the tile's 32-column chunks this warpgroup drains
    constexpr int CHUNKS = BN / 32, GRP_CHUNKS = (CHUNKS + EG - 1) / EG;
    const int c_begin = grp * GRP_CHUNKS, c_end = (c_begin + GRP_CHUNKS < CHUNKS) ? c_begin + GRP_CHUNKS : CHUNKS;
    uint32_t nstore = 0;
    int it = 0;
    for (int tile = tile_first; tile < tile_total; tile += tile_step, ++it) {
      int split, b, mb, nb;
      decode_tile_cl<CL>(p, tile, crank, split, b, mb, nb);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if constexpr (COAL) {
        if (warp == 4 && lane == 0) *pf_started = it + 1;   // lets the prefetcher warp move on to the next tile
      }
      const int row = mb * BM + wq * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * Cfg::ACC_STRIDE;
      if (!p.tma_store) {
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          epilogue_chunk<BN>(p, b, row, nb * BN + c * 32, acc, split == 0);
        }
      } else {
        // coalesced path: each warpgroup stages a 128 x 32 chunk in smem and one thread issues a TMA store
        // every epilogue warp owns a 4 KB staging box (its 32 rows x 32 columns) and issues its own TMA stores: no
        // block-level barrier in the epilogue, a warp never waits for the slowest of its warpgroup (the
        // warpgroup-wide version spent ~16 named-barrier + proxy-fence rounds per 128 x 256 tile)
        uint8_t* buf0 = smem_store + (warp - 4) * 4096;
        const int r = lane;
        const bool leader = (lane == 0);
        const int n_pass = p.D_pre ? 2 : 1;
        // bf16 chunks are 8 KB: two of them double-buffer inside the warpgroup's 16 KB so that a store drains
        // while the next chunk is computed; fp32 chunks (16 KB) are single-buffered
        const bool dbl = !p.d_f32;
        // Coalesced epilogue operands.  With thread = output row, loading the aux (bf16) / residual (fp32) segment of
        // the own row touches 32 different 128-byte lines per instruction; for the K = 192 adapter GEMMs that LSU work
        // was the whole run time.  Here 4 (aux) / 8 (residual) lanes share one row, so an instruction covers complete
        // lines, and the warp hands the pieces to their row owners through ITS OWN 32 rows of the output staging
        // buffer (private to the warp between the "previous store has been read" barrier and the staging of this
        // chunk).  Only for full, 16-byte aligned chunks without a pre-activation output; everything else takes the
        // generic path below.
        // COAL is a separate instantiation (chosen by the host when the operands are 16-byte aligned and there is no
        // pre-activation output) so that the register allocation of the plain path is untouched.
        const bool has_aux = p.aux_op != FB_AUX_NONE, has_res = p.resid != nullptr;
        constexpr bool coalesced_ops = COAL;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = nb * BN + c * 32;
          if (col0 >= p.N) break;
          if (coalesced_ops && split == 0 && col0 + 32 <= p.N) if constexpr (COAL) {
            const int r0 = mb * BM + wq * 32;   // first output row of this warp
            uint4 ax[4];
            float4 rs[8];
            if (has_aux) {
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int rw = r0 + 8 * m + (lane >> 2);
                ax[m] = (rw < p.M) ? __ldg(reinterpret_cast<const uint4*>(p.aux + (long long)rw * p.ld_aux + col0) + (lane & 3))
                                   : make_uint4(0u, 0u, 0u, 0u);
              }
            }
            if (has_res) {
#pragma unroll
              for (int m = 0; m < 8; ++m) {
                const int rw = r0 + 4 * m + (lane >> 3);
                rs[m] = (rw < p.M) ? __ldg(reinterpret_cast<const float4*>(p.resid + (long long)rw * p.ld_resid + col0) + (lane & 7))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
            uint32_t acc[32];
            tmem_ld_32x32(taddr + c * 32, acc);
            tmem_ld_wait();
            float v[32];
            epi_pre(p, row, col0, acc, v);
            if (p.act == FB_ACT_GELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
            } else if (p.act == FB_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
            }
            uint8_t* buf = buf0 + ((dbl && (nstore & 1)) ? 2048 : 0);
            ++nstore;
            if (leader) {                            // the store that last used this buffer has finished reading it
              if (dbl) tma_store_wait_read1(); else tma_store_wait_read0();
            }
            __syncwarp();
            uint8_t* slice = buf;   // this warp's staging box
            if (has_aux) {
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int rr = 8 * m + (lane >> 2), pc = lane & 3;
                *reinterpret_cast<uint4*>(slice + rr * 64 + ((pc ^ ((rr >> 1) & 3)) << 4)) = ax[m];
              }
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const uint4 q = *reinterpret_cast<const uint4*>(slice + lane * 64 + ((u ^ ((lane >> 1) & 3)) << 4));
                const float2 f0 = unpack_bf16(q.x), f1 = unpack_bf16(q.y), f2 = unpack_bf16(q.z), f3 = unpack_bf16(q.w);
                const float a8[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float& x = v[8 * u + e];
                  if (p.aux_op == FB_AUX_MUL_GELU_GRAD) x *= gelu_erf_grad(a8[e]);
                  else if (p.aux_op == FB_AUX_MUL_RELU_MASK) x = (a8[e] > 0.0f) ? x : 0.0f;
                  else x += a8[e];
                }
              }
              __syncwarp();
            }
            if (p.drop_pos == 2) apply_dropout32(p.drop, row, col0, v);
            if (has_res) {
#pragma unroll
              for (int m = 0; m < 8; ++m) {
                const int rr = 4 * m + (lane >> 3), pc = lane & 7;
                *reinterpret_cast<float4*>(slice + rr * 128 + ((pc ^ (rr & 7)) << 4)) = rs[m];
              }
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float4 q = *reinterpret_cast<const float4*>(slice + lane * 128 + ((u ^ (lane & 7)) << 4));
                v[4 * u] += q.x; v[4 * u + 1] += q.y; v[4 * u + 2] += q.z; v[4 * u + 3] += q.w;
              }
              __syncwarp();
            }
            if (!p.d_f32) stage_row_bf16(buf, r, v); else stage_row_f32(buf, r, v);
            fence_proxy_async_smem();
            __syncwarp();
            if (leader) {
              if (p.tma_store == 2)
                tma_reduce_add_2d(&map_d, buf, col0, mb * BM + wq * 32);
              else
                tma_store_2d(&map_d, buf, col0, mb * BM + wq * 32);
              tma_store_commit();
            }
            continue;
          }
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          float v[32];
          if (split == 0) {
            epi_pre(p, row, col0, acc, v);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(acc[i]) * p.alpha;
          }
#pragma unroll 1
          for (int pass = 0; pass < n_pass; ++pass) {
            const bool pre_only = (n_pass == 2 && pass == 0);
            if (!pre_only && split == 0) epi_post(p, row, col0, v);
            uint8_t* buf = buf0 + ((dbl && (nstore & 1)) ? 2048 : 0);
            ++nstore;
            if (leader) {                            // the store that last used this buffer has finished reading it
              if (dbl) tma_store_wait_read1(); else tma_store_wait_read0();
            }
            __syncwarp();
            if (pre_only || !p.d_f32) stage_row_bf16(buf, r, v); else stage_row_f32(buf, r, v);
            fence_proxy_async_smem();
            __syncwarp();
            if (leader) {
              if (p.tma_store == 2 && !pre_only)
                tma_reduce_add_2d(&map_d, buf, col0, mb * BM + wq * 32);   // D += tile (split-K / accumulate)
              else
                tma_store_2d(pre_only ? &map_pre : &map_d, buf, col0, mb * BM + wq * 32);
              tma_store_commit();
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();   // no CTA retires while its peer may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_pdl{-1};   // -1: not yet read from the environment (FB_PDL, default on)
static int pdl_env() {
  static const int v = [] { const char* e = getenv("FB_PDL"); return (e == nullptr || atoi(e) != 0) ? 1 : 0; }();
  return v;
}
bool pdl_enabled() {
  const int v = g_pdl.load(std::memory_order_relaxed);
  return (v < 0 ? pdl_env() : (v && pdl_env())) == 1;
}

int num_sms() {
  static int n = -1;
  if (n < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    n = prop.multiProcessorCount;
  }
  return n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

int make_tmap_store_2d(CUtensorMap* out, const void* ptr, int is_f32, uint64_t cols, uint64_t rows, uint64_t ld_elems) {
  PFN_encodeTiled enc = get_encode();
  FB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  const uint64_t esz = is_f32 ? 4 : 2;
  cuuint64_t d[2] = {cols, rows};
  cuuint64_t st[1] = {ld_elems * esz};
  cuuint32_t bx[2] = {32, 32}, es[2] = {1, 1};   // one epilogue warp's box: 32 columns x its 32 rows
  CUresult r = enc(out, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(store) failed with CUresult %d", (int)r);
  return 0;
}

int make_tmap_bf16_nd(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_elems, const uint32_t* box) {
  PFN_encodeTiled enc = get_encode();
  FB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  FB_REQUIRE(rank >= 2 && rank <= 4, "tensor map rank must be 2..4");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  cuuint64_t d[4];
  cuuint64_t st[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) {
    st[i] = strides_elems[i] * 2;
    FB_REQUIRE(st[i] % 16 == 0, "TMA stride %d (%llu bytes) must be a multiple of 16 bytes", i,
               (unsigned long long)st[i]);
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %d) failed with CUresult %d", rank, (int)r);
  return 0;
}

// 4-D operand map {inner, rows, batch_inner, batch_outer} with a {64, box_rows, 1, 1} box.
static int make_tmap_operand(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t ld,
                             int batch_in, long long stride_in, int batch_out, long long stride_out,
                             uint32_t box_rows) {
  FB_REQUIRE(ld % 8 == 0, "fb_gemm: operand leading dimension must be a multiple of 8 elements (got %llu)",
             (unsigned long long)ld);
  // a zero batch stride means "broadcast": that level is collapsed to size 1 and the kernel always
  // addresses coordinate 0 (GemmParams::*_mul)
  if (stride_in == 0) batch_in = 1;
  if (stride_out == 0) batch_out = 1;
  uint64_t dims[4] = {inner, rows, (uint64_t)std::max(batch_in, 1), (uint64_t)std::max(batch_out, 1)};
  uint64_t strides[3] = {ld, batch_in > 1 ? (uint64_t)stride_in : ld * rows,
                         batch_out > 1 ? (uint64_t)stride_out : ld * rows * dims[2]};
  // dummy strides of size-1 dims still have to be 16-byte multiples: ld is, so ld*rows is
  uint32_t box[4] = {64, box_rows, 1, 1};
  return make_tmap_bf16_nd(out, ptr, 4, dims, strides, box);
}

template <int BN, bool A_MN, bool B_MN, bool COAL, int CL, int EG = 2>
static int launch_gemm_impl(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const CUtensorMap& mpre,
                            const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  static int max_clusters = 0;
  constexpr int NUM_THREADS = gemm_threads(EG);
  constexpr int SMEM = Cfg::smem_bytes(EG);
  static_assert(SMEM <= 232448, "shared memory budget");
  auto kern = gemm_kernel<BN, A_MN, B_MN, COAL, CL, EG>;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (CL > 1) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(CL * 64); qc.blockDim = dim3(NUM_THREADS); qc.dynamicSmemBytes = SMEM;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      FB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &qc));
      FB_REQUIRE(max_clusters > 0, "fb_gemm: no cluster of %d CTAs fits", CL);
    }
    configured = true;
  }
  if (CL == 1) {
    int grid = std::min(p.total_tiles, num_sms());
    FB_REQUIRE(grid > 0, "no CUDA device / empty GEMM");
    FB_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(NUM_THREADS), SMEM, st, ma, mb, md, mpre, p));
  } else {
    const int clusters = std::min(p.total_tiles / CL, max_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    FB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, md, mpre, p));
  }
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// cluster_ok: the host found the shape eligible for the 2-CTA multicast kernel AND built map_b with half-height boxes
template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md, const CUtensorMap& mpre,
                       const GemmParams& p, bool cluster_ok, cudaStream_t st) {
  // coalesced epilogue operands (see the kernel): aligned aux / residual, TMA-store epilogue, no pre-activation output
  const bool has_aux = p.aux_op != FB_AUX_NONE, has_res = p.resid != nullptr;
  // (the GELU-gradient / ReLU-mask epilogues are ALU-bound and measured slightly slower this way: generic path)
  const bool coal =
      p.tma_store && !p.D_pre && (has_aux || has_res) && (!has_res || p.d_f32) && (!has_aux || p.aux_op == FB_AUX_ADD) &&
      (!has_aux || (((reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) && (p.ld_aux & 7) == 0)) &&
      (!has_res || (((reinterpret_cast<uintptr_t>(p.resid) & 15) == 0) && (p.ld_resid & 3) == 0));
  if constexpr (!A_MN && (BN == 256 || BN == 192)) {
    if (cluster_ok) {
      if constexpr (BN == 192) {
        // ALU-heavy epilogues get a third epilogue warpgroup (see gemm_threads)
        static const int eg_env = [] { const char* e = getenv("FB_GEMM_EG"); return e ? atoi(e) : 0; }();
        // (the coalesced-operand instantiation spills at the 128-register budget of 512 threads: only on request)
        const bool heavy = p.tma_store == 1 && !coal && (p.act == FB_ACT_GELU || p.aux_op == FB_AUX_MUL_GELU_GRAD);
        if (eg_env == 3 || (eg_env == 0 && heavy)) {
          if (coal) return launch_gemm_impl<BN, A_MN, B_MN, true, 2, 3>(ma, mb, md, mpre, p, st);
          return launch_gemm_impl<BN, A_MN, B_MN, false, 2, 3>(ma, mb, md, mpre, p, st);
        }
      }
      if (coal) return launch_gemm_impl<BN, A_MN, B_MN, true, 2>(ma, mb, md, mpre, p, st);
      return launch_gemm_impl<BN, A_MN, B_MN, false, 2>(ma, mb, md, mpre, p, st);
    }
  }
  if (coal) return launch_gemm_impl<BN, A_MN, B_MN, true, 1>(ma, mb, md, mpre, p, st);
  return launch_gemm_impl<BN, A_MN, B_MN, false, 1>(ma, mb, md, mpre, p, st);
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& md,
                          const CUtensorMap& mpre, const GemmParams& p, bool cluster_ok, cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(ma, mb, md, mpre, p, cluster_ok, st);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(ma, mb, md, mpre, p, cluster_ok, st);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(ma, mb, md, mpre, p, false, st);
  return launch_gemm<BN, true, true>(ma, mb, md, mpre, p, false, st);
}

static int pick_block_n(int M, int N, int batch, int split_k) {
  const int nsm = std::max(1, num_sms());
  const int mblocks = (M + BM - 1) / BM;
  int best = 64;
  double best_cost = 1e30;
  const int cands[4] = {256, 192, 128, 64};
  for (int bn : cands) {
    const long long tiles = (long long)mblocks * ((N + bn - 1) / bn) * batch * split_k;
    const long long waves = (tiles + nsm - 1) / nsm;
    const double cost = (double)waves * (bn + 24);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// How many 2-CTA clusters of the GEMM kernel can be resident at once (GPC layout dependent); 0 when the query fails.
int cluster_capacity() {
  static int cap = -1;
  if (cap < 0) {
    using Cfg = GemmCfg<256>;
    auto kern = gemm_kernel<256, false, false, false, 2>;
    cap = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(2 * 64); qc.blockDim = dim3(NUM_THREADS); qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) == cudaSuccess) cap = n;
    }
    (void)cudaGetLastError();
  }
  return cap;
}

int gemm(const fb_gemm_args* a, cudaStream_t st) {
  FB_REQUIRE(a != nullptr, "fb_gemm: null args");
  FB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch >= 1, "fb_gemm: bad shape M=%d N=%d K=%d batch=%d", a->M,
             a->N, a->K, a->batch);
  FB_REQUIRE(a->A && a->B && a->D, "fb_gemm: null operand");
  const int split_k = a->split_k > 1 ? a->split_k : 1;
  if (split_k > 1) FB_REQUIRE(a->d_dtype == FB_DT_F32, "fb_gemm: split_k needs an fp32 output");
  if (a->accumulate) FB_REQUIRE(a->d_dtype == FB_DT_F32, "fb_gemm: accumulate needs an fp32 output");
  const int b_out = a->batch, b_in = a->batch_inner > 1 ? a->batch_inner : 1;
  const int batch_total = b_out * b_in;
  if (batch_total > 1)
    FB_REQUIRE(!a->D_pre && !a->aux && !a->resid && !a->rowmask && !a->drop_pos,
               "fb_gemm: batched mode supports bias/act only");
  int bn = a->block_n ? a->block_n : pick_block_n(a->M, a->N, batch_total, split_k);
  {
    // The GELU' x aux epilogue (dX of the FFN down-projection) runs on 128 x 192 tiles with three epilogue warpgroups
    // (launch_gemm), which 128 x 256 tiles cannot have (8 chunks do not split three ways, four groups do not fit the
    // shared-memory budget): 202.5 -> 197.6 us at config 2.  The erf-GELU + pre-activation epilogue of the
    // up-projection measured no gain (194.3 -> 196.3 us: it is not bound by the epilogue warps' issue rate) and keeps
    // 128 x 256 tiles; FB_GEMM_EG=3 forces the three-group kernels for both, FB_GEMM_EG=2 the two-group ones.
    static const int eg_env = [] { const char* e = getenv("FB_GEMM_EG"); return e ? atoi(e) : 0; }();
    if (!a->block_n && bn == 256 && eg_env != 2 && a->N % 192 == 0 && batch_total == 1 && split_k == 1 &&
        !a->a_mn_major && ((a->aux && a->aux_op == FB_AUX_MUL_GELU_GRAD) || (eg_env == 3 && a->act == FB_ACT_GELU)))
      bn = 192;
  }
  FB_REQUIRE(bn == 64 || bn == 128 || bn == 192 || bn == 256, "fb_gemm: block_n must be 64/128/192/256");

  // 2-CTA multicast kernel (see gemm_kernel): big unbatched GEMMs with an even number of M blocks and K-major A
  static const bool cluster_enabled = (getenv("FB_GEMM_CLUSTER") == nullptr) || (atoi(getenv("FB_GEMM_CLUSTER")) != 0);
  const int mblocks_h = (a->M + BM - 1) / BM, nblocks_h = (a->N + bn - 1) / bn;
  // (only when the cluster kernel can occupy (nearly) every SM: on a part whose GPCs leave SMs unpaired the 1-CTA
  // kernel with all SMs is the better choice)
  const bool cluster_ok = cluster_enabled && batch_total == 1 && split_k == 1 && !a->a_mn_major && (bn == 256 || bn == 192) &&
                          (mblocks_h % 2 == 0) && ((long long)mblocks_h * nblocks_h >= 2LL * num_sms()) &&
                          (2 * cluster_capacity() >= num_sms() - 2);

  CUtensorMap ma, mb;
  int rc;
  if (!a->a_mn_major)
    rc = make_tmap_operand(&ma, a->A, a->K, a->M, a->lda, b_in, a->stride_a_in, b_out, a->stride_a, BM);
  else
    rc = make_tmap_operand(&ma, a->A, a->M, a->K, a->lda, b_in, a->stride_a_in, b_out, a->stride_a, BK);
  if (rc) return rc;
  if (!a->b_mn_major)
    rc = make_tmap_operand(&mb, a->B, a->K, a->N, a->ldb, b_in, a->stride_b_in, b_out, a->stride_b,
                           cluster_ok ? bn / 2 : bn);   // the cluster kernel loads half-height boxes
  else
    rc = make_tmap_operand(&mb, a->B, a->N, a->K, a->ldb, b_in, a->stride_b_in, b_out, a->stride_b, BK);
  if (rc) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K; p.batch = batch_total; p.batch_inner = b_in;
  p.stride_d_in = a->stride_d_in;
  p.a_bi_mul = (b_in > 1 && a->stride_a_in != 0) ? 1 : 0;
  p.a_bo_mul = (b_out > 1 && a->stride_a != 0) ? 1 : 0;
  p.b_bi_mul = (b_in > 1 && a->stride_b_in != 0) ? 1 : 0;
  p.b_bo_mul = (b_out > 1 && a->stride_b != 0) ? 1 : 0;
  p.rowmask = a->rowmask;
  p.drop = make_drop(a->drop.p, a->drop.seed, a->drop.offset);
  p.drop_pos = (a->drop.p > 0.0f) ? a->drop_pos : 0;
  p.mblocks = (a->M + BM - 1) / BM;
  p.nblocks = (a->N + bn - 1) / bn;
  p.kblocks = (a->K + BK - 1) / BK;
  p.split_k = std::min(split_k, p.kblocks);
  p.kblocks_per_split = (p.kblocks + p.split_k - 1) / p.split_k;
  p.split_k = (p.kblocks + p.kblocks_per_split - 1) / p.kblocks_per_split;  // no empty splits
  p.total_tiles = p.mblocks * p.nblocks * p.batch * p.split_k;
  p.D = a->D; p.ldd = a->ldd; p.stride_d = a->stride_d; p.d_f32 = (a->d_dtype == FB_DT_F32);
  p.alpha = a->alpha;
  p.bias = a->bias;
  p.stride_bias = (batch_total > 1) ? a->stride_bias : 0;
  p.D_pre = reinterpret_cast<__nv_bfloat16*>(a->D_pre); p.ld_pre = a->ld_pre;
  p.act = a->act;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(a->aux); p.ld_aux = a->ld_aux;
  p.aux_op = a->aux ? a->aux_op : FB_AUX_NONE;
  p.resid = a->resid; p.ld_resid = a->ld_resid;
  p.accumulate = (a->accumulate || split_k > 1) ? 1 : 0;
  {
    // B far larger than L2 and A small enough to stay there: walk the M blocks of an N block first
    static const int mf_env = [] { const char* e = getenv("FB_GEMM_MFAST"); return e ? atoi(e) : -1; }();
    const double a_bytes = 2.0 * a->M * a->K, b_bytes = 2.0 * a->N * a->K;
    const bool want = batch_total == 1 && !cluster_ok && p.mblocks > 1 && b_bytes > 96e6 && a_bytes < 48e6;
    p.m_fast = (mf_env >= 0) ? (mf_env && batch_total == 1 && !cluster_ok) : want;
  }

  // coalesced TMA-store epilogue whenever the output is a plain (unbatched, non-accumulating) matrix with
  // 16-byte aligned rows; otherwise the per-thread direct-store path
  CUtensorMap md, mpre;
  memset(&md, 0, sizeof(md));
  memset(&mpre, 0, sizeof(mpre));
  const uint64_t dsz = p.d_f32 ? 4 : 2;
  const bool d_ok = ((uintptr_t)a->D % 16 == 0) && ((uint64_t)a->ldd * dsz % 16 == 0);
  const bool pre_ok = !a->D_pre || (((uintptr_t)a->D_pre % 16 == 0) && ((uint64_t)a->ld_pre * 2 % 16 == 0));
  p.tma_store = (batch_total == 1 && d_ok && pre_ok && a->M >= 32) ? 1 : 0;
  if (p.tma_store && (p.split_k > 1 || p.accumulate)) p.tma_store = (p.d_f32 && !a->D_pre) ? 2 : 0;
  if (p.tma_store) {
    if ((rc = make_tmap_store_2d(&md, a->D, p.d_f32, a->N, a->M, a->ldd))) return rc;
    if (a->D_pre && (rc = make_tmap_store_2d(&mpre, a->D_pre, 0, a->N, a->M, a->ld_pre))) return rc;
  }
  switch (bn) {
    case 64: return dispatch_major<64>(a->a_mn_major, a->b_mn_major, ma, mb, md, mpre, p, false, st);
    case 128: return dispatch_major<128>(a->a_mn_major, a->b_mn_major, ma, mb, md, mpre, p, false, st);
    case 192: return dispatch_major<192>(a->a_mn_major, a->b_mn_major, ma, mb, md, mpre, p, cluster_ok, st);
    default: return dispatch_major<256>(a->a_mn_major, a->b_mn_major, ma, mb, md, mpre, p, cluster_ok, st);
  }
}

}  // namespace fb

extern "C" {
const char* fb_last_error(void) { return fb::last_error(); }
int fb_abi_version(void) { return FB_ABI_VERSION; }
long long fb_launch_count(void) { return fb::g_launches.load(); }
int fb_gemm_cluster_capacity(void) { return fb::cluster_capacity(); }
int fb_set_pdl(int on) {
  const int prev = fb::pdl_enabled() ? 1 : 0;
  fb::g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}
int fb_gemm(const fb_gemm_args* args, void* stream) { return fb::gemm(args, reinterpret_cast<cudaStream_t>(stream)); }
}
