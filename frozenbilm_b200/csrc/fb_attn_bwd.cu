// Fused disentangled-attention backward for sm_100a (tcgen05 + TMA), head_dim = 64.
//
// Autograd of DisentangledSelfAttention.forward + disentangled_attention_bias + XSoftmax + XDropout
// (reference model/deberta.py:717-947, 100-138, 171-194; torch autograd in the reference, main.py:81).
// With  s[i,j] = scale * (Q_i.K_j + Q_i.PKrel[i-j] + K_j.PQrel[i-j]),  P = softmax(s),  Pd = dropout(P),
// O = Pd.V  and the saved masked scores s2 (log2 domain) / row log-sum-exp of fb_attn_fwd:
//
//   dP   = dO . V^T                                   tcgen05.mma 128x128x64
//   dS'  = P * (dropout_bwd(dP) - rowsum(dO*O))       (registers, thread = query row; dS = scale * dS')
//   dV_J += Pd^T . dO_I        dQ_I += dS' . K_J       dK_J += dS'^T . Q_I        (128x64x128 each)
//   dQ_I += Z1 . PKrev-band    dK_J += Z2 . PQ-band                               (128x64x256 each)
//   dPKrel[band] += Z1^T . Q_I ,  dPQrel[band] += Z2^T . K_J                       (fb_attn_bwd_rel)
//   (scale is applied once, when the accumulators leave TMEM)
//
// Z1 / Z2 are dS' regrouped by RELATIVE POSITION inside a (128-query a, 128-key b) tile pair:
//   Z1[a][k] = dS'[a][b],  k = b - a + 127   (row = query;  PKrev-band[k] = PKrel row of relative position 127 - k)
//   Z2[b][k] = dS'[a][b],  k = a - b + 127   (row = key;    PQ-band[k]    = PQrel row of relative position k - 127)
// so that the positional terms are plain MMAs against one TMA box of a per-relative-position table (the band of
// a tile pair is contiguous; PKrel is read through a row-reversed copy so that both Z's are "row r shifted right
// by 127 - r").
//
// How the shift is made (the expensive part of a naive implementation: a scattered 2-byte store with a swizzled
// address per element, plus re-zeroing 64 KB per tile pair):
//   * 127 - r = 8 (15 - G) + (7 - rho) with G = r >> 3, rho = r & 7.  Z lives in shared memory in the SWIZZLE_NONE
//     K-major core-matrix layout (8 rows x 16 bytes, K-adjacent core matrices LBO = 128 B apart).  Row group G is
//     STORED as a 17-unit window (its 128 values shifted by the fine part 7 - rho, zero elsewhere) at group stride
//     SP; the MMA descriptor is given SBO = SP + LBO, so the tensor core sees group G displaced by G more units:
//     the coarse part 8 (15 - G) of the shift costs nothing.  (Semantics of SWIZZLE_NONE strides and of a skewed
//     SBO, also for the MN-major view Z^T used by fb_attn_bwd_rel, are pinned on hardware by
//     scripts/gpu_umma_probe.py through fb_umma_probe.)
//   * the zero padding between windows is written once per CTA: every tile pair overwrites exactly the same data
//     positions of every row (for Z1 and for Z2), so nothing is re-zeroed and no barrier is needed.
//   * per element the store address is (one of 8 per-thread offsets) + immediate.
//
// fb_attn_bwd:      one CTA per (batch, head); all key blocks x query tiles; dQ (per query tile), dK and dV
//                   (per key block) accumulate in TMEM, so there are no atomics and dQ/dK/dV are written
//                   once, as bf16, straight into the operand buffer of the following dX GEMM.
//                   dS' is also written (bf16 [B,h,L,lp]) for fb_attn_bwd_rel.
// fb_attn_bwd_rel:  gradients of the positional projections.  They are sums over batch AND over all tile
//                   pairs of one diagonal, so a CTA owns (head, diagonal, slice of the batch), keeps the two
//                   [256 x 64] fp32 band accumulators in TMEM across its whole slice and flushes once.
#include <cuda_fp16.h>
#include <cstring>

#include "../../include/frozenbilm_b200.h"
#ifdef FB_DEBUG_SURFACE
#include "../../include/frozenbilm_b200_debug.h"
#endif
#include "fb_common.cuh"

namespace fb {

// ---- skewed SWIZZLE_NONE staging buffer for Z1 / Z2 (bytes) ----
// LBO: K-adjacent core matrices (8 k-elements each).  SP: storage stride of a row group; >= 31.25 LBO so that the
// 32-unit view of a group never sees a neighbour's window, and chosen so that the row groups of a warp hit different
// banks.  SBO = SP + LBO is the group stride the tensor core is told (skew of one unit per group).  PRE: zero units in
// front of group 0.  BYTES >= PRE + 15 SP + 32 LBO.
// fb_attn_bwd only builds "Z1-type" rows (thread = row): LBO = 128 makes those stores conflict-free.  fb_attn_bwd_rel
// also scatters Z2 columns (thread = column): with LBO = 128 the four 8-lane groups of a warp land in the same banks
// (4.5-way conflicts, measured 15 M conflict cycles per launch); LBO = 144 / SP = 4512 makes the Z2 scatter
// conflict-free at 1.5-way for Z1 (bank simulation of both store patterns: DESIGN.md section 3).
struct ZFused {
  static constexpr uint32_t LBO = 128, SP = 4000, SBO = SP + LBO, PRE = 15 * LBO, BYTES = 66048;
};
struct ZRel {
  static constexpr uint32_t LBO = 144, SP = 4512, SBO = SP + LBO, PRE = 15 * LBO, BYTES = 74496;
};
static_assert(ZFused::BYTES >= ZFused::PRE + 15 * ZFused::SP + 32 * ZFused::LBO, "Z buffer (fused)");
static_assert(ZRel::BYTES >= ZRel::PRE + 15 * ZRel::SP + 32 * ZRel::LBO, "Z buffer (rel)");
static_assert(ZFused::SP * 4 >= 125 * ZFused::LBO && ZRel::SP * 4 >= 125 * ZRel::LBO, "row-group windows overlap");
static constexpr uint32_t Z_LBO = ZFused::LBO, Z_SBO = ZFused::SBO, Z_BYTES = ZFused::BYTES;   // fb_attn_bwd

// Z1 rows are queries: this thread owns row a and W of its values, keys b = j_first + e (j_first % 8 == 0).
// Window position of key b: y = b + 7 - rho.
template <int W, class ZL = ZFused>
__device__ __forceinline__ void build_z1(uint8_t* z, int a, int j_first, const uint32_t (&dsb)[W / 2]) {
  const int G = a >> 3, rho = a & 7;
  uint8_t* rowbase = z + ZL::PRE + G * ZL::SP + rho * 16 + (j_first >> 3) * ZL::LBO;
  uint8_t* off[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = 7 - rho + r;
    off[r] = rowbase + (t >> 3) * ZL::LBO + (t & 7) * 2;
  }
#pragma unroll
  for (int e = 0; e < W; ++e) {
    const uint32_t w = dsb[e >> 1];
    *reinterpret_cast<uint16_t*>(off[e & 7] + (e >> 3) * ZL::LBO) = static_cast<uint16_t>((e & 1) ? (w >> 16) : w);
  }
}
// Z2 rows are keys: this thread (query a) contributes one value to each of the rows b = j_first + e, at window
// position y = a + 7 - (b & 7).
template <int W, class ZL = ZFused>
__device__ __forceinline__ void build_z2(uint8_t* z, int a, int j_first, const uint32_t (&dsb)[W / 2]) {
  uint8_t* base = z + ZL::PRE + (j_first >> 3) * ZL::SP;
  uint8_t* off[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int y = a + 7 - r;
    off[r] = base + r * 16 + (y >> 3) * ZL::LBO + (y & 7) * 2;
  }
#pragma unroll
  for (int e = 0; e < W; ++e) {
    const uint32_t w = dsb[e >> 1];
    *reinterpret_cast<uint16_t*>(off[e & 7] + (e >> 3) * ZL::SP) = static_cast<uint16_t>((e & 1) ? (w >> 16) : w);
  }
}

// signal "operands staged in shared memory" from the compute warps to the MMA thread
__device__ __forceinline__ void compute_arrive(uint64_t* bar, int lane) {
  fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ uint4 pack8_scaled(const uint32_t* v, float sc) {
  return make_uint4(pack_bf16(__uint_as_float(v[0]) * sc, __uint_as_float(v[1]) * sc),
                    pack_bf16(__uint_as_float(v[2]) * sc, __uint_as_float(v[3]) * sc),
                    pack_bf16(__uint_as_float(v[4]) * sc, __uint_as_float(v[5]) * sc),
                    pack_bf16(__uint_as_float(v[6]) * sc, __uint_as_float(v[7]) * sc));
}

// Compute warps: NG warpgroups; thread = tile row a (TMEM lane), warpgroup g = columns [CW g, CW g + CW) of the
// 128-wide tile; warp 4 NG = TMA + MMA control, warp 4 NG + 1 = TMEM alloc.
static constexpr int NG = 4, CW = 128 / NG, NCT = 128 * NG;   // column groups, columns per thread, compute threads
static_assert(NG == 4, "the epilogues use 16-column TMEM loads (64 / NG)");
static constexpr int AB_THREADS = NCT + 64;
// shared memory map (bytes); every swizzled operand tile is 1024-aligned (128B swizzle atoms)
static constexpr uint32_t B_OFF_Q = 0, B_OFF_DO = 16384, B_OFF_K = 32768, B_OFF_V = 49152,
                          B_OFF_PK = 65536,                              // PKrev band (32 KB)
                          B_OFF_S0 = 98304,                              // [Pd | dS'] (64 KB); first 32 KB double as
                                                                         // the PQ band once the content MMAs are done
                          B_OFF_Z = B_OFF_S0 + 65536,                    // Z1 / Z2
                          B_OFF_BARS = B_OFF_Z + Z_BYTES,
                          B_OFF_I16 = B_OFF_BARS + 256,                  // 16 x 16 bf16 identity (512 B, SWIZZLE_NONE)
                          AB_SMEM = B_OFF_I16 + 512 + 1024;
static_assert(AB_SMEM <= 232448, "shared memory budget");
// TMEM columns: dP tile, 4 dQ accumulators (L <= 512), dK, dV
static constexpr uint32_t TB_DP = 0, TB_DQ = 128, TB_DK = 384, TB_DV = 448;

struct AttnBwdFusedParams {
  int B, L, h, nb, lp, rel_off, rel_rows;
  const __half* s2; const float* lse;
  const __nv_bfloat16* ctx; const __nv_bfloat16* dctx; long long ldc;
  __nv_bfloat16* dq; long long lddq;
  __nv_bfloat16* dk; long long lddk;
  __nv_bfloat16* dv; long long lddv;
  __nv_bfloat16* ds_out;
  float scale; DropSpec drop;
  long long* dbg;   // optional [pairs][32] clock64 stamps of CTA (0,0): events 0-15 control thread, 16-31 compute warp 0
};
#ifdef FB_DEBUG_SURFACE   // phase stamps exist only in the test-only debug library (include/frozenbilm_b200_debug.h)
#define FB_STAMP(ev)                                                                         \
  do {                                                                                       \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[t * 32 + (ev)] = clock64(); \
  } while (0)
#else
#define FB_STAMP(ev) do { } while (0)
#endif

template <bool REL>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_pk, const __grid_constant__ CUtensorMap map_pq,
                const __grid_constant__ CUtensorMap map_ds, const __grid_constant__ CUtensorMap map_dq,
                const __grid_constant__ CUtensorMap map_dk, const __grid_constant__ CUtensorMap map_dv,
                const __grid_constant__ AttnBwdFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BARS);
  uint64_t* bar_q = bars + 0;    // TMA: Q_I
  uint64_t* bar_do = bars + 1;   // TMA: dO_I
  uint64_t* bar_k = bars + 2;    // TMA: K_J
  uint64_t* bar_v = bars + 3;    // TMA: V_J
  uint64_t* bar_pk = bars + 4;   // TMA: PKrev band
  uint64_t* bar_pq = bars + 5;   // TMA: PQ band
  uint64_t* bar_dp = bars + 6;   // MMA: dP complete
  uint64_t* bar_rdyA = bars + 7;   // compute -> MMA: Pd | dS' staged
  uint64_t* bar_rdyB = bars + 8;   // compute -> MMA: Z1 staged
  uint64_t* bar_rdyC = bars + 9;   // compute -> MMA: Z2 staged
  uint64_t* bar_mmaA = bars + 10;  // MMA: dV, dQ, dK (content terms) complete
  uint64_t* bar_mmaB = bars + 11;  // MMA: dQ += Z1.PKrev complete
  uint64_t* bar_mmaC = bars + 12;  // MMA: dK += Z2.PQ complete
  uint64_t* bar_ep = bars + 16;      // compute -> control: dK_J / dV_J (and finally the dQ tiles) staged in shared memory
  uint64_t* bar_epfree = bars + 17;  // control -> compute: their bulk stores have left shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint8_t* s0 = smem + B_OFF_S0;
  uint8_t* zb = smem + B_OFF_Z;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hh = blockIdx.x, b = blockIdx.y;
  const int nb = p.nb;            // number of 128-row tiles (queries and keys)
  const int npairs = nb * nb;

  if (warp == 4 * NG && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_do);
    if constexpr (REL) { tma_prefetch_desc(&map_pk); tma_prefetch_desc(&map_pq); }
    for (int i = 0; i < 7; ++i) mbar_init(&bars[i], 1);
    mbar_init(bar_rdyA, 4 * NG); mbar_init(bar_rdyB, 4 * NG); mbar_init(bar_rdyC, 4 * NG);
    mbar_init(bar_mmaA, 1); mbar_init(bar_mmaB, 1); mbar_init(bar_mmaC, 1);
    mbar_init(bar_ep, 4 * NG); mbar_init(bar_epfree, 1);
    fence_barrier_init();
  }
  if (warp == 4 * NG + 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_prologue();   // everything below reads what earlier kernels of the stream wrote
  // per-thread row statistics for the query rows I * 128 + a: D = rowsum(dO * O), lse
  float Dr[4] = {0.f, 0.f, 0.f, 0.f}, Lr[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
  if (warp < 4 * NG) {
    if constexpr (REL) {
      // zero the Z staging buffer once: the padding between the row-group windows is never written again
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      for (uint32_t o = threadIdx.x * 16; o < Z_BYTES; o += NCT * 16) *reinterpret_cast<uint4*>(zb + o) = z4;
      if (threadIdx.x < 256) {
        // I16[n][k] as a K-major SWIZZLE_NONE B operand: core matrices (n >> 3, k >> 3), LBO (K) = 128, SBO (N) = 256
        const int n = threadIdx.x >> 4, k = threadIdx.x & 15;
        *reinterpret_cast<uint16_t*>(smem + B_OFF_I16 + (n >> 3) * 256 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
            (n == k) ? static_cast<uint16_t>(0x3F80) : static_cast<uint16_t>(0);
      }
    }
    // Row statistics, computed ONCE per row: column group g takes query tile I = g (NG = 4 covers L <= 512), all 16
    // loads of the row in flight together, and hands the result to the other groups through the (still unused)
    // [Pd | dS'] staging area.  (Every group used to walk every tile with 8 dependent load round trips per tile:
    // 16 % of the kernel's warp samples sat in this prologue.)
    const int a = (warp & 3) * 32 + lane;
    {
      // 8 lanes cover the 128 bytes of one row (4 rows per load instruction: full lines, not 32 scattered 16-byte
      // pieces), partial dot products are combined with three shuffles
      const int I = warp >> 2;
      const int rsub = lane >> 3, piece = lane & 7;
      float2* xs = reinterpret_cast<float2*>(s0);
      uint4 ov[8], gv[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int r = I * 128 + (warp & 3) * 32 + 4 * m + rsub;
        ov[m] = make_uint4(0u, 0u, 0u, 0u);
        gv[m] = ov[m];
        if (I < nb && r < p.L) {
          const long long off = ((long long)b * p.L + r) * p.ldc + hh * 64 + 8 * piece;
          ov[m] = __ldg(reinterpret_cast<const uint4*>(p.ctx + off));
          gv[m] = __ldg(reinterpret_cast<const uint4*>(p.dctx + off));
        }
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const uint32_t ow[4] = {ov[m].x, ov[m].y, ov[m].z, ov[m].w}, gw[4] = {gv[m].x, gv[m].y, gv[m].z, gv[m].w};
        float d = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float2 of = unpack_bf16(ow[w]), gf = unpack_bf16(gw[w]);
          d = fmaf(of.x, gf.x, d);
          d = fmaf(of.y, gf.y, d);
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (piece == 0) xs[I * 128 + (warp & 3) * 32 + 4 * m + rsub].x = d;
      }
      const int r = I * 128 + a;
      xs[I * 128 + a].y = (I < nb && r < p.L) ? p.lse[((long long)b * p.h + hh) * p.L + r] : INFINITY;
    }
    named_bar_sync(1, NCT);
    {
      const float2* xs = reinterpret_cast<const float2*>(s0);
#pragma unroll
      for (int I = 0; I < 4; ++I) {
        const float2 v = xs[I * 128 + a];
        Dr[I] = v.x; Lr[I] = v.y;
      }
    }
    named_bar_sync(1, NCT);   // every thread has its copy: the staging area may be overwritten
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4 * NG) {
    if (elect_one_sync()) {
      // ======================= TMA + MMA control thread =======================
      if constexpr (REL) { if (p.ds_out != nullptr) tma_prefetch_desc(&map_ds); }
      tma_prefetch_desc(&map_dq); tma_prefetch_desc(&map_dk); tma_prefetch_desc(&map_dv);
      // Output tiles are staged by the compute warps as 128B-swizzled [128 x 64] bf16 tiles in the (then free)
      // Pd half of the staging area and written by bulk stores from here: the direct version (thread = row, 32
      // bytes per thread) kept the compute warps ~3000 cycles at every key-block boundary.
      auto store_dkdv = [&](int Jd) {
        mbar_wait(bar_ep, Jd & 1);
        tma_store_4d(&map_dk, s0, 0, hh, Jd * 128, b);
        tma_store_4d(&map_dv, s0 + 16384, 0, hh, Jd * 128, b);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(bar_epfree);
      };
      constexpr uint32_t idesc_dp = make_idesc_bf16(128, 128, 0, 0);   // dP = dO . V^T       (A, B K-major)
      constexpr uint32_t idesc_kn = make_idesc_bf16(128, 64, 0, 1);    // A K-major, B MN-major (dS.K, Z.band)
      constexpr uint32_t idesc_mn = make_idesc_bf16(128, 64, 1, 1);    // A, B MN-major         (Pd^T.dO, dS^T.Q)
      const uint32_t sQ = smem_u32(smem + B_OFF_Q), sDO = smem_u32(smem + B_OFF_DO), sK = smem_u32(smem + B_OFF_K),
                     sV = smem_u32(smem + B_OFF_V), sPK = smem_u32(smem + B_OFF_PK), sS0 = smem_u32(s0),
                     sZ = smem_u32(zb);
      auto load_qdo = [&](int I) {
        mbar_arrive_expect_tx(bar_q, 16384);
        tma_load_4d(smem + B_OFF_Q, &map_q, bar_q, 0, hh, I * 128, b);
        mbar_arrive_expect_tx(bar_do, 16384);
        tma_load_4d(smem + B_OFF_DO, &map_do, bar_do, 0, hh, I * 128, b);
      };
      auto load_k = [&](int J) {
        mbar_arrive_expect_tx(bar_k, 16384);
        tma_load_4d(smem + B_OFF_K, &map_k, bar_k, 0, hh, J * 128, b);
      };
      auto load_v = [&](int J) {
        mbar_arrive_expect_tx(bar_v, 16384);
        tma_load_4d(smem + B_OFF_V, &map_v, bar_v, 0, hh, J * 128, b);
      };
      // PQ band: table row of k = (i - j) + rel_off with i - j = (I - J) 128 + k - 127
      auto pq_band = [&](int I, int J) { return (I - J) * 128 - 127 + p.rel_off; };
      // PKrev band: reversed table row of k = rel_rows - 1 - ((I - J) 128 + 127 - k + rel_off)
      auto pk_band = [&](int I, int J) { return p.rel_rows - 128 - p.rel_off - (I - J) * 128; };
      load_k(0);
      load_v(0);
      load_qdo(0);
      if constexpr (REL) {
        mbar_arrive_expect_tx(bar_pk, 32768);
        tma_load_3d(smem + B_OFF_PK, &map_pk, bar_pk, 0, pk_band(0, 0), hh);
      }
      for (int t = 0; t < npairs; ++t) {
        const int J = t / nb, I = t - J * nb;
        const uint32_t ph = t & 1, phJ = J & 1;
        const bool has_next = (t + 1 < npairs);
        const int Jn = (t + 1) / nb, In = (t + 1) - Jn * nb;
        FB_STAMP(0);
        // ---- dP = dO_I . V_J^T
        mbar_wait(bar_do, ph);
        if (I == 0) mbar_wait(bar_v, phJ);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + TB_DP, make_smem_desc_sw128(sDO + k * 32, 16, 1024),
                    make_smem_desc_sw128(sV + k * 32, 16, 1024), idesc_dp, k > 0);
        umma_commit(bar_dp);
        FB_STAMP(1);
        if (I == 0 && J > 0) store_dkdv(J - 1);   // previous key block's dK / dV (staged during its last pair)
        if (has_next && Jn != J) {
          // last query tile of this key block: V_J has no reader left once dP is done, so the next block's V is
          // fetched now instead of at the block boundary, where its latency sat in front of the next dP
          mbar_wait(bar_dp, ph);
          load_v(Jn);
        }
        // ---- content terms: dV_J += Pd^T.dO_I ; dQ_I += dS'.K_J ; dK_J += dS'^T.Q_I
        mbar_wait(bar_q, ph);
        if (I == 0) mbar_wait(bar_k, phJ);
        mbar_wait(bar_rdyA, ph);
        tc_fence_after();
        FB_STAMP(2);
        if constexpr (REL) {
          // dS' for fb_attn_bwd_rel: the staged [128 x 64] halves ARE 128B-swizzled TMA tiles, so two bulk stores
          // write them (rows >= L clipped by the map) instead of 64 scattered 16-byte stores per warp, which cost
          // ~2000 LSU cycles per tile pair on the critical path
          if (p.ds_out != nullptr) {
            tma_store_4d(&map_ds, s0 + 32768, J * 128, I * 128, hh, b);
            tma_store_4d(&map_ds, s0 + 32768 + 16384, J * 128 + 64, I * 128, hh, b);
            tma_store_commit();
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)   // A = Pd^T: MN-major (keys contiguous), two 64-key chunks 16 KB apart
          umma_bf16(tmem + TB_DV, make_smem_desc_sw128(sS0 + k * 2048, 16384, 1024),
                    make_smem_desc_sw128(sDO + k * 2048, 16, 1024), idesc_mn, (I > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // A = dS': K-major, keys [64 (k/4), +64) in the (k/4)-th 16 KB half
          umma_bf16(tmem + TB_DQ + 64 * I, make_smem_desc_sw128(sS0 + 32768 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(sK + k * 2048, 16, 1024), idesc_kn, (J > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // A = dS'^T: MN-major
          umma_bf16(tmem + TB_DK, make_smem_desc_sw128(sS0 + 32768 + k * 2048, 16384, 1024),
                    make_smem_desc_sw128(sQ + k * 2048, 16, 1024), idesc_mn, (I > 0 || k > 0) ? 1u : 0u);
        umma_commit(bar_mmaA);
        FB_STAMP(3);
        if constexpr (REL) {
          // ---- dS'^T through the tensor core (exact: bf16 x 1.0): D[key][16 j + n] = sum_q dS'[16 j + q][key] I[q][n],
          //      into the dP columns (already consumed); the compute threads read it back with lane = key for Z2
          constexpr uint32_t idesc_t = make_idesc_bf16(128, 16, 1, 0);
          const uint64_t idb = make_smem_desc_none(smem_u32(smem + B_OFF_I16), 128, 256);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            umma_bf16(tmem + TB_DP + 16 * j, make_smem_desc_sw128(sS0 + 32768 + j * 2048, 16384, 1024), idb, idesc_t, 0u);
          // ---- c2p: dQ_I += Z1 . PKrev-band   (K = 256: 16 k-steps of two core matrices)
          mbar_wait(bar_pk, ph);
          mbar_wait(bar_rdyB, ph);
          tc_fence_after();
          FB_STAMP(4);
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_bf16(tmem + TB_DQ + 64 * I, make_smem_desc_none(sZ + k * 2 * Z_LBO, Z_LBO, Z_SBO),
                      make_smem_desc_sw128(sPK + k * 2048, 16, 1024), idesc_kn, 1u);
          tma_store_wait_read0();   // the dS' bulk stores have left shared memory before bar_mmaB releases that half
          umma_commit(bar_mmaB);
          FB_STAMP(5);
        }
        mbar_wait(bar_mmaA, ph);      // Q_I, dO_I, [Pd | dS'] (and K_J, V_J after the last query tile) are free
        FB_STAMP(6);
        if constexpr (REL) {
          mbar_arrive_expect_tx(bar_pq, 32768);
          tma_load_3d(s0, &map_pq, bar_pq, 0, pq_band(I, J), hh);
        }
        if (has_next) {
          load_qdo(In);
          if (Jn != J) load_k(Jn);
        }
        if constexpr (REL) {
          mbar_wait(bar_mmaB, ph);    // the PKrev band and Z1 are free
          FB_STAMP(7);
          if (has_next) {
            mbar_arrive_expect_tx(bar_pk, 32768);
            tma_load_3d(smem + B_OFF_PK, &map_pk, bar_pk, 0, pk_band(In, Jn), hh);
          }
          // ---- p2c: dK_J += Z2 . PQ-band
          mbar_wait(bar_pq, ph);
          FB_STAMP(8);
          mbar_wait(bar_rdyC, ph);
          tc_fence_after();
          FB_STAMP(9);
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_bf16(tmem + TB_DK, make_smem_desc_none(sZ + k * 2 * Z_LBO, Z_LBO, Z_SBO),
                      make_smem_desc_sw128(sS0 + k * 2048, 16, 1024), idesc_kn, 1u);
          umma_commit(bar_mmaC);
          FB_STAMP(10);
          mbar_wait(bar_mmaC, ph);
          FB_STAMP(11);
        }
      }
      store_dkdv(nb - 1);
      mbar_wait(bar_ep, nb & 1);          // the dQ tiles (one per query tile) are staged
      for (int I = 0; I < nb; ++I) tma_store_4d(&map_dq, s0 + I * 16384, 0, hh, I * 128, b);
      tma_store_commit();
      tma_store_wait_all();   // every bulk store is globally complete before the CTA retires
    }
    __syncwarp();
  } else if (warp < 4 * NG) {
    // ============ compute warpgroups: thread = tile row a, warpgroup g = columns [CW g, CW g + CW) ============
    const int g = warp >> 2, wq = warp & 3;
    const int a = wq * 32 + lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    uint64_t* bar_last = REL ? bar_mmaC : bar_mmaA;   // last MMA group of a pair
    constexpr int OC = 64 / NG;                       // output columns per thread in the dQ / dK / dV epilogues
    const bool do_drop = p.drop.thresh16 != 0;

    // This thread's CW saved scores of a pair (fp16 pairs, log2 domain; -inf where masked or padded).  HBM latency
    // under this access pattern is several microseconds, so the segment of pair t+1 is hinted into L2 at the start of
    // pair t and loaded into registers while the last MMA group of pair t runs.
    uint32_t sc[CW / 2];
    constexpr uint32_t NINF2 = 0xFC00FC00u;   // half2(-inf, -inf)
    auto score_ptr = [&](int tt) -> const __half* {
      const int Jt = tt / nb, It = tt - Jt * nb;
      const int it = It * 128 + a;
      if (tt >= npairs || it >= p.L) return nullptr;
      return p.s2 + (((long long)b * p.h + hh) * p.L + it) * p.lp + Jt * 128 + CW * g;
    };
    auto load_scores = [&](int tt) {
      if constexpr (REL) {
        // Coalesced: a direct load (thread = row, 64 contiguous bytes each) touches 32 different lines per
        // instruction.  Instead 4 lanes fetch the 64 bytes of one row (8 rows per instruction), and the warp exchanges
        // the pieces through the 64 bytes per row that its own lanes own in the dS' half of the staging area (free
        // once bar_mmaB has completed: the content / transpose MMAs and the dS' bulk store no longer read it).
        const int Jt = tt / nb, It = tt - Jt * nb;
        if (tt >= npairs) {
#pragma unroll
          for (int u = 0; u < CW / 2; ++u) sc[u] = NINF2;
          return;
        }
        const int c0 = CW * g;
        uint8_t* area = s0 + 32768 + (c0 >> 6) * 16384;       // this column group's 64-key chunk of dS'
        const int ub = (c0 & 63) >> 3;                          // first of its four 16-byte units in a row
        const int rsub = lane >> 2, piece = lane & 3;
        const __half* gsrc = p.s2 + (((long long)b * p.h + hh) * p.L + It * 128 + 32 * wq) * p.lp + Jt * 128 + c0 + 8 * piece;
        const uint4 ninf = make_uint4(NINF2, NINF2, NINF2, NINF2);
        uint4 tmp[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int r = 8 * m + rsub;                           // row inside this warp's 32
          tmp[m] = (It * 128 + 32 * wq + r < p.L) ? __ldg(reinterpret_cast<const uint4*>(gsrc + (long long)r * p.lp)) : ninf;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int ra = 32 * wq + 8 * m + rsub;
          *reinterpret_cast<uint4*>(area + ra * 128 + (((ub + piece) ^ (ra & 7)) << 4)) = tmp[m];
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint4 q = *reinterpret_cast<const uint4*>(area + a * 128 + (((ub + u) ^ (a & 7)) << 4));
          sc[4 * u] = q.x; sc[4 * u + 1] = q.y; sc[4 * u + 2] = q.z; sc[4 * u + 3] = q.w;
        }
        __syncwarp();
      } else {
        const uint4* src = reinterpret_cast<const uint4*>(score_ptr(tt));
        if (src != nullptr) {
#pragma unroll
          for (int u = 0; u < CW / 8; ++u) {
            const uint4 q = __ldg(src + u);
            sc[4 * u] = q.x; sc[4 * u + 1] = q.y; sc[4 * u + 2] = q.z; sc[4 * u + 3] = q.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < CW / 2; ++u) sc[u] = NINF2;
        }
      }
    };
    load_scores(0);

    for (int t = 0; t < npairs; ++t) {
      const int J = t / nb, I = t - J * nb;
      const uint32_t ph = t & 1;
      const int i = I * 128 + a, j0 = J * 128;
      const long long srow = ((long long)b * p.h + hh) * p.L + i;
      const bool stamp = (threadIdx.x == 0);
      if (stamp) FB_STAMP(16);
      {
        const __half* nxt = score_ptr(t + 1);
        if (nxt != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
      }
      float lse = Lr[0], D = Dr[0];
#pragma unroll
      for (int ii = 1; ii < 4; ++ii)
        if (ii == I) { lse = Lr[ii]; D = Dr[ii]; }
      const uint32_t rowkey = drop_row_key(p.drop, static_cast<uint32_t>(srow));
      mbar_wait(bar_dp, ph);
      tc_fence_after();
      if (stamp) FB_STAMP(17);
      uint32_t dsb[CW / 2], pdb[CW / 2];   // dS' and Pd of this thread's CW columns, bf16 pairs
      {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + TB_DP + CW * g, v);
        tmem_ld_wait();
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          float fk[8];
          if (do_drop) {
            drop_factors8(rowkey, static_cast<uint32_t>(j0 + CW * g + 8 * gg) >> 3, p.drop.thresh16, p.drop.inv_keep, fk);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) fk[e] = 1.0f;
          }
          float dsv[8], pdv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int idx = 8 * gg + e;           // column inside this thread's CW
            const float2 s2f = __half22float2(*reinterpret_cast<const __half2*>(&sc[idx >> 1]));
            const float s = (idx & 1) ? s2f.y : s2f.x;
            const float pr = ex2_approx(s - lse);   // s = -inf (masked / padded row) or lse = +inf (dead row) -> 0
            pdv[e] = pr * fk[e];
            dsv[e] = pr * fmaf(__uint_as_float(v[idx]), fk[e], -D);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            dsb[4 * gg + e] = pack_bf16(dsv[2 * e], dsv[2 * e + 1]);
            pdb[4 * gg + e] = pack_bf16(pdv[2 * e], pdv[2 * e + 1]);
          }
        }
      }
      if (stamp) FB_STAMP(18);
      // [Pd | dS'] and Z are free once the previous pair's last MMA group has completed
      if (t > 0) mbar_wait(bar_last, (t - 1) & 1);
      if (I == 0 && J > 0) mbar_wait(bar_epfree, (J - 1) & 1);   // the dK / dV tiles staged there have been stored
      if (stamp) FB_STAMP(19);
      {
        // stage Pd and dS' rows: K-major [row a][64 keys] chunks of 16 KB, 128B swizzle; this thread's 32 columns
        // are the four 16-byte units (col0 >> 3) .. +3 of chunk (col0 >> 6)
        const int col0 = CW * g;
        uint8_t* prow = s0 + (col0 >> 6) * 16384 + a * 128;
        uint8_t* drow = prow + 32768;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
          const int unit = ((((col0 & 63) >> 3) + gg) ^ (a & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + unit) = make_uint4(pdb[4 * gg], pdb[4 * gg + 1], pdb[4 * gg + 2], pdb[4 * gg + 3]);
          *reinterpret_cast<uint4*>(drow + unit) = make_uint4(dsb[4 * gg], dsb[4 * gg + 1], dsb[4 * gg + 2], dsb[4 * gg + 3]);
        }
      }
      if constexpr (REL) {
        // Z1 is built BEFORE the content MMAs are released: shared-memory stores that run while the tensor core
        // streams operands are several times slower (measured), and the MMAs then issue back to back
        build_z1<CW>(zb, a, CW * g, dsb);
        if (stamp) FB_STAMP(20);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar_rdyA); mbar_arrive(bar_rdyB); }
        if (stamp) FB_STAMP(21);
      } else {
        compute_arrive(bar_rdyA, lane);
      }
      if constexpr (REL) {
        if (stamp) FB_STAMP(22);
        mbar_wait(bar_mmaB, ph);      // the tensor core no longer reads Z1; dS'^T is in the dP columns
        tc_fence_after();
        if (stamp) FB_STAMP(23);
        {
          // this thread now owns KEY row a: Z2[a][k] = dS'^T[a][q], k = q - a + 127: the Z1 construction on the transpose
          uint32_t v[32];
          tmem_ld_32x32(lane_base + TB_DP + CW * g, v);
          tmem_ld_wait();
          uint32_t dst[CW / 2];
#pragma unroll
          for (int e = 0; e < CW / 2; ++e) dst[e] = pack_bf16(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          build_z1<CW>(zb, a, CW * g, dst);
        }
        tc_fence_before();
        compute_arrive(bar_rdyC, lane);
        if (stamp) FB_STAMP(24);
      }
      if (I == nb - 1) {
        // ---- key block J is complete: dK_J (x scale), dV_J -> bf16 (thread = key row, warpgroup = OC-column slice)
        mbar_wait(bar_last, ph);
        tc_fence_after();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          uint32_t v[16];
          tmem_ld_32x16(lane_base + (w == 0 ? TB_DK : TB_DV) + OC * g, v);
          tmem_ld_wait();
          const float sc_out = (w == 0) ? p.scale : 1.0f;
          uint8_t* rowp = s0 + w * 16384 + a * 128;          // [128 keys x 64] bf16 tile, 128B swizzle
#pragma unroll
          for (int u = 0; u < OC / 8; ++u)
            *reinterpret_cast<uint4*>(rowp + (((2 * g + u) ^ (a & 7)) << 4)) = pack8_scaled(v + 8 * u, sc_out);
        }
        tc_fence_before();
        compute_arrive(bar_ep, lane);
        tc_fence_before();   // the next key block's first MMA overwrites dK / dV (ordered by bar_rdyA)
      }
      if (stamp) FB_STAMP(25);
      if (t + 1 < npairs) {
        // next pair's scores into registers while the last MMA group runs, then park on its barrier (busy compute
        // warps measurably slow the control thread's MMA issue)
        load_scores(t + 1);
        mbar_wait(bar_last, ph);
      }
      if (stamp) FB_STAMP(26);
    }
    // ---- dQ of every query tile (x scale)
    mbar_wait(bar_last, (npairs - 1) & 1);
    tc_fence_after();
    mbar_wait(bar_epfree, (nb - 1) & 1);   // the last key block's dK / dV have left the staging area
    for (int I = 0; I < nb; ++I) {
      uint32_t v[16];
      tmem_ld_32x16(lane_base + TB_DQ + 64 * I + OC * g, v);
      tmem_ld_wait();
      uint8_t* rowp = s0 + I * 16384 + a * 128;
#pragma unroll
      for (int u = 0; u < OC / 8; ++u)
        *reinterpret_cast<uint4*>(rowp + (((2 * g + u) ^ (a & 7)) << 4)) = pack8_scaled(v + 8 * u, p.scale);
    }
    tc_fence_before();
    compute_arrive(bar_ep, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4 * NG + 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Gradients of the per-relative-position tables (scale applied at the flush):
//   gpk[h, rel_off + d 128 + 127 - k, :] += sum_{b, pairs (I, J = I - d)} Z1^T . Q_I     (c2p, deberta.py:870-881)
//   gpq[h, rel_off + d 128 - 127 + k, :] += sum_{b, pairs}               Z2^T . K_J     (p2c, deberta.py:884-918)
// One CTA = (head, diagonal d, contiguous slice of the (pair, batch) units of that diagonal).  Z1 / Z2 are built
// exactly as in fb_attn_bwd and read by the tensor core through the MN-major (transposed) view of the same buffer.
// ---------------------------------------------------------------------------------------------
static constexpr int AR_THREADS = 320;   // warps 0-7 compute (thread = row, warpgroup = 64-column half), 8 control, 9 TMEM
static constexpr uint32_t R_OFF_Q = 0, R_OFF_K = 16384, R_OFF_Z = 32768, R_OFF_BARS = R_OFF_Z + ZRel::BYTES,
                          AR_SMEM = R_OFF_BARS + 128 + 1024;
static constexpr uint32_t TR_PK = 0, TR_PQ = 128;   // two [256 x 64] accumulators, each two 128-lane halves of 64 columns

struct AttnBwdRelParams {
  int B, L, h, nb, lp, rel_rows, rel_off, units_per_cta;
  const __nv_bfloat16* ds;
  float* gpk; float* gpq;
  float scale;
};

__global__ void __launch_bounds__(AR_THREADS, 2)
attn_bwd_rel_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ AttnBwdRelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R_OFF_BARS);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_rdy1 = bars + 2;   // Z1 staged
  uint64_t* bar_rdy2 = bars + 3;   // Z2 staged
  uint64_t* bar_mma1 = bars + 4;
  uint64_t* bar_mma2 = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  uint8_t* sz = smem + R_OFF_Z;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.nb;
  const int hh = blockIdx.y;
  // blockIdx.x -> (diagonal, slice): diagonals are laid out one after the other, each with ceil(units / units_per_cta) slices
  int d = 0, slice = 0, nunits = 0;
  {
    int x = blockIdx.x;
    for (int dd = -(nb - 1); dd <= nb - 1; ++dd) {
      const int nu = (nb - (dd < 0 ? -dd : dd)) * p.B;
      const int ns = (nu + p.units_per_cta - 1) / p.units_per_cta;
      if (x < ns) { d = dd; slice = x; nunits = nu; break; }
      x -= ns;
    }
  }
  const int u_begin = slice * p.units_per_cta;
  const int u_end = min(u_begin + p.units_per_cta, nunits);
  const int I_first = d > 0 ? d : 0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1);
    mbar_init(bar_rdy1, 8); mbar_init(bar_rdy2, 8);
    mbar_init(bar_mma1, 1); mbar_init(bar_mma2, 1);
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (warp < 8) {
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t o = threadIdx.x * 16; o < ZRel::BYTES; o += 256 * 16) *reinterpret_cast<uint4*>(sz + o) = z4;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_prologue();

  if (warp == 8) {
    if (elect_one_sync()) {
      constexpr uint32_t idesc_mn = make_idesc_bf16(128, 64, 1, 1);
      const uint32_t sQ = smem_u32(smem + R_OFF_Q), sK = smem_u32(smem + R_OFF_K), sZ = smem_u32(sz);
      auto load_q = [&](int u) {
        const int pair = u / p.B, bb = u - pair * p.B;
        mbar_arrive_expect_tx(bar_q, 16384);
        tma_load_4d(smem + R_OFF_Q, &map_q, bar_q, 0, hh, (I_first + pair) * 128, bb);
      };
      auto load_k = [&](int u) {
        const int pair = u / p.B, bb = u - pair * p.B;
        mbar_arrive_expect_tx(bar_k, 16384);
        tma_load_4d(smem + R_OFF_K, &map_k, bar_k, 0, hh, (I_first + pair - d) * 128, bb);
      };
      if (u_begin < u_end) { load_q(u_begin); load_k(u_begin); }
      for (int u = u_begin; u < u_end; ++u) {
        const uint32_t ph = (u - u_begin) & 1;
        const uint32_t acc = (u > u_begin) ? 1u : 0u;
        // ---- band of gpk += Z1^T . Q_I.  A = Z^T: MN-major SWIZZLE_NONE view of the skewed buffer: M = band index k
        //      (8-element groups LBO apart -> SBO field), K = the 128 rows (8-row groups SBO apart -> LBO field);
        //      two M halves of 128 (16 groups each), 8 k-steps of 16 rows
        mbar_wait(bar_q, ph);
        mbar_wait(bar_rdy1, ph);
        tc_fence_after();
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + TR_PK + 64 * mh,
                      make_smem_desc_none(sZ + mh * 16 * ZRel::LBO + k * 2 * ZRel::SBO, ZRel::SBO, ZRel::LBO),
                      make_smem_desc_sw128(sQ + k * 2048, 16, 1024), idesc_mn, (acc || k > 0) ? 1u : 0u);
        umma_commit(bar_mma1);
        mbar_wait(bar_mma1, ph);
        if (u + 1 < u_end) load_q(u + 1);   // the Q tile is free: the next one lands while Z2 is built and multiplied
        // ---- band of gpq += Z2^T . K_J
        mbar_wait(bar_k, ph);
        mbar_wait(bar_rdy2, ph);
        tc_fence_after();
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem + TR_PQ + 64 * mh,
                      make_smem_desc_none(sZ + mh * 16 * ZRel::LBO + k * 2 * ZRel::SBO, ZRel::SBO, ZRel::LBO),
                      make_smem_desc_sw128(sK + k * 2048, 16, 1024), idesc_mn, (acc || k > 0) ? 1u : 0u);
        umma_commit(bar_mma2);
        mbar_wait(bar_mma2, ph);
        if (u + 1 < u_end) load_k(u + 1);
      }
    }
    __syncwarp();
  } else if (warp < 8) {
    const int g = warp >> 2, wq = warp & 3;
    const int a = wq * 32 + lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq * 32) << 16);
    for (int u = u_begin; u < u_end; ++u) {
      const uint32_t ph = (u - u_begin) & 1;
      const int pair = u / p.B, bb = u - pair * p.B;
      const int I = I_first + pair, J = I - d;
      const int i = I * 128 + a;
      if (u + 1 < u_end) {
        // hint the next unit's dS' row segment (one 128-byte line per thread) into L2: its load latency is otherwise
        // the head of every unit's dependency chain
        const int pn = (u + 1) / p.B, bn = (u + 1) - pn * p.B;
        const int in = (I_first + pn) * 128 + a;
        if (in < p.L)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.ds + (((long long)bn * p.h + hh) * p.L + in) * p.lp +
                                                         (I_first + pn - d) * 128 + 64 * g));
      }
      uint32_t dsb[32];
      if (i < p.L) {
        const uint4* src = reinterpret_cast<const uint4*>(p.ds + (((long long)bb * p.h + hh) * p.L + i) * p.lp + J * 128 + 64 * g);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 w = __ldg(src + q);
          dsb[4 * q] = w.x; dsb[4 * q + 1] = w.y; dsb[4 * q + 2] = w.z; dsb[4 * q + 3] = w.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) dsb[q] = 0u;
      }
      if (u > u_begin) mbar_wait(bar_mma2, ph ^ 1);   // previous unit's MMAs no longer read the staging buffer
      build_z1<64, ZRel>(sz, a, 64 * g, dsb);
      compute_arrive(bar_rdy1, lane);
      mbar_wait(bar_mma1, ph);
      build_z2<64, ZRel>(sz, a, 64 * g, dsb);
      compute_arrive(bar_rdy2, lane);
    }
    // ---- flush the band accumulators (thread = band index k = 128 mh + a, warpgroup = 32-column half)
    if (u_begin < u_end) {
      mbar_wait(bar_mma2, (u_end - 1 - u_begin) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int w = 0; w < 4; ++w) {
        const int mh = w & 1;
        const int k = 128 * mh + a;
        // relative position of band index k: Z1 (reversed table) 127 - k, Z2 k - 127
        const int row = p.rel_off + d * 128 + (w < 2 ? 127 - k : k - 127);
        uint32_t v[32];
        tmem_ld_32x32(lane_base + (w < 2 ? TR_PK : TR_PQ) + 64 * mh + 32 * g, v);
        tmem_ld_wait();
        if (k < 255 && row >= 0 && row < p.rel_rows) {
          float4* dst = reinterpret_cast<float4*>((w < 2 ? p.gpk : p.gpq) + ((long long)hh * p.rel_rows + row) * 64 + 32 * g);
#pragma unroll
          for (int e = 0; e < 8; ++e)   // 16-byte vector reductions (red.global.add.v4.f32): 4x fewer L2 atomics
            atomicAdd(dst + e, make_float4(__uint_as_float(v[4 * e]) * p.scale, __uint_as_float(v[4 * e + 1]) * p.scale,
                                           __uint_as_float(v[4 * e + 2]) * p.scale, __uint_as_float(v[4 * e + 3]) * p.scale));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

static int make_rows_map(CUtensorMap* m, const void* ptr, long long ld, int B, int L, int h) {
  const uint64_t dims[4] = {64, (uint64_t)h, (uint64_t)L, (uint64_t)B};
  const uint64_t strides[3] = {64, (uint64_t)ld, (uint64_t)ld * L};
  const uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16_nd(m, ptr, 4, dims, strides, box);
}
static int make_band_map(CUtensorMap* m, const void* ptr, int rel_rows, int h) {
  const uint64_t dims[3] = {64, (uint64_t)rel_rows, (uint64_t)h};
  const uint64_t strides[2] = {64, (uint64_t)rel_rows * 64};
  const uint32_t box[3] = {64, 256, 1};
  return make_tmap_bf16_nd(m, ptr, 3, dims, strides, box);
}

static long long* g_attn_bwd_dbg = nullptr;   // set by fb_attn_bwd_set_debug (profiling aid, normally null)

int attn_bwd(const fb_attn_bwd_args* a, cudaStream_t st) {
  FB_REQUIRE(a && a->Q && a->K && a->V && a->dO && a->O && a->s2 && a->lse && a->dQ && a->dK && a->dV,
             "fb_attn_bwd: null arg");
  FB_REQUIRE(a->B > 0 && a->h > 0 && a->L > 0 && a->L <= 512, "fb_attn_bwd: need 0 < L <= 512 (got %d)", a->L);
  const int lpad = (a->L + 127) / 128 * 128;
  FB_REQUIRE(a->lp == lpad, "fb_attn_bwd: s2 row pitch must be L rounded up to 128 (got %d)", a->lp);
  const bool rel = a->PKrev != nullptr;
  FB_REQUIRE(!rel || a->PQrel, "fb_attn_bwd: PKrev and PQrel must be given together");
  FB_REQUIRE(!rel || (a->rel_off >= lpad && a->rel_rows >= a->rel_off + lpad + 1),
             "fb_attn_bwd: relative tables too small (rel_rows=%d rel_off=%d L=%d)", a->rel_rows, a->rel_off, a->L);
  FB_REQUIRE(a->lddq % 8 == 0 && a->lddk % 8 == 0 && a->lddv % 8 == 0 && a->ldo % 8 == 0,
             "fb_attn_bwd: row pitches must be multiples of 8 elements");
  CUtensorMap mq, mk, mv, mdo, mpk, mpq, mds;
  int rc;
  if ((rc = make_rows_map(&mq, a->Q, a->ldq, a->B, a->L, a->h))) return rc;
  if ((rc = make_rows_map(&mk, a->K, a->ldk, a->B, a->L, a->h))) return rc;
  if ((rc = make_rows_map(&mv, a->V, a->ldv, a->B, a->L, a->h))) return rc;
  if ((rc = make_rows_map(&mdo, a->dO, a->ldo, a->B, a->L, a->h))) return rc;
  if (rel) {
    if ((rc = make_band_map(&mpk, a->PKrev, a->rel_rows, a->h))) return rc;
    if ((rc = make_band_map(&mpq, a->PQrel, a->rel_rows, a->h))) return rc;
  } else {
    mpk = mk; mpq = mk;   // never dereferenced by the plain-attention instantiation
  }
  CUtensorMap mdq, mdk, mdv;
  if ((rc = make_rows_map(&mdq, a->dQ, a->lddq, a->B, a->L, a->h))) return rc;
  if ((rc = make_rows_map(&mdk, a->dK, a->lddk, a->B, a->L, a->h))) return rc;
  if ((rc = make_rows_map(&mdv, a->dV, a->lddv, a->B, a->L, a->h))) return rc;
  mds = mk;
  if (rel && a->ds_out) {
    // dS' [B, h, L, lp] as {key, query, head, batch} with a [64 keys x 128 queries] box (store side)
    const uint64_t dims[4] = {(uint64_t)lpad, (uint64_t)a->L, (uint64_t)a->h, (uint64_t)a->B};
    const uint64_t strides[3] = {(uint64_t)lpad, (uint64_t)lpad * a->L, (uint64_t)lpad * a->L * a->h};
    const uint32_t box[4] = {64, 128, 1, 1};
    if ((rc = make_tmap_bf16_nd(&mds, a->ds_out, 4, dims, strides, box))) return rc;
  }
  AttnBwdFusedParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.h = a->h; p.nb = lpad / 128; p.lp = lpad; p.rel_off = a->rel_off; p.rel_rows = a->rel_rows;
  p.s2 = (const __half*)a->s2; p.lse = a->lse;
  p.ctx = (const __nv_bfloat16*)a->O; p.dctx = (const __nv_bfloat16*)a->dO; p.ldc = a->ldo;
  p.dq = (__nv_bfloat16*)a->dQ; p.lddq = a->lddq;
  p.dk = (__nv_bfloat16*)a->dK; p.lddk = a->lddk;
  p.dv = (__nv_bfloat16*)a->dV; p.lddv = a->lddv;
  p.ds_out = (__nv_bfloat16*)a->ds_out;
  p.scale = a->scale;
  p.drop = make_drop(a->dropout_p, a->seed, a->offset);
  p.dbg = g_attn_bwd_dbg;
  static bool configured = false;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    configured = true;
  }
  dim3 grid(a->h, a->B);
  if (rel)
    FB_CHECK_CUDA(launch_pdl(attn_bwd_kernel<true>, grid, dim3(AB_THREADS), AB_SMEM, st, mq, mk, mv, mdo, mpk, mpq, mds, mdq,
                             mdk, mdv, p));
  else
    FB_CHECK_CUDA(launch_pdl(attn_bwd_kernel<false>, grid, dim3(AB_THREADS), AB_SMEM, st, mq, mk, mv, mdo, mpk, mpq, mds, mdq,
                             mdk, mdv, p));
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attn_bwd_rel(const fb_attn_bwd_rel_args* a, cudaStream_t st) {
  FB_REQUIRE(a && a->ds && a->Q && a->K && a->gpk && a->gpq, "fb_attn_bwd_rel: null arg");
  FB_REQUIRE(a->B > 0 && a->h > 0 && a->L > 0 && a->L <= 512, "fb_attn_bwd_rel: need 0 < L <= 512 (got %d)", a->L);
  const int lpad = (a->L + 127) / 128 * 128;
  FB_REQUIRE(a->lp == lpad, "fb_attn_bwd_rel: ds row pitch must be L rounded up to 128 (got %d)", a->lp);
  FB_REQUIRE(a->rel_off >= lpad && a->rel_rows >= a->rel_off + lpad + 1, "fb_attn_bwd_rel: relative tables too small");
  CUtensorMap mq, mk;
  int rc;
  if ((rc = make_rows_map(&mq, a->Q, a->ldq, a->B, a->L, a->h))) return rc;
  if ((rc = make_rows_map(&mk, a->K, a->ldk, a->B, a->L, a->h))) return rc;
  AttnBwdRelParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.h = a->h; p.nb = lpad / 128; p.lp = lpad; p.rel_rows = a->rel_rows; p.rel_off = a->rel_off;
  p.ds = (const __nv_bfloat16*)a->ds; p.gpk = a->gpk; p.gpq = a->gpq;
  p.scale = a->scale;
  // Slice the (pair, batch) units of every diagonal.  Two CTAs are resident per SM; slices of at most 8 units keep
  // the tail short (the diagonals hold different numbers of units, so "one slice per CTA slot" leaves a second,
  // nearly empty wave) while one flush of the band accumulators (128 KB of fp32 atomics) per 8 units stays cheap.
  const int nb = p.nb;
  const int total_units = nb * nb * a->B;
  int upc = (total_units * a->h + 2 * num_sms() - 1) / (2 * num_sms());
  if (upc > 8) upc = 8;
  if (upc < 1) upc = 1;
  if (a->units_per_cta > 0) upc = a->units_per_cta;
  p.units_per_cta = upc;
  int nslices = 0;
  for (int dd = -(nb - 1); dd <= nb - 1; ++dd) {
    const int nu = (nb - (dd < 0 ? -dd : dd)) * a->B;
    nslices += (nu + upc - 1) / upc;
  }
  FB_CHECK_CUDA(cudaMemsetAsync(a->gpk, 0, sizeof(float) * (size_t)a->h * a->rel_rows * 64, st));
  FB_CHECK_CUDA(cudaMemsetAsync(a->gpq, 0, sizeof(float) * (size_t)a->h * a->rel_rows * 64, st));
  static bool configured = false;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_rel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AR_SMEM));
    configured = true;
  }
  // (plain launch: it follows the two memset nodes above, which are not kernels a dependent launch could overlap)
  attn_bwd_rel_kernel<<<dim3(nslices, a->h), AR_THREADS, AR_SMEM, st>>>(mq, mk, p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb

extern "C" {
int fb_attn_bwd(const fb_attn_bwd_args* args, void* stream) {
  return fb::attn_bwd(args, reinterpret_cast<cudaStream_t>(stream));
}
#ifdef FB_DEBUG_SURFACE
int fb_attn_bwd_set_debug(long long* stamps) {
  fb::g_attn_bwd_dbg = stamps;
  return 0;
}
#endif
int fb_attn_bwd_rel(const fb_attn_bwd_rel_args* args, void* stream) {
  return fb::attn_bwd_rel(args, reinterpret_cast<cudaStream_t>(stream));
}
}
