// fb_umma_probe: a one-CTA tcgen05.mma test bench used by tests/ and scripts/ to pin down shared-memory operand
// layouts (descriptor modes, LBO/SBO semantics) on real hardware before a kernel relies on them.  The host supplies
// raw byte images of the A and B operands, one (a_off, b_off) descriptor start per MMA instruction, the two
// 64-bit shared-memory descriptor templates (everything except the start address) and the 32-bit instruction
// descriptor; the kernel copies the images into shared memory, issues the MMAs into one TMEM accumulator and
// returns it as fp32 [128, N] (row = TMEM lane).  Not on the training path.
#include <cstring>

#include "../../include/frozenbilm_b200_debug.h"
#include "fb_common.cuh"

namespace fb {

static constexpr uint32_t PR_A_BYTES = 131072, PR_B_BYTES = 65536, PR_SMEM = PR_A_BYTES + PR_B_BYTES + 256 + 1024;

struct ProbeParams {
  const uint8_t* a_image; int a_bytes;
  const uint8_t* b_image; int b_bytes;
  int n_mma; int a_off[64]; int b_off[64];
  unsigned long long a_desc, b_desc;
  uint32_t idesc;
  int N;
  float* out;
  long long* cycles;   // optional [2]: clock64 from the first MMA issue to (all issued, all complete)
};

__global__ void __launch_bounds__(160, 1) umma_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + PR_A_BYTES + PR_B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(p.a_image + i);
  for (int i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + PR_A_BYTES + i) = *reinterpret_cast<const uint4*>(p.b_image + i);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 4) {
    if (elect_one_sync()) {
      const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + PR_A_BYTES);
      const long long c0 = clock64();
      for (int i = 0; i < p.n_mma; ++i) {
        const uint64_t ad = p.a_desc | static_cast<uint64_t>(((sa + p.a_off[i]) & 0x3FFFF) >> 4);
        const uint64_t bd = p.b_desc | static_cast<uint64_t>(((sb + p.b_off[i]) & 0x3FFFF) >> 4);
        umma_bf16(tmem, ad, bd, p.idesc, i > 0 ? 1u : 0u);
      }
      umma_commit(bar);
      const long long c1 = clock64();
      mbar_wait(bar, 0);
      const long long c2 = clock64();
      if (p.cycles) { p.cycles[0] = c1 - c0; p.cycles[1] = c2 - c0; }
    }
    __syncwarp();
  } else {
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int c = 0; c < p.N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(lane_base + c, v);
      tmem_ld_wait();
      float* dst = p.out + (long long)(warp * 32 + lane) * p.N + c;
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (c + e < p.N) dst[e] = __uint_as_float(v[e]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

int umma_probe(const fb_umma_probe_args* a, cudaStream_t st) {
  FB_REQUIRE(a && a->a_image && a->b_image && a->out, "fb_umma_probe: null arg");
  FB_REQUIRE(a->a_bytes > 0 && a->a_bytes <= (int)PR_A_BYTES && a->a_bytes % 16 == 0, "fb_umma_probe: bad a_bytes");
  FB_REQUIRE(a->b_bytes > 0 && a->b_bytes <= (int)PR_B_BYTES && a->b_bytes % 16 == 0, "fb_umma_probe: bad b_bytes");
  FB_REQUIRE(a->n_mma > 0 && a->n_mma <= 64 && a->N > 0 && a->N <= 256 && a->N % 32 == 0, "fb_umma_probe: bad n_mma / N");
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  p.a_image = (const uint8_t*)a->a_image; p.a_bytes = a->a_bytes;
  p.b_image = (const uint8_t*)a->b_image; p.b_bytes = a->b_bytes;
  p.n_mma = a->n_mma;
  for (int i = 0; i < a->n_mma; ++i) { p.a_off[i] = a->a_off[i]; p.b_off[i] = a->b_off[i]; }
  p.a_desc = a->a_desc; p.b_desc = a->b_desc; p.idesc = a->idesc; p.N = a->N; p.out = a->out;
  p.cycles = a->cycles;
  static bool configured = false;
  if (!configured) {
    FB_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM));
    configured = true;
  }
  umma_probe_kernel<<<1, 160, PR_SMEM, st>>>(p);
  count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb

extern "C" int fb_umma_probe(const fb_umma_probe_args* args, void* stream) {
  return fb::umma_probe(args, reinterpret_cast<cudaStream_t>(stream));
}
