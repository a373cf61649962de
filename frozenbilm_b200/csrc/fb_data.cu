// Input-side kernels of the hot path: the batch that reaches
// DebertaV2ForMaskedLM.forward is produced on the device, so the training loop has no host-side tensor work and no
// synchronising H2D copies between steps.
//
//   fb_mlm_corrupt    masked-LM corruption of the token ids + labels (reference util/misc.py:14-56 mask_tokens:
//                     Bernoulli(p) over non-special, non-pad tokens; of those 80 % -> [MASK], 10 % -> a uniformly random
//                     id, 10 % unchanged; labels = original id at the selected positions, -100 elsewhere).
//                     The reference draws from torch's global generator; here every decision is a pure function of
//                     (seed, flat token index) so that the CPU oracle (oracle/data_oracle.py) reproduces it bit for bit.
//   fb_gather_frames  frame subsampling / zero padding of per-video CLIP features stored back to back in one fp16
//                     shard (reference datasets/videotext_dataset.py:23-44: more frames than max_feats -> frame
//                     (j * n) // max_feats, fewer -> zero padded, missing video (n = 0) -> zeros, video_len = 0).
#include <cuda_fp16.h>

#include "../../include/frozenbilm_b200.h"
#include "fb_common.cuh"

namespace fb {

// three independent 32-bit draws per token: counter-based, see oracle/data_oracle.py for the same arithmetic
__device__ __forceinline__ uint32_t data_hash(uint32_t idx_lo, uint32_t idx_hi, uint32_t seed_lo, uint32_t seed_hi,
                                              uint32_t stream) {
  uint32_t x = idx_lo * 0x9E3779B1u + seed_lo;
  x ^= (idx_hi + stream) * 0x85EBCA77u + seed_hi;
  x = mix32(x);
  x = mix32(x + stream * 0xC2B2AE3Du + 0x27D4EB2Fu);
  return x;
}

__global__ void mlm_corrupt_kernel(const long long* __restrict__ ids, const uint8_t* __restrict__ special, long long n,
                                   long long pad_id, long long mask_id, long long vocab, uint32_t p_thresh,
                                   uint32_t seed_lo, uint32_t seed_hi, long long* __restrict__ ids_out,
                                   long long* __restrict__ labels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long id = ids[i];
  const bool eligible = !(special != nullptr && special[i] != 0) && !(pad_id >= 0 && id == pad_id);
  const uint32_t lo = (uint32_t)i, hi = (uint32_t)((unsigned long long)i >> 32);
  const uint32_t u0 = data_hash(lo, hi, seed_lo, seed_hi, 0u);
  const bool masked = eligible && (u0 < p_thresh);
  long long out = id;
  if (masked) {
    const uint32_t u1 = data_hash(lo, hi, seed_lo, seed_hi, 1u);
    if (u1 < 3435973837u) {                       // 0.8 * 2^32
      out = mask_id;
    } else if (u1 >= 3865470566u) {               // upper 10 %: random word; the 10 % in between keep the token
      const uint32_t u2 = data_hash(lo, hi, seed_lo, seed_hi, 2u);
      out = (long long)(((unsigned long long)u2 * (unsigned long long)vocab) >> 32);   // uniform in [0, vocab)
    }
  }
  ids_out[i] = out;
  labels[i] = masked ? id : -100;
}

// one CTA per (sample, output frame); 16-byte loads of 8 halves, fp32 or bf16 output
__global__ void __launch_bounds__(128)
gather_frames_kernel(const __half* __restrict__ shard, const long long* __restrict__ offsets,
                     const int* __restrict__ nframes, int F, int D, float* __restrict__ out_f32,
                     __nv_bfloat16* __restrict__ out_bf16, long long* __restrict__ video_len) {
  const int j = blockIdx.x, b = blockIdx.y;
  const int n = nframes[b];
  if (j == 0 && threadIdx.x == 0 && video_len != nullptr) video_len[b] = n < F ? (n > 0 ? n : 0) : F;
  long long src_row = -1;
  if (n > F) src_row = ((long long)j * n) / F;
  else if (j < n) src_row = j;
  const __half* src = src_row >= 0 ? shard + (offsets[b] + src_row) * (long long)D : nullptr;
  const long long o = ((long long)b * F + j) * D;
  const bool vec = (D % 8 == 0) && (src == nullptr || ((reinterpret_cast<uintptr_t>(src) & 15) == 0));
  if (vec) {
    for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
      float v[8];
      if (src != nullptr) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + c));
        const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h2[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.0f;
      }
      if (out_f32 != nullptr) {
        *reinterpret_cast<float4*>(out_f32 + o + c) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(out_f32 + o + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (out_bf16 != nullptr)
        *reinterpret_cast<uint4*>(out_bf16 + o + c) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  } else {
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      const float v = src != nullptr ? __half2float(src[c]) : 0.0f;
      if (out_f32 != nullptr) out_f32[o + c] = v;
      if (out_bf16 != nullptr) out_bf16[o + c] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace fb

extern "C" {
int fb_mlm_corrupt(const long long* ids, const unsigned char* special_mask, long long n, long long pad_id,
                   long long mask_id, long long vocab, float mlm_probability, unsigned long long seed,
                   long long* ids_out, long long* labels_out, void* stream) {
  FB_REQUIRE(ids && ids_out && labels_out && n > 0, "fb_mlm_corrupt: null arg / empty input");
  FB_REQUIRE(vocab > 0 && vocab < (1LL << 31) && mask_id >= 0 && mask_id < vocab, "fb_mlm_corrupt: bad vocab / mask id");
  FB_REQUIRE(mlm_probability >= 0.0f && mlm_probability <= 1.0f, "fb_mlm_corrupt: probability out of range");
  const double t = (double)mlm_probability * 4294967296.0;
  const uint32_t thresh = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
  const unsigned grid = (unsigned)((n + 255) / 256);
  fb::mlm_corrupt_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ids, special_mask, n, pad_id, mask_id, vocab, thresh, (uint32_t)seed, (uint32_t)(seed >> 32), ids_out, labels_out);
  fb::count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int fb_gather_frames(const void* shard_f16, const long long* offsets, const int* nframes, int B, int max_feats, int D,
                     float* out_f32, void* out_bf16, long long* video_len, void* stream) {
  FB_REQUIRE(shard_f16 && offsets && nframes && (out_f32 || out_bf16), "fb_gather_frames: null arg");
  FB_REQUIRE(B > 0 && max_feats > 0 && D > 0, "fb_gather_frames: bad shape");
  fb::gather_frames_kernel<<<dim3(max_feats, B), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(shard_f16), offsets, nframes, max_feats, D, out_f32,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), video_len);
  fb::count_launch();
  FB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}
