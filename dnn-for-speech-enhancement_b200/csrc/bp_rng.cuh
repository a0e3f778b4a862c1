// Counter-based dropout RNG (Philox-4x32-10, Salmon et al. SC'11) keyed on
// (seed ; frame/4, unit, layer, step) so that a mask element depends only on WHICH element it is, never on which
// GPU / tile / thread produced it: 1-GPU and N-GPU runs draw identical masks, and the CPU oracle can regenerate them.
// Replaces curandGenerateUniform (XORWOW seeded from time(NULL), BP_GPU.cu:69,77-78,538,547) + kernDropout
// (DevFunc.cu:34-45): drop where u < p, u in (0,1], no 1/keep rescale.
#pragma once
#include <cstdint>

namespace bp {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Same mapping as cuRAND's uniform: x * 2^-32 + 2^-33  ->  (0, 1]; one fused multiply-add so host == device bit-wise.
__host__ __device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  return fmaf(static_cast<float>(x), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Uniforms for frames 4*fq .. 4*fq+3 of unit `unit` (layer = index of the layer whose INPUT is masked).
__host__ __device__ __forceinline__ void philox_uniform4(uint32_t seed_lo, uint32_t seed_hi, uint32_t fq,
                                                         uint32_t unit, uint32_t layer, uint32_t step,
                                                         float (&u)[4]) {
  uint32_t r[4];
  philox4x32_10(fq, unit, layer, step, seed_lo, seed_hi, r);
  u[0] = u32_to_uniform(r[0]);
  u[1] = u32_to_uniform(r[1]);
  u[2] = u32_to_uniform(r[2]);
  u[3] = u32_to_uniform(r[3]);
}

}  // namespace bp
