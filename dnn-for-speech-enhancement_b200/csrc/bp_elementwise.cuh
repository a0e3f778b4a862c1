// HBM-bound element-wise kernels of the hot path (sm_100a): fused momentum-SGD update, input-layer dropout,
// constant-column fill.  All are pure streaming kernels: 128-bit accesses, grid sized in multiples of the SM count.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "bp_rng.cuh"

namespace bp {

// Rows of the parameter arena that hold biases (weight cost does not apply: the reference passes weightcost = 0
// for biases, BP_GPU.cu:648).  Ranges are in float4 units.
struct SgdBiasRanges {
  int n;
  long long begin4[10];
  long long end4[10];
};

// kernUpdatedelta + kernAccSum fused (DevFunc.cu:313-318, 270-277; call sites BP_GPU.cu:643,648,651,652):
//   delta = momentum*delta - (1-momentum)*lr*(grad/n + weightcost*w);  w = delta + 1.0*w
// evaluated literally in fp32, one rounding per operation (no FMA contraction) so the CPU oracle can match bit-wise.
// Traffic: reads delta,w,grad (12 B) + writes delta,w (8 B) = 20 B per parameter (the reference moves 28).
// w_lo (optional): w - trunc_tf32(w) for the split-precision (3xTF32) GEMMs, +4 B per parameter.
// L2 policy: the momentum deltas and the gradients are touched by nobody else until the next update, so they are
// streamed (evict-first loads/stores) and do not push the weights — which the next bunch's forward and dX GEMMs read —
// out of the 126 MB L2 (the three arenas of the C2 net are 176 MB).
template <bool kHasWC>
__global__ void __launch_bounds__(256)
bp_sgd_kernel(float4* __restrict__ delta, float4* __restrict__ w, const float4* __restrict__ grad, long long begin4,
              long long n4, float nf, float momentum, float one_minus_m_lr, float weightcost, SgdBiasRanges br,
              float4* __restrict__ w_lo, int stream_delta) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  // float4 elements [begin4, n4) of the arena: the update may be issued in two parts (upper layers early, see
  // train_bunch), with the bias ranges staying absolute
  for (long long i = begin4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = __ldcs(grad + i);  // gradient is dead after this read: streaming load
    float4 d = stream_delta ? __ldcs(delta + i) : delta[i];
    float4 x = w[i];
    float wc = weightcost;
    if (kHasWC) {
#pragma unroll 1
      for (int r = 0; r < br.n; ++r)
        if (i >= br.begin4[r] && i < br.end4[r]) wc = 0.0f;
    }
    const float gv[4] = {g.x, g.y, g.z, g.w};
    float dv[4] = {d.x, d.y, d.z, d.w};
    float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float t = __fdiv_rn(gv[k], nf);
      if (kHasWC) t = __fadd_rn(t, __fmul_rn(wc, xv[k]));
      const float nd = __fsub_rn(__fmul_rn(momentum, dv[k]), __fmul_rn(one_minus_m_lr, t));
      dv[k] = nd;
      xv[k] = __fadd_rn(nd, xv[k]);
    }
    if (stream_delta) __stcs(delta + i, make_float4(dv[0], dv[1], dv[2], dv[3]));
    else delta[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    w[i] = make_float4(xv[0], xv[1], xv[2], xv[3]);
    if (w_lo != nullptr) {
      float lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) lo[k] = xv[k] - __uint_as_float(__float_as_uint(xv[k]) & 0xFFFFE000u);
      w_lo[i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// Input-layer dropout applied in place to the device copy of the bunch (BP_GPU.cu:536-540: curandGenerateUniform +
// kernDropout with p = visible_omit on `in`).  Thread = (4 consecutive frames) x (1 input unit); warps run along the
// contiguous unit dimension so every access is coalesced.  Mask key: activation tensor 0.
__global__ void __launch_bounds__(256)
bp_input_dropout_kernel(float* __restrict__ x, long long ldx, int frames, int units, float p, uint32_t seed_lo,
                        uint32_t seed_hi, uint32_t step, int frame0) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  const int fq = blockIdx.y;  // local frame quad
  if (u >= units) return;
  float r[4];
  philox_uniform4(seed_lo, seed_hi, static_cast<uint32_t>(frame0 + fq * 4) >> 2, static_cast<uint32_t>(u), 0u, step,
                  r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int f = fq * 4 + j;
    if (f < frames && r[j] < p) x[static_cast<long long>(f) * ldx + u] = 0.0f;
  }
}

// lo[i] = x[i] - trunc_tf32(x[i]) (split-precision operand preparation for uploaded inputs / weights).
__global__ void __launch_bounds__(256)
bp_split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, long long n4) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = x[i];
    float4 r;
    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    lo[i] = r;
  }
}

// buf[r*ld + col] = value for r in [0, rows): the all-ones input column that makes the dW GEMM also produce the
// bias gradient (kernAccSumrow, DevFunc.cu:224-242).
__global__ void bp_fill_col_kernel(float* __restrict__ buf, long long ld, long long rows, int col, float value) {
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r < rows) buf[r * ld + col] = value;
}

}  // namespace bp
