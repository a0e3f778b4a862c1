// Parameter block, epilogue kinds and tile constants of the GEMM kernels (bp_gemm.cuh, bp_gemm2.cuh) — the part of
// the kernel interface that the host side (bp_runtime.cu, which never instantiates the kernels) shares with the launch
// translation unit (bp_launch.cu).
#pragma once
#include <cstdint>

namespace bp {

// L2 eviction-priority policies for TMA operand loads (createpolicy-encoded).
constexpr unsigned long long kEvictNormal = 0x1000000000000000ull;
constexpr unsigned long long kEvictFirst = 0x12F0000000000000ull;
constexpr unsigned long long kEvictLast = 0x14F0000000000000ull;

enum Epi : int {
  EPI_PLAIN = 0,      // out = acc                               (dW gradient tile, debug)
  EPI_FWD_HID = 1,    // y = act(scale*acc + bias[m]) (+dropout)  (hidden layer forward)
  EPI_FWD_OUT = 2,    // o = scale*acc + bias[m]; optional out2 = o; optional out = gscale*(o - targ); optional sqerr
  EPI_DX = 3,         // out = act'(aux) * acc                    (back-prop through the non-linearity)
};

struct GemmParams {
  int M, N, K;            // logical extents (see header comment)
  int k_splits;           // >1 (EPI_PLAIN, lone-CTA kernel only): the K range is cut into k_splits slices, slice s writes
                          // its partial product to out + s*split_stride; a finishing kernel adds the slices in order
  long long split_stride; // floats between partial-product planes
  int n_begin;            // first column of this launch (multiple of BLOCK_N): columns [n_begin, N) are computed, so a
                          // product can be cut into column slices whose all-reduce starts while the rest computes
  float* out;             // primary output, element (m,n) at out[n*ldo + m]; may be null for EPI_FWD_OUT
  long long ldo;
  const float* bias;      // bias[m]                      (EPI_FWD_*)
  const float* aux;       // targ (EPI_FWD_OUT) or Y (EPI_DX), element (m,n) at aux[n*ldaux + m]
  long long ldaux;
  float* out2;            // raw linear output (EPI_FWD_OUT), element (m,n) at out2[n*ldo2 + m]
  long long ldo2;
  double* sqerr;          // if non-null (EPI_FWD_OUT): += sum (o - targ)^2
  float scale;            // multiplies the accumulator (inference-time keep probability, BP_GPU.cu:705-732)
  float gscale;           // 2/B of kernSubClean (DevFunc.cu:263)
  int act;                // 0 = ReLU (HEAD, DevFunc.cu:67-97), 1 = sigmoid (commented variant :52,:62)
  float drop_p;           // >0: zero y where u < drop_p (kernDropout DevFunc.cu:34-45), no rescale
  uint32_t seed_lo, seed_hi, step, layer;
  int frame0;             // global frame index of column n=0 (data-parallel shard offset), multiple of 4
  int passes;             // 1 = single-pass TF32; 3 = split precision (3xTF32): A*B + A_lo*B + A*B_lo, where X_lo =
                          // X - trunc_tf32(X) is kept as a second fp32 array by whoever writes X (~fp32 accuracy)
  float* out_lo;          // if non-null: out_lo[...] = v - trunc_tf32(v) for every v stored to `out` (same layout)
  // data-parallel reduce-scatter fused into the epilogue (EPI_PLAIN, see bp_peer.cuh): the 32-column chunk starting at
  // column nc goes to scatter[(chunk_base + nc/32) % scatter_n] (a peer's receive slab, same layout as `out`)
  int scatter_n;
  int chunk_base;
  float* scatter[8];
  unsigned long long hint_a, hint_b;  // L2 eviction-priority policy for the A / B operand loads (0 = none)
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // two 32-float (128-byte) swizzle spans per stage
constexpr int GEMM_THREADS = 192;

}  // namespace bp
