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
  EPI_DW_SGD = 4,     // acc = gradient tile, consumed in place: momentum-SGD update of the same tile of the weight and
                      // delta arenas (kernUpdatedelta + kernAccSum, DevFunc.cu:313-318, 270-277); no gradient is stored
  // ReLU nets (BP_RELU_MASK=1, off by default): kernDsigmoid's ReLU' is the predicate y > 0 (DevFunc.cu:81-97), so the
  // forward epilogue can leave one BIT per activation and the dX epilogue read 1/32 of what re-reading Y costs it.
  EPI_FWD_HID_MASK = 5,  // EPI_FWD_HID + relu_mask word (32 frames x 1 unit) per accumulator chunk
  EPI_DX_MASK = 6,       // out = bit ? acc : 0  — EPI_DX for act == 0 with the mask instead of aux (bit-identical)
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
  int stream_out;         // EPI_PLAIN: store with st.global.cs (evict-first), BP_DW_STREAM=1 (default off; untested A/B)
  int l2_prefetch;        // > 0: the producer also issues cp.async.bulk.prefetch.tensor (L2 only) for the k-block this
                          // many steps ahead of the one it loads, and for the first ones before griddepcontrol.wait
                          // (BP_L2_PREFETCH; an L2 prefetch of data a predecessor is still writing is harmless — L2 is
                          // the point of coherence)
  uint32_t dbg_flags;     // measurement aids: bit 0 skip the MMAs (TMA-only), bit 1 skip the loads (MMA-only)
  long long* dbg_trace;   // if non-null, CTA 0 records clock64() per k-block: [0..255] producer slot free,
                          // [256..511] loads issued, [512..767] stage full seen, [768..1023] MMAs issued,
                          // [1024] accumulator ready seen by epilogue, [1025] epilogue done, [1026] CTA start
  // EPI_DW_SGD (single GPU, BP_FUSED_UPDATE=1): the epilogue of the weight-gradient product applies the update to its
  // own tile — element (m,n) of the product is parameter upd_w[n*ldo + m] (same layout as `out`).  Replaces the
  // gradient store (4 B/param) + bp_sgd_kernel (20 B/param) by 16 B/param inside the GEMM, under the next tile's MMAs.
  float* upd_w;
  float* upd_delta;
  float* upd_w_lo;        // 3xTF32: low part of the new weights, or null
  float upd_nf;           // `n` of kernUpdatedelta (global bunch), int promoted to float
  float upd_inv_nf;       // 1/upd_nf if that is exact (power of two: g*2^-k == g/2^k bit for bit), else 0 -> divide
  float upd_momentum, upd_c1, upd_wc;  // c1 = (1-momentum)*lr
  int upd_bias_col;       // product column that is the bias row of the block (weight cost does not apply, BP_GPU.cu:648)
  int upd_prefetch;       // 1: epilogue warps pull their tile's delta/w lines into L2 while the main loop runs
  // EPI_FWD_HID_MASK / EPI_DX_MASK: bit j of relu_mask[(n/32)*ldmask + m] = (Y[n - n%32 + j][m] > 0), Y as stored
  // (after the activation and the dropout mask); one coalesced 128-byte row of words per warp and 32-column chunk.
  uint32_t* relu_mask;
  long long ldmask;
};

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // two 32-float (128-byte) swizzle spans per stage
constexpr int GEMM_THREADS = 192;

}  // namespace bp
