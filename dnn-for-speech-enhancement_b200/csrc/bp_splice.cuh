// Device side of the chunk reader (sm_100a): byte-swap, normalise, 11-frame splice, NAT block, target pick and
// shuffle-scatter of Interface::Readchunk (reference Interface.cc:737-787, 828-853) applied to the RAW Pfile records of
// a chunk.  The host ships ~1 KB per frame (one feature record + one target record) instead of the ~12 KB spliced
// sample, and never touches the samples.  HBM-bound streaming: one CTA per sample, every access coalesced.
//
// Arithmetic is the reader's, operation for operation, so rows are bit-identical with the host reader:
//   x = (bswap(word) - mean[j]) * inv_std[j]      two fp32 roundings, no FMA        (Interface.cc:745-746)
//   nat[j] = (x0 + x1 + ... + x5) / 6.0f           left to right over the segment's first six frames; frames past the
//                                                  end of the chunk's record block count as 0     (Interface.cc:776-779)
//   target row = records[first + targ_offset]      byte-swapped only                 (Interface.cc:815-816, 844-846)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bp {

struct SpliceParams {
  const uint32_t* fea;     // n_records x (2 + dim) big-endian words
  const uint32_t* targ;    // n_records x (2 + out_dim) big-endian words, or null
  const float* mean;       // dim
  const float* inv_std;    // dim
  const int* sample_frame; // n_samples
  const int* sample_seg;   // n_samples (nat only)
  const int* sample_row;   // n_samples or null (identity)
  int dim, ctx, nat, targ_offset, out_dim;
  int n_records, n_samples;
  float* x;                // chunk inputs, row stride ldx
  long long ldx;
  float* t;                // chunk targets, row stride out_dim
  // data-parallel row ownership: rank r keeps rows [r*lb, (r+1)*lb) of every global bunch of B rows
  int world, rank, B, lb;
  int full_rows;           // world > 1: rows >= full_rows (the trailing partial bunch) are dropped
};

__device__ __forceinline__ float splice_be_float(uint32_t w) { return __uint_as_float(__byte_perm(w, 0u, 0x0123u)); }

__global__ void __launch_bounds__(256)
bp_splice_kernel(const SpliceParams p) {
  const int rec_w = p.dim + 2;
  const int span = p.ctx * p.dim;
  for (int i = blockIdx.x; i < p.n_samples; i += gridDim.x) {
    const int d = p.sample_row != nullptr ? __ldg(p.sample_row + i) : i;
    long long row = d;
    if (p.world > 1) {
      if (d >= p.full_rows) continue;
      const int in_b = d % p.B;
      if (in_b / p.lb != p.rank) continue;
      row = static_cast<long long>(d / p.B) * p.lb + (in_b - p.rank * p.lb);
    }
    const int f0 = __ldg(p.sample_frame + i);
    const uint32_t* src = p.fea + static_cast<size_t>(f0) * rec_w + 2;
    float* xr = p.x + row * p.ldx;
    for (int e = threadIdx.x; e < span; e += blockDim.x) {
      const int q = e / p.dim, j = e - q * p.dim;
      const float v = splice_be_float(__ldg(src + static_cast<size_t>(q) * rec_w + j));
      xr[e] = __fmul_rn(__fsub_rn(v, __ldg(p.mean + j)), __ldg(p.inv_std + j));
    }
    if (p.nat) {
      const int seg = __ldg(p.sample_seg + i);
      for (int j = threadIdx.x; j < p.dim; j += blockDim.x) {
        const float m = __ldg(p.mean + j), iv = __ldg(p.inv_std + j);
        float s = 0.0f;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int fr = seg + q;
          float v = 0.0f;
          if (fr < p.n_records)
            v = __fmul_rn(__fsub_rn(splice_be_float(__ldg(p.fea + static_cast<size_t>(fr) * rec_w + 2 + j)), m), iv);
          s = q == 0 ? v : __fadd_rn(s, v);
        }
        xr[span + j] = __fdiv_rn(s, 6.0f);
      }
    }
    if (p.targ != nullptr) {
      const uint32_t* ts = p.targ + static_cast<size_t>(f0 + p.targ_offset) * (p.out_dim + 2) + 2;
      float* tr = p.t + row * p.out_dim;
      for (int k = threadIdx.x; k < p.out_dim; k += blockDim.x) tr[k] = splice_be_float(__ldg(ts + k));
    }
  }
}

}  // namespace bp
