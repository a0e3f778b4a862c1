/* bp_gpu_debug.h — TEST AND BRING-UP AIDS of libbpgpu.so.  NOT part of the drop-in boundary (include/bp_gpu.h): nothing
 * a caller of the reference's BP_GPU needs is declared here.  tests/ use bp_debug_gemm / bp_debug_sgd to compare single
 * kernels with the oracle, bp_debug_plan / bp_debug_chain_plan to pin the host-side kernel choice and tile schedule
 * without a GPU; scripts/ use the rest for measurements recorded under profiles/. */
#ifndef BP_GPU_DEBUG_H_
#define BP_GPU_DEBUG_H_
#include "bp_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Host-only (no GPU needed): the kernel the library would pick for an M x N x K product on a device with num_sms SMs —
 * *pair_n = 0 (128 x 128 tiles on lone CTAs), 128 or 256 (256 x pair_n tiles on CTA pairs) — the number of output
 * tiles, the CTAs launched (persistent, <= one per SM) and the 64-deep k-blocks per tile.  have_b64: the B operand
 * also has a 64-row-box tensor map (hidden-layer forward and dX products); max_pairs <= 0 assumes num_sms / 2. */
int bp_debug_plan(int M, int N, int K, int have_b64, int num_sms, int max_pairs, int* pair_n, int* tiles, int* ctas,
                  int* k_blocks);
/* Stand-alone launch of one fused GEMM (host operands, testing only).
 *  kind 0: fwd  out[n*ldo+m] = act(scale * sum_k W[k*ldw+m] * X[n*ldx+k] + bias[m])     A=W (K x M), B=X (N x K)
 *  kind 1: dX   out[n*ldo+m] = act'(Y[n*ldy+m]) * sum_k W[m*ldw+k] * D[n*ldd+k]         A=W (M x K), B=D (N x K)
 *  kind 2: dW   out[n*ldo+m] = sum_k D[k*ldd+m] * X[k*ldx+n]                            A=D (K x M), B=X (K x N)
 *  kind 3: plain fwd (no bias/activation).   act < 0 means "no activation" for kind 0. */
int bp_debug_gemm(int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* out,
                  int ldo, const float* bias, const float* aux, int ldaux, float scale, int act, int math_mode,
                  float* elapsed_ms);

/* Stand-alone fused SGD update on host arrays of length n (testing only): kernUpdatedelta + kernAccSum
 * (DevFunc.cu:313-318, 270-277). */
int bp_debug_sgd(int n, float* delta, float* weights, const float* grad, int bunch, float momentum, float lrate,
                 float weightcost);

/* Bring-up microbenchmark (testing only): cycles per 128 x bn x 8 TF32 tcgen05.mma on one SM with both operands in
 * shared memory.  combo 0 = A MN-major/B K-major, 1 = K/K, 2 = MN/MN, 3 = K/MN. */
int bp_debug_mma_rate(int combo, int bn, int iters, int mode, double* cyc_issue, double* cyc_total);

/* Bring-up aid of the chained launches (bp_set_option(h, "chain_trace", 1) first): per-tile timestamps of the most
 * recent forward (which = 0) or back-propagation (which = 1) launch, CSV text. */
int bp_debug_chain_trace(bp_handle* h, int which, char* buf, int len);

/* Bring-up aid: GB/s of the GEMM epilogues' store pattern (layout 0: tiles of a row-major matrix, 1: contiguous tiles). */
int bp_debug_store_pattern(int layout, int ld, int mb, int reps, int ctas, double* gbs);

/* Host-only (no GPU needed): the tile schedule the chained launches (csrc/bp_chain.cuh) would run for one train bunch of
 * `rows` frames of the net `layersizes[0..numlayers-1]` on `pairs` CTA pairs.  which = 0 forward, 1 back-propagation.
 * Fills, per scheduled tile in pair-major order: pair[], prod[], mt[], nt[] (at most max_items, *n_items = count), and
 * per product (at most max_prods, *n_prods = count): dep_prod[], dep_all[], m_tiles[], n_tiles[], pair_n[], n_cols[].
 * *makespan_cycles = the cost model's estimate.  Used by tests/test_chain_schedule.py to check that every tile is
 * scheduled exactly once and that "pair sequence + dependencies" is acyclic. */
int bp_debug_chain_plan(int numlayers, const int* layersizes, int rows, int pairs, int which, int passes,
                        int max_items, int* n_items, int* pair, int* prod, int* mt, int* nt, int max_prods, int* n_prods,
                        int* dep_prod, int* dep_all, int* m_tiles, int* n_tiles, int* pair_n, int* n_cols,
                        long long* makespan_cycles);

#ifdef __cplusplus
}
#endif
#endif /* BP_GPU_DEBUG_H_ */
