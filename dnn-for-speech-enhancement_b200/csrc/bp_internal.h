// Host-side internals shared by the two translation units of libbpgpu.so:
//   bp_launch.cu   instantiates and launches the GEMM kernel templates (tensor maps, kernel choice per product,
//                  cluster capacity, split-K finisher, MMA-rate microbenchmark) — the slow one to compile;
//   bp_runtime.cu  everything else: per-rank state, chunk pipeline, forward / train scheduling, data-parallel
//                  exchange, the C ABI of include/bp_gpu.h.
// Nothing here is part of the ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

#include "../../include/bp_gpu.h"
#include "../../include/bp_gpu_debug.h"
#include "bp_gemm_params.h"
#include "bp_chain_params.h"

namespace bp {

extern thread_local std::string g_err;      // message behind bp_last_error()
int fail(int code, const char* fmt, ...);   // sets g_err, returns code

#define CU_TRY(expr)                                                                                            \
  do {                                                                                                          \
    cudaError_t e__ = (expr);                                                                                   \
    if (e__ != cudaSuccess)                                                                                     \
      return ::bp::fail(BP_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define BP_TRY(expr)              \
  do {                            \
    int r__ = (expr);             \
    if (r__ != BP_OK) return r__; \
  } while (0)

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// Tile width along N of the lone-CTA kernel and of the B-operand boxes built for it.
constexpr int kBlockN = 128;
constexpr int kMaxSplits = 8;  // split-K planes of the output-layer product (see out_layer_splits)

// An operand and (split-precision mode only) its low part X_lo = X - trunc_tf32(X), same shape and stride.
struct MapPair {
  CUtensorMap m;
  CUtensorMap lo;
};
// The A operand of a product (128 rows per box).
struct AMaps {
  MapPair full;
};
int make_map1(CUtensorMap* m, const float* base, long long contiguous_extent, long long rows, long long ld,
              int box_outer, bool mn_major, int k_chunks = GEMM_BLOCK_K / 32);
int make_map(MapPair* mp, const float* base, const float* lo_base, long long contiguous_extent, long long rows,
             long long ld, int box_outer, bool mn_major, int k_chunks = GEMM_BLOCK_K / 32);
int make_a_maps(AMaps* am, const float* base, const float* lo_base, long long contiguous_extent, long long rows,
                long long ld, bool mn_major);

// The products of the hot path = (operand majors, epilogue) combinations the kernel templates are instantiated for.
enum Product : int {
  PROD_FWD_HID,     // A = W^T MN-major, B = Y K-major,    EPI_FWD_HID  hidden-layer forward affine
  PROD_FWD_OUT,     //   "                                 EPI_FWD_OUT  output-layer forward (+ loss gradient)
  PROD_FWD_PLAIN,   //   "                                 EPI_PLAIN    tests / probes
  PROD_FWD_SPLITK,  //   "    lone-CTA kernel only,        EPI_PLAIN    split-K planes of the output layer (p.k_splits)
  PROD_DX,          // A = W K-major, B = D K-major,       EPI_DX       back-prop through a layer
  PROD_DW,          // A = D^T MN-major, B = Y^T MN-major, EPI_PLAIN    weight (+bias) gradient
};
// Picks the kernel (lone CTAs / 128- / 256-wide CTA pairs, see pick_kernel in bp_launch.cu) and launches it on `st`.
// b = B map with 128-wide boxes, b64 = the same operand with 64-wide boxes (for 128-wide pairs) or null.
int launch_product(Product prod, cudaStream_t st, int num_sms, const AMaps& a, const MapPair& b, const GemmParams& p,
                   const MapPair* b64 = nullptr);
// Second half of the split-K output-layer product (bp_out_finish_kernel): ws = the planes, ldw their row stride.
int launch_out_finish(cudaStream_t st, const float* ws, long long ldw, const GemmParams& p, long long cells);
void init_cluster_capacity(int num_sms);
// Chained products (bp_chain.cuh): co-resident CTA pairs on the current device; one launch over a schedule.
int chain_max_pairs(int num_sms);
int launch_chain(cudaStream_t st, const ChainArgs& a, int pairs);

// Process-wide scheduling / kernel-choice switches.  Each starts from its environment variable (scripts/SWITCHES.md),
// read once at first use, and can be changed between launches through bp_set_option(h, name, value) — so that
// alternatives are A/B-timed inside ONE process on one resident chunk (scripts/gpu_ab_inproc.py) instead of one process
// per setting.  None of them changes a result bit.
enum Tunable : int {
  TUN_PDL,          // "pdl"          BP_PDL          1   programmatic dependent launch between consecutive GEMMs
  TUN_TMA_HINT,     // "tma_hint"     BP_TMA_HINT     1   L2 eviction-priority hints on operand loads
  TUN_PAIRS,        // "pairs"        BP_PAIRS        1   0 lone CTAs, 1 automatic, 2 always 256-wide, 3 128-wide pairs
  TUN_SGD_STREAM,   // "sgd_stream"   BP_SGD_STREAM   1   streaming access to the momentum deltas in the update
  TUN_SGD_EARLY,    // "sgd_early"    BP_SGD_EARLY    6   blocks per SM of the early update of layers >= 2 (0 = off)
  TUN_SPLITK,       // "splitk"       BP_SPLITK      -1   output-layer K slices: -1 automatic, 0 never, N force
  TUN_COUNT
};
int tunable(Tunable t);
int set_tunable(const char* name, int value);  // BP_OK, or BP_EINVAL (without touching g_err) for an unknown name
int get_tunable(const char* name, int* value);  // BP_OK / BP_EINVAL, as set_tunable

}  // namespace bp
