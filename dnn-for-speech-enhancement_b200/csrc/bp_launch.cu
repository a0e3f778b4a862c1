// libbpgpu.so, launch half: the only translation unit that instantiates the tcgen05 GEMM kernel templates
// (bp_gemm.cuh, bp_gemm2.cuh).  Tensor-map construction, kernel choice per product, cluster capacity, the split-K
// finisher and the MMA-rate microbenchmark live here; bp_runtime.cu reaches them through bp_internal.h.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "bp_gemm.cuh"
#include "bp_gemm2.cuh"
#include "bp_chain.cuh"
#include "bp_internal.h"
#include "bp_microbench.cuh"

namespace bp {

// ------------------------------------------------------------------------------------------------ tunables
namespace {
struct TunableDef {
  const char* name;
  const char* env;
  int dflt;
};
const TunableDef kTunables[TUN_COUNT] = {{"pdl", "BP_PDL", 1},         {"tma_hint", "BP_TMA_HINT", 1},
                                            {"pairs", "BP_PAIRS", 1},     {"sgd_stream", "BP_SGD_STREAM", 1},
                                            {"sgd_early", "BP_SGD_EARLY", 6}, {"splitk", "BP_SPLITK", -1}};
std::atomic<int> g_tunable[TUN_COUNT];
std::once_flag g_tunable_once;
void init_tunables() {
  std::call_once(g_tunable_once, [] {
    for (int i = 0; i < TUN_COUNT; ++i) {
      const char* e = getenv(kTunables[i].env);
      g_tunable[i].store(e ? atoi(e) : kTunables[i].dflt, std::memory_order_relaxed);
    }
  });
}
}  // namespace

int tunable(Tunable t) {
  init_tunables();
  return g_tunable[t].load(std::memory_order_relaxed);
}

int set_tunable(const char* name, int value) {
  init_tunables();
  for (int i = 0; i < TUN_COUNT; ++i)
    if (strcmp(name, kTunables[i].name) == 0) {
      g_tunable[i].store(value, std::memory_order_relaxed);
      return BP_OK;
    }
  return BP_EINVAL;
}

int get_tunable(const char* name, int* value) {
  init_tunables();
  for (int i = 0; i < TUN_COUNT; ++i)
    if (strcmp(name, kTunables[i].name) == 0) {
      *value = g_tunable[i].load(std::memory_order_relaxed);
      return BP_OK;
    }
  return BP_EINVAL;
}

// ------------------------------------------------------------------------------------------------ driver entry
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

static int get_encode() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (!g_encode) return fail(BP_ECUDA, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  return BP_OK;
}

// 3-D view {32 floats, rows (stride ld floats), ld/32 chunks (stride 128 B)} of a row-major fp32 matrix [rows x ld]
// (see the "Operand fetch" paragraph of bp_gemm.cuh).  OOB rows / chunks are zero-filled.
//   K-major operand  (reduction dim contiguous): rows = exact M/N extent, chunks = ceil(K/32); box {32, box_rows, 2};
//                    SWIZZLE_128B.
//   MN-major operand (M/N dim contiguous):       rows = exact K extent,   chunks = ceil(MN/32); box {32, 64, box_mn/32};
//                    SWIZZLE_128B_ATOM_32B (the only layout tcgen05 accepts for MN-major 32-bit operands).
int make_map1(CUtensorMap* m, const float* base, long long contiguous_extent, long long rows, long long ld,
              int box_outer, bool mn_major, int k_chunks) {
  BP_TRY(get_encode());
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld & 31) != 0)
    return fail(BP_EINVAL, "tensor map: base not 16-byte aligned or ld %% 32 != 0 (ld=%lld)", ld);
  const long long chunks = (contiguous_extent + 31) / 32;
  if (chunks * 32 > ld) return fail(BP_EINVAL, "tensor map: extent %lld exceeds ld %lld", contiguous_extent, ld);
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(chunks)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4, 128};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(mn_major ? GEMM_BLOCK_K : box_outer),
                       static_cast<cuuint32_t>(mn_major ? box_outer / 32 : k_chunks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(BP_ECUDA, "cuTensorMapEncodeTiled failed (%d) extent=%lld rows=%lld ld=%lld box=%d mn=%d", (int)r,
                contiguous_extent, rows, ld, box_outer, (int)mn_major);
  return BP_OK;
}

int make_map(MapPair* mp, const float* base, const float* lo_base, long long contiguous_extent, long long rows,
             long long ld, int box_outer, bool mn_major, int k_chunks) {
  BP_TRY(make_map1(&mp->m, base, contiguous_extent, rows, ld, box_outer, mn_major, k_chunks));
  if (lo_base) BP_TRY(make_map1(&mp->lo, lo_base, contiguous_extent, rows, ld, box_outer, mn_major, k_chunks));
  else mp->lo = mp->m;
  return BP_OK;
}

int make_a_maps(AMaps* am, const float* base, const float* lo_base, long long contiguous_extent, long long rows,
                long long ld, bool mn_major) {
  return make_map(&am->full, base, lo_base, contiguous_extent, rows, ld, GEMM_BLOCK_M, mn_major);
}

// ------------------------------------------------------------------------------------------------ GEMM launch
template <bool kAMN, bool kBMN, int kEpi, int BN>
static int launch_gemm_bn(cudaStream_t st, int num_sms, const MapPair& a, const MapPair& b, const GemmParams& p) {
  auto kern = bp_gemm_kernel<kAMN, kBMN, kEpi, BN>;
  constexpr size_t smem = gemm_smem_bytes<BN>();
  static thread_local int configured_dev = -1;
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev;
  }
  const int mt = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int nt = (p.N - p.n_begin + BN - 1) / BN;
  const int tiles = mt * nt;
  if (tiles <= 0 || p.K <= 0) return fail(BP_EINVAL, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  if (p.k_splits > 1 && kEpi != EPI_PLAIN) return fail(BP_EINVAL, "gemm: split-K needs the plain epilogue");
  const int grid = std::min(tiles * std::max(1, p.k_splits), num_sms);
  const bool use_pdl = tunable(TUN_PDL) != 0;  // programmatic dependent launch between consecutive GEMMs (default on)
  GemmParams q = p;
  const bool use_hints = tunable(TUN_TMA_HINT) != 0;  // L2 eviction-priority hints on operand loads (default on)
  if (!use_hints) q.hint_a = q.hint_b = 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  CU_TRY(cudaLaunchKernelEx(&cfg, kern, a.m, b.m, a.lo, b.lo, q));
  return BP_OK;
}

// CTA-pair (cta_group::2) variant: 256 x PAIR_N pair tiles.  Pairs that can be co-resident (GPC granularity) per
// PAIR_N, filled once per device before the first kernel choice (rank_create / bp_debug_gemm).
static int g_max_clusters[2] = {0, 0};
inline int& max_clusters_slot(int pair_n) { return g_max_clusters[pair_n == 256]; }

// Occupancy depends on the cluster size and the shared-memory size only, so one instantiation per PAIR_N answers for all.
template <int PAIR_N>
static void query_cluster_capacity(int num_sms) {
  auto kern = bp_gemm2_kernel<true, false, EPI_FWD_HID, PAIR_N>;
  int cap = num_sms / 2;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm2_smem_bytes<PAIR_N>()) ==
      cudaSuccess) {
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(num_sms / 2 * 2);
    qc.blockDim = dim3(GEMM_THREADS);
    qc.dynamicSmemBytes = gemm2_smem_bytes<PAIR_N>();
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    qc.attrs = qa;
    qc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) == cudaSuccess && n > 0) cap = std::min(n, cap);
  }
  cudaGetLastError();
  max_clusters_slot(PAIR_N) = cap;
}
void init_cluster_capacity(int num_sms) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (g_max_clusters[0] > 0) return;
  query_cluster_capacity<256>(num_sms);
  query_cluster_capacity<128>(num_sms);
  if (getenv("BP_VERBOSE"))
    fprintf(stderr, "libbpgpu: co-resident CTA pairs  128-wide: %d   256-wide: %d\n", g_max_clusters[0],
            g_max_clusters[1]);
}

template <bool kAMN, bool kBMN, int kEpi, int PAIR_N>
static int launch_gemm2(cudaStream_t st, int num_sms, const MapPair& a, const MapPair& b, const GemmParams& p) {
  auto kern = bp_gemm2_kernel<kAMN, kBMN, kEpi, PAIR_N>;
  constexpr size_t smem = gemm2_smem_bytes<PAIR_N>();
  static thread_local int configured_dev = -1;
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_dev = dev;
  }
  init_cluster_capacity(num_sms);
  const int max_clusters = max_clusters_slot(PAIR_N);
  const int mt = (p.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  const int nt = (p.N - p.n_begin + PAIR_N - 1) / PAIR_N;
  const int tiles = mt * nt;
  if (tiles <= 0 || p.K <= 0) return fail(BP_EINVAL, "gemm2: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  const bool use_pdl = tunable(TUN_PDL) != 0;
  GemmParams q = p;
  if (tunable(TUN_TMA_HINT) == 0) q.hint_a = q.hint_b = 0;  // L2 eviction-priority hints on operand loads (default on)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(std::min(tiles, std::min(max_clusters, num_sms / 2)) * 2);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 2 : 1;
  CU_TRY(cudaLaunchKernelEx(&cfg, kern, a.m, b.m, a.lo, b.lo, q));
  return BP_OK;
}

// Kernel choice per product: pair_n = 0 -> 128 x 128 tiles on lone CTAs.
//   BP_PAIRS: 0 = lone CTAs only, 1 = automatic (default), 2 = 256-wide pairs always, 3 = 128-wide pairs whenever the
//             narrow B map exists.
// Automatic: 256 x 256 pair tiles if they fill >= 60 % of the SM pairs, else 256 x 128 pair tiles under the same
// condition, else 128 x 128 tiles on lone CTAs.  Measured in isolation on the B200 (profiles/r1d): 2048x1024x2048 fwd
// 18.4 us on 128-wide pairs vs 25.9 us on 256-wide ones (64 CTAs); 2048x2049x1024 dW 17.6 us on 256-wide pairs vs 19.5 us
// on 128-wide ones — the rule picks the faster one in every product of C2/C3/C5.
inline int pick_kernel(const GemmParams& p, int num_sms, bool have_b64) {
  const int mode = tunable(TUN_PAIRS);
  if (mode == 0) return 0;
  if (mode == 2) return 256;
  if (mode == 3) return have_b64 ? 128 : 0;
  const int mt = (p.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  const int n = p.N - p.n_begin;
  if (mt * ((n + 255) / 256) * 10 >= (num_sms / 2) * 6) return 256;
  if (have_b64 && mt * ((n + 127) / 128) * 10 >= (num_sms / 2) * 6) return 128;
  return 0;
}

// b = B operand map with 128-wide boxes (lone CTAs and 256-wide pairs); b64 = the same operand with 64-wide boxes
// (128-wide pairs: each CTA stages 64 B columns), or null.
template <bool kAMN, bool kBMN, int kEpi>
static int launch_gemm(cudaStream_t st, int num_sms, const AMaps& a, const MapPair& b, const GemmParams& p,
                       const MapPair* b64 = nullptr) {
  const int pair_n = pick_kernel(p, num_sms, b64 != nullptr);
  if (pair_n == 256) return launch_gemm2<kAMN, kBMN, kEpi, 256>(st, num_sms, a.full, b, p);
  if (pair_n == 128) return launch_gemm2<kAMN, kBMN, kEpi, 128>(st, num_sms, a.full, *b64, p);
  return launch_gemm_bn<kAMN, kBMN, kEpi, kBlockN>(st, num_sms, a.full, b, p);
}

// ------------------------------------------------------------------------------------------------ dispatch
int launch_product(Product prod, cudaStream_t st, int num_sms, const AMaps& a, const MapPair& b, const GemmParams& p,
                   const MapPair* b64) {
  switch (prod) {
    case PROD_FWD_HID: return launch_gemm<true, false, EPI_FWD_HID>(st, num_sms, a, b, p, b64);
    case PROD_FWD_OUT: return launch_gemm<true, false, EPI_FWD_OUT>(st, num_sms, a, b, p, b64);
    case PROD_FWD_PLAIN: return launch_gemm<true, false, EPI_PLAIN>(st, num_sms, a, b, p, b64);
    case PROD_FWD_SPLITK: return launch_gemm_bn<true, false, EPI_PLAIN, kBlockN>(st, num_sms, a.full, b, p);
    case PROD_DX: return launch_gemm<false, false, EPI_DX>(st, num_sms, a, b, p, b64);
    case PROD_DW: return launch_gemm<true, true, EPI_PLAIN>(st, num_sms, a, b, p, b64);
  }
  return fail(BP_EINVAL, "launch_product: unknown product %d", (int)prod);
}

// ------------------------------------------------------------------------------------------------ chained products
// CTA pairs of bp_chain_kernel that can be co-resident on this device (the schedule must not use more: its pairs wait
// for each other inside the launch).
int chain_max_pairs(int num_sms) {
  static std::mutex mu;
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev] > 0) return cached[dev];
  int cap = 0;
  if (cudaFuncSetAttribute(bp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem_bytes()) ==
      cudaSuccess) {
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(num_sms / 2 * 2);
    qc.blockDim = dim3(GEMM_THREADS);
    qc.dynamicSmemBytes = chain_smem_bytes();
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = 2;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    qc.attrs = qa;
    qc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, bp_chain_kernel, &qc) == cudaSuccess && n > 0) cap = std::min(n, num_sms / 2);
  }
  cudaGetLastError();
  cached[dev] = cap;
  return cap;
}

int launch_chain(cudaStream_t st, const ChainArgs& a, int pairs) {
  if (pairs <= 0) return fail(BP_EINVAL, "launch_chain: no pairs");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = chain_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU_TRY(cudaLaunchKernelEx(&cfg, bp_chain_kernel, a));
  return BP_OK;
}

int launch_out_finish(cudaStream_t st, const float* ws, long long ldw, const GemmParams& p, long long cells) {
  bp_out_finish_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(ws, ldw, p);
  CU_TRY(cudaGetLastError());
  return BP_OK;
}

}  // namespace bp

extern "C" {
using namespace bp;

// Host-only: which kernel launch_product would pick for an M x N x K product and what its launch looks like — the
// selection rule of pick_kernel made inspectable without a GPU (tests/test_launch_plan.py pins the plans of the
// BASELINE configs).  max_pairs = co-resident CTA pairs to assume (<= 0: num_sms / 2).
int bp_debug_plan(int M, int N, int K, int have_b64, int num_sms, int max_pairs, int* pair_n, int* tiles, int* ctas,
                  int* k_blocks) {
  if (M <= 0 || N <= 0 || K <= 0 || num_sms <= 0 || !pair_n || !tiles || !ctas || !k_blocks)
    return fail(BP_EINVAL, "bp_debug_plan: bad argument");
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.K = K;
  const int pn = pick_kernel(p, num_sms, have_b64 != 0);
  *pair_n = pn;
  *k_blocks = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  if (pn == 0) {
    *tiles = ((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * ((N + kBlockN - 1) / kBlockN);
    *ctas = std::min(*tiles, num_sms);
  } else {
    const int pairs = max_pairs > 0 ? std::min(max_pairs, num_sms / 2) : num_sms / 2;
    *tiles = ((M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M)) * ((N + pn - 1) / pn);
    *ctas = 2 * std::min(*tiles, pairs);
  }
  return BP_OK;
}

// Bring-up aid: cycles per 128 x bn x 8 TF32 MMA (issue, issue+drain) on one SM; combo 0 = A MN/B K, 1 = K/K, 2 = MN/MN,
// 3 = K/MN.  mode bit 0: fence before each group of 8, bit 1: commit after each group.
int bp_debug_mma_rate(int combo, int bn, int iters, int mode, double* cyc_issue, double* cyc_total) {
  long long* d = nullptr;
  long long h[2] = {0, 0};
  int rc = [&]() -> int {
    CU_TRY(cudaMalloc(&d, 16));
    const size_t smem = (128 + 256) * 64 * 4 + 1024;
#define BP_RATE(AM, BM, BN_)                                                                         \
  {                                                                                                  \
    auto k = bp_mma_rate_kernel<AM, BM, BN_>;                                                        \
    CU_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    k<<<1, 64, smem>>>(iters, mode, d);                                                              \
  }
    if (bn == 128) {
      if (combo == 0) BP_RATE(true, false, 128) else if (combo == 1) BP_RATE(false, false, 128)
      else if (combo == 2) BP_RATE(true, true, 128) else BP_RATE(false, true, 128)
    } else if (bn == 256) {
      if (combo == 0) BP_RATE(true, false, 256) else if (combo == 1) BP_RATE(false, false, 256)
      else if (combo == 2) BP_RATE(true, true, 256) else BP_RATE(false, true, 256)
    } else if (bn == 64) {
      if (combo == 0) BP_RATE(true, false, 64) else if (combo == 1) BP_RATE(false, false, 64)
      else if (combo == 2) BP_RATE(true, true, 64) else BP_RATE(false, true, 64)
    } else {
      return fail(BP_EINVAL, "bn must be 64, 128 or 256");
    }
#undef BP_RATE
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    return BP_OK;
  }();
  cudaFree(d);
  if (cyc_issue) *cyc_issue = (double)h[0] / (8.0 * iters);
  if (cyc_total) *cyc_total = (double)h[1] / (8.0 * iters);
  return rc;
}

// Bring-up aid: aggregate GB/s of the epilogue store pattern into a buffer of `mb` megabytes (larger than L2 so that
// every line is a write miss with a dirty victim), `reps` passes.  layout 0: tiles of a row-major [rows x ld] matrix
// (128 row segments of 512 B, ld*4 bytes apart); layout 1: every tile one contiguous 64 KB region.
int bp_debug_store_pattern(int layout, int ld, int mb, int reps, int ctas, double* gbs) {
  if (!gbs || ld < 128 || ld % 128 != 0 || mb <= 0 || reps <= 0 || ctas <= 0) return fail(BP_EINVAL, "bad argument");
  float* buf = nullptr;
  const long long floats = (long long)mb * 1024 * 1024 / 4;
  const long long rows = floats / ld / 128 * 128;      // whole 128-row panels
  const long long n_tiles = rows / 128 * (ld / 128);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = [&]() -> int {
    CU_TRY(cudaMalloc(&buf, floats * 4));
    CU_TRY(cudaMemset(buf, 0, floats * 4));
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    for (int rep = 0; rep <= reps; ++rep) {
      if (rep == 1) CU_TRY(cudaEventRecord(e0));
      if (layout == 0)  // panel = 128 rows of the matrix; tiles of a panel 128 floats apart
        bp_store_pattern_kernel<<<ctas, 128>>>(buf, ld, 128, ld / 128, 128LL * ld, n_tiles, (float)rep);
      else              // same tiles, each contiguous
        bp_store_pattern_kernel<<<ctas, 128>>>(buf, 128, 128 * 128, ld / 128, 128LL * ld, n_tiles, (float)rep);
    }
    CU_TRY(cudaEventRecord(e1));
    CU_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *gbs = (double)n_tiles * 65536.0 * reps / (ms * 1e-3) / 1e9;
    return BP_OK;
  }();
  cudaFree(buf);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return rc;
}

}  // extern "C"
