// libbpgpu.so — C-ABI implementation (see include/bp_gpu.h).  Host side of the B200-native trainer object that
// replaces the reference's BP_GPU class (BP_GPU.h:40-88, BP_GPU.cu).  One `Rank` == one GPU == one NCCL rank.
//
// Data layout in HBM (all fp32):
//   parameter arena   per weight layer l (fan-in K, fan-out N):  (K+1) rows x ldN floats, row k<K = w[k*N + out]
//                     (the reference's own index convention, BP_GPU.cu:188), row K = bias; ldN = roundup(N,32).
//                     Three identical arenas: weights, momentum deltas, gradients (flat, one SGD launch, one
//                     all-reduce region per layer).
//   chunk buffers     x: rows x ldx (ldx = roundup(K0+1,32)), column K0 holds 1.0f;  targets: rows x Nout packed.
//                     Double-buffered so the next chunk's H2D overlaps the current chunk's compute.
//   activations       Y_l: bunch x ldy_l (ldy = roundup(N_l+1,32)), column N_l holds 1.0f (bias-gradient trick);
//                     D_l (= dE/dX_l): bunch x ldd_l (ldd = roundup(N_l,32)).
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bp_gpu.h"
#include "bp_elementwise.cuh"
#include "bp_internal.h"
#include "bp_peer.cuh"
#include "bp_splice.cuh"

namespace bp {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

}  // namespace bp

namespace {

using namespace bp;

// ------------------------------------------------------------------------------------------------ NCCL (lazy dlopen)
struct NcclId128 { char b[128]; };  // layout of ncclUniqueId (passed by value to ncclCommInitRank)
struct NcclApi {
  using Id128 = NcclId128;
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

int load_nccl() {
  std::call_once(g_nccl_once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return;
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.lib, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(g_nccl.lib, "ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(g_nccl.lib, "ncclAllGather"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.lib, "ncclCommDestroy"));
    g_nccl.GetErrorString =
        reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.lib, "ncclGetErrorString"));
  });
  if (!g_nccl.lib || !g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(BP_ECOMM, "libnccl.so.2 not loadable: %s", dlerror() ? dlerror() : "missing symbols");
  return BP_OK;
}
#define NCCL_TRY(expr)                                                                                  \
  do {                                                                                                  \
    int r__ = (expr);                                                                                   \
    if (r__ != 0)                                                                                       \
      return fail(BP_ECOMM, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
  } while (0)
constexpr int kNcclFloat32 = 7;  // ncclFloat32
constexpr int kNcclSum = 0;      // ncclSum
constexpr int kNcclInt8 = 0;     // ncclInt8 / ncclChar

// ------------------------------------------------------------------------------------------------ per-rank state
struct LayerState {
  int K = 0, N = 0;        // fan-in, fan-out
  long long ldN = 0;       // arena row stride
  long long off = 0;       // arena offset (floats) of row 0
  long long size = 0;      // (K+1)*ldN rounded up to 64 floats
  float* y = nullptr;      // activation output Y_l (hidden layers only): rows x ldy
  float* y_lo = nullptr;   // split-precision low part (3xTF32 mode only)
  long long ldy = 0;
  float* d = nullptr;      // dE/dX_l: rows x ldd
  float* d_lo = nullptr;
  long long ldd = 0;
  // tensor maps that do not depend on the chunk
  AMaps w_fwd;         // W^T as MN-major A: {N, K}
  AMaps w_dx;          // W as K-major A:    {N, K}
  MapPair yprev_fwd;   // Y_{l-1} as K-major B (l >= 2): {K, rows}, box {32,kBlockN}
  MapPair yprev_fwd64; // same, 64-row boxes (128-wide CTA pairs)
  MapPair d_dx64;      // D_l as K-major B, 64-row boxes
  MapPair yprev_dw;    // Y_{l-1}^T as MN-major B (l >= 2): {K+1, bunch}, box {32,32}
  MapPair d_dx;        // D_l as K-major B: {N, bunch}, box {32,kBlockN}
  AMaps d_dw;          // D_l^T as MN-major A: {N, bunch}
};

struct ChunkBuf {
  float* x = nullptr;
  float* x_lo = nullptr;   // split-precision low part of x (3xTF32 mode only)
  float* t = nullptr;
  long long cap_rows = 0;
  int rows = 0;            // rows currently resident
  bool has_targ = false;
  int masked_rows = 0;     // rows [0, masked_rows) have had the input-dropout mask applied IN PLACE (BP_GPU.cu:536-540
                           // does the same to its device copy): they can be trained once per upload
  cudaEvent_t uploaded = nullptr;  // recorded on the copy stream after the H2D
  cudaEvent_t split_done = nullptr;  // 3xTF32: x_lo has been derived from x
  cudaEvent_t consumed = nullptr;  // recorded on the compute stream after the last kernel that reads the buffer
  bool consumed_valid = false;
};

constexpr int kPeerFlagGroups = 2;  // counter groups of the peer-memory exchange (Rank::flags)

struct Rank {
  bp_config cfg{};
  int L = 0;               // weight layers
  int local_bunch = 0;
  int num_sms = 0;
  cudaStream_t compute = nullptr, copy = nullptr, comm_stream = nullptr;
  int dw_bn = 128;              // N-tile of the weight-gradient GEMM
  int ar_slices = 1;            // data-parallel: slices per layer gradient (BP_AR_SLICES; >1 measured slower: many
                                // small all-reduces are latency-bound)
  int comm_sms = 0;             // data-parallel: SMs the persistent GEMMs leave free so that NCCL's kernels can run
                                // next to them instead of behind them (BP_COMM_SMS; 16/32 measured no better: the
                                // exposed all-reduce time is that of the last, largest layers' gradients)
  cudaStream_t side = nullptr;  // weight-gradient GEMMs run here, concurrently with the dX chain on `compute`
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_grad = nullptr, ev_comm = nullptr, ev_side = nullptr;
  cudaEvent_t ev_upper = nullptr, ev_comm_upper = nullptr;  // dW / all-reduce of layers >= 2 complete
  cudaEvent_t ev_d[BP_MAXLAYER] = {};  // ev_d[l]: dE/dX_l is complete (recorded on `compute`)
  float *w = nullptr, *dw = nullptr, *g = nullptr;
  float* w_lo = nullptr;   // split-precision low part of the weights (3xTF32 mode only)
  int passes = 1;          // 1 = TF32, 3 = 3xTF32
  long long arena_floats = 0;
  LayerState layer[BP_MAXLAYER];
  long long ldx = 0;
  ChunkBuf chunk[2];
  int cur = 0;             // chunk buffer the resident calls operate on
  float* out_dev = nullptr;
  long long out_cap_rows = 0;
  double* sqerr_dev = nullptr;
  // pipelined decode (bp_decode_raw_submit / _wait): two output slots, results copied out on their own stream so that
  // the D2H of chunk k runs beside the forward pass of chunk k+1
  struct DecodeSlot {
    float* out_dev = nullptr;
    long long cap_rows = 0;
    cudaEvent_t fwd_done = nullptr, out_done = nullptr;
  };
  DecodeSlot dec[2];
  int dec_head = 0, dec_inflight = 0;
  cudaStream_t d2h = nullptr;
  // device staging of a raw chunk (bp_upload_raw_chunk): Pfile records, sample table, norm vectors
  uint32_t* raw_fea = nullptr;
  uint32_t* raw_targ = nullptr;
  int* raw_tab = nullptr;
  float* raw_norm = nullptr;
  size_t raw_fea_cap = 0, raw_targ_cap = 0, raw_tab_cap = 0, raw_norm_cap = 0;  // in elements
  float* splitk_ws = nullptr;  // kMaxSplits partial planes of the output-layer product (bunchsize x ldN_out each)
  void* nccl_comm = nullptr;
  // peer-memory data parallelism (bp_peer.cuh); dp_p2p = 1 once every rank has mapped every other rank's slabs
  struct PeerMem {
    float* w = nullptr;
    float* w_lo = nullptr;
    float* recv = nullptr;
    unsigned long long* flags = nullptr;
    bool ipc = false;  // pointers came from cudaIpcOpenMemHandle (close them on destroy)
  };
  int dp_p2p = 0;
  float* recv = nullptr;                 // world_size slabs of arena_floats: recv + src*arena_floats
  unsigned long long* flags = nullptr;   // kPeerFlagGroups x kMaxPeers counters, group g at flags + g*kMaxPeers:
                                         // 0 gradients of step s landed from src, 1 weights landed from src
  PeerMem peer[kMaxPeers];
  unsigned long long dp_step = 0;
  PeerLayers peer_layers{};
  int chunk_base[BP_MAXLAYER] = {};      // first 32-row ownership chunk of each layer
  uint64_t launches = 0, bunches = 0;
  uint32_t step = 0;
  bool profiling = false;
  static constexpr int kProfCap = 64;        // profiled bunches kept (ring)
  std::vector<cudaEvent_t> pev;              // 7 events per profiled bunch, no host sync while recording
  uint64_t prof_cnt = 0;
  // Launch timeline (bp_set_profiling(h, 2)): one event behind EVERY launch of a bunch, on the stream it was launched
  // on, for the first kTlBunches profiled bunches; bp_get_timeline reports each launch's completion time since the
  // bunch's start mark.  The records sit between the kernels, so programmatic dependent launch does not overlap
  // across them: this is the dependency chain's latency, launch by launch, with warm-in-place caches (what neither the
  // per-class events nor ncu's serialised, cache-flushed replays show).
  static constexpr int kTlCap = 40, kTlBunches = 16;
  int timeline = 0;
  std::vector<cudaEvent_t> tev;              // kTlBunches x kTlCap, created on first use
  const char* tl_label[kTlCap] = {};
  int tl_idx = 0, tl_marks = 0;
  uint64_t tl_bunches = 0;
  // per-bunch sum of squared output error of the two most recent train calls (ring), so that a caller can read
  // call i-1's losses while call i computes (no compute-stream sync on the reading path)
  double* loss_dev[2] = {nullptr, nullptr};
  int loss_cap[2] = {0, 0}, loss_n[2] = {0, 0};
  cudaEvent_t loss_done[2] = {nullptr, nullptr};
  int loss_cur = 0;
  SgdBiasRanges bias_ranges{};
  // Chained products (bp_chain.cuh): the forward products of a train bunch in ONE launch, the dX chain + every dW
  // product in a second one.  Built lazily (after the peer-memory slabs are known), rebuilt when an option changes.
  struct ChainPlan {
    int n_prods = 0, n_items = 0;
    std::vector<ChainProd> h_prods;   // the launch description is assembled from these per bunch (kernel parameters)
    long long makespan = 0;      // the scheduler's estimate, SM cycles
    std::vector<ChainItem> h_items;   // host copies for bp_debug_chain_trace
    std::vector<int> h_pair_off;
    int pair_n_l1 = 128;         // tile width of the product whose B operand is this bunch's input rows
    int max_pair_n = 128;        // widest tile of the plan: decides the ring geometry (4 x 48 KB or 3 x 64 KB)
  };
  static constexpr int kChainCounters = 1024;
  int use_chain = -1;            // -1 automatic: chained launches for data-parallel ranks on peer memory (measured
                                 // +6 % at 8 GPUs on C2, +13 % on C4, and the exchange overlaps the next forward
                                 // launch), one launch per product on a single GPU (C2 5-9 % faster that way, C3
                                 // equal, C4's net 3 % slower: profiles/r2d-f); 0 / 1 force (BP_CHAIN, option "chain")
  bool chain_built = false;
  int chain_pairs = 0;
  ChainPlan chain_fwd, chain_bwd;
  CUtensorMap* chain_maps = nullptr;
  uint32_t* chain_counters = nullptr;   // 2 sets of kChainCounters
  unsigned long long* chain_trace = nullptr;  // bp_set_option("chain_trace", 1): stamps of the most recent launches
  int chain_trace_on = 0;
  int chain_set = 0;

  int gemm_sms() const { return nccl_comm ? std::max(8, num_sms - comm_sms) : num_sms; }
  int Nout() const { return cfg.layersizes[L]; }
  int K0() const { return cfg.layersizes[0]; }
};

int ensure_chunk(Rank* r, ChunkBuf& c, long long rows) {
  if (rows <= c.cap_rows) return BP_OK;
  if (c.x) {
    CU_TRY(cudaStreamSynchronize(r->compute));
    CU_TRY(cudaStreamSynchronize(r->copy));
    CU_TRY(cudaFree(c.x));
    CU_TRY(cudaFree(c.t));
    if (c.x_lo) CU_TRY(cudaFree(c.x_lo));
    c.x = c.t = c.x_lo = nullptr;
  }
  const long long cap = std::max<long long>(rows, r->cfg.bunchsize);
  CU_TRY(cudaMalloc(&c.x, sizeof(float) * cap * r->ldx));
  CU_TRY(cudaMalloc(&c.t, sizeof(float) * cap * r->Nout()));
  CU_TRY(cudaMemsetAsync(c.x, 0, sizeof(float) * cap * r->ldx, r->copy));
  if (r->passes == 3) {
    CU_TRY(cudaMalloc(&c.x_lo, sizeof(float) * cap * r->ldx));
    CU_TRY(cudaMemsetAsync(c.x_lo, 0, sizeof(float) * cap * r->ldx, r->copy));
  }
  bp_fill_col_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, r->copy>>>(c.x, r->ldx, cap, r->K0(), 1.0f);
  CU_TRY(cudaGetLastError());
  r->launches++;
  c.cap_rows = cap;
  c.rows = 0;
  c.consumed_valid = false;
  return BP_OK;
}

void rank_p2p_release(Rank* r);
void chain_release(Rank* r);

int rank_destroy(Rank* r) {
  if (!r) return BP_OK;
  cudaSetDevice(r->cfg.device);
  cudaDeviceSynchronize();
  if (r->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(r->nccl_comm);
  rank_p2p_release(r);
  chain_release(r);
  cudaFree(r->chain_counters);
  cudaFree(r->chain_trace);
  for (auto& c : r->chunk) {
    cudaFree(c.x);
    cudaFree(c.x_lo);
    cudaFree(c.t);
    if (c.uploaded) cudaEventDestroy(c.uploaded);
    if (c.split_done) cudaEventDestroy(c.split_done);
    if (c.consumed) cudaEventDestroy(c.consumed);
  }
  for (int l = 1; l <= r->L; ++l) {
    cudaFree(r->layer[l].y);
    cudaFree(r->layer[l].d);
    cudaFree(r->layer[l].y_lo);
    cudaFree(r->layer[l].d_lo);
  }
  cudaFree(r->w);
  cudaFree(r->w_lo);
  cudaFree(r->dw);
  cudaFree(r->g);
  cudaFree(r->out_dev);
  cudaFree(r->sqerr_dev);
  cudaFree(r->splitk_ws);
  cudaFree(r->raw_fea);
  cudaFree(r->raw_targ);
  cudaFree(r->raw_tab);
  cudaFree(r->raw_norm);
  for (auto e : {r->ev_t0, r->ev_t1, r->ev_grad, r->ev_comm, r->ev_side, r->ev_upper, r->ev_comm_upper})
    if (e) cudaEventDestroy(e);
  for (auto e : r->ev_d)
    if (e) cudaEventDestroy(e);
  for (auto e : r->pev)
    if (e) cudaEventDestroy(e);
  for (auto e : r->tev)
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    cudaFree(r->loss_dev[i]);
    if (r->loss_done[i]) cudaEventDestroy(r->loss_done[i]);
  }
  for (auto& ds : r->dec) {
    cudaFree(ds.out_dev);
    if (ds.fwd_done) cudaEventDestroy(ds.fwd_done);
    if (ds.out_done) cudaEventDestroy(ds.out_done);
  }
  for (auto s : {r->compute, r->copy, r->comm_stream, r->side, r->d2h})
    if (s) cudaStreamDestroy(s);
  delete r;
  return BP_OK;
}

int upload_params(Rank* r, float* const* weights, float* const* bias) {
  for (int l = 1; l <= r->L; ++l) {
    LayerState& ls = r->layer[l];
    CU_TRY(cudaMemcpy2DAsync(r->w + ls.off, ls.ldN * 4, weights[l], size_t(ls.N) * 4, size_t(ls.N) * 4, ls.K,
                             cudaMemcpyHostToDevice, r->compute));
    CU_TRY(cudaMemcpyAsync(r->w + ls.off + (long long)ls.K * ls.ldN, bias[l], size_t(ls.N) * 4,
                           cudaMemcpyHostToDevice, r->compute));
  }
  if (r->passes == 3) {
    bp_split_lo_kernel<<<r->num_sms * 8, 256, 0, r->compute>>>((const float4*)r->w, (float4*)r->w_lo,
                                                               r->arena_floats / 4);
    CU_TRY(cudaGetLastError());
    r->launches++;
  }
  CU_TRY(cudaStreamSynchronize(r->compute));
  return BP_OK;
}

// CUDA loads kernels lazily, at their first launch, and a load may have to wait for the kernels that are running — a
// first launch beside a kernel that waits for it never starts (measured: scripts/probe/coreside_probe.cu,
// profiles/r2g_coreside_probe_lazy_loading.log).  Nothing on the default path depends on that today, but the kernels of
// the data-parallel exchange and the update are loaded up front so that no schedule change can run into it.
int preload_runtime_kernels() {
  cudaFuncAttributes fa;
  CU_TRY(cudaFuncGetAttributes(&fa, bp_peer_signal_kernel));
  CU_TRY(cudaFuncGetAttributes(&fa, bp_peer_wait_kernel));
  CU_TRY(cudaFuncGetAttributes(&fa, bp_peer_sgd_kernel<false>));
  CU_TRY(cudaFuncGetAttributes(&fa, bp_peer_sgd_kernel<true>));
  CU_TRY(cudaFuncGetAttributes(&fa, bp_sgd_kernel<false>));
  CU_TRY(cudaFuncGetAttributes(&fa, bp_sgd_kernel<true>));
  CU_TRY(cudaFuncGetAttributes(&fa, bp_input_dropout_kernel));
  return BP_OK;
}

int rank_create(Rank** out, const bp_config* cfg, float* const* weights, float* const* bias) {
  if (!cfg || !weights || !bias) return fail(BP_EINVAL, "bp_create: null argument");
  if (cfg->numlayers < 2 || cfg->numlayers > BP_MAXLAYER)
    return fail(BP_EINVAL, "bp_create: numlayers=%d out of range [2,%d]", cfg->numlayers, BP_MAXLAYER);
  for (int i = 0; i < cfg->numlayers; ++i)
    if (cfg->layersizes[i] <= 0) return fail(BP_EINVAL, "bp_create: layersizes[%d]=%d", i, cfg->layersizes[i]);
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return fail(BP_EINVAL, "bp_create: rank %d / world %d", cfg->rank, cfg->world_size);
  if (cfg->bunchsize <= 0 || cfg->bunchsize % cfg->world_size != 0)
    return fail(BP_EINVAL, "bp_create: bunchsize %d must be a positive multiple of world_size %d", cfg->bunchsize,
                cfg->world_size);
  if ((cfg->bunchsize / cfg->world_size) % 4 != 0 && cfg->dropoutflag == 1)
    return fail(BP_EINVAL, "bp_create: per-rank bunch (%d) must be a multiple of 4 when dropout is on",
                cfg->bunchsize / cfg->world_size);
  if (cfg->math_mode != BP_MATH_TF32 && cfg->math_mode != BP_MATH_3XTF32)
    return fail(BP_EINVAL, "bp_create: math_mode %d", cfg->math_mode);
  if (cfg->activation != BP_ACT_RELU && cfg->activation != BP_ACT_SIGMOID)
    return fail(BP_EINVAL, "bp_create: activation %d", cfg->activation);

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(BP_ENODEV, "no CUDA device (this library has no CPU fallback)");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(BP_ENODEV, "device %d of %d", cfg->device, ndev);
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(BP_ENODEV, "device %d is sm_%d%d; libbpgpu is built for sm_100a only", cfg->device, prop.major,
                prop.minor);
  CU_TRY(cudaSetDevice(cfg->device));
  BP_TRY(preload_runtime_kernels());

  Rank* r = new Rank();
  r->cfg = *cfg;
  r->L = cfg->numlayers - 1;
  r->local_bunch = cfg->bunchsize / cfg->world_size;
  r->num_sms = prop.multiProcessorCount;
  r->passes = cfg->math_mode == BP_MATH_3XTF32 ? 3 : 1;
  if (const char* e = getenv("BP_AR_SLICES")) r->ar_slices = std::max(1, atoi(e));
  if (const char* e = getenv("BP_COMM_SMS")) r->comm_sms = std::max(0, std::min(64, atoi(e)));
  if (const char* e = getenv("BP_CHAIN")) r->use_chain = atoi(e) < 0 ? -1 : atoi(e) != 0;
  int rc = [&]() -> int {
    CU_TRY(cudaStreamCreateWithFlags(&r->compute, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&r->copy, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&r->comm_stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&r->side, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&r->ev_side, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&r->ev_upper, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&r->ev_comm_upper, cudaEventDisableTiming));
    for (auto& e : r->ev_d) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU_TRY(cudaEventCreate(&r->ev_t0));
    CU_TRY(cudaEventCreate(&r->ev_t1));
    CU_TRY(cudaEventCreateWithFlags(&r->ev_grad, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&r->ev_comm, cudaEventDisableTiming));
    for (auto& c : r->chunk) {
      CU_TRY(cudaEventCreateWithFlags(&c.uploaded, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&c.split_done, cudaEventDisableTiming));
      CU_TRY(cudaEventCreateWithFlags(&c.consumed, cudaEventDisableTiming));
    }
    for (auto& e : r->loss_done) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    r->pev.assign(7 * Rank::kProfCap, nullptr);
    for (auto& e : r->pev) CU_TRY(cudaEventCreate(&e));

    long long off = 0;
    r->bias_ranges.n = 0;
    for (int l = 1; l <= r->L; ++l) {
      LayerState& ls = r->layer[l];
      ls.K = cfg->layersizes[l - 1];
      ls.N = cfg->layersizes[l];
      ls.ldN = round_up(ls.N, 32);
      ls.off = off;
      ls.size = round_up((long long)(ls.K + 1) * ls.ldN, 64);
      r->bias_ranges.begin4[r->bias_ranges.n] = (off + (long long)ls.K * ls.ldN) / 4;
      r->bias_ranges.end4[r->bias_ranges.n] = (off + (long long)(ls.K + 1) * ls.ldN) / 4;
      r->bias_ranges.n++;
      {  // ownership geometry for the peer-memory exchange
        PeerLayers& pl = r->peer_layers;
        const LayerState& lp = r->layer[l - 1];
        const int prev_rows = l == 1 ? 0 : (int)((lp.size + lp.ldN - 1) / lp.ldN);  // rows incl. the 64-float padding
        r->chunk_base[l] = l == 1 ? 0 : r->chunk_base[l - 1] + (prev_rows + 31) / 32;
        pl.begin4[pl.n] = off / 4;
        pl.end4[pl.n] = (off + ls.size) / 4;
        pl.row4[pl.n] = (int)(ls.ldN / 4);
        pl.bias_row[pl.n] = ls.K;
        pl.chunk_base[pl.n] = r->chunk_base[l];
        pl.n++;
      }
      off += ls.size;
    }
    r->arena_floats = off;
    CU_TRY(cudaMalloc(&r->w, off * 4));
    CU_TRY(cudaMalloc(&r->dw, off * 4));
    CU_TRY(cudaMalloc(&r->g, off * 4));
    if (r->passes == 3) {
      CU_TRY(cudaMalloc(&r->w_lo, off * 4));
      CU_TRY(cudaMemsetAsync(r->w_lo, 0, off * 4, r->compute));
    }
    CU_TRY(cudaMemsetAsync(r->w, 0, off * 4, r->compute));
    CU_TRY(cudaMemsetAsync(r->dw, 0, off * 4, r->compute));  // deltas start at zero every run (BP_GPU.cu:137-138,938)
    CU_TRY(cudaMemsetAsync(r->g, 0, off * 4, r->compute));
    CU_TRY(cudaMalloc(&r->sqerr_dev, sizeof(double)));
    CU_TRY(cudaMalloc(&r->splitk_ws, sizeof(float) * kMaxSplits * (size_t)cfg->bunchsize * r->layer[r->L].ldN));

    r->ldx = round_up(r->K0() + 1, 32);
    const long long rows = cfg->bunchsize;  // activation buffers hold a full (global) bunch so CV can use it
    for (int l = 1; l <= r->L; ++l) {
      LayerState& ls = r->layer[l];
      ls.ldd = round_up(ls.N, 32);
      CU_TRY(cudaMalloc(&ls.d, rows * ls.ldd * 4));
      CU_TRY(cudaMemsetAsync(ls.d, 0, rows * ls.ldd * 4, r->compute));
      if (r->passes == 3) {
        CU_TRY(cudaMalloc(&ls.d_lo, rows * ls.ldd * 4));
        CU_TRY(cudaMemsetAsync(ls.d_lo, 0, rows * ls.ldd * 4, r->compute));
      }
      if (l < r->L) {
        ls.ldy = round_up(ls.N + 1, 32);
        CU_TRY(cudaMalloc(&ls.y, rows * ls.ldy * 4));
        CU_TRY(cudaMemsetAsync(ls.y, 0, rows * ls.ldy * 4, r->compute));
        if (r->passes == 3) {  // the ones column's low part is 0: 1.0f truncates exactly
          CU_TRY(cudaMalloc(&ls.y_lo, rows * ls.ldy * 4));
          CU_TRY(cudaMemsetAsync(ls.y_lo, 0, rows * ls.ldy * 4, r->compute));
        }
        bp_fill_col_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, r->compute>>>(ls.y, ls.ldy, rows, ls.N, 1.0f);
        CU_TRY(cudaGetLastError());
        r->launches++;
      }
    }
    for (int l = 1; l <= r->L; ++l) {
      LayerState& ls = r->layer[l];
      const float* wl = r->w + ls.off;
      const float* wlo = r->w_lo ? r->w_lo + ls.off : nullptr;
      BP_TRY(make_a_maps(&ls.w_fwd, wl, wlo, ls.N, ls.K, ls.ldN, true));
      BP_TRY(make_a_maps(&ls.w_dx, wl, wlo, ls.N, ls.K, ls.ldN, false));
      BP_TRY(make_map(&ls.d_dx, ls.d, ls.d_lo, ls.N, r->local_bunch, ls.ldd, kBlockN, false));
      BP_TRY(make_map(&ls.d_dx64, ls.d, ls.d_lo, ls.N, r->local_bunch, ls.ldd, 64, false));
      BP_TRY(make_a_maps(&ls.d_dw, ls.d, ls.d_lo, ls.N, r->local_bunch, ls.ldd, true));
      if (l >= 2) {
        LayerState& lp = r->layer[l - 1];
        BP_TRY(make_map(&ls.yprev_fwd, lp.y, lp.y_lo, ls.K, rows, lp.ldy, kBlockN, false));
        BP_TRY(make_map(&ls.yprev_fwd64, lp.y, lp.y_lo, ls.K, rows, lp.ldy, 64, false));
        BP_TRY(make_map(&ls.yprev_dw, lp.y, lp.y_lo, ls.K + 1, r->local_bunch, lp.ldy, r->dw_bn, true));
      }
    }
    BP_TRY(upload_params(r, weights, bias));
    return BP_OK;
  }();
  if (rc != BP_OK) {
    std::string keep = g_err;
    rank_destroy(r);
    g_err = keep;
    return rc;
  }
  *out = r;
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ chunk upload
// wait_host = false: the caller queues the chunk's bunches first and calls rank_wait_uploaded() before it returns to
// its own caller, so the kernel launches are issued while the DMA runs instead of after it.
int rank_upload(Rank* r, int n_frames, const float* in, const float* targ, bool wait_host = true) {
  if (n_frames <= 0 || !in) return fail(BP_EINVAL, "upload: n_frames=%d in=%p", n_frames, (const void*)in);
  CU_TRY(cudaSetDevice(r->cfg.device));
  const int nb = r->cur ^ 1;
  ChunkBuf& c = r->chunk[nb];
  BP_TRY(ensure_chunk(r, c, n_frames));
  if (c.consumed_valid) CU_TRY(cudaStreamWaitEvent(r->copy, c.consumed, 0));
  const size_t rowb = size_t(r->K0()) * 4;
  CU_TRY(cudaMemcpy2DAsync(c.x, r->ldx * 4, in, rowb, rowb, n_frames, cudaMemcpyHostToDevice, r->copy));
  if (targ)
    CU_TRY(cudaMemcpyAsync(c.t, targ, size_t(n_frames) * r->Nout() * 4, cudaMemcpyHostToDevice, r->copy));
  CU_TRY(cudaEventRecord(c.uploaded, r->copy));  // host buffers are consumed here
  if (r->passes == 3) {
    bp_split_lo_kernel<<<r->num_sms * 8, 256, 0, r->copy>>>((const float4*)c.x, (float4*)c.x_lo,
                                                            (long long)n_frames * r->ldx / 4);
    CU_TRY(cudaGetLastError());
    r->launches++;
    CU_TRY(cudaEventRecord(c.split_done, r->copy));
  }
  c.rows = n_frames;
  c.masked_rows = 0;
  c.has_targ = targ != nullptr;
  r->cur = nb;
  CU_TRY(cudaStreamWaitEvent(r->compute, r->passes == 3 ? c.split_done : c.uploaded, 0));
  // The caller may overwrite its buffers as soon as we return (BP_GPU::train semantics): wait for the DMA only.
  if (wait_host) CU_TRY(cudaEventSynchronize(c.uploaded));
  return BP_OK;
}

// Host wait for the H2D copies of the current chunk (the one rank_upload / rank_upload_raw filled last).
int rank_wait_uploaded(Rank* r) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  CU_TRY(cudaEventSynchronize(r->chunk[r->cur].uploaded));
  return BP_OK;
}

// Raw-chunk upload: the records travel as they lie in the Pfile, the samples are assembled by bp_splice_kernel.
template <typename T>
int grow_dev(Rank* r, T** ptr, size_t* cap, size_t need) {
  if (need <= *cap) return BP_OK;
  CU_TRY(cudaStreamSynchronize(r->copy));  // the previous splice may still be reading the old block
  if (*ptr) CU_TRY(cudaFree(*ptr));
  *ptr = nullptr;
  *cap = 0;
  CU_TRY(cudaMalloc(ptr, sizeof(T) * need));
  *cap = need;
  return BP_OK;
}

int rank_upload_raw(Rank* r, const bp_raw_chunk* rc, bool all_rows, bool wait_host = true) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  // all_rows: this rank assembles every sample (cross-validation / decode run on one device, BP_GPU.cu:440-441)
  const int G = all_rows ? 1 : r->cfg.world_size, B = r->cfg.bunchsize, lb = r->local_bunch;
  const int full_rows = (rc->n_samples / B) * B;
  const int rows = G == 1 ? rc->n_samples : (rc->n_samples / B) * lb;
  if (rows <= 0) return fail(BP_EINVAL, "upload_raw: chunk of %d samples has no full bunch of %d", rc->n_samples, B);
  const int nbuf = r->cur ^ 1;
  ChunkBuf& c = r->chunk[nbuf];
  BP_TRY(ensure_chunk(r, c, rows));
  const int out_dim = r->Nout();
  const size_t fea_words = (size_t)rc->n_records * (rc->fea_dim + 2);
  const size_t targ_words = rc->targ_records ? (size_t)rc->n_records * (out_dim + 2) : 0;
  BP_TRY(grow_dev(r, &r->raw_fea, &r->raw_fea_cap, fea_words));
  if (targ_words) BP_TRY(grow_dev(r, &r->raw_targ, &r->raw_targ_cap, targ_words));
  BP_TRY(grow_dev(r, &r->raw_tab, &r->raw_tab_cap, (size_t)3 * rc->n_samples));
  BP_TRY(grow_dev(r, &r->raw_norm, &r->raw_norm_cap, (size_t)2 * rc->fea_dim));
  if (c.consumed_valid) CU_TRY(cudaStreamWaitEvent(r->copy, c.consumed, 0));
  const size_t ns = (size_t)rc->n_samples;
  CU_TRY(cudaMemcpyAsync(r->raw_fea, rc->fea_records, fea_words * 4, cudaMemcpyHostToDevice, r->copy));
  if (targ_words)
    CU_TRY(cudaMemcpyAsync(r->raw_targ, rc->targ_records, targ_words * 4, cudaMemcpyHostToDevice, r->copy));
  CU_TRY(cudaMemcpyAsync(r->raw_tab, rc->sample_frame, ns * 4, cudaMemcpyHostToDevice, r->copy));
  if (rc->nat) CU_TRY(cudaMemcpyAsync(r->raw_tab + ns, rc->sample_seg, ns * 4, cudaMemcpyHostToDevice, r->copy));
  if (rc->sample_row)
    CU_TRY(cudaMemcpyAsync(r->raw_tab + 2 * ns, rc->sample_row, ns * 4, cudaMemcpyHostToDevice, r->copy));
  CU_TRY(cudaMemcpyAsync(r->raw_norm, rc->mean, (size_t)rc->fea_dim * 4, cudaMemcpyHostToDevice, r->copy));
  CU_TRY(cudaMemcpyAsync(r->raw_norm + rc->fea_dim, rc->inv_std, (size_t)rc->fea_dim * 4, cudaMemcpyHostToDevice,
                         r->copy));
  CU_TRY(cudaEventRecord(c.uploaded, r->copy));  // host buffers are consumed here
  SpliceParams sp{};
  sp.fea = r->raw_fea;
  sp.targ = targ_words ? r->raw_targ : nullptr;
  sp.mean = r->raw_norm;
  sp.inv_std = r->raw_norm + rc->fea_dim;
  sp.sample_frame = r->raw_tab;
  sp.sample_seg = r->raw_tab + ns;
  sp.sample_row = rc->sample_row ? r->raw_tab + 2 * ns : nullptr;
  sp.dim = rc->fea_dim;
  sp.ctx = rc->fea_context;
  sp.nat = rc->nat ? 1 : 0;
  sp.targ_offset = rc->targ_offset;
  sp.out_dim = out_dim;
  sp.n_records = rc->n_records;
  sp.n_samples = rc->n_samples;
  sp.x = c.x;
  sp.ldx = r->ldx;
  sp.t = c.t;
  sp.world = G;
  sp.rank = r->cfg.rank;
  sp.B = B;
  sp.lb = lb;
  sp.full_rows = full_rows;
  const int grid = std::min(rc->n_samples, r->num_sms * 16);
  bp_splice_kernel<<<grid, 256, 0, r->copy>>>(sp);
  CU_TRY(cudaGetLastError());
  r->launches++;
  if (r->passes == 3) {
    bp_split_lo_kernel<<<r->num_sms * 8, 256, 0, r->copy>>>((const float4*)c.x, (float4*)c.x_lo,
                                                            (long long)rows * r->ldx / 4);
    CU_TRY(cudaGetLastError());
    r->launches++;
  }
  CU_TRY(cudaEventRecord(c.split_done, r->copy));  // "rows assembled"
  c.rows = rows;
  c.masked_rows = 0;
  c.has_targ = targ_words != 0;
  r->cur = nbuf;
  CU_TRY(cudaStreamWaitEvent(r->compute, c.split_done, 0));
  if (wait_host) CU_TRY(cudaEventSynchronize(c.uploaded));  // the caller may reuse its buffers (BP_GPU::train semantics)
  return BP_OK;
}

// Number of K slices for the output-layer product (1 = no split).  BP_SPLITK: 0 = never, N>1 = force N, unset = auto:
// as many slices as idle SMs allow, each at least 4 k-blocks deep, at most kMaxSplits.
inline int out_layer_splits(const GemmParams& p, int num_sms) {
  const int mode = tunable(TUN_SPLITK);
  if (mode == 0) return 1;
  const int tiles = ((p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * ((p.N + kBlockN - 1) / kBlockN);
  const int num_kb = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  int s = mode > 1 ? mode : std::min(num_sms / std::max(1, tiles), num_kb / 4);
  s = std::max(1, std::min({s, kMaxSplits, num_kb}));
  const int per = (num_kb + s - 1) / s;
  return (num_kb + per - 1) / per;  // every slice non-empty
}

// Launch timeline mark: an event behind the launch just issued on `st` (see Rank::timeline).
inline void tl_mark(Rank* r, cudaStream_t st, const char* label) {
  if (!r->timeline || r->tl_bunches >= (uint64_t)Rank::kTlBunches || r->tl_idx >= Rank::kTlCap) return;
  if (r->tev.empty()) {
    r->tev.assign((size_t)Rank::kTlBunches * Rank::kTlCap, nullptr);
    for (auto& e : r->tev)
      if (cudaEventCreate(&e) != cudaSuccess) {
        e = nullptr;
        cudaGetLastError();
      }
  }
  cudaEvent_t e = r->tev[(size_t)r->tl_bunches * Rank::kTlCap + r->tl_idx];
  if (!e) return;
  cudaEventRecord(e, st);
  r->tl_label[r->tl_idx++] = label;
}
static const char* const kFwdLabel[BP_MAXLAYER] = {"", "fwd1", "fwd2", "fwd3", "fwd4", "fwd5", "fwd6", "fwd7", "fwd8", "fwd9"};
static const char* const kDxLabel[BP_MAXLAYER] = {"", "dx1", "dx2", "dx3", "dx4", "dx5", "dx6", "dx7", "dx8", "dx9"};
static const char* const kDwLabel[BP_MAXLAYER] = {"", "dw1 (side)", "dw2 (side)", "dw3 (side)", "dw4 (side)", "dw5 (side)",
                                                  "dw6 (side)", "dw7 (side)", "dw8 (side)", "dw9 (side)"};

// kernDropout on the device copy of the input bunch (BP_GPU.cu:536-540): rows [f0, f0+n) of the resident chunk, in place.
int input_dropout(Rank* r, ChunkBuf& c, int f0, int n) {
  const bp_config& cf = r->cfg;
  if (cf.dropoutflag != 1 || !(cf.visible_omit > 0.0f)) return BP_OK;
  const uint32_t seed_lo = (uint32_t)cf.seed, seed_hi = (uint32_t)(cf.seed >> 32);
  const int frame0 = cf.rank * r->local_bunch;
  dim3 grid((r->K0() + 255) / 256, (n + 3) / 4);
  bp_input_dropout_kernel<<<grid, 256, 0, r->compute>>>(c.x + (long long)f0 * r->ldx, r->ldx, n, r->K0(),
                                                        cf.visible_omit, seed_lo, seed_hi, r->step, frame0);
  CU_TRY(cudaGetLastError());
  r->launches++;
  tl_mark(r, r->compute, "input dropout");
  if (r->passes == 3) {  // the same mask on the low part
    bp_input_dropout_kernel<<<grid, 256, 0, r->compute>>>(c.x_lo + (long long)f0 * r->ldx, r->ldx, n, r->K0(),
                                                          cf.visible_omit, seed_lo, seed_hi, r->step, frame0);
    CU_TRY(cudaGetLastError());
    r->launches++;
  }
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ forward
// Forward over rows [f0, f0+n) of the resident chunk.  train=true: masks + D_L; train=false: keep-scaling (CV).
int forward_rows(Rank* r, ChunkBuf& c, int f0, int n, bool train, float* out2, long long ldo2, double* sqerr) {
  // train: sqerr (if non-null) receives this bunch's sum (out - targ)^2 (training-loss monitor, our extension)
  const bp_config& cf = r->cfg;
  const bool drop = cf.dropoutflag == 1;
  const uint32_t seed_lo = (uint32_t)cf.seed, seed_hi = (uint32_t)(cf.seed >> 32);
  const int frame0 = cf.rank * r->local_bunch;
  float* xb = c.x + (long long)f0 * r->ldx;

  if (train) BP_TRY(input_dropout(r, c, f0, n));
  for (int l = 1; l <= r->L; ++l) {
    LayerState& ls = r->layer[l];
    GemmParams p{};
    p.M = ls.N;
    p.N = n;
    p.K = ls.K;
    p.bias = r->w + ls.off + (long long)ls.K * ls.ldN;
    p.act = cf.activation;
    p.scale = 1.0f;
    if (!train && drop) p.scale = 1.0f - (l == 1 ? cf.visible_omit : cf.hid_omit);
    p.seed_lo = seed_lo;
    p.seed_hi = seed_hi;
    p.step = r->step;
    p.layer = (uint32_t)l;
    p.frame0 = frame0;
    p.passes = r->passes;
    p.hint_a = kEvictLast;  // A = the weights: re-read by dX and by the next bunch, keep them in L2
    MapPair xmap, xmap64;
    const MapPair* bmap = &ls.yprev_fwd;
    const MapPair* bmap64 = &ls.yprev_fwd64;
    if (l == 1) {
      const float* xlo = c.x_lo ? c.x_lo + (long long)f0 * r->ldx : nullptr;
      BP_TRY(make_map(&xmap, xb, xlo, ls.K, n, r->ldx, kBlockN, false));
      BP_TRY(make_map(&xmap64, xb, xlo, ls.K, n, r->ldx, 64, false));
      bmap = &xmap;
      bmap64 = &xmap64;
    }
    if (l < r->L) {
      p.out = ls.y;
      p.out_lo = ls.y_lo;
      p.ldo = ls.ldy;
      p.drop_p = (train && drop) ? cf.hid_omit : 0.0f;
      BP_TRY((launch_product(PROD_FWD_HID, r->compute, r->gemm_sms(), ls.w_fwd, *bmap, p, bmap64)));
      if (train) tl_mark(r, r->compute, kFwdLabel[l]);
    } else {
      if (train) {
        p.out = ls.d;
        p.out_lo = ls.d_lo;
        p.ldo = ls.ldd;
        p.aux = c.t + (long long)f0 * ls.N;
        p.ldaux = ls.N;
        p.gscale = 2.0f / (float)cf.bunchsize;  // (2.0f/rows), rows = GLOBAL bunch (DevFunc.cu:263)
        p.sqerr = sqerr;
      } else {
        p.out = nullptr;
        p.out2 = out2;
        p.ldo2 = ldo2;
        if (sqerr) {
          p.aux = c.t + (long long)f0 * ls.N;
          p.ldaux = ls.N;
          p.sqerr = sqerr;
        }
      }
      // The 257-wide output layer makes only ceil(257/128) x ceil(n/128) tiles (24 at bunch 1024) with a long K loop:
      // cut K into slices so the product covers the machine, then add the slices and apply the epilogue in a finisher.
      const int splits = r->splitk_ws ? out_layer_splits(p, r->gemm_sms()) : 1;
      if (splits > 1) {
        GemmParams g = p;
        g.k_splits = splits;
        g.split_stride = (long long)n * ls.ldN;
        g.out = r->splitk_ws;
        g.out_lo = nullptr;
        g.ldo = ls.ldN;
        BP_TRY(launch_product(PROD_FWD_SPLITK, r->compute, r->gemm_sms(), ls.w_fwd, *bmap, g));
        r->launches++;
        if (train) tl_mark(r, r->compute, "out split-K");
        p.k_splits = splits;
        p.split_stride = g.split_stride;
        const long long cells = (long long)n * ls.ldN;
        BP_TRY(launch_out_finish(r->compute, r->splitk_ws, ls.ldN, p, cells));
        if (train) tl_mark(r, r->compute, "out finish");
      } else {
        BP_TRY((launch_product(PROD_FWD_OUT, r->compute, r->gemm_sms(), ls.w_fwd, *bmap, p, bmap64)));
        if (train) tl_mark(r, r->compute, "out");
      }
    }
    r->launches++;
  }
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ peer-memory DP
// BP_DP: "nccl" = all-reduce + replicated update; "p2p" = peer-memory exchange or fail; unset = p2p when every peer
// can be mapped, else nccl (with a note on stderr).
inline int dp_mode_wanted() {
  const char* e = getenv("BP_DP");
  if (!e) return 1;
  if (strcmp(e, "nccl") == 0) return 0;
  return 2;
}

int rank_p2p_alloc(Rank* r) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  const size_t slab = sizeof(float) * (size_t)r->arena_floats * r->cfg.world_size;
  CU_TRY(cudaMalloc(&r->recv, slab));
  CU_TRY(cudaMemset(r->recv, 0, slab));
  CU_TRY(cudaMalloc(&r->flags, sizeof(unsigned long long) * kPeerFlagGroups * kMaxPeers));
  CU_TRY(cudaMemset(r->flags, 0, sizeof(unsigned long long) * kPeerFlagGroups * kMaxPeers));
  CU_TRY(cudaDeviceSynchronize());
  Rank::PeerMem& me = r->peer[r->cfg.rank];
  me.w = r->w;
  me.w_lo = r->w_lo;
  me.recv = r->recv;
  me.flags = r->flags;
  return BP_OK;
}

void rank_p2p_release(Rank* r) {
  for (auto& pm : r->peer) {
    if (pm.ipc)
      for (void* q : {(void*)pm.w, (void*)pm.w_lo, (void*)pm.recv, (void*)pm.flags})
        if (q) cudaIpcCloseMemHandle(q);
    pm = Rank::PeerMem{};
  }
  cudaFree(r->recv);
  cudaFree(r->flags);
  r->recv = nullptr;
  r->flags = nullptr;
  r->dp_p2p = 0;
  cudaGetLastError();
}

// One process per rank: the slabs are exported as CUDA IPC handles, exchanged through the NCCL communicator that
// exists anyway, and mapped with peer access.  Every rank takes the same decision (two agreement rounds), so either
// all ranks use the peer-memory exchange or all use NCCL.
struct IpcPack {
  cudaIpcMemHandle_t w, w_lo, recv, flags;
  int has_lo, ok, pad0, pad1;
};

int rank_p2p_connect_ipc(Rank* r) {
  const int want = dp_mode_wanted();
  if (want == 0) return BP_OK;
  const int W = r->cfg.world_size, me = r->cfg.rank;
  CU_TRY(cudaSetDevice(r->cfg.device));
  IpcPack mine{};
  int ok = (W <= kMaxPeers) && g_nccl.AllGather != nullptr;
  if (ok && rank_p2p_alloc(r) != BP_OK) ok = 0;
  if (ok) {
    ok = cudaIpcGetMemHandle(&mine.w, r->w) == cudaSuccess && cudaIpcGetMemHandle(&mine.recv, r->recv) == cudaSuccess &&
         cudaIpcGetMemHandle(&mine.flags, r->flags) == cudaSuccess &&
         (!r->w_lo || cudaIpcGetMemHandle(&mine.w_lo, r->w_lo) == cudaSuccess);
    cudaGetLastError();
  }
  mine.ok = ok;
  mine.has_lo = r->w_lo != nullptr;
  IpcPack* dev = nullptr;
  CU_TRY(cudaMalloc(&dev, sizeof(IpcPack) * (W + 1)));
  std::vector<IpcPack> all(W);
  CU_TRY(cudaMemcpyAsync(dev + W, &mine, sizeof(IpcPack), cudaMemcpyHostToDevice, r->compute));
  NCCL_TRY(g_nccl.AllGather(dev + W, dev, sizeof(IpcPack), kNcclInt8, r->nccl_comm, r->compute));
  CU_TRY(cudaMemcpyAsync(all.data(), dev, sizeof(IpcPack) * W, cudaMemcpyDeviceToHost, r->compute));
  CU_TRY(cudaStreamSynchronize(r->compute));
  bool all_ok = true;
  for (const IpcPack& q : all) all_ok = all_ok && q.ok;
  float bad = all_ok ? 0.0f : 1.0f;
  if (all_ok) {
    for (int p = 0; p < W && bad == 0.0f; ++p) {
      if (p == me) continue;
      Rank::PeerMem& pm = r->peer[p];
      pm.ipc = true;
      auto open = [&](void** dst, const cudaIpcMemHandle_t& hnd) {
        if (cudaIpcOpenMemHandle(dst, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          *dst = nullptr;
          bad = 1.0f;
          cudaGetLastError();
        }
      };
      open((void**)&pm.w, all[p].w);
      open((void**)&pm.recv, all[p].recv);
      open((void**)&pm.flags, all[p].flags);
      if (all[p].has_lo) open((void**)&pm.w_lo, all[p].w_lo);
    }
  }
  // second round: did everybody manage to map everybody?
  float* dbad = reinterpret_cast<float*>(dev);
  CU_TRY(cudaMemcpyAsync(dbad, &bad, sizeof(float), cudaMemcpyHostToDevice, r->compute));
  NCCL_TRY(g_nccl.AllReduce(dbad, dbad, 1, kNcclFloat32, kNcclSum, r->nccl_comm, r->compute));
  CU_TRY(cudaMemcpyAsync(&bad, dbad, sizeof(float), cudaMemcpyDeviceToHost, r->compute));
  CU_TRY(cudaStreamSynchronize(r->compute));
  CU_TRY(cudaFree(dev));
  r->chain_built = false;  // the dW products' scatter targets are the peers' slabs
  if (bad == 0.0f) {
    r->dp_p2p = 1;
    if (me == 0 && getenv("BP_VERBOSE")) fprintf(stderr, "libbpgpu: gradient exchange over peer memory, %d ranks\n", W);
    return BP_OK;
  }
  rank_p2p_release(r);
  if (want == 2) return fail(BP_ECOMM, "BP_DP=p2p: peer memory of some rank cannot be mapped (CUDA IPC / peer access)");
  if (me == 0) fprintf(stderr, "libbpgpu: peer memory not mappable on every rank, using NCCL all-reduce\n");
  return BP_OK;
}

// One exchange after the local gradient GEMMs of a bunch: publish "my partial gradients have landed everywhere",
// reduce + update the rows this rank owns and push them to every replica, publish "my rows have landed", wait for
// everybody else's rows.  All on the compute stream; the spin waits are bounded (bp_peer.cuh).
int peer_exchange(Rank* r) {
  const bp_config& cf = r->cfg;
  const int W = cf.world_size, me = cf.rank;
  const unsigned long long step = ++r->dp_step;
  PeerFlags fg{}, fw{};
  PeerArenas pa{};
  for (int p = 0; p < W; ++p) {
    fg.slot[p] = r->peer[p].flags + me;
    fw.slot[p] = r->peer[p].flags + kMaxPeers + me;
    pa.w[p] = (float4*)r->peer[p].w;
    pa.w_lo[p] = (float4*)r->peer[p].w_lo;
  }
  bp_peer_signal_kernel<<<1, 32, 0, r->compute>>>(fg, W, step);
  const float nf = (float)cf.bunchsize;
  const float c1 = (1 - cf.momentum) * cf.lrate;
  const int grid = r->num_sms * 8;
  if (cf.weightcost != 0.0f)
    bp_peer_sgd_kernel<true><<<grid, 256, 0, r->compute>>>((float4*)r->dw, (const float4*)r->recv,
                                                           r->arena_floats / 4, pa, r->peer_layers, W, me, nf,
                                                           cf.momentum, c1, cf.weightcost, r->flags, step);
  else
    bp_peer_sgd_kernel<false><<<grid, 256, 0, r->compute>>>((float4*)r->dw, (const float4*)r->recv,
                                                            r->arena_floats / 4, pa, r->peer_layers, W, me, nf,
                                                            cf.momentum, c1, 0.0f, r->flags, step);
  bp_peer_signal_kernel<<<1, 32, 0, r->compute>>>(fw, W, step);
  bp_peer_wait_kernel<<<1, 32, 0, r->compute>>>(r->flags + kMaxPeers, W, step);
  CU_TRY(cudaGetLastError());
  r->launches += 4;
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ chained products
void chain_release(Rank* r) {
  for (Rank::ChainPlan* pl : {&r->chain_fwd, &r->chain_bwd}) {
    *pl = Rank::ChainPlan{};
  }
  cudaFree(r->chain_maps);
  r->chain_maps = nullptr;
  r->chain_built = false;
}

int chain_upload_plan(Rank* r, Rank::ChainPlan& pl, std::vector<ChainProd>& prods, std::vector<ChainShape>& shapes,
                      int which) {
  {  // the tables built from the device state must be the ones the host-only planner (and its CPU tests) describe
    const std::vector<ChainShape> want =
        chain_shapes(r->cfg.layersizes, r->L, r->local_bunch, r->chain_pairs, r->passes, which);
    bool same = want.size() == shapes.size();
    for (size_t i = 0; same && i < want.size(); ++i)
      same = want[i].m_tiles == shapes[i].m_tiles && want[i].n_tiles == shapes[i].n_tiles &&
             want[i].n_cols == shapes[i].n_cols && want[i].kb == shapes[i].kb && want[i].pair_n == shapes[i].pair_n &&
             want[i].dep_prod == shapes[i].dep_prod && want[i].dep_all == shapes[i].dep_all &&
             want[i].prio == shapes[i].prio;
    if (!same) return fail(BP_EINVAL, "chain: product table differs from chain_shapes() (which = %d)", which);
  }
  pl.max_pair_n = 128;
  for (const ChainProd& q : prods) {
    if (q.dep_prod >= 0) prods[q.dep_prod].has_consumer = 1;
    pl.max_pair_n = std::max(pl.max_pair_n, q.pair_n);
  }
  std::vector<ChainItem> items;
  std::vector<int> pair_off;
  pl.makespan = chain_schedule(shapes, r->chain_pairs, items, pair_off);
  if (pl.makespan < 0) return fail(BP_EINVAL, "chain: dependency cycle in the product table");
  pl.n_prods = (int)prods.size();
  pl.n_items = (int)items.size();
  if (pl.n_items > CHAIN_MAX_ITEMS || r->chain_pairs > CHAIN_MAX_PAIRS || pl.n_prods > CHAIN_MAX_PROD)
    return fail(BP_EINVAL, "chain: %d tiles / %d pairs / %d products exceed the launch description", pl.n_items,
                r->chain_pairs, pl.n_prods);
  pl.h_prods = prods;
  pl.h_items = items;
  pl.h_pair_off = pair_off;
  return BP_OK;
}

// Product tables and schedules of a train bunch of r->local_bunch rows.  The parameters mirror forward_rows /
// launch_dx / launch_dw below (the one-launch-per-product path), which stays as the A/B baseline and the fallback.
int chain_build(Rank* r) {
  chain_release(r);
  const bp_config& cf = r->cfg;
  const int n = r->local_bunch;
  r->chain_pairs = chain_max_pairs(r->num_sms);
  if (r->chain_pairs <= 0) return fail(BP_ECUDA, "chain: no co-resident CTA pairs");
  if (!r->chain_counters) {
    CU_TRY(cudaMalloc(&r->chain_counters, 2 * Rank::kChainCounters * sizeof(uint32_t)));
    CU_TRY(cudaMemset(r->chain_counters, 0, 2 * Rank::kChainCounters * sizeof(uint32_t)));
    r->chain_set = 0;
  }
  std::vector<CUtensorMap> maps;
  auto add = [&](const MapPair& mp, int* idx, int* idx_lo) {
    *idx = (int)maps.size();
    maps.push_back(mp.m);
    *idx_lo = (int)maps.size();
    maps.push_back(mp.lo);
  };
  const bool drop = cf.dropoutflag == 1;
  const bool hints = tunable(TUN_TMA_HINT) != 0;
  const int kb_mul = r->passes == 3 ? 3 : 1;
  int cnt = 0;
  auto shape_of = [&](const ChainProd& q, int dep, int dep_all, int prio) {
    ChainShape sh{};
    sh.m_tiles = q.m_tiles;
    sh.n_tiles = q.n_tiles;
    sh.n_cols = q.p.N;
    sh.kb = (q.p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K * kb_mul;
    sh.pair_n = q.pair_n;
    sh.dep_prod = dep;
    sh.dep_all = dep_all;
    sh.prio = prio;
    return sh;
  };
  auto finish_prod = [&](ChainProd& q) {
    q.m_tiles = (q.p.M + 255) / 256;
    q.n_tiles = (q.p.N + q.pair_n - 1) / q.pair_n;
    q.cnt_base = cnt;
    cnt += q.n_tiles;
  };
  // ---- forward: layer l = product l-1
  {
    std::vector<ChainProd> prods;
    std::vector<ChainShape> shapes;
    cnt = 0;
    for (int l = 1; l <= r->L; ++l) {
      LayerState& ls = r->layer[l];
      ChainProd q{};
      GemmParams& p = q.p;
      p.M = ls.N;
      p.N = n;
      p.K = ls.K;
      p.bias = r->w + ls.off + (long long)ls.K * ls.ldN;
      p.act = cf.activation;
      p.scale = 1.0f;
      p.seed_lo = (uint32_t)cf.seed;
      p.seed_hi = (uint32_t)(cf.seed >> 32);
      p.layer = (uint32_t)l;
      p.frame0 = cf.rank * r->local_bunch;
      p.passes = r->passes;
      p.hint_a = hints ? kEvictLast : 0;
      q.amn = 1;
      q.bmn = 0;
      q.pair_n = chain_pair_n(p.M, p.N, r->chain_pairs);
      add(ls.w_fwd.full, &q.map_a, &q.map_a_lo);
      if (l == 1) {
        q.map_b = -1;
        q.map_b_lo = -2;
        r->chain_fwd.pair_n_l1 = q.pair_n;
      } else {
        add(q.pair_n == 256 ? ls.yprev_fwd : ls.yprev_fwd64, &q.map_b, &q.map_b_lo);
      }
      q.dep_prod = l >= 2 ? l - 2 : -1;
      q.dep_all = 0;
      if (l < r->L) {
        q.epi = EPI_FWD_HID;
        p.out = ls.y;
        p.out_lo = ls.y_lo;
        p.ldo = ls.ldy;
        p.drop_p = drop ? cf.hid_omit : 0.0f;
      } else {
        q.epi = EPI_FWD_OUT;
        q.per_bunch = 1;  // targets + loss slot of the bunch
        p.out = ls.d;
        p.out_lo = ls.d_lo;
        p.ldo = ls.ldd;
        p.ldaux = ls.N;
        p.gscale = 2.0f / (float)cf.bunchsize;  // (2.0f/rows), rows = GLOBAL bunch (DevFunc.cu:263)
      }
      finish_prod(q);
      shapes.push_back(shape_of(q, q.dep_prod, 0, l));
      prods.push_back(q);
    }
    if (cnt > Rank::kChainCounters) return fail(BP_EINVAL, "chain: too many tiles per product row (%d counters)", cnt);
    const int pn1 = r->chain_fwd.pair_n_l1;
    BP_TRY(chain_upload_plan(r, r->chain_fwd, prods, shapes, 0));
    r->chain_fwd.pair_n_l1 = pn1;
  }
  // ---- back-propagation: dX_L .. dX_2 (the chain) and dW_L .. dW_1 (the filler)
  {
    std::vector<ChainProd> prods;
    std::vector<ChainShape> shapes;
    cnt = 0;
    int dx_index[BP_MAXLAYER + 1];
    for (auto& v : dx_index) v = -1;
    for (int l = r->L; l >= 2; --l) {  // dX_l: dE/dX_{l-1} = act'(Y_{l-1}) .* (dE/dX_l W_l^T)
      LayerState& ls = r->layer[l];
      LayerState& lp = r->layer[l - 1];
      ChainProd q{};
      GemmParams& p = q.p;
      p.M = ls.K;
      p.N = n;
      p.K = ls.N;
      p.out = lp.d;
      p.out_lo = lp.d_lo;
      p.ldo = lp.ldd;
      p.aux = lp.y;
      p.ldaux = lp.ldy;
      p.act = cf.activation;
      p.passes = r->passes;
      p.hint_a = hints ? kEvictLast : 0;
      q.epi = EPI_DX;
      q.amn = 0;
      q.bmn = 0;
      q.pair_n = chain_pair_n(p.M, p.N, r->chain_pairs);
      add(ls.w_dx.full, &q.map_a, &q.map_a_lo);
      add(q.pair_n == 256 ? ls.d_dx : ls.d_dx64, &q.map_b, &q.map_b_lo);
      q.dep_prod = l < r->L ? dx_index[l + 1] : -1;  // D_l comes from the product that back-propagated through l+1
      q.dep_all = 0;
      finish_prod(q);
      dx_index[l] = (int)prods.size();
      shapes.push_back(shape_of(q, q.dep_prod, 0, 0));
      prods.push_back(q);
    }
    for (int l = r->L; l >= 1; --l) {  // dW_l (+ bias gradient as row K of the block)
      LayerState& ls = r->layer[l];
      ChainProd q{};
      GemmParams& p = q.p;
      p.M = ls.N;
      p.N = ls.K + 1;
      p.K = n;
      p.out = r->g + ls.off;
      p.ldo = ls.ldN;
      p.passes = r->passes;
      if (r->dp_p2p) {  // reduce-scatter fused into the epilogue: chunks go straight to their owners' receive slabs
        p.scatter_n = cf.world_size;
        p.chunk_base = r->chunk_base[l];
        for (int o = 0; o < cf.world_size; ++o)
          p.scatter[o] = r->peer[o].recv + (long long)cf.rank * r->arena_floats + ls.off;
      }
      q.epi = EPI_PLAIN;
      q.amn = 1;
      q.bmn = 1;
      q.pair_n = 256;
      add(ls.d_dw.full, &q.map_a, &q.map_a_lo);
      if (l == 1) {
        q.map_b = -1;
        q.map_b_lo = -2;
      } else {
        add(ls.yprev_dw, &q.map_b, &q.map_b_lo);
      }
      q.dep_prod = l < r->L ? dx_index[l + 1] : -1;  // needs all of D_l
      q.dep_all = 1;
      finish_prod(q);
      shapes.push_back(shape_of(q, q.dep_prod, 1, 1 + (r->L - l)));
      prods.push_back(q);
    }
    if (cnt > Rank::kChainCounters) return fail(BP_EINVAL, "chain: too many tiles per product row (%d counters)", cnt);
    if ((int)prods.size() > CHAIN_MAX_PROD) return fail(BP_EINVAL, "chain: %d products", (int)prods.size());
    BP_TRY(chain_upload_plan(r, r->chain_bwd, prods, shapes, 1));
  }
  CU_TRY(cudaMalloc(&r->chain_maps, maps.size() * sizeof(CUtensorMap)));
  CU_TRY(cudaMemcpy(r->chain_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  r->chain_built = true;
  if (getenv("BP_VERBOSE"))
    fprintf(stderr, "libbpgpu: chained products on %d pairs: forward %d tiles (~%.1f us), back-propagation %d tiles "
            "(~%.1f us by the cost model)\n", r->chain_pairs, r->chain_fwd.n_items, r->chain_fwd.makespan / 1965.0,
            r->chain_bwd.n_items, r->chain_bwd.makespan / 1965.0);
  return BP_OK;
}

int chain_launch(Rank* r, Rank::ChainPlan& pl, const MapPair& dyn_b, const float* targ, double* sqerr) {
  static thread_local ChainArgs a;   // ~20 KB: keep it off the stack of deep call chains
  memset(&a, 0, sizeof a);
  a.n_prods = pl.n_prods;
  for (int i = 0; i < pl.n_prods; ++i) {
    a.prods[i] = pl.h_prods[i];
    a.prods[i].p.step = r->step;
    if (a.prods[i].per_bunch & 1) {
      a.prods[i].p.aux = targ;
      a.prods[i].p.sqerr = sqerr;
    }
  }
  memcpy(a.items, pl.h_items.data(), pl.h_items.size() * sizeof(ChainItem));
  memcpy(a.pair_off, pl.h_pair_off.data(), pl.h_pair_off.size() * sizeof(int));
  a.maps = r->chain_maps;
  a.dyn[0] = dyn_b.m;
  a.dyn[1] = dyn_b.lo;
  a.counters = r->chain_counters + (size_t)r->chain_set * Rank::kChainCounters;
  a.counters_next = r->chain_counters + (size_t)(r->chain_set ^ 1) * Rank::kChainCounters;
  a.n_counters = Rank::kChainCounters;
  a.stage_bytes = CHAIN_A_BYTES + (uint32_t)(pl.max_pair_n / 2) * GEMM_BLOCK_K * 4;
  a.n_stages = (int)(CHAIN_RING_BYTES / a.stage_bytes);
  if (r->chain_trace_on) {
    if (!r->chain_trace) CU_TRY(cudaMalloc(&r->chain_trace, 2 * 4096 * 4 * sizeof(unsigned long long)));
    a.trace = r->chain_trace + (&pl == &r->chain_bwd ? 4096 * 4 : 0);
  }
  BP_TRY(launch_chain(r->compute, a, r->chain_pairs));
  r->chain_set ^= 1;
  r->launches++;
  return BP_OK;
}

// Momentum-SGD update of arena floats [4*begin4, 4*end4) (kernUpdatedelta + kernAccSum, DevFunc.cu:313-318, 270-277).
int launch_sgd_range(Rank* r, long long begin4, long long end4, int blocks_per_sm) {
  const bp_config& cf = r->cfg;
  const float nf = (float)cf.bunchsize;  // `n` of kernUpdatedelta: int promoted to float
  const float c1 = (1 - cf.momentum) * cf.lrate;
  const int sgd_stream = tunable(TUN_SGD_STREAM);
  const int grid = r->num_sms * blocks_per_sm;
  if (cf.weightcost != 0.0f)
    bp_sgd_kernel<true><<<grid, 256, 0, r->compute>>>((float4*)r->dw, (float4*)r->w, (const float4*)r->g, begin4, end4,
                                                      nf, cf.momentum, c1, cf.weightcost, r->bias_ranges,
                                                      (float4*)r->w_lo, sgd_stream);
  else
    bp_sgd_kernel<false><<<grid, 256, 0, r->compute>>>((float4*)r->dw, (float4*)r->w, (const float4*)r->g, begin4, end4,
                                                       nf, cf.momentum, c1, 0.0f, r->bias_ranges, (float4*)r->w_lo,
                                                       sgd_stream);
  CU_TRY(cudaGetLastError());
  r->launches++;
  return BP_OK;
}

int peer_exchange(Rank* r);

// One train bunch as two chained launches (forward products; dX chain + dW products) and the update.
int train_bunch_chain(Rank* r, ChunkBuf& c, int f0, double* loss_slot) {
  const int n = r->local_bunch;
  if (!r->chain_built) BP_TRY(chain_build(r));
  const bool prof = r->profiling;
  int pe = 7 * (int)(r->prof_cnt % Rank::kProfCap);
  auto mark = [&]() { if (prof) cudaEventRecord(r->pev[pe++], r->compute); };
  mark();                                               // 0
  r->tl_idx = 0;
  tl_mark(r, r->compute, "start");
  BP_TRY(input_dropout(r, c, f0, n));
  float* xb = c.x + (long long)f0 * r->ldx;
  const float* xlo = c.x_lo ? c.x_lo + (long long)f0 * r->ldx : nullptr;
  MapPair xfwd, xdw;
  BP_TRY(make_map(&xfwd, xb, xlo, r->K0(), n, r->ldx, r->chain_fwd.pair_n_l1 / 2, false));
  BP_TRY(chain_launch(r, r->chain_fwd, xfwd, c.t + (long long)f0 * r->Nout(), loss_slot));
  tl_mark(r, r->compute, "forward chain");
  mark();                                               // 1: forward done
  BP_TRY(make_map(&xdw, xb, xlo, r->K0() + 1, n, r->ldx, r->dw_bn, true));
  BP_TRY(chain_launch(r, r->chain_bwd, xdw, nullptr, nullptr));
  tl_mark(r, r->compute, "back-propagation chain");
  mark();                                               // 2: dX chain and all dW done
  mark();                                               // 3
  mark();                                               // 4
  mark();                                               // 5
  if (r->dp_p2p) BP_TRY(peer_exchange(r));              // owner-side reduce + update + all-gather
  else BP_TRY(launch_sgd_range(r, 0, r->arena_floats / 4, 8));
  tl_mark(r, r->compute, "update / exchange, end of bunch");
  mark();                                               // 6
  r->step++;
  r->bunches++;
  if (prof) r->prof_cnt++;
  if (r->timeline && r->tl_bunches < (uint64_t)Rank::kTlBunches) {
    r->tl_marks = r->tl_idx;
    r->tl_bunches++;
  }
  return BP_OK;
}

// ------------------------------------------------------------------------------------------------ one train bunch
int train_bunch(Rank* r, ChunkBuf& c, int f0, double* loss_slot) {
  const bp_config& cf = r->cfg;
  const bool want_chain = r->use_chain == 1 || (r->use_chain < 0 && r->dp_p2p);
  if (want_chain && (!r->nccl_comm || r->dp_p2p))
    return train_bunch_chain(r, c, f0, loss_slot);
  const int n = r->local_bunch;
  const bool prof = r->profiling;
  int pe = 7 * (int)(r->prof_cnt % Rank::kProfCap);
  auto mark = [&]() { if (prof) cudaEventRecord(r->pev[pe++], r->compute); };
  mark();                                               // 0
  r->tl_idx = 0;
  tl_mark(r, r->compute, "start");
  BP_TRY(forward_rows(r, c, f0, n, true, nullptr, 0, loss_slot));
  mark();                                               // 1: fwd done
  float* xb = c.x + (long long)f0 * r->ldx;
  // The weight(+bias)-gradient GEMM of layer l needs only dE/dX_l and Y_{l-1}; the dX chain needs only W (the deferred
  // update has not touched it yet).  So dW_l is launched on the side stream as soon as dE/dX_l exists and runs
  // concurrently with the rest of the dX chain, filling the SMs a 128-tile GEMM leaves idle on a 148-SM part.
  // Last layer first, so its all-reduce starts earliest.
  auto launch_dw = [&](int l, cudaEvent_t after) -> int {
    LayerState& ls = r->layer[l];
    CU_TRY(cudaStreamWaitEvent(r->side, after, 0));
    GemmParams p{};
    p.M = ls.N;
    p.N = ls.K + 1;  // + the all-ones column -> row K of the gradient block = bias gradient
    p.K = n;
    p.out = r->g + ls.off;
    p.ldo = ls.ldN;
    p.passes = r->passes;
    MapPair xmap;
    const MapPair* bmap = &ls.yprev_dw;
    if (l == 1) {
      BP_TRY(make_map(&xmap, xb, c.x_lo ? c.x_lo + (long long)f0 * r->ldx : nullptr, ls.K + 1, n, r->ldx, r->dw_bn,
                      true));
      bmap = &xmap;
    }
    // Data-parallel ranks cut the gradient block into row slices (= column slices of the product, contiguous in the
    // arena) so that each slice's all-reduce runs while the next slice is still being computed; only the last
    // slice's reduction is exposed.  A single rank computes the block in one launch.
    if (r->dp_p2p) {  // reduce-scatter fused into the epilogue: chunks go straight to their owners' receive slabs
      p.scatter_n = cf.world_size;
      p.chunk_base = r->chunk_base[l];
      for (int o = 0; o < cf.world_size; ++o)
        p.scatter[o] = r->peer[o].recv + (long long)cf.rank * r->arena_floats + ls.off;
    }
    const int total = p.N;
    int slices = 1;
    if (r->nccl_comm && !r->dp_p2p) slices = std::max(1, std::min(r->ar_slices, (total + kBlockN - 1) / kBlockN));
    const int per = ((total + slices - 1) / slices + kBlockN - 1) / kBlockN * kBlockN;
    for (int b0 = 0; b0 < total; b0 += per) {
      p.n_begin = b0;
      p.N = std::min(total, b0 + per);
      BP_TRY((launch_product(PROD_DW, r->side, r->gemm_sms(), ls.d_dw, *bmap, p)));
      r->launches++;
      tl_mark(r, r->side, kDwLabel[l]);
      if (r->nccl_comm && !r->dp_p2p) {
        float* gs = r->g + ls.off + (long long)b0 * ls.ldN;
        const size_t cnt = (p.N == total) ? (size_t)(ls.size - (long long)b0 * ls.ldN)
                                          : (size_t)(p.N - b0) * (size_t)ls.ldN;
        CU_TRY(cudaEventRecord(r->ev_grad, r->side));
        CU_TRY(cudaStreamWaitEvent(r->comm_stream, r->ev_grad, 0));
        NCCL_TRY(g_nccl.AllReduce(gs, gs, cnt, kNcclFloat32, kNcclSum, r->nccl_comm, r->comm_stream));
      }
    }
    return BP_OK;
  };
  auto mark_upper = [&]() -> int {  // everything the layers >= 2 need for their update has been issued
    CU_TRY(cudaEventRecord(r->ev_upper, r->side));
    if (r->nccl_comm && !r->dp_p2p) CU_TRY(cudaEventRecord(r->ev_comm_upper, r->comm_stream));
    return BP_OK;
  };
  auto launch_dx = [&](int l) -> int {  // dE/dX_{l-1} = act'(Y_{l-1}) .* (dE/dX_l W_l^T), then ev_d[l-1]
    LayerState& ls = r->layer[l];
    LayerState& lp = r->layer[l - 1];
    GemmParams p{};
    p.M = ls.K;  // units of layer l-1
    p.N = n;
    p.K = ls.N;
    p.out = lp.d;
    p.out_lo = lp.d_lo;
    p.ldo = lp.ldd;
    p.aux = lp.y;
    p.ldaux = lp.ldy;
    p.act = cf.activation;
    p.passes = r->passes;
    p.hint_a = kEvictLast;  // A = the weights
    BP_TRY((launch_product(PROD_DX, r->compute, r->gemm_sms(), ls.w_dx, ls.d_dx, p, &ls.d_dx64)));
    r->launches++;
    tl_mark(r, r->compute, kDxLabel[l - 1]);
    CU_TRY(cudaEventRecord(r->ev_d[l - 1], r->compute));
    return BP_OK;
  };
  CU_TRY(cudaEventRecord(r->ev_d[r->L], r->compute));  // dE/dX_L comes out of the forward's last epilogue
  BP_TRY(launch_dw(r->L, r->ev_d[r->L]));
  if (r->L == 2) BP_TRY(mark_upper());
  for (int l = r->L; l >= 2; --l) {
    BP_TRY(launch_dx(l));
    if (l - 1 == 1 && r->L > 2) BP_TRY(mark_upper());  // before dW_1 enters the side stream
    BP_TRY(launch_dw(l - 1, r->ev_d[l - 1]));
  }
  mark();                                               // 2: dX chain issued/done on `compute`
  // The update is HBM-bound, the first layer's gradient GEMM (the largest, and the last to become computable) is
  // tensor-bound and reads neither weights nor deltas: once the dX chain is done the layers >= 2 are updated while dW_1
  // still runs on the side stream.  The early launch leaves thread slots free so the GEMM's CTAs become resident.
  const int sgd_early_blocks = tunable(TUN_SGD_EARLY);  // 0 = one update launch after all gradients (reference order)
  long long tail_end4 = r->arena_floats / 4;
  if (sgd_early_blocks > 0 && r->L >= 2 && !r->dp_p2p) {
    CU_TRY(cudaStreamWaitEvent(r->compute, r->ev_upper, 0));
    if (r->nccl_comm) CU_TRY(cudaStreamWaitEvent(r->compute, r->ev_comm_upper, 0));
    tail_end4 = r->layer[2].off / 4;
    BP_TRY(launch_sgd_range(r, tail_end4, r->arena_floats / 4, sgd_early_blocks));
    tl_mark(r, r->compute, "update, layers >= 2");
  }
  mark();                                               // 3: early update of layers >= 2 done
  CU_TRY(cudaEventRecord(r->ev_side, r->side));
  CU_TRY(cudaStreamWaitEvent(r->compute, r->ev_side, 0));
  tl_mark(r, r->compute, "join (all dW done)");
  mark();                                               // 4: all dW done
  if (r->nccl_comm && !r->dp_p2p) {
    CU_TRY(cudaEventRecord(r->ev_comm, r->comm_stream));
    CU_TRY(cudaStreamWaitEvent(r->compute, r->ev_comm, 0));
  }
  mark();                                               // 5: all-reduce waited
  if (r->dp_p2p) BP_TRY(peer_exchange(r));              // owner-side reduce + update + all-gather
  else BP_TRY(launch_sgd_range(r, 0, tail_end4, 8));
  tl_mark(r, r->compute, "update / exchange, end of bunch");
  mark();                                               // 6: sgd done
  r->step++;
  r->bunches++;
  if (prof) r->prof_cnt++;
  if (r->timeline && r->tl_bunches < (uint64_t)Rank::kTlBunches) {
    r->tl_marks = r->tl_idx;
    r->tl_bunches++;
  }
  return BP_OK;
}

int rank_train_resident(Rank* r, int first_bunch, int n_bunches) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  if (r->cfg.world_size > 1 && !r->nccl_comm && !r->dp_p2p)  // would scale by 2/B_global over B/N rows and diverge
    return fail(BP_ECOMM, "train: rank %d of %d has no communicator (call bp_comm_init first)", r->cfg.rank,
                r->cfg.world_size);
  const bool in_drop = r->cfg.dropoutflag == 1 && r->cfg.visible_omit > 0.0f;
  ChunkBuf& c = r->chunk[r->cur];
  if (!c.x || !c.has_targ) return fail(BP_EINVAL, "train_resident: no resident chunk with targets");
  if (first_bunch < 0 || n_bunches < 0 || (long long)(first_bunch + n_bunches) * r->local_bunch > c.rows)
    return fail(BP_EINVAL, "train_resident: bunches [%d,%d) exceed resident rows %d (bunch %d)", first_bunch,
                first_bunch + n_bunches, c.rows, r->local_bunch);
  const int li = r->loss_cur ^ 1;
  if (n_bunches > r->loss_cap[li]) {
    CU_TRY(cudaStreamSynchronize(r->compute));
    CU_TRY(cudaStreamSynchronize(r->copy));
    if (r->loss_dev[li]) CU_TRY(cudaFree(r->loss_dev[li]));
    r->loss_dev[li] = nullptr;
    r->loss_cap[li] = std::max(n_bunches, 256);
    CU_TRY(cudaMalloc(&r->loss_dev[li], sizeof(double) * r->loss_cap[li]));
  }
  if (n_bunches > 0) CU_TRY(cudaMemsetAsync(r->loss_dev[li], 0, sizeof(double) * n_bunches, r->compute));
  r->loss_n[li] = n_bunches;
  r->loss_cur = li;
  if (in_drop && n_bunches > 0) {
    // the input mask zeroes the resident rows themselves: training them again would stack a second mask on the first
    if (first_bunch * r->local_bunch < c.masked_rows)
      return fail(BP_EINVAL, "train_resident: rows from %d on were already trained with input dropout (visible_omit), "
                  "which masks the resident chunk in place — upload the chunk again", first_bunch * r->local_bunch);
    c.masked_rows = (first_bunch + n_bunches) * r->local_bunch;
  }
  for (int b = 0; b < n_bunches; ++b)
    BP_TRY(train_bunch(r, c, (first_bunch + b) * r->local_bunch, r->loss_dev[li] + b));
  CU_TRY(cudaEventRecord(r->loss_done[li], r->compute));
  CU_TRY(cudaEventRecord(c.consumed, r->compute));
  c.consumed_valid = true;
  return BP_OK;
}

int rank_forward_resident(Rank* r, int first_frame, int n_frames, float* out_host, double* sum_sq) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  ChunkBuf& c = r->chunk[r->cur];
  if (!c.x) return fail(BP_EINVAL, "forward_resident: no resident chunk");
  if (first_frame < 0 || n_frames <= 0 || first_frame + n_frames > c.rows)
    return fail(BP_EINVAL, "forward_resident: rows [%d,%d) exceed resident rows %d", first_frame,
                first_frame + n_frames, c.rows);
  if (sum_sq && !c.has_targ) return fail(BP_EINVAL, "forward_resident: squared error needs resident targets");
  const int no = r->Nout();
  if (out_host && n_frames > r->out_cap_rows) {
    CU_TRY(cudaStreamSynchronize(r->compute));
    if (r->out_dev) CU_TRY(cudaFree(r->out_dev));
    r->out_dev = nullptr;
    CU_TRY(cudaMalloc(&r->out_dev, size_t(n_frames) * no * 4));
    r->out_cap_rows = n_frames;
  }
  if (sum_sq) CU_TRY(cudaMemsetAsync(r->sqerr_dev, 0, sizeof(double), r->compute));
  const int B = r->cfg.bunchsize;
  for (int i = 0; i < n_frames; i += B) {
    const int n = std::min(B, n_frames - i);
    BP_TRY(forward_rows(r, c, first_frame + i, n, false, out_host ? r->out_dev + (long long)i * no : nullptr, no,
                        sum_sq ? r->sqerr_dev : nullptr));
  }
  CU_TRY(cudaEventRecord(c.consumed, r->compute));
  c.consumed_valid = true;
  if (out_host)
    CU_TRY(cudaMemcpyAsync(out_host, r->out_dev, size_t(n_frames) * no * 4, cudaMemcpyDeviceToHost, r->compute));
  if (sum_sq) CU_TRY(cudaMemcpyAsync(sum_sq, r->sqerr_dev, sizeof(double), cudaMemcpyDeviceToHost, r->compute));
  if (out_host || sum_sq) CU_TRY(cudaStreamSynchronize(r->compute));
  return BP_OK;
}

// Pipelined decode of raw chunks: queue "records -> rows (splice kernel) -> forward pass -> D2H into out_host" and
// return once the records have left the caller's buffers; at most two chunks in flight.  The chunk buffers are
// double-buffered already (upload of chunk k+1 beside the forward pass of chunk k); what this adds is a second output
// slot and a D2H stream, so that no stage of chunk k+1 waits for the host to have collected chunk k.
// rc != null: raw records (splice on the device); rc == null: n_rows spliced, normalised rows at `rows` (bp_forward's
// input), for callers that keep the reference's host reader.
int rank_decode_submit(Rank* r, const bp_raw_chunk* rc, int n_rows, const float* rows, float* out_host) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  if (r->dec_inflight >= 2)
    return fail(BP_EINVAL, "decode submit: two chunks already in flight, call the matching wait first");
  Rank::DecodeSlot& ds = r->dec[r->dec_head];
  const int n = rc ? rc->n_samples : n_rows, no = r->Nout();
  if (!r->d2h) CU_TRY(cudaStreamCreateWithFlags(&r->d2h, cudaStreamNonBlocking));
  if (!ds.fwd_done) CU_TRY(cudaEventCreateWithFlags(&ds.fwd_done, cudaEventDisableTiming));
  if (!ds.out_done) CU_TRY(cudaEventCreateWithFlags(&ds.out_done, cudaEventDisableTiming));
  if (n > ds.cap_rows) {  // the slot is idle: its previous chunk has been waited for
    if (ds.out_dev) CU_TRY(cudaFree(ds.out_dev));
    ds.out_dev = nullptr;
    ds.cap_rows = 0;
    CU_TRY(cudaMalloc(&ds.out_dev, size_t(n) * no * 4));
    ds.cap_rows = n;
  }
  // both return when the caller's buffers have been consumed; `compute` waits for the rows
  if (rc) BP_TRY(rank_upload_raw(r, rc, true));
  else BP_TRY(rank_upload(r, n, rows, nullptr));
  ChunkBuf& c = r->chunk[r->cur];
  const int B = r->cfg.bunchsize;
  for (int i = 0; i < n; i += B)
    BP_TRY(forward_rows(r, c, i, std::min(B, n - i), false, ds.out_dev + (long long)i * no, no, nullptr));
  CU_TRY(cudaEventRecord(c.consumed, r->compute));
  c.consumed_valid = true;
  CU_TRY(cudaEventRecord(ds.fwd_done, r->compute));
  CU_TRY(cudaStreamWaitEvent(r->d2h, ds.fwd_done, 0));
  CU_TRY(cudaMemcpyAsync(out_host, ds.out_dev, size_t(n) * no * 4, cudaMemcpyDeviceToHost, r->d2h));
  CU_TRY(cudaEventRecord(ds.out_done, r->d2h));
  r->dec_head ^= 1;
  r->dec_inflight++;
  return BP_OK;
}

// Host wait for the oldest chunk in flight: its enhanced frames are complete in the `out` given at submit time.
int rank_decode_wait(Rank* r) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  if (r->dec_inflight <= 0) return fail(BP_EINVAL, "decode wait: no chunk in flight");
  const int oldest = (r->dec_head + 2 - r->dec_inflight) & 1;
  CU_TRY(cudaEventSynchronize(r->dec[oldest].out_done));
  r->dec_inflight--;
  return BP_OK;
}

int rank_return_weights(Rank* r, float* const* weights, float* const* bias) {
  CU_TRY(cudaSetDevice(r->cfg.device));
  for (int l = 1; l <= r->L; ++l) {
    LayerState& ls = r->layer[l];
    CU_TRY(cudaMemcpy2DAsync(weights[l], size_t(ls.N) * 4, r->w + ls.off, ls.ldN * 4, size_t(ls.N) * 4, ls.K,
                             cudaMemcpyDeviceToHost, r->compute));
    CU_TRY(cudaMemcpyAsync(bias[l], r->w + ls.off + (long long)ls.K * ls.ldN, size_t(ls.N) * 4,
                           cudaMemcpyDeviceToHost, r->compute));
  }
  CU_TRY(cudaStreamSynchronize(r->compute));
  return BP_OK;
}

}  // namespace

// ================================================================================================ C ABI
struct bp_handle {
  std::vector<Rank*> ranks;  // 1 entry for a rank handle; gpu_used entries for an in-process group
};

namespace {

// Run fn(rank_index) on every rank of a group, one host thread per GPU; returns the first failure.
template <class F>
int for_each_rank(bp_handle* h, F fn) {
  if (h->ranks.size() == 1) return fn(0);
  std::vector<int> rc(h->ranks.size(), BP_OK);
  std::vector<std::string> msg(h->ranks.size());
  std::vector<std::thread> th;
  for (size_t i = 0; i < h->ranks.size(); ++i)
    th.emplace_back([&, i] {
      rc[i] = fn((int)i);
      if (rc[i] != BP_OK) msg[i] = g_err;
    });
  for (auto& t : th) t.join();
  for (size_t i = 0; i < rc.size(); ++i)
    if (rc[i] != BP_OK) {
      g_err = msg[i];
      return rc[i];
    }
  return BP_OK;
}

}  // namespace

extern "C" {

const char* bp_last_error(void) { return g_err.c_str(); }
int bp_version(void) { return 100; }

int bp_create_ex(bp_handle** h, const bp_config* cfg, float* const* weights, float* const* bias) {
  if (!h) return fail(BP_EINVAL, "bp_create_ex: null handle pointer");
  *h = nullptr;
  Rank* r = nullptr;
  BP_TRY(rank_create(&r, cfg, weights, bias));
  bp_handle* hh = new bp_handle();
  hh->ranks.push_back(r);
  *h = hh;
  return BP_OK;
}

// One process, one host thread per rank (BPtrain gpu_used=N): peer access is enabled directly.
int group_p2p_connect(bp_handle* h) {
  const int want = dp_mode_wanted();
  const int W = (int)h->ranks.size();
  if (want == 0 || W < 2) return BP_OK;
  bool can = W <= kMaxPeers;
  for (int i = 0; i < W && can; ++i)
    for (int j = 0; j < W && can; ++j) {
      int yes = 1;
      if (i != j && (cudaDeviceCanAccessPeer(&yes, h->ranks[i]->cfg.device, h->ranks[j]->cfg.device) != cudaSuccess ||
                     !yes))
        can = false;
    }
  if (can) {
    int rc = for_each_rank(h, [&](int i) -> int {
      Rank* r = h->ranks[i];
      CU_TRY(cudaSetDevice(r->cfg.device));
      for (int j = 0; j < W; ++j) {
        if (j == i) continue;
        cudaError_t e = cudaDeviceEnablePeerAccess(h->ranks[j]->cfg.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(BP_ECOMM, "peer access %d->%d", i, j);
        cudaGetLastError();
      }
      return rank_p2p_alloc(r);
    });
    can = rc == BP_OK;
  }
  if (!can) {
    for (Rank* r : h->ranks) {
      cudaSetDevice(r->cfg.device);
      rank_p2p_release(r);
    }
    if (want == 2) return fail(BP_ECOMM, "BP_DP=p2p: the devices cannot access each other's memory");
    fprintf(stderr, "libbpgpu: no peer access between the devices, using NCCL all-reduce\n");
    return BP_OK;
  }
  for (int i = 0; i < W; ++i)
    for (int j = 0; j < W; ++j) {
      Rank::PeerMem& pm = h->ranks[i]->peer[j];
      pm.w = h->ranks[j]->w;
      pm.w_lo = h->ranks[j]->w_lo;
      pm.recv = h->ranks[j]->recv;
      pm.flags = h->ranks[j]->flags;
    }
  for (Rank* r : h->ranks) {
    r->dp_p2p = 1;
    r->chain_built = false;
  }
  if (getenv("BP_VERBOSE")) fprintf(stderr, "libbpgpu: gradient exchange over peer memory, %d ranks (one process)\n", W);
  return BP_OK;
}

int bp_create(bp_handle** h, int gpu_used, int numlayers, const int* layersizes, int bunchsize, float lrate,
              float momentum, float weightcost, float* const* weights, float* const* bias, int dropoutflag,
              float visible_omit, float hid_omit) {
  if (!h || !layersizes) return fail(BP_EINVAL, "bp_create: null argument");
  *h = nullptr;
  if (numlayers < 2 || numlayers > BP_MAXLAYER) return fail(BP_EINVAL, "bp_create: numlayers=%d", numlayers);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(BP_ENODEV, "no CUDA device (this library has no CPU fallback)");
  }
  if (gpu_used < 1 || gpu_used > ndev)  // BP_GPU.cu:20-24 "GPU selected out of range"
    return fail(BP_ENODEV, "gpu_used=%d out of range [1,%d]", gpu_used, ndev);
  bp_config cfg{};
  cfg.world_size = gpu_used;
  cfg.numlayers = numlayers;
  for (int i = 0; i < numlayers; ++i) cfg.layersizes[i] = layersizes[i];
  cfg.bunchsize = bunchsize;
  cfg.lrate = lrate;
  cfg.momentum = momentum;
  cfg.weightcost = weightcost;
  cfg.dropoutflag = dropoutflag;
  cfg.visible_omit = visible_omit;
  cfg.hid_omit = hid_omit;
  cfg.activation = BP_ACT_RELU;
  cfg.math_mode = BP_MATH_TF32;
  cfg.seed = 0x5eed5eedULL;
  if (const char* e = getenv("BP_ACTIVATION")) cfg.activation = (strcmp(e, "sigmoid") == 0) ? BP_ACT_SIGMOID : 0;
  if (const char* e = getenv("BP_SEED")) cfg.seed = strtoull(e, nullptr, 0);
  if (const char* e = getenv("BP_MATH")) cfg.math_mode = (strcmp(e, "3xtf32") == 0) ? BP_MATH_3XTF32 : BP_MATH_TF32;

  bp_handle* hh = new bp_handle();
  hh->ranks.assign(gpu_used, nullptr);
  int rc = for_each_rank(hh, [&](int i) -> int {
    bp_config c = cfg;
    c.rank = i;
    c.device = i;
    return rank_create(&hh->ranks[i], &c, weights, bias);
  });
  // The default exchange runs over peer memory and needs no NCCL at all: try it first, and create the NCCL
  // communicator (which needs a loadable libnccl) only when BP_DP=nccl asks for it or the devices cannot map each other.
  if (rc == BP_OK) rc = group_p2p_connect(hh);
  if (rc == BP_OK && gpu_used > 1 && !hh->ranks[0]->dp_p2p) {
    char id[128];
    rc = bp_comm_unique_id(id);
    if (rc == BP_OK)
      rc = for_each_rank(hh, [&](int i) -> int {
        CU_TRY(cudaSetDevice(hh->ranks[i]->cfg.device));
        NcclApi::Id128 uid;
        memcpy(uid.b, id, 128);
        NCCL_TRY(g_nccl.CommInitRank(&hh->ranks[i]->nccl_comm, gpu_used, uid, i));
        return BP_OK;
      });
  }
  if (rc != BP_OK) {
    std::string keep = g_err;
    bp_destroy(hh);
    g_err = keep;
    return rc;
  }
  *h = hh;
  return BP_OK;
}

void bp_destroy(bp_handle* h) {
  if (!h) return;
  for (Rank* r : h->ranks) rank_destroy(r);
  delete h;
}

int bp_upload_chunk(bp_handle* h, int n_frames, const float* in, const float* targ) {
  if (!h) return fail(BP_EINVAL, "null handle");
  if (h->ranks.size() == 1) return rank_upload(h->ranks[0], n_frames, in, targ);
  // In-process group: rank r owns rows [r*lb, (r+1)*lb) of every global bunch.  Re-pack per rank, bunch by bunch.
  const int G = (int)h->ranks.size();
  const int B = h->ranks[0]->cfg.bunchsize, lb = B / G;
  const int nb = n_frames / B;
  if (nb == 0) return fail(BP_EINVAL, "upload: chunk of %d frames has no full bunch of %d", n_frames, B);
  return for_each_rank(h, [&](int i) -> int {
    Rank* r = h->ranks[i];
    const int k0 = r->K0(), no = r->Nout();
    std::vector<float> xin((size_t)nb * lb * k0), tin(targ ? (size_t)nb * lb * no : 0);
    for (int b = 0; b < nb; ++b) {
      memcpy(&xin[(size_t)b * lb * k0], in + ((size_t)b * B + (size_t)i * lb) * k0, sizeof(float) * lb * k0);
      if (targ)
        memcpy(&tin[(size_t)b * lb * no], targ + ((size_t)b * B + (size_t)i * lb) * no, sizeof(float) * lb * no);
    }
    return rank_upload(r, nb * lb, xin.data(), targ ? tin.data() : nullptr);
  });
}

static int check_raw_chunk(bp_handle* h, const bp_raw_chunk* rc) {
  if (!h || !rc) return fail(BP_EINVAL, "null argument");
  Rank* r0 = h->ranks[0];
  if (!rc->fea_records || !rc->mean || !rc->inv_std || !rc->sample_frame || (rc->nat && !rc->sample_seg))
    return fail(BP_EINVAL, "upload_raw: null table");
  if (rc->fea_dim <= 0 || rc->fea_context <= 0 || rc->n_records <= 0 || rc->n_samples <= 0)
    return fail(BP_EINVAL, "upload_raw: fea_dim=%d context=%d records=%d samples=%d", rc->fea_dim, rc->fea_context,
                rc->n_records, rc->n_samples);
  if (rc->fea_dim * (rc->fea_context + (rc->nat ? 1 : 0)) != r0->K0())
    return fail(BP_EINVAL, "upload_raw: fea_dim %d x (context %d + nat %d) != layersizes[0] %d", rc->fea_dim,
                rc->fea_context, rc->nat ? 1 : 0, r0->K0());
  if (rc->targ_records && (rc->targ_offset < 0 || rc->targ_offset >= rc->fea_context))
    return fail(BP_EINVAL, "upload_raw: targ_offset %d outside the context window", rc->targ_offset);
  for (int i = 0; i < rc->n_samples; ++i) {
    const int f = rc->sample_frame[i];
    if (f < 0 || f + rc->fea_context > rc->n_records)
      return fail(BP_EINVAL, "upload_raw: sample %d window [%d,%d) outside the %d records", i, f,
                  f + rc->fea_context, rc->n_records);
    if (rc->nat && (rc->sample_seg[i] < 0 || rc->sample_seg[i] > f))
      return fail(BP_EINVAL, "upload_raw: sample %d segment start %d not in [0,%d]", i, rc->sample_seg[i], f);
    if (rc->sample_row && (rc->sample_row[i] < 0 || rc->sample_row[i] >= rc->n_samples))
      return fail(BP_EINVAL, "upload_raw: sample %d row %d outside [0,%d)", i, rc->sample_row[i], rc->n_samples);
  }
  return BP_OK;
}

int bp_upload_raw_chunk(bp_handle* h, const bp_raw_chunk* rc) {
  BP_TRY(check_raw_chunk(h, rc));
  return for_each_rank(h, [&](int i) { return rank_upload_raw(h->ranks[i], rc, false); });
}

int bp_crossvalid_raw(bp_handle* h, const bp_raw_chunk* rc, float* sum_sq_err, float* out) {
  BP_TRY(check_raw_chunk(h, rc));
  if (sum_sq_err && !rc->targ_records) return fail(BP_EINVAL, "bp_crossvalid_raw: squared error needs target records");
  Rank* r = h->ranks[0];  // device 0 only, like bp_crossvalid
  BP_TRY(rank_upload_raw(r, rc, true));
  double s = 0.0;
  BP_TRY(rank_forward_resident(r, 0, rc->n_samples, out, sum_sq_err ? &s : nullptr));
  if (sum_sq_err) *sum_sq_err = (float)s;
  return BP_OK;
}

int bp_decode_raw_submit(bp_handle* h, const bp_raw_chunk* rc, float* out) {
  BP_TRY(check_raw_chunk(h, rc));
  if (!out) return fail(BP_EINVAL, "bp_decode_raw_submit: null output buffer");
  return rank_decode_submit(h->ranks[0], rc, 0, nullptr, out);  // device 0 only, like bp_crossvalid
}

int bp_forward_submit(bp_handle* h, int n_frames, const float* in, float* out) {
  if (!h || !in || !out || n_frames <= 0) return fail(BP_EINVAL, "bp_forward_submit: bad argument");
  return rank_decode_submit(h->ranks[0], nullptr, n_frames, in, out);
}

int bp_forward_wait(bp_handle* h) { return bp_decode_raw_wait(h); }

int bp_decode_raw_wait(bp_handle* h) {
  if (!h) return fail(BP_EINVAL, "null handle");
  return rank_decode_wait(h->ranks[0]);
}

int bp_train_raw(bp_handle* h, const bp_raw_chunk* rc) {
  if (!h || !rc) return fail(BP_EINVAL, "null argument");
  if (!rc->targ_records) return fail(BP_EINVAL, "bp_train_raw: no target records");
  // NB unlike bp_train, a raw chunk is the GLOBAL chunk for every kind of handle: each rank (thread of a group, or
  // process with its own rank handle) is given the same records and sample table and assembles ITS rows of every
  // global bunch on its device (rank_upload_raw), so the bunch count divides by the global bunch size
  const int B = h->ranks[0]->cfg.bunchsize;
  const int nb = rc->n_samples / B;
  if (rc->n_samples % B) printf("this bunch has only %d samples and is ignored.\n", rc->n_samples % B);  // BP_GPU.cu:317
  if (nb == 0) return BP_OK;
  if (h->ranks.size() == 1) {
    // queue the bunches while the records are still in flight; wait for the DMA last
    BP_TRY(check_raw_chunk(h, rc));
    Rank* r = h->ranks[0];
    BP_TRY(rank_upload_raw(r, rc, false, false));
    const int rc_train = rank_train_resident(r, 0, nb);
    const std::string keep = g_err;
    const int rc_wait = rank_wait_uploaded(r);  // the caller may reuse its buffers only after this
    if (rc_train != BP_OK) g_err = keep;
    return rc_train != BP_OK ? rc_train : rc_wait;
  }
  BP_TRY(bp_upload_raw_chunk(h, rc));
  return bp_train_resident(h, 0, nb);
}

int bp_download_chunk(bp_handle* h, int first_row, int n_rows, float* in, float* targ) {
  if (!h || n_rows <= 0 || first_row < 0) return fail(BP_EINVAL, "bp_download_chunk: bad argument");
  if (h->ranks.size() != 1) return fail(BP_EINVAL, "bp_download_chunk: single-rank handles only");
  Rank* r = h->ranks[0];
  CU_TRY(cudaSetDevice(r->cfg.device));
  ChunkBuf& c = r->chunk[r->cur];
  if (!c.x || first_row + n_rows > c.rows)
    return fail(BP_EINVAL, "bp_download_chunk: rows [%d,%d) exceed resident rows %d", first_row, first_row + n_rows,
                c.rows);
  CU_TRY(cudaStreamSynchronize(r->copy));
  CU_TRY(cudaStreamSynchronize(r->compute));
  if (in)
    CU_TRY(cudaMemcpy2D(in, (size_t)r->K0() * 4, c.x + (long long)first_row * r->ldx, r->ldx * 4, (size_t)r->K0() * 4,
                        n_rows, cudaMemcpyDeviceToHost));
  if (targ) {
    if (!c.has_targ) return fail(BP_EINVAL, "bp_download_chunk: no resident targets");
    CU_TRY(cudaMemcpy(targ, c.t + (long long)first_row * r->Nout(), (size_t)n_rows * r->Nout() * 4,
                      cudaMemcpyDeviceToHost));
  }
  return BP_OK;
}

int bp_train_resident(bp_handle* h, int first_bunch, int n_bunches) {
  if (!h) return fail(BP_EINVAL, "null handle");
  return for_each_rank(h, [&](int i) { return rank_train_resident(h->ranks[i], first_bunch, n_bunches); });
}

int bp_train(bp_handle* h, int n_frames, const float* in, const float* targ) {
  if (!h || !in || !targ) return fail(BP_EINVAL, "bp_train: null argument");
  Rank* r0 = h->ranks[0];
  const int per_call_bunch = (h->ranks.size() == 1) ? r0->local_bunch : r0->cfg.bunchsize;
  const int nb = n_frames / per_call_bunch;
  if (n_frames % per_call_bunch)
    printf("this bunch has only %d samples and is ignored.\n", n_frames % per_call_bunch);  // BP_GPU.cu:317
  if (nb == 0) return BP_OK;
  if (h->ranks.size() == 1) {
    // queue the bunches while the chunk is still in flight; wait for the DMA last
    BP_TRY(rank_upload(r0, n_frames, in, targ, false));
    const int rc_train = rank_train_resident(r0, 0, nb);
    const std::string keep = g_err;
    const int rc_wait = rank_wait_uploaded(r0);  // the caller may overwrite in / targ only after this
    if (rc_train != BP_OK) g_err = keep;
    return rc_train != BP_OK ? rc_train : rc_wait;
  }
  BP_TRY(bp_upload_chunk(h, n_frames, in, targ));
  return bp_train_resident(h, 0, nb);
}

int bp_set_dropout_seed(bp_handle* h, uint64_t seed) {
  if (!h) return fail(BP_EINVAL, "null handle");
  for (Rank* r : h->ranks) {
    r->cfg.seed = seed;
    r->chain_built = false;  // the chained launches' product tables carry the seed
  }
  return BP_OK;
}

int bp_begin_epoch(bp_handle* h, float lrate, float momentum, float weightcost, int reset_dropout_step) {
  if (!h) return fail(BP_EINVAL, "null handle");
  return for_each_rank(h, [&](int i) -> int {
    Rank* r = h->ranks[i];
    CU_TRY(cudaSetDevice(r->cfg.device));
    r->cfg.lrate = lrate;
    r->cfg.momentum = momentum;
    r->cfg.weightcost = weightcost;
    if (reset_dropout_step) r->step = 0;
    // stream order: after the last bunch of the previous epoch, before the first of the next
    CU_TRY(cudaMemsetAsync(r->dw, 0, sizeof(float) * (size_t)r->arena_floats, r->compute));
    return BP_OK;
  });
}

int bp_forward_resident(bp_handle* h, int first_frame, int n_frames, float* out_host, double* sum_sq_err) {
  if (!h) return fail(BP_EINVAL, "null handle");
  return rank_forward_resident(h->ranks[0], first_frame, n_frames, out_host, sum_sq_err);
}

int bp_crossvalid(bp_handle* h, int n_frames, const float* in, const float* targ, float* sum_sq_err) {
  if (!h || !in || !targ || !sum_sq_err) return fail(BP_EINVAL, "bp_crossvalid: null argument");
  if (n_frames <= 0) { *sum_sq_err = 0.0f; return BP_OK; }
  Rank* r = h->ranks[0];  // CV runs on device 0 only, like the reference (BP_GPU.cu:440-441)
  BP_TRY(rank_upload(r, n_frames, in, targ));
  double s = 0.0;
  BP_TRY(rank_forward_resident(r, 0, n_frames, nullptr, &s));
  *sum_sq_err = (float)s;
  return BP_OK;
}

int bp_forward(bp_handle* h, int n_frames, const float* in, float* out) {
  if (!h || !in || !out) return fail(BP_EINVAL, "bp_forward: null argument");
  Rank* r = h->ranks[0];
  BP_TRY(rank_upload(r, n_frames, in, nullptr));
  return rank_forward_resident(r, 0, n_frames, out, nullptr);
}

int bp_return_weights(bp_handle* h, float* const* weights, float* const* bias) {
  if (!h || !weights || !bias) return fail(BP_EINVAL, "bp_return_weights: null argument");
  return rank_return_weights(h->ranks[0], weights, bias);  // device 0, BP_GPU.cu:915
}

int bp_set_option(bp_handle* h, const char* name, int value) {
  if (!h || !name) return fail(BP_EINVAL, "bp_set_option: null argument");
  for (Rank* r : h->ranks) {
    if (strcmp(name, "chain") == 0) r->use_chain = value < 0 ? -1 : value != 0;  // -1 auto, 0 per product, 1 chained
    else if (strcmp(name, "chain_trace") == 0) r->chain_trace_on = value != 0;
    else if (set_tunable(name, value) != BP_OK)                            // process-wide switches (bp_internal.h)
      return fail(BP_EINVAL, "bp_set_option: unknown option '%s'", name);
    r->chain_built = false;  // the product tables capture the switches (hints, streaming stores): rebuild lazily
  }
  return BP_OK;
}

// SM clock as the device sees it: clock64 ticks per globaltimer nanosecond over a ~50 us spin of one thread.
__global__ void bp_clock_probe_kernel(unsigned long long* out) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  do {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  } while (t1 - t0 < 50000ull);
  const long long c1 = clock64();
  out[0] = (unsigned long long)(c1 - c0);
  out[1] = t1 - t0;
}

int bp_get_option(bp_handle* h, const char* name, int* value) {
  if (!h || !name || !value) return fail(BP_EINVAL, "bp_get_option: null argument");
  Rank* r = h->ranks[0];
  if (strcmp(name, "dp_exchange") == 0) {
    *value = r->dp_p2p ? 2 : (r->nccl_comm ? 1 : 0);
    if (h->ranks.size() > 1 && *value == 0) *value = 1;
    return BP_OK;
  }
  if (strcmp(name, "sm_clock_mhz") == 0) {
    CU_TRY(cudaSetDevice(r->cfg.device));
    unsigned long long* d = nullptr;
    unsigned long long hv[2] = {0, 0};
    CU_TRY(cudaMalloc(&d, 16));
    bp_clock_probe_kernel<<<1, 1, 0, r->compute>>>(d);
    cudaError_t e = cudaMemcpyAsync(hv, d, 16, cudaMemcpyDeviceToHost, r->compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->compute);
    cudaFree(d);
    CU_TRY(e);
    *value = hv[1] ? (int)(1000.0 * (double)hv[0] / (double)hv[1] + 0.5) : 0;
    return BP_OK;
  }
  if (strcmp(name, "chain") == 0) { *value = r->use_chain; return BP_OK; }
  if (get_tunable(name, value) == BP_OK) return BP_OK;
  return fail(BP_EINVAL, "bp_get_option: unknown option '%s'", name);
}

// Host-only: the schedule of the chained launches for a net / bunch / pair count (include/bp_gpu_debug.h).
int bp_debug_chain_plan(int numlayers, const int* layersizes, int rows, int pairs, int which, int passes,
                        int max_items, int* n_items, int* pair, int* prod, int* mt, int* nt, int max_prods, int* n_prods,
                        int* dep_prod, int* dep_all, int* m_tiles, int* n_tiles, int* pair_n, int* n_cols,
                        long long* makespan_cycles) {
  if (numlayers < 2 || numlayers > BP_MAXLAYER || !layersizes || rows <= 0 || pairs <= 0 || pairs > CHAIN_MAX_PAIRS ||
      !n_items || !n_prods)
    return fail(BP_EINVAL, "bp_debug_chain_plan: bad argument");
  const std::vector<ChainShape> shapes = chain_shapes(layersizes, numlayers - 1, rows, pairs, passes, which);
  std::vector<ChainItem> items;
  std::vector<int> pair_off;
  const long long ms = chain_schedule(shapes, pairs, items, pair_off);
  if (ms < 0) return fail(BP_EINVAL, "bp_debug_chain_plan: dependency cycle");
  if (makespan_cycles) *makespan_cycles = ms;
  *n_items = (int)items.size();
  *n_prods = (int)shapes.size();
  if ((int)items.size() > max_items || (int)shapes.size() > max_prods)
    return fail(BP_EINVAL, "bp_debug_chain_plan: %d items / %d products exceed the buffers", (int)items.size(),
                (int)shapes.size());
  for (int p = 0; p < pairs; ++p)
    for (int i = pair_off[p]; i < pair_off[p + 1]; ++i) {
      if (pair) pair[i] = p;
      if (prod) prod[i] = items[i].prod;
      if (mt) mt[i] = items[i].mt;
      if (nt) nt[i] = items[i].nt;
    }
  for (size_t q = 0; q < shapes.size(); ++q) {
    if (dep_prod) dep_prod[q] = shapes[q].dep_prod;
    if (dep_all) dep_all[q] = shapes[q].dep_all;
    if (m_tiles) m_tiles[q] = shapes[q].m_tiles;
    if (n_tiles) n_tiles[q] = shapes[q].n_tiles;
    if (pair_n) pair_n[q] = shapes[q].pair_n;
    if (n_cols) n_cols[q] = shapes[q].n_cols;
  }
  return BP_OK;
}

// Bring-up aid: the stamps of the most recent chained launch (which: 0 forward, 1 back-propagation) as CSV text,
// one line per tile: item,pair,prod,mt,nt,deps_ns,first_full_ns,acc_ready_ns,epilogue_done_ns (relative to the earliest).
int bp_debug_chain_trace(bp_handle* h, int which, char* buf, int len) {
  if (!h || !buf || len <= 0) return fail(BP_EINVAL, "bp_debug_chain_trace: null argument");
  Rank* r = h->ranks[0];
  CU_TRY(cudaSetDevice(r->cfg.device));
  if (!r->chain_trace || !r->chain_built) return fail(BP_EINVAL, "bp_debug_chain_trace: no traced launch yet");
  CU_TRY(cudaStreamSynchronize(r->compute));
  Rank::ChainPlan& pl = which ? r->chain_bwd : r->chain_fwd;
  const int n = std::min(pl.n_items, 4096);
  std::vector<unsigned long long> t((size_t)n * 4);
  CU_TRY(cudaMemcpy(t.data(), r->chain_trace + (which ? 4096 * 4 : 0), t.size() * 8, cudaMemcpyDeviceToHost));
  unsigned long long t0 = ~0ull;
  for (auto v : t) if (v && v < t0) t0 = v;
  std::string out;
  char line[160];
  for (int p = 0; p + 1 < (int)pl.h_pair_off.size(); ++p)
    for (int i = pl.h_pair_off[p]; i < pl.h_pair_off[p + 1] && i < n; ++i) {
      const ChainItem& it = pl.h_items[i];
      snprintf(line, sizeof line, "%d,%d,%d,%d,%d,%lld,%lld,%lld,%lld\n", i, p, it.prod, it.mt, it.nt,
               (long long)(t[4 * i] - t0), (long long)(t[4 * i + 1] - t0), (long long)(t[4 * i + 2] - t0),
               (long long)(t[4 * i + 3] - t0));
      out += line;
    }
  if ((int)out.size() + 1 > len) return fail(BP_EINVAL, "bp_debug_chain_trace: buffer too small (%d needed)", (int)out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return BP_OK;
}

int bp_sync(bp_handle* h) {
  if (!h) return fail(BP_EINVAL, "null handle");
  return for_each_rank(h, [&](int i) -> int {
    Rank* r = h->ranks[i];
    CU_TRY(cudaSetDevice(r->cfg.device));
    CU_TRY(cudaStreamSynchronize(r->copy));
    CU_TRY(cudaStreamSynchronize(r->comm_stream));
    CU_TRY(cudaStreamSynchronize(r->side));
    CU_TRY(cudaStreamSynchronize(r->compute));
    return BP_OK;
  });
}

void* bp_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void bp_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int bp_timer_start(bp_handle* h) {
  if (!h) return fail(BP_EINVAL, "null handle");
  Rank* r = h->ranks[0];
  CU_TRY(cudaSetDevice(r->cfg.device));
  CU_TRY(cudaEventRecord(r->ev_t0, r->compute));
  return BP_OK;
}
int bp_timer_stop(bp_handle* h, float* elapsed_ms) {
  if (!h || !elapsed_ms) return fail(BP_EINVAL, "null argument");
  Rank* r = h->ranks[0];
  CU_TRY(cudaSetDevice(r->cfg.device));
  CU_TRY(cudaEventRecord(r->ev_t1, r->compute));
  CU_TRY(cudaEventSynchronize(r->ev_t1));
  CU_TRY(cudaEventElapsedTime(elapsed_ms, r->ev_t0, r->ev_t1));
  return BP_OK;
}

int bp_get_counters(bp_handle* h, uint64_t* kernel_launches, uint64_t* train_bunches) {
  if (!h) return fail(BP_EINVAL, "null handle");
  uint64_t l = 0;
  for (Rank* r : h->ranks) l += r->launches;
  if (kernel_launches) *kernel_launches = l;
  if (train_bunches) *train_bunches = h->ranks[0]->bunches;
  return BP_OK;
}

int bp_set_profiling(bp_handle* h, int on) {
  if (!h) return fail(BP_EINVAL, "null handle");
  for (Rank* r : h->ranks) {
    r->profiling = on == 1;
    r->prof_cnt = 0;
    r->timeline = on == 2;
    r->tl_bunches = 0;
    r->tl_marks = 0;
  }
  return BP_OK;
}

int bp_get_timeline(bp_handle* h, char* labels, int labels_len, float* ms, int max_marks, int* n_marks,
                    int* bunches) {
  if (!h || !labels || labels_len <= 0 || !ms || !n_marks) return fail(BP_EINVAL, "bp_get_timeline: null argument");
  Rank* r = h->ranks[0];
  CU_TRY(cudaSetDevice(r->cfg.device));
  CU_TRY(cudaStreamSynchronize(r->compute));
  CU_TRY(cudaStreamSynchronize(r->side));
  const int nb = (int)r->tl_bunches, nm = std::min(r->tl_marks, max_marks);
  labels[0] = 0;
  *n_marks = 0;
  if (bunches) *bunches = nb;
  if (nb == 0 || nm == 0) return BP_OK;
  std::string all;
  for (int k = 0; k < nm; ++k) {
    double sum = 0.0;
    for (int b = 0; b < nb; ++b) {
      float t = 0.0f;
      cudaEvent_t e0 = r->tev[(size_t)b * Rank::kTlCap], ek = r->tev[(size_t)b * Rank::kTlCap + k];
      if (!e0 || !ek) return fail(BP_ECUDA, "bp_get_timeline: events missing");
      CU_TRY(cudaEventElapsedTime(&t, e0, ek));
      sum += t;
    }
    ms[k] = (float)(sum / nb);
    all += r->tl_label[k] ? r->tl_label[k] : "?";
    all += '\n';
  }
  if ((int)all.size() + 1 > labels_len) return fail(BP_EINVAL, "bp_get_timeline: label buffer too small");
  memcpy(labels, all.c_str(), all.size() + 1);
  *n_marks = nm;
  return BP_OK;
}
int bp_get_profile(bp_handle* h, float ms[6], uint64_t* bunches_profiled) {
  if (!h || !ms) return fail(BP_EINVAL, "null argument");
  Rank* r = h->ranks[0];
  CU_TRY(cudaSetDevice(r->cfg.device));
  CU_TRY(cudaStreamSynchronize(r->compute));
  for (int i = 0; i < 6; ++i) ms[i] = 0.0f;
  const int nb = (int)std::min<uint64_t>(r->prof_cnt, Rank::kProfCap);
  for (int b = 0; b < nb; ++b) {
    cudaEvent_t* e = &r->pev[7 * b];
    float t;
    // marks: 0 start | 1 fwd done | 2 dX done | 3 early update (layers >= 2) done | 4 all dW done |
    //        5 all-reduce waited | 6 update done
    CU_TRY(cudaEventElapsedTime(&t, e[0], e[1])); ms[0] += t;   // forward
    CU_TRY(cudaEventElapsedTime(&t, e[1], e[2])); ms[1] += t;   // dX chain (dW_L..dW_2 run beside it)
    CU_TRY(cudaEventElapsedTime(&t, e[3], e[4])); ms[2] += t;   // exposed tail of the dW GEMMs
    CU_TRY(cudaEventElapsedTime(&t, e[5], e[6])); ms[3] += t;   // final update launch
    CU_TRY(cudaEventElapsedTime(&t, e[4], e[5])); ms[4] += t;   // exposed all-reduce
    CU_TRY(cudaEventElapsedTime(&t, e[2], e[3])); ms[5] += t;   // early update, concurrent with dW_1
  }
  if (bunches_profiled) *bunches_profiled = (uint64_t)nb;
  r->prof_cnt = 0;
  return BP_OK;
}

int bp_train_losses(bp_handle* h, int age, double* out, int max_n, int* n_out) {
  if (!h || !out || max_n < 0 || age < 0 || age > 1) return fail(BP_EINVAL, "bp_train_losses: bad argument");
  const int li0 = h->ranks[0]->loss_cur ^ age;
  int n = std::min(max_n, h->ranks[0]->loss_n[li0]);
  for (int i = 0; i < n; ++i) out[i] = 0.0;
  std::vector<double> tmp(n > 0 ? n : 1);
  for (Rank* r : h->ranks) {
    CU_TRY(cudaSetDevice(r->cfg.device));
    const int li = r->loss_cur ^ age;
    if (n > 0) {
      // wait only for the call that produced these losses, not for whatever was queued after it
      CU_TRY(cudaStreamWaitEvent(r->copy, r->loss_done[li], 0));
      CU_TRY(cudaMemcpyAsync(tmp.data(), r->loss_dev[li], sizeof(double) * n, cudaMemcpyDeviceToHost, r->copy));
      CU_TRY(cudaStreamSynchronize(r->copy));
      for (int i = 0; i < n; ++i) out[i] += tmp[i];
    }
  }
  if (n_out) *n_out = n;
  return BP_OK;
}

int bp_comm_unique_id(char id128[128]) {
  if (!id128) return fail(BP_EINVAL, "null id");
  BP_TRY(load_nccl());
  NCCL_TRY(g_nccl.GetUniqueId(id128));
  return BP_OK;
}
int bp_comm_init(bp_handle* h, const char id128[128]) {
  if (!h || !id128 || h->ranks.size() != 1) return fail(BP_EINVAL, "bp_comm_init: needs a rank handle");
  Rank* r = h->ranks[0];
  if (r->cfg.world_size == 1) return BP_OK;
  BP_TRY(load_nccl());
  CU_TRY(cudaSetDevice(r->cfg.device));
  NcclApi::Id128 uid;
  memcpy(uid.b, id128, 128);
  NCCL_TRY(g_nccl.CommInitRank(&r->nccl_comm, r->cfg.world_size, uid, r->cfg.rank));
  return rank_p2p_connect_ipc(r);
}

int bp_dropout_mask(uint64_t seed, uint32_t step, uint32_t layer, uint32_t frame, uint32_t unit, float p) {
  float u[4];
  philox_uniform4((uint32_t)seed, (uint32_t)(seed >> 32), frame >> 2, unit, layer, step, u);
  return u[frame & 3] < p ? 1 : 0;
}

int bp_debug_gemm(int kind, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* out,
                  int ldo, const float* bias, const float* aux, int ldaux, float scale, int act, int math_mode,
                  float* elapsed_ms) {
  if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return fail(BP_EINVAL, "bp_debug_gemm: bad argument");
  if (math_mode != BP_MATH_TF32 && math_mode != BP_MATH_3XTF32)
    return fail(BP_EINVAL, "bp_debug_gemm: math_mode %d", math_mode);
  int dev = 0;
  CU_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) return fail(BP_ENODEV, "device is sm_%d%d, need sm_100", prop.major, prop.minor);
  // operand shapes (rows x cols as stored on the host, row-major with the given ld)
  long long a_rows, a_cols, b_rows, b_cols;
  bool amn, bmn;
  switch (kind) {
    case 0: case 3: a_rows = K; a_cols = M; b_rows = N; b_cols = K; amn = true; bmn = false; break;
    case 1: a_rows = M; a_cols = K; b_rows = N; b_cols = K; amn = false; bmn = false; break;
    case 2: a_rows = K; a_cols = M; b_rows = K; b_cols = N; amn = true; bmn = true; break;
    default: return fail(BP_EINVAL, "bp_debug_gemm: kind %d", kind);
  }
  const long long dlda = round_up(a_cols, 32), dldb = round_up(b_cols, 32), dldo = round_up(M, 32),
                  dldaux = round_up(M, 32);
  float *dA = nullptr, *dB = nullptr, *dO = nullptr, *dBias = nullptr, *dAux = nullptr, *dAlo = nullptr,
        *dBlo = nullptr;
  cudaStream_t st;
  cudaEvent_t e0, e1;
  int rc = [&]() -> int {
    CU_TRY(cudaStreamCreate(&st));
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    CU_TRY(cudaMalloc(&dA, a_rows * dlda * 4));
    CU_TRY(cudaMalloc(&dB, b_rows * dldb * 4));
    CU_TRY(cudaMalloc(&dO, (long long)N * dldo * 4));
    CU_TRY(cudaMemset(dA, 0, a_rows * dlda * 4));
    CU_TRY(cudaMemset(dB, 0, b_rows * dldb * 4));
    CU_TRY(cudaMemset(dO, 0, (long long)N * dldo * 4));
    CU_TRY(cudaMemcpy2D(dA, dlda * 4, A, size_t(lda) * 4, a_cols * 4, a_rows, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy2D(dB, dldb * 4, B, size_t(ldb) * 4, b_cols * 4, b_rows, cudaMemcpyHostToDevice));
    if (bias) {
      CU_TRY(cudaMalloc(&dBias, size_t(M) * 4));
      CU_TRY(cudaMemcpy(dBias, bias, size_t(M) * 4, cudaMemcpyHostToDevice));
    }
    if (aux) {
      CU_TRY(cudaMalloc(&dAux, (long long)N * dldaux * 4));
      CU_TRY(cudaMemcpy2D(dAux, dldaux * 4, aux, size_t(ldaux) * 4, size_t(M) * 4, N, cudaMemcpyHostToDevice));
    }
    if (math_mode == BP_MATH_3XTF32) {
      CU_TRY(cudaMalloc(&dAlo, a_rows * dlda * 4));
      CU_TRY(cudaMalloc(&dBlo, b_rows * dldb * 4));
      bp_split_lo_kernel<<<1184, 256>>>((const float4*)dA, (float4*)dAlo, a_rows * dlda / 4);
      bp_split_lo_kernel<<<1184, 256>>>((const float4*)dB, (float4*)dBlo, b_rows * dldb / 4);
      CU_TRY(cudaGetLastError());
      CU_TRY(cudaDeviceSynchronize());
    }
    AMaps ma;
    MapPair mb;
    BP_TRY(make_a_maps(&ma, dA, dAlo, a_cols, a_rows, dlda, amn));
    BP_TRY(make_map(&mb, dB, dBlo, b_cols, b_rows, dldb, kBlockN, bmn));
    MapPair mb64;
    BP_TRY(make_map(&mb64, dB, dBlo, b_cols, b_rows, dldb, 64, bmn));
    GemmParams p{};
    p.passes = math_mode == BP_MATH_3XTF32 ? 3 : 1;
    p.M = M; p.N = N; p.K = K;
    p.out = dO; p.ldo = dldo;
    p.bias = dBias;
    p.aux = dAux; p.ldaux = dldaux;
    p.scale = scale;
    p.act = act < 0 ? 0 : act;
    int reps = 1;
    if (const char* e = getenv("BP_DBG_REPS")) reps = std::max(1, atoi(e));
    const int sms = prop.multiProcessorCount;
    if (kind == 0 && !bias) return fail(BP_EINVAL, "kind 0 needs bias");
    if (kind == 1 && !aux) return fail(BP_EINVAL, "kind 1 needs aux (Y)");
    for (int rep = 0; rep <= reps; ++rep) {  // rep 0 is an untimed warm-up when reps > 1
      if (rep == (reps > 1 ? 1 : 0)) CU_TRY(cudaEventRecord(e0, st));
      if (reps == 1 && rep == 1) break;
      if (kind == 0) BP_TRY((launch_product(PROD_FWD_HID, st, sms, ma, mb, p, &mb64)));
      else if (kind == 3) BP_TRY((launch_product(PROD_FWD_PLAIN, st, sms, ma, mb, p, &mb64)));
      else if (kind == 1) BP_TRY((launch_product(PROD_DX, st, sms, ma, mb, p, &mb64)));
      else BP_TRY((launch_product(PROD_DW, st, sms, ma, mb, p, &mb64)));
    }
    CU_TRY(cudaEventRecord(e1, st));
    CU_TRY(cudaStreamSynchronize(st));
    if (elapsed_ms) {
      CU_TRY(cudaEventElapsedTime(elapsed_ms, e0, e1));
      *elapsed_ms /= (float)reps;
    }
    CU_TRY(cudaMemcpy2D(out, size_t(ldo) * 4, dO, dldo * 4, size_t(M) * 4, N, cudaMemcpyDeviceToHost));
    return BP_OK;
  }();
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dBias); cudaFree(dAux); cudaFree(dAlo); cudaFree(dBlo);
  return rc;
}

int bp_debug_sgd(int n, float* delta, float* weights, const float* grad, int bunch, float momentum, float lrate,
                 float weightcost) {
  if (n <= 0 || !delta || !weights || !grad) return fail(BP_EINVAL, "bp_debug_sgd: bad argument");
  const long long n4 = (n + 3) / 4;
  float *d = nullptr, *w = nullptr, *g = nullptr;
  int rc = [&]() -> int {
    CU_TRY(cudaMalloc(&d, n4 * 16));
    CU_TRY(cudaMalloc(&w, n4 * 16));
    CU_TRY(cudaMalloc(&g, n4 * 16));
    CU_TRY(cudaMemset(d, 0, n4 * 16));
    CU_TRY(cudaMemset(w, 0, n4 * 16));
    CU_TRY(cudaMemset(g, 0, n4 * 16));
    CU_TRY(cudaMemcpy(d, delta, size_t(n) * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(w, weights, size_t(n) * 4, cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(g, grad, size_t(n) * 4, cudaMemcpyHostToDevice));
    SgdBiasRanges br{};
    const float c1 = (1 - momentum) * lrate;
    if (weightcost != 0.0f)
      bp_sgd_kernel<true><<<148 * 8, 256>>>((float4*)d, (float4*)w, (const float4*)g, 0, n4, (float)bunch, momentum, c1,
                                            weightcost, br, nullptr, 1);
    else
      bp_sgd_kernel<false><<<148 * 8, 256>>>((float4*)d, (float4*)w, (const float4*)g, 0, n4, (float)bunch, momentum,
                                             c1, 0.0f, br, nullptr, 1);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(delta, d, size_t(n) * 4, cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemcpy(weights, w, size_t(n) * 4, cudaMemcpyDeviceToHost));
    return BP_OK;
  }();
  cudaFree(d); cudaFree(w); cudaFree(g);
  return rc;
}

}  // extern "C"
